// Standalone probe of the tcgen05 conventions the conv conditioner relies on (run on a B200 via gpurun):
//   * K-major SWIZZLE_128B smem descriptors for tf32 A (M=128 rows) and B (N rows), 4 k-steps of K=8 per 128-byte row
//   * A views that start at an arbitrary ROW offset inside a swizzled tile (3x3 conv taps as shifted views)
//   * TMEM allocation, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b lane mapping
//   * accuracy of single-pass TF32 vs the 3-pass hi/lo split against an fp64 reference
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <vector>
#include <cstring>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);       // start address, 16-byte units
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  for (int it = 0; it < 2000000; ++it) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

// smem image helpers: logical (row, k) of a K-major tile with 32 fp32 per row -> byte offset with the 128B swizzle
__device__ __forceinline__ uint32_t sw128_off(int row, int k) { return row * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + ((k & 3) << 2); }

constexpr int AROWS = 160;   // 128 + slack for row offsets

// mode: 0 = single pass (raw fp32 operands, hardware truncates to tf32); 1 = three passes hi*hi + hi*lo + lo*hi
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ D,
                                                    int N, int roff, int base_off, int mode, int* __restrict__ status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                         // AROWS x 128 B
  uint8_t* sAl = sA + AROWS * 128;            // lo part
  uint8_t* sB = sAl + AROWS * 128;            // N x 128 B
  uint8_t* sBl = sB + 128 * 128;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < AROWS * 32; i += 128) {
    const int r = i >> 5, k = i & 31;
    const float v = A[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    *(float*)(sA + sw128_off(r, k)) = v;
    *(float*)(sAl + sw128_off(r, k)) = v - hi;
  }
  for (int i = tid; i < N * 32; i += 128) {
    const int r = i >> 5, k = i & 31;
    const float v = Bm[i];
    const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    *(float*)(sB + sw128_off(r, k)) = v;
    *(float*)(sBl + sw128_off(r, k)) = v - hi;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(sA) + roff * 128, al0 = smem_u32(sAl) + roff * 128, b0 = smem_u32(sB), bl0 = smem_u32(sBl);
    uint32_t acc = 0;
    for (int ks = 0; ks < 4; ++ks) {
      mma_tf32(tmem, make_desc_k_sw128(a0 + ks * 32, base_off), make_desc_k_sw128(b0 + ks * 32, 0), idesc, acc); acc = 1;
      if (mode == 1) {
        mma_tf32(tmem, make_desc_k_sw128(a0 + ks * 32, base_off), make_desc_k_sw128(bl0 + ks * 32, 0), idesc, 1);
        mma_tf32(tmem, make_desc_k_sw128(al0 + ks * 32, base_off), make_desc_k_sw128(b0 + ks * 32, 0), idesc, 1);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  const bool done = mbar_wait_bounded(smem_u32(&bar), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!done) { if (tid == 0) *status = 1; }
  else {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                     "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
                     "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                   : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

static float tf32_trunc(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; }

int main() {
  std::vector<float> A(AROWS * 32), B(128 * 32);
  uint32_t s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f + ((s >> 3) & 0xFF) * 1e-6f; };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  float *dA, *dB, *dD; int* dS;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * 128 * 4); cudaMalloc(&dS, 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 2 * AROWS * 128 + 2 * 128 * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Case { int N, roff, boff, mode; };
  std::vector<Case> cases;
  for (int N : {32, 64, 128, 16, 8}) cases.push_back({N, 0, 0, 0});
  for (int r = 1; r <= 9; ++r) { cases.push_back({32, r, 0, 0}); cases.push_back({32, r, r & 7, 0}); }
  cases.push_back({32, 19, 0, 0}); cases.push_back({64, 21, 0, 1}); cases.push_back({32, 0, 0, 1}); cases.push_back({128, 3, 0, 1});
  int bad = 0;
  for (auto c : cases) {
    cudaMemset(dD, 0, 128 * 128 * 4); cudaMemset(dS, 0, 4);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, c.N, c.roff, c.boff, c.mode, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    std::vector<float> D(128 * c.N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err_t = 0, err_f = 0, mag = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < c.N; ++n) {
        double rt = 0, rf = 0;
        for (int k = 0; k < 32; ++k) {
          const float a = A[(m + c.roff) * 32 + k], b = B[n * 32 + k];
          rt += (double)tf32_trunc(a) * (double)tf32_trunc(b); rf += (double)a * (double)b;
        }
        const double d = D[m * c.N + n];
        err_t = fmax(err_t, fabs(d - rt)); err_f = fmax(err_f, fabs(d - rf)); mag = fmax(mag, fabs(rf));
      }
    const bool ok = e == cudaSuccess && st == 0 && (c.mode == 0 ? err_t < 1e-4 : err_f < 2e-5);
    printf("N=%3d roff=%2d base_off=%d mode=%d : cuda=%s timeout=%d  max|d-ref_tf32|=%.3e  max|d-ref_fp64|=%.3e  (|ref|max %.2f)  %s\n", c.N, c.roff, c.boff,
           c.mode, cudaGetErrorName(e), st, err_t, err_f, mag, ok ? "OK" : "MISMATCH");
    bad += !ok;
    if (e != cudaSuccess) break;
  }
  printf("probe: %d case(s) off\n", bad);
  return 0;
}
