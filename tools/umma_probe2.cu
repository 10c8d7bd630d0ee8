// Second tcgen05 probe (run on a B200 via gpurun) for the conv-conditioner design:
//   (1) correctness of K-major SWIZZLE_128B A views whose 8-row groups are strided by SBO != 1024 B (segment layout:
//       8 output pixels per group, groups 10 stored rows apart) starting at arbitrary row offsets, N in {16,32,64,128};
//   (2) issue-rate of back-to-back SS-mode kind::tf32 MMAs (M=128, K=8) versus N -- is the shared-memory operand read
//       (4 KB of A per instruction) the limiter at small N?
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe2 tools/umma_probe2.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <vector>
#include <cstring>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  for (int it = 0; it < 20000000; ++it) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}

__device__ __forceinline__ uint32_t sw128_off(int row, int k) { return row * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + ((k & 3) << 2); }

constexpr int AROWS = 256;

// A: AROWS x 32 floats (row-major logical), B: N x 32.  Output row m = 8*g + j reads A row  roff + g*gs + j.
// reps > 0: timing mode -- the 4 k-steps are issued `reps` times (accumulating) and elapsed clocks are reported.
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ D,
                                                    int N, int roff, int gs, int reps, int nacc, int* __restrict__ status, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                         // AROWS x 128 B
  uint8_t* sB = sA + AROWS * 128;             // 256 x 128 B
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  for (int i = tid; i < AROWS * 32; i += 128) { const int r = i >> 5, k = i & 31; *(float*)(sA + sw128_off(r, k)) = A[i]; }
  for (int i = tid; i < 256 * 32; i += 128) { const int r = i >> 5, k = i & 31; *(float*)(sB + sw128_off(r, k)) = r < N ? Bm[i] : 0.f; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  long long t0 = 0;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(sA) + roff * 128, b0 = smem_u32(sB);
    const uint32_t sbo = gs * 128;
    uint32_t acc = 0;
    const int R = reps > 0 ? reps : 1;
    t0 = clock64();
    for (int rep = 0; rep < R; ++rep)
      for (int ks = 0; ks < 4; ++ks)
        for (int p = 0; p < nacc; ++p) {
          mma_tf32(tmem + p * N, make_desc(a0 + ks * 32, sbo), make_desc(b0 + ks * 32, 1024), idesc, acc);
          if (p == nacc - 1) acc = 1;
        }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  const bool done = mbar_wait_bounded(smem_u32(&bar), 0);
  if (tid == 0) *cycles = clock64() - t0;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!done) { if (tid == 0) *status = 1; }
  else if (reps == 0) {
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                     "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(taddr) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static float tf32_trunc(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v; }

int main() {
  std::vector<float> A(AROWS * 32), B(256 * 32);
  uint32_t s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f + ((s >> 3) & 0xFF) * 1e-6f; };
  for (auto& v : A) v = rnd();
  for (auto& v : B) v = rnd();
  float *dA, *dB, *dD; int* dS; long long* dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * 256 * 4); cudaMalloc(&dS, 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = AROWS * 128 + 256 * 128 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Case { int N, roff, gs; };
  std::vector<Case> cases;
  for (int N : {16, 32, 64, 128, 256}) { cases.push_back({N, 0, 8}); cases.push_back({N, 3, 10}); }
  for (int r : {0, 1, 5, 11, 23}) { cases.push_back({32, r, 10}); cases.push_back({64, r, 12}); }
  cases.push_back({32, 7, 9}); cases.push_back({128, 2, 6});
  int bad = 0;
  for (auto c : cases) {
    cudaMemset(dD, 0, 128 * 256 * 4); cudaMemset(dS, 0, 4);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, c.N, c.roff, c.gs, 0, 1, dS, dC);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0; cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    std::vector<float> D(128 * c.N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double err_t = 0;
    int skipped = 0;
    for (int m = 0; m < 128; ++m) {
      const int arow = c.roff + (m >> 3) * c.gs + (m & 7);
      if (arow >= AROWS) { ++skipped; continue; }
      for (int n = 0; n < c.N; ++n) {
        double rt = 0;
        for (int k = 0; k < 32; ++k) rt += (double)tf32_trunc(A[arow * 32 + k]) * (double)tf32_trunc(B[n * 32 + k]);
        err_t = fmax(err_t, fabs((double)D[m * c.N + n] - rt));
      }
    }
    const bool ok = e == cudaSuccess && st == 0 && err_t < 1e-4;
    printf("N=%3d roff=%2d group_stride=%2d rows : cuda=%s timeout=%d max|d-ref_tf32|=%.3e (rows skipped %d) %s\n", c.N, c.roff, c.gs,
           cudaGetErrorName(e), st, err_t, skipped, ok ? "OK" : "MISMATCH");
    bad += !ok;
    if (e != cudaSuccess) break;
  }
  printf("probe2 correctness: %d case(s) off\n", bad);
  // ---- issue rate ----
  for (int N : {16, 32, 64, 128, 256})
    for (int nacc : {1, 2, 3, 4, 8}) {
      if (nacc * N > 512) continue;
      const int gs = 10;
      long long best = 1LL << 60;
      const int reps = 256;
      for (int it = 0; it < 5; ++it) {
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, 0, gs, reps, nacc, dS, dC);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("timing launch failed\n"); return 1; }
        long long cyc; cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
        if (cyc < best) best = cyc;
      }
      const double per = (double)best / (reps * 4 * nacc);
      printf("rate: M=128 N=%3d K=8 tf32 SS accumulators=%2d : %.1f clk/MMA  (math floor N/2 = %d; A+B bytes %d -> %.0f B/clk; %.0f MAC/clk)\n", N, nacc, per, N / 2,
             4096 + 32 * N, (4096 + 32 * N) / per, 128.0 * N * 8 / per);
    }
  return 0;
}
