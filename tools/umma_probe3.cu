// Third tcgen05 probe: issue-rate of the instruction patterns the conv conditioner can use for fp32-faithful 3xTF32
// (pixels on M=128, channels on N, K=8 per instruction).  A host-built "program" of MMAs (accumulator column, A/B smem
// offsets, N) is replayed by one thread; clocks per k-step are reported for several orders / accumulator layouts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe3 tools/umma_probe3.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

// smem map (bytes): A_hi tiles at 0 (T x 16 KB), A_lo at 64K, B_hi at 128K (N x 128), B_lo right after B_hi (so [B_hi;B_lo] is one 2N operand)
constexpr uint32_t AHI = 0, ALO = 64 * 1024, BHI = 128 * 1024;

// variant 0: classic 3 passes, one accumulator per tile [0,N); ks outer, tile inner
// variant 1: concat pass (N'=2N -> cols [0,2N)) + lo pass into [0,N); ks outer, tile inner
// variant 2: concat pass + lo pass into separate cols [2N,3N); ks outer, tile inner
// variant 3: like 2 but tile outer, ks inner
// variant 4: like 1 but tile outer, ks inner
// variant 5: like 0 but passes outermost inside a k-step group: for pass: for ks: for tile
template <int N, int T, int V>
__global__ void __launch_bounds__(128) probe_kernel(int reps, long long* __restrict__ cycles, int roff, int sbo, int nw, int gap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar_[4];
  uint64_t& bar = bar_[threadIdx.x >> 5];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024; i += 128) ((float*)base)[i] = 0.001f * (i % 97);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if ((tid & 31) == 0) {
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
  const uint32_t tmem = tmem_base_s + (warp & 1) * 256;
  if ((tid & 31) == 0 && warp < nw) {
    const uint32_t s0 = smem_u32(base);
    constexpr uint32_t idN = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t id2N = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int stride = V == 0 || V == 5 ? N : (V == 2 || V == 3 ? 3 * N : 2 * N);
    const uint64_t dAhi = make_desc(s0 + AHI + roff * 128, sbo), dAlo = make_desc(s0 + ALO + roff * 128, sbo), dBhi = make_desc(s0 + BHI, 1024), dBlo = make_desc(s0 + BHI + N * 128, 1024);
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      if (gap > 0) { const long long tg = clock64(); while (clock64() - tg < gap) {} }
      auto emit = [&](int t, int ks) {
        const uint32_t col = tmem + t * stride; const uint64_t a = (uint64_t)((t * 16384 + ks * 32) >> 4), b = (uint64_t)((ks * 32) >> 4);
        if (V == 0) {
          mma_tf32(col, dAhi + a, dBhi + b, idN, 1); mma_tf32(col, dAhi + a, dBlo + b, idN, 1); mma_tf32(col, dAlo + a, dBhi + b, idN, 1);
        } else {
          mma_tf32(col, dAhi + a, dBhi + b, id2N, 1);
          mma_tf32(col + ((V == 2 || V == 3) ? 2 * N : 0), dAlo + a, dBhi + b, idN, 1);
        }
      };
      if (V == 5) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int t = 0; t < T; ++t) mma_tf32(tmem + t * stride, dAhi + (uint64_t)((t * 16384 + ks * 32) >> 4), dBhi + (uint64_t)((ks * 32) >> 4), idN, 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int t = 0; t < T; ++t) mma_tf32(tmem + t * stride, dAhi + (uint64_t)((t * 16384 + ks * 32) >> 4), dBlo + (uint64_t)((ks * 32) >> 4), idN, 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int t = 0; t < T; ++t) mma_tf32(tmem + t * stride, dAlo + (uint64_t)((t * 16384 + ks * 32) >> 4), dBhi + (uint64_t)((ks * 32) >> 4), idN, 1);
      } else if (V == 3 || V == 4) {
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) emit(t, ks);
      } else {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int t = 0; t < T; ++t) emit(t, ks);
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    cycles[warp] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static int g_roff = 0, g_sbo = 1024, g_nw = 1, g_gap = 0;
template <int N, int T, int V>
void run(long long* dC) {
  constexpr int stride = V == 0 || V == 5 ? N : (V == 2 || V == 3 ? 3 * N : 2 * N);
  if (T * stride > 256) return;
  const size_t smem = 192 * 1024 + 1024;
  cudaFuncSetAttribute(probe_kernel<N, T, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long best = 1LL << 60;
  const int reps = 64;
  for (int it = 0; it < 4; ++it) {
    probe_kernel<N, T, V><<<1, 128, smem>>>(reps, dC, g_roff, g_sbo, g_nw, g_gap);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
    long long c; cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
    if (c < best) best = c;
  }
  const int nmma = 4 * T * (V == 0 || V == 5 ? 3 : 2);
  const double per_kstep = (double)best / (reps * 4 * T);
  printf("gap=%4d issuers=%d roff=%d sbo=%d N=%3d tiles=%d variant=%d : %6.1f clk per (tile,k-step) -> %6.0f useful MAC/clk/SM (%.1f clk/MMA)\n", g_gap, g_nw, g_roff, g_sbo, N, T, V, per_kstep,
         128.0 * N * 8 / per_kstep, (double)best / (reps * nmma));
}
template <int N, int T> void runv(long long* dC) { run<N, T, 0>(dC); run<N, T, 1>(dC); }
template <int N> void runt(long long* dC) { runv<N, 1>(dC); runv<N, 2>(dC); }

int main() {
  long long* dC; cudaMalloc(&dC, 64);
  for (int gap : {0, 50, 100, 200, 300, 400, 600}) { g_gap = gap; run<64, 1, 1>(dC); run<64, 2, 1>(dC); run<128, 1, 1>(dC); }
  return 0;
}
