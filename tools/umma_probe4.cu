// Fourth tcgen05 probe: cycles per kind::f16 MMA (M = 128 or 64, K = 16) as a function of N, operand swizzle (64 B / 128 B rows),
// the A operand's group stride (aligned 8-row groups vs the conv conditioner's GS-row segments) and the number of k-steps that walk
// along one operand row.  One thread issues `reps` x `nmma` instructions back to back and waits for the commit.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe4 tools/umma_probe4.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, int rb) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(rb == 128 ? 2 : 4) << 61);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

// mode 0: every MMA has N columns.  mode 1: the conditioner's pair (N' = 2N into [0,2N), then N into [N,2N)).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
template <int M, int N, int rb, int mode>
__global__ void __launch_bounds__(256) probe_kernel(int reps, long long* __restrict__ cycles, int sbo, int traffic) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 40 * 1024; i += 256) ((uint32_t*)base)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    done = 0;
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
  if (warp == 0) { if (elect_one()) {
    const uint32_t s0 = smem_u32(base);
    const uint32_t idN = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t id2N = (1u << 4) | ((uint32_t)((2 * N) >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t dA = make_desc(s0, sbo, rb), dA2 = make_desc(s0 + 48 * 1024, sbo, rb), dB = make_desc(s0 + 96 * 1024, 8 * rb, rb);
    constexpr int ksn = rb / 32;
    const long long t0 = clock64();
#pragma unroll 1
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) {
        const uint64_t a = dA + (uint64_t)(tap * (rb >> 4)), a2 = dA2 + (uint64_t)(tap * (rb >> 4));
        if (ksn == 4) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (mode == 0) { mma_f16(tmem, a + 2 * ks, dB + 2 * ks, idN, 1); mma_f16(tmem + 256, a2 + 2 * ks, dB + 2 * ks, idN, 1); }
            else { mma_f16(tmem, a + 2 * ks, dB + 2 * ks, id2N, 1); mma_f16(tmem + N, a2 + 2 * ks, dB + 2 * ks, idN, 1); }
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            if (mode == 0) { mma_f16(tmem, a + 2 * ks, dB + 2 * ks, idN, 1); mma_f16(tmem + 256, a2 + 2 * ks, dB + 2 * ks, idN, 1); }
            else { mma_f16(tmem, a + 2 * ks, dB + 2 * ks, id2N, 1); mma_f16(tmem + N, a2 + 2 * ks, dB + 2 * ks, idN, 1); }
          }
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    cycles[0] = clock64() - t0;
    done = 1;
  } } else if (traffic && warp >= 1) {
    // background shared-memory store traffic like the epilogue warps' operand writes (16-byte stores, one row per lane)
    uint8_t* dst = base + 140 * 1024 + (warp - 1) * 4096;
    const int lane = tid & 31;
    uint32_t v = tid;
    while (!done) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        *reinterpret_cast<uint4*>(dst + lane * 128 + (((q ^ lane) & 7) << 4)) = make_uint4(v, v + 1, v + 2, v + 3);
        v = v * 1664525u + 1013904223u;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static long long* dC;
static const size_t smem = 200 * 1024;
template <int M, int N, int rb, int mode>
void run(int sbo, int traffic) {
  const int reps = 32;
  cudaFuncSetAttribute(probe_kernel<M, N, rb, mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long best = 1LL << 60;
  for (int it = 0; it < 3; ++it) {
    probe_kernel<M, N, rb, mode><<<1, 256, smem>>>(reps, dC, sbo, traffic);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
    long long c; cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
    if (c < best) best = c;
  }
  const int nmma = reps * 4 * (rb / 32) * 2;
  printf("M=%3d N=%3d rb=%3d sbo=%4d mode=%d traffic=%d : %6.1f clk/MMA\n", M, N, rb, sbo, mode, traffic, (double)best / nmma);
}
template <int N> void runN() {
  for (int traffic : {0, 1}) {
    run<128, N, 128, 0>(1024, traffic); run<128, N, 128, 0>(1280, traffic); run<128, N, 64, 0>(512, traffic); run<128, N, 64, 0>(640, traffic);
    if (N <= 128) { run<128, N, 128, 1>(1024, traffic); run<128, N, 128, 1>(1280, traffic); run<128, N, 64, 1>(512, traffic); run<128, N, 64, 1>(640, traffic); }
  }
  run<64, N, 128, 0>(1024, 0);
}
int main() {
  cudaMalloc(&dC, 64);
  runN<16>(); runN<32>(); runN<64>(); runN<128>(); runN<256>();
  return 0;
}
