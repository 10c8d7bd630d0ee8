// CTA-level FP32 GEMM over shared-memory-resident activations (used by the conv and ViT conditioners).
//
//   out[n][p] = sum_{c < Kin} sum_{tap < TAPS}  Wt[(c*TAPS + tap)][n] * A[c][ off[tap][p] ]
//
// Activations live in shared memory feature-major, A[c][PS] (pixels / token-rows contiguous), so a warp whose lanes own
// consecutive pixels reads them conflict-free; `off` is a per-thread gather table (identity for 1x1 / linear layers,
// reflect-padded neighbours for the kxk conv).  Weights are packed K-major with the output dimension padded to a
// multiple of 16 (NP) and streamed global->shared in double-buffered cp.async chunks of CC input channels; inside a
// chunk every lane reads its 16 weights as four warp-broadcast 128-bit loads.
// Thread tile: TP pixels (p = tile_base + tp*32 + lane) x 16 output channels, FP32 FMA, fp32 accumulate.
#pragma once
#include "common.cuh"

namespace cfpp {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kTN = 16;   // output channels per thread

// All threads of the CTA must call this (it synchronises).  `active` warps accumulate into acc.
template <int TAPS, int TP, int CC>
__device__ __forceinline__ void cta_gemm(const float* __restrict__ A, int PS, int Kin,
                                         const float* __restrict__ Wt, int NP, float* __restrict__ Wbuf,
                                         const int (&off)[TAPS][TP], int n0, bool active, float (&acc)[TP][kTN]) {
#pragma unroll
  for (int tp = 0; tp < TP; ++tp)
#pragma unroll
    for (int j = 0; j < kTN; ++j) acc[tp][j] = 0.f;

  const int rows_per_chunk = CC * TAPS;
  const int buf_floats = rows_per_chunk * NP;
  const int nchunks = (Kin + CC - 1) / CC;
  auto issue = [&](int c) {
    const int c0 = c * CC;
    const int rows = (Kin - c0 < CC ? Kin - c0 : CC) * TAPS;
    const float4* src = reinterpret_cast<const float4*>(Wt + (int64_t)c0 * TAPS * NP);
    float4* dst = reinterpret_cast<float4*>(Wbuf + (c & 1) * buf_floats);
    const int n4 = rows * NP / 4;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) cp_async16(dst + i, src + i);
    cp_async_commit();
  };
  issue(0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) { issue(c + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
    if (active) {
      const float* Wc = Wbuf + (c & 1) * buf_floats + n0;
      const int c0 = c * CC;
      const int cn = Kin - c0 < CC ? Kin - c0 : CC;
      for (int cc = 0; cc < cn; ++cc) {
        const float* Arow = A + (int64_t)(c0 + cc) * PS;
#pragma unroll
        for (int tap = 0; tap < TAPS; ++tap) {
          const float4* w4 = reinterpret_cast<const float4*>(Wc + (cc * TAPS + tap) * NP);
          const float4 w0 = w4[0], w1 = w4[1], w2 = w4[2], w3 = w4[3];
          const float w[kTN] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
          for (int tp = 0; tp < TP; ++tp) {
            const float a = Arow[off[tap][tp]];
#pragma unroll
            for (int j = 0; j < kTN; ++j) acc[tp][j] = fmaf(a, w[j], acc[tp][j]);
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace cfpp
