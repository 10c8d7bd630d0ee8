// Conv-coupling conditioner (layers/coupling.py:26-29): Conv1x1(Cin->Ch) ReLU Conv KHxKW reflect (Ch->Ch) ReLU Conv1x1(Ch->Cout),
// fused: a CTA keeps S whole samples (so the reflect halo is internal) resident in shared memory through all three
// convolutions; only x0 is read from and h written to HBM.  FP32 FMA path (fp32-faithful; see DESIGN.md for why
// single-pass TF32/BF16 tensor-core operands cannot meet the 1e-4 parity budget).
#include "tile_gemm.cuh"

namespace cfpp {

struct ConvCondArgs {
  const float* x; int64_t x_bstride; float* h;
  const float* w1t; const float* b1; const float* bias1_b;
  const float* w2t; const float* b2; const float* w3t; const float* b3;
  int B, Cin, Ch, Cout, H, W, S, NPh, NP3, PS;
};

__device__ __forceinline__ int reflect_idx(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

template <int KH, int KW, int TP>
__global__ void __launch_bounds__(256) conv_cond_kernel(const ConvCondArgs a) {
  constexpr int TAPS = KH * KW;
  constexpr int CC2 = 36 / TAPS;                      // 3x3: 4 channels, 3x1: 12, 1x1: 36 -> 36 weight rows per chunk
  extern __shared__ float4 sm4[];
  float* R0 = reinterpret_cast<float*>(sm4);
  const int HW = a.H * a.W, PS = a.PS;
  float* R1 = R0 + (int64_t)a.Ch * PS;
  float* Wbuf = R1 + (int64_t)a.Ch * PS;
  const int b0 = blockIdx.x * a.S;
  const int nS = min(a.S, a.B - b0);
  const int P = nS * HW;

  // ---- stage 0: x0 (first Cin channels of each sample) -> R0[c][p] ----
  for (int idx = threadIdx.x; idx < a.Cin * PS; idx += blockDim.x) {
    const int c = idx / PS, p = idx % PS;
    float v = 0.f;
    if (p < P) { const int s = p / HW, hw = p % HW; v = a.x[(int64_t)(b0 + s) * a.x_bstride + (int64_t)c * HW + hw]; }
    R0[idx] = v;
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_nt = a.NPh / kTN;
  const int ntile = warp % n_nt, ptile = warp / n_nt;
  const int n0 = ntile * kTN;
  int off1[1][TP], off2[TAPS][TP];
#pragma unroll
  for (int tp = 0; tp < TP; ++tp) {
    const int p = ptile * 32 * TP + tp * 32 + lane;
    off1[0][tp] = p;
    const int pc = p < P ? p : 0;
    const int s = pc / HW, hw = pc % HW, y = hw / a.W, xx = hw % a.W;
#pragma unroll
    for (int ky = 0; ky < KH; ++ky)
#pragma unroll
      for (int kx = 0; kx < KW; ++kx)
        off2[ky * KW + kx][tp] = s * HW + reflect_idx(y + ky - KH / 2, a.H) * a.W + reflect_idx(xx + kx - KW / 2, a.W);
  }
  __syncthreads();

  float acc[TP][kTN];
  // ---- stage 1: H1 = relu(W1 x0 + b1) -> R1 ----
  cta_gemm<1, TP, 32>(R0, PS, a.Cin, a.w1t, a.NPh, Wbuf, off1, n0, true, acc);
#pragma unroll
  for (int tp = 0; tp < TP; ++tp) {
    const int p = off1[0][tp];
    const float* pb = (a.bias1_b && p < P) ? a.bias1_b + (int64_t)(b0 + p / HW) * a.Ch : nullptr;
#pragma unroll
    for (int j = 0; j < kTN; ++j) {
      const int n = n0 + j;
      if (n < a.Ch) R1[(int64_t)n * PS + p] = fmaxf(acc[tp][j] + (pb ? pb[n] : a.b1[n]), 0.f);
    }
  }
  __syncthreads();
  // ---- stage 2: H2 = relu(conv_kxk_reflect(H1) + b2) -> R0 ----
  cta_gemm<TAPS, TP, CC2>(R1, PS, a.Ch, a.w2t, a.NPh, Wbuf, off2, n0, true, acc);
#pragma unroll
  for (int tp = 0; tp < TP; ++tp) {
    const int p = off1[0][tp];
#pragma unroll
    for (int j = 0; j < kTN; ++j) {
      const int n = n0 + j;
      if (n < a.Ch) R0[(int64_t)n * PS + p] = fmaxf(acc[tp][j] + a.b2[n], 0.f);
    }
  }
  __syncthreads();
  // ---- stage 3: h = W3 H2 + b3 -> HBM ----
  const bool act3 = n0 < a.NP3;
  cta_gemm<1, TP, 32>(R0, PS, a.Ch, a.w3t, a.NP3, Wbuf, off1, n0, act3, acc);
  if (act3) {
#pragma unroll
    for (int tp = 0; tp < TP; ++tp) {
      const int p = off1[0][tp];
      if (p < P) {
        const int s = p / HW, hw = p % HW;
        float* hb = a.h + ((int64_t)(b0 + s) * a.Cout) * HW + hw;
#pragma unroll
        for (int j = 0; j < kTN; ++j) {
          const int n = n0 + j;
          if (n < a.Cout) hb[(int64_t)n * HW] = acc[tp][j] + a.b3[n];
        }
      }
    }
  }
}

template <int KH, int KW, int TP>
static int launch_conv_cond(const ConvCondArgs& a, int nwarps, size_t smem, cudaStream_t st) {
  static DeviceOnce attr_set;
  if (attr_set.first()) { cudaFuncSetAttribute(conv_cond_kernel<KH, KW, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }
  const int blocks = (a.B + a.S - 1) / a.S;
  conv_cond_kernel<KH, KW, TP><<<blocks, nwarps * 32, smem, st>>>(a);
  return check_launch("conv_cond_fwd");
}

}  // namespace cfpp
using namespace cfpp;

extern "C" int cfpp_conv_cond_fwd(const float* x, int64_t x_bstride, float* h,
                                  const float* w1t, const float* b1, const float* bias1_b,
                                  const float* w2t, const float* b2, const float* w3t, const float* b3,
                                  int B, int Cin, int Ch, int Cout, int H, int W, int KH, int KW, void* stream) {
  CFPP_REQUIRE(Cin >= 1 && Ch >= 1 && Cout >= 1 && H >= 1 && W >= 1, "conv_cond: bad dims");
  CFPP_REQUIRE((KH == 3 && KW == 3) || (KH == 3 && KW == 1) || (KH == 1 && KW == 1), "conv_cond: kernel %dx%d unsupported", KH, KW);
  CFPP_REQUIRE((KH == 1 || H >= 2) && (KW == 1 || W >= 2), "conv_cond: reflect padding needs dim >= 2");
  if (B <= 0) return CFPP_OK;
  const int HW = H * W;
  ConvCondArgs a{x, x_bstride, h, w1t, b1, bias1_b, w2t, b2, w3t, b3, B, Cin, Ch, Cout, H, W, 1, 0, 0, 0};
  a.NPh = (Ch + 15) / 16 * 16; a.NP3 = (Cout + 15) / 16 * 16;
  const int n_nt = a.NPh / 16;
  CFPP_REQUIRE(n_nt <= 8, "conv_cond: hidden width %d > 128 unsupported", Ch);
  int S = HW >= 128 ? 1 : 128 / HW;
  if (S > B) S = B;
  int P = S * HW;
  const int TP = P >= 96 ? 4 : 2;
  int n_pt = (P + 32 * TP - 1) / (32 * TP);
  CFPP_REQUIRE(n_pt * n_nt <= 8, "conv_cond: tile %d pixels x %d channels exceeds one CTA (image too large)", P, Ch);
  a.S = S; a.PS = n_pt * 32 * TP;
  const size_t smem = ((size_t)2 * Ch * a.PS + (size_t)2 * 36 * a.NPh) * sizeof(float);
  CFPP_REQUIRE(smem <= 227 * 1024, "conv_cond: needs %zu bytes of shared memory", smem);
  const int nwarps = n_pt * n_nt;
  cudaStream_t st = (cudaStream_t)stream;
#define CFPP_CC(KH_, KW_) (TP == 4 ? launch_conv_cond<KH_, KW_, 4>(a, nwarps, smem, st) : launch_conv_cond<KH_, KW_, 2>(a, nwarps, smem, st))
  if (KH == 3 && KW == 3) return CFPP_CC(3, 3);
  if (KH == 3 && KW == 1) return CFPP_CC(3, 1);
  return CFPP_CC(1, 1);
#undef CFPP_CC
}
