// SimpleViT conditioner of TransCoupling (layers/simple_vit.py:91-127; heads=1, dim_head=64, dim=mlp_dim=T), fused:
// a CTA keeps the token rows of S whole samples resident in shared memory (feature-major) through patch embedding,
// `depth` pre-LN transformer layers and the final LayerNorm; only x0 is read from and h written to HBM.
// GEMMs use the FP32 tile routine of tile_gemm.cuh with weights streamed from L2 (cp.async, double buffered).
#include "tile_gemm.cuh"

namespace cfpp {

struct VitArgs {
  const float* x; int64_t x_bstride; const float* extra; int Cextra; float* h;
  cfpp_vit_desc d; int B, S, PS, NPT, ybuf_rows, qbuf_rows;
};

constexpr int kDH = 64;          // dim_head (simple_vit.py:44)
constexpr int kQKV = 3 * kDH;
constexpr int kVitCC = 16;       // weight rows per cp.async chunk

__host__ __device__ inline int64_t vit_layer_floats(int T) {
  const int64_t NP = (T + 15) / 16 * 16;
  return 4 * (int64_t)T + (int64_t)T * 192 + 64 * NP + 2 * (int64_t)T * NP + 2 * NP;
}

// LayerNorm over the feature axis of rows [0,R): 4 consecutive lanes share a row.  src/dst are [F][PS]; may alias.
__device__ void layer_norm_rows(const float* src, float* dst, int F, int PS, int R, const float* __restrict__ w, const float* __restrict__ b) {
  for (int base = 0; base < R; base += blockDim.x / 4) {
    const int r = base + threadIdx.x / 4, q = threadIdx.x & 3;
    const bool ok = r < R;
    float s = 0.f;
    if (ok) for (int f = q; f < F; f += 4) s += src[f * PS + r];
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2);
    const float mean = s / (float)F;
    float v = 0.f;
    if (ok) for (int f = q; f < F; f += 4) { const float dlt = src[f * PS + r] - mean; v += dlt * dlt; }
    v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
    const float rstd = 1.0f / sqrtf(v / (float)F + 1e-5f);
    if (ok) for (int f = q; f < F; f += 4) dst[f * PS + r] = (src[f * PS + r] - mean) * rstd * w[f] + b[f];
  }
}

// One GEMM over all row tiles: dst-op(epi) applied per thread tile.  All threads must call.
template <int TP, class Epi>
__device__ __forceinline__ void gemm_rounds(const float* A, int PS, int Kin, const float* __restrict__ Wt, int NP, float* Wbuf, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int n_pt = PS / (32 * TP), n_nt = NP / kTN, tiles = n_pt * n_nt;
  for (int t0 = 0; t0 < tiles; t0 += nw) {
    const int tile = t0 + warp;
    const bool active = tile < tiles;
    const int ptile = active ? tile / n_nt : 0, ntile = active ? tile % n_nt : 0;
    int off[1][TP];
#pragma unroll
    for (int tp = 0; tp < TP; ++tp) off[0][tp] = ptile * 32 * TP + tp * 32 + lane;
    float acc[TP][kTN];
    cta_gemm<1, TP, kVitCC>(A, PS, Kin, Wt, NP, Wbuf, off, ntile * kTN, active, acc);
    if (active) {
#pragma unroll
      for (int tp = 0; tp < TP; ++tp)
#pragma unroll
        for (int j = 0; j < kTN; ++j) epi(off[0][tp], ntile * kTN + j, acc[tp][j]);
    }
  }
}

template <int TP>
__global__ void __launch_bounds__(512) vit_cond_kernel(const VitArgs a) {
  extern __shared__ float4 sm4[];
  const cfpp_vit_desc& d = a.d;
  const int PS = a.PS, T = d.T, NPT = a.NPT, ntok = d.n_tok;
  float* X = reinterpret_cast<float*>(sm4);            // [T][PS]   residual stream
  float* Y = X + (int64_t)T * PS;                      // [ybuf_rows][PS]  LN output / attention output / patches
  float* Q = Y + (int64_t)a.ybuf_rows * PS;            // [qbuf_rows][PS]  q|k|v, later the MLP hidden
  float* Wbuf = Q + (int64_t)a.qbuf_rows * PS;
  const int b0 = blockIdx.x * a.S;
  const int nS = min(a.S, a.B - b0);
  const int R = nS * ntok;
  const int HW = d.H * d.W, tw = d.W / d.p2, Ctot = d.Cin + a.Cextra;

  // ---- patchify 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' into Y[f][r] ----
  for (int idx = threadIdx.x; idx < d.patch_dim * PS; idx += blockDim.x) {
    const int f = idx / PS, r = idx % PS;
    float v = 0.f;
    if (r < R) {
      const int s = r / ntok, tok = r % ntok, th = tok / tw, tww = tok % tw;
      const int c = f % Ctot, pp = f / Ctot, i = pp / d.p2, j = pp % d.p2;
      if (c < d.Cin) v = a.x[(int64_t)(b0 + s) * a.x_bstride + (int64_t)c * HW + (th * d.p1 + i) * d.W + (tww * d.p2 + j)];
      else v = a.extra[(int64_t)(b0 + s) * a.Cextra + (c - d.Cin)];
    }
    Y[idx] = v;
  }
  __syncthreads();
  layer_norm_rows(Y, Y, d.patch_dim, PS, R, d.ln0_w, d.ln0_b);
  __syncthreads();
  gemm_rounds<TP>(Y, PS, d.patch_dim, d.pe_wt, NPT, Wbuf, [&](int p, int n, float v) { if (n < T) X[n * PS + p] = v + d.pe_b[n]; });
  __syncthreads();
  layer_norm_rows(X, X, T, PS, R, d.ln1_w, d.ln1_b);
  __syncthreads();
  for (int idx = threadIdx.x; idx < T * PS; idx += blockDim.x) {       // x += pos_embedding (simple_vit.py:122)
    const int f = idx / PS, r = idx % PS;
    if (r < R) X[idx] += d.pos[(r % ntok) * T + f];
  }
  __syncthreads();

  const int64_t lstride = vit_layer_floats(T);
  for (int l = 0; l < d.depth; ++l) {
    const float* Lp = d.layers + l * lstride;
    const float* lna_w = Lp; const float* lna_b = Lp + T;
    const float* wqkv = Lp + 2 * T;
    const float* wo = wqkv + (int64_t)T * kQKV;
    const float* lnf_w = wo + (int64_t)kDH * NPT; const float* lnf_b = lnf_w + T;
    const float* w1 = lnf_b + T; const float* b1 = w1 + (int64_t)T * NPT;
    const float* w2 = b1 + NPT; const float* b2 = w2 + (int64_t)T * NPT;

    // ---- attention block: x += Wo softmax(q k^T / 8) v,  q,k,v = chunk(Wqkv LN(x)) ----
    layer_norm_rows(X, Y, T, PS, R, lna_w, lna_b);
    __syncthreads();
    gemm_rounds<TP>(Y, PS, T, wqkv, kQKV, Wbuf, [&](int p, int n, float v) { Q[n * PS + p] = v; });
    __syncthreads();
    for (int base = 0; base < R; base += blockDim.x / 4) {            // 4 lanes per query row, 16 head dims each
      const int r = base + threadIdx.x / 4, q4 = threadIdx.x & 3;
      const bool ok = r < R;
      const int rr = ok ? r : 0;
      const int r0 = (rr / ntok) * ntok;
      float qv[16], o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) { qv[i] = Q[(q4 * 16 + i) * PS + rr]; o[i] = 0.f; }
      float mx = -INFINITY, den = 0.f;
      for (int j = 0; j < ntok; ++j) {
        const int rj = r0 + j;
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) dot = fmaf(qv[i], Q[(kDH + q4 * 16 + i) * PS + rj], dot);
        dot += __shfl_xor_sync(0xffffffffu, dot, 1); dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        dot *= 0.125f;                                                 // dim_head ** -0.5
        const float nm = fmaxf(mx, dot), corr = expf(mx - nm), pj = expf(dot - nm);
        den = den * corr + pj;
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = fmaf(o[i], corr, pj * Q[(2 * kDH + q4 * 16 + i) * PS + rj]);
        mx = nm;
      }
      const float inv = 1.0f / den;
      if (ok)
#pragma unroll
        for (int i = 0; i < 16; ++i) Y[(q4 * 16 + i) * PS + r] = o[i] * inv;
    }
    __syncthreads();
    gemm_rounds<TP>(Y, PS, kDH, wo, NPT, Wbuf, [&](int p, int n, float v) { if (n < T) X[n * PS + p] += v; });
    __syncthreads();
    // ---- MLP block: x += W2 gelu(W1 LN(x) + b1) + b2 ----
    layer_norm_rows(X, Y, T, PS, R, lnf_w, lnf_b);
    __syncthreads();
    gemm_rounds<TP>(Y, PS, T, w1, NPT, Wbuf, [&](int p, int n, float v) {
      if (n < T) { const float u = v + b1[n]; Q[n * PS + p] = 0.5f * u * (1.0f + erff(u * 0.70710678118654752440f)); }
    });
    __syncthreads();
    gemm_rounds<TP>(Q, PS, T, w2, NPT, Wbuf, [&](int p, int n, float v) { if (n < T) X[n * PS + p] += v + b2[n]; });
    __syncthreads();
  }
  layer_norm_rows(X, X, T, PS, R, d.lnf_w, d.lnf_b);
  __syncthreads();
  // ---- un-patchify 'b (h w) (p1 p2 c) -> b c (h p1) (w p2)', c = T / (p1 p2) ----
  const int Cout = T / (d.p1 * d.p2);
  for (int idx = threadIdx.x; idx < T * PS; idx += blockDim.x) {
    const int f = idx / PS, r = idx % PS;
    if (r >= R) continue;
    const int s = r / ntok, tok = r % ntok, th = tok / tw, tww = tok % tw;
    const int c = f % Cout, pp = f / Cout, i = pp / d.p2, j = pp % d.p2;
    a.h[((int64_t)(b0 + s) * Cout + c) * HW + (th * d.p1 + i) * d.W + (tww * d.p2 + j)] = X[idx];
  }
}

template <int TP>
static int launch_vit(const VitArgs& a, size_t smem, cudaStream_t st) {
  static DeviceOnce attr_set;
  if (attr_set.first()) { cudaFuncSetAttribute(vit_cond_kernel<TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }
  vit_cond_kernel<TP><<<(a.B + a.S - 1) / a.S, 512, smem, st>>>(a);
  return check_launch("vit_cond_fwd");
}

}  // namespace cfpp
using namespace cfpp;

extern "C" int64_t cfpp_vit_layer_floats(int T) { return vit_layer_floats(T); }

extern "C" int cfpp_vit_cond_fwd(const float* x, int64_t x_bstride, const float* extra, int Cextra, float* h,
                                 const cfpp_vit_desc* desc, int B, void* stream) {
  CFPP_REQUIRE(desc != nullptr, "vit_cond: null descriptor");
  const cfpp_vit_desc& d = *desc;
  CFPP_REQUIRE(d.p1 >= 1 && d.p2 >= 1 && d.H % d.p1 == 0 && d.W % d.p2 == 0, "vit_cond: image %dx%d not divisible by patch", d.H, d.W);
  CFPP_REQUIRE(d.n_tok == (d.H / d.p1) * (d.W / d.p2) && d.patch_dim == (d.Cin + Cextra) * d.p1 * d.p2, "vit_cond: inconsistent descriptor");
  CFPP_REQUIRE(d.T >= 4 && d.T % 4 == 0 && d.T <= 256 && d.T % (d.p1 * d.p2) == 0, "vit_cond: T=%d unsupported", d.T);
  CFPP_REQUIRE(d.n_tok <= 128, "vit_cond: %d tokens per sample exceed one CTA tile", d.n_tok);
  CFPP_REQUIRE(Cextra == 0 || extra != nullptr, "vit_cond: extra channels pointer missing");
  if (B <= 0) return CFPP_OK;
  VitArgs a{x, x_bstride, extra, Cextra, h, d, B, 1, 0, (d.T + 15) / 16 * 16, 0, 0};
  a.ybuf_rows = d.T > d.patch_dim ? d.T : d.patch_dim; if (a.ybuf_rows < kDH) a.ybuf_rows = kDH;
  a.qbuf_rows = d.T > kQKV ? d.T : kQKV;
  const int NPmax = a.NPT > kQKV ? a.NPT : kQKV;
  // row-tile candidates: 128 (TP=4), 96 (TP=3), 64 (TP=2); pick the fullest tile that fits in shared memory
  int best_ps = 0; double best_util = -1.0; int best_S = 1;
  for (int ps = 128; ps >= 64; ps -= 32) {
    if (ps < d.n_tok) continue;
    const size_t smem = ((size_t)(d.T + a.ybuf_rows + a.qbuf_rows) * ps + (size_t)2 * kVitCC * NPmax) * sizeof(float);
    if (smem > 220 * 1024) continue;
    int S = ps / d.n_tok; if (S > B) S = B;
    const double util = (double)(S * d.n_tok) / ps + 1e-3 * ps / 128.0;
    if (util > best_util) { best_util = util; best_ps = ps; best_S = S; }
  }
  CFPP_REQUIRE(best_ps > 0, "vit_cond: T=%d with %d tokens does not fit one CTA", d.T, d.n_tok);
  a.PS = best_ps; a.S = best_S;
  const size_t smem = ((size_t)(d.T + a.ybuf_rows + a.qbuf_rows) * a.PS + (size_t)2 * kVitCC * NPmax) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (a.PS == 128) return launch_vit<4>(a, smem, st);
  if (a.PS == 96) return launch_vit<3>(a, smem, st);
  return launch_vit<2>(a, smem, st);
}
