// Training direction of the SimpleViT conditioner (reference layers/simple_vit.py:30-127, used by TransCoupling, coupling.py:100-148):
// an activation-saving forward and the backward, as plain row-major kernels over token rows X (R = B * n_tok rows, F features).
// What torch autograd derives for the reference's nn.LayerNorm / nn.Linear / softmax attention / nn.GELU stack.
// First version: FP32 CUDA cores, one launch per op; every reduction over rows is two-stage in a fixed order (deterministic).
#include <math.h>
#include "common.cuh"

namespace cfpp {
namespace vt {

constexpr int kMaxF = 320;          // widest row a warp keeps in registers (10 values per lane): 2C of the widest ATM coupling
constexpr int kDh = 64;             // attention head width (simple_vit.py: dim_head = 64, heads = 1)

// ---- patchify 'b c (h p1) (w p2) -> (b h w) (p1 p2 c)' and its inverse (simple_vit.py:102,115 / coupling un-patchify) ----------
__global__ void patchify_kernel(const float* __restrict__ x, int64_t bstride, float* __restrict__ tok, int64_t total,
                                int c, int H, int W, int p1, int p2) {
  const int wt = W / p2, n = (H / p1) * wt, PD = p1 * p2 * c;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(idx % PD); const int64_t r = idx / PD;
    const int t = (int)(r % n); const int64_t b = r / n;
    const int ch = f % c, ij = f / c, i = ij / p2, j = ij - i * p2, hh = t / wt, ww = t - hh * wt;
    tok[idx] = x[b * bstride + ((int64_t)ch * H + hh * p1 + i) * W + ww * p2 + j];
  }
}
__global__ void unpatchify_kernel(const float* __restrict__ tok, float* __restrict__ x, int64_t bstride, int64_t total,
                                  int c, int H, int W, int p1, int p2, int accumulate) {
  const int wt = W / p2, n = (H / p1) * wt, PD = p1 * p2 * c;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(idx % PD); const int64_t r = idx / PD;
    const int t = (int)(r % n); const int64_t b = r / n;
    const int ch = f % c, ij = f / c, i = ij / p2, j = ij - i * p2, hh = t / wt, ww = t - hh * wt;
    float* o = x + b * bstride + ((int64_t)ch * H + hh * p1 + i) * W + ww * p2 + j;
    *o = accumulate ? *o + tok[idx] : tok[idx];
  }
}

// ---- LayerNorm over the feature axis, eps = 1e-5, biased variance (nn.LayerNorm); warp per row ------------------------------------
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int64_t R, int F) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= R) return;
  const float* xr = x + row * F;
  float v[kMaxF / 32];
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < kMaxF / 32; ++q) { const int f = lane + 32 * q; v[q] = f < F ? xr[f] : 0.f; s += v[q]; }
  const float m = warp_sum(s) / (float)F;
  float ss = 0.f;
#pragma unroll
  for (int q = 0; q < kMaxF / 32; ++q) { const int f = lane + 32 * q; const float d = f < F ? v[q] - m : 0.f; ss += d * d; }
  const float rs = 1.0f / sqrtf(warp_sum(ss) / (float)F + 1e-5f);
#pragma unroll
  for (int q = 0; q < kMaxF / 32; ++q) { const int f = lane + 32 * q; if (f < F) y[row * F + f] = (v[q] - m) * rs * gamma[f] + beta[f]; }
  if (lane == 0) { mean[row] = m; rstd[row] = rs; }
}

// dx = rstd (dy g - mean(dy g) - xhat mean(dy g xhat)); per-CTA partial dgamma = sum dy xhat, dbeta = sum dy -> part[cta][2][F]
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     float* __restrict__ dx, float* __restrict__ part, int64_t R, int F) {
  __shared__ float red[8][2][kMaxF];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float dg[kMaxF / 32], db[kMaxF / 32];
#pragma unroll
  for (int q = 0; q < kMaxF / 32; ++q) { dg[q] = 0.f; db[q] = 0.f; }
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < R; row += (int64_t)gridDim.x * 8) {
    const float m = mean[row], rs = rstd[row];
    float xh[kMaxF / 32], g[kMaxF / 32];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxF / 32; ++q) {
      const int f = lane + 32 * q;
      if (f < F) {
        const float d = dy[row * F + f];
        xh[q] = (x[row * F + f] - m) * rs; g[q] = d * gamma[f];
        dg[q] += d * xh[q]; db[q] += d;
        s1 += g[q]; s2 += g[q] * xh[q];
      } else { xh[q] = 0.f; g[q] = 0.f; }
    }
    s1 = warp_sum(s1) / (float)F; s2 = warp_sum(s2) / (float)F;
#pragma unroll
    for (int q = 0; q < kMaxF / 32; ++q) { const int f = lane + 32 * q; if (f < F) dx[row * F + f] = rs * (g[q] - s1 - xh[q] * s2); }
  }
#pragma unroll
  for (int q = 0; q < kMaxF / 32; ++q) { red[warp][0][lane + 32 * q] = dg[q]; red[warp][1][lane + 32 * q] = db[q]; }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * F; i += blockDim.x) {
    const int which = i / F, f = i - which * F;
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][which][f];
    part[((int64_t)blockIdx.x * 2 + which) * F + f] = s;
  }
}

// one warp per output: lane-strided partial sums, then the shuffle tree -- a fixed order, so still deterministic
__global__ void __launch_bounds__(256) ln_finish_kernel(const float* __restrict__ part, float* __restrict__ dgamma, float* __restrict__ dbeta, int F, int ctas) {
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < 2 * F; i += gridDim.x * 8) {
    const int which = i / F, f = i - which * F;
    float s = 0.f;
    for (int c = lane; c < ctas; c += 32) s += part[((int64_t)c * 2 + which) * F + f];
    s = warp_sum(s);
    if (lane == 0) (which ? dbeta : dgamma)[f] = s;
  }
}

// out[i] = sum_c part[c][i]: a CTA covers 32 consecutive outputs; warp w sums chunks w, w+8, ... (coalesced rows of 32 floats), the
// eight warp partials are added in warp order -- a fixed order, so still deterministic
__global__ void __launch_bounds__(256) chunk_sum_kernel(const float* __restrict__ part, float* __restrict__ out, int64_t n, int chunks) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int64_t i0 = (int64_t)blockIdx.x * 32; i0 < n; i0 += (int64_t)gridDim.x * 32) {
    const int64_t i = i0 + lane;
    float s = 0.f;
    if (i < n) for (int c = w; c < chunks; c += 8) s += part[(int64_t)c * n + i];
    red[w][lane] = s;
    __syncthreads();
    if (w == 0 && i < n) { float t = 0.f; for (int q = 0; q < 8; ++q) t += red[q][lane]; out[i] = t; }
    __syncthreads();
  }
}

// ---- row GEMMs.  TRANS = 0: OUT[r][j] = sum_i IN[r][i] W[j][i] (+ b[j])   (nn.Linear forward, W is (J, I))
//                  TRANS = 1: OUT[r][j] = sum_i IN[r][i] W[i][j]            (its input gradient, W is (I, J))
// CTA = 64 rows x 64 output columns; IN rows and the weight tile (i-major) live in shared memory; thread = 4 rows x 4 columns:
// per i, four broadcast loads of x and one 128-bit load of four weights feed 16 FMAs.
constexpr int kWS = 68;             // padded width of the weight tile rows (floats): 16-byte aligned, conflict-free 128-bit reads
template <int TRANS>
__global__ void __launch_bounds__(256) rows_linear_kernel(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                                                          float* __restrict__ out, int64_t R, int I, int J, int accumulate, int64_t in_stride) {
  extern __shared__ float4 sm4[];
  float* ws = reinterpret_cast<float*>(sm4);          // [I][kWS]
  float* xs = ws + (size_t)I * kWS;                   // [64][I | 1]
  const int S = I | 1;
  const int64_t r0 = (int64_t)blockIdx.x * 64;
  const int j0 = blockIdx.y * 64;
  const int rows = (int)min((int64_t)64, R - r0), cols = min(64, J - j0);
  for (int i = threadIdx.x; i < 64 * I; i += blockDim.x) { const int r = i / I, k = i - r * I; xs[r * S + k] = r < rows ? in[(r0 + r) * in_stride + k] : 0.f; }
  if (TRANS) {
    for (int e = threadIdx.x; e < I * 64; e += blockDim.x) { const int i = e >> 6, jj = e & 63; ws[i * kWS + jj] = jj < cols ? W[(int64_t)i * J + j0 + jj] : 0.f; }
  } else {
    for (int e = threadIdx.x; e < I * 64; e += blockDim.x) { const int jj = e / I, i = e - jj * I; ws[i * kWS + jj] = jj < cols ? W[(int64_t)(j0 + jj) * I + i] : 0.f; }
  }
  __syncthreads();
  const int tr = threadIdx.x >> 4, tc = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = (bias && 4 * tc + c < cols) ? bias[j0 + 4 * tc + c] : 0.f;
  const float* x0 = xs + (4 * tr) * S;
#pragma unroll 4
  for (int i = 0; i < I; ++i) {
    const float4 w = *reinterpret_cast<const float4*>(ws + i * kWS + 4 * tc);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float xv = x0[a * S + i];
      acc[a][0] = fmaf(xv, w.x, acc[a][0]); acc[a][1] = fmaf(xv, w.y, acc[a][1]);
      acc[a][2] = fmaf(xv, w.z, acc[a][2]); acc[a][3] = fmaf(xv, w.w, acc[a][3]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int r = 4 * tr + a;
    if (r >= rows) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = 4 * tc + c;
      if (j < cols) { float* o = out + (r0 + r) * J + j0 + j; *o = accumulate ? *o + acc[a][c] : acc[a][c]; }
    }
  }
}

// dW[j][i] = sum_r dY[r][j] X[r][i], db[j] = sum_r dY[r][j] over a chunk of rows -> part[chunk][J*I], partb[chunk][J].
// grid (chunks, J tiles x I tiles of 64 x 64); thread = 4 output rows j x 4 input columns i; per row of the chunk two 128-bit shared
// loads feed 16 FMAs.
__global__ void __launch_bounds__(256) rows_wgrad_kernel(const float* __restrict__ X, const float* __restrict__ dY, float* __restrict__ part,
                                                         float* __restrict__ partb, int64_t R, int I, int J, int64_t per_chunk, int itiles) {
  __shared__ float4 sx4[32 * kWS / 4];
  __shared__ float4 sd4[32 * kWS / 4];
  float* sx = reinterpret_cast<float*>(sx4); float* sd = reinterpret_cast<float*>(sd4);
  const int jt = blockIdx.y / itiles, it = blockIdx.y - jt * itiles;
  const int j0 = jt * 64, i0 = it * 64;
  const int tj = threadIdx.x >> 4, ti = threadIdx.x & 15;
  const int64_t rbeg = (int64_t)blockIdx.x * per_chunk, rend = min(R, rbeg + per_chunk);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t rt = rbeg; rt < rend; rt += 32) {
    const int rows = (int)min((int64_t)32, rend - rt);
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * 64; e += blockDim.x) {
      const int r = e >> 6, c = e & 63;
      sx[r * kWS + c] = (r < rows && i0 + c < I) ? X[(rt + r) * I + i0 + c] : 0.f;
      sd[r * kWS + c] = (r < rows && j0 + c < J) ? dY[(rt + r) * J + j0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const float4 g = *reinterpret_cast<const float4*>(sd + r * kWS + 4 * tj);
      const float4 x = *reinterpret_cast<const float4*>(sx + r * kWS + 4 * ti);
      const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        acc[a][0] = fmaf(gv[a], x.x, acc[a][0]); acc[a][1] = fmaf(gv[a], x.y, acc[a][1]);
        acc[a][2] = fmaf(gv[a], x.z, acc[a][2]); acc[a][3] = fmaf(gv[a], x.w, acc[a][3]);
        bacc[a] += gv[a];
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int j = j0 + 4 * tj + a;
    if (j >= J) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) { const int i = i0 + 4 * ti + c; if (i < I) part[((int64_t)blockIdx.x * J + j) * I + i] = acc[a][c]; }
    if (ti == 0 && it == 0 && partb) partb[(int64_t)blockIdx.x * J + j] = bacc[a];
  }
}

// ---- GELU (erf form, nn.GELU()) ------------------------------------------------------------------------------------------------------
__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
  }
}
__global__ void gelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752f));
    const float pdf = 0.39894228040143268f * expf(-0.5f * v * v);
    dx[i] = dy[i] * (cdf + v * pdf);
  }
}

// X[r][f] += pos[r % n][f]   (simple_vit.py:121)
__global__ void add_pos_kernel(float* __restrict__ x, const float* __restrict__ pos, int64_t total, int n, int F) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % F); const int t = (int)((i / F) % n);
    x[i] += pos[t * F + f];
  }
}

// ---- single-head attention per sample: qkv rows (n, 192) = [q | k | v]; P = softmax(q k^T / 8); O = P v ----------------------------
__global__ void __launch_bounds__(128) attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ O, float* __restrict__ P, int n) {
  extern __shared__ float sm[];
  float* q = sm; float* k = q + n * (kDh + 1); float* v = k + n * (kDh + 1); float* s = v + n * (kDh + 1);   // s: n * n
  const int64_t b = blockIdx.x;
  const float* base = qkv + b * n * 3 * kDh;
  for (int i = threadIdx.x; i < n * 3 * kDh; i += blockDim.x) {
    const int t = i / (3 * kDh), f = i - t * 3 * kDh, which = f / kDh, d = f - which * kDh;
    (which == 0 ? q : which == 1 ? k : v)[t * (kDh + 1) + d] = base[i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    float acc = 0.f;
    for (int d = 0; d < kDh; ++d) acc = fmaf(q[i * (kDh + 1) + d], k[j * (kDh + 1) + d], acc);
    s[e] = acc * 0.125f;                                   // dim_head ** -0.5
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float mx = -INFINITY;
    for (int j = 0; j < n; ++j) mx = fmaxf(mx, s[i * n + j]);
    float se = 0.f;
    for (int j = 0; j < n; ++j) { const float e = expf(s[i * n + j] - mx); s[i * n + j] = e; se += e; }
    const float inv = 1.0f / se;
    for (int j = 0; j < n; ++j) { s[i * n + j] *= inv; P[(b * n + i) * n + j] = s[i * n + j]; }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < n * kDh; e += blockDim.x) {
    const int i = e / kDh, d = e - i * kDh;
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc = fmaf(s[i * n + j], v[j * (kDh + 1) + d], acc);
    O[(b * n + i) * kDh + d] = acc;
  }
}

// dV = P^T dO; dP = dO V^T; dS = P (dP - rowsum(dP P)); dq = dS k / 8; dk = dS^T q / 8
__global__ void __launch_bounds__(128) attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ P, const float* __restrict__ dO,
                                                       float* __restrict__ dqkv, int n) {
  extern __shared__ float sm[];
  float* q = sm; float* k = q + n * (kDh + 1); float* v = k + n * (kDh + 1); float* go = v + n * (kDh + 1);
  float* p = go + n * (kDh + 1); float* ds = p + n * n;
  const int64_t b = blockIdx.x;
  const float* base = qkv + b * n * 3 * kDh;
  for (int i = threadIdx.x; i < n * 3 * kDh; i += blockDim.x) {
    const int t = i / (3 * kDh), f = i - t * 3 * kDh, which = f / kDh, d = f - which * kDh;
    (which == 0 ? q : which == 1 ? k : v)[t * (kDh + 1) + d] = base[i];
  }
  for (int i = threadIdx.x; i < n * kDh; i += blockDim.x) go[(i / kDh) * (kDh + 1) + (i % kDh)] = dO[b * n * kDh + i];
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) p[i] = P[b * n * n + i];
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {             // dP
    const int i = e / n, j = e - i * n;
    float acc = 0.f;
    for (int d = 0; d < kDh; ++d) acc = fmaf(go[i * (kDh + 1) + d], v[j * (kDh + 1) + d], acc);
    ds[e] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {                 // dS rows
    float dot = 0.f;
    for (int j = 0; j < n; ++j) dot = fmaf(ds[i * n + j], p[i * n + j], dot);
    for (int j = 0; j < n; ++j) ds[i * n + j] = p[i * n + j] * (ds[i * n + j] - dot);
  }
  __syncthreads();
  float* out = dqkv + b * n * 3 * kDh;
  for (int e = threadIdx.x; e < n * kDh; e += blockDim.x) {
    const int t = e / kDh, d = e - t * kDh;
    float aq = 0.f, ak = 0.f, av = 0.f;
    for (int j = 0; j < n; ++j) {
      aq = fmaf(ds[t * n + j], k[j * (kDh + 1) + d], aq);            // dq[t] = sum_j dS[t][j] k[j]
      ak = fmaf(ds[j * n + t], q[j * (kDh + 1) + d], ak);            // dk[t] = sum_i dS[i][t] q[i]
      av = fmaf(p[j * n + t], go[j * (kDh + 1) + d], av);            // dv[t] = sum_i P[i][t] dO[i]
    }
    out[t * 3 * kDh + d] = aq * 0.125f;
    out[t * 3 * kDh + kDh + d] = ak * 0.125f;
    out[t * 3 * kDh + 2 * kDh + d] = av;
  }
}

inline int grid1d(int64_t n, int per_thread = 1) {
  int64_t blocks = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}
inline int ln_bwd_ctas(int64_t R) { int64_t c = (R + 7) / 8; const int64_t cap = (int64_t)num_sms(); return (int)(c > cap ? cap : (c < 1 ? 1 : c)); }
inline void wgrad_plan(int64_t R, int I, int J, int& chunks, int64_t& per_chunk) {
  const int jb = ((J + 63) / 64) * ((I + 63) / 64);
  int64_t c = ((int64_t)num_sms() * 2 + jb - 1) / jb;
  const int64_t maxc = (R + 31) / 32;
  if (c > maxc) c = maxc; if (c < 1) c = 1;
  per_chunk = ((R + c - 1) / c + 31) / 32 * 32;
  chunks = (int)((R + per_chunk - 1) / per_chunk); if (chunks < 1) chunks = 1;
}

}  // namespace vt
}  // namespace cfpp
using namespace cfpp;
using namespace cfpp::vt;

extern "C" int cfpp_patchify_fwd(const float* x, int64_t x_bstride, float* tok, int B, int c, int H, int W, int p1, int p2, void* stream) {
  CFPP_REQUIRE(c >= 1 && p1 >= 1 && p2 >= 1 && H % p1 == 0 && W % p2 == 0, "patchify: c=%d H=%d W=%d p=(%d,%d)", c, H, W, p1, p2);
  const int64_t total = (int64_t)B * c * H * W;
  if (total <= 0) return CFPP_OK;
  patchify_kernel<<<grid1d(total, 2), 256, 0, (cudaStream_t)stream>>>(x, x_bstride, tok, total, c, H, W, p1, p2);
  return check_launch("patchify_fwd");
}

extern "C" int cfpp_patchify_inv(const float* tok, float* x, int64_t x_bstride, int accumulate, int B, int c, int H, int W, int p1, int p2, void* stream) {
  CFPP_REQUIRE(c >= 1 && p1 >= 1 && p2 >= 1 && H % p1 == 0 && W % p2 == 0, "patchify_inv: c=%d H=%d W=%d p=(%d,%d)", c, H, W, p1, p2);
  const int64_t total = (int64_t)B * c * H * W;
  if (total <= 0) return CFPP_OK;
  unpatchify_kernel<<<grid1d(total, 2), 256, 0, (cudaStream_t)stream>>>(tok, x, x_bstride, total, c, H, W, p1, p2, accumulate);
  return check_launch("patchify_inv");
}

extern "C" int cfpp_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                                  int64_t R, int F, void* stream) {
  CFPP_REQUIRE(F >= 1 && F <= kMaxF, "layernorm: F=%d outside [1,%d]", F, kMaxF);
  if (R <= 0) return CFPP_OK;
  ln_fwd_kernel<<<(unsigned)((R + 7) / 8), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, mean, rstd, R, F);
  return check_launch("layernorm_fwd");
}

extern "C" int64_t cfpp_layernorm_bwd_workspace_floats(int64_t R, int F) { return (int64_t)ln_bwd_ctas(R) * 2 * F; }

extern "C" int cfpp_layernorm_bwd(const float* x, const float* dy, const float* gamma, const float* mean, const float* rstd,
                                  float* dx, float* dgamma, float* dbeta, float* workspace, int64_t R, int F, void* stream) {
  CFPP_REQUIRE(F >= 1 && F <= kMaxF && R >= 1 && workspace, "layernorm_bwd: R=%lld F=%d", (long long)R, F);
  const int ctas = ln_bwd_ctas(R);
  ln_bwd_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(x, dy, gamma, mean, rstd, dx, workspace, R, F);
  int rc = check_launch("layernorm_bwd");
  if (rc != CFPP_OK) return rc;
  ln_finish_kernel<<<(2 * F + 7) / 8, 256, 0, (cudaStream_t)stream>>>(workspace, dgamma, dbeta, F, ctas);
  return check_launch("layernorm_bwd_finish");
}

extern "C" int cfpp_rows_linear_fwd(const float* x, const float* W, const float* bias, float* y, int64_t R, int I, int J, void* stream) {
  CFPP_REQUIRE(I >= 1 && I <= kMaxF && J >= 1, "rows_linear: I=%d J=%d", I, J);
  if (R <= 0) return CFPP_OK;
  const size_t smem = ((size_t)64 * (I | 1) + (size_t)I * kWS) * sizeof(float);
  static DeviceOnce a;
  if (a.first()) { cudaFuncSetAttribute(rows_linear_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (64 * (kMaxF + 1) + kMaxF * kWS) * 4); }
  rows_linear_kernel<0><<<dim3((unsigned)((R + 63) / 64), (J + 63) / 64), 256, smem, (cudaStream_t)stream>>>(x, W, bias, y, R, I, J, 0, I);
  return check_launch("rows_linear_fwd");
}

extern "C" int cfpp_rows_linear_bwd_data(const float* dy, const float* W, float* dx, int accumulate, int64_t R, int I, int J, void* stream) {
  // W is the forward weight (J, I); dx[r][i] = sum_j dy[r][j] W[j][i]
  CFPP_REQUIRE(J >= 1 && I >= 1, "rows_linear_bwd_data: I=%d J=%d", I, J);
  if (R <= 0) return CFPP_OK;
  static DeviceOnce a;
  if (a.first()) { cudaFuncSetAttribute(rows_linear_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (64 * (kMaxF + 1) + kMaxF * kWS) * 4); }
  for (int j0 = 0; j0 < J; j0 += kMaxF) {             // wide outputs (Conv1x1's CN: D*D columns): reduce in slices of 256, accumulating
    const int len = J - j0 < kMaxF ? J - j0 : kMaxF;
    const size_t smem = ((size_t)64 * (len | 1) + (size_t)len * kWS) * sizeof(float);
    rows_linear_kernel<1><<<dim3((unsigned)((R + 63) / 64), (I + 63) / 64), 256, smem, (cudaStream_t)stream>>>(
        dy + j0, W + (int64_t)j0 * I, nullptr, dx, R, len, I, accumulate || j0 > 0, J);
    const int rc = check_launch("rows_linear_bwd_data");
    if (rc != CFPP_OK) return rc;
  }
  return CFPP_OK;
}

extern "C" int64_t cfpp_rows_linear_bwd_weight_workspace_floats(int64_t R, int I, int J) {
  int chunks; int64_t per_chunk;
  wgrad_plan(R, I, J, chunks, per_chunk);
  return (int64_t)chunks * ((int64_t)J * I + J);
}

extern "C" int cfpp_rows_linear_bwd_weight(const float* x, const float* dy, float* dW, float* db, float* workspace,
                                           int64_t R, int I, int J, void* stream) {
  CFPP_REQUIRE(I >= 1 && I <= kMaxF && J >= 1 && R >= 1 && workspace, "rows_linear_bwd_weight: R=%lld I=%d J=%d", (long long)R, I, J);
  int chunks; int64_t per_chunk;
  wgrad_plan(R, I, J, chunks, per_chunk);
  float* part = workspace; float* partb = workspace + (int64_t)chunks * J * I;
  const int itiles = (I + 63) / 64;
  rows_wgrad_kernel<<<dim3(chunks, ((J + 63) / 64) * itiles), 256, 0, (cudaStream_t)stream>>>(x, dy, part, db ? partb : nullptr, R, I, J, per_chunk, itiles);
  int rc = check_launch("rows_linear_bwd_weight");
  if (rc != CFPP_OK) return rc;
  chunk_sum_kernel<<<(unsigned)(((int64_t)J * I + 31) / 32), 256, 0, (cudaStream_t)stream>>>(part, dW, (int64_t)J * I, chunks);
  if ((rc = check_launch("rows_linear_bwd_weight_sum")) != CFPP_OK) return rc;
  if (db) {
    chunk_sum_kernel<<<(J + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partb, db, J, chunks);
    rc = check_launch("rows_linear_bwd_bias_sum");
  }
  return rc;
}

extern "C" int cfpp_gelu_fwd(const float* x, float* y, int64_t n, void* stream) {
  if (n <= 0) return CFPP_OK;
  gelu_fwd_kernel<<<grid1d(n, 4), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  return check_launch("gelu_fwd");
}

extern "C" int cfpp_gelu_bwd(const float* x, const float* dy, float* dx, int64_t n, void* stream) {
  if (n <= 0) return CFPP_OK;
  gelu_bwd_kernel<<<grid1d(n, 4), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, n);
  return check_launch("gelu_bwd");
}

extern "C" int cfpp_add_pos(float* x, const float* pos, int64_t R, int n_tok, int F, void* stream) {
  if (R <= 0) return CFPP_OK;
  add_pos_kernel<<<grid1d(R * F, 4), 256, 0, (cudaStream_t)stream>>>(x, pos, R * F, n_tok, F);
  return check_launch("add_pos");
}

static size_t attn_smem(int n, int mats, int sq) { return ((size_t)mats * n * (kDh + 1) + (size_t)sq * n * n) * sizeof(float); }

extern "C" int cfpp_attention_fwd(const float* qkv, float* O, float* P, int B, int n_tok, void* stream) {
  CFPP_REQUIRE(n_tok >= 1 && attn_smem(n_tok, 4, 2) <= 200 * 1024, "attention: %d tokens exceed shared memory", n_tok);
  if (B <= 0) return CFPP_OK;
  const size_t smem = attn_smem(n_tok, 3, 1);
  if (smem > 48 * 1024) cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  attn_fwd_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(qkv, O, P, n_tok);
  return check_launch("attention_fwd");
}

extern "C" int cfpp_attention_bwd(const float* qkv, const float* P, const float* dO, float* dqkv, int B, int n_tok, void* stream) {
  CFPP_REQUIRE(n_tok >= 1 && attn_smem(n_tok, 4, 2) <= 200 * 1024, "attention: %d tokens exceed shared memory", n_tok);
  if (B <= 0) return CFPP_OK;
  const size_t smem = attn_smem(n_tok, 4, 2);
  if (smem > 48 * 1024) cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  attn_bwd_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(qkv, P, dO, dqkv, n_tok);
  return check_launch("attention_bwd");
}
