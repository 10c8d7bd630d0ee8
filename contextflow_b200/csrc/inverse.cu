// Inverse (sampling) direction of the flow layers and the loss / score epilogue of the experiment loops.
// Reference: layers/coupling.py:68-73,150-155 (Coupling/TransCoupling.reverse), layers/actnorm.py:62-79 (ActNorm.reverse),
// layers/conv1x1.py:59-72 (Conv1x1.reverse: conv with torch.inverse(NN)), layers/transforms.py:14-15, layers/normalize.py:37-41,
// layers/dequantize.py:19-20, layers/augment.py:20-23, layers/distributions/gaussian.py:163-169 (mixture sampling),
// experiment_ad.py:204-211,262-281 and experiment_cl.py:127-133,185-204 (loss + score epilogue).
#include <math.h>
#include "common.cuh"

namespace cfpp {

// ---------------------------------------------------------------------------------------------------------------
// Coupling.reverse: x = cat(z0, (z1 - t) / s), s = exp(2 tanh(r/2)); same single HBM pass as the forward kernel
// (12*C*HW bytes per sample), G threads per sample, 128-bit streaming loads / stores.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float uncouple1(float z1, float t, float r) {
  const float s = expf(2.0f * tanhf(r * 0.5f));
  return (z1 - t) / s;                                   // coupling.py:71: true division by s, as the reference
}

template <bool VEC>
__global__ void __launch_bounds__(256) coupling_inv_kernel(const float* __restrict__ z, const float* __restrict__ h,
                                                           const float* __restrict__ add, float* __restrict__ x,
                                                           int B, int C, int HW, int G) {
  const int spc = blockDim.x / G;
  const int s = threadIdx.x / G, g = threadIdx.x % G;
  const int64_t b = (int64_t)blockIdx.x * spc + s;
  if (b >= B) return;
  const int Ch = C / 2;
  const int64_t n = (int64_t)Ch * HW;
  const float* zb = z + b * 2 * n; const float* hb = h + b * 2 * n; float* xb = x + b * 2 * n;
  const float* ab = add ? add + b * C : nullptr;
  if (VEC) {
    const int n4 = (int)(n / 4), hw4 = HW / 4;
    const float4* z0 = reinterpret_cast<const float4*>(zb); const float4* z1 = z0 + n4;
    const float4* ht = reinterpret_cast<const float4*>(hb); const float4* hr = ht + n4;
    float4* x0 = reinterpret_cast<float4*>(xb); float4* x1 = x0 + n4;
#pragma unroll 2
    for (int i = g; i < n4; i += G) {
      const float4 a0 = ldg_stream(z0 + i), a1 = ldg_stream(z1 + i);
      float4 t = ldg_stream(ht + i), r = ldg_stream(hr + i);
      if (ab) {
        const int ch = i / hw4;
        const float at = ab[ch], ar = ab[Ch + ch];
        t.x += at; t.y += at; t.z += at; t.w += at;
        r.x += ar; r.y += ar; r.z += ar; r.w += ar;
      }
      float4 o;
      o.x = uncouple1(a1.x, t.x, r.x); o.y = uncouple1(a1.y, t.y, r.y);
      o.z = uncouple1(a1.z, t.z, r.z); o.w = uncouple1(a1.w, t.w, r.w);
      stg_stream(x0 + i, a0);
      stg_stream(x1 + i, o);
    }
  } else {
    for (int64_t i = g; i < n; i += G) {
      float t = hb[i], r = hb[n + i];
      if (ab) { const int ch = (int)(i / HW); t += ab[ch]; r += ab[Ch + ch]; }
      xb[i] = zb[i];
      xb[n + i] = uncouple1(zb[n + i], t, r);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ActNorm.reverse (context-free form, actnorm.py:73-78): x = z * exp(logs[d]) + t[d].   8*D*HW bytes per sample.
// ---------------------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256) actnorm_inv_kernel(const float* __restrict__ z, const float* __restrict__ t,
                                                          const float* __restrict__ logs, float* __restrict__ x,
                                                          int64_t total, int D, int HW) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (VEC) {
    const int hw4 = HW / 4;
    const int64_t n4 = total / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const int d = (int)((i / hw4) % D);
      const float e = expf(__ldg(logs + d)), tt = __ldg(t + d);
      const float4 v = ldg_stream(reinterpret_cast<const float4*>(z) + i);
      float4 o;                                           // separate multiply and add: torch evaluates z*exp(logs) then + t
      o.x = __fadd_rn(__fmul_rn(v.x, e), tt); o.y = __fadd_rn(__fmul_rn(v.y, e), tt);
      o.z = __fadd_rn(__fmul_rn(v.z, e), tt); o.w = __fadd_rn(__fmul_rn(v.w, e), tt);
      stg_stream(reinterpret_cast<float4*>(x) + i, o);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const int d = (int)((i / HW) % D);
      x[i] = __fadd_rn(__fmul_rn(z[i], expf(__ldg(logs + d))), __ldg(t + d));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// torch.inverse(NN) (conv1x1.py:70): in-place Gauss-Jordan with partial pivoting in fp64, one CTA, D <= 128.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mat_inverse_kernel(const float* __restrict__ A, int D, float* __restrict__ Ainv,
                                                          int* __restrict__ singular) {
  extern __shared__ double a[];                 // D*D matrix, then col[D], rowk[D]
  double* col = a + (size_t)D * D;
  double* rowk = col + D;
  __shared__ int perm[128];
  __shared__ int piv_row;
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) a[i] = (double)A[i];
  __syncthreads();
  for (int k = 0; k < D; ++k) {
    if (threadIdx.x < 32) {                     // warp 0: arg-max |a[i][k]| over i >= k (first index on ties)
      double best = -1.0; int bi = k;
      for (int i = k + threadIdx.x; i < D; i += 32) { const double v = fabs(a[i * D + k]); if (v > best) { best = v; bi = i; } }
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (threadIdx.x == 0) { piv_row = bi; perm[k] = bi; if (best == 0.0) bad = 1; }
    }
    __syncthreads();
    const int pr = piv_row;
    if (pr != k)
      for (int j = threadIdx.x; j < D; j += blockDim.x) { const double t = a[k * D + j]; a[k * D + j] = a[pr * D + j]; a[pr * D + j] = t; }
    __syncthreads();
    const double pv = a[k * D + k];
    const double inv = pv != 0.0 ? 1.0 / pv : 0.0;
    __syncthreads();
    for (int j = threadIdx.x; j < D; j += blockDim.x) {   // scaled pivot row (unit entry in the pivot column) and the pivot column
      rowk[j] = (j == k ? 1.0 : a[k * D + j]) * inv;
      col[j] = a[j * D + k];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < D * D; e += blockDim.x) {
      const int i = e / D, j = e - i * D;
      if (i == k) a[e] = rowk[j];
      else a[e] = (j == k ? 0.0 : a[e]) - col[i] * rowk[j];
    }
    __syncthreads();
  }
  for (int k = D - 1; k >= 0; --k) {            // undo the row exchanges as column exchanges, in reverse order
    const int pr = perm[k];
    if (pr != k)
      for (int i = threadIdx.x; i < D; i += blockDim.x) { const double t = a[i * D + k]; a[i * D + k] = a[i * D + pr]; a[i * D + pr] = t; }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) Ainv[i] = (float)a[i];
  if (threadIdx.x == 0 && singular) singular[0] = bad;
}

// ---------------------------------------------------------------------------------------------------------------
// The image prologue backwards, one pass (model.py:97-100 read right to left, each layer's .reverse):
//   Augment.reverse (drop the A noise channels), LogitTransform.reverse (sigmoid), Normalization.reverse x2
//   ((v - t) * s), Dequantization.reverse (floor).   y_cont keeps the value before the floor (optional).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoid_f(float v) {     // torch.sigmoid in fp32
  return 1.0f / (1.0f + expf(-v));
}

__global__ void __launch_bounds__(256) prologue_inv_kernel(const float* __restrict__ z, float* __restrict__ x, float* __restrict__ x_cont,
                                                           int64_t total, int n_keep, int n_in,
                                                           float s1, float t1, float s0, float t0, int do_floor) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / n_keep; const int r = (int)(i - b * n_keep);
    float v = sigmoid_f(__ldg(z + b * n_in + r));
    v = __fmul_rn(__fsub_rn(v, t1), s1);
    v = __fmul_rn(__fsub_rn(v, t0), s0);
    if (x_cont) x_cont[i] = v;
    x[i] = do_floor ? floorf(v) : v;
  }
}

__global__ void __launch_bounds__(256) unary_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, int op,
                                                    float a, float b) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = in[i];
    out[i] = op == 0 ? sigmoid_f(v) : op == 1 ? __fmul_rn(__fsub_rn(v, b), a) : op == 2 ? floorf(v) : fmaxf(v, 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Mixture sampling core (gaussian.py:163-166 after the categorical / normal draws): x[b] = mG[m,k_b] + softplus(sG[m,k_b]) * eps[b].
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gmm_sample_kernel(const float* __restrict__ mG, const float* __restrict__ sG,
                                                         const int64_t* __restrict__ comp, const float* __restrict__ eps,
                                                         float* __restrict__ x, int64_t total, int n, int K, int m) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / n; const int r = (int)(i - b * n);
    int k = (int)comp[b]; k = k < 0 ? 0 : (k >= K ? K - 1 : k);
    const int64_t o = ((int64_t)m * K + k) * n + r;
    x[i] = fmaf(softplus_f(__ldg(sG + o)), eps[i], __ldg(mG + o));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Loss / score epilogue: one warp per sample over the M mixture columns.
//   s = dim_inv * logp, NaN -> 0 (experiment_ad.py:204-205); lse = logsumexp_m s; logsigmoid(lse) and logsigmoid(s) sums for
//   cost_uns (:207); weighted cross-entropy numerator / denominator for nn.CrossEntropyLoss(weight) (:208, model.py:294);
//   scores: softmax(s)[:,1] and s[:,-1] (:278), argmax_m s (experiment_cl.py:200).
// Sums are reduced per CTA in a fixed order, then by one finishing CTA in block order: deterministic, no atomics.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float logsigmoid_f(float v) {  // F.logsigmoid: min(v,0) - log1p(exp(-|v|))
  return fminf(v, 0.f) - log1pf(expf(-fabsf(v)));
}

struct ScoreArgs {
  const float* logp; float dim_inv; const int64_t* gt; const float* class_w;
  float* scaled; float* lse; float* softmax1; float* last; int64_t* argmax; double* partial;
  int B, M;
};

constexpr int kScoreWarps = 8;

__global__ void __launch_bounds__(kScoreWarps * 32) score_kernel(ScoreArgs a) {
  __shared__ double red[kScoreWarps][4];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * kScoreWarps + w;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};                    // sum logsigmoid(lse), sum logsigmoid(s), CE numerator, CE denominator
  if (b < a.B) {
    const float* row = a.logp + b * a.M;
    float mx = -INFINITY; int am = 0x7fffffff; float ls_sum = 0.f;
    for (int m = l; m < a.M; m += 32) {
      float s = a.dim_inv * row[m];
      if (s != s) s = 0.f;
      if (a.scaled) a.scaled[b * a.M + m] = s;
      ls_sum += logsigmoid_f(s);
      if (s > mx) { mx = s; am = m; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o); const int oi = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > mx || (om == mx && oi < am)) { mx = om; am = oi; }
    }
    float se = 0.f;
    for (int m = l; m < a.M; m += 32) {
      float s = a.dim_inv * row[m];
      if (s != s) s = 0.f;
      se += expf(s - mx);
    }
    if (am == 0x7fffffff) am = 0;                         // a row of -inf: torch.argmax returns 0
    se = warp_sum(se); ls_sum = warp_sum(ls_sum);
    const float lse = (mx == -INFINITY || mx == INFINITY) ? mx : mx + logf(se);   // torch.logsumexp keeps +-inf rows
    if (l == 0) {
      if (a.lse) a.lse[b] = lse;
      if (a.argmax) a.argmax[b] = am;
      float s1 = 0.f, sl = 0.f;
      if (a.M > 1) { s1 = a.dim_inv * row[1]; if (s1 != s1) s1 = 0.f; }
      sl = a.dim_inv * row[a.M - 1]; if (sl != sl) sl = 0.f;
      if (a.softmax1) a.softmax1[b] = a.M > 1 ? expf(s1 - mx) / se : 1.f;
      if (a.last) a.last[b] = sl;
      acc[0] = logsigmoid_f(lse); acc[1] = ls_sum;
      if (a.gt) {
        int64_t g = a.gt[b]; g = g < 0 ? 0 : (g >= a.M ? a.M - 1 : g);
        float sg = a.dim_inv * row[g]; if (sg != sg) sg = 0.f;
        const float wg = a.class_w ? a.class_w[g] : 1.f;
        acc[2] = (double)wg * (double)(lse - sg); acc[3] = wg;
      }
    }
  }
  if (l == 0) for (int q = 0; q < 4; ++q) red[w][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int i = 0; i < kScoreWarps; ++i) s += red[i][threadIdx.x];
    a.partial[(int64_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
}

// M <= 32: a CTA stages 256 rows through shared memory with coalesced 128-bit-free plain loads (rows are M floats, not vector
// aligned in general), each thread then owns one row; `scaled` goes back through the same tile, coalesced.
constexpr int kScoreRows = 256;

__global__ void __launch_bounds__(kScoreRows) score_rows_kernel(ScoreArgs a) {
  extern __shared__ float tile[];                          // kScoreRows * M
  __shared__ double red[kScoreRows / 32][4];
  const int M = a.M;
  const int64_t b0 = (int64_t)blockIdx.x * kScoreRows;
  const int rows = (int)min((int64_t)kScoreRows, (int64_t)a.B - b0);
  const int n = rows * M;
  const float* src = a.logp + b0 * M;
  for (int i = threadIdx.x; i < n; i += kScoreRows) {
    float s = a.dim_inv * __ldg(src + i);
    if (s != s) s = 0.f;
    tile[i] = s;
  }
  __syncthreads();
  if (a.scaled) for (int i = threadIdx.x; i < n; i += kScoreRows) a.scaled[b0 * M + i] = tile[i];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  if ((int)threadIdx.x < rows) {
    const int64_t b = b0 + threadIdx.x;
    const float* row = tile + threadIdx.x * M;
    float mx = row[0]; int am = 0; float ls_sum = 0.f;
    for (int m = 0; m < M; ++m) { const float s = row[m]; ls_sum += logsigmoid_f(s); if (s > mx) { mx = s; am = m; } }
    float se = 0.f;
    for (int m = 0; m < M; ++m) se += expf(row[m] - mx);
    const float lse = (mx == -INFINITY || mx == INFINITY) ? mx : mx + logf(se);
    if (a.lse) a.lse[b] = lse;
    if (a.argmax) a.argmax[b] = am;
    if (a.softmax1) a.softmax1[b] = M > 1 ? expf(row[1] - mx) / se : 1.f;
    if (a.last) a.last[b] = row[M - 1];
    acc[0] = logsigmoid_f(lse); acc[1] = ls_sum;
    if (a.gt) {
      int64_t g = a.gt[b]; g = g < 0 ? 0 : (g >= M ? M - 1 : g);
      const float wg = a.class_w ? a.class_w[g] : 1.f;
      acc[2] = (double)wg * (double)(lse - row[g]); acc[3] = wg;
    }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double v = acc[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (l == 0) red[w][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double s = 0.0;
    for (int i = 0; i < kScoreRows / 32; ++i) s += red[i][threadIdx.x];
    a.partial[(int64_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
}

__global__ void __launch_bounds__(128) score_finish_kernel(const double* __restrict__ partial, int nblocks, float* __restrict__ sums) {
  __shared__ double red[128];
  const int q = threadIdx.x & 3, lane = threadIdx.x >> 2;         // 32 lanes per quantity, fixed strided order
  double s = 0.0;
  for (int i = lane; i < nblocks; i += 32) s += partial[(int64_t)i * 4 + q];
  red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += red[i * 4 + threadIdx.x];
    sums[threadIdx.x] = (float)t;
  }
}

inline int grid_for(int64_t n, int per_thread = 4) {
  int64_t blocks = (n + 256LL * per_thread - 1) / (256LL * per_thread);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace cfpp
using namespace cfpp;

extern "C" int cfpp_coupling_inv(const float* z, const float* h, const float* add, float* x, int B, int C, int HW, void* stream) {
  CFPP_REQUIRE(C >= 2 && C % 2 == 0 && HW >= 1, "coupling_inv: C=%d must be even, HW=%d", C, HW);
  if (B <= 0) return CFPP_OK;
  const int64_t n = (int64_t)(C / 2) * HW;
  const bool vec = (HW % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(z)) % 16 == 0);
  const int64_t work = vec ? n / 4 : n;
  int G = 32;
  while (G < 256 && G * 4 < work) G <<= 1;
  const int spc = 256 / G;
  const int blocks = (int)((B + spc - 1) / spc);
  if (vec) coupling_inv_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(z, h, add, x, B, C, HW, G);
  else coupling_inv_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(z, h, add, x, B, C, HW, G);
  return check_launch("coupling_inv");
}

extern "C" int cfpp_actnorm_inv(const float* z, float* x, const float* t, const float* logs, int B, int D, int HW, void* stream) {
  CFPP_REQUIRE(D >= 1 && HW >= 1 && t && logs, "actnorm_inv: D=%d HW=%d, t/logs required", D, HW);
  if (B <= 0) return CFPP_OK;
  const int64_t total = (int64_t)B * D * HW;
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(z)) % 16 == 0);
  if (vec) actnorm_inv_kernel<true><<<grid_for(total / 4, 2), 256, 0, (cudaStream_t)stream>>>(z, t, logs, x, total, D, HW);
  else actnorm_inv_kernel<false><<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(z, t, logs, x, total, D, HW);
  return check_launch("actnorm_inv");
}

extern "C" int cfpp_mat_inverse(const float* A, int D, float* Ainv, int* singular, void* stream) {
  CFPP_REQUIRE(D >= 1 && D <= 128, "mat_inverse: D=%d outside [1,128]", D);
  static DeviceOnce attr_set;
  const size_t full = ((size_t)128 * 128 + 2 * 128) * sizeof(double);
  if (attr_set.first()) { cudaFuncSetAttribute(mat_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)full); }
  mat_inverse_kernel<<<1, 256, ((size_t)D * D + 2 * D) * sizeof(double), (cudaStream_t)stream>>>(A, D, Ainv, singular);
  return check_launch("mat_inverse");
}

extern "C" int cfpp_prologue_inv(const float* z, float* x, float* x_cont, int B, int C, int A, int HW,
                                 float s1, float t1, float s0, float t0, int do_floor, void* stream) {
  CFPP_REQUIRE(C >= 1 && A >= 0 && HW >= 1, "prologue_inv: C=%d A=%d HW=%d", C, A, HW);
  if (B <= 0) return CFPP_OK;
  const int64_t total = (int64_t)B * C * HW;
  prologue_inv_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(z, x, x_cont, total, C * HW, (C + A) * HW, s1, t1, s0, t0, do_floor);
  return check_launch("prologue_inv");
}

extern "C" int cfpp_sigmoid_fwd(const float* x, float* y, int64_t n, void* stream) {
  if (n <= 0) return CFPP_OK;
  unary_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, n, 0, 0.f, 0.f);
  return check_launch("sigmoid_fwd");
}

extern "C" int cfpp_normalize_inv(const float* y, float* x, int64_t n, float scale, float translation, void* stream) {
  if (n <= 0) return CFPP_OK;
  unary_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(y, x, n, 1, scale, translation);
  return check_launch("normalize_inv");
}

extern "C" int cfpp_floor_fwd(const float* x, float* y, int64_t n, void* stream) {
  if (n <= 0) return CFPP_OK;
  unary_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, n, 2, 0.f, 0.f);
  return check_launch("floor_fwd");
}

extern "C" int cfpp_relu_fwd(const float* x, float* y, int64_t n, void* stream) {
  if (n <= 0) return CFPP_OK;
  unary_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, n, 3, 0.f, 0.f);
  return check_launch("relu_fwd");
}

extern "C" int cfpp_gmm_sample(const float* mG, const float* sG, const int64_t* comp, const float* eps, float* x,
                               int B, int M, int K, int n_per_sample, int m, void* stream) {
  CFPP_REQUIRE(M >= 1 && K >= 1 && n_per_sample >= 1 && m >= 0 && m < M, "gmm_sample: M=%d K=%d n=%d m=%d", M, K, n_per_sample, m);
  if (B <= 0) return CFPP_OK;
  const int64_t total = (int64_t)B * n_per_sample;
  gmm_sample_kernel<<<grid_for(total), 256, 0, (cudaStream_t)stream>>>(mG, sG, comp, eps, x, total, n_per_sample, K, m);
  return check_launch("gmm_sample");
}

extern "C" int64_t cfpp_score_workspace_bytes(int B) {
  const int64_t blocks = ((int64_t)(B > 0 ? B : 1) + kScoreWarps - 1) / kScoreWarps;
  return blocks * 4 * (int64_t)sizeof(double);
}

extern "C" int cfpp_score_epilogue(const float* logp, float dim_inv, const int64_t* gt, const float* class_w,
                                   float* scaled, float* lse, float* softmax1, float* last, int64_t* argmax, float* sums,
                                   void* workspace, int B, int M, void* stream) {
  CFPP_REQUIRE(M >= 1 && logp && sums && workspace, "score_epilogue: M=%d, logp / sums / workspace required", M);
  CFPP_REQUIRE(B >= 1, "score_epilogue: empty batch (the reference's .mean() of no elements is NaN)");
  ScoreArgs a{logp, dim_inv, gt, class_w, scaled, lse, softmax1, last, argmax, (double*)workspace, B, M};
  int blocks;
  if (M <= 32) {                                           // thread per row, tile staged in shared memory
    blocks = (B + kScoreRows - 1) / kScoreRows;
    score_rows_kernel<<<blocks, kScoreRows, (size_t)kScoreRows * M * sizeof(float), (cudaStream_t)stream>>>(a);
  } else {                                                 // warp per row
    blocks = (B + kScoreWarps - 1) / kScoreWarps;
    score_kernel<<<blocks, kScoreWarps * 32, 0, (cudaStream_t)stream>>>(a);
  }
  int rc = check_launch("score_epilogue");
  if (rc != CFPP_OK) return rc;
  score_finish_kernel<<<1, 128, 0, (cudaStream_t)stream>>>((const double*)workspace, blocks, sums);
  return check_launch("score_finish");
}
