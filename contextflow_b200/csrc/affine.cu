// Invertible 1x1 convolution (shared or per-sample weights), ActNorm, their fusion, slogdet and the ActNorm
// data-dependent initialisation statistics.   Reference: layers/conv1x1.py:28-57, layers/actnorm.py:28-60.
#include "common.cuh"

namespace cfpp {

// ---------------------------------------------------------------------------------------------------------------
// slogdet: LU with partial pivoting in fp64, one CTA.  A is DxD fp32 row-major.
// ---------------------------------------------------------------------------------------------------------------
__global__ void slogdet_kernel(const float* __restrict__ A, int D, float* __restrict__ out) {
  extern __shared__ double lu[];            // D*D
  __shared__ int piv_row;
  __shared__ double piv_val;
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) lu[i] = (double)A[i];
  __syncthreads();
  double logabs = 0.0;                      // only thread 0's copy is used
  for (int k = 0; k < D; ++k) {
    if (threadIdx.x < 32) {                 // warp 0: arg-max |lu[i][k]| over i >= k
      double best = -1.0; int bi = k;
      for (int i = k + threadIdx.x; i < D; i += 32) { double v = fabs(lu[i * D + k]); if (v > best) { best = v; bi = i; } }
      for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (threadIdx.x == 0) { piv_row = bi; piv_val = lu[bi * D + k]; }
    }
    __syncthreads();
    const int pr = piv_row; const double pv = piv_val;
    if (pr != k) for (int j = threadIdx.x; j < D; j += blockDim.x) { double t = lu[k * D + j]; lu[k * D + j] = lu[pr * D + j]; lu[pr * D + j] = t; }
    if (threadIdx.x == 0) logabs += log(fabs(pv));
    __syncthreads();
    if (pv != 0.0) {
      for (int i = k + 1 + threadIdx.x; i < D; i += blockDim.x) {
        const double f = lu[i * D + k] / pv;
        for (int j = k + 1; j < D; ++j) lu[i * D + j] -= f * lu[k * D + j];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)logabs;
}

// ---------------------------------------------------------------------------------------------------------------
// Conv1x1 (+ optional ActNorm epilogue): per sample a (D x D) by (D x HW) product.  A CTA owns NSUB sub-tiles, each
// (one sample, PT pixels); the matrix (shared, or assembled per sample from the raw context matrix c) is kept TRANSPOSED
// in shared memory so that a thread's 4 output rows are one 128-bit load, the pixel tile likewise; each thread owns a
// 4 (rows) x 4 (pixels) register tile: 16 FMA per two LDS.128.
// ---------------------------------------------------------------------------------------------------------------
struct Conv1x1Args {
  const float* x; float* z; float* ldj; const float* NN; const float* logabsdet;
  const float* c; const float* logp_c; int contextflow;
  const float* an_t; const float* an_logs; int an_per_sample; const float* an_logp_c; float an_logp_scale;
  int B, D, HW, PT, NI, TPS, NSUB, tiles_per_sample, WS;
};

template <bool VEC>
__global__ void __launch_bounds__(256) conv1x1_kernel(const Conv1x1Args a) {
  extern __shared__ float4 smem4[];
  const int D = a.D, HW = a.HW, PT = a.PT, WS = a.WS;
  const int nmat = a.c ? a.NSUB : 1;
  float* Wt = reinterpret_cast<float*>(smem4);                 // [nmat][D][WS]   Wt[j][i] = W[i][j]
  float* xs = Wt + (int64_t)nmat * D * WS;                     // [NSUB][D][PT]
  const int64_t total_sub = (int64_t)a.B * a.tiles_per_sample;
  const int64_t st0 = (int64_t)blockIdx.x * a.NSUB;

  // ---- assemble the matrices (transposed) and the pixel tiles: one warp per (matrix row | channel row), lanes along it ----
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int mi = warp; mi < nmat * D; mi += nwarps) {
    const int m = mi / D, i = mi - m * D;
    const int64_t st = st0 + m;
    float* Wm = Wt + (int64_t)m * D * WS + i;
    if (!a.c) {
      for (int j = lane; j < D; j += 32) Wm[j * WS] = a.NN[i * D + j];
    } else if (st < total_sub) {                               // conv1x1.py:36-49
      const float* crow = a.c + ((st / a.tiles_per_sample) * D + i) * D;
      for (int j = lane; j < D; j += 32) {
        const float cij = crow[j];
        float v = (j < i) ? cij : (j == i ? expf(cij) : 0.f);
        if (a.contextflow) v = (v - (i == j ? 1.f : 0.f)) + a.NN[i * D + j];
        Wm[j * WS] = v;
      }
    }
  }
  for (int mj = warp; mj < a.NSUB * D; mj += nwarps) {
    const int m = mj / D, j = mj - m * D;
    const int64_t st = st0 + m;
    float* xr = xs + (int64_t)mj * PT;
    if (st < total_sub) {
      const int64_t b = st / a.tiles_per_sample; const int p0t = (int)(st % a.tiles_per_sample) * PT;
      const float* xg = a.x + (b * D + j) * HW + p0t;
      for (int pl = lane; pl < PT; pl += 32) xr[pl] = (p0t + pl < HW) ? xg[pl] : 0.f;
    } else {
      for (int pl = lane; pl < PT; pl += 32) xr[pl] = 0.f;
    }
  }
  __syncthreads();

  const int m = threadIdx.x / a.TPS, t = threadIdx.x % a.TPS;
  const int64_t st = st0 + m;
  if (m >= a.NSUB || st >= total_sub) return;
  const int64_t b = st / a.tiles_per_sample;
  const int ptile = (int)(st % a.tiles_per_sample);

  // ---- per-sample ldj (one thread per sample) ----
  if (ptile == 0 && t == 0) {
    float l = 0.f;
    if (a.c) {
      float cl = 0.f;
      for (int i = 0; i < D; ++i) cl += a.c[(b * D + i) * D + i];
      l = (float)HW * ((a.contextflow ? a.logabsdet[0] : 0.f) + cl);
      if (a.logp_c) l += a.logp_c[b] * (float)HW;
    } else l = a.logabsdet[0] * (float)HW;
    if (a.an_logs) {
      const float* lg = a.an_logs + (a.an_per_sample ? b * D : 0);
      float sl = 0.f;
      for (int i = 0; i < D; ++i) sl += lg[i];
      l += sl;
      if (a.an_logp_c) l += a.an_logp_scale * a.an_logp_c[b];
    }
    a.ldj[b] = l;
  }

  const int ig = t % a.NI, pg = t / a.NI;
  const int i0 = 4 * ig, p0 = 4 * pg;
  const float* Wm = Wt + (a.c ? (int64_t)m * D * WS : 0) + i0;
  const float* xm = xs + (int64_t)m * D * PT + p0;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
#pragma unroll 4
  for (int j = 0; j < D; ++j) {
    const float4 w = *reinterpret_cast<const float4*>(Wm + j * WS);
    const float4 xv = *reinterpret_cast<const float4*>(xm + j * PT);
    const float wr[4] = {w.x, w.y, w.z, w.w}, xq[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[r][q] = fmaf(wr[r], xq[q], acc[r][q]);
  }
  const float* at = a.an_t ? a.an_t + (a.an_per_sample ? b * D : 0) : nullptr;
  const float* al = a.an_logs ? a.an_logs + (a.an_per_sample ? b * D : 0) : nullptr;
  const int p = ptile * PT + p0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + r;
    if (i >= D) break;
    float v[4] = {acc[r][0], acc[r][1], acc[r][2], acc[r][3]};
    if (al) { const float tt = at[i], e = expf(-al[i]);
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = (v[q] - tt) * e; }
    float* zp = a.z + (b * D + i) * HW + p;
    if (VEC) { if (p < HW) *reinterpret_cast<float4*>(zp) = make_float4(v[0], v[1], v[2], v[3]); }
    else {
#pragma unroll
      for (int q = 0; q < 4; ++q) if (p + q < HW) zp[q] = v[q];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ActNorm alone: elementwise with per-(sample, channel) shift / log-scale.
// ---------------------------------------------------------------------------------------------------------------
__global__ void actnorm_kernel(const float* __restrict__ x, float* __restrict__ z, float* __restrict__ ldj,
                               const float* __restrict__ bt, const float* __restrict__ bl, const float* __restrict__ c,
                               const float* __restrict__ logp_c, float logp_scale, int mode, int B, int D, int HW, bool vec4) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (tid < B) {                                                    // ldj[b] = sum_d logs (actnorm.py:58)
    float s = 0.f;
    for (int d = 0; d < D; ++d) {
      float l = (mode != 2) ? bl[d] : 0.f;
      if (mode != 0) l = c[tid * 2 * D + D + d] + l;
      s += l;
    }
    ldj[tid] = s + (logp_c ? logp_scale * logp_c[tid] : 0.f);
  }
  auto coef = [&](int64_t bd, float& t, float& e) {
    const int d = bd % D; const int64_t b = bd / D;
    float tt = (mode != 2) ? bt[d] : 0.f, ll = (mode != 2) ? bl[d] : 0.f;
    if (mode != 0) { tt = c[b * 2 * D + d] + tt; ll = c[b * 2 * D + D + d] + ll; }
    t = tt; e = expf(-ll);
  };
  if (vec4) {
    const int64_t n4 = (int64_t)B * D * HW / 4;
    const int hw4 = HW / 4;
    for (int64_t i = tid; i < n4; i += stride) {
      float t, e; coef(i / hw4, t, e);
      float4 v = ldg_stream(reinterpret_cast<const float4*>(x) + i);
      v.x = (v.x - t) * e; v.y = (v.y - t) * e; v.z = (v.z - t) * e; v.w = (v.w - t) * e;
      stg_stream(reinterpret_cast<float4*>(z) + i, v);
    }
  } else {
    const int64_t n = (int64_t)B * D * HW;
    for (int64_t i = tid; i < n; i += stride) {
      float t, e; coef(i / HW, t, e);
      z[i] = (x[i] - t) * e;
    }
  }
}

// ActNorm.initialize statistics: one CTA per channel, fp64 two-pass (mean, then unbiased variance).
__global__ void actnorm_stats_kernel(const float* __restrict__ x, float* __restrict__ mean, float* __restrict__ logstd,
                                     int B, int D, int HW) {
  __shared__ double red[32];
  __shared__ double mu_s;
  const int d = blockIdx.x;
  const int64_t n = (int64_t)B * HW;
  auto block_sum = [&](double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    __syncthreads();
    return s;
  };
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += (double)x[((i / HW) * D + d) * (int64_t)HW + (i % HW)];
  const double tot = block_sum(acc);
  if (threadIdx.x == 0) mu_s = tot / (double)n;
  __syncthreads();
  const double mu = mu_s;
  acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = (double)x[((i / HW) * D + d) * (int64_t)HW + (i % HW)] - mu;
    acc += v * v;
  }
  const double ss = block_sum(acc);
  if (threadIdx.x == 0) {
    mean[d] = (float)mu;
    const float sd = (float)sqrt(ss / (double)(n - 1));           // torch.std: unbiased
    logstd[d] = logf(sd + 1e-8f);                                 // actnorm.py:32
  }
}

}  // namespace cfpp
using namespace cfpp;

extern "C" int cfpp_slogdet(const float* A, int D, float* logabsdet, void* stream) {
  CFPP_REQUIRE(D >= 1 && D <= 128, "slogdet: D=%d outside [1,128]", D);
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(slogdet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 8); attr_set = true; }
  slogdet_kernel<<<1, 128, (size_t)D * D * sizeof(double), (cudaStream_t)stream>>>(A, D, logabsdet);
  return check_launch("slogdet");
}

extern "C" int cfpp_conv1x1_fwd(const float* x, float* z, float* ldj, const float* NN, const float* logabsdet,
                                const float* c, const float* logp_c, int contextflow,
                                const float* an_t, const float* an_logs, int an_per_sample, const float* an_logp_c, float an_logp_scale,
                                int B, int D, int HW, void* stream) {
  CFPP_REQUIRE(D >= 1 && D <= 128 && HW >= 1, "conv1x1: D=%d HW=%d unsupported (D<=128)", D, HW);
  CFPP_REQUIRE((an_t == nullptr) == (an_logs == nullptr), "conv1x1: an_t and an_logs must be given together");
  if (B <= 0) return CFPP_OK;
  Conv1x1Args a{x, z, ldj, NN, logabsdet, c, logp_c, contextflow, an_t, an_logs, an_per_sample, an_logp_c, an_logp_scale, B, D, HW};
  const int DP = (D + 3) / 4 * 4;
  a.NI = DP / 4;
  a.WS = DP + 4;                                            // row stride of the transposed matrix (16B aligned, bank-rotated)
  int pt_cap = 4 * (128 / a.NI > 1 ? 128 / a.NI : 1);
  int hw4 = ((HW < 128 ? HW : 128) + 3) / 4 * 4;
  a.PT = hw4 < pt_cap ? hw4 : pt_cap;
  a.TPS = a.NI * (a.PT / 4);
  a.tiles_per_sample = (HW + a.PT - 1) / a.PT;
  const size_t per_sub = ((size_t)D * a.PT + (c ? (size_t)D * a.WS : 0)) * sizeof(float);
  const size_t fixed = c ? 0 : (size_t)D * a.WS * sizeof(float);
  int nsub = 256 / a.TPS; if (nsub < 1) nsub = 1;
  const int by_smem = (int)((100 * 1024 - fixed) / per_sub);
  if (nsub > by_smem) nsub = by_smem < 1 ? 1 : by_smem;
  const int64_t total_sub = (int64_t)B * a.tiles_per_sample;
  if (nsub > total_sub) nsub = (int)total_sub;
  a.NSUB = nsub;
  const size_t smem = fixed + (size_t)nsub * per_sub;
  CFPP_REQUIRE(a.TPS <= 256 && smem <= 200 * 1024, "conv1x1: tile does not fit (D=%d)", D);
  const int threads = (nsub * a.TPS + 31) / 32 * 32;
  const int64_t blocks = (total_sub + nsub - 1) / nsub;
  const bool vec = (HW % 4 == 0) && (reinterpret_cast<uintptr_t>(z) % 16 == 0);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(conv1x1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(conv1x1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  if (vec) conv1x1_kernel<true><<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(a);
  else conv1x1_kernel<false><<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(a);
  return check_launch("conv1x1_fwd");
}

extern "C" int cfpp_actnorm_fwd(const float* x, float* z, float* ldj, const float* base_t, const float* base_logs, const float* c,
                                const float* logp_c, float logp_scale, int mode, int B, int D, int HW, void* stream) {
  CFPP_REQUIRE(mode >= 0 && mode <= 2, "actnorm: mode %d", mode);
  CFPP_REQUIRE(mode == 0 || c != nullptr, "actnorm: context matrix required for mode %d", mode);
  CFPP_REQUIRE(mode == 2 || (base_t && base_logs), "actnorm: base parameters required");
  if (B <= 0) return CFPP_OK;
  const bool vec4 = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(z)) % 16 == 0);
  const int64_t n = (int64_t)B * D * HW / (vec4 ? 4 : 1);
  int64_t blocks = (n + 255) / 256;
  const int64_t minb = (B + 255) / 256, cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < minb) blocks = minb;
  actnorm_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, z, ldj, base_t, base_logs, c, logp_c, logp_scale, mode, B, D, HW, vec4);
  return check_launch("actnorm_fwd");
}

extern "C" int cfpp_actnorm_stats(const float* x, float* mean, float* logstd, int B, int D, int HW, void* stream) {
  CFPP_REQUIRE(B >= 1 && D >= 1 && HW >= 1, "actnorm_stats: bad dims");
  actnorm_stats_kernel<<<D, 256, 0, (cudaStream_t)stream>>>(x, mean, logstd, B, D, HW);
  return check_launch("actnorm_stats");
}
