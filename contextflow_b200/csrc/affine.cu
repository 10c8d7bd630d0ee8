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
// Conv1x1 (+ optional ActNorm epilogue).  One thread per pixel keeps the D input channels in registers; the DxD
// matrix (shared, or assembled per sample from the raw context matrix c) sits in shared memory and is read as
// warp-broadcast float4.  A CTA covers S samples x PT pixels (S>1 only when HW is small).
// ---------------------------------------------------------------------------------------------------------------
struct Conv1x1Args {
  const float* x; float* z; float* ldj; const float* NN; const float* logabsdet;
  const float* c; const float* logp_c; int contextflow;
  const float* an_t; const float* an_logs; int an_per_sample; const float* an_logp_c; float an_logp_scale;
  int B, D, HW, S, PT, tiles_per_sample;
};

template <int DMAX>
__global__ void __launch_bounds__(256) conv1x1_kernel(const Conv1x1Args a) {
  extern __shared__ float4 smem4[];
  float* Wsm = reinterpret_cast<float*>(smem4);
  const int D = a.D, HW = a.HW;
  constexpr int DP = DMAX;                               // padded row length (multiple of 4)
  const int mstride = D * DP + 4;                        // +4 floats: per-sample matrices land on different banks
  const int group = blockIdx.x / a.tiles_per_sample, tile = blockIdx.x % a.tiles_per_sample;
  const int b0 = group * a.S;
  const int nmat = a.c ? a.S : 1;

  // ---- assemble the matrices ----
  for (int idx = threadIdx.x; idx < nmat * D * DP; idx += blockDim.x) {
    const int m = idx / (D * DP), r = idx % (D * DP), i = r / DP, j = r % DP;
    float v = 0.f;
    const int b = b0 + m;
    if (j < D && b < a.B) {
      const float nn = a.NN[i * D + j];
      if (a.c) {                                         // conv1x1.py:36-49
        const float cij = a.c[((int64_t)b * D + i) * D + j];
        v = (j < i) ? cij : (j == i ? expf(cij) : 0.f);
        if (a.contextflow) v = (v - (i == j ? 1.f : 0.f)) + nn;
      } else v = nn;
    }
    Wsm[m * mstride + i * DP + j] = v;
  }
  __syncthreads();

  const int s = threadIdx.x / a.PT, pl = threadIdx.x % a.PT;
  const int b = b0 + s, p = tile * a.PT + pl;
  if (s >= a.S || b >= a.B) return;
  const bool active = p < HW;

  // ---- per-sample ldj (first pixel of the first tile) ----
  if (tile == 0 && pl == 0) {
    float l = 0.f;
    if (a.c) {
      float cl = 0.f;
      for (int i = 0; i < D; ++i) cl += a.c[((int64_t)b * D + i) * D + i];
      l = (float)HW * ((a.contextflow ? a.logabsdet[0] : 0.f) + cl);
      if (a.logp_c) l += a.logp_c[b] * (float)HW;
    } else l = a.logabsdet[0] * (float)HW;
    if (a.an_logs) {
      const float* lg = a.an_logs + (a.an_per_sample ? (int64_t)b * D : 0);
      float sl = 0.f;
      for (int i = 0; i < D; ++i) sl += lg[i];
      l += sl;
      if (a.an_logp_c) l += a.an_logp_scale * a.an_logp_c[b];
    }
    a.ldj[b] = l;
  }
  if (!active) return;

  float xr[DMAX];
  const float* xb = a.x + (int64_t)b * D * HW + p;
#pragma unroll
  for (int j = 0; j < DMAX; ++j) xr[j] = (j < D) ? xb[(int64_t)j * HW] : 0.f;

  const float* Wm = Wsm + (a.c ? s * mstride : 0);
  float* zb = a.z + (int64_t)b * D * HW + p;
  const float* at = a.an_t ? a.an_t + (a.an_per_sample ? (int64_t)b * D : 0) : nullptr;
  const float* al = a.an_logs ? a.an_logs + (a.an_per_sample ? (int64_t)b * D : 0) : nullptr;
  for (int i = 0; i < D; ++i) {
    const float4* wr = reinterpret_cast<const float4*>(Wm + i * DP);
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int j4 = 0; j4 < DMAX / 4; ++j4) {
      const float4 w = wr[j4];
      acc0 = fmaf(w.x, xr[4 * j4 + 0], acc0); acc1 = fmaf(w.y, xr[4 * j4 + 1], acc1);
      acc0 = fmaf(w.z, xr[4 * j4 + 2], acc0); acc1 = fmaf(w.w, xr[4 * j4 + 3], acc1);
    }
    float v = acc0 + acc1;
    if (al) v = (v - at[i]) * expf(-al[i]);
    zb[(int64_t)i * HW] = v;
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

template <int DMAX>
static int launch_conv1x1(const Conv1x1Args& a, cudaStream_t st) {
  const int nmat = a.c ? a.S : 1;
  const size_t smem = ((size_t)nmat * (a.D * DMAX + 4)) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(conv1x1_kernel<DMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
  const int groups = (a.B + a.S - 1) / a.S;
  conv1x1_kernel<DMAX><<<groups * a.tiles_per_sample, a.S * a.PT, smem, st>>>(a);
  return check_launch("conv1x1_fwd");
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
  Conv1x1Args a{x, z, ldj, NN, logabsdet, c, logp_c, contextflow, an_t, an_logs, an_per_sample, an_logp_c, an_logp_scale, B, D, HW, 1, 1, 1};
  const int DMAX = D <= 8 ? 8 : D <= 16 ? 16 : D <= 32 ? 32 : D <= 64 ? 64 : D <= 80 ? 80 : 128;
  // tile shape: up to 128 pixels per sample per CTA; pack samples when HW is small, bounded by shared memory
  int PT = HW >= 128 ? 128 : ((HW + 31) / 32) * 32;
  if (HW < 32) PT = HW;                                   // tiny images: exact, several samples per warp
  int S = 128 / PT; if (S < 1) S = 1;
  if (c) { const int maxS = (int)((160 * 1024) / ((size_t)(D * DMAX + 4) * sizeof(float))); if (S > maxS) S = maxS < 1 ? 1 : maxS; }
  if (S > B) S = B;
  a.S = S; a.PT = PT; a.tiles_per_sample = (HW + PT - 1) / PT;
  cudaStream_t st = (cudaStream_t)stream;
  switch (DMAX) {
    case 8: return launch_conv1x1<8>(a, st);
    case 16: return launch_conv1x1<16>(a, st);
    case 32: return launch_conv1x1<32>(a, st);
    case 64: return launch_conv1x1<64>(a, st);
    case 80: return launch_conv1x1<80>(a, st);
    default: return launch_conv1x1<128>(a, st);
  }
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
