// Gaussian-mixture log-prob (base density and SplitPrior), layers/distributions/gaussian.py:142-161.
//
// out[b,m] = logsumexp_k( logmix[m,k] + sum_e [ -(x_be - mu)^2 / (2 sigma^2) - log sigma - 0.5 log 2pi ] )
// mu = mG[m,k,e] (+ cm[b,m,k,d]),  sigma = softplus(sG[m,k,e] (+ cs[b,m,k,d])),  e = (d,h,w).
//
// A CTA keeps a tile of S samples in shared memory and walks all M*K components over it, so each component's
// parameters are fetched from L2 once per S samples and x is read from HBM exactly once.  Without context the
// sigma-dependent terms are hoisted into a per-parameter table by gmm_prepare_kernel (1/(2 sigma^2), sum_e log sigma),
// leaving 3 FP32 ops per Gaussian evaluation; with per-sample context offsets sigma is a function of (b,m,k,e) and
// is evaluated in place (SFU-bound).
#include "common.cuh"

namespace cfpp {

// logmix[m,k] = log_softmax(log(clamp(softmax(wG[m])/sum, eps, 1-eps)))  (torch Categorical(probs) + MixtureSameFamily)
__device__ float log_mix(const float* __restrict__ wrow, int K, int k) {
  float mx = -INFINITY;
  for (int i = 0; i < K; ++i) mx = fmaxf(mx, wrow[i]);
  float den = 0.f;
  for (int i = 0; i < K; ++i) den += expf(wrow[i] - mx);
  float psum = 0.f;
  for (int i = 0; i < K; ++i) psum += expf(wrow[i] - mx) / den;
  const float eps = 1.1920928955078125e-07f;
  float lmx = -INFINITY;
  for (int i = 0; i < K; ++i) lmx = fmaxf(lmx, logf(fminf(fmaxf(expf(wrow[i] - mx) / den / psum, eps), 1.f - eps)));
  float lse = 0.f, lk = 0.f;
  for (int i = 0; i < K; ++i) {
    const float l = logf(fminf(fmaxf(expf(wrow[i] - mx) / den / psum, eps), 1.f - eps));
    lse += expf(l - lmx);
    if (i == k) lk = l;
  }
  return lk - (lmx + logf(lse));
}

// ws layout: [A: MK*E] [LB: MK]   A = 1/(2 sigma^2), LB = logmix - sum_e log sigma - E*0.5*log(2 pi)
__global__ void gmm_prepare_kernel(const float* __restrict__ sG, const float* __restrict__ wG, float* __restrict__ ws,
                                   int MK, int K, int E, bool with_table) {
  __shared__ float red[32];
  const int mk = blockIdx.x;
  float acc = 0.f;
  if (with_table) {
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      const float sc = softplus_f(sG[(int64_t)mk * E + e]);
      ws[(int64_t)mk * E + e] = 1.f / (2.f * sc * sc);
      acc += logf(sc);
    }
    acc = group_sum(acc, blockDim.x, red);
  }
  if (threadIdx.x == 0)
    ws[(int64_t)MK * E + mk] = log_mix(wG + (mk / K) * K, K, mk % K) - acc - (float)E * kHalfLog2Pi;
}

template <bool CTX, int S, int PK>
__global__ void __launch_bounds__(256) gmm_kernel(const float* __restrict__ x, int64_t x_bstride, const float* __restrict__ mG,
                                                  const float* __restrict__ sG, const float* __restrict__ ws,
                                                  const float* __restrict__ ctx_off, const float* __restrict__ logp_c,
                                                  float logp_scale, float* __restrict__ out, int B, int M, int K, int D, int HW) {
  extern __shared__ float4 sm4[];
  float* xs = reinterpret_cast<float*>(sm4);          // [S][E]
  const int E = D * HW, MK = M * K;
  float* comp = xs + (int64_t)S * E;                  // [S][MK]
  const int b0 = blockIdx.x * S;
  for (int idx = threadIdx.x; idx < S * E; idx += blockDim.x) {
    const int s = idx / E, e = idx % E;
    xs[idx] = (b0 + s < B) ? x[(int64_t)(b0 + s) * x_bstride + e] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* A = ws;
  const float* LB = ws + (int64_t)MK * E;
  const int ngroups = (MK + PK - 1) / PK;
  for (int g = warp; g < ngroups; g += nwarps) {
    float acc[PK][S];
#pragma unroll
    for (int j = 0; j < PK; ++j)
#pragma unroll
      for (int s = 0; s < S; ++s) acc[j][s] = 0.f;
    const int mk0 = g * PK;
    for (int e = lane; e < E; e += 32) {
      float xv[S];
#pragma unroll
      for (int s = 0; s < S; ++s) xv[s] = xs[s * E + e];
      const int d = CTX ? e / HW : 0;
#pragma unroll
      for (int j = 0; j < PK; ++j) {
        const int mk = mk0 + j < MK ? mk0 + j : MK - 1;
        const float mu = mG[(int64_t)mk * E + e];
        if (!CTX) {
          const float a = A[(int64_t)mk * E + e];
#pragma unroll
          for (int s = 0; s < S; ++s) { const float df = xv[s] - mu; acc[j][s] = fmaf(-a * df, df, acc[j][s]); }
        } else {
          const float sg = sG[(int64_t)mk * E + e];
#pragma unroll
          for (int s = 0; s < S; ++s) {
            const int bb = b0 + s < B ? b0 + s : B - 1;
            const float* co = ctx_off + (int64_t)bb * 2 * MK * D + (int64_t)mk * D + d;     // 'b (p m k d)'
            const float sc = softplus_f(sg + co[(int64_t)MK * D]);
            const float df = xv[s] - (mu + co[0]);
            acc[j][s] += -(df * df) / (2.f * sc * sc) - logf(sc);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < PK; ++j)
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float v = warp_sum(acc[j][s]);
        if (lane == 0 && mk0 + j < MK) comp[s * MK + mk0 + j] = v + LB[mk0 + j];
      }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < S * M; idx += blockDim.x) {
    const int s = idx / M, m = idx % M;
    if (b0 + s >= B) continue;
    const float* c = comp + s * MK + m * K;
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, c[k]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(c[k] - mx);
    out[(int64_t)(b0 + s) * M + m] = mx + logf(se) + (logp_c ? logp_scale * logp_c[b0 + s] : 0.f);
  }
}

template <bool CTX, int S, int PK>
static int launch_gmm(const float* x, int64_t bs, const float* mG, const float* sG, const float* ws, const float* co, const float* lp,
                      float lps, float* out, int B, int M, int K, int D, int HW, cudaStream_t st) {
  const int E = D * HW, MK = M * K;
  const size_t smem = ((size_t)S * E + (size_t)S * MK) * sizeof(float);
  static DeviceOnce attr_set;
  if (attr_set.first()) { cudaFuncSetAttribute(gmm_kernel<CTX, S, PK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }
  const int ngroups = (MK + PK - 1) / PK;
  const int nwarps = ngroups < 8 ? ngroups : 8;
  gmm_kernel<CTX, S, PK><<<(B + S - 1) / S, nwarps * 32, smem, st>>>(x, bs, mG, sG, ws, co, lp, lps, out, B, M, K, D, HW);
  return check_launch("gmm_logprob");
}

// ---------------------------------------------------------------------------------------------------------------
// Context given as an embedding-table lookup (embed + eyesample priors, model.py:157,162): the per-sample offsets take
// only as many distinct values as there are context tuples.  Samples are bucketed by context tuple so that every CTA
// tile shares ONE (mean-offset row, sigma table row): the sigma-dependent terms are hoisted per distinct scale context
// and the inner loop is the context-free one plus a broadcast add -- no transcendental per Gaussian evaluation.
// ---------------------------------------------------------------------------------------------------------------
struct GmmTabWs {            // carved out of the caller's workspace
  float* A; float* LB; int* perm; int* tile_key; int* n_tiles;
};

// A[v][mk][e] = 1/(2 sigma^2), LB[v][mk] = logmix - sum_e log sigma - E*0.5*log(2 pi); sigma = softplus(sG + scale_off[v][mk][d])
__global__ void gmm_prepare_tab_kernel(const float* __restrict__ sG, const float* __restrict__ wG, const float* __restrict__ stab,
                                       int width, int soff, float* __restrict__ A, float* __restrict__ LB, int MK, int K, int D, int HW) {
  __shared__ float red[32];
  const int mk = blockIdx.x, v = blockIdx.y, E = D * HW;
  const float* so = stab + (int64_t)v * width + soff + mk * D;
  float acc = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float sc = softplus_f(sG[(int64_t)mk * E + e] + so[e / HW]);
    A[((int64_t)v * MK + mk) * E + e] = 1.f / (2.f * sc * sc);
    acc += logf(sc);
  }
  acc = group_sum(acc, blockDim.x, red);
  if (threadIdx.x == 0) LB[v * MK + mk] = log_mix(wG + (mk / K) * K, K, mk % K) - acc - (float)E * kHalfLog2Pi;
}

// Single-CTA counting sort of the batch by context key; buckets are padded to whole tiles of S samples (perm = -1).
__global__ void __launch_bounds__(1024) gmm_bucket_kernel(const int64_t* __restrict__ ctx, int n_ctx, int card1, int B, int NB, int S,
                                                          int* __restrict__ perm, int* __restrict__ tile_key, int* __restrict__ n_tiles) {
  extern __shared__ int sh[];
  int* cnt = sh; int* base = sh + NB; int* cur = sh + 2 * NB;
  __shared__ int total_tiles;
  for (int k = threadIdx.x; k < NB; k += blockDim.x) { cnt[k] = 0; cur[k] = 0; }
  __syncthreads();
  auto key_of = [&](int b) -> int {
    int k = (int)ctx[(int64_t)b * n_ctx];
    if (n_ctx == 2) k = k * card1 + (int)ctx[(int64_t)b * n_ctx + 1];
    return k < 0 ? 0 : (k >= NB ? NB - 1 : k);
  };
  for (int b = threadIdx.x; b < B; b += blockDim.x) atomicAdd(&cnt[key_of(b)], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < NB; ++k) { base[k] = t; t += (cnt[k] + S - 1) / S; }
    total_tiles = t; *n_tiles = t;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < NB; k += blockDim.x) {
    const int nt = (cnt[k] + S - 1) / S;
    for (int t = 0; t < nt; ++t) tile_key[base[k] + t] = k;
  }
  for (int i = threadIdx.x; i < total_tiles * S; i += blockDim.x) perm[i] = -1;
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int k = key_of(b);
    perm[base[k] * S + atomicAdd(&cur[k], 1)] = b;
  }
}

// Up to 20 warps per tile: at small batches (the reference's default B = 256 gives ~10^2 tiles of 8 samples) the tile's latency is the
// kernel's, so every group of PK components gets its own warp instead of taking turns on 8.
template <int S, int PK>
__global__ void __launch_bounds__(640) gmm_tab_kernel(const float* __restrict__ x, int64_t x_bstride, const float* __restrict__ mG,
                                                      const float* __restrict__ A, const float* __restrict__ LB,
                                                      const int* __restrict__ perm, const int* __restrict__ tile_key,
                                                      const int* __restrict__ n_tiles, const float* __restrict__ mtab, int width, int moff,
                                                      int n_ctx, int card1, const float* __restrict__ logp_c, float logp_scale,
                                                      float* __restrict__ out, int M, int K, int D, int HW) {
  if ((int)blockIdx.x >= *n_tiles) return;
  extern __shared__ float4 sm4[];
  float* xs = reinterpret_cast<float*>(sm4);          // [S][E]
  const int E = D * HW, MK = M * K;
  float* comp = xs + (int64_t)S * E;                  // [S][MK]
  __shared__ int bidx[S];
  const int tile = blockIdx.x;
  const int key = tile_key[tile];
  const int vm = n_ctx == 2 ? key / card1 : key, vs = n_ctx == 2 ? key % card1 : key;
  if (threadIdx.x < S) bidx[threadIdx.x] = perm[tile * S + threadIdx.x];
  __syncthreads();
  for (int idx = threadIdx.x; idx < S * E; idx += blockDim.x) {
    const int s = idx / E, e = idx % E;
    xs[idx] = bidx[s] >= 0 ? x[(int64_t)bidx[s] * x_bstride + e] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* Av = A + (int64_t)vs * MK * E;
  const float* cm = mtab + (int64_t)vm * width + moff;
  const int ngroups = (MK + PK - 1) / PK;
  for (int g = warp; g < ngroups; g += nwarps) {
    float acc[PK][S];
#pragma unroll
    for (int j = 0; j < PK; ++j)
#pragma unroll
      for (int s = 0; s < S; ++s) acc[j][s] = 0.f;
    const int mk0 = g * PK;
    for (int e = lane; e < E; e += 32) {
      float xv[S];
#pragma unroll
      for (int s = 0; s < S; ++s) xv[s] = xs[s * E + e];
      const int d = e / HW;
#pragma unroll
      for (int j = 0; j < PK; ++j) {
        const int mk = mk0 + j < MK ? mk0 + j : MK - 1;
        const float mu = mG[(int64_t)mk * E + e] + cm[mk * D + d];
        const float a = Av[(int64_t)mk * E + e];
#pragma unroll
        for (int s = 0; s < S; ++s) { const float df = xv[s] - mu; acc[j][s] = fmaf(-a * df, df, acc[j][s]); }
      }
    }
#pragma unroll
    for (int j = 0; j < PK; ++j)
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const float v = warp_sum(acc[j][s]);
        if (lane == 0 && mk0 + j < MK) comp[s * MK + mk0 + j] = v + LB[vs * MK + mk0 + j];
      }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < S * M; idx += blockDim.x) {
    const int s = idx / M, m = idx % M, b = bidx[s];
    if (b < 0) continue;
    const float* c = comp + s * MK + m * K;
    float mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmaxf(mx, c[k]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(c[k] - mx);
    out[(int64_t)b * M + m] = mx + logf(se) + (logp_c ? logp_scale * logp_c[b] : 0.f);
  }
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

}  // namespace cfpp
using namespace cfpp;

static bool gmm_tab_plan(int B, int M, int K, int D, int HW, int n_ctx, const int* cards, int* NB, int* Vs, int* S, int64_t* bytes) {
  if (n_ctx < 1 || n_ctx > 2) return false;
  int64_t nb = cards[0]; if (n_ctx == 2) nb *= cards[1];
  if (nb < 1 || nb > 2048) return false;
  const int E = D * HW, MK = M * K;
  const size_t budget = 200 * 1024;
  int s = ((size_t)8 * (E + MK)) * 4 <= budget ? 8 : (((size_t)2 * (E + MK)) * 4 <= budget ? 2 : 0);
  if (!s) return false;
  const int vs = n_ctx == 2 ? cards[1] : cards[0];
  const int64_t max_tiles = (B + s - 1) / s + nb;
  *NB = (int)nb; *Vs = vs; *S = s;
  *bytes = align_up((int64_t)vs * MK * E * 4, 256) + align_up((int64_t)vs * MK * 4, 256) + align_up(max_tiles * s * 4, 256) +
           align_up(max_tiles * 4, 256) + 256;
  return true;
}

extern "C" int64_t cfpp_gmm_ctxtab_workspace_bytes(int B, int M, int K, int D, int HW, int n_ctx, const int* cards) {
  int NB, Vs, S; int64_t bytes;
  return gmm_tab_plan(B, M, K, D, HW, n_ctx, cards, &NB, &Vs, &S, &bytes) ? bytes : -1;
}

extern "C" int cfpp_gmm_logprob_ctxtab_cached(const float* x, int64_t x_bstride, const float* mG, const float* sG, const float* wG,
                                              const int64_t* ctx, int n_ctx, const int* cards, const float* const* tables, int width,
                                              const float* logp_c, float logp_scale, float* out, void* workspace, int64_t workspace_bytes,
                                              int tables_ready, int B, int M, int K, int D, int HW, void* stream);

extern "C" int cfpp_gmm_logprob_ctxtab(const float* x, int64_t x_bstride, const float* mG, const float* sG, const float* wG,
                                       const int64_t* ctx, int n_ctx, const int* cards, const float* const* tables, int width,
                                       const float* logp_c, float logp_scale, float* out, void* workspace, int64_t workspace_bytes,
                                       int B, int M, int K, int D, int HW, void* stream) {
  return cfpp_gmm_logprob_ctxtab_cached(x, x_bstride, mG, sG, wG, ctx, n_ctx, cards, tables, width, logp_c, logp_scale, out, workspace, workspace_bytes,
                                        0, B, M, K, D, HW, stream);
}

// tables_ready != 0: the per-scale-context (1 / (2 sigma^2), log-normaliser) tables at the head of `workspace` were filled by an earlier call with
// the SAME parameters, batch size and workspace -- they depend on the parameters only, so a caller that keeps the workspace per parameter
// version skips their recomputation (M*K x contexts CTAs of softplus / log per forward: a third of this op's time at B = 256)
extern "C" int cfpp_gmm_logprob_ctxtab_cached(const float* x, int64_t x_bstride, const float* mG, const float* sG, const float* wG,
                                              const int64_t* ctx, int n_ctx, const int* cards, const float* const* tables, int width,
                                              const float* logp_c, float logp_scale, float* out, void* workspace, int64_t workspace_bytes,
                                              int tables_ready, int B, int M, int K, int D, int HW, void* stream) {
  int NB, Vs, S; int64_t need;
  if (!gmm_tab_plan(B, M, K, D, HW, n_ctx, cards, &NB, &Vs, &S, &need)) {
    set_error("gmm_ctxtab: unsupported context structure (n_ctx=%d) or tile size", n_ctx);
    return CFPP_ERR_UNSUPPORTED;
  }
  const int MK = M * K, E = D * HW;
  CFPP_REQUIRE((int64_t)width * n_ctx == (int64_t)2 * MK * D, "gmm_ctxtab: embedding width %d x %d features != 2*M*K*D", width, n_ctx);
  CFPP_REQUIRE(workspace && workspace_bytes >= need, "gmm_ctxtab: workspace of %lld bytes required", (long long)need);
  if (B <= 0) return CFPP_OK;
  cudaStream_t st = (cudaStream_t)stream;
  char* p = (char*)workspace;
  GmmTabWs w;
  const int64_t max_tiles = (B + S - 1) / S + NB;
  w.A = (float*)p; p += align_up((int64_t)Vs * MK * E * 4, 256);
  w.LB = (float*)p; p += align_up((int64_t)Vs * MK * 4, 256);
  w.perm = (int*)p; p += align_up(max_tiles * S * 4, 256);
  w.tile_key = (int*)p; p += align_up(max_tiles * 4, 256);
  w.n_tiles = (int*)p;
  // 'b (p m k d)': n_ctx == 1 -> one table row holds [mean | scale]; n_ctx == 2 -> feature 0 is the mean half, feature 1 the scale half
  const float* mtab = tables[0];
  const float* stab = n_ctx == 2 ? tables[1] : tables[0];
  const int moff = 0, soff = n_ctx == 2 ? 0 : MK * D;
  int rc = CFPP_OK;
  if (!tables_ready) {
    gmm_prepare_tab_kernel<<<dim3(MK, Vs), 256, 0, st>>>(sG, wG, stab, width, soff, w.A, w.LB, MK, K, D, HW);
    rc = check_launch("gmm_prepare_tab");
    if (rc) return rc;
  }
  gmm_bucket_kernel<<<1, 1024, 3 * NB * sizeof(int), st>>>(ctx, n_ctx, n_ctx == 2 ? cards[1] : 1, B, NB, S, w.perm, w.tile_key, w.n_tiles);
  rc = check_launch("gmm_bucket");
  if (rc) return rc;
  const size_t smem = ((size_t)S * E + (size_t)S * MK) * sizeof(float);
  const int ngroups = (MK + 3) / 4;
  const int nwarps = ngroups < 20 ? ngroups : 20;
  if (S == 8) {
    static DeviceOnce a8;
    if (a8.first()) { cudaFuncSetAttribute(gmm_tab_kernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024); }
    gmm_tab_kernel<8, 4><<<(int)max_tiles, nwarps * 32, smem, st>>>(x, x_bstride, mG, w.A, w.LB, w.perm, w.tile_key, w.n_tiles, mtab, width, moff,
                                                                  n_ctx, n_ctx == 2 ? cards[1] : 1, logp_c, logp_scale, out, M, K, D, HW);
  } else {
    static DeviceOnce a2;
    if (a2.first()) { cudaFuncSetAttribute(gmm_tab_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024); }
    gmm_tab_kernel<2, 4><<<(int)max_tiles, nwarps * 32, smem, st>>>(x, x_bstride, mG, w.A, w.LB, w.perm, w.tile_key, w.n_tiles, mtab, width, moff,
                                                                  n_ctx, n_ctx == 2 ? cards[1] : 1, logp_c, logp_scale, out, M, K, D, HW);
  }
  return check_launch("gmm_logprob_ctxtab");
}

extern "C" int64_t cfpp_gmm_workspace_floats(int M, int K, int D, int HW) { return (int64_t)M * K * D * HW + (int64_t)M * K; }

extern "C" int cfpp_gmm_logprob(const float* x, int64_t x_bstride, const float* mG, const float* sG, const float* wG,
                                const float* ctx_off, const float* logp_c, float logp_scale, float* out, float* workspace,
                                int B, int M, int K, int D, int HW, void* stream) {
  CFPP_REQUIRE(M >= 1 && K >= 1 && K <= 64 && D >= 1 && HW >= 1, "gmm: bad dims M=%d K=%d D=%d HW=%d", M, K, D, HW);
  CFPP_REQUIRE(workspace != nullptr, "gmm: workspace of cfpp_gmm_workspace_floats() floats required");
  if (B <= 0) return CFPP_OK;
  const int E = D * HW, MK = M * K;
  cudaStream_t st = (cudaStream_t)stream;
  gmm_prepare_kernel<<<MK, 256, 0, st>>>(sG, wG, workspace, MK, K, E, ctx_off == nullptr);
  int rc = check_launch("gmm_prepare");
  if (rc) return rc;
  const size_t budget = 200 * 1024;
  const bool fit8 = ((size_t)8 * (E + MK)) * 4 <= budget, fit2 = ((size_t)2 * (E + MK)) * 4 <= budget;
  CFPP_REQUIRE(fit2, "gmm: event size %d too large for the shared-memory tile", E);
  if (ctx_off) {
    return fit8 ? launch_gmm<true, 8, 2>(x, x_bstride, mG, sG, workspace, ctx_off, logp_c, logp_scale, out, B, M, K, D, HW, st)
                : launch_gmm<true, 2, 4>(x, x_bstride, mG, sG, workspace, ctx_off, logp_c, logp_scale, out, B, M, K, D, HW, st);
  }
  return fit8 ? launch_gmm<false, 8, 4>(x, x_bstride, mG, sG, workspace, nullptr, logp_c, logp_scale, out, B, M, K, D, HW, st)
              : launch_gmm<false, 2, 8>(x, x_bstride, mG, sG, workspace, nullptr, logp_c, logp_scale, out, B, M, K, D, HW, st);
}
