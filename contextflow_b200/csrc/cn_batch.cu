// The CN context networks of a whole forward in one launch (coupling.py:37,45; actnorm.py:21,44; conv1x1.py:22,32).
// Every specialist layer maps its encoded context c (B, C_ctx <= 64) through a tiny MLP or a single nn.Linear; issued one
// nn.Linear at a time that is 60 launches per forward (cfg2), each a few microseconds of work on a fraction of the SMs.  Here
// blockIdx.y selects the job (one chain of 1..3 linear layers) and blockIdx.x a tile of 16 samples whose activations stay in
// shared memory between the layers.  A thread owns one output column for a group of samples: the weight column streams
// through registers once per group (coalesced along the output dimension, L1/L2 resident), the activations arrive as
// warp-broadcast 128-bit shared loads (4 FMA per load).
#include "common.cuh"
#include <algorithm>

namespace cfpp {

constexpr int CN_SPB = 16;          // samples per CTA
constexpr int CN_THREADS = 256;

struct CnBatchArgs {
  cfpp_cn_job job[CFPP_MAX_CN_JOBS];
  const float* in[CFPP_MAX_CN_JOBS];
  float* out[CFPP_MAX_CN_JOBS];
  int stride;                       // shared-memory row stride in floats (multiple of 4, >= every staged width)
};

// Activations of the CTA's 16 samples live in shared memory as float4 [K/4][CN_SPB]: element (sample s, channel k) is component k % 4 of
// entry (k / 4) * CN_SPB + s.  A thread's SPG samples are then SPG consecutive float4 at a compile-time offset from one base register,
// and the inner loop carries two pointers (weights, activations) that advance by a constant -- the first version recomputed
// `row * stride + k` and `k * N + n` per load and spent 45 % of its instructions on integer address arithmetic (ncu, r1u).
template <int SPG>
__device__ __forceinline__ void cn_layer(const float4* __restrict__ src, float* __restrict__ dst, float* __restrict__ gout,
                                         const float* __restrict__ wt, const float* __restrict__ bias, int K, int N, bool last,
                                         int tril, int nS, int s0, int col, int cols_per_pass) {
  const bool tri = last && tril > 0;                       // enumerate only the lower triangle + diagonal: D (D + 1) / 2 columns
  const int ncols = tri ? tril * (tril + 1) / 2 : N;
  const int K4 = (K + 3) >> 2;
  const int64_t N4 = 4 * (int64_t)N;
  for (int t = col; t < ncols; t += cols_per_pass) {
    int n = t;
    if (tri) {
      int i = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
      while (i * (i + 1) / 2 > t) --i;
      while ((i + 1) * (i + 2) / 2 <= t) ++i;
      n = i * tril + (t - i * (i + 1) / 2);
    }
    float acc[SPG];
#pragma unroll
    for (int s = 0; s < SPG; ++s) acc[s] = 0.f;
    const float* wp = wt + n;
    const float4* ap = src + s0;
    int krem = K;
    // the weight column streams from L1 / L2: the four values of step k4 + 1 are requested before the FMAs of step k4 (ncu r2e: 5.7 warps
    // per issue were waiting on these loads when they were consumed right after being issued)
    auto load_w = [&](const float* p, int rem, float (&w)[4]) {
      if (rem >= 4) { w[0] = __ldg(p); w[1] = __ldg(p + N); w[2] = __ldg(p + 2 * N); w[3] = __ldg(p + 3 * N); }
      else {
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = q < rem ? __ldg(p + (int64_t)q * N) : 0.f;
      }
    };
    float wn[4];
    load_w(wp, krem, wn);
    for (int k4 = 0; k4 < K4; ++k4, wp += N4, ap += CN_SPB, krem -= 4) {
      float w[4] = {wn[0], wn[1], wn[2], wn[3]};
      if (k4 + 1 < K4) load_w(wp + N4, krem - 4, wn);
#pragma unroll
      for (int s = 0; s < SPG; ++s) {
        const float4 c = ap[s];
        acc[s] = fmaf(c.x, w[0], acc[s]); acc[s] = fmaf(c.y, w[1], acc[s]);
        acc[s] = fmaf(c.z, w[2], acc[s]); acc[s] = fmaf(c.w, w[3], acc[s]);
      }
    }
    const float bv = bias ? __ldg(bias + n) : 0.f;
    if (last) {
      float* gp = gout + (int64_t)s0 * N + n;
#pragma unroll
      for (int s = 0; s < SPG; ++s) if (s0 + s < nS) gp[(int64_t)s * N] = acc[s] + bv;
    } else {
      float* dp = dst + ((n >> 2) * CN_SPB + s0) * 4 + (n & 3);
#pragma unroll
      for (int s = 0; s < SPG; ++s) dp[4 * s] = fmaxf(acc[s] + bv, 0.f);
    }
  }
}

__global__ void __launch_bounds__(CN_THREADS, 4) cn_batch_kernel(const __grid_constant__ CnBatchArgs a, int B) {
  extern __shared__ float4 cn_smem4[];
  const int rows4 = a.stride >> 2;                           // float4 rows per buffer
  float4* buf0 = cn_smem4;                                   // [rows4][CN_SPB]
  float4* buf1 = buf0 + rows4 * CN_SPB;
  const cfpp_cn_job& J = a.job[blockIdx.y];
  const int64_t b0 = (int64_t)blockIdx.x * CN_SPB;
  const int nS = (int)min((int64_t)CN_SPB, (int64_t)B - b0);
  const int tid = threadIdx.x;
  {
    const int K = J.K, K4 = (K + 3) >> 2;
    const float* in = a.in[blockIdx.y] + b0 * K;
    float* b0f = reinterpret_cast<float*>(buf0);
    for (int idx = tid; idx < K4 * 4 * CN_SPB; idx += CN_THREADS) {          // idx = (k, s), s fastest within a k: coalesced enough for a (16, K) tile
      const int k = idx / CN_SPB, s = idx - k * CN_SPB;
      b0f[((k >> 2) * CN_SPB + s) * 4 + (k & 3)] = (s < nS && k < K) ? __ldg(in + s * K + k) : 0.f;
    }
  }
  __syncthreads();
  float4* src = buf0; float4* dst = buf1;
  float* gout = a.out[blockIdx.y] + b0 * J.N[J.n_layers - 1];
  for (int l = 0; l < J.n_layers; ++l) {
    const int K = l ? J.N[l - 1] : J.K, N = J.N[l];
    const bool last = l == J.n_layers - 1;
    const int ncols = (last && J.tril_dim > 0) ? J.tril_dim * (J.tril_dim + 1) / 2 : N;
    const int cpp = min(CN_THREADS, (ncols + 31) & ~31);
    int groups = CN_THREADS / cpp;                       // 1, 2, 4 or 8 sample groups
    groups = groups >= 8 ? 8 : groups >= 4 ? 4 : groups >= 2 ? 2 : 1;
    const int g = tid / cpp, col = tid - g * cpp;
    float* dstf = reinterpret_cast<float*>(dst);
    if (!last) {                                         // zero the channel padding the next layer's float4 reads touch
      const int N4 = (N + 3) & ~3;
      for (int idx = tid; idx < (N4 - N) * CN_SPB; idx += CN_THREADS) { const int n = N + idx / CN_SPB, sidx = idx % CN_SPB; dstf[((n >> 2) * CN_SPB + sidx) * 4 + (n & 3)] = 0.f; }
    }
    if (g < groups) {
      const int spg = CN_SPB / groups;
      switch (spg) {
        case 16: cn_layer<16>(src, dstf, gout, J.w[l], J.b[l], K, N, last, J.tril_dim, nS, g * 16, col, cpp); break;
        case 8:  cn_layer<8>(src, dstf, gout, J.w[l], J.b[l], K, N, last, J.tril_dim, nS, g * 8, col, cpp); break;
        case 4:  cn_layer<4>(src, dstf, gout, J.w[l], J.b[l], K, N, last, J.tril_dim, nS, g * 4, col, cpp); break;
        default: cn_layer<2>(src, dstf, gout, J.w[l], J.b[l], K, N, last, J.tril_dim, nS, g * 2, col, cpp); break;
      }
    }
    if (!last) { __syncthreads(); float4* t = src; src = dst; dst = t; }
  }
}

}  // namespace cfpp
using namespace cfpp;

extern "C" int cfpp_cn_batch(const cfpp_cn_job* jobs, int n_jobs, const float* const* in, float* const* out, int B, void* stream) {
  CFPP_REQUIRE(n_jobs >= 1 && n_jobs <= CFPP_MAX_CN_JOBS, "cn_batch: n_jobs=%d outside [1,%d]", n_jobs, CFPP_MAX_CN_JOBS);
  CFPP_REQUIRE(jobs && in && out, "cn_batch: null argument");
  int order[CFPP_MAX_CN_JOBS]; double cost[CFPP_MAX_CN_JOBS];
  int width = 4;
  for (int j = 0; j < n_jobs; ++j) {
    const cfpp_cn_job& J = jobs[j];
    CFPP_REQUIRE(J.n_layers >= 1 && J.n_layers <= 3, "cn_batch: job %d has %d layers (1..3)", j, J.n_layers);
    CFPP_REQUIRE(J.K >= 1 && J.K <= CFPP_CN_MAX_WIDTH, "cn_batch: job %d K=%d outside [1,%d]", j, J.K, CFPP_CN_MAX_WIDTH);
    width = std::max(width, J.K);
    double c = 0; int K = J.K;
    for (int l = 0; l < J.n_layers; ++l) {
      CFPP_REQUIRE(J.N[l] >= 1 && J.w[l], "cn_batch: job %d layer %d: N=%d / null weight", j, l, J.N[l]);
      if (l + 1 < J.n_layers) {
        CFPP_REQUIRE(J.N[l] <= CFPP_CN_MAX_WIDTH, "cn_batch: job %d hidden width %d > %d", j, J.N[l], CFPP_CN_MAX_WIDTH);
        width = std::max(width, J.N[l]);
      }
      c += (double)K * J.N[l]; K = J.N[l];
    }
    const int NL = J.N[J.n_layers - 1];
    CFPP_REQUIRE(J.tril_dim == 0 || (J.tril_dim > 0 && (int64_t)J.tril_dim * J.tril_dim == NL), "cn_batch: job %d tril_dim=%d does not match N=%d", j, J.tril_dim, NL);
    CFPP_REQUIRE(in[j] && out[j], "cn_batch: job %d null in/out", j);
    order[j] = j; cost[j] = c;
  }
  if (B <= 0) return CFPP_OK;
  std::stable_sort(order, order + n_jobs, [&](int x, int y) { return cost[x] > cost[y]; });   // heavy chains first: no tail
  CnBatchArgs a;
  for (int j = 0; j < n_jobs; ++j) { a.job[j] = jobs[order[j]]; a.in[j] = in[order[j]]; a.out[j] = out[order[j]]; }
  a.stride = (width + 3) & ~3;
  const size_t smem = 2ull * CN_SPB * a.stride * sizeof(float);
  static DeviceHighWater smem_set;                              // per device: one process may drive several GPUs
  if (smem > 48 * 1024 && smem_set.raise((long long)smem)) {
    cudaError_t e = cudaFuncSetAttribute(cn_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cn_batch: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return CFPP_ERR_CUDA; }
  }
  const int64_t tiles = ((int64_t)B + CN_SPB - 1) / CN_SPB;
  CFPP_REQUIRE(tiles <= 0x7fffffff, "cn_batch: batch too large");
  dim3 grid((unsigned)tiles, n_jobs);
  cn_batch_kernel<<<grid, CN_THREADS, smem, (cudaStream_t)stream>>>(a, B);
  return check_launch("cn_batch");
}
