// Index-only layers (Squeeze, PermuteAxes, channel slice) and the image prologue (Dequantization, Normalization x2,
// LogitTransform, Augment).  All HBM-bound; values are moved bit-exactly by the index kernels.
#include "common.cuh"

namespace cfpp {

// ---- Squeeze: y[b, (c*p1+i)*p2+j, h, w] = x[b, c, h*p1+i, w*p2+j] -------------------------------------------
// One thread per INPUT float2/float pair row segment would complicate generic p; the input row (W floats) is read
// coalesced by consecutive threads and scattered to p2 output planes; the p2 planes' writes are each W/p2-contiguous.
// Grid-stride over input elements keeps reads perfectly coalesced; writes hit p2 distinct 32B-sector streams per warp.
__global__ void squeeze_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int C, int H, int W,
                               int p1, int p2, bool inverse) {
  const int Ho = H / p1, Wo = W / p2;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    // idx enumerates the OUTPUT-of-squeeze layout so that the wide tensor side is always written/read contiguously
    int w = idx % Wo; int64_t r = idx / Wo;
    int h = r % Ho; r /= Ho;
    int j = r % p2; r /= p2;
    int i = r % p1; r /= p1;
    int c = r % C; int64_t b = r / C;
    int64_t in = ((b * C + c) * H + (h * p1 + i)) * (int64_t)W + (w * p2 + j);
    if (!inverse) y[idx] = x[in]; else y[in] = x[idx];
  }
}

// Squeeze with a (2,2) patch and W % 8 == 0 (the image stacks): a thread moves 8 consecutive input floats of one input row
// (two 128-bit streaming loads) to the two output planes j = 0 / 1 (one 128-bit streaming store each); no per-element index math.
// x_gap = floats between the end of one sample's C channels and the start of the next sample (0 for a contiguous tensor; C*H*W when x is
// the first half of the channels of a SplitPrior input, so that the split needs no copy of its own)
__global__ void __launch_bounds__(256) squeeze22_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t units, int H, int W,
                                                        int C, int64_t x_gap) {
  const int W8 = W >> 3, Ho = H >> 1, Wo = W >> 1;
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = u / W8; const int wq = (int)(u - r * W8);
    const int64_t bc = r / H; const int hi = (int)(r - bc * H);
    const int h = hi >> 1, i = hi & 1;
    const float4* src = reinterpret_cast<const float4*>(x + r * W + wq * 8 + (x_gap ? (bc / C) * x_gap : 0));
    const float4 a = ldg_stream(src), b = ldg_stream(src + 1);
    float* dst = y + ((bc * 4 + i * 2) * Ho + h) * (int64_t)Wo + wq * 4;
    stg_stream(reinterpret_cast<float4*>(dst), make_float4(a.x, a.z, b.x, b.z));
    stg_stream(reinterpret_cast<float4*>(dst + (int64_t)Ho * Wo), make_float4(a.y, a.w, b.y, b.w));
  }
}

// The inverse: two 128-bit streaming loads from the planes j = 0 / 1, one interleaved 32-byte row segment out.
__global__ void __launch_bounds__(256) unsqueeze22_kernel(const float* __restrict__ y, float* __restrict__ x, int64_t units, int H, int W) {
  const int W8 = W >> 3, Ho = H >> 1, Wo = W >> 1;
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = u / W8; const int wq = (int)(u - r * W8);
    const int64_t bc = r / H; const int hi = (int)(r - bc * H);
    const int h = hi >> 1, i = hi & 1;
    const float* src = y + ((bc * 4 + i * 2) * Ho + h) * (int64_t)Wo + wq * 4;
    const float4 a = ldg_stream(reinterpret_cast<const float4*>(src));
    const float4 b = ldg_stream(reinterpret_cast<const float4*>(src + (int64_t)Ho * Wo));
    float4* dst = reinterpret_cast<float4*>(x + r * W + wq * 8);
    stg_stream(dst, make_float4(a.x, b.x, a.y, b.y));
    stg_stream(dst + 1, make_float4(a.z, b.z, a.w, b.w));
  }
}

// ---- PermuteAxes (0,2,1,3): y[b,h,c,w] = x[b,c,h,w];  32x32 smem tile transpose over (c,h) per (b,w) ---------
__global__ void permute_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int H, int W) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z / W, w = blockIdx.z % W;
  const int c0 = blockIdx.y * 32, h0 = blockIdx.x * 32;
  const float* xb = x + (int64_t)b * C * H * W;
  float* yb = y + (int64_t)b * C * H * W;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int c = c0 + r, h = h0 + threadIdx.x;
    if (c < C && h < H) tile[r][threadIdx.x] = xb[((int64_t)c * H + h) * W + w];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int h = h0 + r, c = c0 + threadIdx.x;
    if (c < C && h < H) yb[((int64_t)h * C + c) * W + w] = tile[threadIdx.x][r];
  }
}

__global__ void slice_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t total, int C, int HW, int c0, int Cn) {
  const int64_t per = (int64_t)Cn * HW;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = idx / per, r = idx % per;
    y[idx] = x[(b * C + c0) * HW + r];
  }
}

__global__ void place_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t total, int C, int HW, int c0, int Cn) {
  const int64_t per = (int64_t)Cn * HW;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = idx / per, r = idx % per;
    dst[(b * C + c0) * HW + r] = src[idx];
  }
}

// Sliding windows with replication padding (datasets/mtad_data_preprocess.py:58-74) straight into the model layout of
// datasets/mtad_dataloader.py:106-110: x[b, d, t, 0] = float(ts[max(0, end[b] - L + 1 + t), d]).  One CTA per window: the L rows are
// read coalesced along d, transposed through shared memory, written coalesced along t.
template <typename T>
__global__ void __launch_bounds__(256) windows_kernel(const T* __restrict__ ts, const int64_t* __restrict__ end, int64_t end0, int64_t stride,
                                                      float* __restrict__ x, int64_t n_rows, int D, int L) {
  extern __shared__ float tile[];                     // L * (D + 1)
  const int64_t b = blockIdx.x;
  const int64_t e = end ? end[b] : end0 + b * stride;
  for (int i = threadIdx.x; i < L * D; i += blockDim.x) {
    const int t = i / D, d = i - t * D;
    int64_t r = e - L + 1 + t; r = r < 0 ? 0 : (r >= n_rows ? n_rows - 1 : r);
    tile[t * (D + 1) + d] = (float)ts[r * D + d];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L * D; i += blockDim.x) {
    const int d = i / L, t = i - d * L;
    x[b * (int64_t)L * D + i] = tile[t * (D + 1) + d];
  }
}

__global__ void add_kernel(const float* __restrict__ x, const float* __restrict__ u, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = x[i] + u[i];
}
__global__ void normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float s, float t) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = x[i] / s + t;
}

// ---- per-sample kernels with an ldj reduction: one CTA per sample ---------------------------------------------
__global__ void logit_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ ldj, int n) {
  __shared__ float red[32];
  const int64_t b = blockIdx.x;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float v = x[b * n + i];
    float l0 = logf(v), l1 = logf(1.f - v);
    y[b * n + i] = l0 - l1;
    acc += -l0 - l1;
  }
  acc = group_sum(acc, blockDim.x, red);
  if (threadIdx.x == 0) ldj[b] = acc;
}

__global__ void augment_kernel(const float* __restrict__ x, const float* __restrict__ eps, float* __restrict__ y,
                               float* __restrict__ ldj, int C, int A, int HW) {
  __shared__ float red[32];
  const int64_t b = blockIdx.x;
  const int nx = C * HW, ne = A * HW;
  float* yb = y + b * (int64_t)(nx + ne);
  for (int i = threadIdx.x; i < nx; i += blockDim.x) yb[i] = x[b * nx + i];
  float acc = 0.f;
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    float e = eps[b * ne + i];
    yb[nx + i] = e;
    acc += kHalfLog2Pi + 0.5f * e * e;
  }
  acc = group_sum(acc, blockDim.x, red);
  if (threadIdx.x == 0) ldj[b] = acc;
}

__global__ void prologue_kernel(const float* __restrict__ x, const float* __restrict__ u, const float* __restrict__ eps,
                                float* __restrict__ y, float* __restrict__ ldj, int C, int A, int HW,
                                float s0, float t0, float s1, float t1, float ldj_const) {
  __shared__ float red[32];
  const int64_t b = blockIdx.x;
  const int nx = C * HW, ne = A * HW;
  float* yb = y + b * (int64_t)(nx + ne);
  float acc = 0.f;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) {
    float v = x[b * nx + i] + u[b * nx + i];
    v = v / s0 + t0;
    v = v / s1 + t1;
    float l0 = logf(v), l1 = logf(1.f - v);
    yb[i] = l0 - l1;
    acc += -l0 - l1;
  }
  for (int i = threadIdx.x; i < ne; i += blockDim.x) {
    float e = eps[b * ne + i];
    yb[nx + i] = e;
    acc += kHalfLog2Pi + 0.5f * e * e;
  }
  acc = group_sum(acc, blockDim.x, red);
  if (threadIdx.x == 0) ldj[b] = acc + ldj_const;
}

static inline int grid_for(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace cfpp
using namespace cfpp;

extern "C" int cfpp_squeeze_fwd(const float* x, float* y, int B, int C, int H, int W, int p1, int p2, void* stream) {
  CFPP_REQUIRE(B >= 0 && C > 0 && p1 > 0 && p2 > 0 && H % p1 == 0 && W % p2 == 0, "squeeze: bad dims C=%d H=%d W=%d p=(%d,%d)", C, H, W, p1, p2);
  int64_t total = (int64_t)B * C * H * W;
  if (!total) return CFPP_OK;
  if (p1 == 2 && p2 == 2 && W % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0)
    squeeze22_kernel<<<grid_for(total / 8, 256), 256, 0, (cudaStream_t)stream>>>(x, y, total / 8, H, W, C, 0);
  else
    squeeze_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, total, C, H, W, p1, p2, false);
  return check_launch("squeeze_fwd");
}
// Squeeze of the first C channels of every sample of a wider tensor (batch stride x_bstride floats >= C*H*W): SplitPrior's z = x[:, :C]
// (splitprior.py:13) followed by the next block's Squeeze (model.py:125-127) without the copy in between.  2x2 squeezes of 16-byte aligned
// rows only (the BASELINE image stacks); CFPP_ERR_UNSUPPORTED otherwise (the caller slices first).
extern "C" int cfpp_squeeze_strided_fwd(const float* x, int64_t x_bstride, float* y, int B, int C, int H, int W, int p1, int p2, void* stream) {
  CFPP_REQUIRE(B >= 0 && C > 0 && p1 > 0 && p2 > 0 && H % p1 == 0 && W % p2 == 0 && x_bstride >= (int64_t)C * H * W, "squeeze_strided: bad dims");
  int64_t total = (int64_t)B * C * H * W;
  if (!total) return CFPP_OK;
  if (!(p1 == 2 && p2 == 2 && W % 8 == 0 && x_bstride % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0)) {
    set_error("squeeze_strided: only 2x2 squeezes of 16-byte aligned rows (W %% 8 == 0) read through a batch stride");
    return CFPP_ERR_UNSUPPORTED;
  }
  squeeze22_kernel<<<grid_for(total / 8, 256), 256, 0, (cudaStream_t)stream>>>(x, y, total / 8, H, W, C, x_bstride - (int64_t)C * H * W);
  return check_launch("squeeze_fwd");
}
extern "C" int cfpp_squeeze_inv(const float* y, float* x, int B, int C, int H, int W, int p1, int p2, void* stream) {
  CFPP_REQUIRE(B >= 0 && C > 0 && p1 > 0 && p2 > 0 && H % p1 == 0 && W % p2 == 0, "squeeze_inv: bad dims");
  int64_t total = (int64_t)B * C * H * W;
  if (!total) return CFPP_OK;
  if (p1 == 2 && p2 == 2 && W % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0)
    unsqueeze22_kernel<<<grid_for(total / 8, 256), 256, 0, (cudaStream_t)stream>>>(y, x, total / 8, H, W);
  else
    squeeze_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(y, x, total, C, H, W, p1, p2, true);
  return check_launch("squeeze_inv");
}
extern "C" int cfpp_permute_fwd(const float* x, float* y, int B, int C, int H, int W, void* stream) {
  CFPP_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0, "permute: bad dims");
  if (!B) return CFPP_OK;
  CFPP_REQUIRE((int64_t)B * W <= 65535, "permute: B*W=%lld exceeds grid.z", (long long)B * W);
  dim3 grid((H + 31) / 32, (C + 31) / 32, B * W), block(32, 8);
  permute_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, y, C, H, W);
  return check_launch("permute_fwd");
}
extern "C" int cfpp_slice_channels(const float* x, float* y, int B, int C, int HW, int c0, int Cn, void* stream) {
  CFPP_REQUIRE(c0 >= 0 && Cn > 0 && c0 + Cn <= C, "slice: bad channel range");
  int64_t total = (int64_t)B * Cn * HW;
  if (!total) return CFPP_OK;
  slice_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, y, total, C, HW, c0, Cn);
  return check_launch("slice_channels");
}
extern "C" int cfpp_place_channels(const float* src, float* dst, int B, int C, int HW, int c0, int Cn, void* stream) {
  CFPP_REQUIRE(c0 >= 0 && Cn > 0 && c0 + Cn <= C, "place: bad channel range");
  int64_t total = (int64_t)B * Cn * HW;
  if (!total) return CFPP_OK;
  place_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, dst, total, C, HW, c0, Cn);
  return check_launch("place_channels");
}
extern "C" int cfpp_windows_fwd(const void* ts, int ts_is_f64, const int64_t* end, int64_t end0, int64_t stride, float* x,
                                int B, int64_t n_rows, int D, int L, void* stream) {
  CFPP_REQUIRE(n_rows >= 1 && D >= 1 && L >= 1 && (size_t)L * (D + 1) * sizeof(float) <= 48 * 1024, "windows: rows=%lld D=%d L=%d", (long long)n_rows, D, L);
  if (B <= 0) return CFPP_OK;
  const size_t smem = (size_t)L * (D + 1) * sizeof(float);
  if (ts_is_f64) windows_kernel<double><<<B, 256, smem, (cudaStream_t)stream>>>((const double*)ts, end, end0, stride, x, n_rows, D, L);
  else windows_kernel<float><<<B, 256, smem, (cudaStream_t)stream>>>((const float*)ts, end, end0, stride, x, n_rows, D, L);
  return check_launch("windows_fwd");
}
extern "C" int cfpp_add_fwd(const float* x, const float* u, float* y, int64_t n, void* stream) {
  if (n <= 0) return CFPP_OK;
  add_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, u, y, n);
  return check_launch("add_fwd");
}
extern "C" int cfpp_normalize_fwd(const float* x, float* y, int64_t n, float scale, float translation, void* stream) {
  if (n <= 0) return CFPP_OK;
  normalize_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, scale, translation);
  return check_launch("normalize_fwd");
}
extern "C" int cfpp_logit_fwd(const float* x, float* y, float* ldj, int B, int n, void* stream) {
  if (B <= 0) return CFPP_OK;
  logit_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, y, ldj, n);
  return check_launch("logit_fwd");
}
extern "C" int cfpp_augment_fwd(const float* x, const float* eps, float* y, float* ldj, int B, int C, int A, int HW, void* stream) {
  if (B <= 0) return CFPP_OK;
  augment_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, eps, y, ldj, C, A, HW);
  return check_launch("augment_fwd");
}
extern "C" int cfpp_prologue_fwd(const float* x, const float* u, const float* eps, float* y, float* ldj, int B, int C, int A, int HW,
                                 float s0, float t0, float s1, float t1, float ldj_const, void* stream) {
  if (B <= 0) return CFPP_OK;
  CFPP_REQUIRE(A == 0 || eps != nullptr, "prologue: eps required when A>0");
  prologue_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, u, eps, y, ldj, C, A, HW, s0, t0, s1, t1, ldj_const);
  return check_launch("prologue_fwd");
}

// FlowSequential.forward's `logdet += ldj` (layers/flowsequential.py:23): logdet (B,M) += ldj (B,cols), cols in {1, M}.
namespace cfpp {
__global__ void ldj_accumulate_kernel(float* __restrict__ logdet, const float* __restrict__ ldj, int64_t total, int M, int cols) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    logdet[i] += ldj[(i / M) * cols + (cols == 1 ? 0 : i % M)];
}
}  // namespace cfpp
// The whole `logdet += ldj` chain of one forward in a single launch: out = [last +] (((first|0) + t_0) + t_1) + ... in that order,
// i.e. bit-identical to accumulating the terms one launch at a time.
namespace cfpp {
struct LdjSumArgs { const float* term[CFPP_LDJ_SUM_MAX]; int cols[CFPP_LDJ_SUM_MAX]; int n; };
__global__ void ldj_sum_kernel(float* __restrict__ out, const float* __restrict__ first, const float* __restrict__ last, const LdjSumArgs t,
                               int64_t total, int M) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / M; const int m = (int)(i - b * M);
    float acc = first ? first[i] : 0.f;
    for (int k = 0; k < t.n; ++k) acc += t.term[k][t.cols[k] == 1 ? b : i];
    out[i] = last ? last[i] + acc : acc;
  }
}
}  // namespace cfpp
extern "C" int cfpp_ldj_sum(float* out, const float* first, const float* last, const float* const* terms, const int* cols, int n,
                            int B, int M, void* stream) {
  CFPP_REQUIRE(n >= 0 && n <= CFPP_LDJ_SUM_MAX, "ldj_sum: %d terms (max %d per launch)", n, CFPP_LDJ_SUM_MAX);
  cfpp::LdjSumArgs t; t.n = n;
  for (int k = 0; k < n; ++k) {
    CFPP_REQUIRE(cols[k] == 1 || cols[k] == M, "ldj_sum: term %d has %d columns, expected 1 or %d", k, cols[k], M);
    t.term[k] = terms[k]; t.cols[k] = cols[k];
  }
  const int64_t total = (int64_t)B * M;
  if (total <= 0) return CFPP_OK;
  cfpp::ldj_sum_kernel<<<cfpp::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(out, first, last, t, total, M);
  return cfpp::check_launch("ldj_sum");
}

extern "C" int cfpp_ldj_accumulate(float* logdet, const float* ldj, int B, int M, int cols, void* stream) {
  CFPP_REQUIRE(cols == 1 || cols == M, "ldj_accumulate: ldj has %d columns, expected 1 or %d", cols, M);
  const int64_t total = (int64_t)B * M;
  if (total <= 0) return CFPP_OK;
  cfpp::ldj_accumulate_kernel<<<cfpp::grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(logdet, ldj, total, M, cols);
  return cfpp::check_launch("ldj_accumulate");
}
