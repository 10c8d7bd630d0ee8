// Context encoders (model.py:30-90): categorical embedding + surjective dequantisation, one thread per sample; the
// whole 6-layer inner flow (FC, ActNormFC, CouplingFC) x 2 of the variational / argmax / prob encoders runs in registers
// and local memory of that thread (C <= 64).  Also: embedding lookup and the small CN linear layers.
#include "common.cuh"

namespace cfpp {

struct EmbArgs { const float* tab[CFPP_MAX_CTX]; int n_ctx, width; };

__global__ void embed_lookup_kernel(const int64_t* __restrict__ ctx, EmbArgs a, float* __restrict__ out, int64_t total) {
  const int row = a.n_ctx * a.width;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / row; const int j = idx % row, f = j / a.width, jj = j % a.width;
    out[idx] = a.tab[f][ctx[b * a.n_ctx + f] * a.width + jj];
  }
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ void encode_sample(const int64_t* __restrict__ ctx, const float* __restrict__ noise,
                                              float* __restrict__ c_out, float* __restrict__ logp_out,
                                              const cfpp_enc_desc& d, int emit_stage, int b) {
  const int C = d.C, n = d.n_ctx;
  const int64_t* cb = ctx + (int64_t)b * n;
  float* co = c_out + (int64_t)b * C;

  // the embedded (integer / looked-up) input value at output column j, as the surjection sees it
  auto x_in = [&](int j) -> float {
    if (d.emb == CFPP_EMB_ONEHOT) {                       // _embeddings.py:140-150
      int start = 0;
      for (int f = 0; f < n; ++f) { if (j < start + d.card[f]) return (cb[f] == (int64_t)(j - start)) ? 1.f : 0.f; start += d.card[f]; }
      return 0.f;
    } else if (d.emb == CFPP_EMB_EYE) {                   // _embeddings.py:103-109
      return (float)cb[j < n ? j : n - 1];
    } else if (d.emb == CFPP_EMB_DENSE) {
      return d.dense[(int64_t)b * C + j];
    } else {                                              // CatEmbeddings, _embeddings.py:265-283
      const int f = j / d.emb_dim;
      return d.emb_w[f][cb[f] * d.emb_dim + j % d.emb_dim];
    }
  };

  if (d.type == CFPP_ENC_EYESAMPLE) {                     // dequantize.py:133-136
    for (int j = 0; j < C; ++j) co[j] = x_in(j);
    logp_out[b] = 0.f;
    return;
  }
  if (d.type == CFPP_ENC_UNIFORM) {                       // dequantize.py:55-63
    float l = 0.f;
    for (int j = 0; j < C; ++j) {
      co[j] = (x_in(j) + noise[(int64_t)b * C + j]) / d.qbins[j];
      l += d.ldj_per_dim[j] * (float)C;
    }
    logp_out[b] = l;
    return;
  }

  // ---- inner flow: FlowInvSequential.sample (flowsequential.py:60-69) ----
  float x[CFPP_ENC_MAXC], y[CFPP_ENC_MAXC], h1[2 * CFPP_ENC_MAXC], h2[2 * CFPP_ENC_MAXC];
  float logq = 0.f;
  {                                                       // ConditionalGaussianDistribution.sample, gaussian.py:263-270
    auto emb2 = [&](int j) -> float { const int f = j / d.inner_dim; return d.inner_w[f][cb[f] * d.inner_dim + j % d.inner_dim]; };
    for (int j = 0; j < C; ++j) {
      const float mean = emb2(j), ls = emb2(C + j);
      const float v = mean + expf(ls) * noise[(int64_t)b * C + j];
      x[j] = v;
      const float df = v - mean;
      logq += (-kHalfLog2Pi - ls) + (-0.5f * expf(-2.f * ls) * (df * df));
    }
  }
  if (emit_stage == 2) { for (int j = 0; j < C; ++j) co[j] = x[j]; logp_out[b] = logq; return; }
  const int Ch = C / 2, H2 = 2 * C;
  // One thread evaluates the whole inner flow of its sample; weights are warp-uniform global loads.  A scalar inner product costs two
  // loads per FMA (weight + activation) and is load-pipe bound, so when C % 4 == 0 the products are blocked by four: one 128-bit
  // weight load + one activation load per 4 FMAs.
  const bool v4 = (C & 3) == 0;
  for (int L = 0; L < 2; ++L) {
    const float* NN = d.fc[L];                            // FC: conv1x1.py:80-96 (H=W=1)
    if (v4 && (reinterpret_cast<uintptr_t>(NN) & 15) == 0) {
      for (int i = 0; i < C; ++i) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float4* row = reinterpret_cast<const float4*>(NN + i * C);
        for (int j = 0; j < C; j += 4) {
          const float4 w = __ldg(row + (j >> 2));
          a0 = fmaf(w.x, x[j], a0); a1 = fmaf(w.y, x[j + 1], a1); a2 = fmaf(w.z, x[j + 2], a2); a3 = fmaf(w.w, x[j + 3], a3);
        }
        y[i] = (a0 + a1) + (a2 + a3);
      }
    } else {
      for (int i = 0; i < C; ++i) {
        float acc = 0.f;
        for (int j = 0; j < C; ++j) acc = fmaf(NN[i * C + j], x[j], acc);
        y[i] = acc;
      }
    }
    logq -= d.fc_logabsdet[L][0];
    if (emit_stage == L) { for (int j = 0; j < C; ++j) co[j] = y[j]; logp_out[b] = 0.f; return; }
    float sl = 0.f;                                       // ActNormFC: actnorm.py:86-102
    for (int i = 0; i < C; ++i) { const float lg = d.an_logs[L][i]; x[i] = (y[i] - d.an_t[L][i]) * expf(-lg); sl += lg; }
    logq -= sl;
    const bool w4ok = v4 && ((reinterpret_cast<uintptr_t>(d.cw1t[L]) | reinterpret_cast<uintptr_t>(d.cw2t[L]) | reinterpret_cast<uintptr_t>(d.cw3t[L])) & 15) == 0;
    if (w4ok) {                                           // CouplingFC: coupling.py:80-97, four outputs per pass
      for (int o = 0; o < H2; o += 4) {
        float a0 = d.cb1[L][o], a1 = d.cb1[L][o + 1], a2 = d.cb1[L][o + 2], a3 = d.cb1[L][o + 3];
        for (int k = 0; k < Ch; ++k) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(d.cw1t[L] + k * H2 + o)); const float xk = x[k];
          a0 = fmaf(xk, w.x, a0); a1 = fmaf(xk, w.y, a1); a2 = fmaf(xk, w.z, a2); a3 = fmaf(xk, w.w, a3);
        }
        h1[o] = fmaxf(a0, 0.f); h1[o + 1] = fmaxf(a1, 0.f); h1[o + 2] = fmaxf(a2, 0.f); h1[o + 3] = fmaxf(a3, 0.f);
      }
      for (int o = 0; o < H2; o += 4) {
        float a0 = d.cb2[L][o], a1 = d.cb2[L][o + 1], a2 = d.cb2[L][o + 2], a3 = d.cb2[L][o + 3];
        for (int k = 0; k < H2; ++k) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(d.cw2t[L] + k * H2 + o)); const float hk = h1[k];
          a0 = fmaf(hk, w.x, a0); a1 = fmaf(hk, w.y, a1); a2 = fmaf(hk, w.z, a2); a3 = fmaf(hk, w.w, a3);
        }
        h2[o] = fmaxf(a0, 0.f); h2[o + 1] = fmaxf(a1, 0.f); h2[o + 2] = fmaxf(a2, 0.f); h2[o + 3] = fmaxf(a3, 0.f);
      }
      for (int o = 0; o < C; o += 4) {                    // (t | r) = cb3 + h2 W3, kept in y
        float a0 = d.cb3[L][o], a1 = d.cb3[L][o + 1], a2 = d.cb3[L][o + 2], a3 = d.cb3[L][o + 3];
        for (int k = 0; k < H2; ++k) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(d.cw3t[L] + k * C + o)); const float hk = h2[k];
          a0 = fmaf(hk, w.x, a0); a1 = fmaf(hk, w.y, a1); a2 = fmaf(hk, w.z, a2); a3 = fmaf(hk, w.w, a3);
        }
        y[o] = a0; y[o + 1] = a1; y[o + 2] = a2; y[o + 3] = a3;
      }
      float ssum = 0.f;
      for (int o = 0; o < Ch; ++o) {
        const float ls = 2.0f * tanhf(y[Ch + o] * 0.5f);
        x[Ch + o] = fmaf(x[Ch + o], expf(ls), y[o]);
        ssum += ls;
      }
      logq -= ssum;
      continue;
    }
    for (int o = 0; o < H2; ++o) {                        // CouplingFC: coupling.py:80-97
      float acc = d.cb1[L][o];
      for (int k = 0; k < Ch; ++k) acc = fmaf(x[k], d.cw1t[L][k * H2 + o], acc);
      h1[o] = fmaxf(acc, 0.f);
    }
    for (int o = 0; o < H2; ++o) {
      float acc = d.cb2[L][o];
      for (int k = 0; k < H2; ++k) acc = fmaf(h1[k], d.cw2t[L][k * H2 + o], acc);
      h2[o] = fmaxf(acc, 0.f);
    }
    float ssum = 0.f;
    for (int o = 0; o < Ch; ++o) {
      float t = d.cb3[L][o], r = d.cb3[L][Ch + o];
      for (int k = 0; k < H2; ++k) { t = fmaf(h2[k], d.cw3t[L][k * C + o], t); r = fmaf(h2[k], d.cw3t[L][k * C + Ch + o], r); }
      const float ls = 2.0f * tanhf(r * 0.5f);
      x[Ch + o] = fmaf(x[Ch + o], expf(ls), t);
      ssum += ls;
    }
    logq -= ssum;
  }
  // ---- sigmoid flow (activations.py:234-238) and the surjection ----
  const float T = d.temperature[0], logT = logf(T);
  float act = 0.f;
  for (int j = 0; j < C; ++j) {
    const float v = T * x[j];
    act += logT - softplus_f(-v) - softplus_f(v);
    x[j] = sigmoid_f(v);
  }
  if (d.type == CFPP_ENC_VARDEQ) {                        // dequantize.py:107-116
    float l = 0.f;
    for (int j = 0; j < C; ++j) { co[j] = (x_in(j) + x[j]) / d.qbins[j]; l += d.ldj_per_dim[j] * (float)C; }
    logp_out[b] = (l + act) - logq;
  } else if (d.type == CFPP_ENC_ARGMAX) {                 // dequantize.py:239-268: bits MSB first, zero-padded to even
    int j = 0;
    for (int f = 0; f < n; ++f)
      for (int k = d.bits[f] - 1; k >= 0; --k, ++j) { const int bit = (int)((cb[f] >> k) & 1); co[j] = x[j] * (float)(2 * bit - 1); }
    for (; j < C; ++j) co[j] = -x[j];
    logp_out[b] = act - logq;
  } else {                                                // prob sampling, dequantize.py:152-161
    for (int j = 0; j < C; ++j) co[j] = x[j];
    logp_out[b] = act + logq;
  }
}

__global__ void __launch_bounds__(128) ctx_encode_kernel(const int64_t* __restrict__ ctx, const float* __restrict__ noise,
                                                         float* __restrict__ c_out, float* __restrict__ logp_out,
                                                         const cfpp_enc_desc d, int emit_stage, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) encode_sample(ctx, noise, c_out, logp_out, d, emit_stage, b);
}

// All the independent encoders of a run of flow layers in ONE launch: blockIdx.y selects the encoder.  Each layer of a
// specialist model owns its own encoder (36 per forward for the CIFAR/ATM stacks); launched one by one they occupy a
// fraction of the SMs and are latency bound.
struct EncBatchPtrs { const float* noise[CFPP_MAX_ENC_BATCH]; float* c[CFPP_MAX_ENC_BATCH]; float* logp[CFPP_MAX_ENC_BATCH]; };

__global__ void __launch_bounds__(128) ctx_encode_batch_kernel(const int64_t* __restrict__ ctx, const cfpp_enc_desc* __restrict__ descs,
                                                               const EncBatchPtrs p, int B) {
  __shared__ cfpp_enc_desc d;
  const int e = blockIdx.y;
  for (int i = threadIdx.x; i < (int)(sizeof(cfpp_enc_desc) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&d)[i] = reinterpret_cast<const uint32_t*>(descs + e)[i];
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) encode_sample(ctx, p.noise[e], p.c[e], p.logp[e], d, -1, b);
}

// y (B,N) = act(x (B,K) @ wt (K,N) + bias): 64x64 tile, 4x4 per thread, K chunks of 16.
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                                                     float* __restrict__ y, int B, int K, int N, int relu) {
  __shared__ float xs[16][64 + 1];
  __shared__ float wsm[16][64 + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int r = i / 16, k = i % 16;
      xs[k][r] = (r0 + r < B && k0 + k < K) ? x[(int64_t)(r0 + r) * K + k0 + k] : 0.f;
      const int kk = i / 64, c = i % 64;
      wsm[kk][c] = (k0 + kk < K && c0 + c < N) ? wt[(int64_t)(k0 + kk) * N + c0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = xs[k][ty * 4 + i]; w[i] = wsm[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c < N) { float v = acc[i][j] + (bias ? bias[c] : 0.f); y[(int64_t)r * N + c] = relu ? fmaxf(v, 0.f) : v; }
    }
  }
}

}  // namespace cfpp
using namespace cfpp;

extern "C" int cfpp_embed_lookup(const int64_t* ctx, const float* const* tables, int n_ctx, int width, float* out, int B, void* stream) {
  CFPP_REQUIRE(n_ctx >= 1 && n_ctx <= CFPP_MAX_CTX && width >= 1, "embed_lookup: n_ctx=%d width=%d", n_ctx, width);
  if (B <= 0) return CFPP_OK;
  EmbArgs a; a.n_ctx = n_ctx; a.width = width;
  for (int i = 0; i < n_ctx; ++i) a.tab[i] = tables[i];
  const int64_t total = (int64_t)B * n_ctx * width;
  int64_t blocks = (total + 255) / 256, cap = (int64_t)num_sms() * 16;
  embed_lookup_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, (cudaStream_t)stream>>>(ctx, a, out, total);
  return check_launch("embed_lookup");
}

extern "C" int cfpp_ctx_encode(const int64_t* ctx, const float* noise, float* c, float* logp_c, const cfpp_enc_desc* desc,
                               int emit_stage, int B, void* stream) {
  CFPP_REQUIRE(desc != nullptr, "ctx_encode: null descriptor");
  const cfpp_enc_desc& d = *desc;
  CFPP_REQUIRE(d.n_ctx >= 1 && d.n_ctx <= CFPP_MAX_CTX, "ctx_encode: n_ctx=%d", d.n_ctx);
  CFPP_REQUIRE(d.C >= 1 && d.C <= CFPP_ENC_MAXC, "ctx_encode: width C=%d outside [1,%d]", d.C, CFPP_ENC_MAXC);
  CFPP_REQUIRE(d.emb >= 0 && d.emb <= 3 && d.type >= 0 && d.type <= 4, "ctx_encode: bad emb/type");
  const bool flow = d.type >= CFPP_ENC_VARDEQ;
  CFPP_REQUIRE(!flow || emit_stage == 2 || (d.C % 2 == 0), "ctx_encode: inner flow needs an even width");
  CFPP_REQUIRE(d.type == CFPP_ENC_EYESAMPLE || noise != nullptr, "ctx_encode: noise required");
  CFPP_REQUIRE(emit_stage >= -1 && emit_stage <= 2, "ctx_encode: emit_stage");
  if (B <= 0) return CFPP_OK;
  ctx_encode_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx, noise, c, logp_c, d, emit_stage, B);
  return check_launch("ctx_encode");
}

extern "C" int cfpp_ctx_encode_batch(const int64_t* ctx, const cfpp_enc_desc* descs_device, int n_enc, const float* const* noise,
                                     float* const* c_out, float* const* logp_out, int B, void* stream) {
  CFPP_REQUIRE(n_enc >= 1 && n_enc <= CFPP_MAX_ENC_BATCH, "ctx_encode_batch: n_enc=%d outside [1,%d]", n_enc, CFPP_MAX_ENC_BATCH);
  CFPP_REQUIRE(descs_device && noise && c_out && logp_out, "ctx_encode_batch: null argument");
  if (B <= 0) return CFPP_OK;
  EncBatchPtrs p;
  for (int i = 0; i < n_enc; ++i) { p.noise[i] = noise[i]; p.c[i] = c_out[i]; p.logp[i] = logp_out[i]; }
  dim3 grid((B + 127) / 128, n_enc);
  ctx_encode_batch_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(ctx, descs_device, p, B);
  return check_launch("ctx_encode_batch");
}

extern "C" int cfpp_linear_fwd(const float* x, const float* wt, const float* b, float* y, int B, int K, int N, int relu, void* stream) {
  CFPP_REQUIRE(K >= 1 && N >= 1, "linear: K=%d N=%d", K, N);
  if (B <= 0) return CFPP_OK;
  dim3 grid((N + 63) / 64, (B + 63) / 64);
  CFPP_REQUIRE(grid.y <= 65535, "linear: batch too large for one launch");
  linear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, wt, b, y, B, K, N, relu);
  return check_launch("linear_fwd");
}
