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

// ---- tiled form of the inner-flow encoders (vardeq / argmax / probsample) for a compile-time width C ------------------------------
// The one-thread-per-sample walk above keeps x, y, h1, h2 in local memory (runtime C) and reads every weight from global memory at
// dependent-load latency: ncu (r1u) showed 10 long-scoreboard stall cycles per issued instruction.  Here a CTA = 128 samples of ONE encoder:
// the encoder's 2 x (FC, ActNormFC, CouplingFC) parameters are staged once in shared memory (FC transposed to K-major), every layer is
//   out[o..o+3] = bias + sum_k W[k][o..o+3] * in[k]        k unrolled at compile time: `in` lives in registers,
// the four weights arrive as one 128-bit warp-broadcast shared load per 4 FMAs, and the outputs bounce through the thread's own column of a
// shared buffer ([feature][129]: conflict free) so that the next layer reads them back with static register indices.  Noise rows come in
// and c rows go out through the same buffer, coalesced.
constexpr int kEncTile = 128, kEncStride = kEncTile + 1;

template <int C> struct EncFlowLayout {
  static constexpr int Ch = C / 2, H2 = 2 * C;
  static constexpr int fcT = 0, ant = fcT + C * C, ansc = ant + C, w1 = ansc + C, b1 = w1 + Ch * H2, w2 = b1 + H2, b2 = w2 + H2 * H2,
                       w3 = b2 + H2, b3 = w3 + H2 * C, per_layer = b3 + C;
  static constexpr int floats = 2 * per_layer + H2 * kEncStride + 4;          // + {fc_logabsdet[2], sum an_logs[2]}
};

template <int KIN, int NOUT, bool BIAS, bool RELU>
__device__ __forceinline__ void enc_matvec(const float* __restrict__ W, const float* __restrict__ bias, const float (&in)[KIN], float* __restrict__ col) {
#pragma unroll 2
  for (int o = 0; o < NOUT; o += 4) {
    float4 acc = BIAS ? *reinterpret_cast<const float4*>(bias + o) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < KIN; ++k) {
      const float4 w = *reinterpret_cast<const float4*>(W + k * NOUT + o);
      acc.x = fmaf(in[k], w.x, acc.x); acc.y = fmaf(in[k], w.y, acc.y); acc.z = fmaf(in[k], w.z, acc.z); acc.w = fmaf(in[k], w.w, acc.w);
    }
    if (RELU) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    col[(o + 0) * kEncStride] = acc.x; col[(o + 1) * kEncStride] = acc.y; col[(o + 2) * kEncStride] = acc.z; col[(o + 3) * kEncStride] = acc.w;
  }
}

template <int C>
__global__ void __launch_bounds__(kEncTile) ctx_encode_flow_kernel(const int64_t* __restrict__ ctx, const cfpp_enc_desc* __restrict__ descs,
                                                                   const EncBatchPtrs p, int B) {
  using Lo = EncFlowLayout<C>;
  constexpr int Ch = Lo::Ch, H2 = Lo::H2;
  extern __shared__ float4 enc_smem4[];
  __shared__ cfpp_enc_desc d;
  float* sm = reinterpret_cast<float*>(enc_smem4);
  float* bounce = sm + 2 * Lo::per_layer;                   // [H2][kEncStride]
  float* scal = bounce + H2 * kEncStride;                   // fc_logabsdet[0..1], sum an_logs[0..1]
  const int e = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < (int)(sizeof(cfpp_enc_desc) / 4); i += kEncTile)
    reinterpret_cast<uint32_t*>(&d)[i] = reinterpret_cast<const uint32_t*>(descs + e)[i];
  __syncthreads();
  const float* noise = p.noise[e];
  // ---- stage the encoder's parameters and this tile's noise rows ----
  for (int L = 0; L < 2; ++L) {
    float* w = sm + L * Lo::per_layer;
    for (int i = tid; i < C * C; i += kEncTile) { const int r = i / C, c = i - r * C; w[Lo::fcT + c * C + r] = __ldg(d.fc[L] + i); }   // FC: y = NN x (conv1x1.py:80-96)
    for (int i = tid; i < C; i += kEncTile) { w[Lo::ant + i] = __ldg(d.an_t[L] + i); w[Lo::ansc + i] = expf(-__ldg(d.an_logs[L] + i)); w[Lo::b3 + i] = __ldg(d.cb3[L] + i); }
    for (int i = tid; i < Ch * H2; i += kEncTile) w[Lo::w1 + i] = __ldg(d.cw1t[L] + i);
    for (int i = tid; i < H2; i += kEncTile) { w[Lo::b1 + i] = __ldg(d.cb1[L] + i); w[Lo::b2 + i] = __ldg(d.cb2[L] + i); }
    for (int i = tid; i < H2 * H2; i += kEncTile) w[Lo::w2 + i] = __ldg(d.cw2t[L] + i);
    for (int i = tid; i < H2 * C; i += kEncTile) w[Lo::w3 + i] = __ldg(d.cw3t[L] + i);
    if (tid == 32 * L) {                                     // ActNormFC ldj = sum logs (actnorm.py:86-102), in index order like the per-sample walk
      float sl = 0.f;
      for (int i = 0; i < C; ++i) sl += d.an_logs[L][i];
      scal[2 + L] = sl; scal[L] = d.fc_logabsdet[L][0];
    }
  }
  // a CTA walks several 128-sample tiles of its encoder: the parameters are staged once
  const int ntiles = (B + kEncTile - 1) / kEncTile;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
  const int b0 = tile * kEncTile, b = b0 + tid;
  const bool live = b < B;
  for (int idx = tid; idx < kEncTile * C; idx += kEncTile) {
    const int s = idx / C, j = idx - s * C;
    bounce[j * kEncStride + s] = (b0 + s < B) ? __ldg(noise + (int64_t)b0 * C + idx) : 0.f;
  }
  __syncthreads();
  float* col = bounce + tid;
  const int n = d.n_ctx;
  const int64_t* cb = ctx + (int64_t)(live ? b : 0) * n;
  auto x_in = [&](int j) -> float {                          // the embedded input value at output column j (as in encode_sample)
    if (d.emb == CFPP_EMB_ONEHOT) {
      int start = 0;
      for (int f = 0; f < n; ++f) { if (j < start + d.card[f]) return (cb[f] == (int64_t)(j - start)) ? 1.f : 0.f; start += d.card[f]; }
      return 0.f;
    } else if (d.emb == CFPP_EMB_EYE) {
      return (float)cb[j < n ? j : n - 1];
    } else if (d.emb == CFPP_EMB_DENSE) {
      return d.dense[(int64_t)(live ? b : 0) * C + j];
    } else {
      const int f = j / d.emb_dim;
      return d.emb_w[f][cb[f] * d.emb_dim + j % d.emb_dim];
    }
  };
  // ---- ConditionalGaussianDistribution.sample (gaussian.py:263-270) ----
  float x[C], y[C], h[H2];
  float logq = 0.f;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const int f0 = j / d.inner_dim, f1 = (C + j) / d.inner_dim;
    const float mean = d.inner_w[f0][cb[f0] * d.inner_dim + j % d.inner_dim], ls = d.inner_w[f1][cb[f1] * d.inner_dim + (C + j) % d.inner_dim];
    const float v = mean + expf(ls) * col[j * kEncStride];
    x[j] = v;
    const float df = v - mean;
    logq += (-kHalfLog2Pi - ls) + (-0.5f * expf(-2.f * ls) * (df * df));
  }
  // ---- 2 x (FC, ActNormFC, CouplingFC)  (flowsequential.py:60-69) ----
#pragma unroll 1
  for (int L = 0; L < 2; ++L) {
    const float* w = sm + L * Lo::per_layer;
    enc_matvec<C, C, false, false>(w + Lo::fcT, nullptr, x, col);
#pragma unroll
    for (int i = 0; i < C; ++i) x[i] = (col[i * kEncStride] - w[Lo::ant + i]) * w[Lo::ansc + i];
    logq -= scal[L];
    logq -= scal[2 + L];
    float xa[Ch];
#pragma unroll
    for (int i = 0; i < Ch; ++i) xa[i] = x[i];
    enc_matvec<Ch, H2, true, true>(w + Lo::w1, w + Lo::b1, xa, col);            // coupling.py:80-97
#pragma unroll
    for (int i = 0; i < H2; ++i) h[i] = col[i * kEncStride];
    enc_matvec<H2, H2, true, true>(w + Lo::w2, w + Lo::b2, h, col);
#pragma unroll
    for (int i = 0; i < H2; ++i) h[i] = col[i * kEncStride];
    enc_matvec<H2, C, true, false>(w + Lo::w3, w + Lo::b3, h, col);
#pragma unroll
    for (int i = 0; i < C; ++i) y[i] = col[i * kEncStride];
    float ssum = 0.f;
#pragma unroll
    for (int o = 0; o < Ch; ++o) {
      const float ls = 2.0f * tanhf(y[Ch + o] * 0.5f);
      x[Ch + o] = fmaf(x[Ch + o], expf(ls), y[o]);
      ssum += ls;
    }
    logq -= ssum;
  }
  // ---- sigmoid flow (activations.py:234-238) and the surjection ----
  // softplus(-v) + softplus(v) = |v| + 2 log1p(e^-|v|) and sigmoid(v) = {1, e^-|v|} / (1 + e^-|v|): one exponential per element
  const float T = d.temperature[0], logT = logf(T);
  float act = 0.f;
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const float v = T * x[j], av = fabsf(v);
    const float ev = expf(-av);
    act += logT - (av + 2.f * log1pf(ev));
    x[j] = (v >= 0.f ? 1.f : ev) / (1.f + ev);
  }
  float lp;
  if (d.type == CFPP_ENC_VARDEQ) {                          // dequantize.py:107-116
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < C; ++j) { col[j * kEncStride] = (x_in(j) + x[j]) / d.qbins[j]; l += d.ldj_per_dim[j] * (float)C; }
    lp = (l + act) - logq;
  } else if (d.type == CFPP_ENC_ARGMAX) {                   // dequantize.py:239-268: bits MSB first, zero-padded to even
    int f = 0, k = d.bits[0] - 1;
#pragma unroll
    for (int j = 0; j < C; ++j) {
      while (f < n && k < 0) { ++f; k = f < n ? d.bits[f] - 1 : -1; }
      if (f < n) { const int bit = (int)((cb[f] >> k) & 1); col[j * kEncStride] = x[j] * (float)(2 * bit - 1); --k; }
      else col[j * kEncStride] = -x[j];
    }
    lp = act - logq;
  } else {                                                  // prob sampling, dequantize.py:152-161
#pragma unroll
    for (int j = 0; j < C; ++j) col[j * kEncStride] = x[j];
    lp = act + logq;
  }
  if (live) p.logp[e][b] = lp;
  __syncthreads();
  float* c_out = p.c[e];
  for (int idx = tid; idx < kEncTile * C; idx += kEncTile) {
    const int s = idx / C, j = idx - s * C;
    if (b0 + s < B) c_out[(int64_t)b0 * C + idx] = bounce[j * kEncStride + s];
  }
  __syncthreads();
  }
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

extern "C" int cfpp_ctx_encode_flow_supported(int C) { return (C == 8 || C == 20) ? 1 : 0; }

extern "C" int cfpp_ctx_encode_batch_flow(const int64_t* ctx, const cfpp_enc_desc* descs_device, int n_enc, int C, const float* const* noise,
                                          float* const* c_out, float* const* logp_out, int B, void* stream) {
  CFPP_REQUIRE(n_enc >= 1 && n_enc <= CFPP_MAX_ENC_BATCH, "ctx_encode_batch_flow: n_enc=%d outside [1,%d]", n_enc, CFPP_MAX_ENC_BATCH);
  CFPP_REQUIRE(descs_device && noise && c_out && logp_out, "ctx_encode_batch_flow: null argument");
  CFPP_REQUIRE(cfpp_ctx_encode_flow_supported(C), "ctx_encode_batch_flow: width C=%d has no tiled kernel", C);
  for (int i = 0; i < n_enc; ++i) CFPP_REQUIRE(noise[i] && c_out[i] && logp_out[i], "ctx_encode_batch_flow: encoder %d: null noise / output", i);
  if (B <= 0) return CFPP_OK;
  EncBatchPtrs p;
  for (int i = 0; i < n_enc; ++i) { p.noise[i] = noise[i]; p.c[i] = c_out[i]; p.logp[i] = logp_out[i]; }
  // about four resident CTAs per SM in one wave; each CTA stages its encoder's parameters once and walks its share of the sample tiles
  const int ntiles = (B + kEncTile - 1) / kEncTile;
  int per_enc = (4 * num_sms()) / n_enc;
  if (per_enc > ntiles) per_enc = ntiles;
  if (per_enc < 1) per_enc = 1;
  dim3 grid(per_enc, n_enc);
  cudaStream_t st = (cudaStream_t)stream;
#define CFPP_ENCF(C_) do { const int smem = EncFlowLayout<C_>::floats * 4; static DeviceOnce set_; \
    if (set_.first()) { cudaFuncSetAttribute(ctx_encode_flow_kernel<C_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); } \
    ctx_encode_flow_kernel<C_><<<grid, kEncTile, smem, st>>>(ctx, descs_device, p, B); } while (0)
  if (C == 8) CFPP_ENCF(8); else CFPP_ENCF(20);
#undef CFPP_ENCF
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
