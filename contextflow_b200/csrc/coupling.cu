// The fused memory-bound affine-coupling transform (layers/coupling.py:50-66, :131-148).
// One HBM pass: read x and h once (128-bit streaming loads), write z once, reduce log_s into ldj[b] with warp
// shuffles + one shared-memory hop.  Algorithmic traffic: 12*C*HW bytes per sample.
#include "common.cuh"

namespace cfpp {

__device__ __forceinline__ void couple1(float x1, float t, float r, float& z1, float& acc) {
  const float ls = 2.0f * tanhf(r * 0.5f);      // log_s = logs_range * tanh(h / logs_range), logs_range = 2
  z1 = fmaf(x1, expf(ls), t);
  acc += ls;
}

// G threads cooperate on one sample; blockDim/G samples per CTA.
template <bool VEC>
__global__ void __launch_bounds__(256) coupling_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                       const float* __restrict__ add, const float* __restrict__ logp_c,
                                                       float logp_scale, float* __restrict__ z, float* __restrict__ ldj,
                                                       int B, int C, int HW, int G) {
  __shared__ float red[32];
  const int spc = blockDim.x / G;
  const int s = threadIdx.x / G, g = threadIdx.x % G;
  const int64_t b = (int64_t)blockIdx.x * spc + s;
  const bool valid = b < B;
  const int Ch = C / 2;
  const int64_t n = (int64_t)Ch * HW;           // elements per half
  float acc = 0.f;
  if (valid) {
    const float* xb = x + b * 2 * n; const float* hb = h + b * 2 * n; float* zb = z + b * 2 * n;
    const float* ab = add ? add + b * C : nullptr;
    if (VEC) {
      const int n4 = (int)(n / 4), hw4 = HW / 4;
      const float4* x0 = reinterpret_cast<const float4*>(xb); const float4* x1 = x0 + n4;
      const float4* ht = reinterpret_cast<const float4*>(hb); const float4* hr = ht + n4;
      float4* z0 = reinterpret_cast<float4*>(zb); float4* z1 = z0 + n4;
#pragma unroll 2
      for (int i = g; i < n4; i += G) {
        const float4 a0 = ldg_stream(x0 + i), a1 = ldg_stream(x1 + i);
        float4 t = ldg_stream(ht + i), r = ldg_stream(hr + i);
        if (ab) {
          const int ch = i / hw4;
          const float at = ab[ch], ar = ab[Ch + ch];
          t.x += at; t.y += at; t.z += at; t.w += at;
          r.x += ar; r.y += ar; r.z += ar; r.w += ar;
        }
        float4 o;
        couple1(a1.x, t.x, r.x, o.x, acc); couple1(a1.y, t.y, r.y, o.y, acc);
        couple1(a1.z, t.z, r.z, o.z, acc); couple1(a1.w, t.w, r.w, o.w, acc);
        stg_stream(z0 + i, a0);
        stg_stream(z1 + i, o);
      }
    } else {
      for (int64_t i = g; i < n; i += G) {
        float t = hb[i], r = hb[n + i];
        if (ab) { const int ch = (int)(i / HW); t += ab[ch]; r += ab[Ch + ch]; }
        float o;
        couple1(xb[n + i], t, r, o, acc);
        zb[i] = xb[i];
        zb[n + i] = o;
      }
    }
  }
  acc = group_sum(acc, G, red);
  if (valid && g == 0) ldj[b] = acc + (logp_c ? logp_scale * logp_c[b] : 0.f);
}

// MaskedCoupling.forward (layers/ar.py:35-57): the block output carries the identity x on both halves (masked_conv_2d.py:92,98), so
// t = h[:, :C] + x, r = h[:, C:] + x; log_s = 2 tanh(r/2); z = x e^{log_s} + t over ALL C channels; ldj[b] = sum log_s.
// h is (B, 2C, HW).  16*C*HW bytes per sample.
// --contextflow specialist (ar.py:39-42): h += CN(c) (`add`, (B, 2C), broadcast over the pixels), ldj += logp_scale * logp_c[b].
__global__ void __launch_bounds__(256) maf_coupling_kernel(const float* __restrict__ x, const float* __restrict__ h, const float* __restrict__ add,
                                                           const float* __restrict__ logp_c, float logp_scale, float* __restrict__ z,
                                                           float* __restrict__ ldj, int B, int C, int HW, int G) {
  __shared__ float red[32];
  const int spc = blockDim.x / G;
  const int s = threadIdx.x / G, g = threadIdx.x % G;
  const int64_t b = (int64_t)blockIdx.x * spc + s;
  const bool valid = b < B;
  const int64_t n = (int64_t)C * HW;
  float acc = 0.f;
  if (valid) {
    const float* xb = x + b * n; const float* hb = h + b * 2 * n; float* zb = z + b * n;
    const float* ab = add ? add + b * 2 * C : nullptr;
    for (int64_t i = g; i < n; i += G) {
      const float xv = xb[i];
      float at = 0.f, ar = 0.f;
      if (ab) { const int ch = (int)(i / HW); at = ab[ch]; ar = ab[C + ch]; }
      float o;
      couple1(xv, hb[i] + at + xv, hb[n + i] + ar + xv, o, acc);
      zb[i] = o;
    }
  }
  acc = group_sum(acc, G, red);
  if (valid && g == 0) ldj[b] = acc + (logp_c ? logp_scale * logp_c[b] : 0.f);
}

// MaskedCoupling backward: z = x s + t with t = h_t + x, r = h_r + x, s = exp(2 tanh(r/2)), ldj = sum 2 tanh(r/2).
//   dr = (dz x s + dldj[b]) (1 - tanh^2(r/2));  dh = cat(dz, dr);  dx = dz s + dz + dr   (the conditioner's part is added by its own backward).
__global__ void __launch_bounds__(256) maf_coupling_bwd_kernel(const float* __restrict__ x, const float* __restrict__ h, const float* __restrict__ dz,
                                                               const float* __restrict__ dldj, float* __restrict__ dx, float* __restrict__ dh,
                                                               int64_t total, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / n, e = i - b * n;
    const float xv = x[i], g = dz[i];
    const float th = tanhf((h[b * 2 * n + n + e] + xv) * 0.5f);
    const float sc = expf(2.0f * th);
    const float dr = (g * xv * sc + (dldj ? dldj[b] : 0.f)) * (1.0f - th * th);
    dh[b * 2 * n + e] = g;
    dh[b * 2 * n + n + e] = dr;
    dx[i] = g * sc + g + dr;
  }
}

}  // namespace cfpp
using namespace cfpp;

extern "C" int cfpp_coupling_fwd(const float* x, const float* h, const float* add, const float* logp_c, float logp_scale,
                                 float* z, float* ldj, int B, int C, int HW, void* stream) {
  CFPP_REQUIRE(C >= 2 && C % 2 == 0 && HW >= 1, "coupling: C=%d must be even, HW=%d", C, HW);
  if (B <= 0) return CFPP_OK;
  const int64_t n = (int64_t)(C / 2) * HW;
  const bool vec = (HW % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(h) | reinterpret_cast<uintptr_t>(z)) % 16 == 0);
  const int64_t work = vec ? n / 4 : n;
  int G = 32;
  while (G < 256 && G * 4 < work) G <<= 1;
  const int spc = 256 / G;
  const int blocks = (int)((B + spc - 1) / spc);
  if (vec) coupling_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, h, add, logp_c, logp_scale, z, ldj, B, C, HW, G);
  else coupling_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, h, add, logp_c, logp_scale, z, ldj, B, C, HW, G);
  return check_launch("coupling_fwd");
}

extern "C" int cfpp_maf_coupling_ctx_fwd(const float* x, const float* h, const float* add, const float* logp_c, float logp_scale,
                                         float* z, float* ldj, int B, int C, int HW, void* stream) {
  CFPP_REQUIRE(C >= 1 && HW >= 1, "maf_coupling: C=%d HW=%d", C, HW);
  if (B <= 0) return CFPP_OK;
  const int64_t n = (int64_t)C * HW;
  int G = 32;
  while (G < 256 && G * 4 < n) G <<= 1;
  const int spc = 256 / G;
  maf_coupling_kernel<<<(B + spc - 1) / spc, 256, 0, (cudaStream_t)stream>>>(x, h, add, logp_c, logp_scale, z, ldj, B, C, HW, G);
  return check_launch("maf_coupling_fwd");
}

extern "C" int cfpp_maf_coupling_fwd(const float* x, const float* h, float* z, float* ldj, int B, int C, int HW, void* stream) {
  return cfpp_maf_coupling_ctx_fwd(x, h, nullptr, nullptr, 0.f, z, ldj, B, C, HW, stream);
}

extern "C" int cfpp_maf_coupling_bwd(const float* x, const float* h, const float* dz, const float* dldj, float* dx, float* dh,
                                     int B, int C, int HW, void* stream) {
  CFPP_REQUIRE(C >= 1 && HW >= 1, "maf_coupling_bwd: C=%d HW=%d", C, HW);
  if (B <= 0) return CFPP_OK;
  const int64_t n = (int64_t)C * HW, total = (int64_t)B * n;
  int64_t blocks = (total + 255) / 256; const int64_t cap = (int64_t)num_sms() * 16; if (blocks > cap) blocks = cap;
  maf_coupling_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, h, dz, dldj, dx, dh, total, n);
  return check_launch("maf_coupling_bwd");
}
