// Specialist invertible 1x1 convolution with its context network and the ActNorm epilogue in ONE persistent kernel
// (layers/conv1x1.py:31-50 incl. `c = self.CN(c)` at :33, layers/actnorm.py:37-60).
//
// Per sample the layer is z = W_b x with W_b = tril(c,-1) + diag(exp(diag c)) [- I + NN], c = CN(e) = cnw^T e + cnb a (D, D) matrix that
// only exists to be multiplied once.  Materialising c (cfpp_cn_batch -> HBM -> cfpp_conv1x1_fwd) costs 4*D(D+1)/2 bytes written and read
// per sample -- at D = 64 more than the activations themselves -- and the consumer then spends most of its instructions re-assembling
// W_b from it.  Here persistent CTAs (two per SM where shared memory allows) loop over groups of NS <= 4 samples:
//   stage   : the activations of the NEXT group arrive by one bulk-TMA copy per sample (contiguous D*HW floats) behind an mbarrier while
//             the current groups are computed (two to four activation stages);
//   phase A : every thread owns entries t of the packed lower triangle: 1 coalesced weight load per k (cnw is K-major over the packed
//             triangle: (K, T), L2 resident), reused for the group's samples whose encoder rows e sit in shared memory as [k][s] (one 128-bit
//             broadcast load = four samples; ten weight loads in flight per entry); exp on the diagonal, - I + NN, written into the group's W_b tiles; the strictly upper triangle
//             (NN or 0) is written once per CTA;
//   phase B : the register-tiled product of conv1x1_rt.cu (4 rows x PT pixels per thread, operands from shared memory), ActNorm epilogue,
//             128-bit streaming stores; per-sample ldj by one warp per sample in a fixed order (deterministic).
// HBM traffic per sample: 8*D*HW (+ the K-float encoder row): the algorithmic minimum of the layer.
#include "common.cuh"

namespace cfpp {
namespace c1f {

struct Args {
  const float* x; float* z; float* ldj;
  const float* e; const float* cnw; const float* cnb;
  const float* NN; const float* logabsdet; const float* logp_c; int contextflow;
  const float* an_t; const float* an_logs; int an_stride; const float* an_logp_c; float an_logp_scale;
  int B, D, DR, G, HW, K, T, PG, TPS, NS, WS, ngroups, NSTG;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kThreads = 256;
constexpr int kSlots = 4;                                      // sample slots of the encoder tile = samples per group at most
constexpr int kKB = 10;                                        // weight loads in flight per entry of phase A

// PT: pixels per thread (4 or 8); KT: the encoder width K when it is a compile-time case (20, 8), else 0.  Two or three CTAs per SM (the launch sizes shared memory for it): the global-load latencies of one
// CTA's group (encoder rows, the weight columns of phase A) hide behind the other CTAs' products.
template <int PT, int KT>
__global__ void __launch_bounds__(kThreads, 2) conv1x1_ctx_kernel(const Args a) {
  extern __shared__ float4 c1f_smem4[];
  constexpr int SL = kSlots;
  const int D = a.D, DR = a.DR, HW = a.HW, WS = a.WS, NS = a.NS, K = a.K, T = a.T;
  const int KP = (K + kKB - 1) / kKB * kKB;                    // encoder rows padded with zeros to whole blocks of kKB
  float* Ws = reinterpret_cast<float*>(c1f_smem4);             // [NS][DR][WS]; rows >= D and columns >= D stay zero
  const int NSTG = a.NSTG;                                     // activation stages (2..4): NSTG - 1 groups in flight behind mbarriers
  float* Xs = Ws + (size_t)NS * DR * WS;                       // [NSTG][NS][DR][HW]; rows >= D stay zero
  float* Es = Xs + (size_t)NSTG * NS * DR * HW;                // [KP][SL]
  float* dg = Es + (size_t)KP * SL;                            // [NS][DR] raw diagonal of c
  float* sh = dg + (size_t)NS * DR;                            // [NS][DR] ActNorm shift
  float* sc = sh + (size_t)NS * DR;                            // [NS][DR] ActNorm exp(-logs)
  float* lg = sc + (size_t)NS * DR;                            // [NS][DR] ActNorm logs
  uint32_t* tab = reinterpret_cast<uint32_t*>(lg + (size_t)NS * DR);   // [T] (i << 16) | j of packed entry t
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab + ((T + 1) & ~1));  // NSTG mbarriers (8-byte aligned: every block above is a multiple of 8 bytes)
  const int tid = threadIdx.x;
  const uint32_t bar0 = smem_u32(bars);

  // ---- once per CTA: barriers, zero padding, the triangle index table, the context-free part of W ----
  if (tid == 0) { for (int q = 0; q < NSTG; ++q) mbar_init(bar0 + 8 * q, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (int idx = tid; idx < DR * WS; idx += kThreads) {         // the same context-free entries in every sample slot
    const int i = idx / WS, j = idx - i * WS;
    const float v = (a.contextflow && i < D && j < D && j > i) ? __ldg(a.NN + i * D + j) : 0.f;
    for (int s = 0; s < NS; ++s) Ws[(size_t)s * DR * WS + idx] = v;
  }
  if (DR != D) for (int idx = tid; idx < NSTG * NS * (DR - D) * HW; idx += kThreads) {
    const int blk = idx / ((DR - D) * HW), rem = idx - blk * (DR - D) * HW;
    Xs[((size_t)blk * DR + D) * HW + rem] = 0.f;
  }
  for (int idx = tid; idx < KP * SL; idx += kThreads) Es[idx] = 0.f;
  for (int i = tid; i < D; i += kThreads) { const int t0 = i * (i + 1) / 2; for (int j = 0; j <= i; ++j) tab[t0 + j] = ((uint32_t)i << 16) | (uint32_t)j; }
  __syncthreads();

  const uint32_t xbytes = (uint32_t)(D * HW * 4);
  auto prefetch = [&](int grp, int buf) {                     // thread 0 only
    const int64_t b0 = (int64_t)grp * NS;
    const int nS = (int)min((int64_t)NS, (int64_t)a.B - b0);
    mbar_expect_tx(bar0 + 8 * buf, xbytes * nS);
    for (int m = 0; m < nS; ++m)
      bulk_g2s(smem_u32(Xs + ((size_t)buf * NS + m) * DR * HW), a.x + (b0 + m) * (int64_t)D * HW, xbytes, bar0 + 8 * buf);
  };
  if (tid == 0) for (int q = 0; q < NSTG - 1; ++q) if ((int)(blockIdx.x + q * gridDim.x) < a.ngroups) prefetch(blockIdx.x + q * gridDim.x, q);

  // the small per-sample rows of a group (encoder output, ActNorm parameters) are fetched one group ahead into registers
  const int es = tid / K, ek = tid - es * K;                  // this thread's element of the (SL, K) encoder tile (tid < SL * K)
  float e_next = 0.f, t_next[2] = {0.f, 0.f}, l_next[2] = {0.f, 0.f};
  auto fetch_small = [&](int grp) {
    const int64_t b0 = (int64_t)grp * NS;
    const int nS = (int)min((int64_t)NS, (int64_t)a.B - b0);
    e_next = (es < nS && tid < SL * K) ? __ldg(a.e + (b0 + es) * K + ek) : 0.f;
    if (a.an_logs) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int idx = tid + q * kThreads, mm = idx / D, i = idx - mm * D;
        if (mm < nS) { t_next[q] = __ldg(a.an_t + (b0 + mm) * a.an_stride + i); l_next[q] = __ldg(a.an_logs + (b0 + mm) * a.an_stride + i); }
      }
    }
  };
  if ((int)blockIdx.x < a.ngroups) fetch_small(blockIdx.x);

  const int m = tid / a.TPS, tt = tid - m * a.TPS;
  const int g = tt / a.PG, pg = tt - g * a.PG;
  const int p0 = pg * 4, pstep = a.PG * 4;
  const float4* Es4 = reinterpret_cast<const float4*>(Es);
  const float lad = a.contextflow ? __ldg(a.logabsdet) : 0.f;
  const bool teven = (T & 1) == 0 && (reinterpret_cast<uintptr_t>(a.cnw) & 7) == 0;

  int it = 0, buf = 0, use = 0;                                // stage of this iteration, how often the stages have wrapped
  for (int grp = blockIdx.x; grp < a.ngroups; grp += gridDim.x, ++it) {
    const int64_t b0 = (int64_t)grp * NS;
    const int nS = (int)min((int64_t)NS, (int64_t)a.B - b0);
    const bool more = grp + (int)gridDim.x < a.ngroups;
    {   // the stage consumed by the previous iteration (released by its closing barrier) receives the group NSTG - 1 iterations ahead
      const int far = grp + (NSTG - 1) * (int)gridDim.x;
      if (tid == 0 && far < a.ngroups) prefetch(far, buf == 0 ? NSTG - 1 : buf - 1);
    }
    // ---- this group's encoder rows [k][slot] and ActNorm parameters: registers -> shared; then the next group's fetch goes in flight ----
    if (tid < SL * K) Es[ek * SL + es] = e_next;
    if (a.an_logs) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int idx = tid + q * kThreads, mm = idx / D, i = idx - mm * D;
        if (mm < nS) { sh[mm * DR + i] = t_next[q]; sc[mm * DR + i] = expf(-l_next[q]); lg[mm * DR + i] = l_next[q]; }
      }
    }
    if (more) fetch_small(grp + gridDim.x);
    __syncthreads();
    // ---- phase A: c = CN(e) on the packed lower triangle -> W_b tiles (conv1x1.py:33-41); two entries x four samples per thread pass:
    //      per k two weight loads and one 128-bit broadcast load feed eight FMAs (KT = compile-time K: no predicates, no index arithmetic) ----
    for (int t0 = 2 * tid; t0 < T; t0 += 2 * kThreads) {
      const int t1 = t0 + 1 < T ? t0 + 1 : t0;                 // odd T: the last pass computes its single entry twice
      const uint32_t ij0 = tab[t0], ij1 = tab[t1];
      const float bv0 = __ldg(a.cnb + t0), bv1 = __ldg(a.cnb + t1);
      const int i0 = ij0 >> 16, j0 = ij0 & 0xFFFF, i1 = ij1 >> 16, j1 = ij1 & 0xFFFF;
      const float nn0 = a.contextflow ? __ldg(a.NN + i0 * D + j0) : 0.f, nn1 = a.contextflow ? __ldg(a.NN + i1 * D + j1) : 0.f;
      float acc0[SL], acc1[SL];
#pragma unroll
      for (int s = 0; s < SL; ++s) { acc0[s] = 0.f; acc1[s] = 0.f; }
      const float* wp0 = a.cnw + t0;
      const float* wp1 = a.cnw + t1;
      if (KT > 0) {
        float w0[KT > 0 ? KT : 1], w1[KT > 0 ? KT : 1];
        if (teven) {                                             // entries t0, t0 + 1 of one weight row: one 64-bit load (k T + t0 is even)
          const float2* wq = reinterpret_cast<const float2*>(wp0);
          const int T2 = T >> 1;
#pragma unroll
          for (int k = 0; k < KT; ++k) { const float2 w = __ldg(wq + k * T2); w0[k] = w.x; w1[k] = w.y; }
        } else {
#pragma unroll
          for (int k = 0; k < KT; ++k) { w0[k] = __ldg(wp0); w1[k] = __ldg(wp1); wp0 += T; wp1 += T; }
        }
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          const float4 e0 = Es4[k];
          acc0[0] = fmaf(e0.x, w0[k], acc0[0]); acc0[1] = fmaf(e0.y, w0[k], acc0[1]); acc0[2] = fmaf(e0.z, w0[k], acc0[2]); acc0[3] = fmaf(e0.w, w0[k], acc0[3]);
          acc1[0] = fmaf(e0.x, w1[k], acc1[0]); acc1[1] = fmaf(e0.y, w1[k], acc1[1]); acc1[2] = fmaf(e0.z, w1[k], acc1[2]); acc1[3] = fmaf(e0.w, w1[k], acc1[3]);
        }
      } else {
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
          const float w0 = __ldg(wp0), w1 = __ldg(wp1); wp0 += T; wp1 += T;
          const float4 e0 = Es4[k];
          acc0[0] = fmaf(e0.x, w0, acc0[0]); acc0[1] = fmaf(e0.y, w0, acc0[1]); acc0[2] = fmaf(e0.z, w0, acc0[2]); acc0[3] = fmaf(e0.w, w0, acc0[3]);
          acc1[0] = fmaf(e0.x, w1, acc1[0]); acc1[1] = fmaf(e0.y, w1, acc1[1]); acc1[2] = fmaf(e0.z, w1, acc1[2]); acc1[3] = fmaf(e0.w, w1, acc1[3]);
        }
      }
      float* wd0 = Ws + i0 * WS + j0;
      float* wd1 = Ws + i1 * WS + j1;
#pragma unroll
      for (int s = 0; s < SL; ++s) {
        if (s < NS) {
          const float c0 = acc0[s] + bv0, c1 = acc1[s] + bv1;
          float v0 = c0, v1 = c1;
          if (i0 == j0) { dg[s * DR + i0] = c0; v0 = expf(c0); if (a.contextflow) v0 -= 1.f; }
          if (i1 == j1) { dg[s * DR + i1] = c1; v1 = expf(c1); if (a.contextflow) v1 -= 1.f; }
          wd0[(size_t)s * DR * WS] = v0 + nn0;
          wd1[(size_t)s * DR * WS] = v1 + nn1;
        }
      }
    }
    mbar_wait(bar0 + 8 * buf, use & 1);
    __syncthreads();
    // ---- per-sample ldj: one warp per sample, fixed order ----
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int ms = warp; ms < nS; ms += kThreads / 32) {
        const int64_t bb = b0 + ms;
        float part = 0.f, part_an = 0.f;
        for (int i = lane; i < D; i += 32) part += dg[ms * DR + i];
        if (a.an_logs) for (int i = lane; i < D; i += 32) part_an += lg[ms * DR + i];
        part = warp_sum(part); part_an = warp_sum(part_an);
        if (lane == 0) {
          float l = (float)HW * (lad + part);
          if (a.logp_c) l += __ldg(a.logp_c + bb) * (float)HW;
          if (a.an_logs) { l += part_an; if (a.an_logp_c) l += a.an_logp_scale * __ldg(a.an_logp_c + bb); }
          a.ldj[bb] = l;
        }
      }
    }
    // ---- phase B: z = W_b x (rows g + G r, r < 4; pixels 4 (pg + v PG) .. +3 for v < PT/4), ActNorm epilogue ----
    if (m < nS && p0 < HW) {
      const float* Wm = Ws + (size_t)m * DR * WS;
      const float* Xm = Xs + ((size_t)buf * NS + m) * DR * HW + p0;
      float acc[4][PT];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < PT; ++q) acc[r][q] = 0.f;
      const float* wr = Wm + (size_t)g * WS;
      const int rstep = a.G * WS;
#pragma unroll 2
      for (int j0 = 0; j0 < DR; j0 += 4) {
        float4 w[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) w[r] = *reinterpret_cast<const float4*>(wr + r * rstep + j0);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          float xv[PT];
#pragma unroll
          for (int v4 = 0; v4 < PT / 4; ++v4) {
            float4 xq = make_float4(0.f, 0.f, 0.f, 0.f);
            if (PT == 4 || p0 + v4 * pstep < HW) xq = *reinterpret_cast<const float4*>(Xm + (size_t)(j0 + jj) * HW + v4 * pstep);
            xv[4 * v4] = xq.x; xv[4 * v4 + 1] = xq.y; xv[4 * v4 + 2] = xq.z; xv[4 * v4 + 3] = xq.w;
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float wv = jj == 0 ? w[r].x : jj == 1 ? w[r].y : jj == 2 ? w[r].z : w[r].w;
#pragma unroll
            for (int q = 0; q < PT; ++q) acc[r][q] = fmaf(wv, xv[q], acc[r][q]);
          }
        }
      }
      const bool an = a.an_logs != nullptr;
      const int64_t b = b0 + m;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = g + a.G * r;
        if (i >= D) continue;
        if (an) {
          const float t0 = sh[m * DR + i], ee = sc[m * DR + i];
#pragma unroll
          for (int q = 0; q < PT; ++q) acc[r][q] = (acc[r][q] - t0) * ee;
        }
        float* zg = a.z + (b * D + i) * (int64_t)HW + p0;
#pragma unroll
        for (int v4 = 0; v4 < PT / 4; ++v4)
          if (PT == 4 || p0 + v4 * pstep < HW)
            stg_stream(reinterpret_cast<float4*>(zg + v4 * pstep), make_float4(acc[r][4 * v4], acc[r][4 * v4 + 1], acc[r][4 * v4 + 2], acc[r][4 * v4 + 3]));
      }
    }
    __syncthreads();
    if (++buf == NSTG) { buf = 0; ++use; }
  }
}

struct Geo { int PT, G, PG, TPS, NS, DR, WS, occ, nstg; size_t smem; };

// Geometry for (D, HW, K); false when the shape has no plan (the caller keeps the two-kernel route).
static bool plan(int B, int D, int HW, int K, Geo& o) {
  if (D < 4 || D > 128 || HW % 4 != 0 || HW < 4 || K < 1 || K > 64) return false;
  o.DR = (D + 3) / 4 * 4; o.G = o.DR / 4; o.WS = o.DR + 4;
  const int T = D * (D + 1) / 2, KP = (K + kKB - 1) / kKB * kKB;
  if (kSlots * K > kThreads || 2 * kThreads < kSlots * D) return false;   // one encoder element / two ActNorm elements per thread
  bool found = false; double best = 0;
  for (int PT = 4; PT <= 8; PT += 4) {
    const int PG = (HW + PT - 1) / PT;                          // pixel groups (PT = 8: the second piece of the last group may fall beyond HW; guarded)
    const int TPS = o.G * PG;
    if (TPS > kThreads) continue;
    const int ns_max = kThreads / TPS > kSlots ? kSlots : kThreads / TPS;
    auto bytes = [&](int ns, int nstg) {
      return ((size_t)ns * o.DR * o.WS + (size_t)nstg * ns * o.DR * HW + (size_t)KP * kSlots + (size_t)4 * ns * o.DR + (size_t)((T + 1) & ~1)) * 4 + 64;
    };
    for (int NS = ns_max; NS >= 1; --NS) {
      size_t sm = bytes(NS, 2);
      if (sm > 222 * 1024) continue;
      const int occ = sm + 1024 <= 113 * 1024 ? 2 : 1;            // registers (up to 128 x 256 threads) allow two CTAs per SM
      const size_t cap = occ == 2 ? 113 * 1024 - 1024 : 222 * 1024;
      int nstg = 2;                                               // deeper activation prefetch where it is free: HBM latency under load needs ~64 KB in flight per SM
      while (nstg < 4 && bytes(NS, nstg + 1) <= cap && (size_t)(nstg - 1) * NS * D * HW * 4 * occ < 96 * 1024) ++nstg;
      sm = bytes(NS, nstg);
      // busy phase-B threads per SM; 8 pixels per thread = fewer shared loads per FMA; a lone CTA cannot hide its own load latencies
      const double util = (double)NS * TPS / kThreads * (PT == 8 ? 1.15 : 1.0) * (occ == 1 ? 0.5 : 1.0);
      if (!found || util > best) { found = true; best = util; o.PT = PT; o.PG = PG; o.TPS = TPS; o.NS = NS > B ? B : NS; o.smem = sm; o.occ = occ; o.nstg = nstg; }
    }
  }
  return found;
}

}  // namespace c1f
}  // namespace cfpp
using namespace cfpp;

extern "C" int cfpp_conv1x1_ctx_supported(int B, int D, int HW, int K) {
  c1f::Geo g;
  return B > 0 && c1f::plan(B, D, HW, K, g) ? 1 : 0;
}

extern "C" int cfpp_conv1x1_ctx_fwd(const float* x, float* z, float* ldj, const float* e, const float* cnw_tri, const float* cnb_tri,
                                    const float* NN, const float* logabsdet, const float* logp_c, int contextflow,
                                    const float* an_t, const float* an_logs, int an_per_sample, const float* an_logp_c, float an_logp_scale,
                                    int B, int D, int HW, int K, void* stream) {
  CFPP_REQUIRE(x && z && ldj && e && cnw_tri && cnb_tri && NN && logabsdet, "conv1x1_ctx: null argument");
  CFPP_REQUIRE((an_t == nullptr) == (an_logs == nullptr), "conv1x1_ctx: an_t and an_logs must be given together");
  CFPP_REQUIRE(an_t == nullptr || an_per_sample == 1 || an_per_sample == 2, "conv1x1_ctx: the ActNorm epilogue takes per-sample parameters (an_per_sample 1 or 2)");
  CFPP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0, "conv1x1_ctx: x / z must be 16-byte aligned");
  if (B <= 0) return CFPP_OK;
  c1f::Geo g;
  if (!c1f::plan(B, D, HW, K, g)) { set_error("conv1x1_ctx: shape (D %d, HW %d, K %d) has no plan", D, HW, K); return CFPP_ERR_UNSUPPORTED; }
  c1f::Args a{x, z, ldj, e, cnw_tri, cnb_tri, NN, logabsdet, logp_c, contextflow, an_t, an_logs, an_per_sample == 2 ? 2 * D : D, an_logp_c, an_logp_scale,
              B, D, g.DR, g.G, HW, K, D * (D + 1) / 2, g.PG, g.TPS, g.NS, g.WS, (B + g.NS - 1) / g.NS, g.nstg};
  const int slots = num_sms() * g.occ;
  const int grid = a.ngroups < slots ? a.ngroups : slots;
  cudaStream_t st = (cudaStream_t)stream;
#define CFPP_C1F(PT_, KT_) do { static DeviceOnce set_; if (set_.first()) { cudaFuncSetAttribute(c1f::conv1x1_ctx_kernel<PT_, KT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
    cudaFuncSetAttribute(c1f::conv1x1_ctx_kernel<PT_, KT_>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); } \
    c1f::conv1x1_ctx_kernel<PT_, KT_><<<grid, c1f::kThreads, g.smem, st>>>(a); } while (0)
  if (g.PT == 8) { if (K == 20) CFPP_C1F(8, 20); else if (K == 8) CFPP_C1F(8, 8); else CFPP_C1F(8, 0); }
  else { if (K == 20) CFPP_C1F(4, 20); else if (K == 8) CFPP_C1F(4, 8); else CFPP_C1F(4, 0); }
#undef CFPP_C1F
  return check_launch("conv1x1_ctx_fwd");
}
