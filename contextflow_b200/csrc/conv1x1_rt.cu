// Invertible 1x1 convolution (layers/conv1x1.py:28-57) with the fused ActNorm epilogue (layers/actnorm.py:37-60), register-tiled.
//
// Per sample z (D x HW) = W_b (D x D) x (D x HW) with W_b shared (generalist) or assembled from the raw context matrix c
// (tril(c,-1) + diag(exp(diag c)) [- I + NN]).  The layer moves 8*D*HW bytes of activations (+ the lower triangle of c) per sample
// and does D*D*HW MACs on them -- HBM bound as long as the instruction stream stays close to the FMA count.  The first kernel
// (affine.cu: one pixel column per thread, matrix rows as warp-broadcast shared loads) issued one 128-bit shared load per 4 FMAs and
// ran issue-bound at 2-7x the HBM time.  Here a thread owns a 4-row x PT-pixel tile of the output: per 4 input channels it reads 4
// matrix fragments (one 128-bit load per row) and the 4 x PT activation tile (PT/4 128-bit loads per channel) for 16*PT FMAs -- one
// shared load per ~10 FMAs -- with both operands staged in shared memory by coalesced 128-bit global loads.  Rows of a thread are
// g, g+G, g+2G, g+3G (G = D/4 row groups), so that the lanes of a warp read consecutive matrix rows (stride D+4 floats: conflict
// free) and output stores stay coalesced along pixels.
#include "common.cuh"

namespace cfpp {
namespace c1 {

struct Args {
  const float* x; float* z; float* ldj; const float* NN; const float* logabsdet;
  const float* c; const float* logp_c; int contextflow;
  const float* an_t; const float* an_logs; int an_stride; const float* an_logp_c; float an_logp_scale;
  int xq_shift;
  int B, D, DR, G, HW, PTILE, PG, TPS, NS, WS, XSTR, tiles_per_sample;   // WS = matrix row stride (DR + 4), XSTR = activation row stride
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// DC: the channel count when it is one of the compile-time cases (16, 32, 64: every index split below is a shift), 0 = run time
template <int PT, int DC>
__global__ void __launch_bounds__(256, 2) conv1x1_rt_kernel(const Args a) {
  extern __shared__ float4 c1_smem4[];
  const int D = DC ? DC : a.D, DR = DC ? DC : a.DR, HW = a.HW, WS = DR + 4, XSTR = a.XSTR;
  const int nmat = a.c ? a.NS : 1;
  float* Ws = reinterpret_cast<float*>(c1_smem4);            // [nmat][DR][WS]; rows >= D and columns >= D are zero
  float* Xs = Ws + (size_t)nmat * DR * WS;                   // [NS][DR][XSTR]; rows >= D are zero
  float* sh = Xs + (size_t)a.NS * DR * XSTR;                 // [NS][DR] ActNorm shift
  float* sc = sh + (size_t)a.NS * DR;                        // [NS][DR] ActNorm exp(-logs)
  const int64_t grp = blockIdx.x;
  const int64_t sgp = grp / a.tiles_per_sample;
  const int ptile = (int)(grp - sgp * a.tiles_per_sample);
  const int64_t b0 = sgp * a.NS;
  const int nS = (int)min((int64_t)a.NS, (int64_t)a.B - b0);
  const int p_base = ptile * a.PTILE;
  const int npx = min(a.PTILE, HW - p_base);                  // pixels of this tile
  const int tid = threadIdx.x, nthr = blockDim.x;

  // ---- activations: [sample][channel][pixel] -> shared, 128-bit coalesced when aligned (rows >= D and pixels >= npx read as zero) ----
  {
    const bool al = ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0) && (HW % 4 == 0) && (p_base % 4 == 0);
    const int XQ = XSTR >> 2;
    if (al) {
      // asynchronous copies: every thread has all its pieces in flight at once (a register-staged loop keeps one or two), which is
      // what an HBM-bound tile load needs; completion is awaited once, after the matrix pieces below have been issued too
      const int xsh = a.xq_shift;                                        // log2(XQ) when XQ is a power of two, else -1
      for (int idx = tid; idx < a.NS * DR * XQ; idx += nthr) {
        const int row = xsh >= 0 ? idx >> xsh : idx / XQ, q = xsh >= 0 ? idx & (XQ - 1) : idx - row * XQ;
        const int m = row / DR, j = row - m * DR;
        float* dst = Xs + (size_t)row * XSTR + 4 * q;
        if (m < nS && j < D && 4 * q < npx) cp_async16(dst, a.x + ((b0 + m) * D + j) * (int64_t)HW + p_base + 4 * q);
        else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      for (int idx = tid; idx < a.NS * DR * XSTR; idx += nthr) {
        const int q = idx % XSTR; const int row = idx / XSTR; const int j = row % DR, m = row / DR;
        Xs[idx] = (m < nS && j < D && q < npx) ? __ldg(a.x + ((b0 + m) * D + j) * (int64_t)HW + p_base + q) : 0.f;
      }
    }
  }
  // ---- matrices: W = NN (shared) or tril(c,-1) + diag(exp(diag c)) [- I + NN]   (conv1x1.py:36-49), row-major, row stride WS ----
  for (int idx = tid; idx < nmat * DR * (WS - D); idx += nthr) {        // zero the padding columns [D, WS) of every row
    const int row = idx / (WS - D), col = D + idx % (WS - D);
    Ws[(size_t)row * WS + col] = 0.f;
  }
  if (DR != D) for (int idx = tid; idx < nmat * (DR - D) * D; idx += nthr) {   // and the padding rows [D, DR)
    const int mm = idx / ((DR - D) * D), rem = idx % ((DR - D) * D);
    Ws[((size_t)mm * DR + D + rem / D) * WS + rem % D] = 0.f;
  }
  if (!a.c) {
    for (int idx = tid; idx < D * D; idx += nthr) { const int i = idx / D, j = idx - i * D; Ws[i * WS + j] = __ldg(a.NN + idx); }
  } else if ((D & 3) == 0) {
    // raw lower-triangle pieces of c straight into their place in shared memory (asynchronous), transformed in place afterwards
    const int D4 = D >> 2, units = nS * D * D4;
    const float* cb = a.c + b0 * (int64_t)D * D;
    for (int u = tid; u < units; u += nthr) {
      const int row = u / D4, j0 = (u - row * D4) << 2, mm = row / D, i = row - mm * D;
      if (j0 <= i) cp_async16(Ws + ((size_t)mm * DR + i) * WS + j0, cb + (int64_t)u * 4);      // entries strictly above the diagonal are never read
    }
    cp_async_wait_all();
    const float4* n4 = reinterpret_cast<const float4*>(a.NN);
    for (int u = tid; u < units; u += nthr) {                             // each thread transforms the pieces it copied itself
      const int row = u / D4, j0 = (u - row * D4) << 2, mm = row / D, i = row - mm * D;
      float4* wp = reinterpret_cast<float4*>(Ws + ((size_t)mm * DR + i) * WS + j0);
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (j0 <= i) {
        const float4 cv = *wp;
        v[0] = cv.x; v[1] = cv.y; v[2] = cv.z; v[3] = cv.w;
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int j = j0 + q; v[q] = (j < i) ? v[q] : (j == i ? expf(v[q]) : 0.f); }
      }
      if (a.contextflow) {
        const float4 nn = __ldg(n4 + (i * D4 + (j0 >> 2)));
#pragma unroll
        for (int q = 0; q < 4; ++q) if (j0 + q == i) v[q] -= 1.f;        // (no dynamic index: v stays in registers)
        v[0] += nn.x; v[1] += nn.y; v[2] += nn.z; v[3] += nn.w;
      }
      *wp = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
    for (int idx = tid; idx < nS * D * D; idx += nthr) {
      const int mm = idx / (D * D), rem = idx - mm * D * D, i = rem / D, j = rem - i * D;
      const float cij = j <= i ? a.c[(b0 + mm) * (int64_t)D * D + rem] : 0.f;
      float v = (j < i) ? cij : (j == i ? expf(cij) : 0.f);
      if (a.contextflow) v = (v - (i == j ? 1.f : 0.f)) + __ldg(a.NN + rem);
      Ws[((size_t)mm * DR + i) * WS + j] = v;
    }
  }
  if (a.an_logs) {
    for (int idx = tid; idx < a.NS * D; idx += nthr) {
      const int mm = idx / D, i = idx - mm * D;
      const int64_t bb = min(b0 + mm, (int64_t)a.B - 1);
      sh[mm * DR + i] = a.an_t[bb * a.an_stride + i];
      sc[mm * DR + i] = expf(-a.an_logs[bb * a.an_stride + i]);
    }
  }
  // ---- per-sample ldj: one warp per sample (only the CTA of the sample's first pixel tile writes it) ----
  if (ptile == 0) {
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    for (int ms = warp; ms < nS; ms += nwarps) {
      const int64_t bb = b0 + ms;
      float part = 0.f, part_an = 0.f;
      if (a.c) for (int i = lane; i < D; i += 32) part += a.c[(bb * D + i) * (int64_t)D + i];
      if (a.an_logs) for (int i = lane; i < D; i += 32) part_an += a.an_logs[bb * a.an_stride + i];
      part = warp_sum(part); part_an = warp_sum(part_an);
      if (lane == 0) {
        float l;
        if (a.c) { l = (float)HW * ((a.contextflow ? a.logabsdet[0] : 0.f) + part); if (a.logp_c) l += a.logp_c[bb] * (float)HW; }
        else l = a.logabsdet[0] * (float)HW;
        if (a.an_logs) { l += part_an; if (a.an_logp_c) l += a.an_logp_scale * a.an_logp_c[bb]; }
        a.ldj[bb] = l;
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();

  // ---- thread (sample m, row group g, pixel group pg): rows g + G r (r < 4); pixels 4 (pg + v PG) .. +3 for v < PT/4, so that for
  //      every 128-bit load / store the lanes of a warp touch consecutive 16-byte pieces ----
  const int m = tid / a.TPS, t = tid - m * a.TPS;
  if (m >= nS) return;
  const int g = t / a.PG, pg = t - g * a.PG;
  const int p0 = pg * 4, pstep = a.PG * 4;
  if (p0 >= npx) return;
  const float* Wm = Ws + (a.c ? (size_t)m * DR * WS : 0);
  const float* Xm = Xs + (size_t)m * DR * XSTR + p0;
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
        const float4 xq = *reinterpret_cast<const float4*>(Xm + (size_t)(j0 + jj) * XSTR + v4 * pstep);
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
  const bool st16 = ((reinterpret_cast<uintptr_t>(a.z) & 15) == 0) && (HW % 4 == 0) && (p_base % 4 == 0);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = g + a.G * r;
    if (i >= D) continue;
    if (an) {
      const float tt = sh[m * DR + i], e = sc[m * DR + i];
#pragma unroll
      for (int q = 0; q < PT; ++q) acc[r][q] = (acc[r][q] - tt) * e;
    }
    float* zg = a.z + (b * D + i) * (int64_t)HW + p_base + p0;
#pragma unroll
    for (int v4 = 0; v4 < PT / 4; ++v4) {
      const int pp = p0 + v4 * pstep;                                     // first pixel (within the tile) of this 4-pixel piece
      if (st16 && pp + 3 < npx) stg_stream(reinterpret_cast<float4*>(zg + v4 * pstep), make_float4(acc[r][4 * v4], acc[r][4 * v4 + 1], acc[r][4 * v4 + 2], acc[r][4 * v4 + 3]));
      else {
#pragma unroll
        for (int q = 0; q < 4; ++q) if (pp + q < npx) zg[v4 * pstep + q] = acc[r][4 * v4 + q];
      }
    }
  }
}

}  // namespace c1
}  // namespace cfpp
using namespace cfpp;

// Returns CFPP_OK after launching, or CFPP_ERR_UNSUPPORTED (nothing launched) when the tile does not fit: the caller falls back to
// the column-per-thread kernel of affine.cu.
int cfpp_conv1x1_rt_launch(const float* x, float* z, float* ldj, const float* NN, const float* logabsdet, const float* c, const float* logp_c,
                           int contextflow, const float* an_t, const float* an_logs, int an_stride, const float* an_logp_c, float an_logp_scale,
                           int B, int D, int HW, cudaStream_t st) {
  c1::Args a{x, z, ldj, NN, logabsdet, c, logp_c, contextflow, an_t, an_logs, an_stride, an_logp_c, an_logp_scale, 0, B, D};
  a.DR = (D + 3) / 4 * 4; a.G = a.DR / 4; a.HW = HW; a.WS = a.DR + 4;
  const int PT = (HW % 8 == 0 && a.G * (HW / 8) >= 32) ? 8 : 4;          // 8 pixels per thread when that still leaves a warp per sample
  int ptile = (HW + PT - 1) / PT * PT;
  const int max_ptile = (256 / a.G) * PT;                                // one CTA of 256 threads covers at most this many pixels of a sample
  if (max_ptile < PT) return CFPP_ERR_UNSUPPORTED;
  if (ptile > max_ptile) ptile = max_ptile;
  // keep the tile count even across the sample: split HW into equal tiles
  const int tiles = (HW + ptile - 1) / ptile;
  ptile = ((HW + tiles - 1) / tiles + PT - 1) / PT * PT;
  a.PTILE = ptile; a.tiles_per_sample = (HW + ptile - 1) / ptile;
  a.PG = ptile / PT; a.TPS = a.G * a.PG;
  a.XSTR = (ptile + 3) / 4 * 4;
  { const int xq = a.XSTR >> 2; a.xq_shift = -1; for (int sft = 0; sft < 16; ++sft) if ((1 << sft) == xq) a.xq_shift = sft; }
  if (a.TPS > 256) return CFPP_ERR_UNSUPPORTED;
  int ns = 256 / a.TPS; if (ns < 1) ns = 1;
  const size_t per_sample = ((size_t)(c ? a.DR * a.WS : 0) + (size_t)a.DR * a.XSTR + 2 * a.DR) * sizeof(float);
  const size_t fixed = (size_t)(c ? 0 : a.DR * a.WS) * sizeof(float);
  const size_t budget = 100 * 1024;                                       // two CTAs per SM
  while (ns > 1 && fixed + ns * per_sample > budget) --ns;
  if (fixed + ns * per_sample > 200 * 1024) return CFPP_ERR_UNSUPPORTED;
  if (ns > B) ns = B;
  a.NS = ns;
  const int threads = (ns * a.TPS + 31) / 32 * 32;
  const size_t smem = fixed + ns * per_sample;
  const int64_t blocks = (int64_t)((B + ns - 1) / ns) * a.tiles_per_sample;
  if (blocks > 0x7fffffff) return CFPP_ERR_UNSUPPORTED;
#define CFPP_C1RT(PT_, DC_) do { static DeviceOnce set_; if (set_.first()) { cudaFuncSetAttribute(c1::conv1x1_rt_kernel<PT_, DC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
    cudaFuncSetAttribute(c1::conv1x1_rt_kernel<PT_, DC_>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); } \
    c1::conv1x1_rt_kernel<PT_, DC_><<<(unsigned)blocks, threads, smem, st>>>(a); } while (0)
  const int DC = (D == 16 || D == 32 || D == 64) ? D : 0;
  if (PT == 8) { if (DC == 16) CFPP_C1RT(8, 16); else if (DC == 32) CFPP_C1RT(8, 32); else if (DC == 64) CFPP_C1RT(8, 64); else CFPP_C1RT(8, 0); }
  else { if (DC == 16) CFPP_C1RT(4, 16); else if (DC == 32) CFPP_C1RT(4, 32); else if (DC == 64) CFPP_C1RT(4, 64); else CFPP_C1RT(4, 0); }
#undef CFPP_C1RT
  return check_launch("conv1x1_fwd");
}
