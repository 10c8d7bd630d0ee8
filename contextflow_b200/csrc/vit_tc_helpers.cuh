// Shared device helpers of the tensor-core ViT kernels (mbarrier / bulk-TMA / tcgen05 wrappers, fp16 hi-lo operand rows, TMEM row loads).
// The same code as in vit_tc.cu (kept verbatim there); see that file and conv_cond_tc.cu for the conventions they implement.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace cfpp {
namespace vtx {

constexpr int kW = 64;                              // feature width of one operand panel / accumulator chunk
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {       // K-major SWIZZLE_128B, 8-row group stride 1024 bytes
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int n) {           // kind::f16: fp16 operands, fp32 accumulate, K-major A and B, M = 128
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                 "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
// packed fp16 pair {lo half = a, hi half = b}, round to nearest, saturating to the finite range (one F2FP instruction)
__device__ __forceinline__ uint32_t pack_f16x2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// (hi, lo') fp16 pair of two values: hi = v truncated to 11 significant bits (exact in fp16 over its normal range), lo' = rn((v - hi) * 2^11);
// 4 instructions per element
__device__ __forceinline__ void f16_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const float ha = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), hb = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
  hi = pack_f16x2_sat(ha, hb);
  lo = pack_f16x2_sat((a - ha) * kLoScale, (b - hb) * kLoScale);
}
// the thread's activation row (64 channels, zero beyond the live width) -> operand row `row` of the hi / lo regions
__device__ __forceinline__ void store_operand_row(uint8_t* a_hi, uint8_t* a_lo, int row, const float (&v)[kW]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) f16_split2(v[8 * c + 2 * q], v[8 * c + 2 * q + 1], h[q], l[q]);
    const uint32_t off = (uint32_t)row * 128u + (uint32_t)(((c ^ row) & 7) << 4);
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}
// the thread's accumulator row of one 64-wide GEMM output: main block [c0, c0+64) + 2^-11 * cross block [c0+64, c0+128)
__device__ __forceinline__ void load_acc_row(uint32_t taddr, float (&o)[kW]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float v[16], u[16];
    tmem_ld16(taddr + 16 * q, v);
    tmem_ld16(taddr + kW + 16 * q, u);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) o[16 * q + i] = fmaf(u[i], kLoInv, v[i]);
  }
}
// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7) with the SFU exponential / reciprocal: ~14 instructions instead of ~40 for erff
__device__ __forceinline__ float erf_as(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
  const float r = 1.0f - poly * __expf(-ax * ax);
  return copysignf(r, x);
}


}  // namespace vtx
}  // namespace cfpp
