// Shared device/host helpers for libcfpp (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/cfpp.h"

namespace cfpp {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return CFPP_ERR_CUDA; }
  return CFPP_OK;
}

#define CFPP_REQUIRE(cond, ...) do { if (!(cond)) { ::cfpp::set_error(__VA_ARGS__); return CFPP_ERR_ARG; } } while (0)

// Function attributes (dynamic shared-memory limit, carve-out) are per DEVICE: `static DeviceOnce once; if (once.first()) { cudaFuncSetAttribute... }`
// runs its body once for every device a kernel is launched on (one process may drive several GPUs: contextflow_b200/multigpu.py).
struct DeviceOnce {
  std::atomic<unsigned long long> seen{0};
  bool first() {
    int dev = 0; cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return (seen.fetch_or(bit, std::memory_order_relaxed) & bit) == 0;
  }
};
// per-device high-water mark for attributes that grow with the request (`if (hw.raise(bytes)) cudaFuncSetAttribute(..., bytes)`)
struct DeviceHighWater {
  std::atomic<long long> v[64];
  DeviceHighWater() { for (auto& x : v) x.store(0); }
  bool raise(long long want) {
    int dev = 0; cudaGetDevice(&dev);
    std::atomic<long long>& a = v[dev & 63];
    if (want <= a.load(std::memory_order_relaxed)) return false;
    a.store(want, std::memory_order_relaxed);
    return true;
  }
};

inline int num_sms() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over a group of G consecutive threads (G multiple of 32, G <= blockDim). `red` needs blockDim/32 floats.
// Every thread of the group receives the total.  All threads of the block must call (uses __syncthreads).
template <int MAXW = 32>
__device__ __forceinline__ float group_sum(float v, int G, float* red) {
  v = warp_sum(v);
  if (G == 32) return v;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  const int wpg = G >> 5, g0 = (w / wpg) * wpg;
  float s = 0.f;
  for (int i = 0; i < wpg; ++i) s += red[g0 + i];
  __syncthreads();
  return s;
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {   // streaming 128-bit load, no L1 allocation
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float softplus_f(float x) {   // F.softplus(beta=1, threshold=20)
  return x > 20.f ? x : log1pf(expf(x));
}

constexpr float kHalfLog2Pi = 0.91893853320467274178f;

}  // namespace cfpp
