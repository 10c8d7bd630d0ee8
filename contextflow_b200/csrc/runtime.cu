// libcfpp runtime plumbing: error string, version, launch counter.
#include <stdarg.h>
#include "common.cuh"

namespace cfpp {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
}  // namespace cfpp

extern "C" int cfpp_version(void) { return 100; }
extern "C" const char* cfpp_last_error(void) { return cfpp::g_err; }
extern "C" int64_t cfpp_launch_count(void) { return cfpp::g_launches.load(); }

// Asynchronous copy between two devices of this process (or within one) on `stream`, which belongs to the CURRENT device: the device
// that runs a batch slice pulls its input from / pushes its result to the caller's device on its own stream, so no stream of the other
// device takes part (contextflow_b200/multigpu.py).  Peer access is switched on once per ordered device pair when the hardware offers it
// (NVLink / NVSwitch); without it the runtime stages the copy.
extern "C" int cfpp_copy_peer_async(void* dst, int dst_device, const void* src, int src_device, int64_t bytes, void* stream) {
  using namespace cfpp;
  if (bytes <= 0) return CFPP_OK;
  int cur = 0; cudaGetDevice(&cur);
  static std::atomic<unsigned long long> enabled[64];
  for (int other : {dst_device, src_device}) {
    if (other == cur || other < 0 || other >= 64) continue;
    const unsigned long long bit = 1ull << other;
    if ((enabled[cur & 63].fetch_or(bit) & bit) == 0) {
      int can = 0; cudaDeviceCanAccessPeer(&can, cur, other);
      if (can) { cudaError_t e = cudaDeviceEnablePeerAccess(other, 0); if (e != cudaSuccess) (void)cudaGetLastError(); }   // already enabled (torch) is fine
    }
  }
  cudaError_t e = cudaMemcpyPeerAsync(dst, dst_device, src, src_device, (size_t)bytes, (cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("copy_peer_async: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return CFPP_ERR_CUDA; }
  return CFPP_OK;
}
