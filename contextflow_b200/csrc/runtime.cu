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
