// b2s_core.cu — error state, version, launch counter.
#include "b2s_common.cuh"
#include <cstdarg>

namespace b2s {

static thread_local char g_err[512] = "no error";
std::atomic<int64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace b2s

extern "C" int b2s_version(void) { return B2S_VERSION; }
extern "C" const char* b2s_last_error_string(void) { return b2s::g_err; }
extern "C" int64_t b2s_launch_count(void) { return b2s::g_launch_count.load(); }
