// ABI bookkeeping: version, thread-local error string, launch counter.
#include <cstring>

#include "common.cuh"

namespace ou {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<int64_t> g_conv_fallbacks{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace ou

extern "C" int ou_abi_version(void) { return OU_ABI_VERSION; }

extern "C" int ou_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return OU_ERR_INVALID;
  strncpy(buf, ou::g_err, n - 1);
  buf[n - 1] = '\0';
  return OU_OK;
}

extern "C" int64_t ou_launch_count(void) { return ou::g_launches.load(); }

extern "C" int ou_act_dtype(void) { return OU_ACT_IS_BF16; }

extern "C" int64_t ou_conv_fallback_count(void) { return ou::g_conv_fallbacks.load(); }
