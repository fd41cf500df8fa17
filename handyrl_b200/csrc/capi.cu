// Error channel and version of the C ABI (include/hrl_b200.h).
#include "common.cuh"
#include <stdarg.h>

namespace hrl {
static thread_local char g_last_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}
}  // namespace hrl

extern "C" const char *hrl_last_error(void) { return hrl::g_last_error; }
extern "C" int32_t hrl_abi_version(void) { return HRL_ABI_VERSION; }
