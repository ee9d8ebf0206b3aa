// Error reporting + ABI version for libsparsebev_b200.so.
#include "common.cuh"
#include <stdarg.h>

namespace sbev {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace sbev

extern "C" int sbev_abi_version(void) { return 1; }
extern "C" const char* sbev_last_error(void) { return sbev::g_err; }
