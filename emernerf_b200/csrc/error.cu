#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace emer {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace emer

extern "C" const char* emer_last_error(void) { return emer::g_err; }
extern "C" int emer_version(void) { return 1; }
