// Library-wide state of the C-ABI: last-error string, launch counter, version.
#include "common.cuh"
#include <stdarg.h>

namespace sdc {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace sdc

extern "C" int sdc_version(void) { return 100; }
extern "C" const char* sdc_last_error(void) { return sdc::g_err; }
extern "C" int64_t sdc_launch_count(void) { return sdc::g_launches.load(); }
/* launches replayed from a captured CUDA graph are counted by the host layer (one call per replay) */
extern "C" void sdc_count_launches(int64_t n) { sdc::g_launches.fetch_add(n); }
