// Library-level state of libbdetr: error text, compute mode, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "common.cuh"

namespace bdetr {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_mode{BDETR_MODE_FP32};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int current_mode() { return g_mode.load(std::memory_order_relaxed); }
}  // namespace bdetr

extern "C" __attribute__((visibility("default"))) int bdetr_version(void) { return 100; }
extern "C" __attribute__((visibility("default"))) const char *bdetr_last_error(void) { return bdetr::g_err; }
extern "C" __attribute__((visibility("default"))) int bdetr_set_mode(int mode)
{
    if (mode != BDETR_MODE_FP32 && mode != BDETR_MODE_TF32) {
        bdetr::set_error("bdetr_set_mode: unknown mode %d", mode);
        return BDETR_E_UNSUPPORTED;
    }
    bdetr::g_mode.store(mode);
    return BDETR_OK;
}
extern "C" __attribute__((visibility("default"))) int bdetr_get_mode(void) { return bdetr::g_mode.load(); }
extern "C" __attribute__((visibility("default"))) long long bdetr_launch_count(void) { return bdetr::g_launches.load(); }
extern "C" __attribute__((visibility("default"))) void bdetr_reset_launch_count(void) { bdetr::g_launches.store(0); }
