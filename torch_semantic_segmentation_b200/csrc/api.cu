// Library-wide state: version, thread-local error string, launch counter.
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_pdl{-1};

bool tss_pdl_enabled() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = getenv("TSS_PDL");
        v = (e != nullptr && e[0] == '0') ? 0 : 1;
        g_pdl.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}
extern "C" int tss_set_pdl(int enabled) {
    g_pdl.store(enabled ? 1 : 0, std::memory_order_relaxed);
    return TSS_OK;
}

void tss_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void tss_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

extern "C" int tss_version(void) { return TSS_VERSION; }
extern "C" const char* tss_last_error(void) { return g_err; }
extern "C" uint64_t tss_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
