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

// Debugging aid for CUDA-graph capture problems (library.py, TSS_CAPTURE_CHECK=1).
extern "C" int64_t tss_capture_status(int64_t stream_handle) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaError_t e = cudaStreamIsCapturing((cudaStream_t)(uintptr_t)stream_handle, &st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return -(int64_t)e;
    }
    return st == cudaStreamCaptureStatusNone ? 0 : st == cudaStreamCaptureStatusActive ? 1 : 2;
}

#ifdef TSS_TRACE
// libtss_b200_trace.so only (tools/trace_kernels.py): device buffer of TSS_TRACE_MAX_CTAS x TSS_TRACE_SLOTS uint64, or NULL.
static std::atomic<unsigned long long*> g_trace_buffer{nullptr};
unsigned long long* tss_trace_buffer_host() { return g_trace_buffer.load(std::memory_order_relaxed); }
extern "C" int tss_trace_set(void* buffer) {
    g_trace_buffer.store((unsigned long long*)buffer, std::memory_order_relaxed);
    return TSS_OK;
}
#endif
