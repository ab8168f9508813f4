// State of the SIMT emulation + the few host-side library functions the launchers call (csrc/api.cu's role).
#include <stdarg.h>

#include "cuda_runtime.h"

namespace tss_emu {
thread_local Block* block = nullptr;
thread_local unsigned slot_parity = 0;
alignas(1024) unsigned char dyn_smem[256 * 1024];
}  // namespace tss_emu
thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;

static char g_err[512];
static unsigned long long g_launches = 0;
bool tss_pdl_enabled() { return false; }
void tss_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void tss_count_launch(int n) { g_launches += n; }
extern "C" const char* tss_last_error(void) { return g_err; }
extern "C" uint64_t tss_launch_count(void) { return g_launches; }

#include "tcgen05_emu.h"
namespace tss_emu {
MBar mbars[sizeof(dyn_smem) / 8];
std::mutex mbar_mutex;
float tmem[128][512];
}  // namespace tss_emu

// The two tcgen05 GEMMs that are not converted to the emulation yet (csrc/pwconv_tc.cu): the SIMT front (pwconv_simt.cu)
// refers to them for impl 1; a host build reports an error instead of leaving the symbols unresolved.
int tss_pwconv_fwd_tc(const void*, const void*, void*, int64_t, int, int, int64_t, int64_t, const float*, const float*,
                      const void*, int64_t, int, double*, cudaStream_t) {
    tss_set_error("pwconv_fwd (impl 1) is not part of the host emulation");
    return 1;
}
int tss_pwconv_wgrad_tc(const void*, const void*, float*, int64_t, int, int, int64_t, int64_t, cudaStream_t) {
    tss_set_error("pwconv_wgrad (impl 1) is not part of the host emulation");
    return 1;
}
