// Library-wide plumbing: error messages, device checks.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";

void crct_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int crct_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
            crct_set_error("cannot query the SM count of the current device");
            sms = 0;
            return -1;
        }
    }
    return sms;
}

bool crct_pdl_enabled() {
    // opt-in: measured 0.1-0.3 ms per step SLOWER on B200 (profiles/r01_ab_pdl_s14.txt) — see common.cuh
    static const bool on = []() { const char* e = getenv("CRCT_PDL"); return e && e[0] == '1'; }();
    return on;
}

extern "C" CRCT_API const char* crct_last_error(void) { return g_err; }
extern "C" CRCT_API int crct_version(void) { return 100; }

extern "C" CRCT_API int crct_device_check(void) {
    int dev = 0, major = 0;
    CRCT_CUDA(cudaGetDevice(&dev));
    CRCT_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) CRCT_FAIL(CRCT_ERR_ARCH, "device %d has compute capability %d.x; libcrct_b200 needs sm_100 (B200)", dev, major);
    return CRCT_OK;
}
