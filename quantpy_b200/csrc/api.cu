// Error handling, launch bookkeeping and version entry points of the C ABI (include/quantpy_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/quantpy_b200.h"
#include "common.cuh"

namespace qpb {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return QPB_OK;
    set_error("CUDA error %s (%d) in %s", cudaGetErrorString(e), (int)e, what);
    return QPB_ERR_CUDA;
}

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

}  // namespace qpb

extern "C" {

int qpb_abi_version(void) { return QPB_ABI_VERSION; }
const char* qpb_last_error(void) { return qpb::g_err; }
int64_t qpb_launch_count(void) { return qpb::g_launches.load(); }
void qpb_reset_launch_count(void) { qpb::g_launches.store(0); }

}  // extern "C"
