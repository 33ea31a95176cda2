// Error handling, launch bookkeeping and version entry points of the C ABI (include/quantpy_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <utility>

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

void* scratch(cudaStream_t st, int slot, size_t bytes) {
    struct Buf {
        void* p = nullptr;
        size_t n = 0;
    };
    static std::mutex mu;
    static std::map<std::pair<cudaStream_t, int>, Buf> pool;
    std::lock_guard<std::mutex> lock(mu);
    Buf& b = pool[std::make_pair(st, slot)];
    if (b.n < bytes) {
        if (b.p) {
            cudaStreamSynchronize(st);  // earlier kernels on this stream may still read the old buffer
            cudaFree(b.p);
            b.p = nullptr;
            b.n = 0;
        }
        size_t want = bytes < 4096 ? 4096 : bytes;
        if (check_cuda(cudaMalloc(&b.p, want), "scratch cudaMalloc") != QPB_OK) return nullptr;
        b.n = want;
    }
    return b.p;
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

namespace qpb {
// Roofline probe: 8 independent FMA chains per thread, nothing else.  bench.py times it with CUDA
// events to get the FP64 FMA peak of the device it is running on (MEASURED_PEAKS.json has no FP64 line).
__global__ void k_fp64_fma_probe(long iters, double seed, double* __restrict__ sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (long i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 12345.6789) sink[0] = s;  // never true; keeps the chains alive
}
}  // namespace qpb

extern "C" {

int qpb_fp64_fma_probe(int64_t iters_per_thread, double* sink, double* flops_out_host, void* stream) {
    QPB_REQUIRE(iters_per_thread > 0 && sink != nullptr, "bad probe arguments");
    const int blocks = qpb::num_sms() * 8, threads = 256;
    qpb::k_fp64_fma_probe<<<blocks, threads, 0, (cudaStream_t)stream>>>((long)iters_per_thread, 1.0, sink);
    QPB_LAUNCHED("k_fp64_fma_probe");
    if (flops_out_host) *flops_out_host = 2.0 * 8.0 * (double)iters_per_thread * blocks * threads;
    return QPB_OK;
}

int qpb_abi_version(void) { return QPB_ABI_VERSION; }
const char* qpb_last_error(void) { return qpb::g_err; }
int64_t qpb_launch_count(void) { return qpb::g_launches.load(); }
void qpb_reset_launch_count(void) { qpb::g_launches.store(0); }

}  // extern "C"
