// Error handling, launch bookkeeping and version entry points of the C ABI (include/quantpy_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

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

// Grow-only scratch, one buffer per (device, stream, slot).  A buffer that has to grow is not freed (a kernel
// launched earlier on the stream, or a host thread that was handed the pointer a moment ago, may still use it):
// it is retired to a per-device list that lives as long as the process.  Growth is geometric, so the retired
// bytes are bounded by the largest request.
void* scratch(cudaStream_t st, int slot, size_t bytes, bool* fresh) {
    struct Buf {
        void* p = nullptr;
        size_t n = 0;
    };
    struct Key {
        int dev;
        cudaStream_t st;
        int slot;
        bool operator<(const Key& o) const {
            if (dev != o.dev) return dev < o.dev;
            if (st != o.st) return st < o.st;
            return slot < o.slot;
        }
    };
    static std::mutex mu;
    static std::map<Key, Buf> pool;
    static std::vector<void*> retired;
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "scratch cudaGetDevice") != QPB_OK) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    Buf& b = pool[Key{dev, st, slot}];
    if (fresh) *fresh = false;
    if (b.n < bytes) {
        if (fresh) *fresh = true;
        size_t want = bytes < 4096 ? 4096 : bytes;
        if (want < 2 * b.n) want = 2 * b.n;
        void* p = nullptr;
        if (check_cuda(cudaMalloc(&p, want), "scratch cudaMalloc") != QPB_OK) return nullptr;
        if (b.p) retired.push_back(b.p);
        b.p = p;
        b.n = want;
    }
    return b.p;
}

int num_sms() {
    static std::mutex mu;
    static std::map<int, int> per_device;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    std::lock_guard<std::mutex> lock(mu);
    auto it = per_device.find(dev);
    if (it != per_device.end()) return it->second;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    per_device[dev] = sms;
    return sms;
}

// Kernel-selection toggles (tests and profiling only).  Read from the environment ONCE, when the library is
// loaded; qpb_set_option changes them afterwards.  Nothing on the launch path calls getenv.
static std::atomic<int> g_options[QPB_OPT_COUNT_];
static const char* const kOptionEnv[QPB_OPT_COUNT_] = {
    "QPB_NO_TAIL_MERGE", "QPB_NO_HS_FUSION", "QPB_NO_PAULI_KERNEL", "QPB_NO_CONST_KERNEL", "QPB_NO_AXIS_KERNEL",
    "QPB_NO_DMMA_GEMM",  "QPB_NO_ROW_JACOBI", "QPB_NO_PACKED_JACOBI", "QPB_NO_LIN_SMALL",  "QPB_SAMPLER",
    "QPB_MLE_BLOCKS_PER_SM", "QPB_MLE_LANES", "QPB_NO_TILED_MLE",
    "QPB_MLE_PARK_AGE", "QPB_MLE_PARK_LIVE", "QPB_MLE_W_WARPS", "QPB_MLE_PARK_PLATEAU", "QPB_NO_TMA_GEMM",
    "QPB_MLE_TAIL_POLL", "QPB_MLE_TAIL_AGE", "QPB_MLE_ADOPT", "QPB_MLE_MERGE", "QPB_NO_MLE_ORDER",
    "QPB_MLE_PARK_AGE_LO", "QPB_MLE_PARK_AGE_PCT", "QPB_MLE_PARK_AGE_END", "QPB_MLE_PARK_AGE_PCT2",
    "QPB_MLE_REFILL_MIN", "QPB_NO_WARM_JACOBI", "QPB_MLE_SINGLE_WARPS",
    "QPB_SAMPLER_NO_PREFILTER", "QPB_SAMPLER_EXACT_EVERY", "QPB_SAMPLER_THREADS", "QPB_NO_SAMPLE_SORT",
    "QPB_SAMPLER_LANES"};
static const bool g_options_loaded = [] {
    for (int i = 0; i < QPB_OPT_COUNT_; ++i) {
        const char* e = kOptionEnv[i] ? getenv(kOptionEnv[i]) : nullptr;
        int v = 0;
        if (e && *e) {
            if (i == QPB_OPT_SAMPLER) v = !strcmp(e, "alias") ? 1 : (!strcmp(e, "binomial") ? 2 : 0);
            else if (i >= QPB_OPT_MLE_BLOCKS_PER_SM && i != QPB_OPT_NO_TILED_MLE && i != QPB_OPT_NO_TMA_GEMM &&
                     i != QPB_OPT_NO_MLE_ORDER && i != QPB_OPT_NO_WARM_JACOBI) v = atoi(e);
            else v = 1;
        }
        g_options[i].store(v);
    }
    return true;
}();
static std::atomic<long long*> g_trace_ptr{nullptr};
static std::atomic<size_t> g_trace_bytes{0};
long long* debug_trace_buffer(size_t bytes) {
    long long* p = g_trace_ptr.load(std::memory_order_relaxed);
    return (p && g_trace_bytes.load(std::memory_order_relaxed) >= bytes) ? p : nullptr;
}
int option(int which) { return (which >= 0 && which < QPB_OPT_COUNT_) ? g_options[which].load(std::memory_order_relaxed) : 0; }

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
// Shared-memory roofline probe: every warp streams conflict-free 64-bit loads (two 128-byte wavefronts per
// warp instruction) from a 16 KB region; bench.py times it to get the device's wavefront rate.
__global__ void __launch_bounds__(1024) k_smem_probe(long iters, double* __restrict__ sink) {
    __shared__ double buf[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = (double)i;
    __syncthreads();
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int idx = threadIdx.x;
    for (long i = 0; i < iters; ++i) {
        a0 += buf[idx & 2047];
        a1 += buf[(idx + 512) & 2047];
        a2 += buf[(idx + 1024) & 2047];
        a3 += buf[(idx + 1536) & 2047];
        idx += 32;
    }
    const double s = (a0 + a1) + (a2 + a3);
    if (s == 12345.6789) sink[0] = s;
}
}  // namespace qpb

extern "C" {

int qpb_smem_probe(int64_t iters_per_thread, double* sink, double* wavefronts_out_host, void* stream) {
    QPB_REQUIRE(iters_per_thread > 0 && sink != nullptr, "bad probe arguments");
    const int blocks = qpb::num_sms() * 2, threads = 1024;
    qpb::k_smem_probe<<<blocks, threads, 0, (cudaStream_t)stream>>>((long)iters_per_thread, sink);
    QPB_LAUNCHED("k_smem_probe");
    // 4 LDS.64 per thread-iteration = 4 warp instructions per warp-iteration, 2 wavefronts each
    if (wavefronts_out_host) *wavefronts_out_host = 8.0 * (double)iters_per_thread * blocks * (threads / 32);
    return QPB_OK;
}

int qpb_fp64_fma_probe(int64_t iters_per_thread, double* sink, double* flops_out_host, void* stream) {
    QPB_REQUIRE(iters_per_thread > 0 && sink != nullptr, "bad probe arguments");
    const int blocks = qpb::num_sms() * 8, threads = 256;
    qpb::k_fp64_fma_probe<<<blocks, threads, 0, (cudaStream_t)stream>>>((long)iters_per_thread, 1.0, sink);
    QPB_LAUNCHED("k_fp64_fma_probe");
    if (flops_out_host) *flops_out_host = 2.0 * 8.0 * (double)iters_per_thread * blocks * threads;
    return QPB_OK;
}

int qpb_abi_version(void) { return QPB_ABI_VERSION; }
int qpb_set_option(int which, int value) {
    QPB_REQUIRE(which >= 0 && which < QPB_OPT_COUNT_, "unknown option %d", which);
    qpb::g_options[which].store(value);
    return QPB_OK;
}
int qpb_get_option(int which) { return qpb::option(which); }
const char* qpb_last_error(void) { return qpb::g_err; }
int64_t qpb_launch_count(void) { return qpb::g_launches.load(); }
void qpb_reset_launch_count(void) { qpb::g_launches.store(0); }

}  // extern "C"

extern "C" int qpb_debug_set_trace(void* device_buffer, size_t bytes) {
    qpb::g_trace_ptr.store(static_cast<long long*>(device_buffer));
    qpb::g_trace_bytes.store(device_buffer ? bytes : 0);
    return QPB_OK;
}
