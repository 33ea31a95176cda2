// Host-facing end of the bootstrap interval: the calls a binding makes with HOST inputs and outputs.
//
//   qpb_bootstrap_state_interval   BootstrapStateInterval.setup() + cl_to_dist(levels) of
//                                  quantpy/tomography/interval.py:583-612 in ONE call: centre state (Bloch vector
//                                  and matrix) and confidence levels come from host memory, the quantiles go back
//                                  to host memory; everything between stays on the device
//   qpb_quantiles_host             cl_to_dist(levels) alone (interp1d over linspace(0, 1, N), interval.py:611-612)
//                                  on a device-resident sorted array
//
// Both stage their few KB through a page-locked buffer owned by the library (one H2D and one D2H copy per call)
// and end with a stream synchronisation: when they return, the host output is valid.
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "plan.h"

namespace qpb {

// Grow-only page-locked staging buffer per (device, stream); same retirement rule as scratch().  A call that
// returns WITHOUT synchronising (qpb_bootstrap_state_interval with no levels) leaves an upload from this buffer in
// flight, so every user waits for `uploaded` (the event recorded after the last H2D copy) before it writes to it.
struct PinnedStage {
    void* p = nullptr;
    cudaEvent_t uploaded = nullptr;
};
static PinnedStage pinned_stage(cudaStream_t st, size_t bytes) {
    struct Buf {
        void* p = nullptr;
        size_t n = 0;
        cudaEvent_t uploaded = nullptr;
    };
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, Buf> pool;
    static std::vector<void*> retired;
    int dev = 0;
    if (check_cuda(cudaGetDevice(&dev), "pinned_stage cudaGetDevice") != QPB_OK) return {};
    std::lock_guard<std::mutex> lock(mu);
    Buf& b = pool[{dev, st}];
    if (!b.uploaded &&
        check_cuda(cudaEventCreateWithFlags(&b.uploaded, cudaEventDisableTiming), "cudaEventCreate") != QPB_OK)
        return {};
    if (b.n < bytes) {
        size_t want = bytes < 65536 ? 65536 : bytes;
        if (want < 2 * b.n) want = 2 * b.n;
        void* p = nullptr;
        if (check_cuda(cudaHostAlloc(&p, want, cudaHostAllocDefault), "cudaHostAlloc") != QPB_OK) return {};
        if (b.p) retired.push_back(b.p);  // a copy from the old buffer may still be in flight: it is never freed
        b.p = p;
        b.n = want;
    } else if (check_cuda(cudaEventSynchronize(b.uploaded), "pinned_stage wait") != QPB_OK) {
        return {};  // (an event that was never recorded counts as complete)
    }
    return {b.p, b.uploaded};
}

// out[i] = y[lo] + (y[lo + 1] - y[lo]) * (pos - lo), pos = level * (n - 1), lo = min(floor(pos), n - 2): the linear
// interpolant of (linspace(0, 1, n), y) in the very operations of the host formula (no contraction), so a level
// evaluated here and in NumPy on the same sorted array gives the same bits
__global__ void k_quantiles(long n, const double* __restrict__ y, int m, const double* __restrict__ levels,
                            double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    if (n == 1) {
        out[i] = y[0];
        return;
    }
    const double pos = __dmul_rn(levels[i], (double)(n - 1));
    long lo = (long)floor(pos);
    if (lo > n - 2) lo = n - 2;
    if (lo < 0) lo = 0;
    const double frac = __dsub_rn(pos, (double)lo);
    const double y_lo = y[lo], y_hi = y[lo + 1];
    out[i] = __dadd_rn(y_lo, __dmul_rn(__dsub_rn(y_hi, y_lo), frac));
}

static int check_levels(int n_levels, const double* levels_host) {
    for (int i = 0; i < n_levels; ++i) {
        QPB_REQUIRE(!(levels_host[i] < 0.0), "A value in x_new is below the interpolation range.");
        QPB_REQUIRE(!(levels_host[i] > 1.0), "A value in x_new is above the interpolation range.");
    }
    return QPB_OK;
}

}  // namespace qpb

using namespace qpb;

extern "C" {

int qpb_quantiles_host(long long n, const double* sorted_dev, int n_levels, const double* levels_host,
                       double* out_host, void* stream) {
    QPB_REQUIRE(n >= 1, "empty sample");
    QPB_REQUIRE(n_levels >= 0, "negative number of levels");
    if (n_levels == 0) return QPB_OK;
    QPB_REQUIRE(sorted_dev && levels_host && out_host, "NULL buffer");
    int rc = check_levels(n_levels, levels_host);
    if (rc != QPB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t m = (size_t)n_levels;
    const PinnedStage stage = pinned_stage(st, sizeof(double) * 2 * m);
    double* pin = static_cast<double*>(stage.p);
    double* dev = static_cast<double*>(scratch(st, 20, sizeof(double) * 2 * m));
    if (!pin || !dev) return QPB_ERR_NOMEM;
    memcpy(pin, levels_host, sizeof(double) * m);
    QPB_CUDA(cudaMemcpyAsync(dev, pin, sizeof(double) * m, cudaMemcpyHostToDevice, st));
    QPB_CUDA(cudaEventRecord(stage.uploaded, st));
    k_quantiles<<<(n_levels + 255) / 256, 256, 0, st>>>((long)n, sorted_dev, n_levels, dev, dev + m);
    QPB_LAUNCHED("k_quantiles");
    QPB_CUDA(cudaMemcpyAsync(pin + m, dev + m, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    QPB_CUDA(cudaStreamSynchronize(st));
    memcpy(out_host, pin + m, sizeof(double) * m);
    return QPB_OK;
}

int qpb_bootstrap_state_interval(const qpb_state_plan* plan, int B, int P, int O, const double* M_dev,
                                 const double* bloch_host, const double* ref_host, const int32_t* n_shots_host,
                                 uint64_t seed, uint64_t offset, int method, int physical, int init, int max_iter,
                                 double tol, int dist_kind, int n_levels, const double* levels_host,
                                 double* quantiles_host, double* dist_sorted, int32_t* iters_out, void* stream) {
    QPB_REQUIRE(plan != nullptr, "plan is NULL");
    QPB_REQUIRE(B >= 1, "the interval needs at least one resample");
    QPB_REQUIRE(P * O == plan->K, "P*O=%d does not match the plan's K=%d", P * O, plan->K);
    QPB_REQUIRE(M_dev && bloch_host && ref_host && n_shots_host && dist_sorted, "NULL buffer");
    QPB_REQUIRE(n_levels >= 0 && (n_levels == 0 || (levels_host && quantiles_host)), "bad confidence levels");
    int rc = check_levels(n_levels, levels_host);
    if (rc != QPB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t D = (size_t)plan->D, K = (size_t)plan->K, m = (size_t)n_levels;
    const size_t n_in = D + 2 * D + m;  // Bloch vector, complex d x d centre state, levels
    const PinnedStage stage = pinned_stage(st, sizeof(double) * (n_in + m));
    double* pin = static_cast<double*>(stage.p);
    double* dev = static_cast<double*>(scratch(st, 20, sizeof(double) * (n_in + m + K)));
    // per-sample buffers the caller does not see: unsorted distances, counts, the two state buffers
    double* dist = static_cast<double*>(scratch(st, 21, sizeof(double) * (size_t)B));
    int32_t* counts = static_cast<int32_t*>(scratch(st, 22, sizeof(int32_t) * K * (size_t)B));
    void* work = scratch(st, 23, qpb_bootstrap_state_workspace(plan, B, P, O));
    if (!pin || !dev || !dist || !counts || !work) return QPB_ERR_NOMEM;
    memcpy(pin, bloch_host, sizeof(double) * D);
    memcpy(pin + D, ref_host, sizeof(double) * 2 * D);
    if (m) memcpy(pin + 3 * D, levels_host, sizeof(double) * m);
    QPB_CUDA(cudaMemcpyAsync(dev, pin, sizeof(double) * n_in, cudaMemcpyHostToDevice, st));
    QPB_CUDA(cudaEventRecord(stage.uploaded, st));
    double* bloch_dev = dev;
    double* ref_dev = dev + D;
    double* levels_dev = dev + 3 * D;
    double* q_dev = levels_dev + m;
    double* probs_dev = q_dev + m;
    // state.py:109-110: p = clip(2^n M.r, 0, 1)
    rc = qpb_povm_probs(plan->K, plan->D, 1, M_dev, bloch_dev, (double)plan->d, 1, probs_dev, stream);
    if (rc != QPB_OK) return rc;
    rc = qpb_bootstrap_state(plan, B, P, O, probs_dev, n_shots_host, seed, offset, method, physical, init, max_iter,
                             tol, ref_dev, dist_kind, dist, nullptr, counts, iters_out, work, stream);
    if (rc != QPB_OK) return rc;
    rc = qpb_sort_f64(B, dist, dist_sorted, stream);  // interval.py:610
    if (rc != QPB_OK) return rc;
    // no levels, no host output: the call returns with its work queued (the multi-GPU interval goes on to queue the
    // all-gather and the merge behind it, and synchronises once, in qpb_quantiles_host)
    if (!m) return QPB_OK;
    k_quantiles<<<(n_levels + 255) / 256, 256, 0, st>>>((long)B, dist_sorted, n_levels, levels_dev, q_dev);
    QPB_LAUNCHED("k_quantiles");
    QPB_CUDA(cudaMemcpyAsync(pin + n_in, q_dev, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    QPB_CUDA(cudaStreamSynchronize(st));
    memcpy(quantiles_host, pin + n_in, sizeof(double) * m);
    return QPB_OK;
}

}  // extern "C"
