// Host-side orchestration of the state path: qpb_mle_rrr dispatch and the fused bootstrap call
// (sampler -> linear inversion / projection -> R.rho.R -> distance) of
// quantpy/tomography/interval.py:598-609.  All stages are enqueued on one stream with no host
// synchronisation; intermediate counts and states stay in HBM buffers owned by the caller.
#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "plan.h"

using namespace qpb;

extern "C" {

int qpb_mle_rrr(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                double tol, double* rho, int32_t* iters, void* stream) {
    QPB_REQUIRE(plan != nullptr, "plan is NULL");
    QPB_REQUIRE(B >= 0 && max_iter >= 0, "bad arguments B=%d max_iter=%d", B, max_iter);
    QPB_REQUIRE(tol >= 0.0, "tol must be non-negative");
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(counts && rho, "NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_mle_small(plan, B, counts, rho0, max_iter, tol, rho, iters, st);
    if (rc == QPB_ERR_UNSUPPORTED) rc = launch_mle_axis(plan, B, counts, rho0, max_iter, tol, rho, iters, st);
    if (rc == QPB_ERR_UNSUPPORTED && max_iter > 0) rc = launch_mle_tiled(plan, B, counts, rho0, max_iter, tol, rho, iters, st);
    if (rc == QPB_ERR_UNSUPPORTED) rc = launch_mle_generic(plan, B, counts, rho0, max_iter, tol, rho, iters, st);
    return rc;
}

int qpb_lin_project_ordered(const qpb_state_plan* plan, int B, const int32_t* counts, double* rho, int32_t* order,
                            void* stream) {
    QPB_REQUIRE(plan != nullptr, "plan is NULL");
    QPB_REQUIRE(B >= 0, "negative batch");
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(counts && rho && order, "NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const int* dev_order = nullptr;
    int rc = launch_lin_project_small(plan, B, counts, 1, rho, st, &dev_order);
    if (rc == QPB_ERR_UNSUPPORTED) rc = qpb_lin_project(plan, B, counts, 1, rho, stream);
    if (rc != QPB_OK) return rc;
    if (dev_order) {
        QPB_CUDA(cudaMemcpyAsync(order, dev_order, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToDevice, st));
        return QPB_OK;
    }
    return qpb_identity_order(B, order, stream);
}

int qpb_mle_rrr_ordered(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0,
                        const int32_t* order, int max_iter, double tol, double* rho, int32_t* iters, void* stream) {
    QPB_REQUIRE(plan != nullptr, "plan is NULL");
    QPB_REQUIRE(B >= 0 && max_iter >= 0, "bad arguments B=%d max_iter=%d", B, max_iter);
    QPB_REQUIRE(tol >= 0.0, "tol must be non-negative");
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(counts && rho, "NULL buffer");
    if (order && max_iter > 0 && mle_variant(plan) == QPB_MLE_PAULI2)
        return launch_mle_small(plan, B, counts, rho0, max_iter, tol, rho, iters, (cudaStream_t)stream, nullptr, nullptr,
                                nullptr, 1, order);
    return qpb_mle_rrr(plan, B, counts, rho0, max_iter, tol, rho, iters, stream);  // the other kernels take no hint
}

int qpb_mle_variant(const qpb_state_plan* plan) {
    QPB_REQUIRE(plan != nullptr, "plan is NULL");
    return mle_variant(plan);
}

size_t qpb_bootstrap_state_workspace(const qpb_state_plan* plan, int B, int P, int O) {
    if (!plan || B <= 0) return 0;
    (void)P;
    (void)O;
    // two state buffers: start states and reconstructed states
    return 2 * sizeof(double) * 2 * (size_t)plan->D * (size_t)B + 256;
}

int qpb_bootstrap_state(const qpb_state_plan* plan, int B, int P, int O, const double* probs,
                        const int32_t* n_shots_host, uint64_t seed, uint64_t offset, int method, int physical,
                        int init, int max_iter, double tol, const double* ref, int dist_kind, double* dist,
                        double* rho_out, int32_t* counts_out, int32_t* iters_out, void* work, void* stream) {
    QPB_REQUIRE(plan != nullptr, "plan is NULL");
    QPB_REQUIRE(P * O == plan->K, "P*O=%d does not match the plan's K=%d", P * O, plan->K);
    QPB_REQUIRE(method == QPB_METHOD_LIN || method == QPB_METHOD_MLE, "unknown method %d", method);
    QPB_REQUIRE(init == QPB_INIT_LIN || init == QPB_INIT_MIXED, "unknown init %d", init);
    QPB_REQUIRE(B >= 0, "negative batch");
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(probs && n_shots_host && ref && dist, "NULL buffer");
    QPB_REQUIRE(counts_out != nullptr, "counts_out is required (it is the sampler's output buffer)");
    QPB_REQUIRE(work != nullptr, "workspace is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t stride = 2 * (size_t)plan->D * (size_t)B;
    double* buf0 = reinterpret_cast<double*>(((uintptr_t)work + 255) & ~(uintptr_t)255);
    double* buf1 = buf0 + stride;
    double* final_rho = rho_out ? rho_out : buf1;

    int rc = qpb_multinomial(B, P, O, probs, 0, n_shots_host, seed, offset, counts_out, stream);
    if (rc != QPB_OK) return rc;
    if (method == QPB_METHOD_LIN) {
        rc = qpb_lin_project(plan, B, counts_out, physical, final_rho, stream);
        if (rc != QPB_OK) return rc;
        if (iters_out) QPB_CUDA(cudaMemsetAsync(iters_out, 0, sizeof(int32_t) * (size_t)B, st));
    } else {
        const double* start = nullptr;
        const int* order = nullptr;
        const bool pauli2 = max_iter > 0 && mle_variant(plan) == QPB_MLE_PAULI2;
        if (init == QPB_INIT_LIN) {
            // state.py:209: the start is point_estimate("lin") with its default physical=True
            rc = pauli2 ? launch_lin_project_small(plan, B, counts_out, 1, buf0, st, &order) : QPB_ERR_UNSUPPORTED;
            if (rc == QPB_ERR_UNSUPPORTED) rc = qpb_lin_project(plan, B, counts_out, 1, buf0, stream);
            if (rc != QPB_OK) return rc;
            start = buf0;
        }
        if (pauli2 && dist_kind != QPB_DIST_HS) {
            rc = launch_mle_small(plan, B, counts_out, start, max_iter, tol, final_rho, iters_out, st, nullptr, nullptr,
                                  nullptr, 1, order);
            if (rc != QPB_OK) return rc;
            return launch_distance(plan->d, B, final_rho, ref, dist_kind, dist, st);
        }
        if (dist_kind == QPB_DIST_HS && pauli2) {
            // two-qubit Pauli-axis POVMs: the MLE kernel writes the distances itself; the states are stored only if
            // the caller asked for them
            bool done = false;
            rc = launch_mle_small(plan, B, counts_out, start, max_iter, tol, final_rho, iters_out, st, ref, dist, &done,
                                  rho_out ? 1 : 0, order);
            if (rc != QPB_OK) return rc;
            if (done) return QPB_OK;
            return launch_distance(plan->d, B, final_rho, ref, dist_kind, dist, st);
        }
        rc = qpb_mle_rrr(plan, B, counts_out, start, max_iter, tol, final_rho, iters_out, stream);
        if (rc != QPB_OK) return rc;
    }
    return launch_distance(plan->d, B, final_rho, ref, dist_kind, dist, st);
}

}  // extern "C"
