// Internal plan structures and cross-file launch helpers (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct qpb_state_plan {
    int n = 0, d = 0, D = 0, K = 0;
    double* Ar = nullptr;   // [K][D] packed-Hermitian POVM operators
    double* ArT = nullptr;  // [D][K]
    double* LhT = nullptr;  // [K][D] packed-Hermitian linear-inversion map (NULL if no L given)
    double* A_host = nullptr;   // host copy of the Bloch-basis table A [K][D] (structure detection)
    double* Ar_host = nullptr;
    // n = 3, 4 Pauli-axis structure (mle_axis.cu): canonical slot of every count column, 1e-10/c per slot
    bool axis_ok = false;
    int* axis_slots = nullptr;
    double* axis_epsp = nullptr;  // host copy of Ar for kernels that take the table as a launch parameter
};

struct qpb_process_plan {
    int n = 0, d = 0, S = 0, K = 0;
    int d4 = 0;              // d^4 = number of Choi entries
    double* LinvT = nullptr;  // [S*K][2*d4] : (re, im) of Linv[:, col] per input column, row-major Choi
};

namespace qpb {

int launch_lin_project(const qpb_state_plan* plan, int B, const int32_t* counts, int physical, double* rho,
                       cudaStream_t st);
// Register-resident variant for n <= 2 (lin_small.cu); QPB_ERR_UNSUPPORTED otherwise.
// order_out (optional): receives a device array [B] with the order in which the R.rho.R kernel should start the
// samples (likely long runners first, judged by the smallest eigenvalue of the unprojected estimate), or nullptr.
int launch_lin_project_small(const qpb_state_plan* plan, int B, const int32_t* counts, int physical, double* rho,
                             cudaStream_t st, const int** order_out = nullptr);
// Register-resident Jacobi + projection for d = 8, 16 from packed-Hermitian inputs H [B][d*d] (jacobi_rows.cu)
int launch_project_rows(int d, int B, const double* h_in, double* rho, cudaStream_t st);
int launch_mle_generic(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                       double tol, double* rho, int32_t* iters, cudaStream_t st);
int launch_distance(int d, int B, const double* rho, const double* ref, int kind, double* dist, cudaStream_t st);
// Specialised register-resident kernels for n <= 2 (mle_small.cu).  Returns QPB_ERR_UNSUPPORTED when
// the shape has no specialisation, in which case the caller uses the generic kernel.
// hs_ref / hs_dist / hs_done (optional): the kernel may also write the Hilbert-Schmidt distance of every result to
// the complex d x d centre state hs_ref and then sets *hs_done; if it does and hs_store_rho == 0 the states are
// not written (rho must still be a valid buffer: kernels without the fusion use it).
int launch_mle_small(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                     double tol, double* rho, int32_t* iters, cudaStream_t st, const double* hs_ref = nullptr,
                     double* hs_dist = nullptr, bool* hs_done = nullptr, int hs_store_rho = 1,
                     const int* order = nullptr);

// Batched general-POVM R.rho.R at n = 3, 4 on the DMMA GEMM (mle_tiled.cu); QPB_ERR_UNSUPPORTED for small batches
// or when switched off, in which case the caller uses the warp-per-sample kernel.
bool mle_tiled_applicable(const qpb_state_plan* plan);
int launch_mle_tiled(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                     double tol, double* rho, int32_t* iters, cudaStream_t st);

int mle_variant(const qpb_state_plan* plan);
// C [M][N] = (counts [M][Ktot] normalised per group of G columns) * T [Ktot][N] on DMMA (gemm_dmma.cu)
int launch_gemm_counts(int M, int N, int Ktot, int G, const int32_t* counts, const double* T, double* C,
                       cudaStream_t st);
// P [B][K] = clip(scale * X [B][D] M [K][D]^T) on DMMA (gemm_dmma.cu); QPB_ERR_UNSUPPORTED for shapes it does not take
int launch_probs_gemm(int K, int D, int B, const double* Mtab, const double* X, double scale, int clip, double* P,
                      cudaStream_t st);
int axis_plan_setup(qpb_state_plan* plan, const double* A_host);
int launch_mle_axis(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                    double tol, double* rho, int32_t* iters, cudaStream_t st);

}  // namespace qpb
