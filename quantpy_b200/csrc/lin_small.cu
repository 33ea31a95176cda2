// Register-resident linear inversion + physical projection for n <= 2 qubits: ONE THREAD PER SAMPLE.
// Replaces StateTomograph._point_estimate_lin + _make_feasible (quantpy/tomography/state.py:191-202,
// 267-273) for d = 2, 4.  The d x d Hermitian eigenproblem is solved by a fully unrolled cyclic Jacobi
// iteration on registers (same rotation as jacobi.cuh); the inversion table LhT [K][D] is staged in
// shared memory and read with warp-uniform loads.
#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "plan.h"

namespace qpb {

constexpr int kLinThreads = 128;

template <int d>
struct RegMat {
    double re[d][d], im[d][d];
};

// One Jacobi rotation on the (P, Q) pivot of Hermitian A, accumulated into V (A <- G^dagger A G, V <- V G).
template <int d, int P, int Q>
__device__ __forceinline__ void jacobi_rotate(RegMat<d>& A, RegMat<d>& V, double tiny2) {
    const double br = A.re[P][Q], bi = A.im[P][Q];
    const double b2 = br * br + bi * bi;
    if (b2 <= tiny2) return;
    const double iab = fast_rsqrt(b2);
    const double tau = 0.5 * (A.re[Q][Q] - A.re[P][P]) * iab;
    const double x = fma(tau, tau, 1.0);
    const double t = (tau >= 0.0 ? 1.0 : -1.0) * fast_recip(fabs(tau) + x * fast_rsqrt(x));
    const double c = fast_rsqrt(fma(t, t, 1.0));
    const double s = t * c;
    const double sur = s * br * iab, sui = s * bi * iab;  // s * u, u = A_pq / |A_pq|
#pragma unroll
    for (int r = 0; r < d; ++r) {  // columns: X'_rp = c X_rp - s conj(u) X_rq ; X'_rq = s u X_rp + c X_rq
        {
            const double xr = A.re[r][P], xi = A.im[r][P], yr = A.re[r][Q], yi = A.im[r][Q];
            A.re[r][P] = c * xr - (sur * yr + sui * yi);
            A.im[r][P] = c * xi - (sur * yi - sui * yr);
            A.re[r][Q] = (sur * xr - sui * xi) + c * yr;
            A.im[r][Q] = (sur * xi + sui * xr) + c * yi;
        }
        {
            const double xr = V.re[r][P], xi = V.im[r][P], yr = V.re[r][Q], yi = V.im[r][Q];
            V.re[r][P] = c * xr - (sur * yr + sui * yi);
            V.im[r][P] = c * xi - (sur * yi - sui * yr);
            V.re[r][Q] = (sur * xr - sui * xi) + c * yr;
            V.im[r][Q] = (sur * xi + sui * xr) + c * yi;
        }
    }
#pragma unroll
    for (int col = 0; col < d; ++col) {  // rows: A'_p. = c A_p. - s u A_q. ; A'_q. = s conj(u) A_p. + c A_q.
        const double xr = A.re[P][col], xi = A.im[P][col], yr = A.re[Q][col], yi = A.im[Q][col];
        A.re[P][col] = c * xr - (sur * yr - sui * yi);
        A.im[P][col] = c * xi - (sur * yi + sui * yr);
        A.re[Q][col] = (sur * xr + sui * xi) + c * yr;
        A.im[Q][col] = (sur * xi - sui * xr) + c * yi;
    }
    A.re[P][Q] = A.im[P][Q] = A.re[Q][P] = A.im[Q][P] = 0.0;
    A.im[P][P] = A.im[Q][Q] = 0.0;
}

template <int d>
__device__ __forceinline__ void jacobi_sweep(RegMat<d>& A, RegMat<d>& V, double tiny2);
template <>
__device__ __forceinline__ void jacobi_sweep<2>(RegMat<2>& A, RegMat<2>& V, double tiny2) {
    jacobi_rotate<2, 0, 1>(A, V, tiny2);
}
template <>
__device__ __forceinline__ void jacobi_sweep<4>(RegMat<4>& A, RegMat<4>& V, double tiny2) {
    jacobi_rotate<4, 0, 1>(A, V, tiny2);
    jacobi_rotate<4, 2, 3>(A, V, tiny2);
    jacobi_rotate<4, 0, 2>(A, V, tiny2);
    jacobi_rotate<4, 1, 3>(A, V, tiny2);
    jacobi_rotate<4, 0, 3>(A, V, tiny2);
    jacobi_rotate<4, 1, 2>(A, V, tiny2);
}

template <int N>
__global__ void __launch_bounds__(kLinThreads)
k_lin_project_small(int K, int B, const double* __restrict__ LhT, const int32_t* __restrict__ counts, int physical,
                    double* __restrict__ rho) {
    constexpr int d = 1 << N, D = d * d;
    extern __shared__ __align__(16) double tab[];  // [K][D]
    for (int e = threadIdx.x; e < K * D; e += kLinThreads) tab[e] = LhT[e];
    __syncthreads();
    const long b = (long)blockIdx.x * kLinThreads + threadIdx.x;
    if (b >= B) return;
    const int32_t* c = counts + b * K;
    long long tot = 0;
    for (int k = 0; k < K; ++k) tot += c[k];
    const FreqDiv freq((double)tot);
    double h[D];
#pragma unroll
    for (int e = 0; e < D; ++e) h[e] = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const double f = freq((double)c[k]);  // state.py:193
        const double2* row = reinterpret_cast<const double2*>(tab + (size_t)k * D);
#pragma unroll
        for (int e = 0; e < D / 2; ++e) {
            const double2 v = row[e];
            h[2 * e] = fma(v.x, f, h[2 * e]);
            h[2 * e + 1] = fma(v.y, f, h[2 * e + 1]);
        }
    }
    RegMat<d> A;
#pragma unroll
    for (int a = 0; a < d; ++a)
#pragma unroll
        for (int bb = 0; bb < d; ++bb) {
            A.re[a][bb] = (a <= bb) ? h[a * d + bb] : h[bb * d + a];
            A.im[a][bb] = (a == bb) ? 0.0 : ((a < bb) ? h[bb * d + a] : -h[a * d + bb]);
        }
    double2* out = reinterpret_cast<double2*>(rho) + b * D;
    if (!physical) {
#pragma unroll
        for (int a = 0; a < d; ++a)
#pragma unroll
            for (int bb = 0; bb < d; ++bb) out[a * d + bb] = make_double2(A.re[a][bb], A.im[a][bb]);
        return;
    }
    RegMat<d> V;
#pragma unroll
    for (int a = 0; a < d; ++a)
#pragma unroll
        for (int bb = 0; bb < d; ++bb) {
            V.re[a][bb] = (a == bb) ? 1.0 : 0.0;
            V.im[a][bb] = 0.0;
        }
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0, fro = 0.0;
#pragma unroll
        for (int a = 0; a < d; ++a)
#pragma unroll
            for (int bb = 0; bb < d; ++bb) {
                const double m2 = A.re[a][bb] * A.re[a][bb] + A.im[a][bb] * A.im[a][bb];
                fro += m2;
                if (a != bb) off += m2;
            }
        if (off <= 1e-30 * fro || fro == 0.0) break;  // relative off-diagonal norm 1e-15: eigenvalues are second order in it
        jacobi_sweep<d>(A, V, 1e-36 * fro);
    }
    double lam[d], tr = 0.0;
#pragma unroll
    for (int j = 0; j < d; ++j) {
        lam[j] = fmax(A.re[j][j], kClipState);
        tr += lam[j];
    }
    const double inv = 1.0 / tr;
#pragma unroll
    for (int a = 0; a < d; ++a)
#pragma unroll
        for (int bb = 0; bb < d; ++bb) {
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int j = 0; j < d; ++j) {
                re += lam[j] * (V.re[a][j] * V.re[bb][j] + V.im[a][j] * V.im[bb][j]);
                im += lam[j] * (V.im[a][j] * V.re[bb][j] - V.re[a][j] * V.im[bb][j]);
            }
            out[a * d + bb] = make_double2(re * inv, im * inv);
        }
}

int launch_lin_project_small(const qpb_state_plan* plan, int B, const int32_t* counts, int physical, double* rho,
                             cudaStream_t st) {
    if (plan->n > 2) return QPB_ERR_UNSUPPORTED;
    const size_t smem = sizeof(double) * (size_t)plan->K * plan->D;
    if (smem > 96 * 1024) return QPB_ERR_UNSUPPORTED;
    auto kern = plan->n == 1 ? k_lin_project_small<1> : k_lin_project_small<2>;
    if (smem > 48 * 1024) QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (B + kLinThreads - 1) / kLinThreads;
    kern<<<grid, kLinThreads, smem, st>>>(plan->K, B, plan->LhT, counts, physical, rho);
    QPB_LAUNCHED("k_lin_project_small");
    return QPB_OK;
}

}  // namespace qpb
