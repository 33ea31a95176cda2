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

// Scheduling class of a sample for the R.rho.R kernel that follows, from the smallest POSITIVE eigenvalue of the
// UNPROJECTED linear estimate: the smaller it is, the longer the iteration tends to run (CPU study of oracle
// trajectories at 2 qubits: Spearman -0.61 ... -0.77 with the iteration count on states with one or two small
// eigenvalues; every sample above 674 iterations of the bench workload sits in the lowest 2 %).  A negative smallest
// eigenvalue does not count: that direction is clipped away by the projection and the start state already sits on
// that part of the boundary (bench workload: such samples need 54 iterations on average, never more than 83).  Log-spaced (4 classes per octave from
// 2^-24), so no scale has to be known.  Only the ORDER in which samples are started depends on it, never a result.
constexpr int kOrderClasses = 98;
__device__ __forceinline__ int order_class(double minpos) {
    if (!(minpos > 0.0)) return kOrderClasses - 1;
    const int code = (int)((unsigned long long)__double_as_longlong(minpos) >> 50);  // exponent + 2 mantissa bits
    const int lo = (1023 - 24) << 2;
    const int cl = code - lo;
    return cl < 0 ? 0 : (cl > kOrderClasses - 2 ? kOrderClasses - 2 : cl);
}

template <int N>
__device__ __forceinline__ double lin_project_one(int K, long b, const double* __restrict__ tab,
                                                  const int32_t* __restrict__ counts, int physical,
                                                  double* __restrict__ rho);

template <int N>
#ifndef LIN_MIN_BLOCKS
#define LIN_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(kLinThreads, LIN_MIN_BLOCKS)
k_lin_project_small(int K, int B, const double* __restrict__ LhT, const int32_t* __restrict__ counts, int physical,
                    double* __restrict__ rho, unsigned char* __restrict__ class_out, unsigned int* __restrict__ class_hist) {
    constexpr int d = 1 << N, D = d * d;
    extern __shared__ __align__(16) double tab[];  // [K][D]
    __shared__ unsigned int hist_s[kOrderClasses];
    for (int e = threadIdx.x; e < K * D; e += kLinThreads) tab[e] = LhT[e];
    if (class_out)
        for (int e = threadIdx.x; e < kOrderClasses; e += kLinThreads) hist_s[e] = 0u;
    __syncthreads();
    const long b = (long)blockIdx.x * kLinThreads + threadIdx.x;
    if (b < B) {
        const double minpos = lin_project_one<N>(K, b, tab, counts, physical, rho);
        if (class_out) {
            const int cl = order_class(minpos);
            class_out[b] = (unsigned char)cl;
            atomicAdd(&hist_s[cl], 1u);
        }
    }
    if (class_out) {  // uniform
        __syncthreads();
        for (int e = threadIdx.x; e < kOrderClasses; e += kLinThreads)
            if (hist_s[e]) atomicAdd(&class_hist[e], hist_s[e]);
    }
}

// order[pos] = sample index, classes ascending (long runners first); positions inside a class in arrival order
__global__ void __launch_bounds__(256)
k_order_scatter(int B, const unsigned char* __restrict__ cls, const unsigned int* __restrict__ hist,
                unsigned int* __restrict__ fill, int* __restrict__ order) {
    __shared__ unsigned int start_s[kOrderClasses], cnt_s[kOrderClasses], base_s[kOrderClasses];
    const int t = threadIdx.x;
    if (t < kOrderClasses) cnt_s[t] = 0u;
    if (t == 0) {
        unsigned int run = 0;
        for (int e = 0; e < kOrderClasses; ++e) {
            start_s[e] = run;
            run += hist[e];
        }
    }
    __syncthreads();
    const long b = (long)blockIdx.x * 256 + t;
    int cl = 0;
    unsigned int rank = 0;
    if (b < B) {
        cl = cls[b];
        rank = atomicAdd(&cnt_s[cl], 1u);
    }
    __syncthreads();
    if (t < kOrderClasses && cnt_s[t]) base_s[t] = start_s[t] + atomicAdd(&fill[t], cnt_s[t]);
    __syncthreads();
    if (b < B) order[base_s[cl] + rank] = (int)b;
}

// returns the smallest positive eigenvalue of the unprojected estimate (0: none, or physical == 0: not computed)
template <int N>
__device__ __forceinline__ double lin_project_one(int K, long b, const double* __restrict__ tab,
                                                  const int32_t* __restrict__ counts, int physical,
                                                  double* __restrict__ rho) {
    constexpr int d = 1 << N, D = d * d;
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
        return 0.0;
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
    double lam[d], tr = 0.0, minpos = 0.0;  // smallest POSITIVE eigenvalue (0: none)
#pragma unroll
    for (int j = 0; j < d; ++j) {
        const double ev = A.re[j][j];
        if (ev > 0.0 && (minpos == 0.0 || ev < minpos)) minpos = ev;
        lam[j] = fmax(ev, kClipState);
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
    return minpos;
}

__global__ void k_identity_order(int B, int* __restrict__ order) {
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) order[b] = (int)b;
}

}  // namespace qpb

extern "C" int qpb_identity_order(int B, int32_t* order, void* stream) {
    if (B <= 0) return QPB_OK;
    qpb::k_identity_order<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(B, order);
    QPB_LAUNCHED("k_identity_order");
    return QPB_OK;
}

namespace qpb {

int launch_lin_project_small(const qpb_state_plan* plan, int B, const int32_t* counts, int physical, double* rho,
                             cudaStream_t st, const int** order_out) {
    if (order_out) *order_out = nullptr;
    if (plan->n > 2) return QPB_ERR_UNSUPPORTED;
    const size_t smem = sizeof(double) * (size_t)plan->K * plan->D;
    if (smem > 96 * 1024) return QPB_ERR_UNSUPPORTED;
    auto kern = plan->n == 1 ? k_lin_project_small<1> : k_lin_project_small<2>;
    if (smem > 48 * 1024) QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (B + kLinThreads - 1) / kLinThreads;
    // start order for the MLE kernel (only worth two more tiny launches when the batch spans several waves of lanes)
    unsigned char* cls = nullptr;
    unsigned int* hist = nullptr;
    int* order = nullptr;
    if (order_out && physical && B >= 4096 && !option(QPB_OPT_NO_MLE_ORDER)) {
        unsigned char* base = static_cast<unsigned char*>(scratch(st, 10, 1024 + sizeof(int) * (size_t)B + (size_t)B));
        if (!base) return QPB_ERR_NOMEM;
        hist = reinterpret_cast<unsigned int*>(base);           // [kOrderClasses] histogram, then [kOrderClasses] fill
        order = reinterpret_cast<int*>(base + 1024);
        cls = base + 1024 + sizeof(int) * (size_t)B;
        QPB_CUDA(cudaMemsetAsync(hist, 0, 1024, st));
    }
    kern<<<grid, kLinThreads, smem, st>>>(plan->K, B, plan->LhT, counts, physical, rho, cls, hist);
    QPB_LAUNCHED("k_lin_project_small");
    if (cls) {
        k_order_scatter<<<(B + 255) / 256, 256, 0, st>>>(B, cls, hist, hist + 128, order);
        QPB_LAUNCHED("k_order_scatter");
        *order_out = order;
    }
    return QPB_OK;
}

}  // namespace qpb
