// Register-resident R.rho.R maximum likelihood for n <= 2 qubits (d <= 4): ONE THREAD PER SAMPLE.
//
// Each thread keeps rho (packed Hermitian, d*d doubles) and the accumulating R in registers; the
// packed POVM table Ar[K][D] sits in shared memory and is read with warp-uniform (broadcast)
// 128-bit loads; the sample's frequencies f_k sit in shared memory, thread-minor so every access is
// conflict-free.  The iteration never touches HBM.  Iteration counts differ per sample, so warps
// are persistent: a lane whose sample has converged (or hit max_iter) writes it back and pulls the
// next sample index from a global queue while the other lanes of its warp idle only for the
// duration of that refill.
//
// Per iteration and sample (K outcomes, D = d*d):  2*K*D FMA for p_k and R, K divisions,
// (d^3 + d^2(d+1)/2) complex MACs for R rho R -- all on the FP64 pipe, which is the roofline.
#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "plan.h"

#include <cmath>
#include <cstdlib>

namespace qpb {

constexpr int kSmallThreads = 128;
// The FP64 pipe saturates at a few warps per scheduler; fewer resident lanes also mean that more samples
// flow through each lane (better balance) and that the straggler tail runs with less pipe sharing.
constexpr int kMaxBlocksPerSmDefault = 2;
static int max_blocks_per_sm() {
    const int v = option(QPB_OPT_MLE_BLOCKS_PER_SM);
    return v < 1 ? kMaxBlocksPerSmDefault : v;
}
#define kMaxBlocksPerSm max_blocks_per_sm()

// 1/y to within an ulp: hardware seed x0 (relative error e <= 2^-23) and one third-order correction
// x0 (1 + e + e^2), truncation e^3 <= 2^-69 -- three FMAs on the FP64 pipe.
__device__ __forceinline__ double fast_rcp(double y) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(y));
    const double e = fma(-y, x, 1.0);
    const double t = fma(e, e, e);
    return fma(x, t, x);
}

template <int d>
struct Packed {
    // z = H(a, b) of a packed Hermitian array with compile-time indices
    static __device__ __forceinline__ double re(const double (&h)[d * d], int a, int b) {
        return a <= b ? h[a * d + b] : h[b * d + a];
    }
    static __device__ __forceinline__ double im(const double (&h)[d * d], int a, int b) {
        return a == b ? 0.0 : (a < b ? h[b * d + a] : -h[a * d + b]);
    }
};

// rho' = R rho R (unnormalised), packed in / packed out.  Fully unrolled.
template <int d>
__device__ __forceinline__ void rrr_apply(const double (&R)[d * d], const double (&h)[d * d], double (&hn)[d * d]) {
#pragma unroll
    for (int a = 0; a < d; ++a) {
        double sre[d], sim[d];  // row a of S = R rho
#pragma unroll
        for (int b = 0; b < d; ++b) {
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int c = 0; c < d; ++c) {
                const double xr = Packed<d>::re(R, a, c), xi = Packed<d>::im(R, a, c);
                const double yr = Packed<d>::re(h, c, b), yi = Packed<d>::im(h, c, b);
                re = fma(xr, yr, re);
                if (c != b) im = fma(xr, yi, im);
                if (a != c && c != b) re = fma(-xi, yi, re);
                if (a != c) im = fma(xi, yr, im);
            }
            sre[b] = re;
            sim[b] = im;
        }
#pragma unroll
        for (int b = a; b < d; ++b) {  // (S R)[a][b], upper triangle only
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int c = 0; c < d; ++c) {
                const double yr = Packed<d>::re(R, c, b), yi = Packed<d>::im(R, c, b);
                re = fma(sre[c], yr, re);
                if (c != b) re = fma(-sim[c], yi, re);
                if (a != b) {
                    im = fma(sim[c], yr, im);
                    if (c != b) im = fma(sre[c], yi, im);
                }
            }
            hn[a * d + b] = re;
            if (a != b) hn[b * d + a] = im;
        }
    }
}

template <int N>
__global__ void __launch_bounds__(kSmallThreads)
k_mle_rrr_small(int K, int B, const double* __restrict__ Ar, const int32_t* __restrict__ counts,
                const double* __restrict__ rho0, int max_iter, double tol, double* __restrict__ rho,
                int32_t* __restrict__ iters, unsigned int* __restrict__ queue) {
    constexpr int d = 1 << N, D = d * d;
    extern __shared__ __align__(16) double sm[];
    double* tab = sm;                    // [K][D]
    double* fs = sm + (size_t)K * D;     // [K][kSmallThreads]
    const int tid = threadIdx.x, lane = tid & 31;
    for (int e = tid; e < K * D; e += kSmallThreads) tab[e] = Ar[e];
    __syncthreads();

    double h[D];
    int it = 0;
    long b = -1;        // sample owned by this lane, -1 = none
    bool alive = true;  // false once the queue ran dry for this lane
    const double tol2 = tol * tol;

    while (true) {
        // ---- refill lanes without work -------------------------------------------------------
        const bool want = alive && b < 0;
        const unsigned need = __ballot_sync(0xffffffffu, want);
        if (need) {
            unsigned base = 0;
            const int leader = __ffs(need) - 1;
            if (lane == leader) base = atomicAdd(queue, (unsigned)__popc(need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want) {
                const long nb = (long)base + __popc(need & ((1u << lane) - 1u));
                if (nb < B) {
                    b = nb;
                    it = 0;
                    const int32_t* c = counts + b * K;
                    long long tot = 0;
                    for (int k = 0; k < K; ++k) tot += c[k];
                    const double total = (double)tot;
                    const FreqDiv freq(total);
                    for (int k = 0; k < K; ++k) fs[k * kSmallThreads + tid] = freq((double)c[k]);
                    if (rho0) {
                        const double* r0 = rho0 + b * 2 * D;
#pragma unroll
                        for (int a = 0; a < d; ++a)
#pragma unroll
                            for (int bb = 0; bb < d; ++bb)
                                h[a * d + bb] = (a <= bb) ? r0[2 * (a * d + bb)] : r0[2 * (bb * d + a) + 1];
                    } else {
#pragma unroll
                        for (int e = 0; e < D; ++e) h[e] = (e / d == e % d) ? 1.0 / d : 0.0;
                    }
                } else {
                    alive = false;
                }
            }
        }
        if (!__any_sync(0xffffffffu, b >= 0)) break;

        // ---- one R.rho.R iteration (lanes without a sample are predicated off) ---------------
        bool finished = false;
        if (b >= 0) {
            if (max_iter <= 0) {
                finished = true;
            } else {
                double h2[D], R[D];
#pragma unroll
                for (int e = 0; e < D; ++e) {
                    h2[e] = (e / d == e % d) ? h[e] : 2.0 * h[e];
                    R[e] = 0.0;
                }
#pragma unroll 2
                for (int k = 0; k < K; ++k) {
                    double row[D];
                    const double2* src = reinterpret_cast<const double2*>(tab + (size_t)k * D);
#pragma unroll
                    for (int e = 0; e < D / 2; ++e) {
                        const double2 v = src[e];
                        row[2 * e] = v.x;
                        row[2 * e + 1] = v.y;
                    }
                    double p0 = 0.0, p1 = 0.0;
#pragma unroll
                    for (int e = 0; e < D; e += 2) {
                        p0 = fma(row[e], h2[e], p0);
                        p1 = fma(row[e + 1], h2[e + 1], p1);
                    }
                    const double w = fs[k * kSmallThreads + tid] * fast_rcp((p0 + p1) + kLogGuard);
#pragma unroll
                    for (int e = 0; e < D; ++e) R[e] = fma(w, row[e], R[e]);
                }
                double hn[D];
                rrr_apply<d>(R, h, hn);
                double tr = 0.0;
#pragma unroll
                for (int a = 0; a < d; ++a) tr += hn[a * d + a];
                const double inv = fast_rcp(tr);
                double del = 0.0;
#pragma unroll
                for (int e = 0; e < D; ++e) {
                    const double v = hn[e] * inv;
                    const double df = v - h[e];
                    del = fma((e / d == e % d) ? df : 2.0 * df, df, del);
                    h[e] = v;
                }
                ++it;
                finished = (del < tol2) || (it >= max_iter);
            }
        }
        // ---- write back finished samples ------------------------------------------------------
        if (finished) {
            double* out = rho + b * 2 * D;
#pragma unroll
            for (int a = 0; a < d; ++a)
#pragma unroll
                for (int bb = 0; bb < d; ++bb) {
                    double2 z;
                    z.x = Packed<d>::re(h, a, bb);
                    z.y = Packed<d>::im(h, a, bb);
                    reinterpret_cast<double2*>(out)[a * d + bb] = z;
                }
            if (iters) iters[b] = it;
            b = -1;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Compile-time-K variant: the POVM table travels as a __grid_constant__ kernel parameter, i.e. in
// constant bank 0, and the k loop is fully unrolled, so every table entry is an immediate constant
// operand of a DFMA -- no loads, no registers for table rows, and the compiler is free to interleave
// the dot products, reciprocals and rank-one updates of neighbouring outcomes to hide FP64 latency.
// tab2 is the same table with doubled off-diagonal entries (Tr(E rho) in the packed basis).
// ------------------------------------------------------------------------------------------------
template <int N, int K>
struct ConstTables {
    double tab[K * (1 << (2 * N))];
    double tab2[K * (1 << (2 * N))];
};


template <int N, int K>
__global__ void __launch_bounds__(kSmallThreads)
k_mle_rrr_const(const __grid_constant__ ConstTables<N, K> ct, int B, const int32_t* __restrict__ counts,
                const double* __restrict__ rho0, int max_iter, double tol, double* __restrict__ rho,
                int32_t* __restrict__ iters, unsigned int* __restrict__ queue) {
    constexpr int d = 1 << N, D = d * d;
    constexpr int G = (K % 6 == 0) ? 6 : 4;  // outcomes handled per software-pipelined group
    static_assert(K % G == 0, "K must be a multiple of the group size");
    extern __shared__ __align__(16) double sm[];
    double* fs = sm;  // [K][kSmallThreads]
    const int tid = threadIdx.x, lane = tid & 31;

    double h[D];
    int it = 0;
    long b = -1;
    bool alive = true;
    const double tol2 = tol * tol;

    while (true) {
        const bool want = alive && b < 0;
        const unsigned need = __ballot_sync(0xffffffffu, want);
        if (need) {
            unsigned base = 0;
            const int leader = __ffs(need) - 1;
            if (lane == leader) base = atomicAdd(queue, (unsigned)__popc(need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want) {
                const long nb = (long)base + __popc(need & ((1u << lane) - 1u));
                if (nb < B) {
                    b = nb;
                    it = 0;
                    const int32_t* c = counts + b * K;
                    int cc[K];
                    long long tot = 0;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        cc[k] = c[k];
                        tot += cc[k];
                    }
                    const double total = (double)tot;
                    const FreqDiv freq(total);
#pragma unroll
                    for (int k = 0; k < K; ++k) fs[k * kSmallThreads + tid] = freq((double)cc[k]);
                    if (rho0) {
                        const double2* r0 = reinterpret_cast<const double2*>(rho0) + b * D;
#pragma unroll
                        for (int a = 0; a < d; ++a)
#pragma unroll
                            for (int bb = a; bb < d; ++bb) {
                                const double2 z = r0[a * d + bb];
                                h[a * d + bb] = z.x;
                                if (a != bb) h[bb * d + a] = z.y;
                            }
                    } else {
#pragma unroll
                        for (int e = 0; e < D; ++e) h[e] = (e / d == e % d) ? 1.0 / d : 0.0;
                    }
                } else {
                    alive = false;
                }
            }
        }
        if (!__any_sync(0xffffffffu, b >= 0)) break;

        bool finished = false;
        if (b >= 0) {
            if (max_iter <= 0) {
                finished = true;
            } else {
                double R[D];
#pragma unroll
                for (int e = 0; e < D; ++e) R[e] = 0.0;
#pragma unroll
                for (int g = 0; g < K / G; ++g) {
                    double w[G];
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        const int k = g * G + j;
                        double p0 = kLogGuard, p1 = 0.0;
#pragma unroll
                        for (int e = 0; e < D; e += 2) {
                            p0 = fma(ct.tab2[k * D + e], h[e], p0);
                            p1 = fma(ct.tab2[k * D + e + 1], h[e + 1], p1);
                        }
                        w[j] = p0 + p1;
                    }
#pragma unroll
                    for (int j = 0; j < G; ++j) w[j] = fs[(g * G + j) * kSmallThreads + tid] * fast_rcp(w[j]);
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        const int k = g * G + j;
#pragma unroll
                        for (int e = 0; e < D; ++e) R[e] = fma(w[j], ct.tab[k * D + e], R[e]);
                    }
                }
                double hn[D];
                rrr_apply<d>(R, h, hn);
                double tr = 0.0;
#pragma unroll
                for (int a = 0; a < d; ++a) tr += hn[a * d + a];
                const double inv = fast_rcp(tr);
                double del = 0.0;
#pragma unroll
                for (int e = 0; e < D; ++e) {
                    const double v = hn[e] * inv;
                    const double df = v - h[e];
                    del = fma((e / d == e % d) ? df : 2.0 * df, df, del);
                    h[e] = v;
                }
                ++it;
                finished = (del < tol2) || (it >= max_iter);
            }
        }
        if (finished) {
            double2* out = reinterpret_cast<double2*>(rho) + b * D;
#pragma unroll
            for (int a = 0; a < d; ++a)
#pragma unroll
                for (int bb = 0; bb < d; ++bb) {
                    double2 z;
                    z.x = Packed<d>::re(h, a, bb);
                    z.y = Packed<d>::im(h, a, bb);
                    out[a * d + bb] = z;
                }
            if (iters) iters[b] = it;
            b = -1;
        }
    }
}

template <int N, int K>
static int launch_const(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                        double tol, double* rho, int32_t* iters, unsigned int* queue, cudaStream_t st) {
    constexpr int D = 1 << (2 * N), d = 1 << N;
    ConstTables<N, K> ct;  // filled per launch from the plan's host copy; copied into the launch parameters
    for (int k = 0; k < K; ++k)
        for (int e = 0; e < D; ++e) {
            const double v = plan->Ar_host[(size_t)k * D + e];
            ct.tab[k * D + e] = v;
            ct.tab2[k * D + e] = (e / d == e % d) ? v : 2.0 * v;
        }
    const size_t smem = sizeof(double) * (size_t)K * kSmallThreads;
    auto kern = k_mle_rrr_const<N, K>;
    if (smem > 48 * 1024) QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    QPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSmallThreads, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > kMaxBlocksPerSm) per_sm = kMaxBlocksPerSm;
    long blocks = (long)num_sms() * per_sm;
    const long need = ((long)B + kSmallThreads - 1) / kSmallThreads;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(int)blocks, kSmallThreads, smem, st>>>(ct, B, counts, rho0, max_iter, tol, rho, iters, queue);
    QPB_LAUNCHED("k_mle_rrr_const");
    return QPB_OK;
}


// Two-qubit Pauli-axis POVMs have their own table-free kernel (mle_pauli2.cu).
bool plan_is_pauli2(const qpb_state_plan* plan);
int launch_mle_pauli2(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                      double tol, double* rho, int32_t* iters, cudaStream_t st, const double* hs_ref, double* hs_dist,
                      bool* hs_done, int hs_store_rho, const int* order);

// Which kernel qpb_mle_rrr will run for this plan (also exported through qpb_mle_variant for bench.py).
int mle_variant(const qpb_state_plan* plan) {
    if (plan->n > 2) {
        if (plan->axis_ok && !option(QPB_OPT_NO_AXIS_KERNEL)) return QPB_MLE_AXIS;
        return mle_tiled_applicable(plan) ? QPB_MLE_TILED : QPB_MLE_GENERIC;  // (batches below 32 still take GENERIC)
    }
    if (plan->n == 2 && plan->A_host && !option(QPB_OPT_NO_PAULI_KERNEL) && plan_is_pauli2(plan)) return QPB_MLE_PAULI2;
    if (plan->Ar_host && !option(QPB_OPT_NO_CONST_KERNEL)) {
        if ((plan->n == 2 && (plan->K == 36 || plan->K == 16)) || (plan->n == 1 && (plan->K == 6 || plan->K == 4)))
            return QPB_MLE_CONST;
    }
    const size_t smem = sizeof(double) * ((size_t)plan->K * plan->D + (size_t)plan->K * kSmallThreads);
    return smem > 200 * 1024 ? QPB_MLE_GENERIC : QPB_MLE_SMALL;
}

int launch_mle_small(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                     double tol, double* rho, int32_t* iters, cudaStream_t st, const double* hs_ref, double* hs_dist,
                     bool* hs_done, int hs_store_rho, const int* order) {
    if (hs_done) *hs_done = false;
    if (plan->n > 2) return QPB_ERR_UNSUPPORTED;
    const size_t smem = sizeof(double) * ((size_t)plan->K * plan->D + (size_t)plan->K * kSmallThreads);
    if (smem > 200 * 1024) return QPB_ERR_UNSUPPORTED;
    if (plan->n == 2 && plan->A_host && !option(QPB_OPT_NO_PAULI_KERNEL)) {
        const int rc = launch_mle_pauli2(plan, B, counts, rho0, max_iter, tol, rho, iters, st, hs_ref, hs_dist, hs_done,
                                         hs_store_rho, order);
        if (rc != QPB_ERR_UNSUPPORTED) return rc;
    }
    unsigned int* queue = static_cast<unsigned int*>(scratch(st, 6, sizeof(unsigned int)));
    if (!queue) return QPB_ERR_NOMEM;
    QPB_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned int), st));
    if (plan->Ar_host && !option(QPB_OPT_NO_CONST_KERNEL)) {
        int rc = QPB_ERR_UNSUPPORTED;
        if (plan->n == 2 && plan->K == 36) rc = launch_const<2, 36>(plan, B, counts, rho0, max_iter, tol, rho, iters, queue, st);
        else if (plan->n == 2 && plan->K == 16) rc = launch_const<2, 16>(plan, B, counts, rho0, max_iter, tol, rho, iters, queue, st);
        else if (plan->n == 1 && plan->K == 6) rc = launch_const<1, 6>(plan, B, counts, rho0, max_iter, tol, rho, iters, queue, st);
        else if (plan->n == 1 && plan->K == 4) rc = launch_const<1, 4>(plan, B, counts, rho0, max_iter, tol, rho, iters, queue, st);
        if (rc != QPB_ERR_UNSUPPORTED) return rc;
    }
    auto kern = plan->n == 1 ? k_mle_rrr_small<1> : k_mle_rrr_small<2>;
    if (smem > 48 * 1024) QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    QPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSmallThreads, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > kMaxBlocksPerSm) per_sm = kMaxBlocksPerSm;
    long blocks = (long)num_sms() * per_sm;
    const long need = ((long)B + kSmallThreads - 1) / kSmallThreads;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(int)blocks, kSmallThreads, smem, st>>>(plan->K, B, plan->Ar, counts, rho0, max_iter, tol, rho, iters, queue);
    QPB_LAUNCHED("k_mle_rrr_small");
    return QPB_OK;
}

}  // namespace qpb
