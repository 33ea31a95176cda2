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

namespace qpb {

constexpr int kSmallThreads = 128;

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
                    for (int k = 0; k < K; ++k) fs[k * kSmallThreads + tid] = (double)c[k] / total;
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
                    const double w = fs[k * kSmallThreads + tid] / ((p0 + p1) + kLogGuard);
#pragma unroll
                    for (int e = 0; e < D; ++e) R[e] = fma(w, row[e], R[e]);
                }
                double hn[D];
                rrr_apply<d>(R, h, hn);
                double tr = 0.0;
#pragma unroll
                for (int a = 0; a < d; ++a) tr += hn[a * d + a];
                const double inv = 1.0 / tr;
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

int launch_mle_small(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                     double tol, double* rho, int32_t* iters, cudaStream_t st) {
    if (plan->n > 2) return QPB_ERR_UNSUPPORTED;
    const size_t smem = sizeof(double) * ((size_t)plan->K * plan->D + (size_t)plan->K * kSmallThreads);
    if (smem > 200 * 1024) return QPB_ERR_UNSUPPORTED;
    unsigned int* queue = nullptr;
    QPB_CUDA(cudaMallocAsync(&queue, sizeof(unsigned int), st));
    QPB_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned int), st));
    auto kern = plan->n == 1 ? k_mle_rrr_small<1> : k_mle_rrr_small<2>;
    if (smem > 48 * 1024) QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    QPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSmallThreads, smem));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;  // FP64 pipe saturates at 1-2 warps per scheduler; fewer lanes = shorter tail
    long blocks = (long)num_sms() * per_sm;
    const long need = ((long)B + kSmallThreads - 1) / kSmallThreads;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(int)blocks, kSmallThreads, smem, st>>>(plan->K, B, plan->Ar, counts, rho0, max_iter, tol, rho, iters, queue);
    QPB_LAUNCHED("k_mle_rrr_small");
    QPB_CUDA(cudaFreeAsync(queue, st));
    return QPB_OK;
}

}  // namespace qpb
