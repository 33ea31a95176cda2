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
    const char* e = getenv("QPB_MLE_BLOCKS_PER_SM");
    const int v = e ? atoi(e) : kMaxBlocksPerSmDefault;
    return v < 1 ? 1 : v;
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


// ------------------------------------------------------------------------------------------------
// Pauli-axis POVMs on two qubits ('proj', 'proj-set', 'proj4', any shot weights): every effect is
//     E_k = c_k (1 + s1 sigma_a1) (x) (1 + s2 sigma_a2),   a in {X,Y,Z}, s = +-1,
// so with S_ij = Tr(sigma_i (x) sigma_j rho) the probabilities are p_k = c_k (S_00 + s1 S_a0 + s2 S_0b + s1 s2 S_ab)
// and R = sum_k w_k (1 + s1 sigma_a1) (x) (1 + s2 sigma_a2) has Pauli coefficients that are signed sums of
// the weights.  The whole K x D contraction collapses to ~200 additions and there is NO table: the only
// per-iteration loads are the sample's 36 frequencies from shared memory.  Effects are addressed by the
// canonical slot (alpha, beta), alpha = 2*(axis-1) + (sign<0); the plan maps count columns to slots.
// ------------------------------------------------------------------------------------------------
namespace pauli2 {
__host__ __device__ constexpr int phase(int i, int a, int b) {  // sigma_i[a][b] = i^phase, or -1 if zero
    int k = 0;
    for (int j = 0; j < 2; ++j) {
        const int sh = 1 - j;
        const int dig = (i >> (2 * sh)) & 3, aj = (a >> sh) & 1, bj = (b >> sh) & 1;
        if (dig == 0) {
            if (aj != bj) return -1;
        } else if (dig == 1) {
            if (aj == bj) return -1;
        } else if (dig == 2) {
            if (aj == bj) return -1;
            k += (aj == 0) ? 3 : 1;
        } else {
            if (aj != bj) return -1;
            k += 2 * aj;
        }
    }
    return k & 3;
}
__host__ __device__ constexpr int re_of(int ph) { return ph == 0 ? 1 : (ph == 2 ? -1 : 0); }
__host__ __device__ constexpr int im_of(int ph) { return ph == 1 ? 1 : (ph == 3 ? -1 : 0); }
// coefficient of packed h[e] in S_i = Tr(sigma_i rho)
__host__ __device__ constexpr int s_coef(int i, int e) {
    const int a = e / 4, b = e % 4;
    if (a == b) return re_of(phase(i, a, a));
    if (a < b) return 2 * re_of(phase(i, b, a));
    return -2 * im_of(phase(i, a, b));
}
// coefficient of g_i in packed(sum_i g_i sigma_i)[e]
__host__ __device__ constexpr int r_coef(int e, int i) {
    const int a = e / 4, b = e % 4;
    if (a <= b) return re_of(phase(i, a, b));
    return im_of(phase(i, b, a));
}
template <int I, int E>
struct SC {
    static constexpr int v = s_coef(I, E);
};
template <int E, int I>
struct RC {
    static constexpr int v = r_coef(E, I);
};
template <int I, int E = 0>
__device__ __forceinline__ double s_sum(const double (&h)[16], double acc) {
    if constexpr (E == 16) {
        return acc;
    } else {
        if constexpr (SC<I, E>::v != 0) acc = fma((double)SC<I, E>::v, h[E], acc);
        return s_sum<I, E + 1>(h, acc);
    }
}
template <int E, int I = 0>
__device__ __forceinline__ double r_sum(const double (&g)[16], double acc) {
    if constexpr (I == 16) {
        return acc;
    } else {
        if constexpr (RC<E, I>::v != 0) acc = fma((double)RC<E, I>::v, g[I], acc);
        return r_sum<E, I + 1>(g, acc);
    }
}
template <int I = 0>
__device__ __forceinline__ void all_s(const double (&h)[16], double (&s)[16]) {
    if constexpr (I < 16) {
        s[I] = s_sum<I>(h, 0.0);
        all_s<I + 1>(h, s);
    }
}
template <int E = 0>
__device__ __forceinline__ void all_r(const double (&g)[16], double (&R)[16]) {
    if constexpr (E < 16) {
        R[E] = r_sum<E>(g, 0.0);
        all_r<E + 1>(g, R);
    }
}
}  // namespace pauli2

struct PauliParams {
    double epsp[36];      // 1e-10 / c_k per slot (1.0 for unused slots)
    int slot_of_col[36];  // canonical slot of count column k
    int K;
    int uniform;          // all used slots have the same guard (then every entry of epsp holds it)
};

// Tail merging.  The kernel runs ONE 256-thread CTA per SM: warps w and w+4 sit on the same scheduler.  While the
// queue has work both are full; once it is empty each holds a few long-running samples and the two instruction
// streams halve each other's FP64 issue rate.  The upper warp ("donor") therefore hands its remaining samples to
// idle lanes of the lower warp ("receiver") through a shared-memory mailbox as soon as it has at most kDonateMax
// of them, and exits.  A sample's iteration sequence is untouched, so results are bit-identical.
// state word per pair: 0 open | 3 donor writing | 1 mail ready | 4 mail taken | 2 receiver gone
constexpr int kPauliThreads = 256;
constexpr int kDonateMax = 16;
struct TailMail {
    double h[kDonateMax][16];
    long b[kDonateMax];
    int it[kDonateMax];
    int col[kDonateMax];
};

template <bool UNIFORM_GUARD>  // all used slots share one 1e-10/c: fold it into S_00 instead of 36 additions
__global__ void __launch_bounds__(kPauliThreads, 1)
k_mle_rrr_pauli2(const __grid_constant__ PauliParams pp, int B, const int32_t* __restrict__ counts,
                 const double* __restrict__ rho0, int max_iter, double tol, double* __restrict__ rho,
                 int32_t* __restrict__ iters, unsigned int* __restrict__ queue, int merge_tail,
                 const double* __restrict__ hs_ref, double* __restrict__ hs_dist) {
    constexpr int d = 4, D = 16;
    extern __shared__ __align__(16) double sm[];
    double* fs = sm;  // [36][kPauliThreads], column = owning thread at load time
    __shared__ TailMail mail[4];
    __shared__ int pair_state[4];
    __shared__ int mail_count[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pair = warp & 3;
    const bool donor = warp >= 4;
    const int K = pp.K;
    for (int sl = 0; sl < 36; ++sl) fs[sl * kPauliThreads + tid] = 0.0;
    if (tid < 4) pair_state[tid] = 0;
    __syncthreads();

    double h[D];
    int it = 0;
    int col = tid;      // fs column of the sample this lane owns
    long b = -1;
    bool alive = true;  // false once the queue ran dry for this lane
    bool can_merge = merge_tail != 0 && blockDim.x == kPauliThreads;
    bool drained = false;  // warp-uniform: some lane has seen the queue empty (it never refills again)
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
                    col = tid;
                    // all K count loads are issued together (independent, predicated), then normalised
                    const int32_t* c = counts + b * K;
                    int cc[36];
#pragma unroll
                    for (int k = 0; k < 36; ++k) cc[k] = (k < K) ? c[k] : 0;
                    int tot = 0;
#pragma unroll
                    for (int k = 0; k < 36; ++k) tot += cc[k];
                    const FreqDiv freq((double)tot);
#pragma unroll
                    for (int k = 0; k < 36; ++k)
                        if (k < K) fs[pp.slot_of_col[k] * kPauliThreads + tid] = freq((double)cc[k]);
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
        unsigned active = __ballot_sync(0xffffffffu, b >= 0);
        if (!drained) drained = __any_sync(0xffffffffu, !alive);

        // ---- tail merging (only once the queue is empty) ----------------------------------------
        if (can_merge && drained) {
            if (donor) {
                const int nact = __popc(active);
                if (nact > 0 && nact <= kDonateMax) {
                    int old = 0;
                    if (lane == 0) old = atomicCAS(&pair_state[pair], 0, 3);
                    old = __shfl_sync(0xffffffffu, old, 0);
                    if (old == 0) {
                        if (b >= 0) {
                            const int slot = __popc(active & ((1u << lane) - 1u));
#pragma unroll
                            for (int e = 0; e < D; ++e) mail[pair].h[slot][e] = h[e];
                            mail[pair].b[slot] = b;
                            mail[pair].it[slot] = it;
                            mail[pair].col[slot] = col;
                        }
                        __syncwarp();
                        if (lane == 0) {
                            mail_count[pair] = nact;
                            __threadfence_block();
                            atomicExch(&pair_state[pair], 1);
                        }
                        b = -1;
                        active = 0;
                    } else {
                        can_merge = false;  // the receiver has already left: finish alone
                    }
                }
            } else {
                int st = 0;
                if (lane == 0) {
                    st = *(volatile int*)&pair_state[pair];
                    if (active == 0 && st != 1) {
                        // about to leave: close the pair, or wait for a donor that is writing its mail
                        const int old = atomicCAS(&pair_state[pair], 0, 2);
                        st = old;
                        while (st == 3) st = *(volatile int*)&pair_state[pair];
                    }
                }
                st = __shfl_sync(0xffffffffu, st, 0);
                if (st == 1) {
                    __threadfence_block();
                    const int cnt = mail_count[pair];
                    const unsigned idle = ~active;
                    if (__popc(idle) >= cnt) {
                        const int rank = __popc(idle & ((1u << lane) - 1u));
                        if (b < 0 && rank < cnt) {
#pragma unroll
                            for (int e = 0; e < D; ++e) h[e] = mail[pair].h[rank][e];
                            b = mail[pair].b[rank];
                            it = mail[pair].it[rank];
                            col = mail[pair].col[rank];
                        }
                        __syncwarp();
                        if (lane == 0) atomicExch(&pair_state[pair], 4);
                        active = __ballot_sync(0xffffffffu, b >= 0);
                    }
                } else if (st != 0 && st != 3) {
                    can_merge = false;  // pair closed (mail taken, or we closed it ourselves)
                }
            }
        }
        if (active == 0) break;

        // ---- one R.rho.R iteration (lanes without a sample are predicated off) ---------------
        bool finished = false;
        if (b >= 0) {
            if (max_iter <= 0) {
                finished = true;
            } else {
                double S[16], g[16];
                pauli2::all_s(h, S);
                if (UNIFORM_GUARD) S[0] += pp.epsp[0];
#pragma unroll
                for (int e = 0; e < 16; ++e) g[e] = 0.0;
#pragma unroll
                for (int al = 0; al < 6; ++al) {
                    const int a = al / 2 + 1;
                    const bool neg_a = al & 1;
                    double t[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) t[j] = neg_a ? S[j] - S[a * 4 + j] : S[j] + S[a * 4 + j];
                    double w[6];
#pragma unroll
                    for (int be = 0; be < 6; ++be) {
                        const int bq = be / 2 + 1;
                        const double q = (be & 1) ? t[0] - t[bq] : t[0] + t[bq];
                        w[be] = UNIFORM_GUARD ? q : q + pp.epsp[al * 6 + be];
                    }
#pragma unroll
                    for (int be = 0; be < 6; ++be) w[be] = fs[(al * 6 + be) * kPauliThreads + col] * fast_rcp(w[be]);
                    double u[4];
                    u[0] = ((w[0] + w[1]) + (w[2] + w[3])) + (w[4] + w[5]);
                    u[1] = w[0] - w[1];
                    u[2] = w[2] - w[3];
                    u[3] = w[4] - w[5];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        g[j] += u[j];
                        if (neg_a) g[a * 4 + j] -= u[j];
                        else g[a * 4 + j] += u[j];
                    }
                }
                double R[D];
                pauli2::all_r(g, R);
                double hn[D];
                rrr_apply<d>(R, h, hn);
                const double tr = (hn[0] + hn[5]) + (hn[10] + hn[15]);
                const double inv = fast_rcp(tr);
                double dl[4] = {0.0, 0.0, 0.0, 0.0};  // four partial sums: no 16-deep dependent FMA chain
#pragma unroll
                for (int e = 0; e < D; ++e) {
                    const double v = hn[e] * inv;
                    const double df = v - h[e];
                    dl[e & 3] = fma((e / d == e % d) ? df : 2.0 * df, df, dl[e & 3]);
                    h[e] = v;
                }
                const double del = (dl[0] + dl[1]) + (dl[2] + dl[3]);
                ++it;
                finished = (del < tol2) || (it >= max_iter);
            }
        }
        if (finished) {
            if (rho) {
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
            }
            if (hs_dist) {
                // hs_dst(rho, ref) = sqrt(|Tr (rho - ref)^2|) / sqrt 2, no conjugate (quantpy/geometry.py:5-17); same
                // formula as k_distance, evaluated here so that the bootstrap needs neither the state nor a fourth kernel
                double sr = 0.0, si = 0.0;
#pragma unroll
                for (int a = 0; a < d; ++a)
#pragma unroll
                    for (int bb = 0; bb < d; ++bb) {
                        const double ur = Packed<d>::re(h, a, bb) - __ldg(hs_ref + 2 * (a * d + bb));
                        const double ui = Packed<d>::im(h, a, bb) - __ldg(hs_ref + 2 * (a * d + bb) + 1);
                        const double vr = Packed<d>::re(h, bb, a) - __ldg(hs_ref + 2 * (bb * d + a));
                        const double vi = Packed<d>::im(h, bb, a) - __ldg(hs_ref + 2 * (bb * d + a) + 1);
                        sr += ur * vr - ui * vi;
                        si += ur * vi + ui * vr;
                    }
                hs_dist[b] = sqrt(sqrt(sr * sr + si * si)) / sqrt(2.0);
            }
            if (iters) iters[b] = it;
            b = -1;
        }
    }
}

// Recognise a two-qubit Pauli-axis POVM from the Bloch-basis table A [K][16] (host copy).
static bool detect_pauli2(const double* A, int K, PauliParams* pp) {
    if (K < 1 || K > 36) return false;
    bool used[36] = {false};
    for (int sl = 0; sl < 36; ++sl) pp->epsp[sl] = 1.0;
    pp->K = K;
    for (int k = 0; k < K; ++k) {
        const double* r = A + (size_t)k * 16;
        const double c = r[0];
        if (!(c > 0.0)) return false;
        int a1 = 0, a2 = 0, s1 = 0, s2 = 0;
        for (int i = 1; i < 4; ++i) {
            if (r[i * 4] != 0.0) {
                if (a1 || fabs(fabs(r[i * 4]) - c) > 1e-14 * c) return false;
                a1 = i;
                s1 = r[i * 4] > 0 ? 1 : -1;
            }
            if (r[i] != 0.0) {
                if (a2 || fabs(fabs(r[i]) - c) > 1e-14 * c) return false;
                a2 = i;
                s2 = r[i] > 0 ? 1 : -1;
            }
        }
        if (!a1 || !a2) return false;
        for (int i = 1; i < 4; ++i)
            for (int j = 1; j < 4; ++j) {
                const double want = (i == a1 && j == a2) ? s1 * s2 * c : 0.0;
                if (fabs(r[i * 4 + j] - want) > 1e-14 * c) return false;
            }
        const int slot = (2 * (a1 - 1) + (s1 < 0)) * 6 + (2 * (a2 - 1) + (s2 < 0));
        if (used[slot]) return false;
        used[slot] = true;
        pp->slot_of_col[k] = slot;
        pp->epsp[slot] = kLogGuard / c;
    }
    const double first = pp->epsp[pp->slot_of_col[0]];
    pp->uniform = 1;
    for (int k = 0; k < K; ++k)
        if (pp->epsp[pp->slot_of_col[k]] != first) pp->uniform = 0;
    if (pp->uniform)
        for (int sl = 0; sl < 36; ++sl) pp->epsp[sl] = first;  // unused slots have f = 0, any positive guard works
    return true;
}

// Which kernel qpb_mle_rrr will run for this plan (also exported through qpb_mle_variant for bench.py).
int mle_variant(const qpb_state_plan* plan) {
    if (plan->n > 2) return (plan->axis_ok && !getenv("QPB_NO_AXIS_KERNEL")) ? QPB_MLE_AXIS : QPB_MLE_GENERIC;
    if (plan->n == 2 && plan->A_host && !getenv("QPB_NO_PAULI_KERNEL")) {
        PauliParams pp;
        if (detect_pauli2(plan->A_host, plan->K, &pp)) return QPB_MLE_PAULI2;
    }
    if (plan->Ar_host && !getenv("QPB_NO_CONST_KERNEL")) {
        if ((plan->n == 2 && (plan->K == 36 || plan->K == 16)) || (plan->n == 1 && (plan->K == 6 || plan->K == 4)))
            return QPB_MLE_CONST;
    }
    const size_t smem = sizeof(double) * ((size_t)plan->K * plan->D + (size_t)plan->K * kSmallThreads);
    return smem > 200 * 1024 ? QPB_MLE_GENERIC : QPB_MLE_SMALL;
}

int launch_mle_small(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                     double tol, double* rho, int32_t* iters, cudaStream_t st, const double* hs_ref, double* hs_dist,
                     bool* hs_done, int hs_store_rho) {
    if (hs_done) *hs_done = false;
    if (plan->n > 2) return QPB_ERR_UNSUPPORTED;
    const size_t smem = sizeof(double) * ((size_t)plan->K * plan->D + (size_t)plan->K * kSmallThreads);
    if (smem > 200 * 1024) return QPB_ERR_UNSUPPORTED;
    unsigned int* queue = static_cast<unsigned int*>(scratch(st, 0, sizeof(unsigned int)));
    if (!queue) return QPB_ERR_NOMEM;
    QPB_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned int), st));
    if (plan->n == 2 && plan->A_host && !getenv("QPB_NO_PAULI_KERNEL")) {
        PauliParams pp;
        if (detect_pauli2(plan->A_host, plan->K, &pp)) {
            const size_t smem = sizeof(double) * 36 * kPauliThreads;
            auto pk = pp.uniform ? k_mle_rrr_pauli2<true> : k_mle_rrr_pauli2<false>;
            QPB_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            long blocks = num_sms();  // one 8-warp CTA per SM: warps w and w+4 share a scheduler (tail merging)
            const long need = ((long)B + kPauliThreads - 1) / kPauliThreads;
            if (blocks > need) blocks = need;
            const int merge = getenv("QPB_NO_TAIL_MERGE") ? 0 : 1;
            const bool fuse = hs_ref && hs_dist && hs_done && !getenv("QPB_NO_HS_FUSION");
            // with the distance fused and no caller for the states, they are not written at all
            double* rho_dst = (fuse && hs_store_rho == 0) ? nullptr : rho;
            pk<<<(int)blocks, kPauliThreads, smem, st>>>(pp, B, counts, rho0, max_iter, tol, rho_dst, iters, queue, merge,
                                                        fuse ? hs_ref : nullptr, fuse ? hs_dist : nullptr);
            QPB_LAUNCHED("k_mle_rrr_pauli2");
            if (fuse) *hs_done = true;
            return QPB_OK;
        }
    }
    if (plan->Ar_host && !getenv("QPB_NO_CONST_KERNEL")) {
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
