// R.rho.R maximum likelihood for two-qubit Pauli-axis POVMs ('proj', 'proj-set', 'proj4', any shot weights) --
// the kernel behind BASELINE configs[1].
//
// Every effect is E_k = c_k (1 + s1 sigma_a1) (x) (1 + s2 sigma_a2), a in {X,Y,Z}, s = +-1, so with
// S_ij = Tr(sigma_i (x) sigma_j rho) the probabilities are p_k = c_k (S_00 + s1 S_a0 + s2 S_0b + s1 s2 S_ab) and
// R = sum_k w_k (1 + s1 sigma_a1) (x) (1 + s2 sigma_a2) is built from the weights by two 6 -> 2x2 maps (one per
// qubit).  The K x D contraction collapses to ~150 additions and there is NO table: the only per-iteration loads
// are the sample's 36 frequencies.  Effects are addressed by the canonical slot (alpha, beta),
// alpha = 2*(axis-1) + (sign<0); the plan maps count columns to slots.
//
// ONE ARITHMETIC, TWO LANE MAPPINGS.  The iteration is defined as a fixed dataflow graph of IEEE operations
// (explicit __dadd_rn / __dmul_rn / __fma_rn, balanced trees whose additions commute, products as length-4 FMA
// chains in k order).  Two mappings evaluate exactly that graph, so a sample's trajectory -- every bit of every
// iterate, the iteration count, the result -- does not depend on which of them ran it, nor on when it moved from
// one to the other:
//   * thread per sample ("single"): everything in registers, 32 samples per warp.  Cheapest per iteration
//     (681 FP64 instructions per sample-iteration), but one iteration is a ~2400-cycle serial chain for a lone warp.
//   * warp per sample ("W"): the 16 Pauli sums, 36 weights, 24 partial operators and 32 fragment entries are one
//     lane each, exchanged through 1.3 KB of shared memory; each complex 4x4 product is TWO DMMAs (mma.m8n8k4.f64)
//     accumulating into one fragment: [Xr;Xi] x [Yr|Yi], then [-Xi;-Xr] x [Yi|-Yr], which leaves Re (X Y) in
//     columns 0-3 and Im (X Y) in columns 4-7 of rows 0-3, every entry as ONE chain (real terms, then imaginary
//     terms); the step norm is two more (row sums of (m d) x d on a diagonal, a product with ones for their sum).
//     DMMA accumulates exactly like an FMA chain in k order (tools/microbench_fp64.cu: 262144 of 262144 outputs
//     bit-identical), which is what makes the two mappings agree.  ~3x the FP64-pipe cost per iteration, ~5x
//     shorter chain; issue-bound (206 instructions per iteration).
// Iteration counts are heavy-tailed (C2, tol 1e-6: mean 158, p99 578, max 957), so a launch runs the bulk with
// thread-per-sample warps and moves long-running samples (age >= park_age, or whatever is left in a warp once the
// queue is empty) through a global hand-over list to W workers: dedicated warps from the start plus every warp
// that has run out of single-lane work.  Small batches (fewer samples than lanes) run on W workers alone.
#include <cmath>

#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "plan.h"

namespace qpb {

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
// coefficient of packed h[e] in S_i = Tr(sigma_i rho): +-1 on four diagonal entries (I/Z products) or +-2 on two
// off-diagonal entries (anything with an X or Y factor)
__host__ __device__ constexpr int s_coef(int i, int e) {
    const int a = e / 4, b = e % 4;
    if (a == b) return re_of(phase(i, a, a));
    if (a < b) return 2 * re_of(phase(i, b, a));
    return -2 * im_of(phase(i, a, b));
}
struct STerms {
    int n;       // 2 or 4 packed entries
    int e[4];    // ascending
    int neg[4];
    int two;     // all coefficients are +-2
};
__host__ __device__ constexpr STerms s_terms(int i) {
    STerms t{0, {0, 0, 0, 0}, {0, 0, 0, 0}, 0};
    for (int e = 0; e < 16; ++e) {
        const int c = s_coef(i, e);
        if (c != 0) {
            t.e[t.n] = e;
            t.neg[t.n] = c < 0;
            t.two = (c == 2 || c == -2);
            ++t.n;
        }
    }
    return t;
}
}  // namespace pauli2

struct PauliParams {
    double epsp[36];      // 1e-10 / c_k per slot (1.0 for unused slots)
    int slot_of_col[36];  // canonical slot of count column k
    int K;
    int uniform;          // all used slots have the same guard (then every entry of epsp holds it)
};

struct Pauli2Args {
    int B;
    const int32_t* counts;
    const double* rho0;
    const int* order;     // queue position -> sample index (nullptr: identity)
    int max_iter;
    double tol;
    double* rho;
    int32_t* iters;
    const double* hs_ref;
    double* hs_dist;
    // hand-over of long-running samples (global): ctrl[0] sample queue, [1] entries reserved, [2] tickets taken,
    // [3] samples finished
    unsigned int* ctrl;
    double* park_h;       // [cap][16]
    long long* park_b;    // [cap]
    int* park_it;         // [cap]
    unsigned int* park_ready;  // [cap]: == epoch once the entry is complete
    unsigned int park_cap;     // entries of the list; a slot / ticket beyond it is void (the sample stays where it is)
    unsigned int epoch;
    int park_age;         // single-lane samples move to a W worker at this iteration count (INT_MAX: never)
    int park_age_lo;      // ... with a start order: the sample at queue position q moves at park_age_lo + q * park_age_slope
    int park_q2;          //     ... and from queue position park_q2 on (the samples that start after the first wave of
    int park_age_end;     //     lanes) it falls again by park_age_slope2 per position, down to park_age_end: a late start
    float park_age_slope2;//     leaves less time for the slow mapping
    float park_age_slope; //     (capped at park_age): the first positions hold the likely long runners, whose chain
                          //     of thread-per-sample iterations would otherwise end the launch
    int park_plateau;     // ... or from this iteration count on as soon as the step norm has not decreased over the
                          // last 32 iterations (the signature of the few-hundred-iteration plateaus; 0: off)
    int park_live;        // a drained warp with <= park_live live samples hands all of them over (0: never)
    int single_warps;     // warps [0, single_warps) start thread-per-sample, the others are W workers
    int tail_poll;        // once the queue is empty a thread-per-sample warp looks at the hand-over list every tail_poll
                          // iterations (power of two; 0: never): W workers waiting -> it hands over its oldest sample,
                          // entries waiting -> its free lanes adopt them
    int tail_age;         // ... samples younger than this stay where they are
    int adopt;            // free lanes of a drained warp adopt waiting entries (0: W workers only)
    int refill_min;       // lanes without a sample wait until this many of the warp are free (one refill = one queue atomic
                          // and ~1500 cycles of load latency for the whole warp); 1: refill at once
    int merge;            // drained thread-per-sample warps of a CTA pack their samples into fewer warps (0: off)
    long long* trace_s;   // profiling: 4 words per sample (start, hand-over, pick-up, finish times), or null
    long long* trace;     // profiling (qpb_debug_set_trace): 16 words per warp, see tools/pauli2_trace.py; normally null
    int direct;           // W workers take fresh samples from the queue (no single-lane warps)
};

constexpr int kPauliThreadsSingle = 384;  // most thread-per-sample lanes per CTA (12 warps: 3 per scheduler) = stride of the
                                          // frequency columns; the launch decides how many of the 12 warps start that way
constexpr int kPauliWarps = kPauliThreadsSingle / 32;
#ifndef PAULI_CTA_WARPS
#define PAULI_CTA_WARPS 12
#endif
constexpr int kPauliCtaWarps = PAULI_CTA_WARPS;  // warps of a CTA (thread-per-sample + W workers): sets the register budget
constexpr int kCxTop = kPauliWarps, kCxLock = kPauliWarps + 1;  // packing stack: entries, lock (after the live[] words)
constexpr int kPoolWords = 54;             // a sample on the packing stack: 16 state + 36 frequencies + index + age

// ------------------------------------------------------------------------------------------------
// pieces of the dataflow graph shared by both mappings
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
// reciprocal of a probability (+ guard) for the likelihood weight: one Newton step (common.cuh); both lane mappings
// call this one function
__device__ __forceinline__ double weight_recip(double q) { return fast_recip_1step(q); }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double flip(double x, unsigned neg) {  // neg ? -x : x, on the integer pipe
    return __hiloint2double(__double2hiint(x) ^ (int)(neg << 31), __double2loint(x));
}

// hs_dst(rho, ref) = sqrt(|Tr (rho - ref)^2|) / sqrt 2, no conjugate (quantpy/geometry.py:5-20), from the packed state
__device__ __forceinline__ double hs_distance_packed(const double (&h)[16], const double* __restrict__ ref) {
    // term e = (a, b): (rho - ref)[a][b] (rho - ref)[b][a]; summed in the order of k_distance's warp butterfly
    // (element e in lane e: partners e^8, e^4, e^2, e^1), so both paths give the same bits
    double tr[16], ti[16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const double har = a <= b ? h[a * 4 + b] : h[b * 4 + a];
            const double hai = a == b ? 0.0 : (a < b ? h[b * 4 + a] : -h[a * 4 + b]);
            const double ur = har - __ldg(ref + 2 * (a * 4 + b));
            const double ui = hai - __ldg(ref + 2 * (a * 4 + b) + 1);
            const double vr = har - __ldg(ref + 2 * (b * 4 + a));        // Re (rho - ref)[b][a]
            const double vi = -hai - __ldg(ref + 2 * (b * 4 + a) + 1);   // Im (rho - ref)[b][a]
            tr[a * 4 + b] = dadd(0.0, dfma(-ui, vi, dmul(ur, vr)));
            ti[a * 4 + b] = dadd(0.0, dfma(ui, vr, dmul(ur, vi)));
        }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
        for (int e = 0; e < o; ++e) {
            tr[e] = dadd(tr[e], tr[e + o]);
            ti[e] = dadd(ti[e], ti[e + o]);
        }
    const double v = sqrt(sqrt(dfma(ti[0], ti[0], dmul(tr[0], tr[0])))) / sqrt(2.0);
    return v < kZeroBelow ? 0.0 : v;  // geometry.py:17-18
}

// ------------------------------------------------------------------------------------------------
// thread per sample
// ------------------------------------------------------------------------------------------------
namespace single {
template <int I>
__device__ __forceinline__ double s_value(const double (&h)[16]) {
    constexpr pauli2::STerms t = pauli2::s_terms(I);
    const double x0 = t.neg[0] ? -h[t.e[0]] : h[t.e[0]];
    const double x1 = t.neg[1] ? -h[t.e[1]] : h[t.e[1]];
    double s = dadd(x0, x1);
    if constexpr (t.n == 4) {
        const double x2 = t.neg[2] ? -h[t.e[2]] : h[t.e[2]];
        const double x3 = t.neg[3] ? -h[t.e[3]] : h[t.e[3]];
        s = dadd(s, dadd(x2, x3));
    }
    if constexpr (t.two != 0) s = dmul(s, 2.0);
    return s;
}
template <int I = 0>
__device__ __forceinline__ void all_s(const double (&h)[16], double (&s)[16]) {
    if constexpr (I < 16) {
        s[I] = s_value<I>(h);
        all_s<I + 1>(h, s);
    }
}

// length-4 FMA chain in k order starting from an exact zero, i.e. what one DMMA computes; a term whose factor is a
// structural zero is skipped (adding an exact zero changes nothing)
struct Chain {
    double acc = 0.0;
    bool any = false;
    __device__ __forceinline__ void term(double x, double y) {
        acc = any ? dfma(x, y, acc) : dmul(x, y);
        any = true;
    }
};

// one iteration: normalised packed h -> unnormalised packed hn = R h R
template <bool UG>
__device__ __forceinline__ void iterate(const PauliParams& pp, const double (&h)[16], const double* __restrict__ fcol,
                                        int fstride, double (&hn)[16]) {
    double S[16];
    all_s(h, S);
    if (UG) S[0] = dadd(S[0], pp.epsp[0]);
    // weights and the qubit-2 map: N[al] = sum_be w (1 + s2 sigma_b) as (N00, N11, Re N01, Im N01)
    double N[6][4];
#pragma unroll
    for (int al = 0; al < 6; ++al) {
        const int a = al / 2 + 1;
        const bool neg_a = al & 1;
        const double u = neg_a ? dsub(S[0], S[a * 4]) : dadd(S[0], S[a * 4]);
        double v[4];
#pragma unroll
        for (int b = 1; b < 4; ++b) v[b] = neg_a ? dsub(S[b], S[a * 4 + b]) : dadd(S[b], S[a * 4 + b]);
        double w[6];
#pragma unroll
        for (int be = 0; be < 6; ++be) {
            const int b = be / 2 + 1;
            double q = (be & 1) ? dsub(u, v[b]) : dadd(u, v[b]);
            if (!UG) q = dadd(q, pp.epsp[al * 6 + be]);
            w[be] = q;
        }
#pragma unroll
        for (int be = 0; be < 6; ++be) w[be] = dmul(fcol[(al * 6 + be) * fstride], weight_recip(w[be]));
        const double y0 = dadd(dadd(dadd(w[0], w[1]), dadd(w[2], w[3])), dadd(w[4], w[5]));
        const double dz = dsub(w[4], w[5]);
        N[al][0] = dadd(y0, dz);
        N[al][1] = dsub(y0, dz);  // = y0 + (w5 - w4) bit for bit (w5 - w4 is the exact negative of w4 - w5)
        N[al][2] = dsub(w[0], w[1]);
        N[al][3] = dsub(w[3], w[2]);
    }
    // the qubit-1 map: R = sum_al (1 + s1 sigma_a) (x) N[al], as full real / imaginary 4x4 arrays
    double Rr[4][4], Ri[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double y0 = dadd(dadd(dadd(N[0][c], N[1][c]), dadd(N[2][c], N[3][c])), dadd(N[4][c], N[5][c]));
        const double dz = dsub(N[4][c], N[5][c]);
        const double d0 = dadd(y0, dz);  // block (0,0)
        const double d1 = dsub(y0, dz);  // block (1,1): y0 + (N5 - N4), the same bits
        if (c == 0) { Rr[0][0] = d0; Rr[2][2] = d1; }
        if (c == 1) { Rr[1][1] = d0; Rr[3][3] = d1; }
        if (c == 2) { Rr[0][1] = Rr[1][0] = d0; Rr[2][3] = Rr[3][2] = d1; }
        if (c == 3) { Ri[0][1] = d0; Ri[1][0] = -d0; Ri[2][3] = d1; Ri[3][2] = -d1; }
    }
#pragma unroll
    for (int x = 0; x < 4; ++x) Ri[x][x] = 0.0;
    {
        double y1[4], y2[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            y1[c] = dsub(N[0][c], N[1][c]);
            y2[c] = dsub(N[2][c], N[3][c]);
        }
        // block (0,1) = Y1 - i Y2 (rows 0,1; columns 2,3); block (1,0) is its adjoint
        const double e00r = y1[0], e00i = -y2[0];
        const double e11r = y1[1], e11i = -y2[1];
        const double e01r = dadd(y1[2], y2[3]), e01i = dsub(y1[3], y2[2]);
        const double e10r = dsub(y1[2], y2[3]), e10i = -dadd(y1[3], y2[2]);
        Rr[0][2] = Rr[2][0] = e00r; Ri[0][2] = e00i; Ri[2][0] = -e00i;
        Rr[1][3] = Rr[3][1] = e11r; Ri[1][3] = e11i; Ri[3][1] = -e11i;
        Rr[0][3] = Rr[3][0] = e01r; Ri[0][3] = e01i; Ri[3][0] = -e01i;
        Rr[1][2] = Rr[2][1] = e10r; Ri[1][2] = e10i; Ri[2][1] = -e10i;
    }
    // P = rho as full arrays (views of the packed registers)
    double Pr[4][4], Pi[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            Pr[a][b] = a <= b ? h[a * 4 + b] : h[b * 4 + a];
            Pi[a][b] = a == b ? 0.0 : (a < b ? h[b * 4 + a] : -h[a * 4 + b]);
        }
    // hn = (R P) R, upper triangle; row by row so that only one row of S = R P is live.  Every real and every imaginary
    // part is ONE chain: Re (X Y)[a][b] = sum_k Xr Yr, then sum_k (-Xi) Yi;  Im = sum_k Xr Yi, then sum_k Xi Yr -- the
    // order in which two DMMAs accumulating into one fragment produce them in the warp-per-sample mapping
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double Sr[4], Si[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            Chain re, im;
#pragma unroll
            for (int k = 0; k < 4; ++k) re.term(Rr[a][k], Pr[k][b]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (a != k && k != b) re.term(-Ri[a][k], Pi[k][b]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k != b) im.term(Rr[a][k], Pi[k][b]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (a != k) im.term(Ri[a][k], Pr[k][b]);
            Sr[b] = re.acc;
            Si[b] = im.acc;
        }
#pragma unroll
        for (int b = a; b < 4; ++b) {
            Chain re, im;
#pragma unroll
            for (int k = 0; k < 4; ++k) re.term(Sr[k], Rr[k][b]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k != b) re.term(-Si[k], Ri[k][b]);
            hn[a * 4 + b] = re.acc;
            if (a != b) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k != b) im.term(Sr[k], Ri[k][b]);
#pragma unroll
                for (int k = 0; k < 4; ++k) im.term(Si[k], Rr[k][b]);
                hn[b * 4 + a] = im.acc;
            }
        }
    }
}

// normalise, measure the step, advance: returns del = ||h' - h||_F^2
__device__ __forceinline__ double normalise_step(const double (&hn)[16], double (&h)[16]) {
    const double tr = dadd(dadd(hn[0], hn[5]), dadd(hn[10], hn[15]));
    const double inv = fast_recip(tr);
    double dd[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const double x = dmul(hn[e], inv);
        dd[e] = dsub(x, h[e]);
        h[e] = x;
    }
    // ||h' - h||_F^2 = sum_e m_e d_e^2, m = 1 on the diagonal and 2 for a pair (a,b), (b,a): four rows of four entries,
    // each a length-4 FMA chain of (m d) * d in index order, then the four row sums in order -- the arithmetic of the
    // two DMMAs that compute it in the warp-per-sample mapping
    double row[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        Chain c;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = 4 * r + k;
            c.term((e / 4 == e % 4) ? dd[e] : dmul(2.0, dd[e]), dd[e]);  // a product with 1.0 is the operand itself
        }
        row[r] = c.acc;
    }
    return dadd(dadd(dadd(row[0], row[1]), row[2]), row[3]);
}
}  // namespace single

// ------------------------------------------------------------------------------------------------
// warp per sample
// ------------------------------------------------------------------------------------------------
namespace wmode {
// per-warp shared-memory region, in doubles
constexpr int HN = 0, ZERO = 16, S = 18, W = 34, N = 70, X = 94, F = 126, SIZE = 168;

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// component (offset into a 4-vector (M00, M11, Re M01, Im M01)) and sign of the real / imaginary part of entry
// (r2, c2) of a Hermitian 2x2; comp = -1: structurally zero
__device__ __forceinline__ void herm2_component(int r2, int c2, int part, int& comp, unsigned& neg) {
    neg = 0;
    if (r2 == c2) {
        comp = part ? -1 : r2;
    } else {
        comp = part ? 3 : 2;
        if (part && r2 > c2) neg = 1;
    }
}

template <bool UG, bool TRACE>
__device__ void worker(const PauliParams& pp, const Pauli2Args& a, double* __restrict__ wb, const int lane) {
    const unsigned full = 0xffffffffu;
    // ---- per-lane constants of the graph ---------------------------------------------------------
    const int i16 = lane & 15;
    // S stage: lane i16 sums its (two or four) packed entries
    int so0 = ZERO, so1 = ZERO, so2 = ZERO, so3 = ZERO;
    unsigned sneg = 0;
    double sscale = 1.0;
    {
        int cnt = 0;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int c = pauli2::s_coef(i16, e);
            if (c != 0) {
                if (cnt == 0) so0 = HN + e;
                else if (cnt == 1) so1 = HN + e;
                else if (cnt == 2) so2 = HN + e;
                else so3 = HN + e;
                if (c < 0) sneg |= 1u << cnt;
                if (c == 2 || c == -2) sscale = 2.0;
                ++cnt;
            }
        }
    }
    // step norm: entry i16 weighs 1 (diagonal) or 2; after the first DMMA row sum r sits in lane 4 r + r / 2
    const double m_self = (i16 / 4 == i16 % 4) ? 1.0 : 2.0;
    const int row_src = 4 * (lane & 3) + ((lane & 3) >> 1);
    const bool row_odd = (lane >> 2) & 1;
    // B fragment of the first product: [Pr | Pi][k][n], k = lane & 3, n = lane >> 2
    int bo = ZERO;
    unsigned bneg = 0;
    {
        const int k = lane & 3, nn = (lane >> 2) & 3, part = lane >> 4;
        if (!part) bo = HN + (k <= nn ? k * 4 + nn : nn * 4 + k);
        else if (k < nn) bo = HN + nn * 4 + k;
        else if (k > nn) { bo = HN + k * 4 + nn; bneg = 1; }
    }
    // weights: slot k = lane (and 32 + lane for lane < 4)
    int qa0, qb0, qab0, qa1, qb1, qab1;
    unsigned q1neg0, q2neg0, q1neg1, q2neg1;
    {
        const int k0 = lane, k1 = 32 + (lane & 3);
        const int al0 = k0 / 6, be0 = k0 % 6, al1 = k1 / 6, be1 = k1 % 6;
        const int a0 = al0 / 2 + 1, b0 = be0 / 2 + 1, a1 = al1 / 2 + 1, b1 = be1 / 2 + 1;
        qa0 = S + 4 * a0; qb0 = S + b0; qab0 = S + 4 * a0 + b0; q1neg0 = al0 & 1; q2neg0 = be0 & 1;
        qa1 = S + 4 * a1; qb1 = S + b1; qab1 = S + 4 * a1 + b1; q1neg1 = al1 & 1; q2neg1 = be1 & 1;
    }
    const double eps0 = UG ? pp.epsp[0] : pp.epsp[lane];
    const double eps1 = UG ? 0.0 : pp.epsp[32 + (lane & 3)];
    // qubit-2 map: lane (al, c) for lane < 24
    const int n_al = (lane >> 2) < 6 ? (lane >> 2) : 5, n_c = lane & 3;
    const int nxo = W + n_al * 6;
    const int npo = nxo + (n_c == 0 ? 4 : n_c == 1 ? 5 : n_c == 2 ? 0 : 3);
    const int nmo = nxo + (n_c == 0 ? 5 : n_c == 1 ? 4 : n_c == 2 ? 1 : 2);
    // qubit-1 map: lane holds the real (lane < 16) or imaginary part of R[r][c], r = (lane >> 2) & 3, c = lane & 3:
    // val = ((n0 + n1) + (n2 + n3)) + (n4 + n5)  +  (n6 + n7), every n a signed entry of N or an exact zero
    int ro[8];
    unsigned rneg = 0;
    {
        const int r = (lane >> 2) & 3, c = lane & 3, part = lane >> 4;
        const int r1 = r >> 1, r2 = r & 1, c1 = c >> 1, c2 = c & 1;
#pragma unroll
        for (int j = 0; j < 8; ++j) ro[j] = ZERO;
        if (r1 == c1) {  // diagonal block: Y0[comp] +- (N[4] - N[5])[comp]
            int comp;
            unsigned ng;
            herm2_component(r2, c2, part, comp, ng);
            if (comp >= 0) {
#pragma unroll
                for (int j = 0; j < 6; ++j) ro[j] = N + j * 4 + comp;
                ro[6] = N + 4 * 4 + comp;
                ro[7] = N + 5 * 4 + comp;
                const unsigned sgn3 = r1 ? 1u : 0u;  // block (1,1): Y0 - Y3
                rneg = ng ? 0x3fu : 0u;
                rneg |= ((sgn3 ^ ng) << 6) | ((sgn3 ^ ng ^ 1u) << 7);
            }
        } else {
            // block (0,1): E[x][y] = Y1[x][y] - i Y2[x][y]; block (1,0)[x][y] = conj(E[y][x])
            const int x = r1 == 0 ? r2 : c2, y = r1 == 0 ? c2 : r2;
            const unsigned conj = r1 == 0 ? 0u : 1u;
            // Re E[x][y] = Re Y1[x][y] + Im Y2[x][y];  Im E[x][y] = Im Y1[x][y] - Re Y2[x][y]
            int c1comp, c2comp;
            unsigned n1, n2;
            herm2_component(x, y, part, c1comp, n1);        // Y1 part
            herm2_component(x, y, part ^ 1, c2comp, n2);    // Y2: the other part
            if (part) n2 ^= 1u;                             // - Re Y2
            if (part && conj) { n1 ^= 1u; n2 ^= 1u; }
            if (c1comp >= 0) {
                ro[0] = N + 0 * 4 + c1comp;
                ro[1] = N + 1 * 4 + c1comp;
                rneg |= (n1 << 0) | ((n1 ^ 1u) << 1);
            }
            if (c2comp >= 0) {
                ro[6] = N + 2 * 4 + c2comp;
                ro[7] = N + 3 * 4 + c2comp;
                rneg |= (n2 << 6) | ((n2 ^ 1u) << 7);
            }
        }
    }
    // where the combined products go
    const int xo = X + ((lane & 2) << 3) + (lane >> 2) * 4 + 2 * (lane & 1);  // lanes < 16: (Sr | Si) pairs
    int ho0 = -1, ho1 = -1;  // packed destinations of this lane's two entries of the new state (lanes < 16)
    {
        const int i = lane >> 2, jh = lane & 3;
        if (lane < 16) {
            if (jh < 2) {  // real parts, columns 2 jh, 2 jh + 1
                if (i <= 2 * jh) ho0 = HN + i * 4 + 2 * jh;
                if (i <= 2 * jh + 1) ho1 = HN + i * 4 + 2 * jh + 1;
            } else {       // imaginary parts, columns 2 (jh - 2), +1: strict upper triangle only
                const int j0 = 2 * (jh - 2);
                if (i < j0) ho0 = HN + j0 * 4 + i;
                if (i < j0 + 1) ho1 = HN + (j0 + 1) * 4 + i;
            }
        }
    }
    // result layout: lane l holds output double l of the complex 4x4 (re, im interleaved)
    int out_src;
    unsigned out_neg = 0;
    bool out_zero = false;
    {
        const int m = lane >> 1, ra = m >> 2, rb = m & 3, part = lane & 1;
        if (!part) out_src = ra <= rb ? ra * 4 + rb : rb * 4 + ra;
        else if (ra == rb) { out_src = 0; out_zero = true; }
        else if (ra < rb) out_src = rb * 4 + ra;
        else { out_src = ra * 4 + rb; out_neg = 1; }
    }
    const unsigned hi_half = lane < 16 ? 0u : 1u;   // second product: the imaginary half takes -R
    const double tol2 = a.tol * a.tol;
    const int K = pp.K;
    if (lane == 0) wb[ZERO] = 0.0;
    // profiling stamps and counters exist in the TRACE instantiation only (tools/pauli2_trace.py): six 64-bit counters in
    // the loops cost the thread-per-sample mapping registers it does not have
    long long* const tr_w = (TRACE && a.trace) ? a.trace + 16 * ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) : nullptr;
    long long* const trace_s = TRACE ? a.trace_s : nullptr;
    long long tr_samples = 0, tr_its = 0, tr_wait = 0;
    if (tr_w && lane == 0) tr_w[8] = (long long)globaltimer_ns();

    for (;;) {
        // ---- next sample: a handed-over one (ticket order), or a fresh one in direct mode -----------------
        long long b = -1;
        int it = 0;
        int from_park = 0;
        if (lane == 0) {
            if (a.direct) {
                const unsigned idx = atomicAdd(&a.ctrl[0], 1u);
                if (idx < (unsigned)a.B) b = a.order ? (long long)__ldg(a.order + idx) : (long long)idx;
            } else {
                const unsigned ticket = atomicAdd(&a.ctrl[2], 1u);
                unsigned ns = 64;
                for (;;) {
                    unsigned r = 0;
                    if (ticket < a.park_cap)
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(a.park_ready + ticket) : "memory");
                    if (ticket < a.park_cap && r == a.epoch) {
                        b = a.park_b[ticket];
                        it = a.park_it[ticket];
                        from_park = (int)ticket + 1;
                        break;
                    }
                    unsigned done;
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(done) : "l"(a.ctrl + 3) : "memory");
                    if (done >= (unsigned)a.B) break;
                    __nanosleep(ns);
                    if (ns < 2048) ns <<= 1;
                }
            }
        }
        b = __shfl_sync(full, b, 0);
        if (b < 0) {
            if (tr_w && lane == 0) {
                tr_w[9] = (long long)globaltimer_ns();
                tr_w[10] = tr_samples;
                tr_w[11] = tr_its;
                tr_w[12] = tr_wait;
            }
            return;
        }
        it = __shfl_sync(full, it, 0);
        const int tr_it0 = it;
        const long long tr_t0 = tr_w ? (long long)globaltimer_ns() : 0;
        if (trace_s && lane == 0) trace_s[4 * b + (from_park ? 2 : 0)] = tr_t0;
        from_park = __shfl_sync(full, from_park, 0);
        // ---- frequencies by slot, start state -----------------------------------------------------------
        __syncwarp();
        wb[F + lane] = 0.0;
        if (lane < 4) wb[F + 32 + lane] = 0.0;
        const int32_t* crow = a.counts + b * K;
        const int c0 = lane < K ? crow[lane] : 0;
        const int c1 = (32 + lane) < K ? crow[32 + lane] : 0;
        long long tot = (long long)c0 + c1;  // 64-bit, like the thread-per-sample refill: P * shots may pass 2^31
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(full, tot, o);
        const FreqDiv freq((double)tot);
        __syncwarp();
        if (lane < K) wb[F + pp.slot_of_col[lane]] = freq((double)c0);
        if (32 + lane < K) wb[F + pp.slot_of_col[32 + lane]] = freq((double)c1);
        if (lane < 16) {
            double v;
            if (from_park) {
                v = __ldcg(a.park_h + (size_t)(from_park - 1) * 16 + lane);
            } else if (a.rho0) {
                const int ra = lane >> 2, rb = lane & 3;
                v = ra <= rb ? a.rho0[b * 32 + 2 * (ra * 4 + rb)] : a.rho0[b * 32 + 2 * (rb * 4 + ra) + 1];
            } else {
                v = (lane >> 2) == (lane & 3) ? 0.25 : 0.0;
            }
            wb[HN + lane] = v;
        }
        __syncwarp();

        double hprev = wb[HN + i16];
        if (a.max_iter > 0) {
            // `first`: the state in HN is already normalised (start state or handed-over iterate): scale by one,
            // no step to measure.  Branch-free, so that the step-norm butterfly (four dependent shuffles) can be
            // spread over the following phases instead of stalling this one.
            bool first = true;
            double del = 0.0;
            for (;;) {
                // -- P0: normalise, step norm, Pauli sums ---------------------------------------------------
                const double tr = dadd(dadd(wb[HN + 0], wb[HN + 5]), dadd(wb[HN + 10], wb[HN + 15]));
                const double inv = first ? 1.0 : fast_recip(tr);
                const double x0 = dmul(flip(wb[so0], sneg & 1u), inv);
                const double x1 = dmul(flip(wb[so1], (sneg >> 1) & 1u), inv);
                const double x2 = dmul(flip(wb[so2], (sneg >> 2) & 1u), inv);
                const double x3 = dmul(flip(wb[so3], (sneg >> 3) & 1u), inv);
                double s = dmul(dadd(dadd(x0, x1), dadd(x2, x3)), sscale);
                if (UG && i16 == 0) s = dadd(s, eps0);
                wb[S + i16] = s;
                const double bfrag = dmul(flip(wb[bo], bneg), inv);
                const double xe = dmul(wb[HN + i16], inv);
                const double dstep = dsub(xe, hprev);
                // step norm on the tensor pipe: rows of A = (m d) in fours, columns of B = d in fours, so that the
                // diagonal of the product holds the four row sums as FMA chains in index order ...
                double sn0 = 0.0, sn1 = 0.0;
                dmma(sn0, sn1, dmul(m_self, dstep), dstep);
                const double row_sum = row_odd ? sn1 : sn0;  // D[r][r] in lane 4 r + r / 2: element r % 2
                hprev = xe;
                it += first ? 0 : 1;
                __syncwarp();
                // -- P1: weights ------------------------------------------------------------------------------
                {
                    const double s00 = wb[S];
                    const double u = dadd(s00, flip(wb[qa0], q1neg0));
                    const double v = dadd(wb[qb0], flip(wb[qab0], q1neg0));
                    double q = dadd(u, flip(v, q2neg0));
                    if (!UG) q = dadd(q, eps0);
                    const double w0 = dmul(wb[F + lane], weight_recip(q));
                    const double u1 = dadd(s00, flip(wb[qa1], q1neg1));
                    const double v1 = dadd(wb[qb1], flip(wb[qab1], q1neg1));
                    double q1 = dadd(u1, flip(v1, q2neg1));
                    if (!UG) q1 = dadd(q1, eps1);
                    const double w1 = dmul(wb[F + 32 + (lane & 3)], weight_recip(q1));
                    wb[W + lane] = w0;
                    if (lane < 4) wb[W + 32 + lane] = w1;
                }
                // ... and a second product with a matrix of ones adds the four row sums in order, in every lane
                const double row_k = __shfl_sync(full, row_sum, row_src);
                __syncwarp();
                // -- P2: qubit-2 map ----------------------------------------------------------------------------
                {
                    const double2 xa = *reinterpret_cast<const double2*>(wb + nxo);
                    const double2 xb = *reinterpret_cast<const double2*>(wb + nxo + 2);
                    const double2 xc = *reinterpret_cast<const double2*>(wb + nxo + 4);
                    const double y0 = dadd(dadd(dadd(xa.x, xa.y), dadd(xb.x, xb.y)), dadd(xc.x, xc.y));
                    const double dd = dsub(wb[npo], wb[nmo]);
                    const double nv = n_c < 2 ? dadd(y0, dd) : dd;
                    if (lane < 24) wb[N + lane] = nv;
                }
                double pstep = 0.0, pstep_dup = 0.0;
                dmma(pstep, pstep_dup, row_k, 1.0);
                __syncwarp();
                // -- P3: qubit-1 map -> fragment of R; first product ------------------------------------------
                double val;
                {
                    const double n0 = flip(wb[ro[0]], rneg & 1u), n1 = flip(wb[ro[1]], (rneg >> 1) & 1u);
                    const double n2 = flip(wb[ro[2]], (rneg >> 2) & 1u), n3 = flip(wb[ro[3]], (rneg >> 3) & 1u);
                    const double n4 = flip(wb[ro[4]], (rneg >> 4) & 1u), n5 = flip(wb[ro[5]], (rneg >> 5) & 1u);
                    const double n6 = flip(wb[ro[6]], (rneg >> 6) & 1u), n7 = flip(wb[ro[7]], (rneg >> 7) & 1u);
                    val = dadd(dadd(dadd(dadd(n0, n1), dadd(n2, n3)), dadd(n4, n5)), dadd(n6, n7));
                }
                del = first ? 1e300 : pstep;
                // every lane holds the same bits of del and it (lanes 16-31 mirror lanes 0-15): a uniform branch
                if (del < tol2 || it >= a.max_iter) break;
                first = false;
                // S = R P: [Rr; Ri] x [Pr | Pi], then [-Ri; -Rr] x [Pi | -Pr] into the same fragment: rows 0-3 hold
                // Re S = Rr Pr - Ri Pi in columns 0-3 and Im S = Rr Pi + Ri Pr in columns 4-7, each as ONE chain
                const double valx = flip(__shfl_xor_sync(full, val, 16), 1u);  // lanes < 16: -Ri, the others: -Rr
                {
                    const double bfragx = flip(__shfl_xor_sync(full, bfrag, 16), hi_half);
                    double d0 = 0.0, d1 = 0.0;
                    dmma(d0, d1, val, bfrag);
                    dmma(d0, d1, valx, bfragx);
                    if (lane < 16) *reinterpret_cast<double2*>(wb + xo) = make_double2(d0, d1);
                }
                __syncwarp();
                // -- P4: second product -> unnormalised new state ---------------------------------------------
                // T = S R: [Sr; Si] x [Rr | Ri], then [-Si; -Sr] x [Ri | -Rr]; in the B layout (k = column of R, by
                // Hermiticity) the second right operand is valx itself
                {
                    const double a2 = wb[X + lane];
                    const double a2x = flip(wb[X + (lane ^ 16)], 1u);
                    const double b2 = flip(val, hi_half);
                    double d0 = 0.0, d1 = 0.0;
                    dmma(d0, d1, a2, b2);
                    dmma(d0, d1, a2x, valx);
                    if (ho0 >= 0) wb[ho0] = d0;
                    if (ho1 >= 0) wb[ho1] = d1;
                }
                __syncwarp();
            }
        }
        // ---- write back: lane e < 16 holds packed entry e of the result in hprev ------------------------
        {
            double o = __shfl_sync(full, hprev, out_src);
            o = out_zero ? 0.0 : flip(o, out_neg);
            if (a.rho) a.rho[b * 32 + lane] = o;
            if (a.hs_dist) {
                __syncwarp();
                if (lane < 16) wb[HN + lane] = hprev;
                __syncwarp();
                if (lane == 0) {
                    double h[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) h[e] = wb[HN + e];
                    a.hs_dist[b] = hs_distance_packed(h, a.hs_ref);
                }
            }
            if (lane == 0) {
                if (a.iters) a.iters[b] = it;
                if (!a.direct) atomicAdd(&a.ctrl[3], 1u);
            }
            if (trace_s && lane == 0) trace_s[4 * b + 3] = (long long)globaltimer_ns();
            tr_samples += 1;
            tr_its += it - tr_it0;
            if (tr_w) tr_wait += (long long)globaltimer_ns() - tr_t0;  // busy time, in fact
        }
    }
}
}  // namespace wmode

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// frequencies of sample b into the lane's shared-memory column (slot-major, thread-minor)
__device__ __forceinline__ void load_frequencies(const PauliParams& pp, const int32_t* __restrict__ c, int K,
                                                 double* __restrict__ fcol) {
    // all K count loads are issued together (independent, predicated), then normalised
    int cc[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) cc[k] = (k < K) ? c[k] : 0;
    long long tot = 0;
#pragma unroll
    for (int k = 0; k < 36; ++k) tot += cc[k];
    const FreqDiv freq((double)tot);
#pragma unroll
    for (int k = 0; k < 36; ++k)
        if (k < K) fcol[pp.slot_of_col[k] * kPauliThreadsSingle] = freq((double)cc[k]);
}

template <bool UG, bool TRACE>  // UG: all used slots share one 1e-10/c: fold it into S_00 instead of 36 additions
__global__ void __launch_bounds__(32 * kPauliCtaWarps, 1)
k_mle_rrr_pauli2(const __grid_constant__ PauliParams pp, const __grid_constant__ Pauli2Args a) {
    constexpr int D = 16;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* wbase = sm + (size_t)warp * wmode::SIZE;                     // W region of this warp
    double* fs = sm + (size_t)(blockDim.x >> 5) * wmode::SIZE;            // [36][256], thread-per-sample lanes only
    // Packing of the tail (a.merge): once the queue is empty the live samples of a CTA's thread-per-sample warps thin
    // out; a warp whose samples fit into the free lanes of the others pushes them (state, frequencies, index, age)
    // onto a shared-memory stack, the others pop them into their free lanes, and the emptied warp becomes a W worker.
    // live[w] (cx[0..11]): live lanes of warp w as last published, 32 before it saw the queue empty, -1 once it left;
    // cx[kCxTop]: entries on the stack; cx[kCxLock]: lock (all stack / live[-1] transitions happen under it).
    double* pool = fs + 36 * kPauliThreadsSingle;                          // [kPoolWords][32], entry-minor
    volatile int* cx = reinterpret_cast<volatile int*>(pool + kPoolWords * 32);
    if (a.merge) {
        if (tid < kPauliWarps) cx[tid] = tid < a.single_warps ? 32 : -1;
        if (tid == kCxTop || tid == kCxLock) cx[tid] = 0;
        __syncthreads();
    }
    if (warp >= a.single_warps) {
        wmode::worker<UG, TRACE>(pp, a, wbase, lane);
        return;
    }
    const int K = pp.K;
    for (int sl = 0; sl < 36; ++sl) fs[sl * kPauliThreadsSingle + tid] = 0.0;  // own column only: no barrier needed

    double h[D];
    int it = 0, my_age = a.park_age;
    long long b = -1;
    bool alive = true;  // false once the queue ran dry for this lane
    double del_ref = 0.0;   // step norm 32 iterations ago
    bool plateau = false;
    unsigned tick = 0;
    bool dissolved = false;  // this warp pushed its samples onto the packing stack
    long long* const tr_s = (TRACE && a.trace) ? a.trace + 16 * ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) : nullptr;
    long long* const trace_s = TRACE ? a.trace_s : nullptr;
    long long tr_wi = 0, tr_li = 0, tr_wi_tail = 0, tr_li_tail = 0, tr_parked = 0, tr_adopted = 0;
    bool tr_seen_drain = false;
    if (tr_s && lane == 0) tr_s[0] = (long long)globaltimer_ns();
    const double tol2 = a.tol * a.tol;
    const bool hand_over = a.single_warps < (int)(blockDim.x >> 5) || a.park_live > 0 || a.park_age < 0x7fffffff;

    // write the results of finished samples whose lanes satisfy `now` (warp-uniform call, divergent body)
    auto write_back = [&](bool now) {
        const bool wb = now && b < -1;
        const unsigned wbm = __ballot_sync(0xffffffffu, wb);
        if (wbm == 0) return;
        if (wb) {
            const long long ob = -2 - b;
            if (a.rho) {
                double2* out = reinterpret_cast<double2*>(a.rho) + ob * D;
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        double2 z;
                        z.x = x <= y ? h[x * 4 + y] : h[y * 4 + x];
                        z.y = x == y ? 0.0 : (x < y ? h[y * 4 + x] : -h[x * 4 + y]);
                        out[x * 4 + y] = z;
                    }
            }
            if (a.hs_dist) a.hs_dist[ob] = hs_distance_packed(h, a.hs_ref);
            if (a.iters) a.iters[ob] = it;
            if (trace_s) trace_s[4 * ob + 3] = (long long)globaltimer_ns();
            b = -1;
        }
        if (hand_over && lane == 0) atomicAdd(&a.ctrl[3], (unsigned)__popc(wbm));
    };

    while (true) {
        // ---- refill lanes without work -------------------------------------------------------
        const bool want = alive && b < 0;
        const unsigned need = __ballot_sync(0xffffffffu, want);
        const bool go = need && (__popc(need) >= a.refill_min || __ballot_sync(0xffffffffu, b >= 0) == 0);
        // results of finished samples (b = -2 - index, state still in h) are written together with the refill, or at
        // once by lanes that will not get another sample
        write_back(go || !alive);
        if (go) {
            unsigned base = 0;
            const int leader = __ffs(need) - 1;
            if (lane == leader) base = atomicAdd(&a.ctrl[0], (unsigned)__popc(need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want) {
                const long long nb = (long long)base + __popc(need & ((1u << lane) - 1u));
                if (nb < a.B) {
                    b = a.order ? (long long)__ldg(a.order + nb) : nb;
                    my_age = a.park_age;
                    if (a.order) {
                        my_age = min(a.park_age, a.park_age_lo + (int)((float)nb * a.park_age_slope));
                        if (nb > a.park_q2)
                            my_age = max(a.park_age_end, my_age - (int)((float)(nb - a.park_q2) * a.park_age_slope2));
                    }
                    if (trace_s) trace_s[4 * b] = (long long)globaltimer_ns();
                    it = 0;
                    plateau = false;
                    load_frequencies(pp, a.counts + b * K, K, fs + tid);
                    if (a.rho0) {
                        const double2* r0 = reinterpret_cast<const double2*>(a.rho0) + b * D;
#pragma unroll
                        for (int x = 0; x < 4; ++x)
#pragma unroll
                            for (int y = x; y < 4; ++y) {
                                const double2 z = r0[x * 4 + y];
                                h[x * 4 + y] = z.x;
                                if (x != y) h[y * 4 + x] = z.y;
                            }
                    } else {
#pragma unroll
                        for (int e = 0; e < D; ++e) h[e] = (e / 4 == e % 4) ? 0.25 : 0.0;
                    }
                } else {
                    alive = false;
                }
            }
        }
        unsigned active = __ballot_sync(0xffffffffu, b >= 0);
        const bool drained = __any_sync(0xffffffffu, !alive);
        if (active == 0 && !drained) continue;  // cannot happen (a lane without work either refilled or saw the end)
        if (a.merge && drained) {
            write_back(true);  // free lanes are about to be overwritten
            const unsigned full = 0xffffffffu;
            auto lock = [&]() {
                if (lane == 0)
                    while (atomicCAS(const_cast<int*>(cx + kCxLock), 0, 1) != 0) {}
                __syncwarp();
                __threadfence_block();
            };
            auto unlock = [&]() {
                __threadfence_block();
                __syncwarp();
                if (lane == 0) atomicExch(const_cast<int*>(cx + kCxLock), 0);
            };
            auto pop = [&]() {  // under the lock: fill free lanes from the stack
                const int top = cx[kCxTop];
                const unsigned freem = ~active;
                int n = __popc(freem);
                if (n > top) n = top;
                const int rank = __popc(freem & ((1u << lane) - 1u));
                if (b < 0 && rank < n) {
                    const double* src = pool + (top - 1 - rank);
#pragma unroll
                    for (int e = 0; e < D; ++e) h[e] = src[e * 32];
#pragma unroll
                    for (int sl = 0; sl < 36; ++sl) fs[sl * kPauliThreadsSingle + tid] = src[(16 + sl) * 32];
                    b = __double_as_longlong(src[52 * 32]);
                    it = (int)__double_as_longlong(src[53 * 32]);
                    my_age = a.park_age;
                    plateau = false;
                    del_ref = 1e300;
                }
                active = __ballot_sync(full, b >= 0);
                tr_adopted += n;
                if (lane == 0) {
                    cx[kCxTop] = top - n;
                    cx[warp] = __popc(active);
                }
            };
            if (active == 0) {  // leaving: nothing may stay on the stack without a warp to take it
                lock();
                if (!dissolved && cx[kCxTop] > 0) {
                    pop();
                    unlock();
                    continue;
                }
                if (lane == 0) cx[warp] = -1;
                unlock();
                break;
            }
            const int live = __popc(active);
            if (live < 32) {
                if (lane == 0) cx[warp] = live;
                if (cx[kCxTop] > 0) {
                    lock();
                    pop();
                    unlock();
                } else {
                    const int lv = lane < kPauliWarps ? cx[lane] : -1;
                    const int room = __reduce_add_sync(full, (lane != warp && lv >= 0) ? 32 - lv : 0);
                    if (room >= live + a.merge - 1) {
                        lock();
                        const int top = cx[kCxTop];
                        const int lv2 = lane < kPauliWarps ? cx[lane] : -1;
                        const int room2 = __reduce_add_sync(full, (lane != warp && lv2 >= 0) ? 32 - lv2 : 0) - top;
                        if (room2 >= live && top + live <= 32) {
                            const int rank = __popc(active & ((1u << lane) - 1u));
                            if (b >= 0) {
                                double* dst = pool + (top + rank);
#pragma unroll
                                for (int e = 0; e < D; ++e) dst[e * 32] = h[e];
#pragma unroll
                                for (int sl = 0; sl < 36; ++sl) dst[(16 + sl) * 32] = fs[sl * kPauliThreadsSingle + tid];
                                dst[52 * 32] = __longlong_as_double(b);
                                dst[53 * 32] = __longlong_as_double((long long)it);
                                b = -1;
                            }
                            if (lane == 0) {
                                cx[kCxTop] = top + live;
                                cx[warp] = -1;
                            }
                            dissolved = true;
                            tr_parked += live;  // (trace: counted with the handed-over samples)
                            active = 0;
                        }
                        unlock();
                        if (dissolved) break;
                    }
                }
            }
        } else if (active == 0) {
            break;
        }
        if (tr_s && drained && !tr_seen_drain) {
            tr_seen_drain = true;
            if (lane == 0) {
                tr_s[1] = (long long)globaltimer_ns();
                tr_s[7] = __popc(active);
            }
        }
        // ---- hand long-running samples to the W workers ---------------------------------------------------
        if (hand_over) {
            const bool tail = drained && a.tail_poll > 0;
            auto hand_over_entry = [&](unsigned slot) {
                if (slot >= a.park_cap) return;  // list full: the sample stays in this lane
                double2* dst = reinterpret_cast<double2*>(a.park_h + (size_t)slot * 16);
#pragma unroll
                for (int e = 0; e < 8; ++e) __stcg(dst + e, make_double2(h[2 * e], h[2 * e + 1]));
                a.park_b[slot] = b;
                a.park_it[slot] = it;
                if (trace_s) trace_s[4 * b + 1] = (long long)globaltimer_ns();
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.park_ready + slot), "r"(a.epoch) : "memory");
                b = -1;
            };
            // bulk: by age (the W workers are a fixed share of the machine).  Tail: by demand, below.
            bool park = b >= 0 && ((!tail && it >= my_age) || plateau);
            if (drained && __popc(active) <= a.park_live) park = b >= 0;
            if (tail && ((++tick) & (a.tail_poll - 1)) == 0 && __popc(active) > a.park_live) {
                write_back(a.adopt > 0);  // free lanes may adopt entries
                // the queue is empty: look at the hand-over list.  waiting > 0: W workers without an entry -> give
                // them this warp's oldest sample; waiting < 0: entries without a worker -> free lanes adopt them.
                unsigned r = 0, t = 0;
                if (lane == 0) {
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(a.ctrl + 1) : "memory");
                    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(t) : "l"(a.ctrl + 2) : "memory");
                }
                r = __shfl_sync(0xffffffffu, r, 0);
                t = __shfl_sync(0xffffffffu, t, 0);
                if (r > a.park_cap) r = a.park_cap;
                if (t > a.park_cap) t = a.park_cap;
                const int waiting = (int)(t - r);
                if (waiting > 0) {
                    const int age = (b >= 0 && it >= a.tail_age) ? it : -1;
                    const int oldest = __reduce_max_sync(0xffffffffu, age);
                    if (oldest >= 0) {
                        const int who = __ffs(__ballot_sync(0xffffffffu, age == oldest)) - 1;
                        if (lane == who) park = true;
                    }
                } else if (a.adopt > 0 && -waiting >= a.adopt && ~active != 0u) {
                    const unsigned freem = ~active;
                    unsigned n = (unsigned)__popc(freem);
                    if (n > (unsigned)(-waiting)) n = (unsigned)(-waiting);
                    unsigned got = 0;
                    if (lane == 0) got = atomicCAS(&a.ctrl[2], t, t + n);
                    got = __shfl_sync(0xffffffffu, got, 0);
                    if (got == t && b < 0 && (unsigned)__popc(freem & ((1u << lane) - 1u)) < n) {
                        const unsigned ticket = t + (unsigned)__popc(freem & ((1u << lane) - 1u));
                        unsigned ready;
                        do {  // reserved before we looked, so its writer is on the way
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(ready) : "l"(a.park_ready + ticket) : "memory");
                        } while (ready != a.epoch);
                        b = a.park_b[ticket];
                        it = a.park_it[ticket];
                        my_age = a.park_age;
                        plateau = false;
                        del_ref = 1e300;
                        const double2* src = reinterpret_cast<const double2*>(a.park_h + (size_t)ticket * 16);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const double2 z = __ldcg(src + e);
                            h[2 * e] = z.x;
                            h[2 * e + 1] = z.y;
                        }
                        load_frequencies(pp, a.counts + b * K, K, fs + tid);
                    }
                    tr_adopted += __popc(__ballot_sync(0xffffffffu, b >= 0)) - __popc(active);
                    active = __ballot_sync(0xffffffffu, b >= 0);
                }
            }
            const unsigned pm = __ballot_sync(0xffffffffu, park);
            tr_parked += __popc(pm);
            if (pm) {
                unsigned slot0 = 0;
                const int leader = __ffs(pm) - 1;
                if (lane == leader) slot0 = atomicAdd(&a.ctrl[1], (unsigned)__popc(pm));
                slot0 = __shfl_sync(0xffffffffu, slot0, leader);
                if (park) hand_over_entry(slot0 + __popc(pm & ((1u << lane) - 1u)));
                active = __ballot_sync(0xffffffffu, b >= 0);
                if (active == 0) continue;  // everything handed over: refill or leave
            }
        }
        // ---- one R.rho.R iteration (lanes without a sample are predicated off) ---------------
        tr_wi += 1;
        tr_li += __popc(active);
        if (drained) {
            tr_wi_tail += 1;
            tr_li_tail += __popc(active);
        }
        bool finished = false;
        if (b >= 0) {
            if (a.max_iter <= 0) {
                finished = true;
            } else {
                double hn[D];
                single::iterate<UG>(pp, h, fs + tid, kPauliThreadsSingle, hn);
                const double del = single::normalise_step(hn, h);
                ++it;
                finished = (del < tol2) || (it >= a.max_iter);
                if ((it & 31) == 0) {
                    plateau = a.park_plateau > 0 && it >= a.park_plateau && del >= del_ref;
                    del_ref = del;
                }
            }
        }
        if (finished) b = -2 - b;  // result pending: written by write_back()
    }
    if (tr_s && lane == 0) {
        tr_s[2] = (long long)globaltimer_ns();
        tr_s[3] = tr_wi;
        tr_s[4] = tr_li;
        tr_s[5] = tr_wi_tail;
        tr_s[6] = tr_li_tail;
        tr_s[13] = tr_parked;
        tr_s[14] = tr_adopted;
    }
    // out of single-lane work: serve the hand-over list until every sample of the launch is finished
    if (hand_over) wmode::worker<UG, TRACE>(pp, a, wbase, lane);
}

// Warp-per-sample workers only (small batches): few registers, so 32 warps per SM hide each other's latency.
template <bool UG, int THREADS, bool TRACE>
__global__ void __launch_bounds__(THREADS, 1)
k_mle_rrr_pauli2_w(const __grid_constant__ PauliParams pp, const __grid_constant__ Pauli2Args a) {
    extern __shared__ __align__(16) double sm[];
    wmode::worker<UG, TRACE>(pp, a, sm + (size_t)(threadIdx.x >> 5) * wmode::SIZE, threadIdx.x & 31);
}

// Recognise a two-qubit Pauli-axis POVM from the Bloch-basis table A [K][16] (host copy).
bool detect_pauli2(const double* A, int K, PauliParams* pp) {
    if (K < 1 || K > 36) return false;
    bool used[36] = {false};
    for (int sl = 0; sl < 36; ++sl) pp->epsp[sl] = 1.0;
    for (int k = 0; k < 36; ++k) pp->slot_of_col[k] = 0;
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

bool plan_is_pauli2(const qpb_state_plan* plan) {
    if (plan->n != 2 || !plan->A_host) return false;
    PauliParams pp;
    return detect_pauli2(plan->A_host, plan->K, &pp);
}

int launch_mle_pauli2(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                      double tol, double* rho, int32_t* iters, cudaStream_t st, const double* hs_ref, double* hs_dist,
                      bool* hs_done, int hs_store_rho, const int* order) {
    static std::atomic<unsigned int> epoch_counter{1};
    PauliParams pp;
    if (!detect_pauli2(plan->A_host, plan->K, &pp)) return QPB_ERR_UNSUPPORTED;
    QPB_REQUIRE((long long)B < (1ll << 31) - 65536, "batch too large");
    const int sms = num_sms();
    Pauli2Args a;
    a.B = B;
    a.counts = counts;
    a.rho0 = rho0;
    a.order = order;
    a.max_iter = max_iter;
    a.tol = tol;
    a.iters = iters;
    const bool fuse = hs_ref && hs_dist && hs_done && !option(QPB_OPT_NO_HS_FUSION);
    a.rho = (fuse && hs_store_rho == 0) ? nullptr : rho;  // with the distance fused and no caller for the states, they are not written
    a.hs_ref = fuse ? hs_ref : nullptr;
    a.hs_dist = fuse ? hs_dist : nullptr;

    // ---- mapping policy ---------------------------------------------------------------------------------
    // direct (W workers only) while the batch leaves most single-lane lanes empty: its time is the longest
    // chain at ~0.27 us per iteration instead of 1.3 us.  Otherwise thread-per-sample warps plus hand-over.
    const int lanes_opt = option(QPB_OPT_MLE_LANES);
    const bool no_merge = option(QPB_OPT_NO_TAIL_MERGE) != 0 || lanes_opt == 1;
    // direct: at most one sample per W warp and SM-filling (measured: B = 1000 0.45 ms against 0.96 ms hybrid and
    // 1.31 ms thread-per-sample; from B ~ 5000 on the hybrid wins)
    bool direct = lanes_opt == 32 || (lanes_opt == 0 && !option(QPB_OPT_NO_TAIL_MERGE) && (long long)B <= (long long)sms * 24);
    int w_warps = option(QPB_OPT_MLE_W_WARPS);
    if (w_warps == 0) w_warps = kPauliCtaWarps - 8;
    if (w_warps < 0) w_warps = 0;
    if (w_warps > kPauliCtaWarps - 8) w_warps = kPauliCtaWarps - 8;
    int threads, blocks;
    if (direct) {
        a.single_warps = 0;
        a.direct = 1;
        a.park_age = 0x7fffffff;
        a.park_age_lo = 0x7fffffff;
        a.park_age_slope = a.park_age_slope2 = 0.f;
        a.park_q2 = 0x7fffffff;
        a.park_age_end = 0x7fffffff;
        a.park_live = 0;
        a.park_plateau = 0;
        a.tail_poll = a.tail_age = a.adopt = a.merge = 0;
        a.refill_min = 1;
        // warps per CTA: enough that every SM has work, at most 32 (profiling: MLE_BLOCKS_PER_SM overrides)
        int dw = option(QPB_OPT_MLE_BLOCKS_PER_SM) > 0 ? option(QPB_OPT_MLE_BLOCKS_PER_SM) : (int)(((long long)B + sms - 1) / sms);
        const int dw_max = option(QPB_OPT_MLE_W_WARPS) > 4 ? option(QPB_OPT_MLE_W_WARPS) : 16;  // profiling: 16 | 24 | 32
        if (dw > dw_max) dw = dw_max;
        if (dw < 1) dw = 1;
        threads = 32 * dw;
        blocks = (int)(((long long)B + dw - 1) / dw);
        if (blocks > sms) blocks = sms;
    } else {
        a.direct = 0;
        // every SM gets a CTA; only as many thread-per-sample warps as the batch can fill (they run faster alone),
        // the other warps of the 12 serve the hand-over list from the start
        blocks = sms;
        // thread-per-sample warps per CTA (of 12 warps in all): 8 = two per scheduler (default), up to 12
        int max_sw = option(QPB_OPT_MLE_SINGLE_WARPS) > 0 ? option(QPB_OPT_MLE_SINGLE_WARPS) : 8;
        if (max_sw > kPauliWarps) max_sw = kPauliWarps;
        const long long need = ((long long)B + max_sw * 32 - 1) / (max_sw * 32);
        if (blocks > need && no_merge) blocks = (int)need;
        int sw = (int)(((long long)B + (long long)blocks * 32 - 1) / ((long long)blocks * 32));
        if (sw > max_sw) sw = max_sw;
        if (sw < 1) sw = 1;
        a.single_warps = sw;
        if (no_merge) {
            a.park_age = 0x7fffffff;
            a.park_age_lo = 0x7fffffff;
            a.park_age_slope = a.park_age_slope2 = 0.f;
            a.park_q2 = 0x7fffffff;
            a.park_age_end = 0x7fffffff;
            a.park_live = 0;
            a.park_plateau = 0;
            a.tail_poll = a.tail_age = a.adopt = a.merge = 0;
            a.refill_min = 1;
            w_warps = 0;
            a.single_warps = max_sw;
        } else {
            const int pl = option(QPB_OPT_MLE_PARK_PLATEAU);
            a.park_plateau = pl > 0 ? pl : 0;
            // measured on B200 (tools/pauli2_sweep_d.py, C2 workload): the fuller the thread-per-sample lanes, the
            // less W capacity is left, so the hand-over age rises with the batch: 200 / 300 / 450 iterations
            const long long lanes = (long long)sms * max_sw * 32;
            // (tools/pauli2_sweep_n.py: below half a wave of lanes a late general age with an early start for the first
            // queue positions beats the early general age: 12 500 samples 0.788 -> 0.710 ms per step)
            const bool half_wave = (long long)B * 2 <= lanes;
            const int age = half_wave ? 300 : ((long long)B * 2 <= lanes * 3 ? 350 : 500);
            a.park_age = option(QPB_OPT_MLE_PARK_AGE) > 0 ? option(QPB_OPT_MLE_PARK_AGE) : age;
            {
                // measured on B200 (tools/pauli2_sweep_h.py): while most of the batch starts in the first wave of
                // lanes a broad ramp from 250 pays; with several waves the W workers are too few for that
                const bool few_waves = (long long)B * 2 <= lanes * 3;
                const int lo = option(QPB_OPT_MLE_PARK_AGE_LO), pct = option(QPB_OPT_MLE_PARK_AGE_PCT);
                a.park_age_lo = lo > 0 ? lo : (lo < 0 ? a.park_age : (half_wave ? 100 : (few_waves ? 250 : 350)));
                if (a.park_age_lo > a.park_age) a.park_age_lo = a.park_age;
                // (tools/pauli2_sweep_o.py, 5e4 samples: age 350 from 250 over the first 10 % and no demand-driven tail
                // 1.16 -> 1.07 ms per step)
                const double frac = (pct > 0 ? pct : (half_wave ? 50 : 10)) / 100.0;
                a.park_age_slope = (float)((a.park_age - a.park_age_lo) / (frac * (double)B));
                const int end = option(QPB_OPT_MLE_PARK_AGE_END), pct2 = option(QPB_OPT_MLE_PARK_AGE_PCT2);
                a.park_q2 = (int)(lanes < B ? lanes : B);
                a.park_age_end = end > 0 ? end : a.park_age;
                if (a.park_age_end > a.park_age) a.park_age_end = a.park_age;
                const double span = (pct2 > 0 ? pct2 : 50) / 100.0 * (double)(B - a.park_q2) + 1.0;
                a.park_age_slope2 = (float)((a.park_age - a.park_age_end) / span);
            }
            a.park_live = option(QPB_OPT_MLE_PARK_LIVE) > 0 ? option(QPB_OPT_MLE_PARK_LIVE) : (half_wave ? 20 : 5);
            if (a.single_warps < max_sw || a.single_warps + w_warps > kPauliCtaWarps) w_warps = kPauliCtaWarps - a.single_warps;
            const int poll = option(QPB_OPT_MLE_TAIL_POLL);
            a.tail_poll = poll < 0 ? 0 : (poll == 0 ? (((long long)B * 2 <= lanes * 3 && !half_wave) ? 0 : 4) : poll);
            while (a.tail_poll & (a.tail_poll - 1)) a.tail_poll &= a.tail_poll - 1;  // power of two
            a.tail_age = option(QPB_OPT_MLE_TAIL_AGE) > 0 ? option(QPB_OPT_MLE_TAIL_AGE) : 150;
            const int ad = option(QPB_OPT_MLE_ADOPT);
            a.adopt = ad <= 0 ? 0 : ad;      // measured: no gain, off by default
            // measured (tools/pauli2_sweep_j.py, 1e5): 1 -> 1.745, 2 -> 1.708, 3 -> 1.689, 4 -> 1.687, 8 -> 1.693 ms per step
            a.refill_min = option(QPB_OPT_MLE_REFILL_MIN) > 0 ? option(QPB_OPT_MLE_REFILL_MIN) : 6;
            // (with the write-back deferred to the refill, tools/pauli2_sweep_k.py: 4 -> 1.587, 6 -> 1.580 ms; hand-over age
            // 450 ... 700 and the demand-driven tail within 1 % of each other)
            const int mg = option(QPB_OPT_MLE_MERGE);
            a.merge = mg <= 0 ? 0 : mg;      // measured: fuller warps, but the launch ends later (the packed warps keep
                                             // long runners on the slow mapping); off by default
        }
        threads = 32 * (a.single_warps + w_warps);
    }
    if (blocks < 1) blocks = 1;
    // ---- control block and hand-over list -----------------------------------------------------------------
    // slot 0: control words + ready flags (always at the same offset, so a stale word there is an older epoch or
    // zero, never data of another array); slots 7, 8: the hand-over entries
    const bool listed = !(a.direct || no_merge);
    // a sample can be handed over more than once (adoption by a thread-per-sample lane and back), hence 2 B
    const size_t cap = listed ? 2 * (size_t)B + (size_t)blocks * 12 + 64 : 0;
    a.park_cap = (unsigned)cap;
    bool fresh = false;
    unsigned char* base = static_cast<unsigned char*>(scratch(st, 0, 256 + sizeof(unsigned int) * cap, &fresh));
    if (!base) return QPB_ERR_NOMEM;
    a.ctrl = reinterpret_cast<unsigned int*>(base);
    a.park_ready = reinterpret_cast<unsigned int*>(base + 256);
    a.park_h = nullptr;
    a.park_b = nullptr;
    a.park_it = nullptr;
    if (listed) {
        a.park_h = static_cast<double*>(scratch(st, 7, sizeof(double) * 16 * cap));
        unsigned char* bi = static_cast<unsigned char*>(scratch(st, 8, (sizeof(long long) + sizeof(int)) * cap));
        if (!a.park_h || !bi) return QPB_ERR_NOMEM;
        a.park_b = reinterpret_cast<long long*>(bi);
        a.park_it = reinterpret_cast<int*>(bi + sizeof(long long) * cap);
        if (fresh) QPB_CUDA(cudaMemsetAsync(base, 0, 256 + sizeof(unsigned int) * cap, st));  // no flag may hold a future epoch
    }
    QPB_CUDA(cudaMemsetAsync(a.ctrl, 0, 16, st));
    a.trace = debug_trace_buffer(((size_t)blocks * (threads / 32) * 16 + 4 * (size_t)B) * sizeof(long long));
    a.trace_s = a.trace ? a.trace + (size_t)blocks * (threads / 32) * 16 : nullptr;
    a.epoch = epoch_counter.fetch_add(1);
    if (a.epoch == 0) a.epoch = epoch_counter.fetch_add(1);

    const size_t smem = sizeof(double) * ((size_t)(threads / 32) * wmode::SIZE + (a.direct ? 0 : 36 * kPauliThreadsSingle + kPoolWords * 32 + 8));
    const bool tr = a.trace != nullptr;
    auto pk = pp.uniform ? (tr ? k_mle_rrr_pauli2<true, true> : k_mle_rrr_pauli2<true, false>)
                         : (tr ? k_mle_rrr_pauli2<false, true> : k_mle_rrr_pauli2<false, false>);
    if (a.direct) {
        const bool u = pp.uniform;
        if (threads <= 512)
            pk = u ? (tr ? k_mle_rrr_pauli2_w<true, 512, true> : k_mle_rrr_pauli2_w<true, 512, false>)
                   : (tr ? k_mle_rrr_pauli2_w<false, 512, true> : k_mle_rrr_pauli2_w<false, 512, false>);
        else if (threads <= 768)
            pk = u ? (tr ? k_mle_rrr_pauli2_w<true, 768, true> : k_mle_rrr_pauli2_w<true, 768, false>)
                   : (tr ? k_mle_rrr_pauli2_w<false, 768, true> : k_mle_rrr_pauli2_w<false, 768, false>);
        else
            pk = u ? (tr ? k_mle_rrr_pauli2_w<true, 1024, true> : k_mle_rrr_pauli2_w<true, 1024, false>)
                   : (tr ? k_mle_rrr_pauli2_w<false, 1024, true> : k_mle_rrr_pauli2_w<false, 1024, false>);
    }
    QPB_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pk<<<blocks, threads, smem, st>>>(pp, a);
    QPB_LAUNCHED("k_mle_rrr_pauli2");
    if (hs_done) *hs_done = fuse;
    return QPB_OK;
}

}  // namespace qpb
