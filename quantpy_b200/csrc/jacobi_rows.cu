// Register-resident batched Hermitian Jacobi + physical projection for d = 8, 16 (n = 3, 4):
// d LANES PER MATRIX, lane r owns row r of A and row r of the eigenvector matrix V.
//
// Replaces _make_feasible (quantpy/tomography/state.py:267-273) after the DMMA inversion for n = 3.  The
// shared-memory version (k_project_packed, jacobi.cuh) streams A and V through shared memory three times per
// round and runs at 79 % of the shared-memory pipe (profiles/ncu_r1_prof_lin3_v2.txt).  Here the column update
// A <- A G and V <- V G touches only the lane's own registers; the row update A <- G^dagger A needs the partner
// row, which the two lanes of a pair swap with shuffles (2 d doubles), and the d/2 rotations of a round are
// computed by their p-lanes and broadcast (3 doubles each).  Rows and columns are kept in tournament POSITION order
// (see jacobi_round), so every register index is a compile-time constant without unrolling the rounds.  Same
// rotation and stopping rule as jacobi.cuh; the order of the pairs within a sweep differs.
#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "plan.h"

namespace qpb {

template <int d>
struct RowState {
    double ar[d], ai[d];  // row of A
    double vr[d], vi[d];  // row of V
};

template <int d, int G>
__device__ __forceinline__ double gsum_rows(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One round of d/2 simultaneous rotations in POSITION coordinates (circle method of a round-robin tournament):
// rows and columns of A and the columns of V are stored by position, the pairs are always the mirrored positions
// (i, d-1-i), and after the round the occupants of positions 0 .. d-2 move on by one place (position d-1 stays).
// The pairing is therefore the same code every round -- one loop body that stays in the instruction cache (the
// version unrolled over the d-1 rounds spent 3.4 stall cycles per issue on instruction fetch).
template <int d>
__device__ __forceinline__ void jacobi_round(RowState<d>& S, int gl, int gbase, bool done, double tiny2) {
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int H = d / 2;
    const int partner = d - 1 - gl;
    const bool is_p = gl < H;
    // ---- every p-lane computes the rotation of its own pair; the d/2 results are broadcast to the group -------
    double dg = 0.0, br = 0.0, bi = 0.0;  // own diagonal element; A_pq (on the p-lane)
#pragma unroll
    for (int c = 0; c < d; ++c) {
        if (c == gl) dg = S.ar[c];
        if (c == partner) {
            br = S.ar[c];
            bi = S.ai[c];
        }
    }
    double rc[H], rsr[H], rsi[H];
    {
        const double ga = __shfl_sync(kFull, dg, gbase + partner);  // A_qq lives on the partner lane
        const double b2 = br * br + bi * bi;
        double c = 1.0, sr = 0.0, si = 0.0;
        if (!done && is_p && b2 > tiny2) {
            const double iab = fast_rsqrt(b2);
            const double tau = 0.5 * (ga - dg) * iab;
            const double x = fma(tau, tau, 1.0);
            const double root = x * fast_rsqrt(x);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) * fast_recip(fabs(tau) + root);
            c = fast_rsqrt(fma(t, t, 1.0));
            const double s = t * c;
            sr = s * (br * iab);
            si = s * (bi * iab);
        }
#pragma unroll
        for (int k = 0; k < H; ++k) {
            rc[k] = __shfl_sync(kFull, c, gbase + k);
            rsr[k] = __shfl_sync(kFull, sr, gbase + k);
            rsi[k] = __shfl_sync(kFull, si, gbase + k);
        }
    }
    // ---- columns (k, d-1-k) of the lane's own rows of A and V ---------------------------------------
#pragma unroll
    for (int k = 0; k < H; ++k) {
        constexpr int dm1 = d - 1;
        const int p = k, q = dm1 - k;
        const double c = rc[k], sur = rsr[k], sui = rsi[k];
        {
            const double xr = S.ar[p], xi = S.ai[p], yr = S.ar[q], yi = S.ai[q];
            S.ar[p] = c * xr - (sur * yr + sui * yi);
            S.ai[p] = c * xi - (sur * yi - sui * yr);
            S.ar[q] = (sur * xr - sui * xi) + c * yr;
            S.ai[q] = (sur * xi + sui * xr) + c * yi;
        }
        {
            const double xr = S.vr[p], xi = S.vi[p], yr = S.vr[q], yi = S.vi[q];
            S.vr[p] = c * xr - (sur * yr + sui * yi);
            S.vi[p] = c * xi - (sur * yi - sui * yr);
            S.vr[q] = (sur * xr - sui * xi) + c * yr;
            S.vi[q] = (sur * xi + sui * xr) + c * yi;
        }
    }
    // ---- rows: the two lanes of a pair swap their rows --------------------------------------------------
    //   p-lane: x' = c x - s u y          (x = own row, y = partner row)
    //   q-lane: y' = s conj(u) x + c y    (x = partner row, y = own row)
    double c = 1.0, sur = 0.0, sui = 0.0;
#pragma unroll
    for (int k = 0; k < H; ++k)
        if (gl == k || partner == k) {
            c = rc[k];
            sur = rsr[k];
            sui = rsi[k];
        }
    const bool rotated = (sur != 0.0) || (sui != 0.0);
    const double s1 = is_p ? -sur : sur;
#pragma unroll
    for (int col = 0; col < d; ++col) {
        const double or_ = __shfl_sync(kFull, S.ar[col], gbase + partner);
        const double oi = __shfl_sync(kFull, S.ai[col], gbase + partner);
        const double mr = S.ar[col], mi = S.ai[col];
        S.ar[col] = fma(c, mr, fma(s1, or_, sui * oi));
        S.ai[col] = fma(c, mi, fma(s1, oi, -sui * or_));
    }
    // the rotated element is zero by construction, the diagonal real
#pragma unroll
    for (int col = 0; col < d; ++col) {
        if (col == partner && rotated) S.ar[col] = S.ai[col] = 0.0;
        if (col == gl) S.ai[col] = 0.0;
    }
    // ---- next round's seating: positions 0 .. d-2 move on by one place (rows across lanes, columns in registers) -
    const int src = (gl == d - 1) ? gl : (gl + d - 2) % (d - 1);
    double tr_[d], ti_[d];
#pragma unroll
    for (int col = 0; col < d; ++col) {
        tr_[col] = __shfl_sync(kFull, S.ar[col], gbase + src);
        ti_[col] = __shfl_sync(kFull, S.ai[col], gbase + src);
    }
#pragma unroll
    for (int col = 0; col < d - 1; ++col) {
        constexpr int dm1 = d - 1;
        const int from = (col + dm1 - 1) % dm1;
        S.ar[col] = tr_[from];
        S.ai[col] = ti_[from];
    }
    S.ar[d - 1] = tr_[d - 1];
    S.ai[d - 1] = ti_[d - 1];
    {
        double t0 = S.vr[d - 2], t1 = S.vi[d - 2];
#pragma unroll
        for (int col = d - 2; col > 0; --col) {
            S.vr[col] = S.vr[col - 1];
            S.vi[col] = S.vi[col - 1];
        }
        S.vr[0] = t0;
        S.vi[0] = t1;
    }
}

constexpr int kRowsThreads = 128;

template <int d>
__global__ void __launch_bounds__(kRowsThreads, d == 8 ? 4 : 2)
k_project_rows(int B, const double* __restrict__ h_in, double* __restrict__ rho) {
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int G = d, dd = d * d;
    const int tid = threadIdx.x, lane = tid & 31, gl = tid % G, gbase = lane - gl;
    const long groups_per_block = kRowsThreads / G;
    const long stride = (long)gridDim.x * groups_per_block;
    const long first = (long)blockIdx.x * groups_per_block + tid / G;
    const long rounds = (B + stride - 1) / stride;  // every group runs the same number of rounds (warp-wide shuffles)
    for (long it = 0; it < rounds; ++it) {
        const long b = first + it * stride;
        const bool valid = b < B;
        RowState<d> S;
#pragma unroll
        for (int c = 0; c < d; ++c) {
            cplx z;
            z.re = (c == gl) ? 1.0 : 0.0;
            z.im = 0.0;
            if (valid) z = herm_get(h_in + b * dd, d, gl, c);
            S.ar[c] = z.re;
            S.ai[c] = z.im;
            S.vr[c] = (c == gl) ? 1.0 : 0.0;
            S.vi[c] = 0.0;
        }
        bool done = !valid;
        for (int sweep = 0; sweep < 40; ++sweep) {
            double off = 0.0, fro = 0.0;
#pragma unroll
            for (int c = 0; c < d; ++c) {
                const double m2 = S.ar[c] * S.ar[c] + S.ai[c] * S.ai[c];
                fro += m2;
                if (c != gl) off += m2;
            }
            off = gsum_rows<d, G>(off);
            fro = gsum_rows<d, G>(fro);
            if (off <= 1e-30 * fro || fro == 0.0) done = true;
            if (__all_sync(kFull, done)) break;
            const double tiny2 = 1e-36 * fro;
#pragma unroll 1
            for (int round = 0; round < d - 1; ++round) jacobi_round<d>(S, gl, gbase, done, tiny2);
        }
        // ---- clip the spectrum, recompose rho = V diag(lam) V^dagger / Tr  (state.py:269-273) ------------
        double lam_own = 0.0;
#pragma unroll
        for (int c = 0; c < d; ++c)
            if (c == gl) lam_own = S.ar[c];
        lam_own = fmax(lam_own, kClipState);
        double lam[d], tr = 0.0;
#pragma unroll
        for (int j = 0; j < d; ++j) {
            lam[j] = __shfl_sync(kFull, lam_own, gbase + j);
            tr += lam[j];
        }
        const double inv = 1.0 / tr;
        double wr[d], wi[d];  // lam_j V_aj of my row a
#pragma unroll
        for (int j = 0; j < d; ++j) {
            wr[j] = lam[j] * S.vr[j];
            wi[j] = lam[j] * S.vi[j];
        }
        double2* out = reinterpret_cast<double2*>(rho) + b * dd + (long)gl * d;
#pragma unroll
        for (int bb = 0; bb < d; ++bb) {  // rho_{a,bb} = sum_j (lam_j V_aj) conj(V_bb,j): row bb comes from lane bb
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int j = 0; j < d; ++j) {
                const double yr = __shfl_sync(kFull, S.vr[j], gbase + bb), yi = __shfl_sync(kFull, S.vi[j], gbase + bb);
                re += wr[j] * yr + wi[j] * yi;
                im += wi[j] * yr - wr[j] * yi;
            }
            if (valid) out[bb] = make_double2(re * inv, im * inv);
        }
    }
}

int launch_project_rows(int d, int B, const double* h_in, double* rho, cudaStream_t st) {
    if (d != 8 && d != 16) return QPB_ERR_UNSUPPORTED;
    const long groups = kRowsThreads / d;
    long blocks = ((long)B + groups - 1) / groups;
    const long cap = (long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (d == 8) k_project_rows<8><<<(int)blocks, kRowsThreads, 0, st>>>(B, h_in, rho);
    else k_project_rows<16><<<(int)blocks, kRowsThreads, 0, st>>>(B, h_in, rho);
    QPB_LAUNCHED("k_project_rows");
    return QPB_OK;
}

}  // namespace qpb
