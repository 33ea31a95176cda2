// Register-resident Hermitian Jacobi building blocks: d LANES PER MATRIX, lane r owns row r of A and row r of the
// eigenvector matrix V (see jacobi_rows.cu for the description; used by k_project_rows and k_cptp_rows).
#pragma once
#include "common.cuh"

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

}  // namespace qpb
