// Warp-cooperative cyclic Jacobi eigensolver for small complex Hermitian matrices held in
// shared memory (d = 2, 4, 8, 16).  Replaces the LAPACK zheevr calls behind scipy.linalg.eigh at
// quantpy/tomography/state.py:270 and quantpy/tomography/process.py:273.
//
// One warp owns one matrix.  A sweep is d-1 rounds of a round-robin tournament; the d/2 pairs of
// a round are disjoint, so their rotations are computed from the same matrix and applied together
// (columns, then rows, then the eigenvector accumulator).
#pragma once
#include "common.cuh"

namespace qpb {

struct jrot {
    double c, s, ur, ui;  // G = [[c, s*u], [-s*conj(u), c]],  u = A_pq / |A_pq|
    int p, q;
};

__device__ __forceinline__ void jacobi_pair(int d, int round, int i, int& p, int& q) {
    const int m = d - 1;
    int a, b;
    if (i == 0) {
        a = m;
        b = round;
    } else {
        a = (round + i) % m;
        b = (round - i + m) % m;
    }
    p = a < b ? a : b;
    q = a < b ? b : a;
}

// Leading dimension of the shared-memory matrices: d + 1 complex (16 B) elements, so that the d rows of
// one column fall into different banks (with ld = d the column updates were 8- to 16-way bank conflicted and
// the kernel ran at 96 % of the shared-memory pipe, profiles/README_r1.md).
__host__ __device__ inline int jacobi_ld(int d) { return d + 1; }

// A: d x d complex Hermitian, row-major with leading dimension ld = jacobi_ld(d) (destroyed: the diagonal holds
//    the eigenvalues on return).
// V: same layout; on return column j is the eigenvector of eigenvalue A[j][j].  If !WANT_V, V is unused.
// rot: d/2 jrot entries of scratch.  All pointers are shared memory private to this warp.
template <bool WANT_V>
__device__ void warp_jacobi(cplx* __restrict__ A, cplx* __restrict__ V, jrot* __restrict__ rot, int d, int lane) {
    const int dd = d * d;
    const int half = d >> 1;
    const int ld = jacobi_ld(d);
    if (WANT_V) {
        for (int e = lane; e < dd; e += 32) {
            V[(e / d) * ld + e % d].re = (e / d == e % d) ? 1.0 : 0.0;
            V[(e / d) * ld + e % d].im = 0.0;
        }
    }
    __syncwarp();
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0, fro = 0.0;
        for (int e = lane; e < dd; e += 32) {
            const cplx z = A[(e / d) * ld + e % d];
            double m2 = z.re * z.re + z.im * z.im;
            fro += m2;
            if (e / d != e % d) off += m2;
        }
        off = warp_sum(off);
        fro = warp_sum(fro);
        if (off <= 1e-30 * fro || fro == 0.0) break;  // relative off-diagonal norm 1e-15: eigenvalues are second order in it
        const double tiny2 = 1e-36 * fro;
        for (int round = 0; round < d - 1; ++round) {
            if (lane < half) {
                int p, q;
                jacobi_pair(d, round, lane, p, q);
                double al = A[p * ld + p].re, ga = A[q * ld + q].re;
                double br = A[p * ld + q].re, bi = A[p * ld + q].im;
                double b2 = br * br + bi * bi;
                jrot r;
                r.p = p;
                r.q = q;
                if (b2 <= tiny2) {
                    r.c = 1.0; r.s = 0.0; r.ur = 1.0; r.ui = 0.0;
                } else {
                    // same rotation as LAPACK-style Jacobi, with reciprocal square roots instead of sqrt + div
                    const double iab = fast_rsqrt(b2);           // 1 / |beta|
                    const double tau = 0.5 * (ga - al) * iab;
                    const double x = fma(tau, tau, 1.0);
                    const double root = x * fast_rsqrt(x);       // sqrt(1 + tau^2)
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) * fast_recip(fabs(tau) + root);
                    r.c = fast_rsqrt(fma(t, t, 1.0));
                    r.s = t * r.c;
                    r.ur = br * iab;
                    r.ui = bi * iab;
                }
                rot[lane] = r;
            }
            __syncwarp();
            // columns: A <- A G  (and V <- V G)
            for (int w = lane; w < half * d; w += 32) {
                const jrot r = rot[w / d];
                const int row = w % d;
                cplx x = A[row * ld + r.p], y = A[row * ld + r.q];
                // su = s*u ; A'_rp = c x - s conj(u) y ; A'_rq = s u x + c y
                double sur = r.s * r.ur, sui = r.s * r.ui;
                cplx nx, ny;
                nx.re = r.c * x.re - (sur * y.re + sui * y.im);
                nx.im = r.c * x.im - (sur * y.im - sui * y.re);
                ny.re = (sur * x.re - sui * x.im) + r.c * y.re;
                ny.im = (sur * x.im + sui * x.re) + r.c * y.im;
                A[row * ld + r.p] = nx;
                A[row * ld + r.q] = ny;
                if (WANT_V) {
                    cplx vx = V[row * ld + r.p], vy = V[row * ld + r.q];
                    cplx mx, my;
                    mx.re = r.c * vx.re - (sur * vy.re + sui * vy.im);
                    mx.im = r.c * vx.im - (sur * vy.im - sui * vy.re);
                    my.re = (sur * vx.re - sui * vx.im) + r.c * vy.re;
                    my.im = (sur * vx.im + sui * vx.re) + r.c * vy.im;
                    V[row * ld + r.p] = mx;
                    V[row * ld + r.q] = my;
                }
            }
            __syncwarp();
            // rows: A <- G^dagger A :  A'_p. = c A_p. - s u A_q. ;  A'_q. = s conj(u) A_p. + c A_q.
            for (int w = lane; w < half * d; w += 32) {
                const jrot r = rot[w / d];
                const int col = w % d;
                cplx x = A[r.p * ld + col], y = A[r.q * ld + col];
                double sur = r.s * r.ur, sui = r.s * r.ui;
                cplx nx, ny;
                nx.re = r.c * x.re - (sur * y.re - sui * y.im);
                nx.im = r.c * x.im - (sur * y.im + sui * y.re);
                ny.re = (sur * x.re + sui * x.im) + r.c * y.re;
                ny.im = (sur * x.im - sui * x.re) + r.c * y.im;
                A[r.p * ld + col] = nx;
                A[r.q * ld + col] = ny;
            }
            __syncwarp();
            if (lane < half) {
                const jrot r = rot[lane];
                if (r.s != 0.0) {
                    A[r.p * ld + r.q].re = 0.0; A[r.p * ld + r.q].im = 0.0;
                    A[r.q * ld + r.p].re = 0.0; A[r.q * ld + r.p].im = 0.0;
                }
                A[r.p * ld + r.p].im = 0.0;
                A[r.q * ld + r.q].im = 0.0;
            }
            __syncwarp();
        }
    }
}

// Sub-warp variant: G lanes (a power of two, G <= 32, G >= d/2) own one matrix, so a warp solves 32/G matrices in
// lockstep -- at d = 8 the rotation set-up of the one-warp-per-matrix version kept 28 of 32 lanes idle.
// Groups that converge early stay in the loop (predicated off) until every group of the warp has converged.
template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <bool WANT_V, int G>
__device__ void group_jacobi(cplx* __restrict__ A, cplx* __restrict__ V, jrot* __restrict__ rot, int d, int gl,
                             bool valid) {
    const int dd = d * d;
    const int half = d >> 1;
    const int ld = jacobi_ld(d);
    if (WANT_V) {
        for (int e = gl; e < dd; e += G) {
            V[(e / d) * ld + e % d].re = (e / d == e % d) ? 1.0 : 0.0;
            V[(e / d) * ld + e % d].im = 0.0;
        }
    }
    __syncwarp();
    bool done = !valid;
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0.0, fro = 0.0;
        if (!done) {
            for (int e = gl; e < dd; e += G) {
                const cplx z = A[(e / d) * ld + e % d];
                const double m2 = z.re * z.re + z.im * z.im;
                fro += m2;
                if (e / d != e % d) off += m2;
            }
        }
        off = group_sum<G>(off);
        fro = group_sum<G>(fro);
        if (off <= 1e-30 * fro || fro == 0.0) done = true;
        if (__all_sync(0xffffffffu, done)) break;
        const double tiny2 = 1e-36 * fro;
        for (int round = 0; round < d - 1; ++round) {
            if (!done && gl < half) {
                int p, q;
                jacobi_pair(d, round, gl, p, q);
                const double al = A[p * ld + p].re, ga = A[q * ld + q].re;
                const double br = A[p * ld + q].re, bi = A[p * ld + q].im;
                const double b2 = br * br + bi * bi;
                jrot r;
                r.p = p;
                r.q = q;
                if (b2 <= tiny2) {
                    r.c = 1.0; r.s = 0.0; r.ur = 1.0; r.ui = 0.0;
                } else {
                    const double iab = fast_rsqrt(b2);
                    const double tau = 0.5 * (ga - al) * iab;
                    const double x = fma(tau, tau, 1.0);
                    const double root = x * fast_rsqrt(x);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) * fast_recip(fabs(tau) + root);
                    r.c = fast_rsqrt(fma(t, t, 1.0));
                    r.s = t * r.c;
                    r.ur = br * iab;
                    r.ui = bi * iab;
                }
                rot[gl] = r;
            }
            __syncwarp();
            if (!done) {
                for (int w = gl; w < half * d; w += G) {
                    const jrot r = rot[w / d];
                    const int row = w % d;
                    const cplx x = A[row * ld + r.p], y = A[row * ld + r.q];
                    const double sur = r.s * r.ur, sui = r.s * r.ui;
                    cplx nx, ny;
                    nx.re = r.c * x.re - (sur * y.re + sui * y.im);
                    nx.im = r.c * x.im - (sur * y.im - sui * y.re);
                    ny.re = (sur * x.re - sui * x.im) + r.c * y.re;
                    ny.im = (sur * x.im + sui * x.re) + r.c * y.im;
                    A[row * ld + r.p] = nx;
                    A[row * ld + r.q] = ny;
                    if (WANT_V) {
                        const cplx vx = V[row * ld + r.p], vy = V[row * ld + r.q];
                        cplx mx, my;
                        mx.re = r.c * vx.re - (sur * vy.re + sui * vy.im);
                        mx.im = r.c * vx.im - (sur * vy.im - sui * vy.re);
                        my.re = (sur * vx.re - sui * vx.im) + r.c * vy.re;
                        my.im = (sur * vx.im + sui * vx.re) + r.c * vy.im;
                        V[row * ld + r.p] = mx;
                        V[row * ld + r.q] = my;
                    }
                }
            }
            __syncwarp();
            if (!done) {
                for (int w = gl; w < half * d; w += G) {
                    const jrot r = rot[w / d];
                    const int col = w % d;
                    const cplx x = A[r.p * ld + col], y = A[r.q * ld + col];
                    const double sur = r.s * r.ur, sui = r.s * r.ui;
                    cplx nx, ny;
                    nx.re = r.c * x.re - (sur * y.re - sui * y.im);
                    nx.im = r.c * x.im - (sur * y.im + sui * y.re);
                    ny.re = (sur * x.re + sui * x.im) + r.c * y.re;
                    ny.im = (sur * x.im - sui * x.re) + r.c * y.im;
                    A[r.p * ld + col] = nx;
                    A[r.q * ld + col] = ny;
                }
            }
            __syncwarp();
            if (!done && gl < half) {
                const jrot r = rot[gl];
                if (r.s != 0.0) {
                    A[r.p * ld + r.q].re = 0.0; A[r.p * ld + r.q].im = 0.0;
                    A[r.q * ld + r.p].re = 0.0; A[r.q * ld + r.p].im = 0.0;
                }
                A[r.p * ld + r.p].im = 0.0;
                A[r.q * ld + r.q].im = 0.0;
            }
            __syncwarp();
        }
    }
}

}  // namespace qpb
