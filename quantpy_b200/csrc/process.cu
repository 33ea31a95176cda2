// Process-tomography kernels: 'lifp' linear inversion of the Choi matrix and the alternating
// TP/CP projection of quantpy/tomography/process.py:231-289.
//
// Layout: Choi matrices are s x s complex128, s = d^2, row-major, row/column index (i,a) = i*d + a
// with i the input and a the output factor (Choi = sum_ij E_ij (x) Phi(E_ij), channel.py:95-103).
// LinvT [S*K][s*s][2] holds, for every measurement column, the (re,im) contribution to each Choi
// entry (the column-stacking permutation of routines.py:53-61 is folded in at plan creation).
#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "jacobi.cuh"
#include "jacobi_rows.cuh"
#include "plan.h"

#include <cstdlib>

namespace qpb {

// LinvT[col][r*s + c] = Linv[(c*s + r)][col]   (vec index = column stacking)
__global__ void k_lifp_prep(const double* __restrict__ Linv, int s, int cols, double* __restrict__ LinvT) {
    const long ss = (long)s * s, total = ss * cols;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const long col = t / ss, e = t % ss;
        const int r = (int)(e / s), c = (int)(e % s);
        const long src = ((long)c * s + r) * cols + col;
        LinvT[2 * t] = Linv[2 * src];
        LinvT[2 * t + 1] = Linv[2 * src + 1];
    }
}

// choi[b] = sum_col LinvT[col] * freq[b][col];  freq normalised per input state (process.py:285).
// One warp per sample; shared per warp: freq[S*K].
__global__ void k_lifp(int S, int K, int ss, int B, const double* __restrict__ LinvT,
                       const int32_t* __restrict__ counts, double* __restrict__ choi) {
    extern __shared__ double smf[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int cols = S * K;
    double* f = smf + (size_t)warp * cols;
    for (long b = (long)blockIdx.x * nw + warp; b < B; b += (long)gridDim.x * nw) {
        const int32_t* c = counts + b * cols;
        for (int si = 0; si < S; ++si) {
            long long tot = 0;
            for (int k = lane; k < K; k += 32) tot += c[si * K + k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            const double total = (double)tot;
            for (int k = lane; k < K; k += 32) f[si * K + k] = (double)c[si * K + k] / total;
        }
        __syncwarp();
        double* out = choi + b * 2 * ss;
        for (int e = lane; e < 2 * ss; e += 32) {
            double acc = 0.0;
            for (int col = 0; col < cols; ++col) acc += LinvT[(long)col * 2 * ss + e] * f[col];
            out[e] = acc;
        }
        __syncwarp();
    }
}

// 'states' assembly (process.py:316-324): choi[b] = sum_s G_s (x) rho[b][s], with G_s[i][j] the coefficient of input
// state s in the decomposition of the matrix unit E_ij (basis.py:31-34).  One thread per Choi entry.
__global__ void k_choi_from_states(int d, int S, int B, const double* __restrict__ G, const double* __restrict__ rho,
                                   double* __restrict__ choi) {
    const int s = d * d;
    const long total = (long)B * s * s;
    const cplx* Gc = reinterpret_cast<const cplx*>(G);
    const cplx* rc = reinterpret_cast<const cplx*>(rho);
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const long b = t / (s * s);
        const int e = (int)(t % (s * s)), r = e / s, c = e % s;
        const int i = r / d, a = r % d, j = c / d, bb = c % d;
        double re = 0.0, im = 0.0;
        for (int k = 0; k < S; ++k) {
            const cplx g = Gc[(k * d + i) * d + j], z = rc[((b * S + k) * d + a) * d + bb];
            re += g.re * z.re - g.im * z.im;
            im += g.re * z.im + g.im * z.re;
        }
        choi[2 * t] = re;
        choi[2 * t + 1] = im;
    }
}

__host__ __device__ inline size_t cptp_smem_per_warp(int s) {
    return sizeof(cplx) * (5 * (size_t)s * s + 2 * (size_t)s * jacobi_ld(s)) + sizeof(jrot) * (size_t)(s / 2 + 1) +
           sizeof(cplx) * 16;
}

// Alternating projection of process.py:237-257, one warp per Choi matrix.
// check_atol >= 0: first test Channel.is_cptp(atol) (quantpy/channel.py:144-157) and leave matrices that pass
// untouched (iters = 0), as ProcessTomograph._point_estimate_states does (process.py:325-326).
__global__ void k_cptp(int d, int B, const double* __restrict__ choi_in, int n_iter, double tol, double check_atol,
                       double* __restrict__ choi_out, int32_t* __restrict__ iters) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int s = d * d, ss = s * s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned char* base = smraw + (size_t)warp * cptp_smem_per_warp(s);
    cplx* x = reinterpret_cast<cplx*>(base);
    cplx* p = x + ss;
    cplx* q = p + ss;
    cplx* y = q + ss;
    cplx* T = y + ss;
    const int ld = jacobi_ld(s);
    cplx* A = T + ss;        // A, V: padded leading dimension for the Jacobi solve
    cplx* V = A + s * ld;
    cplx* rin = V + s * ld;  // d*d <= 16
    jrot* rot = reinterpret_cast<jrot*>(rin + 16);
    const double invd = 1.0 / d;

    for (long b = (long)blockIdx.x * nw + warp; b < B; b += (long)gridDim.x * nw) {
        const cplx* src = reinterpret_cast<const cplx*>(choi_in) + b * ss;
        for (int e = lane; e < ss; e += 32) {
            x[e] = src[e];
            p[e].re = p[e].im = 0.0;
            q[e].re = q[e].im = 0.0;
            y[e].re = y[e].im = 0.0;
        }
        __syncwarp();
        if (check_atol >= 0.0) {
            // trace preserving: |Tr_out x - I| <= atol + 1e-5 |I| elementwise (np.allclose)
            double bad = 0.0;
            for (int e = lane; e < d * d; e += 32) {
                const int i = e / d, j = e % d;
                double re = 0.0, im = 0.0;
                for (int a = 0; a < d; ++a) {
                    const cplx z = x[(i * d + a) * s + (j * d + a)];
                    re += z.re;
                    im += z.im;
                }
                const double target = (i == j) ? 1.0 : 0.0;
                const double dev = sqrt((re - target) * (re - target) + im * im);
                if (dev > check_atol + 1e-5 * target) bad = 1.0;
            }
            // completely positive: eigenvalues >= -atol
            for (int e = lane; e < ss; e += 32) {
                const int r = e / s, c = e % s, et = c * s + r;
                A[r * ld + c].re = 0.5 * (x[e].re + x[et].re);
                A[r * ld + c].im = 0.5 * (x[e].im - x[et].im);
            }
            __syncwarp();
            warp_jacobi<false>(A, V, rot, s, lane);
            for (int j = lane; j < s; j += 32)
                if (A[j * ld + j].re < -check_atol) bad = 1.0;
            bad = warp_max(bad);
            __syncwarp();
            if (bad == 0.0) {
                cplx* dst0 = reinterpret_cast<cplx*>(choi_out) + b * ss;
                for (int e = lane; e < ss; e += 32) dst0[e] = x[e];
                if (iters && lane == 0) iters[b] = 0;
                __syncwarp();
                continue;
            }
        }
        int it = 0;
        for (it = 1; it <= n_iter; ++it) {
            // ---- y' = TP(x + p) = t + ((I - Tr_out t) (x) I) / d
            for (int e = lane; e < ss; e += 32) {
                T[e].re = x[e].re + p[e].re;
                T[e].im = x[e].im + p[e].im;
            }
            __syncwarp();
            for (int e = lane; e < d * d; e += 32) {
                const int i = e / d, j = e % d;
                double re = 0.0, im = 0.0;
                for (int a = 0; a < d; ++a) {
                    const cplx z = T[(i * d + a) * s + (j * d + a)];
                    re += z.re;
                    im += z.im;
                }
                rin[e].re = ((i == j) ? 1.0 : 0.0) - re;
                rin[e].im = -im;
            }
            __syncwarp();
            double c1r = 0.0, c1i = 0.0;
            for (int e = lane; e < ss; e += 32) {
                const int r = e / s, c = e % s;
                cplx yn = T[e];
                if (r % d == c % d) {
                    const cplx g = rin[(r / d) * d + (c / d)];
                    yn.re += g.re * invd;
                    yn.im += g.im * invd;
                }
                const double dr = yn.re - y[e].re, di = yn.im - y[e].im;
                // conj(y_diff) * q
                c1r += dr * q[e].re + di * q[e].im;
                c1i += dr * q[e].im - di * q[e].re;
                y[e] = yn;
            }
            c1r = warp_sum(c1r);
            c1i = warp_sum(c1i);
            __syncwarp();
            // ---- x' = CP(y' + q): eigh, clip at 1e-12, recompose
            for (int e = lane; e < ss; e += 32) {
                const int r = e / s, c = e % s, et = c * s + r;
                A[r * ld + c].re = 0.5 * ((y[e].re + q[e].re) + (y[et].re + q[et].re));
                A[r * ld + c].im = 0.5 * ((y[e].im + q[e].im) - (y[et].im + q[et].im));
            }
            __syncwarp();
            warp_jacobi<true>(A, V, rot, s, lane);
            double c2r = 0.0, c2i = 0.0, c3 = 0.0;
            for (int e = lane; e < ss; e += 32) {
                const int r = e / s, c = e % s;
                double re = 0.0, im = 0.0;
                for (int j = 0; j < s; ++j) {
                    const double lam = fmax(A[j * ld + j].re, kClipChoi);
                    const cplx u = V[r * ld + j], v = V[c * ld + j];
                    re += lam * (u.re * v.re + u.im * v.im);
                    im += lam * (u.im * v.re - u.re * v.im);
                }
                const double dr = re - x[e].re, di = im - x[e].im;
                c2r += dr * p[e].re + di * p[e].im;
                c2i += dr * p[e].im - di * p[e].re;
                x[e].re = re;
                x[e].im = im;
                const double pr = re - y[e].re, pi = im - y[e].im;  // p_diff = x' - y'
                p[e].re += pr;
                p[e].im += pi;
                q[e].re -= pr;
                q[e].im -= pi;
                c3 += pr * pr + pi * pi;
            }
            c2r = warp_sum(c2r);
            c2i = warp_sum(c2i);
            c3 = warp_sum(c3);
            __syncwarp();
            const double crit = 2.0 * (sqrt(c1r * c1r + c1i * c1i) + sqrt(c2r * c2r + c2i * c2i)) + 2.0 * c3;
            if (crit < tol) break;
        }
        if (it > n_iter) it = n_iter;
        cplx* dst = reinterpret_cast<cplx*>(choi_out) + b * ss;
        for (int e = lane; e < ss; e += 32) dst[e] = x[e];
        if (iters && lane == 0) iters[b] = it;
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// The same alternating projection for two-qubit channels (Choi 16 x 16) with the eigen-decomposition in registers:
// 16 LANES PER MATRIX (two matrices per warp), lane r owns row r.  x, p, q, y live in shared memory by rows (leading
// dimension 17 complex numbers); the CP step loads the lane's row of herm(y + q) into registers and runs the
// round-robin Jacobi of jacobi_rows.cuh (column updates local, row exchange and rotations by shuffles), and
// x' = V max(lam, 1e-12) V^dagger reads the other rows of V as shared-memory broadcasts.  The warp-per-matrix
// kernel above keeps A and V in shared memory and was latency-bound: 1.75 ms for 1000 channels (4 iterations each).
// ------------------------------------------------------------------------------------------------
constexpr int kCptpRowsThreads = 128;

template <int d>
__global__ void __launch_bounds__(kCptpRowsThreads, 1)
k_cptp_rows(int B, const double* __restrict__ choi_in, int n_iter, double tol, double check_atol,
            double* __restrict__ choi_out, int32_t* __restrict__ iters, int option_cold) {
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int s = d * d, G = s, LD = s + 1, MAT = s * LD;
    static_assert(s == 16, "two matrices per warp");
    extern __shared__ __align__(16) unsigned char smraw[];
    const int tid = threadIdx.x, lane = tid & 31, gl = tid % G, gbase = lane - gl;
    const int groups_per_block = kCptpRowsThreads / G;
    cplx* x = reinterpret_cast<cplx*>(smraw) + (size_t)(tid / G) * 6 * MAT;
    cplx* p = x + MAT;
    cplx* q = p + MAT;
    cplx* y = q + MAT;
    cplx* z = y + MAT;  // transposition scratch
    cplx* vp = z + MAT; // eigenvectors of the last CP step (rows)
    const int qi = gl / d, qa = gl % d;  // row r = (input index i, output index a)
    const double invd = 1.0 / d;
    const long stride = (long)gridDim.x * groups_per_block;
    const long first = (long)blockIdx.x * groups_per_block + tid / G;
    const long rounds = (B + stride - 1) / stride;  // every group of a warp runs the same number of rounds (full-mask shuffles)

    auto group_sum = [&](double v) {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        return v;
    };
    // eigenvalues (and eigenvectors in S.v*) of the Hermitian part of the matrix whose rows are in `m`
    // eigenvalues (diagonal of S.a*) and eigenvectors (S.v*, by rows) of the Hermitian part of the matrix whose rows
    // are in `m`.  warm: vp holds the eigenvectors V0 of a nearby matrix (the previous CP step of the same Dykstra
    // iteration sequence): the Jacobi then starts from V0^dagger A V0, which is almost diagonal (2-3 sweeps instead of
    // 7), with the eigenvector accumulator initialised to V0; costs two 16 x 16 products (~1/3 of a sweep).
    auto eigh_rows = [&](const cplx* m, RowState<s>& S, bool skip, bool warm) {
#pragma unroll
        for (int c = 0; c < s; ++c) {
            const cplx a = m[gl * LD + c], t = m[c * LD + gl];
            S.ar[c] = 0.5 * (a.re + t.re);
            S.ai[c] = (c == gl) ? 0.0 : 0.5 * (a.im - t.im);
            S.vr[c] = (c == gl) ? 1.0 : 0.0;
            S.vi[c] = 0.0;
        }
        if (warm) {  // warp-uniform
            __syncwarp();  // everybody has read m
            cplx* t1 = const_cast<cplx*>(m);  // T1 = A V0 overwrites m (row gl only)
#pragma unroll 2
            for (int j = 0; j < s; ++j) {
                double re = 0.0, im = 0.0;
#pragma unroll
                for (int k = 0; k < s; ++k) {
                    const cplx v = vp[k * LD + j];
                    re += S.ar[k] * v.re - S.ai[k] * v.im;
                    im += S.ar[k] * v.im + S.ai[k] * v.re;
                }
                t1[gl * LD + j].re = re;
                t1[gl * LD + j].im = im;
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < s; ++c) S.ar[c] = S.ai[c] = 0.0;
#pragma unroll 2
            for (int k = 0; k < s; ++k) {  // A'[gl][c] = sum_k conj(V0[k][gl]) T1[k][c]
                const cplx u = vp[k * LD + gl];
#pragma unroll
                for (int c = 0; c < s; ++c) {
                    const cplx t = t1[k * LD + c];
                    S.ar[c] += u.re * t.re + u.im * t.im;
                    S.ai[c] += u.re * t.im - u.im * t.re;
                }
            }
#pragma unroll
            for (int c = 0; c < s; ++c) {
                if (c == gl) S.ai[c] = 0.0;
                S.vr[c] = vp[gl * LD + c].re;
                S.vi[c] = vp[gl * LD + c].im;
            }
        }
        bool done = skip;
        for (int sweep = 0; sweep < 40; ++sweep) {
            double off = 0.0, fro = 0.0;
#pragma unroll
            for (int c = 0; c < s; ++c) {
                const double m2 = S.ar[c] * S.ar[c] + S.ai[c] * S.ai[c];
                fro += m2;
                if (c != gl) off += m2;
            }
            off = group_sum(off);
            fro = group_sum(fro);
            if (off <= 1e-30 * fro || fro == 0.0) done = true;
            if (__all_sync(kFull, done)) break;
            const double tiny2 = 1e-36 * fro;
#pragma unroll 1
            for (int round = 0; round < s - 1; ++round) jacobi_round<s>(S, gl, gbase, done, tiny2);
        }
    };

    for (long rnd = 0; rnd < rounds; ++rnd) {
        const long b = first + rnd * stride;
        const bool valid = b < B;
        __syncwarp();
        {
            const cplx* src = reinterpret_cast<const cplx*>(choi_in) + (valid ? b : 0) * s * s + gl * s;
#pragma unroll
            for (int c = 0; c < s; ++c) {
                cplx v;
                v.re = (c == gl) ? invd : 0.0;
                v.im = 0.0;
                if (valid) v = src[c];
                x[gl * LD + c] = v;
                p[gl * LD + c].re = p[gl * LD + c].im = 0.0;
                q[gl * LD + c].re = q[gl * LD + c].im = 0.0;
                y[gl * LD + c].re = y[gl * LD + c].im = 0.0;
            }
        }
        __syncwarp();
        bool active = valid;
        int it = 0;
        auto store_result = [&]() {
            cplx* dst = reinterpret_cast<cplx*>(choi_out) + b * s * s + gl * s;
#pragma unroll
            for (int c = 0; c < s; ++c) dst[c] = x[gl * LD + c];
        };
        if (check_atol >= 0.0) {
            // Channel.is_cptp (channel.py:144-157): |Tr_out x - I| <= atol + 1e-5 |I| elementwise, eigenvalues >= -atol
            double bad = 0.0;
            double tre[d], tim[d];
#pragma unroll
            for (int j = 0; j < d; ++j) {
                const cplx v = x[gl * LD + j * d + qa];
                tre[j] = v.re;
                tim[j] = v.im;
            }
#pragma unroll
            for (int o = 1; o < d; o <<= 1)
#pragma unroll
                for (int j = 0; j < d; ++j) {
                    tre[j] += __shfl_xor_sync(kFull, tre[j], o);
                    tim[j] += __shfl_xor_sync(kFull, tim[j], o);
                }
#pragma unroll
            for (int j = 0; j < d; ++j) {
                const double target = (qi == j) ? 1.0 : 0.0;
                const double dev = sqrt((tre[j] - target) * (tre[j] - target) + tim[j] * tim[j]);
                if (dev > check_atol + 1e-5 * target) bad = 1.0;
            }
            RowState<s> S;
            eigh_rows(x, S, !valid, false);
            double lam_own = 0.0;
#pragma unroll
            for (int c = 0; c < s; ++c)
                if (c == gl) lam_own = S.ar[c];
            if (lam_own < -check_atol) bad = 1.0;
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) bad = fmax(bad, __shfl_xor_sync(kFull, bad, o));
            if (active && bad == 0.0) {
                store_result();
                active = false;
            }
        }
        for (int iter = 1; iter <= n_iter; ++iter) {
            if (!__any_sync(kFull, active)) break;
            // ---- y' = TP(x + p) = t + ((I - Tr_out t) (x) I) / d
            double tr_[s], ti_[s];
#pragma unroll
            for (int c = 0; c < s; ++c) {
                tr_[c] = x[gl * LD + c].re + p[gl * LD + c].re;
                ti_[c] = x[gl * LD + c].im + p[gl * LD + c].im;
            }
            double gre[d], gim[d];  // sum over the output index a of t[(i,a)][(j,a)]: the four lanes of quad i
#pragma unroll
            for (int j = 0; j < d; ++j) {
                gre[j] = 0.0;
                gim[j] = 0.0;
#pragma unroll
                for (int c = 0; c < s; ++c)
                    if (c / d == j && c % d == qa) {
                        gre[j] = tr_[c];
                        gim[j] = ti_[c];
                    }
            }
#pragma unroll
            for (int o = 1; o < d; o <<= 1)
#pragma unroll
                for (int j = 0; j < d; ++j) {
                    gre[j] += __shfl_xor_sync(kFull, gre[j], o);
                    gim[j] += __shfl_xor_sync(kFull, gim[j], o);
                }
            double c1r = 0.0, c1i = 0.0;
#pragma unroll
            for (int c = 0; c < s; ++c) {
                double yr = tr_[c], yi = ti_[c];
                if (c % d == qa) {
                    const int j = c / d;
                    yr += (((qi == j) ? 1.0 : 0.0) - gre[j]) * invd;
                    yi += -gim[j] * invd;
                }
                const cplx yo = y[gl * LD + c], qq = q[gl * LD + c];
                const double dr = yr - yo.re, di = yi - yo.im;
                c1r += dr * qq.re + di * qq.im;  // conj(y_diff) * q
                c1i += dr * qq.im - di * qq.re;
                y[gl * LD + c].re = yr;
                y[gl * LD + c].im = yi;
                z[gl * LD + c].re = yr + qq.re;  // y' + q, Hermitised by eigh_rows
                z[gl * LD + c].im = yi + qq.im;
            }
            __syncwarp();
            // ---- x' = CP(y' + q): eigh, clip at 1e-12, recompose
            RowState<s> S;
            // cold restart every 16 steps: the accumulated V0 loses orthogonality by ~1e-16 per rotation
            eigh_rows(z, S, !active, iter > 1 && ((iter - 1) & 15) != 0 && !option_cold);
            double lam_own = 0.0;
#pragma unroll
            for (int c = 0; c < s; ++c)
                if (c == gl) lam_own = S.ar[c];
            lam_own = fmax(lam_own, kClipChoi);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < s; ++j) {  // vp <- V (rows): this step's recomposition, the next step's warm start
                vp[gl * LD + j].re = S.vr[j];
                vp[gl * LD + j].im = S.vi[j];
            }
            double wr[s], wi[s];  // lam_j V_rj
#pragma unroll
            for (int j = 0; j < s; ++j) {
                const double lam = __shfl_sync(kFull, lam_own, gbase + j);
                wr[j] = lam * S.vr[j];
                wi[j] = lam * S.vi[j];
            }
            __syncwarp();
            double c2r = 0.0, c2i = 0.0, c3 = 0.0;
#pragma unroll 4
            for (int c = 0; c < s; ++c) {
                double re = 0.0, im = 0.0;
#pragma unroll
                for (int j = 0; j < s; ++j) {
                    const cplx v = vp[c * LD + j];  // the same address for all lanes of the group: a broadcast
                    re += wr[j] * v.re + wi[j] * v.im;
                    im += wi[j] * v.re - wr[j] * v.im;
                }
                const cplx xo = x[gl * LD + c], pp = p[gl * LD + c], yy = y[gl * LD + c];
                const double dr = re - xo.re, di = im - xo.im;
                c2r += dr * pp.re + di * pp.im;
                c2i += dr * pp.im - di * pp.re;
                const double pr = re - yy.re, pi = im - yy.im;  // p_diff = x' - y'
                if (active) {
                    x[gl * LD + c].re = re;
                    x[gl * LD + c].im = im;
                    p[gl * LD + c].re = pp.re + pr;
                    p[gl * LD + c].im = pp.im + pi;
                    q[gl * LD + c].re -= pr;
                    q[gl * LD + c].im -= pi;
                }
                c3 += pr * pr + pi * pi;
            }
            c1r = group_sum(c1r);
            c1i = group_sum(c1i);
            c2r = group_sum(c2r);
            c2i = group_sum(c2i);
            c3 = group_sum(c3);
            __syncwarp();
            const double crit = 2.0 * (sqrt(c1r * c1r + c1i * c1i) + sqrt(c2r * c2r + c2i * c2i)) + 2.0 * c3;
            if (active) {
                it = iter;
                if (crit < tol) {
                    store_result();
                    active = false;
                }
            }
        }
        if (active) store_result();  // iteration cap reached (or n_iter == 0)
        if (valid && iters && gl == 0) iters[b] = it;
    }
}

static int launch_cptp(int n, int B, const double* in, int n_iter, double tol, double check_atol, double* out,
                       int32_t* iters, cudaStream_t st) {
    const int d = 1 << n, s = d * d;
    if (s == 16 && !option(QPB_OPT_NO_ROW_JACOBI)) {
        const int groups = kCptpRowsThreads / 16;
        const size_t rsmem = sizeof(cplx) * 6 * 16 * 17 * (size_t)groups;
        QPB_CUDA(cudaFuncSetAttribute(k_cptp_rows<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        long rblocks = ((long)B + groups - 1) / groups;
        const long rcap = (long)num_sms();
        if (rblocks > rcap) rblocks = rcap;
        k_cptp_rows<4><<<(int)rblocks, kCptpRowsThreads, rsmem, st>>>(B, in, n_iter, tol, check_atol, out, iters,
                                                                  option(QPB_OPT_NO_WARM_JACOBI));
        QPB_LAUNCHED("k_cptp_rows");
        return QPB_OK;
    }
    const int warps = (s >= 16) ? 4 : 8;
    const size_t smem = warps * cptp_smem_per_warp(s);
    QPB_REQUIRE(smem <= 227 * 1024, "CPTP projection needs %zu bytes of shared memory", smem);
    if (smem > 48 * 1024) QPB_CUDA(cudaFuncSetAttribute(k_cptp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long blocks = ((long)B + warps - 1) / warps;
    const long cap = (long)num_sms() * (s >= 16 ? 1 : 4);
    if (blocks > cap) blocks = cap;
    k_cptp<<<(int)blocks, warps * 32, smem, st>>>(d, B, in, n_iter, tol, check_atol, out, iters);
    QPB_LAUNCHED("k_cptp");
    return QPB_OK;
}

}  // namespace qpb

using namespace qpb;

extern "C" {

int qpb_process_plan_create(qpb_process_plan** out, int n_qubits, int S, int K, const double* Linv, void* stream) {
    QPB_REQUIRE(out != nullptr, "plan output pointer is NULL");
    QPB_REQUIRE(n_qubits >= 1 && n_qubits <= 2, "process tomography supports 1 or 2 qubits (got %d)", n_qubits);
    QPB_REQUIRE(S >= 1 && K >= 1 && Linv != nullptr, "bad process plan arguments");
    qpb_process_plan* p = new qpb_process_plan();
    p->n = n_qubits;
    p->d = 1 << n_qubits;
    p->S = S;
    p->K = K;
    p->d4 = p->d * p->d * p->d * p->d;
    const size_t bytes = sizeof(double) * 2 * (size_t)p->d4 * S * K;
    int rc = check_cuda(cudaMalloc(&p->LinvT, bytes), "cudaMalloc LinvT");
    if (rc != QPB_OK) {
        delete p;
        return rc;
    }
    const long total = (long)p->d4 * S * K;
    const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    k_lifp_prep<<<grid, 256, 0, (cudaStream_t)stream>>>(Linv, p->d * p->d, S * K, p->LinvT);
    QPB_LAUNCHED("k_lifp_prep");
    *out = p;
    return QPB_OK;
}

int qpb_process_plan_destroy(qpb_process_plan* p) {
    if (!p) return QPB_OK;
    cudaFree(p->LinvT);
    delete p;
    return QPB_OK;
}

int qpb_lifp_cptp(const qpb_process_plan* plan, int B, const int32_t* counts, int cptp, int n_iter, double tol,
                  double* choi, int32_t* iters, void* stream) {
    QPB_REQUIRE(plan != nullptr, "plan is NULL");
    QPB_REQUIRE(B >= 0 && n_iter >= 0, "bad arguments");
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(counts && choi, "NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const int cols = plan->S * plan->K;
    if (!option(QPB_OPT_NO_DMMA_GEMM)) {
        // choi [B][2 d^4] = freq [B][S*K] * LinvT [S*K][2 d^4], frequencies normalised per input state
        int rc = launch_gemm_counts(B, 2 * plan->d4, cols, plan->K, counts, plan->LinvT, choi, st);
        if (rc != QPB_OK) return rc;
    } else {
        const int warps = 8;
        const size_t smem = sizeof(double) * (size_t)warps * cols;
        QPB_REQUIRE(smem <= 200 * 1024, "S*K=%d too large", cols);
        if (smem > 48 * 1024)
            QPB_CUDA(cudaFuncSetAttribute(k_lifp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        long blocks = ((long)B + warps - 1) / warps;
        const long cap = (long)num_sms() * 4;
        if (blocks > cap) blocks = cap;
        k_lifp<<<(int)blocks, warps * 32, smem, st>>>(plan->S, plan->K, plan->d4, B, plan->LinvT, counts, choi);
        QPB_LAUNCHED("k_lifp");
    }
    if (cptp) return launch_cptp(plan->n, B, choi, n_iter, tol, -1.0, choi, iters, st);
    if (iters) QPB_CUDA(cudaMemsetAsync(iters, 0, sizeof(int32_t) * (size_t)B, st));
    return QPB_OK;
}

int qpb_cptp_project(int n_qubits, int B, const double* choi_in, int n_iter, double tol, double* choi_out,
                     int32_t* iters, void* stream) {
    QPB_REQUIRE(n_qubits >= 1 && n_qubits <= 2, "CPTP projection supports 1 or 2 qubits (got %d)", n_qubits);
    QPB_REQUIRE(B >= 0 && n_iter >= 0, "bad arguments");
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(choi_in && choi_out, "NULL buffer");
    return launch_cptp(n_qubits, B, choi_in, n_iter, tol, -1.0, choi_out, iters, (cudaStream_t)stream);
}

int qpb_cptp_project_if_needed(int n_qubits, int B, const double* choi_in, int n_iter, double tol, double atol,
                               double* choi_out, int32_t* iters, void* stream) {
    QPB_REQUIRE(n_qubits >= 1 && n_qubits <= 2, "CPTP projection supports 1 or 2 qubits (got %d)", n_qubits);
    QPB_REQUIRE(B >= 0 && n_iter >= 0 && atol >= 0.0, "bad arguments");
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(choi_in && choi_out, "NULL buffer");
    return launch_cptp(n_qubits, B, choi_in, n_iter, tol, atol, choi_out, iters, (cudaStream_t)stream);
}

int qpb_choi_from_states(int n_qubits, int S, int B, const double* G, const double* rho, double* choi, void* stream) {
    QPB_REQUIRE(n_qubits >= 1 && n_qubits <= 2 && S >= 1 && B >= 0, "bad arguments");
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(G && rho && choi, "NULL buffer");
    const int d = 1 << n_qubits;
    const long total = (long)B * d * d * d * d;
    k_choi_from_states<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, S, B, G, rho, choi);
    QPB_LAUNCHED("k_choi_from_states");
    return QPB_OK;
}

}  // extern "C"
