// State-tomography kernels: operator-table preparation, POVM probabilities, linear inversion with
// physical projection, the generic (any n<=4) R.rho.R maximum-likelihood kernel and distances.
//
// Data layout in HBM
//   Ar  [K][D]  packed-Hermitian POVM operators E_k = sum_i A[k,i] sigma_i         (row per outcome)
//   ArT [D][K]  the same table transposed, so that either contraction reads it coalesced
//   LhT [K][D]  packed-Hermitian linear-inversion map: rho_hat(packed) = sum_k LhT[k][:] f_k
//   counts [B][K] int32, rho [B][d][d] complex128 (row-major, (re,im) pairs)
#include <cmath>
#include <cstdlib>

#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "jacobi.cuh"
#include "plan.h"

namespace qpb {

// ------------------------------------------------------------------------------------------------
// Pauli-string matrix element sigma_i[a,b] in {0, +-1, +-i}; Pauli order quantpy/routines.py:14-19
// (first qubit is the most significant base-4 digit of i and the most significant bit of a, b).
// Returns the power of i (0..3), or -1 when the element is zero.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pauli_phase(int n, int i, int a, int b) {
    int k = 0;
    for (int j = 0; j < n; ++j) {
        const int sh = n - 1 - j;
        const int dig = (i >> (2 * sh)) & 3;
        const int aj = (a >> sh) & 1, bj = (b >> sh) & 1;
        if (dig == 0) {
            if (aj != bj) return -1;
        } else if (dig == 1) {
            if (aj == bj) return -1;
        } else if (dig == 2) {
            if (aj == bj) return -1;
            k += (aj == 0) ? 3 : 1;  // Y[0,1] = -i, Y[1,0] = +i
        } else {
            if (aj != bj) return -1;
            k += 2 * aj;
        }
    }
    return k & 3;
}

// out[v][idx] = scale * packed( sum_i in[v*sv + i*si] sigma_i )[idx]; optionally also outT[idx][v].
__global__ void k_bloch_to_packed(const double* __restrict__ in, int nvec, long sv, long si, int n, double scale,
                                  double* __restrict__ out, double* __restrict__ outT) {
    const int d = 1 << n, D = d * d;
    const long total = (long)nvec * D;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int v = (int)(t / D), idx = (int)(t % D);
        int a = idx / d, b = idx % d;
        const bool want_im = a > b;
        if (want_im) {
            int tmp = a; a = b; b = tmp;  // imaginary part of the upper element
        }
        double acc = 0.0;
        for (int i = 0; i < D; ++i) {
            const int ph = pauli_phase(n, i, a, b);
            if (ph < 0) continue;
            const double x = in[v * sv + i * si];
            if (!want_im) {
                if (ph == 0) acc += x;
                else if (ph == 2) acc -= x;
            } else {
                if (ph == 1) acc += x;
                else if (ph == 3) acc -= x;
            }
        }
        acc *= scale;
        out[(long)v * D + idx] = acc;
        if (outT) outT[(long)idx * nvec + v] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// k1: p[b,k] = scale * M[k,:] . r[b,:]   (state.py:109-110)
// One CTA per 8 samples; r rows staged in shared memory, M read coalesced along i by a warp per k.
// ------------------------------------------------------------------------------------------------
constexpr int PROBS_BT = 8;
__global__ void k_povm_probs(int K, int D, int B, const double* __restrict__ M, const double* __restrict__ r,
                             double scale, int clip, double* __restrict__ p) {
    extern __shared__ double sm[];  // [PROBS_BT][D]
    const int b0 = blockIdx.x * PROBS_BT;
    const int nb = min(PROBS_BT, B - b0);
    for (int e = threadIdx.x; e < nb * D; e += blockDim.x) sm[e] = r[(long)b0 * D + e];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int k = warp; k < K; k += nw) {
        double acc[PROBS_BT];
#pragma unroll
        for (int j = 0; j < PROBS_BT; ++j) acc[j] = 0.0;
        for (int i = lane; i < D; i += 32) {
            const double m = M[(long)k * D + i];
#pragma unroll
            for (int j = 0; j < PROBS_BT; ++j)
                if (j < nb) acc[j] += m * sm[j * D + i];
        }
#pragma unroll
        for (int j = 0; j < PROBS_BT; ++j) {
            double s = warp_sum(acc[j]);
            if (lane == 0 && j < nb) {
                s *= scale;
                if (clip) s = fmin(fmax(s, 0.0), 1.0);
                p[(long)(b0 + j) * K + k] = s;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// helpers on shared-memory d x d complex matrices owned by one warp
// ------------------------------------------------------------------------------------------------
// all matrices below use the padded leading dimension ld = jacobi_ld(d)
__device__ __forceinline__ void warp_matmul(cplx* __restrict__ C, const cplx* __restrict__ X,
                                            const cplx* __restrict__ Y, int d, int lane) {
    const int ld = jacobi_ld(d);
    for (int e = lane; e < d * d; e += 32) {
        const int a = e / d, b = e % d;
        double re = 0.0, im = 0.0;
        for (int c = 0; c < d; ++c) {
            const cplx x = X[a * ld + c], y = Y[c * ld + b];
            re += x.re * y.re - x.im * y.im;
            im += x.re * y.im + x.im * y.re;
        }
        C[a * ld + b].re = re;
        C[a * ld + b].im = im;
    }
}

// C = V diag(f(lambda)) V^dagger, lambda read from the diagonal of A after warp_jacobi.
// mode 0: max(lambda, floor) ; mode 1: sqrt(max(lambda, 0))
__device__ __forceinline__ void warp_recompose(cplx* __restrict__ C, const cplx* __restrict__ A,
                                               const cplx* __restrict__ V, int d, int lane, int mode, double floor_) {
    const int ld = jacobi_ld(d);
    for (int e = lane; e < d * d; e += 32) {
        const int a = e / d, b = e % d;
        double re = 0.0, im = 0.0;
        for (int j = 0; j < d; ++j) {
            double lam = A[j * ld + j].re;
            lam = (mode == 0) ? fmax(lam, floor_) : sqrt(fmax(lam, 0.0));
            const cplx x = V[a * ld + j], y = V[b * ld + j];  // x * conj(y)
            re += lam * (x.re * y.re + x.im * y.im);
            im += lam * (x.im * y.re - x.re * y.im);
        }
        C[a * ld + b].re = re;
        C[a * ld + b].im = im;
    }
}

// ------------------------------------------------------------------------------------------------
// k3-k5: linear inversion + physical projection  (state.py:191-202, 267-273)
// One warp per sample.  shared per warp: f[K] | h[D] | A[dd] | V[dd] | rot[d/2]
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t lin_smem_per_warp(int K, int d) {
    const int dd = d * d, pad = d * jacobi_ld(d);
    return sizeof(double) * (size_t)(K + dd) + sizeof(cplx) * 2 * (size_t)pad + sizeof(jrot) * (size_t)(d / 2 + 1);
}

__global__ void k_lin_project(int n, int K, int B, const double* __restrict__ LhT,
                              const int32_t* __restrict__ counts, const double* __restrict__ h_in, int physical,
                              double* __restrict__ rho) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int d = 1 << n, dd = d * d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int ld = jacobi_ld(d);
    unsigned char* base = smraw + (size_t)warp * lin_smem_per_warp(K, d);
    cplx* A = reinterpret_cast<cplx*>(base);
    cplx* V = A + d * ld;
    double* f = reinterpret_cast<double*>(V + d * ld);
    double* h = f + K;
    jrot* rot = reinterpret_cast<jrot*>(h + dd);

    for (long b = (long)blockIdx.x * nw + warp; b < B; b += (long)gridDim.x * nw) {
        if (h_in) {  // inversion already done by the DMMA GEMM (gemm_dmma.cu)
            for (int idx = lane; idx < dd; idx += 32) h[idx] = h_in[b * dd + idx];
        } else {
            const int32_t* c = counts + b * K;
            long long tot = 0;
            for (int k = lane; k < K; k += 32) tot += c[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            const double total = (double)tot;
            for (int k = lane; k < K; k += 32) f[k] = (double)c[k] / total;  // state.py:193
            __syncwarp();
            for (int idx = lane; idx < dd; idx += 32) {
                double acc = 0.0;
                for (int k = 0; k < K; ++k) acc += LhT[(long)k * dd + idx] * f[k];
                h[idx] = acc;
            }
        }
        __syncwarp();
        for (int e = lane; e < dd; e += 32) A[(e / d) * ld + e % d] = herm_get(h, d, e / d, e % d);
        __syncwarp();
        double* out = rho + b * 2 * dd;
        if (!physical) {
            for (int e = lane; e < dd; e += 32) {
                out[2 * e] = A[(e / d) * ld + e % d].re;
                out[2 * e + 1] = A[(e / d) * ld + e % d].im;
            }
        } else {
            warp_jacobi<true>(A, V, rot, d, lane);
            double tr = 0.0;
            for (int j = 0; j < d; ++j) tr += fmax(A[j * ld + j].re, kClipState);
            const double inv = 1.0 / tr;
            for (int e = lane; e < dd; e += 32) {
                const int a = e / d, bb = e % d;
                double re = 0.0, im = 0.0;
                for (int j = 0; j < d; ++j) {
                    const double lam = fmax(A[j * ld + j].re, kClipState);
                    const cplx x = V[a * ld + j], y = V[bb * ld + j];
                    re += lam * (x.re * y.re + x.im * y.im);
                    im += lam * (x.im * y.re - x.re * y.im);
                }
                out[2 * e] = re * inv;
                out[2 * e + 1] = im * inv;
            }
        }
        __syncwarp();
    }
}

// Projection of already-inverted packed states h_in [B][d*d] (from the DMMA GEMM): G lanes per sample, 32/G samples
// per warp (group_jacobi).  shared per group: A, V (d x ld complex) and d/2 rotations.
__host__ __device__ inline size_t project_smem_per_group(int d) {
    return sizeof(cplx) * 2 * (size_t)d * jacobi_ld(d) + sizeof(jrot) * (size_t)(d / 2 + 1);
}

template <int G>
__global__ void k_project_packed(int d, int B, const double* __restrict__ h_in, int physical, double* __restrict__ rho) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int dd = d * d, ld = jacobi_ld(d);
    const int groups_per_block = blockDim.x / G;
    const int grp = threadIdx.x / G, gl = threadIdx.x % G;
    unsigned char* base = smraw + (size_t)grp * project_smem_per_group(d);
    cplx* A = reinterpret_cast<cplx*>(base);
    cplx* V = A + d * ld;
    jrot* rot = reinterpret_cast<jrot*>(V + d * ld);
    const long stride = (long)gridDim.x * groups_per_block;
    const long first = (long)blockIdx.x * groups_per_block + grp;
    const long rounds = (B + stride - 1) / stride;  // every group runs the same number of rounds (warp-wide syncs)
    for (long r = 0; r < rounds; ++r) {
        const long b = first + r * stride;
        const bool valid = b < B;
        if (valid) {
            const double* h = h_in + b * dd;
            for (int e = gl; e < dd; e += G) A[(e / d) * ld + e % d] = herm_get(h, d, e / d, e % d);
        }
        __syncwarp();
        double* out = rho + b * 2 * dd;
        if (!physical) {
            if (valid)
                for (int e = gl; e < dd; e += G) {
                    out[2 * e] = A[(e / d) * ld + e % d].re;
                    out[2 * e + 1] = A[(e / d) * ld + e % d].im;
                }
        } else {
            group_jacobi<true, G>(A, V, rot, d, gl, valid);
            if (valid) {
                double tr = 0.0;
                for (int j = 0; j < d; ++j) tr += fmax(A[j * ld + j].re, kClipState);
                const double inv = 1.0 / tr;
                for (int e = gl; e < dd; e += G) {
                    const int a = e / d, bb = e % d;
                    double re = 0.0, im = 0.0;
                    for (int j = 0; j < d; ++j) {
                        const double lam = fmax(A[j * ld + j].re, kClipState);
                        const cplx x = V[a * ld + j], y = V[bb * ld + j];
                        re += lam * (x.re * y.re + x.im * y.im);
                        im += lam * (x.im * y.re - x.re * y.im);
                    }
                    out[2 * e] = re * inv;
                    out[2 * e + 1] = im * inv;
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Generic R.rho.R maximum likelihood (any n <= 4): one warp per sample, state in shared memory.
// shared per warp: f[K] | w[K] | h[dd] | h2[dd] | Rh[dd] | hn[dd] | S[dd] cplx
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t mle_smem_per_warp(int K, int d) {
    const int dd = d * d;
    return sizeof(double) * (size_t)(2 * K + 4 * dd) + sizeof(cplx) * (size_t)dd;
}

__global__ void k_mle_rrr_generic(int n, int K, int B, const double* __restrict__ Ar, const double* __restrict__ ArT,
                                  const int32_t* __restrict__ counts, const double* __restrict__ rho0, int max_iter,
                                  double tol, double* __restrict__ rho, int32_t* __restrict__ iters) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int d = 1 << n, dd = d * d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned char* base = smraw + (size_t)warp * mle_smem_per_warp(K, d);
    cplx* S = reinterpret_cast<cplx*>(base);
    double* f = reinterpret_cast<double*>(S + dd);
    double* w = f + K;
    double* h = w + K;
    double* h2 = h + dd;
    double* Rh = h2 + dd;
    double* hn = Rh + dd;

    for (long b = (long)blockIdx.x * nw + warp; b < B; b += (long)gridDim.x * nw) {
        const int32_t* c = counts + b * K;
        long long tot = 0;
        for (int k = lane; k < K; k += 32) tot += c[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        const double total = (double)tot;
        for (int k = lane; k < K; k += 32) f[k] = (double)c[k] / total;  // state.py:227
        for (int e = lane; e < dd; e += 32) {
            const int a = e / d, bb = e % d;
            double v;
            if (rho0) {
                const double* r0 = rho0 + b * 2 * dd;
                v = (a <= bb) ? r0[2 * (a * d + bb)] : r0[2 * (bb * d + a) + 1];
            } else {
                v = (a == bb) ? 1.0 / d : 0.0;
            }
            h[e] = v;
        }
        __syncwarp();
        int it = 0;
        for (it = 1; it <= max_iter; ++it) {
            for (int e = lane; e < dd; e += 32) h2[e] = (e / d == e % d) ? h[e] : 2.0 * h[e];
            __syncwarp();
            // p_k = Tr(E_k rho) ; w_k = f_k / (p_k + 1e-10)
            for (int k = lane; k < K; k += 32) {
                double p = 0.0;
                for (int idx = 0; idx < dd; ++idx) p += ArT[(long)idx * K + k] * h2[idx];
                w[k] = f[k] / (p + kLogGuard);
            }
            __syncwarp();
            // R = sum_k w_k E_k  (packed)
            for (int idx = lane; idx < dd; idx += 32) {
                double acc = 0.0;
                for (int k = 0; k < K; ++k) acc += w[k] * Ar[(long)k * dd + idx];
                Rh[idx] = acc;
            }
            __syncwarp();
            // S = R rho
            for (int e = lane; e < dd; e += 32) {
                const int a = e / d, bb = e % d;
                double re = 0.0, im = 0.0;
                for (int cc = 0; cc < d; ++cc) {
                    const cplx x = herm_get(Rh, d, a, cc), y = herm_get(h, d, cc, bb);
                    re += x.re * y.re - x.im * y.im;
                    im += x.re * y.im + x.im * y.re;
                }
                S[e].re = re;
                S[e].im = im;
            }
            __syncwarp();
            // rho' = S R, upper triangle only (Hermitian), packed
            double tr = 0.0;
            for (int e = lane; e < dd; e += 32) {
                const int a = e / d, bb = e % d;
                if (a > bb) continue;
                double re = 0.0, im = 0.0;
                for (int cc = 0; cc < d; ++cc) {
                    const cplx x = S[a * d + cc], y = herm_get(Rh, d, cc, bb);
                    re += x.re * y.re - x.im * y.im;
                    im += x.re * y.im + x.im * y.re;
                }
                hn[a * d + bb] = re;
                if (a != bb) hn[bb * d + a] = im;
                else tr += re;
            }
            tr = warp_sum(tr);
            __syncwarp();
            const double inv = 1.0 / tr;
            double del = 0.0;
            for (int e = lane; e < dd; e += 32) {
                const double v = hn[e] * inv;
                const double df = v - h[e];
                del += ((e / d == e % d) ? 1.0 : 2.0) * df * df;
                h[e] = v;
            }
            del = sqrt(warp_sum(del));
            __syncwarp();
            if (del < tol) break;
        }
        if (it > max_iter) it = max_iter;
        double* out = rho + b * 2 * dd;
        for (int e = lane; e < dd; e += 32) {
            const cplx z = herm_get(h, d, e / d, e % d);
            out[2 * e] = z.re;
            out[2 * e + 1] = z.im;
        }
        if (iters && lane == 0) iters[b] = it;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// k8: distances (geometry.py:5-56).  One warp per sample; shared per warp: 4 dd cplx + rot
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t dist_smem_per_warp(int d) {
    return sizeof(cplx) * 4 * (size_t)d * jacobi_ld(d) + sizeof(jrot) * (size_t)(d / 2 + 1);
}

__global__ void k_distance(int d, int B, const double* __restrict__ rho, const double* __restrict__ ref, int kind,
                           double* __restrict__ dist) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int dd = d * d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int ld = jacobi_ld(d), pad = d * ld;
    unsigned char* base = smraw + (size_t)warp * dist_smem_per_warp(d);
    cplx* A = reinterpret_cast<cplx*>(base);
    cplx* V = A + pad;
    cplx* T1 = V + pad;
    cplx* T2 = T1 + pad;
    jrot* rot = reinterpret_cast<jrot*>(T2 + pad);
    const cplx* refc = reinterpret_cast<const cplx*>(ref);

    for (long b = (long)blockIdx.x * nw + warp; b < B; b += (long)gridDim.x * nw) {
        const cplx* x = reinterpret_cast<const cplx*>(rho) + b * dd;
        double val;
        if (kind == QPB_DIST_HS) {
            // |Tr (x-ref)(x-ref)| = |sum_ab D_ab D_ba|   (geometry.py:16, no conjugate)
            double re = 0.0, im = 0.0;
            for (int e = lane; e < dd; e += 32) {
                const int a = e / d, bb = e % d;
                const cplx p = x[e], q = refc[e], pt = x[bb * d + a], qt = refc[bb * d + a];
                const double ur = p.re - q.re, ui = p.im - q.im, vr = pt.re - qt.re, vi = pt.im - qt.im;
                // explicit operations: the fused write-back of k_mle_rrr_pauli2 (hs_distance_packed) evaluates the
                // same terms and the same butterfly tree, so fused and unfused bootstraps agree in every bit
                re = __dadd_rn(re, __fma_rn(-ui, vi, __dmul_rn(ur, vr)));
                im = __dadd_rn(im, __fma_rn(ui, vr, __dmul_rn(ur, vi)));
            }
            re = warp_sum(re);
            im = warp_sum(im);
            val = sqrt(sqrt(__fma_rn(im, im, __dmul_rn(re, re)))) / sqrt(2.0);
        } else if (kind == QPB_DIST_TRACE) {
            for (int e = lane; e < dd; e += 32) {
                const int a = e / d, bb = e % d;
                const cplx p = x[e], q = refc[e], pt = x[bb * d + a], qt = refc[bb * d + a];
                A[a * ld + bb].re = 0.5 * ((p.re - q.re) + (pt.re - qt.re));
                A[a * ld + bb].im = 0.5 * ((p.im - q.im) - (pt.im - qt.im));
            }
            __syncwarp();
            warp_jacobi<false>(A, V, rot, d, lane);
            double s = 0.0;
            for (int j = 0; j < d; ++j) s += fabs(A[j * ld + j].re);
            val = 0.5 * s;
        } else {
            // 1 - (Tr sqrt( sqrt(x) ref sqrt(x) ))^2  (geometry.py:52), eigenvalue form
            for (int e = lane; e < dd; e += 32) {
                const int a = e / d, bb = e % d;
                const cplx p = x[e], pt = x[bb * d + a];
                A[a * ld + bb].re = 0.5 * (p.re + pt.re);
                A[a * ld + bb].im = 0.5 * (p.im - pt.im);
                T2[a * ld + bb] = refc[e];
            }
            __syncwarp();
            warp_jacobi<true>(A, V, rot, d, lane);
            warp_recompose(T1, A, V, d, lane, 1, 0.0);  // T1 = sqrt(x)
            __syncwarp();
            warp_matmul(A, T1, T2, d, lane);  // A = sqrt(x) ref
            __syncwarp();
            warp_matmul(V, A, T1, d, lane);  // V = sqrt(x) ref sqrt(x)
            __syncwarp();
            for (int e = lane; e < dd; e += 32) {
                const int a = e / d, bb = e % d;
                const cplx p = V[a * ld + bb], pt = V[bb * ld + a];
                A[a * ld + bb].re = 0.5 * (p.re + pt.re);
                A[a * ld + bb].im = 0.5 * (p.im - pt.im);
            }
            __syncwarp();
            warp_jacobi<false>(A, T1, rot, d, lane);
            double s = 0.0;
            for (int j = 0; j < d; ++j) s += sqrt(fmax(A[j * ld + j].re, 0.0));
            val = 1.0 - s * s;
        }
        if (lane == 0) dist[b] = (val < kZeroBelow) ? 0.0 : val;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int grid_for_warps(long B, int warps_per_block, int max_blocks_per_sm) {
    long need = (B + warps_per_block - 1) / warps_per_block;
    long cap = (long)num_sms() * max_blocks_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

template <typename F>
static int enable_smem(F* fn, size_t bytes) {
    if (bytes > 48 * 1024) {
        QPB_REQUIRE(bytes <= 227 * 1024, "shared memory request %zu exceeds 227 KB", bytes);
        QPB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    return QPB_OK;
}

int launch_lin_project(const qpb_state_plan* plan, int B, const int32_t* counts, int physical, double* rho,
                       cudaStream_t st) {
    const int warps = (plan->n >= 4) ? 4 : 8;
    const size_t smem = warps * lin_smem_per_warp(plan->K, plan->d);
    int rc = enable_smem(k_lin_project, smem);
    if (rc != QPB_OK) return rc;
    const int grid = grid_for_warps(B, warps, 4);
    const double* h_in = nullptr;
    if (plan->n >= 3 && !option(QPB_OPT_NO_DMMA_GEMM)) {
        // batched inversion on the FP64 tensor cores: H [B][D] = freq [B][K] * LhT [K][D]
        double* H = static_cast<double*>(scratch(st, 3, sizeof(double) * (size_t)B * plan->D));
        if (!H) return QPB_ERR_NOMEM;
        rc = launch_gemm_counts(B, plan->D, plan->K, plan->K, counts, plan->LhT, H, st);
        if (rc != QPB_OK) return rc;
        h_in = H;
    }
    // d = 8, 16: one lane per matrix row, state in registers (measured per 1e5 / 1e4 matrices: 1.35 -> 0.95 ms against
    // the shared-memory kernel at d = 8, 2.1 -> 1.4 ms against the warp-per-matrix kernel at d = 16)
    if (h_in && physical && (plan->d == 8 || plan->d == 16) && !option(QPB_OPT_NO_ROW_JACOBI))
        return launch_project_rows(plan->d, B, h_in, rho, st);
    if (h_in && plan->d == 8 && !option(QPB_OPT_NO_PACKED_JACOBI)) {
        // 4 samples per warp, G = d = 8 lanes each (measured: 2.0 -> 1.35 ms per 1e5 matrices; at d = 16 two
        // matrices per warp were slower than one, 2.2 vs 1.9 ms per 1e4, so n = 4 keeps the warp-per-matrix kernel)
        const int d = plan->d;
        const int threads = 256, groups = threads / d;
        const size_t psmem = groups * project_smem_per_group(d);
        auto kern = (d == 8) ? k_project_packed<8> : k_project_packed<16>;
        if (psmem > 48 * 1024) QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        long blocks = ((long)B + groups - 1) / groups;
        const long cap = (long)num_sms() * 4;
        if (blocks > cap) blocks = cap;
        kern<<<(int)blocks, threads, psmem, st>>>(d, B, h_in, physical, rho);
        QPB_LAUNCHED("k_project_packed");
        return QPB_OK;
    }
    k_lin_project<<<grid, warps * 32, smem, st>>>(plan->n, plan->K, B, plan->LhT, counts, h_in, physical, rho);
    QPB_LAUNCHED("k_lin_project");
    return QPB_OK;
}

int launch_mle_generic(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                       double tol, double* rho, int32_t* iters, cudaStream_t st) {
    const int warps = (plan->n >= 4) ? 4 : 8;
    const size_t smem = warps * mle_smem_per_warp(plan->K, plan->d);
    int rc = enable_smem(k_mle_rrr_generic, smem);
    if (rc != QPB_OK) return rc;
    const int grid = grid_for_warps(B, warps, 4);
    k_mle_rrr_generic<<<grid, warps * 32, smem, st>>>(plan->n, plan->K, B, plan->Ar, plan->ArT, counts, rho0, max_iter,
                                                      tol, rho, iters);
    QPB_LAUNCHED("k_mle_rrr_generic");
    return QPB_OK;
}

int launch_distance(int d, int B, const double* rho, const double* ref, int kind, double* dist, cudaStream_t st) {
    const int warps = (d >= 16) ? 4 : 8;
    const size_t smem = warps * dist_smem_per_warp(d);
    int rc = enable_smem(k_distance, smem);
    if (rc != QPB_OK) return rc;
    const int grid = grid_for_warps(B, warps, 4);
    k_distance<<<grid, warps * 32, smem, st>>>(d, B, rho, ref, kind, dist);
    QPB_LAUNCHED("k_distance");
    return QPB_OK;
}

}  // namespace qpb

using namespace qpb;

extern "C" {

int qpb_state_plan_create(qpb_state_plan** out, int n_qubits, int K, const double* A, const double* L, void* stream) {
    QPB_REQUIRE(out != nullptr, "plan output pointer is NULL");
    QPB_REQUIRE(n_qubits >= 1 && n_qubits <= 4, "n_qubits=%d unsupported (1..4)", n_qubits);
    QPB_REQUIRE(K >= 1 && A != nullptr, "POVM table missing");
    cudaStream_t st = (cudaStream_t)stream;
    qpb_state_plan* p = new qpb_state_plan();
    p->n = n_qubits;
    p->d = 1 << n_qubits;
    p->D = p->d * p->d;
    p->K = K;
    const size_t bytes = sizeof(double) * (size_t)K * p->D;
    int rc = check_cuda(cudaMalloc(&p->Ar, bytes), "cudaMalloc Ar");
    if (rc == QPB_OK) rc = check_cuda(cudaMalloc(&p->ArT, bytes), "cudaMalloc ArT");
    if (rc == QPB_OK && L) rc = check_cuda(cudaMalloc(&p->LhT, bytes), "cudaMalloc LhT");
    if (rc != QPB_OK) {
        qpb_state_plan_destroy(p);
        return rc;
    }
    const long total = (long)K * p->D;
    const int grid = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    k_bloch_to_packed<<<grid, 256, 0, st>>>(A, K, p->D, 1, p->n, 1.0, p->Ar, p->ArT);
    QPB_LAUNCHED("k_bloch_to_packed(A)");
    if (L) {
        // column k of L is the Bloch vector contributed by outcome k; fold the 1/2^n of state.py:197
        k_bloch_to_packed<<<grid, 256, 0, st>>>(L, K, 1, K, p->n, 1.0 / p->d, p->LhT, nullptr);
        QPB_LAUNCHED("k_bloch_to_packed(L)");
    }
    if (p->n <= 2) {  // small tables also travel as kernel parameters (mle_small.cu)
        p->Ar_host = new double[(size_t)K * p->D];
        p->A_host = new double[(size_t)K * p->D];
        rc = check_cuda(cudaMemcpyAsync(p->Ar_host, p->Ar, bytes, cudaMemcpyDeviceToHost, st), "copy Ar to host");
        if (rc == QPB_OK) rc = check_cuda(cudaMemcpyAsync(p->A_host, A, bytes, cudaMemcpyDeviceToHost, st), "copy A to host");
        if (rc == QPB_OK) rc = check_cuda(cudaStreamSynchronize(st), "plan sync");
        if (rc != QPB_OK) {
            qpb_state_plan_destroy(p);
            return rc;
        }
    }
    if (p->n >= 3) {  // structure detection for the axis kernel needs the Bloch-basis table on the host once
        double* tmp = new double[(size_t)K * p->D];
        rc = check_cuda(cudaMemcpyAsync(tmp, A, bytes, cudaMemcpyDeviceToHost, st), "copy A to host");
        if (rc == QPB_OK) rc = check_cuda(cudaStreamSynchronize(st), "plan sync");
        if (rc == QPB_OK) rc = axis_plan_setup(p, tmp);
        delete[] tmp;
        if (rc != QPB_OK) {
            qpb_state_plan_destroy(p);
            return rc;
        }
    }
    *out = p;
    return QPB_OK;
}

int qpb_state_plan_destroy(qpb_state_plan* p) {
    if (!p) return QPB_OK;
    cudaFree(p->axis_slots);
    cudaFree(p->axis_epsp);
    delete[] p->Ar_host;
    delete[] p->A_host;
    cudaFree(p->Ar);
    cudaFree(p->ArT);
    cudaFree(p->LhT);
    delete p;
    return QPB_OK;
}

int qpb_povm_probs(int K, int D, int B, const double* M, const double* r, double scale, int clip, double* p,
                   void* stream) {
    QPB_REQUIRE(K > 0 && D > 0 && B >= 0, "bad shape K=%d D=%d B=%d", K, D, B);
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(M && r && p, "NULL buffer");
    {   // thousands of states: a genuine dense contraction -> FP64 tensor cores
        const int rc = launch_probs_gemm(K, D, B, M, r, scale, clip, p, (cudaStream_t)stream);
        if (rc != QPB_ERR_UNSUPPORTED) return rc;
    }
    const size_t smem = sizeof(double) * PROBS_BT * (size_t)D;
    QPB_REQUIRE(smem <= 48 * 1024, "D=%d too large", D);
    const int grid = (B + PROBS_BT - 1) / PROBS_BT;
    k_povm_probs<<<grid, 256, smem, (cudaStream_t)stream>>>(K, D, B, M, r, scale, clip, p);
    QPB_LAUNCHED("k_povm_probs");
    return QPB_OK;
}

int qpb_lin_project(const qpb_state_plan* plan, int B, const int32_t* counts, int physical, double* rho, void* stream) {
    QPB_REQUIRE(plan && plan->LhT, "plan has no linear-inversion table (L was NULL)");
    QPB_REQUIRE(B >= 0, "negative batch");
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(counts && rho, "NULL buffer");
    int rc = option(QPB_OPT_NO_LIN_SMALL) ? QPB_ERR_UNSUPPORTED
                                        : launch_lin_project_small(plan, B, counts, physical, rho, (cudaStream_t)stream);
    if (rc == QPB_ERR_UNSUPPORTED) rc = launch_lin_project(plan, B, counts, physical, rho, (cudaStream_t)stream);
    return rc;
}

int qpb_distance(int dd, int B, const double* rho, const double* ref, int kind, double* dist, void* stream) {
    QPB_REQUIRE(dd == 2 || dd == 4 || dd == 8 || dd == 16, "matrix side %d unsupported", dd);
    QPB_REQUIRE(kind >= 0 && kind <= 2, "unknown distance kind %d", kind);
    QPB_REQUIRE(B >= 0, "negative batch");
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(rho && ref && dist, "NULL buffer");
    return launch_distance(dd, B, rho, ref, kind, dist, (cudaStream_t)stream);
}

}  // extern "C"
