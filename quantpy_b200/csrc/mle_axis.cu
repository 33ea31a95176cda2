// Structured R.rho.R maximum likelihood for n = 3, 4 qubits and Pauli-axis POVMs ('proj', 'proj-set',
// 'proj4', any shot weights): ONE CTA (2 or 4 warps) PER SAMPLE, all state in shared memory.
//
// Every effect is E_k = c_k (x)_q (1 + s_q sigma_{a_q}), a_q in {X,Y,Z}.  With S_i = Tr(sigma_i rho)
// (i = base-4 Pauli string) the probabilities follow from n successive per-qubit "axis" maps
//      digit in {I,X,Y,Z}  ->  slot in {x+,x-,y+,y-,z+,z-} :  out = in[I] +- in[axis]
// and R = sum_k w_k (x)_q (1 + s_q sigma_{a_q}) from the n adjoint maps, so the K x D contraction
// (1296 x 256 at n = 4) costs 2 * sum_q 6^q 4^(n-q) additions instead of 2*K*D multiply-adds, and nothing
// is streamed from L2.  S and R move between the matrix and the Pauli basis by in-place radix-4
// butterflies over the qubits (fast Pauli transform, n stages).  R rho R is two complex d x d products
// out of shared memory.  Effects are addressed by the canonical base-6 slot of their axes/signs.
#include <cmath>
#include <cstdlib>

#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "plan.h"

namespace qpb {

__host__ __device__ constexpr int ipow(int b, int e) { return e == 0 ? 1 : b * ipow(b, e - 1); }

template <int N>
struct Axis {
    static constexpr int d = 1 << N;
    static constexpr int D = d * d;          // 4^N
    static constexpr int K6 = ipow(6, N);    // canonical slots
    static constexpr int ld = d + 1;         // padded leading dimension of the complex matrices (bank conflicts)
    // n = 4 keeps the complex matrices as separate re / im planes with a leading dimension of 20 doubles: the DMMA
    // fragments of the two products (8 rows x 4 consecutive doubles per quarter-warp) then load in two wavefronts;
    // n = 3 keeps interleaved complex numbers (its products stay on the FMA pipe).
    static constexpr bool planar = (N == 4);
    static constexpr int lp = 20;
    static constexpr int mat = planar ? 2 * d * lp : 2 * d * ld;  // doubles per matrix
    // The two contraction buffers hold at most [6^(N-1)][4] reals: the last axis stage, the reciprocals and the
    // first adjoint stage are fused in registers, so the 6^N probabilities are never stored.  The transform /
    // product workspaces alias them (their lifetimes do not overlap), the counts stay int32.
    static constexpr int buf = ipow(6, N - 1) * 4;
    static_assert(mat <= buf, "matrix workspaces must fit in a contraction buffer");
    // bytes of shared memory per sample: rho, R, two buffers, counts
    static constexpr size_t smem_bytes = sizeof(double) * (2 * mat + 2 * buf) + sizeof(int) * K6;
    // resident CTAs per SM the launch bounds ask for (register cap), CTA size
    static constexpr int GS = (N == 4) ? 128 : 64;
    static constexpr int min_blocks = (N == 4) ? 7 : 16;
};

// d x d complex matrix in shared memory: interleaved (re, im) with leading dimension ld, or two planes
template <int N>
struct Mat {
    double* p;
    __device__ __forceinline__ double2 get(int a, int b) const {
        if constexpr (Axis<N>::planar) {
            return make_double2(p[a * Axis<N>::lp + b], p[Axis<N>::d * Axis<N>::lp + a * Axis<N>::lp + b]);
        } else {
            return reinterpret_cast<const double2*>(p)[a * Axis<N>::ld + b];
        }
    }
    __device__ __forceinline__ void set(int a, int b, double2 z) const {
        if constexpr (Axis<N>::planar) {
            p[a * Axis<N>::lp + b] = z.x;
            p[Axis<N>::d * Axis<N>::lp + a * Axis<N>::lp + b] = z.y;
        } else {
            reinterpret_cast<double2*>(p)[a * Axis<N>::ld + b] = z;
        }
    }
    __device__ __forceinline__ const double* re() const { return p; }
    __device__ __forceinline__ const double* im() const { return p + Axis<N>::d * Axis<N>::lp; }
};

// One sample is processed by a group of GS threads: a warp (GS = 32) or a whole CTA (GS = blockDim.x).
template <int GS>
__device__ __forceinline__ void gsync() {
    if constexpr (GS == 32) __syncwarp();
    else __syncthreads();
}
template <int GS>
__device__ __forceinline__ double gsum(double v, double* red, int tid) {
    v = warp_sum(v);
    if constexpr (GS == 32) {
        return v;
    } else {
        __syncthreads();  // red may still be read from the previous reduction
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < GS / 32; ++w) s += red[w];
        return s;
    }
}

// index of matrix element (a, b) in the digit-interleaved order: digit_q = 2 a_q + b_q, first qubit most significant
template <int N>
__device__ __forceinline__ int interleave(int a, int b) {
    int e = 0;
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const int sh = N - 1 - q;
        e = e * 4 + 2 * ((a >> sh) & 1) + ((b >> sh) & 1);
    }
    return e;
}

// Position of transform element i in the workspace: a GF(2)-linear bank swizzle of the low three index bits
// (b0 ^= b4, b1 ^= b3, b2 ^= b4) that makes the eight lanes of a quarter-warp hit eight different 16-byte bank
// groups in every radix-4 stage (lanes differ in index bits {0,1,2}, {0,1,4} or {2,3,4} depending on the stride);
// the plain layout cost 2x (stride 4) and 4x (stride 1) the wavefronts.
__device__ __forceinline__ int tw(int i) { return i ^ (((i >> 4) & 1) * 5) ^ (((i >> 3) & 1) << 1); }

// (row, column) of the matrix element stored at digit-interleaved position e
template <int N>
__device__ __forceinline__ void deinterleave(int e, int& a, int& b) {
    a = 0;
    b = 0;
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const int dig = (e >> (2 * (N - 1 - q))) & 3;
        a = a * 2 + (dig >> 1);
        b = b * 2 + (dig & 1);
    }
}

// Forward fast Pauli transform S_i = Tr(sigma_i M), one radix-4 stage per qubit.  The first stage reads the matrix
// M (row-major, leading dimension ld) directly, the middle stages run in place on the swizzled workspace W, the
// last one writes only what is needed afterwards: the real parts, in natural order, to S.
template <int N, int GS>
__device__ __forceinline__ void pauli_forward(const Mat<N> M, double2* __restrict__ W, double* __restrict__ S,
                                              int lane) {
    constexpr int D = Axis<N>::D;
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const int stride = 1 << (2 * (N - 1 - q));
        for (int g = lane; g < D / 4; g += GS) {
            const int idx = (g / stride) * 4 * stride + (g % stride);
            double2 v0, v1, v2, v3;
            if (q == 0) {
                int al, bl;
                deinterleave<N>(idx, al, bl);  // low bits of (row, column); the first qubit's digit selects the block
                constexpr int H = 1 << (N - 1);
                v0 = M.get(al, bl);
                v1 = M.get(al, bl + H);
                v2 = M.get(al + H, bl);
                v3 = M.get(al + H, bl + H);
            } else {
                v0 = W[tw(idx)];
                v1 = W[tw(idx + stride)];
                v2 = W[tw(idx + 2 * stride)];
                v3 = W[tw(idx + 3 * stride)];
            }
            if (q == N - 1) {
                S[idx] = v0.x + v3.x;        // I
                S[idx + 1] = v1.x + v2.x;    // X
                S[idx + 2] = v2.y - v1.y;    // Y = Re i (v1 - v2)
                S[idx + 3] = v0.x - v3.x;    // Z
            } else {
                W[tw(idx)] = make_double2(v0.x + v3.x, v0.y + v3.y);                     // I
                W[tw(idx + stride)] = make_double2(v1.x + v2.x, v1.y + v2.y);            // X
                W[tw(idx + 2 * stride)] = make_double2(-(v1.y - v2.y), v1.x - v2.x);     // Y = i (v1 - v2)
                W[tw(idx + 3 * stride)] = make_double2(v0.x - v3.x, v0.y - v3.y);        // Z
            }
        }
        gsync<GS>();
    }
}

// Inverse: real Pauli coefficients g (natural order) -> M = sum_i g_i sigma_i (row-major, leading dimension ld).
// g may live in the buffer W aliases, so the first stage reads its quad into registers before anything is written.
template <int N, int GS>
__device__ __forceinline__ void pauli_inverse(const double* __restrict__ gcoef, double2* __restrict__ W,
                                              const Mat<N> M, int lane) {
    constexpr int D = Axis<N>::D;
    static_assert(D / 4 <= GS, "one quad per thread in the first stage");
    {
        constexpr int stride = 1 << (2 * (N - 1));
        double gi = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
        if (lane < D / 4) {
            gi = gcoef[lane];
            gx = gcoef[lane + stride];
            gy = gcoef[lane + 2 * stride];
            gz = gcoef[lane + 3 * stride];
        }
        gsync<GS>();
        if (lane < D / 4) {
            W[tw(lane)] = make_double2(gi + gz, 0.0);                // (0,0)
            W[tw(lane + stride)] = make_double2(gx, -gy);            // (0,1) = gx - i gy
            W[tw(lane + 2 * stride)] = make_double2(gx, gy);         // (1,0) = gx + i gy
            W[tw(lane + 3 * stride)] = make_double2(gi - gz, 0.0);   // (1,1)
        }
        gsync<GS>();
    }
#pragma unroll
    for (int q = 1; q < N; ++q) {
        const int stride = 1 << (2 * (N - 1 - q));
        for (int g = lane; g < D / 4; g += GS) {
            const int idx = (g / stride) * 4 * stride + (g % stride);
            const double2 gi = W[tw(idx)], gx = W[tw(idx + stride)], gy = W[tw(idx + 2 * stride)],
                          gz = W[tw(idx + 3 * stride)];
            const double2 m00 = make_double2(gi.x + gz.x, gi.y + gz.y);
            const double2 m01 = make_double2(gx.x + gy.y, gx.y - gy.x);  // gx - i gy
            const double2 m10 = make_double2(gx.x - gy.y, gx.y + gy.x);  // gx + i gy
            const double2 m11 = make_double2(gi.x - gz.x, gi.y - gz.y);
            if (q == N - 1) {  // idx = 4 * (digits of the other qubits): the last qubit is the lowest bit of row and column
                int a, b;
                deinterleave<N>(idx, a, b);
                M.set(a, b, m00);
                M.set(a, b + 1, m01);
                M.set(a + 1, b, m10);
                M.set(a + 1, b + 1, m11);
            } else {
                W[tw(idx)] = m00;
                W[tw(idx + stride)] = m01;
                W[tw(idx + 2 * stride)] = m10;
                W[tw(idx + 3 * stride)] = m11;
            }
        }
        gsync<GS>();
    }
}

// stage Q of the axis map: in [6^Q][4][4^(N-1-Q)] -> out [6^Q][6][4^(N-1-Q)]; one thread per (prefix, suffix)
// loads the four digits once and writes the six slots
template <int N, int Q, int GS>
__device__ __forceinline__ void axis_forward(const double* __restrict__ in, double* __restrict__ out, int lane) {
    constexpr int P4 = ipow(4, N - 1 - Q), PRE = ipow(6, Q);
    for (int t = lane; t < PRE * P4; t += GS) {
        const int post = t % P4, pre = t / P4;
        const double* src = in + (pre * 4) * P4 + post;
        double* dst = out + (pre * 6) * P4 + post;
        const double x0 = src[0];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double xa = src[(a + 1) * P4];
            dst[(2 * a) * P4] = x0 + xa;
            dst[(2 * a + 1) * P4] = x0 - xa;
        }
    }
    gsync<GS>();
}

// adjoint of stage Q: in [6^Q][6][4^(N-1-Q)] -> out [6^Q][4][4^(N-1-Q)]
template <int N, int Q, int GS>
__device__ __forceinline__ void axis_adjoint(const double* __restrict__ in, double* __restrict__ out, int lane) {
    constexpr int P4 = ipow(4, N - 1 - Q), PRE = ipow(6, Q);
    for (int t = lane; t < PRE * P4; t += GS) {
        const int post = t % P4, pre = t / P4;
        const double* src = in + (pre * 6) * P4 + post;
        double* dst = out + (pre * 4) * P4 + post;
        const double s0 = src[0], s1 = src[P4], s2 = src[2 * P4], s3 = src[3 * P4], s4 = src[4 * P4], s5 = src[5 * P4];
        dst[0] = ((s0 + s1) + (s2 + s3)) + (s4 + s5);
        dst[P4] = s0 - s1;
        dst[2 * P4] = s2 - s3;
        dst[3 * P4] = s4 - s5;
    }
    gsync<GS>();
}

// stages 0 .. N-2 (the last one is fused with the reciprocals, below)
template <int N, int GS, int Q = 0>
__device__ __forceinline__ double* axis_forward_all(double* a, double* b, int lane) {
    if constexpr (Q == N - 1) {
        return a;  // result lives in `a`
    } else {
        axis_forward<N, Q, GS>(a, b, lane);
        return axis_forward_all<N, GS, Q + 1>(b, a, lane);
    }
}
template <int N, int GS, int Q = N - 2>
__device__ __forceinline__ double* axis_adjoint_all(double* a, double* b, int lane) {
    if constexpr (Q < 0) {
        return a;
    } else {
        axis_adjoint<N, Q, GS>(a, b, lane);
        return axis_adjoint_all<N, GS, Q - 1>(b, a, lane);
    }
}

// Last axis stage, weights and first adjoint stage in one pass over [6^(N-1)] prefixes:
//   p(pre, a, +-) = in[pre][I] +- in[pre][a];  w = f / (p + eps);  out[pre][I] = sum w,  out[pre][a] = w+ - w-
template <int N, int GS>
__device__ __forceinline__ void axis_last_fused(const double* __restrict__ in, double* __restrict__ out,
                                                const int* __restrict__ cnt, const double* __restrict__ epsp,
                                                double inv_total, int lane) {
    constexpr int PRE = ipow(6, N - 1);
    const double2* in2 = reinterpret_cast<const double2*>(in);  // buffers are 16-byte aligned, rows are 32 bytes
    double2* out2 = reinterpret_cast<double2*>(out);
    for (int pre = lane; pre < PRE; pre += GS) {
        const double2 lo = in2[pre * 2], hi = in2[pre * 2 + 1];
        const double x0 = lo.x, xs[3] = {lo.y, hi.x, hi.y};
        double u0 = 0.0, dif[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int slot = pre * 6 + 2 * a;
            const double wp = ((double)cnt[slot] * inv_total) * fast_recip((x0 + xs[a]) + epsp[slot]);
            const double wm = ((double)cnt[slot + 1] * inv_total) * fast_recip((x0 - xs[a]) + epsp[slot + 1]);
            u0 += wp + wm;
            dif[a] = wp - wm;
        }
        out2[pre * 2] = make_double2(u0, dif[0]);
        out2[pre * 2 + 1] = make_double2(dif[1], dif[2]);
    }
    gsync<GS>();
}

__device__ __forceinline__ void dmma_884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// C = X * Y for d x d complex matrices in shared memory, Y HERMITIAN (rho or R in both uses).
//
// n = 4: the kernel is bound by shared-memory wavefronts and the products were their largest source, so they run
// on the FP64 tensor cores: warp w owns the 8 x 8 output tile (w / 2, w % 2) and per k-step of 4 issues four
// mma.m8n8k4 (re += Xr Yr + Xi Yi^T..., see below) from four 64-bit fragment loads.  The B fragment is read
// from the TRANSPOSED position with the imaginary part negated (Y[k][n] = conj(Y[n][k])), which gives it the
// same conflict-free address pattern as the A fragment: rows 8 apart in tile, 4 consecutive doubles, lp = 20.
//     C_re = Xr Br - Xi Bi,  C_im = Xr Bi + Xi Br,   Br[k][n] = Yr[n][k],  Bi[k][n] = -Yi[n][k]
// n = 3: 2 x TJ register tiles on the FMA pipe out of interleaved storage (quarter-warps share the column block).
template <int N, int GS>
__device__ __forceinline__ void cmatmul(const Mat<N> C, const Mat<N> X, const Mat<N> Y, int lane) {
    constexpr int d = Axis<N>::d;
    if constexpr (Axis<N>::planar) {
        constexpr int lp = Axis<N>::lp;
        static_assert(d == 16 && GS == 128, "four warps, one 8 x 8 tile each");
        const int w = lane >> 5, l = lane & 31;
        const int ti = w >> 1, tj = w & 1, fr = l >> 2, fc = l & 3;
        const double* xr = X.re() + (8 * ti + fr) * lp + fc;
        const double* xi = X.im() + (8 * ti + fr) * lp + fc;
        const double* yr = Y.re() + (8 * tj + fr) * lp + fc;
        const double* yi = Y.im() + (8 * tj + fr) * lp + fc;
        double cr0 = 0.0, cr1 = 0.0, ci0 = 0.0, ci1 = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < d; k0 += 4) {
            const double ar = xr[k0], ai = xi[k0], br = yr[k0], byi = yi[k0];
            dmma_884(cr0, cr1, ar, br);    // Xr Br
            dmma_884(cr0, cr1, ai, byi);   // -Xi Bi = Xi Yi^T
            dmma_884(ci0, ci1, ai, br);    // Xi Br
            dmma_884(ci0, ci1, ar, -byi);  // Xr Bi
        }
        double* cre = C.p + (8 * ti + fr) * lp + 8 * tj + 2 * fc;
        *reinterpret_cast<double2*>(cre) = make_double2(cr0, cr1);
        *reinterpret_cast<double2*>(cre + d * lp) = make_double2(ci0, ci1);
    } else {
        constexpr int ld = Axis<N>::ld;
        const double2* Xp = reinterpret_cast<const double2*>(X.p);
        const double2* Yp = reinterpret_cast<const double2*>(Y.p);
        double2* Cp = reinterpret_cast<double2*>(C.p);
        constexpr int TI = 1, TJ = 2, ROWS = d / TI;
        constexpr int TASKS = ROWS * (d / TJ);
        for (int t = lane; t < TASKS; t += GS) {
            const int a = t % ROWS, b0 = (t / ROWS) * TJ;
            double re[TI][TJ], im[TI][TJ];
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) re[i][j] = im[i][j] = 0.0;
#pragma unroll 2
            for (int c = 0; c < d; ++c) {
                double2 x[TI];
#pragma unroll
                for (int i = 0; i < TI; ++i) x[i] = Xp[(a + i * ROWS) * ld + c];
#pragma unroll
                for (int j = 0; j < TJ; ++j) {
                    const double2 y = Yp[c * ld + b0 + j];
#pragma unroll
                    for (int i = 0; i < TI; ++i) {
                        re[i][j] = fma(x[i].x, y.x, re[i][j]);
                        re[i][j] = fma(-x[i].y, y.y, re[i][j]);
                        im[i][j] = fma(x[i].x, y.y, im[i][j]);
                        im[i][j] = fma(x[i].y, y.x, im[i][j]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) Cp[(a + i * ROWS) * ld + b0 + j] = make_double2(re[i][j], im[i][j]);
        }
    }
    gsync<GS>();
}

template <int N, int GS>
__global__ void __launch_bounds__(GS, Axis<N>::min_blocks)
k_mle_rrr_axis(int K, int B, const int* __restrict__ slot_of_col, const double* __restrict__ epsp,
               const int32_t* __restrict__ counts, const double* __restrict__ rho0, int max_iter, double tol,
               double* __restrict__ rho_out, int32_t* __restrict__ iters, unsigned int* __restrict__ queue) {
    constexpr int d = Axis<N>::d, D = Axis<N>::D, K6 = Axis<N>::K6, MAT = Axis<N>::mat;
    extern __shared__ __align__(16) double smd[];
    __shared__ double red[32];
    __shared__ unsigned int next_sample;
    const int lane = threadIdx.x;  // index within the group (= CTA)
    constexpr int BUF = Axis<N>::buf;
    const Mat<N> rho{smd}, Rm{smd + MAT};
    double* bufA = smd + 2 * MAT;
    double* bufB = bufA + BUF;
    int* cnt = reinterpret_cast<int*>(bufB + BUF);     // counts in canonical slot order
    double2* W = reinterpret_cast<double2*>(bufB);     // transform workspace (dense D, swizzled): aliases bufB
    const Mat<N> P1{bufB}, T{bufA};                    // R rho and R rho R: alias the contraction buffers

    for (;;) {
        __syncthreads();
        if (lane == 0) next_sample = atomicAdd(queue, 1u);
        __syncthreads();
        const unsigned int b = next_sample;
        if (b >= (unsigned)B) break;
        // ---- load the sample: frequencies into canonical slots, start state
        const int32_t* c = counts + (size_t)b * K;
        long long tot_i = 0;
        for (int k = lane; k < K; k += GS) tot_i += c[k];
        const double total = gsum<GS>((double)tot_i, red, lane);
        const double inv_total = 1.0 / total;
        for (int e = lane; e < K6; e += GS) cnt[e] = 0;
        gsync<GS>();
        for (int k = lane; k < K; k += GS) cnt[slot_of_col[k]] = c[k];
        if (rho0) {
            const double2* r0 = reinterpret_cast<const double2*>(rho0) + (size_t)b * D;
            for (int e = lane; e < D; e += GS) {  // Hermitian part of the start, as the packed kernels take it
                const int a = e / d, bb = e % d;
                const double2 z = r0[e], zt = r0[bb * d + a];
                rho.set(a, bb, (a == bb) ? make_double2(z.x, 0.0) : (a < bb ? z : make_double2(zt.x, -zt.y)));
            }
        } else {
            for (int e = lane; e < D; e += GS) rho.set(e / d, e % d, make_double2((e / d == e % d) ? 1.0 / d : 0.0, 0.0));
        }
        gsync<GS>();
        int it = 0;
        for (it = 1; it <= max_iter; ++it) {
            // S_i = Tr(sigma_i rho), real parts straight into bufA
            pauli_forward<N, GS>(rho, W, bufA, lane);
            double* q = axis_forward_all<N, GS>(bufA, bufB, lane);  // after N-1 stages: [6^(N-1)][4]
            double* u = q == bufA ? bufB : bufA;
            axis_last_fused<N, GS>(q, u, cnt, epsp, inv_total, lane);
            double* g = axis_adjoint_all<N, GS>(u, q, lane);  // Pauli coefficients of R
            pauli_inverse<N, GS>(g, W, Rm, lane);
            cmatmul<N, GS>(P1, Rm, rho, lane);  // P1 = R rho
            cmatmul<N, GS>(T, P1, Rm, lane);    // T = R rho R
            // Hermitise, normalise, step norm
            double tr = 0.0;
            for (int a = lane; a < d; a += GS) tr += T.get(a, a).x;
            tr = gsum<GS>(tr, red, lane);
            const double inv = 1.0 / tr;
            // one thread per element of the upper triangle: reads T(a,b), T(b,a) and the old rho(a,b), writes the new
            // rho(a,b) and its conjugate in place (nobody else touches the pair), counts the pair's step norm twice
            double del = 0.0;
            for (int e = lane; e < D; e += GS) {
                const int a = e / d, bb = e % d;
                if (a > bb) continue;
                const double2 old = rho.get(a, bb);
                if (a == bb) {
                    const double v = T.get(a, a).x * inv;
                    const double dr = v - old.x;
                    del += dr * dr;
                    rho.set(a, a, make_double2(v, 0.0));
                } else {
                    const double2 z = T.get(a, bb), zt = T.get(bb, a);
                    const double2 v = make_double2(0.5 * (z.x + zt.x) * inv, 0.5 * (z.y - zt.y) * inv);
                    const double dr = v.x - old.x, di = v.y - old.y;
                    del += 2.0 * (dr * dr + di * di);
                    rho.set(a, bb, v);
                    rho.set(bb, a, make_double2(v.x, -v.y));
                }
            }
            del = sqrt(gsum<GS>(del, red, lane));
            gsync<GS>();
            if (del < tol) break;
        }
        if (it > max_iter) it = max_iter;
        double2* out = reinterpret_cast<double2*>(rho_out) + (size_t)b * D;
        for (int e = lane; e < D; e += GS) out[e] = rho.get(e / d, e % d);
        if (iters && lane == 0) iters[b] = it;
    }
}

// Recognise an n-qubit Pauli-axis POVM from the Bloch-basis table A [K][4^n] (host); fills slots and eps/c.
static bool detect_axis(const double* A, int n, int K, int* slot_of_col, double* epsp) {
    const int D = 1 << (2 * n);
    int K6 = 1;
    for (int q = 0; q < n; ++q) K6 *= 6;
    if (K < 1 || K > K6) return false;
    for (int s = 0; s < K6; ++s) epsp[s] = 1.0;
    bool* used = new bool[K6]();
    bool ok = true;
    for (int k = 0; k < K && ok; ++k) {
        const double* r = A + (size_t)k * D;
        const double c = r[0];
        if (!(c > 0.0)) { ok = false; break; }
        int axis[8], sign[8], slot = 0;
        for (int q = 0; q < n && ok; ++q) {
            const int stride = 1 << (2 * (n - 1 - q));
            axis[q] = 0;
            for (int dgt = 1; dgt < 4; ++dgt) {
                const double v = r[dgt * stride];
                if (v != 0.0) {
                    if (axis[q] || fabs(fabs(v) - c) > 1e-13 * c) ok = false;
                    axis[q] = dgt;
                    sign[q] = v > 0 ? 1 : -1;
                }
            }
            if (!axis[q]) ok = false;
            if (ok) slot = slot * 6 + 2 * (axis[q] - 1) + (sign[q] < 0);
        }
        for (int i = 0; i < D && ok; ++i) {
            double want = c;
            for (int q = 0; q < n; ++q) {
                const int dgt = (i >> (2 * (n - 1 - q))) & 3;
                if (dgt == 0) continue;
                if (dgt == axis[q]) want *= sign[q];
                else { want = 0.0; break; }
            }
            if (fabs(r[i] - want) > 1e-13 * c) ok = false;
        }
        if (ok && used[slot]) ok = false;
        if (ok) {
            used[slot] = true;
            slot_of_col[k] = slot;
            epsp[slot] = kLogGuard / c;
        }
    }
    delete[] used;
    return ok;
}

template <int N>
static int launch_axis(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                       double tol, double* rho, int32_t* iters, cudaStream_t st) {
    // one CTA of GS threads per sample; shared memory (27.7 KB at n = 4) and the register cap of the launch bounds
    // allow 7 resident CTAs per SM at n = 4, 16 at n = 3
    constexpr int GS = Axis<N>::GS;
    const size_t smem = Axis<N>::smem_bytes;
    QPB_REQUIRE(smem <= 227 * 1024, "axis kernel needs %zu bytes of shared memory", smem);
    auto kern = k_mle_rrr_axis<N, GS>;
    if (smem > 48 * 1024) QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    QPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GS, smem));
    if (per_sm < 1) per_sm = 1;
    long blocks = (long)num_sms() * per_sm;
    if (blocks > B) blocks = B;
    unsigned int* queue = static_cast<unsigned int*>(scratch(st, 0, sizeof(unsigned int)));
    if (!queue) return QPB_ERR_NOMEM;
    QPB_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned int), st));
    kern<<<(int)blocks, GS, smem, st>>>(plan->K, B, plan->axis_slots, plan->axis_epsp, counts, rho0, max_iter,
                                                tol, rho, iters, queue);
    QPB_LAUNCHED("k_mle_rrr_axis");
    return QPB_OK;
}

// Called once at plan creation (n = 3, 4): detect the structure and upload slot map / guards.
int axis_plan_setup(qpb_state_plan* plan, const double* A_host) {
    if (plan->n < 3 || plan->n > 4) return QPB_OK;
    int K6 = 1;
    for (int q = 0; q < plan->n; ++q) K6 *= 6;
    int* slots = new int[plan->K];
    double* epsp = new double[K6];
    int rc = QPB_OK;
    if (detect_axis(A_host, plan->n, plan->K, slots, epsp)) {
        rc = check_cuda(cudaMalloc(&plan->axis_slots, sizeof(int) * plan->K), "cudaMalloc axis_slots");
        if (rc == QPB_OK) rc = check_cuda(cudaMalloc(&plan->axis_epsp, sizeof(double) * K6), "cudaMalloc axis_epsp");
        if (rc == QPB_OK) rc = check_cuda(cudaMemcpy(plan->axis_slots, slots, sizeof(int) * plan->K, cudaMemcpyHostToDevice), "copy slots");
        if (rc == QPB_OK) rc = check_cuda(cudaMemcpy(plan->axis_epsp, epsp, sizeof(double) * K6, cudaMemcpyHostToDevice), "copy epsp");
        if (rc == QPB_OK) plan->axis_ok = true;
    }
    delete[] slots;
    delete[] epsp;
    return rc;
}

int launch_mle_axis(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                    double tol, double* rho, int32_t* iters, cudaStream_t st) {
    if (!plan->axis_ok || option(QPB_OPT_NO_AXIS_KERNEL)) return QPB_ERR_UNSUPPORTED;
    if (plan->n == 3) return launch_axis<3>(plan, B, counts, rho0, max_iter, tol, rho, iters, st);
    if (plan->n == 4) return launch_axis<4>(plan, B, counts, rho0, max_iter, tol, rho, iters, st);
    return QPB_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// The per-sample half of the batched general-POVM iteration (mle_tiled.cu): rho' = R rho R / Tr from the packed R
// rows the GEMM produced, with the same matrix helpers as the axis kernel (DMMA products at n = 4).  One CTA of
// Axis<N>::GS threads per running sample; frozen (converged) samples are skipped until the row map drops them.
// ------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(Axis<N>::GS)
k_rrr_update_mat(int M, const int* __restrict__ map, const double* __restrict__ R, double* __restrict__ H,
                 double* __restrict__ H2, int* __restrict__ its, int* __restrict__ done, int max_iter, double tol) {
    constexpr int d = Axis<N>::d, D = Axis<N>::D, MAT = Axis<N>::mat, GS = Axis<N>::GS;
    extern __shared__ __align__(16) double smd[];
    __shared__ double red[32];
    const int lane = threadIdx.x;
    const Mat<N> rho{smd}, Rm{smd + MAT}, P1{smd + 2 * MAT}, T{smd + 3 * MAT};
    for (long a = blockIdx.x; a < M; a += gridDim.x) {
        const int b = map[a];
        if (done[b]) continue;  // uniform over the CTA
        __syncthreads();
        const double* Rp = R + a * D;
        double* Hp = H + (long)b * D;
        double* H2p = H2 + (long)b * D;
        for (int e = lane; e < D; e += GS) {  // packed Hermitian -> full complex matrices
            const int r = e / d, c = e % d;
            if (r > c) continue;
            const double rr = Rp[r * d + c], ri = (r == c) ? 0.0 : Rp[c * d + r];
            const double hr = Hp[r * d + c], hi = (r == c) ? 0.0 : Hp[c * d + r];
            Rm.set(r, c, make_double2(rr, ri));
            rho.set(r, c, make_double2(hr, hi));
            if (r != c) {
                Rm.set(c, r, make_double2(rr, -ri));
                rho.set(c, r, make_double2(hr, -hi));
            }
        }
        gsync<GS>();
        cmatmul<N, GS>(P1, Rm, rho, lane);  // R rho
        cmatmul<N, GS>(T, P1, Rm, lane);    // R rho R
        double tr = 0.0;
        for (int x = lane; x < d; x += GS) tr += T.get(x, x).x;
        tr = gsum<GS>(tr, red, lane);
        const double inv = 1.0 / tr;
        double del = 0.0;
        for (int e = lane; e < D; e += GS) {
            const int r = e / d, c = e % d;
            if (r > c) continue;
            const double2 old = rho.get(r, c);
            if (r == c) {
                const double v = T.get(r, r).x * inv;
                const double dr = v - old.x;
                del += dr * dr;
                Hp[r * d + r] = v;
                H2p[r * d + r] = v;
            } else {
                const double2 z = T.get(r, c), zt = T.get(c, r);
                const double vr = 0.5 * (z.x + zt.x) * inv, vi = 0.5 * (z.y - zt.y) * inv;
                const double dr = vr - old.x, di = vi - old.y;
                del += 2.0 * (dr * dr + di * di);
                Hp[r * d + c] = vr;
                Hp[c * d + r] = vi;
                H2p[r * d + c] = 2.0 * vr;
                H2p[c * d + r] = 2.0 * vi;
            }
        }
        del = sqrt(gsum<GS>(del, red, lane));
        if (lane == 0) {
            const int it = its[b] + 1;
            its[b] = it;
            if (del < tol || it >= max_iter) done[b] = 1;
        }
    }
}

template <int N>
static int launch_update_mat(int M, const int* map, const double* R, double* H, double* H2, int* its, int* done,
                             int max_iter, double tol, cudaStream_t st) {
    const size_t smem = sizeof(double) * 4 * Axis<N>::mat;
    long blocks = (long)num_sms() * (N == 4 ? 8 : 16);
    if (blocks > M) blocks = M;
    k_rrr_update_mat<N><<<(int)blocks, Axis<N>::GS, smem, st>>>(M, map, R, H, H2, its, done, max_iter, tol);
    QPB_LAUNCHED("k_rrr_update_mat");
    return QPB_OK;
}

int launch_rrr_update_mat(int n, int M, const int* map, const double* R, double* H, double* H2, int* its, int* done,
                          int max_iter, double tol, cudaStream_t st) {
    if (M <= 0) return QPB_OK;
    if (n == 3) return launch_update_mat<3>(M, map, R, H, H2, its, done, max_iter, tol, st);
    if (n == 4) return launch_update_mat<4>(M, map, R, H, H2, its, done, max_iter, tol, st);
    return QPB_ERR_UNSUPPORTED;
}

}  // namespace qpb
