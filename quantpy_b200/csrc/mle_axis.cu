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
    // The two contraction buffers hold at most [6^(N-1)][4] reals: the last axis stage, the reciprocals and the
    // first adjoint stage are fused in registers, so the 6^N probabilities are never stored.  The transform /
    // product workspaces alias them (their lifetimes do not overlap), the counts stay int32.
    static constexpr int buf = ipow(6, N - 1) * 4;
    static_assert(2 * d * ld <= buf, "matrix workspaces must fit in a contraction buffer");
    // bytes of shared memory per sample: rho, R (complex d x ld), two buffers, counts
    static constexpr size_t smem_bytes = sizeof(double) * (2 * 2 * d * ld + 2 * buf) + sizeof(int) * K6;
    // resident CTAs per SM the launch bounds ask for (register cap), CTA size
    static constexpr int GS = (N == 4) ? 128 : 64;
    static constexpr int min_blocks = (N == 4) ? 7 : 16;
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

// forward fast Pauli transform, in place on W (complex, digit-interleaved, swizzled by tw): W[i] <- Tr(sigma_i M)
template <int N, int GS>
__device__ __forceinline__ void pauli_forward(double2* __restrict__ W, int lane) {
    constexpr int D = Axis<N>::D;
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const int stride = 1 << (2 * (N - 1 - q));
        for (int g = lane; g < D / 4; g += GS) {
            const int idx = (g / stride) * 4 * stride + (g % stride);
            const int i0 = tw(idx), i1 = tw(idx + stride), i2 = tw(idx + 2 * stride), i3 = tw(idx + 3 * stride);
            const double2 v0 = W[i0], v1 = W[i1], v2 = W[i2], v3 = W[i3];
            W[i0] = make_double2(v0.x + v3.x, v0.y + v3.y);          // I
            W[i1] = make_double2(v1.x + v2.x, v1.y + v2.y);          // X
            W[i2] = make_double2(-(v1.y - v2.y), v1.x - v2.x);       // Y = i (v1 - v2)
            W[i3] = make_double2(v0.x - v3.x, v0.y - v3.y);          // Z
        }
        gsync<GS>();
    }
}

// inverse: W[i] = g_i (complex) -> W[e(a,b)] = (sum_i g_i sigma_i)[a][b]
template <int N, int GS>
__device__ __forceinline__ void pauli_inverse(double2* __restrict__ W, int lane) {
    constexpr int D = Axis<N>::D;
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const int stride = 1 << (2 * (N - 1 - q));
        for (int g = lane; g < D / 4; g += GS) {
            const int idx = (g / stride) * 4 * stride + (g % stride);
            const int i0 = tw(idx), i1 = tw(idx + stride), i2 = tw(idx + 2 * stride), i3 = tw(idx + 3 * stride);
            const double2 gi = W[i0], gx = W[i1], gy = W[i2], gz = W[i3];
            W[i0] = make_double2(gi.x + gz.x, gi.y + gz.y);          // (0,0)
            W[i1] = make_double2(gx.x + gy.y, gx.y - gy.x);          // (0,1) = gx - i gy
            W[i2] = make_double2(gx.x - gy.y, gx.y + gy.x);          // (1,0) = gx + i gy
            W[i3] = make_double2(gi.x - gz.x, gi.y - gz.y);          // (1,1)
        }
        gsync<GS>();
    }
}

// stage Q of the axis map: in [6^Q][4][4^(N-1-Q)] -> out [6^Q][6][4^(N-1-Q)]
template <int N, int Q, int GS>
__device__ __forceinline__ void axis_forward(const double* __restrict__ in, double* __restrict__ out, int lane) {
    constexpr int P4 = ipow(4, N - 1 - Q), PRE = ipow(6, Q);
    for (int o = lane; o < PRE * 6 * P4; o += GS) {
        const int post = o % P4, al = (o / P4) % 6, pre = o / (6 * P4);
        const double x0 = in[(pre * 4) * P4 + post], xa = in[(pre * 4 + (al >> 1) + 1) * P4 + post];
        out[o] = (al & 1) ? x0 - xa : x0 + xa;
    }
    gsync<GS>();
}

// adjoint of stage Q: in [6^Q][6][4^(N-1-Q)] -> out [6^Q][4][4^(N-1-Q)]
template <int N, int Q, int GS>
__device__ __forceinline__ void axis_adjoint(const double* __restrict__ in, double* __restrict__ out, int lane) {
    constexpr int P4 = ipow(4, N - 1 - Q), PRE = ipow(6, Q);
    for (int o = lane; o < PRE * 4 * P4; o += GS) {
        const int post = o % P4, dig = (o / P4) % 4, pre = o / (4 * P4);
        const double* src = in + (pre * 6) * P4 + post;
        double v;
        if (dig == 0) v = ((src[0] + src[P4]) + (src[2 * P4] + src[3 * P4])) + (src[4 * P4] + src[5 * P4]);
        else v = src[(2 * (dig - 1)) * P4] - src[(2 * (dig - 1) + 1) * P4];
        out[o] = v;
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
    for (int pre = lane; pre < PRE; pre += GS) {
        const double x0 = in[pre * 4];
        double u0 = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double xa = in[pre * 4 + a + 1];
            const int slot = pre * 6 + 2 * a;
            const double wp = ((double)cnt[slot] * inv_total) * fast_recip((x0 + xa) + epsp[slot]);
            const double wm = ((double)cnt[slot + 1] * inv_total) * fast_recip((x0 - xa) + epsp[slot + 1]);
            u0 += wp + wm;
            out[pre * 4 + a + 1] = wp - wm;
        }
        out[pre * 4] = u0;
    }
    gsync<GS>();
}

// C = X * Y for d x d complex matrices in shared memory (row-major, leading dimension ld = d + 1).
// The kernel is bound by shared-memory wavefronts (ncu: 93 % of the LSU data pipe, 40 % of them from the Y loads of
// this routine), so at n = 4 every lane owns a 2 x 4 register tile (rows a and a + d/2): 4 Y loads and 2 X loads
// feed 8 complex multiply-adds.  Lanes of a quarter-warp share the column block (their Y address is one broadcast)
// and differ in the row, whose stride of ld complex numbers keeps the X loads conflict-free.
template <int N, int GS>
__device__ __forceinline__ void cmatmul(double2* __restrict__ C, const double2* __restrict__ X,
                                        const double2* __restrict__ Y, int lane) {
    constexpr int d = Axis<N>::d, ld = Axis<N>::ld;
    // n = 4: 2 x 4 tiles (32 tasks, shared-memory bound); n = 3: 1 x 2 tiles (32 tasks, latency bound: keep lanes busy)
    constexpr int TI = (d >= 16) ? 2 : 1, TJ = (d >= 16) ? 4 : 2, ROWS = d / TI;
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
            for (int i = 0; i < TI; ++i) x[i] = X[(a + i * ROWS) * ld + c];
#pragma unroll
            for (int j = 0; j < TJ; ++j) {
                const double2 y = Y[c * ld + b0 + j];
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
            for (int j = 0; j < TJ; ++j) C[(a + i * ROWS) * ld + b0 + j] = make_double2(re[i][j], im[i][j]);
    }
    gsync<GS>();
}

template <int N, int GS>
__global__ void __launch_bounds__(GS, Axis<N>::min_blocks)
k_mle_rrr_axis(int K, int B, const int* __restrict__ slot_of_col, const double* __restrict__ epsp,
               const int32_t* __restrict__ counts, const double* __restrict__ rho0, int max_iter, double tol,
               double* __restrict__ rho_out, int32_t* __restrict__ iters, unsigned int* __restrict__ queue) {
    constexpr int d = Axis<N>::d, D = Axis<N>::D, K6 = Axis<N>::K6, ld = Axis<N>::ld;
    extern __shared__ __align__(16) double smd[];
    __shared__ double red[32];
    __shared__ unsigned int next_sample;
    const int lane = threadIdx.x;  // index within the group (= CTA)
    constexpr int BUF = Axis<N>::buf;
    double2* rho = reinterpret_cast<double2*>(smd);    // [d][ld]
    double2* Rm = rho + d * ld;                        // [d][ld]
    double* bufA = reinterpret_cast<double*>(Rm + d * ld);
    double* bufB = bufA + BUF;
    int* cnt = reinterpret_cast<int*>(bufB + BUF);     // counts in canonical slot order
    double2* W = reinterpret_cast<double2*>(bufB);     // transform workspace (dense D) / product (padded): aliases bufB

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
                rho[a * ld + bb] = (a == bb) ? make_double2(z.x, 0.0) : (a < bb ? z : make_double2(zt.x, -zt.y));
            }
        } else {
            for (int e = lane; e < D; e += GS)
                rho[(e / d) * ld + e % d] = make_double2((e / d == e % d) ? 1.0 / d : 0.0, 0.0);
        }
        gsync<GS>();
        int it = 0;
        for (it = 1; it <= max_iter; ++it) {
            // S_i = Tr(sigma_i rho)
            for (int e = lane; e < D; e += GS) W[tw(interleave<N>(e / d, e % d))] = rho[(e / d) * ld + e % d];
            gsync<GS>();
            pauli_forward<N, GS>(W, lane);
            for (int e = lane; e < D; e += GS) bufA[e] = W[tw(e)].x;
            gsync<GS>();
            double* q = axis_forward_all<N, GS>(bufA, bufB, lane);  // after N-1 stages: [6^(N-1)][4]
            double* u = q == bufA ? bufB : bufA;
            axis_last_fused<N, GS>(q, u, cnt, epsp, inv_total, lane);
            double* g = axis_adjoint_all<N, GS>(u, q, lane);  // Pauli coefficients of R
            {  // g may live in the buffer W aliases: pass the coefficients through registers
                constexpr int PER = (D + GS - 1) / GS;
                double gv[PER];
#pragma unroll
                for (int j = 0; j < PER; ++j) gv[j] = (lane + j * GS < D) ? g[lane + j * GS] : 0.0;
                gsync<GS>();
#pragma unroll
                for (int j = 0; j < PER; ++j)
                    if (lane + j * GS < D) W[tw(lane + j * GS)] = make_double2(gv[j], 0.0);
                gsync<GS>();
            }
            pauli_inverse<N, GS>(W, lane);
            for (int e = lane; e < D; e += GS) Rm[(e / d) * ld + e % d] = W[tw(interleave<N>(e / d, e % d))];
            gsync<GS>();
            cmatmul<N, GS>(W, Rm, rho, lane);   // W = R rho
            double2* T = reinterpret_cast<double2*>(bufA);  // 2*d*ld doubles fit in one contraction buffer (<= 6^N)
            cmatmul<N, GS>(T, W, Rm, lane);     // T = R rho R
            // Hermitise, normalise, step norm
            double tr = 0.0;
            for (int a = lane; a < d; a += GS) tr += T[a * ld + a].x;
            tr = gsum<GS>(tr, red, lane);
            const double inv = 1.0 / tr;
            double del = 0.0;
            for (int e = lane; e < D; e += GS) {
                const int a = e / d, bb = e % d;
                double2 v;
                if (a == bb) {
                    v = make_double2(T[a * ld + a].x * inv, 0.0);
                } else {
                    const double2 z = T[a * ld + bb], zt = T[bb * ld + a];
                    v = make_double2(0.5 * (z.x + zt.x) * inv, 0.5 * (z.y - zt.y) * inv);
                }
                const double2 old = rho[a * ld + bb];
                const double dr = v.x - old.x, di = v.y - old.y;
                del += dr * dr + di * di;
                W[a * ld + bb] = v;
            }
            del = sqrt(gsum<GS>(del, red, lane));
            gsync<GS>();
            for (int e = lane; e < D; e += GS) rho[(e / d) * ld + e % d] = W[(e / d) * ld + e % d];
            gsync<GS>();
            if (del < tol) break;
        }
        if (it > max_iter) it = max_iter;
        double2* out = reinterpret_cast<double2*>(rho_out) + (size_t)b * D;
        for (int e = lane; e < D; e += GS) out[e] = rho[(e / d) * ld + e % d];
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
    if (!plan->axis_ok || getenv("QPB_NO_AXIS_KERNEL")) return QPB_ERR_UNSUPPORTED;
    if (plan->n == 3) return launch_axis<3>(plan, B, counts, rho0, max_iter, tol, rho, iters, st);
    if (plan->n == 4) return launch_axis<4>(plan, B, counts, rho0, max_iter, tol, rho, iters, st);
    return QPB_ERR_UNSUPPORTED;
}

}  // namespace qpb
