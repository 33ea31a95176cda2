// Structured R.rho.R maximum likelihood for n = 3, 4 qubits and Pauli-axis POVMs ('proj', 'proj-set',
// 'proj4', any shot weights): ONE WARP PER SAMPLE, all state in shared memory.
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
    // doubles of shared memory per warp: rho, R, W (complex d x d each) + two contraction buffers + f
    static constexpr int per_warp = 3 * 2 * D + 3 * K6;
};

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

// forward fast Pauli transform, in place on W (complex, digit-interleaved): W[i] <- Tr(sigma_i M)
template <int N>
__device__ __forceinline__ void pauli_forward(double2* __restrict__ W, int lane) {
    constexpr int D = Axis<N>::D;
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const int stride = 1 << (2 * (N - 1 - q));
        for (int g = lane; g < D / 4; g += 32) {
            const int idx = (g / stride) * 4 * stride + (g % stride);
            const double2 v0 = W[idx], v1 = W[idx + stride], v2 = W[idx + 2 * stride], v3 = W[idx + 3 * stride];
            W[idx] = make_double2(v0.x + v3.x, v0.y + v3.y);                   // I
            W[idx + stride] = make_double2(v1.x + v2.x, v1.y + v2.y);          // X
            W[idx + 2 * stride] = make_double2(-(v1.y - v2.y), v1.x - v2.x);   // Y = i (v1 - v2)
            W[idx + 3 * stride] = make_double2(v0.x - v3.x, v0.y - v3.y);      // Z
        }
        __syncwarp();
    }
}

// inverse: W[i] = g_i (complex) -> W[e(a,b)] = (sum_i g_i sigma_i)[a][b]
template <int N>
__device__ __forceinline__ void pauli_inverse(double2* __restrict__ W, int lane) {
    constexpr int D = Axis<N>::D;
#pragma unroll
    for (int q = 0; q < N; ++q) {
        const int stride = 1 << (2 * (N - 1 - q));
        for (int g = lane; g < D / 4; g += 32) {
            const int idx = (g / stride) * 4 * stride + (g % stride);
            const double2 gi = W[idx], gx = W[idx + stride], gy = W[idx + 2 * stride], gz = W[idx + 3 * stride];
            W[idx] = make_double2(gi.x + gz.x, gi.y + gz.y);                   // (0,0)
            W[idx + stride] = make_double2(gx.x + gy.y, gx.y - gy.x);          // (0,1) = gx - i gy
            W[idx + 2 * stride] = make_double2(gx.x - gy.y, gx.y + gy.x);      // (1,0) = gx + i gy
            W[idx + 3 * stride] = make_double2(gi.x - gz.x, gi.y - gz.y);      // (1,1)
        }
        __syncwarp();
    }
}

// stage Q of the axis map: in [6^Q][4][4^(N-1-Q)] -> out [6^Q][6][4^(N-1-Q)]
template <int N, int Q>
__device__ __forceinline__ void axis_forward(const double* __restrict__ in, double* __restrict__ out, int lane) {
    constexpr int P4 = ipow(4, N - 1 - Q), PRE = ipow(6, Q);
    for (int o = lane; o < PRE * 6 * P4; o += 32) {
        const int post = o % P4, al = (o / P4) % 6, pre = o / (6 * P4);
        const double x0 = in[(pre * 4) * P4 + post], xa = in[(pre * 4 + (al >> 1) + 1) * P4 + post];
        out[o] = (al & 1) ? x0 - xa : x0 + xa;
    }
    __syncwarp();
}

// adjoint of stage Q: in [6^Q][6][4^(N-1-Q)] -> out [6^Q][4][4^(N-1-Q)]
template <int N, int Q>
__device__ __forceinline__ void axis_adjoint(const double* __restrict__ in, double* __restrict__ out, int lane) {
    constexpr int P4 = ipow(4, N - 1 - Q), PRE = ipow(6, Q);
    for (int o = lane; o < PRE * 4 * P4; o += 32) {
        const int post = o % P4, dig = (o / P4) % 4, pre = o / (4 * P4);
        const double* src = in + (pre * 6) * P4 + post;
        double v;
        if (dig == 0) v = ((src[0] + src[P4]) + (src[2 * P4] + src[3 * P4])) + (src[4 * P4] + src[5 * P4]);
        else v = src[(2 * (dig - 1)) * P4] - src[(2 * (dig - 1) + 1) * P4];
        out[o] = v;
    }
    __syncwarp();
}

template <int N, int Q = 0>
__device__ __forceinline__ double* axis_forward_all(double* a, double* b, int lane) {
    if constexpr (Q == N) {
        return a;  // result lives in `a`
    } else {
        axis_forward<N, Q>(a, b, lane);
        return axis_forward_all<N, Q + 1>(b, a, lane);
    }
}
template <int N, int Q = N - 1>
__device__ __forceinline__ double* axis_adjoint_all(double* a, double* b, int lane) {
    if constexpr (Q < 0) {
        return a;
    } else {
        axis_adjoint<N, Q>(a, b, lane);
        return axis_adjoint_all<N, Q - 1>(b, a, lane);
    }
}

// C = X * Y for d x d complex matrices in shared memory (row-major); each lane owns whole output elements
template <int N>
__device__ __forceinline__ void cmatmul(double2* __restrict__ C, const double2* __restrict__ X,
                                        const double2* __restrict__ Y, int lane) {
    constexpr int d = Axis<N>::d, D = Axis<N>::D;
    constexpr int TJ = (D / 32 >= 4) ? 4 : (D / 32 >= 2 ? 2 : 1);  // outputs per lane per pass (same row)
    for (int t = lane; t < D / TJ; t += 32) {
        const int a = t / (d / TJ), b0 = (t % (d / TJ)) * TJ;
        double re[TJ], im[TJ];
#pragma unroll
        for (int j = 0; j < TJ; ++j) re[j] = im[j] = 0.0;
#pragma unroll 4
        for (int c = 0; c < d; ++c) {
            const double2 x = X[a * d + c];
#pragma unroll
            for (int j = 0; j < TJ; ++j) {
                const double2 y = Y[c * d + b0 + j];
                re[j] = fma(x.x, y.x, re[j]);
                re[j] = fma(-x.y, y.y, re[j]);
                im[j] = fma(x.x, y.y, im[j]);
                im[j] = fma(x.y, y.x, im[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < TJ; ++j) C[a * d + b0 + j] = make_double2(re[j], im[j]);
    }
    __syncwarp();
}

template <int N>
__global__ void k_mle_rrr_axis(int K, int B, const int* __restrict__ slot_of_col, const double* __restrict__ epsp_g,
                               const int32_t* __restrict__ counts, const double* __restrict__ rho0, int max_iter,
                               double tol, double* __restrict__ rho_out, int32_t* __restrict__ iters,
                               unsigned int* __restrict__ queue) {
    constexpr int d = Axis<N>::d, D = Axis<N>::D, K6 = Axis<N>::K6;
    extern __shared__ __align__(16) double smd[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    double* epsp = smd;  // [K6], shared by the CTA
    double* base = smd + K6 + (size_t)warp * Axis<N>::per_warp;
    double2* rho = reinterpret_cast<double2*>(base);
    double2* Rm = rho + D;
    double2* W = Rm + D;
    double* bufA = reinterpret_cast<double*>(W + D);
    double* bufB = bufA + K6;
    double* f = bufB + K6;
    for (int e = threadIdx.x; e < K6; e += blockDim.x) epsp[e] = epsp_g[e];
    __syncthreads();
    (void)nw;

    for (;;) {
        unsigned int b = 0;
        if (lane == 0) b = atomicAdd(queue, 1u);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= (unsigned)B) break;
        // ---- load the sample: frequencies into canonical slots, start state
        const int32_t* c = counts + (size_t)b * K;
        long long tot = 0;
        for (int k = lane; k < K; k += 32) tot += c[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        const double total = (double)tot;
        for (int e = lane; e < K6; e += 32) f[e] = 0.0;
        __syncwarp();
        for (int k = lane; k < K; k += 32) f[slot_of_col[k]] = (double)c[k] / total;
        if (rho0) {
            const double2* r0 = reinterpret_cast<const double2*>(rho0) + (size_t)b * D;
            for (int e = lane; e < D; e += 32) {  // Hermitian part of the start, as the packed kernels take it
                const int a = e / d, bb = e % d;
                const double2 z = r0[e], zt = r0[bb * d + a];
                rho[e] = (a == bb) ? make_double2(z.x, 0.0) : (a < bb ? z : make_double2(zt.x, -zt.y));
            }
        } else {
            for (int e = lane; e < D; e += 32) rho[e] = make_double2((e / d == e % d) ? 1.0 / d : 0.0, 0.0);
        }
        __syncwarp();
        int it = 0;
        for (it = 1; it <= max_iter; ++it) {
            // S_i = Tr(sigma_i rho)
            for (int e = lane; e < D; e += 32) W[interleave<N>(e / d, e % d)] = rho[e];
            __syncwarp();
            pauli_forward<N>(W, lane);
            for (int e = lane; e < D; e += 32) bufA[e] = W[e].x;
            __syncwarp();
            double* q = axis_forward_all<N>(bufA, bufB, lane);  // q[slot] = p_slot / c_slot
            for (int e = lane; e < K6; e += 32) {
                const double y = q[e] + epsp[e];
                double x;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(y));
                double er = fma(-y, x, 1.0);
                x = fma(x, er, x);
                er = fma(-y, x, 1.0);
                x = fma(x, er, x);
                q[e] = f[e] * x;
            }
            __syncwarp();
            double* g = axis_adjoint_all<N>(q, q == bufA ? bufB : bufA, lane);  // Pauli coefficients of R
            for (int e = lane; e < D; e += 32) W[e] = make_double2(g[e], 0.0);
            __syncwarp();
            pauli_inverse<N>(W, lane);
            for (int e = lane; e < D; e += 32) Rm[e] = W[interleave<N>(e / d, e % d)];
            __syncwarp();
            cmatmul<N>(W, Rm, rho, lane);   // W = R rho
            double2* T = reinterpret_cast<double2*>(bufA);  // 2*D doubles fit in one contraction buffer (2*4^N <= 6^N for N >= 3)
            cmatmul<N>(T, W, Rm, lane);     // T = R rho R
            // Hermitise, normalise, step norm
            double tr = 0.0;
            for (int a = lane; a < d; a += 32) tr += T[a * d + a].x;
            tr = warp_sum(tr);
            const double inv = 1.0 / tr;
            double del = 0.0;
            for (int e = lane; e < D; e += 32) {
                const int a = e / d, bb = e % d;
                double2 v;
                if (a == bb) {
                    v = make_double2(T[e].x * inv, 0.0);
                } else {
                    const double2 z = T[e], zt = T[bb * d + a];
                    v = make_double2(0.5 * (z.x + zt.x) * inv, 0.5 * (z.y - zt.y) * inv);
                }
                const double dr = v.x - rho[e].x, di = v.y - rho[e].y;
                del += dr * dr + di * di;
                W[e] = v;
            }
            del = sqrt(warp_sum(del));
            __syncwarp();
            for (int e = lane; e < D; e += 32) rho[e] = W[e];
            __syncwarp();
            if (del < tol) break;
        }
        if (it > max_iter) it = max_iter;
        double2* out = reinterpret_cast<double2*>(rho_out) + (size_t)b * D;
        for (int e = lane; e < D; e += 32) out[e] = rho[e];
        if (iters && lane == 0) iters[b] = it;
        __syncwarp();
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
    constexpr int K6 = Axis<N>::K6;
    const int warps = (N == 4) ? 4 : 8;
    const size_t smem = sizeof(double) * ((size_t)K6 + (size_t)warps * Axis<N>::per_warp);
    QPB_REQUIRE(smem <= 227 * 1024, "axis kernel needs %zu bytes of shared memory", smem);
    auto kern = k_mle_rrr_axis<N>;
    if (smem > 48 * 1024) QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    QPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    long blocks = (long)num_sms() * per_sm;
    const long need = ((long)B + warps - 1) / warps;
    if (blocks > need) blocks = need;
    unsigned int* queue = static_cast<unsigned int*>(scratch(st, 0, sizeof(unsigned int)));
    if (!queue) return QPB_ERR_NOMEM;
    QPB_CUDA(cudaMemsetAsync(queue, 0, sizeof(unsigned int), st));
    kern<<<(int)blocks, warps * 32, smem, st>>>(plan->K, B, plan->axis_slots, plan->axis_epsp, counts, rho0, max_iter,
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
