// Shared device/host helpers for libquantpy_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/quantpy_b200.h"

#include <atomic>
#include <cstdio>

#define QPB_OK 0
#define QPB_ERR_INVALID (-1)
#define QPB_ERR_CUDA (-2)
#define QPB_ERR_UNSUPPORTED (-3)
#define QPB_ERR_NOMEM (-4)

namespace qpb {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int num_sms();
int option(int which);  // QPB_OPT_* toggle (include/quantpy_b200.h); cached, no getenv on the launch path
extern std::atomic<int64_t> g_launches;

// Grow-only device scratch, one buffer per (device, stream, slot): kernels of one call reuse it in stream order,
// so no allocation (and no allocator-induced gap after an idle period) happens on the hot path.  Returns nullptr
// and sets the error on failure.  *fresh (optional) is set when the returned buffer was allocated by this call,
// i.e. its contents are undefined.
void* scratch(cudaStream_t st, int slot, size_t bytes, bool* fresh = nullptr);

// Profiling: the buffer registered with qpb_debug_set_trace if it holds at least `bytes`, else nullptr.
long long* debug_trace_buffer(size_t bytes);

// Call right after a <<<>>> launch: counts it and surfaces launch-configuration errors.
#define QPB_LAUNCHED(name)                                     \
    do {                                                       \
        ::qpb::g_launches.fetch_add(1);                        \
        int _rc = ::qpb::check_cuda(cudaGetLastError(), name); \
        if (_rc != QPB_OK) return _rc;                         \
    } while (0)

#define QPB_CUDA(call)                                    \
    do {                                                  \
        int _rc = ::qpb::check_cuda((call), #call);       \
        if (_rc != QPB_OK) return _rc;                    \
    } while (0)

#define QPB_REQUIRE(cond, ...)                            \
    do {                                                  \
        if (!(cond)) {                                    \
            ::qpb::set_error(__VA_ARGS__);                \
            return QPB_ERR_INVALID;                       \
        }                                                 \
    } while (0)

constexpr double kLogGuard = 1e-10;   // quantpy/tomography/state.py:219
constexpr double kClipState = 1e-15;  // quantpy/tomography/state.py:269
constexpr double kClipChoi = 1e-12;   // quantpy/tomography/process.py:272
constexpr double kZeroBelow = 1e-15;  // quantpy/geometry.py:17,35,53

// ---------------------------------------------------------------------------------------------
// Packed-Hermitian layout ("hvec"): a d x d Hermitian matrix is stored as d*d reals, row-major
// H[a][b]:  a==b -> Re rho_aa ;  a<b -> Re rho_ab ;  a>b -> Im rho_ba  (imaginary part of the
// UPPER element (b,a)).  Every kernel keeps states in this form so that Hermiticity is exact.
// ---------------------------------------------------------------------------------------------
struct cplx {
    double re, im;
};

__device__ __forceinline__ cplx herm_get(const double* __restrict__ h, int d, int a, int b) {
    cplx z;
    if (a == b) {
        z.re = h[a * d + a];
        z.im = 0.0;
    } else if (a < b) {
        z.re = h[a * d + b];
        z.im = h[b * d + a];
    } else {
        z.re = h[b * d + a];
        z.im = -h[a * d + b];
    }
    return z;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 1/sqrt(y) and 1/y to ~1 ulp from the hardware seeds (2^-22 / 2^-23) and Newton steps on the FP64 pipe;
// a fraction of the instruction count of IEEE sqrt/div, which dominated the Jacobi rotation set-up.
__device__ __forceinline__ double fast_rsqrt(double y) {
    double x;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(y));
    const double hy = 0.5 * y;
    x = x * fma(-hy * x, x, 1.5);
    x = x * fma(-hy * x, x, 1.5);
    return x;
}
__device__ __forceinline__ double fast_recip(double y) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(y));
    const double e = fma(-y, x, 1.0);
    return fma(x, fma(e, e, e), x);
}

// 1/y with ONE Newton step: relative error = (seed error)^2 <= 2^-44.  For the likelihood weights f/(p + 1e-10) of the
// R.rho.R kernels: a 2^-44 relative error in a weight moves the fixed point by ~2e-14 (Frobenius) and changed no
// iteration count in 20 000 oracle trajectories, four orders below the 1e-10 parity tolerance.
__device__ __forceinline__ double fast_recip_1step(double y) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(y));
    return fma(x, fma(-y, x, 1.0), x);
}

// counts / total for many counts of one sample: ONE IEEE reciprocal, then per count a product and a residual
// correction (q = c r; q + r fma(-q, total, c)).  With r the correctly rounded reciprocal this is the correctly
// rounded quotient (Markstein), i.e. the same bits as the ~35-instruction IEEE division it replaces -- 36 of
// those per refilled lane were 12 % of the MLE kernel's stall samples (profiles/README_r1.md).
struct FreqDiv {
    double total, inv;
    __device__ __forceinline__ explicit FreqDiv(double t) : total(t), inv(1.0 / t) {}
    __device__ __forceinline__ double operator()(double c) const {
        const double q = c * inv;
        return fma(fma(-q, total, c), inv, q);
    }
};

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  key = 64-bit seed, counter = 128 bits.
// ---------------------------------------------------------------------------------------------
struct philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0;
        uint64_t p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    philox4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

}  // namespace qpb
