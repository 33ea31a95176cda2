// Metropolis-Hastings sampling of the likelihood (quantpy/mhmc.py:48-119, used by MHMCStateInterval,
// quantpy/tomography/interval.py:737-759): ONE WARP PER CHAIN, chains are independent.
//
// A chain is sequential by construction (the reference runs exactly one), so the parallelism is inside a step --
// the d x d product L L^dagger, the K probabilities and their logarithms are spread over the lanes -- and across
// chains when the caller runs many (coverage studies, one chain per data set or per seed).  Chain state, proposal
// and the packed-Hermitian L L^dagger live in shared memory; the POVM table is read through L1 (it is shared by
// every chain).  The proposal noise is either supplied by the caller (delta, u per step: the host layer passes
// the legacy np.random draws, so a seeded run reproduces the reference's chain step by step) or generated in
// the kernel from Philox4x32-10 keyed by (seed, global chain index, step) with Box-Muller.
#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "plan.h"

namespace qpb {

constexpr int kMhmcWarps = 4;
constexpr int kMhmcMaxD = 256;  // n <= 4

// packed Cholesky vector (routines.py:84-101): d diagonal entries, then Re and Im of the strict lower triangle in
// np.tril_indices order
__device__ __forceinline__ double chol_re(const double* x, int d, int i, int j) {
    return i == j ? x[i] : x[d + i * (i - 1) / 2 + j];
}
__device__ __forceinline__ double chol_im(const double* x, int d, int i, int j) {
    return i == j ? 0.0 : x[d + d * (d - 1) / 2 + i * (i - 1) / 2 + j];
}

// h = packed-Hermitian L L^dagger of the vector x (all in shared memory); returns sum_k f_k log(p_k + 1e-10) with
// p = Tr(E_k L L^dagger) / Tr(L L^dagger)  (state.py:217-229)
// hw = h with the off-diagonal entries doubled: Tr(E rho) = sum_e packed(E)[e] hw[e] for Hermitian E, rho.
__device__ double log_target(const double* x, double* h, double* hw, int d, int D, int K, const double* __restrict__ Ar,
                             const double* fk, int lane) {
    for (int e = lane; e < D; e += 32) {
        const int a = e / d, b = e % d;
        const int lo = min(a, b), hi = max(a, b);
        double re = 0.0, im = 0.0;  // M_{lo,hi} = sum_c L_{lo,c} conj(L_{hi,c})
        for (int c = 0; c <= lo; ++c) {
            const double ar = chol_re(x, d, lo, c), ai = chol_im(x, d, lo, c);
            const double br = chol_re(x, d, hi, c), bi = chol_im(x, d, hi, c);
            re = fma(ar, br, fma(ai, bi, re));
            im = fma(ai, br, fma(-ar, bi, im));
        }
        const double v = (a <= b) ? re : im;  // [lo][hi] real part, [hi][lo] imaginary part of the upper element
        h[e] = v;
        hw[e] = (a == b) ? v : 2.0 * v;
    }
    __syncwarp();
    double tr = 0.0;
    for (int a = 0; a < d; ++a) tr += h[a * d + a];
    double acc = 0.0;
    for (int k = lane; k < K; k += 32) {
        const double* row = Ar + (size_t)k * D;
        double p = 0.0;
        for (int e = 0; e < D; ++e) p = fma(row[e], hw[e], p);
        acc = fma(fk[k], log(p / tr + kLogGuard), acc);
    }
    __syncwarp();
    return warp_sum(acc);
}

__global__ void __launch_bounds__(kMhmcWarps * 32)
k_mhmc_state(int d, int D, int K, int C, int n_samples, int thinning, int burn_steps, double step,
             const double* __restrict__ Ar, const int32_t* __restrict__ counts, int counts_batched,
             const double* __restrict__ x_init, const double* __restrict__ deltas, const double* __restrict__ uniforms,
             uint32_t k0, uint32_t k1, uint64_t offset, double* __restrict__ samples, int32_t* __restrict__ accepted,
             double* __restrict__ x_final) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* x = sm + (size_t)warp * (4 * D + K);
    double* xp = x + D;
    double* h = xp + D;
    double* hw = h + D;
    double* fk = hw + D;  // the chain's K frequencies
    const long total_steps = (long)burn_steps + (long)n_samples * thinning;
    for (int c = blockIdx.x * kMhmcWarps + warp; c < C; c += gridDim.x * kMhmcWarps) {
        const int32_t* cnt = counts + (counts_batched ? (size_t)c * K : 0);
        long tot = 0;
        for (int k = lane; k < K; k += 32) tot += cnt[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        for (int k = lane; k < K; k += 32) fk[k] = (double)cnt[k] / (double)tot;
        for (int e = lane; e < D; e += 32) x[e] = x_init[(size_t)c * D + e];
        __syncwarp();
        double cur = log_target(x, h, hw, d, D, K, Ar, fk, lane);
        const uint64_t chain = offset + (uint64_t)c;
        int acc_count = 0;
        for (long s = 0; s < total_steps; ++s) {
            // ---- proposal: x' = (x + step delta) / |x + step delta|   (mhmc.py:117-119) ----------
            double nrm2 = 0.0;
            for (int e = lane; e < D; e += 32) {
                double delta;
                if (deltas) {
                    delta = deltas[((size_t)c * total_steps + s) * D + e];
                } else {
                    const philox4 r = philox4x32_10((uint32_t)s, (uint32_t)(s >> 32) ^ ((uint32_t)e << 8), (uint32_t)chain,
                                                    (uint32_t)(chain >> 32), k0, k1);
                    const double u1 = ((double)(((uint64_t)r.x << 21) ^ (r.y >> 11)) + 0.5) * 0x1.0p-53;
                    const double u2 = ((double)(((uint64_t)r.z << 21) ^ (r.w >> 11)) + 0.5) * 0x1.0p-53;
                    delta = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
                }
                const double v = fma(step, delta, x[e]);
                xp[e] = v;
                nrm2 = fma(v, v, nrm2);
            }
            nrm2 = warp_sum(nrm2);
            const double nrm = sqrt(nrm2);
            for (int e = lane; e < D; e += 32) xp[e] = xp[e] / nrm;
            __syncwarp();
            const double prop = log_target(xp, h, hw, d, D, K, Ar, fk, lane);
            // ---- accept with probability min(1, exp(prop - cur)); NaN compares false = reject --------
            double u;
            if (uniforms) {
                u = uniforms[(size_t)c * total_steps + s];
            } else {
                const philox4 r = philox4x32_10((uint32_t)s, (uint32_t)(s >> 32) ^ 0xffffff00u, (uint32_t)chain,
                                                (uint32_t)(chain >> 32), k0, k1);
                u = (double)(((uint64_t)r.x << 21) ^ (r.y >> 11)) * 0x1.0p-53;
            }
            const bool ok = u <= exp(prop - cur);
            if (ok) {
                for (int e = lane; e < D; e += 32) x[e] = xp[e];
                cur = prop;
            }
            __syncwarp();
            if (s >= burn_steps) {
                const long i = s - burn_steps;
                acc_count += ok;
                if (i % thinning == 0) {
                    // L L^dagger of the CURRENT state as a full complex matrix (interval.py:756 does not renormalise)
                    if (!ok) log_target(x, h, hw, d, D, K, Ar, fk, lane);  // h holds the rejected proposal: rebuild
                    double2* out = reinterpret_cast<double2*>(samples) + ((size_t)c * n_samples + i / thinning) * D;
                    for (int e = lane; e < D; e += 32) {
                        const cplx z = herm_get(h, d, e / d, e % d);
                        out[e] = make_double2(z.re, z.im);
                    }
                    __syncwarp();
                }
            }
        }
        if (lane == 0 && accepted) accepted[c] = acc_count;
        if (x_final)
            for (int e = lane; e < D; e += 32) x_final[(size_t)c * D + e] = x[e];
        __syncwarp();
    }
}

}  // namespace qpb

using namespace qpb;

extern "C" int qpb_mhmc_state(const qpb_state_plan* plan, int C, int n_samples, int thinning, int burn_steps, double step,
                              const int32_t* counts, int counts_batched, const double* x_init, const double* deltas,
                              const double* uniforms, uint64_t seed, uint64_t chain_offset, double* samples,
                              int32_t* accepted, double* x_final, void* stream) {
    QPB_REQUIRE(plan, "NULL plan");
    QPB_REQUIRE(C >= 0 && n_samples >= 0 && burn_steps >= 0, "negative size");
    QPB_REQUIRE(thinning >= 1, "thinning must be >= 1");
    QPB_REQUIRE(plan->D <= kMhmcMaxD, "n_qubits=%d unsupported by the MHMC kernel", plan->n);
    QPB_REQUIRE((deltas == nullptr) == (uniforms == nullptr), "deltas and uniforms must be given together");
    if (C == 0) return QPB_OK;
    QPB_REQUIRE(counts && x_init && (samples || n_samples == 0), "NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(double) * (4 * (size_t)plan->D + plan->K) * kMhmcWarps;
    QPB_REQUIRE(smem <= 200 * 1024, "POVM with K=%d outcomes too large for the MHMC kernel", plan->K);
    if (smem > 48 * 1024)
        QPB_CUDA(cudaFuncSetAttribute(k_mhmc_state, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long blocks = ((long)C + kMhmcWarps - 1) / kMhmcWarps;
    const long cap = (long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    k_mhmc_state<<<(int)blocks, kMhmcWarps * 32, smem, st>>>(plan->d, plan->D, plan->K, C, n_samples, thinning, burn_steps,
                                                              step, plan->Ar, counts, counts_batched, x_init, deltas,
                                                              uniforms, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                              chain_offset, samples, accepted, x_final);
    QPB_LAUNCHED("k_mhmc_state");
    return QPB_OK;
}
