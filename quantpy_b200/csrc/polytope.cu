// Polytope-coverage kernels: the per-trial body of the coverage experiments
// test_qst / test_qpt (quantpy/tomography/polytopes/verification.py:9-78) and the helpers they bisect over,
// count_confidence / count_delta (quantpy/tomography/polytopes/utils.py:4-27).
// One thread per (trial, confidence level); the trial's count tables come from the multinomial sampler.
#include "../../include/quantpy_b200.h"
#include "common.cuh"

namespace qpb {

constexpr int kMaxGroups = 256;
struct ShotVecD {
    double n[kMaxGroups];
};

constexpr double kPolyEps = 1e-15;  // EPS of verification.py:16 and utils.py:5

__device__ __forceinline__ double clipped_freq(const int32_t* __restrict__ counts, const double* __restrict__ freq,
                                               long idx, double n) {
    if (freq) return freq[idx];
    return fmin(fmax((double)counts[idx] / n, kPolyEps), 1.0 - kPolyEps);  // verification.py:29
}

// utils.py:4-13 for one trial
__device__ double poly_confidence(double delta, int M, int O, const int32_t* __restrict__ counts,
                                  const double* __restrict__ freq, long base, const ShotVecD& shots) {
    double prod = 1.0;
    for (int m = 0; m < M; ++m) {
        const double n = shots.n[m];
        double s = 0.0;
        for (int o = 0; o < O; ++o) {
            const double f = clipped_freq(counts, freq, base + (long)m * O + o, n);
            const double sh = fmin(fmax(f + delta, kPolyEps), 1.0 - kPolyEps);
            double kl = f * log(f / sh) + (1.0 - f) * log((1.0 - f) / (1.0 - sh));
            if (!(sh < 1.0 - kPolyEps)) kl = INFINITY;
            double e = exp(-n * kl);
            if (fabs(f - 1.0) < 2.0 * kPolyEps) e = 0.0;
            s += e;
        }
        prod *= fmax(1.0 - s, 0.0);
    }
    return prod;
}

__global__ void k_polytope_coverage(int B, int M, int O, int L, const int32_t* __restrict__ counts,
                                    const double* __restrict__ freq, ShotVecD shots,
                                    const double* __restrict__ levels, const double* __restrict__ p_true, int clip_b,
                                    double* __restrict__ delta_out, unsigned char* __restrict__ inside_out) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)B * L) return;
    const long b = t / L;
    const int j = (int)(t % L);
    const long base = b * (long)M * O;
    const double target = levels[j];
    // utils.py:16-27: bisection on [1e-10, 1], stop at width 1e-10, return the last midpoint
    double left = 1e-10, right = 1.0, delta = 0.0;
    while (right - left > 1e-10) {
        delta = 0.5 * (left + right);
        if (poly_confidence(delta, M, O, counts, freq, base, shots) < target + 1e-10) left = delta;
        else right = delta;
    }
    if (delta_out) delta_out[t] = delta;
    if (inside_out && p_true) {
        // verification.py:31-35 / :68-76: the true probabilities must lie below frequencies + delta
        double worst = INFINITY;
        for (int m = 0; m < M; ++m)
            for (int o = 0; o < O; ++o) {
                const long k = (long)m * O + o;
                double bk = clipped_freq(counts, freq, base + k, shots.n[m]) + delta;
                if (clip_b) bk = fmin(fmax(bk, kPolyEps), 1.0 - kPolyEps);
                worst = fmin(worst, bk - p_true[k]);
            }
        inside_out[t] = worst > -kPolyEps ? 1 : 0;
    }
}

__global__ void k_polytope_confidence(int B, int M, int O, int L, const double* __restrict__ freq, ShotVecD shots,
                                      const double* __restrict__ deltas, double* __restrict__ conf_out) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)B * L) return;
    conf_out[t] = poly_confidence(deltas[t % L], M, O, nullptr, freq, (t / L) * (long)M * O, shots);
}

}  // namespace qpb

using namespace qpb;

extern "C" {

int qpb_polytope_coverage(int B, int M, int O, const int32_t* counts, const double* freq,
                          const double* n_shots_host, int L, const double* levels, const double* p_true, int clip_b,
                          double* delta_out, unsigned char* inside_out, void* stream) {
    QPB_REQUIRE(B >= 0 && M >= 1 && O >= 1 && L >= 1, "bad shape B=%d M=%d O=%d L=%d", B, M, O, L);
    QPB_REQUIRE(M <= kMaxGroups, "M=%d exceeds the supported %d POVMs", M, kMaxGroups);
    if (B == 0) return QPB_OK;
    QPB_REQUIRE((counts != nullptr) != (freq != nullptr), "pass exactly one of counts / freq");
    QPB_REQUIRE(n_shots_host && levels, "NULL buffer");
    QPB_REQUIRE(delta_out || (inside_out && p_true), "nothing to compute");
    ShotVecD shots;
    for (int m = 0; m < M; ++m) shots.n[m] = n_shots_host[m];
    const long total = (long)B * L;
    k_polytope_coverage<<<(int)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        B, M, O, L, counts, freq, shots, levels, p_true, clip_b, delta_out, inside_out);
    QPB_LAUNCHED("k_polytope_coverage");
    return QPB_OK;
}

int qpb_polytope_confidence(int B, int M, int O, const double* freq, const double* n_shots_host, int L,
                            const double* deltas, double* conf_out, void* stream) {
    QPB_REQUIRE(B >= 0 && M >= 1 && O >= 1 && L >= 1, "bad shape");
    QPB_REQUIRE(M <= kMaxGroups, "M=%d exceeds the supported %d POVMs", M, kMaxGroups);
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(freq && n_shots_host && deltas && conf_out, "NULL buffer");
    ShotVecD shots;
    for (int m = 0; m < M; ++m) shots.n[m] = n_shots_host[m];
    const long total = (long)B * L;
    k_polytope_confidence<<<(int)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(B, M, O, L, freq, shots,
                                                                                      deltas, conf_out);
    QPB_LAUNCHED("k_polytope_confidence");
    return QPB_OK;
}

}  // extern "C"
