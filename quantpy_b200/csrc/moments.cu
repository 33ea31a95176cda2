// Moments of the weighted squared l2 error of multinomial frequencies:
// replaces l2_mean / l2_variance of quantpy/stats.py:5-53, which MomentInterval
// (quantpy/tomography/interval.py:59-110) evaluates through twelve six-operand einsums.
// The contractions are regrouped by POVM pair (a, b) over the O x O blocks W_ab = W[a,:,b,:]:
//   G_ab = f_a^T W_ab f_b,  u_ab[j] = (f_a^T W_ab)_j,  v_ab[i] = (W_ab f_b)_i,
//   n E = T0 - T1,   n^2 E2 = (T1 - T0)^2 + sum_ab [ G_ab G_ba + G_ab^2 - sum_j f_bj u_ab[j] (v_ba[j] + u_ab[j])
//                              - sum_i f_ai v_ab[i] (u_ba[i] + v_ab[i]) + sum_ij W_ab[ij] (W_ba[ji] + W_ab[ij]) f_ai f_bj ]
// One CTA per frequency table, one thread per POVM pair, O(O^2) work per pair, deterministic block reduction.
#include "../../include/quantpy_b200.h"
#include "common.cuh"

namespace qpb {

__global__ void __launch_bounds__(256)
k_l2_moments(int P, int O, const double* __restrict__ W, const double* __restrict__ freq, double n_trials,
             double* __restrict__ mean_out, double* __restrict__ var_out) {
    const double* f = freq + (size_t)blockIdx.x * P * O;
    double t0 = 0.0, t1 = 0.0, rest = 0.0;
    for (long pair = threadIdx.x; pair < (long)P * P; pair += blockDim.x) {
        const int a = (int)(pair / P), b = (int)(pair % P);
        const double* fa = f + (size_t)a * O;
        const double* fb = f + (size_t)b * O;
        auto Wab = [&](int i, int j) { return W[(((size_t)a * O + i) * P + b) * O + j]; };
        auto Wba = [&](int k, int l) { return W[(((size_t)b * O + k) * P + a) * O + l]; };
        double gab = 0.0, gba = 0.0, acc = 0.0;
        for (int i = 0; i < O; ++i) {  // rows of W_ab: v_ab[i], u_ba[i], elementwise terms
            double v = 0.0, ub = 0.0, el = 0.0;
            for (int j = 0; j < O; ++j) {
                const double w = Wab(i, j), wt = Wba(j, i);
                v = fma(w, fb[j], v);
                ub = fma(fb[j], wt, ub);
                el = fma(w * (wt + w), fb[j], el);
            }
            gab = fma(fa[i], v, gab);
            gba = fma(fa[i], ub, gba);
            acc += fa[i] * (el - v * (ub + v));
        }
        for (int j = 0; j < O; ++j) {  // columns of W_ab: u_ab[j], v_ba[j]
            double u = 0.0, vb = 0.0;
            for (int i = 0; i < O; ++i) {
                u = fma(fa[i], Wab(i, j), u);
                vb = fma(Wba(j, i), fa[i], vb);
            }
            acc -= fb[j] * u * (vb + u);
        }
        rest += gab * gba + gab * gab + acc;
        if (a == b) {
            t1 += gab;
            for (int i = 0; i < O; ++i) t0 = fma(Wab(i, i), fa[i], t0);
        }
    }
    __shared__ double red[3][8];
    t0 = warp_sum(t0);
    t1 = warp_sum(t1);
    rest = warp_sum(rest);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][warp] = t0;
        red[1][warp] = t1;
        red[2][warp] = rest;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            s0 += red[0][w];
            s1 += red[1][w];
            s2 += red[2][w];
        }
        const double mean = (s0 - s1) / n_trials;
        const double second = ((s1 - s0) * (s1 - s0) + s2) / (n_trials * n_trials);
        mean_out[blockIdx.x] = mean;
        var_out[blockIdx.x] = second - mean * mean;
    }
}

}  // namespace qpb

using namespace qpb;

extern "C" int qpb_l2_moments(int B, int P, int O, const double* weights, const double* freq, double n_trials,
                              double* mean_out, double* var_out, void* stream) {
    QPB_REQUIRE(B >= 0 && P >= 1 && O >= 1 && n_trials > 0, "bad arguments B=%d P=%d O=%d", B, P, O);
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(weights && freq && mean_out && var_out, "NULL buffer");
    k_l2_moments<<<B, 256, 0, (cudaStream_t)stream>>>(P, O, weights, freq, n_trials, mean_out, var_out);
    QPB_LAUNCHED("k_l2_moments");
    return QPB_OK;
}
