// Register-resident batched Hermitian Jacobi + physical projection for d = 8, 16 (n = 3, 4):
// d LANES PER MATRIX, lane r owns row r of A and row r of the eigenvector matrix V.
//
// Replaces _make_feasible (quantpy/tomography/state.py:267-273) after the DMMA inversion for n = 3.  The
// shared-memory version (k_project_packed, jacobi.cuh) streams A and V through shared memory three times per
// round and runs at 79 % of the shared-memory pipe (profiles/ncu_r1_prof_lin3_v2.txt).  Here the column update
// A <- A G and V <- V G touches only the lane's own registers; the row update A <- G^dagger A needs the partner
// row, which the two lanes of a pair swap with shuffles (2 d doubles), and the d/2 rotations of a round are
// computed by their p-lanes and broadcast (3 doubles each).  Rows and columns are kept in tournament POSITION order
// (see jacobi_round), so every register index is a compile-time constant without unrolling the rounds.  Same
// rotation and stopping rule as jacobi.cuh; the order of the pairs within a sweep differs.
#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "plan.h"
#include "jacobi_rows.cuh"

// stop when off^2 <= QPB_JACOBI_REL2 * fro^2.  Measured (tools/jacobi_tol_probe.py, max Frobenius error of the projected
// state against numpy eigh / time at C3): 1e-30 -> 1.9e-15 / 1.066 ms, 1e-26 -> 5.0e-14 / 1.025 ms, 1e-22 -> 4.9e-12 / 0.976 ms
#ifndef QPB_JACOBI_REL2
#define QPB_JACOBI_REL2 1e-26
#endif

namespace qpb {

constexpr int kRowsThreads = 128;

template <int d>
__global__ void __launch_bounds__(kRowsThreads, d == 8 ? 4 : 2)
k_project_rows(int B, const double* __restrict__ h_in, double* __restrict__ rho) {
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int G = d, dd = d * d;
    const int tid = threadIdx.x, lane = tid & 31, gl = tid % G, gbase = lane - gl;
    const long groups_per_block = kRowsThreads / G;
    const long stride = (long)gridDim.x * groups_per_block;
    const long first = (long)blockIdx.x * groups_per_block + tid / G;
    const long rounds = (B + stride - 1) / stride;  // every group runs the same number of rounds (warp-wide shuffles)
    for (long it = 0; it < rounds; ++it) {
        const long b = first + it * stride;
        const bool valid = b < B;
        RowState<d> S;
#pragma unroll
        for (int c = 0; c < d; ++c) {
            cplx z;
            z.re = (c == gl) ? 1.0 : 0.0;
            z.im = 0.0;
            if (valid) z = herm_get(h_in + b * dd, d, gl, c);
            S.ar[c] = z.re;
            S.ai[c] = z.im;
            S.vr[c] = (c == gl) ? 1.0 : 0.0;
            S.vi[c] = 0.0;
        }
        bool done = !valid;
        for (int sweep = 0; sweep < 40; ++sweep) {
            double off = 0.0, fro = 0.0;
#pragma unroll
            for (int c = 0; c < d; ++c) {
                const double m2 = S.ar[c] * S.ar[c] + S.ai[c] * S.ai[c];
                fro += m2;
                if (c != gl) off += m2;
            }
            off = gsum_rows<d, G>(off);
            fro = gsum_rows<d, G>(fro);
            if (off <= QPB_JACOBI_REL2 * fro || fro == 0.0) done = true;
            if (__all_sync(kFull, done)) break;
            const double tiny2 = 1e-36 * fro;
#pragma unroll 1
            for (int round = 0; round < d - 1; ++round) jacobi_round<d>(S, gl, gbase, done, tiny2);
        }
        // ---- clip the spectrum, recompose rho = V diag(lam) V^dagger / Tr  (state.py:269-273) ------------
        double lam_own = 0.0;
#pragma unroll
        for (int c = 0; c < d; ++c)
            if (c == gl) lam_own = S.ar[c];
        lam_own = fmax(lam_own, kClipState);
        double lam[d], tr = 0.0;
#pragma unroll
        for (int j = 0; j < d; ++j) {
            lam[j] = __shfl_sync(kFull, lam_own, gbase + j);
            tr += lam[j];
        }
        const double inv = 1.0 / tr;
        double wr[d], wi[d];  // lam_j V_aj of my row a
#pragma unroll
        for (int j = 0; j < d; ++j) {
            wr[j] = lam[j] * S.vr[j];
            wi[j] = lam[j] * S.vi[j];
        }
        double2* out = reinterpret_cast<double2*>(rho) + b * dd + (long)gl * d;
#pragma unroll
        for (int bb = 0; bb < d; ++bb) {  // rho_{a,bb} = sum_j (lam_j V_aj) conj(V_bb,j): row bb comes from lane bb
            double re = 0.0, im = 0.0;
#pragma unroll
            for (int j = 0; j < d; ++j) {
                const double yr = __shfl_sync(kFull, S.vr[j], gbase + bb), yi = __shfl_sync(kFull, S.vi[j], gbase + bb);
                re += wr[j] * yr + wi[j] * yi;
                im += wi[j] * yr - wr[j] * yi;
            }
            if (valid) out[bb] = make_double2(re * inv, im * inv);
        }
    }
}

int launch_project_rows(int d, int B, const double* h_in, double* rho, cudaStream_t st) {
    if (d != 8 && d != 16) return QPB_ERR_UNSUPPORTED;
    const long groups = kRowsThreads / d;
    long blocks = ((long)B + groups - 1) / groups;
    const long cap = (long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (d == 8) k_project_rows<8><<<(int)blocks, kRowsThreads, 0, st>>>(B, h_in, rho);
    else k_project_rows<16><<<(int)blocks, kRowsThreads, 0, st>>>(B, h_in, rho);
    QPB_LAUNCHED("k_project_rows");
    return QPB_OK;
}

}  // namespace qpb
