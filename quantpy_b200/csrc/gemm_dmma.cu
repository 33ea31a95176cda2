// FP64 tensor-core GEMM for the "counts -> estimate" contractions:
//     C[m][n] = sum_k ( counts[m][k] / total[m][k / G] ) * T[k][n]
// m = sample, k = outcome column (G columns share one normalisation: the whole row for state tomography,
// state.py:193; one input state's K outcomes for process tomography, process.py:285), n = packed output entry.
// Used for linear inversion at n >= 3 (T = LhT, K x D) and for 'lifp' (T = LinvT, S*K x 2 d^4), where the
// one-warp-per-sample kernels had to stream the whole table from L2 for every sample.
//
// Tiling: CTA = 64 x 64 outputs, 8 warps in a 4 x 2 grid, warp tile 16 x 32 = 2 x 4 DMMA tiles
// (mma.sync.aligned.m8n8k4 f64, FP64 accumulate), K step 16 through padded shared-memory tiles.
// tcgen05 has no f64 kind, so DMMA via mma.sync is the FP64 tensor path on sm_100a.
#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "plan.h"

namespace qpb {

constexpr int GM = 64, GN = 64, GK = 16;
constexpr int AS = GK + 4;  // padded row stride of the A tile (doubles)
constexpr int BS = GN + 8;  // padded row stride of the B tile

// inv_tot[m][g] = 1 / sum_{k in group g} counts[m][k]; one warp per (m, g)
__global__ void k_group_inverse_totals(int M, int Ktot, int G, const int32_t* __restrict__ counts,
                                       double* __restrict__ inv_tot) {
    const int ng = Ktot / G;
    const long item = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (item >= (long)M * ng) return;
    const int32_t* c = counts + (item / ng) * Ktot + (item % ng) * G;
    long long tot = 0;
    for (int k = lane; k < G; k += 32) tot += c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) inv_tot[item] = 1.0 / (double)tot;
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256)
k_gemm_counts_dmma(int M, int N, int Ktot, int G, const int32_t* __restrict__ counts,
                   const double* __restrict__ inv_tot, const double* __restrict__ T, double* __restrict__ C) {
    __shared__ double As[GM * AS];
    __shared__ double Bs[GK * BS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;           // 4 x 2 warps
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    const int ng = Ktot / G;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int k0 = 0; k0 < Ktot; k0 += GK) {
        // A tile: 64 rows x 16 columns of frequencies (converted from counts on the way in)
#pragma unroll
        for (int e = tid; e < GM * GK; e += 256) {
            const int r = e / GK, c = e % GK;
            const int m = m0 + r, k = k0 + c;
            double v = 0.0;
            if (m < M && k < Ktot) v = (double)counts[(long)m * Ktot + k] * inv_tot[(long)m * ng + k / G];
            As[r * AS + c] = v;
        }
        // B tile: 16 rows x 64 columns of the table
#pragma unroll
        for (int e = tid; e < GK * GN; e += 256) {
            const int r = e / GN, c = e % GN;
            const int k = k0 + r, n = n0 + c;
            Bs[r * BS + c] = (k < Ktot && n < N) ? T[(long)k * N + n] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; kk += 4) {
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = As[(wm * 16 + i * 8 + (lane >> 2)) * AS + kk + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(kk + (lane & 3)) * BS + wn * 32 + j * 8 + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int m = m0 + wm * 16 + i * 8 + (lane >> 2);
            const int n = n0 + wn * 32 + j * 8 + 2 * (lane & 3);
            if (m < M) {
                if (n + 1 < N) {
                    *reinterpret_cast<double2*>(C + (long)m * N + n) = make_double2(acc[i][j][0], acc[i][j][1]);
                } else if (n < N) {
                    C[(long)m * N + n] = acc[i][j][0];
                }
            }
        }
}

// C [M][N] = normalised(counts [M][Ktot]) * T [Ktot][N]; N must be even (packed outputs always are).
int launch_gemm_counts(int M, int N, int Ktot, int G, const int32_t* counts, const double* T, double* C,
                       cudaStream_t st) {
    QPB_REQUIRE(G > 0 && Ktot % G == 0 && (N % 2) == 0, "bad GEMM shape Ktot=%d G=%d N=%d", Ktot, G, N);
    const int ng = Ktot / G;
    double* inv_tot = static_cast<double*>(scratch(st, 2, sizeof(double) * (size_t)M * ng));
    if (!inv_tot) return QPB_ERR_NOMEM;
    const long items = (long)M * ng;
    k_group_inverse_totals<<<(int)((items + 7) / 8), 256, 0, st>>>(M, Ktot, G, counts, inv_tot);
    QPB_LAUNCHED("k_group_inverse_totals");
    dim3 grid((N + GN - 1) / GN, (M + GM - 1) / GM);
    k_gemm_counts_dmma<<<grid, 256, 0, st>>>(M, N, Ktot, G, counts, inv_tot, T, C);
    QPB_LAUNCHED("k_gemm_counts_dmma");
    return QPB_OK;
}

}  // namespace qpb
