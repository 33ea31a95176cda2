// FP64 tensor-core GEMM for the "counts -> estimate" contractions:
//     C[m][n] = sum_k ( counts[m][k] / total[m][k / G] ) * T[k][n]
// m = sample, k = outcome column (G columns share one normalisation: the whole row for state tomography,
// state.py:193; one input state's K outcomes for process tomography, process.py:285), n = packed output entry.
// Used for linear inversion at n >= 3 (T = LhT, K x D) and for 'lifp' (T = LinvT, S*K x 2 d^4), where the
// one-warp-per-sample kernels had to stream the whole table from L2 for every sample.
//
// Tiling: CTA = 64 x 64 outputs, 8 warps in a 4 x 2 grid, warp tile 16 x 32 = 2 x 4 DMMA tiles
// (mma.sync.aligned.m8n8k4 f64, FP64 accumulate).  tcgen05 has no f64 kind, so DMMA via mma.sync is the FP64
// tensor path on sm_100a.  Two kernels: k_gemm_counts_dmma (production: persistent, warp-specialised, operands
// staged by TMA bulk copies into a 3-stage mbarrier ring) and k_gemm_counts_simple (load / barrier / compute,
// for shapes whose rows are not 16-byte multiples, and as the cross-check of the pipeline).
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
k_gemm_counts_simple(int M, int N, int Ktot, int G, const int32_t* __restrict__ counts,
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
            // k order inside a block of 16: step s takes k = 4 c + s from lane column c (the order in which the
            // pipelined kernel's 16-byte count loads deliver them), so both kernels accumulate identically
            const int kcol = 4 * (lane & 3) + (kk >> 2);
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = As[(wm * 16 + i * 8 + (lane >> 2)) * AS + kcol];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kcol * BS + wn * 32 + j * 8 + (lane >> 2)];
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

// ---- pipelined kernel -------------------------------------------------------------------------------------------
// Warp-specialised: warp 8 is the producer, warps 0-7 consume.
//  * table operand (T): the producer streams K-chunks of 32 rows x 64 columns into a 4-stage shared-memory ring with
//    TMA bulk copies (cp.async.bulk, SASS UBLKCP), one 512-byte row per copy, each landing at a padded row stride
//    (68 doubles) that makes the DMMA fragment loads bank-conflict free; an mbarrier per stage counts the bytes in
//    (full) and the consumer warps out (empty).
//  * count operand: never staged.  A lane's fragment elements for four consecutive k-steps are one aligned 16-byte
//    global load (k = 16 t + 4 c + s for step s, lane column c), converted and normalised in registers
//    ((double)c * inv_total) on the way into the DMMA and prefetched one block of 16 ahead.  (A first version staged
//    the counts by TMA as well: 64 row copies of 128 B per chunk made the single producer warp the bottleneck --
//    46 % of the stall samples were consumers waiting on the full barrier, profiles/README_r2.md.)
// CTAs are persistent over (m, n) tiles, n fastest, so the N tiles of one block of count rows run back to back and
// re-read it from L1/L2.
constexpr int KC = 32, STAGES = 4;
constexpr int B_LD = GN + 4;  // doubles per table row in smem: a half-warp's 4 rows x 4 columns hit 16 distinct bank pairs
struct GemmStage {
    double b[KC * B_LD];
};
constexpr int kGemmConsumers = 8;
constexpr int kGemmThreads = 32 * (kGemmConsumers + 1);
constexpr size_t kGemmSmem = sizeof(GemmStage) * STAGES + sizeof(uint64_t) * 2 * STAGES;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ int4 ldg_counts4(const int32_t* p, bool ok) {
    int4 v = make_int4(0, 0, 0, 0);
    if (ok) asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

struct a4 {
    double v[4];
};
__device__ __forceinline__ a4 ldg_f64x4(const double* p, bool ok) {
    a4 r = {{0.0, 0.0, 0.0, 0.0}};
    if (ok) {
        const double2 lo = __ldg(reinterpret_cast<const double2*>(p)), hi = __ldg(reinterpret_cast<const double2*>(p) + 1);
        r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = hi.x; r.v[3] = hi.y;
    }
    return r;
}

// MODE 0: left operand = int32 counts, one normalisation per row (linear inversion);  1: counts, one normalisation
// per group of G columns ('lifp');  2: left operand = float64 rows (`xf`), output scaled by `inv_tot[0]`-free
// epilogue: C = clip(scale * X T) -- the batched POVM-probability contraction p = 2^n M r (state.py:109-110).
// MODE 3: like 2 with the rows of X taken through `rowmap` (the samples of a batched R.rho.R iteration that are still
// running, mle_tiled.cu) and the likelihood-weight epilogue C[m][n] = fq[rowmap[m]][n] / (X T + 1e-10)
// (gradient of state.py:228);  MODE 4: like 2 (rows of X direct), plain epilogue -- R = W A_r of the same iteration.
template <int MODE>
__global__ void __launch_bounds__(kGemmThreads, 2)
k_gemm_counts_dmma(int M, int N, int Ktot, int G, const int32_t* __restrict__ counts, const double* __restrict__ xf,
                   const double* __restrict__ inv_tot, const double* __restrict__ T, double* __restrict__ C,
                   double scale, int clip, const int* __restrict__ rowmap, const double* __restrict__ fq) {
    constexpr bool XF = MODE >= 2;  // float64 left operand
    constexpr bool ONE_GROUP = MODE == 0;
    extern __shared__ __align__(128) unsigned char gsm[];
    GemmStage* stages = reinterpret_cast<GemmStage*>(gsm);
    uint64_t* full = reinterpret_cast<uint64_t*>(gsm + sizeof(GemmStage) * STAGES);
    uint64_t* empty = full + STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_n = (N + GN - 1) / GN, tiles_m = (M + GM - 1) / GM;
    const long tiles = (long)tiles_m * tiles_n;
    const int chunks = (Ktot + KC - 1) / KC;
    const int ng = Ktot / G;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kGemmConsumers);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == kGemmConsumers) {
        // ---------------- producer: table rows by TMA ----------------
        long it = 0;
        for (long t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int n0 = (int)(t % tiles_n) * GN;
            const int cols = min(GN, N - n0);
            for (int c = 0; c < chunks; ++c, ++it) {
                const int s = (int)(it % STAGES);
                const unsigned par = (unsigned)((it / STAGES) & 1);
                const int k0 = c * KC, kc = min(KC, Ktot - k0);
                mbar_wait(&empty[s], par ^ 1u);
                if (lane == 0) mbar_expect_tx(&full[s], (unsigned)(kc * cols * 8));
                __syncwarp();
                // table row k = 16 h + 4 c + s goes to ring row 16 h + 4 s + c: the four lanes of a k-step (c = 0..3)
                // then read four CONSECUTIVE padded rows, i.e. 16 distinct bank pairs per half-warp
                if (lane < kc)
                    tma_bulk_g2s(&stages[s].b[((lane & 16) + 4 * (lane & 3) + ((lane >> 2) & 3)) * B_LD],
                                 T + (long)(k0 + lane) * N + n0, (unsigned)(cols * 8), &full[s]);
            }
        }
        return;
    }
    // ---------------- consumers: 4 x 2 warps, warp tile 16 x 32 = 2 x 4 DMMA tiles ----------------
    const int wm = warp >> 1, wn = warp & 1;
    const int c4 = 4 * (lane & 3);
    const int nblk = (Ktot + 15) >> 4;
    long it = 0;
    for (long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int m0 = (int)(t / tiles_n) * GM, n0 = (int)(t % tiles_n) * GN;
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        const int r0 = wm * 16 + (lane >> 2);                              // tile rows r0 and r0 + 8 of this lane
        const long g0 = min((long)m0 + r0, (long)M - 1), g1 = min((long)m0 + r0 + 8, (long)M - 1);
        const int32_t* row0 = counts + g0 * Ktot + c4;
        const int32_t* row1 = counts + g1 * Ktot + c4;
        const long x0 = (MODE == 3) ? (long)rowmap[g0] : g0, x1 = (MODE == 3) ? (long)rowmap[g1] : g1;
        const double* xrow0 = xf + x0 * Ktot + c4;
        const double* xrow1 = xf + x1 * Ktot + c4;
        double inv0 = 0.0, inv1 = 0.0;
        if (ONE_GROUP) {
            inv0 = inv_tot[g0];
            inv1 = inv_tot[g1];
        }
        int4 cur0 = make_int4(0, 0, 0, 0), cur1 = cur0;
        a4 xc0 = {{0.0, 0.0, 0.0, 0.0}}, xc1 = xc0;
        if (XF) {
            xc0 = ldg_f64x4(xrow0, c4 + 3 < Ktot);
            xc1 = ldg_f64x4(xrow1, c4 + 3 < Ktot);
        } else {
            cur0 = ldg_counts4(row0, c4 + 3 < Ktot);
            cur1 = ldg_counts4(row1, c4 + 3 < Ktot);
        }
        for (int c = 0; c < chunks; ++c, ++it) {
            const int s = (int)(it % STAGES);
            mbar_wait(&full[s], (unsigned)((it / STAGES) & 1));
            const double* bt = stages[s].b + wn * 32 + (lane >> 2);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int blk = 2 * c + half;
                if (blk < nblk) {
                    const int kb = blk * 16;                                // first k of this block
                    const bool more = kb + 16 + c4 + 3 < Ktot;
                    int4 nxt0 = cur0, nxt1 = cur1;
                    a4 xn0 = xc0, xn1 = xc1;
                    if (XF) {
                        xn0 = ldg_f64x4(xrow0 + kb + 16, more);
                        xn1 = ldg_f64x4(xrow1 + kb + 16, more);
                    } else {
                        nxt0 = ldg_counts4(row0 + kb + 16, more);
                        nxt1 = ldg_counts4(row1 + kb + 16, more);
                    }
                    const bool okk = kb + c4 + 3 < Ktot;                    // this lane's four k of the block exist
                    const int cc0[4] = {cur0.x, cur0.y, cur0.z, cur0.w}, cc1[4] = {cur1.x, cur1.y, cur1.z, cur1.w};
#pragma unroll
                    for (int st = 0; st < 4; ++st) {
                        double a0 = XF ? xc0.v[st] : (double)cc0[st], a1 = XF ? xc1.v[st] : (double)cc1[st];
                        if (XF) {
                        } else if (ONE_GROUP) {
                            a0 *= inv0;
                            a1 *= inv1;
                        } else {
                            const int grp = okk ? (kb + c4 + st) / G : 0;
                            a0 *= inv_tot[g0 * ng + grp];
                            a1 *= inv_tot[g1 * ng + grp];
                        }
                        const double* brow = bt + (half * 16 + 4 * st + (lane & 3)) * B_LD;
                        double b[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) b[j] = okk ? brow[j * 8] : 0.0;  // rows past Ktot hold stale bytes
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            dmma_m8n8k4(acc[0][j][0], acc[0][j][1], a0, b[j]);
                            dmma_m8n8k4(acc[1][j][0], acc[1][j][1], a1, b[j]);
                        }
                    }
                    cur0 = nxt0;
                    cur1 = nxt1;
                    xc0 = xn0;
                    xc1 = xn1;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = m0 + wm * 16 + i * 8 + (lane >> 2);
                const int n = n0 + wn * 32 + j * 8 + 2 * (lane & 3);
                double v0 = acc[i][j][0], v1 = acc[i][j][1];
                if (MODE == 2) {
                    v0 *= scale;
                    v1 *= scale;
                    if (clip) {
                        v0 = fmin(fmax(v0, 0.0), 1.0);
                        v1 = fmin(fmax(v1, 0.0), 1.0);
                    }
                }
                if (MODE == 3 && m < M) {
                    const double* fr = fq + (long)rowmap[m] * N + n;
                    if (n < N) v0 = fr[0] / (v0 + kLogGuard);
                    if (n + 1 < N) v1 = fr[1] / (v1 + kLogGuard);
                }
                if (m < M) {
                    if (n + 1 < N) {
                        *reinterpret_cast<double2*>(C + (long)m * N + n) = make_double2(v0, v1);
                    } else if (n < N) {
                        C[(long)m * N + n] = v0;
                    }
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
    // TMA bulk copies need 16-byte rows: Ktot % 4 == 0 and N even hold for every table on the path (6^n, 4^n,
    // S*K with n >= 2 ...); anything else takes the plain tiled kernel
    const bool aligned16 = (Ktot % 4) == 0 && (N % 2) == 0 && ((uintptr_t)counts % 16) == 0 && ((uintptr_t)T % 16) == 0;
    if (!aligned16 || option(QPB_OPT_NO_TMA_GEMM)) {
        dim3 grid((N + GN - 1) / GN, (M + GM - 1) / GM);
        k_gemm_counts_simple<<<grid, 256, 0, st>>>(M, N, Ktot, G, counts, inv_tot, T, C);
        QPB_LAUNCHED("k_gemm_counts_simple");
        return QPB_OK;
    }
    const long tiles = (long)((N + GN - 1) / GN) * ((M + GM - 1) / GM);
    long blocks = (long)num_sms() * 2;
    if (blocks > tiles) blocks = tiles;
    auto kern = ng == 1 ? k_gemm_counts_dmma<0> : k_gemm_counts_dmma<1>;
    QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    kern<<<(int)blocks, kGemmThreads, kGemmSmem, st>>>(M, N, Ktot, G, counts, nullptr, inv_tot, T, C, 1.0, 0, nullptr, nullptr);
    QPB_LAUNCHED("k_gemm_counts_dmma");
    return QPB_OK;
}

// Table transpose for the probability contraction: Mt [D][K] = M [K][D]^T (once per call; the table is at most 2.65 MB)
__global__ void k_transpose_f64(int rows, int cols, const double* __restrict__ in, double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[(long)r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(long)c * rows + r] = tile[threadIdx.x][i];
    }
}

// P [B][K] = clip(scale * X [B][D] * M [K][D]^T) on the FP64 tensor cores: the batched POVM-probability
// contraction (state.py:109-110).  QPB_ERR_UNSUPPORTED when the shape does not fit the pipelined kernel.
int launch_probs_gemm(int K, int D, int B, const double* Mtab, const double* X, double scale, int clip, double* P,
                      cudaStream_t st) {
    if (B < 64 || (D % 4) != 0 || (K % 2) != 0 || ((uintptr_t)X % 16) != 0 || option(QPB_OPT_NO_TMA_GEMM) ||
        option(QPB_OPT_NO_DMMA_GEMM))
        return QPB_ERR_UNSUPPORTED;
    double* Mt = static_cast<double*>(scratch(st, 9, sizeof(double) * (size_t)K * D));
    if (!Mt) return QPB_ERR_NOMEM;
    k_transpose_f64<<<dim3((D + 31) / 32, (K + 31) / 32), dim3(32, 8), 0, st>>>(K, D, Mtab, Mt);
    QPB_LAUNCHED("k_transpose_f64");
    const long tiles = (long)((K + GN - 1) / GN) * ((B + GM - 1) / GM);
    long blocks = (long)num_sms() * 2;
    if (blocks > tiles) blocks = tiles;
    auto kern = k_gemm_counts_dmma<2>;
    QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    kern<<<(int)blocks, kGemmThreads, kGemmSmem, st>>>(B, K, D, D, nullptr, X, nullptr, Mt, P, scale, clip, nullptr, nullptr);
    QPB_LAUNCHED("k_gemm_counts_dmma");
    return QPB_OK;
}

// The two contractions of one batched R.rho.R iteration over the M samples still running (mle_tiled.cu):
//   weights:  W [M][K] = fq[rowmap[m]][k] / (sum_i H2[rowmap[m]][i] ArT[i][k] + 1e-10)      (rowmap, fq given)
//   operator: R [M][D] = sum_k W[m][k] Ar[k][i]                                             (rowmap == nullptr)
// X rows and T rows must be 16-byte aligned, Kd % 4 == 0, N even (4^n and even K always are).
int launch_gemm_f64(int M, int N, int Kd, const double* X, const int* rowmap, const double* fq, const double* T,
                    double* C, cudaStream_t st) {
    QPB_REQUIRE((Kd % 4) == 0 && (N % 2) == 0 && ((uintptr_t)X % 16) == 0 && ((uintptr_t)T % 16) == 0,
                "bad f64 GEMM shape N=%d K=%d", N, Kd);
    if (M <= 0) return QPB_OK;
    const long tiles = (long)((N + GN - 1) / GN) * ((M + GM - 1) / GM);
    long blocks = (long)num_sms() * 2;
    if (blocks > tiles) blocks = tiles;
    auto kern = rowmap ? k_gemm_counts_dmma<3> : k_gemm_counts_dmma<4>;
    QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    kern<<<(int)blocks, kGemmThreads, kGemmSmem, st>>>(M, N, Kd, Kd, nullptr, X, nullptr, T, C, 1.0, 0, rowmap, fq);
    QPB_LAUNCHED("k_gemm_counts_dmma");
    return QPB_OK;
}

}  // namespace qpb
