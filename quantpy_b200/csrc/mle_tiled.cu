// Batched R.rho.R maximum likelihood for UNSTRUCTURED POVMs at n = 3, 4 qubits ('sic', user tables): the two
// K x D contractions of every iteration run as sample-tiled GEMMs on the FP64 tensor cores.
//
//   p_k = Tr(E_k rho) = sum_i ArT[i][k] h2[i]        W [M][K] = f / (H2 [M][D] * ArT [D][K] + 1e-10)     (GEMM 1 + epilogue)
//   R   = sum_k w_k E_k                                R [M][D] = W [M][K] * Ar [K][D]                     (GEMM 2)
//   rho' = R rho R / Tr(R rho R), step norm, stop      one CTA per sample, DMMA products at n = 4          (k_rrr_update_mat)
//
// M = samples still running.  Both GEMMs are k_gemm_counts_dmma (gemm_dmma.cu): persistent 64 x 64 tiles, table
// operand staged by TMA bulk copies into an mbarrier ring, DMMA m8n8k4.  The warp-per-sample kernel this replaces
// (k_mle_rrr_generic, state.cu) streams both tables from L2 once per SAMPLE and iteration (1 MB at n = 4, K = 256);
// here a table tile is read once per 64 samples.  Converged samples are frozen at once (their iteration count and
// state are final) and leave the GEMMs at the next compaction of the row map, every kChunk iterations; that is the
// only host synchronisation (one 4-byte read of the number of running samples).
// Reference: the likelihood of quantpy/tomography/state.py:217-229; oracle: oracle/state.py::mle_rrr.
#include "../../include/quantpy_b200.h"
#include "common.cuh"
#include "plan.h"

namespace qpb {

int launch_gemm_f64(int M, int N, int Kd, const double* X, const int* rowmap, const double* fq, const double* T,
                    double* C, cudaStream_t st);
// rho' = R rho R / Tr, step norm and stopping rule for the running samples (mle_axis.cu: one CTA per sample, DMMA products)
int launch_rrr_update_mat(int n, int M, const int* map, const double* R, double* H, double* H2, int* its, int* done,
                          int max_iter, double tol, cudaStream_t st);

constexpr int kChunk = 8;  // iterations between compactions of the running set

// one warp per sample: frequencies, start state (packed Hermitian) and its off-diagonal-doubled copy
__global__ void k_tiled_init(int d, int K, int B, const int32_t* __restrict__ counts, const double* __restrict__ rho0,
                             double* __restrict__ F, double* __restrict__ H, double* __restrict__ H2,
                             int* __restrict__ map, int* __restrict__ its, int* __restrict__ done) {
    const int dd = d * d, lane = threadIdx.x & 31;
    const long b = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    const int32_t* c = counts + b * K;
    long long tot = 0;
    for (int k = lane; k < K; k += 32) tot += c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    const double total = (double)tot;
    for (int k = lane; k < K; k += 32) F[b * K + k] = (double)c[k] / total;  // state.py:227
    for (int e = lane; e < dd; e += 32) {
        const int a = e / d, bb = e % d;
        double v;
        if (rho0) {
            const double* r0 = rho0 + b * 2 * dd;
            v = (a <= bb) ? r0[2 * (a * d + bb)] : r0[2 * (bb * d + a) + 1];
        } else {
            v = (a == bb) ? 1.0 / d : 0.0;
        }
        H[b * dd + e] = v;
        H2[b * dd + e] = (a == bb) ? v : 2.0 * v;
    }
    if (lane == 0) {
        map[b] = (int)b;
        its[b] = 0;
        done[b] = 0;
    }
}

// compaction of the row map: map_out = the entries of map_in whose sample is still running (block order by arrival:
// the order of the rows has no influence on any result).  *count must be zero on entry.
__global__ void __launch_bounds__(256) k_tiled_compact(int M, const int* __restrict__ map_in, const int* __restrict__ done,
                                                       int* __restrict__ map_out, int* __restrict__ count) {
    __shared__ int warp_off[8];
    __shared__ int base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long a = (long)blockIdx.x * 256 + tid;
    const int b = a < M ? map_in[a] : -1;
    const bool keep = b >= 0 && !done[b];
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_off[warp] = __popc(m);
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) {
            const int c = warp_off[w];
            warp_off[w] = t;
            t += c;
        }
        base = t ? atomicAdd(count, t) : 0;
    }
    __syncthreads();
    if (keep) map_out[base + warp_off[warp] + __popc(m & ((1u << lane) - 1u))] = b;
}

__global__ void k_tiled_finish(int d, int B, const double* __restrict__ H, const int* __restrict__ its,
                               double* __restrict__ rho, int32_t* __restrict__ iters) {
    const int dd = d * d;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * dd) return;
    const long b = i / dd;
    const int e = (int)(i % dd);
    const cplx z = herm_get(H + b * dd, d, e / d, e % d);
    rho[2 * i] = z.re;
    rho[2 * i + 1] = z.im;
    if (iters && e == 0) iters[b] = its[b];
}

bool mle_tiled_applicable(const qpb_state_plan* plan) {
    return plan->n >= 3 && plan->n <= 4 && (plan->K % 2) == 0 && !option(QPB_OPT_NO_TILED_MLE) &&
           !option(QPB_OPT_NO_DMMA_GEMM) && !option(QPB_OPT_NO_TMA_GEMM);
}

int launch_mle_tiled(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                     double tol, double* rho, int32_t* iters, cudaStream_t st) {
    if (!mle_tiled_applicable(plan) || B < 32) return QPB_ERR_UNSUPPORTED;
    const int d = plan->d, D = plan->D, K = plan->K;
    // one scratch block: F [B][K] | W [B][K] | H [B][D] | H2 [B][D] | R [B][D] | map0, map1, its, done [B] | count
    const size_t nd = (size_t)B * D, nk = (size_t)B * K;
    const size_t bytes = sizeof(double) * (2 * nk + 3 * nd) + sizeof(int) * (4 * (size_t)B + 4) + 256;
    unsigned char* base = static_cast<unsigned char*>(scratch(st, 11, bytes));
    if (!base) return QPB_ERR_NOMEM;
    double* F = reinterpret_cast<double*>(base);
    double* W = F + nk;
    double* H = W + nk;
    double* H2 = H + nd;
    double* R = H2 + nd;
    int* map0 = reinterpret_cast<int*>(R + nd);
    int* map1 = map0 + B;
    int* its = map1 + B;
    int* done = its + B;
    int* count = done + B;

    k_tiled_init<<<(B + 7) / 8, 256, 0, st>>>(d, K, B, counts, rho0, F, H, H2, map0, its, done);
    QPB_LAUNCHED("k_tiled_init");
    int M = B;
    int* map = map0;
    int* other = map1;
    for (int done_its = 0; done_its < max_iter && M > 0; done_its += kChunk) {
        const int steps = max_iter - done_its < kChunk ? max_iter - done_its : kChunk;
        for (int s = 0; s < steps; ++s) {
            int rc = launch_gemm_f64(M, K, D, H2, map, F, plan->ArT, W, st);
            if (rc != QPB_OK) return rc;
            rc = launch_gemm_f64(M, D, K, W, nullptr, nullptr, plan->Ar, R, st);
            if (rc != QPB_OK) return rc;
            rc = launch_rrr_update_mat(plan->n, M, map, R, H, H2, its, done, max_iter, tol, st);
            if (rc != QPB_OK) return rc;
        }
        QPB_CUDA(cudaMemsetAsync(count, 0, sizeof(int), st));
        k_tiled_compact<<<(M + 255) / 256, 256, 0, st>>>(M, map, done, other, count);
        QPB_LAUNCHED("k_tiled_compact");
        int m_host = 0;
        QPB_CUDA(cudaMemcpyAsync(&m_host, count, sizeof(int), cudaMemcpyDeviceToHost, st));
        QPB_CUDA(cudaStreamSynchronize(st));
        M = m_host;
        int* t = map;
        map = other;
        other = t;
    }
    const long total = (long)B * D;
    k_tiled_finish<<<(int)((total + 255) / 256), 256, 0, st>>>(d, B, H, its, rho, iters);
    QPB_LAUNCHED("k_tiled_finish");
    return QPB_OK;
}

}  // namespace qpb
