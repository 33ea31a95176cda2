// Keys-only ascending sort of the bootstrap distances (the `dist.sort()` of quantpy/tomography/interval.py:610).
// Plumbing, not a hot kernel: CUB's device radix sort on the caller's stream with the library's scratch pool
// (torch.sort also returns the permutation, which costs three times as much for 1e5 doubles).
#include <cub/device/device_radix_sort.cuh>

#include "../../include/quantpy_b200.h"
#include "common.cuh"

using namespace qpb;

namespace qpb {

// Small inputs (one bootstrap shard of up to 16384 distances: C1, the strong-scaling shards): a bitonic network
// in shared memory, ONE launch of one CTA, instead of the half-dozen launches of a device-wide radix sort.
constexpr int kBitonicMax = 16384;
__global__ void __launch_bounds__(1024)
k_sort_bitonic(int n, int m, const double* __restrict__ in, double* __restrict__ out) {
    extern __shared__ double keys[];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < m; i += nt) keys[i] = i < n ? in[i] : __longlong_as_double(0x7ff0000000000000ll);
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (m >> 1); t += nt) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));  // lower index of the t-th compare-exchange pair
                const int p = i | j;
                const double a = keys[i], b = keys[p];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    keys[i] = b;
                    keys[p] = a;
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < n; i += nt) out[i] = keys[i];
}

// R ascending runs -> one ascending array, by counting: element i of run j lands at
//   i + sum_{j' < j} #(run j' <= x) + sum_{j' > j} #(run j' < x)
// (ties ordered by run, so the map is a permutation).  (R - 1) binary searches per element, one launch; the
// multi-GPU quantile step uses it on the all-gathered, locally sorted shards instead of re-sorting R * width keys
// on every rank.
struct RunTable {
    int len[64];
    int start[64];  // offset of run j in `in`
};
__global__ void k_merge_runs(int R, long total, const __grid_constant__ RunTable rt, const double* __restrict__ in,
                             double* __restrict__ out) {
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    // which run does flat output-independent index g belong to (runs enumerated back to back)
    int j = 0;
    long base = 0;
    while (j < R - 1 && g >= base + rt.len[j]) {
        base += rt.len[j];
        ++j;
    }
    const int i = (int)(g - base);
    const double x = in[rt.start[j] + i];
    long pos = i;
    for (int q = 0; q < R; ++q) {
        if (q == j) continue;
        const double* run = in + rt.start[q];
        int lo = 0, hi = rt.len[q];
        if (q < j) {  // # elements <= x
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (run[mid] <= x) lo = mid + 1;
                else hi = mid;
            }
        } else {      // # elements < x
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (run[mid] < x) lo = mid + 1;
                else hi = mid;
            }
        }
        pos += lo;
    }
    out[pos] = x;
}

}  // namespace qpb

extern "C" int qpb_sort_f64(long long n, const double* in, double* out, void* stream) {
    QPB_REQUIRE(n >= 0, "negative length");
    if (n == 0) return QPB_OK;
    QPB_REQUIRE(in && out && in != out, "sort needs distinct input and output buffers");
    QPB_REQUIRE(n < (1ll << 31), "too many keys");
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= kBitonicMax) {
        int m = 2;
        while (m < n) m <<= 1;
        const size_t smem = sizeof(double) * (size_t)m;
        if (smem > 48 * 1024)
            QPB_CUDA(cudaFuncSetAttribute(k_sort_bitonic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int threads = m / 2 < 1024 ? (m / 2 < 32 ? 32 : m / 2) : 1024;
        k_sort_bitonic<<<1, threads, smem, st>>>((int)n, m, in, out);
        QPB_LAUNCHED("k_sort_bitonic");
        return QPB_OK;
    }
    size_t bytes = 0;
    QPB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, in, out, (int)n, 0, 64, st));
    void* tmp = scratch(st, 5, bytes);
    if (!tmp) return QPB_ERR_NOMEM;
    QPB_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, in, out, (int)n, 0, 64, st));
    return check_cuda(cudaGetLastError(), "cub::DeviceRadixSort::SortKeys");  // library kernels: not counted as ours
}

extern "C" int qpb_merge_sorted_runs(int n_runs, const int32_t* run_len_host, const int64_t* run_start_host,
                                     const double* in, double* out, void* stream) {
    QPB_REQUIRE(n_runs >= 1 && n_runs <= 64, "1..64 runs supported, got %d", n_runs);
    QPB_REQUIRE(run_len_host && run_start_host && in && out && in != out, "NULL or aliased buffer");
    RunTable rt;
    long total = 0;
    for (int j = 0; j < n_runs; ++j) {
        QPB_REQUIRE(run_len_host[j] >= 0 && run_start_host[j] >= 0 && run_start_host[j] < (1ll << 31), "bad run %d", j);
        rt.len[j] = run_len_host[j];
        rt.start[j] = (int)run_start_host[j];
        total += run_len_host[j];
    }
    if (total == 0) return QPB_OK;
    k_merge_runs<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n_runs, total, rt, in, out);
    QPB_LAUNCHED("k_merge_runs");
    return QPB_OK;
}
