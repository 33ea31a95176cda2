// Keys-only ascending sort of the bootstrap distances (the `dist.sort()` of quantpy/tomography/interval.py:610):
// one CTA's bitonic network up to 2048 keys, a four-launch sample sort up to 1M keys, CUB's device radix sort
// beyond (on the caller's stream with the library's scratch pool).  QPB_OPT_NO_SAMPLE_SORT selects the earlier
// path (one-CTA bitonic network up to 16384 keys, radix sort above).
#include <cub/device/device_radix_sort.cuh>

#include "../../include/quantpy_b200.h"
#include "common.cuh"

using namespace qpb;

namespace qpb {

constexpr int kBitonicMax = 16384;  // k_sort_bitonic: up to 2048 keys, up to here with QPB_OPT_NO_SAMPLE_SORT
// ------------------------------------------------------------------------------------------------
// 2049 .. 1M keys (the 1e5 distances of BASELINE configs[1], the strong-scaling shards, configs[3] and [4]): a
// sample sort in FOUR short launches.  The device-wide radix sort needs ten (64-bit keys: eight passes), each
// launch-latency bound at this size -- 0.13 ms for 1e5 keys -- and a bitonic network is a chain of ~100 barriers.
//   k_ss_splitters   one CTA sorts 4 * NB strided samples (bitonic network) and keeps every 4th as a splitter
//                    (NB - 1 of them)
//   k_ss_count       bucket of every key by binary search: 2 j for s_{j-1} < x < s_j ("open"), 2 j + 1 for
//                    x == s_j ("equal": needs no sorting, so heavy duplicates cannot overfill an open bucket)
//   k_ss_scatter     exclusive scan of the bucket sizes, keys to their bucket's range of the output
//   k_ss_sort        one CTA per open bucket: bitonic network in shared memory (<= 4096 keys; the mean is <= 512), and
//                    beyond that -- 8x the mean, never seen with sampled splitters -- ranking by counting through a
//                    scratch array
// Keys are compared as order-preserving 64-bit integers (sign-flipped bit patterns, like a radix sort), so NaNs and
// signed zeros have a defined place and the result is a permutation of the input for any bit patterns.
// ------------------------------------------------------------------------------------------------
constexpr int kSsOversample = 4;
constexpr int kSsMaxBuckets = 2048;         // NB (a power of two); 2 NB bucket ids fit 16 bits
constexpr int kSsCountCap = 64;             // open bucket ranked by counting (c^2 comparisons: only when tiny)
constexpr int kSsBucketCap = 4096;          // open bucket sorted by the shared-memory network
constexpr int kSsTile = 1024;               // keys per CTA of the count / scatter kernels (256 threads x 4)
constexpr int kSmallSortMax = 2048;         // up to here ONE CTA's bitonic network (k_sort_bitonic) is the shortest path
constexpr long long kSampleSortMax = 1ll << 20;  // NB * 512 >= n: the mean open bucket holds at most 512 keys

__device__ __forceinline__ unsigned long long ss_enc(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return b ^ ((b >> 63) ? ~0ull : 0x8000000000000000ull);
}
__device__ __forceinline__ double ss_dec(unsigned long long k) {
    const unsigned long long b = k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull);
    return __longlong_as_double((long long)b);
}

// in-place ascending bitonic network over m (power of two) keys in shared memory.  Stages with partner distance
// j >= 64 are separated by block barriers; the stages j = 32 ... 1 of every phase touch only 64 consecutive keys per
// pair group, so each warp runs them on the chunks it owns with warp barriers alone (1024 keys: 20 block barriers
// instead of 55).
__device__ __forceinline__ void ss_bitonic(unsigned long long* keys, int m, int tid, int nt) {
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int k = 2; k <= m; k <<= 1) {
        int j = k >> 1;
        for (; j >= 64; j >>= 1) {
            for (int t = tid; t < (m >> 1); t += nt) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const unsigned long long a = keys[i], b = keys[p];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                    keys[i] = b;
                    keys[p] = a;
                }
            }
            __syncthreads();
        }
        // j <= 32: chunks of 64 keys (32 pairs), one warp each
        for (int c = warp * 64; c < m; c += nw * 64) {
            for (int jj = j; jj > 0; jj >>= 1) {
                const int i = c + (((lane & ~(jj - 1)) << 1) | (lane & (jj - 1)));
                const int p = i | jj;
                if (p < m) {  // (m < 64: the upper lanes have no pair)
                    const unsigned long long a = keys[i], b = keys[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        keys[i] = b;
                        keys[p] = a;
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
}

// Small inputs (C1, C5 two-qubit): the network in shared memory, one launch of one CTA.
__global__ void __launch_bounds__(1024)
k_sort_bitonic(int n, int m, const double* __restrict__ in, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned long long sm_ss[];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < m; i += nt) sm_ss[i] = i < n ? ss_enc(in[i]) : ~0ull;
    __syncthreads();
    ss_bitonic(sm_ss, m, tid, nt);
    for (int i = tid; i < n; i += nt) out[i] = ss_dec(sm_ss[i]);
}

// position of key i among the c keys in shared memory (ties by index): every lane reads the same word (a broadcast)
__device__ __forceinline__ int ss_rank(const unsigned long long* __restrict__ keys, int c, int i) {
    const unsigned long long x = keys[i];
    int lt = 0, eq = 0;
    int j = 0;
    for (; j + 1 < c; j += 2) {  // c is padded to even by the callers' layout: two keys per 16-byte load
        const ulonglong2 y = *reinterpret_cast<const ulonglong2*>(keys + j);
        lt += (y.x < x) + (y.y < x);
        eq += (y.x == x && j < i) + (y.y == x && j + 1 < i);
    }
    if (j < c) {
        const unsigned long long y = keys[j];
        lt += y < x;
        eq += (y == x && j < i);
    }
    return lt + eq;
}

// hist: [2 NB] bucket sizes, then [2 NB] scatter cursors -- zeroed here, before the kernels that add to them
__global__ void __launch_bounds__(1024)
k_ss_splitters(long n, int nb, const double* __restrict__ in, unsigned long long* __restrict__ splitters,
               unsigned int* __restrict__ hist) {
    extern __shared__ __align__(16) unsigned long long sm_ss[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int ns = nb * kSsOversample;
    for (int i = tid; i < 4 * nb; i += nt) hist[i] = 0u;
    for (int i = tid; i < ns; i += nt) sm_ss[i] = ss_enc(in[(long)(((double)i + 0.5) * (double)n / (double)ns)]);
    __syncthreads();
    ss_bitonic(sm_ss, ns, tid, nt);
    for (int j = tid; j < nb - 1; j += nt) splitters[j] = sm_ss[(j + 1) * kSsOversample];
}

// bucket id of key x: j = #splitters < x; 2 j + 1 if x equals splitter j, else 2 j
__device__ __forceinline__ int ss_bucket(unsigned long long x, const unsigned long long* __restrict__ sp, int nb) {
    int lo = 0, hi = nb - 1;  // first index with sp[idx] >= x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sp[mid] < x) lo = mid + 1;
        else hi = mid;
    }
    return 2 * lo + ((lo < nb - 1 && sp[lo] == x) ? 1 : 0);
}

__global__ void __launch_bounds__(256)
k_ss_count(long n, int nb, const double* __restrict__ in, const unsigned long long* __restrict__ splitters,
           unsigned short* __restrict__ bucket_of, unsigned int* __restrict__ hist) {
    extern __shared__ __align__(16) unsigned long long sm_ss[];
    unsigned long long* sp = sm_ss;                                       // [nb]
    unsigned int* cnt = reinterpret_cast<unsigned int*>(sp + nb);         // [2 nb]
    const int tid = threadIdx.x;
    for (int i = tid; i < nb - 1; i += 256) sp[i] = splitters[i];
    for (int i = tid; i < 2 * nb; i += 256) cnt[i] = 0u;
    __syncthreads();
    const long base = (long)blockIdx.x * kSsTile;
#pragma unroll
    for (int r = 0; r < kSsTile / 256; ++r) {
        const long i = base + r * 256 + tid;
        if (i < n) {
            const int b = ss_bucket(ss_enc(in[i]), sp, nb);
            bucket_of[i] = (unsigned short)b;
            atomicAdd(&cnt[b], 1u);
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * nb; i += 256)
        if (cnt[i]) atomicAdd(&hist[i], cnt[i]);
}

// exclusive scan of hist[0 .. nbk) into offs (shared, [nbk + 1]) by one CTA of 256 threads
__device__ __forceinline__ void ss_scan(const unsigned int* __restrict__ hist, int nbk, unsigned int* offs,
                                        unsigned int* part, int tid) {
    const int per = (nbk + 255) / 256;
    unsigned int s = 0;
    for (int q = 0; q < per; ++q) {
        const int i = tid * per + q;
        if (i < nbk) s += hist[i];
    }
    // exclusive scan of the 256 partial sums: warp scans, then the eight warp totals
    const int lane = tid & 31, warp = tid >> 5;
    unsigned int v = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    if (lane == 31) part[warp] = v;
    __syncthreads();
    unsigned int before = 0;
    for (int w = 0; w < warp; ++w) before += part[w];
    unsigned int run = before + v - s;
    for (int q = 0; q < per; ++q) {
        const int i = tid * per + q;
        if (i < nbk) {
            offs[i] = run;
            run += hist[i];
        }
    }
    if (tid == 255) offs[nbk] = run;
    __syncthreads();
}

__global__ void __launch_bounds__(256)
k_ss_scatter(long n, int nb, const double* __restrict__ in, const unsigned short* __restrict__ bucket_of,
             unsigned int* __restrict__ hist, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned long long sm_ss[];
    unsigned int* offs = reinterpret_cast<unsigned int*>(sm_ss);  // [2 nb + 1]
    unsigned int* cnt = offs + 2 * nb + 1;                         // [2 nb]: tile counts, then tile bases
    unsigned int* part = cnt + 2 * nb;                             // [8]
    const int tid = threadIdx.x;
    unsigned int* cursor = hist + 2 * nb;
    for (int i = tid; i < 2 * nb; i += 256) cnt[i] = 0u;
    ss_scan(hist, 2 * nb, offs, part, tid);
    const long base = (long)blockIdx.x * kSsTile;
    int bk[kSsTile / 256];
    unsigned int rk[kSsTile / 256];
#pragma unroll
    for (int r = 0; r < kSsTile / 256; ++r) {
        const long i = base + r * 256 + tid;
        bk[r] = -1;
        rk[r] = 0;
        if (i < n) {
            bk[r] = bucket_of[i];
            rk[r] = atomicAdd(&cnt[bk[r]], 1u);
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * nb; i += 256)
        if (cnt[i]) cnt[i] = atomicAdd(&cursor[i], cnt[i]);  // start of this tile's keys inside bucket i
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSsTile / 256; ++r) {
        const long i = base + r * 256 + tid;
        if (bk[r] >= 0) out[offs[bk[r]] + cnt[bk[r]] + rk[r]] = in[i];
    }
}

__global__ void __launch_bounds__(512)
k_ss_sort(int nb, const unsigned int* __restrict__ hist, double* __restrict__ out, double* __restrict__ spill) {
    extern __shared__ __align__(16) unsigned long long sm_ss[];
    __shared__ unsigned int s_off[2];
    const int tid = threadIdx.x;
    const int bucket = 2 * blockIdx.x;  // open buckets only
    if (tid < 32) {
        // offset of this bucket = sum of the sizes before it
        unsigned int s = 0;
        for (int i = tid; i < bucket; i += 32) s += hist[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (tid == 0) {
            s_off[0] = s;
            s_off[1] = hist[bucket];
        }
    }
    __syncthreads();
    const unsigned int off = s_off[0], cnt = s_off[1];
    if (cnt <= 1) return;
    double* seg = out + off;
    if (cnt <= (unsigned)kSsCountCap) {
        for (int i = tid; i < (int)cnt; i += 512) sm_ss[i] = ss_enc(seg[i]);
        __syncthreads();
        for (int i = tid; i < (int)cnt; i += 512) seg[ss_rank(sm_ss, (int)cnt, i)] = ss_dec(sm_ss[i]);
        return;
    }
    if (cnt <= (unsigned)kSsBucketCap) {
        int m = 2;
        while (m < (int)cnt) m <<= 1;
        for (int i = tid; i < m; i += 512) sm_ss[i] = i < (int)cnt ? ss_enc(seg[i]) : ~0ull;
        __syncthreads();
        ss_bitonic(sm_ss, m, tid, 512);
        for (int i = tid; i < (int)cnt; i += 512) seg[i] = ss_dec(sm_ss[i]);
        return;
    }
    // oversized bucket: rank every key by counting (ties by position) into the spill array, then copy back
    double* tmp = spill + off;
    for (unsigned int i = tid; i < cnt; i += 512) {
        const unsigned long long x = ss_enc(seg[i]);
        unsigned int rank = 0;
        for (unsigned int j = 0; j < cnt; ++j) {
            const unsigned long long y = ss_enc(seg[j]);
            rank += (y < x) || (y == x && j < i);
        }
        tmp[rank] = seg[i];
    }
    __syncthreads();
    for (unsigned int i = tid; i < cnt; i += 512) seg[i] = tmp[i];
}

static int sample_sort(long long n, const double* in, double* out, cudaStream_t st) {
    int nb = 16;
    while (nb < kSsMaxBuckets && (long long)nb * 512 < n) nb <<= 1;
    // scratch: splitters [nb] u64 | hist + cursors [4 nb] u32 | bucket ids [n] u16 | spill [n] f64
    const size_t o_hist = sizeof(unsigned long long) * (size_t)nb;
    const size_t o_ids = o_hist + sizeof(unsigned int) * 4 * (size_t)nb;
    const size_t o_spill = (o_ids + sizeof(unsigned short) * (size_t)n + 255) & ~(size_t)255;
    unsigned char* base = static_cast<unsigned char*>(scratch(st, 5, o_spill + sizeof(double) * (size_t)n));
    if (!base) return QPB_ERR_NOMEM;
    unsigned long long* splitters = reinterpret_cast<unsigned long long*>(base);
    unsigned int* hist = reinterpret_cast<unsigned int*>(base + o_hist);
    unsigned short* ids = reinterpret_cast<unsigned short*>(base + o_ids);
    double* spill = reinterpret_cast<double*>(base + o_spill);
    const int ns = nb * kSsOversample;
    const size_t smem1 = sizeof(unsigned long long) * (size_t)ns;
    if (smem1 > 48 * 1024) QPB_CUDA(cudaFuncSetAttribute(k_ss_splitters, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    k_ss_splitters<<<1, ns / 2 < 1024 ? ns / 2 : 1024, smem1, st>>>((long)n, nb, in, splitters, hist);
    QPB_LAUNCHED("k_ss_splitters");
    const int tiles = (int)((n + kSsTile - 1) / kSsTile);
    k_ss_count<<<tiles, 256, sizeof(unsigned long long) * nb + sizeof(unsigned int) * 2 * nb, st>>>((long)n, nb, in, splitters, ids, hist);
    QPB_LAUNCHED("k_ss_count");
    k_ss_scatter<<<tiles, 256, sizeof(unsigned int) * (4 * nb + 1 + 8) + 8, st>>>((long)n, nb, in, ids, hist, out);
    QPB_LAUNCHED("k_ss_scatter");
    k_ss_sort<<<nb, 512, sizeof(unsigned long long) * kSsBucketCap, st>>>(nb, hist, out, spill);
    QPB_LAUNCHED("k_ss_sort");
    return QPB_OK;
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
    if (n <= (option(QPB_OPT_NO_SAMPLE_SORT) ? kBitonicMax : kSmallSortMax)) {
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
    if (n <= kSampleSortMax && !option(QPB_OPT_NO_SAMPLE_SORT)) return sample_sort(n, in, out, st);
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
