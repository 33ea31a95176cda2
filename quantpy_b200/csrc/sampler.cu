// k2: Philox multinomial shot sampler, one warp per (resample, POVM).
// Replaces np.random.multinomial(n_m, p_m) at quantpy/tomography/state.py:111-114.
//
// Every shot is an independent categorical draw through a Walker/Vose alias table held in shared
// memory (one 32-bit uniform per shot: the high part of u*O picks the column, the low part is the
// acceptance fraction).  Lanes histogram into lane-private shared-memory bins ([O][32] layout:
// bank == lane, no conflicts and no atomics) when O <= 64, else into one per-warp histogram with
// shared-memory atomics.  Counts therefore sum to n_shots exactly, and, like NumPy, the last
// outcome's probability is the remainder 1 - sum(p[:-1]).
//
// RNG stream: Philox4x32-10, key = seed, counter = (draw block j, POVM m, global sample index):
// results do not depend on the batch size or on how samples are sharded across GPUs.
#include "../../include/quantpy_b200.h"
#include "common.cuh"

#include <cstdlib>

namespace qpb {

constexpr int kMaxPovms = 256;
constexpr int kPrivateBinsMaxO = 64;

struct ShotVec {
    int32_t n[kMaxPovms];
};

// Serial Vose construction by one thread.  q (O doubles) and stack (O ints) are scratch.
__device__ void build_alias(const double* __restrict__ p, int O, double* q, int* stack, uint2* table) {
    double head = 0.0;
    for (int o = 0; o + 1 < O; ++o) {
        double v = fmin(fmax(p[o], 0.0), 1.0);
        q[o] = v;
        head += v;
    }
    q[O - 1] = fmax(1.0 - head, 0.0);  // remainder, as NumPy's multinomial does
    const double norm = (double)O / (head + q[O - 1]);
    int ns = 0, nl = 0;  // small stack grows up from 0, large stack grows down from O-1
    for (int o = 0; o < O; ++o) {
        q[o] *= norm;
        if (q[o] < 1.0) stack[ns++] = o;
        else stack[O - 1 - nl++] = o;
    }
    while (ns > 0 && nl > 0) {
        const int s = stack[--ns];
        const int l = stack[O - nl];
        --nl;
        double t = q[s] * 4294967296.0;
        table[s].x = (t >= 4294967295.0) ? 0xffffffffu : (uint32_t)t;
        table[s].y = (uint32_t)l;
        q[l] = (q[l] + q[s]) - 1.0;
        if (q[l] < 1.0) stack[ns++] = l;
        else stack[O - 1 - nl++] = l;
    }
    int fallback = 0;
    for (int o = 1; o < O; ++o)
        if (p[o] > p[fallback]) fallback = o;
    while (nl > 0) {
        const int l = stack[O - nl];
        --nl;
        table[l].x = 0xffffffffu;
        table[l].y = (uint32_t)l;
    }
    while (ns > 0) {  // only reachable through rounding; never hand mass to a zero-probability outcome
        const int s = stack[--ns];
        const bool zero = !(q[s] > 0.5);
        table[s].x = zero ? 0u : 0xffffffffu;
        table[s].y = (uint32_t)(zero ? fallback : s);
    }
}

__global__ void k_alias_build(const double* __restrict__ p, int rows, int O, double* q, int* stack, uint2* table) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    build_alias(p + (long)r * O, O, q + (long)r * O, stack + (long)r * O, table + (long)r * O);
}

__host__ __device__ inline size_t sampler_smem_per_warp(int O, int batched) {
    size_t s = sizeof(uint2) * (size_t)O;
    s += sizeof(uint32_t) * (size_t)O * (O <= kPrivateBinsMaxO ? 32 : 1);
    if (batched) s += (sizeof(double) + sizeof(int)) * (size_t)O;
    return (s + 15) & ~(size_t)15;
}

template <bool PRIVATE>
__device__ __forceinline__ void tally(uint32_t u, int O, const uint2* __restrict__ table, uint32_t* hist, int lane) {
    const uint64_t x = (uint64_t)u * (uint32_t)O;
    const uint32_t col = (uint32_t)(x >> 32), frac = (uint32_t)x;
    const uint2 e = table[col];
    const uint32_t o = (frac < e.x) ? col : e.y;
    if (PRIVATE) hist[o * 32 + lane] += 1;
    else atomicAdd(&hist[o], 1u);
}

template <bool PRIVATE>
__global__ void k_multinomial(int B, int P, int O, const double* __restrict__ p, int batched,
                              const uint2* __restrict__ tables, ShotVec shots, uint32_t k0, uint32_t k1,
                              uint64_t offset, int32_t* __restrict__ counts) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned char* base = smraw + (size_t)warp * sampler_smem_per_warp(O, batched);
    uint2* table = reinterpret_cast<uint2*>(base);
    uint32_t* hist = reinterpret_cast<uint32_t*>(table + O);
    const int nbins = PRIVATE ? O * 32 : O;

    const long items = (long)B * P;
    for (long item = (long)blockIdx.x * nw + warp; item < items; item += (long)gridDim.x * nw) {
        const long b = item / P;
        const int m = (int)(item % P);
        if (batched) {
            double* q = reinterpret_cast<double*>(hist + nbins);
            int* stack = reinterpret_cast<int*>(q + O);
            if (lane == 0) build_alias(p + item * O, O, q, stack, table);
        } else {
            for (int o = lane; o < O; o += 32) table[o] = tables[(long)m * O + o];
        }
        for (int e = lane; e < nbins; e += 32) hist[e] = 0u;
        __syncwarp();
        const uint32_t n = (uint32_t)shots.n[m];
        const uint32_t ncall = (n + 3u) >> 2;
        const uint64_t sample = offset + (uint64_t)b;
        for (uint32_t j = lane; j < ncall; j += 32) {
            const philox4 r = philox4x32_10(j, (uint32_t)m, (uint32_t)sample, (uint32_t)(sample >> 32), k0, k1);
            const uint32_t left = n - 4u * j;  // >= 1
            tally<PRIVATE>(r.x, O, table, hist, lane);
            if (left > 1) tally<PRIVATE>(r.y, O, table, hist, lane);
            if (left > 2) tally<PRIVATE>(r.z, O, table, hist, lane);
            if (left > 3) tally<PRIVATE>(r.w, O, table, hist, lane);
        }
        __syncwarp();
        int32_t* out = counts + item * O;
        for (int o = lane; o < O; o += 32) {
            uint32_t s;
            if (PRIVATE) {
                s = 0;
#pragma unroll 8
                for (int l = 0; l < 32; ++l) s += hist[o * 32 + ((l + lane) & 31)];
            } else {
                s = hist[o];
            }
            out[o] = (int32_t)s;
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// Conditional-binomial multinomial: counts_o ~ Binomial(n_left, p_o / p_left) for o = 0..O-2, the last
// outcome takes what is left -- the decomposition NumPy's multinomial uses (state.py:112), so the work per
// resample is O(O) instead of O(shots).  One thread per (resample, POVM).
// Binomial variates are exact: sequential inversion (BINV) when n*min(p,1-p) < 10, otherwise Hoermann's BTRS
// transformed rejection with squeeze (86 % of proposals are accepted by two comparisons; the rest are
// tested against the exact log-pmf ratio with Stirling tails accurate to 1e-17).
// Uniforms are 53-bit, strictly inside (0,1), two per Philox4x32-10 block.
// ------------------------------------------------------------------------------------------------
struct PhiloxStream {
    uint32_t k0, k1, c1, c2, c3, j;
    double spare;
    bool have;
    __device__ __forceinline__ double next() {
        if (have) {
            have = false;
            return spare;
        }
        const philox4 r = philox4x32_10(j++, c1, c2, c3, k0, k1);
        const double scale = 1.0 / 9007199254740992.0;  // 2^-53
        spare = ((double)(((uint64_t)(r.z >> 5) << 26) | (uint64_t)(r.w >> 6)) + 0.5) * scale;
        have = true;
        return ((double)(((uint64_t)(r.x >> 5) << 26) | (uint64_t)(r.y >> 6)) + 0.5) * scale;
    }
};

// log(k!) - [(k + 1/2) log(k + 1) - (k + 1) + log(2 pi)/2]: table below 16, 6-term Stirling series above
// (truncation error < 2e-18 for k >= 16).
__constant__ double kStirlingTab[16] = {0.08106146679532726,   0.0413406959554093,    0.02767792568499834,   0.020790672103765093,
                            0.016644691189821193,  0.013876128823070748,  0.011896709945891770,  0.010411265261972096,
                            0.009255462182712733,  0.008330563433362871,  0.007573675487951841,  0.006942840107209530,
                            0.006408994188004207,  0.005951370112758848,  0.005554733551962801,  0.005207655919609640};
__device__ __forceinline__ double stirling_tail(double k) {
    if (k < 16.0) return kStirlingTab[(int)k];
    const double ix = fast_recip(k + 1.0), x2 = ix * ix;
    return (1.0 / 12.0 - (1.0 / 360.0 - (1.0 / 1260.0 - (1.0 / 1680.0 - (1.0 / 1188.0 - 691.0 / 360360.0 * x2) * x2) * x2) * x2) * x2) * ix;
}

__constant__ float kStirlingTabF[16] = {0.08106146679532726f,  0.0413406959554093f,   0.02767792568499834f,  0.020790672103765093f,
                            0.016644691189821193f, 0.013876128823070748f, 0.011896709945891770f, 0.010411265261972096f,
                            0.009255462182712733f, 0.008330563433362871f, 0.007573675487951841f, 0.006942840107209530f,
                            0.006408994188004207f, 0.005951370112758848f, 0.005554733551962801f, 0.005207655919609640f};
__device__ __forceinline__ float stirling_tail_f(double k) {  // float32 copy for the prefilter (absolute error < 1e-8)
    if (k < 16.0) return kStirlingTabF[(int)k];
    const float ix = 1.0f / ((float)k + 1.0f), x2 = ix * ix;
    return (1.0f / 12.0f - (1.0f / 360.0f - (1.0f / 1260.0f) * x2) * x2) * ix;
}

// Binomial(n, p) for 0 < p <= 0.5 and n*p < 10: sequential search from 0 (BINV).
__device__ int binomial_inversion(int n, double p, PhiloxStream& rng) {
    const double q = 1.0 - p;
    const double s = p / q, a = ((double)n + 1.0) * s;
    const double r0 = exp((double)n * log(q));
    for (;;) {
        double u = rng.next(), r = r0;
        int x = 0;
        bool ok = true;
        while (u > r) {
            u -= r;
            ++x;
            if (x > n || x > 200) {  // unreachable but for accumulated rounding in the tail: redraw
                ok = false;
                break;
            }
            r *= (a / (double)x - s);
        }
        if (ok) return x;
    }
}

// Parameters of Hoermann's BTRS transformed-rejection sampler ("The generation of binomial random variates",
// J. Stat. Comput. Simul. 46 (1993)); valid for p <= 0.5 and n*p >= 10.
struct Btrs {
    double a, b, c, vr, alpha, r, m, n;
    __device__ __forceinline__ void setup(int n_, double p) {
        // reciprocals and the square root from the hardware seeds + Newton steps (1 ulp): the constants only
        // have to be consistent between proposal and acceptance test, and IEEE div / sqrt were a third of the
        // instructions of a variate
        n = (double)n_;
        const double q = 1.0 - p, npq = n * p * q;
        const double spq = npq * fast_rsqrt(npq);
        b = 1.15 + 2.53 * spq;
        a = -0.0873 + 0.0248 * b + 0.01 * p;
        c = n * p + 0.5;
        const double ib = fast_recip(b);
        vr = 0.92 - 4.2 * ib;
        alpha = (2.83 + 5.1 * ib) * spq;
        r = p * fast_recip(q);
        m = floor((n + 1.0) * p);
    }
    // One proposal (u, v from the lane's stream): >= 0 accepted inside the squeeze (~86 % of proposals), -1 rejected
    // (outside the support), -2 undecided: the candidate kd needs the exact test below.
    __device__ __forceinline__ int propose(PhiloxStream& rng, double& kd, double& v, double& us) const {
        const double u = rng.next() - 0.5;
        v = rng.next();
        us = 0.5 - fabs(u);
        kd = floor((2.0 * a * fast_recip(us) + b) * u + c);
        if (us >= 0.07 && v <= vr) return (int)kd;
        if (kd < 0.0 || kd > n) return -1;
        return -2;
    }
    // exact test of candidate kd against the log-pmf ratio (reciprocals instead of divisions)
    __device__ __forceinline__ bool accept_exact(double kd, double v, double us) const {
        const double lv = log(v * alpha * fast_recip(a * fast_recip(us * us) + b));
        const double nm = n - m + 1.0, nk = n - kd + 1.0;
        const double bound = (m + 0.5) * log((m + 1.0) * fast_recip(r * nm)) + (n + 1.0) * log(nm * fast_recip(nk)) +
                             (kd + 0.5) * log(nk * r * fast_recip(kd + 1.0)) + stirling_tail(m) + stirling_tail(n - m) -
                             stirling_tail(kd) - stirling_tail(n - kd);
        return lv <= bound;
    }
    // The same comparison in float32 with an error bound: 1 accept, 0 reject, -1 too close to call (then
    // accept_exact decides, so the variate is the one the float64 test alone would give -- the counts keep
    // their bits, tests/test_gpu_state.py::test_binomial_prefilter_changes_no_count).  The three log arguments
    // are 1 + d with small d; each d is formed so that its RELATIVE error is a few float ulps (the cancelling
    // numerators in float64, two FMAs), log1pf / logf are 1-ulp functions, so every term t carries at most
    // ~4e-7 |t|; the Stirling tails are < 0.09 each.  delta is that bound with a factor of five to spare.
    // Four float64 logs and four float64 Stirling series were most of the sampler's FP64-pipe time.
    __device__ __forceinline__ int accept_quick(double kd, double v, double us) const {
        const float usf = (float)us;
        const float lv = logf((float)(v * alpha) / ((float)a / (usf * usf) + (float)b));
        const double nm = n - m + 1.0, nk = n - kd + 1.0;
        const double rnm = r * nm;
        const float d1 = (float)((m + 1.0) - rnm) / (float)rnm;                  // (m+1)/(r nm) - 1
        const float d2 = (float)(kd - m) / (float)nk;                            // nm/nk - 1, numerator exact
        const float d3 = (float)fma(nk, r, -(kd + 1.0)) / (float)(kd + 1.0);     // nk r/(kd+1) - 1
        const float t1 = (float)(m + 0.5) * log1pf(d1);
        const float t2 = (float)(n + 1.0) * log1pf(d2);
        const float t3 = (float)(kd + 0.5) * log1pf(d3);
        const float st = (stirling_tail_f(m) + stirling_tail_f(n - m)) - (stirling_tail_f(kd) + stirling_tail_f(n - kd));
        const float bound = ((t1 + t2) + t3) + st;
        const float delta = 2e-6f * (fabsf(t1) + fabsf(t2) + fabsf(t3) + fabsf(lv)) + 1e-5f;
        if (!(delta < 1.0f)) return -1;  // non-finite or absurdly large terms: no float verdict
        if (lv <= bound - delta) return 1;
        if (lv >= bound + delta) return 0;
        return -1;
    }
};

// About one proposal in seven needs the exact test, but with 32 lanes per warp SOME lane does on almost every trip of
// the loop, and the warp then runs the long path for a handful of lanes.  Undecided candidates therefore wait for
// a trip whose number is a multiple of kExactEvery: the lanes of a warp count trips in lockstep, so the long path
// runs less often with more lanes.  A lane's draws are unchanged (its stream is its own), so the counts are
// bit-identical to testing immediately (checked by hashing 1e5 x 36 counts).  Measured per 1e5 x 36 outcomes:
// period 1: 0.195 ms, 2: 0.183 ms, 3: 0.197 ms, 4: 0.216 ms (waiting lanes cost more than the saved long paths).
constexpr int kExactEvery = 2;

// Conditional-binomial chain over "outcomes" [o_begin, o_end) of one (resample, POVM) item -- the whole table
// (MODE 0), its G groups of consecutive outcomes taken as G coarse outcomes (MODE 1), or the outcomes of ONE group
// with the shots the coarse pass gave it (MODE 2).
//
// Two passes instead of one (G > 1): pass 1 splits the item's shots between G groups of ~O/G consecutive outcomes
// (one thread per item, G - 1 binomials over the group masses), pass 2 runs the chain inside every group (one thread
// per (item, group), O/G - 1 binomials).  The law is the same multinomial (conditioning on group totals is exact), the
// number of binomials per item is the same, but a thread's SERIAL chain -- which is what bounds this kernel: ~2.7 us
// per binomial, 35 of them at 36 outcomes -- is G - 1 + O/G - 1 long instead of O - 1, and pass 2 has G times the
// threads to hide its fixed-latency dependencies with.  G depends on O only and every (item, group) has its own
// Philox stream (counter word 1: bits 16-23 carry the pass / group), so a sample's counts depend neither on the batch
// nor on the shard it is drawn in.  The last outcome of the last group takes the remaining mass 1 - sum(p), as NumPy's
// multinomial does (state.py:112); the last outcome of any other group takes what is left of the group's shots.
template <bool PREFILTER, int MODE>
__global__ void __maxnreg__(88) k_multinomial_binomial(int B, int P, int O, int G, const double* __restrict__ p, int batched,
                                       ShotVec shots, uint32_t k0, uint32_t k1, uint64_t offset,
                                       int32_t* __restrict__ counts, int32_t* __restrict__ group_counts,
                                       int exact_every) {
    const long items = (long)B * P;
    const long work = MODE == 2 ? items * G : items;
    const int gsize = (O + G - 1) / G;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < work; t += (long)gridDim.x * blockDim.x) {
        const long item = MODE == 2 ? t / G : t;
        const int g = MODE == 2 ? (int)(t % G) : 0;
        const long b = item / P;
        const int m = (int)(item % P);
        const double* row = p + (batched ? item : (long)m) * O;
        const uint64_t sample = offset + (uint64_t)b;
        PhiloxStream rng;
        rng.k0 = k0; rng.k1 = k1;
        // bit 31: counter domain separate from the alias sampler; bits 16-23: 0 whole table, 255 group pass, 1 + g group g
        rng.c1 = (uint32_t)m | ((uint32_t)(MODE == 0 ? 0 : (MODE == 1 ? 255 : 1 + g)) << 16) | 0x80000000u;
        rng.c2 = (uint32_t)sample; rng.c3 = (uint32_t)(sample >> 32);
        rng.j = 0; rng.have = false; rng.spare = 0.0;
        // probability of "outcome" o: a table entry, or (MODE 1) the mass of group o, summed in index order
        auto prob = [&](int o) -> double {
            if (MODE != 1) return fmin(fmax(row[o], 0.0), 1.0);
            double s = 0.0;
            const int e = min((o + 1) * gsize, O);
            for (int i = o * gsize; i < e; ++i) s += fmin(fmax(row[i], 0.0), 1.0);
            return s;
        };
        int o_begin = 0, o_end = MODE == 1 ? G : O;
        int left = shots.n[m];  // 32-bit: the shot numbers are int32 (conversions to and from double are single instructions)
        double mass = 1.0, po = 0.0;
        int32_t* out = MODE == 1 ? group_counts + item * G : counts + item * O;
        if (MODE == 2) {
            o_begin = min(g * gsize, O);
            o_end = min(o_begin + gsize, O);
            left = group_counts[item * G + g];
            // the group's mass as pass 1 saw it: its own sum, or for the last group what the others leave of 1
            mass = 0.0;
            if (g == G - 1) {
                for (int gg = 0; gg < G - 1; ++gg) {
                    double s = 0.0;
                    for (int i = gg * gsize; i < (gg + 1) * gsize; ++i) s += fmin(fmax(row[i], 0.0), 1.0);
                    mass += s;
                }
                mass = 1.0 - mass;
            } else {
                for (int i = o_begin; i < o_end; ++i) mass += fmin(fmax(row[i], 0.0), 1.0);
            }
        }
        // Per-lane state machine: every trip of the loop is ONE proposal of the lane's current binomial, so
        // the lanes of a warp walk through their outcome sequences independently instead of waiting for the
        // slowest rejection loop at every outcome.
        Btrs st;
        bool flip = false, ready = false, pending = false;
        double cand = 0.0, cv = 0.0, cus = 0.0;  // undecided candidate and its (v, us)
        int o = o_begin, trip = 0;
        while (o + 1 < o_end) {
            ++trip;
            if (!ready) {  // set up the lane's next binomial, then fall through to its first proposal
                po = prob(o);
                int c = -1;
                if (left <= 0 || !(po > 0.0)) {
                    c = 0;
                } else {
                    const double cond = mass > 0.0 ? fmin(po * fast_recip(mass), 1.0) : 1.0;
                    flip = cond > 0.5;
                    const double r = flip ? 1.0 - cond : cond;
                    if (!(r > 0.0)) c = flip ? left : 0;
                    else if ((double)left * r < 10.0) {
                        const int y = binomial_inversion(left, r, rng);
                        c = flip ? left - y : y;
                    } else {
                        st.setup(left, r);
                        ready = true;
                    }
                }
                if (c >= 0) {
                    out[o] = (int32_t)c;
                    left -= c;
                    mass -= po;
                    ++o;
                }
            }
            int y = -1;
            if (ready && !pending) {
                y = st.propose(rng, cand, cv, cus);
                pending = y == -2;
            }
            if (pending && (trip & (exact_every - 1)) == 0) {  // exact_every is a power of two
                pending = false;
                int verdict = PREFILTER ? st.accept_quick(cand, cv, cus) : -1;
                if (verdict < 0) verdict = st.accept_exact(cand, cv, cus) ? 1 : 0;  // a few per 10^4 candidates
                y = verdict ? (int)cand : -1;
            }
            if (y >= 0) {
                const int c = flip ? left - y : y;
                out[o] = (int32_t)c;
                left -= c;
                mass -= po;
                ++o;
                ready = false;
            }
        }
        if (o_end > o_begin) out[o_end - 1] = (int32_t)left;
    }
}

}  // namespace qpb

using namespace qpb;

extern "C" int qpb_multinomial(int B, int P, int O, const double* p, int p_batched, const int32_t* n_shots_host,
                               uint64_t seed, uint64_t offset, int32_t* counts, void* stream) {
    QPB_REQUIRE(B >= 0 && P >= 1 && O >= 1, "bad shape B=%d P=%d O=%d", B, P, O);
    QPB_REQUIRE(P <= kMaxPovms, "P=%d exceeds the supported %d POVMs per tomograph", P, kMaxPovms);
    if (B == 0) return QPB_OK;
    QPB_REQUIRE(p && n_shots_host && counts, "NULL buffer");
    cudaStream_t st = (cudaStream_t)stream;
    ShotVec shots;
    for (int m = 0; m < P; ++m) {
        QPB_REQUIRE(n_shots_host[m] >= 0, "negative shot count");
        shots.n[m] = n_shots_host[m];
    }
    // O(O) conditional binomials beat O(shots) alias draws unless there are very few shots per outcome
    long max_shots = 0;
    for (int m = 0; m < P; ++m) max_shots = shots.n[m] > max_shots ? shots.n[m] : max_shots;
    const int force = option(QPB_OPT_SAMPLER);
    // measured on B200 (tools/bench_configs.py): the binomial chain is sequential in O (0.0084 ms per outcome at
    // 1e5 threads) while the alias kernel scales with the shots (1.2 ms per 1e4 shots x 1e5 warps)
    // (round 2c, two-pass binomial kernel: 216 outcomes x 1e4 shots 0.875 ms against 1.133 ms alias; 1296 outcomes x 1e4
    // shots 1.04 ms against 0.495 ms alias: the crossover is near 30 shots per outcome)
    bool use_binomial = max_shots > 32L * O;
    if (force == 1) use_binomial = false;
    if (force == 2) use_binomial = true;
    if (use_binomial) {
        const long items = (long)B * P;
        int every = option(QPB_OPT_SAMPLER_EXACT_EVERY) > 0 ? option(QPB_OPT_SAMPLER_EXACT_EVERY) : 0;
        const bool prefilter = !option(QPB_OPT_SAMPLER_NO_PREFILTER);
        // groups per item: a function of O alone (the counts must not depend on the batch), about a dozen outcomes per
        // group.  Measured on B200 at 36 outcomes (tools/sampler_sweep.py, profiles/r2c_sampler_groups.log): 1e5 items
        // 0.197 ms as one chain, 0.196 ms in 3 groups, 0.231 in 6 (the launch is issue-bound there: same number of
        // binomials); 12 500 items 0.097 -> 0.069 ms (chain-bound: 2 + 11 binomials per thread instead of 35)
        // (216 outcomes, 1e5 items: 0.99 ms as one chain, 0.90 / 0.87 / 0.88 / 0.95 / 0.98 / 1.27 ms in 2 / 3 / 4 / 6 / 8 / 16 groups)
        int G = O / 12 < 1 ? 1 : (O / 12 > 4 ? 4 : O / 12);
        if (option(QPB_OPT_SAMPLER_LANES) > 0) G = option(QPB_OPT_SAMPLER_LANES);
        QPB_REQUIRE(G >= 1 && G <= 64, "SAMPLER_LANES (groups per item) must be in 1..64");
        while (G > 1 && (G - 1) * ((O + G - 1) / G) >= O - 1) --G;  // the last group keeps two outcomes of its own
        if (every == 0) every = G > 1 ? 1 : kExactEvery;  // short chains: waiting for an even trip costs more than it saves
        while (every & (every - 1)) every &= every - 1;   // the kernel tests trip & (every - 1)
        typedef void (*kern_t)(int, int, int, int, const double*, int, ShotVec, uint32_t, uint32_t, uint64_t, int32_t*,
                               int32_t*, int);
        static const kern_t table[2][3] = {
            {k_multinomial_binomial<false, 0>, k_multinomial_binomial<false, 1>, k_multinomial_binomial<false, 2>},
            {k_multinomial_binomial<true, 0>, k_multinomial_binomial<true, 1>, k_multinomial_binomial<true, 2>}};
        int32_t* group_counts = nullptr;
        if (G > 1) {
            group_counts = static_cast<int32_t*>(scratch(st, 24, sizeof(int32_t) * (size_t)items * G));
            if (!group_counts) return QPB_ERR_NOMEM;
        }
        // Every thread walks its chain from start to end, so a launch that needs one block more than fits takes twice
        // as long: pick the largest block size whose grid is resident at once, else 128 and a grid-stride loop
        static int occ[2][3][3] = {};  // resident blocks per SM for 128 / 96 / 64 threads
        const int cand[3] = {128, 96, 64};
        for (int mode = (G > 1 ? 1 : 0); mode <= (G > 1 ? 2 : 0); ++mode) {
            kern_t kern = table[prefilter][mode];
            const long work = mode == 2 ? items * G : items;
            int threads = option(QPB_OPT_SAMPLER_THREADS) > 0 ? option(QPB_OPT_SAMPLER_THREADS) : 0;
            if (threads == 0) {
                threads = 128;
                for (int c = 0; c < 3; ++c) {
                    if (occ[prefilter][mode][c] == 0) {
                        int nb = 0;
                        QPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, cand[c], 0));
                        occ[prefilter][mode][c] = nb > 0 ? nb : 1;
                    }
                    if ((work + cand[c] - 1) / cand[c] <= (long)num_sms() * occ[prefilter][mode][c]) {
                        threads = cand[c];
                        break;
                    }
                }
            }
            long blocks = (work + threads - 1) / threads;
            const long cap = (long)num_sms() * 16;
            if (blocks > cap) blocks = cap;
            kern<<<(int)blocks, threads, 0, st>>>(B, P, O, G, p, p_batched, shots, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                  offset, counts, group_counts, every);
            QPB_LAUNCHED("k_multinomial_binomial");
        }
        return QPB_OK;
    }
    uint2* tables = nullptr;
    double* q = nullptr;
    int* stack = nullptr;
    if (!p_batched) {
        const size_t cells = (size_t)P * O;
        unsigned char* base = static_cast<unsigned char*>(scratch(st, 1, (sizeof(uint2) + sizeof(double) + sizeof(int)) * cells + 64));
        if (!base) return QPB_ERR_NOMEM;
        q = reinterpret_cast<double*>(base);
        tables = reinterpret_cast<uint2*>(q + cells);
        stack = reinterpret_cast<int*>(tables + cells);
        k_alias_build<<<(P + 31) / 32, 32, 0, st>>>(p, P, O, q, stack, tables);
        QPB_LAUNCHED("k_alias_build");
    }
    const bool priv = O <= kPrivateBinsMaxO;
    int warps = 8;
    size_t per = sampler_smem_per_warp(O, p_batched);
    while (warps > 1 && warps * per > 96 * 1024) warps >>= 1;
    const size_t smem = warps * per;
    QPB_REQUIRE(smem <= 227 * 1024, "O=%d too large for the sampler", O);
    auto kern = priv ? k_multinomial<true> : k_multinomial<false>;
    if (smem > 48 * 1024) QPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long items = (long)B * P;
    long blocks = (items + warps - 1) / warps;
    const long cap = (long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    kern<<<(int)blocks, warps * 32, smem, st>>>(B, P, O, p, p_batched, tables, shots, (uint32_t)seed,
                                                (uint32_t)(seed >> 32), offset, counts);
    QPB_LAUNCHED("k_multinomial");
    return QPB_OK;
}
