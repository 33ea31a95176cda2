/*
 * quantpy_b200 C ABI -- B200 (sm_100a) kernels for the tomography-bootstrap hot path of
 * nordmtr/quantpy.  The reference has no FFI; its boundary is a Python class API
 * (quantpy/__init__.py:1-23).  Each entry point below replaces the per-sample arithmetic of one
 * reference method; the Python package quantpy_b200 binds them with ctypes (INTEGRATION.md shows
 * the stub a reference maintainer would add).
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless its name ends in _host.  The caller owns all
 *    buffers.  `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are
 *    asynchronous with respect to the host.
 *  - Return value: 0 on success, negative on error (QPB_ERR_*); qpb_last_error() gives the
 *    message for the calling thread.  No exceptions cross the ABI.  Thread-compatible.
 *  - Reference conventions are kept at the boundary: POVM rows and states are real Bloch
 *    (Pauli-coefficient) vectors in the reference's Pauli order (quantpy/routines.py:14-19),
 *    density/Choi matrices are row-major complex128 (re,im pairs), Choi vectorisation is column
 *    stacking (quantpy/routines.py:53-61).  Counts are int32 (the reference's int64 counts are
 *    narrowed by the Python layer; shots per POVM must be < 2^31).
 *  - n = number of qubits (1..4), d = 2^n, D = 4^n, P POVMs of O outcomes, K = P*O, B = batch.
 */
#ifndef QUANTPY_B200_H
#define QUANTPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QPB_ABI_VERSION 1
#if defined(__GNUC__)
#define QPB_API __attribute__((visibility("default")))
#else
#define QPB_API
#endif

enum { QPB_DIST_HS = 0, QPB_DIST_TRACE = 1, QPB_DIST_IF = 2 };       /* quantpy/geometry.py:5-56 */
enum { QPB_METHOD_LIN = 0, QPB_METHOD_MLE = 1 };                     /* state.py:143-189 */
enum { QPB_INIT_LIN = 0, QPB_INIT_MIXED = 1 };                       /* state.py:206-211 */

QPB_API int qpb_abi_version(void);
QPB_API const char* qpb_last_error(void);
/* Number of kernels launched by this library since load / since the last reset (bench bookkeeping). */
/* Kernel-selection toggles for tests and profiling (never needed in production).  Each is initialised ONCE, when
 * the library is loaded, from the environment variable of the same name with the prefix QPB_ (e.g.
 * QPB_NO_TAIL_MERGE=1); qpb_set_option changes it afterwards.  No launch path reads the environment. */
enum {
    QPB_OPT_NO_TAIL_MERGE = 0,   /* pauli2 MLE: no hand-over of long-running samples to warp-per-sample workers */
    QPB_OPT_NO_HS_FUSION = 1,    /* bootstrap: Hilbert-Schmidt distance by k_distance instead of the MLE write-back */
    QPB_OPT_NO_PAULI_KERNEL = 2,
    QPB_OPT_NO_CONST_KERNEL = 3,
    QPB_OPT_NO_AXIS_KERNEL = 4,
    QPB_OPT_NO_DMMA_GEMM = 5,
    QPB_OPT_NO_ROW_JACOBI = 6,
    QPB_OPT_NO_PACKED_JACOBI = 7,
    QPB_OPT_NO_LIN_SMALL = 8,
    QPB_OPT_SAMPLER = 9,         /* 0 auto | 1 alias | 2 conditional binomial (env: "alias" / "binomial") */
    QPB_OPT_MLE_BLOCKS_PER_SM = 10,
    QPB_OPT_MLE_LANES = 11,      /* pauli2 MLE: 0 auto | 1 thread per sample only | 32 warp per sample only */
    QPB_OPT_NO_TILED_MLE = 12,   /* general-POVM MLE at n >= 3: warp-per-sample kernel instead of the DMMA-tiled one */
    QPB_OPT_MLE_PARK_AGE = 13,   /* pauli2 MLE tuning: iteration count at which a sample moves to a warp-per-sample worker (0 = default) */
    QPB_OPT_MLE_PARK_LIVE = 14,  /* pauli2 MLE tuning: a drained warp with at most this many live samples hands them all over (0 = default) */
    QPB_OPT_MLE_W_WARPS = 15,    /* pauli2 MLE tuning: warp-per-sample workers per CTA from the start (0 = default, -1 = none) */
    QPB_OPT_MLE_PARK_PLATEAU = 16, /* pauli2 MLE tuning: first iteration count at which a non-decreasing step norm hands a sample over (0 = default, -1 = never) */
    QPB_OPT_NO_TMA_GEMM = 17,      /* counts GEMM: plain tiled DMMA kernel instead of the TMA/mbarrier pipeline */
    QPB_OPT_MLE_TAIL_POLL = 18,    /* pauli2 MLE tuning: once the queue is empty, thread-per-sample warps look at the hand-over list every this many iterations (0 = default, -1 = never) */
    QPB_OPT_MLE_TAIL_AGE = 19,     /* pauli2 MLE tuning: ... and give waiting workers their oldest sample if it is at least this old (0 = default) */
    QPB_OPT_MLE_ADOPT = 20,        /* pauli2 MLE tuning: ... or adopt waiting entries into free lanes when at least this many wait (0 = default, -1 = never) */
    QPB_OPT_MLE_MERGE = 21,        /* pauli2 MLE tuning: drained thread-per-sample warps of a CTA pack their samples into fewer warps when the others have at least live + this - 1 free lanes (0 = default, -1 = never) */
    QPB_OPT_NO_MLE_ORDER = 22,     /* fused bootstrap: the pauli2 MLE starts the samples in index order instead of likely-long-runners first */
    QPB_OPT_MLE_PARK_AGE_LO = 23,  /* pauli2 MLE tuning, with a start order: hand-over age of the first queue position (0 = default, -1 = same as MLE_PARK_AGE) */
    QPB_OPT_MLE_PARK_AGE_PCT = 24, /* ... rising linearly to MLE_PARK_AGE over this percentage of the queue (0 = default) */
    QPB_OPT_MLE_PARK_AGE_END = 25, /* ... and falling again for the samples started after the first wave of lanes, down to this age (0 = default: no fall) */
    QPB_OPT_MLE_PARK_AGE_PCT2 = 26,/* ... over this percentage of the remaining queue (0 = default) */
    QPB_OPT_MLE_REFILL_MIN = 27,   /* pauli2 MLE tuning: free lanes of a warp wait for this many before they take new samples (0 = default) */
    QPB_OPT_NO_WARM_JACOBI = 28,   /* CPTP projection (Choi 16 x 16): every eigen-decomposition starts from the identity instead of the previous step's eigenvectors */
    QPB_OPT_MLE_SINGLE_WARPS = 29, /* pauli2 MLE tuning: thread-per-sample warps per CTA, of 12 warps in all (0 = default 8; up to 12) */
    QPB_OPT_SAMPLER_NO_PREFILTER = 30, /* binomial sampler: every undecided BTRS candidate goes to the float64 test (no float32 prefilter); same counts */
    QPB_OPT_SAMPLER_EXACT_EVERY = 31,  /* binomial sampler tuning: undecided candidates are tested on loop trips that are multiples of this (0 = default) */
    QPB_OPT_SAMPLER_THREADS = 32,      /* binomial sampler tuning: threads per block (0 = chosen so that the launch is one wave when it can be) */
    QPB_OPT_NO_SAMPLE_SORT = 33,       /* qpb_sort_f64: the earlier path (one-CTA bitonic network up to 16384 keys, device radix sort above) instead of counting / sample sort */
    QPB_OPT_SAMPLER_LANES = 34,        /* binomial sampler: groups of outcomes per (resample, POVM), one thread each after a pass that splits the shots between the groups (0 = by the number of outcomes, 1 = a single chain); changes the random stream, not the law */
    QPB_OPT_COUNT_ = 35
};
QPB_API int qpb_set_option(int which, int value);
QPB_API int qpb_get_option(int which);
QPB_API int64_t qpb_launch_count(void);
/* Profiling only: per-warp time stamps and work counters of the next k_mle_rrr_pauli2 launches are written to this
 * device buffer (16 int64 words per warp; tools/pauli2_trace.py decodes them).  NULL switches it off (the default). */
QPB_API int qpb_debug_set_trace(void* device_buffer, size_t bytes);
QPB_API void qpb_reset_launch_count(void);

/* Roofline probe (bench.py only): runs a pure FP64 FMA kernel on every SM; *flops_out_host receives the
 * number of floating-point operations it performs so that the caller can time it with CUDA events. */
QPB_API int qpb_fp64_fma_probe(int64_t iters_per_thread, double* sink, double* flops_out_host, void* stream);
/* Shared-memory roofline probe (bench.py only): conflict-free 64-bit loads from every SM; *wavefronts_out_host
 * receives the number of 128-byte shared-memory wavefronts it moves. */
QPB_API int qpb_smem_probe(int64_t iters_per_thread, double* sink, double* wavefronts_out_host, void* stream);

/* ---- state plan ---------------------------------------------------------------------------
 * Device-resident operator tables for one (POVM, shot vector) pair, hoisted out of the
 * per-sample loop.  Inputs are exactly the reference's intermediates:
 *   A [K,D]  shot-weighted POVM rows  povm_matrix*n_m/sum(n)        (state.py:193-196, 221-224)
 *   L [D,K]  _left_inv(A) = (A^T A)^-1 A^T                           (routines.py:69-71); may be
 *            NULL if neither 'lin' nor init='lin' is used.
 * The plan converts both to the packed-Hermitian basis on the device.                         */
typedef struct qpb_state_plan qpb_state_plan;
QPB_API int qpb_state_plan_create(qpb_state_plan** plan, int n_qubits, int K, const double* A, const double* L,
                          void* stream);
QPB_API int qpb_state_plan_destroy(qpb_state_plan* plan);

/* k1: p[b,k] = scale * sum_i M[k,i] r[b,i], optionally clipped to [0,1].
 * Replaces np.einsum("ijk,k->ij", povm_matrix, bloch)*2^n + np.clip (state.py:109-110).        */
QPB_API int qpb_povm_probs(int K, int D, int B, const double* M, const double* r, double scale, int clip,
                   double* p, void* stream);

/* k2: counts[b,m,:] ~ Multinomial(n_shots[m], p[m,:]) for b in [0,B), Philox4x32-10 keyed by
 * (seed, offset+b, m): the draw for a global sample index does not depend on B or on how
 * samples are split across GPUs.  Like np.random.multinomial (state.py:111-114) the last outcome
 * takes the remaining probability mass and sum_o counts[b,m,o] == n_shots[m] exactly.
 * p is [P,O] (p_batched=0, shared by all b) or [B,P,O] (p_batched=1).  n_shots_host is a HOST array. */
QPB_API int qpb_multinomial(int B, int P, int O, const double* p, int p_batched, const int32_t* n_shots_host,
                    uint64_t seed, uint64_t offset, int32_t* counts, void* stream);

/* k3-k5: linear inversion + optional projection onto physical states.
 * Replaces StateTomograph._point_estimate_lin + _make_feasible (state.py:191-202, 267-273).
 * counts [B,K] -> rho [B,d,d] complex128.                                                       */
QPB_API int qpb_lin_project(const qpb_state_plan* plan, int B, const int32_t* counts, int physical, double* rho,
                    void* stream);

/* Fused iterative maximum likelihood rho <- R rho R / Tr(R rho R), R = sum_k f_k/(p_k+1e-10) E_k,
 * stopping per batch element when ||rho' - rho||_F < tol or after max_iter iterations.
 * Stands where the reference runs SciPy BFGS over a Cholesky factor (state.py:204-229); see
 * DESIGN.md for why the update differs (BASELINE.json north_star) and how parity is defined.
 * rho0 [B,d,d] start states (NULL = maximally mixed).  rho [B,d,d] out, iters [B] out (may be NULL). */
QPB_API int qpb_mle_rrr(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0, int max_iter,
                double tol, double* rho, int32_t* iters, void* stream);

/* The same two steps with a START ORDER between them.  R.rho.R iteration counts are heavy-tailed and the persistent
 * kernels hand samples to their lanes through a queue; a launch ends sooner when the likely long runners start
 * first.  qpb_lin_project_ordered (physical = 1 implied: the BFGS / R.rho.R start of state.py:209) also returns
 * that order: ascending smallest POSITIVE eigenvalue of the unprojected estimate (the smaller, the longer the
 * iteration tends to run; negative eigenvalues are clipped by the projection and do not count); the identity when the
 * plan's kernels make no use of it.
 * qpb_mle_rrr_ordered takes the samples from the queue in that order (NULL = index order).  The order is a
 * scheduling hint only: every sample's iterates, iteration count and result are bit-identical to qpb_mle_rrr's.
 * qpb_bootstrap_state does this internally. */
QPB_API int qpb_lin_project_ordered(const qpb_state_plan* plan, int B, const int32_t* counts, double* rho,
                            int32_t* order, void* stream);
QPB_API int qpb_mle_rrr_ordered(const qpb_state_plan* plan, int B, const int32_t* counts, const double* rho0,
                        const int32_t* order, int max_iter, double tol, double* rho, int32_t* iters, void* stream);
QPB_API int qpb_identity_order(int B, int32_t* order, void* stream);

/* Which R.rho.R kernel qpb_mle_rrr dispatches to for this plan (bench.py uses it to count flops):
 * GENERIC: warp per sample, any n<=4 | SMALL: thread per sample, table in shared memory |
 * CONST: thread per sample, table in constant bank, unrolled | PAULI2: two-qubit Pauli-axis POVM, no table |
 * AXIS: 3-4 qubit Pauli-axis POVM, CTA per sample, per-qubit axis maps + fast Pauli transform |
 * TILED: unstructured POVM at 3-4 qubits, both contractions of an iteration as sample-tiled DMMA GEMMs with
 *        TMA-staged tables (batches below 32 samples take GENERIC).  */
enum { QPB_MLE_GENERIC = 0, QPB_MLE_SMALL = 1, QPB_MLE_CONST = 2, QPB_MLE_PAULI2 = 3, QPB_MLE_AXIS = 4, QPB_MLE_TILED = 5 };
QPB_API int qpb_mle_variant(const qpb_state_plan* plan);

/* k8: dist[b] = dst(rho[b], ref) with the reference's argument order dst(estimate, centre)
 * (interval.py:609) and its "below 1e-15 -> 0" rule.  dd is the matrix side (d or d^2 for Choi). */
QPB_API int qpb_distance(int dd, int B, const double* rho, const double* ref, int kind, double* dist, void* stream);

/* Fused bootstrap of interval.py:598-609: for each b sample counts from probs [P,O], reconstruct
 * with `method`, and measure the distance to `ref` [d,d].  Outputs dist [B]; rho_out [B,d,d],
 * counts_out [B,K], iters_out [B] are optional (NULL) -- counts_out doubles as workspace and is
 * required unless the library can fuse the sampler (it cannot yet).  work is a caller-owned
 * scratch buffer of qpb_bootstrap_state_workspace(...) bytes.                                   */
QPB_API size_t qpb_bootstrap_state_workspace(const qpb_state_plan* plan, int B, int P, int O);
QPB_API int qpb_bootstrap_state(const qpb_state_plan* plan, int B, int P, int O, const double* probs,
                        const int32_t* n_shots_host, uint64_t seed, uint64_t offset, int method, int physical,
                        int init, int max_iter, double tol, const double* ref, int dist_kind, double* dist,
                        double* rho_out, int32_t* counts_out, int32_t* iters_out, void* work, void* stream);

/* ---- process path -------------------------------------------------------------------------
 * Linv [d^4, S*K] complex128 = _left_inv(_lifp_oper) (process.py:197-209, plain transpose).
 * counts [B,S,K]; frequencies are normalised per input state (process.py:285).
 * choi [B,d^2,d^2] complex128 out; if cptp!=0 the alternating TP/CP projection of
 * process.py:237-278 runs for at most n_iter iterations with stop value tol; iters [B] optional. */
typedef struct qpb_process_plan qpb_process_plan;
QPB_API int qpb_process_plan_create(qpb_process_plan** plan, int n_qubits, int S, int K, const double* Linv,
                            void* stream);
QPB_API int qpb_process_plan_destroy(qpb_process_plan* plan);
QPB_API int qpb_lifp_cptp(const qpb_process_plan* plan, int B, const int32_t* counts, int cptp, int n_iter, double tol,
                  double* choi, int32_t* iters, void* stream);
/* CPTP projection alone (process.py:231-257) on a batch of Choi matrices, in place allowed. */
QPB_API int qpb_cptp_project(int n_qubits, int B, const double* choi_in, int n_iter, double tol, double* choi_out,
                     int32_t* iters, void* stream);

/* 'states' process estimate (process.py:316-327): choi[b] = sum_s G_s (x) rho[b,s] from reconstructed output
 * states rho [B,S,d,d] and the input-basis coefficients G [S,d,d] (all complex128), followed by the conditional
 * projection: matrices that already satisfy Channel.is_cptp(atol) (channel.py:144-157) are returned unchanged
 * with iters = 0, the others go through the alternating projection.  (third "next" row.) */
QPB_API int qpb_choi_from_states(int n_qubits, int S, int B, const double* G, const double* rho, double* choi,
                         void* stream);
QPB_API int qpb_cptp_project_if_needed(int n_qubits, int B, const double* choi_in, int n_iter, double tol, double atol,
                               double* choi_out, int32_t* iters, void* stream);

/* ---- polytope coverage experiments (first "next" row) ---------------------
 * Per-trial body of test_qst / test_qpt (quantpy/tomography/polytopes/verification.py:9-78): for every
 * trial b and confidence level j, delta = count_delta(level_j, frequencies_b, n) by the bisection of
 * quantpy/tomography/polytopes/utils.py:16-27, and inside[b,j] = min_k(freq+delta - p_true) > -1e-15.
 * The trial's tables are either counts [B,M,O] (frequencies = clip(counts/n, 1e-15, 1-1e-15)) or already
 * clipped frequencies freq [B,M,O]; exactly one of the two is non-NULL.  n_shots_host [M] doubles (HOST).
 * levels [L], p_true [M*O] (may be NULL with inside_out NULL), delta_out [B,L], inside_out [B,L] bytes.
 * clip_b: clip freq+delta to [1e-15, 1-1e-15] as test_qst does (test_qpt does not).                      */
QPB_API int qpb_polytope_coverage(int B, int M, int O, const int32_t* counts, const double* freq,
                          const double* n_shots_host, int L, const double* levels, const double* p_true, int clip_b,
                          double* delta_out, unsigned char* inside_out, void* stream);
/* conf_out[b,j] = count_confidence(deltas[j], freq[b], n)  (utils.py:4-13). */
QPB_API int qpb_polytope_confidence(int B, int M, int O, const double* freq, const double* n_shots_host, int L,
                            const double* deltas, double* conf_out, void* stream);

/* ---- moment intervals (second "next" row) -----------------------------
 * mean_out[b], var_out[b] = l2_mean, l2_variance of quantpy/stats.py:5-53 for the frequency tables
 * freq [B,P,O] with the weight tensor weights [P,O,P,O] (interval.py:89) and n_trials shots per POVM.    */
QPB_API int qpb_l2_moments(int B, int P, int O, const double* weights, const double* freq, double n_trials,
                   double* mean_out, double* var_out, void* stream);

/* ---- Metropolis-Hastings likelihood sampling (fourth "next" row) ---------
 * MHMC.sample with normalized_update and a symmetric standard-normal jump (quantpy/mhmc.py:48-119) on the target
 * exp(-StateTomograph._nll(x)) (state.py:217-229), as MHMCStateInterval.setup runs it (interval.py:737-759).
 * C independent chains, one warp each; chain c starts at the packed Cholesky vector x_init[c] [D]
 * (routines.py:84-91), makes burn_steps steps, then n_samples*thinning steps of which every thinning-th state
 * (i % thinning == 0) is returned as L L^dagger in samples [C, n_samples, d, d] complex128.
 * counts [K] (counts_batched = 0) or [C, K]; accepted [C] = accepted proposals of the sampling phase (may be NULL);
 * x_final [C, D] = last state, for warm starts (may be NULL).
 * Noise: deltas [C, burn_steps + n_samples*thinning, D] and uniforms [C, burn_steps + n_samples*thinning] from the
 * caller (both or neither); when NULL it is generated in the kernel from Philox4x32-10 keyed by seed with counter
 * (chain_offset + c, step).                                                                              */
QPB_API int qpb_mhmc_state(const qpb_state_plan* plan, int C, int n_samples, int thinning, int burn_steps, double step,
                   const int32_t* counts, int counts_batched, const double* x_init, const double* deltas,
                   const double* uniforms, uint64_t seed, uint64_t chain_offset, double* samples, int32_t* accepted,
                   double* x_final, void* stream);

/* ---- quantile step ----------------------------------------------------------------------------
 * out = sorted(in) ascending, n float64 keys, in != out (`dist.sort()`, interval.py:610 / 683).  Keys only: the
 * quantile function needs the order statistics, not the permutation.                                   */
QPB_API int qpb_sort_f64(long long n, const double* in, double* out, void* stream);
/* Multi-GPU quantile step: every rank sorts its shard, ONE all-gather collects the sorted shards, and this call
 * merges the n_runs (<= 64) ascending runs in[run_start[j] .. run_start[j] + run_len[j]) into `out` (ascending,
 * sum of the lengths) -- the sorted array the reference gets from `dist.sort()` on the concatenation
 * (interval.py:610, 683), without re-sorting world_size * shard keys on every rank.  in != out.            */
QPB_API int qpb_merge_sorted_runs(int n_runs, const int32_t* run_len_host, const int64_t* run_start_host,
                                  const double* in, double* out, void* stream);

/* cl_to_dist(levels) of interval.py:611-612 -- scipy's interp1d(linspace(0, 1, n), sorted)(levels) -- on a sorted
 * array that lives on the device: HOST levels in, HOST quantiles out (2 * 8 * n_levels bytes cross PCIe, staged
 * through page-locked memory owned by the library); returns after the stream has been synchronised.  A level outside
 * [0, 1] is QPB_ERR_INVALID with interp1d's message.  Evaluates y[lo] + (y[lo+1] - y[lo]) * (pos - lo),
 * pos = level * (n - 1), without contraction: the same bits as the NumPy expression on the same array.       */
QPB_API int qpb_quantiles_host(long long n, const double* sorted_dev, int n_levels, const double* levels_host,
                               double* out_host, void* stream);
/* BootstrapStateInterval.setup() + cl_to_dist(levels) (interval.py:583-612) on one GPU as ONE call with HOST
 * inputs and outputs: Bloch vector [D] and matrix [d,d] complex of the centre state, shot vector and confidence
 * levels are host arrays; the call uploads them (one copy), computes the probabilities clip(2^n M.r) from the
 * unweighted POVM rows M_dev [K,D] (device, state.py:109-110), runs qpb_bootstrap_state on library-owned scratch,
 * sorts the distances into dist_sorted [B] (device, stays there: interval.py:610), interpolates the levels and
 * returns the quantiles in quantiles_host after synchronising the stream.  iters_out [B] (device) is optional.
 * n_levels may be 0 (setup only): there is then no host output and the call returns with its work QUEUED on the
 * stream, not finished -- the multi-GPU interval queues its all-gather and merge behind it and synchronises once,
 * in qpb_quantiles_host (the host inputs have been copied to library-owned page-locked memory by then).   */
QPB_API int qpb_bootstrap_state_interval(const qpb_state_plan* plan, int B, int P, int O, const double* M_dev,
                                 const double* bloch_host, const double* ref_host, const int32_t* n_shots_host,
                                 uint64_t seed, uint64_t offset, int method, int physical, int init, int max_iter,
                                 double tol, int dist_kind, int n_levels, const double* levels_host,
                                 double* quantiles_host, double* dist_sorted, int32_t* iters_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QUANTPY_B200_H */
