/* include/icikt_b200.h -- C ABI of libicikt_b200.so (CUDA, sm_100a).
 *
 * Drop-in boundary for the ONE hot path of ICIKendallTau: the all-pairs / pair-list
 * information-content-informed Kendall-tau-b.  It replaces, in one batched call,
 *
 *   - the R pair loop            R/kendalltau.R:158 + ici_split R/kendalltau.R:280-308
 *     (and kt_split R/kendalltau.R:310-354 for kt_fast),
 *   - the per-pair FFI call      .Call('_ICIKendallTau_ici_kt', ...)  R/RcppExports.R:62-64,
 *                                SEXP _ICIKendallTau_ici_kt(SEXP x6)  src/RcppExports.cpp:83-96,
 *   - the native pair kernel     ici_kt()  src/kendallc.cpp:166-366 and its helpers :5-129,
 *   - the missing-value marking  setup_missing_matrix  R/utils.R:1-23 (fused into the
 *                                per-column preprocessing kernel).
 *
 * (Paths are relative to the reference repository.)  Plain C: pointers and sizes only,
 * no torch / Rcpp types.  There is NO CPU fallback: every entry point returns
 * ICIKT_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Conventions
 *   data      column-major n x C doubles, leading dimension ld >= n (R matrix layout);
 *             n = features = vector length, C = samples = columns being correlated.
 *   missing   a value is missing if it is NaN/NA, or (na_inf) +-Inf, or equal to one of
 *             the finite literals in global_na[] -- exactly R/utils.R:1-23.  For the
 *             plain ici_kt()/kt_fast() semantics pass n_global_na = 0 and na_inf = 0
 *             (NaN only, src/kendallc.cpp:181,190-191).
 *   pairs     0-based column indices.  icikt_all_pairs enumerates utils::combn(C, 2)
 *             order (0,1),(0,2)...(C-2,C-1) and, if include_diag, appends (0,0)...(C-1,C-1)
 *             exactly as setup_comparisons does when !diag_good (R/kendalltau.R:188-194).
 *   outputs   caller-allocated arrays of length P (pair order above).  Degenerate pairs
 *             get NaN in all four doubles and a non-zero status so the R shim can turn them
 *             into NA_real_ and raise the reference's warnings once per class.
 *   counts    optional int64[P][ICIKT_NCOUNTS]: dis, ntie, xtie, ytie, tot, n_entry, b
 *             (bit-exact against the reference's integer intermediates, for tests).
 */
#ifndef ICIKT_B200_H
#define ICIKT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICIKT_ABI_VERSION 1

/* return codes */
#define ICIKT_OK 0
#define ICIKT_ERR_NO_DEVICE (-1)  /* no CUDA device / driver: there is no CPU fallback     */
#define ICIKT_ERR_BAD_ARG (-2)    /* NULL pointer, n < 1, C < 1, bad pair index, ...       */
#define ICIKT_ERR_TOO_LONG (-3)   /* n > icikt_max_n()                                     */
#define ICIKT_ERR_CUDA (-4)       /* CUDA runtime failure, see icikt_last_error()          */
#define ICIKT_ERR_ALLOC (-5)      /* host or device allocation failed                      */

/* per-pair status (src/kendallc.cpp line of the matching early return) */
#define ICIKT_STATUS_OK 0
#define ICIKT_STATUS_ALL_NA 1        /* :190-199  silent NA                                 */
#define ICIKT_STATUS_TOO_SHORT 2     /* :224-231  "The vectors only have a single value"     */
#define ICIKT_STATUS_SINGLE_UNIQUE 3 /* :234-244  "... have only a single unique value"      */
#define ICIKT_STATUS_ALL_TIED 4      /* :291-298  "Ties equal the total"                     */
#define ICIKT_STATUS_UNSUPPORTED 9   /* complete-observations mode only, see ICIKT_PERSPECTIVE_COMPLETE */
#define ICIKT_NSTATUS 10             /* length of a per-class count array                    */

#define ICIKT_PERSPECTIVE_GLOBAL 0 /* any string other than "local", src/kendallc.cpp:180 */
#define ICIKT_PERSPECTIVE_LOCAL 1
/* Not an ICI perspective: plain Kendall-tau-b on the rows present in BOTH columns, what
 * kt_fast(use = "pairwise.complete.obs") computes pair by pair (kt_split drops the rows missing
 * in either column and calls ici_kt on the rest, R/kendalltau.R:323-341).  completeness is 1,
 * counts are those of the shared rows; status 9 = a column's missing rows tie with its
 * minimum in fp64 (|min| >= ~1e15), the caller has to filter that pair on the host.          */
#define ICIKT_PERSPECTIVE_COMPLETE 2

#define ICIKT_ALT_TWO_SIDED 0
#define ICIKT_ALT_LESS 1
#define ICIKT_ALT_GREATER 2
#define ICIKT_ALT_OTHER 3 /* unknown string: p-value stays 0, src/kendallc.cpp:323-332 */

#define ICIKT_NCOUNTS 7
#define ICIKT_COUNT_DIS 0
#define ICIKT_COUNT_NTIE 1
#define ICIKT_COUNT_XTIE 2
#define ICIKT_COUNT_YTIE 3
#define ICIKT_COUNT_TOT 4
#define ICIKT_COUNT_N_ENTRY 5
#define ICIKT_COUNT_B 6 /* rows missing in both columns */

#define ICIKT_KERNEL_TILED 0 /* one CTA per pair, shared-memory bit-partition counting */
#define ICIKT_KERNEL_NAIVE 1 /* one thread per pair, Fenwick tree in global memory
                                (the internal baseline BASELINE.json names)           */

typedef struct icikt_opts {
  int32_t perspective;  /* ICIKT_PERSPECTIVE_*  (ici_kt argument `perspective`)          */
  int32_t alternative;  /* ICIKT_ALT_*          (ici_kt argument `alternative`)          */
  int32_t continuity;   /* 0/1                  (ici_kt argument `continuity`)           */
  int32_t include_diag; /* all-pairs only: append the C (i,i) pairs (!diag_good)         */
  int32_t na_inf;       /* treat +-Inf as missing (Inf present in global_na)             */
  int32_t device;       /* CUDA device ordinal to run on                                 */
  int32_t kernel;       /* ICIKT_KERNEL_*                                                */
  int32_t want_counts;  /* plan API: allocate and fill the int64 counts (one-shot: set from `counts`) */
  int64_t pair_lo;      /* compute only pairs [pair_lo, pair_hi) of the pair order; the   */
  int64_t pair_hi;      /* output arrays are still indexed from 0 = pair_lo.  0,0 = all.  */
} icikt_opts;

/* A pipelined one-shot call (see icikt_stage_table) overlaps the copies with the kernels: columns_ms,
 * pairs_ms and epilogue_ms are then sums over its launches, h2d_ms runs from the start of the call until
 * the last column has landed, d2h_ms from the first result copy to the last, and the parts exceed total_ms. */
typedef struct icikt_timings { /* milliseconds, CUDA events on the library's streams */
  float h2d_ms;      /* host -> device copy of the data matrix                        */
  float columns_ms;  /* per-column preprocessing kernels (K1)                         */
  float pairs_ms;    /* pair kernel (K2) alone                                        */
  float epilogue_ms; /* fp64 epilogue kernel (K3)                                     */
  float d2h_ms;      /* device -> host copy of the results                            */
  float total_ms;    /* first to last event                                           */
  int32_t n_launches; /* kernels launched by the last columns + pairs calls           */
  int32_t reserved;
} icikt_timings;

/* defaults: global, two.sided, no continuity, no diag, NaN-only missing, device 0, tiled */
void icikt_default_opts(icikt_opts* o);

int icikt_abi_version(void);
int icikt_device_count(void);          /* 0 when no usable device                      */
int64_t icikt_max_n(void);             /* longest supported vector (features)          */
const char* icikt_last_error(void);    /* thread-local message of the last failure     */

/* Replaces R/kendalltau.R:158 for the all-pairs case.  P = C*(C-1)/2 (+ C).
 * pvalue/taumax/completeness/status/counts/max_taumax/timings may be NULL.
 * max_taumax receives max(taumax, na.rm = TRUE) over the computed pairs
 * (R/kendalltau.R:368-370), or NaN if every pair is degenerate.                      */
int icikt_all_pairs(const double* data, int64_t n, int64_t C, int64_t ld,
                    const double* global_na, int32_t n_global_na, const icikt_opts* opts,
                    double* raw, double* pvalue, double* taumax, double* completeness,
                    int32_t* status, int64_t* counts, double* max_taumax,
                    icikt_timings* timings);

/* The same over several GPUs of one box in ONE call (what `computation$ncore` workers are to the
 * reference, R/kendalltau.R:250-253): device k of devices[0..n_devices) preprocesses all columns
 * and computes the k-th contiguous slice of the pair order straight into the caller's arrays;
 * max_taumax is the maximum over the devices.  No collective is involved: pairs are independent
 * (SURVEY.md 8e).  devices == NULL means ordinals 0..n_devices-1; opts->device, pair_lo and
 * pair_hi are ignored.  timings (may be NULL) receives the slowest device's figures.            */
int icikt_all_pairs_multi(const double* data, int64_t n, int64_t C, int64_t ld,
                          const double* global_na, int32_t n_global_na, const icikt_opts* opts,
                          const int32_t* devices, int32_t n_devices, double* raw, double* pvalue,
                          double* taumax, double* completeness, int32_t* status, int64_t* counts,
                          double* max_taumax, icikt_timings* timings);

/* Same for an explicit pair list (include_only, ici_kt(x, y) with C = 2 and P = 1,
 * kt_fast incl. (i,i) pairs).  pi[k], pj[k] in [0, C).                               */
int icikt_pair_list(const double* data, int64_t n, int64_t C, int64_t ld,
                    const double* global_na, int32_t n_global_na, const int32_t* pi,
                    const int32_t* pj, int64_t P, const icikt_opts* opts, double* raw,
                    double* pvalue, double* taumax, double* completeness, int32_t* status,
                    int64_t* counts, double* max_taumax, icikt_timings* timings);

/* ici_kendalltau(return_matrix = TRUE) in one call: icikt_all_pairs (pi == NULL; the diagonal
 * pairs are computed iff !diag_good, like setup_comparisons) or icikt_pair_list (pi, pj, P)
 * followed by icikt_plan_download_matrices; see there for the outputs.                      */
int icikt_matrices(const double* data, int64_t n, int64_t C, int64_t ld, const double* global_na,
                   int32_t n_global_na, const int32_t* pi, const int32_t* pj, int64_t P,
                   const icikt_opts* opts, int32_t scale_max, int32_t diag_good,
                   const int32_t* n_good, double* cor, double* raw, double* pvalue, double* taumax,
                   double* completeness, int64_t* status_counts, double* max_taumax,
                   icikt_timings* timings);

/* icikt_matrices (all pairs) over several GPUs of one box in ONE call: the pair order is sliced over
 * the devices like icikt_all_pairs_multi (sharded K1 + peer gather of the tables), then device k
 * fills columns [C*k/N, C*(k+1)/N) of the five matrices -- reading the per-pair results of the other
 * devices in place over NVLink -- and copies its block straight into the caller's arrays, so the
 * 5 x C x C doubles leave over N PCIe links at once.  Needs peer access between the devices; without
 * it (or with one device) the call runs icikt_matrices on devices[0].  n_devices <= 16.          */
int icikt_matrices_multi(const double* data, int64_t n, int64_t C, int64_t ld, const double* global_na,
                         int32_t n_global_na, const icikt_opts* opts, const int32_t* devices,
                         int32_t n_devices, int32_t scale_max, int32_t diag_good, const int32_t* n_good,
                         double* cor, double* raw, double* pvalue, double* taumax, double* completeness,
                         int64_t* status_counts, double* max_taumax, icikt_timings* timings);

/* pairwise_completeness (R/kendalltau.R:563-629): missing[k] = rows missing in column pi[k] or
 * pj[k] (missing_either, :626-629), completeness[k] = 1 - missing/n (:617).  Here, as in
 * setup_missing_matrix (R/utils.R:1-23), NaN/NA rows are missing only if global_na holds a NaN,
 * +-Inf rows only if it holds an Inf.  pi == pj == NULL: all pairs in combn order followed by
 * the C diagonal pairs (diag_good = FALSE, :586), P = C*(C-1)/2 + C outputs; then `matrix`
 * (C x C, may be NULL) receives the symmetric completeness matrix (:598-605).  missing and
 * completeness may be NULL.  No pair kernel runs: bit masks and popc only.                  */
int icikt_pairwise_completeness(const double* data, int64_t n, int64_t C, int64_t ld,
                                const double* global_na, int32_t n_global_na, int32_t device,
                                const int32_t* pi, const int32_t* pj, int64_t P, int32_t* missing,
                                double* completeness, double* matrix);

/* ---- plan API: keeps the matrix, the per-column tables and the results resident in
 * HBM so that repeated runs (benchmarks, several perspectives on one matrix) do not
 * pay the copies.  A plan is bound to one device and is not thread-safe.            */
typedef struct icikt_plan icikt_plan;

/* pi/pj NULL => all pairs (honours opts->include_diag, pair_lo/pair_hi).            */
int icikt_plan_create(icikt_plan** plan, int64_t n, int64_t C, const int32_t* pi,
                      const int32_t* pj, int64_t P, const icikt_opts* opts);
int64_t icikt_plan_num_pairs(const icikt_plan* plan);
/* copy the host matrix into the plan's device buffer (pinned staging inside)        */
int icikt_plan_upload(icikt_plan* plan, const double* data, int64_t ld);
/* or hand over a device pointer (column-major, ld), e.g. a torch tensor's data_ptr  */
int icikt_plan_set_device_matrix(icikt_plan* plan, const double* d_data, int64_t ld);
/* K1: missing marking, sort, dense ranks, tie sums, per-column tables               */
int icikt_plan_columns(icikt_plan* plan, const double* global_na, int32_t n_global_na);
/* Sharded K1 (one process or worker per GPU): the per-column work of columns
 * [col_lo, col_hi) only.  A rank uploads and preprocesses its slice of the columns, the ranks
 * exchange the table slices icikt_plan_tables describes (column c of table k lives at
 * ptr + c * bytes_per_column, on the plan's device; any transport: NCCL all-gather, peer
 * copies), then icikt_plan_columns_finish derives the launch tier from the statistics of all
 * columns and releases icikt_plan_pairs.  `data`/`ld` address the whole matrix.            */
typedef struct icikt_table {
  void* ptr;                /* device pointer to column 0 of the table */
  int64_t bytes_per_column;
} icikt_table;
#define ICIKT_MAX_TABLES 16
int icikt_plan_upload_columns(icikt_plan* plan, const double* data, int64_t ld, int64_t col_lo,
                              int64_t col_hi);
int icikt_plan_columns_range(icikt_plan* plan, const double* global_na, int32_t n_global_na,
                             int64_t col_lo, int64_t col_hi);
/* fills out[0..min(cap, n)) and returns n, the number of tables to exchange (<= ICIKT_MAX_TABLES) */
int icikt_plan_tables(icikt_plan* plan, icikt_table* out, int32_t cap);
int icikt_plan_columns_finish(icikt_plan* plan);
/* K2+K3 over the plan's pairs; results stay on the device                           */
int icikt_plan_pairs(icikt_plan* plan);
/* block until the plan's stream is idle                                             */
int icikt_plan_sync(icikt_plan* plan);
/* copy results to host arrays (any may be NULL)                                     */
int icikt_plan_download(icikt_plan* plan, double* raw, double* pvalue, double* taumax,
                        double* completeness, int32_t* status, int64_t* counts,
                        double* max_taumax);
/* scale_and_reshape (R/kendalltau.R:357-421) on the device: the results of the plan's pairs as
 * symmetric C x C column-major matrices (each may be NULL): cor = raw / max(taumax) if scale_max
 * else raw (:368-372); entries without a computed pair are 0 (:389-396); if diag_good the diagonal
 * is n_good/max(n_good) in cor and raw, 0 in pvalue, 1 in taumax, n_good/n in completeness
 * (:374-386) with n_good[C] from the caller or, if NULL, n minus the column's missing count.
 * Degenerate pairs (status != 0) hold the NaN with R's NA_real_ bit pattern 0x7FF00000000007A2 in all
 * five matrices; status_counts[ICIKT_NSTATUS] (may be NULL) receives the number of pairs per status
 * class so that the host can raise each warning once.                                     */
int icikt_plan_download_matrices(icikt_plan* plan, int32_t scale_max, int32_t diag_good,
                                 const int32_t* n_good, double* cor, double* raw, double* pvalue,
                                 double* taumax, double* completeness, int64_t* status_counts,
                                 double* max_taumax);
/* per-column by-products: n_na[C] (missing count per column, gives n_good and
 * frac_complete of R/kendalltau.R:165-167); may be NULL                             */
int icikt_plan_column_info(icikt_plan* plan, int32_t* n_na);
/* the CUDA stream the plan launches on, as a cudaStream_t cast to void*             */
void* icikt_plan_stream(icikt_plan* plan);
/* timings of the last upload/columns/pairs/download calls                           */
int icikt_plan_timings(icikt_plan* plan, icikt_timings* t);
void icikt_plan_destroy(icikt_plan* plan);

/* Standard normal CDF exactly as the epilogue kernel evaluates it (R nmath pnorm
 * semantics incl. the exact-zero tails), computed ON THE DEVICE for n values.
 * Test hook for the p-value path; lower_tail as in pnorm().                         */
int icikt_pnorm_device(const double* z, int64_t n, int32_t lower_tail, double* out,
                       int32_t device);

/* Host-only helper: the (i, j) columns of pair `index` in the pair order of icikt_all_pairs
 * (utils::combn(C, 2) order, then the diagonal if include_diag).  Used by callers that shard
 * the pair order with pair_lo/pair_hi or scatter results into C x C matrices.             */
int icikt_pair_from_index(int64_t C, int32_t include_diag, int64_t index, int32_t* i, int32_t* j);

/* The one-shot calls keep their device workspace (tables, result buffers, stream) cached
 * between calls with the same shape so that repeated calls do not pay cudaMalloc; this
 * frees it (the R shim calls it from .onUnload).                                      */
void icikt_release_workspace(void);

/* How the one-shot calls (icikt_all_pairs, icikt_matrices) pipeline a large all-pairs job (input of
 * ICIKT_PIPELINE_MIN_BYTES = 32 MB or more, 128 columns or more; ICIKT_NO_PIPELINE=1 switches it off).
 * The reference hands each furrr worker the whole matrix and a slice of the pair list (R/kendalltau.R:
 * 236-253); here the columns are uploaded in chunks [0,f), [f,2f), [2f,4f) ... and the launch that follows a
 * chunk computes every pair whose later column lies in it, so the upload of the next chunk hides behind
 * it; the last chunk is cut into n_blocks row blocks whose results are contiguous in the pair order and
 * are copied out while the next block runs.  This host-only function reports that table for C columns:
 * units[4k..] = (slot, first column, other column of its first pair, pairs), launches[6l..] = (col_lo,
 * col_hi, unit_lo, unit_hi, slot_lo, slot_hi); cta_slots = resident CTAs of the pair kernel (2 x SMs).
 * Returns the number of units (or a negative error); at most cap_units / cap_launches entries are
 * written (either array may be NULL).  Needs no device.                                         */
int64_t icikt_stage_table(int64_t C, int32_t include_diag, int32_t cta_slots, int32_t n_blocks,
                          int64_t cap_units, int64_t* units, int64_t cap_launches, int64_t* launches,
                          int64_t* n_launches);

/* The pair kernel's launch shape for vectors of n rows (host only, no device needed; what the plan selects
 * in icikt_plan_columns): tier 0 / 1 / 2 = no large tie groups / large groups sorted in place / pass B.
 * out[0] warps per CTA, out[1] 8-key runs per thread, out[2] bytes of the per-CTA region (sequence buffers +
 * counter area), out[3] variant (0 two sequence buffers in shared memory, 1 in place, 2 global scratch),
 * out[4] padded length (keys), out[5] in place: rows per staging part of the gather.  n_sm = SMs of the device
 * (148), complete_obs != 0 for the complete-observations mode.                                            */
int icikt_launch_shape(int64_t n, int32_t tier, int32_t n_sm, int32_t complete_obs, int32_t* out);

/* Measures the shared-memory bandwidth of `device` with a conflict-free read+write sweep
 * (the traffic pattern the roofline model of the pair kernel assumes: one 32-bit load and
 * one 32-bit store per element per level).  Returns GB/s through the pointers (either may
 * be NULL): 32-bit accesses and 128-bit accesses.  Benchmark utility, not on the hot path. */
int icikt_measure_smem_bandwidth(int32_t device, double* gbps_32bit, double* gbps_128bit);

/* Measures the sustained instruction-issue rate of `device`, in G warp-instructions per second over
 * the whole GPU, for the two pipes the pair kernel lives on: LOP3 only (integer ALU pipe), IMAD only
 * (FMA pipe), and the two interleaved (what a kernel that balances them can reach).  The pair
 * kernel's executed warp instructions divided by its run time, over the interleaved figure, is the
 * INT/issue roofline fraction bench.py prints.  Any pointer may be NULL.                        */
int icikt_measure_issue_rate(int32_t device, double* alu_only, double* fma_only, double* interleaved);

#ifdef __cplusplus
}
#endif
#endif /* ICIKT_B200_H */
