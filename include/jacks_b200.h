/*
 * jacks_b200 -- C-ABI of the B200-native JACKS inference engine.
 *
 * This is the drop-in boundary for the JACKS hot path.  Every entry point takes plain
 * pointers and sizes (no torch / numpy types) and returns 0 on success or a non-zero
 * JACKS_E_* code; jacks_last_error() then holds a message for the calling thread.
 * The reference (felicityallen/JACKS) is pure Python, so "the reference's FFI for this
 * path" is the set of Python functions below; INTEGRATION.md shows the ctypes stub a
 * maintainer would add to the reference to route them here.
 *
 *   reference function (file:line)                            entry point(s) here
 *   -------------------------------------------------------   ---------------------------------
 *   jacks/jacks/infer.py:18   inferJACKS (gene loop)           jacks_em_plan_create + jacks_em_run_device,
 *   jacks/jacks/infer.py:55   inferJACKSGene (per-gene EM)     jacks_em_batched_host
 *   jacks/jacks/jacks_io.py:416  pseudo-gene EM                same, with a row-gather CSR index
 *   2018_paper_materials/scripts/jacks/run_JACKS_single.py:83-85
 *                             per-line inferJACKS loop        jacks_em_lines_device / jacks_em_lines_host
 *   jacks/jacks/jacks_io.py:69   computeW1PvalsAndFDRs         jacks_kde_pvalues_device / _host
 *   jacks/jacks/jacks_io.py:402-441  load_data_and_run: inferJACKS on the real genes, inferJACKS on the pseudo-genes,
 *                             computeW1PvalsAndFDRs -- one screen, one call     jacks_run_screen_host / _device
 *   jacks/jacks/preprocess.py:77 normalizeLogCounts            jacks_normalize_host
 *   jacks/jacks/preprocess.py:61 condense_normalised_counts    jacks_condense_host
 *   jacks/jacks/preprocess.py:25 calc_posterior_sd             jacks_condense_host (one sample)
 *   jacks/jacks/preprocess.py:132 loadDataAndPreprocess        jacks_preprocess_host (one count table, one upload);
 *                                                             jacks_normalize_device -> jacks_condense_device
 *
 * All floating point data is IEEE float64.  All arrays are dense, C order.  "device"
 * entry points take pointers into the current CUDA device's memory and enqueue work on
 * the given stream without synchronising; "host" entry points take host pointers, do
 * their own host<->device copies (directly when the buffers are pinned, through an
 * internal pinned ring otherwise) and return when the outputs are complete.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * JACKS_E_CUDA.
 */
#ifndef JACKS_B200_H
#define JACKS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JACKS_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define JACKS_API __attribute__((visibility("default")))
#else
#define JACKS_API
#endif

enum {
    JACKS_OK = 0,
    JACKS_E_INVALID = 1,  /* bad argument (null pointer, negative size, inconsistent CSR ...) */
    JACKS_E_CUDA = 2,     /* CUDA runtime error or no usable device */
    JACKS_E_NOMEM = 3,    /* host or device allocation failed */
    JACKS_E_SHAPE = 4     /* "Expecting matched size between control and test data" (infer.py:26) */
};

/* Hyper-parameters of the per-gene EM; defaults are those of inferJACKS / inferJACKSGene
 * (infer.py:18, :55).  Fill with jacks_em_default_params() and override as needed. */
typedef struct JacksEmParams {
    int32_t n_iter;              /* 50   */
    int32_t apply_w_hp;          /* 0    hierarchical prior on w (infer.py:92) */
    double tol;                  /* 0.1  absolute change of the bound that stops the loop */
    double mu0_x;                /* 1.0  */
    double var0_x;               /* 1.0  */
    double mu0_w;                /* 0.0  */
    double var0_w;               /* 1e4  */
    double tau_prior_strength;   /* 0.5  */
} JacksEmParams;

/* Opaque per-device context: streams, scratch, pinned staging ring. */
typedef struct JacksCtx JacksCtx;
/* Opaque batched-EM plan: the gene -> guide-row CSR index on the device, bucketed by guide count. */
typedef struct JacksEmPlan JacksEmPlan;

/* Message of the last failure on the calling thread ("" if none). */
JACKS_API const char* jacks_last_error(void);
JACKS_API int jacks_abi_version(void);
/* Number of CUDA devices visible (0 when there is no driver / device). Never fails. */
JACKS_API int jacks_device_count(void);

JACKS_API int jacks_ctx_create(int device, JacksCtx** out);
JACKS_API void jacks_ctx_destroy(JacksCtx* ctx);
/* Kernel launches enqueued through this context since creation (bench.py's gpu_launches). */
JACKS_API int64_t jacks_ctx_launch_count(const JacksCtx* ctx);

/* Wait for every asynchronous host call issued on this context (JACKS_ASYNC) and release their plans.
 * Returns the first error of the pipeline, if any. */
JACKS_API int jacks_ctx_synchronize(JacksCtx* ctx);

JACKS_API void jacks_em_default_params(JacksEmParams* p);

/* Pinned host memory for callers that want direct DMA on the host entry points. */
JACKS_API int jacks_host_alloc(void** out, uint64_t bytes);
JACKS_API void jacks_host_free(void* p);
/* Page-lock host memory the CALLER owns -- e.g. a shared-memory segment that the rank processes of one node all map, so
 * that every rank delivers its slice of the per-gene results (the dict inferJACKS returns, infer.py:28) over its own
 * PCIe link -- and undo it.  jacks_copy_to_host_async enqueues a device -> host copy on `stream` (a cudaStream_t). */
JACKS_API int jacks_host_register(void* p, uint64_t bytes);
JACKS_API int jacks_host_unregister(void* p);
JACKS_API int jacks_copy_to_host_async(void* dst_host, const void* src_device, uint64_t bytes, void* stream);

/*
 * Build a plan for a gene index (the gene_index dict of inferJACKS, infer.py:18-21, in CSR form).
 *   gene_offsets[n_genes+1]   int64, gene g owns guide_rows[gene_offsets[g] .. gene_offsets[g+1])
 *   guide_rows[sum G]         int32 row numbers into the [N, L, 2] data cubes; rows may repeat
 *                             (pseudo-genes, jacks_io.py:261-262); every row must be < n_rows
 *   n_lines                   L, number of test cell lines
 * Host pointers.  A gene with no guides is rejected with JACKS_E_INVALID (the reference cannot
 * infer it either: nanmedian over an empty axis).
 */
JACKS_API int jacks_em_plan_create(JacksCtx* ctx, const int64_t* gene_offsets, const int32_t* guide_rows,
                         int64_t n_genes, int64_t n_rows, int32_t n_lines, JacksEmPlan** out);
JACKS_API void jacks_em_plan_destroy(JacksEmPlan* plan);

/* Device buffers of one batched EM run.  Optional pointers may be NULL (output not wanted). */
typedef struct JacksEmDeviceIO {
    const double* test;       /* [N, L, 2]  (mean, sd) per guide and line -- testdata of infer.py:18 */
    const double* ctrl;       /* [N, ctrl_lines, 2]; ctrl_lines == L (matched control, infer.py:66-70)
                                 or 1 (one control column shared by all lines) */
    int32_t ctrl_lines;
    int32_t ctrl_rowmean_fill; /* only with ctrl_lines == 1: fill a NaN control sd with the row mean of the
                                  test sds (1-D control branch, infer.py:61-65) instead of the per-element
                                  test sd (matched branch, infer.py:68) */
    const double* fixed_x1;   /* optional [N]: reference efficacies by row (infer.py:22-23, :77-79) */
    const double* fixed_x2;   /* optional [N] */
    double* x1;               /* optional [sum G]   E[x],   in CSR order */
    double* x2;               /* optional [sum G]   E[x^2] */
    double* w1;               /* required [n_genes, L]  E[w] */
    double* w2;               /* optional [n_genes, L]  E[w^2] */
    double* y;                /* optional [sum G, L]  log-fold changes */
    double* tau;              /* optional [sum G, L]  final noise precisions */
    int32_t* n_iters;         /* optional [n_genes]  EM loop passes executed */
    double* bound;            /* optional [n_genes]  final value of the convergence statistic (infer.py:123) */
} JacksEmDeviceIO;

/* Enqueue the batched EM on `stream` (a cudaStream_t, may be NULL for the default stream). */
JACKS_API int jacks_em_run_device(JacksCtx* ctx, const JacksEmPlan* plan, const JacksEmParams* params,
                        const JacksEmDeviceIO* io, void* stream);

/* Host buffers of one batched EM call; same meaning as JacksEmDeviceIO. */
typedef struct JacksEmHostIO {
    const double* test;
    const double* ctrl;
    int64_t n_rows;           /* N */
    int32_t n_lines;          /* L */
    int32_t ctrl_lines;
    int32_t ctrl_rowmean_fill;
    int32_t reserved;
    const double* fixed_x1;
    const double* fixed_x2;
    double* x1;
    double* x2;
    double* w1;
    double* w2;
    double* y;
    double* tau;
    int32_t* n_iters;
    double* bound;
} JacksEmHostIO;

/*
 * inferJACKS in one call on host buffers (infer.py:18-30): uploads the data cubes, runs the EM for
 * every gene of the CSR index, downloads the requested outputs.  Gene ranges are pipelined over three
 * streams (upload / kernels / download), so with pinned buffers the call costs about max(upload, download).
 * `flags`: JACKS_KEEP_RESIDENT -- the uploaded cubes stay on the device inside ctx and a following call with
 *          test == NULL reuses them (the pseudo-gene pass of jacks_io.py:416 reads the same cubes through another
 *          index); JACKS_ASYNC -- return as soon as the work is enqueued; outputs are complete, and the host
 *          buffers may be reused, after jacks_ctx_synchronize().  A following call on the same context queues
 *          behind it, so the real-gene and pseudo-gene passes of one screen overlap their transfers.
 *          JACKS_UPLOAD_USED_ROWS -- upload only the rows the index references (gathered on the host into a compact cube;
 *          any resident screen stays untouched).  For the pseudo-gene index of jacks_io.py:416, which draws from a few
 *          thousand control guides, that is tens of megabytes instead of the screen: issued on a second context it lets
 *          the pseudo-gene kernels and downloads run while the real-gene call is still uploading.
 */
enum { JACKS_KEEP_RESIDENT = 1, JACKS_ASYNC = 2, JACKS_UPLOAD_USED_ROWS = 4 };
JACKS_API int jacks_em_batched_host(JacksCtx* ctx, const JacksEmParams* params,
                          const int64_t* gene_offsets, const int32_t* guide_rows, int64_t n_genes,
                          const JacksEmHostIO* io, int32_t flags);

/*
 * The whole inference job of one screen in ONE call -- what load_data_and_run does between preprocessing and the
 * writers (jacks_io.py:425-441): inferJACKS on the real genes, inferJACKS on the pseudo-genes drawn from the control
 * genes (same cubes, another index, never fixed_x, jacks_io.py:430), and the KDE p-values of the real genes against the
 * pseudo-gene null (computeW1PvalsAndFDRs, pseudo=True).  The pseudo-gene results never leave the device unless asked
 * for: the reference writes no pseudo-gene to any file (jacks_io.py:131) and needs their p-values only for --fdr.
 *
 * Real-gene outputs as in JacksEmDeviceIO / JacksEmHostIO, plus
 *   pvals         [n_genes, L]   optional (needs n_pseudo >= 2)
 *   pseudo_w1     [n_pseudo, L]  device call: required work buffer when n_pseudo > 0; host call: optional download
 *   pseudo_w2, pseudo_n_iters, pseudo_pvals   optional
 *   positive      selection direction of the p-values (--positive)
 */
typedef struct JacksScreenDeviceIO {
    JacksEmDeviceIO real;      /* cubes, fixed_x (real genes only) and the real-gene outputs */
    double* pvals;
    double* pseudo_w1;
    double* pseudo_w2;
    int32_t* pseudo_n_iters;
    double* pseudo_pvals;
    int32_t positive;
    int32_t reserved;
} JacksScreenDeviceIO;

/* Enqueue on `stream`: EM of plan_real, EM of plan_pseudo (may be NULL), null densities, p-values.  No synchronisation. */
JACKS_API int jacks_run_screen_device(JacksCtx* ctx, const JacksEmPlan* plan_real, const JacksEmPlan* plan_pseudo,
                            const JacksEmParams* params, const JacksScreenDeviceIO* io, void* stream);

typedef struct JacksScreenHostIO {
    JacksEmHostIO real;
    double* pvals;
    double* pseudo_w1;
    double* pseudo_w2;
    int32_t* pseudo_n_iters;
    double* pseudo_pvals;
    int32_t positive;
    int32_t reserved;
} JacksScreenHostIO;

/*
 * Host buffers.  Transfers: the screen is uploaded ONCE and starts going up at once (page-locked cubes); the rows the
 * pseudo-gene index references (a few thousand control guides) are pulled beside it by a kernel into a compact cube, so the
 * pseudo-gene EM -- 85 % of the arithmetic at the Avana shape -- runs while the screen is still uploading; the real genes'
 * EM runs beside it on a second stream, range by range behind their rows, each range's outputs going down as soon as it
 * has run; null densities and p-values follow the last pseudo-gene (DESIGN.md section 6 has the timeline).  Pageable
 * cubes take the same route through a pinned ring, upload after the pseudo-gene enqueue.  A control shared by all lines is
 * passed as [N, 1, 2] (ctrl_lines = 1).  Downloaded: the real-gene outputs that are non-NULL and the pseudo-gene outputs
 * that are non-NULL.
 * flags: JACKS_ASYNC, JACKS_KEEP_RESIDENT as for jacks_em_batched_host.
 */
JACKS_API int jacks_run_screen_host(JacksCtx* ctx, const JacksEmParams* params,
                          const int64_t* gene_offsets, const int32_t* guide_rows, int64_t n_genes,
                          const int64_t* pseudo_offsets, const int32_t* pseudo_rows, int64_t n_pseudo,
                          const JacksScreenHostIO* io, int32_t flags);

/*
 * Per-line ("single JACKS") inference: the reference's
 *     for ts in range(L): inferJACKS(gene_index, testdata[:,[ts],:], ctrldata[:,[ts],:], ...)
 * (2018_paper_materials/scripts/jacks/run_JACKS_single.py:83-85, :92-94) in one launch set -- every (gene, line)
 * pair is an independent EM over that line alone (L = 1: infer.py:55-125; the hierarchical prior of infer.py:92
 * never applies).  Same structs as the joint run; the control must be matched (ctrl_lines == L) and the outputs
 * gain a line axis:
 *     w1, w2, bound [n_genes, L] (double), n_iters [n_genes, L] (int32), x1, x2, y, tau [sum G, L].
 * Column ts of every output equals what the reference's ts-th inferJACKS call returns.
 * The host variant honours JACKS_KEEP_RESIDENT and reuses resident cubes when test == NULL; it is synchronous.
 */
JACKS_API int jacks_em_lines_device(JacksCtx* ctx, const JacksEmPlan* plan, const JacksEmParams* params,
                          const JacksEmDeviceIO* io, void* stream);
JACKS_API int jacks_em_lines_host(JacksCtx* ctx, const JacksEmParams* params,
                        const int64_t* gene_offsets, const int32_t* guide_rows, int64_t n_genes,
                        const JacksEmHostIO* io, int32_t flags);

/*
 * KDE p-values of gene effects against the pseudo-gene null (computeW1PvalsAndFDRs with pseudo=True,
 * jacks_io.py:69-89 + scipy.stats.gaussian_kde.integrate_box_1d): for every line l, bandwidth
 * h_l = std(s_l, ddof=1) * n_l^(-1/5) over the non-NaN pseudo values s_l, and
 *   negative selection: p[g,l] = mean_i( Phi((w[g,l]-s_i)/h_l) - Phi((-100-s_i)/h_l) )
 *   positive selection: p[g,l] = mean_i( Phi(( 100-s_i)/h_l) - Phi((w[g,l]-s_i)/h_l) )
 *   w1_genes [n_genes, L], w1_pseudo [n_pseudo, L], pvals [n_genes, L]
 */
JACKS_API int jacks_kde_pvalues_device(JacksCtx* ctx, const double* w1_genes, int64_t n_genes,
                             const double* w1_pseudo, int64_t n_pseudo, int32_t n_lines,
                             int32_t positive, double* pvals, void* stream);
JACKS_API int jacks_kde_pvalues_host(JacksCtx* ctx, const double* w1_genes, int64_t n_genes,
                           const double* w1_pseudo, int64_t n_pseudo, int32_t n_lines,
                           int32_t positive, double* pvals);

/*
 * General KDE evaluation for the branches of computeW1PvalsAndFDRs the command line does not reach
 * (jacks_io.py:76-88, used by 2018_paper_materials/scripts/jacks/run_JACKS_single.py):
 *   kind 0 / 1  cdf as jacks_kde_pvalues_* (negative / positive selection)
 *   kind 2      density, scipy gaussian_kde.evaluate: sum_i exp(-z_i^2/2) / (n h sqrt(2 pi))  -- the two factors
 *               of the local fdr kernel.evaluate(w)/k_all.evaluate(w) (jacks_io.py:88)
 *   leave_one_out != 0: every w_eval[i, l] is itself a member of the null sample and is left out of the kernel
 *               it is evaluated under (jacks_io.py:81-82): n-1 points, bandwidth recomputed without it.
 * NaN null values are dropped (the reference does so for the shared kernels, jacks_io.py:76-78).
 */
JACKS_API int jacks_kde_eval_host(JacksCtx* ctx, const double* w_eval, int64_t n_eval, const double* w_null, int64_t n_null,
                        int32_t n_lines, int32_t kind, int32_t leave_one_out, double* out);

/*
 * Mean-variance error model for one or more samples (calc_posterior_sd, preprocess.py:25-58, as called from
 * condense_normalised_counts, preprocess.py:61-67).
 *   logcounts [n_guides, n_cols]   normalised log2 counts
 *   sample_col_offsets[n_samples+1], sample_cols[...]   CSR list of the replicate columns of each sample
 *   window_override   <= 0: window = min(max(int(n_guides/100), 30), 800) (preprocess.py:35); > 0: use this
 *                     half-width (calc_posterior_sd's windowfracN argument resolved by the caller)
 *   monotonize        non-zero: enforce the reverse running maximum (do_monotonize=True, the default)
 *   data [n_guides, n_samples, 2]  out: (mean over replicates, posterior sd); sd is NaN for a
 *                                  single-replicate sample (preprocess.py:32)
 * Fails with JACKS_E_INVALID when n_guides <= window (the reference raises IndexError there).
 */
JACKS_API int jacks_condense_host(JacksCtx* ctx, const double* logcounts, int64_t n_guides, int32_t n_cols,
                        const int32_t* sample_col_offsets, const int32_t* sample_cols, int32_t n_samples,
                        int32_t window_override, int32_t monotonize, double* data);

/*
 * Per-column normalisation of log counts, normalizeLogCounts (preprocess.py:77-96), optionally fused with the
 * count -> log2(count + prior) map of the loader (preprocess.py:147).
 *   values [n_guides, n_cols]  counts (take_log != 0; the reference eval()s each cell, so non-integers are
 *                              legal) or log counts (take_log == 0)
 *   log2_table (optional, may be NULL): table_len host-computed values log2(k + prior), k = 0..table_len-1.
 *       Integer counts below table_len take their logarithm from the table, i.e. from the caller's own
 *       libm/numpy, which makes the result bit-identical to the reference run on that host; other counts
 *       use the device log2 (<= 1 ulp).
 *   normtype  JACKS_NORM_MEDIAN       x - nanmedian(x)                                    (:80-81)
 *             JACKS_NORM_ZMAD         (x - nanmedian(x)) / (1.4826 * nanmedian|x - med|)    (:82-84)
 *             JACKS_NORM_CTRL_GUIDES  x - nanmedian(x[row_mask != 0])                      (:92-94)
 *             JACKS_NORM_MODE         x - centre of the fullest bin of the 5-tap smoothed 100-bin histogram (:85-91)
 *   row_mask [n_guides] bytes, required for JACKS_NORM_CTRL_GUIDES only.
 */
enum { JACKS_NORM_MEDIAN = 0, JACKS_NORM_ZMAD = 1, JACKS_NORM_CTRL_GUIDES = 2, JACKS_NORM_MODE = 3 };
JACKS_API int jacks_normalize_host(JacksCtx* ctx, const double* values, int64_t n_guides, int32_t n_cols, int32_t take_log,
                         double prior, const double* log2_table, int64_t table_len, int32_t normtype,
                         const uint8_t* row_mask, double* out);

/*
 * The same two steps on DEVICE buffers, enqueued on `stream` without synchronising, so that they chain without leaving the
 * GPU (normalise -> condense -> jacks_em_run_device).  values / log2_table / row_mask / out / logcounts / data are device
 * pointers; the (small) sample index arrays of jacks_condense_device are HOST arrays.  Scratch lives in the context.
 */
JACKS_API int jacks_normalize_device(JacksCtx* ctx, const double* values, int64_t n_guides, int32_t n_cols, int32_t take_log,
                           double prior, const double* log2_table, int64_t table_len, int32_t normtype,
                           const uint8_t* row_mask, double* out, void* stream);
JACKS_API int jacks_condense_device(JacksCtx* ctx, const double* logcounts, int64_t n_guides, int32_t n_cols,
                          const int32_t* sample_col_offsets, const int32_t* sample_cols, int32_t n_samples,
                          int32_t window_override, int32_t monotonize, double* data, void* stream);

/*
 * loadDataAndPreprocess for one count table in one call (preprocess.py:132-171): log2(count + prior), per-column
 * normalisation, the loader's row sort (row_order[i] = input row that becomes output row i: numpy.argsort of the guide
 * ids, preprocess.py:154-157; NULL keeps the input order) and condense_normalised_counts.  The counts go up ONCE, the
 * [N, S, 2] cube comes down once; the normalised, row-sorted log counts are downloaded only if logcounts_out != NULL.
 * row_mask refers to INPUT rows.  Host pointers.
 */
JACKS_API int jacks_preprocess_host(JacksCtx* ctx, const double* counts, int64_t n_guides, int32_t n_cols, int32_t take_log,
                          double prior, const double* log2_table, int64_t table_len, int32_t normtype,
                          const uint8_t* row_mask, const int64_t* row_order, const int32_t* sample_col_offsets,
                          const int32_t* sample_cols, int32_t n_samples, int32_t window_override, int32_t monotonize,
                          double* logcounts_out, double* data);

/*
 * calc_posterior_sd with a guide subset (preprocess.py:25-58 with guideset_indexs; used by inferMissingVariances,
 * preprocess.py:98-119, for single-replicate samples): the mean-variance curve is fitted on the guides with
 * in_set != 0 (window resolved by the caller from the subset size, preprocess.py:35) and interpolated
 * (np.interp, preprocess.py:50-52) at every guide's mean.
 *   data [n_guides, n_rep] (n_rep >= 2), in_set [n_guides] bytes, sd [n_guides] out.
 */
JACKS_API int jacks_posterior_sd_subset_host(JacksCtx* ctx, const double* data, int64_t n_guides, int32_t n_rep,
                                   const uint8_t* in_set, int32_t window, int32_t monotonize, double* sd);

/*
 * Host-side ingest and formatting (no GPU work).
 *
 * jacks_table_scan: count the non-empty records of a delimiter-separated text file after `skip_lines` lines (the
 *   header and anything above it) and return the byte offset where the data starts.
 * jacks_table_read: parse n_records records starting at data_offset: values[r, i] = numeric value of column
 *   sel_cols[i] (NaN and not_numeric[r, i] = 1 when the cell is not a plain decimal literal -- the reference eval()s
 *   cells, preprocess.py:147, so the caller evaluates just those in Python); label_off/label_len [n_records, n_str] =
 *   byte range inside the file of the text columns str_cols[0..n_str) (guide id, gene name; length -1 when the
 *   record has no such field).  Double-quoted fields follow Python's csv 'excel' dialect.
 * jacks_table_write: header, then per row its label (bytes labels[label_off[r] .. label_off[r+1]), which may itself
 *   contain tabs for multi-column labels) followed by n_cols tab-separated values printed like Python's
 *   '%<width>e' % v ('%5e' in writeJacksWResults / writeJacksXResults, '%6e' in writeFoldChanges,
 *   jacks_io.py:91-156); blank[r, c] != 0 leaves the field empty (non-significant entries under --fdr).
 */
JACKS_API int jacks_table_scan(const char* path, int64_t skip_lines, int64_t* n_records, int64_t* data_offset);
JACKS_API int jacks_table_read(const char* path, int64_t data_offset, char delim, int64_t n_records, int32_t n_sel,
                     const int32_t* sel_cols, double* values, uint8_t* not_numeric, int32_t n_str,
                     const int32_t* str_cols, int64_t* label_off, int32_t* label_len);
JACKS_API int jacks_table_write(const char* path, const char* header, int64_t n_rows, int32_t n_cols, const char* labels,
                      const int64_t* label_off, const double* values, const uint8_t* blank, int32_t width);

/*
 * jacks_pseudo_draw: the draws of createPseudoNonessGenes (jacks/jacks/jacks_io.py:256-264) for n_pseudo pseudo-genes,
 *   consuming the stream of Python's global `random` generator draw for draw: per pseudo-gene one
 *   random.choice over ALL genes (its guide count is taken), then per guide one random.choice over the control genes
 *   and one over that gene's guide rows.  random.choice(seq) is seq[_randbelow(len(seq))] with
 *   _randbelow(n) = rejection sampling of getrandbits(n.bit_length()), i.e. the top bits of one MT19937 output per try
 *   (CPython Lib/random.py, Modules/_randommodule.c).
 *   mt_state [625]: the generator's 624 key words and its position, as random.getstate()[1] holds them; updated in
 *   place to the state after the last draw, so random.setstate() continues the caller's stream exactly where the
 *   reference's loop would have left it.
 *   gene_offsets / guide_rows: the gene index (CSR, dict order); ctrl_genes [n_ctrl]: positions of the control genes in
 *   it, in the iteration order of the control-gene list.  out_offsets [n_pseudo + 1], out_rows [rows_cap]: the
 *   pseudo-gene index (CSR).  *rows_needed = number of rows drawn; when it exceeds rows_cap the call returns
 *   JACKS_E_SHAPE, leaves mt_state untouched and the caller retries with a larger buffer.
 *   An empty sequence under a draw is JACKS_E_INVALID (the reference raises IndexError there).
 */
JACKS_API int jacks_pseudo_draw(uint32_t* mt_state, const int64_t* gene_offsets, const int32_t* guide_rows, int64_t n_genes,
                      const int64_t* ctrl_genes, int64_t n_ctrl, int64_t n_pseudo, int64_t* out_offsets,
                      int32_t* out_rows, int64_t rows_cap, int64_t* rows_needed);

/* FP64 FMA-pipe peak of the current device in FLOP/s (dependent-chain-free DFMA microbenchmark);
 * bench.py uses it as the roofline denominator because MEASURED_PEAKS.json has no FP64 entry. */
JACKS_API int jacks_fp64_peak_flops(JacksCtx* ctx, double* flops_out, double* seconds_out);

#ifdef __cplusplus
}
#endif
#endif /* JACKS_B200_H */
