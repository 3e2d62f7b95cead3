/*
 * pagan2_b200.h -- C-ABI of the B200-native PAGAN2 pairwise sequence-graph Viterbi engine.
 *
 * This is the drop-in boundary for ONE path of ariloytynoja/pagan2-msa: the DP fill over the match /
 * x-gap / y-gap log-probability matrices of two sequence graphs and its traceback, i.e. what
 * Viterbi_alignment::align does between src/main/viterbi_alignment.cpp:238 and :383.  The reference
 * has no FFI; the seam is the C++ class surface Node uses (node.cpp:77-159).  A maintainer keeps
 * viterbi_alignment.cpp:191-231 (settings) and :389-392 (build_ancestral_sequence) and replaces the
 * matrix allocation, fill loops, end-corner scan and backtrack by pg2_align_batch() + pg2_expand_path()
 * (see INTEGRATION.md for the stub).
 *
 * Plain C: pointers and sizes only, caller-owned buffers, integer status codes, no exceptions cross
 * this boundary.  All host pointers may be pageable or pinned.  A pg2_ctx is bound to one CUDA device
 * and is not thread-safe; use one ctx per host thread (contexts are independent).
 *
 * There is NO CPU fallback: every entry point that computes fails with PG2_ERR_NO_DEVICE /
 * PG2_ERR_CUDA when the device is unavailable.
 */
#ifndef PAGAN2_B200_H
#define PAGAN2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG2_ABI_VERSION 4

/* ---- status codes ------------------------------------------------------------------------- */
enum {
    PG2_OK = 0,
    PG2_ERR_INVALID = 1,     /* bad argument (null pointer, negative size, malformed CSR, bad handle) */
    PG2_ERR_NO_DEVICE = 2,   /* no CUDA device / wrong architecture (needs sm_100) */
    PG2_ERR_CUDA = 3,        /* a CUDA runtime call failed; pg2_last_error() has the text */
    PG2_ERR_NOMEM = 4,       /* host or device allocation failed */
    PG2_ERR_UNSUPPORTED = 5, /* graph exceeds a packed-pointer limit (in-degree > PG2_MAX_IN_DEGREE) */
    PG2_ERR_CAPACITY = 6     /* caller's step buffer too small; results[].n_steps tell the need */
};

/* per-job status (pg2_result.status) */
enum {
    PG2_JOB_OK = 0,
    PG2_JOB_NO_PATH = 1,     /* end corner is -inf: the band admits no path.  The reference then refills
                                (viterbi_alignment.cpp:298-317); the caller re-submits without band. */
    PG2_JOB_BAD_BAND = 2,    /* band bounds not monotone non-decreasing (tunnel_matrix.h:163-168) */
    PG2_JOB_BAD_GRAPH = 3,   /* an edge does not point to an earlier site / a state is out of range */
    PG2_JOB_BROKEN_PATH = 4  /* traceback met a cell without a predecessor (reference: exit(1),
                                viterbi_alignment.cpp:1167-1171) */
};

/* Basic_alignment::Matrix_pt (basic_alignment.h:107) */
enum { PG2_X_MAT = 0, PG2_Y_MAT = 1, PG2_M_MAT = 2 };

/* job flags */
#define PG2_FLAG_NO_TERMINAL_EDGES 1u            /* --no-terminal-edges (viterbi_alignment.cpp:866,877) */
#define PG2_FLAG_REDUCED_TERMINAL_GAP_PENALTIES 2u /* !--no-reduced-terminal-penalties (basic_alignment.h:627) */

#define PG2_MAX_IN_DEGREE 63 /* backward edges per site representable in a packed traceback pointer */

/* ---- inputs -------------------------------------------------------------------------------- */

/* One sequence graph (reference: Sequence = vector<Site> + vector<Edge>, sequence.h:663-671) as a CSR
 * of BACKWARD edges.  Site 0 is the start site, site n_sites-1 the stop site.  For site s its edges
 * are k in [bwd_off[s], bwd_off[s+1]) IN THE REFERENCE'S LIST ORDER (Site::get_first_bwd_edge /
 * get_next_bwd_edge, sequence.h:395-417) -- the order decides ties.  Every edge_start[k] < s.
 *
 * COMPACT FORM of a plain chain (a leaf or a read as Sequence::create_default_sequence builds it, sequence.cpp:152-303:
 * site s >= 1 is entered by the one edge (s-1 -> s) of weight 1): bwd_off == edge_start == edge_logw == NULL and
 * n_edges == n_sites - 1.  The engine then reads only `state` (nothing else is packed or uploaded for the graph).
 * edge_index may still be given for pg2_expand_path; NULL means edge k has index k + 1 (edge 0 of a default sequence
 * is the dummy first edge). */
typedef struct pg2_graph {
    int32_t n_sites;
    int32_t n_edges;           /* == bwd_off[n_sites] */
    const int32_t *state;      /* [n_sites]  Site::character_state (start/stop sites: -1, never read) */
    const int32_t *bwd_off;    /* [n_sites+1] */
    const int32_t *edge_start; /* [n_edges]  Edge::start_site_index */
    const float *edge_logw;    /* [n_edges]  Edge::log_posterior_weight (float, sequence.h:43) */
    const int32_t *edge_index; /* [n_edges]  Edge::index in the owner's edge vector (for is_used marks) */
} pg2_graph;

/* Substitution/gap parameters of one Evol_model (evol_model.h:55-86), already FLOAT-rounded exactly as
 * the reference's accessors return them. */
typedef struct pg2_model_desc {
    int32_t fas;               /* full alphabet size: 15 DNA, 211 protein, 1892 codon */
    const float *log_score;    /* [fas*fas] column-major: log_score(l,r) = log_score[l + r*fas]
                                  == float(logCharPr->g(l,r)) (db_matrix.h:76-83) */
    float log_gap_open;        /* Evol_model::log_gap_open()      = log_id_prob        */
    float log_gap_ext;         /* Evol_model::log_gap_ext()       = log_ext_prob       */
    float log_gap_end_ext;     /* Evol_model::log_gap_end_ext()   = log_end_ext_prob   */
    float log_gap_break_ext;   /* Evol_model::log_gap_break_ext() (pair-end reads: dead code, kept) */
    float log_non_gap;         /* Evol_model::log_non_gap()       = log_match_prob     */
} pg2_model_desc;

typedef struct pg2_job {
    pg2_graph left;            /* rows i   (Viterbi_alignment::left)  */
    pg2_graph right;           /* columns j (Viterbi_alignment::right) */
    int32_t model;             /* handle from pg2_model_upload */
    uint32_t flags;            /* PG2_FLAG_* */
    const int32_t *upper;      /* anchor band, [left.n_sites-1] inclusive lower j per row, or NULL */
    const int32_t *lower;      /* [left.n_sites-1] inclusive upper j per row, or NULL (both or none) */
} pg2_job;

/* ---- outputs ------------------------------------------------------------------------------- */

/* Result header of one job.  The traceback itself is returned compactly: packed back-pointers (the encoding below
 * needs 14 bits) in WALK order (end corner first: the end pointer, then the pointer of every visited cell), run-length
 * encoded in uint16 words at steps[step_off .. step_off+n_steps): a word with bit 15 clear is a pointer, a word with
 * bit 15 set repeats the previous pointer (word & 0x7fff) more times.  step_off is AUTHORITATIVE: the words of the jobs
 * of a call lie compacted in the caller's step buffer, but in an order the engine chooses (pg2_batch_fetch: job order;
 * pg2_align_batch on large batches: chunk by chunk, jobs grouped by left graph).  The buffer itself must still hold
 * left.n_sites + right.n_sites words per job (pg2_batch_step_capacity), the bound of an incompressible walk; on
 * PG2_ERR_CAPACITY nothing ran and that sum is what the caller must provide (pg2_batch_fetch also writes each job's
 * bound into results[].n_steps).  pg2_expand_path() turns the words into the reference's vector<Path_pointer>. */
typedef struct pg2_result {
    double score;              /* Viterbi log-score == max_end.score (viterbi_alignment.cpp:1558-1566) */
    int64_t cells;             /* in-band DP cells filled */
    int64_t step_off;          /* offset of this job's packed pointers in the caller's step buffer */
    int32_t n_steps;           /* number of encoded uint16 words of this job's walk */
    int32_t status;            /* PG2_JOB_* */
    uint32_t end_ptr;          /* packed end-corner pointer (same encoding as steps[]) */
    int32_t kernel;            /* which fill kernel ran: 0 = wavefront (general fallback), 1 = strip (warp per alignment),
                                  2 = lanes (lane per alignment, jobs sharing the left graph), 3 = pipelined strips (CTA per
                                  alignment, general graphs on both sides, anchor bands) */
} pg2_result;

/* Packed back-pointer: bits 0-1 source matrix (PG2_*_MAT, 3 = none), bits 2-7 ordinal of the LEFT
 * backward edge used (within the site's CSR slice), bits 8-13 ordinal of the RIGHT backward edge. */
#define PG2_PTR_MATRIX(p) ((int)((p) & 3u))
#define PG2_PTR_LEFT(p) ((int)(((p) >> 2) & 63u))
#define PG2_PTR_RIGHT(p) ((int)(((p) >> 8) & 63u))
#define PG2_PTR_NONE 3

/* One element of the reference's path (Path_pointer, basic_alignment.h:52-65), forward order. */
typedef struct pg2_step {
    double score;              /* Matrix_pointer::score carried by the element (-1 for skipped-site steps) */
    int32_t matrix;            /* PG2_*_MAT: column type */
    int32_t x_ind, y_ind;
    int32_t x_edge_ind, y_edge_ind;
    int32_t real_site;         /* 1 real step, 0 = pre-existing gap emitted by insert_preexisting_gap */
} pg2_step;

typedef struct pg2_ctx pg2_ctx;

/* ---- entry points --------------------------------------------------------------------------- */

int pg2_abi_version(void);

/* Text of the last error raised on this thread's most recent failing call (never NULL). */
const char *pg2_last_error(void);

/* Create / destroy an engine bound to CUDA device `device`.  Replaces nothing in the reference (it has
 * no device); one ctx stands where one OpenMP/boost worker thread stood (node.cpp:196-269). */
int pg2_ctx_create(int device, pg2_ctx **out);
void pg2_ctx_destroy(pg2_ctx *ctx);

/* Stage one Evol_model on the device.  Replaces the per-cell getter calls model->log_score(),
 * log_gap_open() ... (evol_model.h:69-83) inside the fill.  Cache handles by distance on the caller's
 * side (node.cpp:70-71 builds one model per node; placement uses a single one). */
int pg2_model_upload(pg2_ctx *ctx, const pg2_model_desc *desc, int32_t *handle_out);
int pg2_model_release(pg2_ctx *ctx, int32_t handle);

/* Align n_jobs independent graph pairs: fill + end corner + traceback.  Replaces
 * viterbi_alignment.cpp:238-296 and :379-383 for each job.  `steps` receives the packed pointers of all
 * jobs back to back (capacity step_cap uint16; a job needs at most left.n_sites + right.n_sites).
 * Returns PG2_OK when the batch ran; per-job outcomes are in results[i].status. */
int pg2_align_batch(pg2_ctx *ctx, int32_t n_jobs, const pg2_job *jobs, pg2_result *results,
                    uint16_t *steps, int64_t step_cap);

/* The same call split in three, for callers that keep a launch batch resident in HBM (and for
 * measurement: bench.py times pg2_batch_run alone for the device-resident figure):
 *   pg2_batch_create  validates O(1) properties, packs the jobs into pinned staging (graphs that share
 *                     host arrays are packed once);
 *   pg2_batch_run     uploads on first use, then validation + fill + traceback kernels; may be repeated;
 *   pg2_batch_fetch   copies results and packed pointers back (step_cap >= pg2_batch_step_capacity).
 * One batch per ctx at a time; the job arrays must stay valid until pg2_batch_create returns. */
typedef struct pg2_batch pg2_batch;
int pg2_batch_create(pg2_ctx *ctx, int32_t n_jobs, const pg2_job *jobs, pg2_batch **out);
int pg2_batch_run(pg2_ctx *ctx, pg2_batch *batch);
int pg2_batch_fetch(pg2_ctx *ctx, pg2_batch *batch, pg2_result *results, uint16_t *steps, int64_t step_cap);
int64_t pg2_batch_step_capacity(const pg2_batch *batch);
void pg2_batch_destroy(pg2_ctx *ctx, pg2_batch *batch);

/* Host-side unpacker: rebuilds the reference's forward path (backtrack_new_path,
 * viterbi_alignment.cpp:1038-1189, incl. the real_site=false steps of insert_preexisting_gap,
 * viterbi_alignment.h:146-193) from one job's packed pointers, replaying the score of every element in
 * the reference's operation order.  out_steps capacity: left.n_sites + right.n_sites.
 * used_left / used_right (optional, may be NULL): edge indices the reference marks is_used(true)
 * (viterbi_alignment.cpp:1054-1057,1079-1101,1128,1155), in marking order; capacities as out_steps.
 * Pure integer/FP64 host work on data the device produced; it runs no DP. */
int pg2_expand_path(const pg2_job *job, const pg2_model_desc *model, const pg2_result *result,
                    const uint16_t *steps, pg2_step *out_steps, int32_t *n_out,
                    int32_t *used_left, int32_t *n_used_left, int32_t *used_right, int32_t *n_used_right);

/* Engine statistics of the last pg2_align_batch on this ctx (for bench.py): device milliseconds of
 * the fill kernels and of the traceback kernel (CUDA events on the engine's stream), bytes copied. */
typedef struct pg2_stats {
    double fill_ms;
    double traceback_ms;
    double h2d_ms, d2h_ms;
    int64_t h2d_bytes, d2h_bytes;
    int64_t cells;
    int64_t traceback_bytes;   /* packed back-pointers written by the fill */
    int32_t fill_launches, traceback_launches;
    int32_t jobs_wavefront, jobs_strip;
    double run_ms;             /* validation + all fill + all traceback launches of the last pg2_batch_run */
    int32_t kernel_launches;   /* kernels launched by the last pg2_batch_run */
    int32_t jobs_strip_groups;
    int32_t jobs_lanes;        /* jobs the lane-per-alignment kernel took (shared row graph) */
    int32_t jobs_pstrip;       /* jobs the pipelined-strip kernel took (CTA per alignment: general x general, banded, small waves) */
    int32_t jobs_band;         /* jobs the band kernel took (warp per anchored alignment of two plain chains) */
    int32_t jobs_pstrip_ring;  /* ... of the pipelined-strip jobs, those that ran with the shared-memory row ring */
    int32_t jobs_lanes_wide;   /* ... of the lane-kernel jobs, those launched in the latency shape (one CTA of 10 warps per SM) */
} pg2_stats;
int pg2_get_stats(pg2_ctx *ctx, pg2_stats *out);

/* Device addresses of the last run's result records (24-byte records {double score; uint32 end_ptr;
 * int32 n_steps; int32 status; int32 raw_steps}, job order) and of the compacted path words (all jobs back to back in
 * job order, *n_steps_total words), for callers that forward results GPU-to-GPU (multi-GPU gather to rank 0 over NCCL,
 * no host bounce).  Waits for the run.  Valid until the next batch is created on this ctx. */
int pg2_batch_device_buffers(pg2_ctx *ctx, pg2_batch *batch, void **results_dev, void **steps_dev, int64_t *n_steps_total);
int pg2_stream_synchronize(pg2_ctx *ctx);

/* Measures the device's FP64 issue rate (the fill kernels' roofline denominator): 1e9 warp-instructions
 * per second chip-wide for independent DADDs, and for the DADD + DSETP + select group of one first-wins
 * candidate update.  Runs two small kernels for a few milliseconds. */
int pg2_measure_fp64_issue(int device, double *dadd_gips, double *cand_gips);

/* Issue-port check behind that roofline: cycles per warp and loop iteration on one SM sub-partition for
 * cycles[0] 8 independent DADDs, cycles[1] 16 independent FADDs, cycles[2] both together.  cycles[2] ~= 24 (the
 * instruction count) rather than cycles[0] + cycles[1] shows that FP64 instructions (2 pipe cycles each) overlap other
 * issue: a kernel with fewer than half FP64 instructions is bound by the issue port, one warp-instruction per cycle.
 * sm_clock_mhz: the clock the cycle counts were derived with. */
int pg2_measure_dispatch_mix(int device, double cycles[3], double *sm_clock_mhz);

/* ---- prefix anchors (host) ------------------------------------------------------------------ */

/* One anchor: seq1[start_1 .. start_1+length) == seq2[start_2 .. start_2+length) (reference: Substring_hit,
 * src/utils/substring_hit.h:32-46, with score == length and both strands plus). */
typedef struct pg2_anchor_hit {
    int32_t start_1, start_2, length;
} pg2_anchor_hit;

/* Replaces Find_anchors::find_long_substrings (src/utils/find_anchors.cpp:35-127; --use-prefix-anchors, called from
 * Viterbi_alignment::define_tunnel, src/main/viterbi_alignment.cpp:70-74): the exact substrings of at least min_length
 * characters that are neighbours in the common suffix order of the two sequences, longest first, greedily thinned so
 * that no two share a site.  Same hits in the same order as the reference (same suffix order, same sort, same greedy
 * walk); linear in the number of candidate hits where the reference's vector::erase loop is quadratic.  Host code (the
 * band the reference derives from the hits is what the device gets, pg2_job.upper / lower).  Returns PG2_ERR_CAPACITY
 * with *n_hits = the number of hits when `cap` is too small. */
int pg2_find_prefix_anchors(const char *seq1, int32_t len1, const char *seq2, int32_t len2, int32_t min_length,
                            pg2_anchor_hit *hits, int32_t cap, int32_t *n_hits);

/* Replaces Find_anchors::define_tunnel (src/utils/find_anchors.cpp:320-435): hits -> the anchor band.  str1 / str2 are the
 * sequence strings WITH the gap characters of skipped sites ('-', Sequence::get_sequence_string(true)); the hits count
 * characters of the gapless strings.  upper / lower receive len1 + 1 values each: row i may use columns upper[i] .. lower[i]
 * (what Viterbi_alignment keeps in upper_bound / lower_bound, viterbi_alignment.h:43-44, and passes on as pg2_job.upper /
 * lower).  `width` = --anchors-offset.  The same values as the reference, in linear time (the reference builds the lower
 * bounds by inserting at the front of a vector).  Host code. */
int pg2_anchor_band(const pg2_anchor_hit *hits, int32_t n_hits, const char *str1, int32_t len1, const char *str2, int32_t len2,
                    int32_t width, int32_t *upper, int32_t *lower);

#ifdef __cplusplus
}
#endif
#endif /* PAGAN2_B200_H */
