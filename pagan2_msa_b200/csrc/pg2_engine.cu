// pg2_engine.cu -- host side of the C-ABI (include/pagan2_b200.h): context, model staging, batch packing,
// launch grouping, result fetch.  No DP arithmetic happens on the host; without a CUDA device every
// computing entry point fails (PG2_ERR_NO_DEVICE) -- there is no CPU fallback.
//
// Launch batch = the unit the schedulers hand over (a guide-tree wave, node.cpp:240-264, or the trial /
// final alignments of many reads, reads_aligner.cpp:983-1216).  Jobs are packed into flat arrays
// (pg2_device.cuh), graphs shared between jobs (same host arrays) are uploaded once, and jobs are cut
// into groups that fit the scratch budget; each group is one fill launch + one traceback launch.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/pagan2_b200.h"
#include "pg2_device.cuh"
#include "pg2_strip_geom.cuh"
#include "pg2_pstrip_geom.cuh"

namespace pg2 {
int pstrip_warps(int n_blocks, int park_cap, bool smalltab, int rr, int K);
long long pstrip_cta_double4(int K, int nw, int max_lx, int ring, int max_slots);
void launch_pstrip_fill(int K, bool smalltab, int nw, int G, int n_jobs, int n_clusters, const DevJob *jobs, const int *job_ids, const DevGraph *graphs,
                        const DevModel *models, const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw,
                        const int4 *d_vrow, const int *d_vlast, const int *d_blo, const int *d_bhi, unsigned *ptrs, DevResult *results,
                        double4 *scratch, int max_lx, int ring, int max_slots, int park_cap, int rr, int *queue, cudaStream_t stream);
void launch_expand_implicit(int n_graphs, const DevGraph *graphs, int *d_off, int *d_estart, float *d_elogw, cudaStream_t stream);
void launch_validate(int n_graphs, int n_jobs, DevGraph *graphs, const DevJob *jobs, const DevModel *models, const int *d_state,
                     const int *d_off, const int *d_estart, const int *d_blo, const int *d_bhi, int *graph_status,
                     DevResult *results, bool few_long, cudaStream_t stream);
void launch_compact_steps(int n_jobs, const DevJob *jobs, const DevResult *results, long long *block_scratch, long long *total,
                          const unsigned short *steps_in, unsigned short *steps_out, cudaStream_t stream);
void launch_wavefront_fill(int n_jobs, int threads, const DevJob *jobs, const int *job_ids, const DevGraph *graphs,
                           const DevModel *models, const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw,
                           const int *d_blo, const int *d_bhi, const int *d_dlo, const long long *d_doff, const int *d_vlast,
                           double4 *scores, unsigned *ptrs, DevResult *results, int max_diag, cudaStream_t stream);
void launch_strip_fill(int K, bool general, bool smalltab, int n_jobs, const DevJob *jobs, const int *job_ids, const DevGraph *graphs, const DevModel *models,
                       const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw, const int4 *d_vrow,
                       unsigned short *ptrs, DevResult *results, double4 *saved_all, long long saved_per_warp, double4 *bcol_all,
                       long long bcol_per_warp, int *queue, int n_warps, cudaStream_t stream);
int strip_warps_per_sm();
void launch_lane_fill(int variant, int W, int n_tasks, const LaneTask *tasks, const DevJob *jobs, const DevGraph *graphs,
                      const DevModel *models, const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw,
                      const int4 *d_vrow, const int *d_vlast, unsigned short *ptrs, DevResult *results, double *scratch, int max_nv,
                      int max_lx, int max_slots, int *queue, int n_ctas, cudaStream_t stream);
int lane_ctas_per_sm(int W);
void launch_band_fill(bool smalltab, int n_jobs, const DevJob *jobs, const int *job_ids, const DevGraph *graphs, const DevModel *models,
                      const int *d_state, const int *d_band4, unsigned *ptrs, DevResult *results, cudaStream_t stream);
void launch_band_traceback(int n_jobs, int max_seg, const int *job_ids, const DevJob *jobs, const int *d_band4, const unsigned *ptrs,
                           DevResult *results, int4 *cand, int2 *act, unsigned short *steps, cudaStream_t stream);
bool strip_eligible(int lx, int ly, bool banded, int l_simple, int r_simple, int l_maxdeg, int r_maxdeg, int fas);
void launch_traceback(int n_jobs, int n_wave, int n_ps, const int *job_ids, const DevJob *jobs, const DevGraph *graphs, const int *d_vlast, const int *d_off,
                      const int *d_estart, const int *d_blo, const int *d_bhi, const int *d_dlo, const long long *d_doff,
                      const unsigned *ptr32, const unsigned short *ptr16, const unsigned *ptrps, unsigned short *steps, DevResult *results,
                      cudaStream_t stream);
}  // namespace pg2

using namespace pg2;

static thread_local std::string g_last_error = "";

static int fail(int code, const char *fmt, const char *a = "", const char *b = "") {
    char buf[512];
    snprintf(buf, sizeof buf, fmt, a, b);
    g_last_error = buf;
    return code;
}

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) return fail(PG2_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// grow-only device buffer
template <class T> struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return PG2_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 64;
        if (cudaMalloc((void **)&p, want * sizeof(T)) != cudaSuccess) {
            cudaGetLastError();
            if (cudaMalloc((void **)&p, n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); p = nullptr; return PG2_ERR_NOMEM; }
            want = n;
        }
        cap = want;
        return PG2_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// grow-only pinned host staging vector
template <class T> struct PinVec {
    T *p = nullptr;
    size_t n = 0, cap = 0;
    bool reserve(size_t want) {
        if (want <= cap) return true;
        size_t nc = std::max(want, cap * 2 + 1024);
        T *q = nullptr;
        if (cudaMallocHost((void **)&q, nc * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return false; }
        if (n) memcpy(q, p, n * sizeof(T));
        if (p) cudaFreeHost(p);
        p = q;
        cap = nc;
        return true;
    }
    T *extend(size_t k) {
        if (!reserve(n + k)) return nullptr;
        T *r = p + n;
        n += k;
        return r;
    }
    void clear() { n = 0; }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = cap = 0; }
};

struct ModelRec {
    bool live = false;
    int fas = 0;
    float *d_table = nullptr;
    DevModel dev;
};

// pipelined strips: rows of the shared-memory row ring a launch group runs with (bits 2-3 of strip_general; 0: parked rows)
static inline int ps_ring_rows(int strip_general) {
    const int tier = (strip_general >> 2) & 7;
    return tier == 0 ? 0 : (4 << tier);  // 8, 16, 32, 64, 128
}

struct Group {
    int kernel;        // 0 wavefront, 1 strip, 2 lanes
    int variant = 0;   // lane kernel: template variant shared by the group's tasks
    int task_first = 0, task_count = 0;  // lane kernel: range in batch->tasks
    int strip_k;       // strip kernel: columns per lane (all jobs of the group share it)
    int strip_general; // strip kernel: general row body needed
    int first, count;  // range in batch->order
    long long cells;   // pointer-buffer entries of the group
    int max_diag;
    int max_slots, max_lx;  // strip kernel per-warp scratch: saved rows, boundary column
    int max_nv = 1;         // lane kernel: longest row program
    int lane_w = LANE_W;    // lane kernel: warps per CTA (LANE_W: three CTAs per SM; LANE_W_WIDE: one, for launches that cannot fill the chip)
    int ps_ring = 1, ps_nw = 1;  // pipelined-strip kernel: boundary ring (virtual rows, power of two), warps per CTA
    int ps_park = 1, ps_blocks = 1;  // ... history slots per block, most blocks of a job
    int phase = 0;          // groups of one phase keep their pointer buffers side by side and share one traceback launch
    long long ptr_off = 0;  // offset of the group's region in its pointer buffer (d_ptr16 or d_ptr32 / d_scores)
};

struct pg2_batch {
    int n_jobs = 0, n_graphs = 0;
    std::vector<DevJob> jobs;
    std::vector<DevGraph> graphs;
    std::vector<int> order;  // job ids sorted by (kernel, -cells)
    std::vector<Group> groups;
    std::vector<LaneTask> tasks;  // lane kernel work items, group by group
    long long total_steps = 0;
    long long total_cells = 0;
    long long n_bcand = 0, n_bact = 0;  // band kernel: walk candidate records / segment records of the batch
    bool few_long = false;              // at most a few thousand graphs, some of them long: validation takes a CTA per graph
    long long h2d_bytes = 0;
    size_t n_off_total = 0, n_edge_total = 0;  // device sizes of d_off / d_estart (staged explicit graphs + implicit chains)
    bool uploaded = false, ran = false, fetch_enqueued = false;
};

constexpr int PIPE_SLOTS = 8;  // chunks of one pg2_align_batch call in flight at once (packing / H2D / kernels / D2H overlap): one slot per
                               // chunk of the default cut, so the host never waits for a traceback that sits behind the next chunks' fills

struct pg2_ctx {
    int device = 0;
    cudaStream_t stream = 0;
    cudaStream_t hi_stream = 0;  // high priority: the traceback of a pipelined chunk must not queue behind the resident
                                 // fill CTAs of the following chunks (they hold every SM's registers until their tail)
    cudaDeviceProp prop;
    std::vector<ModelRec> models;
    bool models_dirty = true;
    size_t scratch_bytes = (size_t)64 << 30;  // pointer/score scratch per launch group (PG2_SCRATCH_MB overrides)
    bool force_wavefront = false;  // PG2_FORCE_WAVEFRONT=1: route every job through the general kernel (tests)
    bool no_lanes = false;         // PG2_NO_LANES=1: keep shared-target jobs on the warp-per-alignment strip kernel (tests)
    bool no_pstrip = false;        // PG2_NO_PSTRIP=1: never use the pipelined-strip kernel (tests: the older kernels stay covered)
    bool pstrip_banded_chains = false;
    bool force_psring = false;     // PG2_FORCE_PSRING=1: the row-ring step for every eligible job, chains too (tests)
    bool no_psring = false;        // PG2_NO_PSRING=1: the pipelined strips keep parked rows in global memory (the older step body; tests)
    bool no_band = false;          // PG2_NO_BAND=1: banded chain x chain jobs stay on the wavefront kernel's chain path (tests)
    int pstrip_cluster_max = 8;     // CTAs (SMs) one pipelined-strip alignment may be spread over (PG2_PSTRIP_CLUSTER; 1 = one CTA per job)  // PG2_PSTRIP_BANDED_CHAINS=1: banded chain x chain jobs too (tests)
    int pstrip_max_jobs = 600;     // strip-eligible jobs of a batch go to the pipelined-strip kernel when there are at most this
                                   // many of them (a warp per alignment cannot fill the chip; PG2_PSTRIP_MAX_JOBS)
    size_t lane_scratch_bytes = (size_t)8 << 30;  // cap of the lane kernel's per-CTA wrap / end-column / parked-row scratch
    int lane_wide_hint = -1;  // pipelined pg2_align_batch: the shape of the whole call's lane launches (1 wide, 0 narrow), decided once
    // staging (pinned) and device arrays of the current batch
    PinVec<int> h_state, h_off, h_estart, h_blo, h_bhi, h_dlo, h_vrow, h_vlast;
    PinVec<float> h_elogw;
    PinVec<long long> h_doff;
    PinVec<int> h_band4;           // band kernel: per job diagonal geometry, row pointer offsets, walk segment table (pg2_band.cu)
    DevBuf<int> d_band4;
    DevBuf<int4> d_bcand;          // band kernel: walk candidates of every segment
    DevBuf<int2> d_bact;           // band kernel: the candidates the paths go through
    DevBuf<int> d_state, d_off, d_estart, d_blo, d_bhi, d_dlo, d_order, d_graph_status, d_vrow, d_vlast, d_queue;
    DevBuf<double4> d_saved, d_bcol;
    DevBuf<double> d_lane_scratch;
    DevBuf<LaneTask> d_tasks;
    DevBuf<float> d_elogw;
    DevBuf<long long> d_doff;
    DevBuf<DevJob> d_jobs;
    DevBuf<DevGraph> d_graphs;
    DevBuf<DevModel> d_models;
    DevBuf<DevResult> d_results;
    DevBuf<double4> d_scores;
    DevBuf<unsigned> d_ptr32;
    DevBuf<unsigned> d_ptrps;      // pipelined-strip kernel: one 32-bit pointer word per cell of its block layout
    DevBuf<double4> d_ps_scratch;  // pipelined-strip kernel: per CTA end columns, boundary rings, parked rows
    DevBuf<unsigned short> d_steps;          // per job a region of lx + ly words (what the walk may need at most)
    DevBuf<unsigned short> d_steps_compact;  // the run-length encoded words of all jobs back to back, job order
    DevBuf<long long> d_step_scan;           // block sums of the compaction scan; the last element is the total
    PinVec<long long> h_step_total;
    DevBuf<unsigned short> d_ptr16;
    PinVec<DevResult> h_results;
    PinVec<DevModel> h_models;  // staging of the model records (ctx-owned, so the upload needs no host wait)
    cudaEvent_t ev[8];
    pg2_stats stats;
    pg2_batch *current = nullptr;
    bool prio_set = false;  // sibling: its stream carries the priority of its pipeline slot
    pg2_ctx *sibling[PIPE_SLOTS - 1] = {};  // further sets of staging / device buffers + streams for pipelined pg2_align_batch calls
    bool borrowed_models = false;  // a sibling shares the primary's model tables and must not free them
};

extern "C" int pg2_abi_version(void) { return PG2_ABI_VERSION; }
extern "C" const char *pg2_last_error(void) { return g_last_error.c_str(); }

extern "C" int pg2_ctx_create(int device, pg2_ctx **out) {
    if (!out) return fail(PG2_ERR_INVALID, "pg2_ctx_create: null out pointer");
    *out = nullptr;
    // Pipelined pg2_align_batch calls keep up to PIPE_SLOTS chunks in flight on two streams each.  With the driver's default of 8
    // hardware work queues several of those streams share a queue, and an upload then waits behind another chunk's traceback that
    // is itself waiting for SM slots (measured: chunk 6 of 8 started 14 ms late).  Only effective before the process creates its
    // CUDA context; a caller that initialised CUDA earlier sets the variable itself (bench.py, engine.py do).
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(PG2_ERR_NO_DEVICE, "no CUDA device visible; this library has no CPU path");
    }
    if (device < 0 || device >= n) return fail(PG2_ERR_NO_DEVICE, "device index out of range");
    CU(cudaSetDevice(device));
    pg2_ctx *c = new pg2_ctx();
    c->device = device;
    if (cudaGetDeviceProperties(&c->prop, device) != cudaSuccess) { delete c; return fail(PG2_ERR_CUDA, "cudaGetDeviceProperties failed"); }
    if (c->prop.major < 10) {
        delete c;
        return fail(PG2_ERR_NO_DEVICE, "device is not sm_100 class; the kernels are built for sm_100a only");
    }
    {
        // the ctx stream sits one level below the traceback stream (slot 0 of a pipelined call: the oldest chunk, see ensure_siblings)
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, std::min(prio_hi + 1, prio_lo)) != cudaSuccess) {
            cudaGetLastError();
            delete c;
            return fail(PG2_ERR_CUDA, "stream creation failed");
        }
        if (cudaStreamCreateWithPriority(&c->hi_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { cudaGetLastError(); c->hi_stream = 0; }
    }
    for (int i = 0; i < 8; i++) cudaEventCreate(&c->ev[i]);
    const char *fw = getenv("PG2_FORCE_WAVEFRONT");
    c->force_wavefront = fw && atoi(fw) != 0;
    const char *nl = getenv("PG2_NO_LANES");
    c->no_lanes = nl && atoi(nl) != 0;
    const char *nps = getenv("PG2_NO_PSTRIP");
    c->no_pstrip = nps && atoi(nps) != 0;
    const char *pbc = getenv("PG2_PSTRIP_BANDED_CHAINS");
    c->pstrip_banded_chains = pbc && atoi(pbc) != 0;
    const char *npr = getenv("PG2_NO_PSRING");
    c->no_psring = npr && atoi(npr) != 0;
    const char *fpr = getenv("PG2_FORCE_PSRING");
    c->force_psring = fpr && atoi(fpr) != 0;
    const char *nb = getenv("PG2_NO_BAND");
    c->no_band = nb && atoi(nb) != 0;
    const char *pcl = getenv("PG2_PSTRIP_CLUSTER");
    if (pcl && atoi(pcl) >= 1 && atoi(pcl) <= 8) c->pstrip_cluster_max = atoi(pcl);
    const char *pmj = getenv("PG2_PSTRIP_MAX_JOBS");
    if (pmj && atoi(pmj) >= 0) c->pstrip_max_jobs = atoi(pmj);
    const char *mb = getenv("PG2_SCRATCH_MB");
    if (mb && atoll(mb) > 0) c->scratch_bytes = (size_t)atoll(mb) << 20;
    size_t cap = c->prop.totalGlobalMem / 100 * 45;  // leave room for inputs, steps and the caller
    if (c->scratch_bytes > cap) c->scratch_bytes = cap;
    memset(&c->stats, 0, sizeof c->stats);
    *out = c;
    return PG2_OK;
}

extern "C" void pg2_ctx_destroy(pg2_ctx *c) {
    if (!c) return;
    for (auto &sib : c->sibling) if (sib) { pg2_ctx_destroy(sib); sib = nullptr; }
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (!c->borrowed_models)
        for (auto &m : c->models) if (m.live && m.d_table) cudaFree(m.d_table);
    c->h_state.release(); c->h_off.release(); c->h_estart.release(); c->h_blo.release(); c->h_bhi.release(); c->h_dlo.release();
    c->h_band4.release(); c->d_band4.release(); c->d_bcand.release(); c->d_bact.release();
    c->h_elogw.release(); c->h_doff.release(); c->h_results.release(); c->h_models.release(); c->h_vrow.release(); c->h_vlast.release();
    c->d_vrow.release(); c->d_vlast.release(); c->d_queue.release(); c->d_saved.release(); c->d_bcol.release();
    c->d_state.release(); c->d_off.release(); c->d_estart.release(); c->d_blo.release(); c->d_bhi.release(); c->d_dlo.release();
    c->d_order.release(); c->d_graph_status.release(); c->d_elogw.release(); c->d_doff.release(); c->d_jobs.release();
    c->d_graphs.release(); c->d_models.release(); c->d_results.release(); c->d_scores.release(); c->d_ptr32.release();
    c->d_ptrps.release(); c->d_ps_scratch.release(); c->d_steps.release(); c->d_steps_compact.release(); c->d_step_scan.release(); c->h_step_total.release(); c->d_ptr16.release(); c->d_lane_scratch.release(); c->d_tasks.release();
    for (int i = 0; i < 8; i++) cudaEventDestroy(c->ev[i]);
    cudaStreamDestroy(c->stream);
    if (c->hi_stream) cudaStreamDestroy(c->hi_stream);
    delete c;
}

extern "C" int pg2_model_upload(pg2_ctx *c, const pg2_model_desc *d, int32_t *handle_out) {
    if (!c || !d || !handle_out || !d->log_score || d->fas <= 0) return fail(PG2_ERR_INVALID, "pg2_model_upload: bad argument");
    CU(cudaSetDevice(c->device));
    ModelRec r;
    r.live = true;
    r.fas = d->fas;
    size_t n = (size_t)d->fas * d->fas;
    if (cudaMalloc((void **)&r.d_table, n * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return fail(PG2_ERR_NOMEM, "model table allocation failed"); }
    CU(cudaMemcpyAsync(r.d_table, d->log_score, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    r.dev.table = r.d_table;
    r.dev.fas = d->fas;
    r.dev.open = d->log_gap_open;
    r.dev.ext = d->log_gap_ext;
    r.dev.end_ext = d->log_gap_end_ext;
    r.dev.brk = d->log_gap_break_ext;
    r.dev.lng = d->log_non_gap;
    int h = -1;
    for (size_t i = 0; i < c->models.size(); i++) if (!c->models[i].live) { h = (int)i; break; }
    if (h < 0) { h = (int)c->models.size(); c->models.push_back(r); } else c->models[h] = r;
    c->models_dirty = true;
    *handle_out = h;
    return PG2_OK;
}

extern "C" int pg2_model_release(pg2_ctx *c, int32_t h) {
    if (!c || h < 0 || h >= (int)c->models.size() || !c->models[h].live) return fail(PG2_ERR_INVALID, "pg2_model_release: bad handle");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->models[h].d_table);
    c->models[h] = ModelRec();
    c->models_dirty = true;
    return PG2_OK;
}

// ------------------------------------------------------------------------------------------------
// batch packing
// ------------------------------------------------------------------------------------------------
struct GraphKey {
    const void *a, *b, *c, *d;
    int n;
    bool operator==(const GraphKey &o) const { return a == o.a && b == o.b && c == o.c && d == o.d && n == o.n; }
};
// Identity table of the graphs of one batch: open addressing on the `state` pointer, sized once (a batch names at
// most 2 graphs per job), 16-byte slots so that the table of a 100 000-job batch stays cache resident; a hit is
// confirmed against the full key of the graph it names.
struct GraphTable {
    struct Slot { const void *a; long long gid; };
    std::vector<Slot> slots;
    size_t mask = 0;
    explicit GraphTable(size_t max_items) {
        size_t cap = 64;
        while (cap < max_items * 2 + 16) cap <<= 1;
        Slot empty = {nullptr, -1};
        slots.assign(cap, empty);
        mask = cap - 1;
    }
    // returns the slot of `k` (gid >= 0) or the empty slot where it belongs (gid < 0)
    Slot &find(const GraphKey &k, const std::vector<const pg2_graph *> &sources) {
        size_t i = ((((size_t)k.a >> 4) * 0x9E3779B97F4A7C15ull) >> 20) & mask;
        for (;;) {
            Slot &s = slots[i];
            if (s.gid < 0) return s;
            if (s.a == k.a) {
                const pg2_graph &g = *sources[(size_t)s.gid];
                if (g.bwd_off == k.b && g.edge_start == k.c && g.edge_logw == k.d && g.n_sites == k.n) return s;
            }
            i = (i + 1) & mask;
        }
    }
};

// Packing step 1 (serial): de-duplicate by host-array identity.  Touches the job records only, never the arrays.
static int intern_graph(pg2_batch *b, const pg2_graph &g, GraphTable &seen, std::vector<const pg2_graph *> &sources, int *gid) {
    const bool compact = !g.bwd_off && !g.edge_start && !g.edge_logw;  // plain chain described by its states only
    if (g.n_sites < 2 || g.n_edges < 0 || !g.state || (compact && g.n_edges != g.n_sites - 1) ||
        (!compact && (!g.bwd_off || (g.n_edges > 0 && (!g.edge_start || !g.edge_logw || !g.edge_index)))))
        return fail(PG2_ERR_INVALID, "graph with null arrays or fewer than 2 sites");
    GraphKey key = {g.state, g.bwd_off, g.edge_start, g.edge_logw, g.n_sites};
    GraphTable::Slot &slot = seen.find(key, sources);
    if (slot.gid >= 0) { *gid = slot.gid; return PG2_OK; }
    DevGraph dg;
    memset(&dg, 0, sizeof dg);
    dg.n_sites = g.n_sites;
    dg.vrow_base = -1;
    dg.np_base = -1;
    dg.n_vrows = g.n_sites - 1;
    dg.vlast_base = -1;
    dg.vplain_base = -1;
    dg.cp_ci_base = -1;
    *gid = (int)b->graphs.size();
    b->graphs.push_back(dg);
    sources.push_back(&g);
    slot.a = key.a;
    slot.gid = *gid;
    return PG2_OK;
}

// Packing step 2 (any thread): one pass over a distinct graph's arrays -- shape summary (which fill kernel may take
// it; the device re-validates everything) and whether the graph is an IMPLICIT chain: site s entered by the one edge
// (s-1 -> s) with log weight +0.0, i.e. a plain leaf or read.  Only the states of an implicit chain are staged and
// uploaded; its CSR is generated on the device (expand_implicit_kernel).  max_indeg = -1 flags malformed offsets,
// -2 an n_edges that does not match bwd_off[n_sites].
static void classify_graph(DevGraph &dg, const pg2_graph &g) {
    const int n_edges = g.n_edges;
    const int *po = g.bwd_off, *pe = g.edge_start;
    const float *pw = g.edge_logw;
    dg.max_span = 1;
    dg.n_extra = 0;
    if (!po) {  // compact form: the caller states that the graph is a plain unit-weight chain
        dg.max_indeg = 1;
        dg.simple = 1;
        dg.zero_w = 1;
        dg.implicit = 1;
        return;
    }
    if (po[g.n_sites] != n_edges) { dg.max_indeg = -2; return; }
    if (n_edges == g.n_sites - 1) {
        // the common case first, as three branch-free passes the compiler vectorises: a plain chain with unit weights has
        // bwd_off = {0, 0, 1, 2, ...}, edge_start = {0, 1, 2, ...} and log weights of +0.0 (all bits clear)
        int diff = po[0];
        for (int s = 1; s <= g.n_sites; s++) diff |= po[s] ^ (s - 1);
        for (int k = 0; k < n_edges; k++) diff |= pe[k] ^ k;
        const uint32_t *wb = reinterpret_cast<const uint32_t *>(pw);
        uint32_t wbits = 0;
        for (int k = 0; k < n_edges; k++) wbits |= wb[k];
        if (diff == 0 && wbits == 0) {
            dg.max_indeg = 1;
            dg.simple = 1;
            dg.zero_w = 1;
            dg.implicit = 1;
            return;
        }
    }
    int simple = 1, maxdeg = 0, maxspan = 1;
    for (int s = 0; s < g.n_sites; s++) {
        int k0 = po[s], k1 = po[s + 1];
        if (k1 < k0 || k1 > n_edges || k0 < 0) { simple = 0; maxdeg = -1; break; }  // malformed: the device flags it
        int deg = k1 - k0;
        if (deg > maxdeg) maxdeg = deg;
        if (s > 0 && (deg != 1 || pe[k0] != s - 1)) {
            simple = 0;
            if (s < g.n_sites - 1)  // (the stop site is no DP site: the end corner reads its sources from the end columns)
                for (int k = k0; k < k1; k++) maxspan = std::max(maxspan, s - pe[k]);
        }
        if (s == 0 && deg != 0) simple = 0;
    }
    dg.max_span = maxspan;
    dg.n_extra = std::max(0, n_edges - (g.n_sites - 1));
    dg.max_indeg = maxdeg;
    dg.simple = simple;
    dg.zero_w = 1;
    if (maxdeg >= 0)
        for (int k = 0; k < n_edges; k++) if (pw[k] != 0.0f || std::signbit(pw[k])) { dg.zero_w = 0; break; }
    dg.implicit = (simple && dg.zero_w && n_edges == g.n_sites - 1) ? 1 : 0;
}

// Packing step 4 (any thread): copy one graph into the staging arrays.
static void copy_graph(pg2_ctx *c, DevGraph &dg, const pg2_graph &g) {
    memcpy(c->h_state.p + dg.state_base, g.state, sizeof(int) * g.n_sites);
    if (dg.implicit) return;
    const int n_edges = g.n_edges;
    int *po = c->h_off.p + dg.off_base, *pe = c->h_estart.p + dg.edge_base;
    float *pw = c->h_elogw.p + dg.edge_base;
    if (dg.max_indeg < 0) {
        // malformed offsets: keep them in range so that no kernel reads out of bounds -- an edgeless graph, flagged bad
        for (int s = 0; s <= g.n_sites; s++) po[s] = 0;
        po[0] = 1;  // off[0] != 0 => validation marks JOB_BAD_GRAPH
        dg.max_indeg = 0;
    } else {
        memcpy(po, g.bwd_off, sizeof(int) * (g.n_sites + 1));
    }
    if (n_edges) {
        memcpy(pe, g.edge_start, sizeof(int) * n_edges);
        memcpy(pw, g.edge_logw, sizeof(float) * n_edges);
    }
}

// CSR of a staged graph as the host sees it: the staged arrays, or the implied chain of an implicit graph
struct HostCsr {
    const int *po, *pe;
    const float *pw;
    bool implicit;
    int off(int s) const { return implicit ? (s > 0 ? s - 1 : 0) : po[s]; }
    int start(int k) const { return implicit ? k : pe[k]; }
    float logw(int k) const { return implicit ? 0.0f : pw[k]; }
};
static HostCsr host_csr(pg2_ctx *c, const DevGraph &dg) {
    HostCsr h;
    h.implicit = dg.implicit != 0;
    h.po = h.implicit ? nullptr : c->h_off.p + dg.off_base;
    h.pe = h.implicit ? nullptr : c->h_estart.p + dg.edge_base;
    h.pw = h.implicit ? nullptr : c->h_elogw.p + dg.edge_base;
    return h;
}

// Row program of a graph used as the strip kernel's ROW graph (pg2_strip_geom.cuh): virtual rows, saved-row
// slots, and the site -> completing-virtual-row map the traceback needs.  Built once per distinct graph.
// Returns PG2_ERR_UNSUPPORTED when the graph needs more saved-row slots than the table can name.
static int build_row_program(pg2_ctx *c, DevGraph &dg) {
    if (dg.vrow_base >= 0) return PG2_OK;
    const int n = dg.n_sites, rows = n - 1;
    const int *ps = c->h_state.p + dg.state_base;
    const HostCsr csr = host_csr(c, dg);
    // Saved-row slots: a DP row p that is the source of an edge p -> s with s - p >= 2 (s a DP row too) must stay
    // addressable until row s is done.  A slot is reused two rows after its last reader (the skewed sweep
    // reads it one step late on the next lane).
    std::vector<int> last_use(n, -1), slot_of(n, -1);
    for (int s = 1; s < rows; s++)
        for (int k = csr.off(s); k < csr.off(s + 1); k++) {
            int p = csr.start(k);
            if (p >= 0 && p < s && s - p >= 2 && last_use[p] < s) last_use[p] = s;
        }
    std::vector<int> free_slots;
    std::vector<std::vector<int> > release(n + 3);
    int n_slots = 0;
    for (int s = 0; s < rows; s++) {
        for (size_t r = 0; r < release[s].size(); r++) free_slots.push_back(release[s][r]);
        if (last_use[s] > 0) {
            int slot;
            if (!free_slots.empty()) { slot = free_slots.back(); free_slots.pop_back(); }
            else slot = n_slots++;
            slot_of[s] = slot;
            // the warp-per-alignment kernel reads a parked row one step late on the next lane; a general column of the
            // pipelined-strip kernel reads the portion of a lane up to PS_HIST - 2 lanes ahead
            release[std::min(last_use[s] + PS_HIST, n + 2)].push_back(slot);
        }
    }
    if (n_slots > STRIP_MAX_SLOTS) return PG2_ERR_UNSUPPORTED;
    dg.n_slots = n_slots;
    // rows the end corner reads: predecessors of the stop site and the last DP row (Y close, :1468-1469)
    std::vector<char> endpred(n, 0);
    endpred[rows - 1 >= 0 ? rows - 1 : 0] = 1;
    for (int k = csr.off(n - 1); k < csr.off(n); k++)
        if (csr.start(k) >= 0 && csr.start(k) < n - 1) endpred[csr.start(k)] = 1;
    int nv = 0;
    for (int s = 0; s < rows; s++) nv += std::max(csr.off(s + 1) - csr.off(s), 1);
    dg.vrow_base = (int)(c->h_vrow.n / 4);
    dg.vlast_base = (int)c->h_vlast.n;
    dg.n_vrows = nv;
    int *vr = c->h_vrow.extend((size_t)nv * 4), *vl = c->h_vlast.extend(n);
    if (!vr || !vl) return PG2_ERR_NOMEM;
    int v = 0;
    for (int s = 0; s < rows; s++) {
        const int k0 = csr.off(s), k1 = csr.off(s + 1), deg = k1 - k0;
        const int st = ps[s] < 0 ? 0 : (ps[s] & VR_STATE_MASK);
        const int tail = (endpred[s] ? VR_ENDPRED : 0) | ((slot_of[s] + 1) << VR_SLOT_SHIFT);
        if (deg == 0) {
            // the start site (and any unreachable site): nothing to accumulate.  Site 0 also carries REG so
            // that it may take the in-place fast row: its sources are the -inf initial strip.
            int info = st | VR_FIRST | VR_LAST | VR_NOEDGE | VR_ZERO_W | (s == 0 ? VR_REG : 0) | tail;
            vr[4 * v] = info; vr[4 * v + 1] = -1; vr[4 * v + 2] = s; vr[4 * v + 3] = 0;
            v++;
        }
        for (int k = k0; k < k1; k++) {
            const int p = csr.start(k);
            const bool reg = (p == s - 1);
            const bool zw = csr.logw(k) == 0.0f && !std::signbit(csr.logw(k));
            int info = st | (k == k0 ? VR_FIRST : 0) | (k == k1 - 1 ? VR_LAST | tail : 0) | (reg ? VR_REG : 0) | (zw ? VR_ZERO_W : 0);
            int src = (reg || p < 0 || p >= n ? 0 : (slot_of[p] < 0 ? 0 : slot_of[p])) | ((k - k0) << 16);
            vr[4 * v] = info; vr[4 * v + 1] = k; vr[4 * v + 2] = s; vr[4 * v + 3] = src;
            v++;
        }
        vl[s] = v - 1;
    }
    vl[n - 1] = v - 1;
    // lane kernel: which virtual rows of each pipeline block are plain interior rows (one unit-weight edge from the
    // row above, not row 0, not read by the end corner, never parked) -- they run in the hot loop
    const int n_blocks = (nv + LANE_B - 1) / LANE_B;
    dg.vplain_base = (int)c->h_vlast.n;
    int *vp = c->h_vlast.extend(n_blocks);
    if (!vp) return PG2_ERR_NOMEM;
    vr = c->h_vrow.p + (size_t)dg.vrow_base * 4;  // extend() may have moved nothing here, but stay safe
    for (int b = 0; b < n_blocks; b++) {
        unsigned m = 0;
        for (int r = 0; r < LANE_B && b * LANE_B + r < nv; r++) {
            const int info = vr[4 * (b * LANE_B + r)];
            const int need = VR_FAST | VR_ZERO_W;
            const int none = VR_ENDPRED | VR_NOEDGE | (int)(~0u << VR_SLOT_SHIFT);
            if ((info & need) == need && !(info & none)) m |= 1u << r;
        }
        vp[b] = (int)m;
    }
    return PG2_OK;
}

// clipped band + anti-diagonal geometry of one banded job (host, O(lx+ly)).  want4: the job may go to the band kernel
// (plain unit-weight chains on both sides): its geometry record is written instead of the wavefront kernel's, unless the
// longest diagonal is beyond what the band kernel takes.  Sets J.kernel = 4 when it did.
static int pack_band(pg2_ctx *c, DevJob &J, const int32_t *upper, const int32_t *lower, bool want4) {
    const int lx = J.lx, ly = J.ly, nd = lx + ly - 1;
    J.band_base = (long long)c->h_blo.n;
    int *blo = c->h_blo.extend(lx), *bhi = c->h_bhi.extend(lx);
    if (!blo || !bhi) return fail(PG2_ERR_NOMEM, "pinned staging allocation failed");
    bool ok = true;
    for (int i = 0; i < lx; i++) {
        blo[i] = upper[i] > 0 ? upper[i] : 0;                  // tunnel_matrix.h:194
        bhi[i] = lower[i] < ly - 1 ? lower[i] : ly - 1;
        if (bhi[i] < blo[i]) ok = false;
        if (i && (blo[i] < blo[i - 1] || bhi[i] < bhi[i - 1])) ok = false;
    }
    if (blo[0] > 0) ok = false;
    if (want4 && ok) {
        // band kernel record: first / last row of every diagonal (two closing entries: "no row"), row pointer offsets,
        // candidate offsets of the walk segments
        const int n_seg = (nd + BAND_SEG - 1) / BAND_SEG;
        const size_t mark = c->h_band4.n;
        if (mark & 1) { if (!c->h_band4.extend(1)) return fail(PG2_ERR_NOMEM, "pinned staging allocation failed"); }  // geometry pairs are copied 8 bytes at a time
        J.b4_base = (long long)c->h_band4.n;
        int *geo = c->h_band4.extend((size_t)2 * (nd + 2) + lx + n_seg + 1);
        if (!geo) return fail(PG2_ERR_NOMEM, "pinned staging allocation failed");
        int *roff = geo + 2 * (nd + 2), *seg = roff + lx;
        long long cells = 0;
        int max_diag = 1, first = 0, last = -1;
        for (int s = 0; s < nd; s++) {
            while (last + 1 < lx && blo[last + 1] + (last + 1) <= s) ++last;
            while (first < lx && bhi[first] + first < s) ++first;
            geo[2 * s] = first;
            geo[2 * s + 1] = last;
            if (last >= first) { cells += last - first + 1; max_diag = std::max(max_diag, last - first + 1); }
        }
        for (int s = nd; s < nd + 2; s++) { geo[2 * s] = lx; geo[2 * s + 1] = lx - 1; }
        long long bytes = 0;
        for (int i = 0; i < lx; i++) { roff[i] = (int)(bytes - blo[i]); bytes += bhi[i] - blo[i] + 1; }
        if (max_diag <= BAND_MAX_DIAG && bytes < 0x7fff0000LL) {
            long long cand = 0;
            for (int k = 0; k < n_seg; k++) {
                seg[k] = (int)cand;
                const int top = (k + 1) * BAND_SEG - 1;
                if (k == n_seg - 1) cand += 1;  // the top segment is entered by the end pointer alone
                else cand += 3LL * (std::max(geo[2 * top + 1] - geo[2 * top] + 1, 0) + std::max(geo[2 * top - 1] - geo[2 * top - 2] + 1, 0));
            }
            seg[n_seg] = (int)cand;
            J.kernel = 4;
            J.n_seg = n_seg;
            J.cells = cells;
            J.max_diag = max_diag;
            J.ptr_cells = 3 * ((bytes + 3) / 4);  // three planes (X, Y, M) of one byte per cell
            J.diag_base = -1;
            return PG2_OK;
        }
        c->h_band4.n = mark;  // too wide for the band kernel: the wavefront kernel's record instead
    }
    J.diag_base = (long long)c->h_dlo.n;
    int *dlo = c->h_dlo.extend(nd + 1);
    long long *doff = c->h_doff.extend(nd + 1);
    if (!dlo || !doff) return fail(PG2_ERR_NOMEM, "pinned staging allocation failed");
    long long cells = 0;
    int max_diag = 1;
    if (ok) {
        // rows on diagonal s: blo[i]+i <= s <= bhi[i]+i, both strictly increasing in i
        int first = 0, last = -1;
        for (int s = 0; s < nd; s++) {
            while (last + 1 < lx && blo[last + 1] + (last + 1) <= s) ++last;
            while (first < lx && bhi[first] + first < s) ++first;
            dlo[s] = first;
            doff[s] = cells;
            if (last >= first) {
                cells += last - first + 1;
                max_diag = std::max(max_diag, last - first + 1);
            }
        }
        dlo[nd] = 0;
        doff[nd] = cells;
    } else {
        // invalid band: the validation kernel reports PG2_JOB_BAD_BAND from the same blo/bhi; the job is
        // skipped by the fill, so any in-range geometry will do
        for (int s = 0; s <= nd; s++) { dlo[s] = 0; doff[s] = s ? 1 : 0; }
        cells = 1;
    }
    J.cells = cells;
    J.max_diag = max_diag;
    return PG2_OK;
}

// Sites of a graph the wavefront kernel cannot take through its plain-cell body (pg2_wavefront.cu): the start site and
// every site whose backward edges are not exactly one edge from the site before it.  Sorted list + plain-site bitmap.
static int build_nonplain(pg2_ctx *c, DevGraph &dg) {
    if (dg.np_base >= 0) return PG2_OK;
    const int rows = dg.n_sites - 1;  // DP sites 0 .. n_sites-2
    const HostCsr csr = host_csr(c, dg);
    std::vector<int> np;
    std::vector<unsigned> mask((size_t)(rows + 31) / 32 + 1, 0u);
    np.push_back(0);
    for (int s = 1; s < rows; s++) {
        const int k0 = csr.off(s), k1 = csr.off(s + 1);
        if (k1 - k0 == 1 && csr.start(k0) == s - 1) mask[(size_t)s >> 5] |= 1u << (s & 31);
        else np.push_back(s);
    }
    dg.np_base = (int)c->h_vlast.n;
    dg.n_np = (int)np.size();
    int *dst = c->h_vlast.extend(np.size());
    if (!dst) return PG2_ERR_NOMEM;
    memcpy(dst, np.data(), np.size() * sizeof(int));
    dg.npmask_base = (int)c->h_vlast.n;
    int *md = c->h_vlast.extend(mask.size());
    if (!md) return PG2_ERR_NOMEM;
    memcpy(md, mask.data(), mask.size() * sizeof(unsigned));
    return PG2_OK;
}

// Column program of a graph used as the COLUMN graph of the pipelined-strip kernel (pg2_pstrip_geom.cuh): general /
// parked / end columns, blocks that start at cut points, history slots.  Built once per distinct graph.  Returns
// PG2_ERR_UNSUPPORTED (remembered in cp_k == 0) when the graph does not fit the kernel's limits: an edge into a general
// column spans more than (PS_HIST - 3) * 4 columns, no cut point within a block's width, more than PS_MAX_END end columns.
static int build_col_program(pg2_ctx *c, DevGraph &dg) {
    if (dg.cp_ci_base >= 0) return dg.cp_k > 0 ? PG2_OK : PG2_ERR_UNSUPPORTED;
    dg.cp_ci_base = 0;
    dg.cp_k = 0;
    const int n = dg.n_sites, cols = n - 1;
    const HostCsr csr = host_csr(c, dg);
    std::vector<char> general((size_t)cols, 0), parked((size_t)cols, 0);
    std::vector<int> forbid((size_t)cols + 2, 0);  // difference array over block boundaries: a block may not start at c when forbid sums > 0
    int maxspan = 1;
    for (int j = 1; j < cols; j++) {
        const int k0 = csr.off(j), k1 = csr.off(j + 1);
        if (k1 < k0) return PG2_ERR_UNSUPPORTED;  // malformed: the general kernel's validation reports it
        if (k1 - k0 == 1 && csr.start(k0) == j - 1) continue;
        general[(size_t)j] = 1;
        int minsrc = j;
        for (int k = k0; k < k1; k++) {
            const int p = csr.start(k);
            if (p < 0 || p >= j) return PG2_ERR_UNSUPPORTED;
            parked[(size_t)p] = 1;
            minsrc = std::min(minsrc, p);
        }
        if (k1 - k0 > PS_MAX_RIGHT_INDEG) return PG2_ERR_UNSUPPORTED;
        if (minsrc < j) {  // every source of a general column lies in the column's own block
            maxspan = std::max(maxspan, j - minsrc);
            forbid[(size_t)minsrc + 1]++;
            forbid[(size_t)j + 1]--;
        }
    }
    int K = maxspan <= (PS_HIST - 3) * 2 ? 2 : (maxspan <= (PS_HIST - 3) * 4 ? 4 : 0);
    // One column per lane for heavily general column graphs (454 / homopolymer read graphs: a third of the columns have two or
    // three backward edges).  A step costs the warp the SUM over a lane's columns of the longest edge loop among the lanes, so
    // half the columns per lane is half the step; the blocks are 32 columns wide, twice as many warps work on the alignment.
    if (K == 2 && maxspan <= PS_HIST - 3 && (long long)dg.n_extra * 5 >= dg.n_sites && !c->no_psring) K = 1;
    if (const char *fk = getenv("PG2_PSTRIP_K")) {
        if (atoi(fk) == 4 && K <= 2) K = 4;
        if (atoi(fk) == 2 && K == 1) K = 2;
        if (atoi(fk) == 1 && K == 2 && maxspan <= PS_HIST - 3) K = 1;
    }
    if (K == 0) return PG2_ERR_UNSUPPORTED;
    // columns the end corner reads: the predecessors of the stop site and the last DP column
    std::vector<int> endcols(1, cols - 1);
    for (int k = csr.off(n - 1); k < csr.off(n); k++) {
        const int p = csr.start(k);
        if (p < 0 || p >= cols) return PG2_ERR_UNSUPPORTED;
        if (std::find(endcols.begin(), endcols.end(), p) == endcols.end()) endcols.push_back(p);
    }
    if ((int)endcols.size() > PS_MAX_END) return PG2_ERR_UNSUPPORTED;
    // blocks: greedy, as wide as the lanes, the parked-column budget and the cut points allow
    // (narrow strips first: fewer cells per step; the wider strips when no cut point lies within 64 columns somewhere)
    std::vector<int> blocks;
    {
        std::vector<char> cut((size_t)cols + 1, 0);
        int run = 0;
        for (int cc = 0; cc <= cols; cc++) { run += forbid[(size_t)cc]; cut[(size_t)cc] = run == 0; }
        for (; K <= 4; K *= 2) {
            blocks.clear();
            int c0 = 0;
            while (c0 < cols) {
                const int limit = std::min(c0 + 32 * K, cols);
                int count = 0, best = -1;
                for (int cc = c0; cc < limit; cc++) {
                    if (parked[(size_t)cc] && ++count > PS_MAX_PARK) break;
                    if (cc + 1 == cols || cut[(size_t)cc + 1]) best = cc + 1;
                }
                if (best < 0) break;
                blocks.push_back(c0);
                blocks.push_back(best);
                c0 = best;
            }
            if (c0 >= cols) break;
        }
        if (K > 4) return PG2_ERR_UNSUPPORTED;
    }
    const int nb = (int)blocks.size() / 2;
    if (nb >= (1 << (31 - PC_BLOCK_SHIFT))) return PG2_ERR_UNSUPPORTED;
    const int n_edges = csr.off(n);
    const int ci_base = (int)c->h_vlast.n;
    int *ci = c->h_vlast.extend((size_t)cols);
    if (!ci) return PG2_ERR_NOMEM;
    int max_park = 1;
    for (int bi = 0; bi < nb; bi++) {
        int slot = 0;
        for (int j = blocks[(size_t)2 * bi]; j < blocks[(size_t)2 * bi + 1]; j++) {
            int w = (general[(size_t)j] ? PC_GENERAL : 0) | (bi << PC_BLOCK_SHIFT);
            if (parked[(size_t)j]) w |= PC_PARKED | (slot++ << PC_SLOT_SHIFT);
            ci[j] = w;
        }
        max_park = std::max(max_park, slot);
    }
    for (size_t e = 0; e < endcols.size(); e++) ci[endcols[e]] |= PC_ENDCOL | ((int)e << PC_END_SHIFT);
    int ei_base = ci_base;
    if (!dg.implicit) {  // per edge: the history slot of its source column (read for edges into general columns only)
        ei_base = (int)c->h_vlast.n;
        int *ei = c->h_vlast.extend((size_t)n_edges + 1);
        if (!ei) return PG2_ERR_NOMEM;
        ci = c->h_vlast.p + ci_base;
        for (int j = 0; j < n; j++)
            for (int k = csr.off(j); k < csr.off(j + 1); k++) {
                const int p = csr.start(k);
                ei[k] = (p >= 0 && p < cols) ? ((ci[p] >> PC_SLOT_SHIFT) & PC_SLOT_MASK) : 0;
            }
    }
    const int blk_base = (int)c->h_vlast.n;
    int *bl = c->h_vlast.extend(blocks.size());
    if (!bl) return PG2_ERR_NOMEM;
    memcpy(bl, blocks.data(), blocks.size() * sizeof(int));
    dg.cp_ci_base = ci_base;
    dg.cp_ei_base = ei_base;
    dg.cp_blk_base = blk_base;
    dg.cp_n_blocks = nb;
    dg.cp_k = K;
    dg.cp_park = max_park;
    return PG2_OK;
}

// Routes one job to the pipelined-strip kernel: row program of the left graph, column program of the right graph, the
// job's block table (column range, virtual-row range inside the band, pointer offsets).  Returns PG2_ERR_UNSUPPORTED when
// the job stays where it was.
static int try_pstrip(pg2_ctx *c, pg2_batch *b, DevJob &J) {
    DevGraph &GL = b->graphs[J.left];
    DevGraph &GR = b->graphs[J.right];
    if (GL.max_indeg > PG2_MAX_IN_DEGREE || GR.max_indeg > PS_MAX_RIGHT_INDEG || GL.max_indeg < 0 || GR.max_indeg < 0) return PG2_ERR_UNSUPPORTED;
    if (c->models[J.model].fas > VR_STATE_MASK || J.lx < 1 || J.ly < 1) return PG2_ERR_UNSUPPORTED;
    int rc = build_row_program(c, GL);
    if (rc != PG2_OK) return rc;
    rc = build_col_program(c, GR);
    if (rc != PG2_OK) return rc;
    const int K = GR.cp_k, nb = GR.cp_n_blocks;
    const int blk_base = (int)c->h_vlast.n;
    int *out = c->h_vlast.extend((size_t)nb * PB_INTS);
    if (!out) return PG2_ERR_NOMEM;
    const int *cb = c->h_vlast.p + GR.cp_blk_base;
    const int *vl = c->h_vlast.p + GL.vlast_base;
    const int *blo = J.banded ? c->h_blo.p + J.band_base : nullptr, *bhi = J.banded ? c->h_bhi.p + J.band_base : nullptr;
    long long words = 0;
    int tallest = 1, ifirst = 0, ilast = -1;
    for (int bi = 0; bi < nb; bi++) {
        const int c0 = cb[2 * bi], c1 = cb[2 * bi + 1];
        int v0 = 0, v1 = GL.n_vrows, i0 = 0;
        if (J.banded) {
            // rows that meet the block: bhi[i] >= c0 - 1 (column c0 - 1 counts: a long-span edge may read it) and blo[i] <= c1 - 1;
            // both bounds are non-decreasing in i
            while (ifirst < J.lx && bhi[ifirst] < c0 - 1) ifirst++;
            while (ilast + 1 < J.lx && blo[ilast + 1] <= c1 - 1) ilast++;
            i0 = ifirst;
            v0 = ifirst == 0 ? 0 : vl[ifirst - 1] + 1;
            v1 = ilast >= ifirst ? vl[ilast] + 1 : v0;
        }
        if (words > 0x7fffffffLL - ps_block_words(v0, v1, K)) { c->h_vlast.n = (size_t)blk_base; return PG2_ERR_UNSUPPORTED; }
        int *e = out + (size_t)bi * PB_INTS;
        e[0] = c0; e[1] = c1; e[2] = v0; e[3] = v1; e[4] = i0; e[5] = (int)words;
        words += ps_block_words(v0, v1, K);
        tallest = std::max(tallest, v1 - v0);
    }
    J.kernel = 3;
    J.strip_k = K;
    J.strip_general = c->models[J.model].fas <= STRIP_SMALL_FAS ? 2 : 0;
    J.n_blocks = nb;
    J.blk_base = blk_base;
    J.ps_ring = tallest;
    J.ptr_cells = words;
    return PG2_OK;
}

// PG2_TIMING=1: wall time of the host packing steps on stderr (tuning aid)
struct PackTimer {
    bool on;
    std::chrono::steady_clock::time_point t;
    PackTimer() : on(getenv("PG2_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
    void lap(const char *what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "pg2 pack: %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

extern "C" int pg2_batch_create(pg2_ctx *c, int32_t n_jobs, const pg2_job *jobs, pg2_batch **out) {
    if (!c || !out || n_jobs < 0 || (n_jobs > 0 && !jobs)) return fail(PG2_ERR_INVALID, "pg2_batch_create: bad argument");
    PackTimer timer;
    *out = nullptr;
    CU(cudaSetDevice(c->device));
    if (c->current) return fail(PG2_ERR_INVALID, "pg2_batch_create: another batch is live on this ctx (destroy it first)");
    pg2_batch *b = new pg2_batch();
    b->n_jobs = n_jobs;
    b->jobs.resize(n_jobs);
    c->h_state.clear(); c->h_off.clear(); c->h_estart.clear(); c->h_elogw.clear(); c->h_vrow.clear(); c->h_vlast.clear();
    c->h_blo.clear(); c->h_bhi.clear(); c->h_dlo.clear(); c->h_doff.clear(); c->h_band4.clear();
    GraphTable seen((size_t)n_jobs * 2);
    b->graphs.reserve((size_t)n_jobs + 16);
    std::vector<const pg2_graph *> sources;
    sources.reserve((size_t)n_jobs + 16);
    // ---- step 1 (serial): argument checks, graph de-duplication; reads the job records only ----
    for (int t = 0; t < n_jobs; t++) {
        const pg2_job &j = jobs[t];
        DevJob &J = b->jobs[t];
        memset(&J, 0, sizeof J);
        if (j.model < 0 || j.model >= (int)c->models.size() || !c->models[j.model].live) { delete b; return fail(PG2_ERR_INVALID, "job names an unknown model handle"); }
        if ((j.upper == nullptr) != (j.lower == nullptr)) { delete b; return fail(PG2_ERR_INVALID, "job band needs both upper and lower (or neither)"); }
        int gl = 0, gr = 0;
        int rc = intern_graph(b, j.left, seen, sources, &gl);
        if (rc == PG2_OK) rc = intern_graph(b, j.right, seen, sources, &gr);
        if (rc != PG2_OK) { delete b; return rc; }
        J.left = gl;
        J.right = gr;
    }
    timer.lap("1 dedup");
    const int ng = (int)b->graphs.size();
    int nthreads = 1;
    {
        size_t sites = 0;
        for (int gi = 0; gi < ng; gi++) sites += (size_t)b->graphs[gi].n_sites;
        if (sites * 16 > ((size_t)8 << 20) && ng >= 64) {
            unsigned hw = std::thread::hardware_concurrency();
            nthreads = (int)std::min<unsigned>(hw ? hw : 4, 4);  // measured: 2-4 threads beat 16 (spawn cost, memory-bound passes)
            const char *pt = getenv("PG2_PACK_THREADS");
            if (pt && atoi(pt) > 0) nthreads = atoi(pt);
        }
    }
    auto parallel_over_graphs = [&](const std::function<void(int)> &fn) {
        if (nthreads <= 1) { for (int gi = 0; gi < ng; gi++) fn(gi); return; }
        std::vector<std::thread> pool;
        for (int w = 0; w < nthreads; w++) {
            int lo = (int)((long long)ng * w / nthreads), hi = (int)((long long)ng * (w + 1) / nthreads);
            pool.emplace_back([&fn, lo, hi]() { for (int gi = lo; gi < hi; gi++) fn(gi); });
        }
        for (auto &th : pool) th.join();
    };
    // ---- step 2 (parallel): one pass over every distinct graph: shape, implicit chains ----
    parallel_over_graphs([&](int gi) {
        if (gi + 2 < ng) {
            // the arrays of a read are small separate allocations: start the next graphs' cache misses early
            const pg2_graph &nx = *sources[gi + 2];
            const int bytes = nx.n_sites * 4;
            for (int o = 0; nx.bwd_off && o < bytes; o += 256) {
                __builtin_prefetch(reinterpret_cast<const char *>(nx.bwd_off) + o);
                __builtin_prefetch(reinterpret_cast<const char *>(nx.edge_start) + o);
                __builtin_prefetch(reinterpret_cast<const char *>(nx.edge_logw) + o);
                __builtin_prefetch(reinterpret_cast<const char *>(nx.state) + o);
            }
        }
        classify_graph(b->graphs[gi], *sources[gi]);
    });
    timer.lap("2 classify graphs");
    // ---- step 3 (serial): bases.  Explicit graphs come first in d_off / d_estart / d_elogw (staged and uploaded),
    //      implicit chains behind them (generated on the device) ----
    {
        long long n_state = 0, e_off = 0, e_edge = 0, i_off = 0, i_edge = 0;
        for (int gi = 0; gi < ng; gi++) {
            DevGraph &dg = b->graphs[gi];
            if (dg.max_indeg == -2) { delete b; return fail(PG2_ERR_INVALID, "graph: n_edges does not match bwd_off[n_sites]"); }
            const pg2_graph &g = *sources[gi];
            dg.state_base = (int)n_state;
            n_state += g.n_sites;
            if (dg.implicit) { dg.off_base = (int)i_off; dg.edge_base = (int)i_edge; i_off += g.n_sites + 1; i_edge += g.n_edges; }
            else { dg.off_base = (int)e_off; dg.edge_base = (int)e_edge; e_off += g.n_sites + 1; e_edge += g.n_edges; }
        }
        if (n_state > 0x7fffffffLL || e_off + i_off > 0x7fffffffLL || e_edge + i_edge > 0x7fffffffLL) {
            delete b;
            return fail(PG2_ERR_INVALID, "batch too large: more than 2^31 sites or edges; split the batch");
        }
        for (int gi = 0; gi < ng; gi++) {
            DevGraph &dg = b->graphs[gi];
            if (dg.implicit) { dg.off_base += (int)e_off; dg.edge_base += (int)e_edge; }
        }
        b->n_off_total = (size_t)(e_off + i_off);
        b->n_edge_total = (size_t)(e_edge + i_edge);
        auto grow = [](auto &v, size_t k) { return k == 0 || v.extend(k) != nullptr; };  // extend(0) of an empty vector is null
        if (!grow(c->h_state, (size_t)n_state) || !grow(c->h_off, (size_t)e_off) || !grow(c->h_estart, (size_t)e_edge) ||
            !grow(c->h_elogw, (size_t)e_edge)) {
            delete b;
            return fail(PG2_ERR_NOMEM, "pinned staging allocation failed");
        }
    }
    // ---- step 4 (parallel): copy the distinct graphs into pinned staging ----
    parallel_over_graphs([&](int gi) { copy_graph(c, b->graphs[gi], *sources[gi]); });
    timer.lap("3-4 bases, copy graphs");
    // ---- phase 3 (serial): bands, kernel choice, row programs ----
    long long step_base = 0;
    int pick_ly = -1, pick_k = 0;  // strip_pick_k of the last read length seen (reads of a batch mostly share it)
    for (int t = 0; t < n_jobs; t++) {
        const pg2_job &j = jobs[t];
        DevJob &J = b->jobs[t];
        int rc;
        J.lx = j.left.n_sites - 1;
        J.ly = j.right.n_sites - 1;
        J.model = j.model;
        J.flags = j.flags;
        J.banded = j.upper != nullptr;
        J.band_base = J.diag_base = -1;
        DevGraph &GL = b->graphs[J.left];
        const DevGraph &GR = b->graphs[J.right];
        if (J.banded) {
            // anchored alignments of two plain unit-weight chains (leaf x leaf): the band kernel, one warp per job
            const bool want4 = !c->force_wavefront && !c->no_band && !c->pstrip_banded_chains && GL.simple && GR.simple && GL.zero_w && GR.zero_w;
            rc = pack_band(c, J, j.upper, j.lower, want4);
            if (rc != PG2_OK) { delete b; return rc; }
        } else {
            J.cells = (long long)J.lx * J.ly;
        }
        if (J.kernel != 4) J.kernel = strip_eligible(J.lx, J.ly, J.banded != 0, GL.simple, GR.simple, GL.max_indeg, GR.max_indeg, c->models[j.model].fas) ? 1 : 0;
        if (c->force_wavefront) J.kernel = 0;
        if (J.kernel == 1) {
            rc = build_row_program(c, GL);
            if (rc == PG2_ERR_UNSUPPORTED) J.kernel = 0;  // too many parked rows: the general kernel takes it
            else if (rc != PG2_OK) { delete b; return fail(rc, "pinned staging allocation failed"); }
        }
        if (J.kernel == 1 && J.ly != pick_ly) { pick_ly = J.ly; pick_k = strip_pick_k(J.ly); }
        J.strip_k = J.kernel == 1 ? pick_k : 0;
        J.strip_general = (J.kernel == 1 && !(GL.simple && GL.zero_w)) ? 1 : 0;
        if (J.kernel == 1 && c->models[j.model].fas <= STRIP_SMALL_FAS) J.strip_general |= 2;  // bit 1: shared-table variant
        if (J.kernel == 4) J.strip_general = c->models[j.model].fas <= STRIP_SMALL_FAS ? 2 : 0;  // bit 1: shared-table variant
        else J.ptr_cells = J.kernel == 1 ? strip_cells(GL.n_vrows, J.ly, J.strip_k) : J.cells;
        if (J.kernel == 4) {
            const int *seg = c->h_band4.p + J.b4_base + 2LL * (J.lx + J.ly - 1 + 2) + J.lx;
            J.cand_base = b->n_bcand;
            J.act_base = b->n_bact;
            b->n_bcand += seg[J.n_seg];
            b->n_bact += J.n_seg;
        }
        J.step_base = step_base;
        J.step_cap = j.left.n_sites + j.right.n_sites;
        step_base += J.step_cap;
        b->total_cells += J.cells;
    }
    b->total_steps = step_base;
    b->n_graphs = (int)b->graphs.size();
    if (b->graphs.size() <= 4096 && n_jobs <= 4096) {
        int longest = 0;
        for (const DevGraph &g : b->graphs) longest = std::max(longest, g.n_sites);
        b->few_long = longest >= 2048;
    }

    timer.lap("5 bands, kernels, row programs");
    // ---- lane kernel tasks: strip-eligible jobs that share the row graph, model and flags, 32 per warp ----
    std::vector<int> lane_order;  // job ids, task by task
    if (!c->no_lanes && !c->force_wavefront) {
        // buckets by (left graph, model, flags); a left graph almost always comes with one (model, flags) pair, so the
        // lookup is a short list per graph instead of a hash table
        struct BucketRef { int model; unsigned flags; int bucket; };
        std::vector<std::vector<BucketRef> > by_graph(b->graphs.size());
        std::vector<std::vector<int> > buckets;
        for (int t = 0; t < n_jobs; t++) {
            const DevJob &J = b->jobs[t];
            if (J.kernel != 1) continue;
            const DevGraph &GL = b->graphs[J.left];
            if ((size_t)lane_cta_doubles(GL.n_vrows, J.lx, GL.n_slots, LANE_W) * sizeof(double) > c->lane_scratch_bytes / 16) continue;
            std::vector<BucketRef> &refs = by_graph[J.left];
            int bi = -1;
            for (const BucketRef &r : refs) if (r.model == J.model && r.flags == (J.flags & 3u)) { bi = r.bucket; break; }
            if (bi < 0) {
                bi = (int)buckets.size();
                refs.push_back({J.model, J.flags & 3u, bi});
                buckets.emplace_back();
            }
            buckets[bi].push_back(t);
        }
        // one launch instead of two when a batch mixes plain-chain and general row graphs: the general variant runs
        // plain rows through the same hot loop, and one task queue has one tail
        bool lane_merge = false;
        {
            bool any_plain = false, any_general = false;
            for (auto &bk : buckets) {
                if ((int)bk.size() < LANE_MIN_JOBS) continue;
                const DevGraph &GL = b->graphs[b->jobs[bk[0]].left];
                ((GL.simple && GL.zero_w) ? any_plain : any_general) = true;
            }
            const char *lm = getenv("PG2_LANE_MERGE");
            lane_merge = any_plain && any_general && !(lm && atoi(lm) == 0);
        }
        for (auto &bk : buckets) {
            if ((int)bk.size() < LANE_MIN_JOBS) continue;
            // lanes of one task sweep max_ly columns: put reads of similar length together
            auto longer = [&](int x, int y) { return b->jobs[x].ly > b->jobs[y].ly; };
            if (!std::is_sorted(bk.begin(), bk.end(), longer)) std::stable_sort(bk.begin(), bk.end(), longer);
            for (size_t pos = 0; pos < bk.size(); pos += 32) {
                const int n = (int)std::min<size_t>(32, bk.size() - pos);
                const DevJob &J0 = b->jobs[bk[pos]];
                const DevGraph &GL = b->graphs[J0.left];
                LaneTask T;
                memset(&T, 0, sizeof T);
                T.left = J0.left;
                T.model = J0.model;
                T.flags = J0.flags;
                T.n_jobs = n;
                T.max_ly = J0.ly;
                T.variant = ((GL.simple && GL.zero_w && !lane_merge) ? 0 : 1) | (c->models[J0.model].fas <= STRIP_SMALL_FAS ? 2 : 0);
                for (int l = 0; l < n; l++) {
                    DevJob &J = b->jobs[bk[pos + l]];
                    T.job_ids[l] = bk[pos + l];
                    if (!b->graphs[J.right].zero_w) T.variant |= 4;
                    J.kernel = 2;
                    J.strip_k = LANE_K;
                    J.lane = l;
                    J.task = (int)b->tasks.size();
                    J.ptr_cells = 0;
                }
                b->tasks.push_back(T);
            }
        }
        // tasks of one variant form a launch; costly tasks first (tail balance)
        std::vector<int> torder(b->tasks.size());
        for (size_t k = 0; k < torder.size(); k++) torder[k] = (int)k;
        auto cost = [&](int k) { const LaneTask &T = b->tasks[k]; return lane_cells(b->graphs[T.left].n_vrows, T.max_ly, LANE_K); };
        std::stable_sort(torder.begin(), torder.end(), [&](int x, int y) {
            if (b->tasks[x].variant != b->tasks[y].variant) return b->tasks[x].variant < b->tasks[y].variant;
            return cost(x) > cost(y);
        });
        std::vector<LaneTask> sorted;
        sorted.reserve(torder.size());
        for (int k : torder) sorted.push_back(b->tasks[k]);
        b->tasks.swap(sorted);
        size_t tpos = 0;
        while (tpos < b->tasks.size()) {
            Group g;
            g.kernel = 2;
            g.variant = b->tasks[tpos].variant;
            g.strip_k = LANE_K;
            g.strip_general = 0;
            g.first = (int)lane_order.size();
            g.count = 0;
            g.cells = 0;
            g.max_diag = 1;
            g.max_slots = 0;
            g.max_lx = 1;
            g.task_first = (int)tpos;
            g.task_count = 0;
            while (tpos < b->tasks.size() && b->tasks[tpos].variant == g.variant) {
                LaneTask &T = b->tasks[tpos];
                const DevGraph &GL = b->graphs[T.left];
                const long long need = lane_cells(GL.n_vrows, T.max_ly, LANE_K);
                if (g.task_count > 0 && (size_t)(g.cells + need) * 2 > c->scratch_bytes) break;
                T.ptr_base = g.cells;
                g.cells += need;
                g.max_slots = std::max(g.max_slots, GL.n_slots);
                g.max_lx = std::max(g.max_lx, GL.n_sites - 1);
                g.max_nv = std::max(g.max_nv, GL.n_vrows);
                for (int l = 0; l < T.n_jobs; l++) {
                    DevJob &J = b->jobs[T.job_ids[l]];
                    J.task = (int)tpos;
                    J.cell_base = T.ptr_base;
                    lane_order.push_back(T.job_ids[l]);
                }
                g.count += T.n_jobs;
                g.task_count++;
                tpos++;
            }
            b->groups.push_back(g);
        }
    }

    // ---- pipelined strips: every job the register-strip kernels cannot take (banded, general right graph), and the
    //      strip-eligible jobs of a batch too small to fill the chip with one warp per alignment ----
    if (!c->force_wavefront && !c->no_pstrip) {
        int n_strip = 0;
        for (int t = 0; t < n_jobs; t++) n_strip += b->jobs[t].kernel == 1;
        const bool small = n_strip <= c->pstrip_max_jobs;
        for (int t = 0; t < n_jobs; t++) {
            DevJob &J = b->jobs[t];
            // plain chains inside a band (anchored leaf x leaf) stay on the wavefront kernel: its chain path keeps a whole
            // anti-diagonal in flight per step, the strips advance one row of the band per step (measured: 152 vs 337 ms
            // per 200 kb pair)
            if (J.kernel == 0 && J.banded && b->graphs[J.left].simple && b->graphs[J.right].simple && !c->pstrip_banded_chains) continue;
            if (J.kernel == 0 || (J.kernel == 1 && small)) {
                const int rc = try_pstrip(c, b, J);
                if (rc == PG2_ERR_NOMEM) { delete b; return fail(rc, "pinned staging allocation failed"); }
            }
        }
    }
    // The row-ring step of the pipelined strips (pg2_pstrip.cu: ps_step_ring: the block's last 32, 64 or 128 rows in shared
    // memory), decided for the whole launch batch so that its K = 2 jobs stay ONE launch group (a wave of leaf pairs and ancestor
    // pairs in two launches runs one after the other).  It pays where the graphs are general enough (measured: C1 ancestors with
    // 10 % extra edges 7.1 -> 5.2 ms, C4 root 5.6 -> 5.0, the pileup step 15.4 -> 8.2); on chains and near-chains (leaf pairs,
    // ancestors with 0.2-2 % extra edges) the register-resident row of ps_step is 10-25 % faster.  The ring must outlast the
    // longest left edge plus the lanes the longest right edge reaches back (a lane is one virtual row ahead of the next).  A
    // 64-row ring leaves room for two warps per CTA, a 128-row ring for one: with a cluster of 8 CTAs that is 16 / 8 column
    // blocks in flight, and a job with more blocks than that is better off with ps_step's four warps per CTA.
    if (!c->no_psring) {
        for (int K = 1; K <= 2; K++) {  // (jobs of different strip widths are different launch groups anyway)
            bool general = c->force_psring;
            int need = 0, blocks = 0;
            for (int t = 0; t < n_jobs; t++) {
                const DevJob &J = b->jobs[t];
                if (J.kernel != 3 || J.strip_k != K) continue;
                const DevGraph &GL = b->graphs[J.left], &GR = b->graphs[J.right];
                if ((long long)GL.n_extra * 25 >= GL.n_sites || (long long)GR.n_extra * 25 >= GR.n_sites) general = true;  // 4 % extra edges
                need = std::max(need, GL.max_span + (GR.max_span + K - 1) / K + 4);
                blocks = std::max(blocks, J.n_blocks);
            }
            int tier = 0;
            if (general && need > 0)
                for (int q = 3; q <= 5; q++)
                    if (need <= (4 << q)) {
                        const size_t ring_bytes = (size_t)((4 << q) + 1) * 3 * K * 33 * sizeof(double);
                        const int warps = (int)std::min<size_t>(4, ((size_t)220 * 1024) / ring_bytes);  // per CTA; 8 CTAs per cluster
                        if (warps >= 4 || c->force_psring || blocks <= 8 * warps) tier = q;
                        break;
                    }
            if (tier)
                for (int t = 0; t < n_jobs; t++) {
                    DevJob &J = b->jobs[t];
                    if (J.kernel == 3 && J.strip_k == K) J.strip_general |= tier << 2;
                }
        }
    }
    for (int t = 0; t < n_jobs; t++) {
        DevJob &J = b->jobs[t];
        if (J.kernel != 0) continue;
        int rc = build_nonplain(c, b->graphs[J.left]);
        if (rc == PG2_OK) rc = build_nonplain(c, b->graphs[J.right]);
        if (rc != PG2_OK) { delete b; return fail(rc, "pinned staging allocation failed"); }
    }
    timer.lap("6 lane tasks, pipelined strips");
    // order: lane jobs (task by task), then strip jobs, then wavefront jobs; larger jobs first inside a class
    // (tail balance)
    b->order = lane_order;
    for (int t = 0; t < n_jobs; t++) if (b->jobs[t].kernel != 2) b->order.push_back(t);
    std::stable_sort(b->order.begin() + lane_order.size(), b->order.end(), [&](int x, int y) {
        const DevJob &A = b->jobs[x], &B = b->jobs[y];
        if (A.kernel != B.kernel) return A.kernel > B.kernel;
        if (A.strip_k != B.strip_k) return A.strip_k < B.strip_k;
        if (A.strip_general != B.strip_general) return A.strip_general < B.strip_general;
        return A.cells > B.cells;
    });
    // groups under the scratch budget: wavefront 36 B/cell (scores + pointer word), strip 2 B/cell
    size_t pos = lane_order.size();
    while (pos < b->order.size()) {
        Group g;
        g.kernel = b->jobs[b->order[pos]].kernel;
        g.strip_k = b->jobs[b->order[pos]].strip_k;
        g.strip_general = b->jobs[b->order[pos]].strip_general;
        g.first = (int)pos;
        g.count = 0;
        g.cells = 0;
        g.max_diag = 1;
        g.max_slots = 0;
        g.max_lx = 1;
        size_t per_cell = g.kernel == 0 ? 36 : (g.kernel >= 3 ? 4 : 2);
        while (pos < b->order.size()) {
            DevJob &J = b->jobs[b->order[pos]];
            if (J.kernel != g.kernel || J.strip_k != g.strip_k || J.strip_general != g.strip_general) break;
            long long padded = (J.ptr_cells + 7) & ~7LL;
            if (g.count > 0 && (size_t)(g.cells + padded) * per_cell > c->scratch_bytes) break;
            J.cell_base = g.cells;
            g.cells += padded;
            g.max_diag = std::max(g.max_diag, J.banded ? J.max_diag : std::min(J.lx, J.ly));
            g.max_slots = std::max(g.max_slots, b->graphs[J.left].n_slots);
            g.max_lx = std::max(g.max_lx, J.lx);
            if (g.kernel == 3) {
                while (g.ps_ring < J.ps_ring) g.ps_ring <<= 1;
                g.ps_park = std::max(g.ps_park, b->graphs[J.right].cp_park);
                g.ps_blocks = std::max(g.ps_blocks, J.n_blocks);
                g.ps_nw = pstrip_warps(g.ps_blocks, g.ps_park, (g.strip_general & 2) != 0, ps_ring_rows(g.strip_general), g.strip_k);
            }
            g.count++;
            pos++;
        }
        b->groups.push_back(g);
    }
    // phases: consecutive groups whose pointer (and score) buffers fit the scratch budget together run their fills
    // back to back and share ONE traceback launch (the walk is latency bound: one launch for many jobs)
    {
        int phase = 0;
        size_t bytes = 0;
        long long off16 = 0, off32 = 0, offps = 0;
        for (auto &g : b->groups) {
            const size_t need = (size_t)g.cells * (g.kernel == 0 ? 36 : (g.kernel >= 3 ? 4 : 2));
            if (bytes > 0 && bytes + need > c->scratch_bytes) { phase++; bytes = 0; off16 = off32 = offps = 0; }
            g.phase = phase;
            long long &off = g.kernel == 0 ? off32 : (g.kernel >= 3 ? offps : off16);  // the band kernel's bytes share the pipelined strips' word buffer
            g.ptr_off = off;
            off += g.cells;
            bytes += need;
            if (g.kernel == 2) {
                for (int t = g.task_first; t < g.task_first + g.task_count; t++) {
                    LaneTask &T = b->tasks[t];
                    T.ptr_base += g.ptr_off;
                    for (int l = 0; l < T.n_jobs; l++) b->jobs[T.job_ids[l]].cell_base = T.ptr_base;
                }
            } else {
                for (int k = g.first; k < g.first + g.count; k++) b->jobs[b->order[k]].cell_base += g.ptr_off;
            }
        }
    }
    timer.lap("7 order, groups, phases");
    c->current = b;
    *out = b;
    return PG2_OK;
}

static int upload_batch(pg2_ctx *c, pg2_batch *b) {
    int rc;
#define ENS(buf, n) if ((rc = (buf).ensure(n)) != PG2_OK) return fail(rc, "device allocation failed (%s)", #buf)
    ENS(c->d_vrow, c->h_vrow.n + 4); ENS(c->d_vlast, c->h_vlast.n + 1); ENS(c->d_queue, 4);
    ENS(c->d_state, c->h_state.n + 1); ENS(c->d_off, b->n_off_total + 1); ENS(c->d_estart, b->n_edge_total + 1); ENS(c->d_elogw, b->n_edge_total + 1);
    ENS(c->d_blo, c->h_blo.n + 1); ENS(c->d_bhi, c->h_bhi.n + 1); ENS(c->d_dlo, c->h_dlo.n + 1); ENS(c->d_doff, c->h_doff.n + 1);
    ENS(c->d_band4, c->h_band4.n + 2); ENS(c->d_bcand, (size_t)b->n_bcand + 1); ENS(c->d_bact, (size_t)b->n_bact + 1);
    ENS(c->d_jobs, b->jobs.size() + 1); ENS(c->d_graphs, b->graphs.size() + 1); ENS(c->d_order, b->order.size() + 1);
    ENS(c->d_graph_status, b->graphs.size() + 1); ENS(c->d_results, b->jobs.size() + 1); ENS(c->d_steps, (size_t)b->total_steps + 1);
    ENS(c->d_steps_compact, (size_t)b->total_steps + 1); ENS(c->d_step_scan, b->jobs.size() / 256 + 4);
    ENS(c->d_models, c->models.size() + 1); ENS(c->d_tasks, b->tasks.size() + 1);
    long long bytes = 0;
#define H2D(dst, src, n, T)                                                                                 \
    if ((n) > 0) { CU(cudaMemcpyAsync((dst).p, (src), (size_t)(n) * sizeof(T), cudaMemcpyHostToDevice, c->stream)); bytes += (long long)(n) * sizeof(T); }
    H2D(c->d_state, c->h_state.p, c->h_state.n, int);
    H2D(c->d_off, c->h_off.p, c->h_off.n, int);
    H2D(c->d_vrow, c->h_vrow.p, c->h_vrow.n, int);
    H2D(c->d_vlast, c->h_vlast.p, c->h_vlast.n, int);
    H2D(c->d_estart, c->h_estart.p, c->h_estart.n, int);
    H2D(c->d_elogw, c->h_elogw.p, c->h_elogw.n, float);
    H2D(c->d_blo, c->h_blo.p, c->h_blo.n, int);
    H2D(c->d_bhi, c->h_bhi.p, c->h_bhi.n, int);
    H2D(c->d_dlo, c->h_dlo.p, c->h_dlo.n, int);
    H2D(c->d_doff, c->h_doff.p, c->h_doff.n, long long);
    H2D(c->d_band4, c->h_band4.p, c->h_band4.n, int);
    H2D(c->d_jobs, b->jobs.data(), b->jobs.size(), DevJob);
    H2D(c->d_graphs, b->graphs.data(), b->graphs.size(), DevGraph);
    H2D(c->d_order, b->order.data(), b->order.size(), int);
    H2D(c->d_tasks, b->tasks.data(), b->tasks.size(), LaneTask);
    if (!c->models.empty()) {
        // staged in pinned memory the ctx owns: the copy is asynchronous and the host goes on packing the next chunk
        c->h_models.clear();
        DevModel *dm = c->h_models.extend(c->models.size());
        if (!dm) return fail(PG2_ERR_NOMEM, "pinned staging allocation failed");
        for (size_t i = 0; i < c->models.size(); i++) dm[i] = c->models[i].dev;
        CU(cudaMemcpyAsync(c->d_models.p, dm, c->models.size() * sizeof(DevModel), cudaMemcpyHostToDevice, c->stream));
        c->models_dirty = false;
    }
    // the CSR of implicit chains (plain leaves and reads) is generated where it is used
    launch_expand_implicit((int)b->graphs.size(), c->d_graphs.p, c->d_off.p, c->d_estart.p, c->d_elogw.p, c->stream);
    b->h2d_bytes = bytes;
    b->uploaded = true;
    return PG2_OK;
#undef ENS
#undef H2D
}

// Upload (if needed) and enqueue validation + every group's fill and traceback on the ctx stream.
// async: enqueue only -- no host round trips, no per-phase timings; batch_wait() completes the run
static int batch_run_impl(pg2_ctx *c, pg2_batch *b, bool async) {
    if (!c || !b || c->current != b) return fail(PG2_ERR_INVALID, "pg2_batch_run: bad batch");
    CU(cudaSetDevice(c->device));
    int rc;
    pg2_stats &st = c->stats;
    if (!b->uploaded) {
        CU(cudaEventRecord(c->ev[0], c->stream));
        if ((rc = upload_batch(c, b)) != PG2_OK) return rc;
        CU(cudaEventRecord(c->ev[1], c->stream));
        st.h2d_ms = 0;
        if (!async) {
            CU(cudaEventSynchronize(c->ev[1]));
            float ms = 0;
            cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]);
            st.h2d_ms = ms;
        }
        st.h2d_bytes = b->h2d_bytes;
    }
    // scratch for the largest phase of each buffer class (strip and lane groups share d_ptr16)
    long long max_w = 0, max_s = 0, max_ps = 0;
    for (auto &g : b->groups) {
        long long &m = g.kernel == 0 ? max_w : (g.kernel >= 3 ? max_ps : max_s);
        m = std::max(m, g.ptr_off + g.cells);
    }
    if (max_ps > 0 && (rc = c->d_ptrps.ensure((size_t)max_ps)) != PG2_OK) return fail(rc, "pointer buffer allocation failed");
    if (max_w > 0) {
        if ((rc = c->d_scores.ensure((size_t)max_w)) != PG2_OK) return fail(rc, "score scratch allocation failed");
        if ((rc = c->d_ptr32.ensure((size_t)max_w)) != PG2_OK) return fail(rc, "pointer buffer allocation failed");
    }
    if (max_s > 0 && (rc = c->d_ptr16.ensure((size_t)max_s)) != PG2_OK) return fail(rc, "pointer buffer allocation failed");
#ifdef PG2_HOST_EMU
    const int resident_warps = 1;
#else
    const int resident_warps = c->prop.multiProcessorCount * strip_warps_per_sm();
#endif
    auto lane_ctas = [&](const Group &g) {
        const size_t lane_resident = (size_t)std::max(c->prop.multiProcessorCount, 1) * lane_ctas_per_sm(g.lane_w);
        const size_t per_cta = (size_t)lane_cta_doubles(g.max_nv, g.max_lx, g.max_slots, g.lane_w) * sizeof(double);
        size_t fit = std::max<size_t>(c->lane_scratch_bytes / per_cta, 1);
        return (int)std::min<size_t>(std::min<size_t>(fit, lane_resident), (size_t)std::max(g.task_count, 1));
    };
    // A lane launch that cannot keep the chip's CTA slots busy for long -- a shard of a strong-scaling run, the trial alignments of
    // a few reads -- waits for its longest task (the root-most target of a placement tree has twice the sites of a leaf and a third
    // of them with several edges: 22 ms against 7 ms for the others).  Such launches take the wide shape: one CTA of LANE_W_WIDE
    // warps per SM, a task's strips in two rounds instead of five.  PG2_LANE_W = 4 / 10 forces a shape, PG2_LANE_WIDE_TASKS sets
    // the limit (tasks per SM below which the wide shape is used).
    {
        const char *fw = getenv("PG2_LANE_W"), *ft = getenv("PG2_LANE_WIDE_TASKS");
        const int forced = fw ? atoi(fw) : 0;
        const int per_sm = (ft && atoi(ft) >= 0) ? atoi(ft) : LANE_WIDE_TASKS_PER_SM;
        for (auto &g : b->groups) {
            if (g.kernel != 2) continue;
            int max_strips = 0;
            for (int t = 0; t < g.task_count; t++) max_strips = std::max(max_strips, (b->tasks[(size_t)g.task_first + t].max_ly + LANE_K - 1) / LANE_K);
            // (the chunks of a pipelined call share the chip: the call decides for all of them, by the tasks it holds in total)
            bool wide = (c->lane_wide_hint >= 0 ? c->lane_wide_hint == 1 : g.task_count <= per_sm * std::max(c->prop.multiProcessorCount, 1)) &&
                        max_strips > LANE_W;
            if (forced == LANE_W) wide = false;
            if (forced == LANE_W_WIDE) wide = true;
            if (wide && (size_t)lane_cta_doubles(g.max_nv, g.max_lx, g.max_slots, LANE_W_WIDE) * sizeof(double) > c->lane_scratch_bytes) wide = false;
            g.lane_w = wide ? LANE_W_WIDE : LANE_W;
        }
    }
    for (auto &g : b->groups)
        if (g.kernel == 2) {
            size_t need = (size_t)lane_cta_doubles(g.max_nv, g.max_lx, g.max_slots, g.lane_w) * lane_ctas(g);
            if ((rc = c->d_lane_scratch.ensure(need)) != PG2_OK) return fail(rc, "lane scratch allocation failed");
        }
    // pipelined strips: one CLUSTER per job in flight.  A launch with few jobs (a guide-tree wave, a pileup step) spreads every
    // job over G SMs with 4 warps per CTA -- one warp per SM sub-partition, all column blocks of a 1.5 k-column alignment in
    // flight at once; a launch that fills the chip by itself keeps one CTA of up to 12 warps per job.
    struct PsShape { int G, nw, clusters; };
    auto ps_shape = [&](const Group &g) {
        PsShape s;
        const int sms = c->prop.multiProcessorCount;
        s.G = 1;
        s.nw = g.ps_nw;
        const int per_job = sms / std::max(g.count, 1);  // SMs one job may take
        if (per_job >= 2 && g.ps_blocks > 4 && c->pstrip_cluster_max > 1) {
            const int nw = std::min(4, g.ps_nw);
            const int want = (g.ps_blocks + nw - 1) / nw;
            const int G = std::min(std::min(c->pstrip_cluster_max, want), per_job);
            if (G > 1) { s.G = G; s.nw = nw; }
        }
        if (s.G > 1) s.clusters = std::max(1, std::min(g.count, sms / s.G));
        else {
            const int per_sm = s.nw > 8 ? 1 : (s.nw > 4 ? 2 : 4);
            s.clusters = std::max(1, std::min(g.count, sms * per_sm));
        }
        return s;
    };
    for (auto &g : b->groups)
        if (g.kernel == 3) {
            const PsShape sh = ps_shape(g);
            const size_t need = (size_t)pstrip_cta_double4(g.strip_k, sh.nw * sh.G, g.max_lx, g.ps_ring, g.max_slots) * (size_t)sh.clusters;
            if ((rc = c->d_ps_scratch.ensure(need)) != PG2_OK) return fail(rc, "pipelined-strip scratch allocation failed");
        }
    for (auto &g : b->groups)
        if (g.kernel == 1) {
            int warps = (std::min(resident_warps, std::max(g.count, 1)) + 3) & ~3;  // whole CTAs of 4 warps: every launched warp owns scratch
            size_t saved = (size_t)std::max(g.max_slots, 1) * 32 * g.strip_k * warps;
            size_t bcol = (size_t)g.max_lx * 2 * warps;
            if ((rc = c->d_saved.ensure(saved)) != PG2_OK) return fail(rc, "saved-row scratch allocation failed");
            if ((rc = c->d_bcol.ensure(bcol)) != PG2_OK) return fail(rc, "boundary-column scratch allocation failed");
        }

    CU(cudaEventRecord(c->ev[7], c->stream));
    launch_validate(b->n_graphs, b->n_jobs, c->d_graphs.p, c->d_jobs.p, c->d_models.p, c->d_state.p, c->d_off.p, c->d_estart.p,
                    c->d_blo.p, c->d_bhi.p, c->d_graph_status.p, c->d_results.p, b->few_long, c->stream);
    st.fill_ms = st.traceback_ms = 0;
    st.fill_launches = st.traceback_launches = 0;
    st.jobs_wavefront = st.jobs_strip = st.jobs_lanes = st.jobs_pstrip = st.jobs_band = st.jobs_pstrip_ring = st.jobs_lanes_wide = 0;
    st.jobs_strip_groups = 0;
    st.cells = b->total_cells;
    st.traceback_bytes = 0;
    // one phase = fills of its groups back to back, then one traceback launch over all its jobs; the events are
    // read after the phase (host round trip per phase, bounded by the scratch budget)
    for (size_t gi = 0; gi < b->groups.size();) {
        size_t ge = gi;
        while (ge < b->groups.size() && b->groups[ge].phase == b->groups[gi].phase) ge++;
        CU(cudaEventRecord(c->ev[2], c->stream));
        int phase_jobs = 0, phase_wave_jobs = 0, phase_ps_jobs = 0;
        for (size_t k = gi; k < ge; k++) {
            const Group &g = b->groups[k];
            const int *ids = c->d_order.p + g.first;
            phase_jobs += g.count;
            if (g.kernel == 0) phase_wave_jobs += g.count;
            if (g.kernel == 3) {
                phase_ps_jobs += g.count;
                const PsShape sh = ps_shape(g);
                launch_pstrip_fill(g.strip_k, (g.strip_general & 2) != 0, sh.nw, sh.G, g.count, sh.clusters, c->d_jobs.p, ids, c->d_graphs.p,
                                   c->d_models.p, c->d_state.p, c->d_off.p, c->d_estart.p, c->d_elogw.p,
                                   reinterpret_cast<const int4 *>(c->d_vrow.p), c->d_vlast.p, c->d_blo.p, c->d_bhi.p, c->d_ptrps.p,
                                   c->d_results.p, c->d_ps_scratch.p, g.max_lx, g.ps_ring, g.max_slots, g.ps_park, ps_ring_rows(g.strip_general),
                                   c->d_queue.p, c->stream);
                st.jobs_pstrip += g.count;
                if (g.strip_k <= 2 && ps_ring_rows(g.strip_general) > 0) st.jobs_pstrip_ring += g.count;
                st.traceback_bytes += g.cells * 4;
            } else if (g.kernel == 4) {
                launch_band_fill((g.strip_general & 2) != 0, g.count, c->d_jobs.p, ids, c->d_graphs.p, c->d_models.p, c->d_state.p, c->d_band4.p,
                                 c->d_ptrps.p, c->d_results.p, c->stream);
                st.jobs_band += g.count;
                st.traceback_bytes += g.cells * 4;
            } else if (g.kernel == 0) {
                int threads = g.max_diag <= 32 ? 32 : g.max_diag <= 64 ? 64 : g.max_diag <= 128 ? 128 : g.max_diag <= 256 ? 256
                              : g.max_diag <= 512 ? 512 : 1024;
                launch_wavefront_fill(g.count, threads, c->d_jobs.p, ids, c->d_graphs.p, c->d_models.p, c->d_state.p, c->d_off.p,
                                      c->d_estart.p, c->d_elogw.p, c->d_blo.p, c->d_bhi.p, c->d_dlo.p, c->d_doff.p, c->d_vlast.p,
                                      c->d_scores.p, c->d_ptr32.p, c->d_results.p, g.max_diag, c->stream);
                st.jobs_wavefront += g.count;
                st.traceback_bytes += g.cells * 4;
            } else if (g.kernel == 2) {
                launch_lane_fill(g.variant, g.lane_w, g.task_count, c->d_tasks.p + g.task_first, c->d_jobs.p, c->d_graphs.p, c->d_models.p,
                                 c->d_state.p, c->d_off.p, c->d_estart.p, c->d_elogw.p, reinterpret_cast<const int4 *>(c->d_vrow.p),
                                 c->d_vlast.p, c->d_ptr16.p, c->d_results.p, c->d_lane_scratch.p, g.max_nv, g.max_lx, g.max_slots, c->d_queue.p,
                                 lane_ctas(g), c->stream);
                st.jobs_lanes += g.count;
                if (g.lane_w == LANE_W_WIDE) st.jobs_lanes_wide += g.count;
                st.jobs_strip_groups++;
                st.traceback_bytes += g.cells * 2;
            } else {
                int warps = (std::min(resident_warps, std::max(g.count, 1)) + 3) & ~3;  // whole CTAs of 4 warps: every launched warp owns scratch
                launch_strip_fill(g.strip_k, (g.strip_general & 1) != 0, (g.strip_general & 2) != 0, g.count, c->d_jobs.p, ids, c->d_graphs.p,
                                  c->d_models.p, c->d_state.p, c->d_off.p, c->d_estart.p, c->d_elogw.p,
                                  reinterpret_cast<const int4 *>(c->d_vrow.p), c->d_ptr16.p, c->d_results.p, c->d_saved.p,
                                  (long long)std::max(g.max_slots, 1) * 32 * g.strip_k, c->d_bcol.p, (long long)g.max_lx, c->d_queue.p,
                                  warps, c->stream);
                st.jobs_strip += g.count;
                st.jobs_strip_groups++;
                st.traceback_bytes += g.cells * 2;
            }
            st.fill_launches++;
        }
        CU(cudaEventRecord(c->ev[3], c->stream));
        // pipelined chunks: the walk goes to the high-priority stream (after the fills, before anything later on the
        // ctx stream), so that its CTAs take the first SM slots that free up
        cudaStream_t tb_stream = (async && c->hi_stream) ? c->hi_stream : c->stream;
        if (tb_stream != c->stream) CU(cudaStreamWaitEvent(tb_stream, c->ev[3], 0));
        launch_traceback(phase_jobs, phase_wave_jobs, phase_ps_jobs, c->d_order.p + b->groups[gi].first, c->d_jobs.p, c->d_graphs.p, c->d_vlast.p,
                         c->d_off.p, c->d_estart.p, c->d_blo.p, c->d_bhi.p, c->d_dlo.p, c->d_doff.p, c->d_ptr32.p, c->d_ptr16.p, c->d_ptrps.p,
                         c->d_steps.p, c->d_results.p, tb_stream);
        for (size_t k = gi; k < ge; k++) {
            const Group &g = b->groups[k];
            if (g.kernel != 4) continue;
            int max_seg = 0;
            for (int q = g.first; q < g.first + g.count; q++) max_seg = std::max(max_seg, b->jobs[b->order[q]].n_seg);
            launch_band_traceback(g.count, max_seg, c->d_order.p + g.first, c->d_jobs.p, c->d_band4.p, c->d_ptrps.p, c->d_results.p,
                                  c->d_bcand.p, c->d_bact.p, c->d_steps.p, tb_stream);
        }
        CU(cudaEventRecord(c->ev[4], tb_stream));
        if (tb_stream != c->stream) CU(cudaStreamWaitEvent(c->stream, c->ev[4], 0));
        if (!async) {
            CU(cudaEventSynchronize(c->ev[4]));
            CU(cudaGetLastError());
            float f = 0, t = 0;
            cudaEventElapsedTime(&f, c->ev[2], c->ev[3]);
            cudaEventElapsedTime(&t, c->ev[3], c->ev[4]);
            st.fill_ms += f;
            st.traceback_ms += t;
        }
        st.traceback_launches++;
        gi = ge;
    }
    // the run-length encoded paths of all jobs back to back (job order): what goes back to the host or to rank 0
    {
        const size_t nb = b->jobs.size() / 256 + 2;
        launch_compact_steps(b->n_jobs, c->d_jobs.p, c->d_results.p, c->d_step_scan.p, c->d_step_scan.p + nb, c->d_steps.p,
                             c->d_steps_compact.p, c->stream);
        CU(cudaEventRecord(c->ev[4], c->stream));  // run_ms includes the compaction
        if (!async) CU(cudaEventSynchronize(c->ev[4]));
    }
    st.kernel_launches = 6 + st.fill_launches + st.traceback_launches + st.jobs_strip_groups;
    st.run_ms = 0;
    if (!async && !b->groups.empty()) {
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[7], c->ev[4]);
        st.run_ms = ms;
    }
    b->ran = true;
    return PG2_OK;
}

extern "C" int pg2_batch_run(pg2_ctx *c, pg2_batch *b) { return batch_run_impl(c, b, false); }

// Results come back in two steps: fetch_enqueue puts the device->host copies of the result records and of the total
// number of path words (into the ctx's pinned staging) on the ctx stream behind the kernels; fetch_complete waits for
// them, copies exactly that many words of the compacted, run-length encoded paths into the caller's buffer and fills
// the caller's pg2_result records (step_off = running sum of n_steps in job order, as the device compacted them).
static int fetch_enqueue(pg2_ctx *c, pg2_batch *b, int64_t step_cap) {
    CU(cudaSetDevice(c->device));
    if (step_cap < b->total_steps) return fail(PG2_ERR_CAPACITY, "step buffer too small");
    c->h_results.clear();
    c->h_step_total.clear();
    DevResult *hr = c->h_results.extend((size_t)b->n_jobs + 1);
    long long *ht = c->h_step_total.extend(1);
    if (!hr || !ht) return fail(PG2_ERR_NOMEM, "pinned staging allocation failed");
    *ht = 0;
    if (b->n_jobs > 0) {
        CU(cudaMemcpyAsync(hr, c->d_results.p, sizeof(DevResult) * b->n_jobs, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(ht, c->d_step_scan.p + (b->jobs.size() / 256 + 2), sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    }
    b->fetch_enqueued = true;
    return PG2_OK;
}

static int fetch_complete(pg2_ctx *c, pg2_batch *b, pg2_result *results, uint16_t *steps) {
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    const DevResult *hr = c->h_results.p;
    const long long total = b->n_jobs > 0 ? c->h_step_total.p[0] : 0;
    if (total < 0 || total > b->total_steps) return fail(PG2_ERR_CUDA, "path compaction returned an impossible word count");
    CU(cudaEventRecord(c->ev[5], c->stream));  // d2h_ms: the copy of the path words
    if (total > 0) CU(cudaMemcpyAsync(steps, c->d_steps_compact.p, sizeof(unsigned short) * (size_t)total, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->ev[6], c->stream));
    CU(cudaStreamSynchronize(c->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]);
    c->stats.d2h_ms = ms;
    c->stats.d2h_bytes = (long long)sizeof(DevResult) * b->n_jobs + (long long)sizeof(unsigned short) * total + (long long)sizeof(long long);
    long long off = 0;
    for (int t = 0; t < b->n_jobs; t++) {
        const DevJob &J = b->jobs[t];
        pg2_result &r = results[t];
        r.score = hr[t].score;
        r.cells = J.cells;
        r.step_off = off;
        r.n_steps = hr[t].n_steps;
        off += hr[t].n_steps;
        r.status = hr[t].status == JOB_UNSUPPORTED ? PG2_JOB_BAD_GRAPH : hr[t].status;
        r.end_ptr = hr[t].end_ptr;
        r.kernel = J.kernel;
        if (hr[t].status == JOB_UNSUPPORTED) { g_last_error = "a graph exceeds PG2_MAX_IN_DEGREE backward edges per site"; }
    }
    if (off != total) return fail(PG2_ERR_CUDA, "path compaction: word counts do not add up");
    return PG2_OK;
}

extern "C" int pg2_batch_fetch(pg2_ctx *c, pg2_batch *b, pg2_result *results, uint16_t *steps, int64_t step_cap) {
    if (!c || !b || c->current != b || !b->ran || (b->n_jobs > 0 && (!results || !steps))) return fail(PG2_ERR_INVALID, "pg2_batch_fetch: bad argument or batch not run");
    if (step_cap < b->total_steps) {
        for (int t = 0; t < b->n_jobs; t++) results[t].n_steps = b->jobs[t].step_cap;
        return fail(PG2_ERR_CAPACITY, "step buffer too small");
    }
    int rc = fetch_enqueue(c, b, step_cap);
    if (rc != PG2_OK) return rc;
    return fetch_complete(c, b, results, steps);
}

extern "C" void pg2_batch_destroy(pg2_ctx *c, pg2_batch *b) {
    if (!b) return;
    if (c) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        if (c->current == b) c->current = nullptr;
    }
    delete b;
}

// One launch batch from host buffers to host buffers.  Large batches are cut into chunks that rotate over the ctx
// and its siblings (further sets of staging / device buffers, each with its own stream): the host packs chunk k+1
// while the device computes chunk k, the kernels of consecutive chunks overlap on the device (the tail of one
// launch is filled by the next), and the (compacted) results of a chunk are collected while later chunks compute.  Jobs that share the left
// graph stay in one chunk, so the lane kernel keeps full tasks.
static int align_batch_single(pg2_ctx *c, int32_t n_jobs, const pg2_job *jobs, pg2_result *results, uint16_t *steps, int64_t step_cap) {
    pg2_batch *b = nullptr;
    PackTimer timer;
    int rc = pg2_batch_create(c, n_jobs, jobs, &b);
    if (rc != PG2_OK) return rc;
    timer.lap("= create");
    rc = pg2_batch_run(c, b);
    timer.lap("= run (upload + kernels)");
    if (rc == PG2_OK) rc = pg2_batch_fetch(c, b, results, steps, step_cap);
    timer.lap("= fetch");
    pg2_batch_destroy(c, b);
    return rc;
}

static int ensure_siblings(pg2_ctx *c, int n_slots) {
    for (int k = 0; k + 1 < n_slots; k++) {
        if (!c->sibling[k]) {
            pg2_ctx *s = nullptr;
            int rc = pg2_ctx_create(c->device, &s);
            if (rc != PG2_OK) return rc;
            s->borrowed_models = true;
            s->scratch_bytes = c->scratch_bytes / PIPE_SLOTS;
            s->force_wavefront = c->force_wavefront;
            s->no_lanes = c->no_lanes;
            s->no_pstrip = c->no_pstrip;
            s->pstrip_max_jobs = c->pstrip_max_jobs;
            s->pstrip_banded_chains = c->pstrip_banded_chains;
            s->pstrip_cluster_max = c->pstrip_cluster_max;
            c->sibling[k] = s;
        }
        // Chunk k of a pipelined call runs on slot k: the earlier chunk's stream gets the higher priority, so that the SM slots
        // a finishing launch gives back go to the OLDEST launch that still has CTAs waiting.  Without this the last three
        // launches of a call share the slots evenly and all end raggedly (measured: ends at 60.3 / 62.2 / 64.4 ms).
        if (!c->sibling[k]->prio_set) {
            int prio_lo = 0, prio_hi = 0;
            cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // numerically lower = higher priority; prio_hi is the tracebacks'
            const int levels = prio_lo - prio_hi;                  // levels below the traceback streams
            if (levels >= 2) {
                const int prio = prio_hi + 1 + ((k + 1) * levels) / PIPE_SLOTS;
                cudaStream_t st = 0;
                if (cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, std::min(prio, prio_lo)) == cudaSuccess) {
                    cudaStreamDestroy(c->sibling[k]->stream);
                    c->sibling[k]->stream = st;
                } else cudaGetLastError();
            }
            c->sibling[k]->prio_set = true;
        }
        c->sibling[k]->models = c->models;  // same device tables
        c->sibling[k]->models_dirty = true;
    }
    return PG2_OK;
}

extern "C" int pg2_align_batch(pg2_ctx *c, int32_t n_jobs, const pg2_job *jobs, pg2_result *results, uint16_t *steps, int64_t step_cap) {
    if (!c || n_jobs < 0 || (n_jobs > 0 && (!jobs || !results || !steps))) return fail(PG2_ERR_INVALID, "pg2_align_batch: bad argument");
    const char *np = getenv("PG2_NO_PIPELINE");
    const char *mj = getenv("PG2_PIPELINE_MIN_JOBS");  // tests force the pipelined path on small batches
    const int min_jobs = (mj && atoi(mj) > 0) ? atoi(mj) : 8192;
    if (n_jobs < min_jobs || (np && atoi(np) != 0)) return align_batch_single(c, n_jobs, jobs, results, steps, step_cap);
    PackTimer timer;
    // group the jobs by left graph (first-appearance order), then cut into chunks by cell count
    long long est_tasks = 0;  // lane tasks the call would form if every group went to the lane kernel
    std::vector<int> perm(n_jobs);
    std::vector<long long> cells_prefix((size_t)n_jobs + 1, 0);
    {
        // open-addressing table on the left graph's `state` pointer -> group id
        // (the table grows with the number of DISTINCT left graphs -- 127 in a placement batch of 100 000 reads)
        size_t cap = 1024;
        struct Slot { const void *key; int gid; };
        std::vector<Slot> table(cap, Slot{nullptr, -1});
        auto probe = [&](const void *key) {
            size_t i = ((((size_t)key >> 4) * 0x9E3779B97F4A7C15ull) >> 24) & (cap - 1);
            while (table[i].gid >= 0 && table[i].key != key) i = (i + 1) & (cap - 1);
            return i;
        };
        std::vector<int> gid(n_jobs), count, gsize;
        std::vector<long long> cells(n_jobs);  // read in job order here; the permuted prefix sum below stays inside this array
        const void *last_key = nullptr;
        int last_gid = -1;
        for (int t = 0; t < n_jobs; t++) {
            const pg2_job &j = jobs[t];
            const void *key = j.left.state;
            if (key != last_key) {
                size_t i = probe(key);
                if (table[i].gid < 0) {
                    if ((count.size() + 1) * 2 > cap) {  // keep the load under one half
                        std::vector<Slot> old_table;
                        old_table.swap(table);
                        cap <<= 2;
                        table.assign(cap, Slot{nullptr, -1});
                        for (const Slot &o : old_table) if (o.gid >= 0) table[probe(o.key)] = o;
                        i = probe(key);
                    }
                    table[i].key = key;
                    table[i].gid = (int)count.size();
                    count.push_back(0);
                    gsize.push_back(j.left.n_sites);
                }
                last_key = key;
                last_gid = table[i].gid;
            }
            gid[t] = last_gid;
            count[last_gid]++;
            cells[t] = (long long)std::max(j.left.n_sites, 1) * std::max(j.right.n_sites, 1);
        }
        // groups with the longest left graph first: the launches of the last chunks then hold the shortest tasks (the
        // tail of the call), as a single launch's longest-first task order would
        std::vector<int> gorder(count.size());
        for (size_t g = 0; g < gorder.size(); g++) gorder[g] = (int)g;
        if (!getenv("PG2_PIPELINE_KEEP_ORDER"))
            std::stable_sort(gorder.begin(), gorder.end(), [&](int a, int b2) { return gsize[a] > gsize[b2]; });
        std::vector<int> start(count.size() + 1, 0);
        {
            int acc = 0;
            for (int g : gorder) { start[g] = acc; acc += count[g]; }
        }
        for (int t = 0; t < n_jobs; t++) perm[start[gid[t]]++] = t;
        for (int k = 0; k < n_jobs; k++) cells_prefix[k + 1] = cells_prefix[k] + cells[perm[k]];
        for (int n : count) est_tasks += (n + 31) / 32;
    }
    const long long total = cells_prefix[n_jobs];
    int n_chunks = (int)std::min<long long>(8, std::max<long long>(2, total / 2500000000LL));
    const char *nc = getenv("PG2_PIPELINE_CHUNKS");
    if (nc && atoi(nc) > 0) n_chunks = atoi(nc);
    // chunk weights: a small first chunk puts the device to work early, a small last chunk keeps the copy back of
    // the final results short; the chunks in between amortise the per-chunk packing and launches
    std::vector<double> weight((size_t)n_chunks, 3.0);
    if (n_chunks >= 4) { weight[0] = 1.0; weight[1] = 2.0; weight[(size_t)n_chunks - 1] = 2.0; }
    if (const char *pw = getenv("PG2_PIPELINE_WEIGHTS")) {  // tuning: comma-separated relative chunk sizes
        std::vector<double> w;
        for (const char *q = pw; *q;) { char *e; double v = strtod(q, &e); if (e == q) break; if (v > 0) w.push_back(v); q = *e ? e + 1 : e; }
        if (w.size() >= 2) { weight = w; n_chunks = (int)w.size(); }
    }
    double wsum = 0;
    for (double w : weight) wsum += w;
    std::vector<int> cut(1, 0);
    double wacc = 0;
    for (int k = 1; k < n_chunks; k++) {
        wacc += weight[(size_t)k - 1];
        const long long want = (long long)((double)total * (wacc / wsum));
        int pos = (int)(std::lower_bound(cells_prefix.begin(), cells_prefix.end(), want) - cells_prefix.begin());
        // cells_prefix has n_jobs + 1 entries: a dominant last job puts the search at n_jobs, where there is nothing to cut
        // (and perm[] ends)
        if (pos >= n_jobs) continue;
        // prefer a boundary between two left graphs; inside a big group cut at a multiple of 32 jobs
        int lo = pos;
        while (lo > cut.back() && jobs[perm[lo]].left.state == jobs[perm[lo - 1]].left.state && pos - lo < 1024) lo--;
        if (lo > cut.back() && jobs[perm[lo]].left.state != jobs[perm[lo - 1]].left.state) pos = lo;
        else pos = cut.back() + ((pos - cut.back()) & ~31);
        if (pos > cut.back() && pos < n_jobs) cut.push_back(pos);
    }
    cut.push_back(n_jobs);
    int n_slots = PIPE_SLOTS;
    const char *ns = getenv("PG2_PIPELINE_SLOTS");
    if (ns && atoi(ns) >= 1 && atoi(ns) <= PIPE_SLOTS) n_slots = atoi(ns);
    n_slots = std::min<int>(n_slots, (int)cut.size() - 1);
    int rc = ensure_siblings(c, n_slots);
    if (rc != PG2_OK) return rc;
    // every slot gets the same share of the scratch budget while chunks are in flight side by side (the siblings were made
    // with budget / PIPE_SLOTS; the primary goes back to its full budget afterwards)
    struct BudgetGuard {
        pg2_ctx *c; size_t saved;
        ~BudgetGuard() { c->scratch_bytes = saved; }
    } budget_guard = {c, c->scratch_bytes};
    if (n_slots > 1) c->scratch_bytes = budget_guard.saved / PIPE_SLOTS;
    // the lane launches of all chunks take one shape: the latency shape when the whole call cannot keep the chip's CTA slots busy
    // for long (batch_run_impl)
    bool wide_call;
    {
        const char *ft = getenv("PG2_LANE_WIDE_TASKS");
        const int per_sm = (ft && atoi(ft) >= 0) ? atoi(ft) : LANE_WIDE_TASKS_PER_SM;
        wide_call = est_tasks <= (long long)per_sm * std::max(c->prop.multiProcessorCount, 1);
    }
    timer.lap("= group + chunk");
    // PG2_TIMING: device timeline of the chunks against one origin (tuning aid)
    cudaEvent_t origin = nullptr;
    const auto host_origin = std::chrono::steady_clock::now();
    if (timer.on) { cudaEventCreate(&origin); cudaEventRecord(origin, c->stream); }

    // Chunk k runs on slot k % n_slots: pack, upload, kernels and (into pinned caller memory) the copy back are all
    // enqueued without waiting; the host only waits for a chunk when its slot is needed again or at the end.
    struct InFlight { pg2_ctx *ctx = nullptr; pg2_batch *batch = nullptr; int lo = 0, hi = 0; long long step_base = 0; };
    InFlight slot[PIPE_SLOTS];
    std::vector<pg2_job> chunk_jobs;
    std::vector<pg2_result> chunk_res;
    long long step_base = 0;
    pg2_stats agg;
    memset(&agg, 0, sizeof agg);
    auto finish = [&](InFlight &f) -> int {
        if (!f.batch) return PG2_OK;
        const int n = f.hi - f.lo;
        chunk_res.resize((size_t)n);
        int r = PG2_OK;
        if (!f.batch->fetch_enqueued) r = fetch_enqueue(f.ctx, f.batch, step_cap - f.step_base);
        if (r == PG2_OK) r = fetch_complete(f.ctx, f.batch, chunk_res.data(), steps + f.step_base);
        if (r == PG2_OK && origin) {
            float t_up = 0, t_k0 = 0, t_fill = 0, t_tb = 0, t_d2h = 0;
            cudaEventElapsedTime(&t_up, origin, f.ctx->ev[0]);
            cudaEventElapsedTime(&t_k0, origin, f.ctx->ev[7]);
            cudaEventElapsedTime(&t_fill, origin, f.ctx->ev[3]);
            cudaEventElapsedTime(&t_tb, origin, f.ctx->ev[4]);
            cudaEventElapsedTime(&t_d2h, origin, f.ctx->ev[6]);
            fprintf(stderr, "pg2 chunk [%6d,%6d): upload starts %6.2f  kernels start %6.2f  fill ends %6.2f  traceback ends %6.2f  d2h ends %6.2f  "
                            "(host: collected at %6.2f ms)\n", f.lo, f.hi, t_up, t_k0, t_fill, t_tb, t_d2h,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_origin).count());
        }
        if (r == PG2_OK) {
            for (int k = 0; k < n; k++) {
                pg2_result &o = results[perm[f.lo + k]];
                o = chunk_res[(size_t)k];
                o.step_off += f.step_base;
            }
            const pg2_stats &st = f.ctx->stats;
            agg.h2d_bytes += st.h2d_bytes; agg.d2h_bytes += st.d2h_bytes; agg.cells += st.cells; agg.traceback_bytes += st.traceback_bytes;
            agg.fill_launches += st.fill_launches; agg.traceback_launches += st.traceback_launches; agg.kernel_launches += st.kernel_launches;
            agg.jobs_wavefront += st.jobs_wavefront; agg.jobs_strip += st.jobs_strip; agg.jobs_lanes += st.jobs_lanes; agg.jobs_pstrip += st.jobs_pstrip; agg.jobs_band += st.jobs_band; agg.jobs_pstrip_ring += st.jobs_pstrip_ring; agg.jobs_lanes_wide += st.jobs_lanes_wide;
            agg.jobs_strip_groups += st.jobs_strip_groups; agg.d2h_ms += st.d2h_ms;
        }
        pg2_batch_destroy(f.ctx, f.batch);
        f.batch = nullptr;
        return r;
    };
    rc = PG2_OK;
    const size_t n_cut = cut.size() - 1;
    for (size_t k = 0; k < n_cut && rc == PG2_OK; k++) {
        InFlight &f = slot[k % (size_t)n_slots];
        rc = finish(f);  // the chunk that used this slot n_slots rounds ago
        if (rc != PG2_OK) break;
        const size_t si = k % (size_t)n_slots;
        f.ctx = si == 0 ? c : c->sibling[si - 1];
        f.lo = cut[k];
        f.hi = cut[k + 1];
        f.step_base = step_base;
        chunk_jobs.resize((size_t)(f.hi - f.lo));
        long long cap = 0;
        for (int i = f.lo; i < f.hi; i++) {
            chunk_jobs[(size_t)(i - f.lo)] = jobs[perm[i]];
            cap += jobs[perm[i]].left.n_sites + jobs[perm[i]].right.n_sites;
        }
        if (step_base + cap > step_cap) { rc = fail(PG2_ERR_CAPACITY, "step buffer too small"); break; }
        step_base += cap;
        rc = pg2_batch_create(f.ctx, f.hi - f.lo, chunk_jobs.data(), &f.batch);
        f.ctx->lane_wide_hint = wide_call ? 1 : 0;
        if (rc == PG2_OK) rc = batch_run_impl(f.ctx, f.batch, true);
        f.ctx->lane_wide_hint = -1;
        if (rc == PG2_OK) rc = fetch_enqueue(f.ctx, f.batch, step_cap - f.step_base);
        if (timer.on)
            fprintf(stderr, "pg2 chunk [%6d,%6d): enqueued at %6.2f ms (host)\n", f.lo, f.hi,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_origin).count());
        if (rc != PG2_OK && f.batch) { pg2_batch_destroy(f.ctx, f.batch); f.batch = nullptr; }
    }
    for (size_t k = 0; k < (size_t)n_slots; k++) {
        // oldest first
        InFlight &f = slot[(n_cut + k) % (size_t)n_slots];
        int r = finish(f);
        if (rc == PG2_OK) rc = r;
    }
    if (rc == PG2_OK) c->stats = agg;
    if (origin) cudaEventDestroy(origin);
    timer.lap("= pipelined chunks");
    return rc;
}

// Device addresses of the batch's result records (pg2_device.cuh DevResult, 24 bytes each, job order) and
// packed-pointer buffer, valid until the next batch is created on this ctx.  For callers that forward
// results GPU-to-GPU (the multi-GPU gather to rank 0 goes over NCCL without a host bounce).
extern "C" int pg2_batch_device_buffers(pg2_ctx *c, pg2_batch *b, void **results_dev, void **steps_dev, int64_t *n_steps_total) {
    if (!c || !b || c->current != b || !b->ran) return fail(PG2_ERR_INVALID, "pg2_batch_device_buffers: batch not run");
    if (results_dev) *results_dev = c->d_results.p;
    if (steps_dev) *steps_dev = c->d_steps_compact.p;
    if (n_steps_total) {
        long long total = 0;
        CU(cudaSetDevice(c->device));
        if (b->n_jobs > 0) {
            CU(cudaMemcpyAsync(&total, c->d_step_scan.p + (b->jobs.size() / 256 + 2), sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        }
        *n_steps_total = total;
    }
    return PG2_OK;
}

extern "C" int pg2_stream_synchronize(pg2_ctx *c) {
    if (!c) return fail(PG2_ERR_INVALID, "pg2_stream_synchronize: null ctx");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return PG2_OK;
}

extern "C" int64_t pg2_batch_step_capacity(const pg2_batch *b) { return b ? b->total_steps : 0; }

extern "C" int pg2_get_stats(pg2_ctx *c, pg2_stats *out) {
    if (!c || !out) return fail(PG2_ERR_INVALID, "pg2_get_stats: bad argument");
    *out = c->stats;
    return PG2_OK;
}
