// pg2_traceback.cu -- validation prepass and the backtrack kernel.
//
// validate_kernel: one warp per distinct graph / per job checks what the reference takes for granted
// (edges point to earlier sites, states index the substitution table, in-degree fits the packed pointer,
// the band is monotone and contains (0,0); tunnel_matrix.h:163-168).  Failing jobs get a status and are
// skipped by the fill kernels -- the host never walks the graphs.
//
// traceback_kernel: walks the packed back-pointers from the end corner to (0,0) exactly as
// backtrack_new_path does (reference src/main/viterbi_alignment.cpp:1038-1189) and emits the pointer of
// every visited cell in walk order.  The host unpacker (pg2_expand.cpp) turns that into the reference's
// vector<Path_pointer>, including the skipped-site steps of insert_preexisting_gap.
#include <string.h>
#include "pg2_device.cuh"
#include "pg2_strip_geom.cuh"
#include "pg2_pstrip_geom.cuh"

namespace pg2 {

// per-lane partial check of one graph: lanes stride over the sites
__device__ __forceinline__ void check_graph_lane(const DevGraph &G, const int *d_off, const int *d_estart, int lane, int nlanes,
                                                 int &bad, int &maxdeg, int &simple) {
    const int *off = d_off + G.off_base;
    const int *es = d_estart + G.edge_base;
    bad = 0; maxdeg = 0; simple = 1;
    if (G.n_sites < 2) { bad = 1; return; }
    for (int s = lane; s < G.n_sites; s += nlanes) {
        int k0 = off[s], k1 = off[s + 1];
        if (k1 < k0 || (s == 0 && k0 != 0)) { bad = 1; break; }
        int deg = k1 - k0;
        if (deg > maxdeg) maxdeg = deg;
        if (s == 0) { if (deg != 0) bad = 1; }
        else if (deg != 1 || es[k0] != s - 1) simple = 0;
        for (int k = k0; k < k1; ++k) {
            int p = es[k];
            if (p < 0 || p >= s) bad = 1;
        }
    }
}

__device__ __forceinline__ void finish_graph(DevGraph *graphs, int *graph_status, int g, int bad, int maxdeg, int simple) {
    graphs[g].max_indeg = maxdeg;
    graphs[g].simple = simple && !bad;
    graph_status[g] = bad ? JOB_BAD_GRAPH : (maxdeg > 63 ? JOB_UNSUPPORTED : JOB_OK);
}

// per-lane partial check of one job; returns 1 when this lane saw a problem of the given kind
__device__ __forceinline__ int check_states_lane(const DevJob &J, const DevGraph *graphs, const DevModel *models, const int *d_state,
                                                 int lane, int nlanes) {
    // states of real sites must index the table (model->log_score, evol_model.h:80)
    const DevGraph GL = graphs[J.left], GR = graphs[J.right];
    int fas = models[J.model].fas;
    int bad = 0;
    for (int s = 1 + lane; s < GL.n_sites - 1; s += nlanes) {
        int st = d_state[GL.state_base + s];
        if (st < 0 || st >= fas) bad = 1;
    }
    for (int s = 1 + lane; s < GR.n_sites - 1; s += nlanes) {
        int st = d_state[GR.state_base + s];
        if (st < 0 || st >= fas) bad = 1;
    }
    return bad;
}

__device__ __forceinline__ int check_band_lane(const DevJob &J, const int *d_blo, const int *d_bhi, int lane, int nlanes) {
    // d_blo/d_bhi hold the CLIPPED band; monotone + (0,0) inside (tunnel_matrix.h:163-168, viterbi_alignment.cpp:729)
    const int *lo = d_blo + J.band_base, *hi = d_bhi + J.band_base;
    int bad = 0;
    for (int i = lane; i < J.lx; i += nlanes) {
        if (lo[i] > hi[i]) bad = 1;
        if (i == 0) { if (lo[0] > 0) bad = 1; }
        else if (lo[i] < lo[i - 1] || hi[i] < hi[i - 1]) bad = 1;
    }
    return bad;
}

__device__ __forceinline__ void finish_job(DevResult *results, int jid, int status) {
    results[jid].status = status;
    results[jid].score = neg_inf();
    results[jid].end_ptr = NO_MAT;
    results[jid].n_steps = 0;
}

#ifndef PG2_HOST_EMU
__global__ void validate_graphs_kernel(int n_graphs, DevGraph *graphs, const int *d_state, const int *d_off, const int *d_estart,
                                       int *graph_status) {
    int g = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
    int lane = threadIdx.x & 31;
    if (g >= n_graphs) return;
    int bad, maxdeg, simple;
    check_graph_lane(graphs[g], d_off, d_estart, lane, 32, bad, maxdeg, simple);
    for (int o = 16; o; o >>= 1) {
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
        maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
        simple &= __shfl_xor_sync(0xffffffffu, simple, o);
    }
    if (lane == 0) finish_graph(graphs, graph_status, g, bad, maxdeg, simple);
}

__global__ void validate_jobs_kernel(int n_jobs, const DevJob *jobs, const DevGraph *graphs, const DevModel *models,
                                     const int *graph_status, const int *d_state, const int *d_blo, const int *d_bhi,
                                     DevResult *results) {
    int jid = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
    int lane = threadIdx.x & 31;
    if (jid >= n_jobs) return;
    const DevJob J = jobs[jid];
    int status = JOB_OK;
    int gs_l = graph_status[J.left], gs_r = graph_status[J.right];
    if (gs_l != JOB_OK) status = gs_l;
    else if (gs_r != JOB_OK) status = gs_r;
    if (status == JOB_OK && __any_sync(0xffffffffu, check_states_lane(J, graphs, models, d_state, lane, 32))) status = JOB_BAD_GRAPH;
    if (status == JOB_OK && J.banded && __any_sync(0xffffffffu, check_band_lane(J, d_blo, d_bhi, lane, 32))) status = JOB_BAD_BAND;
    if (lane == 0) finish_job(results, jid, status);
}
#endif

// Path words are run-length encoded as they are emitted (pagan2_b200.h): a word with bit 15 clear is a packed pointer,
// a word with bit 15 set repeats the previous pointer (word & 0x7fff) more times.  A run of L equal pointers costs
// min(L, 2) words, so the encoding never needs more room than the raw walk (<= Lx + Ly words per job).
struct StepEmit {
    unsigned short *out;
    int n;         // words written
    int raw;       // pointers emitted (what the reference's path length bounds)
    unsigned last; // previous pointer (0xffffffff: none)
    int rep;       // pending repeats of `last`
};
__device__ __forceinline__ void emit_begin(StepEmit &e, unsigned short *out) { e.out = out; e.n = 0; e.raw = 0; e.last = 0xffffffffu; e.rep = 0; }
__device__ __forceinline__ void emit_flush(StepEmit &e) {
    if (e.rep > 0) { e.out[e.n++] = (unsigned short)(0x8000u | (unsigned)e.rep); e.rep = 0; }
}
__device__ __forceinline__ void emit_step(StepEmit &e, unsigned q) {
    q &= 0x3fffu;
    ++e.raw;
    if (q == e.last && e.rep < 0x7fff) { ++e.rep; return; }
    emit_flush(e);
    e.out[e.n++] = (unsigned short)q;
    e.last = q;
}

// Cell pointer lookup for both fill kernels' layouts.
struct TraceCtx {
    const DevJob *J;
    const int *vlast;  // strip layout: per site of the row graph, the virtual row that completed it
    int nv;
    const int *blo, *bhi, *dlo;
    const long long *doff;
    const unsigned *ptr32;        // wavefront layout: anti-diagonal-major, one word per cell
    const unsigned short *ptr16;  // strip layout: row-major [i][j], one half-word per cell
};

// `chain_left`: set when the cell was filled by an in-place fast row of a register-strip kernel, i.e. its only left
// edge is (i-1 -> i): the walk then needs no CSR lookup for the left graph.
__device__ __forceinline__ bool fetch_ptr(const TraceCtx &t, int mat, int i, int j, unsigned &out, bool &chain_left) {
    const DevJob &J = *t.J;
    chain_left = false;
    if (i < 0 || j < 0 || i >= J.lx || j >= J.ly) return false;
    if (J.kernel == 0) {
        long long idx;
        if (J.banded) {
            if (j < t.blo[i] || j > t.bhi[i]) return false;
            int s = i + j;
            idx = t.doff[s] + (i - t.dlo[s]);
        } else {
            int s = i + j;
            idx = diag_cum(s, J.lx, J.ly) + (i - diag_lo(s, J.ly));
        }
        out = word_ptr(t.ptr32[J.cell_base + idx], mat);
    } else {
        if (i == 0 && j == 0) { out = NO_MAT; return true; }  // start corner: no predecessor
        const long long idx = J.kernel == 2 ? lane_ptr_index(t.nv, LANE_K, t.vlast[i], j, J.lane)
                                            : strip_ptr_index(t.nv, J.ly, J.strip_k, t.vlast[i], j);
        const unsigned w = t.ptr16[J.cell_base + idx];
        chain_left = (w & 0x4000u) != 0;
        out = J.kernel == 2 ? lane_decode_ptr(w, mat) : strip_decode_ptr(w, mat);
    }
    return true;
}

__device__ void traceback_one(int jid, const DevJob *jobs, const DevGraph *graphs, const int *d_vlast, const int *d_off,
                              const int *d_estart, const int *d_blo, const int *d_bhi, const int *d_dlo,
                              const long long *d_doff, const unsigned *ptr32, const unsigned short *ptr16, unsigned short *steps,
                              DevResult *results) {
    const DevJob J = jobs[jid];  // by value: the walk's stores must not force a reload of the job record per step
    DevResult *res = results + jid;
    if (res->status != JOB_OK) return;
    const DevGraph GL = graphs[J.left], GR = graphs[J.right];
    const int *l_off = d_off + GL.off_base, *r_off = d_off + GR.off_base;
    const int *l_es = d_estart + GL.edge_base, *r_es = d_estart + GR.edge_base;
    TraceCtx tc;
    tc.J = &J;
    tc.vlast = (J.kernel != 0) ? d_vlast + GL.vlast_base : nullptr;
    tc.nv = GL.n_vrows;
    tc.blo = J.banded ? d_blo + J.band_base : nullptr;
    tc.bhi = J.banded ? d_bhi + J.band_base : nullptr;
    tc.dlo = J.banded ? d_dlo + J.diag_base : nullptr;
    tc.doff = J.banded ? d_doff + J.diag_base : nullptr;
    tc.ptr32 = ptr32;
    tc.ptr16 = ptr16;
    StepEmit em;
    emit_begin(em, steps + J.step_base);

    // end pointer -> last alignment column (viterbi_alignment.cpp:1047-1069)
    unsigned p = res->end_ptr;
    int vit = (int)(p & 3u);
    int i, j;
    if (vit == M_MAT) { i = l_es[l_off[J.lx] + ((p >> 2) & 63u)]; j = r_es[r_off[J.ly] + ((p >> 8) & 63u)]; }
    else if (vit == X_MAT) { i = l_es[l_off[J.lx] + ((p >> 2) & 63u)]; j = J.ly - 1; }
    else if (vit == Y_MAT) { i = J.lx - 1; j = r_es[r_off[J.ly] + ((p >> 8) & 63u)]; }
    else { res->status = JOB_NO_PATH; return; }
    emit_step(em, p);

    int status = JOB_OK;
    // the reference loop ends when (i<1 && j<1) AFTER reading the cell it stands on (:1073-1181)
    // register-strip kernels only take jobs whose right graph is a plain chain: column j is entered from j-1
    const bool chain_right = J.kernel != 0;
    // lane layout: the virtual row of site i travels with the walk (a plain row sits one virtual row below its only
    // source), so plain rows cost one load per step -- the half-word itself.  With 100 000 paths in flight the walk is
    // bound by the rate of 32-byte sectors (one per step), not by latency: a look-ahead window that fetched the next 8
    // rows and diagonal cells at once doubled the sector count and the time (2.1 -> 4.3 ms).
    int v = -1;  // vlast[i], or -1: not known
    for (;;) {
        unsigned q;
        bool chain_left;
        if (J.kernel == 2) {
            if (vit == NO_MAT || i < 0 || j < 0 || i >= J.lx || j >= J.ly) { status = JOB_BROKEN_PATH; break; }
            if (i == 0 && j == 0) { q = NO_MAT; chain_left = false; }  // start corner: no predecessor
            else {
                if (v < 0) v = tc.vlast[i];
                const unsigned w = tc.ptr16[J.cell_base + lane_ptr_index(tc.nv, LANE_K, v, j, J.lane)];
                chain_left = (w & 0x4000u) != 0;
                q = lane_decode_ptr(w, vit);
            }
        } else if (vit == NO_MAT || !fetch_ptr(tc, vit, i, j, q, chain_left)) { status = JOB_BROKEN_PATH; break; }
        if (em.raw >= J.step_cap) { status = JOB_BROKEN_PATH; break; }
        emit_step(em, q);
        int src = (int)(q & 3u);
        if (vit == M_MAT || vit == X_MAT) {
            if (src == NO_MAT) i = -1;
            else if (chain_left) { i -= 1; if (v >= 0) v -= 1; }
            else { i = l_es[l_off[i] + ((q >> 2) & 63u)]; v = -1; }
        }
        if (vit == M_MAT || vit == Y_MAT)
            j = (src == NO_MAT) ? -1 : (chain_right ? j - 1 : r_es[r_off[j] + ((q >> 8) & 63u)]);
        vit = src;
        if (i < 1 && j < 1) break;
    }
    emit_flush(em);
    res->n_steps = em.n;
    res->pad = em.raw;
    res->status = status;
}

// ---- wavefront-layout jobs: one warp per path, the pointer words around the walk fetched a window at a time ----
// The walk is a chain of dependent loads; with one thread per path every step costs two L2/HBM round trips (band
// geometry, then the pointer word) -- 0.5 s for the 400 000 steps of a 200 kb anchored alignment.  Here the warp
// loads the words of diagonals s0-63..s0, rows i0-31..i0 (64 coalesced loads, all in flight at once) into shared memory
// and lane 0 walks inside the window: at least 16 steps per round trip.
constexpr int TW = 32;   // rows per window (one per lane)
constexpr int TWD = 64;  // diagonals per window: a run of 32 match steps goes up 32 rows and 64 diagonals
struct TraceWin {
    unsigned w[TWD * TW];  // [diagonal s0-d][row i0-r]
    long long base[TWD];   // pointer-buffer offset of row lo[d] on diagonal s0-d
    int lo[TWD], hi[TWD];  // in-band rows of diagonal s0-d (hi < lo: none)
    int s0, i0;
};

__device__ __forceinline__ void win_geometry(const DevJob &J, const TraceCtx &t, TraceWin &W, int d) {
    const int s = W.s0 - d;
    if (s < 0) { W.lo[d] = 0; W.hi[d] = -1; W.base[d] = 0; return; }
    if (J.banded) {
        const long long b = t.doff[s];
        W.lo[d] = t.dlo[s];
        W.hi[d] = W.lo[d] + (int)(t.doff[s + 1] - b) - 1;
        W.base[d] = b;
    } else {
        W.lo[d] = diag_lo(s, J.ly);
        W.hi[d] = diag_hi(s, J.lx);
        W.base[d] = diag_cum(s, J.lx, J.ly);
    }
}
__device__ __forceinline__ void win_load(const DevJob &J, const TraceCtx &t, TraceWin &W, int d, int r) {
    const int i = W.i0 - r;
    unsigned v = 0;
    if (i >= W.lo[d] && i <= W.hi[d]) v = t.ptr32[J.cell_base + W.base[d] + (i - W.lo[d])];
    W.w[d * TW + r] = v;
}

// walk state of one path (held by lane 0)
struct TraceState {
    int i, j, vit, status;
    StepEmit em;
    bool done;
};

__device__ __forceinline__ void trace_wave_begin(const DevJob &J, DevResult *res, const int *l_off, const int *r_off, const int *l_es,
                                                 const int *r_es, unsigned short *out, TraceState &st) {
    // end pointer -> last alignment column (viterbi_alignment.cpp:1047-1069)
    const unsigned p = res->end_ptr;
    st.vit = (int)(p & 3u);
    emit_begin(st.em, out);
    st.status = JOB_OK;
    st.done = false;
    st.i = st.j = 0;
    if (st.vit == M_MAT) { st.i = l_es[l_off[J.lx] + ((p >> 2) & 63u)]; st.j = r_es[r_off[J.ly] + ((p >> 8) & 63u)]; }
    else if (st.vit == X_MAT) { st.i = l_es[l_off[J.lx] + ((p >> 2) & 63u)]; st.j = J.ly - 1; }
    else if (st.vit == Y_MAT) { st.i = J.lx - 1; st.j = r_es[r_off[J.ly] + ((p >> 8) & 63u)]; }
    else { st.status = JOB_NO_PATH; st.done = true; return; }
    emit_step(st.em, p);
}

// walks while the cells are inside the window; returns with st.done set, or at a cell outside the window
__device__ __forceinline__ void trace_wave_walk(const DevJob &J, const TraceWin &W, const int *l_off, const int *r_off, const int *l_es,
                                                const int *r_es, unsigned short *out, TraceState &st) {
    for (;;) {
        const int i = st.i, j = st.j;
        if (st.vit == NO_MAT || i < 0 || j < 0 || i >= J.lx || j >= J.ly) { st.status = JOB_BROKEN_PATH; st.done = true; return; }
        const int d = W.s0 - (i + j), r = W.i0 - i;
        if (d < 0 || d >= TWD || r < 0 || r >= TW) return;  // next window
        if (i < W.lo[d] || i > W.hi[d]) { st.status = JOB_BROKEN_PATH; st.done = true; return; }  // outside the band
        if (st.em.raw >= J.step_cap) { st.status = JOB_BROKEN_PATH; st.done = true; return; }
        const unsigned w = W.w[d * TW + r];
        const unsigned q = word_ptr(w, st.vit);
        emit_step(st.em, q);
        const int src = (int)(q & 3u);
        // the reference loop ends when (i<1 && j<1) AFTER reading the cell it stands on (:1073-1181)
        if (st.vit == M_MAT || st.vit == X_MAT)
            st.i = (src == NO_MAT) ? -1 : ((w & WORD_PLAIN_LEFT) ? i - 1 : l_es[l_off[i] + ((q >> 2) & 63u)]);
        if (st.vit == M_MAT || st.vit == Y_MAT)
            st.j = (src == NO_MAT) ? -1 : ((w & WORD_PLAIN_RIGHT) ? j - 1 : r_es[r_off[j] + ((q >> 8) & 63u)]);
        st.vit = src;
        if (st.i < 1 && st.j < 1) { st.done = true; return; }
    }
}

#ifndef PG2_HOST_EMU
__global__ void __launch_bounds__(128) traceback_wave_kernel(int n_jobs, const int *job_ids, const DevJob *jobs, const DevGraph *graphs,
                                                             const int *d_off, const int *d_estart, const int *d_blo, const int *d_bhi,
                                                             const int *d_dlo, const long long *d_doff, const unsigned *ptr32,
                                                             unsigned short *steps, DevResult *results) {
    __shared__ TraceWin wins[4];  // 4 x 9.2 KB
    const int lane = threadIdx.x & 31, t = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (t >= n_jobs) return;
    const int jid = job_ids[t];
    const DevJob J = jobs[jid];
    DevResult *res = results + jid;
    if (J.kernel != 0 || res->status != JOB_OK) return;  // warp-uniform
    TraceWin &W = wins[threadIdx.x >> 5];
    const DevGraph GL = graphs[J.left], GR = graphs[J.right];
    const int *l_off = d_off + GL.off_base, *r_off = d_off + GR.off_base;
    const int *l_es = d_estart + GL.edge_base, *r_es = d_estart + GR.edge_base;
    TraceCtx tc;
    tc.J = &J; tc.vlast = nullptr; tc.nv = 0;
    tc.blo = nullptr; tc.bhi = nullptr;
    tc.dlo = J.banded ? d_dlo + J.diag_base : nullptr;
    tc.doff = J.banded ? d_doff + J.diag_base : nullptr;
    tc.ptr32 = ptr32; tc.ptr16 = nullptr;
    unsigned short *out = steps + J.step_base;
    TraceState st;
    st.i = st.j = 0; st.vit = NO_MAT; st.status = JOB_OK; st.done = false;
    emit_begin(st.em, out);
    if (lane == 0) trace_wave_begin(J, res, l_off, r_off, l_es, r_es, out, st);
    for (;;) {
        const int done = __shfl_sync(0xffffffffu, (int)st.done, 0);
        if (done) break;
        const int bi = __shfl_sync(0xffffffffu, st.i, 0), bj = __shfl_sync(0xffffffffu, st.j, 0);
        if (bi < 0 || bj < 0 || bi >= J.lx || bj >= J.ly) {  // the walk left the matrix: lane 0 reports it
            if (lane == 0) { st.status = JOB_BROKEN_PATH; st.done = true; }
            continue;
        }
        if (lane == 0) { W.s0 = bi + bj; W.i0 = bi; }
        __syncwarp();
        win_geometry(J, tc, W, lane);
        win_geometry(J, tc, W, lane + 32);
        __syncwarp();
        {
            // all 64 loads of the lane in flight at once: addresses first, then the loads, then the stores
            const unsigned *src[TWD];
#pragma unroll
            for (int d = 0; d < TWD; ++d) {
                const int i = W.i0 - lane;
                src[d] = (i >= W.lo[d] && i <= W.hi[d]) ? tc.ptr32 + J.cell_base + W.base[d] + (i - W.lo[d]) : nullptr;
            }
            unsigned v[TWD];
#pragma unroll
            for (int d = 0; d < TWD; ++d) v[d] = src[d] ? __ldg(src[d]) : 0u;
#pragma unroll
            for (int d = 0; d < TWD; ++d) W.w[d * TW + lane] = v[d];
        }
        __syncwarp();
        if (lane == 0) trace_wave_walk(J, W, l_off, r_off, l_es, r_es, out, st);
        __syncwarp();
    }
    if (lane == 0) { emit_flush(st.em); res->n_steps = st.em.n; res->pad = st.em.raw; res->status = st.status; }
}

__global__ void traceback_kernel(int n_jobs, const int *job_ids, const DevJob *jobs, const DevGraph *graphs, const int *d_vlast,
                                 const int *d_off,
                                 const int *d_estart, const int *d_blo, const int *d_bhi, const int *d_dlo,
                                 const long long *d_doff, const unsigned *ptr32, const unsigned short *ptr16, unsigned short *steps,
                                 DevResult *results) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_jobs) return;
    if (jobs[job_ids[t]].kernel == 0 || jobs[job_ids[t]].kernel >= 3) return;  // traceback_wave_kernel / traceback_pstrip_kernel / the band kernel's walk
    traceback_one(job_ids[t], jobs, graphs, d_vlast, d_off, d_estart, d_blo, d_bhi, d_dlo, d_doff, ptr32, ptr16, steps, results);
}
#endif

// ---- pipelined-strip layout (pg2_pstrip_geom.cuh): one warp per path, a window of one block's step rows at a time ----
// The pointer words of a block lie step-major ([step][lane][K], step = virtual row - v0 + lane).  The warp loads the
// PTW_WORDS / (32 K) step rows that end at the walk's current step into shared memory (coalesced, all loads in flight)
// and lane 0 walks inside them; a plain row / plain column flag in the word saves the CSR lookup of the next site.
constexpr int PTW_WORDS = 4096;
struct PsTraceWin {
    unsigned w[PTW_WORDS];
    int b, t_hi;  // block and last step held (b < 0: nothing loaded)
};
struct PsTraceJob {
    const int *blocks, *colinfo, *vlast;
    const unsigned *P;
    int K, rows;  // step rows per window
};
struct PsTraceState {
    TraceState s;
    int v;                      // virtual row that completed site s.i, or -1: not known
    int b, c0, c1, v0, v1, off; // block of column s.j (b < 0: not known)
    int need_t;                 // the step the walk stopped at (to be the last step of the next window)
};

__device__ __forceinline__ void ps_trace_block(const PsTraceJob &tj, PsTraceState &st, int j) {
    if (st.b >= 0 && j >= st.c0 && j < st.c1) return;
    st.b = tj.colinfo[j] >> PC_BLOCK_SHIFT;
    const int *e = tj.blocks + st.b * PB_INTS;
    st.c0 = e[0]; st.c1 = e[1]; st.v0 = e[2]; st.v1 = e[3]; st.off = e[5];
}

// walks while the cells are inside the window; returns with st.s.done set, or at a cell outside the window
__device__ __forceinline__ void ps_trace_walk(const DevJob &J, const PsTraceJob &tj, const PsTraceWin &W, const int *l_off, const int *r_off,
                                              const int *l_es, const int *r_es, PsTraceState &st) {
    TraceState &s = st.s;
    for (;;) {
        const int i = s.i, j = s.j;
        if (s.vit == NO_MAT || i < 0 || j < 0 || i >= J.lx || j >= J.ly) { s.status = JOB_BROKEN_PATH; s.done = true; return; }
        ps_trace_block(tj, st, j);
        if (st.v < 0) st.v = tj.vlast[i];
        if (st.v < st.v0 || st.v >= st.v1) { s.status = JOB_BROKEN_PATH; s.done = true; return; }  // a row the band keeps out of the block
        const int jj = j - st.c0, l = jj / tj.K, k = jj - l * tj.K;
        const int t = st.v - st.v0 + l;
        if (W.b != st.b || t > W.t_hi || t <= W.t_hi - tj.rows) { st.need_t = t; return; }  // next window
        if (s.em.raw >= J.step_cap) { s.status = JOB_BROKEN_PATH; s.done = true; return; }
        const unsigned w = W.w[((W.t_hi - t) * 32 + l) * tj.K + k];
        const unsigned q = word_ptr(w, s.vit);
        emit_step(s.em, q);
        const int src = (int)(q & 3u);
        // the reference loop ends when (i<1 && j<1) AFTER reading the cell it stands on (:1073-1181)
        if (s.vit == M_MAT || s.vit == X_MAT) {
            if (src == NO_MAT) { s.i = -1; st.v = -1; }
            else if (w & PSW_PLAIN_ROW) { s.i = i - 1; st.v -= 1; }
            else { s.i = l_es[l_off[i] + ((q >> 2) & 63u)]; st.v = -1; }
        }
        if (s.vit == M_MAT || s.vit == Y_MAT)
            s.j = (src == NO_MAT) ? -1 : ((w & PSW_PLAIN_COL) ? j - 1 : r_es[r_off[j] + ((q >> 8) & 63u)]);
        s.vit = src;
        if (s.i < 1 && s.j < 1) { s.done = true; return; }
    }
}

// one lane's share of a window load: step rows t_hi, t_hi - 1, ... of block b; rows outside the block read as 0
template <int K>
__device__ __forceinline__ void ps_trace_load(const PsTraceJob &tj, PsTraceWin &W, int b, int t_hi, int lane) {
    const int *e = tj.blocks + b * PB_INTS;
    const int n_steps = e[3] - e[2] + 31;
    const unsigned *base = tj.P + e[5] + lane * K;
    constexpr int ROWS = PTW_WORDS / (32 * K), U = 16;
#pragma unroll 1
    for (int r0 = 0; r0 < ROWS; r0 += U) {
        unsigned v[U][K];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int ts = t_hi - (r0 + u);
            const bool ok = ts >= 0 && ts < n_steps;
#pragma unroll
            for (int k = 0; k < K; ++k) v[u][k] = ok ? __ldg(base + (long long)ts * 32 * K + k) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) W.w[((r0 + u) * 32 + lane) * K + k] = v[u][k];
    }
}

__device__ __forceinline__ void ps_trace_setup(const DevJob &J, const DevGraph &GL, const DevGraph &GR, const int *d_vlast, const unsigned *ptrps,
                                               PsTraceJob &tj) {
    tj.blocks = d_vlast + J.blk_base;
    tj.colinfo = d_vlast + GR.cp_ci_base;
    tj.vlast = d_vlast + GL.vlast_base;
    tj.P = ptrps + J.cell_base;
    tj.K = J.strip_k;
    tj.rows = PTW_WORDS / (32 * J.strip_k);
}

#ifndef PG2_HOST_EMU
__global__ void __launch_bounds__(64) traceback_pstrip_kernel(int n_jobs, const int *job_ids, const DevJob *jobs, const DevGraph *graphs,
                                                              const int *d_vlast, const int *d_off, const int *d_estart, const unsigned *ptrps,
                                                              unsigned short *steps, DevResult *results) {
    __shared__ PsTraceWin wins[2];
    const int lane = threadIdx.x & 31, t = blockIdx.x * 2 + (threadIdx.x >> 5);
    if (t >= n_jobs) return;
    const int jid = job_ids[t];
    const DevJob J = jobs[jid];
    DevResult *res = results + jid;
    if (J.kernel != 3 || res->status != JOB_OK) return;  // warp-uniform
    PsTraceWin &W = wins[threadIdx.x >> 5];
    const DevGraph GL = graphs[J.left], GR = graphs[J.right];
    const int *l_off = d_off + GL.off_base, *r_off = d_off + GR.off_base;
    const int *l_es = d_estart + GL.edge_base, *r_es = d_estart + GR.edge_base;
    PsTraceJob tj;
    ps_trace_setup(J, GL, GR, d_vlast, ptrps, tj);
    unsigned short *out = steps + J.step_base;
    PsTraceState st;
    st.s.i = st.s.j = 0; st.s.vit = NO_MAT; st.s.status = JOB_OK; st.s.done = false;
    st.v = -1; st.b = -1; st.c0 = st.c1 = st.v0 = st.v1 = st.off = 0; st.need_t = 0;
    emit_begin(st.s.em, out);
    if (lane == 0) { W.b = -1; W.t_hi = 0; trace_wave_begin(J, res, l_off, r_off, l_es, r_es, out, st.s); }
    __syncwarp();
    for (;;) {
        if (lane == 0 && !st.s.done) ps_trace_walk(J, tj, W, l_off, r_off, l_es, r_es, st);
        const int done = __shfl_sync(0xffffffffu, (int)st.s.done, 0);
        if (done) break;
        const int nb = __shfl_sync(0xffffffffu, st.b, 0), nt = __shfl_sync(0xffffffffu, st.need_t, 0);
        __syncwarp();
        if (tj.K == 1) ps_trace_load<1>(tj, W, nb, nt, lane);
        else if (tj.K == 2) ps_trace_load<2>(tj, W, nb, nt, lane);
        else ps_trace_load<4>(tj, W, nb, nt, lane);
        if (lane == 0) { W.b = nb; W.t_hi = nt; }
        __syncwarp();
    }
    if (lane == 0) { emit_flush(st.s.em); res->n_steps = st.s.em.n; res->pad = st.s.em.raw; res->status = st.s.status; }
}
#endif

// CSR of an implicit chain: site s >= 1 is entered by the one edge (s-1 -> s) with log weight +0.0
__device__ __forceinline__ void expand_chain(const DevGraph &G, int *d_off, int *d_estart, float *d_elogw, int lane, int nlanes) {
    int *off = d_off + G.off_base, *es = d_estart + G.edge_base;
    float *ew = d_elogw + G.edge_base;
    for (int s = lane; s <= G.n_sites; s += nlanes) off[s] = s > 0 ? s - 1 : 0;
    for (int k = lane; k < G.n_sites - 1; k += nlanes) { es[k] = k; ew[k] = 0.0f; }
}
#ifndef PG2_HOST_EMU
__global__ void expand_implicit_kernel(int n_graphs, const DevGraph *graphs, int *d_off, int *d_estart, float *d_elogw) {
    int g = blockIdx.x * (blockDim.x / 32) + (threadIdx.x / 32);
    if (g >= n_graphs || !graphs[g].implicit) return;
    expand_chain(graphs[g], d_off, d_estart, d_elogw, threadIdx.x & 31, 32);
}
#endif
void launch_expand_implicit(int n_graphs, const DevGraph *graphs, int *d_off, int *d_estart, float *d_elogw, cudaStream_t stream) {
    if (n_graphs <= 0) return;
#ifndef PG2_HOST_EMU
    const int warps = 8;
    expand_implicit_kernel<<<(n_graphs + warps - 1) / warps, warps * 32, 0, stream>>>(n_graphs, graphs, d_off, d_estart, d_elogw);
#else
    (void)stream;
    for (int g = 0; g < n_graphs; ++g)
        if (graphs[g].implicit) expand_chain(graphs[g], d_off, d_estart, d_elogw, 0, 1);
#endif
}

#ifndef PG2_HOST_EMU
// the same checks with a whole CTA per graph / job: a launch batch of a few long graphs (a guide-tree wave of 200 kb
// sequences) would otherwise wait for one warp per graph to stride over 200 000 sites (6.6 + 3 ms per 16 alignments)
__global__ void __launch_bounds__(256) validate_graphs_cta_kernel(int n_graphs, DevGraph *graphs, const int *d_state, const int *d_off,
                                                                  const int *d_estart, int *graph_status) {
    __shared__ int s_maxdeg;
    const int g = blockIdx.x;
    if (threadIdx.x == 0) s_maxdeg = 0;
    __syncthreads();
    int bad, maxdeg, simple;
    check_graph_lane(graphs[g], d_off, d_estart, threadIdx.x, blockDim.x, bad, maxdeg, simple);
    atomicMax(&s_maxdeg, maxdeg);
    bad = __syncthreads_or(bad);
    simple = __syncthreads_and(simple);
    if (threadIdx.x == 0) finish_graph(graphs, graph_status, g, bad, s_maxdeg, simple);
}
__global__ void __launch_bounds__(256) validate_jobs_cta_kernel(int n_jobs, const DevJob *jobs, const DevGraph *graphs, const DevModel *models,
                                                                const int *graph_status, const int *d_state, const int *d_blo,
                                                                const int *d_bhi, DevResult *results) {
    const int jid = blockIdx.x;
    const DevJob J = jobs[jid];
    int status = JOB_OK;
    const int gs_l = graph_status[J.left], gs_r = graph_status[J.right];
    if (gs_l != JOB_OK) status = gs_l;
    else if (gs_r != JOB_OK) status = gs_r;
    const int bad_state = __syncthreads_or(status == JOB_OK ? check_states_lane(J, graphs, models, d_state, threadIdx.x, blockDim.x) : 0);
    if (status == JOB_OK && bad_state) status = JOB_BAD_GRAPH;
    const int bad_band = __syncthreads_or((status == JOB_OK && J.banded) ? check_band_lane(J, d_blo, d_bhi, threadIdx.x, blockDim.x) : 0);
    if (status == JOB_OK && bad_band) status = JOB_BAD_BAND;
    if (threadIdx.x == 0) finish_job(results, jid, status);
}
#endif

// few_long: the batch holds at most a few thousand graphs and some of them are long -- a CTA per graph / job
void launch_validate(int n_graphs, int n_jobs, DevGraph *graphs, const DevJob *jobs, const DevModel *models, const int *d_state,
                     const int *d_off, const int *d_estart, const int *d_blo, const int *d_bhi, int *graph_status,
                     DevResult *results, bool few_long, cudaStream_t stream) {
#ifndef PG2_HOST_EMU
    const int warps = 8;
    if (few_long) {
        if (n_graphs > 0) validate_graphs_cta_kernel<<<n_graphs, 256, 0, stream>>>(n_graphs, graphs, d_state, d_off, d_estart, graph_status);
        if (n_jobs > 0)
            validate_jobs_cta_kernel<<<n_jobs, 256, 0, stream>>>(n_jobs, jobs, graphs, models, graph_status, d_state, d_blo, d_bhi, results);
        return;
    }
    if (n_graphs > 0)
        validate_graphs_kernel<<<(n_graphs + warps - 1) / warps, warps * 32, 0, stream>>>(n_graphs, graphs, d_state, d_off, d_estart,
                                                                                         graph_status);
    if (n_jobs > 0)
        validate_jobs_kernel<<<(n_jobs + warps - 1) / warps, warps * 32, 0, stream>>>(n_jobs, jobs, graphs, models, graph_status,
                                                                                     d_state, d_blo, d_bhi, results);
#else
    (void)stream;
    for (int g = 0; g < n_graphs; ++g) {
        int bad, maxdeg, simple;
        check_graph_lane(graphs[g], d_off, d_estart, 0, 1, bad, maxdeg, simple);
        finish_graph(graphs, graph_status, g, bad, maxdeg, simple);
    }
    for (int jid = 0; jid < n_jobs; ++jid) {
        const DevJob J = jobs[jid];
        int status = JOB_OK;
        if (graph_status[J.left] != JOB_OK) status = graph_status[J.left];
        else if (graph_status[J.right] != JOB_OK) status = graph_status[J.right];
        if (status == JOB_OK && check_states_lane(J, graphs, models, d_state, 0, 1)) status = JOB_BAD_GRAPH;
        if (status == JOB_OK && J.banded && check_band_lane(J, d_blo, d_bhi, 0, 1)) status = JOB_BAD_BAND;
        finish_job(results, jid, status);
    }
#endif
}

// n_wave: how many of the jobs were filled by the wavefront kernel (they get the warp-per-path walk)
void launch_traceback(int n_jobs, int n_wave, int n_ps, const int *job_ids, const DevJob *jobs, const DevGraph *graphs, const int *d_vlast, const int *d_off,
                      const int *d_estart, const int *d_blo, const int *d_bhi, const int *d_dlo, const long long *d_doff,
                      const unsigned *ptr32, const unsigned short *ptr16, const unsigned *ptrps, unsigned short *steps, DevResult *results,
                      cudaStream_t stream) {
    if (n_jobs <= 0) return;
#ifndef PG2_HOST_EMU
    const int threads = 64;
    if (n_ps > 0)
        traceback_pstrip_kernel<<<(n_jobs + 1) / 2, 64, 0, stream>>>(n_jobs, job_ids, jobs, graphs, d_vlast, d_off, d_estart, ptrps, steps, results);
    if (n_wave + n_ps < n_jobs)
        traceback_kernel<<<(n_jobs + threads - 1) / threads, threads, 0, stream>>>(n_jobs, job_ids, jobs, graphs, d_vlast, d_off, d_estart, d_blo,
                                                                                  d_bhi, d_dlo, d_doff, ptr32, ptr16, steps, results);
    if (n_wave > 0)
        traceback_wave_kernel<<<(n_jobs + 3) / 4, 128, 0, stream>>>(n_jobs, job_ids, jobs, graphs, d_off, d_estart, d_blo, d_bhi, d_dlo, d_doff,
                                                                   ptr32, steps, results);
#else
    (void)stream; (void)n_wave; (void)n_ps;
    for (int t = 0; t < n_jobs; ++t) {
        const int jid = job_ids[t];
        const DevJob J = jobs[jid];
        if (J.kernel == 3) {
            // the warp-per-path walk of the pipelined-strip layout, lanes one after the other
            DevResult *res = results + jid;
            if (res->status != JOB_OK) continue;
            const DevGraph GL = graphs[J.left], GR = graphs[J.right];
            const int *l_off = d_off + GL.off_base, *r_off = d_off + GR.off_base;
            const int *l_es = d_estart + GL.edge_base, *r_es = d_estart + GR.edge_base;
            PsTraceJob tj;
            ps_trace_setup(J, GL, GR, d_vlast, ptrps, tj);
            unsigned short *out = steps + J.step_base;
            PsTraceState st;
            st.v = -1; st.b = -1; st.c0 = st.c1 = st.v0 = st.v1 = st.off = 0; st.need_t = 0;
            static thread_local PsTraceWin W;  // 16 KB: not on the stack; one per host thread (several contexts may run side by side)
            W.b = -1; W.t_hi = 0;
            trace_wave_begin(J, res, l_off, r_off, l_es, r_es, out, st.s);
            while (!st.s.done) {
                ps_trace_walk(J, tj, W, l_off, r_off, l_es, r_es, st);
                if (st.s.done) break;
                for (int lane = 0; lane < 32; ++lane) {
                    if (tj.K == 1) ps_trace_load<1>(tj, W, st.b, st.need_t, lane);
                    else if (tj.K == 2) ps_trace_load<2>(tj, W, st.b, st.need_t, lane);
                    else ps_trace_load<4>(tj, W, st.b, st.need_t, lane);
                }
                W.b = st.b; W.t_hi = st.need_t;
            }
            emit_flush(st.s.em);
            res->n_steps = st.s.em.n;
            res->pad = st.s.em.raw;
            res->status = st.s.status;
            continue;
        }
        if (J.kernel == 4) continue;  // launch_band_traceback (pg2_band.cu)
        if (J.kernel != 0) {
            traceback_one(jid, jobs, graphs, d_vlast, d_off, d_estart, d_blo, d_bhi, d_dlo, d_doff, ptr32, ptr16, steps, results);
            continue;
        }
        // the warp-per-path walk, lanes one after the other
        DevResult *res = results + jid;
        if (res->status != JOB_OK) continue;
        const DevGraph GL = graphs[J.left], GR = graphs[J.right];
        const int *l_off = d_off + GL.off_base, *r_off = d_off + GR.off_base;
        const int *l_es = d_estart + GL.edge_base, *r_es = d_estart + GR.edge_base;
        TraceCtx tc;
        tc.J = &J; tc.vlast = nullptr; tc.nv = 0;
        tc.blo = nullptr; tc.bhi = nullptr;
        tc.dlo = J.banded ? d_dlo + J.diag_base : nullptr;
        tc.doff = J.banded ? d_doff + J.diag_base : nullptr;
        tc.ptr32 = ptr32; tc.ptr16 = nullptr;
        unsigned short *out = steps + J.step_base;
        TraceState st;
        TraceWin W;
        trace_wave_begin(J, res, l_off, r_off, l_es, r_es, out, st);
        while (!st.done) {
            if (st.i < 0 || st.j < 0 || st.i >= J.lx || st.j >= J.ly) { st.status = JOB_BROKEN_PATH; break; }
            W.s0 = st.i + st.j; W.i0 = st.i;
            for (int d = 0; d < TWD; ++d) win_geometry(J, tc, W, d);
            for (int d = 0; d < TWD; ++d)
                for (int lane = 0; lane < TW; ++lane) win_load(J, tc, W, d, lane);
            trace_wave_walk(J, W, l_off, r_off, l_es, r_es, out, st);
        }
        emit_flush(st.em);
        res->n_steps = st.em.n;
        res->pad = st.em.raw;
        res->status = st.status;
    }
#endif
}

// ---- compaction: the encoded words of all jobs back to back, in job order ----
// offsets = exclusive prefix sum of n_steps over the job index; three small kernels (block sums, scan of the block sums,
// block-local scan + copy, one warp per 32 jobs).
constexpr int CS_BLOCK = 256;
#ifndef PG2_HOST_EMU
__global__ void __launch_bounds__(CS_BLOCK) steps_block_sum_kernel(int n_jobs, const DevResult *results, long long *block_sum) {
    __shared__ int warp_sum[CS_BLOCK / 32];
    const int t = blockIdx.x * CS_BLOCK + threadIdx.x;
    int v = t < n_jobs ? results[t].n_steps : 0;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int w = 0; w < CS_BLOCK / 32; ++w) s += warp_sum[w];
        block_sum[blockIdx.x] = s;
    }
}
__global__ void steps_scan_blocks_kernel(int n_blocks, long long *block_sum, long long *total) {
    // in place: block_sum[b] becomes the offset of block b; one thread, n_blocks is n_jobs / 256
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long run = 0;
    for (int b = 0; b < n_blocks; ++b) { const long long v = block_sum[b]; block_sum[b] = run; run += v; }
    *total = run;
}
__global__ void __launch_bounds__(CS_BLOCK) steps_compact_kernel(int n_jobs, const DevJob *jobs, const DevResult *results,
                                                                 const long long *block_off, const unsigned short *steps_in,
                                                                 unsigned short *steps_out) {
    __shared__ long long s_off[CS_BLOCK];
    __shared__ int s_n[CS_BLOCK];
    const int t = blockIdx.x * CS_BLOCK + threadIdx.x;
    s_n[threadIdx.x] = t < n_jobs ? results[t].n_steps : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = block_off[blockIdx.x];
        for (int k = 0; k < CS_BLOCK; ++k) { s_off[k] = run; run += s_n[k]; }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp * 32; k < warp * 32 + 32; ++k) {
        const int job = blockIdx.x * CS_BLOCK + k;
        if (job >= n_jobs) break;
        const unsigned short *src = steps_in + jobs[job].step_base;
        unsigned short *dst = steps_out + s_off[k];
        for (int e = lane; e < s_n[k]; e += 32) dst[e] = src[e];
    }
}
#endif
void launch_compact_steps(int n_jobs, const DevJob *jobs, const DevResult *results, long long *block_scratch, long long *total,
                          const unsigned short *steps_in, unsigned short *steps_out, cudaStream_t stream) {
#ifndef PG2_HOST_EMU
    if (n_jobs <= 0) { cudaMemsetAsync(total, 0, sizeof(long long), stream); return; }
    const int nb = (n_jobs + CS_BLOCK - 1) / CS_BLOCK;
    steps_block_sum_kernel<<<nb, CS_BLOCK, 0, stream>>>(n_jobs, results, block_scratch);
    steps_scan_blocks_kernel<<<1, 32, 0, stream>>>(nb, block_scratch, total);
    steps_compact_kernel<<<nb, CS_BLOCK, 0, stream>>>(n_jobs, jobs, results, block_scratch, steps_in, steps_out);
#else
    (void)stream; (void)block_scratch;
    long long run = 0;
    for (int t = 0; t < n_jobs; ++t) {
        const int n = results[t].n_steps;
        memcpy(steps_out + run, steps_in + jobs[t].step_base, sizeof(unsigned short) * (size_t)n);
        run += n;
    }
    *total = run;
#endif
}

}  // namespace pg2
