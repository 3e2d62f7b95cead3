// pg2_lanes.cu -- placement fill kernel: one LANE per alignment, 32 alignments that share the row graph per warp.
//
// Query placement aligns many reads against the same tree node (reads_aligner.cpp:983-1216; the temporary
// node puts the tree node LEFT and the read RIGHT, reads_aligner.h:169-184), so a launch batch holds many jobs
// with the same LEFT graph.  The engine groups them into tasks of 32 (LaneTask, pg2_device.cuh); one warp takes
// a task and every lane aligns its own read against the shared row graph.
//
// Why: all lanes are on the SAME row at the same time, so everything the row graph decides -- in-degree, edge
// spans, edge weights, which rows must be parked for long-span edges -- is warp-uniform.  A site with one edge
// from the row above takes the in-place fast body; a multi-edge site loops over its edges; no lane ever waits
// for another lane's row shape, there is no skew ramp and no shuffle.  (The warp-per-alignment strip kernel,
// pg2_strip.cu, has 32 different rows in flight and falls back to its general body whenever one of them is not
// a plain row.)
//
// Layout.  Each lane sweeps its read in strips of K columns held in registers (pg2_rowmath.cuh), all rows of
// the row graph per strip.  The strip's last column is written to a per-warp boundary column [row][X,Y,M][lane]
// (one coalesced 256 B segment per component) and is the next strip's left neighbour; rows that are the source
// of a long-span edge are parked in a per-warp saved-row scratch [slot][k][X,Y,M][lane].  Back-pointers: one
// uint16 per cell, 8 columns per 128-bit store, lanes interleaved (pg2_strip_geom.cuh: lane_ptr_index).
//
// Arithmetic: the shared row bodies of pg2_rowmath.cuh -- candidate by candidate as the reference
// (src/main/viterbi_alignment.cpp:856-971, 1328-1436, 2029-2219), strict '>' first-wins.
#include "pg2_device.cuh"
#include "pg2_strip_geom.cuh"
#include "pg2_rowmath.cuh"
#ifdef PG2_HOST_EMU
#include <vector>
#endif

namespace pg2 {

struct LaneScratch {
    double *bcol0, *bcol1;  // [row][3][32]
    double *saved;          // [slot][K][3][32]
};

// The whole sweep of ONE lane.  Nothing here talks to another lane; warp-uniformity of the control flow comes
// from the data (the row program and the task are shared by the 32 lanes).
//   active   the lane holds a valid job (inactive lanes run along on dummy columns and store nothing)
//   max_ly   the task's column count (strips are swept for all lanes alike)
template <int K, bool GENERAL, bool SMALLTAB, bool WR>
__device__ __forceinline__ void lane_sweep(StripCtx c, const int lane, const bool active, const int max_ly, const int *r_state,
                                           const float *r_elogw, const LaneScratch sc, uint4 *ptr, DevResult *res) {
    const double ninf = neg_inf();
    const int n_strips = (max_ly + K - 1) / K;
    const int my_last_strip = (c.ly - 1) / K, my_last_k = (c.ly - 1) % K;
    constexpr int Q = K / 8;

    for (int s = 0; s < n_strips; ++s) {
        const int c0 = s * K;
        const bool first = (s == 0);
        double *bprev = (s & 1) ? sc.bcol0 : sc.bcol1;
        double *bcur = (s & 1) ? sc.bcol1 : sc.bcol0;
        c.c_block = c0;
        c.first_block = first;

        LaneState<K> st;
        LaneAcc<K> acc;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int j = c0 + k;
            const bool v = active && j < c.ly;
            st.X[k] = st.Y[k] = st.M[k] = ninf;
            st.extX[k] = (c.term && (j == 0 || j == c.ly - 1)) ? c.end_ext : c.ext;
            st.wr[k] = (WR && v && j >= 1) ? (double)r_elogw[j - 1] : 0.0;  // chain edge (j-1 -> j), CSR position j-1
            st.colbase[k] = (v && j >= 1) ? r_state[j] * c.fas : 0;
            acc.nX[k] = acc.nM[k] = ninf;
            acc.pX[k] = acc.pM[k] = NO_MAT;
        }
        st.penY1 = (c.reduced && first) ? 0.0 : c.open;  // Y move out of column 0 (basic_alignment.h:494)
        st.bX = st.bY = st.bM = ninf;

        const bool store_ptr = active && c0 < c.ly;
        const bool write_bcol = s < my_last_strip;   // a lane past its last strip leaves its end column alone
        const bool write_end = s == my_last_strip;

        // one-row-ahead prefetch of the row program entry and of the row's left neighbour
        int4 vr_n = __ldg(c.l_vrow);
        double nX = ninf, nY = ninf, nM = ninf;
        if (!first) {
            const double *b = bprev + (long long)vr_n.z * 96 + lane;
            nX = b[0]; nY = b[32]; nM = b[64];
        }
        for (int v = 0; v < c.nv; ++v) {
            const int4 vr = vr_n;
            const double rX = nX, rY = nY, rM = nM;
            if (v + 1 < c.nv) {
                vr_n = __ldg(c.l_vrow + v + 1);
                if (!first) {
                    const double *b = bprev + (long long)vr_n.z * 96 + lane;
                    nX = b[0]; nY = b[32]; nM = b[64];
                }
            }
            const int info = vr.x, i = vr.z, sl = info & VR_STATE_MASK;
            unsigned short w[K];
            bool done = true;
            if (!GENERAL || (info & VR_FAST) == VR_FAST) {
                const bool corner = first && i == 0;
                if (!GENERAL || (info & (VR_ZERO_W | VR_NOEDGE))) {
                    fast_row<K, false, WR, SMALLTAB>(c, st, i, sl, 0.0, corner, rX, rY, rM, w);
                } else {
                    const double wl = (double)c.l_elogw[vr.y];
                    fast_row<K, true, WR, SMALLTAB>(c, st, i, sl, wl, corner, rX, rY, rM, w);
                }
            } else {
                if (info & VR_FIRST) {
#pragma unroll
                    for (int k = 0; k < K; ++k) { acc.nX[k] = ninf; acc.nM[k] = ninf; acc.pX[k] = NO_MAT; acc.pM[k] = NO_MAT; }
                }
                if (!(info & VR_NOEDGE)) {
                    const int p = c.l_estart[vr.y];
                    const unsigned ord = ((unsigned)vr.w >> 16) << 2;
                    double sX[K + 1], sY[K + 1], sM[K + 1];
                    if (info & VR_REG) {
                        sX[0] = st.bX; sY[0] = st.bY; sM[0] = st.bM;
#pragma unroll
                        for (int k = 0; k < K; ++k) { sX[k + 1] = st.X[k]; sY[k + 1] = st.Y[k]; sM[k + 1] = st.M[k]; }
                    } else {
                        const int slot = vr.w & 0xffff;
                        const double *row = sc.saved + (long long)slot * K * 96 + lane;
                        sX[0] = sY[0] = sM[0] = ninf;
                        if (!first) {
                            const double *b = bprev + (long long)p * 96 + lane;
                            sX[0] = b[0]; sY[0] = b[32]; sM[0] = b[64];
                        }
#pragma unroll
                        for (int k = 0; k < K; ++k) { sX[k + 1] = row[k * 96]; sY[k + 1] = row[k * 96 + 32]; sM[k + 1] = row[k * 96 + 64]; }
                    }
                    if (info & VR_ZERO_W) {
                        accumulate_edge<K, SMALLTAB, false, WR>(c, st, acc, sl, p, 0.0, ord, sX, sY, sM);
                    } else {
                        const double wl = (double)c.l_elogw[vr.y];
                        accumulate_edge<K, SMALLTAB, true, WR>(c, st, acc, sl, p, wl, ord, sX, sY, sM);
                    }
                }
                done = (info & VR_LAST) != 0;
                if (done) commit_site<K>(c, st, acc, i, first, rX, rY, rM, w);
            }
            if (!done) continue;

            if (store_ptr) {
                uint4 *dst = ptr + (((long long)s * c.nv + v) * Q) * 32 + lane;
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    uint4 o;
                    o.x = (unsigned)w[8 * q + 0] | ((unsigned)w[8 * q + 1] << 16);
                    o.y = (unsigned)w[8 * q + 2] | ((unsigned)w[8 * q + 3] << 16);
                    o.z = (unsigned)w[8 * q + 4] | ((unsigned)w[8 * q + 5] << 16);
                    o.w = (unsigned)w[8 * q + 6] | ((unsigned)w[8 * q + 7] << 16);
                    dst[q * 32] = o;
                }
            }
            if (GENERAL) {
                const int slot = (int)((unsigned)info >> VR_SLOT_SHIFT) - 1;
                if (slot >= 0) {
                    double *row = sc.saved + (long long)slot * K * 96 + lane;
#pragma unroll
                    for (int k = 0; k < K; ++k) { row[k * 96] = st.X[k]; row[k * 96 + 32] = st.Y[k]; row[k * 96 + 64] = st.M[k]; }
                }
            }
            double *b = bcur + (long long)i * 96 + lane;
            if (write_bcol) {
                b[0] = st.X[K - 1]; b[32] = st.Y[K - 1]; b[64] = st.M[K - 1];
            } else if ((info & VR_ENDPRED) && write_end) {
                // rows the end corner reads: keep the lane's LAST column (ly-1) instead
                double vx = ninf, vy = ninf, vm = ninf;
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (k == my_last_k) { vx = st.X[k]; vy = st.Y[k]; vm = st.M[k]; }
                b[0] = vx; b[32] = vy; b[64] = vm;
            }
        }
    }
    if (!active) return;
    // iterate_bwd_edges_for_end_corner (:1440-1552) with a single right edge (ly-1 -> stop)
    const double *lastcol = ((my_last_strip & 1) ? sc.bcol1 : sc.bcol0) + lane;
    double best = ninf;
    unsigned bptr = NO_MAT;
    const int kl0 = c.l_off[c.lx], kl1 = c.l_off[c.lx + 1];
    const double wr = (double)r_elogw[c.ly - 1];  // edge (ly-1 -> ly)
    for (int kl = kl0; kl < kl1; ++kl) {
        const double *v = lastcol + (long long)c.l_estart[kl] * 96;
        double sc_ = __dadd_rn(__dadd_rn(__dadd_rn(v[64], c.lng), (double)c.l_elogw[kl]), wr);
        if (sc_ > best) { best = sc_; bptr = pack_ptr(M_MAT, kl - kl0, 0); }
        sc_ = v[0];  // score_gap_close: + 0
        if (sc_ > best) { best = sc_; bptr = pack_ptr(X_MAT, kl - kl0, 0); }
        if (kl == kl0) {
            sc_ = lastcol[(long long)(c.lx - 1) * 96 + 32];
            if (sc_ > best) { best = sc_; bptr = pack_ptr(Y_MAT, 0, 0); }
        }
    }
    res->score = best;
    res->end_ptr = bptr;
    res->status = (best == ninf) ? JOB_NO_PATH : JOB_OK;
}

__device__ __forceinline__ void lane_make_ctx(StripCtx &c, const LaneTask &T, int ly, const DevGraph &GL, const DevModel &m,
                                              const int *d_off, const int *d_estart, const float *d_elogw, const int4 *d_vrow, int K) {
    c.l_vrow = d_vrow + GL.vrow_base;
    c.nv = GL.n_vrows;
    c.l_off = d_off + GL.off_base;
    c.l_estart = d_estart + GL.edge_base;
    c.l_elogw = d_elogw + GL.edge_base;
    c.table = m.table;
    c.stab = nullptr;
    c.fas = m.fas;
    c.open = (double)m.open;
    c.ext = (double)m.ext;
    c.end_ext = (double)m.end_ext;
    c.lng = (double)m.lng;
    c.lng2 = (double)__fmul_rn(2.0f, m.lng);
    c.term = !(T.flags & FLAG_NO_TERMINAL_EDGES);
    c.reduced = (T.flags & FLAG_REDUCED) != 0;
    c.wr_zero = !(T.variant & 4);
    c.lx = GL.n_sites - 1;
    c.ly = ly;
    c.W = K;
    c.saved = nullptr;
    c.bcol_prev = c.bcol_cur = nullptr;
    c.ptr = nullptr;
    c.c_block = 0;
    c.first_block = true;
}

#ifndef PG2_HOST_EMU
#ifndef PG2_LANE_MINB
#define PG2_LANE_MINB 3
#endif
template <int K, bool GENERAL, bool SMALLTAB, bool WR>
__global__ void __launch_bounds__(128, PG2_LANE_MINB)
lane_fill_kernel(int n_tasks, const LaneTask *tasks, const DevJob *jobs, const DevGraph *graphs, const DevModel *models,
                 const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw, const int4 *d_vrow,
                 unsigned short *ptrs, DevResult *results, double *scratch, long long bcol_doubles, long long warp_doubles,
                 int *queue) {
    __shared__ double2 s_tab[SMALLTAB ? 4 : 1][SMALLTAB ? STRIP_SMALL_FAS * STRIP_SMALL_FAS : 1];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    LaneScratch sc;
    sc.bcol0 = scratch + (long long)warp_global * warp_doubles;
    sc.bcol1 = sc.bcol0 + bcol_doubles;
    sc.saved = sc.bcol1 + bcol_doubles;
    int tab_model = -1;

    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(queue, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= n_tasks) break;
        const LaneTask &T = tasks[q];
        const bool valid = lane < T.n_jobs;
        const int jid = T.job_ids[valid ? lane : 0];
        const DevJob &J = jobs[jid];
        DevResult *res = results + jid;
        const bool active = valid && res->status == JOB_OK;
        if (!__any_sync(0xffffffffu, active)) continue;
        const DevGraph GL = graphs[T.left], GR = graphs[J.right];
        const DevModel m = models[T.model];
        StripCtx c;
        lane_make_ctx(c, T, active ? J.ly : 1, GL, m, d_off, d_estart, d_elogw, d_vrow, K);
        if (SMALLTAB) {
            if (tab_model != T.model) {
                __syncwarp();
                for (int e = lane; e < m.fas * m.fas; e += 32) {
                    double ls = (double)m.table[e];
                    s_tab[wib][e] = make_double2(__dadd_rn(c.lng2, ls), __dadd_rn(c.lng, ls));
                }
                __syncwarp();
                tab_model = T.model;
            }
            c.stab = s_tab[wib];
        }
        lane_sweep<K, GENERAL, SMALLTAB, WR>(c, lane, active, T.max_ly, d_state + GR.state_base, d_elogw + GR.edge_base, sc,
                                            reinterpret_cast<uint4 *>(ptrs + T.ptr_base), res);
        __syncwarp();
    }
}
#endif

int lane_warps_per_sm() { return 4 * 3; }

// Launches one group of lane tasks that share the kernel variant.  `scratch`: n_warps * warp_doubles doubles.
void launch_lane_fill(int variant, int n_tasks, const LaneTask *tasks, const DevJob *jobs, const DevGraph *graphs,
                      const DevModel *models, const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw,
                      const int4 *d_vrow, unsigned short *ptrs, DevResult *results, double *scratch, long long bcol_doubles,
                      long long warp_doubles, int *queue, int n_warps, cudaStream_t stream) {
    if (n_tasks <= 0) return;
    constexpr int K = LANE_K;
#ifndef PG2_HOST_EMU
    cudaMemsetAsync(queue, 0, sizeof(int), stream);
    const int threads = 128;
    const int blocks = (n_warps * 32 + threads - 1) / threads;
#define PG2_LANE_LAUNCH(G, S, W)                                                                                              \
    lane_fill_kernel<K, G, S, W><<<blocks, threads, 0, stream>>>(n_tasks, tasks, jobs, graphs, models, d_state, d_off, d_estart, \
                                                                 d_elogw, d_vrow, ptrs, results, scratch, bcol_doubles,       \
                                                                 warp_doubles, queue)
    switch (variant & 7) {
        case 0: PG2_LANE_LAUNCH(false, false, false); break;
        case 1: PG2_LANE_LAUNCH(true, false, false); break;
        case 2: PG2_LANE_LAUNCH(false, true, false); break;
        case 3: PG2_LANE_LAUNCH(true, true, false); break;
        case 4: PG2_LANE_LAUNCH(false, false, true); break;
        case 5: PG2_LANE_LAUNCH(true, false, true); break;
        case 6: PG2_LANE_LAUNCH(false, true, true); break;
        case 7: PG2_LANE_LAUNCH(true, true, true); break;
    }
#undef PG2_LANE_LAUNCH
#else
    // CPU test emulation: lanes are independent, so each runs its whole sweep in turn
    (void)queue; (void)n_warps; (void)stream;
    LaneScratch sc;
    sc.bcol0 = scratch;
    sc.bcol1 = sc.bcol0 + bcol_doubles;
    sc.saved = sc.bcol1 + bcol_doubles;
    (void)warp_doubles;
    for (int q = 0; q < n_tasks; ++q) {
        const LaneTask &T = tasks[q];
        const DevGraph GL = graphs[T.left];
        const DevModel m = models[T.model];
        std::vector<double2> tab;
        for (int lane = 0; lane < 32; ++lane) {
            const bool valid = lane < T.n_jobs;
            const int jid = T.job_ids[valid ? lane : 0];
            const DevJob &J = jobs[jid];
            DevResult *res = results + jid;
            const bool active = valid && res->status == JOB_OK;
            if (!active) continue;  // an inactive lane stores nothing anybody reads
            const DevGraph GR = graphs[J.right];
            StripCtx c;
            lane_make_ctx(c, T, J.ly, GL, m, d_off, d_estart, d_elogw, d_vrow, K);
            if (variant & 2) {
                if (tab.empty()) {
                    tab.resize((size_t)m.fas * m.fas);
                    for (int e = 0; e < m.fas * m.fas; ++e) {
                        double ls = (double)m.table[e];
                        tab[e] = make_double2(c.lng2 + ls, c.lng + ls);
                    }
                }
                c.stab = tab.data();
            }
            uint4 *ptr = reinterpret_cast<uint4 *>(ptrs + T.ptr_base);
            const int *rs = d_state + GR.state_base;
            const float *rw = d_elogw + GR.edge_base;
            switch (variant & 7) {
                case 0: lane_sweep<K, false, false, false>(c, lane, true, T.max_ly, rs, rw, sc, ptr, res); break;
                case 1: lane_sweep<K, true, false, false>(c, lane, true, T.max_ly, rs, rw, sc, ptr, res); break;
                case 2: lane_sweep<K, false, true, false>(c, lane, true, T.max_ly, rs, rw, sc, ptr, res); break;
                case 3: lane_sweep<K, true, true, false>(c, lane, true, T.max_ly, rs, rw, sc, ptr, res); break;
                case 4: lane_sweep<K, false, false, true>(c, lane, true, T.max_ly, rs, rw, sc, ptr, res); break;
                case 5: lane_sweep<K, true, false, true>(c, lane, true, T.max_ly, rs, rw, sc, ptr, res); break;
                case 6: lane_sweep<K, false, true, true>(c, lane, true, T.max_ly, rs, rw, sc, ptr, res); break;
                case 7: lane_sweep<K, true, true, true>(c, lane, true, T.max_ly, rs, rw, sc, ptr, res); break;
            }
        }
    }
#endif
}

}  // namespace pg2
