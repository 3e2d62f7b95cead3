// pg2_strip.cu -- fast fill kernel: one WARP per alignment, register-resident column strips.
//
// Eligible jobs: unbanded, RIGHT graph a plain chain (leaf / read: in-degree 1, span 1, any edge
// weights), LEFT graph arbitrary (in-degree <= 16, any spans).  This is the query-placement shape
// (reads_aligner.h:169-184: tree node LEFT, read RIGHT) and the leaf-vs-leaf shape of the first tree wave.
//
// Layout.  The right graph's columns are cut into blocks of W = 32*K columns; lane l owns K consecutive
// columns of the block and keeps X/Y/M of its strip for the previous row in registers.  The warp sweeps
// the rows skewed by one row per lane (lane l is on row t-l at step t), so the only per-step exchange is
// the strip's last column handed to lane l+1 with warp shuffles.  Rows that are the source of a
// long-span edge are parked in a per-warp saved-row scratch (L2 resident); the block's last column is
// written to a per-warp boundary column (next block's left neighbour, and what the end corner reads).
// Back-pointers: one uint16 per cell, written step-major so that every step is one coalesced store.
//
// Row program.  The row graph is flattened on the host into VIRTUAL rows, one backward edge each (d_vrow,
// pg2_strip_geom.cuh): a site with three incoming edges takes three steps, a plain site one.  Lanes stay in
// lockstep on virtual rows, so the sweep costs sum(in-degree) + 31 steps instead of sum over steps of the
// largest in-degree among the 32 rows in flight.  Two step bodies: fast_row (the site has exactly one edge,
// from the row above: every row of a leaf graph, ~90 % of an ancestor graph) updates the strip in place;
// general_step accumulates one edge into per-lane accumulators and commits on the site's LAST virtual
// row.  The choice is warp-uniform per step (one __any_sync).
//
// Arithmetic follows the reference candidate by candidate (src/main/viterbi_alignment.cpp:856-971,
// 1328-1436, 2029-2219): same order, same FP64 association, strict '>' (first candidate wins ties).
// "+ 0.0" terms (log_gap_close == 0, zero log edge weights) are dropped: x + 0.0 == x for every x the DP
// can produce (no -0.0 arises: the corner is +0.0 and RN sums give -0.0 only from two -0.0 addends).
// A cell whose candidates are all -inf keeps an arbitrary pointer in fast_row: such a cell cannot lie on
// the Viterbi path, so the traceback never reads it.
#include "pg2_device.cuh"
#include "pg2_strip_geom.cuh"
#include "pg2_rowmath.cuh"
#ifdef PG2_HOST_EMU
#include <vector>
#endif

namespace pg2 {

bool strip_eligible(int lx, int ly, bool banded, int l_simple, int r_simple, int l_maxdeg, int r_maxdeg, int fas) {
    (void)l_simple; (void)r_maxdeg;
    if (banded) return false;
    if (!r_simple) return false;
    if (l_maxdeg > STRIP_MAX_LEFT_INDEG) return false;
    if (lx < 1 || ly < 1) return false;
    if (fas > VR_STATE_MASK) return false;
    return true;
}

// One virtual row of any shape.  `any_saved` (warp-uniform): some lane's edge starts at a parked row this
// step; when false every source is the lane's own previous-row strip and no staging happens.
// Returns true when the site was completed (st now holds row `i`, out_words are valid).
template <int K, bool SMALLTAB>
__device__ __forceinline__ bool general_step(const StripCtx &c, LaneState<K> &st, LaneAcc<K> &acc, int lane, int4 vr,
                                             bool any_saved, double recvX, double recvY, double recvM,
                                             unsigned short *out_words) {
    const double ninf = neg_inf();
    const int info = vr.x, i = vr.z;
    const int sl = info & VR_STATE_MASK;
    if (info & VR_FIRST) {
#pragma unroll
        for (int k = 0; k < K; ++k) { acc.nX[k] = ninf; acc.nM[k] = ninf; acc.pX[k] = NO_MAT; acc.pM[k] = NO_MAT; }
    }
    if (!(info & VR_NOEDGE)) {
        const int p = c.l_estart[vr.y];
        const double wl = (info & VR_ZERO_W) ? 0.0 : (double)c.l_elogw[vr.y];
        const unsigned ord = ((unsigned)vr.w >> 16) << 2;
        double sX[K + 1], sY[K + 1], sM[K + 1];
        sX[0] = st.bX; sY[0] = st.bY; sM[0] = st.bM;
#pragma unroll
        for (int k = 0; k < K; ++k) { sX[k + 1] = st.X[k]; sY[k + 1] = st.Y[k]; sM[k + 1] = st.M[k]; }
        if (any_saved && !(info & VR_REG)) {
            const int slot = vr.w & 0xffff;
            const double4 *row = c.saved + (long long)slot * c.W + lane * K;
            double4 bnd = make_double4(ninf, ninf, ninf, 0.0);
            if (lane > 0) bnd = row[-1];
            else if (!c.first_block) bnd = c.bcol_prev[p];
            sX[0] = bnd.x; sY[0] = bnd.y; sM[0] = bnd.z;
#pragma unroll
            for (int k = 0; k < K; ++k) { double4 v = row[k]; sX[k + 1] = v.x; sY[k + 1] = v.y; sM[k + 1] = v.z; }
        }
        accumulate_edge<K, SMALLTAB>(c, st, acc, sl, p, wl, ord, sX, sY, sM);
    }
    if (!(info & VR_LAST)) return false;
    commit_site<K>(c, st, acc, i, c.c_block + lane * K == 0, recvX, recvY, recvM, out_words);
    return true;
}

// per-lane constants of one column block
template <int K>
__device__ __forceinline__ void strip_init_lane(const StripCtx &c, LaneState<K> &st, int lane, const int *r_state,
                                                const float *r_elogw) {
    const double ninf = neg_inf();
    const int c0 = c.c_block + lane * K;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        int j = c0 + k;
        bool v = j < c.ly;
        st.X[k] = st.Y[k] = st.M[k] = ninf;
        st.extX[k] = (c.term && (j == 0 || j == c.ly - 1)) ? c.end_ext : c.ext;
        // column j >= 1 is entered by the chain edge (j-1 -> j), CSR position j-1
        st.wr[k] = (v && j >= 1) ? (double)r_elogw[j - 1] : 0.0;
        st.colbase[k] = (v && j >= 1) ? r_state[j] * c.fas : 0;
    }
    st.penY1 = (c.reduced && c0 == 0) ? 0.0 : c.open;  // Y move out of column j-1 == 0 (basic_alignment.h:494)
    st.bX = st.bY = st.bM = ninf;
}

// iterate_bwd_edges_for_end_corner (:1440-1552) with a single right edge (ly-1 -> stop): reads the last
// column from the boundary-column scratch.
__device__ __forceinline__ void strip_end_corner(const StripCtx &c, const double4 *lastcol, const float *r_elogw, DevResult *res) {
    const double ninf = neg_inf();
    double best = ninf;
    unsigned ptr = NO_MAT;
    const int kl0 = c.l_off[c.lx], kl1 = c.l_off[c.lx + 1];
    const double wr = (double)r_elogw[c.ly - 1];  // edge (ly-1 -> ly)
    for (int kl = kl0; kl < kl1; ++kl) {
        double4 v = lastcol[c.l_estart[kl]];
        double s = __dadd_rn(__dadd_rn(__dadd_rn(v.z, c.lng), (double)c.l_elogw[kl]), wr);
        if (s > best) { best = s; ptr = pack_ptr(M_MAT, kl - kl0, 0); }
        s = v.x;  // score_gap_close: + 0
        if (s > best) { best = s; ptr = pack_ptr(X_MAT, kl - kl0, 0); }
        if (kl == kl0) {
            s = lastcol[c.lx - 1].y;
            if (s > best) { best = s; ptr = pack_ptr(Y_MAT, 0, 0); }
        }
    }
    res->score = best;
    res->end_ptr = ptr;
    res->status = (best == ninf) ? JOB_NO_PATH : JOB_OK;
}

__device__ __forceinline__ void strip_make_ctx(StripCtx &c, const DevJob &J, const DevGraph &GL, const DevGraph &GR, const DevModel &m,
                                               const int *d_off, const int *d_estart, const float *d_elogw, const int4 *d_vrow,
                                               int K) {
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
    c.term = !(J.flags & FLAG_NO_TERMINAL_EDGES);
    c.reduced = (J.flags & FLAG_REDUCED) != 0;
    c.wr_zero = GR.zero_w != 0;
    c.lx = J.lx;
    c.ly = J.ly;
    c.W = 32 * K;
}

// One step of one lane.  `all_fast`, `any_weights`, `any_saved` are warp-uniform.  Returns true when the
// lane completed a site this step (pointer words valid, st holds the site's row).
template <int K, bool GENERAL, bool SMALLTAB>
__device__ __forceinline__ bool strip_lane_step(const StripCtx &c, LaneState<K> &st, LaneAcc<K> &acc, int lane, int4 vr,
                                                bool all_fast, bool any_weights, bool any_saved, double rX, double rY,
                                                double rM, unsigned short *w) {
    if (!GENERAL || all_fast) {
        const int i = vr.z, sl = vr.x & VR_STATE_MASK;
        const bool corner = (i == 0) && (c.c_block + lane * K == 0);  // row 0 is a fast row
        if (any_weights) {
            const double wl = (vr.x & (VR_ZERO_W | VR_NOEDGE)) ? 0.0 : (double)c.l_elogw[vr.y];
            fast_row<K, true, true, SMALLTAB>(c, st, i, sl, wl, corner, rX, rY, rM, w);
        } else {
            fast_row<K, false, false, SMALLTAB>(c, st, i, sl, 0.0, corner, rX, rY, rM, w);
        }
        return true;
    }
    return general_step<K, SMALLTAB>(c, st, acc, lane, vr, any_saved, rX, rY, rM, w);
}

#ifndef PG2_HOST_EMU
// keeps a per-job constant in a register: without the barrier ptxas re-derives the doubles from the
// float model parameters (F2F) inside the step loop whenever registers get tight
__device__ __forceinline__ void pin(double &v) { asm volatile("" : "+d"(v)); }

#ifndef PG2_STRIP_MINB_SIMPLE
#define PG2_STRIP_MINB_SIMPLE 4
#endif
#ifndef PG2_STRIP_MINB_GENERAL
#define PG2_STRIP_MINB_GENERAL 3
#endif
template <int K, bool GENERAL, bool SMALLTAB>
__global__ void __launch_bounds__(128, GENERAL ? PG2_STRIP_MINB_GENERAL : PG2_STRIP_MINB_SIMPLE)
strip_fill_kernel(int n_jobs, const DevJob *jobs, const int *job_ids, const DevGraph *graphs, const DevModel *models,
                  const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw, const int4 *d_vrow,
                  unsigned short *ptrs, DevResult *results, double4 *saved_all, long long saved_per_warp, double4 *bcol_all,
                  long long bcol_per_warp, int *queue) {
    __shared__ double2 s_tab[SMALLTAB ? 4 : 1][SMALLTAB ? STRIP_SMALL_FAS * STRIP_SMALL_FAS : 1];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double4 *saved = saved_all + (long long)warp_global * saved_per_warp;
    double4 *bcol0 = bcol_all + (long long)warp_global * bcol_per_warp * 2;
    double4 *bcol1 = bcol0 + bcol_per_warp;
    const double ninf = neg_inf();
    int tab_model = -1;

    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(queue, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= n_jobs) break;
        const int jid = job_ids[q];
        const DevJob &J = jobs[jid];
        DevResult *res = results + jid;
        if (res->status != JOB_OK) continue;
        const DevGraph GL = graphs[J.left], GR = graphs[J.right];
        const DevModel m = models[J.model];
        StripCtx c;
        strip_make_ctx(c, J, GL, GR, m, d_off, d_estart, d_elogw, d_vrow, K);
        pin(c.open); pin(c.ext); pin(c.end_ext); pin(c.lng);
        c.saved = saved;
        c.ptr = ptrs + J.cell_base;
        if (SMALLTAB) {
            if (tab_model != J.model) {
                __syncwarp();
                for (int e = lane; e < m.fas * m.fas; e += 32) {
                    double ls = (double)m.table[e];
                    s_tab[wib][e] = make_double2(__dadd_rn(c.lng2, ls), __dadd_rn(c.lng, ls));
                }
                __syncwarp();
                tab_model = J.model;
            }
            c.stab = s_tab[wib];
        }
        const int *r_state = d_state + GR.state_base;
        const float *r_elogw = d_elogw + GR.edge_base;
        const int n_blocks = (c.ly + c.W - 1) / c.W;
        constexpr int KS = (K + 1) & ~1;
        // GENERAL == false: the engine only sends jobs whose left graph is a plain chain with unit weights,
        // so every row is a fast row and the general body is not even compiled in
        const bool left_fast = !GENERAL || (GL.simple != 0 && GL.zero_w != 0);

        for (int b = 0; b < n_blocks; ++b) {
            c.c_block = b * c.W;
            c.first_block = (b == 0);
            c.bcol_prev = (b & 1) ? bcol0 : bcol1;
            c.bcol_cur = (b & 1) ? bcol1 : bcol0;
            LaneState<K> st;
            LaneAcc<K> acc;
            strip_init_lane<K>(c, st, lane, r_state, r_elogw);
#pragma unroll
            for (int k = 0; k < K; ++k) { pin(st.extX[k]); acc.nX[k] = acc.nM[k] = ninf; acc.pX[k] = acc.pM[k] = NO_MAT; }
            pin(st.penY1);
            // the lane that owns the block's last valid column feeds the boundary column: every row when
            // another block follows, else only the rows the end corner reads (predecessors of the stop site)
            const int last_col = min(c.c_block + c.W, c.ly) - 1;
            const int last_lane = (last_col - c.c_block) / K, last_k = (last_col - c.c_block) % K;
            const int bcol_need = (b + 1 < n_blocks) ? ~0 : VR_ENDPRED;
            // lane 0 of the first block has no left neighbour: adding -inf to whatever the shuffle delivers
            // (its own column) yields the -inf boundary on the FP64 pipe, no selects
            double lane0_mask = (lane == 0 && c.first_block) ? ninf : 0.0;
            pin(lane0_mask);
            unsigned short *out = c.ptr + ((long long)b * (c.nv + 31) * 32 + lane) * KS;
            const int n_steps = c.nv + 31;
            for (int t = 0; t < n_steps; ++t) {
                // hand the previous step's last column to the next lane
                double rX = __dadd_rn(__shfl_up_sync(0xffffffffu, st.X[K - 1], 1), lane0_mask);
                double rY = __dadd_rn(__shfl_up_sync(0xffffffffu, st.Y[K - 1], 1), lane0_mask);
                double rM = __dadd_rn(__shfl_up_sync(0xffffffffu, st.M[K - 1], 1), lane0_mask);
                const int v = t - lane;
                const bool active = (v >= 0 && v < c.nv);
                int4 vr = make_int4(VR_FAST | VR_ZERO_W, -1, 0, 0);
                if (active) vr = __ldg(c.l_vrow + v);
                bool all_fast = true, any_weights = !c.wr_zero, any_saved = false;
                if (GENERAL && !left_fast) {
                    all_fast = !__any_sync(0xffffffffu, (vr.x & VR_FAST) != VR_FAST);
                    any_weights = __any_sync(0xffffffffu, !(vr.x & (VR_ZERO_W | VR_NOEDGE))) || !c.wr_zero;
                    any_saved = __any_sync(0xffffffffu, active && !(vr.x & (VR_REG | VR_NOEDGE)));
                }
                if (active) {
                    if (lane == 0 && !c.first_block) { double4 bv = c.bcol_prev[vr.z]; rX = bv.x; rY = bv.y; rM = bv.z; }
                    unsigned short w[KS];
                    if (KS > K) w[KS - 1] = 0;
                    const bool done = strip_lane_step<K, GENERAL, SMALLTAB>(c, st, acc, lane, vr, all_fast, any_weights, any_saved,
                                                                            rX, rY, rM, w);
                    if (done) {
                        // coalesced pointer store: KS half-words per lane, step-major
                        unsigned *dst = reinterpret_cast<unsigned *>(out + (long long)t * 32 * KS);
#pragma unroll
                        for (int h = 0; h < KS / 2; ++h) dst[h] = (unsigned)w[2 * h] | ((unsigned)w[2 * h + 1] << 16);
                        // park rows that feed long-span edges
                        if (GENERAL) {
                            const int slot = (int)((unsigned)vr.x >> VR_SLOT_SHIFT) - 1;
                            if (slot >= 0) {
                                double4 *row = c.saved + (long long)slot * c.W + lane * K;
#pragma unroll
                                for (int k = 0; k < K; ++k) row[k] = make_double4(st.X[k], st.Y[k], st.M[k], 0.0);
                            }
                        }
                        if (lane == last_lane && (vr.x & bcol_need)) {
                            double vx = ninf, vy = ninf, vm = ninf;
#pragma unroll
                            for (int k = 0; k < K; ++k) if (k == last_k) { vx = st.X[k]; vy = st.Y[k]; vm = st.M[k]; }
                            c.bcol_cur[vr.z] = make_double4(vx, vy, vm, 0.0);
                        }
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
        if (lane == 0) {
            const double4 *lastcol = ((n_blocks - 1) & 1) ? bcol1 : bcol0;
            strip_end_corner(c, lastcol, r_elogw, res);
        }
        __syncwarp();
    }
}
#endif

// CPU test emulation of one warp (tests/emu): the same step bodies, lanes run one after the other inside
// a step with the shuffle replaced by a snapshot of the previous step.
template <int K>
static void strip_emulate_job(const DevJob &J, const DevGraph &GL, const DevGraph &GR, const DevModel &m, const int *d_state,
                              const int *d_off, const int *d_estart, const float *d_elogw, const int4 *d_vrow,
                              unsigned short *ptrs, DevResult *res, double4 *saved, double4 *bcol0, double4 *bcol1) {
#ifdef PG2_HOST_EMU
    const double ninf = neg_inf();
    StripCtx c;
    strip_make_ctx(c, J, GL, GR, m, d_off, d_estart, d_elogw, d_vrow, K);
    c.saved = saved;
    c.ptr = ptrs + J.cell_base;
    std::vector<double2> tab;
    const bool smalltab = m.fas <= STRIP_SMALL_FAS;
    if (smalltab) {
        tab.resize((size_t)m.fas * m.fas);
        for (int e = 0; e < m.fas * m.fas; ++e) {
            double ls = (double)m.table[e];
            tab[e] = make_double2(c.lng2 + ls, c.lng + ls);
        }
        c.stab = tab.data();
    }
    const int *r_state = d_state + GR.state_base;
    const float *r_elogw = d_elogw + GR.edge_base;
    const int n_blocks = (c.ly + c.W - 1) / c.W;
    const int KS = strip_ks(K);
    const bool left_fast = GL.simple != 0 && GL.zero_w != 0;
    for (int b = 0; b < n_blocks; ++b) {
        c.c_block = b * c.W;
        c.first_block = (b == 0);
        c.bcol_prev = (b & 1) ? bcol0 : bcol1;
        c.bcol_cur = (b & 1) ? bcol1 : bcol0;
        LaneState<K> st[32];
        LaneAcc<K> acc[32];
        for (int l = 0; l < 32; ++l) {
            strip_init_lane<K>(c, st[l], l, r_state, r_elogw);
            for (int k = 0; k < K; ++k) { acc[l].nX[k] = acc[l].nM[k] = ninf; acc[l].pX[k] = acc[l].pM[k] = NO_MAT; }
        }
        const int last_col = (c.c_block + c.W < c.ly ? c.c_block + c.W : c.ly) - 1;
        const int last_lane = (last_col - c.c_block) / K, last_k = (last_col - c.c_block) % K;
        const int bcol_need = (b + 1 < n_blocks) ? ~0 : VR_ENDPRED;
        for (int t = 0; t < c.nv + 31; ++t) {
            double sx[32], sy[32], sm[32];
            int4 vr[32];
            bool all_fast = true, any_weights = !c.wr_zero, any_saved = false;
            for (int l = 0; l < 32; ++l) {
                sx[l] = st[l].X[K - 1]; sy[l] = st[l].Y[K - 1]; sm[l] = st[l].M[K - 1];
                const int v = t - l;
                const bool active = v >= 0 && v < c.nv;
                vr[l] = active ? c.l_vrow[v] : make_int4(VR_FAST | VR_ZERO_W, -1, 0, 0);
                if (!left_fast) {
                    if ((vr[l].x & VR_FAST) != VR_FAST) all_fast = false;
                    if (!(vr[l].x & (VR_ZERO_W | VR_NOEDGE))) any_weights = true;
                    if (active && !(vr[l].x & (VR_REG | VR_NOEDGE))) any_saved = true;
                }
            }
            for (int l = 0; l < 32; ++l) {
                const int v = t - l;
                if (v < 0 || v >= c.nv) continue;
                double rX, rY, rM;
                if (l == 0) {
                    if (c.first_block) rX = rY = rM = ninf;
                    else { double4 bv = c.bcol_prev[vr[l].z]; rX = bv.x; rY = bv.y; rM = bv.z; }
                } else { rX = sx[l - 1]; rY = sy[l - 1]; rM = sm[l - 1]; }
                unsigned short w[8];
                bool done;
                if (smalltab) done = strip_lane_step<K, true, true>(c, st[l], acc[l], l, vr[l], all_fast, any_weights, any_saved, rX, rY, rM, w);
                else done = strip_lane_step<K, true, false>(c, st[l], acc[l], l, vr[l], all_fast, any_weights, any_saved, rX, rY, rM, w);
                if (!done) continue;
                unsigned short *out = c.ptr + (((long long)b * (c.nv + 31) + t) * 32 + l) * KS;
                for (int k = 0; k < K; ++k) out[k] = w[k];
                const int slot = (int)((unsigned)vr[l].x >> VR_SLOT_SHIFT) - 1;
                if (slot >= 0) {
                    double4 *row = c.saved + (long long)slot * c.W + l * K;
                    for (int k = 0; k < K; ++k) row[k] = make_double4(st[l].X[k], st[l].Y[k], st[l].M[k], 0.0);
                }
                if (l == last_lane && (vr[l].x & bcol_need))
                    c.bcol_cur[vr[l].z] = make_double4(st[l].X[last_k], st[l].Y[last_k], st[l].M[last_k], 0.0);
            }
        }
    }
    strip_end_corner(c, ((n_blocks - 1) & 1) ? bcol1 : bcol0, r_elogw, res);
#else
    (void)J; (void)GL; (void)GR; (void)m; (void)d_state; (void)d_off; (void)d_estart; (void)d_elogw; (void)d_vrow; (void)ptrs;
    (void)res; (void)saved; (void)bcol0; (void)bcol1;
#endif
}

// Launch one group of strip jobs that share the strip width K.  saved/bcol scratch is per resident warp.
int strip_warps_per_sm() { return 16; }

void launch_strip_fill(int K, bool general, bool smalltab, int n_jobs, const DevJob *jobs, const int *job_ids, const DevGraph *graphs, const DevModel *models,
                       const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw, const int4 *d_vrow,
                       unsigned short *ptrs, DevResult *results, double4 *saved_all, long long saved_per_warp, double4 *bcol_all,
                       long long bcol_per_warp, int *queue, int n_warps, cudaStream_t stream) {
    if (n_jobs <= 0) return;
#ifndef PG2_HOST_EMU
    cudaMemsetAsync(queue, 0, sizeof(int), stream);
    const int threads = 128;
    const int blocks = (n_warps * 32 + threads - 1) / threads;
#define PG2_STRIP_LAUNCH(KK, G, S)                                                                                            \
    strip_fill_kernel<KK, G, S><<<blocks, threads, 0, stream>>>(n_jobs, jobs, job_ids, graphs, models, d_state, d_off, d_estart, \
                                                                d_elogw, d_vrow, ptrs, results, saved_all, saved_per_warp,  \
                                                                bcol_all, bcol_per_warp, queue)
#define PG2_STRIP_CASE(KK)                                         \
    case KK:                                                       \
        if (general && smalltab) PG2_STRIP_LAUNCH(KK, true, true);  \
        else if (general) PG2_STRIP_LAUNCH(KK, true, false);        \
        else if (smalltab) PG2_STRIP_LAUNCH(KK, false, true);       \
        else PG2_STRIP_LAUNCH(KK, false, false);                    \
        break;
    switch (K) {
        PG2_STRIP_CASE(2) PG2_STRIP_CASE(3) PG2_STRIP_CASE(4) PG2_STRIP_CASE(5) PG2_STRIP_CASE(6) PG2_STRIP_CASE(8)
        default: break;
    }
#undef PG2_STRIP_CASE
#undef PG2_STRIP_LAUNCH
#else
    (void)queue; (void)n_warps; (void)stream; (void)general; (void)smalltab;
    for (int q = 0; q < n_jobs; ++q) {
        const int jid = job_ids[q];
        const DevJob &J = jobs[jid];
        DevResult *res = results + jid;
        if (res->status != JOB_OK) continue;
        double4 *bcol0 = bcol_all, *bcol1 = bcol_all + bcol_per_warp;
        switch (K) {
#define PG2_STRIP_CASE(KK) case KK: strip_emulate_job<KK>(J, graphs[J.left], graphs[J.right], models[J.model], d_state, d_off, d_estart, d_elogw, d_vrow, ptrs, res, saved_all, bcol0, bcol1); break;
            PG2_STRIP_CASE(2) PG2_STRIP_CASE(3) PG2_STRIP_CASE(4) PG2_STRIP_CASE(5) PG2_STRIP_CASE(6) PG2_STRIP_CASE(8)
#undef PG2_STRIP_CASE
            default: break;
        }
    }
#endif
}

}  // namespace pg2
