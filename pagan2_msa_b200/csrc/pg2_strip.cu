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
// Arithmetic follows the reference candidate by candidate (src/main/viterbi_alignment.cpp:856-971,
// 1328-1436, 2029-2219): same order, same FP64 association, strict '>' (first candidate wins ties).
// "+ 0.0" terms (log_gap_close == 0) are dropped: x + 0.0 == x for every x the DP can produce.
#include "pg2_device.cuh"
#include "pg2_strip_geom.cuh"

namespace pg2 {

bool strip_eligible(int lx, int ly, bool banded, int l_simple, int r_simple, int l_maxdeg, int r_maxdeg, int fas) {
    (void)l_simple; (void)r_maxdeg; (void)fas;
    if (banded) return false;
    if (!r_simple) return false;
    if (l_maxdeg > STRIP_MAX_LEFT_INDEG) return false;
    if (lx < 1 || ly < 1) return false;
    return true;
}

template <int K> struct LaneState {
    double X[K], Y[K], M[K];     // own strip, row this lane handled in the previous step
    double bX, bY, bM;           // left neighbour column (c0-1) of that same row
    double extX[K];              // log_gap_ext / log_gap_end_ext per column (X moves, :864-868)
    double penY[K];              // gap-open penalty for Y moves from column j-1 (0 when j-1 == 0 and reduced)
    double wr[K];                // log weight of the column edge into site j
    int colbase[K];              // state_r[j] * fas
    bool valid[K];
};

struct StripCtx {
    // row graph (left)
    const int *l_state, *l_off, *l_estart, *l_slot;
    const float *l_elogw;
    // model
    const float *table;
    int fas;
    double open, ext, end_ext, lng, lng2;
    bool term, reduced;
    int lx, ly;
    // per-warp scratch
    double4 *saved;   // [n_slots][W]
    double4 *bcol_prev, *bcol_cur;  // [lx]
    unsigned short *ptr;  // job's pointer region
    int W, c_block;   // block width, first column of the block
    bool first_block;
};

// One row of one lane's strip.  recv* = (X,Y,M) of column c0-1 on this row (from lane l-1, or the
// previous block's boundary column for lane 0).  On return st holds this row.
template <int K>
__device__ __forceinline__ void strip_row(const StripCtx &c, LaneState<K> &st, int lane, int i, double recvX, double recvY,
                                          double recvM, unsigned short *out_words) {
    const double ninf = neg_inf();
    const int c0 = c.c_block + lane * K;
    double nX[K], nM[K], nY[K];
    unsigned pX[K], pM[K], pY[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { nX[k] = ninf; nM[k] = ninf; pX[k] = NO_MAT; pM[k] = NO_MAT; }

    const int k0 = c.l_off[i], k1 = c.l_off[i + 1];
    const int sl = (i > 0) ? c.l_state[i] : 0;
    // substitution terms of this row (only rows i >= 1 have candidates)
    double mlog[K], xlog[K];
    if (k1 > k0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double ls = st.valid[k] ? (double)__ldg(c.table + sl + st.colbase[k]) : 0.0;
            mlog[k] = __dadd_rn(c.lng2, ls);   // 2*log_non_gap + log_score  (:1364)
            xlog[k] = __dadd_rn(c.lng, ls);    // close(=0) + log_non_gap + log_score (:1366-1367)
        }
    }
    for (int e = k0; e < k1; ++e) {
        const int p = c.l_estart[e];
        const double wl = (double)c.l_elogw[e];
        const unsigned ord = (unsigned)(e - k0) << 2;
        // source row p: columns c0-1 .. c0+K-1
        double sX[K + 1], sY[K + 1], sM[K + 1];
        if (p == i - 1) {
            sX[0] = st.bX; sY[0] = st.bY; sM[0] = st.bM;
#pragma unroll
            for (int k = 0; k < K; ++k) { sX[k + 1] = st.X[k]; sY[k + 1] = st.Y[k]; sM[k + 1] = st.M[k]; }
        } else {
            const int slot = c.l_slot[p];
            const double4 *row = c.saved + (long long)slot * c.W + lane * K;
            double4 b;
            if (lane > 0) b = row[-1];
            else if (!c.first_block) b = c.bcol_prev[p];
            else b = make_double4(ninf, ninf, ninf, 0.0);
            sX[0] = b.x; sY[0] = b.y; sM[0] = b.z;
#pragma unroll
            for (int k = 0; k < K; ++k) { double4 v = row[k]; sX[k + 1] = v.x; sY[k + 1] = v.y; sM[k + 1] = v.z; }
        }
        const double pen = (c.reduced && p == 0) ? 0.0 : c.open;  // get_log_gap_open_penalty (basic_alignment.h:490-513)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            // X(i,j) from (p,j): ext, double, open  (:2116-2211)
            double s = __dadd_rn(sX[k + 1], st.extX[k]);
            if (s > nX[k]) { nX[k] = s; pX[k] = X_MAT | ord; }
            s = __dadd_rn(sY[k + 1], c.open);
            if (s > nX[k]) { nX[k] = s; pX[k] = Y_MAT | ord; }
            s = __dadd_rn(__dadd_rn(sM[k + 1], c.lng), pen);
            if (s > nX[k]) { nX[k] = s; pX[k] = M_MAT | ord; }
            // M(i,j) from (p,j-1): M, X, Y  (:2029-2112)
            s = __dadd_rn(__dadd_rn(__dadd_rn(sM[k], mlog[k]), wl), st.wr[k]);
            if (s > nM[k]) { nM[k] = s; pM[k] = M_MAT | ord; }
            s = __dadd_rn(__dadd_rn(__dadd_rn(sX[k], xlog[k]), wl), st.wr[k]);
            if (s > nM[k]) { nM[k] = s; pM[k] = X_MAT | ord; }
            s = __dadd_rn(__dadd_rn(__dadd_rn(sY[k], xlog[k]), wl), st.wr[k]);
            if (s > nM[k]) { nM[k] = s; pM[k] = Y_MAT | ord; }
        }
    }
    // column 0 has no M (compute_fwd_scores :956-969); cell (0,0) is the start corner (:725-733)
    if (c0 == 0) {
        nM[0] = (i == 0) ? 0.0 : ninf;
        pM[0] = NO_MAT;
    }
    // Y(i,j) from (i,j-1): ext, double, open -- sequential along the row
    const double extY = (c.term && (i == 0 || i == c.lx - 1)) ? c.end_ext : c.ext;
    double lX = recvX, lY = recvY, lM = recvM;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double best = ninf;
        unsigned ptr = NO_MAT;
        double s = __dadd_rn(lY, extY);
        if (s > best) { best = s; ptr = Y_MAT; }
        s = __dadd_rn(lX, c.open);
        if (s > best) { best = s; ptr = X_MAT; }
        s = __dadd_rn(__dadd_rn(lM, c.lng), st.penY[k]);
        if (s > best) { best = s; ptr = M_MAT; }
        nY[k] = best;
        pY[k] = ptr;
        lX = nX[k]; lY = best; lM = nM[k];
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        st.X[k] = nX[k]; st.Y[k] = nY[k]; st.M[k] = nM[k];
        out_words[k] = (unsigned short)strip_word(pX[k], pY[k], pM[k]);
    }
    st.bX = recvX; st.bY = recvY; st.bM = recvM;
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
        st.valid[k] = v;
        st.X[k] = st.Y[k] = st.M[k] = ninf;
        st.extX[k] = (c.term && (j == 0 || j == c.ly - 1)) ? c.end_ext : c.ext;
        st.penY[k] = (c.reduced && j == 1) ? 0.0 : c.open;
        // column j >= 1 is entered by the chain edge (j-1 -> j), CSR position j-1
        st.wr[k] = (v && j >= 1) ? (double)r_elogw[j - 1] : 0.0;
        st.colbase[k] = (v && j >= 1) ? r_state[j] * c.fas : 0;
    }
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

__device__ __forceinline__ void strip_make_ctx(StripCtx &c, const DevJob &J, const DevGraph &GL, const DevModel &m, const int *d_state,
                                               const int *d_off, const int *d_estart, const float *d_elogw, const int *d_slot,
                                               int K) {
    c.l_state = d_state + GL.state_base;
    c.l_off = d_off + GL.off_base;
    c.l_estart = d_estart + GL.edge_base;
    c.l_elogw = d_elogw + GL.edge_base;
    c.l_slot = d_slot + GL.state_base;
    c.table = m.table;
    c.fas = m.fas;
    c.open = (double)m.open;
    c.ext = (double)m.ext;
    c.end_ext = (double)m.end_ext;
    c.lng = (double)m.lng;
    c.lng2 = (double)__fmul_rn(2.0f, m.lng);
    c.term = !(J.flags & FLAG_NO_TERMINAL_EDGES);
    c.reduced = (J.flags & FLAG_REDUCED) != 0;
    c.lx = J.lx;
    c.ly = J.ly;
    c.W = 32 * K;
}

#ifndef PG2_HOST_EMU
template <int K>
__global__ void __launch_bounds__(128)
strip_fill_kernel(int n_jobs, const DevJob *jobs, const int *job_ids, const DevGraph *graphs, const DevModel *models,
                  const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw, const int *d_slot,
                  unsigned short *ptrs, DevResult *results, double4 *saved_all, long long saved_per_warp, double4 *bcol_all,
                  long long bcol_per_warp, int *queue) {
    const int lane = threadIdx.x & 31;
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double4 *saved = saved_all + (long long)warp_global * saved_per_warp;
    double4 *bcol0 = bcol_all + (long long)warp_global * bcol_per_warp * 2;
    double4 *bcol1 = bcol0 + bcol_per_warp;
    const double ninf = neg_inf();

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
        strip_make_ctx(c, J, GL, m, d_state, d_off, d_estart, d_elogw, d_slot, K);
        c.saved = saved;
        c.ptr = ptrs + J.cell_base;
        const int *r_state = d_state + GR.state_base;
        const float *r_elogw = d_elogw + GR.edge_base;
        const int n_blocks = (c.ly + c.W - 1) / c.W;
        constexpr int KS = (K + 1) & ~1;

        for (int b = 0; b < n_blocks; ++b) {
            c.c_block = b * c.W;
            c.first_block = (b == 0);
            c.bcol_prev = (b & 1) ? bcol0 : bcol1;
            c.bcol_cur = (b & 1) ? bcol1 : bcol0;
            LaneState<K> st;
            strip_init_lane<K>(c, st, lane, r_state, r_elogw);
            // lane that owns the block's last valid column writes the boundary column
            const int last_col = min(c.c_block + c.W, c.ly) - 1;
            const int last_lane = (last_col - c.c_block) / K, last_k = (last_col - c.c_block) % K;
            unsigned short *out = c.ptr + ((long long)b * (c.lx + 31) * 32 + lane) * KS;
            const int n_steps = c.lx + 31;
            for (int t = 0; t < n_steps; ++t) {
                // hand the previous step's last column to the next lane
                double rX = __shfl_up_sync(0xffffffffu, st.X[K - 1], 1);
                double rY = __shfl_up_sync(0xffffffffu, st.Y[K - 1], 1);
                double rM = __shfl_up_sync(0xffffffffu, st.M[K - 1], 1);
                const int i = t - lane;
                if (i >= 0 && i < c.lx) {
                    if (lane == 0) {
                        if (c.first_block) { rX = rY = rM = ninf; }
                        else { double4 v = c.bcol_prev[i]; rX = v.x; rY = v.y; rM = v.z; }
                    }
                    unsigned short w[KS];
                    if (KS > K) w[KS - 1] = 0;
                    strip_row<K>(c, st, lane, i, rX, rY, rM, w);
                    // coalesced pointer store: KS half-words per lane, step-major
                    unsigned *dst = reinterpret_cast<unsigned *>(out + (long long)t * 32 * KS);
#pragma unroll
                    for (int h = 0; h < KS / 2; ++h) dst[h] = (unsigned)w[2 * h] | ((unsigned)w[2 * h + 1] << 16);
                    // park rows that feed long-span edges
                    const int slot = c.l_slot[i];
                    if (slot >= 0) {
                        double4 *row = c.saved + (long long)slot * c.W + lane * K;
#pragma unroll
                        for (int k = 0; k < K; ++k) row[k] = make_double4(st.X[k], st.Y[k], st.M[k], 0.0);
                    }
                    if (lane == last_lane) {
                        double vx = ninf, vy = ninf, vm = ninf;
#pragma unroll
                        for (int k = 0; k < K; ++k) if (k == last_k) { vx = st.X[k]; vy = st.Y[k]; vm = st.M[k]; }
                        c.bcol_cur[i] = make_double4(vx, vy, vm, 0.0);
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

// CPU test emulation of one warp (tests/emu): the same strip_row / init / end-corner bodies, lanes run
// one after the other inside a step with the shuffle replaced by a snapshot of the previous step.
template <int K>
static void strip_emulate_job(const DevJob &J, const DevGraph &GL, const DevGraph &GR, const DevModel &m, const int *d_state,
                              const int *d_off, const int *d_estart, const float *d_elogw, const int *d_slot,
                              unsigned short *ptrs, DevResult *res, double4 *saved, double4 *bcol0, double4 *bcol1) {
#ifdef PG2_HOST_EMU
    const double ninf = neg_inf();
    StripCtx c;
    strip_make_ctx(c, J, GL, m, d_state, d_off, d_estart, d_elogw, d_slot, K);
    c.saved = saved;
    c.ptr = ptrs + J.cell_base;
    const int *r_state = d_state + GR.state_base;
    const float *r_elogw = d_elogw + GR.edge_base;
    const int n_blocks = (c.ly + c.W - 1) / c.W;
    const int KS = strip_ks(K);
    for (int b = 0; b < n_blocks; ++b) {
        c.c_block = b * c.W;
        c.first_block = (b == 0);
        c.bcol_prev = (b & 1) ? bcol0 : bcol1;
        c.bcol_cur = (b & 1) ? bcol1 : bcol0;
        LaneState<K> st[32];
        for (int l = 0; l < 32; ++l) strip_init_lane<K>(c, st[l], l, r_state, r_elogw);
        const int last_col = (c.c_block + c.W < c.ly ? c.c_block + c.W : c.ly) - 1;
        const int last_lane = (last_col - c.c_block) / K, last_k = (last_col - c.c_block) % K;
        for (int t = 0; t < c.lx + 31; ++t) {
            double sx[32], sy[32], sm[32];
            for (int l = 0; l < 32; ++l) { sx[l] = st[l].X[K - 1]; sy[l] = st[l].Y[K - 1]; sm[l] = st[l].M[K - 1]; }
            for (int l = 0; l < 32; ++l) {
                const int i = t - l;
                if (i < 0 || i >= c.lx) continue;
                double rX, rY, rM;
                if (l == 0) {
                    if (c.first_block) rX = rY = rM = ninf;
                    else { double4 v = c.bcol_prev[i]; rX = v.x; rY = v.y; rM = v.z; }
                } else { rX = sx[l - 1]; rY = sy[l - 1]; rM = sm[l - 1]; }
                unsigned short w[8];
                strip_row<K>(c, st[l], l, i, rX, rY, rM, w);
                unsigned short *out = c.ptr + (((long long)b * (c.lx + 31) + t) * 32 + l) * KS;
                for (int k = 0; k < K; ++k) out[k] = w[k];
                const int slot = c.l_slot[i];
                if (slot >= 0) {
                    double4 *row = c.saved + (long long)slot * c.W + l * K;
                    for (int k = 0; k < K; ++k) row[k] = make_double4(st[l].X[k], st[l].Y[k], st[l].M[k], 0.0);
                }
                if (l == last_lane) c.bcol_cur[i] = make_double4(st[l].X[last_k], st[l].Y[last_k], st[l].M[last_k], 0.0);
            }
        }
    }
    strip_end_corner(c, ((n_blocks - 1) & 1) ? bcol1 : bcol0, r_elogw, res);
#else
    (void)J; (void)GL; (void)GR; (void)m; (void)d_state; (void)d_off; (void)d_estart; (void)d_elogw; (void)d_slot; (void)ptrs;
    (void)res; (void)saved; (void)bcol0; (void)bcol1;
#endif
}

// Launch one group of strip jobs that share the strip width K.  saved/bcol scratch is per resident warp.
int strip_warps_per_sm() { return 16; }

void launch_strip_fill(int K, int n_jobs, const DevJob *jobs, const int *job_ids, const DevGraph *graphs, const DevModel *models,
                       const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw, const int *d_slot,
                       unsigned short *ptrs, DevResult *results, double4 *saved_all, long long saved_per_warp, double4 *bcol_all,
                       long long bcol_per_warp, int *queue, int n_warps, cudaStream_t stream) {
    if (n_jobs <= 0) return;
#ifndef PG2_HOST_EMU
    cudaMemsetAsync(queue, 0, sizeof(int), stream);
    const int threads = 128;
    const int blocks = (n_warps * 32 + threads - 1) / threads;
#define PG2_STRIP_CASE(KK)                                                                                                   \
    case KK:                                                                                                                 \
        strip_fill_kernel<KK><<<blocks, threads, 0, stream>>>(n_jobs, jobs, job_ids, graphs, models, d_state, d_off, d_estart, \
                                                              d_elogw, d_slot, ptrs, results, saved_all, saved_per_warp,    \
                                                              bcol_all, bcol_per_warp, queue);                               \
        break;
    switch (K) {
        PG2_STRIP_CASE(2) PG2_STRIP_CASE(3) PG2_STRIP_CASE(4) PG2_STRIP_CASE(5) PG2_STRIP_CASE(6) PG2_STRIP_CASE(8)
        default: break;
    }
#undef PG2_STRIP_CASE
#else
    (void)queue; (void)n_warps; (void)stream;
    for (int q = 0; q < n_jobs; ++q) {
        const int jid = job_ids[q];
        const DevJob &J = jobs[jid];
        DevResult *res = results + jid;
        if (res->status != JOB_OK) continue;
        double4 *bcol0 = bcol_all, *bcol1 = bcol_all + bcol_per_warp;
        switch (K) {
#define PG2_STRIP_CASE(KK) case KK: strip_emulate_job<KK>(J, graphs[J.left], graphs[J.right], models[J.model], d_state, d_off, d_estart, d_elogw, d_slot, ptrs, res, saved_all, bcol0, bcol1); break;
            PG2_STRIP_CASE(2) PG2_STRIP_CASE(3) PG2_STRIP_CASE(4) PG2_STRIP_CASE(5) PG2_STRIP_CASE(6) PG2_STRIP_CASE(8)
#undef PG2_STRIP_CASE
            default: break;
        }
    }
#endif
}

}  // namespace pg2
