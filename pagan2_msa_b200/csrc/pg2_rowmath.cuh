// pg2_rowmath.cuh -- the per-row arithmetic shared by the register-strip fill kernels (pg2_strip.cu: warp per
// alignment; pg2_lanes.cu: lane per alignment).  A "strip" is K consecutive DP columns of one row, held in
// registers; both kernels sweep a strip down the rows of the LEFT graph and differ only in where the strip's
// left-neighbour column comes from (a shuffle from the previous lane / a per-warp boundary column).
//
// Arithmetic follows the reference candidate by candidate (src/main/viterbi_alignment.cpp:856-971,
// 1328-1436, 2029-2219): same order, same FP64 association, strict '>' (first candidate wins ties).
#pragma once
#include "pg2_device.cuh"
#include "pg2_strip_geom.cuh"

namespace pg2 {

template <int K> struct LaneState {
    double X[K], Y[K], M[K];     // own strip, row this lane handled in the previous step
    double bX, bY, bM;           // left neighbour column (c0-1) of that same row
    double extX[K];              // log_gap_ext / log_gap_end_ext per column (X moves, :864-868)
    double wr[K];                // log weight of the column edge into site j
    double penY1;                // gap-open penalty of the Y move out of column 0 (k == 1 of lane 0, block 0)
    int colbase[K];              // state_r[j] * fas (0 for padding columns)
};

struct StripCtx {
    // row graph (left)
    const int4 *l_vrow;
    const int *l_off, *l_estart;
    const float *l_elogw;
    int nv;                      // virtual rows
    // model
    const float *table;          // global float table (any alphabet)
    const double2 *stab;         // shared {2*lng + ls, lng + ls} table, or nullptr when the alphabet is too big
    int fas;
    double open, ext, end_ext, lng, lng2;
    bool term, reduced, wr_zero;
    int lx, ly;
    // per-warp scratch
    double4 *saved;   // [n_slots][W]
    double4 *bcol_prev, *bcol_cur;  // [lx]
    unsigned short *ptr;  // job's pointer region
    int W, c_block;   // block width, first column of the block
    bool first_block;
};

// {m_log, x_log} of one cell: 2*log_non_gap + log_score and log_non_gap + log_score (:1363-1367).
// SMALLTAB: both terms come precomputed from the per-warp shared table (alphabets up to STRIP_SMALL_FAS).
template <bool SMALLTAB>
__device__ __forceinline__ void subst_terms(const StripCtx &c, int sl, int colbase, double &mlog, double &xlog) {
    if (SMALLTAB) {
        double2 v = c.stab[sl + colbase];
        mlog = v.x;
        xlog = v.y;
    } else {
        double ls = (double)__ldg(c.table + sl + colbase);
        mlog = __dadd_rn(c.lng2, ls);
        xlog = __dadd_rn(c.lng, ls);
    }
}

// Fast-row pointer half-word: bit 14 set, bits 0-5 the raw comparison outcomes
//   bit0/1 X: (double > ext), (open > max of the first two)      candidates in order X, Y, M
//   bit2/3 Y: (open > double), (that winner > ext)                candidates in order Y, X, M
//   bit4/5 M: (X > M), (Y > max of the first two)                 candidates in order M, X, Y
// decoded by strip_fast_ptr() in pg2_strip_geom.cuh.  A cell whose candidates are all -inf gets
// arbitrary bits: it cannot lie on the Viterbi path.
//
// Row with a single backward edge from row i-1: in-place update of the lane's strip.
// WL / WR: add the left edge's log weight `wl` / the column edges' log weights st.wr[k].  A dropped "+ 0.0" is
// exact: x + 0.0 == x for every x the DP produces (no -0.0 arises).
template <int K, bool WL, bool WR, bool SMALLTAB>
__device__ __forceinline__ void fast_row(const StripCtx &c, LaneState<K> &st, int i, int sl, double wl, bool corner, double recvX,
                                         double recvY, double recvM, unsigned short *out_words) {
    const double pen = (c.reduced && i == 1) ? 0.0 : c.open;  // open penalty out of row p = i-1 (basic_alignment.h:490-513)
    unsigned bits[K];
    // descending k: X(i,j) reads (i-1,j), M(i,j) reads (i-1,j-1); both still hold row i-1
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
        const double qX = k ? st.X[k - 1] : st.bX, qY = k ? st.Y[k - 1] : st.bY, qM = k ? st.M[k - 1] : st.bM;
        // X: ext, double, open (:2116-2211)
        double a = __dadd_rn(st.X[k], st.extX[k]);
        double b = __dadd_rn(st.Y[k], c.open);
        double d = __dadd_rn(__dadd_rn(st.M[k], c.lng), pen);
        bool p1 = b > a;
        double t = p1 ? b : a;
        bool p2 = d > t;
        const double nx = p2 ? d : t;
        unsigned w = (p1 ? 1u : 0u) | (p2 ? 2u : 0u);
        // M: from M, X, Y (:2029-2112)
        double mlog, xlog;
        subst_terms<SMALLTAB>(c, sl, st.colbase[k], mlog, xlog);
        a = __dadd_rn(qM, mlog);
        b = __dadd_rn(qX, xlog);
        d = __dadd_rn(qY, xlog);
        if (WL) { a = __dadd_rn(a, wl); b = __dadd_rn(b, wl); d = __dadd_rn(d, wl); }
        if (WR) { a = __dadd_rn(a, st.wr[k]); b = __dadd_rn(b, st.wr[k]); d = __dadd_rn(d, st.wr[k]); }
        p1 = b > a;
        t = p1 ? b : a;
        p2 = d > t;
        const double nm = p2 ? d : t;
        w |= (p1 ? 16u : 0u) | (p2 ? 32u : 0u);
        bits[k] = w;
        st.X[k] = nx;
        st.M[k] = nm;
    }
    // Column 0 needs no special case for i > 0: its M sources are the -inf boundary, so M(i,0) = -inf falls
    // out of the arithmetic.  Row 0 has no edges but its sources are the -inf initial strip, so X(0,j) =
    // M(0,j) = -inf fall out as well; only the start corner M(0,0) = 0 (:725-733) is planted here.
    if (corner) st.M[0] = 0.0;
    // Y(i,j) from (i,j-1): ext, double, open.  The two chain-independent candidates are folded first;
    // (g > a ? g : a) with g = first-wins(double, open) equals the sequential first-wins over all three.
    const double extY = (c.term && (i == 0 || i == c.lx - 1)) ? c.end_ext : c.ext;
    double lX = recvX, lY = recvY, lM = recvM;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double penY = (k == 1) ? st.penY1 : c.open;
        const double b = __dadd_rn(lX, c.open);
        const double d = __dadd_rn(__dadd_rn(lM, c.lng), penY);
        const bool p2 = d > b;
        const double g = p2 ? d : b;
        const double a = __dadd_rn(lY, extY);
        const bool p1 = g > a;
        const double ny = p1 ? g : a;
        out_words[k] = (unsigned short)(bits[k] | (p2 ? 4u : 0u) | (p1 ? 8u : 0u) | 0x4000u);
        lX = st.X[k];
        lM = st.M[k];
        lY = ny;
        st.Y[k] = ny;
    }
    st.bX = recvX; st.bY = recvY; st.bM = recvM;
}

// Per-lane accumulators of a site whose edges are spread over several virtual rows.
template <int K> struct LaneAcc {
    double nX[K], nM[K];
    unsigned pX[K], pM[K];
};

// candidates of ONE backward edge into the accumulators; (sX,sY,sM)[0] is column c0-1 of the source row,
// [k+1] column c0+k
template <int K, bool SMALLTAB, bool WL = true, bool WR = true>
__device__ __forceinline__ void accumulate_edge(const StripCtx &c, const LaneState<K> &st, LaneAcc<K> &acc, int sl, int p,
                                                double wl, unsigned ord, const double *sX, const double *sY, const double *sM) {
    const double pen = (c.reduced && p == 0) ? 0.0 : c.open;  // get_log_gap_open_penalty (basic_alignment.h:490-513)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = __dadd_rn(sX[k + 1], st.extX[k]);                  // X: ext, double, open (:2116-2211)
        if (s > acc.nX[k]) { acc.nX[k] = s; acc.pX[k] = X_MAT | ord; }
        s = __dadd_rn(sY[k + 1], c.open);
        if (s > acc.nX[k]) { acc.nX[k] = s; acc.pX[k] = Y_MAT | ord; }
        s = __dadd_rn(__dadd_rn(sM[k + 1], c.lng), pen);
        if (s > acc.nX[k]) { acc.nX[k] = s; acc.pX[k] = M_MAT | ord; }
        double mlog, xlog;
        subst_terms<SMALLTAB>(c, sl, st.colbase[k], mlog, xlog);
        s = __dadd_rn(sM[k], mlog);  // M: from M, X, Y (:2029-2112)
        if (WL) s = __dadd_rn(s, wl);
        if (WR) s = __dadd_rn(s, st.wr[k]);
        if (s > acc.nM[k]) { acc.nM[k] = s; acc.pM[k] = M_MAT | ord; }
        s = __dadd_rn(sX[k], xlog);
        if (WL) s = __dadd_rn(s, wl);
        if (WR) s = __dadd_rn(s, st.wr[k]);
        if (s > acc.nM[k]) { acc.nM[k] = s; acc.pM[k] = X_MAT | ord; }
        s = __dadd_rn(sY[k], xlog);
        if (WL) s = __dadd_rn(s, wl);
        if (WR) s = __dadd_rn(s, st.wr[k]);
        if (s > acc.nM[k]) { acc.nM[k] = s; acc.pM[k] = Y_MAT | ord; }
    }
}

// Last virtual row of a site: the accumulators hold X(i,.) and M(i,.); run the Y chain along the strip
// (Y(i,j) from (i,j-1): ext, double, open), emit the general-form pointer words and make row i the lane's
// current row.  `col0`: the strip starts at DP column 0, which has no M ((0,0) is the start corner,
// :725-733, :956-969).
template <int K>
__device__ __forceinline__ void commit_site(const StripCtx &c, LaneState<K> &st, LaneAcc<K> &acc, int i, bool col0, double recvX,
                                            double recvY, double recvM, unsigned short *out_words) {
    const double ninf = neg_inf();
    if (col0) {
        acc.nM[0] = (i == 0) ? 0.0 : ninf;
        acc.pM[0] = NO_MAT;
    }
    const double extY = (c.term && (i == 0 || i == c.lx - 1)) ? c.end_ext : c.ext;
    double lX = recvX, lY = recvY, lM = recvM;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double penY = (k == 1) ? st.penY1 : c.open;
        double best = ninf;
        unsigned ptr = NO_MAT;
        double s = __dadd_rn(lY, extY);
        if (s > best) { best = s; ptr = Y_MAT; }
        s = __dadd_rn(lX, c.open);
        if (s > best) { best = s; ptr = X_MAT; }
        s = __dadd_rn(__dadd_rn(lM, c.lng), penY);
        if (s > best) { best = s; ptr = M_MAT; }
        out_words[k] = (unsigned short)strip_word(acc.pX[k], ptr, acc.pM[k]);
        lX = acc.nX[k]; lY = best; lM = acc.nM[k];
        st.X[k] = acc.nX[k]; st.Y[k] = best; st.M[k] = acc.nM[k];
    }
    st.bX = recvX; st.bY = recvY; st.bM = recvM;
}

}  // namespace pg2
