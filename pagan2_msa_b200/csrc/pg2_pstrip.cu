// pg2_pstrip.cu -- pipelined-strip fill kernel: one CTA per alignment, the column blocks of the alignment pipelined over
// the warps of the CTA.  General graphs on BOTH sides, with or without an anchor band.
//
// This is the kernel of the guide-tree waves (node.cpp:240-264: few, large alignments; from the second wave on both
// children are ancestor graphs), of the pileup (reads_aligner.cpp: one growing root against one 454 read graph at a
// time) and of anchored alignments (Tunnel_matrix bands, utils/tunnel_matrix.h:45-344).  Geometry, column program and
// block table: pg2_pstrip_geom.cuh.
//
// A block of <= 32*K columns is swept by one warp: lane l owns K consecutive columns and keeps X / Y / M of its strip
// for the row it handled last in registers; the lanes are skewed by one virtual row (lane l is on virtual row t - l at
// step t), so the strip's last column is handed to lane l + 1 by shuffles.  The row graph is the left graph's row
// program (pg2_strip_geom.cuh: one virtual row per backward edge; rows that start long-span edges are parked in a
// per-warp scratch).  Column j of the right graph is PLAIN when its only backward edge comes from column j - 1 (its
// sources are then the neighbouring registers) and GENERAL otherwise: a general column walks its backward edges and
// reads the cells (i, pr), (i - 1, pr) from a shared-memory history of the block's PARKED columns, (pl, pr) of a
// long-span left edge from the parked row.
//
// The warp of block b + 1 follows the warp of block b: the last column of block b goes to a boundary-column ring in
// global memory (one entry per virtual row), a progress counter in shared memory says how far it is valid, and lane 0
// of block b + 1 fetches its entries a few steps ahead.  No CTA-wide barrier runs inside an alignment.
//
// Arithmetic follows the reference candidate by candidate (src/main/viterbi_alignment.cpp:856-971, 1328-1436,
// 2029-2219): same order, same FP64 association, strict '>' (first candidate wins ties).  "+ 0.0" terms are dropped
// (exact: the DP never produces -0.0).
#include "pg2_device.cuh"
#include "pg2_strip_geom.cuh"
#include "pg2_pstrip_geom.cuh"
#ifdef PG2_HOST_EMU
#include <vector>
#else
#include <cooperative_groups.h>
#endif

namespace pg2 {

// warp-uniform constants of one job and of the block being swept
struct PsCtx {
    // row graph (left)
    const int4 *l_vrow;
    const int *l_off, *l_estart;
    const float *l_elogw;
    int nv;
    // column graph (right)
    const int *r_state, *r_off, *r_estart, *r_einfo, *colinfo;
    const float *r_elogw;
    // model
    const float *table;
    const double2 *stab;
    int fas;
    double open, ext, end_ext, lng, lng2;
    bool term, reduced, banded, weights;
    int lx, ly;
    const int *blo, *bhi;
    // block
    int c0, c1, v0, v1, i0;
    // scratch
    double4 *saved;      // [n_slots][1 + 32*K]: entry 0 is column c0 - 1
    double *hist;        // shared: [PS_HIST][park_cap][3]
    int park_cap;        // history slots per block (the launch group's largest block need)
    double4 *endstore;   // [PS_MAX_END][lx]
    int saved_stride;    // 1 + 32*K
    // row ring (ps_step_ring): the last rr rows of the block in shared memory, [rr + 1][X,Y,M][K][33] doubles; entry l + 1 of a
    // [k] line is lane l's column k, entry 0 of line K - 1 the column left of the block; row rr is the "no row" of -inf
    double *rowring;
    int rr;              // ring rows (a power of two), 0: the kernel runs ps_step (parked rows in global memory, column history)
};

// the three scalars every candidate needs, in registers (everything else of PsCtx is read from shared memory on demand)
struct PsHot {
    double open, ext, lng;
    unsigned flags;      // PH_*
    int lx1;             // lx - 1
    unsigned stab_s;     // shared-memory address of the {2 lng + ls, lng + ls} table (SMALLTAB)
};
constexpr unsigned PH_TERM = 1u, PH_REDUCED = 2u, PH_BANDED = 4u, PH_WEIGHTS = 8u;

template <int K> struct PsLane {
    double X[K], Y[K], M[K];   // own strip, row handled last
    double Mo[K];              // (M + log_non_gap) + log_gap_open: what a gap move out of the cell starts from.  The reference
                               // forms it twice per cell (X move :2190-2211, Y move); the reduced terminal penalty
                               // (basic_alignment.h:490-513) only ever meets a finite M at the start corner, where it is planted
    double bX, bY, bM;         // column to the left of the strip, same row
    double extX[K];            // X-extension term per column (:864-868)
    double wr[K];              // log weight of the edge into a plain column
    int colbase[K];            // state_r[j] * fas
    int cinfo[K];              // column info words; -1 for padding columns (j >= c1)
    int j0;                    // first column of the strip
};

template <int K> struct PsAcc {
    double nX[K], nM[K];
    unsigned pX[K], pM[K];
};

template <bool SMALLTAB>
__device__ __forceinline__ void ps_subst(const PsCtx &c, const PsHot &h3, int sl, int colbase, double &mlog, double &xlog) {
    if (SMALLTAB) {
#ifdef PG2_HOST_EMU
        const double2 v = c.stab[sl + colbase];
        mlog = v.x;
        xlog = v.y;
#else
        // (an explicit shared-memory load: through the pointer kept in the shared-memory context it would be a generic one)
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(mlog), "=d"(xlog) : "r"(h3.stab_s + (unsigned)(sl + colbase) * 16u));
#endif
    } else {
        const double ls = (double)__ldg(c.table + sl + colbase);
        mlog = __dadd_rn(c.lng2, ls);
        xlog = __dadd_rn(c.lng, ls);
    }
}

// double4 through L2 (written by another lane or warp of the CTA a moment ago: not to be served from a stale L1 line)
__device__ __forceinline__ double4 ps_ldcg(const double4 *p) {
    const double2 a = __ldcg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldcg(reinterpret_cast<const double2 *>(p) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// first-wins running maximum: the candidate replaces the best only when strictly greater (basic_alignment.h:449-462)
__device__ __forceinline__ void ps_cand(double s, unsigned code, double &best, unsigned &ptr) {
    const bool p = s > best;
    best = p ? s : best;
    ptr = p ? code : ptr;
}

// history of the parked columns: entry (site i, slot) holds X, Y, M of cell (i, parked column)
__device__ __forceinline__ double *ps_hist(const PsCtx &c, int i, int slot) {
    return c.hist + ((((i + PS_HIST) & (PS_HIST - 1)) * c.park_cap + slot) * 3);
}

// M of a GENERAL column for one left edge: every backward edge pr -> j, sources (p, pr) from the history (the edge starts
// at the row above) or from the parked row
__device__ __forceinline__ void ps_general_m(const PsCtx &c, int i, int j, bool reg, const double4 *prow, double mlog, double xlog, double wl,
                                             unsigned lord, double &best, unsigned &ptr) {
    const double ninf = neg_inf();
    const int kr0 = c.r_off[j], kr1 = c.r_off[j + 1];
    for (int kr = kr0; kr < kr1; ++kr) {
        double vx = ninf, vy = ninf, vm = ninf;
        if (reg) {
            const double *h = ps_hist(c, i - 1, c.r_einfo[kr]);
            vx = h[0]; vy = h[1]; vm = h[2];
        } else if (prow) {
            const double4 v = ps_ldcg(prow + (c.r_estart[kr] - c.c0) + 1);
            vx = v.x; vy = v.y; vm = v.z;
        }
        const double wrk = (double)c.r_elogw[kr];
        const unsigned code = lord | ((unsigned)(kr - kr0) << 8);
        ps_cand(__dadd_rn(__dadd_rn(__dadd_rn(vm, mlog), wl), wrk), M_MAT | code, best, ptr);   // :2029-2112
        ps_cand(__dadd_rn(__dadd_rn(__dadd_rn(vx, xlog), wl), wrk), X_MAT | code, best, ptr);
        ps_cand(__dadd_rn(__dadd_rn(__dadd_rn(vy, xlog), wl), wrk), Y_MAT | code, best, ptr);
    }
}

// Y of a GENERAL column: every backward edge pr -> j, sources (i, pr) from the history
__device__ __forceinline__ void ps_general_y(const PsCtx &c, const PsHot &h3, int i, int j, double extY, double &best, unsigned &ptr) {
    const int kr0 = c.r_off[j], kr1 = c.r_off[j + 1];
    for (int kr = kr0; kr < kr1; ++kr) {
        const double *h = ps_hist(c, i, c.r_einfo[kr]);
        const double penY = ((h3.flags & PH_REDUCED) && c.r_estart[kr] == 0) ? 0.0 : h3.open;
        const unsigned ord = (unsigned)(kr - kr0) << 2;
        ps_cand(__dadd_rn(h[1], extY), Y_MAT | ord, best, ptr);
        ps_cand(__dadd_rn(h[0], h3.open), X_MAT | ord, best, ptr);
        ps_cand(__dadd_rn(__dadd_rn(h[2], h3.lng), penY), M_MAT | ord, best, ptr);
    }
}

// One virtual row (one backward edge of the left site) of one lane.  Returns true when the site was completed: st then
// holds row i (cells outside the band forced to -inf) and out[] the pointer words of its K cells.
//   any_saved (warp-uniform): some lane's edge starts at a parked row this step
//   rX, rY, rM: cell (i, j0 - 1), the strip's left neighbour in the row being completed
template <int K, bool SMALLTAB>
__device__ __forceinline__ bool ps_step(const PsCtx &c, const PsHot &h3, PsLane<K> &st, PsAcc<K> &acc, int lane, int4 vr, bool any_saved,
                                        double rX, double rY, double rM, unsigned *out) {
    const double ninf = neg_inf();
    const int info = vr.x, i = vr.z;
    const int sl = info & VR_STATE_MASK;
    const bool first = (info & VR_FIRST) != 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        acc.nX[k] = first ? ninf : acc.nX[k];
        acc.nM[k] = first ? ninf : acc.nM[k];
        acc.pX[k] = first ? (unsigned)NO_MAT : acc.pX[k];
        acc.pM[k] = first ? (unsigned)NO_MAT : acc.pM[k];
    }
    const bool reg = (info & VR_REG) != 0;
    const bool edge = !(info & VR_NOEDGE);
    const double4 *prow = nullptr;
    // An edge that starts at a parked row: the lane's strip registers take the parked row for this step and get the
    // previous row back afterwards (nothing is copied on steps without such an edge)
    double tX[K], tY[K], tM[K], tMo[K], tbX = 0, tbY = 0, tbM = 0;
    const bool swap = any_saved && edge && !reg;
    if (any_saved) {
        if (swap) {
            const int p = c.l_estart[vr.y];
            // a row above the block's first row lies outside the band for every column of the block (and for c0 - 1)
            const bool pvalid = p >= c.i0;
            const double pen = ((h3.flags & PH_REDUCED) && p == 0) ? 0.0 : h3.open;  // get_log_gap_open_penalty (basic_alignment.h:490-513)
            const double4 *row = c.saved + (long long)(vr.w & 0xffff) * c.saved_stride;
            prow = pvalid ? row : nullptr;
            tbX = st.bX; tbY = st.bY; tbM = st.bM;
            double4 v = make_double4(ninf, ninf, ninf, 0.0);
            if (pvalid) v = ps_ldcg(row + (st.j0 - c.c0));
            st.bX = v.x; st.bY = v.y; st.bM = v.z;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                tX[k] = st.X[k]; tY[k] = st.Y[k]; tM[k] = st.M[k]; tMo[k] = st.Mo[k];
                v = make_double4(ninf, ninf, ninf, 0.0);
                if (pvalid) v = ps_ldcg(row + (st.j0 - c.c0) + k + 1);
                st.X[k] = v.x; st.Y[k] = v.y; st.M[k] = v.z;
                st.Mo[k] = __dadd_rn(__dadd_rn(v.z, h3.lng), pen);
            }
        }
    }
    if (edge) {
        const double wl = (info & VR_ZERO_W) ? 0.0 : (double)c.l_elogw[vr.y];
        const unsigned lord = ((unsigned)vr.w >> 16) << 2;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            // X: ext, double, open out of (p, j) (:2116-2211)
            ps_cand(__dadd_rn(st.X[k], st.extX[k]), X_MAT | lord, acc.nX[k], acc.pX[k]);
            ps_cand(__dadd_rn(st.Y[k], h3.open), Y_MAT | lord, acc.nX[k], acc.pX[k]);
            ps_cand(st.Mo[k], M_MAT | lord, acc.nX[k], acc.pX[k]);
            // M: from M, X, Y of (p, pr) for every backward edge pr -> j (:1353-1436, :2029-2112)
            double mlog, xlog;
            ps_subst<SMALLTAB>(c, h3, sl, st.colbase[k], mlog, xlog);
            if (!(st.cinfo[k] & PC_GENERAL)) {
                const double qM = k ? st.M[k - 1] : st.bM, qX = k ? st.X[k - 1] : st.bX, qY = k ? st.Y[k - 1] : st.bY;
                double a = __dadd_rn(qM, mlog), b = __dadd_rn(qX, xlog), d = __dadd_rn(qY, xlog);
                if (h3.flags & PH_WEIGHTS) {
                    a = __dadd_rn(__dadd_rn(a, wl), st.wr[k]);
                    b = __dadd_rn(__dadd_rn(b, wl), st.wr[k]);
                    d = __dadd_rn(__dadd_rn(d, wl), st.wr[k]);
                }
                ps_cand(a, M_MAT | lord, acc.nM[k], acc.pM[k]);
                ps_cand(b, X_MAT | lord, acc.nM[k], acc.pM[k]);
                ps_cand(d, Y_MAT | lord, acc.nM[k], acc.pM[k]);
            } else if (st.cinfo[k] >= 0) {
                ps_general_m(c, i, st.j0 + k, reg, prow, mlog, xlog, wl, lord, acc.nM[k], acc.pM[k]);
            }
        }
    }
    if (any_saved) {
        if (swap) {
            st.bX = tbX; st.bY = tbY; st.bM = tbM;
#pragma unroll
            for (int k = 0; k < K; ++k) { st.X[k] = tX[k]; st.Y[k] = tY[k]; st.M[k] = tM[k]; st.Mo[k] = tMo[k]; }
        }
    }
    if (!(info & VR_LAST)) return false;

    // ---- the site's last virtual row: Y chain along the strip, band mask, pointer words ----
    if (st.j0 == 0) {  // DP column 0 has no M; (0,0) is the start corner (:725-733, :956-969)
        acc.nM[0] = (i == 0) ? 0.0 : ninf;
        acc.pM[0] = NO_MAT;
    }
    const double extY = ((h3.flags & PH_TERM) && (i == 0 || i == h3.lx1)) ? c.end_ext : h3.ext;
    int blo = 0, bhi = 0x7fffffff;
    if (h3.flags & PH_BANDED) { blo = c.blo[i]; bhi = c.bhi[i]; }
    const unsigned plain_row = ((info & VR_FAST) == VR_FAST && edge) ? PSW_PLAIN_ROW : 0u;
    // Y: ext, double, open out of (i, j - 1) (:2116-2211 with the roles of X and Y swapped).  The two candidates that do
    // not depend on the chain are folded first: (g > a ? g : a) with g = first-wins(double, open) equals the sequential
    // first-wins over ext, double, open
    double lXo = __dadd_rn(rX, h3.open), lMo = __dadd_rn(__dadd_rn(rM, h3.lng), h3.open), lY = rY;
    if (st.j0 == 1 && i == 0 && (h3.flags & PH_REDUCED)) lMo = __dadd_rn(__dadd_rn(rM, h3.lng), 0.0);  // the neighbour is the start corner
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = st.j0 + k;
        double ny;
        unsigned py, plain_col = 0;
        if (!(st.cinfo[k] & PC_GENERAL)) {
            const bool p2 = lMo > lXo;
            const double g = p2 ? lMo : lXo;
            const double a = __dadd_rn(lY, extY);
            const bool p1 = g > a;
            ny = p1 ? g : a;
            py = p1 ? (p2 ? (unsigned)M_MAT : (unsigned)X_MAT) : (unsigned)Y_MAT;
            py = (ny == ninf) ? (unsigned)NO_MAT : py;  // no candidate: the cell keeps no pointer (it is never walked)
            plain_col = j > 0 ? PSW_PLAIN_COL : 0u;
        } else {
            ny = ninf;
            py = NO_MAT;
            if (st.cinfo[k] >= 0) ps_general_y(c, h3, i, j, extY, ny, py);
        }
        double nx = acc.nX[k], nm = acc.nM[k];
        if (j < blo || j > bhi) { nx = ninf; ny = ninf; nm = ninf; }  // Tunnel_slice::at: -inf outside the band
        double nmo = __dadd_rn(__dadd_rn(nm, h3.lng), h3.open);
        if (j == 0 && i == 0 && (h3.flags & PH_REDUCED)) nmo = __dadd_rn(__dadd_rn(nm, h3.lng), 0.0);  // the start corner's gap moves
        out[k] = cell_word(acc.pX[k], py, acc.pM[k]) | plain_row | plain_col;
        st.X[k] = nx; st.Y[k] = ny; st.M[k] = nm; st.Mo[k] = nmo;
        if (st.cinfo[k] >= 0 && (st.cinfo[k] & (PC_PARKED | PC_ENDCOL))) {
            if (st.cinfo[k] & PC_PARKED) {
                double *h = ps_hist(c, i, (st.cinfo[k] >> PC_SLOT_SHIFT) & PC_SLOT_MASK);
                h[0] = nx; h[1] = ny; h[2] = nm;
            }
            if ((st.cinfo[k] & PC_ENDCOL) && (info & VR_ENDPRED))
                c.endstore[(long long)((st.cinfo[k] >> PC_END_SHIFT) & 3) * c.lx + i] = make_double4(nx, ny, nm, 0.0);
        }
        lXo = __dadd_rn(nx, h3.open); lMo = nmo; lY = ny;
    }
    st.bX = rX; st.bY = rY; st.bM = rM;
    return true;
}

// ---- ps_step_ring: the same virtual row with every source row in shared memory -------------------------------------------
// ps_step keeps the row above in registers and fetches any other source row (a long-span edge) from a parked copy in global
// memory; the lanes of a warp are on 32 different rows, so on ancestor graphs almost every step runs the swap code for a few
// lanes and the general-column loops for a few others (ncu, 1.2 k x 1.2 k root: 11.7 of 32 threads per instruction, 1 150
// instructions per step).  Here the block keeps its last rr rows in a shared-memory ring: EVERY source cell -- (p, j) for X,
// (p, j - 1) or (p, pr) for M, (i, pr) for the Y of a general column -- is one shared-memory load at an address computed from the
// row program entry, whatever the edge; there is no swap, no parked row, no column history and no branch on the kind of row.
// Hazards: lane l reads columns of lanes l' <= l; lane l' is l - l' virtual rows ahead and overwrites ring row (i' & (rr - 1)) as
// it completes site i', so rr must exceed the longest left span plus the lane distance of the longest right span plus 2 (the
// engine picks rr per launch group, try_pstrip).
template <int K> __device__ __forceinline__ int ps_rr_index(int row, int mat, int k, int l1) {
#ifdef PG2_HOST_EMU
    // the CPU test build checks every ring address the step computes (rows 0 .. 128, the "no row" included)
    if (row < 0 || row > 128 || mat < 0 || mat > 2 || k < 0 || k >= K || l1 < 0 || l1 > 32) abort();
#endif
    return ((row * 3 + mat) * K + k) * 33 + l1;
}

template <int K, bool SMALLTAB>
__device__ __forceinline__ bool ps_step_ring(const PsCtx &c, const PsHot &h3, PsLane<K> &st, PsAcc<K> &acc, int lane, int4 vr,
                                             double rX, double rY, double rM, unsigned *out) {
    const double ninf = neg_inf();
    const int info = vr.x, i = vr.z;
    const int sl = info & VR_STATE_MASK;
    const bool first = (info & VR_FIRST) != 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        acc.nX[k] = first ? ninf : acc.nX[k];
        acc.nM[k] = first ? ninf : acc.nM[k];
        acc.pX[k] = first ? (unsigned)NO_MAT : acc.pX[k];
        acc.pM[k] = first ? (unsigned)NO_MAT : acc.pM[k];
    }
    const bool edge = !(info & VR_NOEDGE);
    int p = i - 1;
    if (edge && !(info & VR_REG)) p = c.l_estart[vr.y];
    // a row above the block's first row lies outside the band for every column of the block (and for c0 - 1): the "no row"
    const bool pvalid = edge && p >= c.i0;
    const int rp = pvalid ? (p & (c.rr - 1)) : c.rr;
    const double pen = ((h3.flags & PH_REDUCED) && p == 0) ? 0.0 : h3.open;  // get_log_gap_open_penalty (basic_alignment.h:490-513)
    const double wl = (info & VR_ZERO_W) ? 0.0 : (double)c.l_elogw[edge ? vr.y : 0];
    const unsigned lord = ((unsigned)vr.w >> 16) << 2;
    const double *R = c.rowring;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        // X: ext, double, open out of (p, j) (:2116-2211)
        const double sX = R[ps_rr_index<K>(rp, 0, k, lane + 1)], sY = R[ps_rr_index<K>(rp, 1, k, lane + 1)], sM = R[ps_rr_index<K>(rp, 2, k, lane + 1)];
        ps_cand(__dadd_rn(sX, st.extX[k]), X_MAT | lord, acc.nX[k], acc.pX[k]);
        ps_cand(__dadd_rn(sY, h3.open), Y_MAT | lord, acc.nX[k], acc.pX[k]);
        ps_cand(__dadd_rn(__dadd_rn(sM, h3.lng), pen), M_MAT | lord, acc.nX[k], acc.pX[k]);
        // M: from M, X, Y of (p, pr) for every backward edge pr -> j (:1353-1436, :2029-2112)
        double mlog, xlog;
        ps_subst<SMALLTAB>(c, h3, sl, st.colbase[k], mlog, xlog);
        if (!(st.cinfo[k] & PC_GENERAL)) {
            const int kk = k ? k - 1 : K - 1, l1 = k ? lane + 1 : lane;  // column j - 1: the lane's own, or the last one of the lane before
            const double qX = R[ps_rr_index<K>(rp, 0, kk, l1)], qY = R[ps_rr_index<K>(rp, 1, kk, l1)], qM = R[ps_rr_index<K>(rp, 2, kk, l1)];
            double a = __dadd_rn(qM, mlog), b = __dadd_rn(qX, xlog), d = __dadd_rn(qY, xlog);
            if (h3.flags & PH_WEIGHTS) {
                a = __dadd_rn(__dadd_rn(a, wl), st.wr[k]);
                b = __dadd_rn(__dadd_rn(b, wl), st.wr[k]);
                d = __dadd_rn(__dadd_rn(d, wl), st.wr[k]);
            }
            ps_cand(a, M_MAT | lord, acc.nM[k], acc.pM[k]);
            ps_cand(b, X_MAT | lord, acc.nM[k], acc.pM[k]);
            ps_cand(d, Y_MAT | lord, acc.nM[k], acc.pM[k]);
        } else if (st.cinfo[k] >= 0 && edge) {
            const int j = st.j0 + k, kr0 = c.r_off[j], kr1 = c.r_off[j + 1];
            for (int kr = kr0; kr < kr1; ++kr) {
                const int cc = c.r_estart[kr] - c.c0;  // >= 0: blocks start at cut points of the column graph
                const int kk = cc % K, l1 = cc / K + 1;
                const double vx = R[ps_rr_index<K>(rp, 0, kk, l1)], vy = R[ps_rr_index<K>(rp, 1, kk, l1)], vm = R[ps_rr_index<K>(rp, 2, kk, l1)];
                const double wrk = (double)c.r_elogw[kr];
                const unsigned code = lord | ((unsigned)(kr - kr0) << 8);
                ps_cand(__dadd_rn(__dadd_rn(__dadd_rn(vm, mlog), wl), wrk), M_MAT | code, acc.nM[k], acc.pM[k]);
                ps_cand(__dadd_rn(__dadd_rn(__dadd_rn(vx, xlog), wl), wrk), X_MAT | code, acc.nM[k], acc.pM[k]);
                ps_cand(__dadd_rn(__dadd_rn(__dadd_rn(vy, xlog), wl), wrk), Y_MAT | code, acc.nM[k], acc.pM[k]);
            }
        }
    }
    if (!(info & VR_LAST)) return false;

    // ---- the site's last virtual row: Y chain along the strip, band mask, pointer words, the row goes into the ring ----
    if (st.j0 == 0) {  // DP column 0 has no M; (0,0) is the start corner (:725-733, :956-969)
        acc.nM[0] = (i == 0) ? 0.0 : ninf;
        acc.pM[0] = NO_MAT;
    }
    const double extY = ((h3.flags & PH_TERM) && (i == 0 || i == h3.lx1)) ? c.end_ext : h3.ext;
    int blo = 0, bhi = 0x7fffffff;
    if (h3.flags & PH_BANDED) { blo = c.blo[i]; bhi = c.bhi[i]; }
    const unsigned plain_row = ((info & VR_FAST) == VR_FAST && edge) ? PSW_PLAIN_ROW : 0u;
    const int ri = i & (c.rr - 1);
    double *W = c.rowring;
    if (lane == 0) {  // the column left of the block, row i (rX, rY, rM: the boundary entry lane 0 fetched)
        W[ps_rr_index<K>(ri, 0, K - 1, 0)] = rX; W[ps_rr_index<K>(ri, 1, K - 1, 0)] = rY; W[ps_rr_index<K>(ri, 2, K - 1, 0)] = rM;
    }
    // cell (i, j0 - 1): the lane to the left completed row i one step ago and left it in the ring -- no shuffle
    rX = W[ps_rr_index<K>(ri, 0, K - 1, lane)]; rY = W[ps_rr_index<K>(ri, 1, K - 1, lane)]; rM = W[ps_rr_index<K>(ri, 2, K - 1, lane)];
    double lXo = __dadd_rn(rX, h3.open), lMo = __dadd_rn(__dadd_rn(rM, h3.lng), h3.open), lY = rY;
    if (st.j0 == 1 && i == 0 && (h3.flags & PH_REDUCED)) lMo = __dadd_rn(__dadd_rn(rM, h3.lng), 0.0);  // the neighbour is the start corner
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = st.j0 + k;
        double ny;
        unsigned py, plain_col = 0;
        if (!(st.cinfo[k] & PC_GENERAL)) {
            const bool p2 = lMo > lXo;
            const double g = p2 ? lMo : lXo;
            const double a = __dadd_rn(lY, extY);
            const bool p1 = g > a;
            ny = p1 ? g : a;
            py = p1 ? (p2 ? (unsigned)M_MAT : (unsigned)X_MAT) : (unsigned)Y_MAT;
            py = (ny == ninf) ? (unsigned)NO_MAT : py;  // no candidate: the cell keeps no pointer (it is never walked)
            plain_col = j > 0 ? PSW_PLAIN_COL : 0u;
        } else {
            ny = ninf;
            py = NO_MAT;
            if (st.cinfo[k] >= 0) {
                // Y of a general column: every backward edge pr -> j, sources (i, pr) from the ring (written by this lane a
                // moment ago or by a lane to the left on an earlier step)
                const int kr0 = c.r_off[j], kr1 = c.r_off[j + 1];
                for (int kr = kr0; kr < kr1; ++kr) {
                    const int pr = c.r_estart[kr], cc = pr - c.c0;
                    const int kk = cc % K, l1 = cc / K + 1;
                    const double vx = W[ps_rr_index<K>(ri, 0, kk, l1)], vy = W[ps_rr_index<K>(ri, 1, kk, l1)], vm = W[ps_rr_index<K>(ri, 2, kk, l1)];
                    const double penY = ((h3.flags & PH_REDUCED) && pr == 0) ? 0.0 : h3.open;
                    const unsigned ord = (unsigned)(kr - kr0) << 2;
                    ps_cand(__dadd_rn(vy, extY), Y_MAT | ord, ny, py);
                    ps_cand(__dadd_rn(vx, h3.open), X_MAT | ord, ny, py);
                    ps_cand(__dadd_rn(__dadd_rn(vm, h3.lng), penY), M_MAT | ord, ny, py);
                }
            }
        }
        double nx = acc.nX[k], nm = acc.nM[k];
        if (j < blo || j > bhi) { nx = ninf; ny = ninf; nm = ninf; }  // Tunnel_slice::at: -inf outside the band
        double nmo = __dadd_rn(__dadd_rn(nm, h3.lng), h3.open);
        if (j == 0 && i == 0 && (h3.flags & PH_REDUCED)) nmo = __dadd_rn(__dadd_rn(nm, h3.lng), 0.0);  // the start corner's gap moves
        out[k] = cell_word(acc.pX[k], py, acc.pM[k]) | plain_row | plain_col;
        st.X[k] = nx; st.Y[k] = ny; st.M[k] = nm;
        W[ps_rr_index<K>(ri, 0, k, lane + 1)] = nx; W[ps_rr_index<K>(ri, 1, k, lane + 1)] = ny; W[ps_rr_index<K>(ri, 2, k, lane + 1)] = nm;
        if (st.cinfo[k] >= 0 && (st.cinfo[k] & PC_ENDCOL) && (info & VR_ENDPRED))
            c.endstore[(long long)((st.cinfo[k] >> PC_END_SHIFT) & 3) * c.lx + i] = make_double4(nx, ny, nm, 0.0);
        lXo = __dadd_rn(nx, h3.open); lMo = nmo; lY = ny;
    }
    return true;
}

// per-lane constants of one block
template <int K>
__device__ __forceinline__ void ps_init_lane(const PsCtx &c, PsLane<K> &st, int lane) {
    const double ninf = neg_inf();
    st.j0 = c.c0 + lane * K;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = st.j0 + k;
        const bool v = j < c.c1;
        st.X[k] = st.Y[k] = st.M[k] = st.Mo[k] = ninf;
        st.extX[k] = (c.term && (j == 0 || j == c.ly - 1)) ? c.end_ext : c.ext;
        st.cinfo[k] = v ? c.colinfo[j] : -1;
        const bool plain = v && j >= 1 && !(st.cinfo[k] & PC_GENERAL);
        st.wr[k] = plain ? (double)c.r_elogw[c.r_off[j]] : 0.0;
        st.colbase[k] = (v && j >= 1) ? c.r_state[j] * c.fas : 0;
    }
    st.bX = st.bY = st.bM = ninf;
}

// iterate_bwd_edges_for_end_corner (:1440-1552) from the stored end columns.  Run by one thread after the last block.
__device__ void ps_end_corner(const PsCtx &c, const double4 *endstore, DevResult *res) {
    const int kl0 = c.l_off[c.lx], kl1 = c.l_off[c.lx + 1], kr0 = c.r_off[c.ly], kr1 = c.r_off[c.ly + 1];
    const double ninf = neg_inf();
    double best = ninf;
    unsigned ptr = NO_MAT;
    auto cell = [&](int p, int q) {
        const int ci = c.colinfo[q];
        return ps_ldcg(endstore + (long long)((ci >> PC_END_SHIFT) & 3) * c.lx + p);
    };
    if (kl1 > kl0 && kr1 > kr0) {
        auto m_pair = [&](int kl, int kr) {
            const double4 v = cell(c.l_estart[kl], c.r_estart[kr]);
            const double s = __dadd_rn(__dadd_rn(__dadd_rn(v.z, c.lng), (double)c.l_elogw[kl]), (double)c.r_elogw[kr]);
            if (s > best) { best = s; ptr = pack_ptr(M_MAT, kl - kl0, kr - kr0); }
        };
        auto x_close = [&](int kl) {  // score_gap_close :2221-2255, close penalty 0
            const double s = cell(c.l_estart[kl], c.ly - 1).x;
            if (s > best) { best = s; ptr = pack_ptr(X_MAT, kl - kl0, 0); }
        };
        auto y_close = [&](int kr) {
            const double s = cell(c.lx - 1, c.r_estart[kr]).y;
            if (s > best) { best = s; ptr = pack_ptr(Y_MAT, 0, kr - kr0); }
        };
        m_pair(kl0, kr0);
        x_close(kl0);
        y_close(kr0);
        for (int kr = kr0 + 1; kr < kr1; ++kr) { m_pair(kl0, kr); y_close(kr); }
        for (int kl = kl0 + 1; kl < kl1; ++kl) {
            m_pair(kl, kr0);
            x_close(kl);
            for (int kr = kr0 + 1; kr < kr1; ++kr) { m_pair(kl, kr); y_close(kr); }
        }
    }
    res->score = best;
    res->end_ptr = ptr;
    res->status = (best == ninf) ? JOB_NO_PATH : JOB_OK;
}

// entries of the end columns the end corner may read: the rows that start an edge into the left stop site, and row lx - 1
__device__ __forceinline__ void ps_end_init(const PsCtx &c, double4 *endstore, int tid, int nthreads) {
    const double4 empty = make_double4(neg_inf(), neg_inf(), neg_inf(), 0.0);
    const int kl0 = c.l_off[c.lx], n = c.l_off[c.lx + 1] - kl0;
    for (int e = tid; e < (n + 1) * PS_MAX_END; e += nthreads) {
        const int which = e / PS_MAX_END, slot = e - which * PS_MAX_END;
        const int row = which < n ? c.l_estart[kl0 + which] : c.lx - 1;
        endstore[(long long)slot * c.lx + row] = empty;
    }
}

__device__ __forceinline__ void ps_make_ctx(PsCtx &c, const DevJob &J, const DevGraph &GL, const DevGraph &GR, const DevModel &m,
                                            const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw,
                                            const int4 *d_vrow, const int *d_vlast, const int *d_blo, const int *d_bhi) {
    c.l_vrow = d_vrow + GL.vrow_base;
    c.nv = GL.n_vrows;
    c.l_off = d_off + GL.off_base;
    c.l_estart = d_estart + GL.edge_base;
    c.l_elogw = d_elogw + GL.edge_base;
    c.r_state = d_state + GR.state_base;
    c.r_off = d_off + GR.off_base;
    c.r_estart = d_estart + GR.edge_base;
    c.r_elogw = d_elogw + GR.edge_base;
    c.r_einfo = d_vlast + GR.cp_ei_base;
    c.colinfo = d_vlast + GR.cp_ci_base;
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
    c.banded = J.banded != 0;
    c.weights = !(GL.zero_w && GR.zero_w);
    c.lx = J.lx;
    c.ly = J.ly;
    c.blo = c.banded ? d_blo + J.band_base : nullptr;
    c.bhi = c.banded ? d_bhi + J.band_base : nullptr;
}

__device__ __forceinline__ void ps_load_block(PsCtx &c, const int *blk) {
    c.c0 = blk[0]; c.c1 = blk[1]; c.v0 = blk[2]; c.v1 = blk[3]; c.i0 = blk[4];
}

#ifndef PG2_HOST_EMU
// Progress counters of the block pipeline live in global memory (the warps of one alignment may sit on several SMs of a
// thread-block cluster).  `wide`: producer and consumer may be on different SMs -- gpu scope; otherwise cta scope.
__device__ __forceinline__ int ps_load_acquire(const int *p, bool wide) {
    int v;
    // (relaxed, not acquire: a gpu-scope acquire invalidates the SM's L1 on every poll; what the flag guards is read with
    // cp.async.cg / ld.cg, which go to L2 anyway, and is issued only after the polling loop has seen the value)
    if (wide) asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else asm volatile("ld.acquire.cta.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ps_store_release(int *p, int v, bool wide) {
    if (wide) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    else asm volatile("st.release.cta.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int K, bool SMALLTAB, bool RING>
__global__ void __launch_bounds__(PS_MAX_WARPS * 32, 1)
pstrip_fill_kernel(int n_jobs, const DevJob *jobs, const int *job_ids, const DevGraph *graphs, const DevModel *models, const int *d_state,
                   const int *d_off, const int *d_estart, const float *d_elogw, const int4 *d_vrow, const int *d_vlast, const int *d_blo,
                   const int *d_bhi, unsigned *ptrs, DevResult *results, double4 *scratch, long long cta_d4, long long end_d4, int ring,
                   int max_slots, int park_cap, int rr, int *queue) {
    extern __shared__ __align__(16) unsigned char ps_smem[];
    // One alignment is swept by the warps of a whole thread-block CLUSTER: G CTAs of nw warps on G SMs (a guide-tree wave
    // holds a handful of alignments and 148 SMs; with 4 warps per CTA every warp has an SM sub-partition to itself).
    // Consecutive column blocks go to consecutive CTAs of the cluster: pipeline slot gw = w * G + rank takes blocks gw, gw + NW, ...
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int G = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const bool wide = G > 1;
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NW = G * nw, gw = w * G + rank;
    double *hist_all = reinterpret_cast<double *>(ps_smem);
    // per warp: the parked-column history (ps_step) or the row ring (ps_step_ring)
    // (an even count: the double2 table behind the warps' regions is read 16 bytes at a time)
    const int hist_doubles = ((RING ? (rr + 1) * 3 * K * 33 : PS_HIST * park_cap * 3) + 1) & ~1;
    double2 *s_tab = reinterpret_cast<double2 *>(hist_all + (size_t)nw * hist_doubles);
    __shared__ int s_job;
    // the warp-uniform constants live in shared memory, one copy per warp (the block fields differ): in registers they
    // would cost every lane some sixty registers
    __shared__ PsCtx s_ctx[PS_MAX_WARPS];
    // boundary column entries on their way in (cp.async, PS_PREFETCH steps ahead), a ring of 8 per warp: {X, Y, M, -}
    __shared__ __align__(16) double s_bnd[PS_MAX_WARPS][8][4];
    // ps_step: the strip's last column on its way to the next lane.  (A __shfl_up_sync here sits in a loop whose trip count
    // comes from memory; ptxas wraps every such shuffle in a WARPSYNC.COLLECTIVE call -- a fifth of the kernel's time, ncu.)
    __shared__ double s_xch[PS_MAX_WARPS][3][33];
    const double ninf = neg_inf();
    const int ring_mask = ring - 1;
    double4 *cl_scratch = scratch + (long long)(blockIdx.x / G) * cta_d4;  // one region per cluster
    double4 *endstore = cl_scratch;  // [PS_MAX_END][max lx]
    const long long warp_d4 = 2LL * ring + (long long)(max_slots > 0 ? max_slots : 1) * (1 + 32 * K);
    double4 *my = cl_scratch + end_d4 + (long long)gw * warp_d4;            // this pipeline slot: 2 boundary rings, parked rows
    double4 *prev = cl_scratch + end_d4 + (long long)((gw + NW - 1) % NW) * warp_d4;
    int *g_prog = reinterpret_cast<int *>(cl_scratch + end_d4 + (long long)NW * warp_d4);  // progress of every pipeline slot
    int tab_model = -1;

    for (;;) {
        cluster.sync();  // every CTA is done with the previous job (its end corner included)
        // the job index travels through global memory, not through rank 0's shared memory: a CTA that leaves the loop must not be
        // read by its cluster mates afterwards
        if (rank == 0) {
            if (threadIdx.x == 0) g_prog[NW] = atomicAdd(queue, 1);
            for (int e = threadIdx.x; e < NW; e += blockDim.x) g_prog[e] = 0;
            __threadfence();
        }
        cluster.sync();
        if (threadIdx.x == 0) s_job = ps_load_acquire(g_prog + NW, wide);
        __syncthreads();
        const int q = s_job;
        if (q >= n_jobs) break;
        const int jid = job_ids[q];
        const DevJob &J = jobs[jid];
        DevResult *res = results + jid;
        if (res->status != JOB_OK) continue;  // rejected by the validation kernel (cluster-uniform)
        const DevGraph GL = graphs[J.left], GR = graphs[J.right];
        const DevModel m = models[J.model];
        PsCtx &c = s_ctx[w];
        if (lane == 0) {
            ps_make_ctx(c, J, GL, GR, m, d_state, d_off, d_estart, d_elogw, d_vrow, d_vlast, d_blo, d_bhi);
            c.hist = hist_all + (size_t)w * hist_doubles;
            c.rowring = c.hist;
            c.rr = RING ? rr : 0;
            c.park_cap = park_cap;
            c.endstore = endstore;
            c.saved = my + 2LL * ring;
            c.saved_stride = 1 + 32 * K;
            c.stab = SMALLTAB ? s_tab : nullptr;
        }
        __syncwarp();
        if (SMALLTAB) {
            if (tab_model != J.model) {  // block-uniform
                for (int e = threadIdx.x; e < m.fas * m.fas; e += blockDim.x) {
                    const double ls = (double)m.table[e];
                    s_tab[e] = make_double2(__dadd_rn(c.lng2, ls), __dadd_rn(c.lng, ls));
                }
                tab_model = J.model;
                __syncthreads();
            }
        }
        const int *blocks = d_vlast + J.blk_base;
        const int n_blocks = J.n_blocks;
        const int stride = c.nv + 2;  // progress values of one round
        unsigned *P = ptrs + J.cell_base;
        // rows the end corner reads: -inf until a block writes them (a row above a banded block never is)
        if (rank == 0) { ps_end_init(c, endstore, threadIdx.x, blockDim.x); __threadfence(); }
        cluster.sync();

        for (int b = gw; b < n_blocks; b += NW) {
            const int round = b / NW;
            __syncwarp();
            if (lane == 0) ps_load_block(c, blocks + b * PB_INTS);
            __syncwarp();
            const int ptr_off = blocks[b * PB_INTS + 5];
            int pv0 = 0, pv1 = 0;
            if (b > 0) { pv0 = blocks[(b - 1) * PB_INTS + 2]; pv1 = blocks[(b - 1) * PB_INTS + 3]; }
            // the producer of this block's left boundary: slot gw - 1 in this round, or the last slot one round earlier
            const int *prod = g_prog + (gw + NW - 1) % NW;
            const int prod_base = (gw == 0 ? round - 1 : round) * stride;
            const double4 *bcol_prev = prev + (long long)(((gw == 0 ? round - 1 : round) & 1) ? ring : 0);
            double4 *bcol_cur = my + (long long)((round & 1) ? ring : 0);
            PsLane<K> st;
            PsAcc<K> acc;
            PsHot h3;
            h3.open = c.open; h3.ext = c.ext; h3.lng = c.lng;
            h3.flags = (c.term ? PH_TERM : 0u) | (c.reduced ? PH_REDUCED : 0u) | (c.banded ? PH_BANDED : 0u) | (c.weights ? PH_WEIGHTS : 0u);
            h3.lx1 = c.lx - 1;
            h3.stab_s = SMALLTAB ? (unsigned)__cvta_generic_to_shared(s_tab) : 0u;
            ps_init_lane<K>(c, st, lane);
#pragma unroll
            for (int k = 0; k < K; ++k) { acc.nX[k] = acc.nM[k] = ninf; acc.pX[k] = acc.pM[k] = NO_MAT; }
            for (int e = lane; e < hist_doubles; e += 32) c.hist[e] = ninf;
            __syncwarp();
            const int last_col = c.c1 - 1 - c.c0, last_lane = last_col / K, last_k = last_col % K;
            const bool feeds_next = b + 1 < n_blocks;
            const int v0 = c.v0, v1 = c.v1;
            // boundary column: lane 0 starts the copy of the entry of virtual row v into the ring PS_PREFETCH steps before it
            // is used; an entry the previous block does not hold (rows outside its share of the band) is -inf
            int avail = 0;  // producer progress seen so far
            double *ringp = &s_bnd[w][0][0];
            auto issue = [&](int v) {
                if (lane == 0) {
                    double *dst = ringp + ((v - v0) & 7) * 4;
                    if (b == 0 || v < pv0 || v >= pv1) {
                        dst[0] = ninf; dst[1] = ninf; dst[2] = ninf;
                    } else {
                        const int need = prod_base + v + 1;
                        while (avail < need) {
                            avail = ps_load_acquire(prod, wide);
                            if (avail < need) __nanosleep(64);  // the waiting warp leaves the issue slots to the warps that work
                        }
                        const double4 *src = bcol_prev + (v & ring_mask);
                        const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + 16), "l"(reinterpret_cast<const char *>(src) + 16) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
#pragma unroll
            for (int d = 0; d < PS_PREFETCH; ++d) issue(v0 + d);
            unsigned *out = P + ptr_off + lane * K;
            const int n_steps = (v1 - v0) + last_lane;
            // the row program entry of the NEXT step is fetched a step ahead
            const int4 vr_idle = make_int4(VR_FAST | VR_ZERO_W, -1, 0, 0);
            int4 vr_next = vr_idle;
            if (lane == 0 && v1 > v0) vr_next = __ldg(c.l_vrow + v0);
            for (int t = 0; t < n_steps; ++t) {
                double rX = 0.0, rY = 0.0, rM = 0.0;
                if (!RING) {  // (the row-ring step reads its left neighbour from the ring)
                    s_xch[w][0][lane + 1] = st.X[K - 1]; s_xch[w][1][lane + 1] = st.Y[K - 1]; s_xch[w][2][lane + 1] = st.M[K - 1];
                    __syncwarp();
                    rX = s_xch[w][0][lane]; rY = s_xch[w][1][lane]; rM = s_xch[w][2][lane];
                }
                asm volatile("cp.async.wait_group %0;" ::"n"(PS_PREFETCH - 1) : "memory");
                if (lane == 0) { const double *e = ringp + (t & 7) * 4; rX = e[0]; rY = e[1]; rM = e[2]; }
                issue(v0 + t + PS_PREFETCH);
                const int v = v0 + t - lane;
                const bool active = (v >= v0 && v < v1 && lane <= last_lane);
                const int4 vr = vr_next;
                {
                    const int vn = v + 1;
                    vr_next = vr_idle;
                    if (vn >= v0 && vn < v1 && lane <= last_lane) vr_next = __ldg(c.l_vrow + vn);
                }
                const bool any_saved = RING ? false : __any_sync(0xffffffffu, active && !(vr.x & (VR_REG | VR_NOEDGE)));
                if (active) {
                    unsigned wds[K];
                    const bool done = RING ? ps_step_ring<K, SMALLTAB>(c, h3, st, acc, lane, vr, rX, rY, rM, wds)
                                           : ps_step<K, SMALLTAB>(c, h3, st, acc, lane, vr, any_saved, rX, rY, rM, wds);
                    if (done) {
                        unsigned *dst = out + (long long)t * 32 * K;
                        if (K == 2) *reinterpret_cast<uint2 *>(dst) = make_uint2(wds[0], wds[1]);
                        else if (K == 4) *reinterpret_cast<uint4 *>(dst) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
                        else {
#pragma unroll
                            for (int k = 0; k < K; ++k) dst[k] = wds[k];
                        }
                        const int slot = RING ? -1 : (int)((unsigned)vr.x >> VR_SLOT_SHIFT) - 1;
                        if (slot >= 0) {  // park the row for long-span edges
                            double4 *row = c.saved + (long long)slot * c.saved_stride + (st.j0 - c.c0);
                            if (lane == 0) row[0] = make_double4(st.bX, st.bY, st.bM, 0.0);
#pragma unroll
                            for (int k = 0; k < K; ++k) row[k + 1] = make_double4(st.X[k], st.Y[k], st.M[k], 0.0);
                        }
                        if (lane == last_lane && feeds_next) {
                            double vx = ninf, vy = ninf, vm = ninf;
#pragma unroll
                            for (int k = 0; k < K; ++k) if (k == last_k) { vx = st.X[k]; vy = st.Y[k]; vm = st.M[k]; }
                            bcol_cur[v & ring_mask] = make_double4(vx, vy, vm, 0.0);
                        }
                    }
                    // progress: every 8th virtual row (a release store orders the thread's earlier stores: not every step)
                    if (lane == last_lane && feeds_next && ((v & 7) == 7 || v == v1 - 1)) ps_store_release(g_prog + gw, round * stride + (v + 1), wide);
                }
                __syncwarp();
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            if (lane == 0) ps_store_release(g_prog + gw, round * stride + c.nv + 1, wide);
        }
        __threadfence();  // the end columns this CTA wrote, before the cluster meets
        cluster.sync();
        if (rank == 0 && threadIdx.x == 0) ps_end_corner(c, endstore, res);
    }
}
#endif

// smem of one CTA: the parked-column histories of its warps (+ the shared substitution table)
// (rr > 0: the row ring of ps_step_ring, K = 1 or 2)
static size_t ps_smem_bytes(int nw, int park_cap, bool smalltab, int rr, int K) {
    const size_t per_warp = ((rr > 0 ? (size_t)(rr + 1) * 3 * K * 33 : (size_t)PS_HIST * park_cap * 3) + 1) & ~(size_t)1;
    return (size_t)nw * per_warp * sizeof(double) + (smalltab ? STRIP_SMALL_FAS * STRIP_SMALL_FAS * sizeof(double2) : 0);
}
// warps per CTA: as many as the job has blocks, the kernel allows and the histories / row rings fit
int pstrip_warps(int n_blocks, int park_cap, bool smalltab, int rr, int K) {
    int nw = n_blocks < PS_MAX_WARPS ? n_blocks : PS_MAX_WARPS;
    if (K > 2) rr = 0;
    const size_t budget = rr > 0 ? (size_t)224 * 1024 : (size_t)PS_SMEM_BUDGET;  // (two warps with 64-row rings: 210 KB)
    while (nw > 1 && ps_smem_bytes(nw, park_cap, smalltab, rr, K) > budget) --nw;
    return nw < 1 ? 1 : nw;
}

#ifdef PG2_HOST_EMU
// CPU test emulation of one job: the same step body; the blocks one after the other (one of the interleavings the
// progress counters admit), lanes one after the other inside a step with the shuffle replaced by a snapshot.
template <int K>
static void ps_emulate_job(const DevJob &J, const DevGraph &GL, const DevGraph &GR, const DevModel &m, const int *d_state, const int *d_off,
                           const int *d_estart, const float *d_elogw, const int4 *d_vrow, const int *d_vlast, const int *d_blo,
                           const int *d_bhi, unsigned *ptrs, DevResult *res, double4 *scratch, long long end_d4, int ring, int max_slots,
                           int park_cap, int rr) {
    const double ninf = neg_inf();
    PsCtx c;
    ps_make_ctx(c, J, GL, GR, m, d_state, d_off, d_estart, d_elogw, d_vrow, d_vlast, d_blo, d_bhi);
    std::vector<double> hist(rr > 0 ? (size_t)(rr + 1) * 3 * K * 33 : (size_t)PS_HIST * park_cap * 3);
    c.hist = hist.data();
    c.rowring = hist.data();
    c.rr = rr;
    c.park_cap = park_cap;
    c.endstore = scratch;
    double4 *ringbuf[2] = {scratch + end_d4, scratch + end_d4 + ring};
    c.saved = scratch + end_d4 + 2LL * ring;
    c.saved_stride = 1 + 32 * K;
    (void)max_slots;
    std::vector<double2> tab;
    const bool smalltab = m.fas <= STRIP_SMALL_FAS;
    if (smalltab) {
        tab.resize((size_t)m.fas * m.fas);
        for (int e = 0; e < m.fas * m.fas; ++e) {
            const double ls = (double)m.table[e];
            tab[e] = make_double2(c.lng2 + ls, c.lng + ls);
        }
        c.stab = tab.data();
    }
    const int *blocks = d_vlast + J.blk_base;
    unsigned *P = ptrs + J.cell_base;
    const int ring_mask = ring - 1;
    PsHot h3;
    h3.open = c.open; h3.ext = c.ext; h3.lng = c.lng;
    h3.flags = (c.term ? PH_TERM : 0u) | (c.reduced ? PH_REDUCED : 0u) | (c.banded ? PH_BANDED : 0u) | (c.weights ? PH_WEIGHTS : 0u);
    h3.lx1 = c.lx - 1;
    h3.stab_s = 0;
    ps_end_init(c, c.endstore, 0, 1);
    for (int b = 0; b < J.n_blocks; ++b) {
        ps_load_block(c, blocks + b * PB_INTS);
        const int ptr_off = blocks[b * PB_INTS + 5];
        int pv0 = 0, pv1 = 0;
        if (b > 0) { pv0 = blocks[(b - 1) * PB_INTS + 2]; pv1 = blocks[(b - 1) * PB_INTS + 3]; }
        const double4 *bcol_prev = ringbuf[(b + 1) & 1];
        double4 *bcol_cur = ringbuf[b & 1];
        PsLane<K> st[32];
        PsAcc<K> acc[32];
        for (int l = 0; l < 32; ++l) {
            ps_init_lane<K>(c, st[l], l);
            for (int k = 0; k < K; ++k) { acc[l].nX[k] = acc[l].nM[k] = ninf; acc[l].pX[k] = acc[l].pM[k] = NO_MAT; }
        }
        for (auto &h : hist) h = ninf;
        const int last_col = c.c1 - 1 - c.c0, last_lane = last_col / K, last_k = last_col % K;
        const bool feeds_next = b + 1 < J.n_blocks;
        const int n_steps = (c.v1 - c.v0) + last_lane;
        for (int t = 0; t < n_steps; ++t) {
            double sx[32], sy[32], sm[32];
            int4 vr[32];
            bool any_saved = false;
            for (int l = 0; l < 32; ++l) {
                sx[l] = st[l].X[K - 1]; sy[l] = st[l].Y[K - 1]; sm[l] = st[l].M[K - 1];
                const int v = c.v0 + t - l;
                const bool active = v >= c.v0 && v < c.v1 && l <= last_lane;
                vr[l] = active ? c.l_vrow[v] : make_int4(VR_FAST | VR_ZERO_W, -1, 0, 0);
                if (active && !(vr[l].x & (VR_REG | VR_NOEDGE))) any_saved = true;
            }
            for (int l = 0; l < 32; ++l) {
                const int v = c.v0 + t - l;
                if (v < c.v0 || v >= c.v1 || l > last_lane) continue;
                double rX, rY, rM;
                if (l == 0) {
                    rX = rY = rM = ninf;
                    if (b > 0 && v >= pv0 && v < pv1) { const double4 bv = bcol_prev[v & ring_mask]; rX = bv.x; rY = bv.y; rM = bv.z; }
                } else { rX = sx[l - 1]; rY = sy[l - 1]; rM = sm[l - 1]; }
                unsigned wds[K];
                bool done;
                if (rr > 0) {
                    if (smalltab) done = ps_step_ring<K, true>(c, h3, st[l], acc[l], l, vr[l], rX, rY, rM, wds);
                    else done = ps_step_ring<K, false>(c, h3, st[l], acc[l], l, vr[l], rX, rY, rM, wds);
                } else if (smalltab) done = ps_step<K, true>(c, h3, st[l], acc[l], l, vr[l], any_saved, rX, rY, rM, wds);
                else done = ps_step<K, false>(c, h3, st[l], acc[l], l, vr[l], any_saved, rX, rY, rM, wds);
                if (!done) continue;
                unsigned *dst = P + ptr_off + ((long long)t * 32 + l) * K;
                for (int k = 0; k < K; ++k) dst[k] = wds[k];
                const int slot = rr > 0 ? -1 : (int)((unsigned)vr[l].x >> VR_SLOT_SHIFT) - 1;
                if (slot >= 0) {
                    double4 *row = c.saved + (long long)slot * c.saved_stride + (st[l].j0 - c.c0);
                    if (l == 0) row[0] = make_double4(st[l].bX, st[l].bY, st[l].bM, 0.0);
                    for (int k = 0; k < K; ++k) row[k + 1] = make_double4(st[l].X[k], st[l].Y[k], st[l].M[k], 0.0);
                }
                if (l == last_lane && feeds_next)
                    bcol_cur[v & ring_mask] = make_double4(st[l].X[last_k], st[l].Y[last_k], st[l].M[last_k], 0.0);
            }
        }
    }
    ps_end_corner(c, c.endstore, res);
}
#endif

int pstrip_max_warps() { return PS_MAX_WARPS; }

// per-cluster scratch in double4: end columns, then per pipeline slot (nw = warps of the whole cluster) two boundary rings and
// the parked rows, then the progress counters of the slots
long long pstrip_cta_double4(int K, int nw, int max_lx, int ring, int max_slots) {
    return (long long)PS_MAX_END * max_lx + (long long)nw * (2LL * ring + (long long)(max_slots > 0 ? max_slots : 1) * (1 + 32 * K)) + (nw + 7) / 8 + 1;
}

// Launches one group of pipelined-strip jobs that share the strip width K and the table variant: n_clusters clusters of G CTAs
// of nw warps each, one alignment per cluster at a time.
void launch_pstrip_fill(int K, bool smalltab, int nw, int G, int n_jobs, int n_clusters, const DevJob *jobs, const int *job_ids, const DevGraph *graphs,
                        const DevModel *models, const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw,
                        const int4 *d_vrow, const int *d_vlast, const int *d_blo, const int *d_bhi, unsigned *ptrs, DevResult *results,
                        double4 *scratch, int max_lx, int ring, int max_slots, int park_cap, int rr, int *queue, cudaStream_t stream) {
    if (n_jobs <= 0) return;
    if (K > 2) rr = 0;  // the row ring is built for K = 1 and 2
    const long long end_d4 = (long long)PS_MAX_END * max_lx;
    const long long cta_d4 = pstrip_cta_double4(K, nw * G, max_lx, ring, max_slots);
#ifndef PG2_HOST_EMU
    cudaMemsetAsync(queue, 0, sizeof(int), stream);
    const int smem = (int)ps_smem_bytes(nw, park_cap, smalltab, rr, K);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(n_clusters * G));
    cfg.blockDim = dim3((unsigned)(nw * 32));
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)G;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define PG2_PS_LAUNCH(KK, S, RG)                                                                                                  \
    do {                                                                                                                          \
        cudaFuncSetAttribute(pstrip_fill_kernel<KK, S, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                   \
        cudaLaunchKernelEx(&cfg, pstrip_fill_kernel<KK, S, RG>, n_jobs, jobs, job_ids, graphs, models, d_state, d_off, d_estart,   \
                           d_elogw, d_vrow, d_vlast, d_blo, d_bhi, ptrs, results, scratch, cta_d4, end_d4, ring, max_slots,       \
                           park_cap, rr, queue);                                                                                  \
    } while (0)
    if (K == 1 && rr > 0) { if (smalltab) PG2_PS_LAUNCH(1, true, true); else PG2_PS_LAUNCH(1, false, true); }
    else if (K == 1) { if (smalltab) PG2_PS_LAUNCH(1, true, false); else PG2_PS_LAUNCH(1, false, false); }
    else if (K == 2 && rr > 0) { if (smalltab) PG2_PS_LAUNCH(2, true, true); else PG2_PS_LAUNCH(2, false, true); }
    else if (K == 2) { if (smalltab) PG2_PS_LAUNCH(2, true, false); else PG2_PS_LAUNCH(2, false, false); }
    else { if (smalltab) PG2_PS_LAUNCH(4, true, false); else PG2_PS_LAUNCH(4, false, false); }
#undef PG2_PS_LAUNCH
#else
    (void)queue; (void)stream; (void)n_clusters; (void)nw; (void)G; (void)smalltab; (void)cta_d4;
    for (int q = 0; q < n_jobs; ++q) {
        const int jid = job_ids[q];
        const DevJob &J = jobs[jid];
        DevResult *res = results + jid;
        if (res->status != JOB_OK) continue;
        if (K == 1) ps_emulate_job<1>(J, graphs[J.left], graphs[J.right], models[J.model], d_state, d_off, d_estart, d_elogw, d_vrow, d_vlast,
                                      d_blo, d_bhi, ptrs, res, scratch, end_d4, ring, max_slots, park_cap, rr);
        else if (K == 2) ps_emulate_job<2>(J, graphs[J.left], graphs[J.right], models[J.model], d_state, d_off, d_estart, d_elogw, d_vrow, d_vlast,
                                      d_blo, d_bhi, ptrs, res, scratch, end_d4, ring, max_slots, park_cap, rr);
        else ps_emulate_job<4>(J, graphs[J.left], graphs[J.right], models[J.model], d_state, d_off, d_estart, d_elogw, d_vrow, d_vlast, d_blo,
                               d_bhi, ptrs, res, scratch, end_d4, ring, max_slots, park_cap, 0);
    }
#endif
}

}  // namespace pg2
