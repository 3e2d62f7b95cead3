// pg2_lanes.cu -- placement fill kernel: one LANE per alignment, 32 alignments that share the row graph per
// warp, the column strips of those alignments pipelined over the warps of a CTA.
//
// Query placement aligns many reads against the same tree node (reads_aligner.cpp:983-1216; the temporary
// node puts the tree node LEFT and the read RIGHT, reads_aligner.h:169-184), so a launch batch holds many jobs
// with the same LEFT graph.  The engine groups them into tasks of 32 (LaneTask, pg2_device.cuh); one CTA takes
// a task and every lane aligns its own read against the shared row graph.
//
// Why: all lanes are on the SAME row at the same time, so everything the row graph decides -- in-degree, edge
// spans, edge weights, which rows must be parked for long-span edges -- is warp-uniform.  A site with one
// unit-weight edge from the row above (98 % of the sites of a PAGAN ancestor graph) takes the in-place fast
// body; any other site parks the row above and accumulates its edges one by one.  No lane ever waits for
// another lane's row shape, there is no skew ramp and no shuffle.
//
// Layout.  A lane holds a strip of K = 8 columns of its read in registers and sweeps it down the rows.  The
// strips of a task are dealt round-robin to the W = 4 warps of the CTA, which run as a software pipeline over
// blocks of B = 8 virtual rows: warp w works on a block of its strip as soon as its predecessor warp has
// published the same block of the strip to the left.  That boundary column travels through a shared-memory
// ring [channel][slot][row][X,Y,M][lane], D = 2 blocks deep, guarded by per-warp progress counters (a producer
// may run up to D blocks ahead of its consumer; nobody waits at a CTA-wide barrier).  Only every fourth strip
// boundary (warp 3 -> warp 0 of the next round) goes through a global wrap buffer and comes back by cp.async
// one block ahead.
// Back-pointers: one uint16 per cell, 8 columns per 128-bit store, lanes interleaved (pg2_strip_geom.cuh).
//
// Arithmetic: candidate by candidate as the reference (src/main/viterbi_alignment.cpp:856-971, 1328-1436,
// 2029-2219), same FP64 association, strict '>' (first candidate wins ties).  The term (M + log_non_gap) +
// log_gap_open, which the reference computes twice per cell (X move out of the cell, :2190-2211, and Y move
// out of it), is computed once and kept next to M ("Mo").  The reduced terminal penalty (basic_alignment.h:
// 490-513) only ever meets a finite M at the start corner, so it is planted there.
#include "pg2_device.cuh"
#include "pg2_strip_geom.cuh"
#ifdef PG2_HOST_EMU
#include <vector>
#endif
#ifdef PG2_LANE_TIMING
#include <cstdio>
#endif

#ifndef PG2_LANE_MINB
#define PG2_LANE_MINB 3
#endif
// rows of the hot loop per iteration (tuning: -DPG2_LANE_UNROLL=2)
#define PG2_PRAGMA_(x) _Pragma(#x)
#define PG2_PRAGMA(x) PG2_PRAGMA_(x)
#ifdef PG2_LANE_UNROLL
#define PG2_LANE_ROW_UNROLL PG2_PRAGMA(unroll PG2_LANE_UNROLL)
#else
#define PG2_LANE_ROW_UNROLL
#endif

namespace pg2 {

// warp-uniform per-task constants
struct LaneCtx {
    const int4 *l_vrow;
    const int *l_vplain;         // per pipeline block: bit r set when virtual row r of the block is a plain interior row
    const int *l_off, *l_estart;
    const float *l_elogw;
    int nv, lx, n_slots;
    const float *table;          // global float table (any alphabet)
    const double2 *stab;         // shared {2*lng + ls, lng + ls} table (alphabets up to STRIP_SMALL_FAS)
    unsigned stab_s;             // its shared-memory address: the kernel loads with ld.shared (through the generic pointer every row
                                 // re-derives the shared window in the uniform datapath: four issue slots)
    int fas;
    double open, ext, end_ext, lng, lng2;
    bool term, reduced;
};

// per-lane state of the strip the lane's warp is sweeping
template <int K> struct LState {
    double X[K], Y[K], M[K], Mo[K];  // row handled last; Mo = (M + lng) + open
    double bX, bY, bM;               // left neighbour column (c0-1) of that row
    double wr[K];                    // log weight of the column edge into site j (WR variants only)
    int tab[K];                      // SMALLTAB: byte offset of column j's table column; else state_r[j] * fas
};

// per-lane geometry of the job
struct LaneGeom {
    int ly;            // DP columns of this lane's read (1 for an idle lane)
    int last_strip;    // strip that holds column ly-1
    int last_k;        // its position in that strip
    bool active;
};

// s = max(b, a) with the first-wins rule (b replaces a only when strictly greater); the outcome is OR-ed into w
// as one bit.  One DSETP + one 64-bit select + one predicated LOP3.
__device__ __forceinline__ double sel_gt(double b, double a, unsigned &w, unsigned bit) {
#ifdef PG2_HOST_EMU
    if (b > a) { w |= bit; return b; }
    return a;
#else
    double t;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %2, %3;\n\tselp.f64 %0, %2, %3, p;\n\t@p or.b32 %1, %1, %4;\n\t}"
        : "=d"(t), "+r"(w)
        : "d"(b), "d"(a), "r"(bit));
    return t;
#endif
}

// running maximum of a general row: if (s > best) { best = s; ptr = (ptr & keep) | code; }
__device__ __forceinline__ void acc_gt(double s, double &best, unsigned &ptr, unsigned keep, unsigned code) {
#ifdef PG2_HOST_EMU
    if (s > best) { best = s; ptr = (ptr & keep) | code; }
#else
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %2, %0;\n\tselp.f64 %0, %2, %0, p;\n\t@p lop3.b32 %1, %1, %3, %4, 0xEA;\n\t}"
        : "+d"(best), "+r"(ptr)
        : "d"(s), "r"(keep), "r"(code));
#endif
}

// {m_log, x_log} of one cell: 2*log_non_gap + log_score and log_non_gap + log_score (:1363-1367).
// rowoff / tabk: byte offsets into the shared double2 table (SMALLTAB) or element offsets into the float table.
template <bool SMALLTAB>
__device__ __forceinline__ void lane_subst(const LaneCtx &c, int rowoff, int tabk, double &mlog, double &xlog) {
    if (SMALLTAB) {
#ifdef PG2_HOST_EMU
        const double2 v = *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(c.stab) + (rowoff + tabk));
        mlog = v.x;
        xlog = v.y;
#else
        asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(mlog), "=d"(xlog) : "r"(c.stab_s + (unsigned)(rowoff + tabk)));
#endif
    } else {
        const double ls = (double)__ldg(c.table + rowoff + tabk);
        mlog = __dadd_rn(c.lng2, ls);
        xlog = __dadd_rn(c.lng, ls);
    }
}

// Y(i,j) from (i,j-1) along the strip: ext, double, open (:2116-2211 with the roles of X and Y swapped).  The
// two chain-independent candidates are folded first; (g > a ? g : a) with g = first-wins(double, open) equals
// the sequential first-wins over all three.  Outcome bits go to P2 (open beat double) and P1 (that winner beat
// ext) of the cell's half-word.  Also makes (rX,rY,rM) the row's left neighbour.
template <int K, unsigned P2, unsigned P1>
__device__ __forceinline__ void lane_y_chain(const LaneCtx &c, LState<K> &st, double extY, double rX, double rY, double rM,
                                             unsigned *w) {
    double lXo = __dadd_rn(rX, c.open);
    double lMo = __dadd_rn(__dadd_rn(rM, c.lng), c.open);
    double lY = rY;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const unsigned sh = (k & 1) * 16;
        const double g = sel_gt(lMo, lXo, w[k >> 1], P2 << sh);
        const double a = __dadd_rn(lY, extY);
        const double ny = sel_gt(g, a, w[k >> 1], P1 << sh);
        lXo = __dadd_rn(st.X[k], c.open);
        lMo = st.Mo[k];
        lY = ny;
        st.Y[k] = ny;
    }
    st.bX = rX; st.bY = rY; st.bM = rM;
}

// Fast-row pointer half-word: bit 14 set, bits 0-5 the raw comparison outcomes
//   bit0/1 X: (double > ext), (open > max of the first two)      candidates in order X, Y, M
//   bit2/3 Y: (open > double), (that winner > ext)                candidates in order Y, X, M
//   bit4/5 M: (X > M), (Y > max of the first two)                 candidates in order M, X, Y
// (lane_decode_ptr, pg2_strip_geom.cuh).  A cell whose candidates are all -inf gets arbitrary bits: it cannot
// lie on the Viterbi path.
//
// Row whose only backward edge comes from the row above: in-place update of X and M of the strip (the Y chain
// follows, lane_y_chain).
//   WL      the edge carries a log weight (wl)           EXTARR  ex[k] holds the X-extension term of column k
//                                                        (else every column extends with log_gap_ext)
template <int K, bool WL, bool WR, bool SMALLTAB, bool EXTARR>
__device__ __forceinline__ void lane_xm_row(const LaneCtx &c, LState<K> &st, int sl, double wl, const double *ex, unsigned *w) {
#pragma unroll
    for (int h = 0; h < K / 2; ++h) w[h] = 0x40004000u;
    const int rowoff = SMALLTAB ? sl * 16 : sl;
    // descending k: X(i,j) reads (i-1,j), M(i,j) reads (i-1,j-1); both still hold row i-1
#pragma unroll
    for (int k = K - 1; k >= 0; --k) {
        const unsigned sh = (k & 1) * 16;
        const double qX = k ? st.X[k - 1] : st.bX, qY = k ? st.Y[k - 1] : st.bY, qM = k ? st.M[k - 1] : st.bM;
        // X: ext, double, open (:2116-2211)
        double a = __dadd_rn(st.X[k], EXTARR ? ex[k] : c.ext);
        double b = __dadd_rn(st.Y[k], c.open);
        double t = sel_gt(b, a, w[k >> 1], 1u << sh);
        const double nx = sel_gt(st.Mo[k], t, w[k >> 1], 2u << sh);
        // M: from M, X, Y (:2029-2112)
        double mlog, xlog;
        lane_subst<SMALLTAB>(c, rowoff, st.tab[k], mlog, xlog);
        a = __dadd_rn(qM, mlog);
        b = __dadd_rn(qX, xlog);
        double d = __dadd_rn(qY, xlog);
        if (WL) { a = __dadd_rn(a, wl); b = __dadd_rn(b, wl); d = __dadd_rn(d, wl); }
        if (WR) { a = __dadd_rn(a, st.wr[k]); b = __dadd_rn(b, st.wr[k]); d = __dadd_rn(d, st.wr[k]); }
        t = sel_gt(b, a, w[k >> 1], 16u << sh);
        const double nm = sel_gt(d, t, w[k >> 1], 32u << sh);
        st.X[k] = nx;
        st.M[k] = nm;
        st.Mo[k] = __dadd_rn(__dadd_rn(nm, c.lng), c.open);
    }
    // Column 0 needs no special case for i > 0: its M sources are the -inf boundary, so M(i,0) = -inf falls
    // out of the arithmetic.  Row 0 has no edges but its sources are the -inf initial strip, so X(0,j) =
    // M(0,j) = -inf fall out as well; the start corner is planted by the caller (lane_slow_row).
}

// A parked row of one warp: [column 0 = c0-1, column k+1 = c0+k][X, Y, M, Mo][lane].
template <int K>
__device__ __forceinline__ void lane_park(const LaneCtx &c, const LState<K> &st, double *slot) {
    slot[0] = st.bX; slot[32] = st.bY; slot[64] = st.bM;
    slot[96] = __dadd_rn(__dadd_rn(st.bM, c.lng), c.open);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double *p = slot + (k + 1) * 128;
        p[0] = st.X[k]; p[32] = st.Y[k]; p[64] = st.M[k]; p[96] = st.Mo[k];
    }
}

// Pointer accumulators of a site whose backward edges are visited one virtual row at a time (several edges, a
// long-span edge, or no edge at all): X pointer (mat | ordinal << 2) << 4 | M pointer, mat in bits 0-1 and
// ordinal in bits 10-13 -- the layout of the general-form half-word (lane_decode_ptr).
template <int K> struct LAcc {
    unsigned pXM[K];
};
constexpr unsigned LANE_PX_KEEP = ~0x3f0u, LANE_PM_KEEP = ~0x3c03u;
__device__ __forceinline__ unsigned lane_px(unsigned mat, unsigned ord) { return (mat | (ord << 2)) << 4; }
__device__ __forceinline__ unsigned lane_pm(unsigned mat, unsigned ord) { return mat | (ord << 10); }

// One edge of such a site.  The row above is parked on the site's first virtual row (slot n_slots of the warp's
// scratch), so st.X / st.M become the score accumulators and every edge reads its source row from a slot: one
// copy of the code, no extra registers in flight.  Returns true on the site's last virtual row.
template <int K, bool WR, bool SMALLTAB>
__device__ __forceinline__ bool lane_general_vrow(const LaneCtx &c, LState<K> &st, LAcc<K> &acc, const double *ex, int4 vr, double *slots) {
    const double ninf = neg_inf();
    const int info = vr.x, sl = info & VR_STATE_MASK;
    const int rowoff = SMALLTAB ? sl * 16 : sl;
    if (info & VR_FIRST) {
        lane_park<K>(c, st, slots + (long long)c.n_slots * LANE_SLOT_DOUBLES);
#pragma unroll
        for (int k = 0; k < K; ++k) { st.X[k] = ninf; st.M[k] = ninf; acc.pXM[k] = lane_px(NO_MAT, 0) | lane_pm(NO_MAT, 0); }
    }
    if (!(info & VR_NOEDGE)) {
        const double wl = (info & VR_ZERO_W) ? 0.0 : (double)c.l_elogw[vr.y];  // + 0.0 is exact
        const unsigned ord = (unsigned)vr.w >> 16;
        const int slot = (info & VR_REG) ? c.n_slots : (vr.w & 0xffff);
        const double *row = slots + (long long)slot * LANE_SLOT_DOUBLES;
        // source row: column c0-1 in [0], column c0+k in [k+1]; all loads in flight at once (one L2 round trip)
        double sX[K + 1], sY[K + 1], sM[K], sMo[K];
#pragma unroll
        for (int k = 0; k <= K; ++k) { sX[k] = row[k * 128]; sY[k] = row[k * 128 + 32]; }
#pragma unroll
        for (int k = 0; k < K; ++k) { sM[k] = row[k * 128 + 64]; sMo[k] = row[(k + 1) * 128 + 96]; }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            // X: ext, double, open (:2116-2211); the open candidate was formed when the source row was made
            acc_gt(__dadd_rn(sX[k + 1], ex[k]), st.X[k], acc.pXM[k], LANE_PX_KEEP, lane_px(X_MAT, ord));
            acc_gt(__dadd_rn(sY[k + 1], c.open), st.X[k], acc.pXM[k], LANE_PX_KEEP, lane_px(Y_MAT, ord));
            acc_gt(sMo[k], st.X[k], acc.pXM[k], LANE_PX_KEEP, lane_px(M_MAT, ord));
            // M: from M, X, Y (:2029-2112), ((score + log) + wl) + wr
            double mlog, xlog;
            lane_subst<SMALLTAB>(c, rowoff, st.tab[k], mlog, xlog);
            double a = __dadd_rn(__dadd_rn(sM[k], mlog), wl);
            double b = __dadd_rn(__dadd_rn(sX[k], xlog), wl);
            double d = __dadd_rn(__dadd_rn(sY[k], xlog), wl);
            if (WR) { a = __dadd_rn(a, st.wr[k]); b = __dadd_rn(b, st.wr[k]); d = __dadd_rn(d, st.wr[k]); }
            acc_gt(a, st.M[k], acc.pXM[k], LANE_PM_KEEP, lane_pm(M_MAT, ord));
            acc_gt(b, st.M[k], acc.pXM[k], LANE_PM_KEEP, lane_pm(X_MAT, ord));
            acc_gt(d, st.M[k], acc.pXM[k], LANE_PM_KEEP, lane_pm(Y_MAT, ord));
        }
    }
    return (info & VR_LAST) != 0;
}

// per-lane constants of one strip
template <int K, bool WR, bool SMALLTAB>
__device__ __forceinline__ void lane_strip_init(const LaneCtx &c, LState<K> &st, const LaneGeom &g, int c0, const int *r_state,
                                                const float *r_elogw) {
    const double ninf = neg_inf();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = c0 + k;
        const bool v = g.active && j < g.ly;
        st.X[k] = st.Y[k] = st.M[k] = st.Mo[k] = ninf;
        st.wr[k] = (WR && v && j >= 1) ? (double)r_elogw[j - 1] : 0.0;  // chain edge (j-1 -> j), CSR position j-1
        const int sr = (v && j >= 1) ? r_state[j] : 0;
        st.tab[k] = SMALLTAB ? sr * c.fas * 16 : sr * c.fas;
    }
    st.bX = st.bY = st.bM = ninf;
}

// warp-wide OR of a per-lane condition (the CPU test emulation runs one lane at a time: a lane for which the
// condition is false computes the same values on either side of the branch it selects)
__device__ __forceinline__ bool lane_any(bool x) {
#ifdef PG2_HOST_EMU
    return x;
#else
    return __any_sync(0xffffffffu, x);
#endif
}

// X-extension term of the K columns of strip s (viterbi_alignment.cpp:864-868): log_gap_end_ext on the first and
// the last DP column when terminal gaps are cheap, log_gap_ext elsewhere
template <int K>
__device__ __forceinline__ void lane_ext_terms(const LaneCtx &c, const LaneGeom &g, int s, double *ex) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = s * K + k;
        ex[k] = (c.term && (j == 0 || j == g.ly - 1)) ? c.end_ext : c.ext;
    }
}

// A run of n plain interior rows starting at virtual row v (one unit-weight edge from the row above, not row 0,
// not read by the end corner, never parked): the hot loop, nothing but the row body, the boundary hand-over and
// the pointer store.  ex[k]: X-extension term of column k; ring_in / out / dst point at the run's first row.
template <int K, bool SMALLTAB, bool WR>
__device__ __forceinline__ void lane_fast_run(const LaneCtx &c, LState<K> &st, const double *ex, int v, int n, bool first, bool store_ptr,
                                              const double *ring_in, double *out, uint4 *dst) {
    const double ninf = neg_inf();
    constexpr int Q = K / 8;
    // the row states come through a per-thread running pointer, one row ahead and without a test (the row after the run is at
    // worst the padding entry behind the last row program, pg2_engine.cu): as warp-uniform index arithmetic the fetch costs
    // nine issue slots per row (measured with the table loads above: 54.3 -> 53.4 ms per 100 000 reads)
    const int *pinfo = &c.l_vrow[v].x;
#ifndef PG2_HOST_EMU
    asm volatile("" : "+l"(pinfo));
#endif
    int info_n = __ldg(pinfo);
    PG2_LANE_ROW_UNROLL
    for (int r = 0; r < n; ++r) {
        const int sl = info_n & VR_STATE_MASK;
        pinfo += 4;
        info_n = __ldg(pinfo);
        double rX = ninf, rY = ninf, rM = ninf;
        if (!first) { rX = ring_in[r * 96]; rY = ring_in[r * 96 + 32]; rM = ring_in[r * 96 + 64]; }
        unsigned w[K / 2];
        lane_xm_row<K, false, WR, SMALLTAB, true>(c, st, sl, 0.0, ex, w);
        lane_y_chain<K, 4u, 8u>(c, st, c.ext, rX, rY, rM, w);
        if (store_ptr) {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                uint4 o;
                o.x = w[4 * q]; o.y = w[4 * q + 1]; o.z = w[4 * q + 2]; o.w = w[4 * q + 3];
                dst[(r * Q + q) * 32] = o;
            }
        }
        if (out) {
            double *b = out + r * 96;
            b[0] = st.X[K - 1]; b[32] = st.Y[K - 1]; b[64] = st.M[K - 1];
        }
    }
}

// One virtual row of any other kind -- a row of the first strip or of a strip that holds a last column, row 0, a
// row the end corner reads, a row that is parked for long-span edges, a weighted edge, one edge of a multi-edge
// site: the generic body.  Returns true while a multi-edge site stays open.
template <int K, bool GENERAL, bool SMALLTAB, bool WR>
__device__ __forceinline__ bool lane_slow_row(const LaneCtx &c, LState<K> &st, LAcc<K> &acc, const double *ex, const LaneGeom &g, int s, int v, bool first,
                                              bool store_ptr, bool is_last, const double *ring_in, double *out, uint4 *dst, double *slots,
                                              double *endcol) {
    const double ninf = neg_inf();
    constexpr int Q = K / 8;
    const int4 vr = __ldg(c.l_vrow + v);
    const int info = vr.x, i = vr.z;
    const double extY = (c.term && (i == 0 || i == c.lx - 1)) ? c.end_ext : c.ext;
    unsigned w[K / 2];
    if (GENERAL && (info & VR_FAST) != VR_FAST) {
        // one edge of a general site; after its last virtual row the accumulators are row i of the strip
        if (!lane_general_vrow<K, WR, SMALLTAB>(c, st, acc, ex, vr, slots)) return true;
#pragma unroll
        for (int k = 0; k < K; ++k) st.Mo[k] = __dadd_rn(__dadd_rn(st.M[k], c.lng), c.open);
#pragma unroll
        for (int h = 0; h < K / 2; ++h) w[h] = acc.pXM[2 * h] | (acc.pXM[2 * h + 1] << 16);
    } else {
        // one edge from the row above (or row 0): in place, adding the edge's log weight (+ 0.0 is exact)
        const double wl = (!GENERAL || (info & (VR_ZERO_W | VR_NOEDGE))) ? 0.0 : (double)c.l_elogw[vr.y];
        lane_xm_row<K, GENERAL, WR, SMALLTAB, true>(c, st, info & VR_STATE_MASK, wl, ex, w);
        if (first && i == 0) {
            // the start corner M(0,0) = 0 (:725-733), with the open penalty its two gap moves pay
            // (get_log_gap_open_penalty, basic_alignment.h:490-513)
            st.M[0] = 0.0;
            st.Mo[0] = __dadd_rn(__dadd_rn(0.0, c.lng), c.reduced ? 0.0 : c.open);
        }
    }
    double rX = ninf, rY = ninf, rM = ninf;
    if (!first) { rX = ring_in[0]; rY = ring_in[32]; rM = ring_in[64]; }
    lane_y_chain<K, 4u, 8u>(c, st, extY, rX, rY, rM, w);
    if (store_ptr) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            uint4 o;
            o.x = w[4 * q]; o.y = w[4 * q + 1]; o.z = w[4 * q + 2]; o.w = w[4 * q + 3];
            dst[q * 32] = o;
        }
    }
    if (out) { out[0] = st.X[K - 1]; out[32] = st.Y[K - 1]; out[64] = st.M[K - 1]; }
    if (GENERAL) {
        const int slot = (int)((unsigned)info >> VR_SLOT_SHIFT) - 1;
        if (slot >= 0) lane_park<K>(c, st, slots + (long long)slot * LANE_SLOT_DOUBLES);
    }
    if ((info & VR_ENDPRED) && is_last) {
        // rows the end corner reads: the lane's LAST column (ly-1)
        double vx = ninf, vy = ninf, vm = ninf;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (k == g.last_k) { vx = st.X[k]; vy = st.Y[k]; vm = st.M[k]; }
        double *b = endcol + (long long)i * 96;
        b[0] = vx; b[32] = vy; b[64] = vm;
    }
    return false;
}

// One pipeline block of one lane: virtual rows [v0, v1) of strip `s`, as runs of plain rows and single other rows.
// Code size matters here: the SM's instruction cache holds about 2000 instructions, so there is exactly one copy
// of the hot loop and one generic body for every other kind of row.
//   ring_in   left neighbour column of these rows, [row - v0][X,Y,M][lane] (already offset by the lane); unused
//             for the first strip (-inf boundary)
//   ring_out  where this strip's last column goes for the next strip, same shape; wrap_out (global, indexed by
//             virtual row) instead when the next strip belongs to the next round; both null for the last strip
//   plain     bit r set: virtual row v0 + r is a plain interior row (the host builds the masks, build_row_program)
template <int K, bool GENERAL, bool SMALLTAB, bool WR>
__device__ __forceinline__ void lane_block(const LaneCtx &c, LState<K> &st, const LaneGeom &g, int s, int v0, int v1, unsigned plain,
                                           const double *ring_in, double *ring_out, double *wrap_out, double *slots, double *endcol,
                                           uint4 *ptr) {
    const bool first = (s == 0);
    const bool store_ptr = g.active && s * K < g.ly;
    const bool is_last = g.active && s == g.last_strip;
    constexpr int Q = K / 8;
#ifndef PG2_HOST_EMU
    // the next block's row program (one 128-byte line) on its way to L1 while this block computes
    if (v1 < c.nv) asm volatile("prefetch.global.L1 [%0];" ::"l"(c.l_vrow + v1));
#endif
    // only the first strip and a strip that holds some lane's last column have columns with their own
    // X-extension term (warp-uniform test)
    double ex[K];
    if (lane_any(first || is_last)) {
        lane_ext_terms<K>(c, g, s, ex);
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) ex[k] = c.ext;
    }
    double *out = wrap_out ? wrap_out + (long long)v0 * 96 : ring_out;
    uint4 *dst = ptr + ((long long)s * c.nv + v0) * Q * 32;
    // a site whose virtual rows straddle the block boundary keeps its pointer accumulators in the warp's scratch
    LAcc<K> acc;
    unsigned *acc_area = reinterpret_cast<unsigned *>(slots + (long long)(c.n_slots + 1) * LANE_SLOT_DOUBLES);
#pragma unroll
    for (int k = 0; k < K; ++k) acc.pXM[k] = 0;
    if (GENERAL && !(plain & 1u) && !(__ldg(&c.l_vrow[v0].x) & VR_FIRST)) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc.pXM[k] = acc_area[k * 64];
    }
    bool open_site = false;
    int r = 0;
    const int n = v1 - v0;
    while (r < n) {
        const int run = __ffs((int)~(plain >> r)) - 1;  // plain rows from r on
        if (run > 0) {
            lane_fast_run<K, SMALLTAB, WR>(c, st, ex, v0 + r, run, first, store_ptr, ring_in + r * 96, out ? out + r * 96 : nullptr,
                                           dst + r * Q * 32);
            r += run;
        } else {
            open_site = lane_slow_row<K, GENERAL, SMALLTAB, WR>(c, st, acc, ex, g, s, v0 + r, first, store_ptr, is_last, ring_in + r * 96,
                                                                out ? out + r * 96 : nullptr, dst + r * Q * 32, slots, endcol);
            r += 1;
        }
    }
    if (GENERAL && open_site) {
#pragma unroll
        for (int k = 0; k < K; ++k) acc_area[k * 64] = acc.pXM[k];
    }
}

// iterate_bwd_edges_for_end_corner (:1440-1552) with a single right edge (ly-1 -> stop)
__device__ __forceinline__ void lane_end_corner(const LaneCtx &c, const LaneGeom &g, const double *endcol, const float *r_elogw,
                                                DevResult *res) {
    const double ninf = neg_inf();
    double best = ninf;
    unsigned bptr = NO_MAT;
    const int kl0 = c.l_off[c.lx], kl1 = c.l_off[c.lx + 1];
    const double wr = (double)r_elogw[g.ly - 1];  // edge (ly-1 -> ly)
    for (int kl = kl0; kl < kl1; ++kl) {
        const double *v = endcol + (long long)c.l_estart[kl] * 96;
        double s = __dadd_rn(__dadd_rn(__dadd_rn(v[64], c.lng), (double)c.l_elogw[kl]), wr);
        if (s > best) { best = s; bptr = pack_ptr(M_MAT, kl - kl0, 0); }
        s = v[0];  // score_gap_close: + 0
        if (s > best) { best = s; bptr = pack_ptr(X_MAT, kl - kl0, 0); }
        if (kl == kl0) {
            s = endcol[(long long)(c.lx - 1) * 96 + 32];
            if (s > best) { best = s; bptr = pack_ptr(Y_MAT, 0, 0); }
        }
    }
    res->score = best;
    res->end_ptr = bptr;
    res->status = (best == ninf) ? JOB_NO_PATH : JOB_OK;
}

__device__ __forceinline__ void lane_make_ctx(LaneCtx &c, const LaneTask &T, const DevGraph &GL, const DevModel &m, const int *d_off,
                                              const int *d_estart, const float *d_elogw, const int4 *d_vrow, const int *d_vlast) {
    c.l_vrow = d_vrow + GL.vrow_base;
    c.l_vplain = d_vlast + GL.vplain_base;
    c.nv = GL.n_vrows;
    c.l_off = d_off + GL.off_base;
    c.l_estart = d_estart + GL.edge_base;
    c.l_elogw = d_elogw + GL.edge_base;
    c.lx = GL.n_sites - 1;
    c.n_slots = GL.n_slots;
    c.table = m.table;
    c.stab = nullptr;
    c.stab_s = 0;
    c.fas = m.fas;
    c.open = (double)m.open;
    c.ext = (double)m.ext;
    c.end_ext = (double)m.end_ext;
    c.lng = (double)m.lng;
    c.lng2 = (double)__fmul_rn(2.0f, m.lng);
    c.term = !(T.flags & FLAG_NO_TERMINAL_EDGES);
    c.reduced = (T.flags & FLAG_REDUCED) != 0;
}

__device__ __forceinline__ void lane_make_geom(LaneGeom &g, bool active, int ly) {
    g.active = active;
    g.ly = active ? ly : 1;
    g.last_strip = (g.ly - 1) / LANE_K;
    g.last_k = (g.ly - 1) % LANE_K;
}

// Pipeline schedule of one task: warp w handles strips w, w + W, ...; its work items are numbered u = 0, 1, ...;
// item u is block b = u % period of round r = u / period, i.e. of strip r * W + w (items with b >= n_blocks are
// empty: the period is padded so that warp 0 never prefetches a wrap row of the round in progress).
//   warp w > 0 may start item u when warp w-1 has published item u (same block, strip to the left);
//   warp w < W-1 may write ring slot u % D when warp w+1 has published item u - D;
//   warp 0 may prefetch the wrap rows of item u (round >= 1) when warp W-1 has published item u - period.
struct LaneSched {
    int n_strips, rounds, n_blocks, period, items;
};
template <int W> __device__ __forceinline__ LaneSched lane_schedule(int nv, int max_ly) {
    LaneSched s;
    s.n_strips = (max_ly + LANE_K - 1) / LANE_K;
    s.rounds = (s.n_strips + W - 1) / W;
    s.n_blocks = (nv + LANE_B - 1) / LANE_B;
    s.period = s.n_blocks > W + 1 ? s.n_blocks : W + 1;
    s.items = s.rounds * s.period;
    return s;
}

// shared-memory ring: channels 0 .. W-2 (warp c -> warp c+1) are LANE_D blocks deep, channel W-1 (the wrap
// prefetch of warp 0) two blocks; one block = [LANE_B rows][X,Y,M][32 lanes] doubles
constexpr int LANE_BLOCK_DOUBLES = LANE_B * 96;
template <int W> struct LaneRing {
    static constexpr int doubles = ((W - 1) * LANE_D + 2) * LANE_BLOCK_DOUBLES;
    __device__ __forceinline__ static int offset(int channel, int u) {
        return channel < W - 1 ? (channel * LANE_D + u % LANE_D) * LANE_BLOCK_DOUBLES : ((W - 1) * LANE_D + (u & 1)) * LANE_BLOCK_DOUBLES;
    }
};

#ifndef PG2_HOST_EMU
// Progress counters (shared memory, one per pipeline stage).  A stage publishes "item u done" with a release
// store after its lanes have synchronised; a waiting stage polls with acquire loads.  CTA scope is all that
// is needed: producer and consumer run on the same SM.
__device__ __forceinline__ void lane_wait(const int *p, int target) {
    if (target <= 0) return;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    int v;
    for (;;) {
#if defined(PG2_LANE_FENCE_NONE)
        asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
#else
        asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
#endif
        if (v >= target) break;
#ifdef PG2_LANE_SLEEP
        __nanosleep(PG2_LANE_SLEEP);
#endif
    }
}
__device__ __forceinline__ void lane_publish(int *p, int value, int lane) {
    __syncwarp();
    if (lane == 0) {
        const unsigned a = (unsigned)__cvta_generic_to_shared(p);
#if defined(PG2_LANE_FENCE_NONE)
        asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"(a), "r"(value) : "memory");
#else
        asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(a), "r"(value) : "memory");
#endif
    }
}

// CTAs placed on each SM so far (never reset: only its value modulo W is used)
__device__ int g_lane_sm_rotation[256];

// W: warps per CTA = strips of one task in flight.  LANE_W (4, three CTAs per SM) is the throughput shape: a launch with more
// tasks than the chip holds CTAs keeps twelve warps per SM busy whatever W is, and short pipelines waste less of a task's last
// round.  LANE_W_WIDE (10, one CTA per SM) is the latency shape for launches that cannot fill the chip (a shard of a
// strong-scaling run, the trial alignments of a few reads): a task's 19 strips take two rounds instead of five, so the longest
// task -- what such a launch waits for -- ends 2.5 times sooner.
template <int K, int W, bool GENERAL, bool SMALLTAB, bool WR>
__global__ void __launch_bounds__(W * 32, (W == LANE_W ? PG2_LANE_MINB : 1))
lane_fill_kernel(int n_tasks, const LaneTask *tasks, const DevJob *jobs, const DevGraph *graphs, const DevModel *models,
                 const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw, const int4 *d_vrow, const int *d_vlast,
                 unsigned short *ptrs, DevResult *results, double *scratch, long long wrap_doubles, long long endcol_doubles,
                 long long slot_doubles, int *queue) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);
    double2 *s_tab = reinterpret_cast<double2 *>(ring + LaneRing<W>::doubles);
    __shared__ int s_task;
    __shared__ int s_progress[W];
    const int lane = threadIdx.x & 31;
    // pipeline stage of this warp.  Hardware warp k of every CTA lives on SM sub-partition k % 4; rotating the
    // stages by the CTA index puts one warp of each stage on every sub-partition (stage 0 carries the wrap
    // prefetch, the last stage idles in a short last round), so the sub-partitions stay evenly loaded.
#if defined(PG2_LANE_NOROT)
    const int w = threadIdx.x >> 5;
#elif !defined(PG2_LANE_ROT_SM)
    const int w = ((threadIdx.x >> 5) + blockIdx.x) % W;
#else
    // (tuning variant, measured and not adopted: the rotation counted per SM instead of taken from the CTA index, for the case
    // that several launches share the chip -- e2e 70.0 vs 69.4 ms, resident fill 55.2 vs 54.1 ms: the stage index stops being
    // a uniform-datapath value)
    __shared__ int s_rot;
    if (threadIdx.x == 0) {
        unsigned smid;
        asm("mov.u32 %0, %%smid;" : "=r"(smid));
        s_rot = atomicAdd(&g_lane_sm_rotation[smid & 255u], 1);
    }
    __syncthreads();
    const int w = ((threadIdx.x >> 5) + s_rot) % W;
#endif
    double *wrap = scratch + (long long)blockIdx.x * (wrap_doubles + endcol_doubles + W * slot_doubles);
    double *endcol = wrap + wrap_doubles;
    double *slots = endcol + endcol_doubles + (long long)w * slot_doubles + lane;
    int *progress = s_progress;
    int tab_model = -1;

    for (;;) {
        __syncthreads();  // the previous task's end corners are done with endcol; s_task / s_progress may be rewritten
        if (threadIdx.x == 0) s_task = atomicAdd(queue, 1);
        if (threadIdx.x < W) s_progress[threadIdx.x] = 0;
        __syncthreads();
        const int q = s_task;
        if (q >= n_tasks) break;
        const LaneTask &T = tasks[q];
        const bool valid = lane < T.n_jobs;
        const int jid = T.job_ids[valid ? lane : 0];
        const DevJob &J = jobs[jid];
        DevResult *res = results + jid;
        LaneGeom g;
        lane_make_geom(g, valid && res->status == JOB_OK, J.ly);
        const DevGraph GL = graphs[T.left], GR = graphs[J.right];
        const DevModel m = models[T.model];
        LaneCtx c;
        lane_make_ctx(c, T, GL, m, d_off, d_estart, d_elogw, d_vrow, d_vlast);
        // keep the doubles in registers: ptxas otherwise re-derives them from the float model parameters
        // (F2F) inside the row loop whenever registers get tight
        asm volatile("" : "+d"(c.open), "+d"(c.ext), "+d"(c.lng));
        if (SMALLTAB) {
            if (tab_model != T.model) {  // block-uniform
                for (int e = threadIdx.x; e < m.fas * m.fas; e += blockDim.x) {
                    const double ls = (double)m.table[e];
                    s_tab[e] = make_double2(__dadd_rn(c.lng2, ls), __dadd_rn(c.lng, ls));
                }
                tab_model = T.model;
                __syncthreads();
            }
            c.stab = s_tab;
#ifndef PG2_HOST_EMU
            c.stab_s = (unsigned)__cvta_generic_to_shared(s_tab);
#endif
        }
        const int *r_state = d_state + GR.state_base;
        const float *r_elogw = d_elogw + GR.edge_base;
        uint4 *ptr = reinterpret_cast<uint4 *>(ptrs + T.ptr_base) + lane;
        const LaneSched sch = lane_schedule<W>(c.nv, T.max_ly);
        LState<K> st;
        lane_strip_init<K, WR, SMALLTAB>(c, st, g, 0, r_state, r_elogw);

#ifdef PG2_LANE_TIMING
        long long t_in = 0, t_out = 0, t_blk = 0, t_pre = 0, t_all = clock64();
#define PG2_T0 const long long tt0_ = clock64()
#define PG2_T1(acc) acc += clock64() - tt0_
#else
#define PG2_T0
#define PG2_T1(acc)
#endif
        int plain_n = __ldg(c.l_vplain);  // the next item's row mask, fetched one item ahead
        for (int u = 0; u < sch.items; ++u) {
            const int r = u / sch.period, b = u - r * sch.period, s = r * W + w;
            const unsigned plain = (unsigned)plain_n;
            {
                const int b1 = (b + 1 == sch.period) ? 0 : b + 1;
                if (b1 < sch.n_blocks) plain_n = __ldg(c.l_vplain + b1);
            }
            if (w == 0) {
                // start bringing the NEXT item's wrap rows (written by warp W-1 one round earlier) into the ring;
                // the copy lands while this item is being computed
                const int u1 = u + 1, r1 = u1 / sch.period, b1 = u1 - r1 * sch.period;
                if (r1 >= 1 && r1 < sch.rounds && b1 < sch.n_blocks) {
                    lane_wait(progress + W - 1, u1 - sch.period + 1);
                    const int v0 = b1 * LANE_B, v1 = min(v0 + LANE_B, c.nv);
                    double *dst = ring + LaneRing<W>::offset(W - 1, u1) + lane;
                    const double *src = wrap + (long long)v0 * 96 + lane;
                    for (int e = 0; e < (v1 - v0) * 3; ++e) {
                        const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + e * 32);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src + e * 32) : "memory");
                    }
                }
            }
            if (s < sch.n_strips && b < sch.n_blocks) {
                const bool has_next = s + 1 < sch.n_strips;
                { PG2_T0; if (w > 0) lane_wait(progress + w - 1, u + 1); PG2_T1(t_in); }        // input block published
                { PG2_T0; if (has_next && w < W - 1) lane_wait(progress + w + 1, u - LANE_D + 1); PG2_T1(t_out); }  // output slot drained
                if (b == 0) lane_strip_init<K, WR, SMALLTAB>(c, st, g, s * K, r_state, r_elogw);
                const int v0 = b * LANE_B, v1 = min(v0 + LANE_B, c.nv);
                const double *ring_in = ring + LaneRing<W>::offset((w + W - 1) % W, u) + lane;
                double *ring_out = (has_next && w != W - 1) ? ring + LaneRing<W>::offset(w, u) + lane : nullptr;
                double *wrap_out = (has_next && w == W - 1) ? wrap + lane : nullptr;
                { PG2_T0; lane_block<K, GENERAL, SMALLTAB, WR>(c, st, g, s, v0, v1, plain, ring_in, ring_out, wrap_out, slots, endcol + lane, ptr); PG2_T1(t_blk); }
            }
            { PG2_T0; if (w == 0) asm volatile("cp.async.wait_all;" ::: "memory"); PG2_T1(t_pre); }
            lane_publish(progress + w, u + 1, lane);
        }
#ifdef PG2_LANE_TIMING
        if (lane == 0 && (blockIdx.x % 97) == 5 && q < 2000)
            printf("task %d cta %d stage %d: total %lld in_wait %lld out_wait %lld block %lld cpwait %lld items %d\n", q, blockIdx.x, w,
                   clock64() - t_all, t_in, t_out, t_blk, t_pre, sch.items);
#endif
        __syncthreads();
        if (w == 0 && g.active) lane_end_corner(c, g, endcol + lane, r_elogw, res);
    }
}
#endif

// CTAs of one SM: three of LANE_W warps, one of LANE_W_WIDE
int lane_ctas_per_sm(int W) { return W == LANE_W ? PG2_LANE_MINB : 1; }

#ifdef PG2_HOST_EMU
// CPU test emulation of ONE CTA: the same items and block bodies; warp w runs item t - w at step t, warps and
// lanes one after the other -- one of the interleavings the progress counters admit (a producer's ring slot
// is read one step after it was written and rewritten LANE_D steps later).
template <int W>
static void lane_emulate(int variant, int n_tasks, const LaneTask *tasks, const DevJob *jobs, const DevGraph *graphs, const DevModel *models,
                         const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw, const int4 *d_vrow,
                         const int *d_vlast, unsigned short *ptrs, DevResult *results, double *scratch, long long wrap_doubles,
                         long long endcol_doubles, long long slot_doubles) {
    constexpr int K = LANE_K;
    double *wrap = scratch, *endcol = wrap + wrap_doubles, *slots0 = endcol + endcol_doubles;
    std::vector<double> ring((size_t)LaneRing<W>::doubles);
    for (int q = 0; q < n_tasks; ++q) {
        const LaneTask &T = tasks[q];
        const DevGraph GL = graphs[T.left];
        const DevModel m = models[T.model];
        LaneCtx c;
        lane_make_ctx(c, T, GL, m, d_off, d_estart, d_elogw, d_vrow, d_vlast);
        std::vector<double2> tab;
        if (variant & 2) {
            tab.resize((size_t)m.fas * m.fas);
            for (int e = 0; e < m.fas * m.fas; ++e) {
                const double ls = (double)m.table[e];
                tab[e] = make_double2(c.lng2 + ls, c.lng + ls);
            }
            c.stab = tab.data();
        }
        const LaneSched sch = lane_schedule<W>(c.nv, T.max_ly);
        std::vector<LState<K> > st((size_t)W * 32);
        LaneGeom geom[32];
        const int *r_state[32];
        const float *r_elogw[32];
        DevResult *res[32];
        for (int lane = 0; lane < 32; ++lane) {
            const bool valid = lane < T.n_jobs;
            const int jid = T.job_ids[valid ? lane : 0];
            const DevJob &J = jobs[jid];
            res[lane] = results + jid;
            lane_make_geom(geom[lane], valid && res[lane]->status == JOB_OK, J.ly);
            const DevGraph GR = graphs[J.right];
            r_state[lane] = d_state + GR.state_base;
            r_elogw[lane] = d_elogw + GR.edge_base;
        }
        for (int t = 0; t < sch.items + W - 1; ++t) {
            for (int w = 0; w < W; ++w) {
                const int u = t - w;
                if (u < 0 || u >= sch.items) continue;
                const int r = u / sch.period, b = u - r * sch.period, s = r * W + w;
                if (s < sch.n_strips && b < sch.n_blocks) {
                    const int v0 = b * LANE_B, v1 = v0 + LANE_B < c.nv ? v0 + LANE_B : c.nv;
                    const bool has_next = s + 1 < sch.n_strips;
                    for (int lane = 0; lane < 32; ++lane) {
                        LState<K> &S = st[(size_t)w * 32 + lane];
#define PG2_LANE_EMU(G, SM, WRV)                                                                                      \
    do {                                                                                                               \
        if (b == 0) lane_strip_init<K, WRV, SM>(c, S, geom[lane], s * K, r_state[lane], r_elogw[lane]);                \
        lane_block<K, G, SM, WRV>(c, S, geom[lane], s, v0, v1, (unsigned)c.l_vplain[b],                                                    \
                                  ring.data() + LaneRing<W>::offset((w + W - 1) % W, u) + lane,                 \
                                  (has_next && w != W - 1) ? ring.data() + LaneRing<W>::offset(w, u) + lane : nullptr, \
                                  (has_next && w == W - 1) ? wrap + lane : nullptr,                               \
                                  slots0 + (long long)w * slot_doubles + lane, endcol + lane,                          \
                                  reinterpret_cast<uint4 *>(ptrs + T.ptr_base) + lane);                                \
    } while (0)
                        switch (variant & 7) {
                            case 0: PG2_LANE_EMU(false, false, false); break;
                            case 1: PG2_LANE_EMU(true, false, false); break;
                            case 2: PG2_LANE_EMU(false, true, false); break;
                            case 3: PG2_LANE_EMU(true, true, false); break;
                            case 4: PG2_LANE_EMU(false, false, true); break;
                            case 5: PG2_LANE_EMU(true, false, true); break;
                            case 6: PG2_LANE_EMU(false, true, true); break;
                            case 7: PG2_LANE_EMU(true, true, true); break;
                        }
#undef PG2_LANE_EMU
                    }
                }
                if (w == 0) {
                    const int u1 = u + 1, r1 = u1 / sch.period, b1 = u1 - r1 * sch.period;
                    if (r1 >= 1 && r1 < sch.rounds && b1 < sch.n_blocks) {
                        const int v0 = b1 * LANE_B, v1 = v0 + LANE_B < c.nv ? v0 + LANE_B : c.nv;
                        memcpy(ring.data() + LaneRing<W>::offset(W - 1, u1), wrap + (long long)v0 * 96,
                               sizeof(double) * (size_t)(v1 - v0) * 96);
                    }
                }
            }
        }
        for (int lane = 0; lane < 32; ++lane)
            if (geom[lane].active) lane_end_corner(c, geom[lane], endcol + lane, r_elogw[lane], res[lane]);
    }
}
#else
template <int W>
static void lane_launch(int variant, int n_tasks, const LaneTask *tasks, const DevJob *jobs, const DevGraph *graphs, const DevModel *models,
                        const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw, const int4 *d_vrow,
                        const int *d_vlast, unsigned short *ptrs, DevResult *results, double *scratch, long long wrap_doubles,
                        long long endcol_doubles, long long slot_doubles, int *queue, int n_ctas, cudaStream_t stream) {
    constexpr int K = LANE_K;
    const int smem = LaneRing<W>::doubles * (int)sizeof(double) + ((variant & 2) ? STRIP_SMALL_FAS * STRIP_SMALL_FAS * (int)sizeof(double2) : 0);
#define PG2_LANE_LAUNCH(G, S, WRV)                                                                                                \
    do {                                                                                                                          \
        cudaFuncSetAttribute(lane_fill_kernel<K, W, G, S, WRV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);               \
        lane_fill_kernel<K, W, G, S, WRV><<<n_ctas, W * 32, smem, stream>>>(n_tasks, tasks, jobs, graphs, models, d_state, d_off, \
                                                                            d_estart, d_elogw, d_vrow, d_vlast, ptrs, results,    \
                                                                            scratch,                                              \
                                                                            wrap_doubles, endcol_doubles, slot_doubles, queue);   \
    } while (0)
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
}
#endif

// Launches one group of lane tasks that share the kernel variant, W (LANE_W or LANE_W_WIDE) warps per CTA.
// `scratch`: n_ctas * lane_cta_doubles(.., W) doubles.
void launch_lane_fill(int variant, int W, int n_tasks, const LaneTask *tasks, const DevJob *jobs, const DevGraph *graphs,
                      const DevModel *models, const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw,
                      const int4 *d_vrow, const int *d_vlast, unsigned short *ptrs, DevResult *results, double *scratch, int max_nv,
                      int max_lx, int max_slots, int *queue, int n_ctas, cudaStream_t stream) {
    if (n_tasks <= 0) return;
    const long long wrap_doubles = (long long)max_nv * 96, endcol_doubles = (long long)max_lx * 96;
    const long long slot_doubles = (long long)(max_slots + 2) * LANE_SLOT_DOUBLES;
#ifndef PG2_HOST_EMU
    cudaMemsetAsync(queue, 0, sizeof(int), stream);
    if (W == LANE_W_WIDE)
        lane_launch<LANE_W_WIDE>(variant, n_tasks, tasks, jobs, graphs, models, d_state, d_off, d_estart, d_elogw, d_vrow, d_vlast, ptrs, results,
                                 scratch, wrap_doubles, endcol_doubles, slot_doubles, queue, n_ctas, stream);
    else
        lane_launch<LANE_W>(variant, n_tasks, tasks, jobs, graphs, models, d_state, d_off, d_estart, d_elogw, d_vrow, d_vlast, ptrs, results,
                            scratch, wrap_doubles, endcol_doubles, slot_doubles, queue, n_ctas, stream);
#else
    (void)queue; (void)n_ctas; (void)stream;
    if (W == LANE_W_WIDE)
        lane_emulate<LANE_W_WIDE>(variant, n_tasks, tasks, jobs, graphs, models, d_state, d_off, d_estart, d_elogw, d_vrow, d_vlast, ptrs, results,
                                  scratch, wrap_doubles, endcol_doubles, slot_doubles);
    else
        lane_emulate<LANE_W>(variant, n_tasks, tasks, jobs, graphs, models, d_state, d_off, d_estart, d_elogw, d_vrow, d_vlast, ptrs, results,
                             scratch, wrap_doubles, endcol_doubles, slot_doubles);
#endif
}

}  // namespace pg2
