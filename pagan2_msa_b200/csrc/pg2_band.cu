// pg2_band.cu -- anchored alignments of two plain unit-weight chains (leaf x leaf inside the reference's anchor band,
// BASELINE config 5: 200 kb x 200 kb, ~48 in-band cells per row): ONE WARP per alignment, no CTA barrier.
//
// The band of such a job is 400 000 anti-diagonals of ~24 cells.  Cells of one anti-diagonal are independent; the work
// per diagonal is tiny, so what bounds the fill is the LATENCY of one diagonal step.  The wavefront kernel spends a CTA
// barrier, a global geometry fetch and ~40 index instructions per step (375 ns); here a step is
//   * one pass of the warp over the rows of the diagonal (lane = row - first row, further passes when the diagonal is longer
//     than 32 cells), scores of the two previous diagonals in shared-memory rings indexed by ROW (no offset arithmetic),
//   * M of the NEXT diagonal computed one step ahead (its sources are complete one step earlier), so the chain that links
//     two steps is ring load -> DADD -> two first-wins updates -> ring store,
//   * band geometry, row and column states staged into shared-memory rings by cp.async two chunks of 32 steps ahead,
//   * one pointer BYTE per cell (three 2-bit source matrices), row-major inside the band.
// The walk over those bytes is parallel as well (the path of a 200 kb pair has 400 000 steps): the diagonals are cut into
// segments of BAND_SEG; for every segment every possible entry state (cell on its top two diagonals x matrix) is walked to
// the segment's lower edge by its own thread, a per-job thread then links the segments the real path visits, and the
// segments emit their run-length encoded words side by side.
//
// Restates compute_fwd_scores / iterate_bwd_edges_for_gap / iterate_bwd_edges_for_match / iterate_bwd_edges_for_end_corner
// and backtrack_new_path for in-degree-1 graphs (reference src/main/viterbi_alignment.cpp:856-971, 1038-1189, 1328-1552,
// 2029-2255; band semantics utils/tunnel_matrix.h:85-98).  Candidate order, FP64 association and strict '>' as in the
// other kernels (pg2_wavefront.cu: wave_cell_plain_core is the same cell, one thread per cell and a barrier per diagonal).
#include "pg2_device.cuh"
#include "pg2_strip_geom.cuh"
#ifdef PG2_HOST_EMU
#include <vector>
#endif

namespace pg2 {

constexpr int BAND_R = 512;   // row ring: rows live around a diagonal, rows staged ahead (BAND_MAX_DIAG + 2 BAND_CHUNK + 8 + BAND_THREADS < BAND_R)
constexpr int BAND_C = 1024;  // column ring
constexpr int BAND_G = 256;   // geometry ring: diagonals s0-48 .. s0+159
constexpr int BAND_CHUNK = 48;  // steps between two staging points; a multiple of 3, so that the ring slots of a step are constants
constexpr int BAND_RM = BAND_R - 1, BAND_CM = BAND_C - 1, BAND_GM = BAND_G - 1;
constexpr int BAND_THREADS = 96;  // one warp per matrix: X, Y, M

struct BandSm {
    double *x, *y;  // [3][BAND_R] X, Y of the row's cell on diagonals s, s-1, s-2 (slots rotate)
    double *m;      // [3][BAND_R] written at step s: M of the row's cell on diagonal s+1
    int *rstate;    // [BAND_R] state of the row's site
    int *rroff;     // [BAND_R] byte offset of the row's pointers minus its first in-band column
    int *cstate;    // [BAND_C] state of the column's site
    int *geo;       // [BAND_G][2] first / last in-band row of a diagonal
    double2 *stab;  // [fas * fas] {2 lng + ls, lng + ls} (SMALLTAB)
};

struct BandConst {
    double open, ext, end_ext, lng, lng2;
    const float *table;
    int fas, lx, ly;
    bool term, reduced;
};

struct BandStep {
    int s, lo0, hi0, lo1, hi1;  // diagonal, its rows, the rows of diagonal s+1
    int b0, b1, b2;             // ring offsets of diagonals s, s-1, s-2
};

__host__ __device__ inline size_t band_smem_bytes() {
    return (size_t)3 * 3 * BAND_R * 8 + (size_t)BAND_R * 8 + (size_t)BAND_C * 4 + (size_t)BAND_G * 8 +
           (size_t)STRIP_SMALL_FAS * STRIP_SMALL_FAS * 16;
}
__host__ __device__ inline void band_carve(BandSm &sm, unsigned char *base) {
    sm.stab = reinterpret_cast<double2 *>(base); base += (size_t)STRIP_SMALL_FAS * STRIP_SMALL_FAS * 16;
    sm.x = reinterpret_cast<double *>(base); base += (size_t)3 * BAND_R * 8;
    sm.y = reinterpret_cast<double *>(base); base += (size_t)3 * BAND_R * 8;
    sm.m = reinterpret_cast<double *>(base); base += (size_t)3 * BAND_R * 8;
    sm.geo = reinterpret_cast<int *>(base); base += (size_t)BAND_G * 8;
    sm.rstate = reinterpret_cast<int *>(base); base += (size_t)BAND_R * 4;
    sm.rroff = reinterpret_cast<int *>(base); base += (size_t)BAND_R * 4;
    sm.cstate = reinterpret_cast<int *>(base);
}

// first-wins running maximum (basic_alignment.h:449-462)
__device__ __forceinline__ void band_cand(double s, unsigned code, double &best, unsigned &ptr) {
    const bool p = s > best;
    best = p ? s : best;
    ptr = p ? code : ptr;
}
// The first-wins maximum of three candidates in their order (c1, c2, c3), two selects deep instead of three: the later
// two are folded first (a tie keeps c2), the result replaces c1 only when strictly greater (a tie keeps c1) -- the same
// winner as the sequential scan; no candidate above -inf leaves the cell without a pointer.  The candidates must not be NaN
// (a NaN never wins the reference's `>`, but it would get through the fold): the reference's tables hold NaN for state pairs
// it never scores, band_ls() turns them into -inf, which never wins either.
__device__ __forceinline__ double band_max3(double c1, double c2, double c3, unsigned p1, unsigned p2, unsigned p3, unsigned &ptr) {
    const bool q = c3 > c2;
    const double g = q ? c3 : c2;
    const unsigned pg = q ? p3 : p2;
    const bool w = g > c1;
    const double best = w ? g : c1;
    ptr = w ? pg : p1;
    ptr = (best == neg_inf()) ? (unsigned)NO_MAT : ptr;
    return best;
}
// a pointer byte, written when the cell lies inside the band (a predicated store: no branch in the step)
__device__ __forceinline__ void band_store_ptr(unsigned char *plane, int off, unsigned p, bool v) {
#ifdef PG2_HOST_EMU
    if (v) plane[off] = (unsigned char)p;
#else
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.u8 [%0], %1;\n\t}" ::"l"(plane + off), "r"(p), "r"((unsigned)v) : "memory");
#endif
}

// One row of step s, one matrix.  The step writes the rows [lo0 - 1, hi0 + 3] of every ring slot: the computed value where
// the cell lies inside the band -- X, Y of cell (i, s - i) for lo0 <= i <= hi0, M of cell (i, s + 1 - i) for lo1 <= i <= hi1
// (its sources are the cells of diagonal s - 1, complete since the previous step) -- and -inf elsewhere.  Every ring entry a
// later step reads was written by the step it belongs to, so reads need no band test (Tunnel_slice::at returns -inf outside
// the band).  Rows past the range (the tail of the last pass) write -inf to entries nothing reads before they are rewritten.
// The body has no branch: a step is a chain of dependent instructions of ONE warp, and a divergent branch costs it more
// than the work it skips.
//
// X of (i, j): ext, double, open out of (i-1, j)   (:2116-2211)
template <bool BOUNDARY>
__device__ __forceinline__ void band_x_cell(const BandSm &sm, const BandConst &k, const BandStep &t, int i, unsigned char *PX) {
    const int r = i & BAND_RM, ru = (i - 1) & BAND_RM, j = t.s - i;
    const int off = sm.rroff[r] + j;
    const double a = sm.x[t.b1 + ru], b = sm.y[t.b1 + ru], c = sm.m[t.b2 + ru];
    double extx = k.ext, pen = k.open;
    if (BOUNDARY) {
        if (k.term && (j == 0 || j == k.ly - 1)) extx = k.end_ext;  // terminal gap extension (:864-868)
        if (k.reduced && i == 1) pen = 0.0;                         // get_log_gap_open_penalty (basic_alignment.h:490-513)
    }
    unsigned p;
    const double best = band_max3(__dadd_rn(a, extx), __dadd_rn(b, k.open), __dadd_rn(__dadd_rn(c, k.lng), pen), X_MAT, Y_MAT, M_MAT, p);
    const bool v = i >= t.lo0 && i <= t.hi0;
    sm.x[t.b0 + r] = v ? best : neg_inf();
    band_store_ptr(PX, off, p, v);
}
// Y of (i, j): ext, double, open out of (i, j-1)
template <bool BOUNDARY>
__device__ __forceinline__ void band_y_cell(const BandSm &sm, const BandConst &k, const BandStep &t, int i, unsigned char *PY) {
    const int r = i & BAND_RM, j = t.s - i;
    const int off = sm.rroff[r] + j;
    const double a = sm.y[t.b1 + r], b = sm.x[t.b1 + r], c = sm.m[t.b2 + r];
    double exty = k.ext, pen = k.open;
    if (BOUNDARY) {
        if (k.term && (i == 0 || i == k.lx - 1)) exty = k.end_ext;  // (:875-879)
        if (k.reduced && j == 1) pen = 0.0;
    }
    unsigned p;
    const double best = band_max3(__dadd_rn(a, exty), __dadd_rn(b, k.open), __dadd_rn(__dadd_rn(c, k.lng), pen), Y_MAT, X_MAT, M_MAT, p);
    const bool v = i >= t.lo0 && i <= t.hi0;
    sm.y[t.b0 + r] = v ? best : neg_inf();
    band_store_ptr(PY, off, p, v);
}
__device__ __forceinline__ double band_ls(float v) { return v != v ? neg_inf() : (double)v; }
// the substitution terms of cell (i, s + 1 - i)
template <bool SMALLTAB>
__device__ __forceinline__ double2 band_subst(const BandSm &sm, const BandConst &k, int s, int i) {
    const int sl = sm.rstate[i & BAND_RM], sr = sm.cstate[(s + 1 - i) & BAND_CM];
    if (SMALLTAB) return sm.stab[sl + sr * k.fas];
    const double ls = band_ls(__ldg(k.table + (size_t)sl + (size_t)sr * (size_t)k.fas));
    return make_double2(__dadd_rn(k.lng2, ls), __dadd_rn(k.lng, ls));
}
// M of (i, j+1): from M, X, Y of (i-1, j)   (:2029-2112)
__device__ __forceinline__ void band_m_cell(const BandSm &sm, const BandStep &t, int i, double2 sub, unsigned char *PM) {
    const int r = i & BAND_RM, ru = (i - 1) & BAND_RM, j = t.s - i;
    const int off = sm.rroff[r] + j + 1;
    const double a = sm.m[t.b2 + ru], b = sm.x[t.b1 + ru], c = sm.y[t.b1 + ru];
    unsigned p;
    const double best = band_max3(__dadd_rn(a, sub.x), __dadd_rn(b, sub.y), __dadd_rn(c, sub.y), M_MAT, X_MAT, Y_MAT, p);
    const bool v = i >= t.lo1 && i <= t.hi1;
    sm.m[t.b0 + r] = v ? best : neg_inf();
    band_store_ptr(PM, off, p, v);
}

// layout of a job's record in d_band4 (ints): [2 (nd + 2)] first / last row of every diagonal (+ two closing entries),
// [lx] pointer byte offset of every row minus its first in-band column, [n_seg + 1] first candidate of every walk segment
__host__ __device__ inline long long band_geo_off(const DevJob &J) { return J.b4_base; }
__host__ __device__ inline long long band_roff_off(const DevJob &J) { return J.b4_base + 2LL * (J.lx + J.ly - 1 + 2); }
__host__ __device__ inline long long band_seg_off(const DevJob &J) { return band_roff_off(J) + J.lx; }
// the pointers: three planes (X, Y, M) of one byte per in-band cell, row-major; a job's region holds J.ptr_cells words
__host__ __device__ inline long long band_plane_bytes(const DevJob &J) { return J.ptr_cells / 3 * 4; }

__device__ __forceinline__ void band_make_const(BandConst &k, const DevJob &J, const DevModel &m) {
    k.open = (double)m.open; k.ext = (double)m.ext; k.end_ext = (double)m.end_ext; k.lng = (double)m.lng;
    k.lng2 = (double)__fmul_rn(2.0f, m.lng);  // 2*model->log_non_gap() stays float (:1364)
    k.table = m.table; k.fas = m.fas; k.lx = J.lx; k.ly = J.ly;
    k.term = !(J.flags & FLAG_NO_TERMINAL_EDGES);
    k.reduced = (J.flags & FLAG_REDUCED) != 0;
}

// initialise_array_corner (:725-733) and the pointers no step writes: (0,0), and M of the cells of diagonal 1
__device__ __forceinline__ void band_init_corner(const BandSm &sm, const DevJob &J, const int *g_geo, const int *g_roff, unsigned char *P8) {
    const long long plane = band_plane_bytes(J);
    sm.m[2 * BAND_R + 0] = 0.0;  // M(0,0): "diagonal -1" sits in slot 2 (written a step ahead, like every M)
    for (int q = 0; q < 3; ++q) P8[q * plane + g_roff[0]] = (unsigned char)NO_MAT;
    if (J.lx + J.ly - 1 > 1)
        for (int i = g_geo[2]; i <= g_geo[3]; ++i) P8[2 * plane + g_roff[i] + (1 - i)] = (unsigned char)NO_MAT;
}

// the last cell -> Viterbi score and end pointer (iterate_bwd_edges_for_end_corner :1440-1552, one edge on each side)
__device__ __forceinline__ void band_end_corner(const BandConst &k, double X, double Y, double M, DevResult *res) {
    double best = neg_inf();
    unsigned ptr = NO_MAT;
    band_cand(__dadd_rn(M, k.lng), pack_ptr(M_MAT, 0, 0), best, ptr);
    band_cand(X, pack_ptr(X_MAT, 0, 0), best, ptr);
    band_cand(Y, pack_ptr(Y_MAT, 0, 0), best, ptr);
    res->score = best;
    res->end_ptr = ptr;
    res->status = (best == neg_inf()) ? JOB_NO_PATH : JOB_OK;
}

#ifndef PG2_HOST_EMU
__device__ __forceinline__ void band_cp4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void band_cp8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// One step of one warp (ROLE 0 X, 1 Y, 2 M); B0, B1, B2: the ring slots of diagonals s, s-1, s-2 -- constants, the step
// loop is unrolled three times.  Straight-line code: the first pass of the warp over the diagonal's rows, a rarely taken loop
// for diagonals longer than a warp, the CTA barrier.  The geometry of diagonal s + 2 is fetched at step s; the M warp fetches
// the substitution terms of its first pass one step ahead (two dependent shared-memory loads that do not wait for the other
// warps' results).
template <int ROLE, bool SMALLTAB, bool BOUNDARY, int B0, int B1, int B2>
__device__ __forceinline__ void band_step(const BandSm &sm, const BandConst &k, BandStep &t, int s, int lane, int &lo2, int &hi2,
                                          double2 &sub_next, unsigned char *PQ) {
    t.s = s; t.lo0 = t.lo1; t.hi0 = t.hi1; t.lo1 = lo2; t.hi1 = hi2;
    lo2 = sm.geo[2 * ((s + 2) & BAND_GM)];
    hi2 = sm.geo[2 * ((s + 2) & BAND_GM) + 1];
    t.b0 = B0 * BAND_R; t.b1 = B1 * BAND_R; t.b2 = B2 * BAND_R;
    const int i0 = t.lo0 - 1 + lane;
    if (ROLE == 2) {
        const double2 sub = sub_next;
        sub_next = band_subst<SMALLTAB>(sm, k, s + 1, t.lo1 - 1 + lane);
        band_m_cell(sm, t, i0, sub, PQ);
    } else if (ROLE == 0) band_x_cell<BOUNDARY>(sm, k, t, i0, PQ);
    else band_y_cell<BOUNDARY>(sm, k, t, i0, PQ);
    if (t.hi0 - t.lo0 + 4 >= 32) {  // rows lo0-1 .. hi0+3 do not fit one pass
        for (int i = i0 + 32; i <= t.hi0 + 3 + lane; i += 32) {
            if (ROLE == 2) band_m_cell(sm, t, i, band_subst<SMALLTAB>(sm, k, s, i), PQ);
            else if (ROLE == 0) band_x_cell<BOUNDARY>(sm, k, t, i, PQ);
            else band_y_cell<BOUNDARY>(sm, k, t, i, PQ);
        }
    }
    __syncthreads();
}
// the steps s_first .. s_first + n - 1 of a chunk; s_first = 1 (mod 3): step s writes slot s % 3
template <int ROLE, bool SMALLTAB, bool BOUNDARY>
__device__ __forceinline__ void band_run_chunk(const BandSm &sm, const BandConst &k, int s_first, int n, int lane, unsigned char *PQ) {
    BandStep t;
    t.lo1 = sm.geo[2 * (s_first & BAND_GM)];
    t.hi1 = sm.geo[2 * (s_first & BAND_GM) + 1];
    int lo2 = sm.geo[2 * ((s_first + 1) & BAND_GM)], hi2 = sm.geo[2 * ((s_first + 1) & BAND_GM) + 1];
    double2 sub_next = make_double2(0.0, 0.0);
    if (ROLE == 2) sub_next = band_subst<SMALLTAB>(sm, k, s_first, t.lo1 - 1 + lane);
    int s = s_first;
    for (; n >= 3; n -= 3, s += 3) {
        band_step<ROLE, SMALLTAB, BOUNDARY, 1, 0, 2>(sm, k, t, s, lane, lo2, hi2, sub_next, PQ);
        band_step<ROLE, SMALLTAB, BOUNDARY, 2, 1, 0>(sm, k, t, s + 1, lane, lo2, hi2, sub_next, PQ);
        band_step<ROLE, SMALLTAB, BOUNDARY, 0, 2, 1>(sm, k, t, s + 2, lane, lo2, hi2, sub_next, PQ);
    }
    if (n >= 1) band_step<ROLE, SMALLTAB, BOUNDARY, 1, 0, 2>(sm, k, t, s, lane, lo2, hi2, sub_next, PQ);
    if (n >= 2) band_step<ROLE, SMALLTAB, BOUNDARY, 2, 1, 0>(sm, k, t, s + 1, lane, lo2, hi2, sub_next, PQ);
}

template <bool SMALLTAB>
__global__ void __launch_bounds__(BAND_THREADS, 1)
band_fill_kernel(const DevJob *jobs, const int *job_ids, const DevGraph *graphs, const DevModel *models, const int *d_state,
                 const int *d_band4, unsigned *ptrs, DevResult *results) {
    extern __shared__ __align__(16) unsigned char band_smem[];
    const int jid = job_ids[blockIdx.x];
    const DevJob J = jobs[jid];
    DevResult *res = results + jid;
    if (res->status != JOB_OK) return;  // rejected by the validation kernel (CTA-uniform)
    const int tid = threadIdx.x, lane = tid & 31, role = tid >> 5;  // role: 0 X, 1 Y, 2 M
    const DevModel m = models[J.model];
    BandSm sm;
    band_carve(sm, band_smem);
    BandConst k;
    band_make_const(k, J, m);
    const int *l_state = d_state + graphs[J.left].state_base, *r_state = d_state + graphs[J.right].state_base;
    const int *g_geo = d_band4 + band_geo_off(J), *g_roff = d_band4 + band_roff_off(J);
    unsigned char *P8 = reinterpret_cast<unsigned char *>(ptrs + J.cell_base);
    unsigned char *PQ = P8 + role * band_plane_bytes(J);  // this warp's pointer plane
    const int nd = J.lx + J.ly - 1;
    const double ninf = neg_inf();

    for (int e = tid; e < 3 * BAND_R; e += BAND_THREADS) { sm.x[e] = ninf; sm.y[e] = ninf; sm.m[e] = ninf; }
    for (int e = tid; e < BAND_R; e += BAND_THREADS) { sm.rstate[e] = 0; sm.rroff[e] = 0; }
    for (int e = tid; e < BAND_C; e += BAND_THREADS) sm.cstate[e] = 0;
    if (SMALLTAB)
        for (int e = tid; e < m.fas * m.fas; e += BAND_THREADS) {
            const double ls = band_ls(m.table[e]);
            sm.stab[e] = make_double2(__dadd_rn(k.lng2, ls), __dadd_rn(k.lng, ls));
        }
    // the geometry of diagonals 0 .. 2 BAND_CHUNK + 15
    for (int t = tid; t < 2 * BAND_CHUNK + 16 && t < nd + 2; t += BAND_THREADS) band_cp8(sm.geo + 2 * (t & BAND_GM), g_geo + 2 * t);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (tid == 0) band_init_corner(sm, J, g_geo, g_roff, P8);
    int fr = 0, fc = 0;  // rows / columns staged so far
    // chunk c: steps s0 + 1 .. s0 + BAND_CHUNK, s0 = c BAND_CHUNK (the last step is nd - 1)
    for (int s0 = 0; s0 + 1 < nd; s0 += BAND_CHUNK) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        // staging, two chunks ahead: the geometry of diagonals s0 + 2 CHUNK + 16 .., the rows up to hi(s0) + 2 CHUNK + 8, the
        // columns up to s0 + 2 CHUNK + 8 - lo(s0) (a diagonal moves its row range by at most one row per step)
        const int lo_s0 = sm.geo[2 * (s0 & BAND_GM)], hi_s0 = sm.geo[2 * (s0 & BAND_GM) + 1];
        {
            const int t = s0 + 2 * BAND_CHUNK + 16 + tid;
            if (tid < BAND_CHUNK && t < nd + 2) band_cp8(sm.geo + 2 * (t & BAND_GM), g_geo + 2 * t);
            const int row_target = hi_s0 + 2 * BAND_CHUNK + 8, col_target = s0 + 2 * BAND_CHUNK + 8 - lo_s0;
            for (; fr < row_target && fr < J.lx; fr += BAND_THREADS) {
                const int i = fr + tid;
                // (the start sites carry no state of the alphabet: row 0 and column 0 keep the ring's 0 -- their M is -inf whatever the term)
                if (i < J.lx) { if (i > 0) band_cp4(sm.rstate + (i & BAND_RM), l_state + i); band_cp4(sm.rroff + (i & BAND_RM), g_roff + i); }
            }
            for (; fc < col_target && fc < J.ly; fc += BAND_THREADS) {
                const int j = fc + tid;
                if (j > 0 && j < J.ly) band_cp4(sm.cstate + (j & BAND_CM), r_state + j);
            }
            if (s0 == 0) { asm volatile("cp.async.wait_all;" ::: "memory"); __syncthreads(); }
        }
        const int s_first = s0 + 1, n = min(BAND_CHUNK, nd - 1 - s0), s_last = s0 + n;
        // a chunk that may touch row 0 / 1 / lx-1 or column 0 / 1 / ly-1 takes the bodies with the terminal terms (the row and
        // column ranges of a diagonal only move forward: the chunk's first and last step bound them)
        const int hi_end = sm.geo[2 * ((s_last + 1) & BAND_GM) + 1];
        const bool bnd = lo_s0 <= 1 || hi_end >= J.lx - 2 || s0 - hi_s0 <= 1 || s_last + 1 - lo_s0 >= J.ly - 2;
        if (role == 0) {
            if (bnd) band_run_chunk<0, SMALLTAB, true>(sm, k, s_first, n, lane, PQ);
            else band_run_chunk<0, SMALLTAB, false>(sm, k, s_first, n, lane, PQ);
        } else if (role == 1) {
            if (bnd) band_run_chunk<1, SMALLTAB, true>(sm, k, s_first, n, lane, PQ);
            else band_run_chunk<1, SMALLTAB, false>(sm, k, s_first, n, lane, PQ);
        } else band_run_chunk<2, SMALLTAB, false>(sm, k, s_first, n, lane, PQ);
    }
    if (tid == 0) {
        // diagonal nd-1 sits in slot (nd-1) % 3, diagonal nd-2 (which holds the M of nd-1) in slot (nd-2) % 3; "diagonal -1" in slot 2
        const int b1 = ((nd - 1) % 3) * BAND_R, b2 = ((nd + 1) % 3) * BAND_R;
        const int r = (J.lx - 1) & BAND_RM;
        const int lo = g_geo[2 * (nd - 1)], hi = g_geo[2 * (nd - 1) + 1];
        double X = ninf, Y = ninf, M = ninf;
        if (J.lx - 1 >= lo && J.lx - 1 <= hi) { X = sm.x[b1 + r]; Y = sm.y[b1 + r]; M = sm.m[b2 + r]; }
        band_end_corner(k, X, Y, M, res);
    }
}
#endif

void launch_band_fill(bool smalltab, int n_jobs, const DevJob *jobs, const int *job_ids, const DevGraph *graphs, const DevModel *models,
                      const int *d_state, const int *d_band4, unsigned *ptrs, DevResult *results, cudaStream_t stream) {
    if (n_jobs <= 0) return;
#ifndef PG2_HOST_EMU
    const int smem = (int)band_smem_bytes();
    if (smalltab) {
        cudaFuncSetAttribute(band_fill_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        band_fill_kernel<true><<<n_jobs, BAND_THREADS, smem, stream>>>(jobs, job_ids, graphs, models, d_state, d_band4, ptrs, results);
    } else {
        cudaFuncSetAttribute(band_fill_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        band_fill_kernel<false><<<n_jobs, BAND_THREADS, smem, stream>>>(jobs, job_ids, graphs, models, d_state, d_band4, ptrs, results);
    }
#else
    // CPU test emulation: the same cell bodies, diagonals in order, the three matrices and the rows of a step one after the
    // other (they are independent: a step reads the slots of diagonals s-1, s-2 and writes the slot of s)
    (void)stream;
    for (int b = 0; b < n_jobs; ++b) {
        const int jid = job_ids[b];
        const DevJob J = jobs[jid];
        DevResult *res = results + jid;
        if (res->status != JOB_OK) continue;
        const DevModel m = models[J.model];
        std::vector<unsigned char> mem(band_smem_bytes() + 16);
        BandSm sm;
        band_carve(sm, mem.data());
        BandConst k;
        band_make_const(k, J, m);
        const int *l_state = d_state + graphs[J.left].state_base, *r_state = d_state + graphs[J.right].state_base;
        const int *g_geo = d_band4 + band_geo_off(J), *g_roff = d_band4 + band_roff_off(J);
        unsigned char *P8 = reinterpret_cast<unsigned char *>(ptrs + J.cell_base);
        const long long plane = band_plane_bytes(J);
        const int nd = J.lx + J.ly - 1;
        const double ninf = neg_inf();
        for (int e = 0; e < 3 * BAND_R; ++e) { sm.x[e] = ninf; sm.y[e] = ninf; sm.m[e] = ninf; }
        for (int e = 0; e < BAND_R; ++e) { sm.rstate[e] = 0; sm.rroff[e] = 0; }
        for (int e = 0; e < BAND_C; ++e) sm.cstate[e] = 0;
        const bool smalltab_job = m.fas <= STRIP_SMALL_FAS;
        if (smalltab_job)
            for (int e = 0; e < m.fas * m.fas; ++e) {
                const double ls = band_ls(m.table[e]);
                sm.stab[e] = make_double2(__dadd_rn(k.lng2, ls), __dadd_rn(k.lng, ls));
            }
        band_init_corner(sm, J, g_geo, g_roff, P8);
        int fr = 0, fc = 0, b0 = BAND_R, b1 = 0, b2 = 2 * BAND_R;
        for (int s = 1; s < nd; ++s) {
            BandStep t;
            t.s = s; t.lo0 = g_geo[2 * s]; t.hi0 = g_geo[2 * s + 1]; t.lo1 = g_geo[2 * (s + 1)]; t.hi1 = g_geo[2 * (s + 1) + 1];
            t.b0 = b0; t.b1 = b1; t.b2 = b2;
            // staged as far ahead as the kernel's chunks may reach (the ring sizes are part of what is tested)
            for (; fr < J.lx && fr <= t.hi0 + 2 * BAND_CHUNK + 8 + BAND_THREADS; ++fr) { if (fr > 0) sm.rstate[fr & BAND_RM] = l_state[fr]; sm.rroff[fr & BAND_RM] = g_roff[fr]; }
            for (; fc < J.ly && fc <= s + 2 * BAND_CHUNK + 8 + BAND_THREADS - t.lo0; ++fc) if (fc > 0) sm.cstate[fc & BAND_CM] = r_state[fc];
            const bool bnd = t.lo0 <= 1 || t.hi0 >= J.lx - 1 || s - t.hi0 <= 1 || s - t.lo0 >= J.ly - 1;
            const int i_end = t.hi0 + 3 + 31;  // the last pass of the kernel runs all its lanes
            for (int i = i_end; i >= t.lo0 - 1; --i) {
                if (bnd) { band_x_cell<true>(sm, k, t, i, P8); band_y_cell<true>(sm, k, t, i, P8 + plane); }
                else { band_x_cell<false>(sm, k, t, i, P8); band_y_cell<false>(sm, k, t, i, P8 + plane); }
                band_m_cell(sm, t, i, smalltab_job ? band_subst<true>(sm, k, s, i) : band_subst<false>(sm, k, s, i), P8 + 2 * plane);
            }
            const int freed = b2; b2 = b1; b1 = b0; b0 = freed;
        }
        const int r = (J.lx - 1) & BAND_RM;
        const int lo = g_geo[2 * (nd - 1)], hi = g_geo[2 * (nd - 1) + 1];
        double X = ninf, Y = ninf, M = ninf;
        if (J.lx - 1 >= lo && J.lx - 1 <= hi) { X = sm.x[b1 + r]; Y = sm.y[b1 + r]; M = sm.m[b2 + r]; }
        band_end_corner(k, X, Y, M, res);
    }
#endif
}

// ---- the walk: segments of BAND_SEG diagonals, every entry state of a segment walked by its own thread ----------------
// Segment k holds the diagonals [k SEG, (k+1) SEG).  A walk step goes down one diagonal (X, Y) or two (M), so the path
// enters segment k on its top diagonal (k+1) SEG - 1 or on the one below.  Candidate e of segment k: cell c = e / 3 of those
// two diagonals (top first), matrix e % 3.  Its record: where the walk leaves the segment (the candidate index of segment
// k-1, BAND_DONE, BAND_BROKEN), how many run-length encoded words and how many pointers it emits on the way.
constexpr int BAND_DONE = -1, BAND_BROKEN = -2, BAND_UNUSED = -3;

struct BandTrace {
    const unsigned char *P8;
    long long plane;
    const int *geo, *roff, *seg;
    int lx, ly, n_seg;
};

// run-length encoder of pg2_traceback.cu (StepEmit), counting only when out == nullptr
struct BandEmit {
    unsigned short *out;
    int n, raw, rep;
    unsigned last;
};
__device__ __forceinline__ void band_emit_flush(BandEmit &e) {
    if (e.rep > 0) { if (e.out) e.out[e.n] = (unsigned short)(0x8000u | (unsigned)e.rep); ++e.n; e.rep = 0; }
}
__device__ __forceinline__ void band_emit_step(BandEmit &e, unsigned q) {
    ++e.raw;
    if (q == e.last && e.rep < 0x7fff) { ++e.rep; return; }
    band_emit_flush(e);
    if (e.out) e.out[e.n] = (unsigned short)q;
    ++e.n;
    e.last = q;
}

__device__ __forceinline__ int band_diag_len(const BandTrace &T, int s) { return T.geo[2 * s + 1] - T.geo[2 * s] + 1; }

// candidate index of the state (i, j, vit) that entered segment k (i + j is its top diagonal or the one below)
__device__ __forceinline__ int band_entry_index(const BandTrace &T, int k, int i, int j, int vit) {
    const int top = (k + 1) * BAND_SEG - 1, s = i + j;
    const int c = (s == top ? 0 : band_diag_len(T, top)) + (i - T.geo[2 * s]);
    return T.seg[k] + 3 * c + vit;
}

// backtrack_new_path (:1073-1181) from (i, j, vit) down to the lower edge of segment k
__device__ __forceinline__ int band_walk(const BandTrace &T, int k, int i, int j, int vit, BandEmit &em) {
    const int floor_s = k * BAND_SEG;
    for (;;) {
        if (vit == NO_MAT || i < 0 || j < 0 || i >= T.lx || j >= T.ly) return BAND_BROKEN;
        const int s = i + j;
        if (i < T.geo[2 * s] || i > T.geo[2 * s + 1]) return BAND_BROKEN;  // outside the band
        if (s < floor_s) return band_entry_index(T, k - 1, i, j, vit);
        const unsigned q = (unsigned)T.P8[vit * T.plane + T.roff[i] + j] & 3u;  // = pack_ptr(source matrix, 0, 0)
        band_emit_step(em, q);
        // the reference loop ends when (i<1 && j<1) AFTER reading the cell it stands on
        if (vit == M_MAT || vit == X_MAT) i = (q == NO_MAT) ? -1 : i - 1;
        if (vit == M_MAT || vit == Y_MAT) j = (q == NO_MAT) ? -1 : j - 1;
        vit = (int)q;
        if (i < 1 && j < 1) return BAND_DONE;
    }
}

__device__ __forceinline__ void band_trace_setup(BandTrace &T, const DevJob &J, const int *d_band4, const unsigned *ptrs) {
    T.P8 = reinterpret_cast<const unsigned char *>(ptrs + J.cell_base);
    T.plane = band_plane_bytes(J);
    T.geo = d_band4 + band_geo_off(J);
    T.roff = d_band4 + band_roff_off(J);
    T.seg = d_band4 + band_seg_off(J);
    T.lx = J.lx; T.ly = J.ly; T.n_seg = J.n_seg;
}

// candidate e of segment k -> its state; false when e is past the segment's candidates
__device__ __forceinline__ bool band_candidate_state(const BandTrace &T, int k, int e, int &i, int &j, int &vit) {
    if (e >= T.seg[k + 1] - T.seg[k]) return false;
    const int top = (k + 1) * BAND_SEG - 1;
    int c = e / 3;
    vit = e - 3 * c;
    const int len0 = band_diag_len(T, top);
    const int s = c < len0 ? top : top - 1;
    if (c >= len0) c -= len0;
    i = T.geo[2 * s] + c;
    j = s - i;
    return true;
}

// the path's first state: the end pointer (viterbi_alignment.cpp:1047-1069; both stop sites have one edge, from the last site)
__device__ __forceinline__ void band_top_state(const DevJob &J, const DevResult *res, int &i, int &j, int &vit, unsigned &p) {
    p = res->end_ptr;
    vit = (int)(p & 3u);
    i = J.lx - 1;
    j = J.ly - 1;
}

// walks candidate e of segment k (the top segment has one: the end pointer) and records where it leaves the segment
__device__ __forceinline__ void band_trace_candidate(const DevJob &J, const BandTrace &T, const DevResult *res, int k, int e, int4 *cand) {
    BandEmit em;
    em.out = nullptr; em.n = 0; em.raw = 0; em.rep = 0; em.last = 0xffffffffu;
    int i, j, vit, slot;
    if (k == T.n_seg - 1) {
        if (e != 0) return;
        unsigned p;
        band_top_state(J, res, i, j, vit, p);
        slot = T.seg[T.n_seg - 1];  // the record behind the last segment's candidates
        if (vit == NO_MAT) { cand[slot] = make_int4(BAND_BROKEN, 0, 0, JOB_NO_PATH); return; }
        band_emit_step(em, p);
    } else {
        if (!band_candidate_state(T, k, e, i, j, vit)) return;
        slot = T.seg[k] + e;
    }
    const int exit_code = band_walk(T, k, i, j, vit, em);
    band_emit_flush(em);
    cand[slot] = make_int4(exit_code, em.n, em.raw, 0);
}

// links the segments the path visits: act[k] = {candidate the path enters segment k with, offset of the segment's words}
__device__ __forceinline__ void band_trace_stitch(const DevJob &J, const BandTrace &T, DevResult *res, const int4 *cand, int2 *act) {
    for (int k = 0; k < T.n_seg; ++k) act[k] = make_int2(BAND_UNUSED, 0);
    int4 c = cand[T.seg[T.n_seg - 1]];
    if (c.w == JOB_NO_PATH) { res->status = JOB_NO_PATH; return; }
    act[T.n_seg - 1] = make_int2(0, 0);
    int words = c.y, raw = c.z, k = T.n_seg - 2;
    while (c.x >= 0 && k >= 0) {
        act[k] = make_int2(c.x, words);
        c = cand[c.x];
        words += c.y;
        raw += c.z;
        --k;
    }
    res->n_steps = words;
    res->pad = raw;
    res->status = (c.x == BAND_DONE && raw <= J.step_cap) ? JOB_OK : JOB_BROKEN_PATH;
}

// segment k of the real path writes its words
__device__ __forceinline__ void band_trace_emit(const DevJob &J, const BandTrace &T, const DevResult *res, int k, const int2 *act,
                                                unsigned short *steps) {
    const int2 a = act[k];
    if (a.x == BAND_UNUSED) return;
    BandEmit em;
    em.out = steps + J.step_base + a.y; em.n = 0; em.raw = 0; em.rep = 0; em.last = 0xffffffffu;
    int i, j, vit;
    if (k == T.n_seg - 1) {
        unsigned p;
        band_top_state(J, res, i, j, vit, p);
        band_emit_step(em, p);
    } else if (!band_candidate_state(T, k, a.x - T.seg[k], i, j, vit)) return;
    band_walk(T, k, i, j, vit, em);
    band_emit_flush(em);
}

#ifndef PG2_HOST_EMU
__global__ void __launch_bounds__(128) band_trace_candidates_kernel(const int *job_ids, const DevJob *jobs, const int *d_band4,
                                                                    const unsigned *ptrs, const DevResult *results, int4 *cand_all) {
    const int jid = job_ids[blockIdx.y];
    const DevJob J = jobs[jid];
    const DevResult *res = results + jid;
    const int k = blockIdx.x;
    if (J.kernel != 4 || res->status != JOB_OK || k >= J.n_seg) return;
    BandTrace T;
    band_trace_setup(T, J, d_band4, ptrs);
    int4 *cand = cand_all + J.cand_base;
    const int n = k == J.n_seg - 1 ? 1 : T.seg[k + 1] - T.seg[k];
    for (int e = threadIdx.x; e < n; e += blockDim.x) band_trace_candidate(J, T, res, k, e, cand);
}
__global__ void band_trace_stitch_kernel(int n_jobs, const int *job_ids, const DevJob *jobs, const int *d_band4, const unsigned *ptrs,
                                         DevResult *results, const int4 *cand_all, int2 *act_all) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_jobs) return;
    const int jid = job_ids[t];
    const DevJob J = jobs[jid];
    DevResult *res = results + jid;
    if (J.kernel != 4 || res->status != JOB_OK) return;
    BandTrace T;
    band_trace_setup(T, J, d_band4, ptrs);
    band_trace_stitch(J, T, res, cand_all + J.cand_base, act_all + J.act_base);
}
__global__ void __launch_bounds__(128) band_trace_emit_kernel(const int *job_ids, const DevJob *jobs, const int *d_band4, const unsigned *ptrs,
                                                              const DevResult *results, const int2 *act_all, unsigned short *steps) {
    const int jid = job_ids[blockIdx.y];
    const DevJob J = jobs[jid];
    const DevResult *res = results + jid;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    // (a path the stitch found broken keeps the words of the segments it did reach, like the serial walks)
    if (J.kernel != 4 || (res->status != JOB_OK && res->status != JOB_BROKEN_PATH) || k >= J.n_seg) return;
    BandTrace T;
    band_trace_setup(T, J, d_band4, ptrs);
    band_trace_emit(J, T, res, k, act_all + J.act_base, steps);
}
#endif

// n_jobs: the jobs of the phase (all kernels; the band jobs among them are picked by J.kernel); max_seg: most segments of a band job
void launch_band_traceback(int n_jobs, int max_seg, const int *job_ids, const DevJob *jobs, const int *d_band4, const unsigned *ptrs,
                           DevResult *results, int4 *cand, int2 *act, unsigned short *steps, cudaStream_t stream) {
    if (n_jobs <= 0 || max_seg <= 0) return;
#ifndef PG2_HOST_EMU
    band_trace_candidates_kernel<<<dim3((unsigned)max_seg, (unsigned)n_jobs), 128, 0, stream>>>(job_ids, jobs, d_band4, ptrs, results, cand);
    band_trace_stitch_kernel<<<(n_jobs + 31) / 32, 32, 0, stream>>>(n_jobs, job_ids, jobs, d_band4, ptrs, results, cand, act);
    band_trace_emit_kernel<<<dim3((unsigned)(max_seg + 127) / 128, (unsigned)n_jobs), 128, 0, stream>>>(job_ids, jobs, d_band4, ptrs, results, act, steps);
#else
    (void)stream;
    for (int t = 0; t < n_jobs; ++t) {
        const int jid = job_ids[t];
        const DevJob J = jobs[jid];
        DevResult *res = results + jid;
        if (J.kernel != 4 || res->status != JOB_OK) continue;
        BandTrace T;
        band_trace_setup(T, J, d_band4, ptrs);
        for (int k = 0; k < J.n_seg; ++k) {
            const int n = k == J.n_seg - 1 ? 1 : T.seg[k + 1] - T.seg[k];
            for (int e = 0; e < n; ++e) band_trace_candidate(J, T, res, k, e, cand + J.cand_base);
        }
        band_trace_stitch(J, T, res, cand + J.cand_base, act + J.act_base);
        if (res->status != JOB_OK && res->status != JOB_BROKEN_PATH) continue;
        for (int k = 0; k < J.n_seg; ++k) band_trace_emit(J, T, res, k, act + J.act_base, steps);
    }
#endif
}

}  // namespace pg2
