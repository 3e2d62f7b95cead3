// pg2_wavefront.cu -- general fill kernel: one CTA per alignment, anti-diagonal wavefront.
//
// Handles every job shape (arbitrary in-degree and edge spans on both graphs, anchor bands).  The scores of
// the two previous anti-diagonals live in a shared-memory ring: the predecessors (i-1,j), (i,j-1), (i-1,j-1) --
// every predecessor of a plain site, 97 % of the sites of a PAGAN ancestor -- cost one shared-memory read, so a
// diagonal step is a barrier plus L1-resident CSR lookups instead of three chains of L2 round trips.  All
// scores also go to an anti-diagonal-major double4 scratch in HBM/L2 for long-span edges (skipped when both
// graphs are plain chains: nothing then reads further back than two diagonals).  One packed 32-bit word of
// back-pointers per cell is streamed out for the traceback kernel.
//
// Restates compute_fwd_scores / iterate_bwd_edges_for_gap / iterate_bwd_edges_for_match /
// iterate_bwd_edges_for_end_corner (reference src/main/viterbi_alignment.cpp:856-971, 1328-1552,
// 2029-2255).  Candidate order and FP64 association follow the reference exactly; ties keep the first
// candidate (strict '>', basic_alignment.h:449-462).
#include "pg2_device.cuh"
#ifdef PG2_HOST_EMU
#include <vector>
#endif

namespace pg2 {

constexpr int WAVE_RING_MAX = 2816;      // longest diagonal the shared-memory ring takes: 3 slots x 3 doubles per cell, 198 KB
constexpr int WAVE_RING_DOUBLES = 25344;  // the ring's budget: 198 KB of doubles
constexpr int WAVE_RING_DEPTH_MAX = 64;

// scores of the last `depth` diagonals in shared memory: buf[slot][X,Y,M][cap].  Depth 3 serves the predecessors of plain
// sites; a job with short diagonals gets as many slots as the budget holds (a 400-column read graph: 21), so that the
// sources of multi-edge sites -- cell (p, q) with s - (p + q) < depth, i.e. every edge pair whose spans add up to less than
// the depth -- come from shared memory as well instead of the L2-resident scratch (the pileup step on this kernel: 23.6 ->
// 19.5 ms; what remains is the divergent edge-pair loops of one SM's warps).
struct WaveRing {
    double *buf;         // nullptr: the diagonals of this job do not fit; every read goes to the global scratch
    int *range;          // [depth][2] first / last row held by every slot (last < first: the slot holds no diagonal)
    int cap;             // cells per diagonal the ring holds
    int depth, cur;      // slots; slot of the diagonal being computed (diagonal s - d sits d slots before it, cyclically)
    int s;               // diagonal being computed
    int off0, off1, off2;  // offsets (in doubles) of the ring slots of diagonals s, s-1, s-2
    int lo0, lo1, hi1, lo2, hi2;  // first row of diagonal s; row ranges of diagonals s-1 and s-2
    bool global_scores;  // false: both graphs are plain chains, the global scratch is never read
};

struct WaveCtx {
    WaveRing ring;
    // two-pass diagonals (ring only): sorted non-plain sites and plain-site bitmaps of both graphs, cursors at the
    // first listed site inside the current diagonal's row / column range
    const int *np_l, *np_r;
    const unsigned *mask_l, *mask_r;
    int n_np_l, n_np_r, cur_l, cur_r;
    bool two_pass;
    bool chain_job;  // both graphs are plain chains and a diagonal never has more cells than the CTA has threads
    const DevJob *job;
    const int *l_state, *l_off, *l_estart;
    const float *l_elogw;
    const int *r_state, *r_off, *r_estart;
    const float *r_elogw;
    const int *blo, *bhi, *dlo;
    const long long *doff;
    double4 *scores;
    int lx, ly;
    bool banded;
};

__device__ __forceinline__ long long cell_index(const WaveCtx &c, int p, int q) {
    int s = p + q;
    if (c.banded) return c.doff[s] + (p - c.dlo[s]);
    return diag_cum(s, c.lx, c.ly) + (p - diag_lo(s, c.ly));
}

// Tunnel_slice::at (utils/tunnel_matrix.h:85-98): -inf outside the band
__device__ __forceinline__ double4 load_cell(const WaveCtx &c, int p, int q) {
    if (c.ring.buf) {
        // rows lo..hi of a diagonal are exactly its in-band cells (the band is monotone)
        const int ds = c.ring.s - (p + q);
        if (ds == 1 || ds == 2) {
            const int lo = ds == 1 ? c.ring.lo1 : c.ring.lo2, hi = ds == 1 ? c.ring.hi1 : c.ring.hi2;
            if (p < lo || p > hi) {
                double ninf = neg_inf();
                return make_double4(ninf, ninf, ninf, 0.0);
            }
            const double *b = c.ring.buf + (ds == 1 ? c.ring.off1 : c.ring.off2) + (p - lo);
            return make_double4(b[0], b[c.ring.cap], b[2 * c.ring.cap], 0.0);
        }
        if (ds > 2 && ds < c.ring.depth) {  // an older diagonal the ring still holds
            int slot = c.ring.cur - ds;
            if (slot < 0) slot += c.ring.depth;
            const int lo = c.ring.range[2 * slot], hi = c.ring.range[2 * slot + 1];
            if (p < lo || p > hi) {
                double ninf = neg_inf();
                return make_double4(ninf, ninf, ninf, 0.0);
            }
            const double *b = c.ring.buf + slot * 3 * c.ring.cap + (p - lo);
            return make_double4(b[0], b[c.ring.cap], b[2 * c.ring.cap], 0.0);
        }
    }
    if (c.banded && (q < c.blo[p] || q > c.bhi[p])) {
        double ninf = neg_inf();
        return make_double4(ninf, ninf, ninf, 0.0);
    }
    const double4 *ptr = c.scores + cell_index(c, p, q);
    double2 a = __ldcg(reinterpret_cast<const double2 *>(ptr));
    double2 b = __ldcg(reinterpret_cast<const double2 *>(ptr) + 1);
    return make_double4(a.x, a.y, b.x, b.y);
}

// One gap cell (X when is_x, else Y).  Returns score, writes packed pointer (mat | ord<<2).
template <bool IS_X>
__device__ __forceinline__ double gap_cell(const WaveCtx &c, const DevModel &m, int i, int j, bool end_gap, bool reduced,
                                           unsigned &ptr_out) {
    const int *off = IS_X ? c.l_off : c.r_off;
    const int *es = IS_X ? c.l_estart : c.r_estart;
    int site = IS_X ? i : j;
    int k0 = off[site], k1 = off[site + 1];
    double best = neg_inf();
    unsigned ptr = NO_MAT;
    double ext = (double)(end_gap ? m.end_ext : m.ext);
    double open = (double)m.open, lng = (double)m.lng;
    for (int k = k0; k < k1; ++k) {
        int p = es[k];
        double4 v = IS_X ? load_cell(c, p, j) : load_cell(c, i, p);
        double same = IS_X ? v.x : v.y, other = IS_X ? v.y : v.x;
        unsigned ord = (unsigned)(k - k0) << 2;
        double s = __dadd_rn(same, ext);                       // score_gap_ext    :2116-2149
        if (s > best) { best = s; ptr = (IS_X ? X_MAT : Y_MAT) | ord; }
        s = __dadd_rn(__dadd_rn(other, 0.0), open);            // score_gap_double :2158-2180
        if (s > best) { best = s; ptr = (IS_X ? Y_MAT : X_MAT) | ord; }
        double pen = (reduced && p == 0) ? 0.0 : open;         // get_log_gap_open_penalty basic_alignment.h:490
        s = __dadd_rn(__dadd_rn(v.z, lng), pen);               // score_gap_open   :2190-2211
        if (s > best) { best = s; ptr = M_MAT | ord; }
    }
    ptr_out = ptr;
    return best;
}

__device__ __forceinline__ void match_pairs(const WaveCtx &c, int kl0, int kl1, int kr0, int kr1, double m_log, double x_log,
                                            double y_log, bool end_corner, double &best, unsigned &ptr) {
    for (int kl = kl0; kl < kl1; ++kl) {
        int pl = c.l_estart[kl];
        double wl = (double)c.l_elogw[kl];
        for (int kr = kr0; kr < kr1; ++kr) {
            int pr = c.r_estart[kr];
            double wr = (double)c.r_elogw[kr];
            double4 v = load_cell(c, pl, pr);
            unsigned ord = ((unsigned)(kl - kl0) << 2) | ((unsigned)(kr - kr0) << 8);
            double s = __dadd_rn(__dadd_rn(__dadd_rn(v.z, m_log), wl), wr);  // score_m_match :2029-2056
            if (s > best) { best = s; ptr = M_MAT | ord; }
            if (!end_corner) {
                s = __dadd_rn(__dadd_rn(__dadd_rn(v.x, x_log), wl), wr);     // score_x_match :2058-2084
                if (s > best) { best = s; ptr = X_MAT | ord; }
                s = __dadd_rn(__dadd_rn(__dadd_rn(v.y, y_log), wl), wr);     // score_y_match :2086-2112
                if (s > best) { best = s; ptr = Y_MAT | ord; }
            }
        }
    }
}

// iterate_bwd_edges_for_end_corner (:1440-1552).  Run by one thread after the last anti-diagonal.
__device__ void end_corner(const WaveCtx &c, const DevModel &m, DevResult *res) {
    int kl0 = c.l_off[c.lx], kl1 = c.l_off[c.lx + 1], kr0 = c.r_off[c.ly], kr1 = c.r_off[c.ly + 1];
    double best = neg_inf();
    unsigned ptr = NO_MAT;
    if (kl1 > kl0 && kr1 > kr0) {
        double m_log = (double)m.lng;
        // The reference interleaves M-pair and gap-close candidates; every candidate is compared with
        // strict '>' against the running maximum, so the visiting order below is what matters.
        auto m_pair = [&](int kl, int kr) {
            double4 v = load_cell(c, c.l_estart[kl], c.r_estart[kr]);
            double s = __dadd_rn(__dadd_rn(__dadd_rn(v.z, m_log), (double)c.l_elogw[kl]), (double)c.r_elogw[kr]);
            if (s > best) { best = s; ptr = pack_ptr(M_MAT, kl - kl0, kr - kr0); }
        };
        auto x_close = [&](int kl) {  // score_gap_close :2221-2255, close penalty 0
            double4 v = load_cell(c, c.l_estart[kl], c.ly - 1);
            double s = __dadd_rn(v.x, 0.0);
            if (s > best) { best = s; ptr = pack_ptr(X_MAT, kl - kl0, 0); }
        };
        auto y_close = [&](int kr) {
            double4 v = load_cell(c, c.lx - 1, c.r_estart[kr]);
            double s = __dadd_rn(v.y, 0.0);
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
    res->status = (best == neg_inf()) ? JOB_NO_PATH : JOB_OK;
}

// Per-job setup shared by the kernel and the CPU test emulation.
__device__ __forceinline__ void make_wave_ctx(WaveCtx &c, const DevJob &J, const DevGraph &GL, const DevGraph &GR, const int *d_state,
                                              const int *d_off, const int *d_estart, const float *d_elogw, const int *d_blo,
                                              const int *d_bhi, const int *d_dlo, const long long *d_doff, double4 *scores) {
    c.job = &J;
    c.l_state = d_state + GL.state_base;
    c.l_off = d_off + GL.off_base;
    c.l_estart = d_estart + GL.edge_base;
    c.l_elogw = d_elogw + GL.edge_base;
    c.r_state = d_state + GR.state_base;
    c.r_off = d_off + GR.off_base;
    c.r_estart = d_estart + GR.edge_base;
    c.r_elogw = d_elogw + GR.edge_base;
    c.lx = J.lx;
    c.ly = J.ly;
    c.banded = J.banded != 0;
    c.blo = c.banded ? d_blo + J.band_base : nullptr;
    c.bhi = c.banded ? d_bhi + J.band_base : nullptr;
    c.dlo = c.banded ? d_dlo + J.diag_base : nullptr;
    c.doff = c.banded ? d_doff + J.diag_base : nullptr;
    c.scores = scores + J.cell_base;
    c.ring.buf = nullptr;
    c.ring.range = nullptr;
    c.ring.cap = 0;
    c.ring.depth = 3;
    c.ring.cur = 1;
    c.ring.global_scores = true;
    c.two_pass = false;
    c.chain_job = false;
}

// the ring holds this job's diagonals: serve near reads from it; plain chains on both sides never read the scratch
// slots the ring of a launch group gets: as many as the budget holds, at least the three the plain predecessors need
__host__ __device__ inline int wave_ring_depth(int cap) {
    const int d = WAVE_RING_DOUBLES / (3 * (cap > 0 ? cap : 1));
    return d < 3 ? 3 : (d > WAVE_RING_DEPTH_MAX ? WAVE_RING_DEPTH_MAX : d);
}
__device__ __forceinline__ void wave_use_ring(WaveCtx &c, double *buf, int *range, int cap, int depth, const DevGraph &GL, const DevGraph &GR,
                                              const int *d_vlast) {
    // two passes pay when the listed sites are a handful (plain leaves inside an anchor band): the general body's
    // latency is then off the diagonal's critical path; with a few per cent of listed sites the single pass is faster
    if (GL.np_base >= 0 && GR.np_base >= 0 && GL.n_np + GR.n_np <= 2 + (GL.n_sites + GR.n_sites) / 400) {
        c.two_pass = true;
        c.np_l = d_vlast + GL.np_base; c.n_np_l = GL.n_np;
        c.np_r = d_vlast + GR.np_base; c.n_np_r = GR.n_np;
        c.mask_l = reinterpret_cast<const unsigned *>(d_vlast + GL.npmask_base);
        c.mask_r = reinterpret_cast<const unsigned *>(d_vlast + GR.npmask_base);
        c.cur_l = c.cur_r = 0;
    }
    c.ring.buf = buf;
    c.ring.range = range;
    c.ring.cap = cap;
    c.ring.depth = depth;
    c.ring.global_scores = !(GL.simple && GR.simple);
    c.ring.lo1 = 0; c.ring.hi1 = 0;   // diagonal 0 is the start corner
    c.ring.lo2 = 0; c.ring.hi2 = -1;  // there is no diagonal -1
    // the first diagonal computed is s = 1: it goes to slot 1, diagonal 0 sits in slot 0, "diagonal -1" in the last slot
    c.ring.cur = 1;
    c.ring.off0 = 3 * cap; c.ring.off1 = 0; c.ring.off2 = (depth - 1) * 3 * cap;
}
// `writer`: the one thread that records the diagonal's row range for the reads of later diagonals (they see it after this
// diagonal's barrier; the slot's previous tenant, diagonal s - depth, is out of every reader's reach)
__device__ __forceinline__ void wave_ring_begin(WaveCtx &c, int s, int ilo, int ihi, bool writer) {
    c.ring.s = s;
    c.ring.lo0 = ilo;
    if (writer && c.ring.range) { c.ring.range[2 * c.ring.cur] = ilo; c.ring.range[2 * c.ring.cur + 1] = ihi; }
}
__device__ __forceinline__ void wave_ring_end(WaveCtx &c, int ilo, int ihi) {
    c.ring.lo2 = c.ring.lo1; c.ring.hi2 = c.ring.hi1;
    c.ring.lo1 = ilo; c.ring.hi1 = ihi;
    c.ring.cur = c.ring.cur + 1 == c.ring.depth ? 0 : c.ring.cur + 1;
    c.ring.off2 = c.ring.off1;
    c.ring.off1 = c.ring.off0;
    c.ring.off0 = c.ring.cur * 3 * c.ring.cap;
}

// initialise_array_corner (:725-733)
__device__ __forceinline__ void wave_init(const WaveCtx &c, unsigned *P) {
    double ninf = neg_inf();
    double2 *s0 = reinterpret_cast<double2 *>(c.scores);
    s0[0] = make_double2(ninf, ninf);
    s0[1] = make_double2(0.0, 0.0);
    if (c.ring.buf) {  // slot 0 = diagonal 0; the other slots hold nothing yet
        c.ring.buf[0] = ninf; c.ring.buf[c.ring.cap] = ninf; c.ring.buf[2 * c.ring.cap] = 0.0;
        for (int d = 0; d < c.ring.depth; ++d) { c.ring.range[2 * d] = 0; c.ring.range[2 * d + 1] = d == 0 ? 0 : -1; }
    }
    P[0] = cell_word(NO_MAT, NO_MAT, NO_MAT);
}

__device__ __forceinline__ void diag_geometry(const WaveCtx &c, int s, int &ilo, int &ihi, long long &base) {
    if (c.banded) {
        ilo = c.dlo[s];
        base = c.doff[s];
        ihi = ilo + (int)(c.doff[s + 1] - base) - 1;
    } else {
        ilo = diag_lo(s, c.ly);
        ihi = diag_hi(s, c.lx);
        base = diag_cum(s, c.lx, c.ly);
    }
}

// compute_fwd_scores (:856-971) for one cell
__device__ __forceinline__ void wave_cell(const WaveCtx &c, const DevModel &m, unsigned flags, float lng2, int i, int j, long long idx,
                                          unsigned *P) {
    const bool term = !(flags & FLAG_NO_TERMINAL_EDGES);
    const bool reduced = (flags & FLAG_REDUCED) != 0;
    const double ninf = neg_inf();
    double sx = ninf, sy = ninf, sm = ninf;
    unsigned px = NO_MAT, py = NO_MAT, pm = NO_MAT;
    if (i > 0) sx = gap_cell<true>(c, m, i, j, term && (j == 0 || j == c.ly - 1), reduced, px);
    if (j > 0) sy = gap_cell<false>(c, m, i, j, term && (i == 0 || i == c.lx - 1), reduced, py);
    if (i > 0 && j > 0) {
        double ls = (double)__ldg(m.table + (size_t)c.l_state[i] + (size_t)c.r_state[j] * (size_t)m.fas);
        double m_log = __dadd_rn((double)lng2, ls);
        double x_log = __dadd_rn((double)m.lng, ls);
        match_pairs(c, c.l_off[i], c.l_off[i + 1], c.r_off[j], c.r_off[j + 1], m_log, x_log, x_log, false, sm, pm);
    }
    if (c.ring.buf) {
        double *b = c.ring.buf + c.ring.off0 + (i - c.ring.lo0);
        b[0] = sx; b[c.ring.cap] = sy; b[2 * c.ring.cap] = sm;
    }
    if (c.ring.global_scores) {
        double2 *dst = reinterpret_cast<double2 *>(c.scores + idx);
        dst[0] = make_double2(sx, sy);
        dst[1] = make_double2(sm, 0.0);
    }
    // gap_cell returns (mat | ord<<2) for both X and Y, which is the in-word form.  Bits 30 / 31: the only backward
    // edge of left site i / right site j comes from the site before it -- the walk then needs no CSR lookup.
    unsigned plain = 0;
    if (i > 0) { const int k0 = c.l_off[i]; if (c.l_off[i + 1] - k0 == 1 && c.l_estart[k0] == i - 1) plain |= WORD_PLAIN_LEFT; }
    if (j > 0) { const int k0 = c.r_off[j]; if (c.r_off[j + 1] - k0 == 1 && c.r_estart[k0] == j - 1) plain |= WORD_PLAIN_RIGHT; }
    P[idx] = cell_word(px, py, pm) | plain;
}

// a cell of the two previous diagonals (ds = 1 or 2) from the ring; rows outside the diagonal's range are outside the band
__device__ __forceinline__ void ring_near(const WaveCtx &c, int ds, int p, double &x, double &y, double &m) {
    const int lo = ds == 1 ? c.ring.lo1 : c.ring.lo2, hi = ds == 1 ? c.ring.hi1 : c.ring.hi2;
    if (p < lo || p > hi) { x = y = m = neg_inf(); return; }
    const double *b = c.ring.buf + (ds == 1 ? c.ring.off1 : c.ring.off2) + (p - lo);
    x = b[0]; y = b[c.ring.cap]; m = b[2 * c.ring.cap];
}

__device__ __forceinline__ bool site_plain(const unsigned *mask, int s) { return (mask[s >> 5] >> (s & 31)) & 1u; }

// compute_fwd_scores for a cell whose left site i and right site j are both plain (one backward edge each, from i-1 and
// j-1; i, j >= 1): the same candidates in the same order as gap_cell / match_pairs with single-trip loops, all three
// source cells from the ring.
__device__ __forceinline__ void wave_cell_plain_core(const WaveCtx &c, const DevModel &m, unsigned flags, float lng2, int i, int j, long long idx,
                                                     unsigned *P, int sl, int sr, double wl, double wr) {
    const bool term = !(flags & FLAG_NO_TERMINAL_EDGES);
    const bool reduced = (flags & FLAG_REDUCED) != 0;
    const double ninf = neg_inf();
    const double open = (double)m.open, lng = (double)m.lng;
    double ux, uy, um, lx_, ly_, lm, dx, dy, dm;
    ring_near(c, 1, i - 1, ux, uy, um);  // (i-1, j)
    ring_near(c, 1, i, lx_, ly_, lm);    // (i, j-1)
    ring_near(c, 2, i - 1, dx, dy, dm);  // (i-1, j-1)
    const double ls = (double)__ldg(m.table + (size_t)sl + (size_t)sr * (size_t)m.fas);
    // X: ext, double, open from (i-1, j)   (gap_cell<true>)
    double sx = ninf, s;
    unsigned px = NO_MAT;
    s = __dadd_rn(ux, (double)((term && j == c.ly - 1) ? m.end_ext : m.ext));
    if (s > sx) { sx = s; px = X_MAT; }
    s = __dadd_rn(__dadd_rn(uy, 0.0), open);
    if (s > sx) { sx = s; px = Y_MAT; }
    s = __dadd_rn(__dadd_rn(um, lng), (reduced && i == 1) ? 0.0 : open);
    if (s > sx) { sx = s; px = M_MAT; }
    // Y: ext, double, open from (i, j-1)   (gap_cell<false>)
    double sy = ninf;
    unsigned py = NO_MAT;
    s = __dadd_rn(ly_, (double)((term && i == c.lx - 1) ? m.end_ext : m.ext));
    if (s > sy) { sy = s; py = Y_MAT; }
    s = __dadd_rn(__dadd_rn(lx_, 0.0), open);
    if (s > sy) { sy = s; py = X_MAT; }
    s = __dadd_rn(__dadd_rn(lm, lng), (reduced && j == 1) ? 0.0 : open);
    if (s > sy) { sy = s; py = M_MAT; }
    // M: from M, X, Y of (i-1, j-1)   (match_pairs)
    const double m_log = __dadd_rn((double)lng2, ls), x_log = __dadd_rn(lng, ls);
    double sm = ninf;
    unsigned pm = NO_MAT;
    s = __dadd_rn(__dadd_rn(__dadd_rn(dm, m_log), wl), wr);
    if (s > sm) { sm = s; pm = M_MAT; }
    s = __dadd_rn(__dadd_rn(__dadd_rn(dx, x_log), wl), wr);
    if (s > sm) { sm = s; pm = X_MAT; }
    s = __dadd_rn(__dadd_rn(__dadd_rn(dy, x_log), wl), wr);
    if (s > sm) { sm = s; pm = Y_MAT; }
    double *b = c.ring.buf + c.ring.off0 + (i - c.ring.lo0);
    b[0] = sx; b[c.ring.cap] = sy; b[2 * c.ring.cap] = sm;
    if (c.ring.global_scores) {
        double2 *dst = reinterpret_cast<double2 *>(c.scores + idx);
        dst[0] = make_double2(sx, sy);
        dst[1] = make_double2(sm, 0.0);
    }
    P[idx] = cell_word(px, py, pm) | WORD_PLAIN_LEFT | WORD_PLAIN_RIGHT;
}
__device__ __forceinline__ void wave_cell_plain(const WaveCtx &c, const DevModel &m, unsigned flags, float lng2, int i, int j, long long idx,
                                                unsigned *P) {
    wave_cell_plain_core(c, m, flags, lng2, i, j, idx, P, c.l_state[i], c.r_state[j], (double)c.l_elogw[c.l_off[i]],
                         (double)c.r_elogw[c.r_off[j]]);
}

// Chain x chain jobs (anchored leaf x leaf, 400 000 short diagonals): a thread keeps the cell with the same offset on
// every diagonal, so from one diagonal to the next exactly one of its row / column moves on by one site.  The state and
// edge weight of the current and of the NEXT row and column stay in registers; the load for the site after that is issued
// a diagonal ahead, so no global load sits on the diagonal's critical path (nor does the band geometry, see the kernel).
struct ChainTrack {
    int i, j;  // the row / column the fields below describe (-2: nothing loaded)
    int sl, sl_n, sr, sr_n;
    double wl, wl_n, wr, wr_n;
};
__device__ __forceinline__ void chain_site(const int *state, const float *elogw, int n, int s, int &st, double &w) {
    const int q = s < 1 ? 1 : (s > n - 1 ? n - 1 : s);  // a plain chain: site q >= 1 is entered by edge q-1
    st = state[q];
    w = (double)elogw[q - 1];
}
__device__ __forceinline__ void wave_chain_cell(const WaveCtx &c, const DevModel &m, unsigned flags, float lng2, int s, int ilo, int ihi,
                                                long long base, unsigned *P, int tid, ChainTrack &t) {
    const int i = ilo + tid, j = s - i;
    if (i == t.i + 1) { t.sl = t.sl_n; t.wl = t.wl_n; chain_site(c.l_state, c.l_elogw, c.lx, i + 1, t.sl_n, t.wl_n); }
    else if (i != t.i) { chain_site(c.l_state, c.l_elogw, c.lx, i, t.sl, t.wl); chain_site(c.l_state, c.l_elogw, c.lx, i + 1, t.sl_n, t.wl_n); }
    if (j == t.j + 1) { t.sr = t.sr_n; t.wr = t.wr_n; chain_site(c.r_state, c.r_elogw, c.ly, j + 1, t.sr_n, t.wr_n); }
    else if (j != t.j) { chain_site(c.r_state, c.r_elogw, c.ly, j, t.sr, t.wr); chain_site(c.r_state, c.r_elogw, c.ly, j + 1, t.sr_n, t.wr_n); }
    t.i = i; t.j = j;
    if (i > ihi) return;
    if (i >= 1 && j >= 1) wave_cell_plain_core(c, m, flags, lng2, i, j, base + tid, P, t.sl, t.sr, t.wl, t.wr);
    else wave_cell(c, m, flags, lng2, i, j, base + tid, P);  // first row / column: the general body
}

// One anti-diagonal, the share of thread `tid` of `nthreads`.  With the ring: first every cell whose two sites are
// plain (one lean body, no CSR walk, no divergence), then the cells of the listed non-plain rows and columns that
// cross this diagonal through the general body -- a few per cent of the cells, gathered into the first threads
// instead of dragging every warp through the edge loops.  Cells of one diagonal are independent.
__device__ __forceinline__ void wave_diagonal(WaveCtx &c, const DevModel &m, unsigned flags, float lng2, int s, int ilo, int ihi, long long base,
                                              unsigned *P, int tid, int nthreads) {
    if (!c.two_pass) {
        for (int i = ilo + tid; i <= ihi; i += nthreads) wave_cell(c, m, flags, lng2, i, s - i, base + (i - ilo), P);
        return;
    }
    for (int i = ilo + tid; i <= ihi; i += nthreads)
        if (site_plain(c.mask_l, i) && site_plain(c.mask_r, s - i)) wave_cell_plain(c, m, flags, lng2, i, s - i, base + (i - ilo), P);
    // listed rows inside [ilo, ihi] and listed columns inside [s - ihi, s - ilo]; both ranges only move forward
    const int jlo = s - ihi, jhi = s - ilo;
    while (c.cur_l < c.n_np_l && c.np_l[c.cur_l] < ilo) ++c.cur_l;
    while (c.cur_r < c.n_np_r && c.np_r[c.cur_r] < jlo) ++c.cur_r;
    for (int e = c.cur_l + tid; e < c.n_np_l; e += nthreads) {
        const int i = c.np_l[e];
        if (i > ihi) break;
        wave_cell(c, m, flags, lng2, i, s - i, base + (i - ilo), P);
    }
    for (int e = c.cur_r + tid; e < c.n_np_r; e += nthreads) {
        const int j = c.np_r[e];
        if (j > jhi) break;
        const int i = s - j;
        if (site_plain(c.mask_l, i)) wave_cell(c, m, flags, lng2, i, j, base + (i - ilo), P);  // else: done with its row above
    }
}

#ifndef PG2_HOST_EMU
__global__ void __launch_bounds__(1024, 1)
wavefront_fill_kernel(const DevJob *jobs, const int *job_ids, const DevGraph *graphs, const DevModel *models, const int *d_state,
                      const int *d_off, const int *d_estart, const float *d_elogw, const int *d_blo, const int *d_bhi,
                      const int *d_dlo, const long long *d_doff, const int *d_vlast, double4 *scores, unsigned *ptrs, DevResult *results,
                      int ring_cap, int ring_depth) {
    extern __shared__ __align__(16) double wave_smem[];
    const int jid = job_ids[blockIdx.x];
    const DevJob &J = jobs[jid];
    DevResult *res = results + jid;
    if (res->status != JOB_OK) return;  // rejected by the validation kernel
    const DevGraph GL = graphs[J.left], GR = graphs[J.right];
    const DevModel m = models[J.model];
    WaveCtx c;
    make_wave_ctx(c, J, GL, GR, d_state, d_off, d_estart, d_elogw, d_blo, d_bhi, d_dlo, d_doff, scores);
    if (ring_cap > 0) wave_use_ring(c, wave_smem, reinterpret_cast<int *>(wave_smem + (size_t)ring_depth * 3 * ring_cap), ring_cap, ring_depth, GL, GR, d_vlast);
    unsigned *P = ptrs + J.cell_base;
    const unsigned flags = J.flags;
    const float lng2 = __fmul_rn(2.0f, m.lng);  // 2*model->log_non_gap() stays float (:1364)

    if (threadIdx.x == 0) wave_init(c, P);
    __syncthreads();

    const int n_diag = c.lx + c.ly - 1;
    if (ring_cap > 0 && GL.simple && GR.simple && ring_cap <= (int)blockDim.x) {
        // chain x chain: per-thread site tracking; the band geometry of 32 diagonals is fetched by the 32 lanes of a
        // warp in one round trip and handed out by shuffles, so no global load sits between two diagonals
        c.chain_job = true;
        ChainTrack track;
        track.i = track.j = -2;
        const int lane = (int)threadIdx.x & 31;
        for (int s0 = 1; s0 < n_diag; s0 += 32) {
            int g_lo, g_hi;
            long long g_base;
            diag_geometry(c, min(s0 + lane, n_diag - 1), g_lo, g_hi, g_base);
            const int nd = min(32, n_diag - s0);
            for (int d = 0; d < nd; ++d) {
                const int s = s0 + d;
                const int ilo = __shfl_sync(0xffffffffu, g_lo, d), ihi = __shfl_sync(0xffffffffu, g_hi, d);
                const long long base = __shfl_sync(0xffffffffu, g_base, d);
                wave_ring_begin(c, s, ilo, ihi, threadIdx.x == 0);
                wave_chain_cell(c, m, flags, lng2, s, ilo, ihi, base, P, (int)threadIdx.x, track);
                __syncthreads();
                wave_ring_end(c, ilo, ihi);
            }
        }
    } else
    for (int s = 1; s < n_diag; ++s) {
        int ilo, ihi;
        long long base;
        diag_geometry(c, s, ilo, ihi, base);
        wave_ring_begin(c, s, ilo, ihi, threadIdx.x == 0);
        wave_diagonal(c, m, flags, lng2, s, ilo, ihi, base, P, (int)threadIdx.x, (int)blockDim.x);
        __syncthreads();  // diagonal s is complete; the ring slot of diagonal s-2 may be overwritten
        wave_ring_end(c, ilo, ihi);
    }
    wave_ring_begin(c, n_diag, 0, -1, false);  // the end corner reads the cells of the last diagonals from the ring
    if (threadIdx.x == 0) end_corner(c, m, res);
}
#endif

void launch_wavefront_fill(int n_jobs, int threads, const DevJob *jobs, const int *job_ids, const DevGraph *graphs,
                           const DevModel *models, const int *d_state, const int *d_off, const int *d_estart, const float *d_elogw,
                           const int *d_blo, const int *d_bhi, const int *d_dlo, const long long *d_doff, const int *d_vlast,
                           double4 *scores, unsigned *ptrs, DevResult *results, int max_diag, cudaStream_t stream) {
    if (n_jobs <= 0) return;
    // ring of three diagonals x {X,Y,M} in shared memory when the group's longest diagonal fits
    const int ring_cap = max_diag <= WAVE_RING_MAX ? (max_diag > 0 ? max_diag : 1) : 0;
    const int ring_depth = wave_ring_depth(ring_cap);
#ifndef PG2_HOST_EMU
    const int smem = ring_cap > 0 ? ring_cap * 3 * ring_depth * (int)sizeof(double) + 2 * ring_depth * (int)sizeof(int) : 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(wavefront_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    wavefront_fill_kernel<<<n_jobs, threads, smem, stream>>>(jobs, job_ids, graphs, models, d_state, d_off, d_estart, d_elogw, d_blo,
                                                             d_bhi, d_dlo, d_doff, d_vlast, scores, ptrs, results, ring_cap, ring_depth);
#else
    // CPU test emulation: anti-diagonals in order, cells of one diagonal in any order
    (void)threads; (void)stream;
    for (int b = 0; b < n_jobs; ++b) {
        const int jid = job_ids[b];
        const DevJob &J = jobs[jid];
        DevResult *res = results + jid;
        if (res->status != JOB_OK) continue;
        const DevGraph GL = graphs[J.left], GR = graphs[J.right];
        const DevModel m = models[J.model];
        WaveCtx c;
        make_wave_ctx(c, J, GL, GR, d_state, d_off, d_estart, d_elogw, d_blo, d_bhi, d_dlo, d_doff, scores);
        std::vector<double> ring((size_t)ring_cap * 3 * ring_depth + 1);
        std::vector<int> ring_range((size_t)2 * ring_depth);
        if (ring_cap > 0) wave_use_ring(c, ring.data(), ring_range.data(), ring_cap, ring_depth, GL, GR, d_vlast);
        unsigned *P = ptrs + J.cell_base;
        const float lng2 = __fmul_rn(2.0f, m.lng);
        wave_init(c, P);
        const int n_diag = c.lx + c.ly - 1;
        // the kernel's thread count for this group: a power of two >= max_diag (pg2_engine.cu)
        int emu_threads = 32;
        while (emu_threads < max_diag && emu_threads < 1024) emu_threads *= 2;
        if (ring_cap > 0 && GL.simple && GR.simple && ring_cap <= emu_threads) {
            c.chain_job = true;
            std::vector<ChainTrack> track((size_t)emu_threads);
            for (auto &t : track) t.i = t.j = -2;
            for (int s = 1; s < n_diag; ++s) {
                int ilo, ihi;
                long long base;
                diag_geometry(c, s, ilo, ihi, base);
                wave_ring_begin(c, s, ilo, ihi, true);
                for (int tid = emu_threads - 1; tid >= 0; --tid) wave_chain_cell(c, m, J.flags, lng2, s, ilo, ihi, base, P, tid, track[(size_t)tid]);
                wave_ring_end(c, ilo, ihi);
            }
        } else
        for (int s = 1; s < n_diag; ++s) {
            int ilo, ihi;
            long long base;
            diag_geometry(c, s, ilo, ihi, base);
            wave_ring_begin(c, s, ilo, ihi, true);
            // a handful of emulated threads, last first: the shares of a diagonal are independent of each other
            for (int tid = 4; tid >= 0; --tid) wave_diagonal(c, m, J.flags, lng2, s, ilo, ihi, base, P, tid, 5);
            wave_ring_end(c, ilo, ihi);
        }
        wave_ring_begin(c, n_diag, 0, -1, false);
        end_corner(c, m, res);
    }
#endif
}

}  // namespace pg2
