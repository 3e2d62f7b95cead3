// pg2_strip_geom.cuh -- pointer-buffer geometry of the strip kernel, shared by the fill kernel, the
// traceback kernel and the engine's scratch accounting.
#pragma once
#include "pg2_device.cuh"

namespace pg2 {

constexpr int STRIP_MAX_LEFT_INDEG = 16;

// d_vrow[]: the row program of a graph used as the ROW graph of the strip kernel.  A DP row with d
// backward edges becomes d consecutive virtual rows, one edge each, so that a warp whose lanes are on
// different rows never waits for the lane with the highest in-degree: every lane does exactly one edge per
// step.  Entry = int4 {info, edge, site, src}:
//   info bits 0-11  character state of the site (0 for the start site)
//        bit  12    FIRST virtual row of the site (accumulators start at -inf)
//        bit  13    LAST virtual row of the site (Y chain, pointers, commit)
//        bit  14    REG: the edge starts at the site just above (source row is in registers)
//        bit  15    ZERO_W: the edge's log weight is +0.0
//        bit  16    ENDPRED: the end corner reads this row's last column (set on the LAST virtual row)
//        bit  17    NOEDGE: the site has no backward edge (the start site; unreachable sites)
//        bits 18-31 saved-row slot + 1 when the site is the source of a long-span edge (on the LAST row)
//   edge  CSR position of the edge (graph-relative), -1 when NOEDGE
//   site  the DP row
//   src   bits 0-15 saved-row slot of the edge's start site (when not REG), bits 16-23 edge ordinal in the site
constexpr int VR_STATE_MASK = 0xfff;
constexpr int VR_FIRST = 1 << 12;
constexpr int VR_LAST = 1 << 13;
constexpr int VR_REG = 1 << 14;
constexpr int VR_ZERO_W = 1 << 15;
constexpr int VR_ENDPRED = 1 << 16;
constexpr int VR_NOEDGE = 1 << 17;
constexpr int VR_SLOT_SHIFT = 18;
constexpr int VR_FAST = VR_FIRST | VR_LAST | VR_REG;  // all three: the in-place fast row applies
constexpr int STRIP_MAX_SLOTS = 16000;
constexpr int STRIP_SMALL_FAS = 16;       // alphabets up to this size use the shared double2 table (DNA: 15)

// uint16 cell word: X ptr bits 0-5 (mat | lord<<2), Y ptr bits 6-7 (mat), M ptr bits 8-13 (mat | lord<<2)
__host__ __device__ inline unsigned strip_word(unsigned px, unsigned py, unsigned pm) {
    return (px & 0x3fu) | ((py & 3u) << 6) | ((pm & 0x3fu) << 8);
}

__host__ __device__ inline int strip_ks(int K) { return (K + 1) & ~1; }  // half-words per lane slot (even)

// Geometry shared with the engine (scratch accounting) and the traceback (cell lookup).
__host__ __device__ inline int strip_pick_k(int ly) {
    // smallest padded width among the compiled strip widths
    const int ks[6] = {2, 3, 4, 5, 6, 8};
    int best = 8;
    long long best_cost = -1;
    for (int a = 0; a < 6; ++a) {
        int W = 32 * ks[a];
        long long blocks = (ly + W - 1) / W;
        long long cost = blocks * W + blocks * 24;  // columns computed (+ a per-block sweep overhead)
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ks[a]; }
    }
    return best;
}
// nv = virtual rows of the row graph; v = the virtual row that completes DP row i
__host__ __device__ inline long long strip_cells(int nv, int ly, int K) {
    int W = 32 * K;
    long long blocks = (ly + W - 1) / W;
    return blocks * (long long)(nv + 31) * 32 * strip_ks(K);
}
__host__ __device__ inline long long strip_ptr_index(int nv, int ly, int K, int v, int j) {
    (void)ly;
    int W = 32 * K;
    int b = j / W, jj = j - b * W;
    int l = jj / K, k = jj - l * K;
    return (((long long)b * (nv + 31) + (v + l)) * 32 + l) * strip_ks(K) + k;
}

// ---- lane kernel (pg2_lanes.cu): 32 alignments that share the row graph, one per lane -----------------
// Pointer buffer of one task: [strip][virtual row][K/8][lane] uint4, i.e. 8 half-words (8 columns) per lane
// per 128-bit store, lanes interleaved so that a warp-wide store is one contiguous 512 B segment.
constexpr int LANE_K = 8;             // columns per lane strip (multiple of 8)
constexpr int LANE_W = 4;             // warps per CTA = strips of one task in flight (three CTAs per SM: the throughput shape)
constexpr int LANE_W_WIDE = 10;       // ... of a launch that cannot fill the chip (one CTA per SM: a task ends 2.5 times sooner)
constexpr int LANE_WIDE_TASKS_PER_SM = 10;  // launches with at most this many tasks per SM take the wide shape (measured on the
                                      // placement workload: 848 tasks 29.2 -> 19.3 ms, 1 622 tasks 33.8 vs 34.7 ms, 3 175 tasks 56 vs 65 ms)
#ifndef PG2_LANE_B
#define PG2_LANE_B 8
#endif
#ifndef PG2_LANE_D
#define PG2_LANE_D 2
#endif
constexpr int LANE_B = PG2_LANE_B;    // virtual rows per pipeline block
constexpr int LANE_D = PG2_LANE_D;    // blocks a warp may run ahead of the warp that consumes its boundary column
constexpr int LANE_MIN_JOBS = 16;     // fewer jobs on one row graph than this stay on the strip kernel
constexpr int LANE_SLOT_DOUBLES = (LANE_K + 1) * 4 * 32;  // one parked row of one warp: [K+1 columns][X,Y,M,Mo][lane]
__host__ __device__ inline long long lane_cells(int nv, int max_ly, int K) {  // half-words per task
    long long strips = (max_ly + K - 1) / K;
    return strips * (long long)nv * 32 * K;
}
__host__ __device__ inline long long lane_ptr_index(int nv, int K, int v, int j, int lane) {
    int s = j / K, k = j - s * K;
    return ((((long long)s * nv + v) * (K / 8) + (k >> 3)) * 32 + lane) * 8 + (k & 7);
}
// per-CTA global scratch of the lane kernel, in doubles: the wrap column [virtual row][X,Y,M][lane] (strip
// boundary handed from the CTA's last warp to its first), the end column [row][X,Y,M][lane] (what the end
// corner reads) and per warp n_slots + 2 parked rows (rows that start long-span edges, the row above an open
// general site, and the pointer accumulators of a site that straddles two pipeline blocks).
__host__ __device__ inline long long lane_cta_doubles(int max_nv, int max_lx, int n_slots, int W) {
    return (long long)max_nv * 96 + (long long)max_lx * 96 + (long long)W * (n_slots + 2) * LANE_SLOT_DOUBLES;
}

// Decodes one pointer of a strip-kernel half-word into the API encoding (mat | lord<<2 | rord<<8).
__host__ __device__ inline unsigned strip_decode_ptr(unsigned w, int mat) {
    if (w & 0x4000u) {  // fast row: raw comparison bits, single left edge (ordinal 0)
        if (mat == X_MAT) { unsigned p1 = w & 1u, p2 = (w >> 1) & 1u; return p2 ? M_MAT : (p1 ? Y_MAT : X_MAT); }
        if (mat == Y_MAT) { unsigned p2 = (w >> 2) & 1u, p1 = (w >> 3) & 1u; return p1 ? (p2 ? M_MAT : X_MAT) : Y_MAT; }
        unsigned p1 = (w >> 4) & 1u, p2 = (w >> 5) & 1u;
        return p2 ? Y_MAT : (p1 ? X_MAT : M_MAT);
    }
    // general row: X ptr bits 0-5 (mat | lord<<2), Y ptr bits 6-7 (mat), M ptr bits 8-13 (mat | lord<<2)
    if (mat == X_MAT) return w & 0x3fu;
    if (mat == Y_MAT) return (w >> 6) & 3u;
    return (w >> 8) & 0x3fu;
}

// Lane-kernel half-words: fast rows as above (bit 14 set).  General rows (bit 14 clear) keep the raw outcome bits
// of the Y chain where fast rows have them (bits 2-3), the X pointer (mat | ordinal << 2) in bits 4-9, and the M
// pointer split: source matrix in bits 0-1, edge ordinal in bits 10-13.
__host__ __device__ inline unsigned lane_decode_ptr(unsigned w, int mat) {
    if ((w & 0x4000u) || mat == Y_MAT) return strip_decode_ptr(w | 0x4000u, mat);
    if (mat == X_MAT) return (w >> 4) & 0x3fu;
    return (w & 3u) | (((w >> 10) & 0xfu) << 2);
}

}  // namespace pg2
