// pg2_strip_geom.cuh -- pointer-buffer geometry of the strip kernel, shared by the fill kernel, the
// traceback kernel and the engine's scratch accounting.
#pragma once
#include "pg2_device.cuh"

namespace pg2 {

constexpr int STRIP_MAX_LEFT_INDEG = 16;

// d_rowinfo[]: one word per site of a graph used as the ROW graph of the strip kernel
//   bits 0-11  character state (0 for the start/stop sites)
//   bit  12    fast row: exactly one backward edge, from the site just above (also set for site 0)
//   bit  13    every backward edge of the site has log weight +0.0
//   bit  14    the end corner reads this row's last column
//   bits 16-31 saved-row slot + 1 when the site is the source of a long-span edge, else 0
constexpr int ROWINFO_STATE_MASK = 0xfff;
constexpr int ROWINFO_FAST = 1 << 12;
constexpr int ROWINFO_ZERO_W = 1 << 13;
constexpr int ROWINFO_ENDPRED = 1 << 14;  // bit 14: the end corner reads this row (predecessor of the stop site, or the last row)
constexpr int STRIP_SMALL_FAS = 16;       // alphabets up to this size use the shared double2 table (DNA: 15)
constexpr int ROWINFO_SLOT_SHIFT = 16;

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
__host__ __device__ inline long long strip_cells(int lx, int ly, int K) {
    int W = 32 * K;
    long long blocks = (ly + W - 1) / W;
    return blocks * (long long)(lx + 31) * 32 * strip_ks(K);
}
__host__ __device__ inline long long strip_ptr_index(int lx, int ly, int K, int i, int j) {
    int W = 32 * K;
    int b = j / W, jj = j - b * W;
    int l = jj / K, k = jj - l * K;
    return (((long long)b * (lx + 31) + (i + l)) * 32 + l) * strip_ks(K) + k;
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

}  // namespace pg2
