// pg2_strip_geom.cuh -- pointer-buffer geometry of the strip kernel, shared by the fill kernel, the
// traceback kernel and the engine's scratch accounting.
#pragma once
#include "pg2_device.cuh"

namespace pg2 {

constexpr int STRIP_MAX_LEFT_INDEG = 16;

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

}  // namespace pg2
