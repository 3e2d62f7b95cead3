// pg2_device.cuh -- device-side data layout shared by the fill, traceback and validation kernels.
//
// HBM layout of one launch batch (all arrays are flat, one allocation each, jobs index into them):
//   d_state[]   int32   Site::character_state of every distinct graph, back to back
//   d_off[]     int32   CSR offsets (graph-relative, n_sites+1 per graph)
//   d_estart[]  int32   Edge::start_site_index per backward edge, reference list order
//   d_elogw[]   float   Edge::log_posterior_weight per backward edge
//   d_vrow[]    int4    strip kernel row program: one entry per (DP row, backward edge) of a row graph
//   d_vlast[]   int32   per site of a row graph: index of the virtual row that completes the site
//   d_blo/d_bhi int32   clipped anchor band per row (banded jobs only; tunnel_matrix.h:194)
//   d_dlo       int32   first row on each anti-diagonal   (banded jobs only)
//   d_doff      int64   cell offset of each anti-diagonal (banded jobs only)
//   scores      double4 {X,Y,M,-} per in-band cell, ANTI-DIAGONAL-MAJOR (wavefront kernel scratch)
//   ptrs        uint32 / uint16 packed back-pointers per in-band cell (streamed once, read by traceback)
//   steps       uint32  packed pointers along each job's Viterbi path, walk order
#pragma once
#include <stdint.h>
#ifdef PG2_HOST_EMU
#include "pg2_emu_runtime.h"  // tests/emu: CPU test build only, never the product
#else
#include <cuda_runtime.h>
#endif

namespace pg2 {

constexpr int X_MAT = 0, Y_MAT = 1, M_MAT = 2, NO_MAT = 3;
constexpr unsigned FLAG_NO_TERMINAL_EDGES = 1u, FLAG_REDUCED = 2u;

// job status (mirrors PG2_JOB_* in include/pagan2_b200.h)
constexpr int JOB_OK = 0, JOB_NO_PATH = 1, JOB_BAD_BAND = 2, JOB_BAD_GRAPH = 3, JOB_BROKEN_PATH = 4,
              JOB_UNSUPPORTED = 5;

struct DevGraph {
    int n_sites;
    int state_base;  // into d_state
    int off_base;    // into d_off
    int edge_base;   // into d_estart / d_elogw
    int max_indeg;   // filled by the validation kernel
    int simple;      // 1: every site s>=1 has exactly one backward edge, from s-1 (plain leaf / read graph)
    int n_slots;     // saved-row slots the strip kernel needs when this graph is the row graph
    int zero_w;      // 1: every edge has log weight +0.0 (weight 1)
    int vrow_base;   // into d_vrow (int4 units); -1 until the graph is used as a strip row graph
    int n_vrows;     // virtual rows: one per (DP row, backward edge), edgeless rows count once
    int vlast_base;  // into d_vlast: per site, the virtual row that completes it
    int vplain_base; // into d_vlast: per block of LANE_B virtual rows, bit r set when row r is a plain interior row
    int implicit;    // 1: plain chain with unit weights whose CSR (d_off / d_estart / d_elogw) is generated on the device
    int np_base;     // into d_vlast: sorted list of the DP sites that are NOT plain (site 0, and every site whose backward
    int n_np;        //   edges are not exactly one edge from the site before it); -1 until a wavefront job uses the graph
    int npmask_base; // into d_vlast: bit s & 31 of word s >> 5 set when DP site s is plain
    // column program of a graph used as the COLUMN graph of the pipelined-strip kernel (pg2_pstrip_geom.cuh); -1 until built
    int cp_ci_base;  // into d_vlast: one info word per DP column (general / parked / end column, slots, block index)
    int cp_ei_base;  // into d_vlast: per backward edge (CSR order), the history slot of its source column
    int cp_blk_base; // into d_vlast: column range [c0, c1) of every block, 2 ints per block
    int cp_n_blocks;
    int cp_k;        // columns per lane the program was cut for (0: the graph cannot be a pstrip column graph)
    int cp_park;     // parked columns of the block that has most of them (sizes the kernel's shared-memory history)
    int max_span;    // longest edge into a DP site (site - start site); 1 for plain chains (host, classify_graph)
    int n_extra;     // backward edges beyond one per site (host, classify_graph): how far the graph is from a chain
};

struct DevModel {
    const float *table;  // [fas*fas] column-major log_score[l + r*fas]
    int fas;
    float open, ext, end_ext, brk, lng;
};

struct DevJob {
    int lx, ly;          // DP matrix dims = n_sites-1 of left / right (viterbi_alignment.cpp:243)
    int left, right;     // indices into the DevGraph array
    int model;           // index into the DevModel array
    unsigned flags;
    int banded;
    short kernel;        // 0 wavefront, 1 strip, 2 lanes, 3 pipelined strips, 4 band (warp per banded chain x chain job)
    short strip_general; // strip kernel: 1 = left graph needs the general row body, 0 = plain unit-weight chain
    long long band_base; // into d_blo / d_bhi (lx entries)
    long long diag_base; // into d_dlo / d_doff (lx+ly-1 entries)
    long long cell_base; // into the group's score / ptr buffers
    long long cells;     // in-band DP cells (the algorithmic count)
    long long ptr_cells; // pointer-buffer entries this job occupies in its kernel's layout
    long long step_base; // into the steps buffer
    int step_cap;
    int strip_k;         // strip / lane kernel: columns per lane
    int lane;            // lane kernel: the job's lane in its task (cell_base = the task's pointer region)
    int task;            // lane kernel: task index
    int max_diag;        // banded jobs: cells on the longest in-band anti-diagonal
    int n_blocks;        // pipelined-strip kernel: column blocks of the job
    int blk_base;        // pipelined-strip kernel: into d_vlast, PB_INTS ints per block
    int ps_ring;         // pipelined-strip kernel: virtual rows of the job's tallest block
    int n_seg;           // band kernel: walk segments (BAND_SEG diagonals each)
    long long b4_base;   // band kernel: the job's geometry record in d_band4 (pg2_band.cu: band_geo_off / band_roff_off / band_seg_off)
    long long cand_base; // band kernel: the job's walk candidates (int4 records)
    long long act_base;  // band kernel: per segment, the candidate the path enters it with and its word offset
};

constexpr int BAND_SEG = 256;       // diagonals per walk segment of the band kernel
constexpr int BAND_MAX_DIAG = 300;  // longest anti-diagonal the band kernel's row ring takes (pg2_band.cu: BAND_R, staging reach)

// Lane kernel work item: up to 32 alignments that share the LEFT (row) graph, model and flags; every right
// graph is a plain chain.  One warp takes one task, one alignment per lane.
struct LaneTask {
    int left;            // DevGraph index of the shared row graph
    int model;
    unsigned flags;
    int n_jobs;          // valid lanes (1..32)
    int max_ly;          // longest DP column count among the lanes
    int variant;         // bit 0 general row body, bit 1 shared small table, bit 2 right graphs carry weights
    long long ptr_base;  // into the group's pointer buffer (half-words)
    int job_ids[32];
};

struct DevResult {
    double score;
    unsigned end_ptr;
    int n_steps;
    int status;
    int pad;
};

// ---- packed back-pointers -------------------------------------------------------------------
// API encoding of ONE pointer: bits 0-1 source matrix, 2-7 left edge ordinal, 8-13 right edge ordinal.
__host__ __device__ inline unsigned pack_ptr(int mat, int lord, int rord) {
    return (unsigned)mat | ((unsigned)lord << 2) | ((unsigned)rord << 8);
}
// wavefront kernel cell word: X pointer in bits 0-7 (mat | lord<<2), Y pointer in bits 8-15
// (mat | rord<<2), M pointer in bits 16-29 (mat | lord<<2 | rord<<8).
__host__ __device__ inline unsigned cell_word(unsigned px, unsigned py, unsigned pm) {
    return (px & 0xffu) | ((py & 0xffu) << 8) | ((pm & 0x3fffu) << 16);
}
// bits 30 / 31 of a cell word: left site i / right site j has exactly one backward edge, from the site before it
constexpr unsigned WORD_PLAIN_LEFT = 1u << 30, WORD_PLAIN_RIGHT = 1u << 31;
__host__ __device__ inline unsigned word_ptr(unsigned w, int mat) {
    if (mat == X_MAT) return w & 0xffu;                                    // mat | lord<<2
    if (mat == Y_MAT) { unsigned y = (w >> 8) & 0xffu; return (y & 3u) | ((y >> 2) << 8); }
    return (w >> 16) & 0x3fffu;
}

// ---- anti-diagonal-major cell indexing (unbanded closed form) --------------------------------
// cells with i+j < s in an n x m matrix
__host__ __device__ inline long long diag_cum(long long s, long long n, long long m) {
    long long a = n < m ? n : m, b = n < m ? m : n;
    if (s <= a) return s * (s + 1) / 2;
    if (s <= b) return a * (a + 1) / 2 + (s - a) * a;
    long long t = s - b;
    return a * (a + 1) / 2 + (b - a) * a + t * (a - 1) - t * (t - 1) / 2;
}
__host__ __device__ inline int diag_lo(int s, int m) { return s - (m - 1) > 0 ? s - (m - 1) : 0; }
__host__ __device__ inline int diag_hi(int s, int n) { return s < n - 1 ? s : n - 1; }

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }

}  // namespace pg2
