// pg2_pstrip_geom.cuh -- geometry of the pipelined-strip fill kernel (pg2_pstrip.cu), shared by the kernel, the engine's
// packing (column programs, block tables, scratch accounting) and the traceback (cell lookup).
//
// One CTA aligns one job.  The columns of the RIGHT graph are cut into BLOCKS of at most 32*K consecutive columns; a
// block is swept by one warp exactly like a block of the warp-per-alignment strip kernel (pg2_strip.cu: lane l owns K
// columns, the lanes are skewed by one virtual row, the strip's last column travels by shuffle), and consecutive blocks
// are PIPELINED over the warps of the CTA: the warp of block b+1 follows the warp of block b through a boundary
// column in global memory guarded by a progress counter in shared memory.  Both graphs may be general:
//   rows     the left graph's row program (d_vrow: one virtual row per backward edge, parked rows for long spans),
//   columns  the right graph's COLUMN PROGRAM below: a column whose backward edges are not exactly one edge from the
//            column before it is GENERAL and reads its sources (i, pr) / (i-1, pr) from a small shared-memory history
//            of the PARKED columns (columns that are the source of an edge into a general column), (pl, pr) of a
//            long-span left edge from the parked row.
// Blocks start at CUT POINTS of the right graph (no edge other than (c-1 -> c) crosses into the block), so every
// source of a general column lies in its own block; a block holds at most PS_MAX_PARK parked columns.
// An anchor band restricts every block to the rows that meet it; cells outside the band are forced to -inf.
#pragma once
#include "pg2_device.cuh"

namespace pg2 {

constexpr int PS_MAX_WARPS = 12;     // warps per CTA = blocks of one job in flight
constexpr int PS_HIST = 16;          // history depth (sites) of the parked columns: lane distance of an edge <= PS_HIST - 2
constexpr int PS_MAX_PARK = 64;      // parked columns per block at most (7 slot bits); a graph's own maximum sizes the history
constexpr int PS_MAX_END = 4;        // distinct columns the end corner reads (predecessors of the right stop site and ly-1)
constexpr int PS_MAX_RIGHT_INDEG = 63;
constexpr int PS_SMEM_BUDGET = 200 * 1024;  // shared memory of one CTA the histories may take
constexpr int PS_PREFETCH = 4;       // steps the boundary column is fetched ahead

// column info word (one per DP column of a right graph, d_vlast pool):
//   bit 0      GENERAL: the column's backward edges are not exactly one edge from the column before it
//   bit 1      PARKED: some general column reads this column; bits 5-11 its slot in the block's history
//   bit 2      ENDCOL: the end corner reads this column; bits 3-4 its end slot
//   bits 12-   index of the column's block
constexpr int PC_GENERAL = 1, PC_PARKED = 2, PC_ENDCOL = 4;
constexpr int PC_END_SHIFT = 3, PC_SLOT_SHIFT = 5, PC_SLOT_MASK = 127, PC_BLOCK_SHIFT = 12;

// One block of a job (d_vlast pool, 6 ints per block, job by job):
//   c0, c1     columns [c0, c1)
//   v0, v1     virtual rows [v0, v1) of the row program the block sweeps (all of them without a band)
//   i0         first DP row (site) of the block: rows above it are outside the band for every column of the block
//   ptr_off    offset of the block's pointer words in the job's region (32-bit words)
constexpr int PB_INTS = 6;

// pointer words of one block: [step][lane][K], step = (v - v0) + lane
__host__ __device__ inline long long ps_block_words(int v0, int v1, int K) { return (long long)(v1 - v0 + 31) * 32 * K; }
__host__ __device__ inline long long ps_word_index(int ptr_off, int v0, int c0, int K, int v, int j) {
    const int jj = j - c0, l = jj / K, k = jj - l * K;
    return (long long)ptr_off + ((long long)(v - v0 + l) * 32 + l) * K + k;
}

// 32-bit pointer word of this kernel: bits 0-29 as the wavefront kernel's cell word (X pointer mat | lord<<2 in bits
// 0-7, Y pointer mat | rord<<2 in bits 8-15, M pointer mat | lord<<2 | rord<<8 in bits 16-29); bit 30: the row is a
// plain row (its only left edge comes from the row above, which sits one virtual row higher); bit 31: the column is a
// plain column (its only right edge comes from the column before it).
constexpr unsigned PSW_PLAIN_ROW = 1u << 30, PSW_PLAIN_COL = 1u << 31;

// per-warp global scratch of the kernel, in double4 units: the boundary column ring [ring] (written by the warp for
// the warp of the next block), parked rows [n_slots][32*K]
__host__ __device__ inline long long ps_warp_double4(int ring, int n_slots, int K) {
    return (long long)ring + (long long)(n_slots > 0 ? n_slots : 1) * 32 * K;
}

}  // namespace pg2
