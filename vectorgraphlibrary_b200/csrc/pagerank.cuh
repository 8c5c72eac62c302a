// pagerank.cuh — internal interface of the PageRank sweep (pagerank.cu), shared with the partitioned driver
// (partition.cu). Not part of the C ABI.
#pragma once
#include "common.cuh"

#define PR_THREADS 256
#define PR_WARPS (PR_THREADS / 32)
#define PR_TASK_MAX_ROWS 31       // lanes 0..nrows hold the task's row boundaries
#define PR_MAX_PEERS 8
// tunables (profiles/r1_pr_ab.txt has the A/B runs); -D overrides are for developer A/B builds only
#ifndef PR_COLD_ID
#define PR_COLD_ID 49152          // ids at or above this are gathered without allocating in L1
#endif
#ifndef PR_TASK_EDGES
#define PR_TASK_EDGES 4096        // target edges per warp task
#endif
#ifndef PR_PIECE_EDGES
#define PR_PIECE_EDGES 8192       // rows with at least this many edges are cut into pieces of this size
#endif
#ifndef PR_MIN_CTAS
#define PR_MIN_CTAS 5             // __launch_bounds__ minimum resident CTAs per SM (48 registers)
#endif
#ifndef PR_VE_SEGS_PER_WARP
#define PR_VE_SEGS_PER_WARP 16    // 32-row segments of the padded tail copy handled by one warp
#endif
#define PR_ZERO_ROWS_PER_CTA 4096 // rows without out-edges handled by one CTA

// a warp's unit of work in the heavy region (rows with degree >= 32)
struct PrTask
{
    int64_t e0;         // first edge position
    int32_t e_len;      // number of edges
    int32_t row0;       // first row
    int32_t nrows;      // 1..31 complete rows, or 1 for a piece of a long row
    int32_t slot;       // -1: complete rows; else this piece's slot in the partial-sum array
    int32_t slot_first; // first slot of the long row
    int32_t npieces;    // pieces of the long row
};

struct PrParams
{
    const int64_t *ptr;
    const int32_t *adj;
    const float *contrib_in;       // indexed by column id (whole vector)
    const float *inv;              // indexed by local row
    float *contrib_out;            // indexed by local row (this rank's slice of the next vector)
    float *peer_out[PR_MAX_PEERS]; // the same slice inside the peers' copies of the next vector (partitioned graphs)
    float *rank_out;               // written on the final sweep only (may be NULL otherwise), indexed by local row
    const double *dangling_in;
    double *dangling_out;
    const PrTask *tasks;
    float *piece_partial;
    int32_t *piece_count;          // one arrival counter per long row (row < tier_border[0])
    int32_t ntasks;
    int32_t heavy_blocks;          // blocks [0, heavy_blocks) run warp tasks
    int32_t rows;                  // local rows
    int32_t col_of_row0;           // column id of local row 0 (0 on one GPU); row r is column col_of_row0 + r
    int32_t npeers;
    int32_t interleave;            // mix heavy and tail blocks in the grid
    float k, d, v_as_float;
    uint32_t vp, vp_mask;          // partitioned graphs: rows per rank (slice stride) and vp - 1 when vp is a power of two
    uint32_t cold_local;           // gathers of local rows at or above this bypass L1 (PR_COLD_ID / ranks)
    // tail: rows [tail_first, zero_first) have degree 1..31 and are read from the padded column-major copy
    const int32_t *ve_adj;         // segment s: ve_adj[ve_ptr[s] + j*32 + lane] = j-th neighbour of row tail_first + 32 s + lane
    const int64_t *ve_ptr;         // ve_segments + 1 offsets
    int32_t ve_segments;
    int32_t tail_first, zero_first;
    int32_t tail_blocks;           // blocks [heavy_blocks, heavy_blocks + tail_blocks) run tail segments, the rest zero rows
};

struct L2Pol
{
    uint64_t stream, keep;
};

int vglb_pr_prepare(vglb_ctx *ctx, vglb_graph *g, int iters);
int64_t vglb_pr_plan(const vglb_graph *g, int32_t rows, PrParams *P);
int vglb_pr_launch_sweep(vglb_ctx *ctx, const PrParams &P, int64_t nblocks);
