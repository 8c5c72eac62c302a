// bfs.cu — direction-optimising BFS over the device VectCSR (sm_100a): top-down advance with degree-binned sparse
// queues, bottom-up advance over a dense bitmap, and the conversions between the two frontier forms.
//
// Reference: BFS::fast_vgl_top_down (algorithms/bfs/bfs.hpp:5-51) = per level one scatter (advance) + one
// generate_new_frontier; the DO heuristic gpu_change_state (algorithms/bfs/change_state/change_state.hpp:100-141,
// ALPHA 15 / BETA 18 :5-6); the dead hand-wired DO code (hardwired_do_bfs.hpp:859-1009) for the loop structure.
// Levels are unique whatever direction is taken: source = 1, unreachable = -1 (change_state.h:21-23).
//
// B200 design
//   * the next frontier is emitted INSIDE the advance (the thread that wins the atomicOr on the visited bitmap
//     enqueues), instead of the reference's separate 2-pass generate_new_frontier over all V
//     (multicore/generate_new_frontier.hpp:35-108 = 34 % of its BFS time);
//   * sparse frontier = three queues binned by degree tier; the tier of a vertex is a comparison of its id with two
//     borders because ids are degree-sorted: CTA per vertex (deg >= 4096), warp per vertex (deg >= 32), 8 lanes per
//     vertex (the tail) — replaces the ve / vc / collective kernels of multicore/advance_sparse.hpp:7-249;
//   * dense frontier = bitmap (V/8 bytes: 8 MiB at scale 26, L2-resident); bottom-up assigns one warp to one 32-vertex
//     word of the visited bitmap, so next-frontier and visited words are written without atomics; each lane probes the
//     first few in-neighbours of its vertex, longer rows are finished warp-cooperatively with ballot early exit;
//     in-rows list sources in ascending id = descending out-degree, so the likeliest parents come first;
//   * queue <-> bitmap conversion: ballot/popc + warp-aggregated atomics (north_star (c)).
// One host read-back of 8 counters per level drives termination and the direction switch.
#include <stdlib.h>

#include "common.cuh"
#include "frontier.cuh"

#define BFS_THREADS 256
#define BFS_SMALL_LANES 8
#define BFS_BU_PROBE 4

// NT threads (a CTA, a warp or an 8-lane group) expand the out-row [s,e) of one frontier vertex.
// Every thread of the warp must call this together (ballots inside); `e <= s` for idle groups.
// PART = the graph is one rank's part: `visited` is the replicated bitmap (read only here) and discoveries are only
// marked in the candidate bitmap `levels` points to (reinterpreted) — owners resolve them after the exchange.
template <int NT, bool PART>
__device__ __forceinline__ void td_expand(const int32_t *__restrict__ adj, int64_t s, int64_t e, int tid,
                                          uint32_t *__restrict__ visited, int32_t *__restrict__ levels, int32_t next_level,
                                          int32_t b0, int32_t b1, const TierQueues &nq, unsigned long long *counters)
{
    for (int64_t p = s + tid;; p += NT)
    {
        const bool active = p < e;
        if (!__any_sync(0xffffffffu, active)) break;
        bool won = false;
        int32_t v = 0;
        if (active)
        {
            v = adj[p];
            const uint32_t bit = 1u << (v & 31);
            if (PART)
            {
                uint32_t *cand = reinterpret_cast<uint32_t *>(levels);
                if (!(visited[v >> 5] & bit) && !(cand[v >> 5] & bit)) atomicOr(&cand[v >> 5], bit);
            }
            else
            {
                if (!(visited[v >> 5] & bit))
                {
                    const uint32_t old = atomicOr(&visited[v >> 5], bit);
                    won = !(old & bit);
                }
                if (won) levels[v] = next_level;
            }
        }
        if (!PART) enqueue_binned(won, v, b0, b1, nq, counters);
    }
}

template <bool PART>
__global__ void __launch_bounds__(BFS_THREADS)
bfs_td_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ adj, TierQueues cq, int32_t n_big, int32_t n_mid,
              int32_t n_small, int32_t blocks_mid, int32_t blocks_small, uint32_t *__restrict__ visited,
              int32_t *__restrict__ levels, int32_t next_level, int32_t b0, int32_t b1, TierQueues nq,
              unsigned long long *counters)
{
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long edges = 0;
    if (b < n_big)
    {
        const int32_t u = cq.q[0][b];
        const int64_t s = ptr[u], e = ptr[u + 1];
        if (threadIdx.x == 0) edges = e - s;
        td_expand<BFS_THREADS, PART>(adj, s, e, threadIdx.x, visited, levels, next_level, b0, b1, nq, counters);
    }
    else if (b < n_big + blocks_mid)
    {
        const int nwarps = blocks_mid * (BFS_THREADS / 32);
        for (int i = (b - n_big) * (BFS_THREADS / 32) + warp; i < n_mid; i += nwarps)
        {
            const int32_t u = cq.q[1][i];
            const int64_t s = ptr[u], e = ptr[u + 1];
            if (lane == 0) edges += e - s;
            td_expand<32, PART>(adj, s, e, lane, visited, levels, next_level, b0, b1, nq, counters);
        }
    }
    else
    {
        constexpr int G = BFS_SMALL_LANES;
        constexpr int GROUPS = BFS_THREADS / G;
        const int ngroups = blocks_small * GROUPS;
        const int gid = threadIdx.x / G, gl = threadIdx.x % G;
        // all groups of a warp iterate together (ballots in td_expand): pad the trip count to a multiple of the stride
        const int first = (b - n_big - blocks_mid) * GROUPS + gid;
        for (int base = first - gid; base < n_small; base += ngroups)
        {
            const int i = base + gid;
            int64_t s = 0, e = 0;
            if (i < n_small)
            {
                const int32_t u = cq.q[2][i];
                s = ptr[u];
                e = ptr[u + 1];
                if (gl == 0) edges += e - s;
            }
            td_expand<G, PART>(adj, s, e, gl, visited, levels, next_level, b0, b1, nq, counters);
        }
    }
    edges = warp_sum_i64(edges);
    if (lane == 0 && edges) atomicAdd(&counters[C_EDGES], (unsigned long long)edges);
}

// sum of out-degrees of the freshly built next frontier (m_f of the direction heuristic)
__global__ void bfs_queue_degree_kernel(const int64_t *__restrict__ ptr, TierQueues q, int32_t n0, int32_t n1, int32_t n2,
                                        unsigned long long *counters)
{
    const int n = n0 + n1 + n2;
    long long sum = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int32_t u = i < n0 ? q.q[0][i] : (i < n0 + n1 ? q.q[1][i - n0] : q.q[2][i - n0 - n1]);
        sum += ptr[u + 1] - ptr[u];
    }
    sum = warp_sum_i64(sum);
    if ((threadIdx.x & 31) == 0 && sum) atomicAdd(&counters[C_MF], (unsigned long long)sum);
}

// bottom-up step: one warp per 32-vertex word of the visited bitmap
__global__ void __launch_bounds__(BFS_THREADS)
bfs_bu_kernel(const int64_t *__restrict__ in_ptr, const int32_t *__restrict__ in_adj, int32_t V,
              uint32_t *__restrict__ visited, const uint32_t *__restrict__ cur_bm, uint32_t *__restrict__ next_bm,
              int32_t *__restrict__ levels, int32_t next_level, unsigned long long *counters)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nwords = ((int64_t)V + 31) >> 5;
    long long edges = 0, rows = 0;
    int found_total = 0;
    for (int64_t w = warp; w < nwords; w += nwarps)
    {
        const uint32_t vis = visited[w];
        const int32_t v = (int32_t)(w << 5) + lane;
        const uint32_t valid = (w == nwords - 1 && (V & 31)) ? ((1u << (V & 31)) - 1u) : 0xffffffffu;
        const uint32_t unv = ~vis & valid;
        if (unv == 0)
        {
            if (lane == 0) next_bm[w] = 0;
            continue;
        }
        const bool mine = (unv >> lane) & 1u;
        bool found = false;
        int64_t s = 0, e = 0;
        if (mine)
        {
            s = in_ptr[v];
            e = in_ptr[v + 1];
            rows++;
            const int64_t pe = min(e, s + BFS_BU_PROBE);
            for (int64_t p = s; p < pe; p++)
            {
                edges++;
                if (bm_test(cur_bm, in_adj[p]))
                {
                    found = true;
                    break;
                }
            }
        }
        // rows longer than the probe: finished by the whole warp, 32 neighbours at a time
        unsigned pending = __ballot_sync(0xffffffffu, mine && !found && (s + BFS_BU_PROBE < e));
        while (pending)
        {
            const int src_lane = __ffs(pending) - 1;
            pending &= pending - 1;
            const int64_t ps = __shfl_sync(0xffffffffu, s, src_lane) + BFS_BU_PROBE;
            const int64_t pe = __shfl_sync(0xffffffffu, e, src_lane);
            bool hit = false;
            for (int64_t p0 = ps; p0 < pe; p0 += 32)
            {
                const int64_t p = p0 + lane;
                bool h = false;
                if (p < pe)
                {
                    edges++;
                    h = bm_test(cur_bm, in_adj[p]);
                }
                if (__any_sync(0xffffffffu, h))
                {
                    hit = true;
                    break;
                }
            }
            if (lane == src_lane && hit) found = true;
        }
        const uint32_t fmask = __ballot_sync(0xffffffffu, found);
        if (found) levels[v] = next_level;
        if (lane == 0)
        {
            next_bm[w] = fmask;
            if (fmask) visited[w] = vis | fmask;
            found_total += __popc(fmask);
        }
    }
    edges = warp_sum_i64(edges);
    rows = warp_sum_i64(rows);
    if (lane == 0)
    {
        if (edges) atomicAdd(&counters[C_EDGES], (unsigned long long)edges);
        if (rows) atomicAdd(&counters[C_ROWS], (unsigned long long)rows);
        if (found_total) atomicAdd(&counters[C_FOUND], (unsigned long long)found_total);
    }
}

// sparse queues -> dense bitmap (bitmap must be zeroed before)
__global__ void bfs_queue_to_bitmap_kernel(TierQueues q, int32_t n0, int32_t n1, int32_t n2, uint32_t *__restrict__ bm)
{
    const int n = n0 + n1 + n2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int32_t u = i < n0 ? q.q[0][i] : (i < n0 + n1 ? q.q[1][i - n0] : q.q[2][i - n0 - n1]);
        atomicOr(&bm[u >> 5], 1u << (u & 31));
    }
}

// dense bitmap -> degree-binned sparse queues: popc per word, warp scan, one atomic per warp and tier
__global__ void bfs_bitmap_to_queue_kernel(const uint32_t *__restrict__ bm, int32_t V, int32_t b0, int32_t b1, TierQueues q,
                                           unsigned long long *counters)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwords = ((int64_t)V + 31) >> 5;
    const int64_t nwords_padded = (nwords + 31) & ~(int64_t)31;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords_padded; w += (int64_t)gridDim.x * blockDim.x)
    {
        const uint32_t word = w < nwords ? bm[w] : 0u;
        if (!__any_sync(0xffffffffu, word != 0)) continue;
        const int32_t base = (int32_t)(w << 5);
        const uint32_t m0 = below_border_mask(base, b0);
        const uint32_t m1 = below_border_mask(base, b1);
        const uint32_t part[3] = {word & m0, word & m1 & ~m0, word & ~m1};
#pragma unroll
        for (int t = 0; t < 3; t++)
        {
            const int cnt = __popc(part[t]);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total == 0) continue;
            unsigned long long wbase = 0;
            if (lane == 0) wbase = atomicAdd(&counters[C_NEXT_BIG + t], (unsigned long long)total);
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            int64_t pos = (int64_t)wbase + incl - cnt;
            uint32_t bits = part[t];
            while (bits)
            {
                const int bpos = __ffs(bits) - 1;
                bits &= bits - 1;
                q.q[t][pos++] = base + bpos;
            }
        }
    }
}

__global__ void bfs_init_kernel(int32_t *levels, uint32_t *visited, int32_t source, int32_t *queue_slot)
{
    levels[source] = VGLB_FIRST_LEVEL_VERTEX;
    visited[source >> 5] = 1u << (source & 31);
    *queue_slot = source;
}


// ---- 1D-partitioned BFS (one rank's part; every rank runs the same host loop on the same allreduced counters) -------
// Reference: the NEC backend replicates levels and MPI_Allreduce(MAX)es the whole int array after every advance
// (vgl_compute_api/common/mpi_exchange.hpp:155-271). Here only bitmaps travel:
//   top-down : each rank expands the frontier vertices it owns and marks every unvisited destination in a full-length
//              candidate bitmap; an all-to-all hands slice q of every rank's candidates to rank q, which ORs them, masks
//              its visited slice, writes the levels it owns and queues its new frontier vertices (bfs_combine_kernel);
//   bottom-up: each rank scans the in-rows of the unvisited vertices it owns against the replicated frontier bitmap
//              (same kernel as on one GPU, pointed at the owned slice);
//   both     : the owned slices of the new frontier are allgathered (V/8 bytes in total) and ORed into the replicated
//              visited bitmap; {found, m_f, edges, rows} are allreduced and drive termination and the alpha/beta switch.

__global__ void bfs_part_init_kernel(uint32_t *visited, uint32_t *cur_bm, int32_t source_col, int32_t *levels_local,
                                     int32_t local_row /* -1 unless owned */, int32_t *queue_slot)
{
    visited[source_col >> 5] = 1u << (source_col & 31);
    cur_bm[source_col >> 5] = 1u << (source_col & 31);
    if (local_row >= 0)
    {
        levels_local[local_row] = VGLB_FIRST_LEVEL_VERTEX;
        *queue_slot = local_row;
    }
}

// one thread per word of the owned slice: OR of the P received candidate slices, minus the visited ones
__global__ void __launch_bounds__(256)
bfs_combine_kernel(const uint32_t *__restrict__ stage, int32_t P, int32_t wslice, const uint32_t *__restrict__ visited_slice,
                   uint32_t *__restrict__ next_slice, const int64_t *__restrict__ ptr, int32_t *__restrict__ levels_local,
                   int32_t next_level, int32_t b0, int32_t b1, TierQueues nq, unsigned long long *counters)
{
    const int lane = threadIdx.x & 31;
    const int32_t wpad = (wslice + 31) & ~31;
    long long mf = 0;
    int found = 0;
    for (int32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < wpad; w += gridDim.x * blockDim.x)
    {
        uint32_t bits = 0;
        if (w < wslice)
        {
            for (int p = 0; p < P; p++) bits |= stage[(int64_t)p * wslice + w];
            bits &= ~visited_slice[w];
            next_slice[w] = bits;
            found += __popc(bits);
        }
        while (__any_sync(0xffffffffu, bits != 0))
        {
            const bool has = bits != 0;
            int32_t row = 0;
            if (has)
            {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                row = (w << 5) + b;
                levels_local[row] = next_level;
                mf += ptr[row + 1] - ptr[row];
            }
            enqueue_binned(has, row, b0, b1, nq, counters);
        }
    }
    mf = warp_sum_i64(mf);
    found = (int)warp_sum_i64(found);
    if (lane == 0)
    {
        if (mf) atomicAdd(&counters[C_MF], (unsigned long long)mf);
        if (found) atomicAdd(&counters[C_FOUND], (unsigned long long)found);
    }
}

__global__ void bfs_or_kernel(uint32_t *__restrict__ visited, const uint32_t *__restrict__ next_bm, int64_t words)
{
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (int64_t)gridDim.x * blockDim.x)
    {
        const uint32_t n = next_bm[w];
        if (n) visited[w] |= n;
    }
}

__global__ void bfs_copy_counters_kernel(unsigned long long *c)
{
    if (threadIdx.x < C_COUNT) c[C_COUNT + threadIdx.x] = c[threadIdx.x];
}

static int bfs_partitioned(vglb_ctx *ctx, vglb_graph *g, int32_t source, int32_t *d_levels, const vglb_bfs_opts *opts,
                           vglb_stats *stats)
{
    VGLB_REQUIRE(source >= 0 && source < g->cols, "vglb_bfs: source column out of range");
    const bool dopt = opts && opts->direction_optimising;
    if (dopt && !g->d_in_ptr)
    {
        vglb_set_error("vglb_bfs: direction-optimising BFS needs a graph built with VGLB_GRAPH_WITH_INCOMING");
        return VGLB_EINVAL;
    }
    const long long alpha = (opts && opts->alpha > 0) ? opts->alpha : 15;
    const long long beta = (opts && opts->beta > 0) ? opts->beta : 18;
    CUDA_TRY(cudaSetDevice(ctx->device));
    vglb_comm *comm = g->comm;
    const int32_t P = g->part_world, rank = g->part_rank, vp = g->vp, rows = g->V;
    const int32_t wslice = vp / 32;
    const int64_t words = g->cols / 32, my = (int64_t)rank * wslice;
    for (int i = 0; i < 3; i++)
        if (!g->d_part_bm[i]) CUDA_TRY(cudaMalloc(&g->d_part_bm[i], (size_t)(words + 32) * 4));
    if (!g->d_part_stage) CUDA_TRY(cudaMalloc(&g->d_part_stage, (size_t)(words + 32) * 4));
    if (!g->d_queue[0]) CUDA_TRY(cudaMalloc(&g->d_queue[0], ((size_t)vp + 3) * 4));
    const int64_t launches0 = ctx->launches;
    const int32_t b0 = g->tier_border[0], b1 = g->tier_border[1];
    unsigned long long *d_cnt = (unsigned long long *)ctx->d_counters;
    unsigned long long *h_cnt = (unsigned long long *)ctx->h_counters;
    cudaStream_t st = ctx->stream;
    TierQueues cq;
    cq.q[0] = g->d_queue[0];
    cq.q[1] = g->d_queue[0] + b0;
    cq.q[2] = g->d_queue[0] + b1;
    uint32_t *visited = g->d_part_bm[0], *cur_bm = g->d_part_bm[1], *next_bm = g->d_part_bm[2];

    CUDA_TRY(cudaEventRecord(ctx->ev_start, st));
    if (rows > 0) CUDA_TRY(cudaMemsetAsync(d_levels, 0xFF, (size_t)rows * 4, st));
    CUDA_TRY(cudaMemsetAsync(visited, 0, (size_t)words * 4, st));
    CUDA_TRY(cudaMemsetAsync(cur_bm, 0, (size_t)words * 4, st));
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 2 * C_COUNT * 8, st));
    const bool own_source = source / vp == rank;
    const int32_t src_row = own_source ? source - rank * vp : -1;
    VGLB_REQUIRE(!own_source || src_row < rows, "vglb_bfs: source is a padding column");
    const int src_tier = src_row < 0 ? 0 : (src_row < b0 ? 0 : (src_row < b1 ? 1 : 2));
    bfs_part_init_kernel<<<1, 1, 0, st>>>(visited, cur_bm, source, d_levels, src_row, cq.q[src_tier]);
    KERNEL_TRY();
    ctx->launches++;

    int32_t n[3] = {0, 0, 0};
    if (own_source) n[src_tier] = 1;
    long long n_cur = 1, visited_total = 1;
    bool bottom_up = false;
    int32_t level = VGLB_FIRST_LEVEL_VERTEX;
    int64_t tot_edges = 0, tot_rows = 0, tot_frontier_bytes = 0, levels_run = 0;
    int32_t bu_levels = 0;
    const long long Vg = g->V_orig;
    const long long factor = (g->E_global / (Vg > 0 ? Vg : 1)) / 2 > 0 ? (g->E_global / Vg) / 2 : 1;
    const int max_blocks = ctx->sm_count * 16;
    int rc;

    while (n_cur > 0)
    {
        if (!bottom_up)
        {
            CUDA_TRY(cudaMemsetAsync(next_bm, 0, (size_t)words * 4, st)); // candidates
            const int blocks_mid = (int)min((int64_t)max_blocks, ceil_div64(n[1], BFS_THREADS / 32));
            const int blocks_small = (int)min((int64_t)max_blocks, ceil_div64(n[2], BFS_THREADS / BFS_SMALL_LANES));
            const int64_t grid = (int64_t)n[0] + blocks_mid + blocks_small;
            if (grid > 0)
            {
                bfs_td_kernel<true><<<(unsigned)grid, BFS_THREADS, 0, st>>>(g->d_out_ptr, g->d_out_adj, cq, n[0], n[1], n[2], blocks_mid,
                                                                          blocks_small, visited, (int32_t *)next_bm, level + 1, b0, b1,
                                                                          cq, d_cnt);
                KERNEL_TRY();
                ctx->launches++;
            }
            rc = vglb_comm_alltoall_async(comm, next_bm, g->d_part_stage, (size_t)wslice * 4);
            if (rc != VGLB_OK) return rc;
            bfs_combine_kernel<<<(unsigned)min((int64_t)max_blocks, ceil_div64(wslice, 256)), 256, 0, st>>>(
                g->d_part_stage, P, wslice, visited + my, next_bm + my, g->d_out_ptr, d_levels, level + 1, b0, b1, cq, d_cnt);
            KERNEL_TRY();
            ctx->launches++;
            tot_frontier_bytes += (int64_t)wslice * 4 * (2 * P + 2);
        }
        else
        {
            CUDA_TRY(cudaMemsetAsync(next_bm + my, 0, (size_t)wslice * 4, st));
            bfs_bu_kernel<<<ctx->sm_count * 8, BFS_THREADS, 0, st>>>(g->d_in_ptr, g->d_in_adj, rows, visited + my, cur_bm,
                                                                    next_bm + my, d_levels, level + 1, d_cnt);
            KERNEL_TRY();
            ctx->launches++;
            bu_levels++;
            tot_frontier_bytes += (int64_t)wslice * 4 * 3;
        }
        rc = vglb_comm_allgather_async(comm, next_bm, (size_t)wslice * 4);
        if (rc != VGLB_OK) return rc;
        bfs_or_kernel<<<(unsigned)min((int64_t)max_blocks, ceil_div64(words, 256)), 256, 0, st>>>(visited, next_bm, words);
        KERNEL_TRY();
        bfs_copy_counters_kernel<<<1, 32, 0, st>>>(d_cnt);
        KERNEL_TRY();
        ctx->launches += 2;
        rc = vglb_comm_allreduce_async(comm, d_cnt + C_COUNT, C_COUNT, VGLB_DT_I64, VGLB_OP_SUM);
        if (rc != VGLB_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, 2 * C_COUNT * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 2 * C_COUNT * 8, st));
        levels_run++;
        const unsigned long long *gl = h_cnt + C_COUNT; // global sums
        const long long n_next = (long long)gl[C_FOUND];
        tot_edges += (int64_t)h_cnt[C_EDGES];
        tot_rows += bottom_up ? (int64_t)h_cnt[C_ROWS] : (int64_t)n[0] + n[1] + n[2];
        tot_frontier_bytes += words * 4 * 3; // allgathered frontier written, ORed into visited
        visited_total += n_next;
        if (n_next == 0) break;

        bool next_bu = bottom_up;
        if (dopt)
        {
            const long long unvisited = Vg - visited_total;
            if (!bottom_up && n_cur < n_next)
            {
                if ((long long)gl[C_MF] >= (unvisited * factor + Vg) / alpha) next_bu = true;
            }
            else if (bottom_up && n_cur >= n_next)
            {
                if (n_next < (unvisited * factor + Vg) / (factor * beta)) next_bu = false;
            }
        }
        if (!next_bu)
        {
            if (!bottom_up)
            {
                n[0] = (int32_t)h_cnt[C_NEXT_BIG]; n[1] = (int32_t)h_cnt[C_NEXT_MID]; n[2] = (int32_t)h_cnt[C_NEXT_SMALL];
            }
            else
            {
                bfs_bitmap_to_queue_kernel<<<(unsigned)min((int64_t)max_blocks, ceil_div64(wslice, 256)), 256, 0, st>>>(
                    next_bm + my, rows, b0, b1, cq, d_cnt);
                KERNEL_TRY();
                ctx->launches++;
                CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, 3 * 8, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                n[0] = (int32_t)h_cnt[0]; n[1] = (int32_t)h_cnt[1]; n[2] = (int32_t)h_cnt[2];
                CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 2 * C_COUNT * 8, st));
            }
        }
        uint32_t *t = cur_bm; cur_bm = next_bm; next_bm = t;
        bottom_up = next_bu;
        n_cur = n_next;
        level++;
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_stop, st));
    CUDA_TRY(cudaEventSynchronize(ctx->ev_stop));
    if (stats)
    {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
        memset(stats, 0, sizeof(*stats));
        stats->seconds = ms * 1e-3;
        stats->iterations = levels_run;
        stats->edges_inspected = tot_edges;      // this rank's
        stats->vertices_processed = tot_rows;
        stats->frontier_bytes = tot_frontier_bytes;
        stats->algorithmic_bytes = 8 * tot_edges + 12 * tot_rows + tot_frontier_bytes;
        stats->kernel_launches = ctx->launches - launches0;
        stats->bottom_up_levels = bu_levels;
    }
    return VGLB_OK;
}

static int bfs_prepare(vglb_ctx *ctx, vglb_graph *g)
{
    if (g->bfs_ready) return VGLB_OK;
    const size_t words = ((size_t)g->V + 31) / 32 + 32;
    CUDA_TRY(cudaMalloc(&g->d_visited, words * 4));
    CUDA_TRY(cudaMalloc(&g->d_front_bm[0], words * 4));
    CUDA_TRY(cudaMalloc(&g->d_front_bm[1], words * 4));
    CUDA_TRY(cudaMalloc(&g->d_queue[0], ((size_t)g->V + 3) * 4));
    CUDA_TRY(cudaMalloc(&g->d_queue[1], ((size_t)g->V + 3) * 4));
    g->bfs_ready = 1;
    return VGLB_OK;
}

extern "C" int vglb_bfs(vglb_ctx *ctx, vglb_graph *g, int32_t source, int32_t *d_levels, const vglb_bfs_opts *opts,
                        vglb_stats *stats)
{
    VGLB_REQUIRE(ctx != NULL && g != NULL && d_levels != NULL, "vglb_bfs: NULL argument");
    if (g->comm) return bfs_partitioned(ctx, g, source, d_levels, opts, stats);
    VGLB_REQUIRE(source >= 0 && source < g->V, "vglb_bfs: source out of range");
    const bool dopt = opts && opts->direction_optimising;
    if (dopt && !g->d_in_ptr)
    {
        vglb_set_error("vglb_bfs: direction-optimising BFS needs a graph built with VGLB_GRAPH_WITH_INCOMING");
        return VGLB_EINVAL;
    }
    const long long alpha = (opts && opts->alpha > 0) ? opts->alpha : 15; // change_state.hpp:5-6
    const long long beta = (opts && opts->beta > 0) ? opts->beta : 18;
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = bfs_prepare(ctx, g);
    if (rc != VGLB_OK) return rc;
    const int64_t launches0 = ctx->launches;
    const int32_t V = g->V;
    const int32_t b0 = g->tier_border[0], b1 = g->tier_border[1];
    const size_t words = ((size_t)V + 31) / 32;
    unsigned long long *d_cnt = (unsigned long long *)ctx->d_counters;
    unsigned long long *h_cnt = (unsigned long long *)ctx->h_counters;
    cudaStream_t st = ctx->stream;

    // queue regions: capacity of a tier's queue = population of the tier
    auto regions = [&](int32_t *base) {
        TierQueues q;
        q.q[0] = base;
        q.q[1] = base + b0;
        q.q[2] = base + b1;
        return q;
    };
    TierQueues cq = regions(g->d_queue[0]), nq = regions(g->d_queue[1]);
    uint32_t *cur_bm = g->d_front_bm[0], *next_bm = g->d_front_bm[1];

    CUDA_TRY(cudaEventRecord(ctx->ev_start, st));
    CUDA_TRY(cudaMemsetAsync(d_levels, 0xFF, (size_t)V * 4, st)); // UNVISITED_VERTEX = -1
    CUDA_TRY(cudaMemsetAsync(g->d_visited, 0, words * 4, st));
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, C_COUNT * 8, st));
    const int src_tier = source < b0 ? 0 : (source < b1 ? 1 : 2);
    bfs_init_kernel<<<1, 1, 0, st>>>(d_levels, g->d_visited, source, cq.q[src_tier]);
    KERNEL_TRY();
    ctx->launches++;

    int32_t n[3] = {0, 0, 0};
    n[src_tier] = 1;
    long long n_cur = 1, visited_total = 1;
    bool bottom_up = false;
    int32_t level = VGLB_FIRST_LEVEL_VERTEX;
    int64_t tot_edges = 0, tot_rows = 0, tot_frontier_bytes = 0, levels_run = 0;
    int32_t bu_levels = 0;
    const long long factor = (g->E / (V > 0 ? V : 1)) / 2 > 0 ? (g->E / V) / 2 : 1; // change_state.hpp:104
    const int max_blocks = ctx->sm_count * 16;

    while (n_cur > 0)
    {
        if (!bottom_up)
        {
            const int blocks_mid = (int)min((int64_t)max_blocks, ceil_div64(n[1], BFS_THREADS / 32));
            const int blocks_small = (int)min((int64_t)max_blocks, ceil_div64(n[2], BFS_THREADS / BFS_SMALL_LANES));
            const int64_t grid = (int64_t)n[0] + blocks_mid + blocks_small;
            bfs_td_kernel<false><<<(unsigned)grid, BFS_THREADS, 0, st>>>(g->d_out_ptr, g->d_out_adj, cq, n[0], n[1], n[2], blocks_mid,
                                                                 blocks_small, g->d_visited, d_levels, level + 1, b0, b1, nq,
                                                                 d_cnt);
            KERNEL_TRY();
            ctx->launches++;
            tot_rows += n_cur;
        }
        else
        {
            bfs_bu_kernel<<<ctx->sm_count * 8, BFS_THREADS, 0, st>>>(g->d_in_ptr, g->d_in_adj, V, g->d_visited, cur_bm,
                                                                    next_bm, d_levels, level + 1, d_cnt);
            KERNEL_TRY();
            ctx->launches++;
            bu_levels++;
            tot_frontier_bytes += (int64_t)words * 4 * 3; // visited read, frontier read, next written
        }
        CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, C_COUNT * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        levels_run++;
        int32_t nn[3] = {(int32_t)h_cnt[C_NEXT_BIG], (int32_t)h_cnt[C_NEXT_MID], (int32_t)h_cnt[C_NEXT_SMALL]};
        const long long n_next = bottom_up ? (long long)h_cnt[C_FOUND] : (long long)nn[0] + nn[1] + nn[2];
        const long long in_lvl = (long long)h_cnt[C_EDGES];
        tot_edges += in_lvl;
        if (bottom_up) tot_rows += (long long)h_cnt[C_ROWS];
        else tot_frontier_bytes += 8 * n_next; // queue written now, read next level
        visited_total += n_next;
        if (n_next == 0) break;

        // direction switch — gpu_change_state (change_state.hpp:100-141) evaluated on the device counters; the
        // top-down test uses the edges the NEXT top-down level would inspect (m_f) instead of the level just done
        bool next_bu = bottom_up;
        if (dopt)
        {
            const long long unvisited = (long long)V - visited_total;
            if (!bottom_up && n_cur < n_next)
            {
                CUDA_TRY(cudaMemsetAsync(d_cnt + C_MF, 0, 8, st));
                bfs_queue_degree_kernel<<<(unsigned)min((int64_t)max_blocks, ceil_div64(n_next, 256)), 256, 0, st>>>(
                    g->d_out_ptr, nq, nn[0], nn[1], nn[2], d_cnt);
                KERNEL_TRY();
                ctx->launches++;
                CUDA_TRY(cudaMemcpyAsync(h_cnt + C_MF, d_cnt + C_MF, 8, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                const long long m_f = (long long)h_cnt[C_MF];
                if (m_f >= (unvisited * factor + V) / alpha) next_bu = true;
            }
            else if (bottom_up && n_cur >= n_next)
            {
                if (n_next < (unvisited * factor + V) / (factor * beta)) next_bu = false;
            }
        }
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, C_COUNT * 8, st));
        if (!bottom_up && !next_bu)
        {
            TierQueues t = cq; cq = nq; nq = t;
            n[0] = nn[0]; n[1] = nn[1]; n[2] = nn[2];
        }
        else if (!bottom_up && next_bu)
        {
            CUDA_TRY(cudaMemsetAsync(cur_bm, 0, words * 4, st));
            bfs_queue_to_bitmap_kernel<<<(unsigned)min((int64_t)max_blocks, ceil_div64(n_next, 256)), 256, 0, st>>>(
                nq, nn[0], nn[1], nn[2], cur_bm);
            KERNEL_TRY();
            ctx->launches++;
            tot_frontier_bytes += (int64_t)words * 4 + 4 * n_next;
        }
        else if (bottom_up && next_bu)
        {
            uint32_t *t = cur_bm; cur_bm = next_bm; next_bm = t;
        }
        else
        {
            bfs_bitmap_to_queue_kernel<<<(unsigned)min((int64_t)max_blocks, ceil_div64((int64_t)words, 256)), 256, 0, st>>>(
                next_bm, V, b0, b1, cq, d_cnt);
            KERNEL_TRY();
            ctx->launches++;
            CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, 3 * 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            n[0] = (int32_t)h_cnt[0]; n[1] = (int32_t)h_cnt[1]; n[2] = (int32_t)h_cnt[2];
            CUDA_TRY(cudaMemsetAsync(d_cnt, 0, C_COUNT * 8, st));
            tot_frontier_bytes += (int64_t)words * 4 + 4 * n_next;
        }
        bottom_up = next_bu;
        n_cur = n_next;
        level++;
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_stop, st));
    CUDA_TRY(cudaEventSynchronize(ctx->ev_stop));
    if (stats)
    {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
        memset(stats, 0, sizeof(*stats));
        stats->seconds = ms * 1e-3;
        stats->iterations = levels_run;
        stats->edges_inspected = tot_edges;
        stats->vertices_processed = tot_rows;
        stats->frontier_bytes = tot_frontier_bytes;
        // B = sum_levels [ 8 e_l + 12 f_l + g_l ]
        stats->algorithmic_bytes = 8 * tot_edges + 12 * tot_rows + tot_frontier_bytes;
        stats->kernel_launches = ctx->launches - launches0;
        stats->bottom_up_levels = bu_levels;
    }
    return VGLB_OK;
}
