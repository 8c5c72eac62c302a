// sssp.cu — SSSP as frontier Bellman-Ford over the device VectCSR (sm_100a).
//
// Reference: ShortestPaths::vgl_dijkstra_partial_active (algorithms/sssp/shortest_paths.hpp:7-78): per round
//   compute(prev = dist) over all V, scatter(work_frontier, if dist[dst] > dist[src]+w then dist[dst] = dist[src]+w),
//   generate_new_frontier(dist != prev) over all V. The multicore relax is a racy non-atomic read-modify-write
//   (SURVEY §0 hazard 1); the GPU reference is racy too (sssp/gpu_shortest_paths.hpp:156-167). The parity target is
//   the unique min-plus fixed point in fp32 = ShortestPaths::seq_dijkstra (seq_shortest_paths.hpp:8-68), bit-exact:
//   every candidate is one fp32 add `dist[src] + w`, fp32 add is monotone, so the minimum over paths does not depend
//   on the relaxation order. Unreachable = FLT_MAX - 100 = FLT_MAX in fp32 (shortest_paths.hpp:22).
//
// B200 design
//   * distances of non-negative floats order like their bit patterns: relax = atomicMin on the uint32 view, after a
//     plain (L2-resident) read that filters the candidates that cannot win;
//   * the next frontier is emitted inside the advance: the thread that lowers dist[v] marks v in a per-round bitmap
//     (atomicOr); the first marker enqueues v into the queue of its degree tier. No `prev` copy, no V-pass GNF —
//     the reference's two full-V passes per round (shortest_paths.hpp:40-44,58-66) disappear;
//   * same degree-binned queues / CTA-warp-8-lane tiers as bfs.cu; column indices and weights are streamed together
//     (ld.global.nc.L1::no_allocate + L2 evict-first), the distance vector is the L2-resident gather target.
// HBM roofline: algorithmic bytes = sum_rounds [ 12 e_i (index + weight + dist[dst]) + 24 f_i + 8 n(F_{i+1}) ].
#include <float.h>
#include <stdlib.h>

#include "common.cuh"
#include "frontier.cuh"

#define SSSP_THREADS 256
#define SSSP_SMALL_LANES 8

// PART = one rank's part of a partitioned graph: `dist` is the replicated vector (column ids), winners are not queued
// here — owners find their changed vertices after the allreduce(min).
template <int NT, bool PART>
__device__ __forceinline__ void sssp_expand(const int32_t *__restrict__ adj, const float *__restrict__ wgt, int64_t s,
                                            int64_t e, int tid, float du, uint32_t *__restrict__ dist,
                                            uint32_t *__restrict__ mark, int32_t b0, int32_t b1, const TierQueues &nq,
                                            unsigned long long *counters, uint64_t pol_stream)
{
    for (int64_t p = s + tid;; p += NT)
    {
        const bool active = p < e;
        if (!__any_sync(0xffffffffu, active)) break;
        bool won = false;
        int32_t v = 0;
        if (active)
        {
            v = ld_stream_s32(adj + p, pol_stream);
            const float w = ld_stream_f32(wgt + p, pol_stream);
            const uint32_t cand = __float_as_uint(__fadd_rn(du, w)); // shortest_paths.hpp:50-53
            if (cand < dist[v])
            {
                const uint32_t old = atomicMin(&dist[v], cand);
                if (!PART && cand < old)
                {
                    const uint32_t bit = 1u << (v & 31);
                    const uint32_t m = atomicOr(&mark[v >> 5], bit);
                    won = !(m & bit);
                }
            }
        }
        if (!PART) enqueue_binned(won, v, b0, b1, nq, counters);
    }
}

template <bool PART>
__global__ void __launch_bounds__(SSSP_THREADS)
sssp_relax_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ adj, const float *__restrict__ wgt,
                  TierQueues cq, int32_t n_big, int32_t n_mid, int32_t n_small, int32_t blocks_mid, int32_t blocks_small,
                  uint32_t *__restrict__ dist, uint32_t *__restrict__ mark, int32_t b0, int32_t b1, TierQueues nq,
                  unsigned long long *counters, int32_t col0)
{
    const uint64_t pol = l2_policy_evict_first();
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long edges = 0;
    if (b < n_big)
    {
        const int32_t u = cq.q[0][b];
        const int64_t s = ptr[u], e = ptr[u + 1];
        const float du = __uint_as_float(dist[col0 + u]);
        if (threadIdx.x == 0) edges = e - s;
        sssp_expand<SSSP_THREADS, PART>(adj, wgt, s, e, threadIdx.x, du, dist, mark, b0, b1, nq, counters, pol);
    }
    else if (b < n_big + blocks_mid)
    {
        const int nwarps = blocks_mid * (SSSP_THREADS / 32);
        for (int i = (b - n_big) * (SSSP_THREADS / 32) + warp; i < n_mid; i += nwarps)
        {
            const int32_t u = cq.q[1][i];
            const int64_t s = ptr[u], e = ptr[u + 1];
            const float du = __uint_as_float(dist[col0 + u]);
            if (lane == 0) edges += e - s;
            sssp_expand<32, PART>(adj, wgt, s, e, lane, du, dist, mark, b0, b1, nq, counters, pol);
        }
    }
    else
    {
        constexpr int G = SSSP_SMALL_LANES;
        constexpr int GROUPS = SSSP_THREADS / G;
        const int ngroups = blocks_small * GROUPS;
        const int gid = threadIdx.x / G, gl = threadIdx.x % G;
        for (int base = (b - n_big - blocks_mid) * GROUPS; base < n_small; base += ngroups)
        {
            const int i = base + gid;
            int64_t s = 0, e = 0;
            float du = 0.f;
            if (i < n_small)
            {
                const int32_t u = cq.q[2][i];
                s = ptr[u];
                e = ptr[u + 1];
                du = __uint_as_float(dist[col0 + u]);
                if (gl == 0) edges += e - s;
            }
            sssp_expand<G, PART>(adj, wgt, s, e, gl, du, dist, mark, b0, b1, nq, counters, pol);
        }
    }
    edges = warp_sum_i64(edges);
    if (lane == 0 && edges) atomicAdd(&counters[C_EDGES], (unsigned long long)edges);
}


// ---- 1D-partitioned SSSP ---------------------------------------------------------------------------------------------
// Each rank relaxes the out-edges of the frontier vertices it owns into its replica of the distance vector (atomicMin),
// the replicas are combined with one allreduce(min) on the uint32 view per round — the reference MPI_Allreduce(MIN)es
// the same array after its advance (mpi_exchange.hpp:155-271, gpu_shortest_paths.hpp:133-196) — and every owner queues
// the vertices of its slice whose distance dropped (generate_new_frontier dist != prev, shortest_paths.hpp:58-66).

__global__ void sssp_part_init_kernel(uint32_t *__restrict__ dist, int64_t cols, int32_t source_col,
                                      uint32_t *__restrict__ prev, int32_t vp, int32_t src_row, int32_t *queue_slot)
{
    const uint32_t inf_bits = __float_as_uint(FLT_MAX - 100.0f);
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < cols; v += (int64_t)gridDim.x * blockDim.x)
    {
        dist[v] = v == source_col ? 0u : inf_bits;
        if (v < vp) prev[v] = (int32_t)v == src_row ? 0u : inf_bits;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && src_row >= 0) *queue_slot = src_row;
}

__global__ void __launch_bounds__(256)
sssp_part_frontier_kernel(const uint32_t *__restrict__ dist_slice, uint32_t *__restrict__ prev, int32_t rows, int32_t b0,
                          int32_t b1, TierQueues nq, unsigned long long *counters)
{
    const int32_t rpad = (rows + 31) & ~31;
    for (int32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < rpad; r += gridDim.x * blockDim.x)
    {
        bool changed = false;
        if (r < rows)
        {
            const uint32_t d = dist_slice[r];
            if (d < prev[r])
            {
                prev[r] = d;
                changed = true;
            }
        }
        enqueue_binned(changed, r, b0, b1, nq, counters);
    }
}

__global__ void sssp_copy_counters_kernel(unsigned long long *c)
{
    if (threadIdx.x < C_COUNT) c[C_COUNT + threadIdx.x] = c[threadIdx.x];
}

static int sssp_partitioned(vglb_ctx *ctx, vglb_graph *g, const float *d_weights, int32_t source, float *d_dist,
                            vglb_stats *stats)
{
    VGLB_REQUIRE(source >= 0 && source < g->cols, "vglb_sssp: source column out of range");
    CUDA_TRY(cudaSetDevice(ctx->device));
    vglb_comm *comm = g->comm;
    const int32_t rank = g->part_rank, vp = g->vp, rows = g->V, col0 = g->col_of_row0;
    if (!g->d_part_vec) CUDA_TRY(cudaMalloc(&g->d_part_vec, (size_t)g->cols * 4));
    if (!g->d_part_prev) CUDA_TRY(cudaMalloc(&g->d_part_prev, (size_t)vp * 4));
    if (!g->d_queue[0]) CUDA_TRY(cudaMalloc(&g->d_queue[0], ((size_t)vp + 3) * 4));
    if (!g->d_queue[1]) CUDA_TRY(cudaMalloc(&g->d_queue[1], ((size_t)vp + 3) * 4));
    const int64_t launches0 = ctx->launches;
    const int32_t b0 = g->tier_border[0], b1 = g->tier_border[1];
    unsigned long long *d_cnt = (unsigned long long *)ctx->d_counters;
    unsigned long long *h_cnt = (unsigned long long *)ctx->h_counters;
    cudaStream_t st = ctx->stream;
    uint32_t *dist = g->d_part_vec;
    auto regions = [&](int32_t *base) {
        TierQueues q;
        q.q[0] = base;
        q.q[1] = base + b0;
        q.q[2] = base + b1;
        return q;
    };
    TierQueues cq = regions(g->d_queue[0]), nq = regions(g->d_queue[1]);

    CUDA_TRY(cudaEventRecord(ctx->ev_start, st));
    const bool own_source = source / vp == rank;
    const int32_t src_row = own_source ? source - col0 : -1;
    VGLB_REQUIRE(!own_source || src_row < rows, "vglb_sssp: source is a padding column");
    const int src_tier = src_row < 0 ? 0 : (src_row < b0 ? 0 : (src_row < b1 ? 1 : 2));
    sssp_part_init_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(dist, g->cols, source, g->d_part_prev, vp, src_row, cq.q[src_tier]);
    KERNEL_TRY();
    ctx->launches++;
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 2 * C_COUNT * 8, st));
    int32_t n[3] = {0, 0, 0};
    if (own_source) n[src_tier] = 1;
    long long n_cur = 1;
    int64_t tot_edges = 0, tot_rows = 0, tot_next = 0, rounds = 0;
    const int max_blocks = ctx->sm_count * 16;
    int rc;
    while (n_cur > 0)
    {
        const int blocks_mid = (int)min((int64_t)max_blocks, ceil_div64(n[1], SSSP_THREADS / 32));
        const int blocks_small = (int)min((int64_t)max_blocks, ceil_div64(n[2], SSSP_THREADS / SSSP_SMALL_LANES));
        const int64_t grid = (int64_t)n[0] + blocks_mid + blocks_small;
        if (grid > 0)
        {
            sssp_relax_kernel<true><<<(unsigned)grid, SSSP_THREADS, 0, st>>>(g->d_out_ptr, g->d_out_adj, d_weights, cq, n[0], n[1], n[2],
                                                                            blocks_mid, blocks_small, dist, NULL, b0, b1, nq, d_cnt, col0);
            KERNEL_TRY();
            ctx->launches++;
        }
        rc = vglb_comm_allreduce_async(comm, dist, (size_t)g->cols, VGLB_DT_U32, VGLB_OP_MIN);
        if (rc != VGLB_OK) return rc;
        if (rows > 0)
        {
            sssp_part_frontier_kernel<<<(unsigned)min((int64_t)max_blocks, ceil_div64(rows, 256)), 256, 0, st>>>(
                dist + col0, g->d_part_prev, rows, b0, b1, nq, d_cnt);
            KERNEL_TRY();
        }
        sssp_copy_counters_kernel<<<1, 32, 0, st>>>(d_cnt);
        KERNEL_TRY();
        ctx->launches += 2;
        rc = vglb_comm_allreduce_async(comm, d_cnt + C_COUNT, C_COUNT, VGLB_DT_I64, VGLB_OP_SUM);
        if (rc != VGLB_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, 2 * C_COUNT * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 2 * C_COUNT * 8, st));
        rounds++;
        tot_rows += (int64_t)n[0] + n[1] + n[2];
        tot_edges += (int64_t)h_cnt[C_EDGES];
        n[0] = (int32_t)h_cnt[C_NEXT_BIG];
        n[1] = (int32_t)h_cnt[C_NEXT_MID];
        n[2] = (int32_t)h_cnt[C_NEXT_SMALL];
        const unsigned long long *gl = h_cnt + C_COUNT;
        n_cur = (long long)(gl[C_NEXT_BIG] + gl[C_NEXT_MID] + gl[C_NEXT_SMALL]);
        tot_next += (int64_t)n[0] + n[1] + n[2];
        TierQueues t = cq; cq = nq; nq = t;
    }
    if (rows > 0) CUDA_TRY(cudaMemcpyAsync(d_dist, dist + col0, (size_t)rows * 4, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaEventRecord(ctx->ev_stop, st));
    CUDA_TRY(cudaEventSynchronize(ctx->ev_stop));
    if (stats)
    {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
        memset(stats, 0, sizeof(*stats));
        stats->seconds = ms * 1e-3;
        stats->iterations = rounds;
        stats->edges_inspected = tot_edges; // this rank's
        stats->vertices_processed = tot_rows;
        stats->frontier_bytes = 8 * tot_next + rounds * g->cols * 4 * 2; // queues + the allreduced vector (read + written)
        stats->algorithmic_bytes = 12 * tot_edges + 24 * tot_rows + stats->frontier_bytes;
        stats->kernel_launches = ctx->launches - launches0;
    }
    return VGLB_OK;
}

__global__ void sssp_init_kernel(uint32_t *__restrict__ dist, int32_t V, int32_t source, int32_t *queue_slot)
{
    const uint32_t inf_bits = __float_as_uint(FLT_MAX - 100.0f); // shortest_paths.hpp:22
    for (int32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x)
        dist[v] = v == source ? 0u : inf_bits;
    if (blockIdx.x == 0 && threadIdx.x == 0) *queue_slot = source;
}

extern "C" int vglb_sssp(vglb_ctx *ctx, vglb_graph *g, const float *d_weights, int32_t source, float *d_dist,
                         vglb_stats *stats)
{
    VGLB_REQUIRE(ctx != NULL && g != NULL && d_dist != NULL, "vglb_sssp: NULL argument");
    VGLB_REQUIRE(d_weights != NULL || g->E == 0, "vglb_sssp: NULL weights");
    if (g->comm) return sssp_partitioned(ctx, g, d_weights, source, d_dist, stats);
    VGLB_REQUIRE(source >= 0 && source < g->V, "vglb_sssp: source out of range");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t words = ((size_t)g->V + 31) / 32;
    if (!g->d_queue[0])
    {
        CUDA_TRY(cudaMalloc(&g->d_queue[0], ((size_t)g->V + 3) * 4));
        CUDA_TRY(cudaMalloc(&g->d_queue[1], ((size_t)g->V + 3) * 4));
    }
    if (!g->d_visited) CUDA_TRY(cudaMalloc(&g->d_visited, (words + 32) * 4));
    const int64_t launches0 = ctx->launches;
    const int32_t V = g->V, b0 = g->tier_border[0], b1 = g->tier_border[1];
    unsigned long long *d_cnt = (unsigned long long *)ctx->d_counters;
    unsigned long long *h_cnt = (unsigned long long *)ctx->h_counters;
    cudaStream_t st = ctx->stream;
    uint32_t *dist = (uint32_t *)d_dist, *mark = g->d_visited;
    auto regions = [&](int32_t *base) {
        TierQueues q;
        q.q[0] = base;
        q.q[1] = base + b0;
        q.q[2] = base + b1;
        return q;
    };
    TierQueues cq = regions(g->d_queue[0]), nq = regions(g->d_queue[1]);

    CUDA_TRY(cudaEventRecord(ctx->ev_start, st));
    const int src_tier = source < b0 ? 0 : (source < b1 ? 1 : 2);
    sssp_init_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(dist, V, source, cq.q[src_tier]);
    KERNEL_TRY();
    ctx->launches++;
    CUDA_TRY(cudaMemsetAsync(d_cnt, 0, C_COUNT * 8, st));
    int32_t n[3] = {0, 0, 0};
    n[src_tier] = 1;
    long long n_cur = 1;
    int64_t tot_edges = 0, tot_rows = 0, tot_next = 0, rounds = 0;
    const int max_blocks = ctx->sm_count * 16;
    while (n_cur > 0)
    {
        CUDA_TRY(cudaMemsetAsync(mark, 0, words * 4, st));
        const int blocks_mid = (int)min((int64_t)max_blocks, ceil_div64(n[1], SSSP_THREADS / 32));
        const int blocks_small = (int)min((int64_t)max_blocks, ceil_div64(n[2], SSSP_THREADS / SSSP_SMALL_LANES));
        const int64_t grid = (int64_t)n[0] + blocks_mid + blocks_small;
        sssp_relax_kernel<false><<<(unsigned)grid, SSSP_THREADS, 0, st>>>(g->d_out_ptr, g->d_out_adj, d_weights, cq, n[0], n[1],
                                                                         n[2], blocks_mid, blocks_small, dist, mark, b0, b1, nq,
                                                                         d_cnt, 0);
        KERNEL_TRY();
        ctx->launches++;
        CUDA_TRY(cudaMemcpyAsync(h_cnt, d_cnt, C_COUNT * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, C_COUNT * 8, st));
        rounds++;
        tot_rows += n_cur;
        tot_edges += (int64_t)h_cnt[C_EDGES];
        n[0] = (int32_t)h_cnt[C_NEXT_BIG];
        n[1] = (int32_t)h_cnt[C_NEXT_MID];
        n[2] = (int32_t)h_cnt[C_NEXT_SMALL];
        n_cur = (long long)n[0] + n[1] + n[2];
        tot_next += n_cur;
        TierQueues t = cq; cq = nq; nq = t;
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_stop, st));
    CUDA_TRY(cudaEventSynchronize(ctx->ev_stop));
    if (stats)
    {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
        memset(stats, 0, sizeof(*stats));
        stats->seconds = ms * 1e-3;
        stats->iterations = rounds;
        stats->edges_inspected = tot_edges;
        stats->vertices_processed = tot_rows;
        stats->frontier_bytes = 8 * tot_next;
        stats->algorithmic_bytes = 12 * tot_edges + 24 * tot_rows + 8 * tot_next;
        stats->kernel_launches = ctx->launches - launches0;
    }
    return VGLB_OK;
}
