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

template <int NT>
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
                if (cand < old)
                {
                    const uint32_t bit = 1u << (v & 31);
                    const uint32_t m = atomicOr(&mark[v >> 5], bit);
                    won = !(m & bit);
                }
            }
        }
        enqueue_binned(won, v, b0, b1, nq, counters);
    }
}

__global__ void __launch_bounds__(SSSP_THREADS)
sssp_relax_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ adj, const float *__restrict__ wgt,
                  TierQueues cq, int32_t n_big, int32_t n_mid, int32_t n_small, int32_t blocks_mid, int32_t blocks_small,
                  uint32_t *__restrict__ dist, uint32_t *__restrict__ mark, int32_t b0, int32_t b1, TierQueues nq,
                  unsigned long long *counters)
{
    const uint64_t pol = l2_policy_evict_first();
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long edges = 0;
    if (b < n_big)
    {
        const int32_t u = cq.q[0][b];
        const int64_t s = ptr[u], e = ptr[u + 1];
        const float du = __uint_as_float(dist[u]);
        if (threadIdx.x == 0) edges = e - s;
        sssp_expand<SSSP_THREADS>(adj, wgt, s, e, threadIdx.x, du, dist, mark, b0, b1, nq, counters, pol);
    }
    else if (b < n_big + blocks_mid)
    {
        const int nwarps = blocks_mid * (SSSP_THREADS / 32);
        for (int i = (b - n_big) * (SSSP_THREADS / 32) + warp; i < n_mid; i += nwarps)
        {
            const int32_t u = cq.q[1][i];
            const int64_t s = ptr[u], e = ptr[u + 1];
            const float du = __uint_as_float(dist[u]);
            if (lane == 0) edges += e - s;
            sssp_expand<32>(adj, wgt, s, e, lane, du, dist, mark, b0, b1, nq, counters, pol);
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
                du = __uint_as_float(dist[u]);
                if (gl == 0) edges += e - s;
            }
            sssp_expand<G>(adj, wgt, s, e, gl, du, dist, mark, b0, b1, nq, counters, pol);
        }
    }
    edges = warp_sum_i64(edges);
    if (lane == 0 && edges) atomicAdd(&counters[C_EDGES], (unsigned long long)edges);
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
        sssp_relax_kernel<<<(unsigned)grid, SSSP_THREADS, 0, st>>>(g->d_out_ptr, g->d_out_adj, d_weights, cq, n[0], n[1],
                                                                  n[2], blocks_mid, blocks_small, dist, mark, b0, b1, nq,
                                                                  d_cnt);
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
