// pagerank.cu — PageRank "pull" sweep as a vectorised segmented reduction over the degree-sorted CSR (sm_100a).
//
// Reference semantics = the MULTICORE build (algorithms/pr/pr.hpp:7-148):
//     r'[u] = k + d * ( sum_{(u->v) in E, v != u} r[v] * inv[v]  +  D ),   inv[v] = 1/indeg_noloops(v) (0 if none)
//     D = sum_{v : indeg_noloops(v) == 0} r[v] / V,   r0 = 1/V,   d = 0.85f,   k = (1-d)/V,   exactly `iters` sweeps.
// The reference does this with four operator calls per sweep (compute save_old_ranks :85-90, reduce dangling :94-103,
// scatter edge_op+post :105-124, reduce ranks_sum :130-135) = 4 V-passes + one E-pass with TWO gathers per edge.
//
// B200 design: ONE kernel per sweep, no atomics on the rank vector, no host sync inside the loop.
//   * contrib[v] = r[v]*inv[v] is produced by the epilogue of the previous sweep, so the E-pass gathers one fp32 per
//     edge (the product is the same fp32 multiply the reference does per edge, pr.hpp:112-115);
//   * rows are degree-sorted, so the load-balancing tiers are contiguous id ranges decided by blockIdx alone
//     (replaces the ve / vc / collective tiers of multicore/advance_all_active.hpp:7-229):
//         degree >= 4096 : one CTA per row, int4 column-index loads, block reduction
//         32..4095       : one warp per row, int4 column-index loads, shuffle reduction
//         16..31, 8..15, 4..7, 2..3, <=1 : 16 / 8 / 4 / 2 / 1 lanes per row (no divergence: neighbours in id have
//                          neighbouring degrees), rows of one warp are adjacent in the adjacency array => coalesced
//   * column indices are streamed (ld.global.nc.L1::no_allocate.L2::evict_first), the gathered contribution vector
//     is kept L2-resident (ld.global.nc.L2::evict_last): 64 MB at scale 24 fits the 126 MB L2;
//   * the epilogue fuses post-op (:118-121), next sweep's contribution, and the NEXT sweep's dangling mass (summed in
//     fp64: block reduction + one atomicAdd(double) per CTA), so `reduce` never returns to the host.
// Row sums are fp32 trees instead of the reference's sequential fp32 (difference ~1e-7 relative, tolerance 1e-6).
// HBM roofline: algorithmic bytes per sweep = 8E (index + gathered value per edge) + 16V (row pointer, inv read,
// contribution write) [+4V rank write on the last sweep].
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

#define PR_THREADS 256
#define PR_WARP_ROWS_PER_WARP 8   // rows handled by one warp of the warp tier
#define PR_GROUP_PASSES 16        // passes of a CTA over its rows in the sub-warp tiers

struct PrParams
{
    const int64_t *ptr;
    const int32_t *adj;
    const float *contrib_in;
    const float *inv;
    float *contrib_out;
    float *rank_out; // written on the final sweep only (may be NULL otherwise)
    const double *dangling_in;
    double *dangling_out;
    int32_t V;
    float k, d, v_as_float;
    int32_t tier_border[VGLB_NUM_TIERS]; // first row NOT in tier t
    int32_t block_start[VGLB_NUM_TIERS]; // first block of tier t (tier 7 shares tier 6's kernel path)
};

struct L2Pol
{
    uint64_t stream, keep;
};

__device__ __forceinline__ void pr_epilogue(const PrParams &P, int32_t row, float sum, float dang, double &dang_local)
{
    const float inv_r = P.inv[row];
    // k + d * (rank + dangling) — pr.hpp:118-121, no FMA contraction (the x86-64 reference build has none)
    const float rank = __fadd_rn(P.k, __fmul_rn(P.d, __fadd_rn(sum, dang)));
    P.contrib_out[row] = __fmul_rn(rank, inv_r);
    if (P.rank_out) P.rank_out[row] = rank;
    if (inv_r == 0.0f) dang_local += (double)__fdiv_rn(rank, P.v_as_float); // pr.hpp:94-101
}

// sum over one row with `nthreads` cooperating threads (tid in [0,nthreads)), int4 body + scalar head/tail
template <int NTHREADS>
__device__ __forceinline__ float pr_row_partial(const PrParams &P, const L2Pol &pol, int32_t row, int64_t s, int64_t e, int tid)
{
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const int64_t s4 = (s + 3) & ~(int64_t)3;
    const int64_t e4 = e & ~(int64_t)3;
    if (s4 >= e4)
    {
        for (int64_t p = s + tid; p < e; p += NTHREADS)
        {
            const int32_t v = ld_stream_s32(P.adj + p, pol.stream);
            if (v != row) acc0 += ld_gather_f32(P.contrib_in + v, pol.keep);
        }
        return acc0;
    }
    // head (< 4 elements) and tail (< 4 elements)
    if (s + tid < s4)
    {
        const int32_t v = ld_stream_s32(P.adj + s + tid, pol.stream);
        if (v != row) acc1 += ld_gather_f32(P.contrib_in + v, pol.keep);
    }
    if (e4 + tid < e)
    {
        const int32_t v = ld_stream_s32(P.adj + e4 + tid, pol.stream);
        if (v != row) acc2 += ld_gather_f32(P.contrib_in + v, pol.keep);
    }
    const int4 *adj4 = reinterpret_cast<const int4 *>(P.adj);
    const int64_t q_end = e4 >> 2;
    int64_t q = (s4 >> 2) + tid;
    // two vectors (8 gathers) in flight per thread
    for (; q + NTHREADS < q_end; q += 2 * NTHREADS)
    {
        const int4 a = ld_stream_v4(adj4 + q, pol.stream);
        const int4 b = ld_stream_v4(adj4 + q + NTHREADS, pol.stream);
        const float a0 = a.x != row ? ld_gather_f32(P.contrib_in + a.x, pol.keep) : 0.f;
        const float a1 = a.y != row ? ld_gather_f32(P.contrib_in + a.y, pol.keep) : 0.f;
        const float a2 = a.z != row ? ld_gather_f32(P.contrib_in + a.z, pol.keep) : 0.f;
        const float a3 = a.w != row ? ld_gather_f32(P.contrib_in + a.w, pol.keep) : 0.f;
        const float b0 = b.x != row ? ld_gather_f32(P.contrib_in + b.x, pol.keep) : 0.f;
        const float b1 = b.y != row ? ld_gather_f32(P.contrib_in + b.y, pol.keep) : 0.f;
        const float b2 = b.z != row ? ld_gather_f32(P.contrib_in + b.z, pol.keep) : 0.f;
        const float b3 = b.w != row ? ld_gather_f32(P.contrib_in + b.w, pol.keep) : 0.f;
        acc0 += a0 + b0;
        acc1 += a1 + b1;
        acc2 += a2 + b2;
        acc3 += a3 + b3;
    }
    if (q < q_end)
    {
        const int4 a = ld_stream_v4(adj4 + q, pol.stream);
        if (a.x != row) acc0 += ld_gather_f32(P.contrib_in + a.x, pol.keep);
        if (a.y != row) acc1 += ld_gather_f32(P.contrib_in + a.y, pol.keep);
        if (a.z != row) acc2 += ld_gather_f32(P.contrib_in + a.z, pol.keep);
        if (a.w != row) acc3 += ld_gather_f32(P.contrib_in + a.w, pol.keep);
    }
    return (acc0 + acc1) + (acc2 + acc3);
}

// G lanes per row, rows [row0, row1) of this CTA
template <int G>
__device__ __forceinline__ void pr_group_tier(const PrParams &P, const L2Pol &pol, int32_t row0, int32_t row1, float dang, double &dang_local)
{
    constexpr int GROUPS = PR_THREADS / G;
    const int gid = threadIdx.x / G, gl = threadIdx.x % G;
    for (int32_t base = row0; base < row1; base += GROUPS)
    {
        const int32_t row = base + gid;
        float acc = 0.f;
        if (row < row1)
        {
            const int64_t s = P.ptr[row], e = P.ptr[row + 1];
            // degree is in [G, 2G): at most two column indices per lane, both loads issued before the gathers
            const int64_t p0 = s + gl, p1 = p0 + G;
            int32_t v0 = row, v1 = row;
            if (p0 < e) v0 = ld_stream_s32(P.adj + p0, pol.stream);
            if (p1 < e) v1 = ld_stream_s32(P.adj + p1, pol.stream);
            if (v0 != row) acc += ld_gather_f32(P.contrib_in + v0, pol.keep);
            if (v1 != row) acc += ld_gather_f32(P.contrib_in + v1, pol.keep);
            for (int64_t p = p1 + G; p < e; p += G) // only when a caller passes rows with degree >= 2G
            {
                const int32_t v = ld_stream_s32(P.adj + p, pol.stream);
                if (v != row) acc += ld_gather_f32(P.contrib_in + v, pol.keep);
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (gl == 0 && row < row1) pr_epilogue(P, row, acc, dang, dang_local);
    }
}

__global__ void __launch_bounds__(PR_THREADS) pr_sweep_kernel(const __grid_constant__ PrParams P)
{
    L2Pol pol;
    pol.stream = l2_policy_evict_first();
    pol.keep = l2_policy_evict_last();
    __shared__ float s_part[PR_THREADS / 32];
    __shared__ double s_dang[PR_THREADS / 32];
    const int b = blockIdx.x;
    const float dang = (float)(*P.dangling_in);
    double dang_local = 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (b < P.block_start[1])
    {
        // tier 0: one CTA per row
        const int32_t row = b;
        const int64_t s = P.ptr[row], e = P.ptr[row + 1];
        float acc = pr_row_partial<PR_THREADS>(P, pol, row, s, e, threadIdx.x);
        acc = warp_sum_f32(acc);
        if (lane == 0) s_part[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0)
        {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < PR_THREADS / 32; w++) t += s_part[w];
            pr_epilogue(P, row, t, dang, dang_local);
        }
    }
    else if (b < P.block_start[2])
    {
        // tier 1: one warp per row, PR_WARP_ROWS_PER_WARP rows per warp
        constexpr int ROWS = (PR_THREADS / 32) * PR_WARP_ROWS_PER_WARP;
        const int32_t row0 = P.tier_border[0] + (b - P.block_start[1]) * ROWS;
        const int32_t row1 = min(row0 + ROWS, P.tier_border[1]);
        for (int32_t row = row0 + warp; row < row1; row += PR_THREADS / 32)
        {
            const int64_t s = P.ptr[row], e = P.ptr[row + 1];
            float acc = pr_row_partial<32>(P, pol, row, s, e, lane);
            acc = warp_sum_f32(acc);
            if (lane == 0) pr_epilogue(P, row, acc, dang, dang_local);
        }
    }
    else
    {
        int t = 2;
#pragma unroll
        for (int i = 3; i < VGLB_NUM_TIERS - 1; i++)
            if (b >= P.block_start[i]) t = i;
        const int32_t tier_first = P.tier_border[t - 1];
        const int32_t tier_last = (t == VGLB_NUM_TIERS - 2) ? P.V : P.tier_border[t];
        const int G = 32 >> (t - 1); // t=2:16, 3:8, 4:4, 5:2, 6:1
        const int32_t rows_per_cta = (PR_THREADS / G) * PR_GROUP_PASSES;
        const int32_t row0 = tier_first + (b - P.block_start[t]) * rows_per_cta;
        const int32_t row1 = min(row0 + rows_per_cta, tier_last);
        switch (t)
        {
        case 2: pr_group_tier<16>(P, pol, row0, row1, dang, dang_local); break;
        case 3: pr_group_tier<8>(P, pol, row0, row1, dang, dang_local); break;
        case 4: pr_group_tier<4>(P, pol, row0, row1, dang, dang_local); break;
        case 5: pr_group_tier<2>(P, pol, row0, row1, dang, dang_local); break;
        default: pr_group_tier<1>(P, pol, row0, row1, dang, dang_local); break;
        }
    }
    // next sweep's dangling mass: fp64 block reduction, one atomic per CTA that has any
    dang_local = warp_sum_f64(dang_local);
    if (lane == 0) s_dang[warp] = dang_local;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < PR_THREADS / 32; w++) t += s_dang[w];
        if (t != 0.0) atomicAdd(P.dangling_out, t);
    }
}

// inv[v] = (float)(1.0 / indeg_noloops[v]) or 0 — pr.hpp:66-73 (double division then narrowing, like the reference)
__global__ void pr_inverse_degree_kernel(const int32_t *__restrict__ indeg, int32_t V, float *__restrict__ inv)
{
    int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V)
    {
        const int32_t d = indeg[v];
        inv[v] = d == 0 ? 0.0f : (float)(1.0 / (double)d);
    }
}

// r0 = 1/V (pr.hpp:40-45): contrib0 = r0*inv, dangling[0] = sum over inv==0 of r0/V
__global__ void pr_init_kernel(const float *__restrict__ inv, int32_t V, float r0, float v_as_float,
                               float *__restrict__ contrib, double *__restrict__ dangling0)
{
    __shared__ double s_dang[PR_THREADS / 32];
    double local = 0.0;
    for (int32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x)
    {
        const float i = inv[v];
        contrib[v] = __fmul_rn(r0, i);
        if (i == 0.0f) local += (double)__fdiv_rn(r0, v_as_float);
    }
    local = warp_sum_f64(local);
    if ((threadIdx.x & 31) == 0) s_dang[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double t = 0.0;
        for (int w = 0; w < PR_THREADS / 32; w++) t += s_dang[w];
        if (t != 0.0) atomicAdd(dangling0, t);
    }
}

__global__ void pr_fill_kernel(float *a, int32_t n, float val)
{
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = val;
}

static int pr_prepare(vglb_ctx *ctx, vglb_graph *g, int iters)
{
    if (!g->d_pr_inv)
    {
        int32_t *d_indeg = NULL;
        CUDA_TRY(cudaMalloc(&d_indeg, (size_t)g->V * 4));
        int rc = vglb_graph_indegree_noloops(ctx, g, d_indeg);
        if (rc != VGLB_OK) { cudaFree(d_indeg); return rc; }
        CUDA_TRY(cudaMalloc(&g->d_pr_inv, (size_t)g->V * 4));
        CUDA_TRY(cudaMalloc(&g->d_pr_contrib[0], (size_t)g->V * 4));
        CUDA_TRY(cudaMalloc(&g->d_pr_contrib[1], (size_t)g->V * 4));
        pr_inverse_degree_kernel<<<(unsigned)ceil_div64(g->V, 256), 256, 0, ctx->stream>>>(d_indeg, g->V, g->d_pr_inv);
        KERNEL_TRY();
        ctx->launches++;
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_indeg);
    }
    if (g->pr_dangling_slots < iters + 1)
    {
        cudaFree(g->d_pr_dangling);
        g->d_pr_dangling = NULL;
        CUDA_TRY(cudaMalloc(&g->d_pr_dangling, (size_t)(iters + 1) * sizeof(double)));
        g->pr_dangling_slots = iters + 1;
    }
    return VGLB_OK;
}

extern "C" int vglb_pagerank(vglb_ctx *ctx, vglb_graph *g, int iters, float damping, float *d_ranks, vglb_stats *stats)
{
    VGLB_REQUIRE(ctx != NULL && g != NULL && d_ranks != NULL, "vglb_pagerank: NULL argument");
    VGLB_REQUIRE(iters >= 0 && iters < (1 << 20), "vglb_pagerank: bad iteration count");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    int rc = pr_prepare(ctx, g, iters);
    if (rc != VGLB_OK) return rc;

    const int32_t V = g->V;
    PrParams P;
    P.ptr = g->d_out_ptr;
    P.adj = g->d_out_adj;
    P.inv = g->d_pr_inv;
    P.V = V;
    P.d = damping;
    P.k = (float)((1.0 - (double)damping) / (double)((float)V)); // pr.hpp:37-38
    P.v_as_float = (float)V;
    // block ranges per tier
    int64_t nblocks = 0;
    for (int t = 0; t < VGLB_NUM_TIERS; t++) P.tier_border[t] = g->tier_border[t];
    for (int t = 0; t < VGLB_NUM_TIERS - 1; t++)
    {
        P.block_start[t] = (int32_t)nblocks;
        const int32_t first = t == 0 ? 0 : g->tier_border[t - 1];
        const int32_t last = (t == VGLB_NUM_TIERS - 2) ? V : g->tier_border[t]; // tier 6 also takes degree-0 rows
        const int64_t rows = last - first;
        int64_t rows_per_cta;
        if (t == 0) rows_per_cta = 1;
        else if (t == 1) rows_per_cta = (PR_THREADS / 32) * PR_WARP_ROWS_PER_WARP;
        else rows_per_cta = (int64_t)(PR_THREADS / (32 >> (t - 1))) * PR_GROUP_PASSES;
        nblocks += ceil_div64(rows, rows_per_cta);
    }
    P.block_start[VGLB_NUM_TIERS - 1] = (int32_t)nblocks;
    VGLB_REQUIRE(nblocks < 0x7fffffffLL, "vglb_pagerank: grid too large");

    CUDA_TRY(cudaEventRecord(ctx->ev_start, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(g->d_pr_dangling, 0, (size_t)(iters + 1) * sizeof(double), ctx->stream));
    const float r0 = (float)(1.0 / (double)V); // pr.hpp:42
    if (iters == 0)
    {
        pr_fill_kernel<<<(unsigned)ceil_div64(V, 256), 256, 0, ctx->stream>>>(d_ranks, V, r0);
        KERNEL_TRY();
        ctx->launches++;
    }
    else
    {
        pr_init_kernel<<<ctx->sm_count * 4, PR_THREADS, 0, ctx->stream>>>(g->d_pr_inv, V, r0, P.v_as_float,
                                                                        g->d_pr_contrib[0], g->d_pr_dangling);
        KERNEL_TRY();
        ctx->launches++;
    }
    for (int it = 0; it < iters; it++)
    {
        P.contrib_in = g->d_pr_contrib[it & 1];
        P.contrib_out = g->d_pr_contrib[(it + 1) & 1];
        P.rank_out = (it == iters - 1) ? d_ranks : NULL;
        P.dangling_in = g->d_pr_dangling + it;
        P.dangling_out = g->d_pr_dangling + it + 1;
        pr_sweep_kernel<<<(unsigned)nblocks, PR_THREADS, 0, ctx->stream>>>(P);
        KERNEL_TRY();
        ctx->launches++;
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_stop, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(ctx->ev_stop));
    if (stats)
    {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
        memset(stats, 0, sizeof(*stats));
        stats->seconds = ms * 1e-3;
        stats->iterations = iters;
        stats->edges_inspected = (int64_t)iters * g->E;
        stats->vertices_processed = (int64_t)iters * V;
        stats->algorithmic_bytes = (int64_t)iters * (8 * g->E + 16 * (int64_t)V) + (iters > 0 ? 4 * (int64_t)V : 0);
        stats->kernel_launches = ctx->launches - launches0;
    }
    return VGLB_OK;
}
