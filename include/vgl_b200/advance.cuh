// vgl_b200/advance.cuh — header-only, lambda-generic device side of the B200 backend (sm_100a).
//
// These templates are the "advance / compute / reduce / generate_new_frontier" operators of VGL's GraphAbstractions
// (vgl_compute_api/common/graph_abstractions.h:96-152) for ARBITRARY user functors: device lambdas cannot cross the
// C ABI of libvgl_b200, so the generic path is instantiated in the caller's translation unit (nvcc --extended-lambda)
// by include/vgl_b200/graph_abstractions_b200.cuh, while the data structures, compaction and the fused algorithms
// live behind the C ABI. libvgl_b200 itself instantiates them for CC (cc.cu).
//
// Functor contracts = the reference's (architecture_independent_api.h:17-30):
//   edge_op  (int src_id, int dst_id, int local_edge_pos, long long global_edge_pos, int vector_index)
//   vertex op(int src_id, int connections_count, int vector_index)       pre: before, post: after all edges of src
// vector_index is the lane id. Ops of one vertex may run on up to a whole CTA: edge ops must be atomic where the
// reference's GPU lambdas are (VGL_SRC_ID_ADD, architecture_independent_api.h:47-51).
//
// Load balancing replaces the reference's block / virtual-warp kernels and six streams
// (vgl_compute_api/gpu/advance_csr.hpp:78-165,222-305): ids are degree-sorted, so a tier is a contiguous id range
// (all-active) or one of three degree-binned queues (sparse), and one launch covers all tiers.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

namespace vglb
{

constexpr int kAdvThreads = 256;
constexpr int kNumTiers = 8;
constexpr int kWarpRowsPerWarp = 8;
constexpr int kGroupPasses = 16;

struct CsrView
{
    const int64_t *ptr;
    const int32_t *adj;
    int32_t V;
    int32_t tier_border[kNumTiers]; // first id whose degree is below {4096,32,16,8,4,2,1,0}
};

struct AllActivePlan
{
    int32_t block_start[kNumTiers];
    int64_t blocks;
};

inline AllActivePlan plan_all_active(const CsrView &g)
{
    AllActivePlan P;
    int64_t nb = 0;
    for (int t = 0; t < kNumTiers - 1; t++)
    {
        P.block_start[t] = (int32_t)nb;
        const int32_t first = t == 0 ? 0 : g.tier_border[t - 1];
        const int32_t last = (t == kNumTiers - 2) ? g.V : g.tier_border[t];
        const int64_t rows = last - first;
        int64_t per = t == 0 ? 1 : (t == 1 ? (kAdvThreads / 32) * kWarpRowsPerWarp : (int64_t)(kAdvThreads / (32 >> (t - 1))) * kGroupPasses);
        nb += (rows + per - 1) / per;
    }
    P.block_start[kNumTiers - 1] = (int32_t)nb;
    P.blocks = nb;
    return P;
}

struct NoVertexOp
{
    __device__ __forceinline__ void operator()(int, int, int) const {}
};

// one row processed by NT cooperating threads (NT = CTA, 32, 16, 8, 4, 2 or 1); `sync` separates pre / edges / post
template <int NT, class EdgeOp, class PreOp, class PostOp, class Sync>
__device__ __forceinline__ void advance_row(const CsrView &g, int32_t row, int tid, long long edge_shift, EdgeOp &edge_op,
                                            PreOp &pre, PostOp &post, Sync sync)
{
    const int64_t s = g.ptr[row], e = g.ptr[row + 1];
    const int deg = (int)(e - s);
    const int lane = threadIdx.x & 31;
    if (tid == 0) pre(row, deg, lane);
    sync();
    for (int64_t p = s + tid; p < e; p += NT) edge_op(row, g.adj[p], (int)(p - s), edge_shift + p, lane);
    sync();
    if (tid == 0) post(row, deg, lane);
}

// lanes of the warp that form this thread's G-lane row group
template <int G>
__device__ __forceinline__ unsigned group_lane_mask()
{
    return G >= 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
}

template <int G, class EdgeOp, class PreOp, class PostOp>
__device__ __forceinline__ void advance_group_rows(const CsrView &g, int32_t row0, int32_t row1, long long edge_shift,
                                                   EdgeOp &edge_op, PreOp &pre, PostOp &post)
{
    constexpr int GROUPS = kAdvThreads / G;
    const int gid = threadIdx.x / G, gl = threadIdx.x % G;
    // the G lanes of a group always take the same branch (they share `row`), so the lanes that must meet at the
    // pre / edges / post barriers are exactly the group's lanes — named explicitly, never __activemask(), which only
    // reports whoever happens to be converged and would let a lane start its edge ops before lane 0 has run pre()
    const unsigned group_mask = group_lane_mask<G>();
    for (int32_t base = row0; base < row1; base += GROUPS)
    {
        const int32_t row = base + gid;
        if (row < row1) advance_row<G>(g, row, gl, edge_shift, edge_op, pre, post, [group_mask] { __syncwarp(group_mask); });
    }
}

// ALL_ACTIVE advance over every vertex (advance_worker ALL_ACTIVE branch, multicore/advance_worker.hpp:204-319)
// The reference passes a second functor triple for the low-degree ("collective") region
// (graph_abstractions.h:96-118, multicore/advance_all_active.hpp:150-229); here it serves the rows with < 32 edges.
template <class EdgeOp, class PreOp, class PostOp, class CEdgeOp, class CPreOp, class CPostOp>
__global__ void __launch_bounds__(kAdvThreads)
advance_all_active_kernel(const CsrView g, const AllActivePlan P, long long edge_shift, EdgeOp edge_op, PreOp pre, PostOp post,
                          CEdgeOp c_edge_op, CPreOp c_pre, CPostOp c_post)
{
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (b < P.block_start[1])
    {
        advance_row<kAdvThreads>(g, b, threadIdx.x, edge_shift, edge_op, pre, post, [] { __syncthreads(); });
    }
    else if (b < P.block_start[2])
    {
        constexpr int ROWS = (kAdvThreads / 32) * kWarpRowsPerWarp;
        const int32_t row0 = g.tier_border[0] + (b - P.block_start[1]) * ROWS;
        const int32_t row1 = min(row0 + ROWS, g.tier_border[1]);
        for (int32_t row = row0 + warp; row < row1; row += kAdvThreads / 32)
            advance_row<32>(g, row, lane, edge_shift, edge_op, pre, post, [] { __syncwarp(); });
    }
    else
    {
        int t = 2;
#pragma unroll
        for (int i = 3; i < kNumTiers - 1; i++)
            if (b >= P.block_start[i]) t = i;
        const int32_t first = g.tier_border[t - 1];
        const int32_t last = (t == kNumTiers - 2) ? g.V : g.tier_border[t];
        const int G = 32 >> (t - 1);
        const int32_t per = (kAdvThreads / G) * kGroupPasses;
        const int32_t row0 = first + (b - P.block_start[t]) * per;
        const int32_t row1 = min(row0 + per, last);
        switch (t)
        {
        case 2: advance_group_rows<16>(g, row0, row1, edge_shift, c_edge_op, c_pre, c_post); break;
        case 3: advance_group_rows<8>(g, row0, row1, edge_shift, c_edge_op, c_pre, c_post); break;
        case 4: advance_group_rows<4>(g, row0, row1, edge_shift, c_edge_op, c_pre, c_post); break;
        case 5: advance_group_rows<2>(g, row0, row1, edge_shift, c_edge_op, c_pre, c_post); break;
        default: advance_group_rows<1>(g, row0, row1, edge_shift, c_edge_op, c_pre, c_post); break;
        }
    }
}

// SPARSE advance over three degree-binned id queues (advance_worker SPARSE branch, multicore/advance_sparse.hpp:7-249)
struct SparseFrontierView
{
    const int32_t *q[3]; // big (CTA per vertex) / mid (warp per vertex) / small (8 lanes per vertex)
    int32_t n[3];
    int32_t blocks_mid, blocks_small;
};

template <class EdgeOp, class PreOp, class PostOp, class CEdgeOp, class CPreOp, class CPostOp>
__global__ void __launch_bounds__(kAdvThreads)
advance_sparse_kernel(const CsrView g, const SparseFrontierView F, long long edge_shift, EdgeOp edge_op, PreOp pre, PostOp post,
                      CEdgeOp c_edge_op, CPreOp c_pre, CPostOp c_post)
{
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (b < F.n[0])
    {
        advance_row<kAdvThreads>(g, F.q[0][b], threadIdx.x, edge_shift, edge_op, pre, post, [] { __syncthreads(); });
    }
    else if (b < F.n[0] + F.blocks_mid)
    {
        const int nwarps = F.blocks_mid * (kAdvThreads / 32);
        for (int i = (b - F.n[0]) * (kAdvThreads / 32) + warp; i < F.n[1]; i += nwarps)
            advance_row<32>(g, F.q[1][i], lane, edge_shift, edge_op, pre, post, [] { __syncwarp(); });
    }
    else
    {
        constexpr int G = 8;
        constexpr int GROUPS = kAdvThreads / G;
        const int ngroups = F.blocks_small * GROUPS;
        const int gid = threadIdx.x / G, gl = threadIdx.x % G;
        const unsigned group_mask = group_lane_mask<G>(); // the loop condition is uniform per group
        for (int i = (b - F.n[0] - F.blocks_mid) * GROUPS + gid; i < F.n[2]; i += ngroups)
            advance_row<G>(g, F.q[2][i], gl, edge_shift, c_edge_op, c_pre, c_post, [group_mask] { __syncwarp(group_mask); });
    }
}

// advance over a CSR whose rows are not degree-sorted (the incoming direction shares the outgoing numbering, so ids say
// nothing about in-degrees): warp per row, optionally restricted to an id list.
template <class EdgeOp, class PreOp, class PostOp>
__global__ void __launch_bounds__(kAdvThreads)
advance_unsorted_kernel(const CsrView g, const int32_t *__restrict__ ids, int32_t n, long long edge_shift, EdgeOp edge_op,
                        PreOp pre, PostOp post)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps)
        advance_row<32>(g, ids ? ids[i] : (int32_t)i, lane, edge_shift, edge_op, pre, post, [] { __syncwarp(); });
}

// compute (common/compute.hpp:62-85): map over all vertices / over a sparse id list
template <class ComputeOp>
__global__ void compute_all_active_kernel(const int64_t *__restrict__ ptr, int32_t V, ComputeOp op)
{
    for (int32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x)
        op(v, (int)(ptr[v + 1] - ptr[v]), (int)(threadIdx.x & 31));
}

template <class ComputeOp>
__global__ void compute_sparse_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ ids, int32_t n, ComputeOp op)
{
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int32_t v = ids[i];
        op(v, (int)(ptr[v + 1] - ptr[v]), (int)(threadIdx.x & 31));
    }
}

// generate_new_frontier, flag pass (common/generate_new_frontier.hpp:4-43): flags[v] = cond(v, deg) as a bitmap word per
// warp via ballot; the compaction into queues is vglb_gnf_from_bitmap behind the C ABI.
template <class Cond>
__global__ void gnf_bitmap_kernel(const int64_t *__restrict__ ptr, int32_t V, uint32_t *__restrict__ bitmap, Cond cond)
{
    const int32_t padded = (V + 31) & ~31;
    for (int32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < padded; v += gridDim.x * blockDim.x)
    {
        const bool in = v < V && cond(v, (int)(ptr[v + 1] - ptr[v])) > 0;
        const uint32_t word = __ballot_sync(0xffffffffu, in);
        if ((threadIdx.x & 31) == 0) bitmap[v >> 5] = word;
    }
}

// reduce (common/reduce.hpp:4-67): block reduction + one atomic per CTA; T in {int, float, double}, accumulated in Acc
template <class Acc, class ReduceOp>
__global__ void reduce_sum_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ ids, int32_t n, Acc *out,
                                  ReduceOp op)
{
    __shared__ Acc s_part[32];
    Acc local = 0;
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int32_t v = ids ? ids[i] : i;
        local += (Acc)op(v, (int)(ptr[v + 1] - ptr[v]), (int)(threadIdx.x & 31));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        Acc t = 0;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; w++) t += s_part[w];
        atomicAdd(out, t);
    }
}

template <class ReduceOp>
__global__ void reduce_max_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ ids, int32_t n, int *out,
                                  ReduceOp op)
{
    int local = INT_MIN;
    for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const int32_t v = ids ? ids[i] : i;
        local = max(local, (int)op(v, (int)(ptr[v + 1] - ptr[v]), (int)(threadIdx.x & 31)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local = max(local, __shfl_xor_sync(0xffffffffu, local, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, local);
}

} // namespace vglb
