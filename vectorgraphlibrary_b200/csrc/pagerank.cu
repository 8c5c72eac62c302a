// pagerank.cu — PageRank "pull" sweep as a vectorised segmented reduction over the degree-sorted CSR (sm_100a).
//
// Reference semantics = the MULTICORE build (algorithms/pr/pr.hpp:7-148):
//     r'[u] = k + d * ( sum_{(u->v) in E, v != u} r[v] * inv[v]  +  D ),   inv[v] = 1/indeg_noloops(v) (0 if none)
//     D = sum_{v : indeg_noloops(v) == 0} r[v] / V,   r0 = 1/V,   d = 0.85f,   k = (1-d)/V,   exactly `iters` sweeps.
// The reference does this with four operator calls per sweep (compute save_old_ranks :85-90, reduce dangling :94-103,
// scatter edge_op+post :105-124, reduce ranks_sum :130-135) = 4 V-passes + one E-pass with TWO gathers per edge.
//
// B200 design: ONE persistent kernel per sweep (one 1024-thread CTA per SM: 1 producer warp + 31 consumer warps), no
// atomics on the rank vector, no host sync inside the loop.
//   * contrib[v] = r[v]*inv[v] is produced by the epilogue of the previous sweep, so the E-pass gathers one fp32 per
//     edge (the product is the same fp32 multiply the reference does per edge, pr.hpp:112-115);
//   * everything that is STREAMED (column indices, row pointers, inverse degrees) reaches the SM through the TMA:
//     rows are cut into CHUNKS of whole rows (<= 4096 edges, <= 1024 rows; boundaries precomputed once per graph by
//     binary search on the row pointers), the producer warp fetches a chunk ticket (one atomic per chunk, heaviest
//     chunks first) and issues three cp.async.bulk copies per chunk into a 3-stage shared-memory ring guarded by
//     full/empty mbarriers. The consumers therefore see only ONE long-latency operation per edge, the gather itself
//     (measured before this change: 60 % of all stall samples were long-scoreboard waits on the ptr -> adj -> gather
//     -> inv chain, and the sweep ran at 1 edge per clock per SM whatever the cache behaviour was);
//   * the gather: a divergent 4-byte gather costs a 32-byte sector of L2 bandwidth. Ids are degree-sorted and in/out
//     degrees of RMAT/Kronecker graphs are correlated, so the first ids are by far the most gathered ones (scale 22:
//     54 % of all edge targets are among the first 32768 ids): every CTA stages the head of the contribution vector in
//     the shared memory the ring leaves free (~140 KB) and serves those gathers from there — north_star (b)
//     "shared-memory staging of hub adjacency" applied to the side of the hubs that is actually hot, their VALUES;
//   * load balance inside a chunk (replaces the ve / vc / collective tiers of multicore/advance_all_active.hpp:7-229):
//     rows are degree-sorted, so a chunk holds rows of nearly equal degree d; G = pow2 >= d/8 lanes (1..32) work on one
//     row, so every lane issues up to EIGHT independent gathers before the first use; the 992/G rows of one pass are
//     adjacent in shared memory. Rows with more than 2048 edges are not staged: the whole CTA streams such a row with
//     int4 loads straight from global memory (8 gathers in flight per thread);
//   * the gathered contribution vector is kept L2-resident (evict-last; 64 MB at scale 24 fits the 126 MB L2), the
//     streamed arrays are evict-first;
//   * the epilogue fuses post-op (:118-121), next sweep's contribution, and the NEXT sweep's dangling mass (summed in
//     fp64, one atomicAdd(double) per warp), so `reduce` never returns to the host.
// Row sums are fp32 trees instead of the reference's sequential fp32 (difference ~1e-7 relative, tolerance 1e-6).
// HBM roofline: algorithmic bytes per sweep = 8E (index + gathered value per edge) + 16V (row pointer, inv read,
// contribution write) [+4V rank write on the last sweep].
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

#define PR_THREADS 1024
#define PR_CONSUMERS (PR_THREADS - 32)
#define PR_CWARPS (PR_CONSUMERS / 32)
#define PR_STAGES 3
#define PR_CAP 4096               // edges per chunk (capacity of a ring stage)
#define PR_ROWCAP 992             // rows per chunk (= consumer threads: one pass at one lane per row)
#define PR_BIG_DEGREE 2048        // rows with more edges are streamed by the whole CTA, one row at a time
#define PR_MID_DEGREE 256         // chunk stride: 2048 edges above this degree, 3840 below (stride + degree <= PR_CAP)
#define PR_SMALL_DEGREE 4         // rows up to this degree: fixed chunks of PR_ROWCAP rows (992 * 4 <= PR_CAP)
#define PR_PER_LANE 8             // independent gathers per lane and step
#define PR_BATCH 16               // tasks per ticket (descriptors prefetched by the lanes of the producer warp)

enum { PR_TASK_DONE = 0, PR_TASK_CHUNK = 1, PR_TASK_BIGROW = 2 };

struct __align__(16) PrStage
{
    int32_t adj[PR_CAP + 8];
    int64_t ptr[PR_ROWCAP + 4];
    float inv[PR_ROWCAP + 8];
    int32_t desc[8]; // type, first row, row count, adj offset, ptr offset, inv offset, first edge (lo, hi; 16-byte aligned down)
};

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
    const int32_t *chunk_row; // chunk k covers rows [chunk_row[k], chunk_row[k+1])
    unsigned int *ticket;     // this sweep's task ticket counter (zeroed before the sweep)
    int32_t V;
    int32_t hot;              // number of leading contributions staged in shared memory (multiple of 4)
    int32_t big_rows;         // rows [0, big_rows) have more than PR_BIG_DEGREE edges: tasks [0, big_rows)
    int32_t chunks;           // tasks [big_rows, big_rows + chunks)
    float k, d, v_as_float;
};

struct L2Pol
{
    uint64_t stream, keep;
};

extern __shared__ __align__(16) unsigned char pr_smem[];

// ---- mbarrier / TMA (1-D bulk copy) primitives -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!done);
}
// global -> shared bulk copy through the TMA; bytes and both addresses are multiples of 16
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :
                 : "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void consumer_bar_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(PR_CONSUMERS) : "memory");
}

__device__ __forceinline__ float pr_gather(const PrParams &P, const float *s_hot, const L2Pol &pol, int32_t v)
{
    return v < P.hot ? s_hot[v] : ld_gather_f32(P.contrib_in + v, pol.keep);
}

__device__ __forceinline__ void pr_epilogue(const PrParams &P, int32_t row, float inv_r, float sum, float dang, double &dang_local)
{
    // k + d * (rank + dangling) — pr.hpp:118-121, no FMA contraction (the x86-64 reference build has none)
    const float rank = __fadd_rn(P.k, __fmul_rn(P.d, __fadd_rn(sum, dang)));
    P.contrib_out[row] = __fmul_rn(rank, inv_r);
    if (P.rank_out) P.rank_out[row] = rank;
    if (inv_r == 0.0f) dang_local += (double)__fdiv_rn(rank, P.v_as_float); // pr.hpp:94-101
}

// one staged chunk: G lanes per row, 992/G rows per pass, every lane gathers up to 8 values per step
template <int G>
__device__ __forceinline__ void pr_chunk_rows(const PrParams &P, const float *s_hot, const L2Pol &pol, const PrStage &S, int ct,
                                              float dang, double &dang_local)
{
    constexpr int GROUPS = PR_CONSUMERS / G;
    const int32_t ra = S.desc[1], nrows = S.desc[2];
    const int32_t *s_adj = S.adj + S.desc[3];
    const int64_t *s_ptr = S.ptr + S.desc[4];
    const float *s_inv = S.inv + S.desc[5];
    const int64_t e0 = s_ptr[0]; // first edge of the chunk
    const int gid = ct / G, gl = ct % G;
    for (int32_t base = 0; base < nrows; base += GROUPS)
    {
        const int32_t r = base + gid;
        float acc = 0.f;
        if (r < nrows)
        {
            const int32_t row = ra + r;
            const int32_t s = (int32_t)(s_ptr[r] - e0), e = (int32_t)(s_ptr[r + 1] - e0);
            for (int32_t pb = s + gl; pb < e; pb += PR_PER_LANE * G)
            {
                int32_t v[PR_PER_LANE];
#pragma unroll
                for (int j = 0; j < PR_PER_LANE; j++)
                {
                    const int32_t p = pb + j * G;
                    v[j] = p < e ? s_adj[p] : row;
                }
                float a[PR_PER_LANE];
#pragma unroll
                for (int j = 0; j < PR_PER_LANE; j++) a[j] = v[j] != row ? pr_gather(P, s_hot, pol, v[j]) : 0.f;
                acc += ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
            }
        }
        __syncwarp();
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (gl == 0 && r < nrows) pr_epilogue(P, ra + r, s_inv[r], acc, dang, dang_local);
    }
}

// a row with more than PR_BIG_DEGREE edges: all consumer threads stream it with int4 loads from global memory
__device__ __forceinline__ float pr_big_row_partial(const PrParams &P, const float *s_hot, const L2Pol &pol, int32_t row, int64_t s,
                                                    int64_t e, int ct)
{
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const int64_t s4 = (s + 3) & ~(int64_t)3;
    const int64_t e4 = e & ~(int64_t)3;
    // head (< 4 elements) and tail (< 4 elements)
    if (s + ct < s4)
    {
        const int32_t v = ld_stream_s32(P.adj + s + ct, pol.stream);
        if (v != row) acc1 += pr_gather(P, s_hot, pol, v);
    }
    if (e4 + ct < e)
    {
        const int32_t v = ld_stream_s32(P.adj + e4 + ct, pol.stream);
        if (v != row) acc2 += pr_gather(P, s_hot, pol, v);
    }
    const int4 *adj4 = reinterpret_cast<const int4 *>(P.adj);
    const int64_t q_end = e4 >> 2;
    int64_t q = (s4 >> 2) + ct;
    // two vectors (8 gathers) in flight per thread
    for (; q + PR_CONSUMERS < q_end; q += 2 * PR_CONSUMERS)
    {
        const int4 a = ld_stream_v4(adj4 + q, pol.stream);
        const int4 b = ld_stream_v4(adj4 + q + PR_CONSUMERS, pol.stream);
        const float a0 = a.x != row ? pr_gather(P, s_hot, pol, a.x) : 0.f;
        const float a1 = a.y != row ? pr_gather(P, s_hot, pol, a.y) : 0.f;
        const float a2 = a.z != row ? pr_gather(P, s_hot, pol, a.z) : 0.f;
        const float a3 = a.w != row ? pr_gather(P, s_hot, pol, a.w) : 0.f;
        const float b0 = b.x != row ? pr_gather(P, s_hot, pol, b.x) : 0.f;
        const float b1 = b.y != row ? pr_gather(P, s_hot, pol, b.y) : 0.f;
        const float b2 = b.z != row ? pr_gather(P, s_hot, pol, b.z) : 0.f;
        const float b3 = b.w != row ? pr_gather(P, s_hot, pol, b.w) : 0.f;
        acc0 += a0 + b0;
        acc1 += a1 + b1;
        acc2 += a2 + b2;
        acc3 += a3 + b3;
    }
    if (q < q_end)
    {
        const int4 a = ld_stream_v4(adj4 + q, pol.stream);
        if (a.x != row) acc0 += pr_gather(P, s_hot, pol, a.x);
        if (a.y != row) acc1 += pr_gather(P, s_hot, pol, a.y);
        if (a.z != row) acc2 += pr_gather(P, s_hot, pol, a.z);
        if (a.w != row) acc3 += pr_gather(P, s_hot, pol, a.w);
    }
    return (acc0 + acc1) + (acc2 + acc3);
}

__global__ void __launch_bounds__(PR_THREADS, 1) pr_sweep_kernel(const __grid_constant__ PrParams P)
{
    // shared memory: [ring stages][full/empty barriers][big-row partials][hot contributions]
    PrStage *stages = reinterpret_cast<PrStage *>(pr_smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(pr_smem + PR_STAGES * sizeof(PrStage));
    uint64_t *empty = full + PR_STAGES;
    float *s_part = reinterpret_cast<float *>(empty + PR_STAGES); // [PR_STAGES][32]
    float *s_hot = s_part + PR_STAGES * 32;

    L2Pol pol;
    pol.stream = l2_policy_evict_first();
    pol.keep = l2_policy_evict_last();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < PR_STAGES; s++)
        {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], PR_CWARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // stage the hot prefix of the contribution vector (read through L2, which every CTA shares)
    {
        const float4 *src4 = reinterpret_cast<const float4 *>(P.contrib_in);
        float4 *dst4 = reinterpret_cast<float4 *>(s_hot);
        for (int i = threadIdx.x; i < (P.hot >> 2); i += PR_THREADS) dst4[i] = __ldg(src4 + i);
    }
    __syncthreads();

    const int32_t total = P.big_rows + P.chunks;
    if (warp == 0)
    {
        // ---- producer warp: hands out tasks in batches and feeds the ring through the TMA. Every lane prefetches the
        // descriptor of one task of the NEXT batch (chunk rows, first/last edge) while the current batch is issued, so
        // the ticket atomic and the two dependent loads are off the consumers' critical path ----
        int stage = 0;
        uint32_t phase = 0;
        auto fetch = [&](int32_t &base, int32_t &ra, int32_t &rb, int64_t &ea, int64_t &eb) {
            base = 0;
            if (lane == 0) base = (int32_t)atomicAdd(P.ticket, (unsigned int)PR_BATCH);
            base = __shfl_sync(0xffffffffu, base, 0);
            const int32_t task = base + lane;
            ra = rb = 0;
            ea = eb = 0;
            if (lane < PR_BATCH && task < total && task >= P.big_rows)
            {
                const int32_t k = task - P.big_rows;
                ra = P.chunk_row[k];
                rb = P.chunk_row[k + 1];
                ea = P.ptr[ra];
                eb = P.ptr[rb];
            }
        };
        int32_t base, ra, rb, nbase, nra, nrb;
        int64_t ea, eb, nea, neb;
        fetch(base, ra, rb, ea, eb);
        for (;;)
        {
            if (base >= total)
            {
                if (lane == 0)
                {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    stages[stage].desc[0] = PR_TASK_DONE;
                    mbar_arrive(&full[stage]);
                }
                break;
            }
            fetch(nbase, nra, nrb, nea, neb);
            for (int i = 0; i < PR_BATCH && base + i < total; i++)
            {
                if (lane == i)
                {
                    const int32_t task = base + i;
                    mbar_wait(&empty[stage], phase ^ 1u);
                    PrStage &S = stages[stage];
                    if (task < P.big_rows)
                    {
                        S.desc[0] = PR_TASK_BIGROW;
                        S.desc[1] = task;
                        mbar_arrive(&full[stage]);
                    }
                    else
                    {
                        const int64_t ea_al = ea & ~(int64_t)3;
                        const int32_t ra_ptr = ra & ~1, ra_inv = ra & ~3;
                        const uint32_t adj_bytes = (uint32_t)(((eb - ea_al) * 4 + 15) & ~(int64_t)15);
                        const uint32_t ptr_bytes = (uint32_t)((((rb + 1 - ra_ptr) * 8) + 15) & ~15);
                        const uint32_t inv_bytes = (uint32_t)((((rb - ra_inv) * 4) + 15) & ~15);
                        S.desc[0] = PR_TASK_CHUNK;
                        S.desc[1] = ra;
                        S.desc[2] = rb - ra;
                        S.desc[3] = (int32_t)(ea - ea_al);
                        S.desc[4] = ra - ra_ptr;
                        S.desc[5] = ra - ra_inv;
                        mbar_arrive_expect_tx(&full[stage], adj_bytes + ptr_bytes + inv_bytes);
                        if (adj_bytes) tma_load_1d(S.adj, P.adj + ea_al, adj_bytes, &full[stage], pol.stream);
                        tma_load_1d(S.ptr, P.ptr + ra_ptr, ptr_bytes, &full[stage], pol.stream);
                        if (inv_bytes) tma_load_1d(S.inv, P.inv + ra_inv, inv_bytes, &full[stage], pol.stream);
                    }
                }
                // lanes must issue strictly in order: a lane running a full ring ahead would see its parity wait on a
                // stage succeed against the PREVIOUS phase and overwrite a stage that is still being read
                __syncwarp();
                if (++stage == PR_STAGES)
                {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            base = nbase; ra = nra; rb = nrb; ea = nea; eb = neb;
        }
        return;
    }

    // ---- consumers ----
    const int ct = threadIdx.x - 32;
    const int cwarp = warp - 1;
    const float dang = (float)(*P.dangling_in);
    double dang_local = 0.0;
    int stage = 0;
    uint32_t phase = 0;
    for (;;)
    {
        mbar_wait(&full[stage], phase);
        const PrStage &S = stages[stage];
        const int type = S.desc[0];
        if (type == PR_TASK_DONE) break;
        if (type == PR_TASK_CHUNK)
        {
            // rows of a chunk have nearly equal degrees: G = pow2 >= (largest degree) / 8 lanes per row
            const int64_t *s_ptr = S.ptr + S.desc[4];
            const int32_t dmax = (int32_t)(s_ptr[1] - s_ptr[0]);
            if (dmax > 128) pr_chunk_rows<32>(P, s_hot, pol, S, ct, dang, dang_local);
            else if (dmax > 64) pr_chunk_rows<16>(P, s_hot, pol, S, ct, dang, dang_local);
            else if (dmax > 32) pr_chunk_rows<8>(P, s_hot, pol, S, ct, dang, dang_local);
            else if (dmax > 16) pr_chunk_rows<4>(P, s_hot, pol, S, ct, dang, dang_local);
            else if (dmax > 8) pr_chunk_rows<2>(P, s_hot, pol, S, ct, dang, dang_local);
            else pr_chunk_rows<1>(P, s_hot, pol, S, ct, dang, dang_local);
        }
        else
        {
            const int32_t row = S.desc[1];
            const int64_t s = P.ptr[row], e = P.ptr[row + 1];
            float acc = pr_big_row_partial(P, s_hot, pol, row, s, e, ct);
            acc = warp_sum_f32(acc);
            float *part = s_part + stage * 32;
            if (lane == 0) part[cwarp] = acc;
            consumer_bar_sync();
            if (cwarp == 0)
            {
                float t = lane < PR_CWARPS ? part[lane] : 0.f;
                t = warp_sum_f32(t);
                if (lane == 0) pr_epilogue(P, row, P.inv[row], t, dang, dang_local);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == PR_STAGES)
        {
            stage = 0;
            phase ^= 1u;
        }
    }
    // next sweep's dangling mass: fp64 warp reduction, one atomic per warp that has any
    dang_local = warp_sum_f64(dang_local);
    if (lane == 0 && dang_local != 0.0) atomicAdd(P.dangling_out, dang_local);
}

#define PR_RING_BYTES (PR_STAGES * sizeof(PrStage) + 2 * PR_STAGES * sizeof(uint64_t) + PR_STAGES * 32 * sizeof(float))
#define PR_SMEM_MAX (227 * 1024)
#define PR_HOT_MAX ((int)((PR_SMEM_MAX - PR_RING_BYTES) / 4) & ~3)

// number of rows with degree >= threshold (rows are degree-sorted descending)
__device__ int32_t pr_rows_with_degree_at_least(const int64_t *ptr, int32_t V, int64_t threshold)
{
    int32_t lo = 0, hi = V; // first row with degree < threshold
    while (lo < hi)
    {
        const int32_t mid = lo + ((hi - lo) >> 1);
        if (ptr[mid + 1] - ptr[mid] >= threshold) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// region borders: [0] rows > PR_BIG_DEGREE | [1] rows > PR_MID_DEGREE | [2] rows > PR_SMALL_DEGREE | V; and their first edges
__global__ void pr_region_kernel(const int64_t *__restrict__ ptr, int32_t V, int64_t *__restrict__ out /* [3 rows][3 edges] */)
{
    const int64_t thr[3] = {PR_BIG_DEGREE + 1, PR_MID_DEGREE + 1, PR_SMALL_DEGREE + 1};
    if (threadIdx.x < 3)
    {
        const int32_t r = pr_rows_with_degree_at_least(ptr, V, thr[threadIdx.x]);
        out[threadIdx.x] = r;
        out[3 + threadIdx.x] = ptr[r];
    }
}

// chunk k of a region starts at the first row r in [first, last] with ptr[r] - ptr[first] >= k * stride (stride = 0:
// fixed chunks of PR_ROWCAP rows)
__global__ void pr_chunk_rows_kernel(const int64_t *__restrict__ ptr, int32_t first, int32_t last, int32_t stride,
                                     int32_t nchunks, int32_t *__restrict__ chunk_row)
{
    const int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nchunks) return;
    if (stride == 0)
    {
        chunk_row[k] = first + k * PR_ROWCAP;
        return;
    }
    const int64_t target = ptr[first] + (int64_t)k * stride;
    int32_t lo = first, hi = last; // smallest r in [first, last] with ptr[r] >= target
    while (lo < hi)
    {
        const int32_t mid = lo + ((hi - lo) >> 1);
        if (ptr[mid] >= target) hi = mid;
        else lo = mid + 1;
    }
    chunk_row[k] = lo;
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
    __shared__ double s_dang[8];
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
        for (int w = 0; w < 8; w++) t += s_dang[w];
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
        // +16 bytes: these arrays are read in 16-byte units (TMA bulk copies / float4 staging of the hot prefix)
        CUDA_TRY(cudaMalloc(&g->d_pr_inv, (size_t)g->V * 4 + 16));
        CUDA_TRY(cudaMalloc(&g->d_pr_contrib[0], (size_t)g->V * 4 + 16));
        CUDA_TRY(cudaMalloc(&g->d_pr_contrib[1], (size_t)g->V * 4 + 16));
        pr_inverse_degree_kernel<<<(unsigned)ceil_div64(g->V, 256), 256, 0, ctx->stream>>>(d_indeg, g->V, g->d_pr_inv);
        KERNEL_TRY();
        ctx->launches++;
        // chunk table: regions of the degree-sorted row range, then chunk boundaries inside each region
        int64_t *d_reg = (int64_t *)(ctx->d_counters + 48);
        int64_t reg[6];
        pr_region_kernel<<<1, 32, 0, ctx->stream>>>(g->d_out_ptr, g->V, d_reg);
        KERNEL_TRY();
        ctx->launches++;
        CUDA_TRY(cudaMemcpyAsync(reg, d_reg, sizeof(reg), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        const int32_t R0 = (int32_t)reg[0], R1 = (int32_t)reg[1], R2 = (int32_t)reg[2], V = g->V;
        const int64_t nA = ceil_div64(reg[4] - reg[3], PR_CAP - PR_BIG_DEGREE);
        const int64_t nB = ceil_div64(reg[5] - reg[4], PR_CAP - PR_MID_DEGREE);
        const int64_t nC = ceil_div64((int64_t)V - R2, PR_ROWCAP);
        VGLB_REQUIRE(nA + nB + nC < 0x7fffffffLL - V, "vglb_pagerank: too many chunks");
        g->pr_big_rows = R0;
        g->pr_chunks = (int32_t)(nA + nB + nC);
        CUDA_TRY(cudaMalloc(&g->d_pr_chunk_row, ((size_t)g->pr_chunks + 1) * 4));
        const int32_t first[3] = {R0, R1, R2}, last[3] = {R1, R2, V};
        const int32_t stride[3] = {PR_CAP - PR_BIG_DEGREE, PR_CAP - PR_MID_DEGREE, 0};
        const int64_t cnt[3] = {nA, nB, nC};
        int64_t off = 0;
        for (int r = 0; r < 3; r++)
        {
            if (cnt[r] > 0)
            {
                pr_chunk_rows_kernel<<<(unsigned)ceil_div64(cnt[r], 256), 256, 0, ctx->stream>>>(
                    g->d_out_ptr, first[r], last[r], stride[r], (int32_t)cnt[r], g->d_pr_chunk_row + off);
                KERNEL_TRY();
                ctx->launches++;
            }
            off += cnt[r];
        }
        CUDA_TRY(cudaMemcpyAsync(g->d_pr_chunk_row + off, &g->V, 4, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_indeg);
        CUDA_TRY(cudaFuncSetAttribute(pr_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM_MAX));
    }
    if (g->pr_dangling_slots < iters + 1)
    {
        cudaFree(g->d_pr_dangling);
        g->d_pr_dangling = NULL;
        // per sweep: one fp64 dangling slot and one ticket counter
        CUDA_TRY(cudaMalloc(&g->d_pr_dangling, (size_t)(iters + 1) * (sizeof(double) + sizeof(unsigned int))));
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
    P.chunk_row = g->d_pr_chunk_row;
    P.big_rows = g->pr_big_rows;
    P.chunks = g->pr_chunks;
    P.V = V;
    int64_t hot_max = PR_HOT_MAX;
    if (const char *e = getenv("VGLB_PR_HOT")) hot_max = atol(e) < PR_HOT_MAX ? (atol(e) & ~3L) : PR_HOT_MAX; // tuning knob
    P.hot = (int32_t)(((int64_t)V < hot_max ? ((int64_t)V + 3) & ~3LL : hot_max));
    P.d = damping;
    P.k = (float)((1.0 - (double)damping) / (double)((float)V)); // pr.hpp:37-38
    P.v_as_float = (float)V;
    unsigned int *tickets = (unsigned int *)(g->d_pr_dangling + (iters + 1));
    const size_t smem = PR_RING_BYTES + (size_t)P.hot * 4;

    CUDA_TRY(cudaEventRecord(ctx->ev_start, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(g->d_pr_dangling, 0, (size_t)(iters + 1) * (sizeof(double) + sizeof(unsigned int)), ctx->stream));
    const float r0 = (float)(1.0 / (double)V); // pr.hpp:42
    if (iters == 0)
    {
        pr_fill_kernel<<<(unsigned)ceil_div64(V, 256), 256, 0, ctx->stream>>>(d_ranks, V, r0);
        KERNEL_TRY();
        ctx->launches++;
    }
    else
    {
        pr_init_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(g->d_pr_inv, V, r0, P.v_as_float, g->d_pr_contrib[0],
                                                                  g->d_pr_dangling);
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
        P.ticket = tickets + it;
        pr_sweep_kernel<<<ctx->sm_count, PR_THREADS, smem, ctx->stream>>>(P);
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
