// cc.cu — connected components by min-label hooking + pointer jumping over the device VectCSR (sm_100a).
//
// Reference: ConnectedComponents::vgl_shiloach_vishkin (algorithms/cc/shiloach_vishkin.hpp:7-88):
//   compute(comp[v] = v)                                           (:19-24, v = SCATTER-sorted id)
//   while (hook_changes):
//       scatter(all-active, if comp[src] < comp[dst] then comp[dst] = comp[src])   (:30-53, racy but monotone)
//       while (jump_changes): compute(comp[v] = comp[comp[v]])     (:55-76)
// The unique fixed point is comp[v] = min sorted id over {v} U ancestors(v), the component minimum on
// symmetric graphs, so any schedule of monotone hooks and jumps reaches bit-identical labels.
//
// B200 design: the hook is the lambda-generic all-active advance of include/vgl_b200/advance.cuh (the same template
// the GraphAbstractionsB200 shim instantiates for user lambdas) with an atomicMin edge op — one launch covering all
// degree tiers; hooks are applied in place, so a label travels many hops within one round (Gauss-Seidel), and the
// whole jump loop of the reference collapses into ONE kernel that chases every vertex to its current root
// (comp[x] <= x always holds, so the chase terminates and concurrent writes only shorten it). The convergence flag is
// a device word read back once per round (the reference returns a reduce<int> to the host per round, :47-52).
// HBM roofline: algorithmic bytes = sum_hook [ 8E + 12V ] + sum_jump 8V.
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "vgl_b200/advance.cuh"

struct CcHookOp
{
    int32_t *comp;
    int *changed;
    __device__ __forceinline__ void operator()(int src, int dst, int, long long, int) const
    {
        const int32_t cs = comp[src];
        if (cs < comp[dst])                       // shiloach_vishkin.hpp:38-46
        {
            if (atomicMin(&comp[dst], cs) > cs) *changed = 1;
        }
    }
};

__global__ void cc_init_kernel(int32_t *__restrict__ comp, int32_t V)
{
    for (int32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) comp[v] = v;
}

// all jump rounds of shiloach_vishkin.hpp:55-76 at once: comp[v] = root of v's current label chain
__global__ void cc_jump_kernel(int32_t *__restrict__ comp, int32_t V)
{
    for (int32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x)
    {
        const int32_t c = comp[v];
        int32_t r = c;
        for (;;)
        {
            const int32_t n = comp[r];
            if (n == r) break;
            r = n;
        }
        if (r != c) comp[v] = r;
    }
}


// ---- 1D-partitioned CC -----------------------------------------------------------------------------------------------
// Labels stay the reference's SCATTER-sorted ids (so the result is bit-identical to one GPU): the replicated label
// vector is indexed by column id and initialised with the sorted id of every column. Each rank hooks over the rows it
// owns into its replica, one allreduce(min) per round combines the replicas (the reference exchanges the whole array
// after its advance too, mpi_exchange.hpp:155-271), and every rank then runs the same pointer-jump over the whole
// vector (deterministic: a chain's root never changes during the jump), so the replicas stay identical.
struct CcHookPartOp
{
    int32_t *comp; // by column id
    int *changed;
    int32_t col0;
    __device__ __forceinline__ void operator()(int src, int dst, int, long long, int) const
    {
        const int32_t cs = comp[col0 + src];
        if (cs < comp[dst])
        {
            if (atomicMin(&comp[dst], cs) > cs) *changed = 1;
        }
    }
};

__device__ __forceinline__ int32_t cc_col_of_sorted(int32_t s, int32_t P, int32_t vp) { return (s % P) * vp + s / P; }

__global__ void cc_part_init_kernel(int32_t *__restrict__ comp, int64_t cols, int32_t P, int32_t vp)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += (int64_t)gridDim.x * blockDim.x)
        comp[c] = (int32_t)(c % vp) * P + (int32_t)(c / vp);
}

__global__ void cc_part_jump_kernel(int32_t *__restrict__ comp, int64_t cols, int32_t P, int32_t vp)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += (int64_t)gridDim.x * blockDim.x)
    {
        const int32_t l = comp[c];
        int32_t r = l;
        for (;;)
        {
            const int32_t n = comp[cc_col_of_sorted(r, P, vp)];
            if (n == r) break;
            r = n;
        }
        if (r != l) comp[c] = r;
    }
}

static int cc_partitioned(vglb_ctx *ctx, vglb_graph *g, int32_t *d_labels, vglb_stats *stats)
{
    CUDA_TRY(cudaSetDevice(ctx->device));
    vglb_comm *comm = g->comm;
    const int64_t launches0 = ctx->launches;
    const int32_t rows = g->V, P = g->part_world, vp = g->vp;
    cudaStream_t st = ctx->stream;
    int *d_changed = (int *)(ctx->d_counters + 40);
    int *h_changed = (int *)(ctx->h_counters + 40);
    if (!g->d_part_vec) CUDA_TRY(cudaMalloc(&g->d_part_vec, (size_t)g->cols * 4));
    int32_t *comp = (int32_t *)g->d_part_vec;

    vglb::CsrView view;
    view.ptr = g->d_out_ptr;
    view.adj = g->d_out_adj;
    view.V = rows;
    for (int t = 0; t < VGLB_NUM_TIERS; t++) view.tier_border[t] = g->tier_border[t];
    const vglb::AllActivePlan plan = vglb::plan_all_active(view);
    VGLB_REQUIRE(plan.blocks < 0x7fffffffLL, "vglb_cc: grid too large");
    CcHookPartOp hook{comp, d_changed, g->col_of_row0};
    vglb::NoVertexOp none;

    CUDA_TRY(cudaEventRecord(ctx->ev_start, st));
    cc_part_init_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(comp, g->cols, P, vp);
    KERNEL_TRY();
    ctx->launches++;
    int64_t hook_rounds = 0;
    int rc;
    for (;;)
    {
        CUDA_TRY(cudaMemsetAsync(d_changed, 0, sizeof(int), st));
        if (plan.blocks > 0)
        {
            vglb::advance_all_active_kernel<<<(unsigned)plan.blocks, vglb::kAdvThreads, 0, st>>>(view, plan, 0LL, hook, none, none, hook, none, none);
            KERNEL_TRY();
            ctx->launches++;
        }
        hook_rounds++;
        rc = vglb_comm_allreduce_async(comm, comp, (size_t)g->cols, VGLB_DT_I32, VGLB_OP_MIN);
        if (rc != VGLB_OK) return rc;
        rc = vglb_comm_allreduce_async(comm, d_changed, 1, VGLB_DT_I32, VGLB_OP_MAX);
        if (rc != VGLB_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(h_changed, d_changed, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (!*h_changed) break;
        cc_part_jump_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(comp, g->cols, P, vp);
        KERNEL_TRY();
        ctx->launches++;
    }
    if (rows > 0) CUDA_TRY(cudaMemcpyAsync(d_labels, comp + g->col_of_row0, (size_t)rows * 4, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaEventRecord(ctx->ev_stop, st));
    CUDA_TRY(cudaEventSynchronize(ctx->ev_stop));
    if (stats)
    {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
        memset(stats, 0, sizeof(*stats));
        stats->seconds = ms * 1e-3;
        stats->iterations = hook_rounds;
        stats->edges_inspected = hook_rounds * g->E;
        stats->vertices_processed = hook_rounds * (int64_t)rows;
        stats->frontier_bytes = hook_rounds * g->cols * 4 * 2; // the allreduced label vector
        stats->algorithmic_bytes = hook_rounds * (8 * g->E + 12 * (int64_t)rows) + (hook_rounds - 1) * 8 * g->cols + stats->frontier_bytes;
        stats->kernel_launches = ctx->launches - launches0;
    }
    return VGLB_OK;
}

extern "C" int vglb_cc(vglb_ctx *ctx, vglb_graph *g, int32_t *d_labels, vglb_stats *stats)
{
    VGLB_REQUIRE(ctx != NULL && g != NULL && d_labels != NULL, "vglb_cc: NULL argument");
    if (g->comm) return cc_partitioned(ctx, g, d_labels, stats);
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    const int32_t V = g->V;
    cudaStream_t st = ctx->stream;
    int *d_changed = (int *)(ctx->d_counters + 40);
    int *h_changed = (int *)(ctx->h_counters + 40);

    vglb::CsrView view;
    view.ptr = g->d_out_ptr;
    view.adj = g->d_out_adj;
    view.V = V;
    for (int t = 0; t < VGLB_NUM_TIERS; t++) view.tier_border[t] = g->tier_border[t];
    const vglb::AllActivePlan plan = vglb::plan_all_active(view);
    VGLB_REQUIRE(plan.blocks < 0x7fffffffLL, "vglb_cc: grid too large");
    CcHookOp hook{d_labels, d_changed};
    vglb::NoVertexOp none;

    CUDA_TRY(cudaEventRecord(ctx->ev_start, st));
    cc_init_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(d_labels, V);
    KERNEL_TRY();
    ctx->launches++;
    int64_t hook_rounds = 0;
    for (;;)
    {
        CUDA_TRY(cudaMemsetAsync(d_changed, 0, sizeof(int), st));
        if (plan.blocks > 0)
        {
            vglb::advance_all_active_kernel<<<(unsigned)plan.blocks, vglb::kAdvThreads, 0, st>>>(view, plan, 0LL, hook, none, none, hook, none, none);
            KERNEL_TRY();
            ctx->launches++;
        }
        hook_rounds++;
        CUDA_TRY(cudaMemcpyAsync(h_changed, d_changed, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (!*h_changed) break;
        cc_jump_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(d_labels, V);
        KERNEL_TRY();
        ctx->launches++;
    }
    CUDA_TRY(cudaEventRecord(ctx->ev_stop, st));
    CUDA_TRY(cudaEventSynchronize(ctx->ev_stop));
    if (stats)
    {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_stop));
        memset(stats, 0, sizeof(*stats));
        stats->seconds = ms * 1e-3;
        stats->iterations = hook_rounds;
        stats->edges_inspected = hook_rounds * g->E;
        stats->vertices_processed = hook_rounds * (int64_t)V;
        stats->algorithmic_bytes = hook_rounds * (8 * g->E + 12 * (int64_t)V) + (hook_rounds - 1) * 8 * (int64_t)V;
        stats->kernel_launches = ctx->launches - launches0;
    }
    return VGLB_OK;
}
