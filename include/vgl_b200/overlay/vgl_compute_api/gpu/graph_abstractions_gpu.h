// vgl_compute_api/gpu/graph_abstractions_gpu.h — B200 (sm_100a) implementation of VGL's GPU backend slot.
//
// HOW IT PLUGS IN. The reference selects its backend at compile time: `-D __USE_GPU__` makes
// architecture_independent_api.h:41-43 define VGL_GRAPH_ABSTRACTIONS as GraphAbstractionsGPU and makes
// vgl_compute_api/common/graph_abstractions.h:188-190 do `#include "vgl_compute_api/gpu/graph_abstractions_gpu.h"` — a path
// that is resolved through the -I list (it does not exist relative to common/). Putting this directory first,
//     nvcc -D __USE_GPU__ -I <repo>/include/vgl_b200/overlay -I <repo>/include -I <VGL checkout> ... -lvgl_b200
// replaces the whole vgl_compute_api/gpu directory of the reference by this file WITHOUT touching the checkout: graph_library.h,
// the data structures (VGL_Graph, VerticesArray, EdgesArray, VGL_Frontier — managed memory under __USE_GPU__,
// memory_API.hpp:3-15) and every algorithms/*.hpp compile unchanged and run on this backend. The class keeps the name
// GraphAbstractionsGPU because the reference names it directly: `friend class GraphAbstractionsGPU` in every frontier
// container (frontier_vect_csr.h:50-52) and `GraphAbstractionsGPU graph_API(_graph, SCATTER)` in
// algorithms/sssp/gpu_shortest_paths.hpp:12,75,138. oracle/Makefile builds the same harness twice — once against the
// reference's own gpu directory, once against this overlay — and tests/test_gpu_dropin.py compares both with the oracle.
//
// Interface = vgl_compute_api/gpu/graph_abstractions_gpu.h:16-190 of the reference (ctor, scatter / gather in the 8- and
// 1-functor forms, compute, reduce<T>, generate_new_frontier, the per-container workers the common dispatchers call,
// friend class GraphAbstractions), recipe manuals/add_new_architecture.txt.
//
// WHAT RUNS WHERE. Device lambdas cannot cross a C ABI, so the lambda-templated kernels are header code
// (include/vgl_b200/advance.cuh: merge-path load-balanced advance, compute, reduce, filter) instantiated in the caller's
// translation unit; everything that does not depend on a lambda is libvgl_b200 behind the C ABI (include/vgl_b200.h):
// the degree-tier analysis of each direction's CSR (vglb_graph_borrow_csr: the reference's managed arrays are used in
// place, nothing is copied), the one-pass order-preserving compaction of generate_new_frontier (vglb_gnf_from_bitmap,
// written straight into the reference frontier's ids[]), add_vertex / add_group_of_vertices bookkeeping
// (vglb_frontier_set_ids), stream and counters.
//
// Host-visible behaviour is the reference's: every operator returns after the device has finished (the reference ends each
// in cudaDeviceSynchronize; algorithms read managed flags on the host right after, e.g. gpu_shiloach_vishkin.hpp:55),
// errors are `throw const char*`.
#pragma once

#include <cuda_runtime.h>

#include <map>

#include "vgl_b200.h"
#include "vgl_b200/advance.cuh"

#include "vector_register/vector_registers.h"

// lane / warp id helpers of the reference's gpu/helpers.hpp:5-19 (user lambdas may call them)
__forceinline__ __device__ unsigned lane_id()
{
    unsigned ret;
    asm volatile("mov.u32 %0, %laneid;" : "=r"(ret));
    return ret;
}

__forceinline__ __device__ unsigned warp_id()
{
    unsigned ret;
    asm volatile("mov.u32 %0, %warpid;" : "=r"(ret));
    return ret;
}

class GraphAbstractionsGPU : public GraphAbstractions
{
private:
    // one direction of the attached VGL_Graph as libvgl_b200 sees it
    struct Direction
    {
        VectorCSRGraph *container = NULL;
        vglb_graph *graph = NULL;
        vglb_graph_info info;
        long long hub_edges = 0; // edges of the rows with >= 4096 edges
    };
    // what the backend keeps per reference frontier object: the compaction state behind the C ABI plus a token that tells
    // whether the frontier still is what this backend generated (host-side add_vertex / add_group_of_vertices /
    // set_all_active overwrite the reference's own fields, see refresh())
    struct FrontierState
    {
        vglb_frontier *frontier = NULL;
        vglb_graph *graph = NULL;
        int token = 0;
        vglb_frontier_info info;
    };

    vglb_ctx *ctx;
    cudaStream_t stream;
    Direction directions[2]; // [SCATTER] outgoing, [GATHER] incoming
    std::map<BaseFrontier *, FrontierState> frontiers;
    uint32_t *filter_bitmap;
    double *reduce_buffer;
    int max_blocks;
    int next_token;
    bool use_safe_stores;

    static void check(int _rc)
    {
        if (_rc != VGLB_OK) throw vglb_last_error(); // const char*, like the reference's throw "literal"
    }
    static void launch_check()
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) throw cudaGetErrorString(e); // SAFE_KERNEL_CALL (cuda_error_handling.h:15-27)
    }
    void finish() // the reference ends every operator in cudaDeviceSynchronize (gpu/advance_vect_csr.hpp:131)
    {
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) throw cudaGetErrorString(e);
    }

    Direction &direction_of(VectorCSRGraph &_graph)
    {
        for (int d = 0; d < 2; d++)
            if (directions[d].container == &_graph) return attach(directions[d]);
        throw "Error in GraphAbstractionsGPU (B200): the container does not belong to the attached VGL_Graph";
    }
    Direction &attach(Direction &_dir)
    {
        if (_dir.graph == NULL)
        {
            VectorCSRGraph *c = _dir.container;
            check(vglb_graph_borrow_csr(ctx, c->get_vertices_count(), c->get_edges_count(), (const int64_t *)c->get_vertex_pointers(),
                                        c->get_adjacent_ids(), &_dir.graph));
            check(vglb_graph_get_info(_dir.graph, &_dir.info));
            int64_t e = 0;
            if (_dir.info.tier_border[0] > 0)
                check(vglb_memcpy_d2h(ctx, &e, _dir.info.d_out_ptr + _dir.info.tier_border[0], sizeof(e)));
            _dir.hub_edges = e;
        }
        return _dir;
    }
    vglb::CsrView view(const Direction &_dir) const
    {
        vglb::CsrView v;
        v.ptr = _dir.info.d_out_ptr;
        v.adj = _dir.info.d_out_adj;
        v.V = _dir.info.vertices;
        for (int t = 0; t < vglb::kNumTiers; t++) v.tier_border[t] = _dir.info.tier_border[t];
        v.max_degree = _dir.info.max_degree;
        return v;
    }

    // The reference frontier object is the authority (the algorithm holds it and modifies it on the host); this brings the
    // backend's view of it up to date. A frontier this backend generated carries the backend's token in the reference's
    // collective_part_size (negative: every host-side modifier of modification.hpp:5-145 overwrites it with a value >= 0).
    FrontierState &refresh(FrontierVectorCSR &_frontier, Direction &_dir)
    {
        FrontierState &st = frontiers[&_frontier];
        if (st.frontier != NULL && st.graph != _dir.graph) // the frontier moved to the other direction
        {
            check(vglb_frontier_destroy(ctx, st.frontier));
            st.frontier = NULL;
        }
        if (st.frontier == NULL)
        {
            check(vglb_frontier_create_borrowed(ctx, _dir.graph, _frontier.ids, &st.frontier));
            st.graph = _dir.graph;
            st.token = 0;
        }
        if (_frontier.sparsity_type == ALL_ACTIVE_FRONTIER) return st;
        const bool ours = st.token != 0 && _frontier.collective_part_size == -st.token;
        if (!ours)
        {
            // made on the host: clear() then add_vertex / add_group_of_vertices wrote `size` ascending ids into ids[]
            check(vglb_frontier_clear(ctx, st.frontier));
            check(vglb_frontier_set_ids(ctx, st.frontier, _frontier.ids, _frontier.size, 1));
            check(vglb_frontier_get_info(ctx, st.frontier, &st.info));
            _frontier.neighbours_count = (int)st.info.neighbours;
            stamp(_frontier, st);
        }
        return st;
    }
    void stamp(FrontierVectorCSR &_frontier, FrontierState &_st)
    {
        _st.token = next_token++;
        if (next_token > 1000000000) next_token = 1;
        _frontier.collective_part_size = -_st.token;
    }

    // compute inner implementation
    template <typename ComputeOperation, typename GraphContainer, typename FrontierContainer>
    void compute_worker(GraphContainer &_graph, FrontierContainer &_frontier, ComputeOperation &&compute_op);

    // reduce inner implementation
    template <typename _T, typename ReduceOperation, typename GraphContainer, typename FrontierContainer>
    void reduce_worker(GraphContainer &_graph, FrontierContainer &_frontier, ReduceOperation &&reduce_op, REDUCE_TYPE _reduce_type,
                       _T &_result);

    // advance inner implementation: the VectorCSR container is this backend's format ...
    template <typename EdgeOperation, typename VertexPreprocessOperation, typename VertexPostprocessOperation,
              typename CollectiveEdgeOperation, typename CollectiveVertexPreprocessOperation,
              typename CollectiveVertexPostprocessOperation>
    void advance_worker(VectorCSRGraph &_graph, FrontierVectorCSR &_frontier, EdgeOperation &&edge_op,
                        VertexPreprocessOperation &&vertex_preprocess_op, VertexPostprocessOperation &&vertex_postprocess_op,
                        CollectiveEdgeOperation &&collective_edge_op,
                        CollectiveVertexPreprocessOperation &&collective_vertex_preprocess_op,
                        CollectiveVertexPostprocessOperation &&collective_vertex_postprocess_op, bool _inner_mpi_processing);

    // ... the other containers (EdgesList, CSR, CSR_VG) are outside the hot path this backend covers
    template <typename EdgeOperation, typename VertexPreprocessOperation, typename VertexPostprocessOperation,
              typename CollectiveEdgeOperation, typename CollectiveVertexPreprocessOperation,
              typename CollectiveVertexPostprocessOperation, typename GraphContainer, typename FrontierContainer>
    void advance_worker(GraphContainer &_graph, FrontierContainer &_frontier, EdgeOperation &&edge_op,
                        VertexPreprocessOperation &&vertex_preprocess_op, VertexPostprocessOperation &&vertex_postprocess_op,
                        CollectiveEdgeOperation &&collective_edge_op,
                        CollectiveVertexPreprocessOperation &&collective_vertex_preprocess_op,
                        CollectiveVertexPostprocessOperation &&collective_vertex_postprocess_op, bool _inner_mpi_processing)
    {
        throw "Error in GraphAbstractionsGPU (B200): only the VECTOR_CSR_GRAPH container is supported";
    }

public:
    // attaches graph-processing API to the specific graph
    GraphAbstractionsGPU(VGL_Graph &_graph, TraversalDirection _initial_traversal = SCATTER);
    ~GraphAbstractionsGPU();
    GraphAbstractionsGPU(const GraphAbstractionsGPU &) = delete;

    // generate new frontier implementation; public since it launches a kernel with a device lambda
    template <typename FilterCondition>
    void generate_new_frontier_worker(VectorCSRGraph &_graph, FrontierVectorCSR &_frontier, FilterCondition &&filter_cond);

    template <typename FilterCondition, typename GraphContainer, typename FrontierContainer>
    void generate_new_frontier_worker(GraphContainer &_graph, FrontierContainer &_frontier, FilterCondition &&filter_cond)
    {
        throw "Error in GraphAbstractionsGPU (B200): only the VECTOR_CSR_GRAPH container is supported";
    }

    // performs user-defined "edge_op" operation over all OUTGOING edges, neighbouring specified frontier
    template <typename EdgeOperation, typename VertexPreprocessOperation, typename VertexPostprocessOperation,
              typename CollectiveEdgeOperation, typename CollectiveVertexPreprocessOperation,
              typename CollectiveVertexPostprocessOperation>
    void scatter(VGL_Graph &_graph, VGL_Frontier &_frontier, EdgeOperation &&edge_op, VertexPreprocessOperation &&vertex_preprocess_op,
                 VertexPostprocessOperation &&vertex_postprocess_op, CollectiveEdgeOperation &&collective_edge_op,
                 CollectiveVertexPreprocessOperation &&collective_vertex_preprocess_op,
                 CollectiveVertexPostprocessOperation &&collective_vertex_postprocess_op)
    {
        this->common_scatter(_graph, _frontier, edge_op, vertex_preprocess_op, vertex_postprocess_op, collective_edge_op,
                             collective_vertex_preprocess_op, collective_vertex_postprocess_op, this);
    }

    template <typename EdgeOperation>
    void scatter(VGL_Graph &_graph, VGL_Frontier &_frontier, EdgeOperation &&edge_op)
    {
        vglb::NoVertexOp none; // no vertex ops: hub rows may be split over several CTAs
        scatter(_graph, _frontier, edge_op, none, none, edge_op, none, none);
    }

    // performs user-defined "edge_op" operation over all INCOMING edges, neighbouring specified frontier
    template <typename EdgeOperation, typename VertexPreprocessOperation, typename VertexPostprocessOperation,
              typename CollectiveEdgeOperation, typename CollectiveVertexPreprocessOperation,
              typename CollectiveVertexPostprocessOperation>
    void gather(VGL_Graph &_graph, VGL_Frontier &_frontier, EdgeOperation &&edge_op, VertexPreprocessOperation &&vertex_preprocess_op,
                VertexPostprocessOperation &&vertex_postprocess_op, CollectiveEdgeOperation &&collective_edge_op,
                CollectiveVertexPreprocessOperation &&collective_vertex_preprocess_op,
                CollectiveVertexPostprocessOperation &&collective_vertex_postprocess_op)
    {
        this->common_gather(_graph, _frontier, edge_op, vertex_preprocess_op, vertex_postprocess_op, collective_edge_op,
                            collective_vertex_preprocess_op, collective_vertex_postprocess_op, this);
    }

    template <typename EdgeOperation>
    void gather(VGL_Graph &_graph, VGL_Frontier &_frontier, EdgeOperation &&edge_op)
    {
        vglb::NoVertexOp none;
        gather(_graph, _frontier, edge_op, none, none, edge_op, none, none);
    }

    // performs user-defined "compute_op" operation for each element in the given frontier
    template <typename ComputeOperation>
    void compute(VGL_Graph &_graph, VGL_Frontier &_frontier, ComputeOperation &&compute_op)
    {
        this->common_compute(_graph, _frontier, compute_op, this);
    }

    // performs reduction using user-defined "reduce_op" operation for each element in the given frontier
    template <typename _T, typename ReduceOperation>
    _T reduce(VGL_Graph &_graph, VGL_Frontier &_frontier, ReduceOperation &&reduce_op, REDUCE_TYPE _reduce_type)
    {
        _T result = 0;
        this->common_reduce(_graph, _frontier, reduce_op, _reduce_type, result, this);
        return result;
    }

    // creates new frontier, which satisfy user-defined "cond" condition
    template <typename FilterCondition>
    void generate_new_frontier(VGL_Graph &_graph, VGL_Frontier &_frontier, FilterCondition &&filter_cond)
    {
        this->common_generate_new_frontier(_graph, _frontier, filter_cond, this);
    }

    void enable_safe_stores() { use_safe_stores = true; }
    void disable_safe_stores() { use_safe_stores = false; }

    friend class GraphAbstractions;
};

/////////////////////////////////////////////////////////////////////////////////////////////////////////////////////

namespace vglb_overlay
{
// one libvgl_b200 context per process and device: a backend object is created per algorithm call (bfs.hpp:60)
inline vglb_ctx *context()
{
    static vglb_ctx *contexts[64] = {NULL};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) throw "Error in GraphAbstractionsGPU (B200): no CUDA device (there is no CPU fallback)";
    if (contexts[dev & 63] == NULL && vglb_init(dev, &contexts[dev & 63]) != VGLB_OK) throw vglb_last_error();
    return contexts[dev & 63];
}
} // namespace vglb_overlay

GraphAbstractionsGPU::GraphAbstractionsGPU(VGL_Graph &_graph, TraversalDirection _initial_traversal)
{
    processed_graph_ptr = &_graph;
    current_traversal_direction = _initial_traversal;
    if (_graph.get_container_type() != VECTOR_CSR_GRAPH)
        throw "Error in GraphAbstractionsGPU (B200): only the VECTOR_CSR_GRAPH container is supported";
    ctx = vglb_overlay::context();
    stream = (cudaStream_t)vglb_stream(ctx);
    directions[0].container = (VectorCSRGraph *)_graph.get_outgoing_data();
    directions[1].container = (VectorCSRGraph *)_graph.get_incoming_data();
    filter_bitmap = NULL;
    reduce_buffer = NULL;
    next_token = 1;
    use_safe_stores = false;
    check(vglb_malloc(ctx, ((size_t)_graph.get_vertices_count() / 32 + 2) * 4, (void **)&filter_bitmap));
    check(vglb_malloc(ctx, 16, (void **)&reduce_buffer));
    max_blocks = 148 * 16;
    cudaDeviceProp prop;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) max_blocks = prop.multiProcessorCount * 16;
}

GraphAbstractionsGPU::~GraphAbstractionsGPU()
{
    cudaStreamSynchronize(stream);
    for (auto &kv : frontiers)
        if (kv.second.frontier) vglb_frontier_destroy(ctx, kv.second.frontier);
    for (int d = 0; d < 2; d++)
        if (directions[d].graph) vglb_graph_free(ctx, directions[d].graph); // borrowed arrays stay the reference's
    vglb_free(ctx, filter_bitmap);
    vglb_free(ctx, reduce_buffer);
}

/////////////////////////////////////////////////////////////////////////////////////////////////////////////////////

template <typename EdgeOperation, typename VertexPreprocessOperation, typename VertexPostprocessOperation,
          typename CollectiveEdgeOperation, typename CollectiveVertexPreprocessOperation, typename CollectiveVertexPostprocessOperation>
void GraphAbstractionsGPU::advance_worker(VectorCSRGraph &_graph, FrontierVectorCSR &_frontier, EdgeOperation &&edge_op,
                                          VertexPreprocessOperation &&vertex_preprocess_op,
                                          VertexPostprocessOperation &&vertex_postprocess_op,
                                          CollectiveEdgeOperation &&collective_edge_op,
                                          CollectiveVertexPreprocessOperation &&collective_vertex_preprocess_op,
                                          CollectiveVertexPostprocessOperation &&collective_vertex_postprocess_op,
                                          bool _inner_mpi_processing)
{
    Timer tm;
    tm.start();
    Direction &dir = direction_of(_graph);
    FrontierState &st = refresh(_frontier, dir);
    // EdgesArray layout [outgoing CSR | outgoing VE | incoming CSR | incoming VE] (compute_process_shift,
    // common/graph_abstractions.hpp:18-27): global_edge_pos = shift of the direction + CSR position
    const long long process_shift = (long long)compute_process_shift(current_traversal_direction, CSR_STORAGE);
    const vglb::CsrView g = view(dir);
    typedef typename std::decay<VertexPreprocessOperation>::type PreOp;
    typedef typename std::decay<VertexPostprocessOperation>::type PostOp;
    if (_frontier.sparsity_type == ALL_ACTIVE_FRONTIER)
    {
        const vglb::AllActivePlan plan = vglb::plan_all_active<PreOp, PostOp>(g, dir.info.edges, dir.hub_edges);
        if (plan.blocks > 0)
        {
            vglb::advance_all_active_kernel<<<(unsigned)plan.blocks, vglb::kAdvThreads, 0, stream>>>(g, plan, process_shift, edge_op,
                                                                                                  vertex_preprocess_op, vertex_postprocess_op);
            launch_check();
        }
    }
    else if (_frontier.size > 0)
    {
        vglb::SparseFrontierView F;
        F.ids = st.info.d_ids;
        F.n_hub = st.info.tier_size[0];
        F.n_mid = st.info.tier_size[1];
        F.n_small = st.info.tier_size[2];
        const long long grid = vglb::plan_sparse<PreOp, PostOp>(g, F, max_blocks);
        vglb::advance_sparse_kernel<<<(unsigned)grid, vglb::kAdvThreads, 0, stream>>>(g, F, process_shift, edge_op, vertex_preprocess_op,
                                                                                    vertex_postprocess_op);
        launch_check();
    }
    finish();
    tm.end();
    const size_t work = _frontier.sparsity_type == ALL_ACTIVE_FRONTIER ? (size_t)dir.info.edges : (size_t)st.info.neighbours;
    performance_stats.update_advance_stats(tm.get_time(), work * (INT_ELEMENTS_PER_EDGE) * sizeof(int), work);
}

/////////////////////////////////////////////////////////////////////////////////////////////////////////////////////

template <typename ComputeOperation, typename GraphContainer, typename FrontierContainer>
void GraphAbstractionsGPU::compute_worker(GraphContainer &_graph, FrontierContainer &_frontier, ComputeOperation &&compute_op)
{
    Direction &dir = direction_of((VectorCSRGraph &)_graph);
    const int64_t *ptr = dir.info.d_out_ptr;
    const int vertices_count = dir.info.vertices;
    if (_frontier.get_sparsity_type() == ALL_ACTIVE_FRONTIER)
    {
        const long long blocks = ((long long)vertices_count + 255) / 256;
        vglb::compute_all_active_kernel<<<(unsigned)(blocks < max_blocks ? blocks : max_blocks), 256, 0, stream>>>(ptr, vertices_count, compute_op);
        launch_check();
    }
    else if (_frontier.get_size() > 0)
    {
        FrontierState &st = refresh((FrontierVectorCSR &)_frontier, dir);
        const int n = _frontier.get_size();
        const long long blocks = ((long long)n + 255) / 256;
        vglb::compute_sparse_kernel<<<(unsigned)(blocks < max_blocks ? blocks : max_blocks), 256, 0, stream>>>(ptr, st.info.d_ids, n, compute_op);
        launch_check();
    }
    finish();
}

/////////////////////////////////////////////////////////////////////////////////////////////////////////////////////

template <typename _T, typename ReduceOperation, typename GraphContainer, typename FrontierContainer>
void GraphAbstractionsGPU::reduce_worker(GraphContainer &_graph, FrontierContainer &_frontier, ReduceOperation &&reduce_op,
                                         REDUCE_TYPE _reduce_type, _T &_result)
{
    Direction &dir = direction_of((VectorCSRGraph &)_graph);
    const int64_t *ptr = dir.info.d_out_ptr;
    const bool all = _frontier.get_sparsity_type() == ALL_ACTIVE_FRONTIER;
    const int n = all ? dir.info.vertices : _frontier.get_size();
    const int32_t *ids = NULL;
    if (!all && n > 0) ids = refresh((FrontierVectorCSR &)_frontier, dir).info.d_ids;
    const long long blocks = ((long long)n + 255) / 256;
    const unsigned grid = (unsigned)(blocks < max_blocks / 2 ? (blocks > 0 ? blocks : 1) : max_blocks / 2);
    if (_reduce_type == REDUCE_SUM)
    {
        // sums are accumulated in double on the device whatever _T is (block tree + one atomic per CTA); the reference GPU
        // path stages reduce_op values in a double[V] buffer and calls thrust::reduce (gpu/reduce.hpp:152-182)
        check(vglb_memset(ctx, reduce_buffer, 0, sizeof(double)));
        if (n > 0)
        {
            vglb::reduce_sum_kernel<double><<<grid, 256, 0, stream>>>(ptr, ids, n, reduce_buffer, reduce_op);
            launch_check();
        }
        double r = 0.0;
        check(vglb_memcpy_d2h(ctx, &r, reduce_buffer, sizeof(double)));
        _result = (_T)r;
    }
    else if (_reduce_type == REDUCE_MAX)
    {
        const int init = INT_MIN;
        check(vglb_memcpy_h2d(ctx, reduce_buffer, &init, sizeof(int)));
        if (n > 0)
        {
            vglb::reduce_max_kernel<<<grid, 256, 0, stream>>>(ptr, ids, n, (int *)reduce_buffer, reduce_op);
            launch_check();
        }
        int r = 0;
        check(vglb_memcpy_d2h(ctx, &r, reduce_buffer, sizeof(int)));
        _result = (_T)r;
    }
    else
        throw "Error in GraphAbstractionsGPU::reduce_worker: unsupported reduce type";
}

/////////////////////////////////////////////////////////////////////////////////////////////////////////////////////

template <typename FilterCondition>
void GraphAbstractionsGPU::generate_new_frontier_worker(VectorCSRGraph &_graph, FrontierVectorCSR &_frontier, FilterCondition &&filter_cond)
{
    Timer tm;
    tm.start();
    _frontier.set_direction(current_traversal_direction);
    Direction &dir = direction_of(_graph);
    FrontierState &st = frontiers[&_frontier];
    if (st.frontier != NULL && st.graph != dir.graph)
    {
        check(vglb_frontier_destroy(ctx, st.frontier));
        st.frontier = NULL;
    }
    if (st.frontier == NULL)
    {
        check(vglb_frontier_create_borrowed(ctx, dir.graph, _frontier.ids, &st.frontier));
        st.graph = dir.graph;
    }
    const int vertices_count = dir.info.vertices;
    // filter pass: the reference frontier's int flags[] and a bitmap word per warp (ballot) ...
    const long long blocks = ((long long)vertices_count + 255) / 256;
    vglb::gnf_bitmap_kernel<<<(unsigned)(blocks < max_blocks ? blocks : max_blocks), 256, 0, stream>>>(dir.info.d_out_ptr, vertices_count,
                                                                                                filter_bitmap, _frontier.flags, filter_cond);
    launch_check();
    // ... then ONE pass behind the C ABI: order-preserving compaction into the frontier's own ids[] (ascending ids keep the
    // degree tiers contiguous prefixes), size, neighbour count and tier populations in the same kernel
    check(vglb_gnf_from_bitmap(ctx, st.frontier, filter_bitmap));
    check(vglb_frontier_get_info(ctx, st.frontier, &st.info));
    _frontier.size = st.info.size;
    _frontier.neighbours_count = (int)st.info.neighbours; // (an int in the reference, base_frontier.h:14)
    _frontier.vector_engine_part_size = st.info.tier_size[0];
    _frontier.vector_core_part_size = st.info.tier_size[1];
    if (st.info.size == vertices_count)
    {
        _frontier.sparsity_type = ALL_ACTIVE_FRONTIER; // gpu/generate_new_frontier.hpp:131-135
        _frontier.vector_engine_part_type = _frontier.vector_core_part_type = _frontier.collective_part_type = ALL_ACTIVE_FRONTIER;
        _frontier.collective_part_size = st.info.tier_size[2];
        st.token = 0;
    }
    else
    {
        _frontier.sparsity_type = SPARSE_FRONTIER;
        _frontier.vector_engine_part_type = _frontier.vector_core_part_type = _frontier.collective_part_type = SPARSE_FRONTIER;
        stamp(_frontier, st);
    }
    tm.end();
    performance_stats.update_gnf_time(tm);
}

/////////////////////////////////////////////////////////////////////////////////////////////////////////////////////
