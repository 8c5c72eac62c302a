// vgl_b200/graph_abstractions_b200.cuh — the user-lambda operator API of VGL on top of libvgl_b200 (header-only, C++17,
// compiled by nvcc --extended-lambda in the CALLER's translation unit because device lambdas cannot cross a C ABI).
//
// It is the class a maintainer plugs into the reference's backend slot
//     #define VGL_GRAPH_ABSTRACTIONS GraphAbstractionsB200          (architecture_independent_api.h:33-43)
// following manuals/add_new_architecture.txt; the public surface mirrors vgl_compute_api/gpu/graph_abstractions_gpu.h
// :118-190 — scatter / gather in the 8-functor and 1-functor forms, compute, reduce<T>, generate_new_frontier,
// change_traversal_direction — with the reference's functor contracts (architecture_independent_api.h:17-30):
//     edge op    (int src_id, int dst_id, int local_edge_pos, long long global_edge_pos, int vector_index)
//     vertex op  (int src_id, int connections_count, int vector_index)
//     filter     (int src_id, int connections_count) -> IN_FRONTIER_FLAG / NOT_IN_FRONTIER_FLAG
//     reduce op  (int src_id, int connections_count, int vector_index) -> T
// Lambdas capture arrays BY VALUE (VGL_LAMBDA_CAP on the GPU, :65-69): VerticesArrayB200 / EdgesArrayB200 copy
// shallowly like the reference's (vertices_array.hpp:20-28) and index with a __host__ __device__ operator[].
//
// What differs from the reference GPU backend, by design:
//   * graph, arrays and frontiers live in cudaMalloc'ed HBM owned through the C ABI (no managed memory);
//   * operators are enqueued on ONE stream and do not end in cudaDeviceSynchronize; only reduce (returns a scalar) and
//     generate_new_frontier (returns the new size) wait for the device;
//   * one launch per advance covers all degree tiers (ids are degree-sorted: a tier is an id range or a prefix of the
//     ascending sparse id list), instead of 3-6 launches on 6 streams;
//   * the incoming CSR shares the SCATTER numbering, so change_traversal_direction never permutes user arrays.
// Errors follow the reference convention: `throw const char*` (common/advance.hpp:19-26).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "../vgl_b200.h"
#include "advance.cuh"

// lambda signature macros, GPU flavour of architecture_independent_api.h:17-30
#define __VGLB_COMPUTE_ARGS__ __device__(int src_id, int connections_count, int vector_index)
#define __VGLB_SCATTER_ARGS__ __device__(int src_id, int dst_id, int local_edge_pos, long long int global_edge_pos, int vector_index)
#define __VGLB_GATHER_ARGS__ __VGLB_SCATTER_ARGS__
#define __VGLB_ADVANCE_ARGS__ __VGLB_SCATTER_ARGS__
#define __VGLB_ADVANCE_PREPROCESS_ARGS__ __device__(int src_id, int connections_count, int vector_index)
#define __VGLB_ADVANCE_POSTPROCESS_ARGS__ __device__(int src_id, int connections_count, int vector_index)
#define __VGLB_GNF_ARGS__ __device__(int src_id, int connections_count)->int
#define __VGLB_REDUCE_INT_ARGS__ __device__(int src_id, int connections_count, int vector_index)->int
#define __VGLB_REDUCE_FLT_ARGS__ __device__(int src_id, int connections_count, int vector_index)->float
#define __VGLB_REDUCE_DBL_ARGS__ __device__(int src_id, int connections_count, int vector_index)->double

namespace vglb
{

constexpr int IN_FRONTIER_FLAG = 1;     // framework_types.h:164-165
constexpr int NOT_IN_FRONTIER_FLAG = 0;
enum REDUCE_TYPE { REDUCE_SUM = 0, REDUCE_MAX = 1 }; // framework_types.h:138-144
enum TraversalDirection { SCATTER = VGLB_SCATTER, GATHER = VGLB_GATHER, ORIGINAL = VGLB_ORIGINAL };
enum FrontierSparsityType { ALL_ACTIVE_FRONTIER = VGLB_FRONTIER_ALL_ACTIVE, DENSE_FRONTIER = VGLB_FRONTIER_DENSE, SPARSE_FRONTIER = VGLB_FRONTIER_SPARSE };

inline void check(int rc)
{
    if (rc != VGLB_OK) throw vglb_last_error(); // const char*, like the reference's throw "literal"
}

// VGL_RUNTIME::init_library (vgl_runtime.hpp:5-16)
class RuntimeB200
{
public:
    vglb_ctx *ctx = nullptr;
    explicit RuntimeB200(int device = 0) { check(vglb_init(device, &ctx)); }
    ~RuntimeB200() { vglb_finalize(ctx); }
    RuntimeB200(const RuntimeB200 &) = delete;
    cudaStream_t stream() const { return (cudaStream_t)vglb_stream(ctx); }
    void synchronize() const { check(vglb_synchronize(ctx)); }
};

// VGL_Graph with VECTOR_CSR_GRAPH containers (vgl_graph.hpp:57-68)
class GraphB200
{
public:
    RuntimeB200 &rt;
    vglb_graph *handle = nullptr;
    vglb_graph_info info;
    GraphB200(RuntimeB200 &_rt, int _vertices, long long _edges, const int *_src, const int *_dst, bool _on_device = false,
              int _flags = VGLB_GRAPH_WITH_INCOMING)
        : rt(_rt)
    {
        check(vglb_graph_from_edges(rt.ctx, _vertices, _edges, _src, _dst, _on_device ? 1 : 0, _flags, &handle));
        check(vglb_graph_get_info(handle, &info));
    }
    ~GraphB200() { vglb_graph_free(rt.ctx, handle); }
    GraphB200(const GraphB200 &) = delete;
    int get_vertices_count() const { return info.vertices; }
    long long get_edges_count() const { return info.edges; }
    // VGL_Graph::reorder(int, from, to) (vgl_graph/reorder.hpp:3-40): one id between ORIGINAL and SCATTER numbering
    int reorder(int _vertex, TraversalDirection _from, TraversalDirection _to) const
    {
        if (_from == _to) return _vertex;
        if (_vertex < 0 || _vertex >= info.vertices) throw "Error in GraphB200::reorder : vertex id out of range";
        const int32_t *map = (_from == ORIGINAL) ? info.d_orig_to_sorted : info.d_sorted_to_orig;
        int32_t out = 0;
        check(vglb_memcpy_d2h(rt.ctx, &out, map + _vertex, sizeof(out)));
        return out;
    }
};

// VerticesArray<T> (vertices_array.h:16-77): flat T[V] in SCATTER numbering; copies are shallow (captured by lambdas)
template <typename _T>
class VerticesArrayB200
{
    static_assert(sizeof(_T) == 4, "device vertex arrays hold 4-byte elements (reorder is a 4-byte permute)");
    _T *ptr = nullptr;
    int size_ = 0;
    bool is_copy = false;
    GraphB200 *graph = nullptr;

public:
    explicit VerticesArrayB200(GraphB200 &_graph) : size_(_graph.get_vertices_count()), graph(&_graph)
    {
        check(vglb_malloc(graph->rt.ctx, (size_t)size_ * sizeof(_T), (void **)&ptr));
    }
    __host__ __device__ VerticesArrayB200(const VerticesArrayB200 &_o) : ptr(_o.ptr), size_(_o.size_), is_copy(true), graph(_o.graph) {}
    __host__ __device__ ~VerticesArrayB200()
    {
#ifndef __CUDA_ARCH__
        if (!is_copy && ptr) vglb_free(graph->rt.ctx, ptr);
#endif
    }
    __host__ __device__ inline _T &operator[](int _idx) const { return ptr[_idx]; }
    __host__ __device__ inline _T *get_ptr() const { return ptr; }
    int size() const { return size_; }
    // result in ORIGINAL numbering on the host: VerticesArray::reorder(ORIGINAL) + move_to_host
    std::vector<_T> to_host_original() const
    {
        _T *tmp = nullptr;
        check(vglb_malloc(graph->rt.ctx, (size_t)size_ * sizeof(_T), (void **)&tmp));
        check(vglb_varray_reorder_u32(graph->rt.ctx, graph->handle, (const uint32_t *)ptr, (uint32_t *)tmp, VGLB_SCATTER, VGLB_ORIGINAL));
        std::vector<_T> h((size_t)size_);
        check(vglb_memcpy_d2h(graph->rt.ctx, h.data(), tmp, (size_t)size_ * sizeof(_T)));
        vglb_free(graph->rt.ctx, tmp);
        return h;
    }
};

// EdgesArray<T> (edges_array.h): one value per edge, indexed by global_edge_pos = [outgoing CSR | incoming CSR]
template <typename _T>
class EdgesArrayB200
{
    _T *ptr = nullptr;
    long long size_ = 0;
    bool is_copy = false;
    GraphB200 *graph = nullptr;

public:
    explicit EdgesArrayB200(GraphB200 &_graph) : size_(2 * _graph.get_edges_count()), graph(&_graph)
    {
        check(vglb_malloc(graph->rt.ctx, (size_t)(size_ > 0 ? size_ : 1) * sizeof(_T), (void **)&ptr));
    }
    __host__ __device__ EdgesArrayB200(const EdgesArrayB200 &_o) : ptr(_o.ptr), size_(_o.size_), is_copy(true), graph(_o.graph) {}
    __host__ __device__ ~EdgesArrayB200()
    {
#ifndef __CUDA_ARCH__
        if (!is_copy && ptr) vglb_free(graph->rt.ctx, ptr);
#endif
    }
    __host__ __device__ inline _T &operator[](long long _idx) const { return ptr[_idx]; }
    __host__ __device__ inline _T *get_ptr() const { return ptr; }
    // deterministic weights of the outgoing direction (EdgesArray::set_all_random twin), mirrored to the incoming segment the
    // way set_all_random does (vect_csr_edges_array.hpp:49-65)
    void set_synthetic_weights(unsigned long long _seed)
    {
        static_assert(sizeof(_T) == 4, "synthetic weights are fp32");
        check(vglb_earray_fill_synthetic_weights(graph->rt.ctx, graph->handle, _seed, (float *)ptr));
        if (graph->info.has_incoming) mirror_outgoing_to_incoming();
    }
    // VGL_Graph::copy_outgoing_to_incoming_edges (vgl_graph/reorder.hpp:229-233): values of the outgoing segment [0, E) ->
    // the same edges' positions in the incoming segment [E, 2E), which gather-direction operators index with global_edge_pos
    void mirror_outgoing_to_incoming()
    {
        static_assert(sizeof(_T) == 4, "the mirror moves 4-byte elements");
        check(vglb_earray_mirror_out_to_in_u32(graph->rt.ctx, graph->handle, (const uint32_t *)ptr, (uint32_t *)(ptr + graph->get_edges_count())));
    }
};

// VGL_Frontier (frontier.h:13-54)
class FrontierB200
{
public:
    GraphB200 &graph;
    vglb_frontier *handle = nullptr;
    explicit FrontierB200(GraphB200 &_graph) : graph(_graph) { check(vglb_frontier_create(graph.rt.ctx, graph.handle, &handle)); }
    ~FrontierB200() { vglb_frontier_destroy(graph.rt.ctx, handle); }
    FrontierB200(const FrontierB200 &) = delete;
    void set_all_active() { check(vglb_frontier_set_all_active(graph.rt.ctx, handle)); }
    void clear() { check(vglb_frontier_clear(graph.rt.ctx, handle)); }
    void add_vertex(int _v) { check(vglb_frontier_add_vertex(graph.rt.ctx, handle, _v)); }
    // add_group_of_vertices (modification.hpp:88-145): host ids, sorted ascending first like the reference (Sorter::sort)
    void add_group_of_vertices(int *_vertex_ids, int _number_of_vertices)
    {
        std::sort(_vertex_ids, _vertex_ids + _number_of_vertices);
        check(vglb_frontier_set_ids(graph.rt.ctx, handle, _vertex_ids, _number_of_vertices, 0));
    }
    vglb_frontier_info get_info() const
    {
        vglb_frontier_info fi;
        check(vglb_frontier_get_info(graph.rt.ctx, handle, &fi));
        return fi;
    }
    int size() const { return get_info().size; }
    long long get_neighbours_count() const { return get_info().neighbours; }
    FrontierSparsityType get_sparsity_type() const { return (FrontierSparsityType)get_info().sparsity_type; }
};

class GraphAbstractionsB200
{
    GraphB200 &graph;
    vglb_ctx *ctx;
    cudaStream_t stream;
    TraversalDirection current_traversal_direction;
    double *reduce_buffer = nullptr;  // device scalar (the reference keeps a double[V] buffer, graph_abstractions_gpu.hpp:10-17)
    uint32_t *filter_bitmap = nullptr; // output of the filter pass of generate_new_frontier
    int max_blocks;
    long long hub_edges_ = -1;

    CsrView view(bool incoming) const
    {
        CsrView v;
        v.ptr = incoming ? graph.info.d_in_ptr : graph.info.d_out_ptr;
        v.adj = incoming ? graph.info.d_in_adj : graph.info.d_out_adj;
        v.V = graph.info.vertices;
        for (int t = 0; t < kNumTiers; t++) v.tier_border[t] = graph.info.tier_border[t];
        v.max_degree = graph.info.max_degree;
        return v;
    }
    static void launch_check()
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) throw cudaGetErrorString(e); // SAFE_KERNEL_CALL (cuda_error_handling.h:15-27)
    }
    long long hub_edges() // edges of the rows with >= 4096 edges (row pointer at the first tier border), read once
    {
        if (hub_edges_ < 0)
        {
            int64_t e = 0;
            if (graph.info.tier_border[0] > 0) check(vglb_memcpy_d2h(ctx, &e, graph.info.d_out_ptr + graph.info.tier_border[0], sizeof(e)));
            hub_edges_ = e;
        }
        return hub_edges_;
    }

    template <typename EdgeOperation, typename VertexPreprocessOperation, typename VertexPostprocessOperation>
    void advance_worker(FrontierB200 &_frontier, bool _incoming, EdgeOperation &edge_op, VertexPreprocessOperation &vertex_preprocess_op,
                        VertexPostprocessOperation &vertex_postprocess_op)
    {
        const vglb_frontier_info fi = _frontier.get_info();
        const long long edge_shift = _incoming ? graph.info.edges : 0; // EdgesArray segments: [outgoing | incoming]
        if (_incoming)
        {
            if (!graph.info.has_incoming) throw "Error in GraphAbstractionsB200::gather : the graph was built without the incoming direction";
            // rows of the incoming CSR are not degree-sorted (it shares the SCATTER numbering): no id range says anything about
            // in-degrees, so every row is a "small / mid" row of a sparse list — warp batches of 32 rows walked flat, rows with
            // >= 32 in-edges by the whole warp
            const bool all = fi.sparsity_type == VGLB_FRONTIER_ALL_ACTIVE;
            const int n = all ? graph.info.vertices : fi.size;
            if (n == 0) return;
            const long long blocks = ((long long)n + kAdvThreads - 1) / kAdvThreads;
            advance_unsorted_kernel<<<(unsigned)(blocks < max_blocks ? blocks : max_blocks), kAdvThreads, 0, stream>>>(
                view(true), all ? nullptr : fi.d_ids, n, edge_shift, edge_op, vertex_preprocess_op, vertex_postprocess_op);
            launch_check();
            return;
        }
        const CsrView g = view(false);
        if (fi.sparsity_type == VGLB_FRONTIER_ALL_ACTIVE)
        {
            const AllActivePlan plan = plan_all_active<VertexPreprocessOperation, VertexPostprocessOperation>(g, graph.info.edges, hub_edges());
            if (plan.blocks == 0) return;
            advance_all_active_kernel<<<(unsigned)plan.blocks, kAdvThreads, 0, stream>>>(g, plan, edge_shift, edge_op, vertex_preprocess_op,
                                                                                       vertex_postprocess_op);
            launch_check();
            return;
        }
        // DENSE and SPARSE: the ascending id list; its tiers are contiguous prefixes because ids are degree-sorted
        if (fi.size == 0) return;
        SparseFrontierView F;
        F.ids = fi.d_ids;
        F.n_hub = fi.tier_size[0];
        F.n_mid = fi.tier_size[1];
        F.n_small = fi.tier_size[2];
        const long long grid = plan_sparse<VertexPreprocessOperation, VertexPostprocessOperation>(g, F, max_blocks);
        advance_sparse_kernel<<<(unsigned)grid, kAdvThreads, 0, stream>>>(g, F, edge_shift, edge_op, vertex_preprocess_op, vertex_postprocess_op);
        launch_check();
    }

public:
    // attaches the graph-processing API to a graph (graph_abstractions_gpu.h:118-120)
    explicit GraphAbstractionsB200(GraphB200 &_graph, TraversalDirection _initial_traversal = SCATTER)
        : graph(_graph), ctx(_graph.rt.ctx), stream(_graph.rt.stream()), current_traversal_direction(_initial_traversal)
    {
        check(vglb_malloc(ctx, 16, (void **)&reduce_buffer));
        check(vglb_malloc(ctx, ((size_t)graph.get_vertices_count() / 32 + 2) * 4, (void **)&filter_bitmap));
        cudaDeviceProp prop;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaGetDeviceProperties(&prop, dev);
        max_blocks = prop.multiProcessorCount * 16;
    }
    ~GraphAbstractionsB200()
    {
        vglb_free(ctx, reduce_buffer);
        vglb_free(ctx, filter_bitmap);
    }
    GraphAbstractionsB200(const GraphAbstractionsB200 &) = delete;

    // change_traversal_direction (common/graph_abstractions.hpp:80-92): nothing to permute on the device layout
    void change_traversal_direction(TraversalDirection _new_direction) { current_traversal_direction = _new_direction; }
    template <typename _T, typename... Arrays>
    void change_traversal_direction(TraversalDirection _new_direction, VerticesArrayB200<_T> &, Arrays &...) { current_traversal_direction = _new_direction; }
    void enable_safe_stores() {}
    void disable_safe_stores() {}

    // scatter: advance over the outgoing edges of the frontier (graph_abstractions.h:96-118)
    template <typename EdgeOperation, typename VertexPreprocessOperation, typename VertexPostprocessOperation,
              typename CollectiveEdgeOperation, typename CollectiveVertexPreprocessOperation, typename CollectiveVertexPostprocessOperation>
    void scatter(GraphB200 &_graph, FrontierB200 &_frontier, EdgeOperation &&edge_op, VertexPreprocessOperation &&vertex_preprocess_op,
                 VertexPostprocessOperation &&vertex_postprocess_op, CollectiveEdgeOperation &&collective_edge_op,
                 CollectiveVertexPreprocessOperation &&collective_vertex_preprocess_op,
                 CollectiveVertexPostprocessOperation &&collective_vertex_postprocess_op)
    {
        if (current_traversal_direction != SCATTER) throw "Error in GraphAbstractions::scatter : wrong traversal direction"; // common/advance.hpp:19-26
        if (&_graph != &graph) throw "Error in GraphAbstractionsB200::scatter : API object is attached to another graph";
        // the "collective" triple serves the low-degree region on SX-Aurora; like the reference's GPU backend
        // (gpu/advance_vect_csr.hpp:56-141) one triple runs everywhere
        (void)collective_edge_op; (void)collective_vertex_preprocess_op; (void)collective_vertex_postprocess_op;
        advance_worker(_frontier, false, edge_op, vertex_preprocess_op, vertex_postprocess_op);
    }
    template <typename EdgeOperation>
    void scatter(GraphB200 &_graph, FrontierB200 &_frontier, EdgeOperation &&edge_op)
    {
        NoVertexOp none;
        scatter(_graph, _frontier, edge_op, none, none, edge_op, none, none);
    }

    // gather: advance over the incoming edges of the frontier (graph_abstractions.h:120-142)
    template <typename EdgeOperation, typename VertexPreprocessOperation, typename VertexPostprocessOperation,
              typename CollectiveEdgeOperation, typename CollectiveVertexPreprocessOperation, typename CollectiveVertexPostprocessOperation>
    void gather(GraphB200 &_graph, FrontierB200 &_frontier, EdgeOperation &&edge_op, VertexPreprocessOperation &&vertex_preprocess_op,
                VertexPostprocessOperation &&vertex_postprocess_op, CollectiveEdgeOperation &&collective_edge_op,
                CollectiveVertexPreprocessOperation &&collective_vertex_preprocess_op,
                CollectiveVertexPostprocessOperation &&collective_vertex_postprocess_op)
    {
        if (current_traversal_direction != GATHER) throw "Error in GraphAbstractions::gather : wrong traversal direction";
        if (&_graph != &graph) throw "Error in GraphAbstractionsB200::gather : API object is attached to another graph";
        (void)collective_edge_op; (void)collective_vertex_preprocess_op; (void)collective_vertex_postprocess_op;
        advance_worker(_frontier, true, edge_op, vertex_preprocess_op, vertex_postprocess_op);
    }
    template <typename EdgeOperation>
    void gather(GraphB200 &_graph, FrontierB200 &_frontier, EdgeOperation &&edge_op)
    {
        NoVertexOp none;
        gather(_graph, _frontier, edge_op, none, none, edge_op, none, none);
    }

    // compute: vertex map over the frontier (common/compute.hpp:62-85)
    template <typename ComputeOperation>
    void compute(GraphB200 &_graph, FrontierB200 &_frontier, ComputeOperation &&compute_op)
    {
        const vglb_frontier_info fi = _frontier.get_info();
        const int64_t *ptr = current_traversal_direction == GATHER && _graph.info.has_incoming ? _graph.info.d_in_ptr : _graph.info.d_out_ptr;
        const bool all = fi.sparsity_type == VGLB_FRONTIER_ALL_ACTIVE;
        const int n = all ? _graph.get_vertices_count() : fi.size;
        if (n == 0) return;
        const long long blocks = ((long long)n + 255) / 256;
        const unsigned grid = (unsigned)(blocks < max_blocks ? blocks : max_blocks);
        if (all) compute_all_active_kernel<<<grid, 256, 0, stream>>>(ptr, n, compute_op);
        else compute_sparse_kernel<<<grid, 256, 0, stream>>>(ptr, fi.d_ids, n, compute_op);
        launch_check();
    }

    // reduce: sum / max of reduce_op over the frontier, returned to the host (common/reduce.hpp:4-67)
    template <typename _T, typename ReduceOperation>
    _T reduce(GraphB200 &_graph, FrontierB200 &_frontier, ReduceOperation &&reduce_op, REDUCE_TYPE _reduce_type)
    {
        const vglb_frontier_info fi = _frontier.get_info();
        const int64_t *ptr = _graph.info.d_out_ptr;
        const bool all = fi.sparsity_type == VGLB_FRONTIER_ALL_ACTIVE;
        const int n = all ? _graph.get_vertices_count() : fi.size;
        const int32_t *ids = all ? nullptr : fi.d_ids;
        const long long blocks = ((long long)n + 255) / 256;
        const unsigned grid = (unsigned)(blocks < max_blocks / 2 ? (blocks > 0 ? blocks : 1) : max_blocks / 2);
        if (_reduce_type == REDUCE_SUM)
        {
            // sums are accumulated in double on the device (the reference GPU path does the same, gpu/reduce.hpp:152-182)
            check(vglb_memset(ctx, reduce_buffer, 0, sizeof(double)));
            if (n > 0)
            {
                reduce_sum_kernel<double><<<grid, 256, 0, stream>>>(ptr, ids, n, reduce_buffer, reduce_op);
                launch_check();
            }
            double r = 0.0;
            check(vglb_memcpy_d2h(ctx, &r, reduce_buffer, sizeof(double)));
            return (_T)r;
        }
        if (_reduce_type == REDUCE_MAX)
        {
            const int init = INT_MIN;
            check(vglb_memcpy_h2d(ctx, reduce_buffer, &init, sizeof(int)));
            if (n > 0)
            {
                reduce_max_kernel<<<grid, 256, 0, stream>>>(ptr, ids, n, (int *)reduce_buffer, reduce_op);
                launch_check();
            }
            int r = 0;
            check(vglb_memcpy_d2h(ctx, &r, reduce_buffer, sizeof(int)));
            return (_T)r;
        }
        throw "Error in GraphAbstractionsB200::reduce : unsupported reduce type";
    }

    // generate_new_frontier: filter -> bitmap (ballot per warp) -> one-pass compaction behind the C ABI
    // (common/generate_new_frontier.hpp:4-43)
    template <typename FilterCondition>
    void generate_new_frontier(GraphB200 &_graph, FrontierB200 &_frontier, FilterCondition &&filter_cond)
    {
        const int V = _graph.get_vertices_count();
        const long long blocks = ((long long)V + 255) / 256;
        gnf_bitmap_kernel<<<(unsigned)(blocks < max_blocks ? blocks : max_blocks), 256, 0, stream>>>(_graph.info.d_out_ptr, V, filter_bitmap,
                                                                                              (int32_t *)nullptr, filter_cond);
        launch_check();
        check(vglb_gnf_from_bitmap(ctx, _frontier.handle, filter_bitmap));
    }
};

} // namespace vglb
