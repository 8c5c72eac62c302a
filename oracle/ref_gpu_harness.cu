/*
 * ref_gpu_harness.cu — TEST INFRASTRUCTURE, not product code.
 *
 * extern "C" wrapper around the reference's GPU build (`-D __USE_GPU__`): it includes the UNMODIFIED
 * /root/reference/graph_library.h where it lies and calls the reference's own algorithm entry points
 *     BFS::vgl_top_down                         (algorithms/bfs/bfs.hpp:56-86)
 *     PageRank::vgl_page_rank                   (algorithms/pr/gpu_pr.hpp:7-175, PULL and PUSH)
 *     ShortestPaths::vgl_dijkstra               (algorithms/sssp/gpu_shortest_paths.hpp:7-230: all-active push / pull, partial-active)
 *     ConnectedComponents::vgl_shiloach_vishkin (algorithms/cc/gpu_shiloach_vishkin.hpp:7-80)
 *     HITS::vgl_hits                            (algorithms/hits/hits.hpp:5-100 — not one of the four: pre-ops, both directions, reduce)
 * oracle/Makefile compiles THIS SAME FILE twice, differing only in the include path:
 *     oracle/_ref/libvgl_dropin.so   -I include/vgl_b200/overlay first: vgl_compute_api/gpu = the B200 backend of this repo
 *                                    (the drop-in proof: reference call sites, unchanged, on our operators)
 *     oracle/_ref/libvgl_refgpu.so   the reference's own CUDA backend recompiled for sm_100 (a second stated baseline)
 * Results are returned in ORIGINAL vertex numbering. Used by tests/test_gpu_dropin.py and bench.py's baseline legs.
 */
#define INT_ELEMENTS_PER_EDGE 5.0
#define VECTOR_ENGINE_THRESHOLD_VALUE 2147483646
#define VECTOR_CORE_THRESHOLD_VALUE 5 * VECTOR_LENGTH

#include "graph_library.h"
#include "vglb_synth.h"
#include <cstring>
#include <fcntl.h>
#include <unistd.h>

namespace
{

struct GpuRefGraph
{
    VGL_Graph *graph;
    long long edges;
    int vertices;
};

bool g_inited = false;

struct StdoutSilencer
{
    int saved;
    StdoutSilencer()
    {
        fflush(stdout);
        cout.flush();
        saved = dup(1);
        int devnull = open("/dev/null", O_WRONLY);
        dup2(devnull, 1);
        close(devnull);
    }
    ~StdoutSilencer()
    {
        fflush(stdout);
        cout.flush();
        dup2(saved, 1);
        close(saved);
    }
};

template <typename T>
void dump_original(VerticesArray<T> &arr, T *out, int n)
{
    cudaDeviceSynchronize();
    arr.move_to_host();
    arr.reorder(ORIGINAL);
    cudaDeviceSynchronize();
    memcpy(out, arr.get_ptr(), sizeof(T) * (size_t)n);
}

double mteps_to_seconds(double mteps, long long edges) { return mteps > 0 ? (double)edges / (mteps * 1e6) : 0.0; }

#define GUARDED(body)                                                             \
    try                                                                           \
    {                                                                             \
        body                                                                      \
    }                                                                             \
    catch (const char *e) { fprintf(stderr, "vglgpu: %s\n", e); return -1; }      \
    catch (string e) { fprintf(stderr, "vglgpu: %s\n", e.c_str()); return -1; }

} // namespace

extern "C" {

/* 1 when vgl_compute_api/gpu came from include/vgl_b200/overlay (the B200 backend), 0 for the reference's own */
int vglgpu_is_b200_backend()
{
#ifdef VGL_B200_H
    return 1;
#else
    return 0;
#endif
}

void *vglgpu_graph_create(int vertices, long long edges, const int *src, const int *dst)
{
    try
    {
        StdoutSilencer quiet;
        if (!g_inited)
        {
            char arg0[] = "vglgpu";
            char *argv[] = {arg0, NULL};
            VGL_RUNTIME::init_library(1, argv);
            g_inited = true;
        }
        EdgesContainer ec(vertices, edges);
        memcpy(ec.get_src_ids(), src, sizeof(int) * (size_t)edges);
        memcpy(ec.get_dst_ids(), dst, sizeof(int) * (size_t)edges);
        GpuRefGraph *rg = new GpuRefGraph;
        rg->graph = new VGL_Graph(VECTOR_CSR_GRAPH);
        rg->graph->import(ec);
        rg->edges = edges;
        rg->vertices = vertices;
        return rg;
    }
    catch (const char *e) { fprintf(stderr, "vglgpu: %s\n", e); return NULL; }
    catch (string e) { fprintf(stderr, "vglgpu: %s\n", e.c_str()); return NULL; }
}

void vglgpu_graph_destroy(void *h)
{
    GpuRefGraph *rg = (GpuRefGraph *)h;
    if (!rg) return;
    cudaDeviceSynchronize();
    delete rg->graph;
    delete rg;
}

/* BFS::vgl_top_down through VGL_GRAPH_ABSTRACTIONS; returns seconds of the reference's timed region */
double vglgpu_bfs(void *h, int source_orig, int *levels_orig)
{
    GpuRefGraph *rg = (GpuRefGraph *)h;
    GUARDED(
        StdoutSilencer quiet;
        VGL_Graph &g = *rg->graph;
        VerticesArray<int> levels(g, SCATTER);
        int src = g.reorder(source_orig, ORIGINAL, SCATTER);
        double mteps = BFS::vgl_top_down(g, levels, src);
        dump_original(levels, levels_orig, rg->vertices);
        return mteps_to_seconds(mteps, rg->edges);)
}

/* PageRank::vgl_page_rank, GPU semantics; traversal 0 = PULL (default of the reference), 1 = PUSH */
double vglgpu_pagerank(void *h, int iters, int traversal, float *ranks_orig)
{
    GpuRefGraph *rg = (GpuRefGraph *)h;
    GUARDED(
        StdoutSilencer quiet;
        VGL_Graph &g = *rg->graph;
        VerticesArray<float> ranks(g, SCATTER);
        double mteps = PageRank::vgl_page_rank(g, ranks, 1.0e-4f, iters, traversal == 0 ? PULL_TRAVERSAL : PUSH_TRAVERSAL);
        dump_original(ranks, ranks_orig, rg->vertices);
        return mteps > 0 ? (double)iters * (double)rg->edges / (mteps * 1e6) : 0.0;)
}

/* ShortestPaths::vgl_dijkstra; mode 0 = ALL_ACTIVE PUSH, 1 = ALL_ACTIVE PULL, 2 = PARTIAL_ACTIVE (push).
 * Weights: vglb_edge_weight(orig_src, orig_dst, seed) for every position of the outgoing AND the incoming CSR (the weight is a
 * function of the edge's end points, so both directions are filled from the hash). The reference would mirror out -> in with
 * VGL_Graph::copy_outgoing_to_incoming_edges, but under __USE_GPU__ that call moves BYTES, not elements
 * (vect_csr/reorder.hpp:72-74 passes the char* casts to the templated cuda_reorder_gather_copy) — a defect of the reference's GPU
 * build that is independent of the backend, so the harness does not go through it. */
double vglgpu_sssp(void *h, unsigned long long weight_seed, int source_orig, float *dist_orig, int mode)
{
    GpuRefGraph *rg = (GpuRefGraph *)h;
    GUARDED(
        StdoutSilencer quiet;
        VGL_Graph &g = *rg->graph;
        cudaDeviceSynchronize();
        g.move_to_host();
        VectorCSRGraph *out = (VectorCSRGraph *)g.get_outgoing_data();
        VectorCSRGraph *in = (VectorCSRGraph *)g.get_incoming_data();
        EdgesArray<float> weights(g);
        weights.set_all_constant(0);
        float *w = weights.get_ptr();
        const long long *ptr = out->get_vertex_pointers();
        const int *adj = out->get_adjacent_ids();
        const long long E = rg->edges;
        for (int v = 0; v < rg->vertices; v++)
        {
            int ov = g.reorder(v, SCATTER, ORIGINAL);
            for (long long p = ptr[v]; p < ptr[v + 1]; p++) w[p] = vglb_edge_weight(ov, g.reorder(adj[p], SCATTER, ORIGINAL), weight_seed);
        }
        float *w_out_ve = w + E;
        float *w_in = w_out_ve + out->get_edges_count_in_ve();
        float *w_in_ve = w_in + E;
        out->get_ve_ptr()->copy_array_from_csr_to_ve(w_out_ve, w);
        const long long *iptr = in->get_vertex_pointers();
        const int *iadj = in->get_adjacent_ids();
        for (int v = 0; v < rg->vertices; v++) /* row v of the incoming CSR (GATHER numbering) lists the sources of edges u -> v */
        {
            int ov = g.reorder(v, GATHER, ORIGINAL);
            for (long long p = iptr[v]; p < iptr[v + 1]; p++) w_in[p] = vglb_edge_weight(g.reorder(iadj[p], GATHER, ORIGINAL), ov, weight_seed);
        }
        in->get_ve_ptr()->copy_array_from_csr_to_ve(w_in_ve, w_in);

        VerticesArray<float> dist(g, SCATTER);
        int src = g.reorder(source_orig, ORIGINAL, SCATTER);
        double mteps;
        if (mode == 0)
            mteps = ShortestPaths::vgl_dijkstra(g, weights, dist, src, ALL_ACTIVE, PUSH_TRAVERSAL);
        else if (mode == 1)
            mteps = ShortestPaths::vgl_dijkstra(g, weights, dist, g.reorder(source_orig, ORIGINAL, GATHER), ALL_ACTIVE, PULL_TRAVERSAL);
        else
            mteps = ShortestPaths::vgl_dijkstra(g, weights, dist, src, PARTIAL_ACTIVE, PUSH_TRAVERSAL);
        dump_original(dist, dist_orig, rg->vertices);
        return mteps_to_seconds(mteps, rg->edges);)
}

double vglgpu_cc(void *h, int *labels_orig)
{
    GpuRefGraph *rg = (GpuRefGraph *)h;
    GUARDED(
        StdoutSilencer quiet;
        VGL_Graph &g = *rg->graph;
        VerticesArray<int> comp(g, SCATTER);
        double mteps = ConnectedComponents::vgl_shiloach_vishkin(g, comp);
        dump_original(comp, labels_orig, rg->vertices);
        return mteps_to_seconds(mteps, rg->edges);)
}

/* HITS::vgl_hits (gather with a pre-op, scatter with a pre-op, reduce<double>, compute; both directions in one run) and the
 * reference's own sequential check HITS::seq_hits on the same graph */
double vglgpu_hits(void *h, int steps, float *auth_orig, float *hub_orig, float *seq_auth_orig, float *seq_hub_orig)
{
    GpuRefGraph *rg = (GpuRefGraph *)h;
    GUARDED(
        StdoutSilencer quiet;
        VGL_Graph &g = *rg->graph;
        VerticesArray<float> auth(g, SCATTER);
        VerticesArray<float> hub(g, SCATTER);
        double mteps = HITS::vgl_hits(g, auth, hub, steps);
        dump_original(auth, auth_orig, rg->vertices);
        dump_original(hub, hub_orig, rg->vertices);
        if (seq_auth_orig && seq_hub_orig)
        {
            g.move_to_host();
            VerticesArray<float> sa(g, SCATTER);
            VerticesArray<float> sh(g, SCATTER);
            HITS::seq_hits(g, sa, sh, steps);
            dump_original(sa, seq_auth_orig, rg->vertices);
            dump_original(sh, seq_hub_orig, rg->vertices);
        }
        return mteps > 0 ? (double)steps * (double)rg->edges / (mteps * 1e6) : 0.0;)
}

/* frontier.add_group_of_vertices (modification.hpp:88-145) + compute + reduce through the backend: marks[v] = 7 for every listed
 * vertex, returns the reduce<int>(REDUCE_SUM) of the degrees of the group (or -1) */
long long vglgpu_group_mark(void *h, const int *ids_orig, int n, int *marks_orig)
{
    GpuRefGraph *rg = (GpuRefGraph *)h;
    GUARDED(
        StdoutSilencer quiet;
        VGL_Graph &g = *rg->graph;
        VGL_GRAPH_ABSTRACTIONS graph_API(g);
        VGL_FRONTIER frontier(g);
        VerticesArray<int> marks(g, SCATTER);
        graph_API.change_traversal_direction(SCATTER, marks, frontier);
        g.move_to_device();
        marks.move_to_device();
        frontier.move_to_device();
        frontier.set_all_active();
        auto zero = [marks] __VGL_COMPUTE_ARGS__ { marks[src_id] = 0; };
        graph_API.compute(g, frontier, zero);
        std::vector<int> ids(n);
        for (int i = 0; i < n; i++) ids[i] = g.reorder(ids_orig[i], ORIGINAL, SCATTER);
        frontier.clear();
        frontier.add_group_of_vertices(ids.data(), n);
        auto mark = [marks] __VGL_COMPUTE_ARGS__ { marks[src_id] = 7; };
        graph_API.compute(g, frontier, mark);
        auto degree = [] __VGL_REDUCE_INT_ARGS__ { return connections_count; };
        long long total = graph_API.reduce<int>(g, frontier, degree, REDUCE_SUM);
        dump_original(marks, marks_orig, rg->vertices);
        return total;)
}

} // extern "C"
