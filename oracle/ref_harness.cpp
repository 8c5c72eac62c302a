/*
 * ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
 *
 * Thin extern "C" wrapper around the UNMODIFIED reference (VectorGraphLibrary, multicore/OpenMP build). It includes
 * /root/reference/graph_library.h where it lies (no reference source is copied into this repo) and is compiled by
 * oracle/Makefile into oracle/_ref/libvgl_ref_<profile>.so. One shared object per app profile, because the
 * reference's tier thresholds are compile-time macros set per app before the include
 * (apps/bfs/bfs.cpp:3-7, apps/pr/pr.cpp:3-5, apps/sssp/sssp.cpp:3-12, apps/cc/cc.cpp:3-5).
 *
 * Used by: tests/ (pin the C oracle against the real reference), tests/golden/make_golden.py and
 * bench.py --impl reference / cpu_baseline (the reference's own CPU path, timed). Never by the product.
 *
 * Entry points mirror what the reference apps do (apps/bfs/bfs.cpp:15-62 etc.) with deterministic inputs:
 *   graph import  : EdgesContainer filled from arrays -> VGL_Graph(VECTOR_CSR_GRAPH).import  (vgl_graph.hpp:57-68)
 *   BFS           : BFS::vgl_top_down / BFS::seq_top_down                 (algorithms/bfs/bfs.hpp:56-86, seq_bfs.hpp:13-55)
 *   PageRank      : PageRank::vgl_page_rank(graph, ranks, 1e-4, iters)    (algorithms/pr/pr.hpp:7-148)
 *   SSSP          : seq_dijkstra | vgl_dijkstra(ALL_ACTIVE|PARTIAL_ACTIVE, PUSH)  (algorithms/sssp/*.hpp)
 *   CC            : ConnectedComponents::vgl_shiloach_vishkin             (algorithms/cc/shiloach_vishkin.hpp:7-88)
 * Results are reordered to ORIGINAL vertex numbering before they are returned.
 */
#if defined(REF_PROFILE_BFS)
#define INT_ELEMENTS_PER_EDGE 4.0
#define VECTOR_CORE_THRESHOLD_VALUE 2*VECTOR_LENGTH
#define COLLECTIVE_FRONTIER_TYPE_CHANGE_THRESHOLD 0.35
#elif defined(REF_PROFILE_PR)
#define INT_ELEMENTS_PER_EDGE 5.0
#define VECTOR_ENGINE_THRESHOLD_VALUE 2147483646
#define VECTOR_CORE_THRESHOLD_VALUE 5*VECTOR_LENGTH
#elif defined(REF_PROFILE_SSSP)
#define INT_ELEMENTS_PER_EDGE 5.0
#define VECTOR_ENGINE_THRESHOLD_VALUE VECTOR_LENGTH*MAX_SX_AURORA_THREADS*128
#define VECTOR_CORE_THRESHOLD_VALUE 5*VECTOR_LENGTH
#elif defined(REF_PROFILE_CC)
#define INT_ELEMENTS_PER_EDGE 5.0
#define VECTOR_ENGINE_THRESHOLD_VALUE VECTOR_LENGTH*MAX_SX_AURORA_THREADS*128
#define VECTOR_CORE_THRESHOLD_VALUE 5*VECTOR_LENGTH
#else
#error "define one of REF_PROFILE_{BFS,PR,SSSP,CC}"
#endif

#include "graph_library.h"
#include "vglb_synth.h"
#include <cstring>
#include <unistd.h>
#include <fcntl.h>

namespace {

struct RefGraph
{
    VGL_Graph *graph;
    long long edges;
    int vertices;
};

bool g_inited = false;

/* The reference chats on stdout (PR prints "ranks sum" every iteration, pr.hpp:135). Silence it during calls. */
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
    arr.reorder(ORIGINAL);
    memcpy(out, arr.get_ptr(), sizeof(T) * (size_t)n);
}

double mteps_to_seconds(double mteps, long long edges)
{
    return mteps > 0 ? (double)edges / (mteps * 1e6) : 0.0;
}

} // namespace

extern "C" {

int vglref_max_threads() { return omp_get_max_threads(); }

/* Build the reference graph (both directions) from an edge list. Needs OMP_NUM_THREADS >= 2. */
void *vglref_graph_create(int vertices, long long edges, const int *src, const int *dst)
{
    try
    {
        StdoutSilencer quiet;
        if (!g_inited)
        {
            char arg0[] = "vglref";
            char *argv[] = {arg0, NULL};
            VGL_RUNTIME::init_library(1, argv);
            g_inited = true;
        }
        EdgesContainer ec(vertices, edges);
        memcpy(ec.get_src_ids(), src, sizeof(int) * (size_t)edges);
        memcpy(ec.get_dst_ids(), dst, sizeof(int) * (size_t)edges);
        RefGraph *rg = new RefGraph;
        rg->graph = new VGL_Graph(VECTOR_CSR_GRAPH);
        rg->graph->import(ec);
        rg->edges = edges;
        rg->vertices = vertices;
        return rg;
    }
    catch (const char *e) { fprintf(stderr, "vglref: %s\n", e); return NULL; }
    catch (string e) { fprintf(stderr, "vglref: %s\n", e.c_str()); return NULL; }
}

void vglref_graph_destroy(void *h)
{
    RefGraph *rg = (RefGraph *)h;
    if (!rg) return;
    delete rg->graph;
    delete rg;
}

/* Dump the reference's degree-sorted CSR of one direction (0 = SCATTER/outgoing, 1 = GATHER/incoming) and the
 * ORIGINAL -> sorted id map, to pin our own graph builder against it. */
int vglref_graph_layout(void *h, int direction, long long *row_ptr, int *adj, int *orig_to_sorted,
                        int *ve_threshold_vertex, int *vc_threshold_vertex)
{
    RefGraph *rg = (RefGraph *)h;
    TraversalDirection dir = direction == 0 ? SCATTER : GATHER;
    VectorCSRGraph *c = (VectorCSRGraph *)rg->graph->get_direction_data(dir);
    memcpy(row_ptr, c->get_vertex_pointers(), sizeof(long long) * ((size_t)rg->vertices + 1));
    memcpy(adj, c->get_adjacent_ids(), sizeof(int) * (size_t)rg->edges);
    for (int v = 0; v < rg->vertices; v++)
        orig_to_sorted[v] = rg->graph->reorder(v, ORIGINAL, dir);
    *ve_threshold_vertex = c->get_vector_engine_threshold_vertex();
    *vc_threshold_vertex = c->get_vector_core_threshold_vertex();
    return 0;
}

/* mode 0: BFS::vgl_top_down (multicore operators); mode 1: BFS::seq_top_down. Returns seconds of the timed region. */
double vglref_bfs(void *h, int source_orig, int *levels_orig, int mode)
{
    RefGraph *rg = (RefGraph *)h;
    try
    {
        StdoutSilencer quiet;
        VGL_Graph &g = *rg->graph;
        VerticesArray<int> levels(g, SCATTER);
        int src = g.reorder(source_orig, ORIGINAL, SCATTER);
        double mteps = mode == 0 ? BFS::vgl_top_down(g, levels, src) : BFS::seq_top_down(g, levels, src);
        dump_original(levels, levels_orig, rg->vertices);
        return mteps_to_seconds(mteps, rg->edges);
    }
    catch (const char *e) { fprintf(stderr, "vglref: %s\n", e); return -1; }
}

/* PageRank::vgl_page_rank, multicore semantics. Returns seconds for `iters` iterations. */
double vglref_pagerank(void *h, int iters, float *ranks_orig)
{
    RefGraph *rg = (RefGraph *)h;
    try
    {
        StdoutSilencer quiet;
        VGL_Graph &g = *rg->graph;
        VerticesArray<float> ranks(g, SCATTER);
        double mteps = PageRank::vgl_page_rank(g, ranks, 1.0e-4f, iters); /* = iters * E / t / 1e6, pr.hpp:147 */
        dump_original(ranks, ranks_orig, rg->vertices);
        return mteps > 0 ? (double)iters * (double)rg->edges / (mteps * 1e6) : 0.0;
    }
    catch (const char *e) { fprintf(stderr, "vglref: %s\n", e); return -1; }
}

/* mode 0: seq_dijkstra (authoritative), 1: vgl_dijkstra ALL_ACTIVE PUSH, 2: PARTIAL_ACTIVE PUSH (racy; timing only).
 * Weights: vglb_edge_weight(orig_src, orig_dst, weight_seed) for every out-CSR position, mirrored to the VE copy and
 * the incoming direction the way EdgesArray::set_all_random does (vect_csr_edges_array.hpp:49-65). */
double vglref_sssp(void *h, unsigned long long weight_seed, int source_orig, float *dist_orig, int mode)
{
    RefGraph *rg = (RefGraph *)h;
    try
    {
        StdoutSilencer quiet;
        VGL_Graph &g = *rg->graph;
        VectorCSRGraph *out = (VectorCSRGraph *)g.get_outgoing_data();
        VectorCSRGraph *in = (VectorCSRGraph *)g.get_incoming_data();
        EdgesArray<float> weights(g);
        weights.set_all_constant(0);
        float *w = weights.get_ptr();
        const long long *ptr = out->get_vertex_pointers();
        const int *adj = out->get_adjacent_ids();
        const long long E = rg->edges;
        #pragma omp parallel for schedule(dynamic, 1024)
        for (int v = 0; v < rg->vertices; v++)
        {
            int ov = g.reorder(v, SCATTER, ORIGINAL);
            for (long long p = ptr[v]; p < ptr[v + 1]; p++)
                w[p] = vglb_edge_weight(ov, g.reorder(adj[p], SCATTER, ORIGINAL), weight_seed);
        }
        float *w_out_ve = w + E;
        float *w_in = w_out_ve + out->get_edges_count_in_ve();
        float *w_in_ve = w_in + E;
        out->get_ve_ptr()->copy_array_from_csr_to_ve(w_out_ve, w);
        g.copy_outgoing_to_incoming_edges(w, w_in);
        in->get_ve_ptr()->copy_array_from_csr_to_ve(w_in_ve, w_in);

        VerticesArray<float> dist(g, SCATTER);
        int src = g.reorder(source_orig, ORIGINAL, SCATTER);
        double mteps;
        if (mode == 0)
            mteps = ShortestPaths::seq_dijkstra(g, weights, dist, src);
        else if (mode == 1)
            mteps = ShortestPaths::vgl_dijkstra(g, weights, dist, src, ALL_ACTIVE, PUSH_TRAVERSAL);
        else
            mteps = ShortestPaths::vgl_dijkstra(g, weights, dist, src, PARTIAL_ACTIVE, PUSH_TRAVERSAL);
        dump_original(dist, dist_orig, rg->vertices);
        return mteps_to_seconds(mteps, rg->edges);
    }
    catch (const char *e) { fprintf(stderr, "vglref: %s\n", e); return -1; }
}

/* ConnectedComponents::vgl_shiloach_vishkin. labels_orig[v] = label value (a SCATTER-sorted id) of ORIGINAL vertex v. */
double vglref_cc(void *h, int *labels_orig)
{
    RefGraph *rg = (RefGraph *)h;
    try
    {
        StdoutSilencer quiet;
        VGL_Graph &g = *rg->graph;
        VerticesArray<int> comp(g, SCATTER);
        double mteps = ConnectedComponents::vgl_shiloach_vishkin(g, comp);
        dump_original(comp, labels_orig, rg->vertices);
        return mteps_to_seconds(mteps, rg->edges);
    }
    catch (const char *e) { fprintf(stderr, "vglref: %s\n", e); return -1; }
}

/* ---- file formats (pin libvgl_b200's readers / writers against the reference's own) ----
 * EdgesContainer::save_to_binary_file (vgl_runtime/graph_generation/edges_container.h:58-76) and
 * VGL_Graph::save_to_binary_file / load_from_binary_file (vgl_graph.hpp:109-161). */
int vglref_edges_save(const char *path, int vertices, long long edges, const int *src, const int *dst)
{
    try
    {
        StdoutSilencer quiet;
        EdgesContainer ec(vertices, edges);
        memcpy(ec.get_src_ids(), src, sizeof(int) * (size_t)edges);
        memcpy(ec.get_dst_ids(), dst, sizeof(int) * (size_t)edges);
        return ec.save_to_binary_file(path) ? 0 : 1;
    }
    catch (const char *e) { fprintf(stderr, "vglref: %s\n", e); return 2; }
}

/* import of an .el_container file, the `-import <file>` path of the apps (cmd_parser.hpp:64-68) */
void *vglref_graph_create_from_edges_file(const char *path)
{
    try
    {
        StdoutSilencer quiet;
        if (!g_inited)
        {
            char arg0[] = "vglref";
            char *argv[] = {arg0, NULL};
            VGL_RUNTIME::init_library(1, argv);
            g_inited = true;
        }
        EdgesContainer ec;
        if (!ec.load_from_binary_file(path)) return NULL;
        RefGraph *rg = new RefGraph;
        rg->graph = new VGL_Graph(VECTOR_CSR_GRAPH);
        rg->graph->import(ec);
        rg->edges = ec.get_edges_count();
        rg->vertices = ec.get_vertices_count();
        return rg;
    }
    catch (const char *e) { fprintf(stderr, "vglref: %s\n", e); return NULL; }
    catch (string e) { fprintf(stderr, "vglref: %s\n", e.c_str()); return NULL; }
}

int vglref_graph_save(void *h, const char *path)
{
    RefGraph *rg = (RefGraph *)h;
    StdoutSilencer quiet;
    return rg->graph->save_to_binary_file(path) ? 0 : 1;
}

void *vglref_graph_load(const char *path)
{
    try
    {
        StdoutSilencer quiet;
        if (!g_inited)
        {
            char arg0[] = "vglref";
            char *argv[] = {arg0, NULL};
            VGL_RUNTIME::init_library(1, argv);
            g_inited = true;
        }
        RefGraph *rg = new RefGraph;
        rg->graph = new VGL_Graph(VECTOR_CSR_GRAPH);
        if (!rg->graph->load_from_binary_file(path))
        {
            delete rg->graph;
            delete rg;
            return NULL;
        }
        rg->edges = rg->graph->get_edges_count();
        rg->vertices = rg->graph->get_vertices_count();
        return rg;
    }
    catch (const char *e) { fprintf(stderr, "vglref: %s\n", e); return NULL; }
    catch (string e) { fprintf(stderr, "vglref: %s\n", e.c_str()); return NULL; }
}

int vglref_graph_vertices(void *h) { return ((RefGraph *)h)->vertices; }
long long vglref_graph_edges(void *h) { return ((RefGraph *)h)->edges; }

} // extern "C"
