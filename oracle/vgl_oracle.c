/*
 * vgl_oracle.c — TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C CPU restatement of the reference's frontier-processing hot path (VectorGraphLibrary, multicore build):
 * VectCSR import, BFS, PageRank (multicore semantics), SSSP and CC. Each function cites the reference file:line it
 * follows. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (vectorgraphlibrary_b200) never does.
 *
 * Pinning: the reference ships no golden vectors, so this oracle is pinned against outputs of the
 * reference itself, run in the build container through oracle/_ref/libvgl_ref_*.so (tests/test_oracle_vs_reference.py)
 * and against the fixtures those runs produced (tests/golden/, generator script committed beside them).
 *
 * Unlike the reference (32-bit edge counters) every edge counter here is 64-bit, so this oracle also
 * serves the configs the reference cannot run (E >= 2^31).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <math.h>
#include <omp.h>
#include "vglb_synth.h"

#define UNVISITED_VERTEX (-1)   /* algorithms/bfs/change_state/change_state.h:21-23 */
#define FIRST_LEVEL_VERTEX 1

/* ------------------------------------------------------------------------------------------------------------ */
/* Synthetic inputs (definitions live in include/vglb_synth.h; this is only the loop over edge indices).          */
/* ------------------------------------------------------------------------------------------------------------ */

void vglo_generate_edges(int kind, int scale, int64_t edges, uint64_t seed, int a, int b, int c,
                         int32_t *src, int32_t *dst)
{
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < edges; i++)
        vglb_gen_edge(kind, scale, seed, (uint64_t)i, a, b, c, &src[i], &dst[i]);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* VectCSR import — vgl_datastructures/graphs/undirected_containers/vect_csr/import.hpp:257-337                     */
/*   degree histogram (:5-57), stable sort by degree DESC (:61-99 + sorter.h:55-92 std::stable_sort), renumber       */
/*   (edges_container.h:163-213), stable sort edges by new src (edges_container.h:101-161), row pointers from src    */
/*   run boundaries (import.hpp:103-153). Stable sorts are restated as counting sorts (same permutation).            */
/*   Outputs: row_ptr[V+1], adj[E] in sorted numbering; fwd[orig]=sorted, bwd[sorted]=orig;                          */
/*   edge_order[p] = index in the input edge list of CSR position p (may be NULL).                                   */
/* ------------------------------------------------------------------------------------------------------------ */

int vglo_build_vect_csr(int32_t V, int64_t E, const int32_t *src, const int32_t *dst,
                        int64_t *row_ptr, int32_t *adj, int32_t *fwd, int32_t *bwd, int64_t *edge_order)
{
    int32_t *deg = (int32_t *)calloc((size_t)V, sizeof(int32_t));
    if (!deg) return -1;
    int32_t maxdeg = 0;
    for (int64_t i = 0; i < E; i++) deg[src[i]]++;
    for (int32_t v = 0; v < V; v++) if (deg[v] > maxdeg) maxdeg = deg[v];

    /* stable sort of ids by degree descending == counting sort over (maxdeg - deg), ids ascending inside a bucket */
    int64_t *bucket = (int64_t *)calloc((size_t)maxdeg + 2, sizeof(int64_t));
    if (!bucket) { free(deg); return -1; }
    for (int32_t v = 0; v < V; v++) bucket[(maxdeg - deg[v]) + 1]++;
    for (int32_t k = 0; k <= maxdeg; k++) bucket[k + 1] += bucket[k];
    for (int32_t v = 0; v < V; v++)
    {
        int64_t pos = bucket[maxdeg - deg[v]]++;
        bwd[pos] = v;          /* backward(i) = original */
        fwd[v] = (int32_t)pos; /* forward(orig) = sorted */
    }
    free(bucket);

    /* row pointers in sorted numbering */
    row_ptr[0] = 0;
    for (int32_t s = 0; s < V; s++) row_ptr[s + 1] = row_ptr[s] + deg[bwd[s]];
    free(deg);

    /* stable sort edges by new src: counting sort keeps input order inside a row */
    int64_t *cursor = (int64_t *)malloc(((size_t)V + 1) * sizeof(int64_t));
    if (!cursor) return -1;
    memcpy(cursor, row_ptr, ((size_t)V + 1) * sizeof(int64_t));
    for (int64_t i = 0; i < E; i++)
    {
        int64_t p = cursor[fwd[src[i]]]++;
        adj[p] = fwd[dst[i]];
        if (edge_order) edge_order[p] = i;
    }
    free(cursor);
    return 0;
}

/* Tier borders — vect_csr/nec_api.hpp:5-50: index after the last vertex whose degree is >= threshold. */
void vglo_estimate_thresholds(int32_t V, const int64_t *row_ptr, int32_t ve_threshold_value, int32_t vc_threshold_value,
                              int32_t *ve_threshold_vertex, int32_t *vc_threshold_vertex)
{
    int32_t ve = 0, vc = 0;
    for (int32_t v = 0; v < V; v++)
    {
        int64_t d = row_ptr[v + 1] - row_ptr[v];
        int64_t dn = (v == V - 1) ? 0 : row_ptr[v + 2] - row_ptr[v + 1];
        if (d >= ve_threshold_value && dn < ve_threshold_value) ve = v + 1;
        else if (d >= vc_threshold_value && dn < vc_threshold_value) vc = v + 1;
    }
    *ve_threshold_vertex = ve;
    *vc_threshold_vertex = vc;
}

/* Vertex-array permutation — vgl_graph/reorder.hpp:3-170: out_orig[v] = in_sorted[fwd[v]] (4-byte elements). */
void vglo_reorder_to_original_u32(int32_t V, const int32_t *fwd, const uint32_t *in_sorted, uint32_t *out_orig)
{
    for (int32_t v = 0; v < V; v++) out_orig[v] = in_sorted[fwd[v]];
}

/* Edge weights for every CSR position (see vglb_edge_weight; harness twin: oracle/ref_harness.cpp vglref_sssp). */
void vglo_edge_weights(int32_t V, const int64_t *row_ptr, const int32_t *adj, const int32_t *bwd, uint64_t seed, float *w)
{
    #pragma omp parallel for schedule(dynamic, 1024)
    for (int32_t v = 0; v < V; v++)
        for (int64_t p = row_ptr[v]; p < row_ptr[v + 1]; p++)
            w[p] = vglb_edge_weight(bwd[v], bwd[adj[p]], seed);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* BFS — algorithms/bfs/bfs.hpp:5-51 (level-synchronous top-down) == algorithms/bfs/seq_bfs.hpp:13-55 (queue);     */
/* levels are unique: source = 1, unreachable = -1. Ids are sorted (SCATTER) ids.                                  */
/* ------------------------------------------------------------------------------------------------------------ */

int vglo_bfs(int32_t V, const int64_t *row_ptr, const int32_t *adj, int32_t source, int32_t *levels,
             int64_t *edges_inspected)
{
    int32_t *queue = (int32_t *)malloc((size_t)V * sizeof(int32_t));
    if (!queue) return -1;
    for (int32_t v = 0; v < V; v++) levels[v] = UNVISITED_VERTEX;
    int64_t head = 0, tail = 0, inspected = 0;
    levels[source] = FIRST_LEVEL_VERTEX;
    queue[tail++] = source;
    while (head < tail)
    {
        int32_t s = queue[head++];
        for (int64_t p = row_ptr[s]; p < row_ptr[s + 1]; p++)
        {
            int32_t d = adj[p];
            inspected++;
            if (levels[d] == UNVISITED_VERTEX)
            {
                levels[d] = levels[s] + 1;
                queue[tail++] = d;
            }
        }
    }
    free(queue);
    if (edges_inspected) *edges_inspected = inspected;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* PageRank, MULTICORE semantics — algorithms/pr/pr.hpp:7-148:                             */
/*   r'[u] = k + d*( sum_{(u->v) in E, v != u} r[v]*inv[v] + D ),  inv[v] = 1/indeg_noloops(v) (0 if none),          */
/*   D = sum_{v: indeg_noloops(v)==0} r[v]/V,  r0 = 1/V,  d = 0.85f,  k = (1-d)/V, exactly `iters` sweeps.           */
/* fp32 restatement: row sums sequential in CSR order (pr.hpp:109-116, one thread per row because                    */
/* apps/pr/pr.cpp:4 empties the ve tier); dangling sum = OpenMP static-chunk fp32 reduction over `threads` chunks     */
/* (pr.hpp:94-103 -> multicore/reduce.hpp:18-31), chunks combined in thread order.                                   */
/* indeg_noloops is indexed by SCATTER id (the reference computes it in GATHER order then reorders, pr.hpp:28-75).    */
/* ------------------------------------------------------------------------------------------------------------ */

void vglo_indegree_noloops(int32_t V, const int64_t *row_ptr, const int32_t *adj, int32_t *indeg_noloops)
{
    memset(indeg_noloops, 0, (size_t)V * sizeof(int32_t));
    for (int32_t u = 0; u < V; u++)
        for (int64_t p = row_ptr[u]; p < row_ptr[u + 1]; p++)
            if (adj[p] != u) indeg_noloops[adj[p]]++;
}

int vglo_pagerank_f32(int32_t V, const int64_t *row_ptr, const int32_t *adj, const int32_t *indeg_noloops,
                      int iters, int threads, float *ranks)
{
    float *old = (float *)malloc((size_t)V * sizeof(float));
    float *inv = (float *)malloc((size_t)V * sizeof(float));
    if (!old || !inv) return -1;
    const float d = 0.85f;
    const float k = (float)((1.0 - d) / ((float)V));                       /* pr.hpp:37-38 */
    for (int32_t v = 0; v < V; v++)
    {
        ranks[v] = (float)(1.0 / V);                                       /* pr.hpp:42 */
        inv[v] = (float)(1.0 / indeg_noloops[v]);                          /* pr.hpp:68-71 */
        if (indeg_noloops[v] == 0) inv[v] = 0;
    }
    if (threads < 1) threads = 1;
    for (int it = 0; it < iters; it++)
    {
        for (int32_t v = 0; v < V; v++) { old[v] = ranks[v]; ranks[v] = 0; }  /* pr.hpp:85-90 */
        /* dangling: schedule(static) chunks like libgomp: q = V/T, first V%T chunks one longer */
        float dangling = 0.0f;
        {
            int32_t q = V / threads, t = V % threads;
            for (int tid = 0; tid < threads; tid++)
            {
                int32_t len = q + (tid < t ? 1 : 0);
                int32_t start = tid < t ? tid * (q + 1) : tid * q + t;
                float part = 0.0f;
                for (int32_t v = start; v < start + len; v++)
                {
                    float val = 0.0f;
                    if (indeg_noloops[v] == 0) val = old[v] / V;            /* pr.hpp:94-101 */
                    part += val;
                }
                dangling += part;
            }
        }
        #pragma omp parallel for schedule(dynamic, 256)
        for (int32_t u = 0; u < V; u++)
        {
            float acc = 0.0f;
            for (int64_t p = row_ptr[u]; p < row_ptr[u + 1]; p++)
            {
                int32_t v = adj[p];
                if (u != v) acc += old[v] * inv[v];                         /* pr.hpp:109-116 */
            }
            ranks[u] = k + d * (acc + dangling);                            /* pr.hpp:118-121 */
        }
    }
    free(old);
    free(inv);
    return 0;
}

/* Attribution model of the B200 kernel's "reference-order" mode (vglb_pagerank_ex, dangling_mode = 1): fp32 state and
 * the reference's T-chunk sequential fp32 dangling sum (above), but every row sum accumulated in fp64 and rounded once
 * (what a pairwise / tree fp32 sum approaches). Separates the two sources of the reference's fp32 drift:
 * rel_l1(this, vglo_pagerank_f32) is the row-sum-order share, rel_l1(this, vglo_pagerank_f64) the dangling-sum share. */
int vglo_pagerank_f32_tree_rows(int32_t V, const int64_t *row_ptr, const int32_t *adj, const int32_t *indeg_noloops,
                                int iters, int threads, float *ranks)
{
    float *old = (float *)malloc((size_t)V * sizeof(float));
    float *inv = (float *)malloc((size_t)V * sizeof(float));
    if (!old || !inv) return -1;
    const float d = 0.85f;
    const float k = (float)((1.0 - d) / ((float)V));
    for (int32_t v = 0; v < V; v++)
    {
        ranks[v] = (float)(1.0 / V);
        inv[v] = (float)(1.0 / indeg_noloops[v]);
        if (indeg_noloops[v] == 0) inv[v] = 0;
    }
    if (threads < 1) threads = 1;
    for (int it = 0; it < iters; it++)
    {
        for (int32_t v = 0; v < V; v++) { old[v] = ranks[v]; ranks[v] = 0; }
        float dangling = 0.0f;
        {
            int32_t q = V / threads, t = V % threads;
            for (int tid = 0; tid < threads; tid++)
            {
                int32_t len = q + (tid < t ? 1 : 0);
                int32_t start = tid < t ? tid * (q + 1) : tid * q + t;
                float part = 0.0f;
                for (int32_t v = start; v < start + len; v++)
                    if (indeg_noloops[v] == 0) part += old[v] / V;
                dangling += part;
            }
        }
        #pragma omp parallel for schedule(dynamic, 256)
        for (int32_t u = 0; u < V; u++)
        {
            double acc = 0.0;
            for (int64_t p = row_ptr[u]; p < row_ptr[u + 1]; p++)
            {
                int32_t v = adj[p];
                if (u != v) acc += (double)(old[v] * inv[v]);
            }
            ranks[u] = k + d * ((float)acc + dangling);
        }
    }
    free(old);
    free(inv);
    return 0;
}

/* Same recurrence evaluated in fp64 with fp32 constants widened: the attribution reference of SURVEY §8c. */
int vglo_pagerank_f64(int32_t V, const int64_t *row_ptr, const int32_t *adj, const int32_t *indeg_noloops,
                      int iters, double *ranks)
{
    double *old = (double *)malloc((size_t)V * sizeof(double));
    if (!old) return -1;
    const double d = (double)0.85f;
    const double k = (1.0 - d) / (double)V;
    for (int32_t v = 0; v < V; v++) ranks[v] = 1.0 / V;
    for (int it = 0; it < iters; it++)
    {
        double dangling = 0.0;
        for (int32_t v = 0; v < V; v++)
        {
            old[v] = ranks[v];
            if (indeg_noloops[v] == 0) dangling += old[v] / V;
        }
        #pragma omp parallel for schedule(dynamic, 256)
        for (int32_t u = 0; u < V; u++)
        {
            double acc = 0.0;
            for (int64_t p = row_ptr[u]; p < row_ptr[u + 1]; p++)
            {
                int32_t v = adj[p];
                if (u != v) acc += old[v] / (double)indeg_noloops[v];
            }
            ranks[u] = k + d * (acc + dangling);
        }
    }
    free(old);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* SSSP — algorithms/sssp/seq_shortest_paths.hpp:8-68 (binary-heap Dijkstra, fp32, inf = FLT_MAX - 100 == FLT_MAX). */
/* The frontier Bellman-Ford of shortest_paths.hpp:7-78 converges to the same unique min-plus fixed point.         */
/* ------------------------------------------------------------------------------------------------------------ */

typedef struct { float key; int32_t v; } heap_item;

static void heap_push(heap_item **h, int64_t *n, int64_t *cap, float key, int32_t v)
{
    if (*n == *cap) { *cap = *cap ? *cap * 2 : 1024; *h = (heap_item *)realloc(*h, (size_t)*cap * sizeof(heap_item)); }
    int64_t i = (*n)++;
    heap_item *a = *h;
    while (i > 0)
    {
        int64_t p = (i - 1) / 2;
        if (a[p].key < key || (a[p].key == key && a[p].v <= v)) break;
        a[i] = a[p];
        i = p;
    }
    a[i].key = key; a[i].v = v;
}

static heap_item heap_pop(heap_item *a, int64_t *n)
{
    heap_item top = a[0], last = a[--(*n)];
    int64_t i = 0;
    for (;;)
    {
        int64_t c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && (a[c + 1].key < a[c].key || (a[c + 1].key == a[c].key && a[c + 1].v < a[c].v))) c++;
        if (last.key < a[c].key || (last.key == a[c].key && last.v <= a[c].v)) break;
        a[i] = a[c];
        i = c;
    }
    a[i] = last;
    return top;
}

int vglo_sssp(int32_t V, const int64_t *row_ptr, const int32_t *adj, const float *w, int32_t source, float *dist,
              int64_t *edges_relaxed)
{
    const float inf_val = FLT_MAX - 100;                                   /* seq_shortest_paths.hpp:24 */
    for (int32_t v = 0; v < V; v++) dist[v] = inf_val;
    heap_item *heap = NULL;
    int64_t n = 0, cap = 0, relaxed = 0;
    heap_push(&heap, &n, &cap, 0.0f, source);
    dist[source] = 0;
    while (n > 0)
    {
        heap_item it = heap_pop(heap, &n);
        int32_t s = it.v;
        if (it.key > dist[s]) continue; /* stale entry: relaxing from it cannot improve anything (monotone fp32 +) */
        for (int64_t p = row_ptr[s]; p < row_ptr[s + 1]; p++)
        {
            int32_t t = adj[p];
            float cand = dist[s] + w[p];
            relaxed++;
            if (dist[t] > cand)
            {
                dist[t] = cand;
                heap_push(&heap, &n, &cap, cand, t);
            }
        }
    }
    free(heap);
    if (edges_relaxed) *edges_relaxed = relaxed;
    return 0;
}

/* Frontier Bellman-Ford exactly as shortest_paths.hpp:7-78 structures it (sequential, so race-free): used to count
 * the iterations / relaxed edges the reference's partial-active variant performs. */
int vglo_sssp_frontier_bf(int32_t V, const int64_t *row_ptr, const int32_t *adj, const float *w, int32_t source,
                          float *dist, int64_t *edges_relaxed, int32_t *iterations)
{
    const float inf_val = FLT_MAX - 100;
    float *prev = (float *)malloc((size_t)V * sizeof(float));
    int32_t *frontier = (int32_t *)malloc((size_t)V * sizeof(int32_t));
    if (!prev || !frontier) return -1;
    for (int32_t v = 0; v < V; v++) dist[v] = inf_val;
    dist[source] = 0;
    int64_t fsize = 1, relaxed = 0;
    int32_t iters = 0;
    frontier[0] = source;
    while (fsize > 0)
    {
        memcpy(prev, dist, (size_t)V * sizeof(float));
        for (int64_t i = 0; i < fsize; i++)
        {
            int32_t s = frontier[i];
            for (int64_t p = row_ptr[s]; p < row_ptr[s + 1]; p++)
            {
                int32_t t = adj[p];
                float cand = dist[s] + w[p];
                relaxed++;
                if (dist[t] > cand) dist[t] = cand;
            }
        }
        fsize = 0;
        for (int32_t v = 0; v < V; v++) if (dist[v] != prev[v]) frontier[fsize++] = v;
        iters++;
    }
    free(prev);
    free(frontier);
    if (edges_relaxed) *edges_relaxed = relaxed;
    if (iterations) *iterations = iters;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* CC — algorithms/cc/shiloach_vishkin.hpp:7-88: comp[v] = v (sorted id); hook comp[dst] = min over edges           */
/* (src,dst) of comp[src]; jump comp[v] = comp[comp[v]]; until nothing changes. The unique fixed point is            */
/* comp[v] = min sorted id over {v} U ancestors(v) = component minimum on symmetric graphs.             */
/* ------------------------------------------------------------------------------------------------------------ */

int vglo_cc(int32_t V, const int64_t *row_ptr, const int32_t *adj, int32_t *comp, int32_t *hook_rounds)
{
    for (int32_t v = 0; v < V; v++) comp[v] = v;
    int hook_changes = 1;
    int32_t rounds = 0;
    while (hook_changes)
    {
        hook_changes = 0;
        for (int32_t s = 0; s < V; s++)
        {
            for (int64_t p = row_ptr[s]; p < row_ptr[s + 1]; p++)
            {
                int32_t t = adj[p];
                int32_t sv = comp[s], tv = comp[t];
                if (sv < tv) { comp[t] = sv; hook_changes = 1; }
            }
        }
        int jump_changes = 1;
        while (jump_changes)
        {
            jump_changes = 0;
            for (int32_t v = 0; v < V; v++)
            {
                int32_t c = comp[v], cc = comp[c];
                if (c != cc) { comp[v] = cc; jump_changes = 1; }
            }
        }
        rounds++;
    }
    if (hook_rounds) *hook_rounds = rounds;
    return 0;
}

/* Parity predicates — vgl_runtime/helpers/verify_results/verify_results.h:9-28 (exact int / fp32 equality count). */
int64_t vglo_count_mismatch_u32(int64_t n, const uint32_t *a, const uint32_t *b)
{
    int64_t bad = 0;
    for (int64_t i = 0; i < n; i++) bad += (a[i] != b[i]);
    return bad;
}

double vglo_rel_l1_f32(int64_t n, const float *a, const float *ref)
{
    double num = 0, den = 0;
    for (int64_t i = 0; i < n; i++) { num += fabs((double)a[i] - (double)ref[i]); den += fabs((double)ref[i]); }
    return den > 0 ? num / den : num;
}
