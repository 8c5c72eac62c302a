// shim_algorithms.cu — BFS, SSSP, CC and PageRank written against the user-lambda operator API of
// include/vgl_b200/graph_abstractions_b200.cuh, the way the reference's algorithms/* are written against
// VGL_GRAPH_ABSTRACTIONS (operator sequences follow algorithms/bfs/bfs.hpp:5-51, sssp/shortest_paths.hpp:7-78,
// cc/shiloach_vishkin.hpp:7-88, pr/pr.hpp:7-148). Test driver: builds a seeded synthetic graph, runs the four
// algorithms through the generic (lambda) path and dumps the results in ORIGINAL numbering; tests/test_gpu_shim.py
// compares them with the oracle. Also exercises the misuse errors (throw const char*).
//
//   shim_algorithms <kind> <scale> <edge_factor> <seed> <source_original_id> <weight_seed> <pr_iters> <out_dir>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vgl_b200/graph_abstractions_b200.cuh"

using namespace vglb;

#define UNVISITED_VERTEX (-1)
#define FIRST_LEVEL_VERTEX 1

// ---- BFS, top-down: per level one scatter + one generate_new_frontier ------------------------------------------------
static int bfs_top_down(GraphB200 &graph, GraphAbstractionsB200 &graph_API, FrontierB200 &frontier, VerticesArrayB200<int> &levels,
                        int source_vertex)
{
    auto init_levels = [levels, source_vertex] __VGLB_COMPUTE_ARGS__ {
        levels[src_id] = (src_id == source_vertex) ? FIRST_LEVEL_VERTEX : UNVISITED_VERTEX;
    };
    frontier.set_all_active();
    graph_API.compute(graph, frontier, init_levels);
    frontier.clear();
    frontier.add_vertex(source_vertex);
    int current_level = FIRST_LEVEL_VERTEX;
    while (frontier.size() > 0)
    {
        auto edge_op = [levels, current_level] __VGLB_SCATTER_ARGS__ {
            if (levels[src_id] == current_level && levels[dst_id] == UNVISITED_VERTEX) levels[dst_id] = current_level + 1;
        };
        graph_API.scatter(graph, frontier, edge_op);
        auto on_next_level = [levels, current_level] __VGLB_GNF_ARGS__ {
            return levels[src_id] == current_level + 1 ? IN_FRONTIER_FLAG : NOT_IN_FRONTIER_FLAG;
        };
        graph_API.generate_new_frontier(graph, frontier, on_next_level);
        current_level++;
    }
    return current_level;
}

// ---- SSSP, frontier Bellman-Ford (partial-active push) ----------------------------------------------------------------
static int sssp_partial_active(GraphB200 &graph, GraphAbstractionsB200 &graph_API, FrontierB200 &work_frontier, FrontierB200 &all_active,
                               EdgesArrayB200<float> &weights, VerticesArrayB200<float> &distances, int source_vertex)
{
    VerticesArrayB200<float> prev_distances(graph);
    const float inf_val = FLT_MAX - 100.0f; // shortest_paths.hpp:22
    auto init_distances = [distances, source_vertex, inf_val] __VGLB_COMPUTE_ARGS__ {
        distances[src_id] = (src_id == source_vertex) ? 0.0f : inf_val;
    };
    all_active.set_all_active();
    graph_API.compute(graph, all_active, init_distances);
    work_frontier.clear();
    work_frontier.add_vertex(source_vertex);
    int iterations = 0;
    while (work_frontier.size() > 0)
    {
        auto save_old_distances = [distances, prev_distances] __VGLB_COMPUTE_ARGS__ { prev_distances[src_id] = distances[src_id]; };
        graph_API.compute(graph, all_active, save_old_distances);
        auto edge_op_push = [distances, weights] __VGLB_SCATTER_ARGS__ {
            const float cand = distances[src_id] + weights[global_edge_pos];
            // non-negative floats order like their bit patterns: the reference's racy "if (d[dst] > cand) d[dst] = cand"
            // (shortest_paths.hpp:46-54) made atomic
            if (cand < distances[dst_id]) atomicMin((unsigned int *)&distances[dst_id], __float_as_uint(cand));
        };
        graph_API.scatter(graph, work_frontier, edge_op_push);
        auto changes_occurred = [distances, prev_distances] __VGLB_GNF_ARGS__ {
            return distances[src_id] != prev_distances[src_id] ? IN_FRONTIER_FLAG : NOT_IN_FRONTIER_FLAG;
        };
        graph_API.generate_new_frontier(graph, work_frontier, changes_occurred);
        iterations++;
    }
    return iterations;
}

// ---- CC, min-label hooking + pointer jumping ---------------------------------------------------------------------------
static int cc_shiloach_vishkin(GraphB200 &graph, GraphAbstractionsB200 &graph_API, FrontierB200 &frontier, VerticesArrayB200<int> &components)
{
    VerticesArrayB200<int> before(graph);
    frontier.set_all_active();
    auto init_components = [components] __VGLB_COMPUTE_ARGS__ { components[src_id] = src_id; };
    graph_API.compute(graph, frontier, init_components);
    int rounds = 0;
    for (;;)
    {
        auto remember = [components, before] __VGLB_COMPUTE_ARGS__ { before[src_id] = components[src_id]; };
        graph_API.compute(graph, frontier, remember);
        auto hook = [components] __VGLB_SCATTER_ARGS__ {
            const int src_val = components[src_id];
            if (src_val < components[dst_id]) atomicMin(&components[dst_id], src_val);
        };
        graph_API.scatter(graph, frontier, hook);
        rounds++;
        auto hook_changes = [components, before] __VGLB_REDUCE_INT_ARGS__ { return components[src_id] != before[src_id] ? 1 : 0; };
        if (graph_API.reduce<int>(graph, frontier, hook_changes, REDUCE_SUM) == 0) break;
        for (;;)
        {
            graph_API.compute(graph, frontier, remember);
            auto jump = [components] __VGLB_COMPUTE_ARGS__ {
                const int c = components[src_id];
                const int cc = components[c];
                if (cc != c) components[src_id] = cc;
            };
            graph_API.compute(graph, frontier, jump);
            if (graph_API.reduce<int>(graph, frontier, hook_changes, REDUCE_SUM) == 0) break;
        }
    }
    return rounds;
}

// ---- PageRank, multicore semantics: r'[u] = k + d * (sum_{u->v, v != u} r[v] / indeg_noloops(v) + dangling) ------------
// operators used: compute (x4 setup, x1 per sweep), gather (self-loop count over the incoming direction), reduce<float>
// (dangling mass), scatter in the 8-functor form with a post-op
static void page_rank(GraphB200 &graph, GraphAbstractionsB200 &graph_API, FrontierB200 &everything, VerticesArrayB200<float> &ranks,
                      int sweeps)
{
    const int V = graph.get_vertices_count();
    const float damping = 0.85f;
    const float teleport = (float)((1.0 - damping) / ((float)V));
    VerticesArrayB200<int> self_loops(graph), in_degree(graph);
    VerticesArrayB200<float> inv_in_degree(graph), previous(graph);
    everything.set_all_active();

    // in-degrees come from the incoming direction: connections_count of a vertex there, minus its self loops
    graph_API.change_traversal_direction(GATHER);
    auto start = [ranks, self_loops, in_degree, V] __VGLB_COMPUTE_ARGS__ {
        ranks[src_id] = (float)(1.0 / V);
        self_loops[src_id] = 0;
        in_degree[src_id] = connections_count;
    };
    graph_API.compute(graph, everything, start);
    auto count_self_loops = [self_loops] __VGLB_GATHER_ARGS__ {
        if (dst_id == src_id) atomicAdd(&self_loops[src_id], 1);
    };
    graph_API.gather(graph, everything, count_self_loops);
    auto invert = [inv_in_degree, in_degree, self_loops] __VGLB_COMPUTE_ARGS__ {
        const int without_loops = in_degree[src_id] - self_loops[src_id];
        in_degree[src_id] = without_loops;
        inv_in_degree[src_id] = without_loops > 0 ? (float)(1.0 / without_loops) : 0.0f;
    };
    graph_API.compute(graph, everything, invert);
    graph_API.change_traversal_direction(SCATTER);

    for (int sweep = 0; sweep < sweeps; sweep++)
    {
        auto shift = [previous, ranks] __VGLB_COMPUTE_ARGS__ {
            previous[src_id] = ranks[src_id];
            ranks[src_id] = 0.0f;
        };
        graph_API.compute(graph, everything, shift);
        // rank mass of the vertices nobody points to, spread evenly
        auto dangling_share = [in_degree, previous, V] __VGLB_REDUCE_FLT_ARGS__ {
            return in_degree[src_id] == 0 ? previous[src_id] / V : 0.0f;
        };
        const float dangling = graph_API.reduce<float>(graph, everything, dangling_share, REDUCE_SUM);
        // several lanes work on one row here, so the accumulation into src_id is atomic (VGL_SRC_ID_ADD on the reference GPU)
        auto pull = [ranks, previous, inv_in_degree] __VGLB_SCATTER_ARGS__ {
            if (dst_id != src_id) atomicAdd(&ranks[src_id], previous[dst_id] * inv_in_degree[dst_id]);
        };
        auto finish = [ranks, teleport, damping, dangling] __VGLB_ADVANCE_POSTPROCESS_ARGS__ {
            ranks[src_id] = teleport + damping * (ranks[src_id] + dangling);
        };
        NoVertexOp nothing;
        graph_API.scatter(graph, everything, pull, nothing, finish, pull, nothing, finish);
    }
}

// wall-clock of one algorithm through the lambda API (device drained on both sides)
struct Stopwatch
{
    RuntimeB200 &rt;
    std::chrono::steady_clock::time_point t0;
    explicit Stopwatch(RuntimeB200 &_rt) : rt(_rt)
    {
        rt.synchronize();
        t0 = std::chrono::steady_clock::now();
    }
    double ms()
    {
        rt.synchronize();
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

template <typename T>
static void dump(const std::string &dir, const char *name, const std::vector<T> &v)
{
    const std::string path = dir + "/" + name;
    FILE *f = fopen(path.c_str(), "wb");
    if (!f || fwrite(v.data(), sizeof(T), v.size(), f) != v.size())
    {
        fprintf(stderr, "cannot write %s\n", path.c_str());
        exit(2);
    }
    fclose(f);
}

int main(int argc, char **argv)
{
    if (argc != 9)
    {
        fprintf(stderr, "usage: %s kind scale edge_factor seed source_orig weight_seed pr_iters out_dir\n", argv[0]);
        return 2;
    }
    const int kind = atoi(argv[1]), scale = atoi(argv[2]), ef = atoi(argv[3]);
    const unsigned long long seed = strtoull(argv[4], NULL, 0), weight_seed = strtoull(argv[6], NULL, 0);
    const int source_orig = atoi(argv[5]), pr_iters = atoi(argv[7]);
    const std::string out_dir = argv[8];
    try
    {
        const int V = 1 << scale;
        const long long E = (long long)ef << scale;
        std::vector<int> src((size_t)E), dst((size_t)E);
        check(vglb_generate_edges_host(kind, scale, E, seed, 57, 19, 19, src.data(), dst.data()));
        RuntimeB200 runtime(0);
        GraphB200 graph(runtime, V, E, src.data(), dst.data());
        GraphAbstractionsB200 graph_API(graph);
        FrontierB200 frontier(graph), all_active(graph);
        const int source_vertex = graph.reorder(source_orig, ORIGINAL, SCATTER);

        VerticesArrayB200<int> levels(graph);
        Stopwatch sw_bfs(runtime);
        const int last_level = bfs_top_down(graph, graph_API, frontier, levels, source_vertex);
        const double ms_bfs = sw_bfs.ms();
        dump(out_dir, "bfs_levels.bin", levels.to_host_original());

        EdgesArrayB200<float> weights(graph);
        weights.set_synthetic_weights(weight_seed);
        VerticesArrayB200<float> distances(graph);
        Stopwatch sw_sssp(runtime);
        const int sssp_rounds = sssp_partial_active(graph, graph_API, frontier, all_active, weights, distances, source_vertex);
        const double ms_sssp = sw_sssp.ms();
        dump(out_dir, "sssp_dist.bin", distances.to_host_original());

        VerticesArrayB200<int> components(graph);
        Stopwatch sw_cc(runtime);
        const int cc_rounds = cc_shiloach_vishkin(graph, graph_API, frontier, components);
        const double ms_cc = sw_cc.ms();
        dump(out_dir, "cc_labels.bin", components.to_host_original());

        VerticesArrayB200<float> page_ranks(graph);
        Stopwatch sw_pr(runtime);
        page_rank(graph, graph_API, frontier, page_ranks, pr_iters);
        const double ms_pr = sw_pr.ms();
        dump(out_dir, "pr_ranks.bin", page_ranks.to_host_original());

        // misuse must throw const char*, like the reference (common/advance.hpp:19-26, modification.hpp:33-36)
        int caught = 0;
        auto noop = [] __VGLB_SCATTER_ARGS__ {};
        try { graph_API.gather(graph, frontier, noop); } catch (const char *) { caught++; }   // SCATTER direction is current
        frontier.clear();
        frontier.add_vertex(0);
        try { frontier.add_vertex(1); } catch (const char *) { caught++; }                     // non-empty frontier
        // reduce max over a sparse frontier + the frontier's bookkeeping
        auto deg_op = [] __VGLB_REDUCE_INT_ARGS__ { return connections_count; };
        frontier.set_all_active();
        const int max_deg = graph_API.reduce<int>(graph, frontier, deg_op, REDUCE_MAX);
        const long long deg_sum = (long long)graph_API.reduce<double>(graph, frontier, deg_op, REDUCE_SUM);
        runtime.synchronize();
        printf("SHIM_TIMES_MS bfs=%.3f sssp=%.3f cc=%.3f pagerank=%.3f (lambda API, wall clock)\n", ms_bfs, ms_sssp, ms_cc, ms_pr);
        printf("SHIM_OK V=%d E=%lld bfs_last_level=%d sssp_rounds=%d cc_rounds=%d errors_caught=%d max_degree=%d degree_sum=%lld info_max_degree=%d\n",
               V, E, last_level, sssp_rounds, cc_rounds, caught, max_deg, deg_sum, graph.info.max_degree);
    }
    catch (const char *msg)
    {
        fprintf(stderr, "VGL error: %s\n", msg);
        return 1;
    }
    return 0;
}
