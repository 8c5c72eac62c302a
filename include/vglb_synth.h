/*
 * vglb_synth.h — deterministic synthetic inputs (graphs, edge weights, sources).
 *
 * The reference's generators are time-seeded (vgl_runtime/graph_generation/graph_generation.hpp:128,
 * vgl_runtime/graph_generation/edges_container.h:215-233, helpers/random_generator/common_generator.hpp:8-11),
 * so identical inputs on the CPU oracle and on the GPU must come from a counter-based generator of our own.
 * Every function here is pure integer arithmetic of (seed, index): the host build (gcc), the reference harness
 * (g++) and the device build (nvcc) produce bit-identical edges and weights.
 *
 * Distributions mirror the reference:
 *   RMAT       : per bit level one quadrant draw with (a,b,c,d) in percent, default 57/19/19/5
 *                (vgl_runtime/vgl_runtime.hpp:36, graph_generation.hpp:104-190), then a vertex-label permutation
 *                (edges_container.h:215-233 random_shuffle_edges).
 *   Kronecker  : Graph500 initiator A,B,C = .57,.19,.19 drawn as two conditional uniforms per level
 *                (external_libraries/graph_500_generator.py:17-30) + label permutation.
 *   Uniform    : src,dst ~ U[0,V) iid (graph_generation.hpp:5-48).
 *   Weights    : fp32 in [0,100) = 100 * u24 * 2^-24, u24 from hash(orig_src, orig_dst, seed)
 *                (reference: U[0,MAX_WEIGHT=100], settings.h:93, common_generator.hpp:22-36).
 */
#ifndef VGLB_SYNTH_H
#define VGLB_SYNTH_H

#include <stdint.h>

#ifdef __CUDACC__
#define VGLB_HD __host__ __device__ __forceinline__
#else
#define VGLB_HD static inline
#endif

#define VGLB_GEN_RMAT 0
#define VGLB_GEN_KRONECKER 1
#define VGLB_GEN_UNIFORM 2

#define VGLB_MASTER_SEED 0xB200ULL

/* splitmix64 finalizer: a bijective 64-bit mixer used as a counter-based RNG. */
VGLB_HD uint64_t vglb_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* Bijection of [0, 2^bits): stands in for the reference's random_shuffle of vertex labels. */
VGLB_HD uint32_t vglb_permute_label(uint32_t v, int bits, uint64_t seed)
{
    if (bits <= 0) return 0;
    const uint32_t mask = (bits >= 32) ? 0xFFFFFFFFu : ((1u << bits) - 1u);
    const uint64_t k = vglb_mix64(seed ^ 0x5DEECE66DULL);
    const uint32_t m1 = ((uint32_t)k) | 1u, c1 = (uint32_t)(k >> 32);
    const uint64_t k2 = vglb_mix64(k);
    const uint32_t m2 = ((uint32_t)k2) | 1u, c2 = (uint32_t)(k2 >> 32);
    const int sh = (bits + 1) / 2;
    v = (v * m1 + c1) & mask;
    v ^= v >> sh;
    v = (v * m2 + c2) & mask;
    v ^= v >> sh;
    v = (v * m1 + c2) & mask;
    v ^= v >> sh;
    return v & mask;
}

/* thresholds on a 2^32 scale from integer percents (exact integer math, no floats). */
VGLB_HD uint32_t vglb_pct_threshold(int pct)
{
    return (uint32_t)(((uint64_t)pct << 32) / 100u);
}

/*
 * Edge number `idx` of a 2^scale-vertex graph. a,b,c are percents (d = 100-a-b-c).
 * kind: VGLB_GEN_RMAT | VGLB_GEN_KRONECKER | VGLB_GEN_UNIFORM.
 */
VGLB_HD void vglb_gen_edge(int kind, int scale, uint64_t seed, uint64_t idx, int a, int b, int c,
                           int32_t *src_out, int32_t *dst_out)
{
    const uint64_t key = vglb_mix64(seed ^ vglb_mix64(idx));
    uint32_t s = 0, d = 0;
    if (kind == VGLB_GEN_UNIFORM)
    {
        const uint32_t mask = (scale >= 32) ? 0xFFFFFFFFu : ((1u << scale) - 1u);
        const uint64_t r = vglb_mix64(key);
        s = ((uint32_t)r) & mask;
        d = ((uint32_t)(r >> 32)) & mask;
        *src_out = (int32_t)s;
        *dst_out = (int32_t)d;
        return;
    }
    if (kind == VGLB_GEN_RMAT)
    {
        const uint32_t ta = vglb_pct_threshold(a), tab = vglb_pct_threshold(a + b),
                       tabc = vglb_pct_threshold(a + b + c);
        uint64_t r64 = 0;
        for (int level = 0; level < scale; level++)
        {
            if ((level & 1) == 0) r64 = vglb_mix64(key + (uint64_t)(level >> 1) * 0xD1342543DE82EF95ULL);
            const uint32_t r = (level & 1) ? (uint32_t)(r64 >> 32) : (uint32_t)r64;
            const uint32_t ib = (r >= tab) ? 1u : 0u;                       /* lower half of the matrix */
            const uint32_t jb = (r >= ta && r < tab) || (r >= tabc) ? 1u : 0u; /* right half */
            s = (s << 1) | ib;
            d = (d << 1) | jb;
        }
    }
    else /* VGLB_GEN_KRONECKER */
    {
        /* ii_bit = u1 > A+B ; jj_bit = u2 > (ii ? C/(1-A-B) : A/(A+B)) — graph_500_generator.py:26-29 */
        const uint32_t tab = vglb_pct_threshold(a + b);
        const uint32_t t_i0 = (uint32_t)(((uint64_t)a << 32) / (uint64_t)(a + b));
        const uint32_t t_i1 = (uint32_t)(((uint64_t)c << 32) / (uint64_t)(100 - a - b));
        for (int level = 0; level < scale; level++)
        {
            const uint64_t r64 = vglb_mix64(key + (uint64_t)level * 0xD1342543DE82EF95ULL);
            const uint32_t u1 = (uint32_t)r64, u2 = (uint32_t)(r64 >> 32);
            const uint32_t ib = (u1 >= tab) ? 1u : 0u;
            const uint32_t jb = (u2 >= (ib ? t_i1 : t_i0)) ? 1u : 0u;
            s = (s << 1) | ib;
            d = (d << 1) | jb;
        }
    }
    *src_out = (int32_t)vglb_permute_label(s, scale, seed);
    *dst_out = (int32_t)vglb_permute_label(d, scale, seed);
}

/* fp32 weight in [0,100) of the edge between ORIGINAL ids (src,dst): 100 * u24 / 2^24, exact in fp32 steps. */
VGLB_HD float vglb_edge_weight(int32_t orig_src, int32_t orig_dst, uint64_t seed)
{
    const uint64_t h = vglb_mix64(seed ^ (((uint64_t)(uint32_t)orig_src << 32) | (uint64_t)(uint32_t)orig_dst));
    const uint32_t u24 = (uint32_t)(h >> 40);
    return 100.0f * ((float)u24 * (1.0f / 16777216.0f));
}

/* k-th candidate source (ORIGINAL id) for a run; callers skip candidates with out-degree 0
 * (reference: select_random_nz_vertex per round, apps/bfs/bfs.cpp:38). */
VGLB_HD int32_t vglb_source_candidate(uint64_t seed, uint64_t k, int32_t vertices_count)
{
    return (int32_t)(vglb_mix64(seed ^ (0xA5A5A5A5ULL + k * 0x9E3779B97F4A7C15ULL)) % (uint64_t)vertices_count);
}

#endif /* VGLB_SYNTH_H */
