// dev/pr_lab.cu — DEVELOPER MICROBENCHMARK (not product, not test): what bounds a 4-byte gather on B200?
// Builds the bench graph through the C ABI of libvgl_b200, then streams its adjacency with "flat" kernels that only
// differ in how the gathered value is fetched. Prints ms and G gathers/s per variant.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I include dev/pr_lab.cu \
//        -L vectorgraphlibrary_b200 -lvgl_b200 -Xlinker -rpath,'$ORIGIN/../vectorgraphlibrary_b200' -o dev/pr_lab
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "vgl_b200.h"
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
#define VK(x) do { int r_ = (x); if (r_) { printf("vglb %s -> %d %s\n", #x, r_, vglb_last_error()); exit(1); } } while (0)

__device__ __forceinline__ int4 ld_stream_v4(const int4 *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_nc(const float *p)
{
    float r;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_nc_noalloc(const float *p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_nc_l1_evict_last(const float *p)
{
    float r;
    asm volatile("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_nc_l1_evict_first(const float *p)
{
    float r;
    asm volatile("ld.global.nc.L1::evict_first.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

enum Mode
{
    M_STREAM = 0,   // no gather: sum of the indices (cost of streaming the adjacency alone)
    M_LDG = 1,      // ld.global.nc of c[v]                       (the product's gather)
    M_LDG_NOALLOC,  // ld.global.nc.L1::no_allocate of c[v]
    M_UNIFORM,      // c[hash(v,p) % V]: skew-free gather over the whole 64 MB vector
    M_L1SMALL,      // c[v & 8191]: L1-resident gather (32 KB)
    M_LDS,          // smem[v % H]
    M_MIXED,        // v < H ? smem[v] : ldg(c[v])                (branchy hot/cold)
    M_COAL,         // c[(p*4+k) % V] coalesced gather (upper bound)
    M_LDG64,        // 8-byte gather: ((float2*)c)[v>>1]
    M_HOTONLY,      // v < H ? smem[v] : 0  (what the hot half costs alone)
    M_COLDONLY,     // v < H ? 0 : ldg      (what the cold half costs alone)
    M_EVICT_SPLIT,  // v < H ? ld.L1::evict_last : ld.L1::evict_first   (H = id threshold, no smem)
    M_NOALLOC_COLD, // v < H ? ld : ld.L1::no_allocate
    M_EVICT_FIRST_COLD, // v < H ? ld : ld.L1::evict_first
    M_UNIPAIR,      // uniform random, lanes 2k/2k+1 share a 128 B line (different sectors)
    M_UNIQUAD,      // uniform random, 4 lanes share a 128 B line (4 sectors of it)
    M_UNISAMESEC,   // uniform random, lanes 2k/2k+1 share a 32 B sector
};

extern __shared__ __align__(16) float s_tab[];

template <int MODE>
__device__ __forceinline__ float fetch(const float *__restrict__ c, int32_t v, uint32_t V, int H, int64_t p)
{
    if (MODE == M_STREAM) return __int_as_float(v & 0x3fffff);
    if (MODE == M_LDG) return ld_nc(c + v);
    if (MODE == M_LDG_NOALLOC) return ld_nc_noalloc(c + v);
    if (MODE == M_UNIFORM) return ld_nc(c + (hash32((uint32_t)v * 2654435761u + (uint32_t)p) & (V - 1)));
    if (MODE == M_L1SMALL) return ld_nc(c + (v & 8191));
    if (MODE == M_LDS) return s_tab[(uint32_t)v % (uint32_t)H];
    if (MODE == M_MIXED) return v < H ? s_tab[v] : ld_nc(c + v);
    if (MODE == M_COAL) return ld_nc(c + (p & (V - 1)));
    if (MODE == M_LDG64) { float2 t = *reinterpret_cast<const float2 *>(c + (v & ~1)); return t.x + t.y; }
    if (MODE == M_HOTONLY) return v < H ? s_tab[v] : 0.f;
    if (MODE == M_COLDONLY) return v < H ? 0.f : ld_nc(c + v);
    if (MODE == M_EVICT_SPLIT) return v < H ? ld_nc_l1_evict_last(c + v) : ld_nc_l1_evict_first(c + v);
    if (MODE == M_NOALLOC_COLD) return v < H ? ld_nc(c + v) : ld_nc_noalloc(c + v);
    if (MODE == M_EVICT_FIRST_COLD) return v < H ? ld_nc(c + v) : ld_nc_l1_evict_first(c + v);
    if (MODE == M_UNIPAIR || MODE == M_UNIQUAD || MODE == M_UNISAMESEC)
    {
        const unsigned lane = threadIdx.x & 31;
        const unsigned grp = MODE == M_UNIQUAD ? (lane >> 2) : (lane >> 1);
        const unsigned sub = MODE == M_UNIQUAD ? (lane & 3) : (lane & 1);
        // all lanes of a group derive the same line from the group leader's index
        const int32_t vl = __shfl_sync(0xffffffffu, v, MODE == M_UNIQUAD ? (lane & ~3u) : (lane & ~1u));
        const uint32_t line = hash32((uint32_t)vl * 2654435761u + (uint32_t)(p >> 7) + grp) & ((V - 1) >> 5);
        const uint32_t off = MODE == M_UNISAMESEC ? sub : sub * 8;
        return ld_nc(c + (line << 5) + off);
    }
    return 0.f;
}

template <int MODE, int UNROLL>
__global__ void flat_kernel(const int4 *__restrict__ adj4, int64_t n4, const float *__restrict__ c, uint32_t V, int H,
                            float *__restrict__ out)
{
    if (MODE == M_LDS || MODE == M_MIXED || MODE == M_HOTONLY || MODE == M_COLDONLY)
    {
        for (int i = threadIdx.x; i < H; i += blockDim.x) s_tab[i] = c[i];
        __syncthreads();
    }
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * UNROLL;
    for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; q0 < n4; q0 += stride)
    {
        int4 a[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            const int64_t q = q0 + (int64_t)u * blockDim.x;
            a[u] = q < n4 ? ld_stream_v4(adj4 + q) : make_int4(0, 0, 0, 0);
        }
        float f[UNROLL][4];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            const int64_t p = (q0 + (int64_t)u * blockDim.x) * 4;
            f[u][0] = fetch<MODE>(c, a[u].x, V, H, p);
            f[u][1] = fetch<MODE>(c, a[u].y, V, H, p + 1);
            f[u][2] = fetch<MODE>(c, a[u].z, V, H, p + 2);
            f[u][3] = fetch<MODE>(c, a[u].w, V, H, p + 3);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            acc0 += f[u][0]; acc1 += f[u][1]; acc2 += f[u][2]; acc3 += f[u][3];
        }
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = (acc0 + acc1) + (acc2 + acc3);
}

// chunk-partitioned adjacency: every chunk of CH edges lists its hot targets (v < H) first; hot_count[chunk]
#define CH 2048
__global__ void partition_chunks_kernel(const int32_t *__restrict__ adj, int64_t E, int H, int32_t *__restrict__ out,
                                        int32_t *__restrict__ hot_count)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nchunks = (E + CH - 1) / CH;
    if (warp >= nchunks) return;
    const int64_t base = warp * CH;
    const int n = (int)min((int64_t)CH, E - base);
    int hot = 0;
    for (int i = lane; i < ((n + 31) & ~31); i += 32)
    {
        const bool h = i < n && adj[base + i] < H;
        hot += __popc(__ballot_sync(0xffffffffu, h));
    }
    int hpos = 0, cpos = hot;
    for (int i = lane; i < ((n + 31) & ~31); i += 32)
    {
        const bool in = i < n;
        const int32_t v = in ? adj[base + i] : 0;
        const bool h = in && v < H;
        const unsigned mh = __ballot_sync(0xffffffffu, h), mc = __ballot_sync(0xffffffffu, in && !h);
        const unsigned below = (1u << lane) - 1u;
        if (h) out[base + hpos + __popc(mh & below)] = v;
        else if (in) out[base + cpos + __popc(mc & below)] = v;
        hpos += __popc(mh);
        cpos += __popc(mc);
    }
    if (lane == 0) hot_count[warp] = hot;
}

// one warp per chunk: hot prefix through shared memory, cold suffix through ld.global.nc; 128 edges per warp-iteration
template <int UNROLL>
__global__ void part_kernel(const int32_t *__restrict__ adjp, const int32_t *__restrict__ hot_count, int64_t E,
                            const float *__restrict__ c, int H, float *__restrict__ out)
{
    for (int i = threadIdx.x; i < H; i += blockDim.x) s_tab[i] = c[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t nchunks = E / CH; // lab: E is a multiple of CH
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    for (int64_t ch = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ch < nchunks; ch += nwarps)
    {
        const int4 *a4 = reinterpret_cast<const int4 *>(adjp + ch * CH);
        const int hot = hot_count[ch];
#pragma unroll 1
        for (int it = 0; it < CH / 128; it += UNROLL)
        {
            int4 a[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) a[u] = ld_stream_v4(a4 + (it + u) * 32 + lane);
#pragma unroll
            for (int u = 0; u < UNROLL; u++)
            {
                const int pos = (it + u) * 128 + lane * 4;
                if (pos + 128 - lane * 4 <= hot) // whole warp-iteration is hot (uniform branch)
                {
                    acc0 += s_tab[a[u].x]; acc1 += s_tab[a[u].y]; acc2 += s_tab[a[u].z]; acc3 += s_tab[a[u].w];
                }
                else if (pos - lane * 4 >= hot) // whole warp-iteration is cold
                {
                    acc0 += ld_nc(c + a[u].x); acc1 += ld_nc(c + a[u].y); acc2 += ld_nc(c + a[u].z); acc3 += ld_nc(c + a[u].w);
                }
                else
                {
                    acc0 += pos + 0 < hot ? s_tab[a[u].x] : ld_nc(c + a[u].x);
                    acc1 += pos + 1 < hot ? s_tab[a[u].y] : ld_nc(c + a[u].y);
                    acc2 += pos + 2 < hot ? s_tab[a[u].z] : ld_nc(c + a[u].z);
                    acc3 += pos + 3 < hot ? s_tab[a[u].w] : ld_nc(c + a[u].w);
                }
            }
        }
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = (acc0 + acc1) + (acc2 + acc3);
}

// distributed shared memory: a cluster of CS CTAs holds c[0 .. CS*H); v below that is read from CTA v / H's smem
template <int UNROLL>
__global__ void dsmem_kernel(const int4 *__restrict__ adj4, int64_t n4, const float *__restrict__ c, int H, int CS,
                             float *__restrict__ out)
{
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    for (int i = threadIdx.x; i < H; i += blockDim.x) s_tab[i] = c[(int64_t)rank * H + i];
    cluster.sync();
    const uint32_t span = (uint32_t)H * (uint32_t)CS;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * UNROLL;
    auto get = [&](int32_t v) -> float {
        if ((uint32_t)v < span)
        {
            const unsigned r = (uint32_t)v / (uint32_t)H;
            const float *remote = cluster.map_shared_rank(s_tab, r);
            return remote[(uint32_t)v - r * (uint32_t)H];
        }
        return ld_nc(c + v);
    };
    for (int64_t q0 = (int64_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; q0 < n4; q0 += stride)
    {
        int4 a[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            const int64_t q = q0 + (int64_t)u * blockDim.x;
            a[u] = q < n4 ? ld_stream_v4(adj4 + q) : make_int4(0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            const int64_t q = q0 + (int64_t)u * blockDim.x;
            if (q < n4)
            {
                acc0 += get(a[u].x); acc1 += get(a[u].y); acc2 += get(a[u].z); acc3 += get(a[u].w);
            }
        }
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = (acc0 + acc1) + (acc2 + acc3);
    cluster.sync();
}

// tile-per-CTA: gather phase (flat, coalesced int4, U vectors per thread) -> values staged in shared memory -> reduction
// phase over fake rows of DEG consecutive edges, G lanes per row (stands in for the segmented reduction)
template <int U, int DEG, int G>
__global__ void __launch_bounds__(256) staged_kernel(const int4 *__restrict__ adj4, int64_t n4, const float *__restrict__ c,
                                                     float *__restrict__ out)
{
    constexpr int T = 256 * U * 4;
    __shared__ __align__(16) float s_val[T];
    const int64_t q0 = (int64_t)blockIdx.x * (256 * U);
    int4 a[U];
#pragma unroll
    for (int u = 0; u < U; u++) a[u] = q0 + u * 256 + threadIdx.x < n4 ? ld_stream_v4(adj4 + q0 + u * 256 + threadIdx.x) : make_int4(0, 0, 0, 0);
    float4 f[U];
#pragma unroll
    for (int u = 0; u < U; u++)
    {
        f[u].x = ld_nc(c + a[u].x); f[u].y = ld_nc(c + a[u].y); f[u].z = ld_nc(c + a[u].z); f[u].w = ld_nc(c + a[u].w);
    }
#pragma unroll
    for (int u = 0; u < U; u++) reinterpret_cast<float4 *>(s_val)[u * 256 + threadIdx.x] = f[u];
    __syncthreads();
    constexpr int ROWS = T / DEG;
    const int gid = threadIdx.x / G, gl = threadIdx.x % G;
    for (int r = gid; r < ROWS; r += 256 / G)
    {
        float acc = 0.f;
        for (int j = gl; j < DEG; j += G) acc += s_val[r * DEG + j];
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (gl == 0) out[(int64_t)blockIdx.x * ROWS + r] = acc;
    }
}

// the same, persistent: grid = resident CTAs, each looping over tiles with stride gridDim.x
template <int U, int DEG, int G, int STAGE>
__global__ void __launch_bounds__(256) staged_persistent_kernel(const int4 *__restrict__ adj4, int64_t n4, const float *__restrict__ c,
                                                                float *__restrict__ out)
{
    constexpr int T = 256 * U * 4;
    __shared__ __align__(16) float s_val[STAGE ? T : 4];
    const int64_t ntiles = (n4 + 256 * U - 1) / (256 * U);
    float keep = 0.f;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
        const int64_t q0 = tile * (256 * U);
        int4 a[U];
#pragma unroll
        for (int u = 0; u < U; u++) a[u] = q0 + u * 256 + threadIdx.x < n4 ? ld_stream_v4(adj4 + q0 + u * 256 + threadIdx.x) : make_int4(0, 0, 0, 0);
        float4 f[U];
#pragma unroll
        for (int u = 0; u < U; u++)
        {
            f[u].x = ld_nc(c + a[u].x); f[u].y = ld_nc(c + a[u].y); f[u].z = ld_nc(c + a[u].z); f[u].w = ld_nc(c + a[u].w);
        }
        if (STAGE)
        {
#pragma unroll
            for (int u = 0; u < U; u++) reinterpret_cast<float4 *>(s_val)[u * 256 + threadIdx.x] = f[u];
            __syncthreads();
            constexpr int ROWS = T / DEG;
            const int gid = threadIdx.x / G, gl = threadIdx.x % G;
            for (int r = gid; r < ROWS; r += 256 / G)
            {
                float acc = 0.f;
                for (int j = gl; j < DEG; j += G) acc += s_val[r * DEG + j];
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (gl == 0) out[tile * ROWS + r] = acc;
            }
            __syncthreads();
        }
        else
        {
#pragma unroll
            for (int u = 0; u < U; u++) keep += (f[u].x + f[u].y) + (f[u].z + f[u].w);
        }
    }
    if (!STAGE) out[(int64_t)blockIdx.x * 256 + threadIdx.x] = keep;
}

template <int U, int DEG, int G, int STAGE>
static void run_staged_persistent(const int32_t *adj, int64_t E, const float *c, float *d_out, int ctas_per_sm)
{
    const int64_t n4 = E / 4;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, staged_persistent_kernel<U, DEG, G, STAGE>, 256, 0));
    const int per_sm = ctas_per_sm > 0 ? std::min(occ, ctas_per_sm) : 0;
    const int64_t ntiles = (n4 + 256 * U - 1) / (256 * U);
    const int grid = per_sm > 0 ? 148 * per_sm : (int)ntiles;
    const float ms = time_ms([&] { staged_persistent_kernel<U, DEG, G, STAGE><<<grid, 256>>>((const int4 *)adj, n4, c, d_out); });
    printf("%s tile %5d edges, rows of %4d, %2d lanes/row, grid %6d (%d/SM, occ %d) : %8.3f ms  %7.1f Gedge/s\n",
           STAGE ? "PERSIST-STAGED" : "PERSIST-NOSTAGE", 256 * U * 4, DEG, G, grid, per_sm, occ, ms, E / ms * 1e-6);
    fflush(stdout);
}

template <int U, int DEG, int G>
static void run_staged(const int32_t *adj, int64_t E, const float *c, float *d_out)
{
    const int64_t n4 = E / 4;
    const int grid = (int)((n4 + 256 * U - 1) / (256 * U));
    const float ms = time_ms([&] { staged_kernel<U, DEG, G><<<grid, 256>>>((const int4 *)adj, n4, c, d_out); });
    printf("STAGED tile %5d edges, fake rows of %4d, %2d lanes/row : %8.3f ms  %7.1f Gedge/s\n", 256 * U * 4, DEG, G, ms, E / ms * 1e-6);
    fflush(stdout);
}

__global__ void fill_random_kernel(float *c, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        c[i] = (float)(hash32((uint32_t)i) >> 8) * (1.0f / 16777216.0f);
}

static cudaEvent_t ev0, ev1;
template <class F>
static float time_ms(F launch, int reps = 5)
{
    launch();
    CK(cudaDeviceSynchronize());
    std::vector<float> t;
    for (int r = 0; r < reps; r++)
    {
        CK(cudaEventRecord(ev0));
        launch();
        CK(cudaEventRecord(ev1));
        CK(cudaEventSynchronize(ev1));
        float ms;
        CK(cudaEventElapsedTime(&ms, ev0, ev1));
        t.push_back(ms);
    }
    CK(cudaGetLastError());
    std::sort(t.begin(), t.end());
    return t[t.size() / 2];
}

static double checksum(const float *d_out, size_t n)
{
    std::vector<float> h(n);
    CK(cudaMemcpy(h.data(), d_out, n * 4, cudaMemcpyDeviceToHost));
    double s = 0;
    for (float x : h) s += x;
    return s;
}

template <int MODE, int UNROLL>
static void run_flat(const char *name, const int32_t *adj, int64_t E, const float *c, int32_t V, int H, int threads,
                     int ctas_per_sm, float *d_out)
{
    const size_t smem = (MODE == M_LDS || MODE == M_MIXED || MODE == M_HOTONLY || MODE == M_COLDONLY) ? (size_t)H * 4 : 0;
    CK(cudaFuncSetAttribute(flat_kernel<MODE, UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, flat_kernel<MODE, UNROLL>, threads, smem));
    const int per_sm = std::min(occ, ctas_per_sm);
    if (per_sm < 1) { printf("%-34s cannot launch\n", name); return; }
    const int grid = 148 * per_sm;
    CK(cudaMemset(d_out, 0, (size_t)grid * threads * 4));
    const float ms = time_ms([&] { flat_kernel<MODE, UNROLL><<<grid, threads, smem>>>((const int4 *)adj, E / 4, c, (uint32_t)V, H, d_out); });
    printf("%-34s thr %4d x %d/SM unroll %d H %6d : %8.3f ms  %7.1f Gedge/s  sum %.6e\n", name, threads, per_sm, UNROLL, H, ms,
           E / ms * 1e-6, checksum(d_out, (size_t)grid * threads));
    fflush(stdout);
}

int main(int argc, char **argv)
{
    const int scale = argc > 1 ? atoi(argv[1]) : 24;
    const int ef = 16;
    const int32_t V = 1 << scale;
    const int64_t E = (int64_t)ef << scale;
    vglb_ctx *ctx;
    VK(vglb_init(0, &ctx));
    int32_t *d_src, *d_dst;
    VK(vglb_malloc(ctx, E * 4, (void **)&d_src));
    VK(vglb_malloc(ctx, E * 4, (void **)&d_dst));
    VK(vglb_generate_edges_device(ctx, 0, scale, E, 0xB200, 57, 19, 19, d_src, d_dst));
    vglb_graph *g;
    VK(vglb_graph_from_edges(ctx, V, E, d_src, d_dst, 1, 0, &g));
    VK(vglb_free(ctx, d_src));
    VK(vglb_free(ctx, d_dst));
    vglb_graph_info info;
    VK(vglb_graph_get_info(g, &info));
    printf("graph scale %d V %d E %lld maxdeg %d tiers:", scale, V, (long long)E, info.max_degree);
    for (int t = 0; t < 8; t++) printf(" %d", info.tier_border[t]);
    printf("\n");
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));

    // the product
    float *d_ranks;
    CK(cudaMalloc(&d_ranks, (size_t)V * 4));
    for (int r = 0; r < 3; r++)
    {
        vglb_stats st;
        VK(vglb_pagerank(ctx, g, 20, 0.85f, d_ranks, &st));
        printf("product vglb_pagerank: %.3f ms per sweep, %.1f Gedge/s\n", st.seconds * 1e3 / 20, 20.0 * E / st.seconds * 1e-9);
    }

    float *c, *d_out;
    CK(cudaMalloc(&c, (size_t)V * 4));
    CK(cudaMalloc(&d_out, (size_t)E * 4 + (size_t)148 * 16 * 1024 * 4));
    fill_random_kernel<<<148 * 8, 256>>>(c, V);
    const int32_t *adj = info.d_out_adj;

    // fraction of targets below H (how hot the hubs' values are)
    {
        std::vector<int32_t> h((size_t)E);
        CK(cudaMemcpy(h.data(), adj, (size_t)E * 4, cudaMemcpyDeviceToHost));
        const int Hs[] = {8192, 16384, 24576, 32768, 49152, 57344, 98304, 196608, 393216, 786432, 1572864, 3145728};
        int64_t cnt[12] = {0};
        for (int64_t i = 0; i < E; i++)
            for (int k = 0; k < 12; k++)
                if (h[i] < Hs[k]) cnt[k]++;
        for (int k = 0; k < 12; k++) printf("targets < %8d : %.3f\n", Hs[k], (double)cnt[k] / E);
    }

#define FLAT(MODE, U, H, T, C) run_flat<MODE, U>(#MODE, adj, E, c, V, H, T, C, d_out)
    FLAT(M_LDG, 2, 0, 256, 8);
    run_staged<4, 128, 32>(adj, E, c, d_out);
    run_staged_persistent<4, 128, 32, 0>(adj, E, c, d_out, 8);
    run_staged_persistent<4, 128, 32, 0>(adj, E, c, d_out, 4);
    run_staged_persistent<4, 128, 32, 0>(adj, E, c, d_out, 0);
    run_staged_persistent<4, 128, 32, 1>(adj, E, c, d_out, 8);
    run_staged_persistent<4, 128, 32, 1>(adj, E, c, d_out, 4);
    run_staged_persistent<4, 128, 32, 1>(adj, E, c, d_out, 0);
    run_staged_persistent<2, 128, 32, 0>(adj, E, c, d_out, 8);
    run_staged_persistent<2, 128, 32, 1>(adj, E, c, d_out, 8);
    run_staged_persistent<2, 128, 32, 0>(adj, E, c, d_out, 0);
    run_staged_persistent<1, 128, 32, 0>(adj, E, c, d_out, 8);
    run_staged_persistent<1, 128, 32, 0>(adj, E, c, d_out, 0);
    // partitioned chunks
    if (argc > 2)
    {
        int32_t *adjp, *hot_count;
        CK(cudaMalloc(&adjp, (size_t)E * 4));
        CK(cudaMalloc(&hot_count, (size_t)(E / CH + 1) * 4));
        const int Hs[] = {49152, 24576, 12288};
        const int thr[] = {1024, 512, 256};
        for (int k = 0; k < 3; k++)
        {
            const int H = Hs[k];
            const int64_t nchunks = E / CH;
            partition_chunks_kernel<<<(unsigned)((nchunks * 32 + 255) / 256), 256>>>(adj, E, H, adjp, hot_count);
            CK(cudaDeviceSynchronize());
            const size_t smem = (size_t)H * 4;
            CK(cudaFuncSetAttribute(part_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute(part_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int occ = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, part_kernel<2>, thr[k], smem));
            const int grid = 148 * occ;
            float ms = time_ms([&] { part_kernel<2><<<grid, thr[k], smem>>>(adjp, hot_count, E, c, H, d_out); });
            printf("%-34s thr %4d x %d/SM unroll 2 H %6d : %8.3f ms  %7.1f Gedge/s  sum %.6e\n", "PARTITIONED", thr[k], occ, H, ms,
                   E / ms * 1e-6, checksum(d_out, (size_t)grid * thr[k]));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, part_kernel<4>, thr[k], smem));
            ms = time_ms([&] { part_kernel<4><<<148 * occ, thr[k], smem>>>(adjp, hot_count, E, c, H, d_out); });
            printf("%-34s thr %4d x %d/SM unroll 4 H %6d : %8.3f ms  %7.1f Gedge/s  sum %.6e\n", "PARTITIONED", thr[k], occ, H, ms,
                   E / ms * 1e-6, checksum(d_out, (size_t)148 * occ * thr[k]));
            fflush(stdout);
        }
        cudaFree(adjp);
        cudaFree(hot_count);
    }

    // distributed shared memory
    if (argc > 2)
    {
        const int css[] = {2, 4, 8, 16};
        for (int k = 0; k < 4; k++)
        {
            const int CS = css[k], H = 49152, threads = 1024;
            const size_t smem = (size_t)H * 4;
            auto kern = dsmem_kernel<2>;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (CS > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3(threads);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            cfg.gridDim = dim3(CS);
            int nclusters = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
            if (e != cudaSuccess || nclusters < 1) { printf("DSMEM cluster %d: not launchable (%s)\n", CS, cudaGetErrorString(e)); cudaGetLastError(); continue; }
            cfg.gridDim = dim3(nclusters * CS);
            const int4 *a4 = (const int4 *)adj;
            const int64_t n4 = E / 4;
            CK(cudaMemset(d_out, 0, (size_t)cfg.gridDim.x * threads * 4));
            const float ms = time_ms([&] { CK(cudaLaunchKernelEx(&cfg, kern, a4, n4, (const float *)c, H, CS, d_out)); });
            printf("DSMEM cluster %2d (%3d clusters, %3d CTAs) span %7d : %8.3f ms  %7.1f Gedge/s  sum %.6e\n", CS, nclusters,
                   nclusters * CS, H * CS, ms, E / ms * 1e-6, checksum(d_out, (size_t)cfg.gridDim.x * threads));
            fflush(stdout);
        }
    }
    return 0;
}
