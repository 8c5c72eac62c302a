// dev/pr_bin_lab.cu — DEVELOPER MICROBENCHMARK (not product, not test): column-binned gather for the PageRank sweep.
// Question: pr_sweep_kernel is bound by the SM's L1-miss request port (one 128-byte line request per clock and SM; ~56 % of
// the gathers of RMAT-24 miss). If the edges of the heavy rows are regrouped by COLUMN BIN (bins of H consecutive column ids,
// whose slice of the contribution vector sits in shared memory while the bin is processed), how fast is
//   (a) the binned part:  2-byte local column ids, run-start flags, gathers from shared memory, a segmented sum per
//                         (bin,row) run, one partial-sum store per run;
//   (b) what is left:     the flat stream of the edges that were not binned (low-degree rows, columns beyond the bins)?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I include dev/pr_bin_lab.cu \
//        -L vectorgraphlibrary_b200 -lvgl_b200 -Xlinker -rpath,'$ORIGIN/../vectorgraphlibrary_b200' -o dev/pr_bin_lab
//   ./dev/pr_bin_lab <scale> <min degree of a binned row> <H> <number of bins>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "vgl_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
#define VK(x) do { int r_ = (x); if (r_) { printf("vglb %s -> %d %s\n", #x, r_, vglb_last_error()); exit(1); } } while (0)

#define CHUNK 4096 // slots per chunk (8 warp steps of 512)
#define STEP 512
#define PIECE 4096 // longer runs are cut into pieces with their own slots
#define STAGE 128  // a step closes at most 512 / ALIGN runs
#define ALIGN 4    // runs start at multiples of ALIGN slots (padding slots gather a zero)

__device__ __forceinline__ int4 ld_stream_v4(const int4 *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_stream_v4u(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
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

// flat stream of 4-byte column ids, hot ids (< Hot) through L1, the rest without allocating — what is left of the sweep
template <int UNROLL>
__global__ void flat_kernel(const int4 *__restrict__ adj4, int64_t n4, const float *__restrict__ c, int Hot, float *__restrict__ out)
{
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
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
        {
            acc0 += a[u].x < Hot ? ld_nc(c + a[u].x) : ld_nc_noalloc(c + a[u].x);
            acc1 += a[u].y < Hot ? ld_nc(c + a[u].y) : ld_nc_noalloc(c + a[u].y);
            acc2 += a[u].z < Hot ? ld_nc(c + a[u].z) : ld_nc_noalloc(c + a[u].z);
            acc3 += a[u].w < Hot ? ld_nc(c + a[u].w) : ld_nc_noalloc(c + a[u].w);
        }
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = (acc0 + acc1) + (acc2 + acc3);
}

extern __shared__ __align__(16) float s_slice[];

// Binned part. Edges sorted by (bin, row); a (bin,row) run starts at a multiple of ALIGN slots (padded with slots whose column
// is H: s_slice[H] = 0) and runs longer than PIECE are cut into pieces with their own partial-sum slots. wcol = column - bin * H
// as uint16. A warp step covers 512 slots, 16 consecutive ones per lane, stored as two 256-slot halves so that both 16-byte
// loads of a warp are contiguous. Everything about the run structure is static and precomputed per lane and step (meta):
//   bits 0..3   a run starts at slot 4g of the lane          bits 4..11  number of run starts in the lanes before (this step)
//   bits 12..16 segmented-scan mask: add the value of lane - 2^k in round k
//   bit 17      some lane before this one has a start         bits 18..25 run starts in the whole step
// step_run0[s] = number of starts before step s; run_slot[r] = where run r's sum goes. Every bin is padded to whole chunks; the
// padding starts with a flagged dummy run that is never closed.
// One persistent CTA per SM takes a contiguous range of chunks; the warp that owns a chunk emits the runs that START in it:
// it ignores the leading part of the run that started earlier and runs on past the end of the chunk until the next start.
struct StepData
{
    uint4 a, b;
    unsigned m;
};
__device__ __forceinline__ StepData load_step(const uint4 *__restrict__ wcol8, const uint32_t *__restrict__ meta, int64_t s, int lane)
{
    StepData d;
    d.a = ld_stream_v4u(wcol8 + s * 64 + lane);
    d.b = ld_stream_v4u(wcol8 + s * 64 + 32 + lane);
    d.m = __ldg(meta + s * 32 + lane);
    return d;
}

template <bool EMIT>
__global__ void __launch_bounds__(1024, 1) bin_kernel(const uint4 *__restrict__ wcol8, const uint32_t *__restrict__ meta, const int32_t *__restrict__ step_run0,
                                                       const int32_t *__restrict__ run_slot, const int32_t *__restrict__ bin_chunk0, const int32_t *__restrict__ cta_chunk0, int nbins, int H,
                                                       const float *__restrict__ c, float *__restrict__ slot, float *__restrict__ sink, long long *__restrict__ cta_ns)
{
    const unsigned FULL = 0xffffffffu;
    long long t_begin = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_begin));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int k_lo = cta_chunk0[blockIdx.x], k_hi = cta_chunk0[blockIdx.x + 1];
    constexpr int SPC = CHUNK / STEP;
    float sunk = 0.f;
    float *stage = s_slice + H + 4 + warp * STAGE;
    int bin = 0;
    while (bin < nbins && bin_chunk0[bin + 1] <= k_lo) bin++;
    for (; bin < nbins && bin_chunk0[bin] < k_hi; bin++)
    {
        const int b_lo = max(k_lo, bin_chunk0[bin]), b_hi = min(k_hi, bin_chunk0[bin + 1]);
        const int bin_end_step = bin_chunk0[bin + 1] * SPC;
        __syncthreads();
        {
            const float4 *src = reinterpret_cast<const float4 *>(c + (int64_t)bin * H);
            float4 *dst = reinterpret_cast<float4 *>(s_slice);
            for (int i = threadIdx.x; i < H / 4; i += blockDim.x) dst[i] = src[i];
            if (threadIdx.x < 4) s_slice[H + threadIdx.x] = 0.f;
        }
        __syncthreads();
        for (int k = b_lo + warp; k < b_hi; k += nwarps)
        {
            const int r_first = step_run0[k * SPC], r_end = step_run0[(k + 1) * SPC]; // runs that start in this chunk
            if (r_first == r_end) continue;
            float carry = 0.f; // sum of the open run so far (lanes of earlier steps)
            int s = k * SPC;
            StepData cur = load_step(wcol8, meta, s, lane);
            for (; s < bin_end_step; s++)
            {
                StepData nxt;
                if (s + 1 < bin_end_step) nxt = load_step(wcol8, meta, s + 1, lane);
                else nxt = cur;
                const int run0 = step_run0[s];
                const unsigned m = cur.m;
                float g[4];
                g[0] = (s_slice[cur.a.x & 0xffff] + s_slice[cur.a.x >> 16]) + (s_slice[cur.a.y & 0xffff] + s_slice[cur.a.y >> 16]);
                g[1] = (s_slice[cur.a.z & 0xffff] + s_slice[cur.a.z >> 16]) + (s_slice[cur.a.w & 0xffff] + s_slice[cur.a.w >> 16]);
                g[2] = (s_slice[cur.b.x & 0xffff] + s_slice[cur.b.x >> 16]) + (s_slice[cur.b.y & 0xffff] + s_slice[cur.b.y >> 16]);
                g[3] = (s_slice[cur.b.z & 0xffff] + s_slice[cur.b.z >> 16]) + (s_slice[cur.b.w & 0xffff] + s_slice[cur.b.w >> 16]);
                const int pre = (m >> 4) & 0xff, total = (m >> 18) & 0xff;
                // in-lane pass: head = sum before the first start, tail = open run; the sums of the runs that close inside the lane
                // are staged in shared memory by their index within the step (start t closes run run0 + t - 1)
                float run = 0.f, head = 0.f;
                int j = 0;
#pragma unroll
                for (int e = 0; e < 4; e++)
                {
                    if ((m >> e) & 1)
                    {
                        if (j == 0) head = run;
                        else if (EMIT) stage[pre + j] = run;
                        else sunk += run;
                        run = 0.f;
                        j++;
                    }
                    run += g[e];
                }
                if (j == 0) head = run;
                // segmented inclusive scan over lanes of (lane with a start ? tail : whole lane), static masks
                float v = run;
#pragma unroll
                for (int o = 0; o < 5; o++)
                {
                    const float vu = __shfl_up_sync(FULL, v, 1 << o);
                    if ((m >> (12 + o)) & 1) v += vu;
                }
                float ex = __shfl_up_sync(FULL, v, 1);
                if (lane == 0) ex = 0.f;
                if (j > 0)
                {
                    const float first_close = (((m >> 17) & 1) ? ex : carry + ex) + head; // closes run run0 + pre - 1
                    if (EMIT) stage[pre] = first_close;
                    else sunk += first_close;
                }
                const float v31 = __shfl_sync(FULL, v, 31);
                carry = total > 0 ? v31 : carry + v31;
                if (EMIT && total > 0)
                {
                    __syncwarp();
                    for (int t0 = 0; t0 < total; t0 += 64)
                    {
                        const int ta = t0 + lane, tb = ta + 32;
                        const int ra = run0 + ta - 1, rb = run0 + tb - 1;
                        const int sa = (ta < total && ra >= r_first && ra < r_end) ? __ldg(run_slot + ra) : -1;
                        const int sb = (tb < total && rb >= r_first && rb < r_end) ? __ldg(run_slot + rb) : -1;
                        if (sa >= 0) slot[sa] = stage[ta];
                        if (sb >= 0) slot[sb] = stage[tb];
                    }
                    __syncwarp();
                }
                if (s + 1 >= (k + 1) * SPC && run0 + total > r_end) break; // run r_end has started: every run of this chunk is closed
                cur = nxt;
            }
        }
    }
    if (!EMIT) sink[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = sunk;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        long long t_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
        cta_ns[blockIdx.x] = t_end - t_begin;
    }
}

__global__ void fill_random_kernel(float *c, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    {
        uint32_t x = (uint32_t)i;
        x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
        c[i] = (float)(x >> 8) * (1.0f / 16777216.0f);
    }
}

static cudaEvent_t ev0, ev1;
template <class F>
static float time_ms(F launch, int reps = 7)
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
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    std::vector<int64_t> ptr((size_t)V + 1);
    std::vector<int32_t> adj((size_t)E);
    CK(cudaMemcpy(ptr.data(), info.d_out_ptr, ((size_t)V + 1) * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(adj.data(), info.d_out_adj, (size_t)E * 4, cudaMemcpyDeviceToHost));
    printf("graph scale %d V %d E %lld maxdeg %d\n", scale, V, (long long)E, info.max_degree);

    float *c, *d_out, *d_slot;
    CK(cudaMalloc(&c, (size_t)V * 4));
    CK(cudaMalloc(&d_out, (size_t)148 * 16 * 1024 * 4));
    fill_random_kernel<<<148 * 8, 256>>>(c, V);
    int32_t *d_adj_main;
    CK(cudaMalloc(&d_adj_main, (size_t)E * 4 + 64));

    // baseline: the whole adjacency, flat
    {
        CK(cudaMemcpy(d_adj_main, adj.data(), (size_t)E * 4, cudaMemcpyHostToDevice));
        const float ms = time_ms([&] { flat_kernel<2><<<148, 1024>>>((const int4 *)d_adj_main, E / 4, c, 49152, d_out); });
        printf("flat, all edges (hot 49152 through L1)                  : %8.3f ms  %7.1f Gedge/s\n", ms, E / ms * 1e-6);
    }

    struct Cfg { int tb, H, nb; };
    std::vector<Cfg> cfgs;
    if (argc > 4) cfgs.push_back({atoi(argv[2]), atoi(argv[3]), atoi(argv[4])});
    else cfgs.push_back({32, 49152, 32});
    const double per_run = getenv("LAB_RUN_COST") ? atof(getenv("LAB_RUN_COST")) : 0.0;
    for (const Cfg &cf : cfgs)
    {
        const int H = cf.H, NB = cf.nb;
        int32_t Rb = 0;
        while (Rb < V && ptr[Rb + 1] - ptr[Rb] >= cf.tb) Rb++;
        const int64_t span = (int64_t)H * NB;
        std::vector<int32_t> cnt((size_t)NB * Rb, 0);
        std::vector<int32_t> main_adj;
        main_adj.reserve((size_t)E);
        for (int32_t r = 0; r < V; r++)
            for (int64_t p = ptr[r]; p < ptr[r + 1]; p++)
            {
                const int32_t v = adj[p];
                if (r < Rb && v < span && v != r) cnt[(size_t)(v / H) * Rb + r]++;
                else main_adj.push_back(v);
            }
        while (main_adj.size() % 4) main_adj.push_back(0);
        // layout: runs at multiples of ALIGN, pieces of long runs, bins padded to whole chunks
        std::vector<int32_t> bin_chunk0(NB + 1, 0);
        std::vector<int64_t> run_off((size_t)NB * Rb);
        std::vector<int64_t> flag_pos;
        std::vector<int32_t> run_slot, slot_of_extra;
        int32_t nslots = NB * Rb + 1;
        int64_t pos = 0, nruns_real = 0, moved = 0;
        for (int b = 0; b < NB; b++)
        {
            bin_chunk0[b] = (int32_t)(pos / CHUNK);
            for (int32_t r = 0; r < Rb; r++)
            {
                const int64_t n = cnt[(size_t)b * Rb + r];
                run_off[(size_t)b * Rb + r] = pos;
                if (n == 0) continue;
                nruns_real++;
                moved += n;
                for (int64_t o = 0; o < n; o += PIECE)
                {
                    flag_pos.push_back(pos + o);
                    if (o == 0) run_slot.push_back(b * Rb + r);
                    else
                    {
                        run_slot.push_back(nslots++);
                        slot_of_extra.push_back(b * Rb + r);
                    }
                }
                pos += (n + ALIGN - 1) / ALIGN * ALIGN;
            }
            flag_pos.push_back(pos);
            run_slot.push_back(NB * Rb);
            pos = (pos + ALIGN + CHUNK - 1) / CHUNK * CHUNK;
        }
        bin_chunk0[NB] = (int32_t)(pos / CHUNK);
        const int64_t W = pos;
        auto phys = [](int64_t q) { // logical position -> stored position (two 256-slot halves per 512-slot step)
            const int64_t st = q / STEP;
            const int i = (int)(q % STEP), ln = i / 16, sub = i % 16;
            return st * STEP + (sub / 8) * 256 + ln * 8 + (sub % 8);
        };
        std::vector<uint16_t> wcol((size_t)W, (uint16_t)H);
        {
            std::vector<int64_t> fillp(run_off);
            for (int32_t r = 0; r < Rb; r++)
                for (int64_t p = ptr[r]; p < ptr[r + 1]; p++)
                {
                    const int32_t v = adj[p];
                    if (v < span && v != r)
                    {
                        const int b = v / H;
                        wcol[(size_t)phys(fillp[(size_t)b * Rb + r]++)] = (uint16_t)(v - b * H);
                    }
                }
        }
        const int64_t nsteps = W / STEP;
        std::vector<uint32_t> meta((size_t)nsteps * 32, 0);
        std::vector<int32_t> step_run0((size_t)nsteps + 1, 0);
        {
            std::vector<uint8_t> lane_flags((size_t)nsteps * 32, 0);
            for (int64_t q : flag_pos) lane_flags[(size_t)(q / 16)] |= (uint8_t)(1u << ((q % 16) / ALIGN));
            int32_t runs = 0;
            for (int64_t st = 0; st < nsteps; st++)
            {
                step_run0[(size_t)st] = runs;
                int pre[32], total = 0, seg = -1, segs[32];
                for (int l = 0; l < 32; l++)
                {
                    pre[l] = total;
                    const int f = lane_flags[(size_t)st * 32 + l];
                    total += __builtin_popcount(f);
                    if (f) seg = l;
                    segs[l] = seg;
                }
                for (int l = 0; l < 32; l++)
                {
                    const int f = lane_flags[(size_t)st * 32 + l];
                    uint32_t m = (uint32_t)f | ((uint32_t)pre[l] << 4) | ((uint32_t)total << 18);
                    const int start = segs[l] < 0 ? 0 : segs[l];
                    for (int o = 0; o < 5; o++)
                        if (!f && l - (1 << o) >= start) m |= 1u << (12 + o);
                    if (l > 0 && segs[l - 1] >= 0) m |= 1u << 17;
                    meta[(size_t)st * 32 + l] = m;
                }
                runs += total;
            }
            step_run0[(size_t)nsteps] = runs;
        }
        printf("binned rows: degree >= %d -> %d rows; H %d x %d bins = columns < %lld; binned edges %lld (%.1f %%) in %lld slots, runs %lld (avg %.1f edges), left %zu edges\n",
               cf.tb, Rb, H, NB, (long long)span, (long long)moved, 100.0 * moved / E, (long long)W, (long long)nruns_real, (double)moved / nruns_real, main_adj.size());
        std::vector<int32_t> cta_chunk0(149, 0);
        {
            const size_t nch = (size_t)(W / CHUNK);
            std::vector<double> cum(nch + 1, 0.0);
            for (size_t k = 0; k < nch; k++) cum[k + 1] = cum[k] + 8.0 + per_run * (step_run0[(k + 1) * (CHUNK / STEP)] - step_run0[k * (CHUNK / STEP)]);
            for (int b = 0; b <= 148; b++)
                cta_chunk0[b] = (int32_t)(std::lower_bound(cum.begin(), cum.end(), cum[nch] * b / 148.0) - cum.begin());
            cta_chunk0[148] = (int32_t)nch;
        }
        uint16_t *d_wcol;
        uint32_t *d_meta;
        int32_t *d_step_run0, *d_run_slot, *d_bin_chunk0, *d_cta_chunk0;
        long long *d_cta_ns;
        CK(cudaMalloc(&d_wcol, (size_t)W * 2));
        CK(cudaMalloc(&d_meta, meta.size() * 4));
        CK(cudaMalloc(&d_step_run0, step_run0.size() * 4));
        CK(cudaMalloc(&d_run_slot, run_slot.size() * 4));
        CK(cudaMalloc(&d_bin_chunk0, bin_chunk0.size() * 4));
        CK(cudaMalloc(&d_cta_chunk0, 149 * 4));
        CK(cudaMalloc(&d_cta_ns, 148 * 8));
        CK(cudaMalloc(&d_slot, (size_t)nslots * 4));
        CK(cudaMemcpy(d_wcol, wcol.data(), (size_t)W * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_step_run0, step_run0.data(), step_run0.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_run_slot, run_slot.data(), run_slot.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_bin_chunk0, bin_chunk0.data(), bin_chunk0.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_cta_chunk0, cta_chunk0.data(), 149 * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_adj_main, main_adj.data(), main_adj.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(d_slot, 0, (size_t)nslots * 4));
        const int smem = (H + 4) * 4 + 32 * STAGE * 4;
        CK(cudaFuncSetAttribute(bin_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(bin_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        auto stat = [](long long *a) { long long mn = a[0], mx = a[0], sm = 0; for (int i = 0; i < 148; i++) { mn = std::min(mn, a[i]); mx = std::max(mx, a[i]); sm += a[i]; } printf(" [CTA us min %.0f avg %.0f max %.0f]", mn * 1e-3, sm * 1e-3 / 148, mx * 1e-3); };
        const float ms_bin = time_ms([&] {
            bin_kernel<true><<<148, 1024, smem>>>((const uint4 *)d_wcol, d_meta, d_step_run0, d_run_slot, d_bin_chunk0, d_cta_chunk0, NB, H, c, d_slot, d_out, d_cta_ns);
        });
        long long ns[148], ns2[148];
        CK(cudaMemcpy(ns, d_cta_ns, sizeof(ns), cudaMemcpyDeviceToHost));
        const float ms_bin_noemit = time_ms([&] {
            bin_kernel<false><<<148, 1024, smem>>>((const uint4 *)d_wcol, d_meta, d_step_run0, d_run_slot, d_bin_chunk0, d_cta_chunk0, NB, H, c, d_slot, d_out, d_cta_ns);
        });
        CK(cudaMemcpy(ns2, d_cta_ns, sizeof(ns2), cudaMemcpyDeviceToHost));
        const float ms_main = time_ms([&] { flat_kernel<2><<<148, 1024>>>((const int4 *)d_adj_main, (int64_t)main_adj.size() / 4, c, 49152, d_out); });
        // check the slots against the host
        std::vector<float> hc((size_t)V), hslot((size_t)nslots);
        CK(cudaMemcpy(hc.data(), c, (size_t)V * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hslot.data(), d_slot, hslot.size() * 4, cudaMemcpyDeviceToHost));
        double max_rel = 0;
        int64_t bad = 0;
        for (int b = 0; b < NB; b++)
            for (int32_t r = 0; r < Rb; r += 1)
            {
                const int64_t q0 = run_off[(size_t)b * Rb + r], n = cnt[(size_t)b * Rb + r];
                double s = 0;
                for (int64_t q = q0; q < q0 + n; q++) s += hc[(size_t)b * H + wcol[(size_t)phys(q)]];
                double got = hslot[(size_t)b * Rb + r];
                if (n > PIECE)
                    for (size_t x = 0; x < slot_of_extra.size(); x++)
                        if (slot_of_extra[x] == b * Rb + r) got += hslot[(size_t)NB * Rb + 1 + x];
                const double rel = fabs(got - s) / (fabs(s) + 1e-30);
                if (n == 0 ? got != 0.0 : rel > 1e-4) bad++;
                if (n > 0) max_rel = std::max(max_rel, rel);
            }
        printf("   run cost %.4f: bin kernel %8.3f ms (%7.1f Gedge/s)", per_run, ms_bin, moved / ms_bin * 1e-6);
        stat(ns);
        printf("; without the stores %8.3f ms", ms_bin_noemit);
        stat(ns2);
        printf("  + rest %8.3f ms  = %8.3f ms   [slots wrong: %lld, max rel %.2e]\n", ms_main, ms_bin + ms_main, (long long)bad, max_rel);
        printf("      CTA us (with stores):");
        for (int b = 0; b < 148; b += 6) printf(" %.0f", ns[b] * 1e-3);
        printf("\n");
        fflush(stdout);
        cudaFree(d_wcol); cudaFree(d_meta); cudaFree(d_step_run0); cudaFree(d_run_slot); cudaFree(d_bin_chunk0); cudaFree(d_cta_chunk0); cudaFree(d_slot); cudaFree(d_cta_ns);
    }
    return 0;
}
