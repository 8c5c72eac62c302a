// dsmem_lab.cu — microbenchmark: random 4-byte gathers from distributed shared memory of a thread-block cluster versus
// the same gathers from an L2-resident global array. Question: can a cluster-resident "hot" slice of the PageRank
// contribution vector (cluster x ~200 KB) serve gathers faster than the L1-miss path (1 line request / clk / SM)?
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

#define THREADS 512
#define SMEM_FLOATS (48 * 1024) // 192 KB per CTA

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int CLUSTER>
__global__ void __launch_bounds__(THREADS, 1) dsmem_gather(const float *__restrict__ src, float *out, int iters, int mode)
{
    extern __shared__ float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    for (int i = threadIdx.x; i < SMEM_FLOATS; i += THREADS) smem[i] = src[rank * SMEM_FLOATS + i];
    cluster.sync();
    const uint32_t total = CLUSTER * SMEM_FLOATS;
    float acc = 0.f;
    uint32_t h = hash32(blockIdx.x * THREADS + threadIdx.x + 1);
    for (int it = 0; it < iters; it++)
    {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            h = h * 1664525u + 1013904223u;
            const uint32_t idx = (h >> 8) % total;
            if (mode == 0)
            {
                const float *remote = cluster.map_shared_rank(smem, idx / SMEM_FLOATS);
                v[k] = remote[idx % SMEM_FLOATS];
            }
            else if (mode == 1) v[k] = smem[idx % SMEM_FLOATS];           // local shared memory only
            else v[k] = __ldg(src + idx);                                    // global (L2 / L1)
        }
#pragma unroll
        for (int k = 0; k < 8; k++) acc += v[k];
    }
    cluster.sync();
    if (acc == 12345.678f) out[0] = acc;
}

template <int CLUSTER>
static void run(const float *d_src, float *d_out, int sms)
{
    const int iters = 2000;
    cudaFuncSetAttribute(dsmem_gather<CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_FLOATS * 4);
    if (CLUSTER > 8) cudaFuncSetAttribute(dsmem_gather<CLUSTER>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int mode = 0; mode < 3; mode++)
    {
        cudaLaunchConfig_t cfg = {};
        int grid = (sms / CLUSTER) * CLUSTER;
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = SMEM_FLOATS * 4;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = CLUSTER; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        int maxc = 0;
        cudaOccupancyMaxActiveClusters(&maxc, dsmem_gather<CLUSTER>, &cfg);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaError_t err = cudaLaunchKernelEx(&cfg, dsmem_gather<CLUSTER>, d_src, d_out, 10, mode);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        err = cudaLaunchKernelEx(&cfg, dsmem_gather<CLUSTER>, d_src, d_out, iters, mode);
        cudaEventRecord(e1);
        cudaError_t e2 = cudaDeviceSynchronize();
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gathers = (double)grid * THREADS * iters * 8;
        printf("cluster %2d mode %s: grid %3d (max active clusters %d) %8.3f ms  %7.1f G gathers/s  = %.2f gathers/clk/SM @1.9GHz  [%s %s]\n", CLUSTER,
               mode == 0 ? "dsmem " : mode == 1 ? "local " : "global", grid, maxc, ms, gathers / ms / 1e6, gathers / (ms * 1e-3) / grid / 1.9e9,
               cudaGetErrorString(err), cudaGetErrorString(e2));
    }
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const size_t n = (size_t)16 * SMEM_FLOATS;
    float *d_src, *d_out;
    cudaMalloc(&d_src, n * 4);
    cudaMalloc(&d_out, 4);
    cudaMemset(d_src, 0, n * 4);
    run<2>(d_src, d_out, prop.multiProcessorCount);
    run<4>(d_src, d_out, prop.multiProcessorCount);
    run<8>(d_src, d_out, prop.multiProcessorCount);
    run<16>(d_src, d_out, prop.multiProcessorCount);
    return 0;
}
