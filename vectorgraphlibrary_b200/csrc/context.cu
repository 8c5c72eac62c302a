// context.cu — runtime layer of libvgl_b200: device selection, stream, HBM allocations, error reporting.
// Replaces VGL_RUNTIME::init_library / select_device (vgl_runtime/vgl_runtime.hpp:5-16,
// helpers/gpu_API/select_device.cuh:5-8) and MemoryAPI (helpers/memory_API/memory_API.hpp:3-101): explicit
// cudaMalloc'ed HBM instead of cudaMallocManaged + prefetch hints, one stream per context instead of the
// reference's six streams + cudaDeviceSynchronize after every operator.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

static thread_local char g_error[512] = "";

void vglb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

extern "C" const char *vglb_last_error(void) { return g_error; }

extern "C" int vglb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int vglb_init(int device, vglb_ctx **out_ctx)
{
    VGLB_REQUIRE(out_ctx != NULL, "vglb_init: out_ctx is NULL");
    int n = vglb_device_count();
    if (n <= 0 || device < 0 || device >= n)
    {
        vglb_set_error("vglb_init: no usable CUDA device %d (found %d); libvgl_b200 has no CPU fallback", device, n);
        return VGLB_ENODEVICE;
    }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    vglb_ctx *ctx = (vglb_ctx *)calloc(1, sizeof(vglb_ctx));
    if (!ctx) return VGLB_ENOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreate(&ctx->ev_start));
    CUDA_TRY(cudaEventCreate(&ctx->ev_stop));
    CUDA_TRY(cudaMallocHost((void **)&ctx->h_counters, 64 * sizeof(int64_t)));
    CUDA_TRY(cudaMalloc((void **)&ctx->d_counters, 64 * sizeof(int64_t)));
    CUDA_TRY(cudaMemset(ctx->d_counters, 0, 64 * sizeof(int64_t)));
    ctx->flush_bytes = ctx->l2_bytes * 2 > (size_t)(256u << 20) ? ctx->l2_bytes * 2 : (size_t)(256u << 20);
    ctx->d_flush = NULL;
    *out_ctx = ctx;
    return VGLB_OK;
}

extern "C" int vglb_finalize(vglb_ctx *ctx)
{
    if (!ctx) return VGLB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->d_flush) cudaFree(ctx->d_flush);
    cudaFree(ctx->d_counters);
    cudaFreeHost(ctx->h_counters);
    cudaEventDestroy(ctx->ev_start);
    cudaEventDestroy(ctx->ev_stop);
    cudaStreamDestroy(ctx->stream);
    free(ctx);
    return VGLB_OK;
}

extern "C" int vglb_synchronize(vglb_ctx *ctx)
{
    VGLB_REQUIRE(ctx != NULL, "vglb_synchronize: ctx is NULL");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VGLB_OK;
}

extern "C" void *vglb_stream(vglb_ctx *ctx) { return ctx ? (void *)ctx->stream : NULL; }

extern "C" int vglb_malloc(vglb_ctx *ctx, size_t bytes, void **d_ptr)
{
    VGLB_REQUIRE(ctx != NULL && d_ptr != NULL, "vglb_malloc: NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 16);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        vglb_set_error("vglb_malloc: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        return VGLB_ENOMEM;
    }
    return VGLB_OK;
}

extern "C" int vglb_free(vglb_ctx *ctx, void *d_ptr)
{
    VGLB_REQUIRE(ctx != NULL, "vglb_free: ctx is NULL");
    if (d_ptr) CUDA_TRY(cudaFree(d_ptr));
    return VGLB_OK;
}

extern "C" int vglb_memcpy_h2d(vglb_ctx *ctx, void *d_dst, const void *h_src, size_t bytes)
{
    VGLB_REQUIRE(ctx != NULL, "vglb_memcpy_h2d: ctx is NULL");
    CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VGLB_OK;
}

extern "C" int vglb_memcpy_d2h(vglb_ctx *ctx, void *h_dst, const void *d_src, size_t bytes)
{
    VGLB_REQUIRE(ctx != NULL, "vglb_memcpy_d2h: ctx is NULL");
    CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return VGLB_OK;
}

extern "C" int vglb_memcpy_d2d(vglb_ctx *ctx, void *d_dst, const void *d_src, size_t bytes)
{
    VGLB_REQUIRE(ctx != NULL, "vglb_memcpy_d2d: ctx is NULL");
    CUDA_TRY(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return VGLB_OK;
}

extern "C" int vglb_memset(vglb_ctx *ctx, void *d_dst, int byte_value, size_t bytes)
{
    VGLB_REQUIRE(ctx != NULL, "vglb_memset: ctx is NULL");
    CUDA_TRY(cudaMemsetAsync(d_dst, byte_value, bytes, ctx->stream));
    return VGLB_OK;
}

extern "C" int vglb_host_alloc_pinned(size_t bytes, void **h_ptr)
{
    VGLB_REQUIRE(h_ptr != NULL, "vglb_host_alloc_pinned: NULL argument");
    cudaError_t e = cudaMallocHost(h_ptr, bytes ? bytes : 16);
    if (e != cudaSuccess)
    {
        cudaGetLastError();
        vglb_set_error("vglb_host_alloc_pinned(%zu): %s", bytes, cudaGetErrorString(e));
        return VGLB_ENOMEM;
    }
    return VGLB_OK;
}

extern "C" int vglb_host_free_pinned(void *h_ptr)
{
    if (h_ptr) CUDA_TRY(cudaFreeHost(h_ptr));
    return VGLB_OK;
}

__global__ void flush_l2_kernel(uint4 *buf, size_t n16, uint32_t tag)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) buf[i] = make_uint4(tag, tag + 1, tag + 2, tag + 3);
}

extern "C" int vglb_flush_l2(vglb_ctx *ctx)
{
    VGLB_REQUIRE(ctx != NULL, "vglb_flush_l2: ctx is NULL");
    if (!ctx->d_flush) CUDA_TRY(cudaMalloc(&ctx->d_flush, ctx->flush_bytes));
    static uint32_t tag = 1;
    flush_l2_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint4 *)ctx->d_flush, ctx->flush_bytes / 16, tag++);
    KERNEL_TRY();
    return VGLB_OK;
}

// ---- synthetic inputs --------------------------------------------------------------------------------------------

__global__ void generate_edges_kernel(int kind, int scale, int64_t edges, uint64_t seed, int a, int b, int c,
                                      int32_t *__restrict__ src, int32_t *__restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < edges; i += stride)
    {
        int32_t s, d;
        vglb_gen_edge(kind, scale, seed, (uint64_t)i, a, b, c, &s, &d);
        src[i] = s;
        dst[i] = d;
    }
}

extern "C" int vglb_generate_edges_device(vglb_ctx *ctx, int kind, int scale, int64_t edges, uint64_t seed, int a,
                                          int b, int c, int32_t *d_src, int32_t *d_dst)
{
    VGLB_REQUIRE(ctx != NULL && d_src != NULL && d_dst != NULL, "vglb_generate_edges_device: NULL argument");
    VGLB_REQUIRE(scale >= 1 && scale <= 30 && edges >= 0, "vglb_generate_edges_device: bad scale/edges");
    VGLB_REQUIRE(kind >= 0 && kind <= 2 && a > 0 && b >= 0 && c >= 0 && a + b + c < 100,
                 "vglb_generate_edges_device: bad kind or probabilities");
    if (edges == 0) return VGLB_OK;
    generate_edges_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(kind, scale, edges, seed, a, b, c, d_src, d_dst);
    KERNEL_TRY();
    ctx->launches++;
    return VGLB_OK;
}

extern "C" int vglb_generate_edges_host(int kind, int scale, int64_t edges, uint64_t seed, int a, int b, int c,
                                        int32_t *h_src, int32_t *h_dst)
{
    VGLB_REQUIRE(h_src != NULL && h_dst != NULL, "vglb_generate_edges_host: NULL argument");
    VGLB_REQUIRE(scale >= 1 && scale <= 30 && edges >= 0, "vglb_generate_edges_host: bad scale/edges");
    VGLB_REQUIRE(kind >= 0 && kind <= 2 && a > 0 && b >= 0 && c >= 0 && a + b + c < 100,
                 "vglb_generate_edges_host: bad kind or probabilities");
    for (int64_t i = 0; i < edges; i++) vglb_gen_edge(kind, scale, seed, (uint64_t)i, a, b, c, &h_src[i], &h_dst[i]);
    return VGLB_OK;
}
