// tmp.cuh -- stream-ordered temporaries (cudaMallocAsync pool; retained between calls).
#pragma once
#include "common.cuh"

template <typename T>
struct DevTmp {
    T *p = nullptr;
    cudaStream_t s;
    cudaError_t err = cudaSuccess;
    DevTmp(euler_ctx *ctx, size_t n) : s(ctx->stream)
    {
        err = cudaMallocAsync((void **)&p, (n ? n : 1) * sizeof(T), s);
        if (err != cudaSuccess) p = nullptr;
    }
    ~DevTmp()
    {
        if (p) cudaFreeAsync(p, s);
    }
    DevTmp(const DevTmp &) = delete;
    DevTmp &operator=(const DevTmp &) = delete;
    operator T *() const { return p; }
    T *get() const { return p; }
    bool ok() const { return p != nullptr; }
};

#define TMP_CHECK(ctx, t)                                                                          \
    do {                                                                                           \
        if (!(t).ok()) return euler_fail((ctx), EULER_ERR_NOMEM, "%s:%d device temp alloc failed: %s", __FILE__, \
                                         __LINE__, cudaGetErrorString((t).err));                   \
    } while (0)

// read one u64 from device (synchronises the stream)
static inline int read_u64(euler_ctx *ctx, const u64 *d, u64 *h)
{
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pinned, d, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *h = ctx->h_pinned[0];
    return EULER_OK;
}
static inline int read_u64s(euler_ctx *ctx, const u64 *d, u64 *h, int n)
{
    if (n < 0 || n > EULER_PINNED_WORDS) return euler_fail(ctx, EULER_ERR_ARG, "read_u64s: %d words exceed the pinned buffer", n);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pinned, d, n * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) h[i] = ctx->h_pinned[i];
    return EULER_OK;
}
