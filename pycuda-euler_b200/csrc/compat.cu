// compat.cu -- per-step entry points kept for callers of the reference's fine-grained wrappers:
// the three bucketed-table phases of pygpuhash (phase1 :37-52, copyToBucket :95-127, bucketSort
// :187-232) and the ten Shiloach-Vishkin sub-steps of pycomponent (:36-654).  The product path
// uses the open-addressing table and the union-find components instead; these kernels restate
// each step's semantics for sm_100a so the step-level API stays usable.
#include "kernels.h"
#include "tmp.cuh"

#define CB 256
#define C0 0x01010101ull
#define C1 0x12345678ull
#define LARGE_PRIME 1900813ull
#define MAX_BUCKET_ITEM 520

__device__ __forceinline__ u32 hash_h(u64 key, u32 bucketCount) { return (u32)(((C0 + C1 * key) % LARGE_PRIME) % bucketCount); }

__global__ void __launch_bounds__(CB) compat_phase1_kernel(const u64 *__restrict__ keys, u32 *__restrict__ offset, u64 n,
                                                            u32 *count, u32 bucketCount)
{
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;  // (B4: the reference's floor grid drops the tail; every key is processed here)
    offset[t] = atomicInc(count + hash_h(keys[t], bucketCount), 0xffffffffu);
}
__global__ void __launch_bounds__(CB) compat_copy_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ values,
                                                          const u32 *__restrict__ offset, u64 n, const u32 *__restrict__ start,
                                                          u32 bucketCount, u64 *__restrict__ bufK, u32 *__restrict__ bufV)
{
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const u64 key = keys[t];
    const u32 index = start[hash_h(key, bucketCount)] + offset[t];
    bufK[index] = key;
    bufV[index] = values[t];
}
// one warp per bucket, rank = number of smaller keys (keys are distinct)
__global__ void __launch_bounds__(32) compat_bucket_sort_kernel(const u64 *__restrict__ bufK, const u32 *__restrict__ bufV,
                                                                 const u32 *__restrict__ start, const u32 *__restrict__ bucketSize,
                                                                 u64 *__restrict__ TK, u32 *__restrict__ TV)
{
    __shared__ u64 keys[MAX_BUCKET_ITEM];
    const u32 b = blockIdx.x;
    const u32 off = start[b];
    u32 size = bucketSize[b];
    if (size > MAX_BUCKET_ITEM) size = MAX_BUCKET_ITEM;  // B5: overflow beyond 520 is dropped, not written out of bounds
    for (u32 i = threadIdx.x; i < size; i += 32) keys[i] = bufK[off + i];
    __syncwarp();
    for (u32 i = threadIdx.x; i < size; i += 32) {
        const u64 k = keys[i];
        u32 rank = 0;
        for (u32 j = 0; j < size; j++) rank += keys[j] < k;
        TK[(u64)b * MAX_BUCKET_ITEM + rank] = k;
        TV[(u64)b * MAX_BUCKET_ITEM + rank] = bufV[off + i];
    }
}

// ---- Shiloach-Vishkin sub-steps ---------------------------------------------------------------
struct CcArgs {
    const euler_succ_vertex *v;
    u32 *prevD, *D, *Q, *t1, *val1, *t2, *val2, *flag;
    u32 n, s;
};
__global__ void __launch_bounds__(CB) compat_cc_kernel(int step, CcArgs a)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 n = a.n;
    if (t >= n) return;
    switch (step) {
    case 0: a.D[t] = t; a.Q[t] = 0; break;                                              // :42-43
    case 1: a.D[t] = a.prevD[a.prevD[t]]; break;                                         // :94
    case 2: if (a.D[t] != a.prevD[t]) a.Q[a.D[t]] = a.s; break;                          // :155-158
    case 3: {                                                                            // :222-242
        a.t1[t] = n; a.t2[t] = n;
        const u32 n1 = a.v[t].n1, n2 = a.v[t].n2;
        if (a.D[t] == a.prevD[t] && n1 < n && a.D[n1] < a.D[t]) { a.t1[t] = a.D[t]; a.val1[t] = a.D[n1]; }
        if (a.D[t] == a.prevD[t] && n2 < n && a.D[n2] < a.D[t]) { a.t2[t] = a.D[t]; a.val2[t] = a.D[n2]; }
        break;
    }
    case 4:                                                                              // :316-332
        if (a.t1[t] < n) { atomicMin(a.D + a.t1[t], a.val1[t]); atomicExch(a.Q + a.val1[t], a.s); }
        if (a.t2[t] < n) { atomicMin(a.D + a.t2[t], a.val2[t]); atomicExch(a.Q + a.val2[t], a.s); }
        break;
    case 5: {                                                                            // :402-416
        a.t1[t] = n; a.t2[t] = n;
        const u32 d = a.D[t], n1 = a.v[t].n1, n2 = a.v[t].n2;
        const bool stagnant_root = d == a.D[d] && a.Q[d] < a.s;
        if (stagnant_root && n1 < n && d != a.D[n1]) { a.t1[t] = d; a.val1[t] = a.D[n1]; }
        if (stagnant_root && n2 < n && d != a.D[n2]) { a.t2[t] = d; a.val2[t] = a.D[n2]; }
        break;
    }
    case 6:                                                                              // :486-497
        if (a.t1[t] < n) atomicMin(a.D + a.t1[t], a.val1[t]);
        if (a.t2[t] < n) atomicMin(a.D + a.t2[t], a.val2[t]);
        break;
    case 7: a.val1[t] = a.D[a.D[t]]; break;                                               // :558
    case 8: a.D[t] = a.val1[t]; break;                                                   // :606
    case 9: if (a.Q[t] == a.s) atomicExch(a.flag, 1u); break;                            // :649-651
    }
}

template <typename T>
static int up(euler_ctx *ctx, DevTmp<T> &d, const T *h, u64 n)
{
    if (!d.ok()) return euler_fail(ctx, EULER_ERR_NOMEM, "device temp alloc failed");
    if (n && h) CUDA_TRY(ctx, cudaMemcpyAsync(d.get(), h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    else if (n) CUDA_TRY(ctx, cudaMemsetAsync(d.get(), 0, n * sizeof(T), ctx->stream));
    return EULER_OK;
}
template <typename T>
static int down(euler_ctx *ctx, T *h, const T *d, u64 n)
{
    if (n && h) CUDA_TRY(ctx, cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    return EULER_OK;
}

extern "C" {

int euler_compat_phase1(euler_ctx *ctx, const uint64_t *keys, uint64_t n, uint32_t bucketCount, uint32_t *offset,
                        uint32_t *count)
{
    if (!ctx || !bucketCount) return EULER_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    DevTmp<u64> dk(ctx, n);
    DevTmp<u32> doff(ctx, n), dcnt(ctx, bucketCount);
    EULER_TRY(up(ctx, dk, (const u64 *)keys, n));
    EULER_TRY(up(ctx, doff, (const u32 *)nullptr, n));
    EULER_TRY(up(ctx, dcnt, (const u32 *)count, bucketCount));
    if (n) compat_phase1_kernel<<<grid_for(n, CB), CB, 0, ctx->stream>>>(dk, doff, n, dcnt, bucketCount);
    CUDA_TRY(ctx, cudaGetLastError());
    EULER_TRY(down(ctx, offset, doff.get(), n));
    EULER_TRY(down(ctx, count, dcnt.get(), bucketCount));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return EULER_OK;
}

int euler_compat_copy_to_bucket(euler_ctx *ctx, const uint64_t *keys, const uint32_t *values, const uint32_t *offset,
                                uint64_t n, const uint32_t *start, uint32_t bucketCount, uint64_t *bufferK, uint32_t *bufferV)
{
    if (!ctx || !bucketCount) return EULER_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    DevTmp<u64> dk(ctx, n), dbk(ctx, n);
    DevTmp<u32> dv(ctx, n), doff(ctx, n), dst(ctx, bucketCount), dbv(ctx, n);
    EULER_TRY(up(ctx, dk, (const u64 *)keys, n)); EULER_TRY(up(ctx, dv, values, n)); EULER_TRY(up(ctx, doff, offset, n));
    EULER_TRY(up(ctx, dst, start, bucketCount));
    EULER_TRY(up(ctx, dbk, (const u64 *)nullptr, n)); EULER_TRY(up(ctx, dbv, (const u32 *)nullptr, n));
    if (n) compat_copy_kernel<<<grid_for(n, CB), CB, 0, ctx->stream>>>(dk, dv, doff, n, dst, bucketCount, dbk, dbv);
    CUDA_TRY(ctx, cudaGetLastError());
    EULER_TRY(down(ctx, (u64 *)bufferK, dbk.get(), n));
    EULER_TRY(down(ctx, bufferV, dbv.get(), n));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return EULER_OK;
}

int euler_compat_bucket_sort(euler_ctx *ctx, const uint64_t *bufferK, const uint32_t *bufferV, uint64_t n,
                             const uint32_t *start, const uint32_t *bucketSize, uint32_t bucketCount, uint64_t *TK,
                             uint32_t *TV)
{
    if (!ctx || !bucketCount) return EULER_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const u64 tl = (u64)bucketCount * MAX_BUCKET_ITEM;
    DevTmp<u64> dbk(ctx, n), dTK(ctx, tl);
    DevTmp<u32> dbv(ctx, n), dst(ctx, bucketCount), dsz(ctx, bucketCount), dTV(ctx, tl);
    EULER_TRY(up(ctx, dbk, (const u64 *)bufferK, n)); EULER_TRY(up(ctx, dbv, bufferV, n));
    EULER_TRY(up(ctx, dst, start, bucketCount)); EULER_TRY(up(ctx, dsz, bucketSize, bucketCount));
    EULER_TRY(up(ctx, dTK, (const u64 *)nullptr, tl)); EULER_TRY(up(ctx, dTV, (const u32 *)nullptr, tl));
    compat_bucket_sort_kernel<<<bucketCount, 32, 0, ctx->stream>>>(dbk, dbv, dst, dsz, dTK, dTV);
    CUDA_TRY(ctx, cudaGetLastError());
    EULER_TRY(down(ctx, (u64 *)TK, dTK.get(), tl));
    EULER_TRY(down(ctx, TV, dTV.get(), tl));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return EULER_OK;
}

// step: 0 init, 1 s1p1, 2 s1p2, 3 s2p1, 4 s2p2, 5 s3p1, 6 s3p2, 7 s4p1, 8 s4p2, 9 s5.
// Every array is in/out (NULL = zero-filled scratch); flag is one u32.
int euler_compat_cc_step(euler_ctx *ctx, int step, uint32_t n, uint32_t s, const euler_succ_vertex *v, uint32_t *prevD,
                         uint32_t *D, uint32_t *Q, uint32_t *t1, uint32_t *val1, uint32_t *t2, uint32_t *val2,
                         uint32_t *flag)
{
    if (!ctx || step < 0 || step > 9) return EULER_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!n) return EULER_OK;
    DevTmp<euler_succ_vertex> dv(ctx, n);
    DevTmp<u32> dp(ctx, n), dD(ctx, n), dQ(ctx, n), d1(ctx, n), dv1(ctx, n), d2(ctx, n), dv2(ctx, n), df(ctx, 1);
    EULER_TRY(up(ctx, dv, v, n)); EULER_TRY(up(ctx, dp, (const u32 *)prevD, n)); EULER_TRY(up(ctx, dD, (const u32 *)D, n));
    EULER_TRY(up(ctx, dQ, (const u32 *)Q, n)); EULER_TRY(up(ctx, d1, (const u32 *)t1, n));
    EULER_TRY(up(ctx, dv1, (const u32 *)val1, n)); EULER_TRY(up(ctx, d2, (const u32 *)t2, n));
    EULER_TRY(up(ctx, dv2, (const u32 *)val2, n)); EULER_TRY(up(ctx, df, (const u32 *)flag, 1));
    CcArgs a = {dv, dp, dD, dQ, d1, dv1, d2, dv2, df, n, s};
    compat_cc_kernel<<<grid_for(n, CB), CB, 0, ctx->stream>>>(step, a);
    CUDA_TRY(ctx, cudaGetLastError());
    EULER_TRY(down(ctx, D, dD.get(), n)); EULER_TRY(down(ctx, Q, dQ.get(), n)); EULER_TRY(down(ctx, t1, d1.get(), n));
    EULER_TRY(down(ctx, val1, dv1.get(), n)); EULER_TRY(down(ctx, t2, d2.get(), n)); EULER_TRY(down(ctx, val2, dv2.get(), n));
    EULER_TRY(down(ctx, flag, df.get(), 1));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return EULER_OK;
}

}  // extern "C"
