// scan.cuh -- single-pass exclusive prefix sum (decoupled look-back), u32 items -> u32 prefix,
// u64 grand total.  Replaces the pycuda.scan.ExclusiveScanKernel call sites of the reference
// (pygpuhash.py:290, pydebruijn.py:560-573, pyeulertour.py:748,774).
//
// The input is a functor so per-slot weights (table occupancy, palindrome-aware strand counts)
// are computed on the fly instead of being materialised: one read of the source, one write of
// the prefix.
#pragma once
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_ITEMS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

#define SCAN_FLAG_AGG (1ull << 62)
#define SCAN_FLAG_INC (2ull << 62)
#define SCAN_VAL_MASK ((1ull << 62) - 1)

struct ScanInU32 {
    const u32 *p;
    __device__ __forceinline__ u32 operator()(u64 i) const { return p[i]; }
};

__device__ __forceinline__ u64 ld_volatile_u64(const u64 *p)
{
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u64(u64 *p, u64 v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// state[0..ntiles) tile descriptors, state[ntiles] dynamic tile counter; all zero on entry.
template <typename InFn>
__global__ void __launch_bounds__(SCAN_THREADS) scan_exclusive_kernel(InFn in, u64 n, u32 *out, u64 *state,
                                                                        u64 ntiles, u64 *total)
{
    __shared__ u64 s_tile;
    __shared__ u64 s_warp[SCAN_THREADS / 32];
    __shared__ u64 s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) s_tile = atomicAdd(state + ntiles, 1ull);
    __syncthreads();
    const u64 tile = s_tile;
    const u64 base = tile * SCAN_TILE + (u64)tid * SCAN_ITEMS;

    u32 v[SCAN_ITEMS];
    u64 tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const u64 idx = base + i;
        v[i] = idx < n ? in(idx) : 0u;
        tsum += v[i];
    }
    // block exclusive scan of per-thread sums
    u64 incl = tsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    u64 warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        const u64 t = s_warp[w];
        if (w < warp) warp_off += t;
        block_sum += t;
    }
    const u64 thread_excl = warp_off + incl - tsum;

    // decoupled look-back by warp 0
    if (warp == 0) {
        u64 prefix = 0;
        if (tile == 0) {
            if (lane == 0) st_volatile_u64(state + 0, SCAN_FLAG_INC | block_sum);
        } else {
            if (lane == 0) st_volatile_u64(state + tile, SCAN_FLAG_AGG | block_sum);
            long long look = (long long)tile - 1;
            while (true) {
                const long long idx = look - lane;
                u64 s;
                if (idx >= 0) {
                    do { s = ld_volatile_u64(state + idx); } while ((s >> 62) == 0);
                } else {
                    s = SCAN_FLAG_INC;  // virtual tile before tile 0: inclusive 0
                }
                const unsigned inc_mask = __ballot_sync(0xffffffffu, (s >> 62) == 2);
                u64 val = s & SCAN_VAL_MASK;
                if (inc_mask) {
                    const int first = __ffs(inc_mask) - 1;
                    if (lane > first) val = 0;
                }
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
                prefix += val;
                if (inc_mask) break;
                look -= 32;
            }
            if (lane == 0) st_volatile_u64(state + tile, SCAN_FLAG_INC | ((prefix + block_sum) & SCAN_VAL_MASK));
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (tile == ntiles - 1 && total) *total = prefix + block_sum;
        }
    }
    __syncthreads();
    u64 run = s_prefix + thread_excl;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const u64 idx = base + i;
        if (idx < n) out[idx] = (u32)run;
        run += v[i];
    }
}

// host launcher: d_total (device u64) receives the grand total; state scratch is managed by ctx.
int scan_state_reserve(euler_ctx *ctx, u64 ntiles, u64 **state);

template <typename InFn>
static int scan_exclusive(euler_ctx *ctx, InFn in, u64 n, u32 *d_out, u64 *d_total)
{
    if (n == 0) {
        if (d_total) CUDA_TRY(ctx, cudaMemsetAsync(d_total, 0, sizeof(u64), ctx->stream));
        return EULER_OK;
    }
    const u64 ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    u64 *state = nullptr;
    EULER_TRY(scan_state_reserve(ctx, ntiles, &state));
    CUDA_TRY(ctx, cudaMemsetAsync(state, 0, (ntiles + 1) * sizeof(u64), ctx->stream));
    scan_exclusive_kernel<InFn><<<(unsigned)ntiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, d_out, state, ntiles, d_total);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
