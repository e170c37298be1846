// scan.cuh -- single-pass exclusive prefix sum (decoupled look-back).  Replaces the
// pycuda.scan.ExclusiveScanKernel call sites of the reference (pygpuhash.py:290,
// pydebruijn.py:560-573, pyeulertour.py:748,774).
//
// Layout: a tile is 8 warps x 16 rows x 32 lanes.  Lane i of a warp owns items row*32 + i of the
// warp's 512-item segment, so every load and store is a fully coalesced 128 B (or 256 B) row; the
// in-warp scan is a shuffle scan per row with the row total carried forward.
//
// A Policy supplies the items and receives the prefixes, so weights are computed on the fly
// (table occupancy, strand multiplicity) and results can be written in any fused form:
//   typedef T                                    u32 (one sum per tile < 2^32) or u64 (two u32 sums packed hi|lo)
//   __device__ T    load(u64 idx) const          value of item idx
//   __device__ void store(u64 idx, u64 excl, u64 value, bool valid) const
// store() is called by all 32 lanes of a row together (valid == false past the end), so it may use
// warp shuffles.
#pragma once
#include "common.cuh"

#define SCAN_THREADS 256
#define SCAN_WARPS (SCAN_THREADS / 32)
#define SCAN_ROWS 16
#define SCAN_TILE (SCAN_THREADS * SCAN_ROWS)

#define SCAN_ST_AGG 1ull
#define SCAN_ST_INC 2ull

struct __align__(16) ScanState {
    u64 flag;
    u64 value;
};

__device__ __forceinline__ ScanState ld_state(const ScanState *p)
{
    ScanState s;
    asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(s.flag), "=l"(s.value) : "l"(p));
    return s;
}
__device__ __forceinline__ void st_state(ScanState *p, u64 flag, u64 value)
{
    asm volatile("st.volatile.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(flag), "l"(value) : "memory");
}

// Decoupled look-back, run by one full warp of the block that owns `tile`: publishes the tile's
// aggregate, sums the predecessors' states back to the nearest inclusive prefix, publishes the
// inclusive prefix and returns the exclusive one (all lanes).  The last tile writes *total.
__device__ __forceinline__ u64 scan_lookback(ScanState *state, u64 tile, u64 block_sum, int lane, u64 ntiles, u64 *total)
{
    u64 prefix = 0;
    if (tile == 0) {
        if (lane == 0) st_state(state, SCAN_ST_INC, block_sum);
    } else {
        if (lane == 0) st_state(state + tile, SCAN_ST_AGG, block_sum);
        long long look = (long long)tile - 1;
        while (true) {
            const long long idx = look - lane;
            ScanState s;
            if (idx >= 0) {
                do { s = ld_state(state + idx); } while (s.flag == 0);
            } else {
                s.flag = SCAN_ST_INC;  // virtual tile before tile 0
                s.value = 0;
            }
            const unsigned inc_mask = __ballot_sync(0xffffffffu, s.flag == SCAN_ST_INC);
            u64 val = s.value;
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
        if (lane == 0) st_state(state + tile, SCAN_ST_INC, prefix + block_sum);
    }
    if (lane == 0 && tile == ntiles - 1 && total) *total = prefix + block_sum;
    return prefix;
}

// state[0..ntiles): tile descriptors; counter: dynamic tile index; all zero on entry.
template <typename P>
__global__ void __launch_bounds__(SCAN_THREADS, 4) scan_exclusive_kernel(P p, u64 n, ScanState *state, u64 *counter,
                                                                        u64 ntiles, u64 *total)
{
    __shared__ u64 s_tile;
    __shared__ u64 s_warp[SCAN_WARPS];
    __shared__ u64 s_prefix;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) s_tile = atomicAdd(counter, 1ull);
    __syncthreads();
    const u64 tile = s_tile;
    typedef typename P::T T;
    constexpr int ROWS = 16;
    const u64 base = tile * (u64)(SCAN_THREADS * ROWS) + (u64)warp * (32 * ROWS) + lane;
  // u32 for plain sums, u64 for packed pairs: halves the register footprint
    T v[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const u64 idx = base + (u64)r * 32;
        v[r] = idx < n ? p.load(idx) : (T)0;
    }
    // warp total of the segment (the per-row prefixes are recomputed in the store pass instead of being
    // kept: 16 fewer live registers per thread, which buys a whole extra resident block)
    u64 carry = 0;   // 64-bit even for the u32 policies: a 512-item segment of large weights must not wrap before it is widened
#pragma unroll
    for (int r = 0; r < ROWS; r++) carry += v[r];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) carry += __shfl_xor_sync(0xffffffffu, carry, d);
    if (lane == 0) s_warp[warp] = carry;
    __syncthreads();
    u64 warp_off = 0, block_sum = 0;
#pragma unroll
    for (int w = 0; w < SCAN_WARPS; w++) {
        const u64 t = s_warp[w];
        if (w < warp) warp_off += t;
        block_sum += t;
    }

    // decoupled look-back by warp 0
    if (warp == 0) {
        const u64 prefix = scan_lookback(state, tile, block_sum, lane, ntiles, total);
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    u64 off = s_prefix + warp_off;
#pragma unroll
    for (int r = 0; r < ROWS; r++) {
        const u64 idx = base + (u64)r * 32;
        T inc = v[r];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const T t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        p.store(idx, off + inc - v[r], v[r], idx < n);
        off += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// ---- stock policies ----------------------------------------------------------------------------
// u32 array in -> u32 exclusive prefix out (the ExclusiveScanKernel(uintc,"a+b",0) call sites)
struct ScanU32 {
    typedef u32 T;
    const u32 *in;
    u32 *out;
    __device__ __forceinline__ u32 load(u64 i) const { return in[i]; }
    __device__ __forceinline__ void store(u64 i, u64 ex, u64, bool valid) const
    {
        if (valid) out[i] = (u32)ex;
    }
};
// weight functor in -> u32 prefix out
template <typename F>
struct ScanFn {
    typedef u32 T;
    F f;
    u32 *out;
    __device__ __forceinline__ u32 load(u64 i) const { return f(i); }
    __device__ __forceinline__ void store(u64 i, u64 ex, u64, bool valid) const
    {
        if (valid) out[i] = (u32)ex;
    }
};

int scan_state_reserve(euler_ctx *ctx, u64 ntiles, ScanState **state, u64 **counter);

// `total` (device u64, may be NULL) receives the grand total.
template <typename P>
static int scan_run(euler_ctx *ctx, P p, u64 n, u64 *d_total)
{
    if (n == 0) {
        if (d_total) CUDA_TRY(ctx, cudaMemsetAsync(d_total, 0, sizeof(u64), ctx->stream));
        return EULER_OK;
    }
    constexpr u64 tile_items = (u64)SCAN_THREADS * 16;
    const u64 ntiles = (n + tile_items - 1) / tile_items;
    ScanState *state = nullptr;
    u64 *counter = nullptr;
    EULER_TRY(scan_state_reserve(ctx, ntiles, &state, &counter));
    CUDA_TRY(ctx, cudaMemsetAsync(state, 0, (ntiles + 1) * sizeof(ScanState), ctx->stream));
    scan_exclusive_kernel<P><<<(unsigned)ntiles, SCAN_THREADS, 0, ctx->stream>>>(p, n, state, counter, ntiles, d_total);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// compatibility helpers used across the library
struct ScanInU32 {
    const u32 *p;
    __device__ __forceinline__ u32 operator()(u64 i) const { return p[i]; }
};
template <typename F>
static int scan_exclusive(euler_ctx *ctx, F f, u64 n, u32 *d_out, u64 *d_total)
{
    return scan_run(ctx, ScanFn<F>{f, d_out}, n, d_total);
}
