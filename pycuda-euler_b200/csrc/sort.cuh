// sort.cuh -- stable LSD radix sort of u64 keys with an optional u32 payload (8-bit digits,
// warp match-any ranking).  Used for canonical (ascending-key) ids, circuit-edge ordering
// (the reference's host np.sort, pyeulertour.py:792) and contig ordering; not on the timed
// encode+hash+graph path unless EULER_RUN_CANONICAL_IDS is requested.
#pragma once
#include "common.cuh"
#include "scan.cuh"

#define RS_THREADS 256
#define RS_WARPS (RS_THREADS / 32)
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)

__device__ __forceinline__ void rs_load(const u64 *keys, u64 n, u64 tile_base, int warp, int lane, u64 (&k)[RS_ITEMS],
                                        bool (&ok)[RS_ITEMS], u64 (&idx)[RS_ITEMS])
{
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        idx[j] = tile_base + (u64)warp * (32 * RS_ITEMS) + (u64)j * 32 + lane;
        ok[j] = idx[j] < n;
        k[j] = ok[j] ? keys[idx[j]] : ~0ull;
    }
}

// per-block digit histogram: hist[d * nblocks + block]
static __global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const u64 *keys, u64 n, int shift, u32 *hist, u32 nblocks)
{
    __shared__ u32 cnt[256];
    const int tid = threadIdx.x;
    cnt[tid] = 0;
    __syncthreads();
    const u64 tile_base = (u64)blockIdx.x * RS_TILE;
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        const u64 i = tile_base + (u64)j * RS_THREADS + tid;
        if (i < n) atomicAdd(&cnt[(keys[i] >> shift) & 0xff], 1u);
    }
    __syncthreads();
    hist[(u64)tid * nblocks + blockIdx.x] = cnt[tid];
}

static __global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const u64 *keys_in, const u32 *vals_in, u64 *keys_out,
                                                                 u32 *vals_out, u64 n, int shift, const u32 *base,
                                                                 u32 nblocks)
{
    __shared__ u32 cnt[RS_WARPS][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const u64 tile_base = (u64)blockIdx.x * RS_TILE;
    u64 k[RS_ITEMS], idx[RS_ITEMS];
    bool ok[RS_ITEMS];
    rs_load(keys_in, n, tile_base, warp, lane, k, ok, idx);
    const unsigned lt = (1u << lane) - 1u;
    // phase 1: per-warp digit counts
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        const u32 d = ok[j] ? (u32)((k[j] >> shift) & 0xff) : 256u + 0u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (ok[j] && (peers & lt) == 0) cnt[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // digit `tid`: exclusive scan over warps, seeded with the global base of (digit, block)
        u32 run = base[(u64)tid * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const u32 t = cnt[w][tid];
            cnt[w][tid] = run;
            run += t;
        }
    }
    __syncthreads();
    // phase 2: stable rank and scatter
#pragma unroll
    for (int j = 0; j < RS_ITEMS; j++) {
        const u32 d = ok[j] ? (u32)((k[j] >> shift) & 0xff) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        u32 pos = 0;
        if (ok[j]) pos = cnt[warp][d] + __popc(peers & lt);
        __syncwarp();
        if (ok[j] && (peers & lt) == 0) cnt[warp][d] += __popc(peers);
        __syncwarp();
        if (ok[j]) {
            keys_out[pos] = k[j];
            if (vals_in) vals_out[pos] = vals_in[idx[j]];
        }
    }
}

// Sort keys[0..n) (and vals, may be NULL) ascending on bits [0, nbits). keys_tmp/vals_tmp are
// scratch of the same size.  The sorted data always ends up back in keys/vals.  n < 2^32.
static int radix_sort_pairs(euler_ctx *ctx, u64 *keys, u32 *vals, u64 n, int nbits, u64 *keys_tmp, u32 *vals_tmp,
                            u32 *hist /* 256*nblocks u32 scratch */)
{
    if (n < 2) return EULER_OK;
    const u32 nblocks = (u32)((n + RS_TILE - 1) / RS_TILE);
    u64 *kin = keys, *kout = keys_tmp;
    u32 *vin = vals, *vout = vals_tmp;
    for (int shift = 0; shift < nbits; shift += 8) {
        rs_hist_kernel<<<nblocks, RS_THREADS, 0, ctx->stream>>>(kin, n, shift, hist, nblocks);
        CUDA_TRY(ctx, cudaGetLastError());
        EULER_TRY(scan_exclusive(ctx, ScanInU32{hist}, (u64)256 * nblocks, hist, (u64 *)nullptr));
        rs_scatter_kernel<<<nblocks, RS_THREADS, 0, ctx->stream>>>(kin, vin, kout, vout, n, shift, hist, nblocks);
        CUDA_TRY(ctx, cudaGetLastError());
        u64 *tk = kin; kin = kout; kout = tk;
        u32 *tv = vin; vin = vout; vout = tv;
    }
    if (kin != keys) {
        CUDA_TRY(ctx, cudaMemcpyAsync(keys, kin, n * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
        if (vals) CUDA_TRY(ctx, cudaMemcpyAsync(vals, vin, n * sizeof(u32), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return EULER_OK;
}
