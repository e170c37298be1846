// encode.cu -- encoder kernels and the fused encode+count kernel (the hot kernel of the path).
#include "encode.cuh"
#include "kernels.h"

// ---- read-start bitmap: bit t set iff a read starts at byte t ---------------------------------
__global__ void mark_starts_kernel(const u64 *__restrict__ off, u64 nreads, u32 *__restrict__ bits)
{
    const u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nreads) return;
    const u64 t = off[j];
    atomicOr(bits + (t >> 5), 1u << (t & 31));
}

int enc_mark_starts(euler_ctx *ctx, const u64 *d_off, u64 nreads, u64 n_bases, u32 *d_bits)
{
    const u64 words = n_bases / 32 + 2;
    CUDA_TRY(ctx, cudaMemsetAsync(d_bits, 0, words * sizeof(u32), ctx->stream));
    if (nreads) {
        mark_starts_kernel<<<grid_for(nreads, 256), 256, 0, ctx->stream>>>(d_off, nreads, d_bits);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return EULER_OK;
}

// ---- module-level encoder: one value per window start byte ------------------------------------
__global__ void __launch_bounds__(256) encode_positions_kernel(const uint4 *__restrict__ buf16, u64 n_bases,
                                                                const u32 *__restrict__ start_bits, u32 l,
                                                                u64 *__restrict__ out_fwd, u64 *__restrict__ out_rc,
                                                                unsigned char *__restrict__ out_valid, u64 ntiles)
{
    const int lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 tile = warp; tile < ntiles; tile += nwarps) {
        const long long chunk = (long long)(tile * ENC_ADV) - ENC_HALO + lane;
        const Chunk c = load_chunk(buf16, chunk, n_bases, start_bits);
        for_each_window(c, l, lane, [&](int i, u64 key) {
            const u64 start = (u64)chunk * 16 + i + 1 - l;
            out_fwd[start] = key;
            if (out_rc) out_rc[start] = revcomp64(key, l);
            if (out_valid) out_valid[start] = 1;
        });
    }
}

int enc_positions(euler_ctx *ctx, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, u64 *d_fwd, u64 *d_rc,
                  unsigned char *d_valid)
{
    CUDA_TRY(ctx, cudaMemsetAsync(d_fwd, 0, n_bases * sizeof(u64), ctx->stream));
    if (d_rc) CUDA_TRY(ctx, cudaMemsetAsync(d_rc, 0, n_bases * sizeof(u64), ctx->stream));
    if (d_valid) CUDA_TRY(ctx, cudaMemsetAsync(d_valid, 0, n_bases, ctx->stream));
    if (!n_bases) return EULER_OK;
    const u64 nchunks = (n_bases + 15) / 16;
    const u64 ntiles = (nchunks + ENC_ADV - 1) / ENC_ADV;
    const unsigned grid = (unsigned)min((u64)ctx->num_sms * 8, (ntiles + 7) / 8);
    encode_positions_kernel<<<grid ? grid : 1, 256, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, d_fwd,
                                                                      d_rc, d_valid, ntiles);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

__global__ void compute_kmers_kernel(const u64 *__restrict__ lmers, u64 n, u64 mask, u64 *__restrict__ pk,
                                     u64 *__restrict__ sk)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 x = lmers[i];
    pk[i] = (x & (mask << 2)) >> 2;  // LMER_PREFIX pyencode.py:109
    sk[i] = x & mask;                // LMER_SUFFIX pyencode.py:110
}

int enc_compute_kmers(euler_ctx *ctx, const u64 *d_lmers, u64 n, u64 mask, u64 *d_pk, u64 *d_sk)
{
    if (!n) return EULER_OK;
    compute_kmers_kernel<<<grid_for(n, 256), 256, 0, ctx->stream>>>(d_lmers, n, mask, d_pk, d_sk);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

// ---- fused encode + canonicalise + count (hot kernel) -----------------------------------------
// Every forward window contributes +1 to its canonical key min(x, rc(x)); the both-strand table of
// the reference (eulercuda.py:141-161) is count[x] = count[rc x] = c, or 2c for palindromes.
// Table: SoA keys u64[cap] / counts u32[cap], linear probing, EMPTY = all-ones (never canonical).
#define CNT_BLOCK 256

__global__ void __launch_bounds__(CNT_BLOCK) count_canonical_kernel(const uint4 *__restrict__ buf16, u64 n_bases,
                                                                     const u32 *__restrict__ start_bits, u32 l,
                                                                     u64 *__restrict__ tab_keys, u32 *__restrict__ tab_cnt,
                                                                     u64 cap, u64 ntiles, u64 *__restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u64 max_probe = cap < 8192 ? cap : 8192;
    u32 nl_tot = 0, nk_tot = 0;
    bool overflow = false;

    for (u64 tile = warp; tile < ntiles; tile += nwarps) {
        const long long chunk = (long long)(tile * ENC_ADV) - ENC_HALO + lane;
        const Chunk c = load_chunk(buf16, chunk, n_bases, start_bits);
        u64 keys[16];
        u32 okmask = 0;
        const u32 cnts = for_each_window(c, l, lane, [&](int i, u64 key) {
            const u64 rc = revcomp64(key, l);
            keys[i] = key < rc ? key : rc;
            okmask |= 1u << i;
        });
        nl_tot += cnts & 0xffffu;
        nk_tot += cnts >> 16;

        // Two batches of 8 keys.  Probing proceeds in warp-uniform rounds: every round first issues
        // the loads of all still-pending keys of the batch (8 independent requests per lane in
        // flight), then resolves them, so lanes that need another probe take it together instead
        // of serialising one lane at a time.
#pragma unroll
        for (int b = 0; b < 16; b += 8) {
            u64 slot[8], cur[8];
            u32 pend = (okmask >> b) & 0xffu;
#pragma unroll
            for (int i = 0; i < 8; i++) slot[i] = hash_slot(keys[b + i], cap);
            u64 probes = 0;
            while (__any_sync(0xffffffffu, pend != 0)) {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (pend & (1u << i)) cur[i] = ld_cg_u64(tab_keys + slot[i]);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (pend & (1u << i)) {
                        const u64 key = keys[b + i];
                        bool done = cur[i] == key;
                        if (!done && cur[i] == EULER_EMPTY_KEY) {
                            const u64 old = atomicCAS(tab_keys + slot[i], EULER_EMPTY_KEY, key);
                            done = (old == EULER_EMPTY_KEY) || (old == key);
                        }
                        if (done) {
                            atomicAdd(tab_cnt + slot[i], 1u);
                            pend &= ~(1u << i);
                        } else if (++slot[i] == cap) {
                            slot[i] = 0;
                        }
                    }
                }
                if (++probes >= max_probe && pend) { overflow = true; pend = 0; }
            }
        }
    }
    // per-warp reduction of the window counters
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        nl_tot += __shfl_xor_sync(0xffffffffu, nl_tot, d);
        nk_tot += __shfl_xor_sync(0xffffffffu, nk_tot, d);
    }
    if (lane == 0) {
        if (nl_tot) atomicAdd(stats + 0, (u64)nl_tot);
        if (nk_tot) atomicAdd(stats + 1, (u64)nk_tot);
    }
    if (overflow) atomicOr((unsigned long long *)(stats + 2), 1ull);
}

int enc_count_canonical(euler_ctx *ctx, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, u64 *tab_keys,
                        u32 *tab_cnt, u64 cap, u64 *d_stats)
{
    if (!n_bases) return EULER_OK;
    const u64 nchunks = (n_bases + 15) / 16;
    const u64 ntiles = (nchunks + ENC_ADV - 1) / ENC_ADV;
    const u64 warps_per_block = CNT_BLOCK / 32;
    u64 grid = (u64)ctx->num_sms * 4;
    const u64 need = (ntiles + warps_per_block - 1) / warps_per_block;
    if (grid > need) grid = need;
    count_canonical_kernel<<<(unsigned)(grid ? grid : 1), CNT_BLOCK, 0, ctx->stream>>>(
        (const uint4 *)d_buf, n_bases, d_bits, l, tab_keys, tab_cnt, cap, ntiles, d_stats);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
