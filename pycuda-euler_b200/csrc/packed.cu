// packed.cu -- L2-sized count table for the fused encode+count kernel.
//
// The SoA table (8 B key + 4 B count per slot, load 0.55) is 100 MB for the 4.6 Mbp workload; the
// profiles show ~55 % of its probes and REDs missing L2 and the kernel sitting at the DRAM
// random-sector rate.  Here a slot is ONE 64-bit word:
//
//     [ count : cb ][ remainder : 64-b ][ displacement : 5 ][ valid : 1 ]        cb = b - 6
//
// With a bijective hash H, key' = H(key); the top b bits of key' choose the home bucket (4 slots =
// one 32 B sector, nb = 2^b buckets) and only the remaining 64-b bits are stored (quotienting): the
// bucket index plus the displacement (how many buckets past home, at most 31) restore them.  A hit is
// one 256-bit load and one 64-bit atomicAdd on the SAME sector, the table is 8 B per slot, and the
// count lives in the top bits so that it wraps harmlessly: the thread whose add wraps it records
// 2^cb occurrences in a small side table.  After counting, unpack_kernel expands the words into the
// SoA arrays the graph stage reads (same slot numbering), so nothing downstream changes.
// Out-of-range displacement or a full side table set a flag and the pipeline re-runs with the SoA
// kernel (no silent loss).
#include "encode.cuh"
#include "kernels.h"

#define PK_BLOCK 256
#define PK_C 0x9E3779B97F4A7C15ull
#define PK_CINV 0xF1DE83E19937733Dull   // PK_C * PK_CINV == 1 (mod 2^64)
#define PK_DISP_BITS 5
#define PK_MAXDISP (1 << PK_DISP_BITS)

__device__ __forceinline__ u64 pk_hash(u64 key) { return (key ^ (key >> 29)) * PK_C; }
__device__ __forceinline__ u64 pk_unhash(u64 y)
{
    const u64 t = y * PK_CINV;
    return t ^ (t >> 29) ^ (t >> 58);
}

struct PackedGeom {
    u32 b;        // log2(number of buckets)
    u32 rb;       // remainder bits = 64 - b
    u32 lowbits;  // rb + PK_DISP_BITS + 1
    u64 lowmask, rmask, inc, cmask;
    u32 nb_mask;
};
__host__ __device__ __forceinline__ PackedGeom packed_geom(u32 b)
{
    PackedGeom g;
    g.b = b;
    g.rb = 64 - b;
    g.lowbits = g.rb + PK_DISP_BITS + 1;
    g.lowmask = (1ull << g.lowbits) - 1ull;
    g.rmask = (1ull << g.rb) - 1ull;
    g.inc = 1ull << g.lowbits;
    g.cmask = (1ull << (64 - g.lowbits)) - 1ull;
    g.nb_mask = (1u << b) - 1u;
    return g;
}

// the add that wrapped the count field owes 2^cb occurrences to the side table
__device__ __forceinline__ void pk_side_add(u64 key, u64 *side_keys, u32 *side_cnt, u64 side_cap, u64 *stats)
{
    const u64 slot = table_insert(side_keys, side_cap, key, side_cap / EULER_BUCKET);
    if (slot == EULER_NO_SLOT) atomicOr((unsigned long long *)(stats + 2), 4ull);
    else atomicAdd(side_cnt + slot, 1u);
}

__global__ void __launch_bounds__(PK_BLOCK, 4) count_packed_kernel(const uint4 *__restrict__ buf16, u64 n_bases,
                                                                    const u32 *__restrict__ start_bits, u32 l,
                                                                    u64 *__restrict__ tab, PackedGeom g, u64 *__restrict__ side_keys,
                                                                    u32 *__restrict__ side_cnt, u64 side_cap, u64 ntiles,
                                                                    u64 *__restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    const u32 k = l - 1;
    const u32 top = 2 * (l - 1);
    const u64 kmask = key_mask_d(l);
    u32 nl_tot = 0, nk_tot = 0;
    bool overflow = false;

    for (u64 tile = warp; tile < ntiles; tile += nwarps) {
        const long long chunk = (long long)(tile * ENC_ADV) - ENC_HALO + lane;
        const Chunk c = load_chunk(buf16, chunk, n_bases, start_bits);
        const u32 p1 = __shfl_up_sync(0xffffffffu, c.codes, 1), p2 = __shfl_up_sync(0xffffffffu, c.codes, 2);
        const u32 v1 = __shfl_up_sync(0xffffffffu, c.vmask, 1), v2 = __shfl_up_sync(0xffffffffu, c.vmask, 2);
        const u32 s1 = __shfl_up_sync(0xffffffffu, c.smask, 1), s2 = __shfl_up_sync(0xffffffffu, c.smask, 2);
        u64 f = ((u64)p2 << 32) | p1;
        u64 rc = revcomp64(f & kmask, l);
        const u32 pv = (v2 << 16) | v1, ps = (s2 << 16) | s1;
        u32 vrun = (pv == 0xffffffffu) ? 32u : (u32)__ffs(~pv) - 1u;
        u32 srun = ps ? (u32)__ffs(ps) - 1u : 32u;
        u32 codes = c.codes;
        u32 vm = (lane < ENC_HALO) ? 0u : (c.vmask << 16);
        u32 sm = c.smask << 16;
        if (lane < ENC_HALO) vrun = 0;

#pragma unroll 1
        for (int b4 = 0; b4 < 4; b4++) {
            u64 key[4];
            u32 pend = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const u32 cc = codes >> 30;
                codes <<= 2;
                const bool valid = (vm >> 31) != 0, start = (sm >> 31) != 0;
                vm <<= 1;
                sm <<= 1;
                f = (f << 2) | cc;
                rc = (rc >> 2) | ((u64)(3u - cc) << top);
                vrun = valid ? vrun + 1u : 0u;
                srun = start ? 0u : srun + 1u;
                nk_tot += (vrun >= k && srun + 1u >= k) ? 1u : 0u;
                const bool ok = vrun >= l && srun + 1u >= l;
                nl_tot += ok ? 1u : 0u;
                const u64 fm = f & kmask;
                key[i] = fm < rc ? fm : rc;
                pend |= (ok ? 1u : 0u) << i;
            }
            u32 home[4];
            u64 ident[4];   // (remainder << 6) | valid; the displacement is OR-ed in per round
            K4 q[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const u64 h = pk_hash(key[i]);
                home[i] = (u32)(h >> g.rb);
                ident[i] = ((h & g.rmask) << (PK_DISP_BITS + 1)) | 1ull;
            }
            // warp-uniform rounds: round d probes bucket home + d
            for (u32 d = 0; __any_sync(0xffffffffu, pend != 0); d++) {
                if (d >= PK_MAXDISP) { overflow = true; break; }
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (pend & (1u << i)) q[i] = ld_bucket_cg(tab + (u64)((home[i] + d) & g.nb_mask) * EULER_BUCKET);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    if (!(pend & (1u << i))) continue;
                    u64 *bk = tab + (u64)((home[i] + d) & g.nb_mask) * EULER_BUCKET;
                    const u64 want = ident[i] | ((u64)d << 1);
                    int hit = -1;
                    bool fresh = false;
#pragma unroll
                    for (int j = 0; j < EULER_BUCKET; j++) {
                        if (hit >= 0) continue;
                        const u64 v = q[i].k[j];
                        if ((v & g.lowmask) == want) hit = j;
                        else if (v == 0) {
                            const u64 old = atomicCAS(bk + j, 0ull, want | g.inc);
                            if (old == 0) { hit = j; fresh = true; }
                            else if ((old & g.lowmask) == want) hit = j;
                        }
                    }
                    if (hit >= 0) {
                        if (!fresh) {
                            const u64 old = atomicAdd(bk + hit, g.inc);
                            if ((old >> g.lowbits) == g.cmask) pk_side_add(key[i], side_keys, side_cnt, side_cap, stats);
                        }
                        pend &= ~(1u << i);
                    }
                }
            }
        }
    }
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

// packed words -> SoA (keys, counts), same slot numbering
__global__ void __launch_bounds__(PK_BLOCK) unpack_kernel(const u64 *__restrict__ tab, PackedGeom g, u64 cap,
                                                           const u64 *__restrict__ side_keys, const u32 *__restrict__ side_cnt,
                                                           u64 side_cap, const u64 *__restrict__ side_used, u64 *__restrict__ keys,
                                                           u32 *__restrict__ cnt)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    const u64 v = tab[i];
    if (v == 0) {
        keys[i] = EULER_EMPTY_KEY;
        cnt[i] = 0;
        return;
    }
    const u32 d = (u32)(v >> 1) & (PK_MAXDISP - 1);
    const u64 rem = (v >> (PK_DISP_BITS + 1)) & g.rmask;
    const u32 home = ((u32)(i / EULER_BUCKET) - d) & g.nb_mask;
    const u64 key = pk_unhash(((u64)home << g.rb) | rem);
    u64 n = v >> g.lowbits;
    if (*side_used) {
        const u64 s = table_find(side_keys, side_cap, key);
        if (s != EULER_NO_SLOT) n += (u64)side_cnt[s] << (64 - g.lowbits);
    }
    keys[i] = key;
    cnt[i] = (u32)n;
}

__global__ void side_used_kernel(const u64 *__restrict__ side_keys, u64 side_cap, u64 *flag)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < side_cap && side_keys[i] != EULER_EMPTY_KEY) *flag = 1ull;
}

// tab: u64[cap] zeroed by the caller, cap = 4 << b.  side tables: EMPTY-filled keys / zero counts.
int enc_count_packed(euler_ctx *ctx, const void *d_buf, u64 n_bases, const u32 *d_bits, u32 l, u64 *tab, u32 b,
                     u64 *side_keys, u32 *side_cnt, u64 side_cap, u64 *d_stats)
{
    if (!n_bases) return EULER_OK;
    const u64 nchunks = (n_bases + 15) / 16;
    const u64 ntiles = (nchunks + ENC_ADV - 1) / ENC_ADV;
    u64 grid = (u64)ctx->num_sms * 4;
    const u64 need = (ntiles + PK_BLOCK / 32 - 1) / (PK_BLOCK / 32);
    if (grid > need) grid = need;
    count_packed_kernel<<<(unsigned)(grid ? grid : 1), PK_BLOCK, 0, ctx->stream>>>((const uint4 *)d_buf, n_bases, d_bits, l, tab,
                                                                                  packed_geom(b), side_keys, side_cnt, side_cap,
                                                                                  ntiles, d_stats);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}

int enc_unpack(euler_ctx *ctx, const u64 *tab, u32 b, const u64 *side_keys, const u32 *side_cnt, u64 side_cap, u64 *d_side_used,
               u64 *keys, u32 *cnt)
{
    const u64 cap = (u64)EULER_BUCKET << b;
    CUDA_TRY(ctx, cudaMemsetAsync(d_side_used, 0, sizeof(u64), ctx->stream));
    side_used_kernel<<<grid_for(side_cap, 256), 256, 0, ctx->stream>>>(side_keys, side_cap, d_side_used);
    unpack_kernel<<<grid_for(cap, PK_BLOCK), PK_BLOCK, 0, ctx->stream>>>(tab, packed_geom(b), cap, side_keys, side_cnt, side_cap,
                                                                        d_side_used, keys, cnt);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
