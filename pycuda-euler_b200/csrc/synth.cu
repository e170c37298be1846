// synth.cu -- deterministic synthetic reads on device: counter-based splitmix64
// genome / read starts / strands / substitution errors, reproducible on the CPU for any subset.
#include "kernels.h"

__device__ __forceinline__ u64 splitmix64_at(u64 seed, u64 ctr)
{
    u64 z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ u32 genome_base(u64 i)
{
    return (u32)((splitmix64_at(0x5EED0001ull, i >> 5) >> (2 * (i & 31))) & 3);
}

// one warp per read; lanes stride over the bases
__global__ void __launch_bounds__(256) synth_reads_kernel(u64 G, u32 L, u64 thr, u64 first, u64 nreads, char *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const u64 warp = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 jj = warp; jj < nreads; jj += nwarps) {
        const u64 j = first + jj;
        const u64 h = splitmix64_at(0x5EED0002ull, j);
        const u64 st = (h >> 1) % (G - L + 1);
        const bool rcs = h & 1;
        char *o = out + jj * L;
        for (u32 p = lane; p < L; p += 32) {
            u32 b = rcs ? 3u - genome_base(st + L - 1 - p) : genome_base(st + p);
            if (thr) {
                const u64 e = splitmix64_at(0x5EED0003ull, j * (u64)L + p);
                if ((e & 0xffffffffull) < thr) b = (b + 1 + (u32)((e >> 32) % 3)) & 3;
            }
            o[p] = "ACGT"[b];
        }
    }
}

int synth_reads(euler_ctx *ctx, u64 G, u32 L, u32 err_ppm, u64 first, u64 nreads, void *d_out)
{
    if (!nreads) return EULER_OK;
    if (L == 0 || G < L) return euler_fail(ctx, EULER_ERR_ARG, "synth: genome shorter than read");
    const u64 thr = ((u64)err_ppm << 32) / 1000000ull;
    u64 grid = (nreads + 7) / 8;
    if (grid > (u64)ctx->num_sms * 16) grid = (u64)ctx->num_sms * 16;
    synth_reads_kernel<<<(unsigned)grid, 256, 0, ctx->stream>>>(G, L, thr, first, nreads, (char *)d_out);
    CUDA_TRY(ctx, cudaGetLastError());
    return EULER_OK;
}
