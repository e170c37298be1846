// ingest.cu -- FASTA / FASTQ ingestion on device: the step in front of the encoder,
// host Python in the reference (eulercuda.read_fasta :439-447, read_fastq :44-56).  The raw file
// bytes go to the GPU once; line splitting, header / '+' / quality skipping and the read-offset
// array are three scans and two scatters, and the reads never exist on the host.
//
// Semantics of the reference readers: FASTA -- every line that does not start with '>' is one read
// (so a blank line is an empty read and multi-line records are NOT joined); FASTQ -- lines 1, 5, 9, ...
// (index % 4 == 1).  '\n' and a trailing '\r' are not part of a read.
#include "kernels.h"
#include "scan.cuh"
#include "tmp.cuh"

#define IB 256

struct NewlineIn {
    const unsigned char *b;
    __device__ __forceinline__ u32 operator()(u64 i) const { return b[i] == '\n' ? 1u : 0u; }
};

__global__ void __launch_bounds__(IB) line_starts_kernel(const unsigned char *__restrict__ b, u64 n, const u32 *__restrict__ lineid,
                                                          u32 *__restrict__ linestart)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0 || b[i - 1] == '\n') linestart[lineid[i]] = (u32)i;
}

__device__ __forceinline__ bool line_is_read(const unsigned char *b, const u32 *linestart, u32 lid, int fastq)
{
    if (fastq) return (lid & 3u) == 1u;
    return b[linestart[lid]] != '>';
}

struct KeepIn {
    const unsigned char *b;
    const u32 *lineid, *linestart;
    int fastq;
    __device__ __forceinline__ u32 operator()(u64 i) const
    {
        const unsigned char c = b[i];
        if (c == '\n' || c == '\r') return 0u;
        return line_is_read(b, linestart, lineid[i], fastq) ? 1u : 0u;
    }
};
struct ReadLineIn {
    const unsigned char *b;
    const u32 *linestart;
    int fastq;
    __device__ __forceinline__ u32 operator()(u64 lid) const { return line_is_read(b, linestart, (u32)lid, fastq) ? 1u : 0u; }
};

__global__ void __launch_bounds__(IB) ingest_scatter_kernel(const unsigned char *__restrict__ b, u64 n, const u32 *__restrict__ lineid,
                                                             const u32 *__restrict__ linestart, int fastq,
                                                             const u32 *__restrict__ outpos, unsigned char *__restrict__ out)
{
    const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned char c = b[i];
    if (c == '\n' || c == '\r') return;
    if (line_is_read(b, linestart, lineid[i], fastq)) out[outpos[i]] = c;
}

__global__ void __launch_bounds__(IB) ingest_offsets_kernel(const unsigned char *__restrict__ b, const u32 *__restrict__ linestart,
                                                             u64 nlines, int fastq, const u32 *__restrict__ readidx,
                                                             const u32 *__restrict__ outpos, u64 *__restrict__ off)
{
    const u64 lid = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (lid >= nlines) return;
    if (line_is_read(b, linestart, (u32)lid, fastq)) off[readidx[lid]] = outpos[linestart[lid]];
}

// d_file: n raw bytes.  Outputs (ctx-owned by the caller): d_reads (>= n bytes), d_off (>= lines + 1).
// Returns counts through nreads / nbases (synchronises).
int ingest_parse(euler_ctx *ctx, const unsigned char *d_file, u64 n, int fastq, unsigned char *d_reads, u64 *d_off, u64 off_cap,
                 u64 *nreads, u64 *nbases)
{
    *nreads = 0; *nbases = 0;
    if (!n) {
        CUDA_TRY(ctx, cudaMemsetAsync(d_off, 0, sizeof(u64), ctx->stream));
        return EULER_OK;
    }
    if (n >= 0xffffffffull) return euler_fail(ctx, EULER_ERR_RANGE, "ingest: file of %llu bytes needs chunking (u32 positions)", n);
    DevTmp<u32> lineid(ctx, n), outpos(ctx, n + 1);
    DevTmp<u64> totals(ctx, 4);
    TMP_CHECK(ctx, lineid); TMP_CHECK(ctx, outpos); TMP_CHECK(ctx, totals);
    EULER_TRY(scan_exclusive(ctx, NewlineIn{d_file}, n, lineid.get(), totals.get()));
    u64 nl = 0;
    EULER_TRY(read_u64(ctx, totals, &nl));
    unsigned char last = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_pinned, d_file + n - 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    last = *(unsigned char *)ctx->h_pinned;
    const u64 nlines = nl + (last == '\n' ? 0 : 1);
    if (nlines + 1 > off_cap) return euler_fail(ctx, EULER_ERR_ARG, "ingest: offset buffer too small");
    DevTmp<u32> linestart(ctx, nlines + 1), readidx(ctx, nlines + 1);
    TMP_CHECK(ctx, linestart); TMP_CHECK(ctx, readidx);
    line_starts_kernel<<<grid_for(n, IB), IB, 0, ctx->stream>>>(d_file, n, lineid, linestart);
    EULER_TRY(scan_exclusive(ctx, KeepIn{d_file, lineid, linestart, fastq}, n, outpos.get(), totals.get() + 1));
    EULER_TRY(scan_exclusive(ctx, ReadLineIn{d_file, linestart, fastq}, nlines, readidx.get(), totals.get() + 2));
    u64 h[3];
    EULER_TRY(read_u64s(ctx, totals, h, 3));
    const u64 nb = h[1], nr = h[2];
    ingest_scatter_kernel<<<grid_for(n, IB), IB, 0, ctx->stream>>>(d_file, n, lineid, linestart, fastq, outpos, d_reads);
    ingest_offsets_kernel<<<grid_for(nlines, IB), IB, 0, ctx->stream>>>(d_file, linestart, nlines, fastq, readidx, outpos, d_off);
    CUDA_TRY(ctx, cudaGetLastError());
    // off[R] = total kept bytes
    ctx->h_pinned[16] = nb;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_off + nr, ctx->h_pinned + 16, sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *nreads = nr;
    *nbases = nb;
    return EULER_OK;
}
