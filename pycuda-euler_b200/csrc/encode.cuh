// encode.cuh -- 2-bit read encoder: 128-bit loads of the ASCII byte stream, SIMD-in-register
// packing, warp-shuffle sliding windows.  Replaces encodeLmerDevice (pyencode.py:45-76),
// computeKmerDevice (:112-135) and the intended encodeLmerComplementDevice (:173-209).
//
// Tile geometry: a warp covers 32 chunks of 16 bytes.  Lanes [0, HALO) only supply left context
// (the l-1 bases before a window's last base); lanes [HALO, 32) each own the 16 windows that END
// in their chunk.  HALO = 2 for l <= 32.  Consecutive tiles advance by 32-HALO chunks, so every
// load is a 16-byte aligned vector load and no state is carried between tiles.
#pragma once
#ifdef EULER_SIMT_EMU   // CPU tests of the kernels (tests/host/simt_emu.h)
#include "simt_emu.h"
#else
#include "common.cuh"
#endif

#define ENC_HALO 2
#define ENC_ADV (32 - ENC_HALO)

struct Chunk {
    u32 codes;  // 16 bases, first base in bits 31:30
    u32 vmask;  // bit (15-i) = base i is ACGT and inside the buffer
    u32 smask;  // bit (15-i) = a read starts at base i
};

// four ASCII bytes -> 8 bits of codes (first byte in bits 7:6) and 4 valid bits (first byte in bit 3)
__device__ __forceinline__ void pack4(u32 w, u32 &codes8, u32 &valid4)
{
    const u32 up = w & 0xDFDFDFDFu;  // fold case
    const u32 ok = __vcmpeq4(up, 0x41414141u) | __vcmpeq4(up, 0x43434343u) | __vcmpeq4(up, 0x47474747u) |
                   __vcmpeq4(up, 0x54545454u);
    // A=0x41 C=0x43 G=0x47 T=0x54: code = ((c>>1) ^ (c>>2)) & 3  -> A0 C1 G2 T3 (pyencode.py:42 codeF)
    const u32 t = ((w >> 1) ^ (w >> 2)) & 0x03030303u;
    codes8 = (t * 0x40100401u) >> 24;
    valid4 = ((ok & 0x01010101u) * 0x08040201u) >> 24;
}

__device__ __forceinline__ Chunk load_chunk(const uint4 *buf16, long long chunk, u64 n_bases, const u32 *start_bits)
{
    Chunk c = {0u, 0u, 0u};
    if (chunk < 0) return c;
    const u64 pos = (u64)chunk * 16;
    if (pos >= n_bases) return c;
    uint4 w;
    if (pos + 16 <= n_bases) {
        w = ld_stream_v4(buf16 + chunk);
    } else {  // ragged tail: byte loads, pad with 0 (invalid)
        const unsigned char *b = (const unsigned char *)buf16 + pos;
        u32 t[4] = {0, 0, 0, 0};
        for (u32 i = 0; pos + i < n_bases; i++) t[i >> 2] |= (u32)b[i] << (8 * (i & 3));
        w = make_uint4(t[0], t[1], t[2], t[3]);
    }
    u32 c0, c1, c2, c3, v0, v1, v2, v3;
    pack4(w.x, c0, v0);
    pack4(w.y, c1, v1);
    pack4(w.z, c2, v2);
    pack4(w.w, c3, v3);
    c.codes = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    c.vmask = (v0 << 12) | (v1 << 8) | (v2 << 4) | v3;
    const u32 word = __ldg(start_bits + (chunk >> 1));
    const u32 half = (chunk & 1) ? (word >> 16) : (word & 0xffffu);
    c.smask = __brev(half) >> 16;
    return c;
}

// Windows ending in this lane's chunk.  Calls f(i, fwd_key) for every valid l-window whose last
// base is base i of the chunk; returns (#valid l-windows) | (#valid (l-1)-windows << 16).
template <typename F>
__device__ __forceinline__ u32 for_each_window(const Chunk &c, u32 l, int lane, F f)
{
    const u32 p1 = __shfl_up_sync(0xffffffffu, c.codes, 1), p2 = __shfl_up_sync(0xffffffffu, c.codes, 2);
    const u32 v1 = __shfl_up_sync(0xffffffffu, c.vmask, 1), v2 = __shfl_up_sync(0xffffffffu, c.vmask, 2);
    const u32 s1 = __shfl_up_sync(0xffffffffu, c.smask, 1), s2 = __shfl_up_sync(0xffffffffu, c.smask, 2);
    if (lane < ENC_HALO) return 0;
    const u64 A = ((u64)p2 << 32) | p1;
    const u64 VM = ((u64)v2 << 32) | ((u64)v1 << 16) | c.vmask;
    const u64 SM = ((u64)s2 << 32) | ((u64)s1 << 16) | c.smask;
    const u64 lm = (l >= 32) ? 0xffffffffull : ((1ull << l) - 1);  // l valid bases
    const u64 lm1 = lm >> 1;                                        // no read start in the last l-1
    const u64 km = lm >> 1, km1 = lm >> 2;                          // same for k = l-1
    const u64 kmask = key_mask_d(l);
    u32 nl = 0, nk = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int sh = 15 - i;
        const u64 vw = VM >> sh, sw = SM >> sh;
        if ((vw & km) == km && (sw & km1) == 0) nk++;
        if ((vw & lm) == lm && (sw & lm1) == 0) {
            nl++;
            const u64 key = ((A << (32 - 2 * sh)) | ((u64)c.codes >> (2 * sh))) & kmask;
            f(i, key);
        }
    }
    return nl | (nk << 16);
}
