// common.cuh -- shared device/host helpers for libeuler_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/euler_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;

#define EULER_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define EULER_NO_ID 0xFFFFFFFFu
#define EULER_SMS 148

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct Pipeline;

struct euler_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    // pinned host staging for small device->host reads
    u64 *h_pinned = nullptr;
    Pipeline *pipe = nullptr;
    DevBuf scan_state;  // tile descriptors of the single-pass scan
    DevBuf cg_buf;      // circuit edges of the last tour_circuit_edges call
    DevBuf text_buf;    // contig text of the last emission
    cudaEvent_t ev[8] = {};
    int num_sms = EULER_SMS;
    size_t l2_bytes = 0;
    size_t persist_max = 0;
};

int euler_fail(euler_ctx *ctx, int code, const char *fmt, ...);

#define CUDA_TRY(ctx, expr)                                                                        \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return euler_fail((ctx), EULER_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr,    \
                              cudaGetErrorString(e_));                                             \
    } while (0)

#define EULER_TRY(expr)                                                                            \
    do {                                                                                           \
        int r_ = (expr);                                                                           \
        if (r_ != EULER_OK) return r_;                                                             \
    } while (0)

// grow-only device buffer
int dev_reserve(euler_ctx *ctx, DevBuf &b, size_t bytes);
void dev_free(DevBuf &b);

template <typename T>
struct DevArr {
    DevBuf b;
    T *ptr() const { return (T *)b.p; }
    int reserve(euler_ctx *ctx, size_t n) { return dev_reserve(ctx, b, (n ? n : 1) * sizeof(T)); }
    void free() { dev_free(b); }
};

static inline unsigned grid_for(u64 n, unsigned block)
{
    u64 g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > 0x7fffffffull) g = 0x7fffffffull;
    return (unsigned)g;
}

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ u64 mix64(u64 x)
{
    // murmur3 fmix64
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// slot in [0, cap) from a 64-bit hash (Lemire fastrange)
__device__ __forceinline__ u64 hash_slot(u64 key, u64 cap) { return __umul64hi(mix64(key), cap); }

__device__ __forceinline__ u64 key_mask_d(u32 len) { return len >= 32 ? ~0ull : ((1ull << (2 * len)) - 1ull); }

// reverse complement of a len-mer packed MSB-first in the low 2*len bits
__device__ __forceinline__ u64 revcomp64(u64 x, u32 len)
{
    u64 y = __brevll(~x);                                                  // reverse all bits of the complement
    y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);  // restore bit order inside each base
    return y >> (64 - 2 * len);
}

__device__ __forceinline__ u64 ld_cg_u64(const u64 *p)
{
    u64 v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ u32 ld_cg_u32(const u32 *p)
{
    u32 v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// streaming 128-bit load (read-once input): non-coherent path, no L1 allocation
__device__ __forceinline__ uint4 ld_stream_v4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ---- open-addressing table primitives (linear probing, out-of-band-free sentinel) --------------
// Canonical l-mers never equal all-ones (canon(T^32)=A^32=0) and k-mers use <= 62 bits, so
// 0xFFFF... is a safe EMPTY for every key this library stores.

// returns slot of `key` after inserting it if absent; EULER_NO_SLOT on overflow
#define EULER_NO_SLOT 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ u64 table_insert(u64 *keys, u64 cap, u64 key, u64 max_probe)
{
    u64 slot = hash_slot(key, cap);
    for (u64 probe = 0; probe < max_probe; probe++) {
        u64 k = ld_cg_u64(keys + slot);
        if (k == key) return slot;
        if (k == EULER_EMPTY_KEY) {
            u64 old = atomicCAS(keys + slot, EULER_EMPTY_KEY, key);
            if (old == EULER_EMPTY_KEY || old == key) return slot;
        }
        slot++;
        if (slot == cap) slot = 0;
    }
    return EULER_NO_SLOT;
}

__device__ __forceinline__ u64 table_find(const u64 *keys, u64 cap, u64 key)
{
    u64 slot = hash_slot(key, cap);
    for (u64 probe = 0; probe < cap; probe++) {
        u64 k = __ldg(keys + slot);
        if (k == key) return slot;
        if (k == EULER_EMPTY_KEY) return EULER_NO_SLOT;
        slot++;
        if (slot == cap) slot = 0;
    }
    return EULER_NO_SLOT;
}

#endif  // __CUDACC__
