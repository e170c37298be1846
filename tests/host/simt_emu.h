// simt_emu.h -- a small SIMT emulator for CPU tests of the CUDA kernels (test infrastructure only).
//
// A kernel source compiled with -DEULER_SIMT_EMU includes this header instead of the CUDA ones.  Every CUDA thread
// becomes an OS thread; blocks run one after the other (a block never depends on a LATER block in the kernels tested
// here: work and output tickets are taken with atomics).  What is emulated:
//   * threadIdx / blockIdx / blockDim / gridDim, static and dynamic shared memory (one block at a time, so `static`
//     storage is the block's shared memory), __syncthreads (threads that have left the kernel count as arrived);
//   * warp collectives (__shfl*_sync, __ballot_sync, __any_sync, __syncwarp) as rendezvous of the warp's live lanes
//     through a per-warp mailbox.  Every rendezvous carries the identity of the operation (kind + source line of
//     the call): lanes that meet at DIFFERENT collectives abort the run with a message -- on hardware that is the
//     undefined behaviour / hang of a warp whose lanes disagree about "warp-uniform" control flow;
//   * atomics on shared and global memory (GCC __atomic builtins), volatile loads / stores, fences, bit intrinsics.
// Lanes are NOT in lock step between collectives (like independent thread scheduling at its most adversarial), so
// code that relies on implicit warp synchrony fails here as well.  A deadlock shows as a hang: callers run the
// test binary under a timeout.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/euler_b200.h"   // euler_vertex: the artefact struct the kernels write

typedef unsigned long long u64;
typedef unsigned int u32;

// ---- the constants the kernels take from common.cuh / kernels.h (tests/test_cpu_simt.py checks they agree) -------------
#define EULER_EMPTY_KEY 0xFFFFFFFFFFFFFFFFull
#define EULER_NO_ID 0xFFFFFFFFu
#define BKT_FLAG_REGION 0x10u
#define BKT_FLAG_TABLE 0x20u
#define BKT_FLAG_OUTPUT 0x40u
#define BKT_FLAG_BOUNDARY 0x80u
#define BKT_FLAG_INTERNAL 0x100u
#define BKT_MAX_CAP 7424u
#define BKT_REDO_CAP 128u

// ---- qualifiers ------------------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

struct alignas(16) uint4 { u32 x, y, z, w; };
struct alignas(16) ulonglong2 { u64 x, y; };
static inline uint4 make_uint4(u32 x, u32 y, u32 z, u32 w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline ulonglong2 make_ulonglong2(u64 x, u64 y) { ulonglong2 r; r.x = x; r.y = y; return r; }

namespace simt {

struct Dim { unsigned x, y, z; };

// barrier whose participants may leave for good (a thread that returns from the kernel)
class Barrier {
public:
    void init(int n) { expected_ = n; arrived_ = 0; gen_ = 0; }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m_);
        const unsigned g = gen_;
        if (++arrived_ == expected_) { arrived_ = 0; gen_++; cv_.notify_all(); }
        else cv_.wait(lk, [&] { return gen_ != g; });
    }
    void drop()
    {
        std::unique_lock<std::mutex> lk(m_);
        expected_--;
        if (expected_ > 0 && arrived_ == expected_) { arrived_ = 0; gen_++; cv_.notify_all(); }
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    int expected_ = 0, arrived_ = 0;
    unsigned gen_ = 0;
};

struct Warp {
    Barrier bar;
    u64 slot[32];
    u32 op[32];
    volatile bool live[32];
};

struct Block {
    Barrier bar;
    std::vector<Warp> warps;
    unsigned char *dyn = nullptr;
};

inline Block *&cur_block() { static Block *b = nullptr; return b; }
struct Tls { Dim tid, bid; int lane, warp; };
inline Tls &tls() { static thread_local Tls t; return t; }
inline Dim &block_dim() { static Dim d; return d; }
inline Dim &grid_dim() { static Dim d; return d; }
inline unsigned char *dyn_smem() { return cur_block()->dyn; }

// SIMT_EMU_JITTER=n: every lane yields at random before one in n collectives / atomics, to vary the interleavings
inline unsigned jitter_period() { static const unsigned n = getenv("SIMT_EMU_JITTER") ? (unsigned)atoi(getenv("SIMT_EMU_JITTER")) : 0u; return n; }
inline void jitter()
{
    const unsigned n = jitter_period();
    if (!n) return;
    static thread_local unsigned long long s = 0x9E3779B97F4A7C15ull ^ (unsigned long long)(uintptr_t)&s;
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    if (s % n == 0)
        for (unsigned i = 0; i < 1 + (s >> 20) % 4; i++) std::this_thread::yield();
}

[[noreturn]] inline void die(const char *what, u32 a, u32 b)
{
    fprintf(stderr, "SIMT EMU: %s (op %08x vs %08x) in block %u, warp %d, lane %d\n", what, a, b, tls().bid.x, tls().warp, tls().lane);
    fflush(stderr);
    _Exit(97);
}

// one warp rendezvous: publish (value, op), wait for the live lanes, let `read` look at the mailbox, wait again
template <typename R>
inline auto collective(u64 v, u32 op, R read) -> decltype(read((const Warp *)nullptr))
{
    Warp &w = cur_block()->warps[tls().warp];
    const int lane = tls().lane;
    jitter();
    w.slot[lane] = v;
    w.op[lane] = op;
    w.bar.wait();
    for (int j = 0; j < 32; j++)
        if (w.live[j] && w.op[j] != op) die("lanes of a warp met at different collectives", op, w.op[j]);
    auto r = read((const Warp *)&w);
    w.bar.wait();
    return r;
}

// launch: body(block index) is called by every thread of every block; blocks run one after the other
template <typename F>
inline void launch(unsigned grid, unsigned block, size_t smem, F body)
{
    if (block % 32) { fprintf(stderr, "SIMT EMU: block size must be a multiple of 32\n"); _Exit(98); }
    block_dim() = Dim{block, 1, 1};
    grid_dim() = Dim{grid, 1, 1};
    std::vector<unsigned char> dyn(smem + 64);
    for (unsigned b = 0; b < grid; b++) {
        Block blk;
        blk.bar.init((int)block);
        blk.warps = std::vector<Warp>(block / 32);
        for (auto &w : blk.warps) {
            w.bar.init(32);
            for (int j = 0; j < 32; j++) { w.live[j] = true; w.op[j] = 0; w.slot[j] = 0; }
        }
        // garbage on entry, like real shared memory
        memset(dyn.data(), 0xA5, dyn.size());
        blk.dyn = (unsigned char *)(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
        cur_block() = &blk;
        std::vector<std::thread> th;
        th.reserve(block);
        for (unsigned t = 0; t < block; t++)
            th.emplace_back([&, t] {
                Tls &me = tls();
                me.tid = Dim{t, 0, 0};
                me.bid = Dim{b, 0, 0};
                me.lane = (int)(t & 31);
                me.warp = (int)(t >> 5);
                body();
                // leaving the kernel: this thread no longer takes part in barriers
                Warp &w = blk.warps[me.warp];
                w.live[me.lane] = false;
                w.bar.drop();
                blk.bar.drop();
            });
        for (auto &x : th) x.join();
        cur_block() = nullptr;
    }
}

}  // namespace simt

#define threadIdx (simt::tls().tid)
#define blockIdx (simt::tls().bid)
#define blockDim (simt::block_dim())
#define gridDim (simt::grid_dim())

// ---- block and warp collectives ------------------------------------------------------------------------------------------
#define SIMT_OP(kind) (((u32)(kind) << 24) | (u32)__LINE__)
static inline void simt_syncthreads() { simt::cur_block()->bar.wait(); }
#define __syncthreads() simt_syncthreads()

static inline void simt_syncwarp(u32 op) { simt::collective(0, op, [](const simt::Warp *) { return 0; }); }
#define __syncwarp() simt_syncwarp(SIMT_OP(1))

template <typename T>
static inline T simt_shfl(T v, int src, u32 op)
{
    return (T)simt::collective((u64)v, op, [&](const simt::Warp *w) { return w->slot[src & 31]; });
}
template <typename T>
static inline T simt_shfl_up(T v, unsigned d, u32 op)
{
    const int lane = simt::tls().lane;
    return (T)simt::collective((u64)v, op, [&](const simt::Warp *w) { return lane >= (int)d ? w->slot[lane - (int)d] : (u64)v; });
}
template <typename T>
static inline T simt_shfl_xor(T v, int m, u32 op)
{
    const int lane = simt::tls().lane;
    return (T)simt::collective((u64)v, op, [&](const simt::Warp *w) { return w->slot[(lane ^ m) & 31]; });
}
static inline unsigned simt_ballot(bool p, u32 op)
{
    return simt::collective(p ? 1ull : 0ull, op, [&](const simt::Warp *w) {
        unsigned m = 0;
        for (int j = 0; j < 32; j++)
            if (w->live[j] && w->slot[j]) m |= 1u << j;
        return m;
    });
}
// (mask is always the full warp in the kernels tested; the emulator checks that the live lanes all arrive)
#define __shfl_sync(mask, v, src) simt_shfl((v), (int)(src), SIMT_OP(2))
#define __shfl_up_sync(mask, v, d) simt_shfl_up((v), (unsigned)(d), SIMT_OP(3))
#define __shfl_xor_sync(mask, v, m) simt_shfl_xor((v), (int)(m), SIMT_OP(4))
#define __ballot_sync(mask, p) simt_ballot((p), SIMT_OP(5))
#define __any_sync(mask, p) (simt_ballot((p), SIMT_OP(6)) != 0u)

// ---- atomics, memory ------------------------------------------------------------------------------------------------------
static inline u32 atomicAdd(u32 *p, u32 v) { simt::jitter(); return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { simt::jitter(); return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline u32 atomicOr(u32 *p, u32 v) { simt::jitter(); return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { }
    return old;
}
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long cmp, unsigned long long v)
{
    simt::jitter();
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;   // the old value either way
}
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { std::this_thread::yield(); }
static inline long long clock64() { return 0; }
static inline u32 ld_vol_u32(const u32 *p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline u64 ld_vol_u64(const u64 *p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
static inline void st_vol_u32(u32 *p, u32 v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
static inline void st_vol_u64(u64 *p, u64 v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
static inline uint4 ld_stream_v4(const uint4 *p) { return *p; }

// ---- bit intrinsics -------------------------------------------------------------------------------------------------------
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline u32 __umulhi(u32 a, u32 b) { return (u32)(((u64)a * (u64)b) >> 32); }
static inline unsigned __brev(unsigned x)
{
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(x);
}
static inline u32 __ldg(const u32 *p) { return *p; }
// per-byte compare: 0xff where the bytes are equal
static inline u32 __vcmpeq4(u32 a, u32 b)
{
    u32 r = 0;
    for (int i = 0; i < 4; i++)
        if (((a >> (8 * i)) & 0xffu) == ((b >> (8 * i)) & 0xffu)) r |= 0xffu << (8 * i);
    return r;
}
static inline u64 key_mask_d(u32 len) { return len >= 32 ? ~0ull : ((1ull << (2 * len)) - 1ull); }
static inline unsigned long long __brevll(unsigned long long x) { return ((unsigned long long)__brev((unsigned)x) << 32) | __brev((unsigned)(x >> 32)); }
