// TEST HARNESS ONLY -- host emulation of the handful of CUDA constructs the kernels use, so that the
// kernel SOURCE (hal_b200/csrc/*.cuh) can be compiled with g++ and its control flow checked against the
// CPU oracle in the "-m 'not gpu'" test tier.  One OS thread per CUDA thread; warp collectives are
// implemented with a pthread barrier per warp.  Nothing in the product library is built from this file:
// hal_b200/libhalgpu.so contains the sm_100a build only and fails loudly without a GPU.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <pthread.h>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __constant__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(x) alignas(x)
#define __shared__ static /* blocks run one at a time in the emulator, so one static copy == per-block shared memory */

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct longlong2 {
    long long x, y;
};
struct ulonglong2 {
    unsigned long long x, y;
};

namespace simt {
struct Warp {
    pthread_barrier_t bar;
    uint64_t slot[32];
};
struct Block {
    pthread_barrier_t bar;
    std::vector<Warp> warps;
    uint8_t *smem;
};
struct ThreadCtx {
    dim3 tid, bid, bdim, gdim;
    Warp *warp;
    Block *block;
    int lane;
};
inline ThreadCtx &ctx() {
    static thread_local ThreadCtx c;
    return c;
}
inline uint8_t *dynamicSmem() { return ctx().block->smem; }

template <class T> inline uint64_t toBits(T v) {
    uint64_t b = 0;
    static_assert(sizeof(T) <= 8, "shuffle payload too wide");
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T> inline T fromBits(uint64_t b) {
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}
inline void warpBarrier() { pthread_barrier_wait(&ctx().warp->bar); }

inline pthread_mutex_t &launchMutex() {
    static pthread_mutex_t m = PTHREAD_MUTEX_INITIALIZER;
    return m;
}
template <class K, class P> void launch(K kernel, dim3 grid, dim3 block, size_t smemBytes, const P &param) {
    // one kernel at a time in the whole process: __shared__ variables are statics, and the multi-rank tests launch from
    // several host threads
    struct Guard {
        Guard() { pthread_mutex_lock(&launchMutex()); }
        ~Guard() { pthread_mutex_unlock(&launchMutex()); }
    } guard;
    const unsigned nb = grid.x, nt = block.x, nw = (nt + 31) / 32;
    std::vector<Block> blocks(nb);
    std::vector<std::vector<uint8_t>> smems(nb);
    for (unsigned b = 0; b < nb; ++b) {
        smems[b].assign(smemBytes + 64, 0);
        blocks[b].smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smems[b].data()) + 63) & ~uintptr_t(63));
        blocks[b].warps = std::vector<Warp>(nw);
        pthread_barrier_init(&blocks[b].bar, nullptr, nt);
        for (unsigned w = 0; w < nw; ++w) {
            unsigned lanes = (w + 1) * 32 <= nt ? 32 : nt - w * 32;
            pthread_barrier_init(&blocks[b].warps[w].bar, nullptr, lanes);
        }
    }
    for (unsigned b = 0; b < nb; ++b) { // one block at a time (see __shared__)
        std::vector<std::thread> threads;
        for (unsigned t = 0; t < nt; ++t) {
            threads.emplace_back([&, b, t]() {
                ThreadCtx &c = ctx();
                c.tid = dim3(t); c.bid = dim3(b); c.bdim = block; c.gdim = grid;
                c.block = &blocks[b];
                c.warp = &blocks[b].warps[t / 32];
                c.lane = t % 32;
                kernel(param);
            });
        }
        for (auto &th : threads) th.join();
    }
}
} // namespace simt

#define threadIdx (simt::ctx().tid)
#define blockIdx (simt::ctx().bid)
#define blockDim (simt::ctx().bdim)
#define gridDim (simt::ctx().gdim)

inline void __syncwarp(unsigned = 0xffffffffu) { simt::warpBarrier(); }
inline void __syncthreads() { pthread_barrier_wait(&simt::ctx().block->bar); }
inline unsigned __ballot_sync(unsigned, bool pred) {
    simt::Warp *w = simt::ctx().warp;
    w->slot[simt::ctx().lane] = pred ? 1 : 0;
    simt::warpBarrier();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (unsigned)(w->slot[i] & 1) << i;
    simt::warpBarrier();
    return m;
}
inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0; }
template <class T> inline T __shfl_sync(unsigned, T v, int src) {
    simt::Warp *w = simt::ctx().warp;
    w->slot[simt::ctx().lane] = simt::toBits(v);
    simt::warpBarrier();
    T r = simt::fromBits<T>(w->slot[src & 31]);
    simt::warpBarrier();
    return r;
}
template <class T> inline T __shfl_up_sync(unsigned, T v, int delta) {
    simt::Warp *w = simt::ctx().warp;
    const int lane = simt::ctx().lane;
    w->slot[lane] = simt::toBits(v);
    simt::warpBarrier();
    T r = lane >= delta ? simt::fromBits<T>(w->slot[lane - delta]) : v;
    simt::warpBarrier();
    return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline long long __double_as_longlong(double v) { return simt::fromBits<long long>(simt::toBits(v)); }
inline double __longlong_as_double(long long v) { return simt::fromBits<double>(simt::toBits(v)); }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
template <class T> inline T __ldg(const T *p) { return *p; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
