// Thin runtime layer under the engine: device memory, copies, kernel launch, events, and the two
// library primitives used for batch plumbing (radix sort of the interval keys, exclusive scan of the
// per-interval counts).  The product build maps it onto the CUDA runtime + CUB; the tests/simt harness
// build (HALGPU_SIMT_EMUL) maps it onto host memory and the thread-based warp emulator so that the
// engine's control flow (plans, retry ladder, CSR assembly) is covered by the CPU-only test tier.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

#if defined(HALGPU_SIMT_EMUL)
#include "simt_emul.h"
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <chrono>
#include <numeric>
#include <vector>
#else
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/iterator/reverse_iterator.h>
#include <cuda_runtime.h>
#include <cstring>
#endif

namespace halgpu {
namespace rt {

struct GpuError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

extern unsigned long long g_launches; // kernels launched by this library in this process

#if defined(HALGPU_SIMT_EMUL)

typedef int Stream;
inline void setDevice(int) {}
inline Stream createStream() { return 0; }
inline void destroyStream(Stream) {}
inline void *dmalloc(size_t n) { // 256-byte aligned like cudaMalloc; calloc keeps large untouched buffers lazily zeroed
    void *raw = std::calloc((n ? n : 1) + 256 + sizeof(void *), 1);
    if (!raw) throw GpuError("out of memory");
    const uintptr_t a = (reinterpret_cast<uintptr_t>(raw) + sizeof(void *) + 255) & ~(uintptr_t)255;
    reinterpret_cast<void **>(a)[-1] = raw;
    return reinterpret_cast<void *>(a);
}
inline void dfree(void *p) { if (p) std::free(reinterpret_cast<void **>(p)[-1]); }
inline void *dmallocAsync(size_t n, Stream) { return dmalloc(n); }
inline void dfreeAsync(void *p, Stream) { dfree(p); }
inline void *hostAlloc(size_t n) { return std::malloc(n ? n : 1); }
inline void hostFree(void *p) { std::free(p); }
inline void h2d(void *d, const void *h, size_t n, Stream) { if (n) std::memcpy(d, h, n); }
inline void d2h(void *h, const void *d, size_t n, Stream) { if (n) std::memcpy(h, d, n); }
inline void d2d(void *dst, const void *src, size_t n, Stream) { if (n) std::memcpy(dst, src, n); }
inline void dmemset(void *d, int v, size_t n, Stream) { if (n) std::memset(d, v, n); }
inline void sync(Stream) {}
inline int smCount() { return 2; }
inline void retainPool(int) {}
// peer memory (multi.cu): the harness's ranks are threads of one process, "device" memory is host memory
struct IpcHandle { uint8_t b[64]; };
inline void ipcExport(void *p, IpcHandle &h, uint64_t &offset) { std::memset(&h, 0, sizeof(h)); std::memcpy(h.b, &p, sizeof(p)); offset = 0; }
inline void *ipcOpen(const IpcHandle &h) { void *p; std::memcpy(&p, h.b, sizeof(p)); return p; }
inline void ipcClose(void *) {}
inline void deviceUuid(int, uint8_t out[16]) { std::memset(out, 0, 16); }
inline int deviceByUuid(const uint8_t[16]) { return 0; }
inline bool enablePeerAccess(int, int) { return true; }
inline void copyFromPeer(void *dst, const void *src, size_t n, Stream) { if (n) std::memcpy(dst, src, n); }
struct Event {
    std::chrono::steady_clock::time_point t;
    void record(Stream) { t = std::chrono::steady_clock::now(); }
    void wait(Stream) const {}
    void hostWait() const {}
    static float elapsedMs(const Event &a, const Event &b) { return std::chrono::duration<float, std::milli>(b.t - a.t).count(); }
};
template <class K, class P> inline void launch(K kernel, unsigned grid, unsigned block, size_t smem, Stream, const P &param) {
    ++g_launches;
    const unsigned cap = block >= 256 ? 1u : (block >= 128 ? 2u : 4u); // kernels are grid-stride; keep the emulated thread count small
    if (grid > cap) grid = cap;
    simt::launch(kernel, dim3(grid), dim3(block), smem, param);
}
template <class K> inline void allowSmem(K, size_t) {}
inline void sortPairsU64U32(const uint64_t *keysIn, uint64_t *keysOut, const uint32_t *valsIn, uint32_t *valsOut, size_t n, int, Stream) {
    std::vector<uint32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0u);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return keysIn[a] < keysIn[b]; });
    for (size_t i = 0; i < n; ++i) { keysOut[i] = keysIn[idx[i]]; valsOut[i] = valsIn[idx[i]]; }
}
inline void exclusiveScanU32(const uint32_t *in, uint64_t *out, size_t n, Stream) { // writes n+1 entries
    uint64_t s = 0;
    for (size_t i = 0; i < n; ++i) { out[i] = s; s += in[i]; }
    out[n] = s;
}
// positions i in [0, n) with flag[i] != 0, ascending; *dCount receives how many (out must hold n entries)
inline void selectFlagged(const uint8_t *flag, int64_t *out, unsigned long long *dCount, size_t n, Stream) {
    unsigned long long c = 0;
    for (size_t i = 0; i < n; ++i) if (flag[i]) out[c++] = (int64_t)i;
    *dCount = c;
}
// The *Tmp primitives take their temporary storage from the caller (tmp == nullptr: only report the size needed).
inline void sortPairsU64U64Tmp(void *tmp, size_t &tmpBytes, const uint64_t *keysIn, uint64_t *keysOut, const uint64_t *valsIn, uint64_t *valsOut,
                               size_t n, int beginBit, int endBit, Stream) {
    if (tmp == nullptr) { tmpBytes = 1; return; }
    std::vector<uint32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0u);
    const uint64_t mask = (endBit >= 64 ? ~0ull : ((1ull << endBit) - 1ull)) & ~((1ull << beginBit) - 1ull);
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return (keysIn[a] & mask) < (keysIn[b] & mask); });
    for (size_t i = 0; i < n; ++i) { keysOut[i] = keysIn[idx[i]]; valsOut[i] = valsIn[idx[i]]; }
}
inline void sortKeysU64Tmp(void *tmp, size_t &tmpBytes, const uint64_t *keysIn, uint64_t *keysOut, size_t n, int beginBit, int endBit, Stream) {
    if (tmp == nullptr) { tmpBytes = 1; return; }
    const uint64_t mask = (endBit >= 64 ? ~0ull : ((1ull << endBit) - 1ull)) & ~((1ull << beginBit) - 1ull);
    std::vector<uint64_t> v(keysIn, keysIn + n);
    std::stable_sort(v.begin(), v.end(), [&](uint64_t a, uint64_t b) { return (a & mask) < (b & mask); });
    std::copy(v.begin(), v.end(), keysOut);
}
// out[i] = sum of (in[j] & mask) for j < i, i in [0, n]  (n + 1 entries are read and written)
inline void exclusiveScanMaskedTmp(void *tmp, size_t &tmpBytes, const unsigned long long *in, uint64_t *out, size_t n, uint64_t mask, Stream) {
    if (tmp == nullptr) { tmpBytes = 1; return; }
    uint64_t s = 0;
    for (size_t i = 0; i <= n; ++i) { out[i] = s; s += in[i] & mask; }
}
// out[i] = min(in[i], in[i + 1], ..., in[n - 1])
inline void suffixMinU32Tmp(void *tmp, size_t &tmpBytes, const uint32_t *in, uint32_t *out, size_t n, Stream) {
    if (tmp == nullptr) { tmpBytes = 1; return; }
    uint32_t m = 0xffffffffu;
    for (size_t i = n; i-- > 0;) { m = std::min(m, in[i]); out[i] = m; }
}

#else // ---------------------------------------------------------------- CUDA

typedef cudaStream_t Stream;
inline void check(cudaError_t e, const char *what) {
    if (e != cudaSuccess) throw GpuError(std::string(what) + ": " + cudaGetErrorString(e));
}
inline void setDevice(int d) { check(cudaSetDevice(d), "cudaSetDevice"); }
inline Stream createStream() { Stream s; check(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate"); return s; }
inline void destroyStream(Stream s) { cudaStreamDestroy(s); }
inline void *dmalloc(size_t n) { void *p = nullptr; check(cudaMalloc(&p, n ? n : 1), "cudaMalloc"); return p; }
inline void dfree(void *p) { if (p) cudaFree(p); }
// per-batch scratch: stream-ordered pool allocations (no device-wide synchronisation, memory is retained by the pool)
inline void *dmallocAsync(size_t n, Stream s) { void *p = nullptr; check(cudaMallocAsync(&p, n ? n : 1, s), "cudaMallocAsync"); return p; }
inline void dfreeAsync(void *p, Stream s) { if (p) cudaFreeAsync(p, s); }
// pinned host buffers for results are recycled: cudaMallocHost of a few hundred MB costs ~100 ms
void *hostAlloc(size_t n);
void hostFree(void *p);
inline void h2d(void *d, const void *h, size_t n, Stream s) { if (n) check(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s), "H2D copy"); }
inline void d2h(void *h, const void *d, size_t n, Stream s) { if (n) check(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s), "D2H copy"); }
inline void d2d(void *dst, const void *src, size_t n, Stream s) { if (n) check(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, s), "D2D copy"); }
inline void dmemset(void *d, int v, size_t n, Stream s) { if (n) check(cudaMemsetAsync(d, v, n, s), "memset"); }
inline void sync(Stream s) { check(cudaStreamSynchronize(s), "stream synchronize"); }
inline int smCount() {
    int dev = 0, n = 0;
    check(cudaGetDevice(&dev), "cudaGetDevice");
    check(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev), "SM count");
    return n;
}
inline void retainPool(int device) { // keep freed stream-ordered memory in the pool across synchronisations
    cudaMemPool_t pool;
    check(cudaDeviceGetDefaultMemPool(&pool, device), "cudaDeviceGetDefaultMemPool");
    uint64_t keep = ~0ull;
    check(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep), "cudaMemPoolSetAttribute");
}
// ---- peer memory (multi.cu): the other ranks' result buffers are read straight over NVLink ----
struct IpcHandle { uint8_t b[64]; };
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IpcHandle carries a cudaIpcMemHandle_t");
// handle of the cudaMalloc allocation that starts at p
inline void ipcExport(void *p, IpcHandle &h, uint64_t &offset) {
    cudaIpcMemHandle_t ch;
    check(cudaIpcGetMemHandle(&ch, p), "cudaIpcGetMemHandle");
    std::memcpy(h.b, &ch, 64);
    offset = 0; // (callers export whole cudaMalloc allocations: the DeviceCache's buffers)
}
inline void *ipcOpen(const IpcHandle &h) { // (another process's allocation; peer access is enabled as needed)
    cudaIpcMemHandle_t ch;
    std::memcpy(&ch, h.b, 64);
    void *p = nullptr;
    check(cudaIpcOpenMemHandle(&p, ch, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return p;
}
inline void ipcClose(void *p) { if (p) cudaIpcCloseMemHandle(p); }
inline void deviceUuid(int dev, uint8_t out[16]) {
    cudaDeviceProp prop;
    check(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties");
    std::memcpy(out, prop.uuid.bytes, 16);
}
inline int deviceByUuid(const uint8_t uuid[16]) { // ordinal among this process's visible devices, -1: not visible
    int n = 0;
    check(cudaGetDeviceCount(&n), "cudaGetDeviceCount");
    for (int d = 0; d < n; ++d) {
        uint8_t u[16];
        deviceUuid(d, u);
        if (std::memcmp(u, uuid, 16) == 0) return d;
    }
    return -1;
}
inline bool enablePeerAccess(int myDev, int peerDev) { // the current device is myDev
    if (myDev == peerDev) return true;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, myDev, peerDev) != cudaSuccess || !can) { cudaGetLastError(); return false; }
    const cudaError_t e = cudaDeviceEnablePeerAccess(peerDev, 0);
    cudaGetLastError();
    return e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
}
inline void copyFromPeer(void *dst, const void *src, size_t n, Stream s) { // copy engines, no SMs
    if (n) check(cudaMemcpyAsync(dst, src, n, cudaMemcpyDefault, s), "peer copy");
}
struct Event {
    cudaEvent_t e = nullptr;
    Event() { check(cudaEventCreate(&e), "cudaEventCreate"); }
    ~Event() { if (e) cudaEventDestroy(e); }
    Event(const Event &) = delete;
    Event &operator=(const Event &) = delete;
    void record(Stream s) { check(cudaEventRecord(e, s), "cudaEventRecord"); }
    void wait(Stream s) const { check(cudaStreamWaitEvent(s, e, 0), "cudaStreamWaitEvent"); } // s waits for this event
    void hostWait() const { check(cudaEventSynchronize(e), "cudaEventSynchronize"); }
    static float elapsedMs(const Event &a, const Event &b) { float ms = 0; check(cudaEventElapsedTime(&ms, a.e, b.e), "cudaEventElapsedTime"); return ms; }
};
template <class K, class P> inline void launch(K kernel, unsigned grid, unsigned block, size_t smem, Stream s, const P &param) {
    ++g_launches;
    kernel<<<grid, block, smem, s>>>(param);
    check(cudaGetLastError(), "kernel launch");
}
template <class K> inline void allowSmem(K kernel, size_t bytes) {
    check(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), "cudaFuncSetAttribute(smem)");
}
inline void sortPairsU64U32(const uint64_t *keysIn, uint64_t *keysOut, const uint32_t *valsIn, uint32_t *valsOut, size_t n, int endBit, Stream s) {
    size_t tmpBytes = 0;
    check(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keysIn, keysOut, valsIn, valsOut, (int64_t)n, 0, endBit, s), "radix sort (size)");
    void *tmp = dmallocAsync(tmpBytes, s);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keysIn, keysOut, valsIn, valsOut, (int64_t)n, 0, endBit, s);
    g_launches += 1;
    dfreeAsync(tmp, s);
    check(e, "radix sort");
}
struct U32toU64 {
    __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)v; }
};
inline void exclusiveScanU32(const uint32_t *in, uint64_t *out, size_t n, Stream s) { // writes n+1 entries
    auto it = thrust::make_transform_iterator(in, U32toU64());
    size_t tmpBytes = 0;
    check(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, it, out, (int64_t)(n + 1), s), "scan (size)");
    void *tmp = dmallocAsync(tmpBytes, s);
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, it, out, (int64_t)(n + 1), s);
    g_launches += 1;
    dfreeAsync(tmp, s);
    check(e, "scan");
}
// positions i in [0, n) with flag[i] != 0, ascending; *dCount receives how many (out must hold n entries)
inline void selectFlagged(const uint8_t *flag, int64_t *out, unsigned long long *dCount, size_t n, Stream s) {
    cub::CountingInputIterator<int64_t> it(0);
    size_t tmpBytes = 0;
    check(cub::DeviceSelect::Flagged(nullptr, tmpBytes, it, flag, out, dCount, (int64_t)n, s), "select (size)");
    void *tmp = dmallocAsync(tmpBytes, s);
    cudaError_t e = cub::DeviceSelect::Flagged(tmp, tmpBytes, it, flag, out, dCount, (int64_t)n, s);
    g_launches += 1;
    dfreeAsync(tmp, s);
    check(e, "select");
}
// The *Tmp primitives take their temporary storage from the caller (tmp == nullptr: only report the size needed), so a
// batch allocates nothing once the engine's buffer cache is warm.
inline void sortPairsU64U64Tmp(void *tmp, size_t &tmpBytes, const uint64_t *keysIn, uint64_t *keysOut, const uint64_t *valsIn, uint64_t *valsOut,
                               size_t n, int beginBit, int endBit, Stream s) {
    check(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, keysIn, keysOut, valsIn, valsOut, (int64_t)n, beginBit, endBit, s), "radix sort");
    if (tmp != nullptr) g_launches += 1;
}
inline void sortKeysU64Tmp(void *tmp, size_t &tmpBytes, const uint64_t *keysIn, uint64_t *keysOut, size_t n, int beginBit, int endBit, Stream s) {
    check(cub::DeviceRadixSort::SortKeys(tmp, tmpBytes, keysIn, keysOut, (int64_t)n, beginBit, endBit, s), "radix sort (keys)");
    if (tmp != nullptr) g_launches += 1;
}
struct MaskU64 {
    uint64_t mask;
    __host__ __device__ uint64_t operator()(unsigned long long v) const { return (uint64_t)v & mask; }
};
inline void exclusiveScanMaskedTmp(void *tmp, size_t &tmpBytes, const unsigned long long *in, uint64_t *out, size_t n, uint64_t mask, Stream s) {
    auto it = thrust::make_transform_iterator(in, MaskU64{mask});
    check(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, it, out, (int64_t)(n + 1), s), "scan");
    if (tmp != nullptr) g_launches += 1;
}
struct MinU32 {
    __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a < b ? a : b; }
};
inline void suffixMinU32Tmp(void *tmp, size_t &tmpBytes, const uint32_t *in, uint32_t *out, size_t n, Stream s) {
    auto rin = thrust::make_reverse_iterator(in + n);
    auto rout = thrust::make_reverse_iterator(out + n);
    check(cub::DeviceScan::InclusiveScan(tmp, tmpBytes, rin, rout, MinU32(), (int64_t)n, s), "suffix min");
    if (tmp != nullptr) g_launches += 1;
}
#endif

} // namespace rt
} // namespace halgpu
