// Collective layer under the multi-GPU entry points: one all-gather per batch over NCCL / NVLink.
//
// The product build resolves NCCL at run time (dlopen "libnccl.so.2": inside a torch process that is the NCCL torch
// already loaded, in a plain C++ client the system library), so libhalgpu.so itself has no link-time NCCL dependency and
// the single-GPU path works on machines without NCCL.  The tests/simt harness build (HALGPU_SIMT_EMUL) replaces it with
// an in-process rendezvous between host threads -- one thread per emulated rank -- so the packing / offset logic of the
// multi-GPU path is covered by the CPU-only test tier (world size 2).
#pragma once
#include "rt.hpp"
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#if !defined(HALGPU_SIMT_EMUL)
#include <dlfcn.h>
#include <nccl.h>
#else
#include <condition_variable>
#endif

namespace halgpu {
namespace rt {

#if defined(HALGPU_SIMT_EMUL)

struct CommGroup { // shared by the ranks of one communicator
    std::mutex m;
    std::condition_variable cv;
    int nranks = 0, arrived = 0, generation = 0;
    std::vector<const void *> send;
    void barrier() {
        std::unique_lock<std::mutex> g(m);
        const int gen = generation;
        if (++arrived == nranks) { arrived = 0; ++generation; cv.notify_all(); }
        else cv.wait(g, [&] { return generation != gen; });
    }
};
struct Comm {
    CommGroup *group;
    int nranks, rank;
};
inline std::map<std::string, CommGroup *> &commRegistry() { static std::map<std::string, CommGroup *> r; return r; }
inline std::mutex &commRegistryMutex() { static std::mutex m; return m; }
inline void commUniqueId(uint8_t id[128]) {
    static int counter = 0;
    std::lock_guard<std::mutex> g(commRegistryMutex());
    std::memset(id, 0, 128);
    std::snprintf(reinterpret_cast<char *>(id), 128, "halgpu-emul-comm-%d", ++counter);
}
inline Comm *commInit(int nranks, int rank, const uint8_t id[128]) {
    std::lock_guard<std::mutex> g(commRegistryMutex());
    CommGroup *&grp = commRegistry()[std::string(reinterpret_cast<const char *>(id))];
    if (grp == nullptr) { grp = new CommGroup; grp->nranks = nranks; grp->send.assign((size_t)nranks, nullptr); }
    if (grp->nranks != nranks || rank < 0 || rank >= nranks) throw GpuError("communicator: inconsistent rank / size");
    return new Comm{grp, nranks, rank};
}
inline void commDestroy(Comm *c) { delete c; }
inline void commGroupStart() {}
inline void commGroupEnd() {}
inline void commAllGather(Comm *c, const void *send, void *recv, size_t bytesPerRank, Stream) {
    c->group->send[(size_t)c->rank] = send;
    c->group->barrier();
    for (int r = 0; r < c->nranks; ++r) std::memcpy(static_cast<uint8_t *>(recv) + (size_t)r * bytesPerRank, c->group->send[(size_t)r], bytesPerRank);
    c->group->barrier();
}

inline void commAllGatherSmall(Comm *c, const void *send, void *recv, size_t bytesPerRank, Stream s) { commAllGather(c, send, recv, bytesPerRank, s); }

// ragged all-gather: rank r contributes count[r] elements of elemBytes; they land at recv + r * slotElems * elemBytes
// (slotElems > 0) or packed back to back in rank order (slotElems == 0)
inline void commAllGatherV(Comm *c, const void *send, void *recv, const std::vector<uint64_t> &count, size_t elemBytes, uint64_t slotElems, Stream) {
    c->group->send[(size_t)c->rank] = send;
    c->group->barrier();
    uint64_t at = 0;
    for (int r = 0; r < c->nranks; ++r) {
        const uint64_t base = slotElems ? (uint64_t)r * slotElems : at;
        std::memcpy(static_cast<uint8_t *>(recv) + base * elemBytes, c->group->send[(size_t)r], count[(size_t)r] * elemBytes);
        at += count[(size_t)r];
    }
    c->group->barrier();
}

#else // ---------------------------------------------------------------- NCCL

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*getUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*commInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*commDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*commSplit)(ncclComm_t, int, int, ncclComm_t *, void *) = nullptr; // optional (NCCL >= 2.18)
    ncclResult_t (*allGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*groupStart)() = nullptr;
    ncclResult_t (*groupEnd)() = nullptr;
    const char *(*getErrorString)(ncclResult_t) = nullptr;
    static NcclApi &get() {
        static NcclApi api;
        static std::once_flag once;
        std::call_once(once, [] {
            const char *names[] = {"libnccl.so.2", "libnccl.so"};
            for (const char *n : names) {
                api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
                if (api.lib) break;
            }
            if (!api.lib) return;
            auto sym = [&](const char *s) { return dlsym(api.lib, s); };
            api.getUniqueId = reinterpret_cast<decltype(api.getUniqueId)>(sym("ncclGetUniqueId"));
            api.commInitRank = reinterpret_cast<decltype(api.commInitRank)>(sym("ncclCommInitRank"));
            api.commDestroy = reinterpret_cast<decltype(api.commDestroy)>(sym("ncclCommDestroy"));
            api.commSplit = reinterpret_cast<decltype(api.commSplit)>(sym("ncclCommSplit"));
            api.allGather = reinterpret_cast<decltype(api.allGather)>(sym("ncclAllGather"));
            api.broadcast = reinterpret_cast<decltype(api.broadcast)>(sym("ncclBroadcast"));
            api.groupStart = reinterpret_cast<decltype(api.groupStart)>(sym("ncclGroupStart"));
            api.groupEnd = reinterpret_cast<decltype(api.groupEnd)>(sym("ncclGroupEnd"));
            api.getErrorString = reinterpret_cast<decltype(api.getErrorString)>(sym("ncclGetErrorString"));
        });
        if (!api.lib || !api.getUniqueId || !api.commInitRank || !api.commDestroy || !api.allGather || !api.broadcast || !api.groupStart || !api.groupEnd) {
            throw GpuError("NCCL (libnccl.so.2) is not available: the multi-GPU entry points need it");
        }
        return api;
    }
    void check(ncclResult_t r, const char *what) const {
        if (r != ncclSuccess) throw GpuError(std::string(what) + ": " + (getErrorString ? getErrorString(r) : "NCCL error"));
    }
};
struct Comm {
    ncclComm_t c;
    int nranks, rank;
    // a second communicator over the same ranks for the per-batch 32-byte headers: NCCL runs the operations of ONE communicator
    // in issue order, so on `c` the header of batch k+1 would wait behind the record gather of batch k.  Equal to c when
    // ncclCommSplit is unavailable.
    ncclComm_t small;
};
static_assert(NCCL_UNIQUE_ID_BYTES == 128, "halgpu_comm_unique_id hands out 128-byte ids");
inline void commUniqueId(uint8_t id[128]) {
    NcclApi &a = NcclApi::get();
    ncclUniqueId u;
    a.check(a.getUniqueId(&u), "ncclGetUniqueId");
    std::memcpy(id, u.internal, 128);
}
inline Comm *commInit(int nranks, int rank, const uint8_t id[128]) {
    NcclApi &a = NcclApi::get();
    ncclUniqueId u;
    std::memcpy(u.internal, id, 128);
    ncclComm_t c;
    a.check(a.commInitRank(&c, nranks, u, rank), "ncclCommInitRank");
    ncclComm_t small = c;
    if (a.commSplit != nullptr && nranks > 1 && std::getenv("HALGPU_ONE_COMM") == nullptr) {
        ncclComm_t c2 = nullptr;
        if (a.commSplit(c, 0, rank, &c2, nullptr) == ncclSuccess && c2 != nullptr) small = c2;
    }
    return new Comm{c, nranks, rank, small};
}
inline void commDestroy(Comm *c) {
    if (c == nullptr) return;
    if (c->small != c->c) NcclApi::get().commDestroy(c->small);
    NcclApi::get().commDestroy(c->c);
    delete c;
}
inline void commGroupStart() { NcclApi &a = NcclApi::get(); a.check(a.groupStart(), "ncclGroupStart"); }
inline void commGroupEnd() { NcclApi &a = NcclApi::get(); a.check(a.groupEnd(), "ncclGroupEnd"); }
inline void commAllGather(Comm *c, const void *send, void *recv, size_t bytesPerRank, Stream s) {
    NcclApi &a = NcclApi::get();
    a.check(a.allGather(send, recv, bytesPerRank, ncclUint8, c->c, s), "ncclAllGather");
}

inline void commAllGatherSmall(Comm *c, const void *send, void *recv, size_t bytesPerRank, Stream s) {
    NcclApi &a = NcclApi::get();
    a.check(a.allGather(send, recv, bytesPerRank, ncclUint8, c->small, s), "ncclAllGather (header)");
}

// ragged all-gather (see the harness version above): one broadcast per rank with its exact size; inside an ncclGroup NCCL
// fuses them into one launch
inline void commAllGatherV(Comm *c, const void *send, void *recv, const std::vector<uint64_t> &count, size_t elemBytes, uint64_t slotElems, Stream s) {
    NcclApi &a = NcclApi::get();
    uint64_t at = 0;
    for (int r = 0; r < c->nranks; ++r) {
        const uint64_t base = slotElems ? (uint64_t)r * slotElems : at;
        uint8_t *dst = static_cast<uint8_t *>(recv) + base * elemBytes;
        if (count[(size_t)r] > 0) a.check(a.broadcast(r == c->rank ? send : dst, dst, count[(size_t)r] * elemBytes, ncclUint8, r, c->c, s), "ncclBroadcast");
        at += count[(size_t)r];
    }
}

#endif

} // namespace rt
} // namespace halgpu
