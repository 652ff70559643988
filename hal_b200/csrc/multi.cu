// Multi-GPU liftover: the staged index is replicated, every rank lifts its own shard of the interval
// batch, and ONE all-gather of the output interval buffer over NCCL / NVLink leaves the whole batch's result -- CSR
// offsets and records, in rank order -- on every GPU.  No collective sits inside the walk (intervals are independent:
// liftover/impl/halBlockLiftover.cpp:47, halLiftover.cpp:51), so the overlap that matters is between the gather of batch k
// and the lift of batch k+1: begin() returns as soon as the local lift is done and the collectives are enqueued on the
// communicator's own stream; end() waits for them.  The partitioning precedent in the reference is hal2mafMP.py:63-79.
//
// Per batch: a 32-byte header per rank (its interval and record counts) is all-gathered first so that every rank knows
// every shard's size; then the per-interval offsets (32-bit on the wire) and the records travel.  When all shards hold
// the same number of records -- the usual case for equal shards of collinear data -- that is one ncclAllGather straight
// from the engine's result buffer into the final array (no staging copy on either side); ragged shards use the same
// collective as one ncclGroup of per-rank broadcasts with exact sizes, which also lands every shard at its final place.
#include "../../include/halgpu.h"
#include "comm.hpp"
#include "engine.hpp"
#include <array>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <unistd.h>

using namespace halgpu;

namespace halgpu {

struct PackOffParams {
    const uint64_t *off64; // n + 1 local CSR offsets
    uint32_t *off32;       // n entries on the wire (the exclusive prefix; the total travels in the header)
    int64_t n;
};
__global__ void packOffKernel(const PackOffParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) p.off32[i] = (uint32_t)p.off64[i];
}

// Compact wire form of an output record, 16 bytes instead of 32 (the all-gather is NVLink-bound: bytes are time):
//   word 0: start (40 bits) | end - start (24 bits)
//   word 1: src_start (40) | tgt_seq (16) | n_frag (4) | strand (2) | src_strand (2)      strands: '+' 0, '-' 1, '.' 2
// A batch travels compact only if EVERY record of EVERY rank fits (fitsKernel clears this rank's flag in the header that is
// exchanged before the gather); otherwise the 32-byte records travel as they are.
__device__ __forceinline__ bool recFits(const halgpu_lift_rec &r) {
    return r.start >= 0 && r.start < (1ll << 40) && r.end >= r.start && r.end - r.start < (1ll << 24) && r.src_start >= 0 && r.src_start < (1ll << 40) &&
           r.tgt_seq >= 0 && r.tgt_seq < (1 << 16) && r.n_frag < 16;
}
__device__ __forceinline__ unsigned long long strandCode(uint8_t c) { return c == '+' ? 0ull : (c == '-' ? 1ull : 2ull); }
__device__ __forceinline__ uint8_t strandChar(unsigned long long c) { return c == 0 ? '+' : (c == 1 ? '-' : '.'); }
struct WireParams {
    const halgpu_lift_rec *recs;   // fits / pack: this rank's records
    unsigned long long *wire;      // pack: 2 words per record; unpack: all ranks' words
    halgpu_lift_rec *out;          // unpack
    unsigned long long *fitFlag;   // fits: cleared when a record does not fit
    int64_t n;
};
// (unpack: wire and out are advanced by the caller to the shard's first record)
// pack + "does every record fit": one pass over this rank's records.  Runs BEFORE the header exchange (the header carries the
// flag); when some rank's records do not fit the wire words are simply not used.
__global__ void packRecKernel(const WireParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    bool ok = true;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        const halgpu_lift_rec r = p.recs[i];
        ok = ok && recFits(r) && (r.strand == '+' || r.strand == '-' || r.strand == '.') && (r.src_strand == '+' || r.src_strand == '-' || r.src_strand == '.');
        ulonglong2 w;
        w.x = (unsigned long long)r.start | ((unsigned long long)(r.end - r.start) << 40);
        w.y = (unsigned long long)r.src_start | ((unsigned long long)(uint32_t)r.tgt_seq << 40) | ((unsigned long long)r.n_frag << 56) |
              (strandCode(r.strand) << 60) | (strandCode(r.src_strand) << 62);
        reinterpret_cast<ulonglong2 *>(p.wire)[i] = w;
    }
    if (!ok) *p.fitFlag = 0ull;
}
__global__ void unpackRecKernel(const WireParams p) { // one 16-byte load and two 16-byte stores per record
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const unsigned long long m40 = (1ull << 40) - 1ull;
    static_assert(sizeof(halgpu_lift_rec) == 32, "record layout");
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        const ulonglong2 w = reinterpret_cast<const ulonglong2 *>(p.wire)[i];
        const unsigned long long a = w.x, b = w.y;
        ulonglong2 lo, hi;
        lo.x = a & m40;                 // start
        lo.y = lo.x + (a >> 40);        // end
        hi.x = b & m40;                 // src_start
        hi.y = ((b >> 40) & 0xffffull)                                   // tgt_seq (32 bits)
               | ((unsigned long long)strandChar((b >> 60) & 3ull) << 32) // strand
               | ((unsigned long long)strandChar(b >> 62) << 40)          // src_strand
               | (((b >> 56) & 0xfull) << 48);                            // n_frag (16 bits)
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(p.out + i);
        dst[0] = lo; dst[1] = hi;
    }
}
struct IdentityOffParams {
    uint64_t *off64;
    int64_t n; // writes n + 1 entries
};
__global__ void identityOffKernel(const IdentityOffParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= p.n; i += step) p.off64[i] = (uint64_t)i;
}

#define HG_MAX_RANKS 64
struct UnpackOffParams {
    const uint32_t *off32;   // all ranks' wire offsets, rank r's at off32 + wireBase[r]
    uint64_t *off64;         // nTotal + 1 global CSR offsets
    int64_t ivBase[HG_MAX_RANKS + 1], wireBase[HG_MAX_RANKS];
    uint64_t recBase[HG_MAX_RANKS + 1];
    int32_t nranks;
};
__global__ void unpackOffKernel(const UnpackOffParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const int64_t nTotal = p.ivBase[p.nranks];
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= nTotal; g += step) {
        if (g == nTotal) { p.off64[g] = p.recBase[p.nranks]; continue; }
        int r = 0;
        while (r + 1 < p.nranks && g >= p.ivBase[r + 1]) ++r;
        p.off64[g] = p.recBase[r] + (uint64_t)p.off32[p.wireBase[r] + (g - p.ivBase[r])];
    }
}

} // namespace halgpu

// per-batch header of one rank (16 words), all-gathered before the records move
enum { H_N = 0, H_NREC = 1, H_FITS = 2 /* cleared on the device by packRecKernel */, H_IDENTITY = 3,
       H_PTR_SLOT = 4 /* this rank's send slot: records at its start ... */, H_OFF_OFFS = 5 /* ... 32-bit offsets at this byte offset */,
       H_IPC_SLOT = 8 /* .. 15: the slot's IPC handle */, H_WORDS = 16 };

struct PeerInfo { // exchanged once, when the communicator is made
    uint64_t pid, token; // same pid + token: a rank of this very process (its pointers are used as they are)
    uint8_t uuid[16];    // the GPU it runs on
};

struct halgpu_comm {
    halgpu_ctx *ctx = nullptr;
    rt::Comm *comm = nullptr;
    rt::Stream stream{};         // record transfers
    std::vector<rt::Stream> pullStreams; // peer-memory gather: copies from several peers in flight at once (one copy engine
                                         // stream moved 0.45 TB/s at 8 ranks, gpurun_out/bench_n_n8_timeline.err)
    rt::Stream hdrStream{};      // per-batch headers (own communicator: comm.hpp)
    uint64_t *hostHdr = nullptr; // pinned: H_WORDS per rank
    bool timeline = false;       // HALGPU_GATHER_TIMELINE=1: end() prints when each phase of the batch ran on the device
    std::unique_ptr<rt::Event> origin;
    // peer-memory gather: every rank reads the other ranks' result buffers straight over NVLink with the copy engines
    bool pull = false;
    std::vector<PeerInfo> peers;
    std::vector<uint8_t> sameProcess;
    std::map<std::pair<int, std::array<uint8_t, 64>>, void *> opened; // (rank, IPC handle) -> mapping in this process
    // What the other ranks read is a SEND SLOT owned by the communicator, not whatever buffer the engine's cache handed out:
    // few allocations, exported once, so the peers map each of them once (the first copy out of a newly mapped allocation
    // costs milliseconds).  A slot stays busy until the header exchange after the batch's end().
    struct SendSlot { uint8_t *buf; size_t cap; bool busy; rt::IpcHandle ipc; };
    std::vector<SendSlot> slots;
    uint64_t exchanges = 0; // header exchanges completed
    struct Zombie { uint64_t tag; int slot; };
    std::vector<Zombie> zombies;
};

struct halgpu_gather {
    halgpu_comm *cm = nullptr;
    LiftOutput local;              // this rank's result, alive until every rank has read it
    std::vector<uint64_t> n, nRec; // per rank
    uint32_t *wireOff = nullptr;   // all ranks' 32-bit offsets
    uint32_t *sendOff = nullptr;                                 // NCCL path: from the cache; peer-memory path: inside the send slot
    unsigned long long *sendWire = nullptr, *recvWire = nullptr; // compact records (16 bytes each), when every rank's fit
    int slot = -1;                                               // peer-memory path: this batch's send slot
    bool sendOffFromCache = false;
    bool compact = false, identity = false, pulled = false;
    uint64_t *offsets = nullptr;   // global CSR
    halgpu_lift_rec *recs = nullptr;
    std::unique_ptr<rt::Event> ready, done;
    std::unique_ptr<rt::Event> tl[3]; // timeline: lift done / gather starts / unpack done
    std::unique_ptr<rt::Event> pullDone[3];
    float kernelMs = 0, fastMs = 0;
    size_t nComplex = 0, nRetry = 0;
    int launches = 0;
};

namespace {
int failMsg(char **err, const std::string &msg) {
    if (err != nullptr) {
        *err = static_cast<char *>(std::malloc(msg.size() + 1));
        if (*err != nullptr) std::memcpy(*err, msg.c_str(), msg.size() + 1);
    }
    return 1;
}
template <class F> int guardedCall(char **err, F f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        return failMsg(err, e.what());
    } catch (...) {
        return failMsg(err, "unknown error");
    }
}
unsigned gridOf(int64_t n, unsigned block) {
    int64_t g = (n + block - 1) / block;
    if (g > 148 * 16) g = 148 * 16;
    return (unsigned)(g < 1 ? 1 : g);
}
} // namespace

namespace {
uint64_t processToken() {
    static const uint64_t t = (uint64_t)std::chrono::steady_clock::now().time_since_epoch().count() ^ ((uint64_t)getpid() << 32);
    return t;
}
// all ranks exchange `bytes` bytes each over the header communicator (host in, host out)
void exchangeSmall(halgpu_comm *c, const void *mine, void *all, size_t bytes) {
    DeviceCache &cache = c->ctx->impl->cache();
    const int W = c->comm->nranks;
    uint8_t *d = static_cast<uint8_t *>(cache.take((size_t)(W + 1) * bytes));
    rt::h2d(d + (size_t)W * bytes, mine, bytes, c->hdrStream);
    rt::commAllGatherSmall(c->comm, d + (size_t)W * bytes, d, bytes, c->hdrStream);
    rt::d2h(all, d, (size_t)W * bytes, c->hdrStream);
    rt::sync(c->hdrStream);
    cache.give(d);
}
void releaseZombies(halgpu_comm *c, bool all) {
    size_t keep = 0;
    for (halgpu_comm::Zombie &z : c->zombies) {
        if (all || z.tag < c->exchanges) c->slots[(size_t)z.slot].busy = false;
        else c->zombies[keep++] = z;
    }
    c->zombies.resize(keep);
}
// the send side of a batch goes back: cache buffers at once, a send slot after the next header exchange (the other ranks
// may still be copying out of it; every rank gets to that exchange only after it has ended this batch)
void releaseSendSide(halgpu_gather *g) {
    halgpu_comm *c = g->cm;
    DeviceCache &cache = c->ctx->impl->cache();
    if (g->slot >= 0) {
        halgpu_comm::Zombie z;
        z.tag = c->exchanges; z.slot = g->slot;
        c->zombies.push_back(z);
        if (g->sendOffFromCache) cache.give(g->sendOff);
    } else {
        cache.give(g->sendOff); cache.give(g->sendWire);
    }
    g->slot = -1; g->sendOff = nullptr; g->sendWire = nullptr;
}
int acquireSlot(halgpu_comm *c, size_t bytes) {
    for (size_t i = 0; i < c->slots.size(); ++i) {
        if (!c->slots[i].busy && c->slots[i].cap >= bytes) { c->slots[i].busy = true; return (int)i; }
    }
    // (smaller idle slots stay allocated until the communicator is freed: a peer may still have them mapped)
    halgpu_comm::SendSlot sl;
    sl.cap = (bytes + bytes / 4 + 4095) & ~(size_t)4095;
    sl.buf = static_cast<uint8_t *>(rt::dmalloc(sl.cap));
    sl.busy = true;
    uint64_t off = 0;
    rt::ipcExport(sl.buf, sl.ipc, off);
    c->slots.push_back(sl);
    return (int)c->slots.size() - 1;
}
} // namespace

extern "C" {

int halgpu_comm_unique_id(uint8_t id[128], char **err) {
    if (id == nullptr) return failMsg(err, "halgpu_comm_unique_id: null argument");
    return guardedCall(err, [&] { rt::commUniqueId(id); });
}


int halgpu_comm_init(halgpu_ctx *ctx, int nranks, int rank, const uint8_t id[128], halgpu_comm **out, char **err) {
    if (ctx == nullptr || id == nullptr || out == nullptr) return failMsg(err, "halgpu_comm_init: null argument");
    *out = nullptr;
    if (nranks < 1 || nranks > HG_MAX_RANKS || rank < 0 || rank >= nranks) return failMsg(err, "halgpu_comm_init: bad rank / size (at most 64 ranks)");
    return guardedCall(err, [&] {
        rt::setDevice(ctx->impl->device());
        std::unique_ptr<halgpu_comm> c(new halgpu_comm);
        c->ctx = ctx;
        c->comm = rt::commInit(nranks, rank, id);
        c->stream = rt::createStream();
        c->hdrStream = rt::createStream();
        for (int i = 0; i < std::min(3, nranks - 1); ++i) c->pullStreams.push_back(rt::createStream());
        c->timeline = std::getenv("HALGPU_GATHER_TIMELINE") != nullptr;
        if (c->timeline) { c->origin.reset(new rt::Event); c->origin->record(ctx->impl->stream()); }
        c->hostHdr = static_cast<uint64_t *>(rt::hostAlloc((size_t)nranks * H_WORDS * sizeof(uint64_t)));
        // who the other ranks are, and whether every rank can read every other rank's memory (NVLink / PCIe peer access);
        // HALGPU_GATHER_NCCL=1 (measurement switch) keeps the records on ncclAllGather
        c->peers.resize((size_t)nranks);
        c->sameProcess.assign((size_t)nranks, 0);
        PeerInfo me;
        std::memset(&me, 0, sizeof(me));
        me.pid = (uint64_t)getpid(); me.token = processToken();
        rt::deviceUuid(ctx->impl->device(), me.uuid);
        if (nranks > 1) exchangeSmall(c.get(), &me, c->peers.data(), sizeof(PeerInfo));
        else c->peers[0] = me;
        // Measured (profiles/r02_bench_n{2,4,8}_*.json): 2 ranks 1.04 ms per step either way, 4 ranks 1.44 ms against 1.56 ms through
        // ncclAllGather, 8 ranks 2.84 against 2.97 ms (one copy stream, expansion still in end()).
        uint64_t capable = 1;
        if (std::getenv("HALGPU_GATHER_PULL") != nullptr) capable = 1;
        if (std::getenv("HALGPU_GATHER_NCCL") != nullptr) capable = 0;
        for (int r = 0; r < nranks; ++r) {
            const PeerInfo &p = c->peers[(size_t)r];
            c->sameProcess[(size_t)r] = p.pid == me.pid && p.token == me.token;
            if (r == rank) continue;
            const int dev = rt::deviceByUuid(p.uuid);
            if (dev < 0 || !rt::enablePeerAccess(ctx->impl->device(), dev)) capable = 0;
        }
        std::vector<uint64_t> caps((size_t)nranks, capable);
        if (nranks > 1) exchangeSmall(c.get(), &capable, caps.data(), sizeof(uint64_t));
        c->pull = nranks > 1;
        for (uint64_t v : caps) c->pull = c->pull && v == 1;
        if (std::getenv("HALGPU_DEBUG")) fprintf(stderr, "[halgpu] communicator rank %d of %d: peer-memory gather %s\n", rank, nranks, c->pull ? "on" : "off (NCCL all-gather)");
        *out = c.release();
    });
}

void halgpu_comm_free(halgpu_comm *c) {
    if (c == nullptr) return;
    try {
        rt::setDevice(c->ctx->impl->device());
        if (c->pull) { // nobody reads anybody's buffers any more once every rank is here (the call is collective)
            uint64_t mine = 0;
            std::vector<uint64_t> all((size_t)c->comm->nranks);
            exchangeSmall(c, &mine, all.data(), sizeof(uint64_t));
            for (auto &kv : c->opened) rt::ipcClose(kv.second);
            c->opened.clear();
            exchangeSmall(c, &mine, all.data(), sizeof(uint64_t)); // (every mapping is closed before any owner frees its memory)
        }
    } catch (...) {}
    releaseZombies(c, true);
    for (halgpu_comm::SendSlot &sl : c->slots) rt::dfree(sl.buf);
    rt::commDestroy(c->comm);
    rt::destroyStream(c->stream);
    rt::destroyStream(c->hdrStream);
    for (rt::Stream ps : c->pullStreams) rt::destroyStream(ps);
    rt::hostFree(c->hostHdr);
    delete c;
}

int halgpu_comm_rank(const halgpu_comm *c) { return c ? c->comm->rank : -1; }
int halgpu_comm_size(const halgpu_comm *c) { return c ? c->comm->nranks : 0; }

int halgpu_liftover_allgather_begin(halgpu_comm *cm, int src, int tgt, int coalescenceLimit, uint32_t flags, size_t n, const int64_t *dStart,
                                    const int64_t *dEnd, const uint8_t *dStrand, halgpu_gather **out, char **err) {
    if (cm == nullptr || out == nullptr) return failMsg(err, "halgpu_liftover_allgather_begin: null argument");
    *out = nullptr;
    if ((flags & (HALGPU_PSL | HALGPU_RAW_FRAGMENTS)) != 0) return failMsg(err, "halgpu_liftover_allgather: PSL counts / raw fragments are not gathered");
    return guardedCall(err, [&] {
        Context &C = *cm->ctx->impl;
        rt::setDevice(C.device());
        const int W = cm->comm->nranks, me = cm->comm->rank;
        std::unique_ptr<halgpu_gather> g(new halgpu_gather);
        g->cm = cm;
        DeviceCache &cache = C.cache();
        try {
            // the wire form of this batch's records, decided before the lift: with the peer-memory gather of 32-byte records the
            // lift's record pool IS the send slot, so a batch the lane kernel finishes alone is read by the peers where it lies
            bool wantCompact = W >= 4;
            if (std::getenv("HALGPU_GATHER_WIRE32") != nullptr) wantCompact = false;
            if (std::getenv("HALGPU_GATHER_WIRE16") != nullptr) wantCompact = true;
            ExternalPool ext;
            size_t slotOffRegion = 0; // where the 32-bit offsets go inside the slot
            if (cm->pull && !wantCompact) {
                const int coalForPlan = (flags & (HALGPU_NO_DUPES | HALGPU_COLUMN_LIFTOVER)) != 0 ? -1 : coalescenceLimit;
                slotOffRegion = (C.poolBytesFor(src, tgt, n, coalForPlan) + 255) & ~(size_t)255;
                g->slot = acquireSlot(cm, slotOffRegion + (size_t)(n + 1) * 4);
                ext.buf = cm->slots[(size_t)g->slot].buf; ext.bytes = slotOffRegion;
            }
            C.liftover(src, tgt, flags, n, dStart, dEnd, dStrand, g->local, 0, nullptr, coalescenceLimit, ext.buf ? &ext : nullptr); // returns with the engine's stream idle
            g->kernelMs = g->local.kernelMs; g->fastMs = g->local.fastMs; g->nComplex = g->local.nComplex; g->nRetry = g->local.nRetry;
            g->launches = g->local.launches;
            if (g->local.nRec >= 0xffffffffull) throw HalError("a shard produced 2^32 or more records; use smaller batches");
            // 1. this rank's records in compact wire form (packRecKernel also decides whether they all fit), on the engine's
            //    stream.  The compact form halves the link bytes and costs one pack pass here and one expansion pass over ALL
            //    ranks' records (step 3): worth it from 4 ranks up (every rank receives (W - 1) / W of the result), while 2
            //    ranks are faster with the 32-byte records moved as they are.  HALGPU_GATHER_WIRE32 / HALGPU_GATHER_WIRE16
            //    force either form (measurement switches).
            uint64_t *dHdr = static_cast<uint64_t *>(cache.take((size_t)(W + 1) * H_WORDS * 8));
            const bool myIdentity = g->local.fastMs > 0 && g->local.nComplex == 0 && g->local.nRec == n;
            size_t recRegion = (std::max<uint64_t>(g->local.nRec, 1) * (wantCompact ? 16 : sizeof(halgpu_lift_rec)) + 255) & ~(size_t)255;
            uint8_t *slotRecs = nullptr; // peer-memory path, 32-byte form: where a COPY of the records has to go (NULL: they are there)
            if (cm->pull) {
                if (wantCompact) {
                    g->slot = acquireSlot(cm, recRegion + (myIdentity ? 0 : (size_t)(n + 1) * 4));
                    g->sendWire = reinterpret_cast<unsigned long long *>(cm->slots[(size_t)g->slot].buf);
                } else if (g->local.recsExternal) {
                    recRegion = slotOffRegion; // (the records lie at the start of the slot already)
                } else if (recRegion <= slotOffRegion) { // gathered by the engine into a cache buffer (complex intervals): copy
                    recRegion = slotOffRegion;
                    slotRecs = cm->slots[(size_t)g->slot].buf;
                } else { // more records than the slot was sized for: take a larger one (nobody has been told about the first)
                    cm->slots[(size_t)g->slot].busy = false;
                    g->slot = acquireSlot(cm, recRegion + (size_t)(n + 1) * 4);
                    slotRecs = cm->slots[(size_t)g->slot].buf;
                }
                if (!myIdentity) g->sendOff = reinterpret_cast<uint32_t *>(cm->slots[(size_t)g->slot].buf + recRegion);
            } else {
                if (wantCompact) g->sendWire = static_cast<unsigned long long *>(cache.take(std::max<uint64_t>(g->local.nRec, 1) * 16));
                if (!myIdentity) g->sendOff = static_cast<uint32_t *>(cache.take((size_t)(n + 1) * 4)); // (32-bit offsets on the wire)
            }
            uint64_t mine[H_WORDS];
            std::memset(mine, 0, sizeof(mine));
            mine[H_N] = (uint64_t)n; mine[H_NREC] = (uint64_t)g->local.nRec; mine[H_FITS] = wantCompact ? 1u : 0u; mine[H_IDENTITY] = myIdentity ? 1u : 0u;
            if (cm->pull) { // where the other ranks find this rank's shard
                mine[H_PTR_SLOT] = (uint64_t)(uintptr_t)cm->slots[(size_t)g->slot].buf;
                mine[H_OFF_OFFS] = (uint64_t)recRegion;
                std::memcpy(&mine[H_IPC_SLOT], cm->slots[(size_t)g->slot].ipc.b, 64);
            }
            rt::h2d(dHdr + (size_t)W * H_WORDS, mine, sizeof(mine), C.stream());
            if (cm->timeline) { g->tl[0].reset(new rt::Event); g->tl[0]->record(C.stream()); }
            if (wantCompact && g->local.nRec > 0) {
                WireParams wp;
                std::memset(&wp, 0, sizeof(wp));
                wp.recs = g->local.recs; wp.wire = g->sendWire; wp.n = (int64_t)g->local.nRec;
                wp.fitFlag = reinterpret_cast<unsigned long long *>(dHdr + (size_t)W * H_WORDS + H_FITS);
                rt::launch(packRecKernel, gridOf(wp.n, 256), 256, 0, C.stream(), wp);
            }
            if (slotRecs != nullptr && g->local.nRec > 0) rt::d2d(slotRecs, g->local.recs, (size_t)g->local.nRec * sizeof(halgpu_lift_rec), C.stream());
            auto packOffsets = [&] {
                PackOffParams pp;
                pp.off64 = g->local.offsets; pp.off32 = g->sendOff; pp.n = (int64_t)n;
                rt::launch(packOffKernel, gridOf((int64_t)n, 256), 256, 0, C.stream(), pp);
            };
            if (!myIdentity) packOffsets();
            g->ready.reset(new rt::Event);
            g->done.reset(new rt::Event);
            g->ready->record(C.stream());
            // 2. every rank learns every shard's size (and buffers): one 128-byte header per rank over the header communicator
            //    on its own stream -- the only host round trip of the gather, and not queued behind the previous batch's
            //    records.  With the peer-memory gather a rank's header also says "my buffers are complete": the engine's
            //    stream is drained first.
            if (cm->pull) rt::sync(C.stream());
            g->ready->wait(cm->hdrStream);
            rt::commAllGatherSmall(cm->comm, dHdr + (size_t)W * H_WORDS, dHdr, H_WORDS * 8, cm->hdrStream);
            rt::d2h(cm->hostHdr, dHdr, (size_t)W * H_WORDS * 8, cm->hdrStream);
            rt::sync(cm->hdrStream);
            cache.give(dHdr);
            ++cm->exchanges;
            releaseZombies(cm, false); // every rank has finished the batches it ended before this exchange
            const uint64_t *H = cm->hostHdr;
            g->n.resize((size_t)W); g->nRec.resize((size_t)W);
            uint64_t nTotal = 0, recTotal = 0, maxN = 0;
            bool uniform = true, anyIdentity = false;
            g->compact = true; g->identity = true;
            for (int r = 0; r < W; ++r) {
                const uint64_t *h = H + (size_t)r * H_WORDS;
                g->n[(size_t)r] = h[H_N]; g->nRec[(size_t)r] = h[H_NREC];
                g->compact = g->compact && h[H_FITS] == 1;
                g->identity = g->identity && h[H_IDENTITY] == 1;
                anyIdentity = anyIdentity || h[H_IDENTITY] == 1;
                nTotal += g->n[(size_t)r]; recTotal += g->nRec[(size_t)r];
                maxN = std::max(maxN, g->n[(size_t)r]);
                uniform = uniform && g->n[(size_t)r] == g->n[0] && g->nRec[(size_t)r] == g->nRec[0];
            }
            // (a rank whose offsets are the identity has packed none: in the rare mixed batch they are packed now and travel by NCCL)
            // (likewise a batch that wanted the compact form while some rank's records do not fit it)
            g->pulled = cm->pull && !(anyIdentity && !g->identity) && g->compact == wantCompact;
            if (std::getenv("HALGPU_DEBUG")) fprintf(stderr, "[halgpu] gather rank %d: compact %d identity %d uniform %d records %llu, %s\n", me, (int)g->compact, (int)g->identity, (int)uniform, (unsigned long long)recTotal, g->pulled ? "peer copies" : "NCCL");
            g->offsets = static_cast<uint64_t *>(cache.take((size_t)(nTotal + 2) * 8));
            g->recs = static_cast<halgpu_lift_rec *>(cache.take(std::max<uint64_t>(recTotal, 1) * sizeof(halgpu_lift_rec)));
            if (!g->identity) {
                g->wireOff = static_cast<uint32_t *>(cache.take((size_t)(maxN + 1) * 4 * (size_t)W));
                if (g->sendOff == nullptr) { // my offsets are the identity, some other rank's are not (NCCL path)
                    g->sendOff = static_cast<uint32_t *>(cache.take((size_t)(n + 1) * 4));
                    g->sendOffFromCache = true;
                    packOffsets();
                    g->ready->record(C.stream());
                }
            }
            if (g->compact) g->recvWire = static_cast<unsigned long long *>(cache.take(std::max<uint64_t>(recTotal, 1) * 16));
            // 3. the records (and offsets) of every rank, and their expansion into the result's form, all on the communicator's
            //    streams: the engine's stream is free for the next batch's lift, end() only waits
            g->ready->wait(cm->stream);
            if (cm->timeline) { g->tl[1].reset(new rt::Event); g->tl[1]->record(cm->stream); }
            UnpackOffParams up;
            std::memset(&up, 0, sizeof(up));
            up.off32 = g->wireOff; up.off64 = g->offsets; up.nranks = W;
            for (int r = 0; r < W; ++r) {
                up.ivBase[r + 1] = up.ivBase[r] + (int64_t)g->n[(size_t)r];
                up.recBase[r + 1] = up.recBase[r] + g->nRec[(size_t)r];
                up.wireBase[r] = uniform ? up.ivBase[r] : (int64_t)r * (int64_t)(maxN + 1);
            }
            if (g->identity) { // offsets[i] = i: nothing to wait for
                IdentityOffParams ip;
                ip.off64 = g->offsets; ip.n = up.ivBase[W];
                rt::launch(identityOffKernel, gridOf(ip.n + 1, 256), 256, 0, cm->stream, ip);
            }
            auto unpackShard = [&](const unsigned long long *wire, uint64_t firstRec, uint64_t count, rt::Stream st) {
                if (count == 0) return;
                WireParams wp;
                std::memset(&wp, 0, sizeof(wp));
                wp.wire = const_cast<unsigned long long *>(wire); wp.out = g->recs + firstRec; wp.n = (int64_t)count;
                rt::launch(unpackRecKernel, gridOf(wp.n, 256), 256, 0, st, wp);
            };
            const void *sendRecs = g->compact ? (const void *)g->sendWire : (const void *)g->local.recs;
            uint8_t *recvRecs = g->compact ? reinterpret_cast<uint8_t *>(g->recvWire) : reinterpret_cast<uint8_t *>(g->recs);
            const size_t recBytes = g->compact ? 16 : sizeof(halgpu_lift_rec);
            if (g->pulled) {
                // Peer-memory gather: this rank copies every other rank's shard out of that rank's own buffer, straight to its
                // final place, with the copy engines over NVLink (no SMs, no staging, no NCCL kernel competing with the next
                // batch's lift).  Rank me starts with rank me + 1 and goes round, so that at any time every rank serves one reader.
                uint64_t recAt[HG_MAX_RANKS + 1];
                recAt[0] = 0;
                for (int r = 0; r < W; ++r) recAt[r + 1] = recAt[r] + g->nRec[(size_t)r];
                auto peerSlot = [&](int r) -> const uint8_t * {
                    const uint64_t *h = H + (size_t)r * H_WORDS;
                    if (cm->sameProcess[(size_t)r]) return reinterpret_cast<const uint8_t *>((uintptr_t)h[H_PTR_SLOT]);
                    std::array<uint8_t, 64> key;
                    std::memcpy(key.data(), &h[H_IPC_SLOT], 64);
                    auto it = cm->opened.find(std::make_pair(r, key));
                    if (it == cm->opened.end()) {
                        rt::IpcHandle ih;
                        std::memcpy(ih.b, key.data(), 64);
                        it = cm->opened.emplace(std::make_pair(r, key), rt::ipcOpen(ih)).first;
                    }
                    return static_cast<const uint8_t *>(it->second);
                };
                const size_t nPull = cm->pullStreams.size();
                for (size_t i = 0; i < nPull; ++i) g->ready->wait(cm->pullStreams[i]);
                for (int k = 0; k < W; ++k) {
                    const int r = (me + 1 + k) % W; // (my own shard last: a local copy)
                    const rt::Stream cs = (r == me || nPull == 0) ? cm->stream : cm->pullStreams[(size_t)k % nPull];
                    if (g->nRec[(size_t)r] > 0) {
                        if (r == me && g->compact) { // my own shard is expanded straight out of the send slot
                            unpackShard(g->sendWire, recAt[r], g->nRec[(size_t)r], cs);
                        } else {
                            const void *src = r == me ? sendRecs : (const void *)peerSlot(r);
                            rt::copyFromPeer(recvRecs + recAt[r] * recBytes, src, (size_t)g->nRec[(size_t)r] * recBytes, cs);
                            // (each shard is expanded behind its own copy, while the next peer's copy is in flight)
                            if (g->compact) unpackShard(g->recvWire + 2 * recAt[r], recAt[r], g->nRec[(size_t)r], cs);
                        }
                    }
                    if (!g->identity && g->n[(size_t)r] > 0) {
                        const void *src = r == me ? (const void *)g->sendOff : (const void *)(peerSlot(r) + H[(size_t)r * H_WORDS + H_OFF_OFFS]);
                        const uint64_t at = uniform ? (uint64_t)r * g->n[0] : (uint64_t)r * (maxN + 1);
                        rt::copyFromPeer(g->wireOff + at, src, (size_t)g->n[(size_t)r] * 4, cs);
                    }
                }
                for (size_t i = 0; i < nPull; ++i) { // the communicator's stream (and with it `done`) waits for every copy
                    g->pullDone[i].reset(new rt::Event);
                    g->pullDone[i]->record(cm->pullStreams[i]);
                    g->pullDone[i]->wait(cm->stream);
                }
            } else {
                rt::commGroupStart();
                if (uniform) {
                    if (!g->identity) rt::commAllGather(cm->comm, g->sendOff, g->wireOff, (size_t)g->n[0] * 4, cm->stream);
                    rt::commAllGather(cm->comm, sendRecs, recvRecs, (size_t)g->nRec[0] * recBytes, cm->stream);
                } else {
                    if (!g->identity) rt::commAllGatherV(cm->comm, g->sendOff, g->wireOff, g->n, 4, (maxN + 1), cm->stream);
                    rt::commAllGatherV(cm->comm, sendRecs, recvRecs, g->nRec, recBytes, 0, cm->stream);
                }
                rt::commGroupEnd();
                if (g->compact) unpackShard(g->recvWire, 0, recTotal, cm->stream);
            }
            if (!g->identity) rt::launch(unpackOffKernel, gridOf(up.ivBase[W] + 1, 256), 256, 0, cm->stream, up);
            if (cm->timeline) { g->tl[2].reset(new rt::Event); g->tl[2]->record(cm->stream); }
            g->done->record(cm->stream);
            (void)me;
        } catch (...) {
            try { rt::sync(cm->stream); rt::sync(cm->hdrStream); } catch (...) {}
            C.release(g->local.offsets); if (!g->local.recsExternal) C.release(g->local.recs); C.release(g->local.psl);
            releaseSendSide(g.get());
            cache.give(g->wireOff); cache.give(g->offsets); cache.give(g->recs); cache.give(g->recvWire);
            throw;
        }
        *out = g.release();
    });
}

int halgpu_liftover_allgather_end(halgpu_gather *g, halgpu_lift_result **out, size_t *nPerRank, size_t *nRecPerRank, char **err) {
    if (g == nullptr || out == nullptr) return failMsg(err, "halgpu_liftover_allgather_end: null argument");
    *out = nullptr;
    halgpu_comm *cm = g->cm;
    Context &C = *cm->ctx->impl;
    const int rc = guardedCall(err, [&] {
        rt::setDevice(C.device());
        const int W = cm->comm->nranks;
        uint64_t nTotal = 0, recTotal = 0;
        for (int r = 0; r < W; ++r) { nTotal += g->n[(size_t)r]; recTotal += g->nRec[(size_t)r]; }
        g->done->hostWait(); // everything of this batch ran on the communicator's streams (begin)
        if (cm->timeline && g->tl[0] && g->tl[1] && g->tl[2]) {
            const rt::Event &o = *cm->origin;
            fprintf(stderr, "[halgpu] gather timeline rank %d (ms since the communicator was made): lift done %.3f | gather + expansion %.3f .. %.3f | lift kernels %.3f ms, compact %d, %s\n",
                    cm->comm->rank, rt::Event::elapsedMs(o, *g->tl[0]), rt::Event::elapsedMs(o, *g->tl[1]), rt::Event::elapsedMs(o, *g->tl[2]),
                    g->kernelMs, (int)g->compact, g->pulled ? "peer copies" : "NCCL");
        }
        halgpu_lift_result *r = static_cast<halgpu_lift_result *>(std::calloc(1, sizeof(halgpu_lift_result)));
        r->n = (size_t)nTotal; r->n_rec = (size_t)recTotal; r->offsets = g->offsets; r->recs = g->recs; r->on_device = 1;
        r->kernel_ms = g->kernelMs; r->fast_ms = g->fastMs; r->n_complex = g->nComplex; r->n_retry = g->nRetry; r->launches = g->launches + 2;
        r->owner = cm->ctx;
        g->offsets = nullptr; g->recs = nullptr;
        for (int k = 0; k < W; ++k) {
            if (nPerRank) nPerRank[k] = (size_t)g->n[(size_t)k];
            if (nRecPerRank) nRecPerRank[k] = (size_t)g->nRec[(size_t)k];
        }
        *out = r;
    });
    if (rc != 0) { try { rt::sync(cm->stream); } catch (...) {} }
    C.release(g->local.offsets); if (!g->local.recsExternal) C.release(g->local.recs); C.release(g->local.psl);
    releaseSendSide(g);
    C.cache().give(g->wireOff); C.cache().give(g->offsets); C.cache().give(g->recs); C.cache().give(g->recvWire);
    delete g;
    return rc;
}

} // extern "C"
