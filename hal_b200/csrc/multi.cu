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
#include <cstdio>
#include <cstdlib>
#include <cstring>

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
__global__ void unpackRecKernel(const WireParams p) {
    const int64_t step = (int64_t)gridDim.x * blockDim.x;
    const unsigned long long m40 = (1ull << 40) - 1ull;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += step) {
        const unsigned long long a = p.wire[2 * i], b = p.wire[2 * i + 1];
        halgpu_lift_rec r;
        r.start = (int64_t)(a & m40); r.end = r.start + (int64_t)(a >> 40);
        r.src_start = (int64_t)(b & m40); r.tgt_seq = (int32_t)((b >> 40) & 0xffffull); r.n_frag = (uint16_t)((b >> 56) & 0xfull);
        r.strand = strandChar((b >> 60) & 3ull); r.src_strand = strandChar(b >> 62);
        p.out[i] = r;
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

struct halgpu_comm {
    halgpu_ctx *ctx = nullptr;
    rt::Comm *comm = nullptr;
    rt::Stream stream{};         // record gathers
    rt::Stream hdrStream{};      // per-batch headers (own communicator: comm.hpp)
    uint64_t *hostHdr = nullptr; // pinned: 4 words per rank
    bool timeline = false;       // HALGPU_GATHER_TIMELINE=1: end() prints when each phase of the batch ran on the device
    std::unique_ptr<rt::Event> origin;
};

struct halgpu_gather {
    halgpu_comm *cm = nullptr;
    LiftOutput local;              // this rank's result, alive until the collectives have read it
    std::vector<uint64_t> n, nRec; // per rank
    uint32_t *wireOff = nullptr;   // all ranks' 32-bit offsets
    uint32_t *sendOff = nullptr;
    unsigned long long *sendWire = nullptr, *recvWire = nullptr; // compact records (16 bytes each), when every rank's fit
    bool compact = false, identity = false;
    uint64_t *offsets = nullptr;   // global CSR
    halgpu_lift_rec *recs = nullptr;
    std::unique_ptr<rt::Event> ready, done;
    std::unique_ptr<rt::Event> tl[3]; // timeline: lift done / gather starts / unpack done
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
        c->timeline = std::getenv("HALGPU_GATHER_TIMELINE") != nullptr;
        if (c->timeline) { c->origin.reset(new rt::Event); c->origin->record(ctx->impl->stream()); }
        c->hostHdr = static_cast<uint64_t *>(rt::hostAlloc((size_t)nranks * 4 * sizeof(uint64_t)));
        *out = c.release();
    });
}

void halgpu_comm_free(halgpu_comm *c) {
    if (c == nullptr) return;
    rt::commDestroy(c->comm);
    rt::destroyStream(c->stream);
    rt::destroyStream(c->hdrStream);
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
            C.liftover(src, tgt, flags, n, dStart, dEnd, dStrand, g->local, 0, nullptr, coalescenceLimit); // returns with the engine's stream idle
            g->kernelMs = g->local.kernelMs; g->fastMs = g->local.fastMs; g->nComplex = g->local.nComplex; g->nRetry = g->local.nRetry;
            g->launches = g->local.launches;
            if (g->local.nRec >= 0xffffffffull) throw HalError("a shard produced 2^32 or more records; use smaller batches");
            // 1. this rank's records in compact wire form (packRecKernel also decides whether they all fit), on the engine's
            //    stream.  The compact form halves the link bytes and costs one pack pass here and one unpack pass over ALL
            //    ranks' records in end(): worth it from 4 ranks up (every rank receives (W - 1) / W of the result), while 2
            //    ranks are faster with the 32-byte records gathered in place.  HALGPU_GATHER_WIRE32 / HALGPU_GATHER_WIRE16
            //    force either form (measurement switches).
            uint64_t *dHdr = static_cast<uint64_t *>(cache.take((size_t)(W + 1) * 32));
            const bool myIdentity = g->local.fastMs > 0 && g->local.nComplex == 0 && g->local.nRec == n;
            bool wantCompact = W >= 4;
            if (std::getenv("HALGPU_GATHER_WIRE32") != nullptr) wantCompact = false;
            if (std::getenv("HALGPU_GATHER_WIRE16") != nullptr) wantCompact = true;
            // header: intervals, records, "all my records fit the compact wire form", "my offsets are the identity"
            uint64_t mine[4] = {(uint64_t)n, (uint64_t)g->local.nRec, wantCompact ? 1u : 0u, myIdentity ? 1u : 0u};
            rt::h2d(dHdr + (size_t)W * 4, mine, 32, C.stream());
            if (cm->timeline) { g->tl[0].reset(new rt::Event); g->tl[0]->record(C.stream()); }
            if (wantCompact) {
                g->sendWire = static_cast<unsigned long long *>(cache.take(std::max<uint64_t>(g->local.nRec, 1) * 16));
                if (g->local.nRec > 0) {
                    WireParams wp;
                    std::memset(&wp, 0, sizeof(wp));
                    wp.recs = g->local.recs; wp.wire = g->sendWire; wp.n = (int64_t)g->local.nRec;
                    wp.fitFlag = reinterpret_cast<unsigned long long *>(dHdr + (size_t)W * 4 + 2);
                    rt::launch(packRecKernel, gridOf(wp.n, 256), 256, 0, C.stream(), wp);
                }
            }
            if (!myIdentity) { // (32-bit offsets on the wire; not needed when every rank's offsets turn out to be the identity)
                g->sendOff = static_cast<uint32_t *>(cache.take((size_t)(n + 1) * 4));
                PackOffParams pp;
                pp.off64 = g->local.offsets; pp.off32 = g->sendOff; pp.n = (int64_t)n;
                rt::launch(packOffKernel, gridOf((int64_t)n, 256), 256, 0, C.stream(), pp);
            }
            g->ready.reset(new rt::Event);
            g->done.reset(new rt::Event);
            g->ready->record(C.stream());
            // 2. every rank learns every shard's size: one 32-byte header per rank over the header communicator on its own
            //    stream -- the only host round trip of the gather, and not queued behind the previous batch's records
            g->ready->wait(cm->hdrStream);
            rt::commAllGatherSmall(cm->comm, dHdr + (size_t)W * 4, dHdr, 32, cm->hdrStream);
            rt::d2h(cm->hostHdr, dHdr, (size_t)W * 32, cm->hdrStream);
            rt::sync(cm->hdrStream);
            cache.give(dHdr);
            g->n.resize((size_t)W); g->nRec.resize((size_t)W);
            uint64_t nTotal = 0, recTotal = 0, maxN = 0;
            bool uniform = true;
            g->compact = true; g->identity = true;
            for (int r = 0; r < W; ++r) {
                g->n[(size_t)r] = cm->hostHdr[4 * r]; g->nRec[(size_t)r] = cm->hostHdr[4 * r + 1];
                g->compact = g->compact && cm->hostHdr[4 * r + 2] == 1;
                g->identity = g->identity && cm->hostHdr[4 * r + 3] == 1;
                nTotal += g->n[(size_t)r]; recTotal += g->nRec[(size_t)r];
                maxN = std::max(maxN, g->n[(size_t)r]);
                uniform = uniform && g->n[(size_t)r] == g->n[0] && g->nRec[(size_t)r] == g->nRec[0];
            }
            if (std::getenv("HALGPU_DEBUG")) fprintf(stderr, "[halgpu] gather rank %d: compact %d identity %d uniform %d records %llu\n", me, (int)g->compact, (int)g->identity, (int)uniform, (unsigned long long)recTotal);
            g->offsets = static_cast<uint64_t *>(cache.take((size_t)(nTotal + 2) * 8));
            g->recs = static_cast<halgpu_lift_rec *>(cache.take(std::max<uint64_t>(recTotal, 1) * sizeof(halgpu_lift_rec)));
            if (!g->identity) {
                g->wireOff = static_cast<uint32_t *>(cache.take((size_t)(maxN + 1) * 4 * (size_t)W));
                if (g->sendOff == nullptr) { // my offsets are the identity, some other rank's are not
                    g->sendOff = static_cast<uint32_t *>(cache.take((size_t)(n + 1) * 4));
                    PackOffParams pp;
                    pp.off64 = g->local.offsets; pp.off32 = g->sendOff; pp.n = (int64_t)n;
                    rt::launch(packOffKernel, gridOf((int64_t)n, 256), 256, 0, C.stream(), pp);
                    g->ready->record(C.stream());
                }
            }
            if (g->compact) g->recvWire = static_cast<unsigned long long *>(cache.take(std::max<uint64_t>(recTotal, 1) * 16));
            // 3. the gather on the communicator's stream: offsets and records, fused into one NCCL group
            g->ready->wait(cm->stream);
            if (cm->timeline) { g->tl[1].reset(new rt::Event); g->tl[1]->record(cm->stream); }
            const void *sendRecs = g->compact ? (const void *)g->sendWire : (const void *)g->local.recs;
            void *recvRecs = g->compact ? (void *)g->recvWire : (void *)g->recs;
            const size_t recBytes = g->compact ? 16 : sizeof(halgpu_lift_rec);
            rt::commGroupStart();
            if (uniform) {
                if (!g->identity) rt::commAllGather(cm->comm, g->sendOff, g->wireOff, (size_t)g->n[0] * 4, cm->stream);
                rt::commAllGather(cm->comm, sendRecs, recvRecs, (size_t)g->nRec[0] * recBytes, cm->stream);
            } else {
                if (!g->identity) rt::commAllGatherV(cm->comm, g->sendOff, g->wireOff, g->n, 4, (maxN + 1), cm->stream);
                rt::commAllGatherV(cm->comm, sendRecs, recvRecs, g->nRec, recBytes, 0, cm->stream);
            }
            rt::commGroupEnd();
            g->done->record(cm->stream);
            (void)me;
        } catch (...) {
            try { rt::sync(cm->stream); rt::sync(cm->hdrStream); } catch (...) {}
            C.release(g->local.offsets); C.release(g->local.recs); C.release(g->local.psl);
            cache.give(g->sendOff); cache.give(g->wireOff); cache.give(g->offsets); cache.give(g->recs);
            cache.give(g->sendWire); cache.give(g->recvWire);
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
        UnpackOffParams up;
        std::memset(&up, 0, sizeof(up));
        up.off32 = g->wireOff; up.off64 = g->offsets; up.nranks = W;
        uint64_t maxN = 0;
        bool uniform = true;
        for (int r = 0; r < W; ++r) { maxN = std::max(maxN, g->n[(size_t)r]); uniform = uniform && g->n[(size_t)r] == g->n[0] && g->nRec[(size_t)r] == g->nRec[0]; }
        for (int r = 0; r < W; ++r) {
            up.ivBase[r + 1] = up.ivBase[r] + (int64_t)g->n[(size_t)r];
            up.recBase[r + 1] = up.recBase[r] + g->nRec[(size_t)r];
            up.wireBase[r] = uniform ? up.ivBase[r] : (int64_t)r * (int64_t)(maxN + 1);
        }
        g->done->wait(C.stream()); // the engine's stream continues once the gather has landed
        if (g->identity) {
            IdentityOffParams ip;
            ip.off64 = g->offsets; ip.n = up.ivBase[W];
            rt::launch(identityOffKernel, gridOf(ip.n + 1, 256), 256, 0, C.stream(), ip);
        } else {
            rt::launch(unpackOffKernel, gridOf(up.ivBase[W] + 1, 256), 256, 0, C.stream(), up);
        }
        if (g->compact) {
            WireParams wp;
            std::memset(&wp, 0, sizeof(wp));
            wp.wire = g->recvWire; wp.out = g->recs; wp.n = (int64_t)up.recBase[W];
            if (wp.n > 0) rt::launch(unpackRecKernel, gridOf(wp.n, 256), 256, 0, C.stream(), wp);
        }
        if (cm->timeline) { g->tl[2].reset(new rt::Event); g->tl[2]->record(C.stream()); }
        rt::sync(C.stream());
        if (cm->timeline && g->tl[0] && g->tl[1] && g->tl[2]) {
            const rt::Event &o = *cm->origin;
            fprintf(stderr, "[halgpu] gather timeline rank %d (ms since the communicator was made): lift done %.3f | gather %.3f .. %.3f | unpack done %.3f | lift kernels %.3f ms, compact %d\n",
                    cm->comm->rank, rt::Event::elapsedMs(o, *g->tl[0]), rt::Event::elapsedMs(o, *g->tl[1]), rt::Event::elapsedMs(o, *g->done),
                    rt::Event::elapsedMs(o, *g->tl[2]), g->kernelMs, (int)g->compact);
        }
        halgpu_lift_result *r = static_cast<halgpu_lift_result *>(std::calloc(1, sizeof(halgpu_lift_result)));
        r->n = (size_t)up.ivBase[W]; r->n_rec = (size_t)up.recBase[W]; r->offsets = g->offsets; r->recs = g->recs; r->on_device = 1;
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
    C.release(g->local.offsets); C.release(g->local.recs); C.release(g->local.psl);
    C.cache().give(g->sendOff); C.cache().give(g->wireOff); C.cache().give(g->offsets); C.cache().give(g->recs);
    C.cache().give(g->sendWire); C.cache().give(g->recvWire);
    delete g;
    return rc;
}

} // extern "C"
