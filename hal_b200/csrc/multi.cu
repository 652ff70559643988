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
    rt::Stream stream{};
    uint64_t *hostHdr = nullptr; // pinned: 4 words per rank
};

struct halgpu_gather {
    halgpu_comm *cm = nullptr;
    LiftOutput local;              // this rank's result, alive until the collectives have read it
    std::vector<uint64_t> n, nRec; // per rank
    uint32_t *wireOff = nullptr;   // all ranks' 32-bit offsets
    uint32_t *sendOff = nullptr;
    uint64_t *offsets = nullptr;   // global CSR
    halgpu_lift_rec *recs = nullptr;
    std::unique_ptr<rt::Event> ready, done;
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
        c->hostHdr = static_cast<uint64_t *>(rt::hostAlloc((size_t)nranks * 4 * sizeof(uint64_t)));
        *out = c.release();
    });
}

void halgpu_comm_free(halgpu_comm *c) {
    if (c == nullptr) return;
    rt::commDestroy(c->comm);
    rt::destroyStream(c->stream);
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
            // 1. every rank learns every shard's size (one 32-byte header per rank; the only host round trip of the gather)
            uint64_t *dHdr = static_cast<uint64_t *>(cache.take((size_t)(W + 1) * 32));
            uint64_t mine[4] = {(uint64_t)n, (uint64_t)g->local.nRec, 0, 0};
            rt::h2d(dHdr + (size_t)W * 4, mine, 32, cm->stream);
            rt::commAllGather(cm->comm, dHdr + (size_t)W * 4, dHdr, 32, cm->stream);
            rt::d2h(cm->hostHdr, dHdr, (size_t)W * 32, cm->stream);
            rt::sync(cm->stream);
            cache.give(dHdr);
            g->n.resize((size_t)W); g->nRec.resize((size_t)W);
            uint64_t nTotal = 0, recTotal = 0, maxN = 0;
            bool uniform = true;
            for (int r = 0; r < W; ++r) {
                g->n[(size_t)r] = cm->hostHdr[4 * r]; g->nRec[(size_t)r] = cm->hostHdr[4 * r + 1];
                nTotal += g->n[(size_t)r]; recTotal += g->nRec[(size_t)r];
                maxN = std::max(maxN, g->n[(size_t)r]);
                uniform = uniform && g->n[(size_t)r] == g->n[0] && g->nRec[(size_t)r] == g->nRec[0];
            }
            // 2. this rank's offsets in wire form
            g->sendOff = static_cast<uint32_t *>(cache.take((size_t)(maxN + 1) * 4));
            g->wireOff = static_cast<uint32_t *>(cache.take((size_t)(maxN + 1) * 4 * (size_t)W));
            g->offsets = static_cast<uint64_t *>(cache.take((size_t)(nTotal + 2) * 8));
            g->recs = static_cast<halgpu_lift_rec *>(cache.take(std::max<uint64_t>(recTotal, 1) * sizeof(halgpu_lift_rec)));
            PackOffParams pp;
            pp.off64 = g->local.offsets; pp.off32 = g->sendOff; pp.n = (int64_t)n;
            rt::launch(packOffKernel, gridOf((int64_t)n, 256), 256, 0, C.stream(), pp);
            g->ready.reset(new rt::Event);
            g->done.reset(new rt::Event);
            g->ready->record(C.stream());
            g->ready->wait(cm->stream);
            // 3. the gather: offsets and records, fused into one NCCL group
            rt::commGroupStart();
            if (uniform) {
                rt::commAllGather(cm->comm, g->sendOff, g->wireOff, (size_t)g->n[0] * 4, cm->stream);
                rt::commAllGather(cm->comm, g->local.recs, g->recs, (size_t)g->nRec[0] * sizeof(halgpu_lift_rec), cm->stream);
            } else {
                rt::commAllGatherV(cm->comm, g->sendOff, g->wireOff, g->n, 4, (maxN + 1), cm->stream);
                rt::commAllGatherV(cm->comm, g->local.recs, g->recs, g->nRec, sizeof(halgpu_lift_rec), 0, cm->stream);
            }
            rt::commGroupEnd();
            g->done->record(cm->stream);
            (void)me;
        } catch (...) {
            try { rt::sync(cm->stream); } catch (...) {}
            C.release(g->local.offsets); C.release(g->local.recs); C.release(g->local.psl);
            cache.give(g->sendOff); cache.give(g->wireOff); cache.give(g->offsets); cache.give(g->recs);
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
        rt::launch(unpackOffKernel, gridOf(up.ivBase[W] + 1, 256), 256, 0, C.stream(), up);
        rt::sync(C.stream());
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
    delete g;
    return rc;
}

} // extern "C"
