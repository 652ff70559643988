// C ABI (include/halgpu.h) over the engine.  Error convention of blockViz/inc/halBlockViz.h: non-zero
// return + malloc'd message in *err.
#include "../../include/halgpu.h"
#include "engine.hpp"
#include <cstdlib>
#include <cstring>

using namespace halgpu;

namespace {
int fail(char **err, const std::string &msg) {
    if (err != nullptr) {
        *err = static_cast<char *>(std::malloc(msg.size() + 1));
        if (*err != nullptr) std::memcpy(*err, msg.c_str(), msg.size() + 1);
    }
    return 1;
}
template <class F> int guarded(char **err, F f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        return fail(err, e.what());
    } catch (...) {
        return fail(err, "unknown error");
    }
}
const GenomeInfo *genome(const halgpu_ctx *ctx, int g) {
    if (ctx == nullptr) return nullptr;
    const auto &G = ctx->impl->file().genomes();
    return (g >= 0 && g < (int)G.size()) ? &G[g] : nullptr;
}
} // namespace

static_assert(sizeof(halgpu_frag) == sizeof(halgpu_lift_rec), "fragment records share the result buffer");

extern "C" {

int halgpu_open(const char *path, int device, halgpu_ctx **out, char **err) {
    if (out == nullptr || path == nullptr) return fail(err, "halgpu_open: null argument");
    *out = nullptr;
    return guarded(err, [&] {
        std::unique_ptr<halgpu_ctx> c(new halgpu_ctx);
        c->impl.reset(new Context(path, device));
        const auto &G = c->impl->file().genomes();
        c->seqTables.resize(G.size());
        for (size_t g = 0; g < G.size(); ++g) {
            for (const SequenceInfo &s : G[g].sequences) {
                halgpu_seq q;
                q.name = s.name.c_str(); q.start = s.start; q.length = s.length; q.num_top = s.numTop; q.num_bottom = s.numBottom;
                c->seqTables[g].push_back(q);
            }
        }
        *out = c.release();
    });
}

void halgpu_close(halgpu_ctx *ctx) { delete ctx; }

int halgpu_num_genomes(const halgpu_ctx *ctx) { return ctx ? (int)ctx->impl->file().genomes().size() : 0; }
const char *halgpu_genome_name(const halgpu_ctx *ctx, int g) { const GenomeInfo *i = genome(ctx, g); return i ? i->name.c_str() : nullptr; }
int halgpu_genome_id(const halgpu_ctx *ctx, const char *name) { return (ctx && name) ? ctx->impl->file().genomeId(name) : -1; }
int halgpu_genome_parent(const halgpu_ctx *ctx, int g) { const GenomeInfo *i = genome(ctx, g); return i ? i->parent : -1; }
int halgpu_genome_num_children(const halgpu_ctx *ctx, int g) { const GenomeInfo *i = genome(ctx, g); return i ? (int)i->children.size() : 0; }
int halgpu_genome_child(const halgpu_ctx *ctx, int g, int slot) {
    const GenomeInfo *i = genome(ctx, g);
    return (i && slot >= 0 && slot < (int)i->children.size()) ? i->children[slot] : -1;
}
int64_t halgpu_genome_length(const halgpu_ctx *ctx, int g) { const GenomeInfo *i = genome(ctx, g); return i ? i->length : -1; }
int64_t halgpu_genome_num_top(const halgpu_ctx *ctx, int g) { const GenomeInfo *i = genome(ctx, g); return i ? i->numTop : -1; }
int64_t halgpu_genome_num_bottom(const halgpu_ctx *ctx, int g) { const GenomeInfo *i = genome(ctx, g); return i ? i->numBottom : -1; }
const char *halgpu_newick(const halgpu_ctx *ctx) { return ctx ? ctx->impl->file().newick().c_str() : nullptr; }
int halgpu_sequence_table(const halgpu_ctx *ctx, int g, const halgpu_seq **out, size_t *n) {
    if (genome(ctx, g) == nullptr || out == nullptr || n == nullptr) return 1;
    *out = ctx->seqTables[g].data();
    *n = ctx->seqTables[g].size();
    return 0;
}
int halgpu_mrca(const halgpu_ctx *ctx, int a, int b) {
    if (genome(ctx, a) == nullptr || genome(ctx, b) == nullptr) return -1;
    return ctx->impl->file().mrca(a, b);
}
size_t halgpu_staged_bytes(const halgpu_ctx *ctx) { return ctx ? ctx->impl->stagedBytes() : 0; }
void *halgpu_stream(const halgpu_ctx *ctx) { return ctx ? (void *)(uintptr_t)ctx->impl->stream() : nullptr; }

static int liftDeviceWithBase(halgpu_ctx *ctx, int src, int tgt, int coalescenceLimit, uint32_t flags, size_t n, const int64_t *dStart,
                              const int64_t *dEnd, const uint8_t *dStrand, uint64_t offsetBase, halgpu_lift_result **out, char **err);

int halgpu_liftover_device(halgpu_ctx *ctx, int src, int tgt, int coalescenceLimit, uint32_t flags, size_t n,
                           const int64_t *dStart, const int64_t *dEnd, const uint8_t *dStrand,
                           halgpu_lift_result **out, char **err) {
    return liftDeviceWithBase(ctx, src, tgt, coalescenceLimit, flags, n, dStart, dEnd, dStrand, 0, out, err);
}

static int liftDeviceWithBase(halgpu_ctx *ctx, int src, int tgt, int coalescenceLimit, uint32_t flags, size_t n, const int64_t *dStart,
                              const int64_t *dEnd, const uint8_t *dStrand, uint64_t offsetBase, halgpu_lift_result **out, char **err) {
    if (ctx == nullptr || out == nullptr) return fail(err, "halgpu_liftover: null argument");
    *out = nullptr;
    if (coalescenceLimit < -1 || coalescenceLimit >= halgpu_num_genomes(ctx)) return fail(err, "halgpu_liftover: coalescence limit genome out of range");
    return guarded(err, [&] {
        rt::setDevice(ctx->impl->device());
        LiftOutput lo;
        ctx->impl->liftover(src, tgt, flags, n, dStart, dEnd, dStrand, lo, offsetBase, nullptr, coalescenceLimit);
        halgpu_lift_result *r = static_cast<halgpu_lift_result *>(std::calloc(1, sizeof(halgpu_lift_result)));
        r->n = n; r->n_rec = lo.nRec; r->offsets = lo.offsets; r->recs = lo.recs; r->on_device = 1;
        r->kernel_ms = lo.kernelMs; r->launches = lo.launches; r->n_retry = lo.nRetry; r->psl = lo.psl;
        r->fast_ms = lo.fastMs; r->n_complex = lo.nComplex; r->owner = ctx;
        *out = r;
    });
}

int halgpu_liftover(halgpu_ctx *ctx, int src, int tgt, int coalescenceLimit, uint32_t flags, size_t n,
                    const int64_t *start, const int64_t *endIncl, const uint8_t *strand, halgpu_lift_result **out,
                    char **err) {
    if (ctx == nullptr || out == nullptr || (n > 0 && (start == nullptr || endIncl == nullptr))) {
        return fail(err, "halgpu_liftover: null argument");
    }
    *out = nullptr;
    return guarded(err, [&] {
        rt::setDevice(ctx->impl->device());
        rt::Stream s = ctx->impl->stream();
        (void)s;
        if (genome(ctx, src) == nullptr || genome(ctx, tgt) == nullptr) throw HalError("genome index out of range");
        // (interval bounds are validated by the kernel itself: an out-of-range interval makes the call fail)
        // Pipeline: the batch is cut into chunks; while chunk k is being lifted on the engine's stream, chunk k+1's inputs
        // travel host->device and chunk k-1's records travel device->host on the copy stream.
        rt::Stream cs = ctx->impl->copyStream(), cb = ctx->impl->copyBackStream();
        size_t nChunks = n >= (4u << 20) ? 8 : (n >= (1u << 20) ? 4 : 1);
        if (const char *forced = std::getenv("HALGPU_HOST_CHUNKS")) { // test hook
            const long f = std::atol(forced);
            if (f >= 1 && (size_t)f <= std::max<size_t>(n, 1)) nChunks = (size_t)f;
        }
        std::vector<size_t> lo(nChunks + 1);
        for (size_t k = 0; k <= nChunks; ++k) lo[k] = n * k / nChunks;
        void *dS = nullptr, *dE = nullptr, *dT = nullptr;
        std::vector<rt::Event> inReady(nChunks);
        auto upload = [&](size_t k) {
            const size_t a = lo[k], c = lo[k + 1] - lo[k];
            rt::h2d((int64_t *)dS + a, start + a, c * 8, cs);
            rt::h2d((int64_t *)dE + a, endIncl + a, c * 8, cs);
            if (strand) rt::h2d((uint8_t *)dT + a, strand + a, c, cs);
            inReady[k].record(cs);
        };
        halgpu_lift_result *r = static_cast<halgpu_lift_result *>(std::calloc(1, sizeof(halgpu_lift_result)));
        if (r == nullptr) throw HalError("out of memory");
        std::vector<halgpu_lift_result *> parts;
        size_t recCap = n + n / 4 + 4096, nRec = 0;
        const bool wantPsl = (flags & HALGPU_PSL) != 0;
        try {
            DeviceCache &cache = ctx->impl->cache();
            dS = cache.take(n * 8); dE = cache.take(n * 8); dT = strand ? cache.take(n) : nullptr;
            r->n = n;
            r->offsets = static_cast<uint64_t *>(rt::hostAlloc((n + 1) * sizeof(uint64_t)));
            r->recs = static_cast<halgpu_lift_rec *>(rt::hostAlloc(recCap * sizeof(halgpu_lift_rec)));
            if (wantPsl) r->psl = static_cast<uint32_t *>(rt::hostAlloc(recCap * 16));
            upload(0);
            for (size_t k = 0; k < nChunks; ++k) {
                if (k + 1 < nChunks) upload(k + 1);
                inReady[k].wait(s); // the engine's stream waits for this chunk's inputs only
                const size_t a = lo[k], c = lo[k + 1] - lo[k];
                halgpu_lift_result *dev = nullptr;
                char *e2 = nullptr;
                const int rc = liftDeviceWithBase(ctx, src, tgt, coalescenceLimit, flags, c, (const int64_t *)dS + a, (const int64_t *)dE + a,
                                                  strand ? (const uint8_t *)dT + a : nullptr, nRec, &dev, &e2);
                if (rc != 0) {
                    std::string m = e2 ? e2 : "liftover failed";
                    std::free(e2);
                    throw HalError(m);
                }
                parts.push_back(dev);
                if (nRec + dev->n_rec > recCap) { // grow the pinned result (rare: more than 1.25 lines per interval)
                    rt::sync(cb);
                    const size_t newCap = std::max(recCap * 2, nRec + dev->n_rec + (n - lo[k + 1]) * 2);
                    halgpu_lift_rec *nr = static_cast<halgpu_lift_rec *>(rt::hostAlloc(newCap * sizeof(halgpu_lift_rec)));
                    std::memcpy(nr, r->recs, nRec * sizeof(halgpu_lift_rec));
                    rt::hostFree(r->recs);
                    r->recs = nr;
                    if (wantPsl) {
                        uint32_t *np = static_cast<uint32_t *>(rt::hostAlloc(newCap * 16));
                        std::memcpy(np, r->psl, nRec * 16);
                        rt::hostFree(r->psl);
                        r->psl = np;
                    }
                    recCap = newCap;
                }
                // halgpu_liftover_device returns with the engine's stream idle, so the copy stream may read the result now
                rt::d2h(r->offsets + a, dev->offsets, (c + (k + 1 == nChunks ? 1 : 0)) * sizeof(uint64_t), cb);
                rt::d2h(r->recs + nRec, dev->recs, dev->n_rec * sizeof(halgpu_lift_rec), cb);
                if (wantPsl) rt::d2h(r->psl + 4 * nRec, dev->psl, dev->n_rec * 16, cb);
                nRec += dev->n_rec;
                r->kernel_ms += dev->kernel_ms; r->launches += dev->launches; r->n_retry += dev->n_retry;
                r->fast_ms += dev->fast_ms; r->n_complex += dev->n_complex;
            }
            rt::sync(cs);
            rt::sync(cb);
            r->n_rec = nRec; // (offsets already carry each chunk's base: added on the device)
        } catch (...) {
            try { rt::sync(cs); rt::sync(cb); } catch (...) {}
            for (halgpu_lift_result *d : parts) halgpu_free_result(d);
            ctx->impl->release(dS); ctx->impl->release(dE); ctx->impl->release(dT);
            halgpu_free_result(r);
            throw;
        }
        for (halgpu_lift_result *d : parts) halgpu_free_result(d);
        ctx->impl->release(dS); ctx->impl->release(dE); ctx->impl->release(dT);
        *out = r;
    });
}

int halgpu_columns_depth_device(halgpu_ctx *ctx, int ref, int64_t first, int64_t last, int64_t step, const int *targets,
                                size_t nt, uint32_t flags, int32_t *dOut, float *kernelMs, char **err) {
    if (ctx == nullptr || dOut == nullptr || (nt > 0 && targets == nullptr)) return fail(err, "halgpu_columns_depth: null argument");
    return guarded(err, [&] {
        rt::setDevice(ctx->impl->device());
        std::vector<int> t(targets, targets + nt);
        ctx->impl->depth(ref, first, last, step, t, flags, dOut, kernelMs);
    });
}

int halgpu_columns_depth(halgpu_ctx *ctx, int ref, int64_t first, int64_t last, int64_t step, const int *targets, size_t nt,
                         uint32_t flags, int32_t *out, float *kernelMs, char **err) {
    if (ctx == nullptr || out == nullptr) return fail(err, "halgpu_columns_depth: null argument");
    if (step < 1 || last < first) return fail(err, "halgpu_columns_depth: empty or malformed range");
    return guarded(err, [&] {
        rt::setDevice(ctx->impl->device());
        rt::Stream s = ctx->impl->stream();
        const size_t n = (size_t)((last - first) / step + 1);
        int32_t *d = static_cast<int32_t *>(rt::dmallocAsync(n * sizeof(int32_t), s));
        char *e2 = nullptr;
        const int rc = halgpu_columns_depth_device(ctx, ref, first, last, step, targets, nt, flags, d, kernelMs, &e2);
        if (rc == 0) {
            rt::d2h(out, d, n * sizeof(int32_t), s);
            rt::sync(s);
        }
        rt::dfreeAsync(d, s);
        if (rc != 0) {
            std::string m = e2 ? e2 : "depth failed";
            std::free(e2);
            throw HalError(m);
        }
    });
}

int halgpu_column_runs(halgpu_ctx *ctx, int ref, int64_t first, int64_t last, const int *targets, size_t nt, uint32_t flags,
                       halgpu_col_runs **out, char **err) {
    return halgpu_column_runs_in_sweep(ctx, ref, first, last, first, targets, nt, flags, out, err);
}

int halgpu_column_runs_in_sweep(halgpu_ctx *ctx, int ref, int64_t first, int64_t last, int64_t sweepFirst, const int *targets, size_t nt,
                                uint32_t flags, halgpu_col_runs **out, char **err) {
    if (ctx == nullptr || out == nullptr || (nt > 0 && targets == nullptr)) return fail(err, "halgpu_column_runs: null argument");
    *out = nullptr;
    static_assert(sizeof(halgpu_col_row) == sizeof(ColRowRec), "row record layout");
    return guarded(err, [&] {
        rt::setDevice(ctx->impl->device());
        std::vector<int> t(targets, targets + nt);
        halgpu_col_runs *r = static_cast<halgpu_col_runs *>(std::calloc(1, sizeof(halgpu_col_runs)));
        try {
            ctx->impl->columnRuns(ref, first, last, t, flags, *r, sweepFirst);
        } catch (...) {
            halgpu_free_col_runs(r);
            throw;
        }
        *out = r;
    });
}

int halgpu_wiggle_liftover(halgpu_ctx *ctx, int src, int tgt, uint32_t flags, size_t nRuns, const int64_t *first, const int64_t *last,
                           const int64_t *valOff, const double *vals, size_t nVals, size_t nPre, const int64_t *prePos,
                           const double *preVal, halgpu_wig_result **out, char **err) {
    if (ctx == nullptr || out == nullptr || (nRuns > 0 && (first == nullptr || last == nullptr || valOff == nullptr || vals == nullptr)) ||
        (nPre > 0 && (prePos == nullptr || preVal == nullptr))) {
        return fail(err, "halgpu_wiggle_liftover: null argument");
    }
    *out = nullptr;
    return guarded(err, [&] {
        rt::setDevice(ctx->impl->device());
        WigOutput wo;
        ctx->impl->wiggle(src, tgt, flags, nRuns, first, last, valOff, vals, nVals, nPre, prePos, preVal, wo);
        halgpu_wig_result *r = static_cast<halgpu_wig_result *>(std::calloc(1, sizeof(halgpu_wig_result)));
        r->n = wo.n; r->pos = wo.pos; r->val = wo.val; r->kernel_ms = wo.kernelMs; r->launches = wo.launches; r->n_retry = wo.nRetry;
        *out = r;
    });
}

int halgpu_maf_text(halgpu_ctx *ctx, size_t nRows, const halgpu_maf_row *rows, size_t nPieces, const halgpu_maf_piece *pieces, const char *prefix,
                    size_t prefixBytes, size_t outBytes, char *out, float *kernelMs, char **err) {
    if (ctx == nullptr || (nRows > 0 && rows == nullptr) || (nPieces > 0 && pieces == nullptr) || (prefixBytes > 0 && prefix == nullptr) ||
        (outBytes > 0 && out == nullptr)) {
        return fail(err, "halgpu_maf_text: null argument");
    }
    return guarded(err, [&] {
        rt::setDevice(ctx->impl->device());
        ctx->impl->mafText(nRows, rows, nPieces, pieces, prefix, prefixBytes, outBytes, out, kernelMs);
    });
}

void halgpu_free_wig_result(halgpu_wig_result *r) {
    if (r == nullptr) return;
    rt::hostFree(r->pos);
    rt::hostFree(r->val);
    std::free(r);
}

void halgpu_free_col_runs(halgpu_col_runs *r) {
    if (r == nullptr) return;
    rt::hostFree(r->run_col);
    rt::hostFree(r->row_offset);
    rt::hostFree(r->rows);
    rt::hostFree(r->run_class);
    std::free(r);
}

const uint8_t *halgpu_genome_dna(const halgpu_ctx *ctx, int g) {
    const GenomeInfo *i = genome(ctx, g);
    return i ? i->dna : nullptr;
}

int halgpu_genome_metadata(const halgpu_ctx *ctx, int g, size_t index, const char **key, const char **value) {
    const GenomeInfo *i = genome(ctx, g);
    if (i == nullptr || key == nullptr || value == nullptr || index >= i->metadata.size()) return 1;
    auto it = i->metadata.begin();
    std::advance(it, (long)index);
    *key = it->first.c_str();
    *value = it->second.c_str();
    return 0;
}
const void *halgpu_genome_top_segments(const halgpu_ctx *ctx, int g) {
    const GenomeInfo *i = genome(ctx, g);
    return i ? i->top : nullptr;
}
const void *halgpu_genome_bottom_segments(const halgpu_ctx *ctx, int g, size_t *stride) {
    const GenomeInfo *i = genome(ctx, g);
    if (i && stride) *stride = i->bottomStride;
    return i ? i->bottom : nullptr;
}

void halgpu_free_result(halgpu_lift_result *r) {
    if (r == nullptr) return;
    if (r->on_device) { // the buffers return to the owning context's cache: no cudaFree, no device synchronisation
        halgpu_ctx *owner = static_cast<halgpu_ctx *>(r->owner);
        if (owner != nullptr) {
            owner->impl->release(r->offsets);
            owner->impl->release(r->recs);
            owner->impl->release(r->psl);
        }
    } else {
        rt::hostFree(r->offsets);
        rt::hostFree(r->recs);
        rt::hostFree(r->psl);
    }
    std::free(r);
}

void halgpu_free_string(char *s) { std::free(s); }

void *halgpu_host_alloc(size_t bytes) {
    try {
        return rt::hostAlloc(bytes);
    } catch (...) {
        return nullptr;
    }
}
void halgpu_host_free(void *p) { rt::hostFree(p); }

uint64_t halgpu_launch_count(void) { return rt::g_launches; }

} // extern "C"
