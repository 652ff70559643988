// C ABI (include/halgpu.h) over the engine.  Error convention of blockViz/inc/halBlockViz.h: non-zero
// return + malloc'd message in *err.
#include "../../include/halgpu.h"
#include "engine.hpp"
#include <cstdlib>
#include <cstring>

using namespace halgpu;

struct halgpu_ctx {
    std::unique_ptr<Context> impl;
    std::vector<std::vector<halgpu_seq>> seqTables;
};

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

int halgpu_liftover_device(halgpu_ctx *ctx, int src, int tgt, int coalescenceLimit, uint32_t flags, size_t n,
                           const int64_t *dStart, const int64_t *dEnd, const uint8_t *dStrand,
                           halgpu_lift_result **out, char **err) {
    if (ctx == nullptr || out == nullptr) return fail(err, "halgpu_liftover: null argument");
    *out = nullptr;
    if (coalescenceLimit != -1 && coalescenceLimit != halgpu_mrca(ctx, src, tgt)) {
        return fail(err, "halgpu_liftover: --coalescenceLimit other than the MRCA is not supported");
    }
    return guarded(err, [&] {
        rt::setDevice(ctx->impl->device());
        LiftOutput lo;
        ctx->impl->liftover(src, tgt, flags, n, dStart, dEnd, dStrand, lo);
        halgpu_lift_result *r = static_cast<halgpu_lift_result *>(std::calloc(1, sizeof(halgpu_lift_result)));
        r->n = n; r->n_rec = lo.nRec; r->offsets = lo.offsets; r->recs = lo.recs; r->on_device = 1;
        r->kernel_ms = lo.kernelMs; r->launches = lo.launches; r->n_retry = lo.nRetry; r->psl = lo.psl;
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
        // bounds check against the source genome (the CLI layer reports per-line errors; this is the last line of defence)
        const GenomeInfo *S = genome(ctx, src);
        if (S == nullptr || genome(ctx, tgt) == nullptr) throw HalError("genome index out of range");
        for (size_t i = 0; i < n; ++i) {
            if (start[i] < 0 || endIncl[i] < start[i] || endIncl[i] >= S->length) {
                throw HalError("interval " + std::to_string(i) + " is outside genome " + S->name);
            }
        }
        void *dS = rt::dmallocAsync(n * 8, s), *dE = rt::dmallocAsync(n * 8, s), *dT = strand ? rt::dmallocAsync(n, s) : nullptr;
        halgpu_lift_result *dev = nullptr;
        char *e2 = nullptr;
        rt::h2d(dS, start, n * 8, s);
        rt::h2d(dE, endIncl, n * 8, s);
        if (strand) rt::h2d(dT, strand, n, s);
        int rc = halgpu_liftover_device(ctx, src, tgt, coalescenceLimit, flags, n, (const int64_t *)dS, (const int64_t *)dE,
                                        (const uint8_t *)dT, &dev, &e2);
        rt::dfreeAsync(dS, s); rt::dfreeAsync(dE, s); rt::dfreeAsync(dT, s);
        if (rc != 0) {
            std::string m = e2 ? e2 : "liftover failed";
            std::free(e2);
            throw HalError(m);
        }
        halgpu_lift_result *r = static_cast<halgpu_lift_result *>(std::calloc(1, sizeof(halgpu_lift_result)));
        *r = *dev;
        r->on_device = 0;
        r->offsets = static_cast<uint64_t *>(rt::hostAlloc((n + 1) * sizeof(uint64_t)));
        r->recs = static_cast<halgpu_lift_rec *>(rt::hostAlloc(std::max<size_t>(dev->n_rec, 1) * sizeof(halgpu_lift_rec)));
        rt::d2h(r->offsets, dev->offsets, (n + 1) * sizeof(uint64_t), s);
        rt::d2h(r->recs, dev->recs, dev->n_rec * sizeof(halgpu_lift_rec), s);
        if (dev->psl != nullptr) {
            r->psl = static_cast<uint32_t *>(rt::hostAlloc(std::max<size_t>(dev->n_rec, 1) * 16));
            rt::d2h(r->psl, dev->psl, dev->n_rec * 16, s);
        }
        rt::sync(s);
        halgpu_free_result(dev);
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
    if (ctx == nullptr || out == nullptr || (nt > 0 && targets == nullptr)) return fail(err, "halgpu_column_runs: null argument");
    *out = nullptr;
    static_assert(sizeof(halgpu_col_row) == sizeof(ColRowRec), "row record layout");
    return guarded(err, [&] {
        rt::setDevice(ctx->impl->device());
        std::vector<int> t(targets, targets + nt);
        halgpu_col_runs *r = static_cast<halgpu_col_runs *>(std::calloc(1, sizeof(halgpu_col_runs)));
        try {
            ctx->impl->columnRuns(ref, first, last, t, flags, *r);
        } catch (...) {
            halgpu_free_col_runs(r);
            throw;
        }
        *out = r;
    });
}

void halgpu_free_col_runs(halgpu_col_runs *r) {
    if (r == nullptr) return;
    rt::hostFree(r->run_col);
    rt::hostFree(r->row_offset);
    rt::hostFree(r->rows);
    std::free(r);
}

const uint8_t *halgpu_genome_dna(const halgpu_ctx *ctx, int g) {
    const GenomeInfo *i = genome(ctx, g);
    return i ? i->dna : nullptr;
}

void halgpu_free_result(halgpu_lift_result *r) {
    if (r == nullptr) return;
    if (r->on_device) {
        rt::dfree(r->offsets);
        rt::dfree(r->recs);
        rt::dfree(r->psl);
    } else {
        rt::hostFree(r->offsets);
        rt::hostFree(r->recs);
        rt::hostFree(r->psl);
    }
    std::free(r);
}

void halgpu_free_string(char *s) { std::free(s); }

uint64_t halgpu_launch_count(void) { return rt::g_launches; }

} // extern "C"
