// TEST HARNESS ONLY -- a stand-in for libhalgpu.so that implements just the entry points the halLiftover CLI's text
// layer touches, with a synthetic, deterministic "mapping" (no HAL file, no GPU).  It lets the CPU test tier push
// millions of BED lines through hal_b200/csrc/host/{gpu_liftover,bed_fast}.cpp to check that the multi-threaded text
// path and the serial BedLine path print the same bytes, and to time the text layer alone.  Never shipped or loaded by
// the product.
//
// Genomes: "S" (source) and "T" (target), three sequences each.  Interval [s, e] of S maps to k = s % 3 lines:
// line j covers T positions [s + 7j, e + 7j], strand '+' for even j and '-' for odd j, src_start = s + j.
#include "../../include/halgpu.h"
#include <cstdlib>
#include <cstring>
#include <string>

struct halgpu_ctx { int dummy; };
static const halgpu_seq kSrc[3] = {{"chrA", 0, 400000000, 0, 0}, {"chrB", 400000000, 300000000, 0, 0}, {"scaffold_17 x", 700000000, 1000, 0, 0}};
static const halgpu_seq kTgt[3] = {{"tA", 0, 500000000, 0, 0}, {"tB", 500000000, 400000000, 0, 0}, {"tC", 900000000, 100000000, 0, 0}};

extern "C" {
int halgpu_open(const char *, int, halgpu_ctx **out, char **) { *out = new halgpu_ctx; return 0; }
void halgpu_close(halgpu_ctx *c) { delete c; }
int halgpu_genome_id(const halgpu_ctx *, const char *name) { return !std::strcmp(name, "S") ? 0 : (!std::strcmp(name, "T") ? 1 : -1); }
const char *halgpu_genome_name(const halgpu_ctx *, int g) { return g == 0 ? "S" : "T"; }
int halgpu_sequence_table(const halgpu_ctx *, int g, const halgpu_seq **out, size_t *n) { *out = g == 0 ? kSrc : kTgt; *n = 3; return 0; }
void *halgpu_host_alloc(size_t n) { return std::malloc(n ? n : 1); }
void halgpu_host_free(void *p) { std::free(p); }
void halgpu_free_string(char *s) { std::free(s); }
void halgpu_free_result(halgpu_lift_result *r) {
    if (r) { std::free(r->offsets); std::free(r->recs); std::free(r); }
}
int halgpu_liftover(halgpu_ctx *, int, int, int, uint32_t, size_t n, const int64_t *s, const int64_t *e, const uint8_t *st,
                    halgpu_lift_result **out, char **) {
    halgpu_lift_result *r = static_cast<halgpu_lift_result *>(std::calloc(1, sizeof *r));
    r->n = n;
    r->offsets = static_cast<uint64_t *>(std::malloc((n + 1) * 8));
    uint64_t tot = 0;
    for (size_t i = 0; i < n; ++i) { r->offsets[i] = tot; tot += (uint64_t)(s[i] % 3); }
    r->offsets[n] = tot;
    r->n_rec = tot;
    r->recs = static_cast<halgpu_lift_rec *>(std::malloc((tot + 1) * sizeof(halgpu_lift_rec)));
    for (size_t i = 0; i < n; ++i) {
        for (int j = 0; j < (int)(s[i] % 3); ++j) {
            halgpu_lift_rec &q = r->recs[r->offsets[i] + j];
            const int64_t a = s[i] + 7 * j, b = e[i] + 7 * j;
            q.tgt_seq = a >= kTgt[2].start ? 2 : (a >= kTgt[1].start ? 1 : 0);
            q.start = a - kTgt[q.tgt_seq].start;
            q.end = b + 1 - kTgt[q.tgt_seq].start;
            q.src_start = s[i] + j;
            const bool rev = (j & 1) != 0;
            const uint8_t in = st ? st[i] : '+';
            q.strand = in == '.' ? '.' : ((rev != (in == '-')) ? '-' : '+');
            q.src_strand = in;
            q.n_frag = 1;
        }
    }
    *out = r;
    return 0;
}
}
