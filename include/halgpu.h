/*
 * halgpu.h -- C ABI of the B200-native HAL liftover / column-extraction hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.
 * Conventions follow the reference's in-tree C precedent, blockViz/inc/halBlockViz.h: functions
 * return 0 on success / non-zero on error with a malloc'd message in *err (caller frees it with
 * halgpu_free_string; err may be NULL), results are freed by the caller with halgpu_free_*.
 * A context is safe for one host thread at a time (the reference objects are single-thread only,
 * blockViz/impl/halBlockViz.cpp:29-38); it owns one CUDA stream.
 *
 * Each entry point names the reference interface it replaces.
 */
#ifndef HALGPU_H
#define HALGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct halgpu_ctx halgpu_ctx;

/* One chromosome / scaffold of a genome (replaces hal::Sequence getName/getStartPosition/
 * getSequenceLength, api/inc/halSequence.h; backed by MMapSequenceData, mmapSequenceData.h:21-30). */
typedef struct halgpu_seq {
    const char *name;  /* owned by the context */
    int64_t start;     /* first base in forward genome coordinates */
    int64_t length;
    int64_t num_top;   /* top / bottom segments of this sequence */
    int64_t num_bottom;
} halgpu_seq;

/* One lifted interval == one output BedLine of BlockLiftover::liftInterval
 * (liftover/impl/halBlockLiftover.cpp:82-105).  32 bytes, one DRAM sector. */
typedef struct halgpu_lift_rec {
    int64_t start;       /* target start, relative to its sequence */
    int64_t end;         /* exclusive */
    int64_t src_start;   /* BedLine::_srcStart: source start in forward GENOME coordinates */
    int32_t tgt_seq;     /* index into halgpu_sequence_table(tgtGenome) */
    uint8_t strand;      /* '+', '-' or '.' */
    uint8_t src_strand;  /* '+', '-' or '.' */
    uint16_t n_frag;     /* mapped fragments merged into this line (saturates at 65535) */
} halgpu_lift_rec;

/* One mapped fragment (HALGPU_RAW_FRAGMENTS): `length` source bases starting at src_start (forward GENOME coordinates of the
 * source genome) are homologous to the `length` target bases starting at tgt_start (forward GENOME coordinates of the target
 * genome, NOT sequence relative); with bit 1 of flags set the target piece is read on the reverse strand (its first base
 * pairs with the LAST source base), bit 0 likewise for the source piece ('-' input intervals).  32 bytes. */
typedef struct halgpu_frag {
    int64_t src_start, tgt_start, length;
    uint32_t flags; /* bit 0: source reversed, bit 1: target reversed */
    uint32_t pad;
} halgpu_frag;

/* Result of one liftover batch, CSR by input interval, lines of one interval in the reference's
 * output order (stable by src_start, liftover/impl/halLiftover.cpp:90). */
typedef struct halgpu_lift_result {
    size_t n;                 /* number of input intervals */
    size_t n_rec;             /* total output lines */
    uint64_t *offsets;        /* n+1 entries */
    halgpu_lift_rec *recs;    /* n_rec entries */
    int on_device;            /* 1: offsets/recs are device pointers (halgpu_liftover_device) */
    /* measurement: device time of the mapping kernel(s) of this batch and launch count */
    float kernel_ms;
    int launches;
    size_t n_retry;           /* intervals that needed the large-scratch re-launch */
    /* HALGPU_PSL only (else NULL): per output line the four PSL base counts of BlockLiftover::readPSLInfo
     * (liftover/impl/halBlockLiftover.cpp:115-162): matches, misMatches, repMatches, nCount -- 4 x uint32 per record */
    uint32_t *psl;
    /* measurement: device time of the one-lane-per-interval kernel alone (part of kernel_ms; 0 when the batch did not use
     * it) and the number of intervals it left to the one-warp-per-interval walk */
    float fast_ms;
    size_t n_complex;
    void *owner;              /* private: the context whose buffer cache the device buffers return to */
} halgpu_lift_result;

enum {
    HALGPU_NO_DUPES = 1u,     /* halLiftover --noDupes (liftover/impl/halLiftoverMain.cpp:24) */
    HALGPU_NO_SORT = 2u,      /* do not reorder the batch by source position inside the call */
    HALGPU_PSL = 4u,          /* also compare source and target DNA of every mapped fragment (halLiftover --outPSL) */
    HALGPU_RAW_FRAGMENTS = 16u, /* return the mapped fragments of every interval -- halMapSegment's output for each of its source
                                 * segments, before insertAndBreakOverlaps and extractSegment -- as halgpu_frag records in
                                 * result->recs (same size as halgpu_lift_rec; cast), in no particular order within an
                                 * interval.  For callers that refine and merge across intervals themselves (halSynteny lifts
                                 * whole chromosomes).  Not with HALGPU_PSL or HALGPU_COLUMN_LIFTOVER. */
    HALGPU_NO_FAST = 64u,       /* walk every interval piece by piece (one warp per interval) even where the whole interval is one
                                 * collinear run that the one-lane-per-interval kernel would map in one step per genome;
                                 * same results, for measurements and tests */
    HALGPU_SEED_BOTTOM = 32u,   /* take the source segments from the genome's BOTTOM array even if it has a top array
                                 * (BlockMapper::map does so when the source genome is the MRCA, halBlockMapper.cpp:76-83) */
    HALGPU_COLUMN_LIFTOVER = 8u /* hal::ColumnLiftover::liftInterval semantics (liftover/impl/halColumnLiftover.cpp:21-92)
                                 * instead of BlockLiftover's: per target sequence and strand the maximal runs of target
                                 * bases homologous to the interval, forward runs first; src_start = -1; with
                                 * HALGPU_NO_DUPES only canonical paralogs are followed upward (ColumnIterator noDupes) */
};

/* ---- open / stage (replaces openHalAlignment + MMapAlignment ctor, api/impl/halAlignmentInstance.cpp:133,
 *      api/mmap_impl/mmapAlignment.cpp:10-19,74-80): mmap the HAL-MMAP file, parse it and stage the
 *      segment index arrays and packed DNA into the HBM of `device` once. ---- */
int halgpu_open(const char *mmap_hal_path, int device, halgpu_ctx **out, char **err);
void halgpu_close(halgpu_ctx *ctx);

/* ---- read-only object model (replaces Alignment::getNumGenomes/getRootName/getParentName/getChildNames/
 *      getNewickTree, api/inc/halAlignment.h, and Genome::getSequence*/
int halgpu_num_genomes(const halgpu_ctx *ctx);
const char *halgpu_genome_name(const halgpu_ctx *ctx, int genome);
int halgpu_genome_id(const halgpu_ctx *ctx, const char *name); /* -1 if absent (openGenome == NULL) */
int halgpu_genome_parent(const halgpu_ctx *ctx, int genome);   /* -1 for the root */
int halgpu_genome_num_children(const halgpu_ctx *ctx, int genome);
int halgpu_genome_child(const halgpu_ctx *ctx, int genome, int slot);
int64_t halgpu_genome_length(const halgpu_ctx *ctx, int genome);
int64_t halgpu_genome_num_top(const halgpu_ctx *ctx, int genome);
int64_t halgpu_genome_num_bottom(const halgpu_ctx *ctx, int genome);
const char *halgpu_newick(const halgpu_ctx *ctx);
int halgpu_sequence_table(const halgpu_ctx *ctx, int genome, const halgpu_seq **out, size_t *n);
int halgpu_mrca(const halgpu_ctx *ctx, int genome_a, int genome_b); /* getLowestCommonAncestor, halCommon.cpp:123 */
size_t halgpu_staged_bytes(const halgpu_ctx *ctx);                  /* bytes resident in HBM */
void *halgpu_stream(const halgpu_ctx *ctx);                         /* the cudaStream_t all work runs on */

/* ---- liftover (replaces halMapSegment + BlockMapper::extractSegment as driven by
 *      BlockLiftover::liftInterval, liftover/impl/halBlockLiftover.cpp:46-113).
 *      Inputs are forward GENOME coordinates: src_start = bed.start + seq.start,
 *      src_end_incl = bed.end - 1 + seq.start (halBlockLiftover.cpp:48-49); strand may be NULL ('+').
 *      coalescence_limit: -1 (== the MRCA, the CLI default) or a genome index (halLiftover --coalescenceLimit): with an
 *      ancestor of the MRCA, paralogs that coalesce below it are mapped too (mapRecursiveParalogies,
 *      api/impl/halSegmentMapper.cpp:525-576; ignored with HALGPU_NO_DUPES like in the reference, :619); a genome that
 *      is not the MRCA or one of its ancestors fails with the reference's "Hit root genome ..." message.
 *      Host buffers in, host result out;
 *      the host<->device copies are part of the call.
 *      result->recs[i].n_frag counts the pieces THIS library merged into the line (a whole collinear run counts once); it is
 *      a diagnostic, the reference has no such field. ---- */
int halgpu_liftover(halgpu_ctx *ctx, int src_genome, int tgt_genome, int coalescence_limit, uint32_t flags,
                    size_t n, const int64_t *src_start, const int64_t *src_end_incl, const uint8_t *strand,
                    halgpu_lift_result **out, char **err);

/* Same, with the three input arrays already resident in device memory of the context's GPU and the
 * result left in device memory (result->on_device == 1).  The result's device buffers belong to the context: free the
 * result (halgpu_free_result) before halgpu_close. */
int halgpu_liftover_device(halgpu_ctx *ctx, int src_genome, int tgt_genome, int coalescence_limit, uint32_t flags,
                           size_t n, const int64_t *d_src_start, const int64_t *d_src_end_incl,
                           const uint8_t *d_strand, halgpu_lift_result **out, char **err);

/* ---- alignment depth (replaces the ColumnIterator sweep of halAlignmentDepth::printSequence,
 *      alignmentDepth/halAlignmentDepth.cpp:215-308, for one contiguous reference range).
 *      Positions first, first+step, ... <= last are forward GENOME coordinates of ref_genome; depth_out receives
 *      one int32 per position: (#distinct genomes with >= 1 aligned base) - 1, or (#aligned bases) - 1 with
 *      HALGPU_COUNT_DUPES (:262-280).  targets (n_targets == 0: all genomes) restricts the genomes counted and
 *      the traversal to their spanning tree (api/impl/halColumnIterator.cpp:47-51, 802-803). ---- */
enum {
    HALGPU_COUNT_DUPES = 1u,   /* halAlignmentDepth --countDupes */
    HALGPU_NO_ANCESTORS = 2u,  /* --noAncestors */
    HALGPU_COL_NO_DUPES = 4u   /* ColumnIterator noDupes */
};
int halgpu_columns_depth(halgpu_ctx *ctx, int ref_genome, int64_t first, int64_t last, int64_t step, const int *targets,
                         size_t n_targets, uint32_t flags, int32_t *depth_out, float *kernel_ms, char **err);
/* same with depth_out in device memory */
int halgpu_columns_depth_device(halgpu_ctx *ctx, int ref_genome, int64_t first, int64_t last, int64_t step,
                                const int *targets, size_t n_targets, uint32_t flags, int32_t *d_depth_out,
                                float *kernel_ms, char **err);

/* ---- column runs (replaces the ColumnIterator sweep consumed by MafExport::convertSequence,
 *      maf/impl/halMafExport.cpp:46-81: seq->getColumnIterator(...) + toRight() per base, default flags
 *      unique=false, maxRefGap=0).  The reference columns first..last (forward GENOME coordinates, one
 *      reference sequence) are returned as maximal runs of consecutive columns whose rows are the same
 *      sequences and strands advancing collinearly: run r covers columns run_col[r] .. run_col[r+1]-1
 *      (relative to `first`); its FIRST column's rows are rows[row_offset[r] .. row_offset[r+1]) in ColumnMap
 *      order (genome name, sequence index, discovery order -- api/inc/halColumnIterator.h:45-54); in column
 *      run_col[r]+j every row sits at pos + j (forward strand) or pos - j (rev). ---- */
typedef struct halgpu_col_row {
    int64_t pos;     /* forward genome coordinate in its genome */
    int32_t seq;     /* index into halgpu_sequence_table(genome) */
    int16_t genome;
    uint8_t rev;     /* DnaIterator::getReversed() */
    uint8_t pad;
} halgpu_col_row;

typedef struct halgpu_col_runs {
    size_t n_cols, n_runs, n_rows;
    int64_t *run_col;       /* n_runs + 1 entries, run_col[n_runs] == n_cols */
    uint64_t *row_offset;   /* n_runs + 1 entries */
    halgpu_col_row *rows;   /* n_rows entries */
    float kernel_ms;        /* device time of the column-walk kernels */
    uint8_t *run_class;     /* HALGPU_COL_UNIQUE only (else NULL), n_runs entries: 0 = columns MafExport writes; 1 = walked by
                             * the iterator but not written (left-most reference base left of the sweep start: isCanonicalOnRef
                             * false) -- their sequences still become ColumnMap keys; 2 = skipped without being walked (the
                             * position was visited through an earlier column of the sweep) */
} halgpu_col_runs;

enum {
    HALGPU_ONLY_ORTHOLOGS = 8u, /* hal2maf --onlyOrthologs; also HALGPU_NO_ANCESTORS, HALGPU_COL_NO_DUPES */
    HALGPU_COL_UNIQUE = 16u     /* hal2maf --unique: ColumnIterator(unique = true) + MafExport's isCanonicalOnRef test
                                 * (maf/impl/halMafExport.cpp:52,61; api/impl/halColumnIterator.cpp:208-212,749-762) */
};

int halgpu_column_runs(halgpu_ctx *ctx, int ref_genome, int64_t first, int64_t last, const int *targets, size_t n_targets,
                       uint32_t flags, halgpu_col_runs **out, char **err);
/* same for a range that is one chunk of a longer sweep which started at sweep_first <= first (only matters with
 * HALGPU_COL_UNIQUE, whose visit cache spans the whole sweep of one convertSequence call) */
int halgpu_column_runs_in_sweep(halgpu_ctx *ctx, int ref_genome, int64_t first, int64_t last, int64_t sweep_first, const int *targets,
                                size_t n_targets, uint32_t flags, halgpu_col_runs **out, char **err);
void halgpu_free_col_runs(halgpu_col_runs *runs);
/* packed DNA of a genome as staged (host pointer into the mapped file, 2 bases per byte, even index = high nibble;
 * replaces Genome::getDnaIterator for bulk text emission, api/inc/halDnaIterator.h:131-138) */
const uint8_t *halgpu_genome_dna(const halgpu_ctx *ctx, int genome);
/* raw segment arrays of a genome (host pointers into the mapped file; num + 1 records, the last one a sentinel whose
 * start is the genome length): top records are 40 B (api/mmap_impl/mmapTopSegmentData.h:40-44), bottom records
 * *stride B (mmapBottomSegmentData.h:35-52); the first int64 of every record is its start position.  Replaces
 * Genome::getTopSegmentIterator / getBottomSegmentIterator for host code that only needs segment boundaries. */
/* metadata of a genome (replaces Genome::getMetaData()->getMap(), api/inc/halMetaData.h): entry `index` in key order;
 * returns 0 and sets *key / *value (owned by the context) while index < the number of entries, else 1 */
int halgpu_genome_metadata(const halgpu_ctx *ctx, int genome, size_t index, const char **key, const char **value);
const void *halgpu_genome_top_segments(const halgpu_ctx *ctx, int genome);
const void *halgpu_genome_bottom_segments(const halgpu_ctx *ctx, int genome, size_t *stride);

/* ---- MAF row text (replaces the per-base appendColumn / DnaIterator::getBase loop and the row printer of MafBlock,
 *      maf/impl/halMafBlock.cpp:370-395,452-456,499-519; api/inc/halDnaIterator.h:131-138).  The caller (the block state
 *      machine of csrc/host/maf_export.cpp) describes every row of the finished blocks: where its bytes go in the output,
 *      its formatted prefix, and its text as pieces -- runs of gaps or of consecutive bases of one genome, read forward or
 *      reverse-complemented from the staged packed DNA.  The whole text is produced on the device and copied to `out`
 *      (host memory; page-locked memory from halgpu_host_alloc makes the copy run at full PCIe rate). ---- */
typedef struct halgpu_maf_row {
    uint64_t out_offset;     /* first byte of the row in the output */
    uint32_t prefix_offset;  /* the row's prefix inside `prefix` ("a\n" of a block's first row included) */
    uint32_t prefix_len;
    uint32_t first_piece, num_pieces;
    int32_t genome;
    uint32_t tail_newlines;  /* '\n' characters after the text: 1, or 2 for the last row of a block that is followed by a blank line */
} halgpu_maf_row;
typedef struct halgpu_maf_piece {
    int64_t pos;         /* forward genome coordinate of the first base (walks DOWN for reverse-complemented pieces) */
    int64_t count_kind;  /* (characters << 2) | kind: 0 gap run, 1 bases forward, 2 bases reverse-complemented */
} halgpu_maf_piece;
int halgpu_maf_text(halgpu_ctx *ctx, size_t n_rows, const halgpu_maf_row *rows, size_t n_pieces, const halgpu_maf_piece *pieces,
                    const char *prefix, size_t prefix_bytes, size_t out_bytes, char *out, float *kernel_ms, char **err);

/* ---- wiggle liftover (replaces the mapping core of hal::WiggleLiftover -- mapSegment / mapFragments and the
 *      WiggleTiles<double> accumulator, liftover/impl/halWiggleLiftover.cpp:98-158, liftover/inc/halWiggleTiles.h; the
 *      --append preload of WiggleLoader::visitLine, halWiggleLoader.cpp:37-48).
 *      Input: n_runs source ranges [run_first, run_last_incl] in forward GENOME coordinates of src_genome.  Run i carries
 *      one value per base, vals[val_offset[i] + (p - run_first[i])] for its base p, when val_offset[i] >= 0; or the single
 *      value vals[~val_offset[i]] for all its bases when val_offset[i] < 0 (a wiggle line with a span).  Every base of
 *      the target genome that some source base maps to (halMapSegment; HALGPU_NO_DUPES as in halgpu_liftover) receives
 *      max(value, what it holds), a base nothing was written to holding 0.0 for that comparison; the n_preload
 *      (preload_pos, preload_val) pairs (distinct target GENOME positions) are stored first as they are.
 *      Output: the target bases that hold a value, ascending, with their values (host memory, halgpu_free_wig_result).
 *      No arithmetic is done on the values, so they are bit-exact copies of inputs (or +0.0). ---- */
typedef struct halgpu_wig_result {
    size_t n;
    int64_t *pos;   /* forward genome coordinates of tgt_genome, ascending */
    double *val;
    float kernel_ms;
    int launches;
    size_t n_retry;
} halgpu_wig_result;
int halgpu_wiggle_liftover(halgpu_ctx *ctx, int src_genome, int tgt_genome, uint32_t flags, size_t n_runs,
                           const int64_t *run_first, const int64_t *run_last_incl, const int64_t *val_offset,
                           const double *vals, size_t n_vals, size_t n_preload, const int64_t *preload_pos,
                           const double *preload_val, halgpu_wig_result **out, char **err);
void halgpu_free_wig_result(halgpu_wig_result *res);

void halgpu_free_result(halgpu_lift_result *res);
void halgpu_free_string(char *s);

/* page-locked host staging buffers for the input arrays of halgpu_liftover (the copies then run at full PCIe rate and
 * overlap the kernel); returns NULL on failure.  Buffers are recycled by the library; free with halgpu_host_free.
 * (The reference has no counterpart: its BedScanner hands one line at a time to liftInterval, halBedScanner.cpp:40-61.) */
void *halgpu_host_alloc(size_t bytes);
void halgpu_host_free(void *p);

/* ---- multi-GPU liftover (the reference's only parallel driver partitions work the same way across
 *      processes: maf/hal2mafMP.py:63-79).  One process -- or host thread -- per GPU, each with its own context on the same
 *      HAL file (the staged index is replicated).  Every rank lifts its own shard of the interval batch; ONE all-gather of
 *      the output interval buffer over NVLink then leaves the result of the WHOLE batch, shards in rank order, in the
 *      device memory of every rank: every rank copies the other ranks' shards out of their buffers (peer memory: CUDA IPC
 *      between processes, peer access between threads of one process); NCCL carries the per-batch 128-byte headers and is
 *      the fallback gather where peer access is not available.  The communicator is bootstrapped like NCCL's: rank 0
 *      creates a 128-byte id (halgpu_comm_unique_id) and hands it to the other ranks by any out-of-band means (a file, MPI,
 *      torch.distributed); all ranks then call halgpu_comm_init.  NCCL is resolved at run time (libnccl.so.2); without it
 *      these calls fail and everything else works.
 *      begin() returns once this rank's shard is lifted and the transfers are enqueued on the communicator's own streams;
 *      end() waits for them and hands out the gathered result (device memory of the context; halgpu_free_result).  A
 *      caller that begins batch k+1 before ending batch k overlaps the gather of k with the lift of k+1.  begin / end /
 *      halgpu_comm_free are collective: every rank must issue them in the same order. ---- */
typedef struct halgpu_comm halgpu_comm;
typedef struct halgpu_gather halgpu_gather;
int halgpu_comm_unique_id(uint8_t id[128], char **err);
int halgpu_comm_init(halgpu_ctx *ctx, int nranks, int rank, const uint8_t id[128], halgpu_comm **out, char **err);
void halgpu_comm_free(halgpu_comm *comm); /* before halgpu_close of its context */
int halgpu_comm_rank(const halgpu_comm *comm);
int halgpu_comm_size(const halgpu_comm *comm);
int halgpu_liftover_allgather_begin(halgpu_comm *comm, int src_genome, int tgt_genome, int coalescence_limit, uint32_t flags,
                                    size_t n, const int64_t *d_src_start, const int64_t *d_src_end_incl, const uint8_t *d_strand,
                                    halgpu_gather **out, char **err);
/* n_per_rank / n_rec_per_rank (optional, halgpu_comm_size entries): intervals / records each rank contributed; the
 * gather handle is consumed whether or not the call succeeds */
int halgpu_liftover_allgather_end(halgpu_gather *gather, halgpu_lift_result **out, size_t *n_per_rank, size_t *n_rec_per_rank,
                                  char **err);

/* number of CUDA kernels this library has launched in this process (bench.py "gpu_launches") */
uint64_t halgpu_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
