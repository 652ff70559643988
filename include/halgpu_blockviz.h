/*
 * halgpu_blockviz.h -- the reference's block-visualisation C API (blockViz/inc/halBlockViz.h, the interface the UCSC browser
 * links against) as exported by hal_b200/libhalBlockVizGpu.so on top of the B200 context of include/halgpu.h.
 * Same names, struct layouts, argument meaning and error convention (NULL / -1 plus a malloc'd message in *errStr; with
 * errStr == NULL the reference throws, this library aborts with the message).  The ABI is identical, so client code built
 * against the reference's own halBlockViz.h links against this library unchanged; this header exists for builds without the
 * reference tree.
 *
 * Implemented on the GPU path: halGetBlocksInTargetRange[_filterByChrom] (all three duplication modes, sequence modes,
 * coalescence limit, reversed target range, mapBackAdjacencies) and halGetMaf / halGetMAF with maxRefGap == 0.  Host-only queries: halOpen (a HAL-MMAP file), halClose, halCloseGenome, halGetSpecies,
 * halGetPossibleCoalescenceLimits, halGetChroms, halGetDna, halGetMaxLODQueryLength.  Not implemented (return the failure
 * value with a message): maxRefGap > 0 (gapped column iterators).
 */
#ifndef HAL_BLOCK_VIZ_H
#define HAL_BLOCK_VIZ_H
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef long hal_int_t;

struct hal_target_range_t { /* one reference range of a target-duplication list */
    struct hal_target_range_t *next;
    hal_int_t tStart;
    hal_int_t size;
};

struct hal_target_dupe_list_t { /* homologous ranges of the reference that share query sequence */
    struct hal_target_dupe_list_t *next;
    hal_int_t id;
    struct hal_target_range_t *tRange;
    char *qChrom;
};

struct hal_block_results_t {
    struct hal_block_t *mappedBlocks;
    struct hal_target_dupe_list_t *targetDupeBlocks;
};

struct hal_block_t { /* all coordinates forward-strand, sequence relative */
    struct hal_block_t *next;
    char *qChrom;
    hal_int_t tStart;
    hal_int_t qStart;
    hal_int_t size;
    char strand;
    char *qSequence; /* query DNA if requested */
    char *tSequence; /* target DNA if requested */
};

struct hal_species_t {
    struct hal_species_t *next;
    char *name;
    hal_int_t length;
    hal_int_t numChroms;
    char *parentName;
    double parentBranchLength;
};

struct hal_chromosome_t {
    struct hal_chromosome_t *next;
    char *name;
    hal_int_t length;
};

struct hal_metadata_t {
    struct hal_metadata_t *next;
    char *key;
    char *value;
};

typedef enum { HAL_NO_DUPS = 0, HAL_QUERY_DUPS, HAL_QUERY_AND_TARGET_DUPS } hal_dup_type_t;
typedef enum { HAL_NO_SEQUENCE = 0, HAL_LOD0_SEQUENCE, HAL_FORCE_LOD0_SEQUENCE } hal_seqmode_type_t;

int halOpenHalOrLod(char *lodFilePath, char **errStr);
int halOpenLOD(char *lodFilePath, char **errStr);
int halOpen(char *halFilePath, char **errStr);
int halClose(int halHandle, char **errStr);
int halCloseGenome(int halHandle, const char *genomeName, char **errStr);

void halFreeBlockResults(struct hal_block_results_t *results);
void halFreeBlocks(struct hal_block_t *block);
void halFreeTargetDupeLists(struct hal_target_dupe_list_t *dupes);
void halFreeSpeciesList(struct hal_species_t *species);
void halFreeChromList(struct hal_chromosome_t *chromosome);
void halFreeMetadataList(struct hal_metadata_t *metadata);

struct hal_species_t *halGetPossibleCoalescenceLimits(int halHandle, const char *qSpecies, const char *tSpecies, char **errStr);

struct hal_block_results_t *halGetBlocksInTargetRange(int halHandle, char *qSpecies, char *tSpecies, char *tChrom, hal_int_t tStart,
                                                      hal_int_t tEnd, hal_int_t tReversed, hal_seqmode_type_t seqMode,
                                                      hal_dup_type_t dupMode, int mapBackAdjacencies, const char *coalescenceLimitName,
                                                      char **errStr);
struct hal_block_results_t *halGetBlocksInTargetRange_filterByChrom(int halHandle, char *qSpecies, char *tSpecies, char *tChrom,
                                                                    hal_int_t tStart, hal_int_t tEnd, hal_int_t tReversed,
                                                                    hal_seqmode_type_t seqMode, hal_dup_type_t dupMode,
                                                                    int mapBackAdjacencies, char *qChrom,
                                                                    const char *coalescenceLimitName, char **errStr);

hal_int_t halGetMaf(FILE *outFile, int halHandle, struct hal_species_t *qSpeciesNames, char *tSpecies, char *tChrom, hal_int_t tStart,
                    hal_int_t tEnd, int maxRefGap, int maxBlockLength, int doDupes, char **errStr);
hal_int_t halGetMAF(FILE *outFile, int halHandle, struct hal_species_t *qSpeciesNames, char *tSpecies, char *tChrom, hal_int_t tStart,
                    hal_int_t tEnd, int doDupes, char **errStr);

struct hal_species_t *halGetSpecies(int halHandle, char **errStr);
struct hal_chromosome_t *halGetChroms(int halHandle, char *speciesName, char **errStr);
char *halGetDna(int halHandle, char *speciesName, char *chromName, hal_int_t start, hal_int_t end, char **errStr);
hal_int_t halGetMaxLODQueryLength(int halHandle, char **errStr);
struct hal_metadata_t *halGetGenomeMetadata(int halHandle, const char *genomeName, char **errStr);

#ifdef __cplusplus
}
#endif
#endif
