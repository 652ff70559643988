// TEST HARNESS -- prints everything the blockViz C API returns for one query, as text, so that the reference's
// libHalBlockViz (compiled into oracle/_ref/blockVizCli from /root/reference) and this repo's GPU implementation of the same
// API (hal_b200/libhalBlockVizGpu.so) can be compared byte for byte.  It only uses the public C API; it is compiled
// against the reference's halBlockViz.h (oracle build) or include/halgpu_blockviz.h (-DHALGPU_BLOCKVIZ_HEADER).
//
// usage: blockVizCli <hal> species | meta <genome> | chroms <genome> | dna <genome> <chrom> <start> <end> | limits <q> <t> | maxlod
//        blockVizCli <hal> maf <tSpecies> <tChrom> <tStart> <tEnd> <maxRefGap> <maxBlockLength> <doDupes> <q1,q2,...>   (MAF to stdout)
//        blockVizCli <hal> blocks <qSpecies> <tSpecies> <tChrom> <tStart> <tEnd> <tReversed> <seqMode> <dupMode> <adj> <limit|-> [qChromFilter]
#ifdef HALGPU_BLOCKVIZ_HEADER
#include "halgpu_blockviz.h" // this repo's declaration of the same API
#else
#include "halBlockViz.h" // the reference's header (oracle build)
#endif
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: see source\n"); return 2; }
    char *err = NULL;
    int h = halOpenHalOrLod(argv[1], &err); // a HAL file or a level-of-detail list
    if (h < 0) { printf("ERROR %s\n", err ? err : "?"); return 1; }
    const std::string cmd = argv[2];
    int rc = 0;
    if (cmd == "species") {
        hal_species_t *s = halGetSpecies(h, &err);
        if (!s) { printf("ERROR %s\n", err ? err : "?"); rc = 1; }
        for (hal_species_t *p = s; p; p = p->next) printf("%s\t%ld\t%ld\t%s\t%g\n", p->name, p->length, p->numChroms, p->parentName, p->parentBranchLength);
        halFreeSpeciesList(s);
    } else if (cmd == "meta") {
        hal_metadata_t *m = halGetGenomeMetadata(h, argv[3], &err);
        if (!m && err) { printf("ERROR %s\n", err); rc = 1; }
        for (hal_metadata_t *p = m; p; p = p->next) printf("%s=%s\n", p->key, p->value);
        halFreeMetadataList(m);
    } else if (cmd == "chroms") {
        hal_chromosome_t *c = halGetChroms(h, argv[3], &err);
        if (!c) { printf("ERROR %s\n", err ? err : "?"); rc = 1; }
        for (hal_chromosome_t *p = c; p; p = p->next) printf("%s\t%ld\n", p->name, p->length);
        halFreeChromList(c);
    } else if (cmd == "dna") {
        char *d = halGetDna(h, argv[3], argv[4], atol(argv[5]), atol(argv[6]), &err);
        if (!d) { printf("ERROR %s\n", err ? err : "?"); rc = 1; } else { printf("%s\n", d); free(d); }
    } else if (cmd == "limits") {
        hal_species_t *s = halGetPossibleCoalescenceLimits(h, argv[3], argv[4], &err);
        if (!s) printf("NONE %s\n", err ? err : "");
        for (hal_species_t *p = s; p; p = p->next) printf("%s\t%ld\t%ld\t%s\t%g\n", p->name, p->length, p->numChroms, p->parentName, p->parentBranchLength);
        halFreeSpeciesList(s);
    } else if (cmd == "maxlod") {
        printf("%ld\n", halGetMaxLODQueryLength(h, &err));
    } else if (cmd == "maf" && argc >= 11) {
        hal_species_t *head = NULL, *tail = NULL;
        std::string names = argv[10];
        size_t at = 0;
        while (at <= names.size()) {
            size_t c = names.find(',', at);
            if (c == std::string::npos) c = names.size();
            hal_species_t *s = (hal_species_t *)calloc(1, sizeof(hal_species_t));
            s->name = strdup(names.substr(at, c - at).c_str());
            if (!head) head = s; else tail->next = s;
            tail = s;
            at = c + 1;
        }
        hal_int_t n = halGetMaf(stdout, h, head, argv[3], argv[4], atol(argv[5]), atol(argv[6]), atoi(argv[7]), atoi(argv[8]), atoi(argv[9]), &err);
        fflush(stdout);
        if (n < 0) { printf("ERROR %s\n", err ? err : "?"); rc = 1; } else printf("RETURNED %ld\n", n);
        halFreeSpeciesList(head);
    } else if (cmd == "blocks" && argc >= 13) {
        const char *limit = strcmp(argv[12], "-") ? argv[12] : NULL;
        hal_block_results_t *r;
        if (argc > 13) {
            r = halGetBlocksInTargetRange_filterByChrom(h, argv[3], argv[4], argv[5], atol(argv[6]), atol(argv[7]), atol(argv[8]),
                                                        (hal_seqmode_type_t)atoi(argv[9]), (hal_dup_type_t)atoi(argv[10]), atoi(argv[11]), argv[13], limit, &err);
        } else {
            r = halGetBlocksInTargetRange(h, argv[3], argv[4], argv[5], atol(argv[6]), atol(argv[7]), atol(argv[8]), (hal_seqmode_type_t)atoi(argv[9]),
                                          (hal_dup_type_t)atoi(argv[10]), atoi(argv[11]), limit, &err);
        }
        if (!r) { printf("ERROR %s\n", err ? err : "?"); rc = 1; }
        else {
            for (hal_block_t *b = r->mappedBlocks; b; b = b->next)
                printf("B\t%s\t%ld\t%ld\t%ld\t%c\t%s\t%s\n", b->qChrom, b->tStart, b->qStart, b->size, b->strand, b->qSequence ? b->qSequence : "-", b->tSequence ? b->tSequence : "-");
            for (hal_target_dupe_list_t *d = r->targetDupeBlocks; d; d = d->next) {
                printf("D\t%ld\t%s", d->id, d->qChrom);
                for (hal_target_range_t *t = d->tRange; t; t = t->next) printf("\t%ld:%ld", t->tStart, t->size);
                printf("\n");
            }
            halFreeBlockResults(r);
        }
    } else { fprintf(stderr, "bad command\n"); rc = 2; }
    halClose(h, &err);
    return rc;
}
