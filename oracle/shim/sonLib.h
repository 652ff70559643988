/* TEST INFRASTRUCTURE ONLY (oracle build shim) -- not part of the product path.
 *
 * Minimal stand-in for the subset of sonLib's stTree / stString API that the
 * reference's hot-path sources call (sonLib itself is an un-vendored sibling
 * checkout of the reference, see /root/reference/include.mk:24-25).  Only what
 * the mmap back end, ColumnIterator and MafBlock need: newick parse/print with
 * "%g" branch lengths, label/parent/child accessors and client data.
 * Call sites: api/mmap_impl/mmapAlignment.h:43-230, api/impl/halColumnIterator.cpp:404-555,
 * maf/impl/halMafBlock.cpp:137-199,484-496.
 */
#ifndef ORACLE_SHIM_SONLIB_H
#define ORACLE_SHIM_SONLIB_H
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct stTree {
    double branchLength;
    std::vector<stTree *> kids;
    char *label;
    void *clientData;
    stTree *parent;
};

static inline char *stString_copy(const char *s) {
    if (s == NULL) return NULL;
    size_t n = strlen(s) + 1;
    char *r = (char *)malloc(n);
    memcpy(r, s, n);
    return r;
}
static inline stTree *stTree_construct(void) {
    stTree *t = new stTree;
    t->branchLength = INFINITY;
    t->label = NULL;
    t->clientData = NULL;
    t->parent = NULL;
    return t;
}
static inline void stTree_destruct(stTree *t) {
    if (t == NULL) return;
    for (size_t i = 0; i < t->kids.size(); i++) stTree_destruct(t->kids[i]);
    free(t->label);
    delete t;
}
static inline stTree *stTree_getParent(stTree *t) { return t->parent; }
static inline void stTree_setParent(stTree *t, stTree *p) {
    if (t->parent != NULL) {
        std::vector<stTree *> &k = t->parent->kids;
        for (size_t i = 0; i < k.size(); i++)
            if (k[i] == t) { k.erase(k.begin() + i); break; }
    }
    t->parent = p;
    if (p != NULL) p->kids.push_back(t);
}
static inline int64_t stTree_getChildNumber(stTree *t) { return (int64_t)t->kids.size(); }
static inline stTree *stTree_getChild(stTree *t, int64_t i) { return t->kids[(size_t)i]; }
static inline void stTree_setChild(stTree *t, int64_t i, stTree *c) { t->kids[(size_t)i] = c; }
static inline const char *stTree_getLabel(stTree *t) { return t->label; }
static inline void stTree_setLabel(stTree *t, const char *l) {
    char *c = stString_copy(l);
    free(t->label);
    t->label = c;
}
static inline double stTree_getBranchLength(stTree *t) { return t->branchLength; }
static inline void stTree_setBranchLength(stTree *t, double d) { t->branchLength = d; }
static inline void *stTree_getClientData(stTree *t) { return t->clientData; }
static inline void stTree_setClientData(stTree *t, void *d) { t->clientData = d; }
static inline stTree *stTree_findChild(stTree *t, const char *label) {
    if (t->label != NULL && strcmp(t->label, label) == 0) return t;
    for (size_t i = 0; i < t->kids.size(); i++) {
        stTree *r = stTree_findChild(t->kids[i], label);
        if (r != NULL) return r;
    }
    return NULL;
}
static inline bool stTree_equals(stTree *a, stTree *b) {
    if ((a->label == NULL) != (b->label == NULL)) return false;
    if (a->label != NULL && strcmp(a->label, b->label) != 0) return false;
    if (a->branchLength != b->branchLength) return false;
    if (a->kids.size() != b->kids.size()) return false;
    for (size_t i = 0; i < a->kids.size(); i++)
        if (!stTree_equals(a->kids[i], b->kids[i])) return false;
    return true;
}
static inline void stTreeShim_print(stTree *t, std::string &out) {
    if (!t->kids.empty()) {
        out += '(';
        for (size_t i = 0; i < t->kids.size(); i++) {
            if (i) out += ',';
            stTreeShim_print(t->kids[i], out);
        }
        out += ')';
    }
    if (t->label != NULL) out += t->label;
    if (t->branchLength != INFINITY) {
        char buf[64];
        snprintf(buf, sizeof buf, ":%g", t->branchLength);
        out += buf;
    }
}
static inline char *stTree_getNewickTreeString(stTree *t) {
    std::string s;
    stTreeShim_print(t, s);
    s += ';';
    return stString_copy(s.c_str());
}
static inline stTree *stTreeShim_parse(const char *&p) {
    stTree *t = stTree_construct();
    while (*p == ' ' || *p == '\n' || *p == '\t') p++;
    if (*p == '(') {
        p++;
        for (;;) {
            stTree *c = stTreeShim_parse(p);
            stTree_setParent(c, t);
            while (*p == ' ' || *p == '\n' || *p == '\t') p++;
            if (*p == ',') { p++; continue; }
            if (*p == ')') { p++; break; }
            break; /* malformed: stop */
        }
    }
    const char *s = p;
    while (*p && *p != ':' && *p != ',' && *p != ')' && *p != ';' && *p != '(') p++;
    if (p > s) {
        std::string l(s, p - s);
        stTree_setLabel(t, l.c_str());
    }
    if (*p == ':') {
        p++;
        char *e;
        t->branchLength = strtod(p, &e);
        p = e;
    }
    return t;
}
static inline stTree *stTree_parseNewickString(const char *s) {
    const char *p = s;
    return stTreeShim_parse(p);
}
#endif
