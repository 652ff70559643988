/* TEST INFRASTRUCTURE ONLY (oracle build shim): the handful of CuTest entry points the reference's
 * unit-test sources use (api/tests/*.cpp, liftover/tests/halLiftoverTests.cpp), so those files can be
 * compiled where they lie and run against the oracle build.  Included inside extern "C" { } by the
 * reference headers, hence everything here is plain C compatible with C++. */
#ifndef ORACLE_SHIM_CUTEST_H
#define ORACLE_SHIM_CUTEST_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct CuString { char *buffer; size_t length, size; } CuString;
typedef struct CuTest CuTest;
typedef void (*CuTestFunction)(CuTest *);
struct CuTest { const char *name; CuTestFunction function; int failed; int ran; char message[1024]; };
typedef struct CuSuite { int count; CuTest *list[2048]; int failCount; } CuSuite;

static inline CuString *CuStringNew(void) {
    CuString *s = (CuString *)calloc(1, sizeof(CuString));
    s->size = 1 << 20;
    s->buffer = (char *)calloc(1, s->size);
    return s;
}
static inline void CuStringAppend(CuString *s, const char *t) {
    size_t n = strlen(t);
    if (s->length + n + 1 < s->size) { memcpy(s->buffer + s->length, t, n + 1); s->length += n; }
}
static inline CuSuite *CuSuiteNew(void) { return (CuSuite *)calloc(1, sizeof(CuSuite)); }
static inline void CuSuiteAdd(CuSuite *suite, const char *name, CuTestFunction f) {
    CuTest *t = (CuTest *)calloc(1, sizeof(CuTest));
    t->name = name; t->function = f;
    suite->list[suite->count++] = t;
}
#define SUITE_ADD_TEST(SUITE, TEST) CuSuiteAdd(SUITE, #TEST, TEST)
static inline void CuSuiteAddSuite(CuSuite *dst, CuSuite *src) {
    for (int i = 0; i < src->count; i++) dst->list[dst->count++] = src->list[i];
}
static inline void CuSuiteRun(CuSuite *s) {
    for (int i = 0; i < s->count; i++) {
        CuTest *t = s->list[i];
        t->function(t);
        t->ran = 1;
        if (t->failed) s->failCount++;
    }
}
static inline void CuSuiteSummary(CuSuite *s, CuString *out) {
    for (int i = 0; i < s->count; i++) CuStringAppend(out, s->list[i]->failed ? "F" : ".");
    CuStringAppend(out, "\n\n");
}
static inline void CuSuiteDetails(CuSuite *s, CuString *out) {
    char buf[1400];
    if (s->failCount == 0) {
        snprintf(buf, sizeof buf, "OK (%d %s)\n", s->count, s->count == 1 ? "test" : "tests");
        CuStringAppend(out, buf);
        return;
    }
    for (int i = 0; i < s->count; i++)
        if (s->list[i]->failed) {
            snprintf(buf, sizeof buf, "FAILED %s: %s\n", s->list[i]->name, s->list[i]->message);
            CuStringAppend(out, buf);
        }
    snprintf(buf, sizeof buf, "\n!!!FAILURES!!!\nRuns: %d Fails: %d\n", s->count, s->failCount);
    CuStringAppend(out, buf);
}
static inline void CuFail_Line(CuTest *t, const char *file, int line, const char *m) {
    if (!t->failed) snprintf(t->message, sizeof t->message, "%s:%d: %s", file, line, m ? m : "");
    t->failed = 1;
}
#define CuFail(tc, ms) CuFail_Line((tc), __FILE__, __LINE__, (ms))
#define CuAssertTrue(tc, cond) do { if (!(cond)) CuFail_Line((tc), __FILE__, __LINE__, "assert failed: " #cond); } while (0)
#define CuAssertIntEquals(tc, ex, ac) do { if ((long)(ex) != (long)(ac)) CuFail_Line((tc), __FILE__, __LINE__, "int mismatch: " #ex " vs " #ac); } while (0)
#define CuAssertStrEquals(tc, ex, ac) do { if (strcmp((ex), (ac)) != 0) CuFail_Line((tc), __FILE__, __LINE__, "str mismatch: " #ex " vs " #ac); } while (0)
#endif
