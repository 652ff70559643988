/* TEST INFRASTRUCTURE ONLY (oracle build shim) for sonLib's commonC.h: getTempFile + stString_print. */
#ifndef ORACLE_SHIM_COMMONC_H
#define ORACLE_SHIM_COMMONC_H
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
static inline char *getTempFile(void) {
    char tmpl[] = "/tmp/halOracleXXXXXX";
    int fd = mkstemp(tmpl);
    if (fd >= 0) close(fd);
    return strdup(tmpl);
}
static inline char *stString_print(const char *fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return strdup(buf);
}
#endif
