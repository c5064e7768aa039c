/* error.c -- thread-local last-error string (ndt_b200.h: the library never exits) */
#include <stdarg.h>
#include <stdio.h>
#include "ndt_internal.h"

static __thread char g_err[512] = "";

int ndt_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

const char *ndt_b200_last_error(void) { return g_err; }
const char *ndt_b200_version(void) { return "ndt_b200 0.1 (sm_100a)"; }
