/* ndt_internal.h -- shared between the C and CUDA translation units of libndt_b200 */
#ifndef NDT_INTERNAL_H
#define NDT_INTERNAL_H
#include "ndt_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* records a thread-local message and returns `code` so callers can `return ndt_set_error(...)` */
int ndt_set_error(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
#ifdef __cplusplus
}
#endif
#endif
