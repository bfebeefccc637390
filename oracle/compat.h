/* Force-included when compiling the reference sources on Linux.
 * The reference's OBJ loader (src/trianglemesh.cpp:114-116) uses the MSVC-only
 * errno_t / fopen_s; this shim supplies them so the untouched sources compile.
 * Test infrastructure only. */
#ifndef FLIP_ORACLE_COMPAT_H
#define FLIP_ORACLE_COMPAT_H
#ifdef __cplusplus
#include <cstdio>
#include <cerrno>
typedef int errno_t;
static inline errno_t fopen_s(FILE **f, const char *name, const char *mode) {
    *f = fopen(name, mode);
    return *f ? 0 : errno;
}
#endif
#endif
