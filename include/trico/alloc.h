/* Shim for /root/reference/trico/alloc.h:12-30.  Buffers handed out by the library (raw codec
 * results, float/double attribute lists) come from malloc, so callers release them with
 * trico_free exactly as with the reference. */
#ifndef TRICO_B200_ALLOC_SHIM_H
#define TRICO_B200_ALLOC_SHIM_H
#include <stdlib.h>
#define trico_malloc(size) malloc(size)
#define trico_calloc(num, size) calloc((num), (size))
#define trico_realloc(ptr, size) realloc((ptr), (size))
#define trico_free(ptr) free(ptr)
#endif
