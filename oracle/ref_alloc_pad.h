/* TEST INFRASTRUCTURE ONLY - force-included (-include) when compiling the UNMODIFIED reference
 * sources into oracle/_ref/libtrico_ref.so.
 *
 * The reference under-sizes its own output buffer: trico_compress allocates
 *   4n + 3(n+7)/8 + (n&7)                      (floating_point_stream_compression.c:95)
 * but can write 5 + 4n + 3*ceil(n/8) + (8 - n%8)%8 bytes (5-byte stream header :120-126, pad
 * slots :196-204), i.e. up to 12 bytes more for incompressible or very short inputs; glibc then
 * aborts in the shrinking realloc (:209) with "corrupted size vs. prev_size".  The double
 * variant has the same shape (:585).  Padding every allocation by 64 bytes lets the reference
 * run on those inputs without touching a line of its algorithm. */
#include <stdlib.h>
#define malloc(s) malloc((s) + 64)
#define realloc(p, s) realloc((p), (s) + 64)
