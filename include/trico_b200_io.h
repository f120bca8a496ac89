/*
 * trico_b200_io.h - the mesh front-end of the trico tools on the B200.
 *
 * Drop-in for the two STL readers of the reference's trico_io library and for the normal
 * recomputation of its decoder tool; the heavy part - vertex de-duplication, a sort + unique +
 * index remap over the 3T facet corners - runs as hand-written sm_100a kernels
 * (trico_b200/csrc/stl.cuh).  No CPU fallback: every entry point returns 0 without a CUDA device.
 *
 *   trico_read_stl        replaces /root/reference/trico_io/iostl.c:141-195 (declared iostl.h)
 *   trico_read_stl_full   replaces /root/reference/trico_io/iostl.c:197-259
 *   (both call trico_remove_duplicate_vertices, iostl.c:70-138: quicksort :60-68 under the
 *    comparator :8-19, walk :107-137)
 *   trico_write_stl       replaces /root/reference/trico_io/iostl.c:261-320
 *   trico_b200_triangle_normals   replaces the loop at /root/reference/tools/trico_decoder/main.c:439-470
 *
 * Same names, arguments, return values and ownership as the reference: buffers handed back come
 * from malloc (callers release them with trico_free = free, trico/alloc.h:12-30).
 */
#ifndef TRICO_B200_IO_H
#define TRICO_B200_IO_H

#include <stdint.h>
#include <stddef.h>
#include "trico_b200_device.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Binary STL file -> indexed mesh with duplicate vertices removed: vertices in lexicographic
 * (x, y, z) order, 3 indices per triangle in file order.  Returns 0 for a missing file, an ASCII
 * ("solid") file or a truncated one, exactly as the reference. */
TB200_API int trico_read_stl(uint32_t* nr_of_vertices, float** vertices, uint32_t* nr_of_triangles, uint32_t** triangles,
                             const char* filename);
/* The same, plus the facet normals (3 floats per triangle) and attribute words of the file. */
TB200_API int trico_read_stl_full(uint32_t* nr_of_vertices, float** vertices, uint32_t* nr_of_triangles, uint32_t** triangles,
                                  float** normals, uint16_t** attributes, const char* filename);

/* Indexed mesh -> binary STL file, byte for byte the file the reference writes (iostl.c:261-320): NULL normals
 * are written as zeros, NULL attributes as zero words.  The facets are gathered on the device. */
TB200_API int trico_write_stl(const float* vertices, const uint32_t* triangles, const uint32_t nr_of_triangles,
                              const float* triangle_normals, const uint16_t* attributes, const char* filename);

/* Triangle normals from an indexed mesh, host buffers, bit-identical to the reference decoder's loop. */
TB200_API int trico_b200_triangle_normals(const float* vertices, uint32_t nr_of_vertices, const uint32_t* triangles,
                                          uint32_t nr_of_triangles, float* triangle_normals);

/* ---- device level (what bench.py and the tests time with device-resident buffers) ---- */

/* d_facets: the 50-byte facet records of a binary STL file (the file from byte 84 on), on the device.
 * d_vertices must hold 9 * ntriangles floats (the worst case: nothing shared), d_triangles
 * 3 * ntriangles indices; d_normals (3 * ntriangles floats) and d_attributes (ntriangles words) may be
 * NULL.  *nr_of_vertices receives the vertex count (the call synchronises the context's stream).
 * ntriangles <= 1431655765 (3T corner ids are 32-bit, as in the reference). */
TB200_API int tb200_stl_dedup(tb200_ctx* ctx, const uint8_t* d_facets, uint32_t ntriangles, float* d_vertices,
                              uint32_t* d_triangles, float* d_normals, uint16_t* d_attributes, uint32_t* nr_of_vertices);
/* bytes of device scratch tb200_stl_dedup allocates (and frees) for ntriangles facets */
TB200_API uint64_t tb200_stl_dedup_scratch_bytes(uint32_t ntriangles);
/* how many of the twelve 8-bit sort passes the last tb200_stl_dedup of this thread ran (the rest had one digit) */
TB200_API int tb200_stl_last_sort_passes(void);
/* d_facets (50 * ntriangles bytes, 2-byte aligned) <- the facet records of trico_write_stl; d_normals and
 * d_attributes may be NULL */
TB200_API int tb200_stl_facets(tb200_ctx* ctx, const float* d_vertices, const uint32_t* d_triangles, uint32_t ntriangles,
                               const float* d_normals, const uint16_t* d_attributes, uint8_t* d_facets);
TB200_API int tb200_triangle_normals(tb200_ctx* ctx, const float* d_vertices, const uint32_t* d_triangles, uint32_t ntriangles,
                                     float* d_normals);

#ifdef __cplusplus
}
#endif
#endif
