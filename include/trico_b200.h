/*
 * trico_b200.h - the trico C API, served by hand-written sm_100a CUDA kernels.
 *
 * Drop-in for the reference library's three public headers:
 *   /root/reference/trico/trico.h:36-94                           archive API (54 functions)
 *   /root/reference/trico/transpose_aos_to_soa.h:12-38            14 transposes
 *   /root/reference/trico/floating_point_stream_compression.h:11-17  4 raw codec functions
 * Same names, argument meaning, ownership and 1/0/NULL error convention.  The shims under
 * include/trico/ let reference callers keep their `#include <trico/trico.h>` lines.
 *
 * Differences a caller can observe (DESIGN.md "quirk policy"):
 *   - archives written here carry version 1 (chunked streams); version-0 archives written by the
 *     reference are read by the legacy kernels.  An archive with no stream still reports version 0.
 *   - trico_write_uv_per_*_double emit the double tags 6/8 the readers expect (the reference emits
 *     the float tags 5/7, trico.c:622,:627, and cannot read its own output).
 *   - trico_read_attributes_uint8 fills *attrib (the reference decompresses into the pointer
 *     variable itself, trico.c:1439).
 *   - any pointer argument may be a CUDA device pointer; the copy is then skipped.
 *   - there is no CPU fallback: without a usable CUDA device every data call fails (0 / NULL).
 */
#ifndef TRICO_B200_H
#define TRICO_B200_H

#include <stdint.h>

#if defined(__cplusplus)
extern "C" {
#endif

#ifndef TRICO_API
#define TRICO_API __attribute__((visibility("default")))
#endif

/* wire values of the per-stream type byte: trico/trico.h:11-34 */
enum trico_stream_type
  {
  trico_empty = 0,
  trico_vertex_float_stream = 1,
  trico_vertex_double_stream = 2,
  trico_triangle_uint32_stream = 3,
  trico_triangle_uint64_stream = 4,
  trico_uv_per_vertex_float_stream = 5,
  trico_uv_per_vertex_double_stream = 6,
  trico_uv_per_triangle_float_stream = 7,
  trico_uv_per_triangle_double_stream = 8,
  trico_vertex_normal_float_stream = 9,
  trico_vertex_normal_double_stream = 10,
  trico_triangle_normal_float_stream = 11,
  trico_triangle_normal_double_stream = 12,
  trico_vertex_color_stream = 13,
  trico_triangle_color_stream = 14,
  trico_attribute_float_stream = 15,
  trico_attribute_double_stream = 16,
  trico_attribute_uint8_stream = 17,
  trico_attribute_uint16_stream = 18,
  trico_attribute_uint32_stream = 19,
  trico_attribute_uint64_stream = 20
  };

/* ---- archive lifetime: trico.c:126, :158, :183 ---- */
TRICO_API void* trico_open_archive_for_writing(uint64_t initial_buffer_size);
TRICO_API void* trico_open_archive_for_reading(const uint8_t* data, uint64_t data_size);   /* borrows `data` until close */
TRICO_API void trico_close_archive(void* archive);

/* ---- writer state: trico.c:191, :197 ---- */
TRICO_API uint8_t* trico_get_buffer_pointer(void* archive);      /* host memory, valid until the next write / close */
TRICO_API uint64_t trico_get_size(void* archive);

/* ---- reader state: trico.c:203, :209, :860-941, :1670 ---- */
TRICO_API uint32_t trico_get_version(void* archive);
TRICO_API enum trico_stream_type trico_get_next_stream_type(void* archive);
TRICO_API uint32_t trico_get_number_of_vertices(void* archive);
TRICO_API uint32_t trico_get_number_of_triangles(void* archive);
TRICO_API uint32_t trico_get_number_of_uvs(void* archive);
TRICO_API uint32_t trico_get_number_of_normals(void* archive);
TRICO_API uint32_t trico_get_number_of_colors(void* archive);
TRICO_API uint32_t trico_get_number_of_attributes(void* archive);
TRICO_API int trico_skip_next_stream(void* archive);

/* ---- writers.  vec3 float/double: trico.c:215/:380 (x,y,z FPC streams) ---- */
TRICO_API int trico_write_vertices(void* archive, const float* vertices, uint32_t nr_of_vertices);
TRICO_API int trico_write_vertices_double(void* archive, const double* vertices, uint32_t nr_of_vertices);
TRICO_API int trico_write_vertex_normals(void* archive, const float* normals, uint32_t nr_of_normals);
TRICO_API int trico_write_vertex_normals_double(void* archive, const double* normals, uint32_t nr_of_normals);
TRICO_API int trico_write_triangle_normals(void* archive, const float* normals, uint32_t nr_of_normals);
TRICO_API int trico_write_triangle_normals_double(void* archive, const double* normals, uint32_t nr_of_normals);
/* index streams, 3 indices per triangle, byte planes + LZ4: trico.c:323, :444 */
TRICO_API int trico_write_triangles(void* archive, const uint32_t* tria_indices, uint32_t nr_of_triangles);
TRICO_API int trico_write_triangles_long(void* archive, const uint64_t* tria_indices, uint32_t nr_of_triangles);
/* vec2: trico.c:534, :582.  The per-triangle float writer stores 3*nr_of_uv_positions (trico.c:579) */
TRICO_API int trico_write_uv_per_vertex(void* archive, const float* uv, uint32_t nr_of_uv_positions);
TRICO_API int trico_write_uv_per_vertex_double(void* archive, const double* uv, uint32_t nr_of_uv_positions);
TRICO_API int trico_write_uv_per_triangle(void* archive, const float* uv, uint32_t nr_of_uv_positions);
TRICO_API int trico_write_uv_per_triangle_double(void* archive, const double* uv, uint32_t nr_of_uv_positions);
/* colours (u32 RGBA) and attribute lists: trico.c:698, :279, :301, :630, :657, :755, :770 */
TRICO_API int trico_write_vertex_colors(void* archive, const uint32_t* color, uint32_t nr_of_colors);
TRICO_API int trico_write_triangle_colors(void* archive, const uint32_t* color, uint32_t nr_of_colors);
TRICO_API int trico_write_attributes_float(void* archive, const float* attrib, uint32_t nr_of_attribs);
TRICO_API int trico_write_attributes_double(void* archive, const double* attrib, uint32_t nr_of_attribs);
TRICO_API int trico_write_attributes_uint8(void* archive, const uint8_t* attrib, uint32_t nr_of_attribs);
TRICO_API int trico_write_attributes_uint16(void* archive, const uint16_t* attrib, uint32_t nr_of_attribs);
TRICO_API int trico_write_attributes_uint32(void* archive, const uint32_t* attrib, uint32_t nr_of_attribs);
TRICO_API int trico_write_attributes_uint64(void* archive, const uint64_t* attrib, uint32_t nr_of_attribs);

/* ---- readers: trico.c:943-1668.  The caller pre-allocates count*arity elements and passes the
 * address of its pointer; NULL skips the stream.  Exception kept from the reference: the float /
 * double attribute readers malloc the result and overwrite *attrib (trico.c:1377, :1408). ---- */
TRICO_API int trico_read_vertices(void* archive, float** vertices);
TRICO_API int trico_read_vertices_double(void* archive, double** vertices);
TRICO_API int trico_read_vertex_normals(void* archive, float** normals);
TRICO_API int trico_read_vertex_normals_double(void* archive, double** normals);
TRICO_API int trico_read_triangle_normals(void* archive, float** normals);
TRICO_API int trico_read_triangle_normals_double(void* archive, double** normals);
TRICO_API int trico_read_triangles(void* archive, uint32_t** triangles);
TRICO_API int trico_read_triangles_long(void* archive, uint64_t** triangles);
TRICO_API int trico_read_uv_per_vertex(void* archive, float** uv);
TRICO_API int trico_read_uv_per_vertex_double(void* archive, double** uv);
TRICO_API int trico_read_uv_per_triangle(void* archive, float** uv);
TRICO_API int trico_read_uv_per_triangle_double(void* archive, double** uv);
TRICO_API int trico_read_vertex_colors(void* archive, uint32_t** color);
TRICO_API int trico_read_triangle_colors(void* archive, uint32_t** color);
TRICO_API int trico_read_attributes_float(void* archive, float** attrib);
TRICO_API int trico_read_attributes_double(void* archive, double** attrib);
TRICO_API int trico_read_attributes_uint8(void* archive, uint8_t** attrib);
TRICO_API int trico_read_attributes_uint16(void* archive, uint16_t** attrib);
TRICO_API int trico_read_attributes_uint32(void* archive, uint32_t** attrib);
TRICO_API int trico_read_attributes_uint64(void* archive, uint64_t** attrib);

/* ---- raw FPC codec, reference stream format, byte-identical output:
 * floating_point_stream_compression.h:11-17.  *out is malloc'd (free with free()). ---- */
TRICO_API void trico_compress(uint32_t* nr_of_compressed_bytes, uint8_t** out, const float* input, const uint32_t number_of_floats, uint32_t hash1_size_exponent, uint32_t hash2_size_exponent);
TRICO_API void trico_decompress(uint32_t* number_of_floats, float** out, const uint8_t* compressed);
TRICO_API void trico_compress_double_precision(uint32_t* nr_of_compressed_bytes, uint8_t** out, const double* input, const uint32_t number_of_doubles, uint64_t hash1_size_exponent, uint64_t hash2_size_exponent);
TRICO_API void trico_decompress_double_precision(uint32_t* number_of_doubles, double** out, const uint8_t* compressed);

/* ---- transposes into caller-allocated arrays: transpose_aos_to_soa.h:12-38 ---- */
TRICO_API void trico_transpose_xyz_aos_to_soa(float** x, float** y, float** z, const float* vertices, uint32_t nr_of_vertices);
TRICO_API void trico_transpose_xyz_soa_to_aos(float** vertices, const float* x, const float* y, const float* z, uint32_t nr_of_vertices);
TRICO_API void trico_transpose_xyz_aos_to_soa_double_precision(double** x, double** y, double** z, const double* vertices, uint32_t nr_of_vertices);
TRICO_API void trico_transpose_xyz_soa_to_aos_double_precision(double** vertices, const double* x, const double* y, const double* z, uint32_t nr_of_vertices);
TRICO_API void trico_transpose_uv_aos_to_soa(float** u, float** v, const float* uv, uint32_t nr_of_uv_positions);
TRICO_API void trico_transpose_uv_soa_to_aos(float** uv, const float* u, const float* v, uint32_t nr_of_uv_positions);
TRICO_API void trico_transpose_uv_aos_to_soa_double_precision(double** u, double** v, const double* uv, uint32_t nr_of_uv_positions);
TRICO_API void trico_transpose_uv_soa_to_aos_double_precision(double** uv, const double* u, const double* v, uint32_t nr_of_uv_positions);
TRICO_API void trico_transpose_uint16_aos_to_soa(uint8_t** b1, uint8_t** b2, const uint16_t* indices, uint32_t nr_of_indices);
TRICO_API void trico_transpose_uint16_soa_to_aos(uint16_t** indices, const uint8_t* b1, const uint8_t* b2, uint32_t nr_of_indices);
TRICO_API void trico_transpose_uint32_aos_to_soa(uint8_t** b1, uint8_t** b2, uint8_t** b3, uint8_t** b4, const uint32_t* indices, uint32_t nr_of_indices);
TRICO_API void trico_transpose_uint32_soa_to_aos(uint32_t** indices, const uint8_t* b1, const uint8_t* b2, const uint8_t* b3, const uint8_t* b4, uint32_t nr_of_indices);
TRICO_API void trico_transpose_uint64_aos_to_soa(uint8_t** b1, uint8_t** b2, uint8_t** b3, uint8_t** b4, uint8_t** b5, uint8_t** b6, uint8_t** b7, uint8_t** b8, const uint64_t* indices, uint32_t nr_of_indices);
TRICO_API void trico_transpose_uint64_soa_to_aos(uint64_t** indices, const uint8_t* b1, const uint8_t* b2, const uint8_t* b3, const uint8_t* b4, const uint8_t* b5, const uint8_t* b6, const uint8_t* b7, const uint8_t* b8, uint32_t nr_of_indices);

/* ---- extensions (not in the reference) ---- */
/* text of the last failure on this thread ("" if none) */
TRICO_API const char* trico_b200_last_error(void);
/* chunk geometry for streams written to this archive from now on: log2 of values per FPC chunk
 * and log2 of bytes per LZ4 plane block; 0 keeps the default (9 / 8 / 14). */
TRICO_API int trico_b200_set_chunking(void* archive, int fpc_log2_values, int lz4_log2_bytes);
/* kernels launched on behalf of this archive so far */
TRICO_API uint64_t trico_b200_launch_count(void* archive);
/* 0: write this archive in the reference's own format (readable by an unmodified reference decoder),
 * 1: the chunked container (default; TRICO_B200_FORMAT=0 changes the default for unmodified callers).
 * Must be called before the first trico_write_*. */
TRICO_API int trico_b200_set_format(void* archive, int version);

#if defined(__cplusplus)
}
#endif
#endif /* TRICO_B200_H */
