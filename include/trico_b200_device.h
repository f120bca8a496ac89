/*
 * trico_b200_device.h - device-level C ABI of the B200 trico hot path.
 *
 * Plain C, plain pointers and sizes.  Everything here runs hand-written sm_100a kernels; there is
 * no CPU fallback: every entry point returns 0 when CUDA is unavailable or a launch fails.
 *
 * This is the layer the archive API (trico_b200.h, the drop-in for the reference's
 * trico/trico.h:36-94) is built on, and what bench.py times with device-resident buffers.
 * Pointer arguments named d_* must be device pointers.
 *
 * Reference functions each entry point replaces are cited as /root/reference paths.
 */
#ifndef TRICO_B200_DEVICE_H
#define TRICO_B200_DEVICE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef TB200_API
#define TB200_API __attribute__((visibility("default")))
#endif

typedef struct tb200_ctx tb200_ctx;

/* One context = one device + one CUDA stream + a workspace (look-back descriptors, tickets,
 * legacy predictor tables).  `cuda_stream` may be NULL (the context creates its own stream) or an
 * existing cudaStream_t (e.g. torch's current stream) so callers can time with their own events.
 * Returns NULL if no CUDA device is usable. */
TB200_API tb200_ctx* tb200_ctx_create(int device, void* cuda_stream);
TB200_API void tb200_ctx_destroy(tb200_ctx* ctx);
TB200_API void* tb200_ctx_stream(tb200_ctx* ctx);
TB200_API int tb200_ctx_device(tb200_ctx* ctx);
/* cudaSetDevice(ctx's device) for the calling thread (streams, events and pinned memory created by
 * the helpers below belong to the current device) */
TB200_API int tb200_ctx_make_current(tb200_ctx* ctx);
TB200_API int tb200_set_device(int device);
/* frees the workspaces the context has grown (they are re-created on demand) */
TB200_API void tb200_ctx_trim(tb200_ctx* ctx);
TB200_API int tb200_ctx_sync(tb200_ctx* ctx);                 /* 1 ok, 0 CUDA error */
TB200_API const char* tb200_last_error(void);                /* text of the last failure in this thread */
/* Number of kernels launched through this context so far (bench.py's gpu_launches). */
TB200_API uint64_t tb200_ctx_launch_count(tb200_ctx* ctx);

/* ---- v1 stream geometry (DESIGN.md "v1 wire format") ---- */
#define TB200_V1_FIXED_BYTES 15u      /* u8 type, u32 count, u8 codec_info, u8 log2_chunk, u64 payload_bytes */
/* codec of a stream type: 1 = FPC, 2 = LZ4 byte planes, 0 = not a stream type (trico.h:11-34) */
TB200_API int tb200_stream_layout(int type, int* wordsize, int* ncomp, int* per_count);
TB200_API uint64_t tb200_v1_nchunks(int type, uint32_t count, int log2_chunk);
/* worst-case size of a whole v1 stream (fixed header + size table + payload) */
TB200_API uint64_t tb200_v1_stream_bound(int type, uint32_t count, int log2_chunk);
TB200_API int tb200_default_log2_chunk(int type, uint32_t count);

/* ---- chunked FPC: replaces trico_transpose_*_aos_to_soa + trico_compress[_double_precision]
 *      (trico/transpose_aos_to_soa.c:8-82, trico/floating_point_stream_compression.c:86,:576) ---- */
/* d_in: AoS, n*ncomp words of `wordsize` bytes.  Writes u16 sizes[nchunks] and the packed chunk
 * payloads; the payload byte count goes to d_total (aligned u64) and, little-endian, to the 8
 * bytes at d_total_field (may be NULL).  Asynchronous on the context stream. */
TB200_API int tb200_fpc_encode(tb200_ctx* ctx, int wordsize, int ncomp, const void* d_in, uint64_t n, int log2_chunk,
                     int e1, int e2, uint8_t* d_sizes, uint8_t* d_payload, uint8_t* d_total_field, uint64_t* d_total);
/* inverse: trico_decompress[_double_precision] + trico_transpose_*_soa_to_aos
 * (floating_point_stream_compression.c:212,:803; transpose_aos_to_soa.c:18-82) */
TB200_API int tb200_fpc_decode(tb200_ctx* ctx, int wordsize, int ncomp, const uint8_t* d_sizes, const uint8_t* d_payload,
                     uint64_t payload_bytes, uint64_t n, int log2_chunk, int e1, int e2, void* d_out);

/* ---- reference-format (v0) FPC streams, whole stream = one predictor chain ---- */
/* Encodes `nstreams` component streams (component c = d_in[j*stride + c]) into slots of
 * out_stride bytes at d_out; d_nbytes[c] receives each stream's length (5-byte header included).
 * Byte-identical to trico_compress / trico_compress_double_precision. */
TB200_API int tb200_fpc_encode_v0(tb200_ctx* ctx, int wordsize, const void* d_in, uint32_t n, uint32_t stride, int nstreams,
                        int e1, int e2, uint8_t* d_out, uint64_t out_stride, uint32_t* d_nbytes);
TB200_API uint64_t tb200_fpc_v0_bound(int wordsize, uint32_t n);
/* Decodes `nstreams` v0 streams found at d_base + offsets[c] (host array of offsets);
 * lengths[c] = bytes of the device buffer that are readable from the stream's first byte on (the
 * kernel fetches the stream in 16-byte pieces and never reads beyond this; NULL: at least 4 KiB
 * past the last byte of every stream are readable); hash_info[c] = first byte of each stream (host
 * knows it from the archive) selects table sizes.  Element j of stream c goes to d_out[j*stride + c]. */
TB200_API int tb200_fpc_decode_v0(tb200_ctx* ctx, int wordsize, const uint8_t* d_base, const uint64_t* offsets, const uint64_t* lengths,
                        const uint8_t* hash_info, int nstreams, uint32_t expect_n, void* d_out, uint32_t stride);

/* ---- chunked byte-plane + LZ4: replaces trico_transpose_uint{16,32,64}_aos_to_soa +
 *      LZ4_compress_default (transpose_aos_to_soa.c:84-147, lz4/lz4.c:1271) ---- */
TB200_API int tb200_lz4_encode(tb200_ctx* ctx, int wordsize, const void* d_in, uint64_t n, int log2_chunk,
                     uint8_t* d_sizes, uint8_t* d_payload, uint8_t* d_total_field, uint64_t* d_total);
/* inverse: LZ4_decompress_safe (lz4/lz4.c:2078) + trico_transpose_uint*_soa_to_aos */
TB200_API int tb200_lz4_decode(tb200_ctx* ctx, int wordsize, const uint8_t* d_sizes, const uint8_t* d_payload,
                     uint64_t payload_bytes, uint64_t n, int log2_chunk, void* d_out);
/* same without the stream synchronisation: malformed blocks set *d_status (a device word the caller
 * zeroed) to 1; the caller reads it after its own synchronisation */
TB200_API int tb200_lz4_decode_async(tb200_ctx* ctx, int wordsize, const uint8_t* d_sizes, const uint8_t* d_payload,
                     uint64_t payload_bytes, uint64_t n, int log2_chunk, void* d_out, uint32_t* d_status);
/* reference-format (v0) planes: `nplanes` whole-plane LZ4 blocks at d_base + offsets[p] of
 * nbytes[p] compressed bytes, each decoding to n bytes; merged into n elements of nplanes bytes. */
/* reference-format planes from the GPU: plane p of the n wordsize-byte elements as ONE LZ4 block
 * (what LZ4_compress_default over the plane is to the reference's writers, trico.c:346) at
 * d_out + p * out_stride, its size in d_nbytes[p]; out_stride >= tb200_lz4_v0_bound(n) */
TB200_API uint64_t tb200_lz4_v0_bound(uint64_t n);
TB200_API int tb200_lz4_encode_v0(tb200_ctx* ctx, int wordsize, const void* d_in, uint64_t n, uint8_t* d_out, uint64_t out_stride, uint64_t* d_nbytes);
TB200_API int tb200_lz4_decode_v0(tb200_ctx* ctx, int nplanes, const uint8_t* d_base, const uint64_t* offsets,
                        const uint32_t* nbytes, uint64_t n, void* d_out);

/* ---- whole streams ---- */
/* Encodes one v1 stream (header + size table + payload, assembled in-kernel) into d_out.
 * The stream's total byte count is written to *d_stream_bytes (device u64).  `count` is the value
 * stored in the stream header (trico.c:221). */
TB200_API int tb200_encode_stream(tb200_ctx* ctx, int type, const void* d_data, uint32_t count, int log2_chunk,
                        uint8_t* d_out, uint64_t out_cap, uint64_t* d_stream_bytes);
/* Decodes one v1 stream whose 15 header bytes are given in host memory (`header`) and whose full
 * bytes (starting at the type byte) are on the device at d_stream. */
TB200_API int tb200_decode_stream(tb200_ctx* ctx, const uint8_t* header, const uint8_t* d_stream, uint64_t stream_bytes, void* d_out);

/* ---- batches of streams (many small meshes: BASELINE C5) ----
 * The streams of a batch run concurrently on several CUDA streams of the context and nothing is
 * synchronised or read back per stream.  tb200_encode_streams: stream i (types[i], d_data[i],
 * counts[i], default chunking) is encoded into its slot of d_arena (>= tb200_batch_arena_bytes),
 * then the streams are packed back to back into d_packed, in order: the bytes tb200_encode_stream
 * produces one by one.  d_sizes: n device u64 (stream sizes); d_prefix: 2n + 2 device u64 - entries
 * [0, n] receive the offsets of the streams in d_packed and the total, the rest is scratch.
 * tb200_decode_streams: stream i = sizes[i] bytes at d_packed + offsets[i] (host arrays) with its
 * 15 header bytes at headers + 15 i (host); malformed LZ4 blocks set *d_status (device word the
 * caller zeroed) - checked by the caller after its own synchronisation. */
TB200_API uint64_t tb200_batch_arena_bytes(int n, const int* types, const uint32_t* counts);
TB200_API int tb200_encode_streams(tb200_ctx* ctx, int n, const int* types, const void* const* d_data, const uint32_t* counts,
                        uint8_t* d_arena, uint64_t arena_cap, uint8_t* d_packed, uint64_t packed_cap,
                        uint64_t* d_sizes, uint64_t* d_prefix);
TB200_API int tb200_decode_streams(tb200_ctx* ctx, int n, const uint8_t* headers, const uint8_t* d_packed, const uint64_t* offsets,
                        const uint64_t* sizes, void* const* d_out, uint32_t* d_status);

/* ---- multi-GPU: chunk-sharded streams; one process per GPU, NCCL loaded at run time ----
 * A stream is cut into contiguous chunk-aligned shares, one per rank.  Every rank encodes its share
 * with the ordinary kernels (no data-path collective); an all-gather exchanges the compressed sizes;
 * optionally the shares are moved device to device to their final offsets on the root (grouped
 * ncclSend / ncclRecv).  The assembled stream is byte-identical to tb200_encode_stream on one GPU. */
typedef struct tb200_comm tb200_comm;
TB200_API int tb200_comm_unique_id(uint8_t* id128);                 /* rank 0 creates it, the host distributes the 128 bytes */
TB200_API tb200_comm* tb200_comm_create(tb200_ctx* ctx, int rank, int world, const uint8_t* id128);
TB200_API void tb200_comm_destroy(tb200_comm* comm);
TB200_API int tb200_comm_rank(tb200_comm* comm);
TB200_API int tb200_comm_world(tb200_comm* comm);
/* share of `rank`: units [*first, *first + *n) of a stream of `count` units (vertices, triangles, ...) */
TB200_API int tb200_shard_range(int type, uint32_t count, int log2_chunk, int rank, int world, uint32_t* first, uint32_t* n);
/* collective; see device_api_comm.inc.  ms (may be NULL): [0] encode + size exchange, [1] assembly, device time */
TB200_API int tb200_encode_stream_sharded(tb200_ctx* ctx, tb200_comm* comm, int type, const void* d_local, uint32_t count_local,
                        uint32_t count_total, int log2_chunk, int root, int assemble,
                        uint8_t* d_out, uint64_t out_cap, uint64_t* d_stream_bytes, float* ms);
/* this rank's encoded share of the last sharded call (size table, payload): device pointers valid until the next call */
TB200_API int tb200_comm_local_share(tb200_comm* comm, const uint8_t** d_sizes, uint64_t* table_bytes, const uint8_t** d_payload, uint64_t* payload_bytes);
/* decodes units [first, first + n) of a whole v1 stream: what a rank does with the chunk range it is handed */
TB200_API int tb200_decode_stream_range(tb200_ctx* ctx, const uint8_t* header, const uint8_t* d_stream, uint64_t stream_bytes,
                        uint32_t first, uint32_t n, void* d_out);

/* ---- standalone transposes (the 14 exported trico_transpose_* symbols) ---- */
TB200_API int tb200_deinterleave(tb200_ctx* ctx, int wordsize, int ncomp, const void* d_aos, uint64_t n, void* const* d_comp);
TB200_API int tb200_interleave(tb200_ctx* ctx, int wordsize, int ncomp, void* d_aos, uint64_t n, const void* const* d_comp);

/* ---- memory helpers (thin wrappers so C / ctypes callers need no CUDA headers) ---- */
TB200_API void* tb200_device_alloc(uint64_t bytes);
TB200_API void tb200_device_free(void* d);
TB200_API void* tb200_host_alloc_pinned(uint64_t bytes);
TB200_API void tb200_host_free_pinned(void* h);
TB200_API int tb200_memcpy_h2d(tb200_ctx* ctx, void* d, const void* h, uint64_t bytes);   /* async on ctx stream */
TB200_API int tb200_memcpy_d2h(tb200_ctx* ctx, void* h, const void* d, uint64_t bytes);   /* async on ctx stream */
TB200_API int tb200_memset_d(tb200_ctx* ctx, void* d, int value, uint64_t bytes);         /* async on ctx stream */
TB200_API int tb200_device_count(void);
/* 1 if p points to device (or managed) memory, 0 for ordinary / pinned host memory */
TB200_API int tb200_pointer_is_device(const void* p);
/* timing on the context stream: event handles are opaque */
TB200_API void* tb200_event_create(void);
TB200_API void tb200_event_destroy(void* ev);
TB200_API int tb200_event_record(tb200_ctx* ctx, void* ev);
TB200_API float tb200_event_elapsed_ms(void* start, void* stop);   /* synchronises on `stop` */

/* extra streams / events (the archive layer overlaps H2D, kernels and D2H with them) */
TB200_API void* tb200_stream_create(void);
TB200_API void tb200_stream_destroy(void* stream);
TB200_API int tb200_memcpy_h2d_on(void* stream, void* d, const void* h, uint64_t bytes);
TB200_API int tb200_memcpy_d2h_on(void* stream, void* h, const void* d, uint64_t bytes);
TB200_API int tb200_event_record_on(void* stream, void* ev);
TB200_API int tb200_stream_wait_event(void* stream, void* ev);
TB200_API int tb200_event_sync(void* ev);
TB200_API int tb200_stream_sync(void* stream);
TB200_API void* tb200_event_create_notiming(void);
TB200_API int tb200_pointer_is_pinned_host(const void* p);

/* exponents of the chunked FPC writer (hash_info byte 0x12): 4-entry FCM, 16-entry DFCM tables */
#define TB200_V1_E1 2
#define TB200_V1_E2 4

#ifdef __cplusplus
}
#endif
#endif
