/* zng_b200.h -- C ABI of the B200 (sm_100a) implementation of zlib-ng's chunked-DEFLATE hot path.
 *
 * This is the shim the C11 host library (include/zlib-ng-b200.h: zng_deflateInit2 / zng_deflate /
 * zng_crc32 ...) calls, and what a maintainer of the reference would bind at the accelerator seam
 * the reference already has for IBM Z DFLTCC (deflate.c:72-106 hook macros; DEFLATE_HOOK at
 * deflate.c:1039 replaces the whole strategy call).  Plain pointers and sizes only; `stream` is a
 * cudaStream_t passed as void* (NULL = the legacy default stream).  All `d_` pointers are device
 * memory on the context's device, all `h_` pointers are host memory.
 *
 * Return values are zlib-ng's (zlib-ng.h.in:180-188): 0 = Z_OK, negative = error; CUDA failures
 * map to ZNG_B200_CUDA_ERROR and zng_b200_last_error() names them.  There is no CPU fallback:
 * every entry point fails when the device or the kernels are unavailable.
 *
 * Frozen parameters: windowBits 15, memLevel 8, Z_DEFAULT_STRATEGY,
 * chunk <= 65536 bytes, levels 1 (deflate_quick), 2 (deflate_fast) and 3..6 (deflate_medium; below level 5 it runs
 * without its look-ahead branch, deflate_medium.c:151,234; levels 5-6 look ahead and run fizzle_matches).  Level 6 is
 * zlib-ng's Z_DEFAULT_COMPRESSION.
 */
#ifndef ZNG_B200_H
#define ZNG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZNG_B200_OK             0
#define ZNG_B200_STREAM_ERROR (-2)   /* bad / unsupported parameters (Z_STREAM_ERROR) */
#define ZNG_B200_DATA_ERROR   (-3)
#define ZNG_B200_MEM_ERROR    (-4)   /* allocation failed (Z_MEM_ERROR) */
#define ZNG_B200_BUF_ERROR    (-5)   /* output capacity too small (Z_BUF_ERROR) */
#define ZNG_B200_NOT_RESUMABLE (-6)   /* zng_b200_inflate_stream_feed: a header the resumable decoder leaves to the one-shot path */
#define ZNG_B200_CUDA_ERROR   (-100) /* a CUDA call failed; see zng_b200_last_error() */

#define ZNG_B200_CHUNK_MAX 65536u

/* flush modes accepted for chunk compression (zlib-ng.h.in:165-171) */
#define ZNG_B200_SYNC_FLUSH 2
#define ZNG_B200_FULL_FLUSH 3
#define ZNG_B200_FINISH     4

typedef struct zng_b200_ctx zng_b200_ctx;

/* ---- context ------------------------------------------------------------------------------ */
int         zng_b200_device_count(void);
/* device < 0: the calling thread's current device.  One context per host thread (like one
 * zng_stream per thread, README of the reference); contexts are independent. */
int         zng_b200_ctx_create(zng_b200_ctx **ctx, int device);
void        zng_b200_ctx_destroy(zng_b200_ctx *ctx);
int         zng_b200_ctx_device(const zng_b200_ctx *ctx);
int         zng_b200_ctx_sm_count(const zng_b200_ctx *ctx);
const char *zng_b200_last_error(const zng_b200_ctx *ctx);
int         zng_b200_sync(zng_b200_ctx *ctx, void *stream);

/* pinned host memory for the host-buffer entry points (cudaHostAlloc / cudaFreeHost) */
void       *zng_b200_host_alloc(size_t bytes);
void        zng_b200_host_free(void *p);

/* ---- K1/K2: chunk compression, device resident ------------------------------------------- */
/* Size of one chunk's output slot: an upper bound of what zng_deflate can emit for `chunk_len`
 * input bytes at levels 1-2 (deflate.c:709-781 zng_deflateBound: 9 bits per literal + block and
 * flush overhead), rounded so that slots stay 16-byte aligned with read slack for the gather. */
size_t      zng_b200_deflate_bound(size_t chunk_len);

/* Replaces: configuration_table[level].func(s, flush) called from zng_deflate (deflate.c:1036-1043)
 * for ceil(n/chunk) independent chunks, each on a stream whose window and hash state are empty --
 * i.e. the byte output of "zng_deflateReset; zng_deflate(Z_FULL_FLUSH)" per chunk (flush = 3; also
 * 2), or of "zng_deflate(Z_FINISH)" per chunk (flush = 4: BFINAL set, no trailing stored block).
 *   d_in        n input bytes
 *   d_out       ceil(n/chunk) slots of out_stride bytes (out_stride >= zng_b200_deflate_bound(chunk),
 *               multiple of 16; d_out 16-byte aligned); slot i receives chunk i's raw-deflate bytes
 *   d_sizes[i]  compressed size of chunk i
 *   d_crcs[i]   zng_crc32(0, chunk i)    (may be NULL)       crc32.c:27-41
 *   d_adlers[i] zng_adler32(1, chunk i)  (may be NULL)       adler32.c:15-28
 * level 1 = deflate_quick (deflate_quick.c:47-130), level 2 = deflate_fast (deflate_fast.c:19-104),
 * levels 3..6 = deflate_medium (deflate_medium.c:22-278) with configuration_table's {nice, chain} (deflate.c:160-163). */
int zng_b200_deflate_chunks(zng_b200_ctx *ctx, const void *d_in, size_t n, uint32_t chunk, int level, int flush,
                            void *d_out, size_t out_stride, uint32_t *d_sizes, uint32_t *d_crcs,
                            uint32_t *d_adlers, void *stream);

/* Replaces: per chunk, a fresh zng_deflateInit2(level, Z_DEFLATED, -15, 8, 0) + zng_deflateSetDictionary(the 32768
 * stream bytes in front of the chunk) (deflate.c:456-512) + one zng_deflate(flush) -- pigz's default, dependent-chunk
 * mode: chunks stay independent units of work, but chunk i may reference the last 32 KiB of chunk
 * i-1, so the concatenation (Z_SYNC_FLUSH joins) is one raw-deflate stream that must be inflated in order.  The first
 * chunk of the buffer has no dictionary.  Levels 1..6, chunk 65536 only; other arguments as zng_b200_deflate_chunks.  Level 1 runs
 * the speculative warp parser on absolute positions (K1p); levels 2-6 run the reference's own window state per chunk, with its two real
 * slides and refills (K2w, csrc/deflate_window.cu). */
int zng_b200_deflate_chunks_primed(zng_b200_ctx *ctx, const void *d_in, size_t n, uint32_t chunk, int level, int flush,
                                   void *d_out, size_t out_stride, uint32_t *d_sizes, uint32_t *d_crcs,
                                   uint32_t *d_adlers, void *stream);

/* The same for a buffer that is NOT the start of the stream: have_halo != 0 says the 32768 bytes in front of d_in are readable and
 * are the stream bytes that precede it, so the first chunk is primed like every other (a rank's shard of a multi-GPU dependent
 * stream; zng_b200_halo_exchange below delivers those bytes from the previous rank). */
int zng_b200_deflate_chunks_primed_at(zng_b200_ctx *ctx, const void *d_in, size_t n, uint32_t chunk, int level, int flush, int have_halo,
                                      void *d_out, size_t out_stride, uint32_t *d_sizes, uint32_t *d_crcs,
                                      uint32_t *d_adlers, void *stream);

/* Debug aid (ZLIB_DEBUG's Tracevv token trace analogue, deflate_p.h:39-44,72): as above for level 1,
 * additionally writing the LZ77 token stream of chunk i to d_tokens[i*tok_stride ...]: a literal is
 * its byte value, a match is 0x80000000 | len << 16 | dist, the list ends with 0x40000000. */
int zng_b200_deflate_chunks_trace(zng_b200_ctx *ctx, const void *d_in, size_t n, uint32_t chunk, int level, int flush,
                                  void *d_out, size_t out_stride, uint32_t *d_sizes, uint32_t *d_tokens,
                                  uint32_t tok_stride, void *stream);

/* ---- stream assembly ---------------------------------------------------------------------- */
/* d_offsets[i] = base + sum_{j<i} d_sizes[j] for i in 0..nchunks (nchunks+1 entries). */
int zng_b200_chunk_offsets(zng_b200_ctx *ctx, const uint32_t *d_sizes, uint32_t nchunks, uint64_t base,
                           uint64_t *d_offsets, void *stream);
/* Copy chunk i's d_sizes[i] bytes from its slot to d_dst + d_offsets[i] (the concatenation is the
 * raw-deflate stream; what flush_pending (deflate.c:789-807) does one chunk at a time). */
int zng_b200_gather_chunks(zng_b200_ctx *ctx, const void *d_slots, size_t out_stride, const uint32_t *d_sizes,
                           const uint64_t *d_offsets, uint32_t nchunks, void *d_dst, void *stream);

/* ---- multi-GPU stream assembly (csrc/multi.cu) ---------------------------------------------- */
/* One rank per GPU, each with its own context; rank r owns a contiguous chunk range of one stream.  Replaces, across ranks,
 * what a single zng_stream accumulates over its zng_deflate(Z_FULL_FLUSH) calls: total_out (deflate.c:789-807), the running
 * CRC-32 (deflate.c:1021,1092 via crc32_combine, crc32_braid_comb.c:16-24) and the gzip trailer (deflate.c:1091-1096).
 * The only collective is one NCCL allgather of the per-chunk (compressed size, crc32) pairs, 8 bytes per chunk.
 * NCCL is bound at run time (libnccl.so.2); without it these entry points return ZNG_B200_STREAM_ERROR. */
typedef struct zng_b200_comm zng_b200_comm;
#define ZNG_B200_COMM_ID_BYTES 128
/* rank 0: a fresh NCCL unique id (128 bytes) that the caller hands to every rank by its own means (MPI, a file, a socket) */
int  zng_b200_comm_unique_id(void *id, size_t cap);
/* every rank, collectively: a communicator of `nranks` ranks on the context's device (nranks == 1: no NCCL needed) */
int  zng_b200_comm_create(zng_b200_ctx *ctx, int nranks, int rank, const void *id, zng_b200_comm **out);
/* wrap a communicator the application already has (an ncclComm_t passed as void*); it is not destroyed with the wrapper */
int  zng_b200_comm_adopt(zng_b200_ctx *ctx, void *nccl_comm, zng_b200_comm **out);
void zng_b200_comm_destroy(zng_b200_comm *comm);
int  zng_b200_comm_size(const zng_b200_comm *comm);
int  zng_b200_comm_rank(const zng_b200_comm *comm);
const char *zng_b200_comm_error(const zng_b200_comm *comm);
/* Collective.  In: this rank's per-chunk sizes and CRC-32s (device), its chunk count and byte count (every chunk of the stream
 * but the very last is `chunk` bytes).  Out: d_offsets_local[i] = byte offset of this rank's chunk i in the whole stream,
 * i in 0..nchunks_local (base + exclusive scan over ALL ranks' chunks); *h_stream_end = base + total compressed bytes;
 * *h_crc32 = zng_crc32 of the whole input (fold of all chunk CRCs); *h_total_in = input bytes of all ranks. */
int  zng_b200_stream_index_multi(zng_b200_comm *comm, const uint32_t *d_sizes, const uint32_t *d_crcs, uint32_t nchunks_local,
                                 uint32_t chunk, size_t n_local, uint64_t base, uint64_t *d_offsets_local,
                                 uint64_t *h_stream_end, uint32_t *h_crc32, uint64_t *h_total_in, void *stream);
/* Collective, dependent (primed) mode: rank r sends the last 32768 bytes of its shard to rank r+1 and receives rank r-1's into d_halo,
 * which the caller places directly in front of its input ([d_halo, d_halo + 32768 + n_local) contiguous); then
 * zng_b200_deflate_chunks_primed_at(..., have_halo = rank > 0, ...) makes the N-GPU dependent stream equal the 1-GPU one. */
int  zng_b200_halo_exchange(zng_b200_comm *comm, const void *d_in, size_t n_local, void *d_halo, void *stream);
/* Collective: the whole pigz-style step for this rank's n_local bytes -- chunk compression (64 KiB chunks, Z_FULL_FLUSH ends),
 * the allgather, scan and fold above, and the gather of this rank's chunks into d_packed.  The gzip file is
 *   header (10 bytes, rank 0) | rank 0's bytes | rank 1's bytes | ... | "03 00" crc32 isize (10 bytes, at *h_file_bytes - 10)
 * and this rank's *h_my_bytes packed bytes belong at file offset *h_my_offset.  d_slots / d_sizes / d_crcs as for
 * zng_b200_deflate_chunks; d_offsets_local: nchunks_local + 1 entries; h_header10 / h_trailer10 are filled on every rank. */
int  zng_b200_gzip_multi(zng_b200_comm *comm, const void *d_in, size_t n_local, int level,
                         void *d_slots, size_t out_stride, uint32_t *d_sizes, uint32_t *d_crcs, uint64_t *d_offsets_local,
                         void *d_packed, size_t packed_cap, uint64_t *h_my_offset, uint64_t *h_my_bytes,
                         uint64_t *h_file_bytes, uint8_t *h_header10, uint8_t *h_trailer10, void *stream);

/* ---- K3: checksums ------------------------------------------------------------------------ */
/* Per-tile zng_crc32(0, .) / zng_adler32(1, .) of consecutive tile_bytes-sized pieces (<= 65536). */
int zng_b200_checksum_chunks(zng_b200_ctx *ctx, const void *d_in, size_t n, uint32_t tile_bytes,
                             uint32_t *d_crcs, uint32_t *d_adlers, void *stream);
/* *d_result = crc32_combine-fold of per-tile CRCs of an n-byte buffer onto `init`
 * (crc32_braid_comb.c:16-24; equals zng_crc32_z(init, buf, n)). */
int zng_b200_crc32_fold(zng_b200_ctx *ctx, const uint32_t *d_crcs, uint32_t ntiles, uint32_t tile_bytes, size_t n,
                        uint32_t init, uint32_t *d_result, void *stream);
/* adler32_combine-fold (adler32.c:32-54; equals zng_adler32_z(init, buf, n)). */
int zng_b200_adler32_fold(zng_b200_ctx *ctx, const uint32_t *d_adlers, uint32_t ntiles, uint32_t tile_bytes, size_t n,
                          uint32_t init, uint32_t *d_result, void *stream);
/* Replaces functable.crc32 / functable.adler32 (functable.h:27,33) for a device buffer:
 * *d_result = zng_crc32_z(init, buf, n) / zng_adler32_z(init, buf, n). */
int zng_b200_crc32(zng_b200_ctx *ctx, const void *d_buf, size_t n, uint32_t init, uint32_t *d_result, void *stream);
int zng_b200_adler32(zng_b200_ctx *ctx, const void *d_buf, size_t n, uint32_t init, uint32_t *d_result, void *stream);

/* ---- K4: batched inflate of independent members -------------------------------------------- */
/* Replaces: zng_inflateInit2(&s, window_bits); zng_inflate(&s, Z_FINISH) (inflate.c:219-255, :476-1201, with
 * inflate_fast inffast_tpl.h:53-318 and zng_inflate_table inftrees.c:30-295 underneath) for n_members
 * independent members, each on a fresh stream.  window_bits as in zng_inflateInit2: -15 raw, 15 zlib,
 * 31 gzip, 47 auto-detect.
 *   d_in / d_in_off     member i occupies d_in[d_in_off[i] .. d_in_off[i+1])   (n_members + 1 offsets)
 *   d_out / d_out_off   member i's output goes to d_out[d_out_off[i] ..], capacity d_out_off[i+1]-d_out_off[i]
 *   d_sizes[i]          bytes produced (strm.total_out)
 *   d_checks[i]         strm.adler after the call: crc32 (gzip) / adler32 (zlib) of the output (may be NULL)
 *   d_status[i]         what zng_inflate returns: Z_STREAM_END 1, Z_NEED_DICT 2, Z_DATA_ERROR -3, Z_BUF_ERROR -5
 *   d_in_used[i]        bytes consumed (strm.total_in) for members that reached Z_STREAM_END (may be NULL)
 *   d_detail[i]         low byte: id of strm->msg (zng_b200_inflate_msg), 0x100: output capacity reached,
 *                       0x200: input exhausted (may be NULL) */
int zng_b200_inflate_members(zng_b200_ctx *ctx, const void *d_in, const uint64_t *d_in_off, uint32_t n_members,
                             int window_bits, void *d_out, const uint64_t *d_out_off, uint32_t *d_sizes,
                             uint32_t *d_checks, int32_t *d_status, uint32_t *d_in_used, uint32_t *d_detail, void *stream);
/* the reference's strm->msg string for a detail value (NULL when there is none) */
const char *zng_b200_inflate_msg(uint32_t detail);

/* ---- host-buffer entry points (what zng_deflate / zng_crc32 of the host library call) ----- */
/* Compress h_in[0..n) as ceil(n/chunk) chunks and write the concatenated raw-deflate stream to
 * h_out.  final != 0: the last chunk is compressed with Z_FINISH semantics (BFINAL block, no
 * stored block); otherwise every chunk ends with the Z_FULL_FLUSH marker.  *crc32 / *adler32
 * (optional) receive zng_crc32_z(0,...) / zng_adler32_z(1,...) of the whole input.  Host<->device
 * copies are pipelined with the kernels; pinned buffers (zng_b200_host_alloc) avoid staging.  Level 1 from 32 MiB on
 * is streamed: one persistent parse kernel takes the chunks as the copy engine delivers them and every finished 32 MiB
 * slab is emitted, packed and copied out next to it (environment: ZNG_B200_STREAMED=0 selects the slab pipeline that
 * levels 2-6 use, ZNG_B200_TRACE=1 prints the timeline). */
int zng_b200_deflate_host(zng_b200_ctx *ctx, const void *h_in, size_t n, uint32_t chunk, int level, int final,
                          void *h_out, size_t out_cap, size_t *out_len, uint32_t *crc32, uint32_t *adler32);

/* Replaces: what zng_deflate does on a raw level-1 stream after zng_deflateSetDictionary (deflate.c:456-512), pigz's
 * dependent mode -- every 65536-byte piece of h_in is compressed by a fresh stream primed with the 32768 bytes in front of it
 * (h_dict, 32768 bytes, for the first piece; NULL = the first piece has no dictionary).  Pieces end with the sync-flush
 * marker, final != 0: Z_FINISH on the last one.  Synchronous (not pipelined). */
int zng_b200_deflate_host_primed(zng_b200_ctx *ctx, const void *h_dict, const void *h_in, size_t n, int final,
                                 void *h_out, size_t out_cap, size_t *out_len, uint32_t *crc32, uint32_t *adler32);
/* The same for levels 1..6 (deflate_quick / deflate_fast / deflate_medium after zng_deflateSetDictionary): levels 2-6 run on the
 * reference's own window state per chunk, real slides included (csrc/deflate_window.cu). */
int zng_b200_deflate_host_primed_level(zng_b200_ctx *ctx, const void *h_dict, const void *h_in, size_t n, int level, int final,
                                       void *h_out, size_t out_cap, size_t *out_len, uint32_t *crc32, uint32_t *adler32);
/* Host-buffer form of zng_b200_inflate_members (same arrays, all in host memory; offsets are copied to the
 * device with the data, results come back in the h_ arrays). */
int zng_b200_inflate_members_host(zng_b200_ctx *ctx, const void *h_in, const uint64_t *h_in_off, uint32_t n_members,
                                  int window_bits, void *h_out, const uint64_t *h_out_off, uint32_t *h_sizes,
                                  uint32_t *h_checks, int32_t *h_status, uint32_t *h_in_used, uint32_t *h_detail);
/* Inflate ONE stream (raw -15 / zlib 15 / gzip 31 / auto 47) held in host memory: the call zng_inflate of the host
 * library makes.  A stream whose pieces end with Z_FULL_FLUSH (what zng_deflate of this library and pigz-style tools
 * emit) is split at its 00 00 FF FF markers on the device and its segments are decoded in parallel, one warp each
 * (inflateSync's marker search, inflate.c:1290-1360, applied to every marker at once); any other stream, and every
 * error, goes through the one-member path above and reports exactly what the reference's zng_inflate(Z_FINISH) reports.
 *   *status   Z_STREAM_END 1 / Z_NEED_DICT 2 / Z_DATA_ERROR -3 / Z_BUF_ERROR -5
 *   *detail   as d_detail above; 0x400 with Z_BUF_ERROR: cap was too small and *out_len holds the size needed
 *             (nothing was written) */
int zng_b200_inflate_stream_host(zng_b200_ctx *ctx, const void *h_in, size_t n, int window_bits, void *h_out, size_t cap,
                                 size_t *out_len, size_t *in_used, uint32_t *check, int32_t *status, uint32_t *detail);
/* Resumable inflate of ONE stream fed piecewise -- what zng_inflate(Z_NO_FLUSH) of the host library calls (inflate.c:476: the
 * reference's inflate() is resumable at byte granularity; here the granularity is the deflate block).  Each feed decodes from the last
 * block boundary the previous one reached to the last boundary inside what has arrived; decoded output is handed over at once,
 * consumed input is dropped, the check value and the trailer are verified at the end.  Linear work, piecewise output.
 *   *status    0 fed (more input, or more room in h_out, needed) / 1 stream end and every byte delivered / -3 data error (*detail =
 *              message id for zng_b200_inflate_msg) / ZNG_B200_NOT_RESUMABLE: FDICT / FHCRC / malformed header -- nothing consumed,
 *              use zng_b200_inflate_stream_host / 2: more than the backlog limit (256 MiB, env ZNG_B200_INFLATE_BACKLOG) of decoded
 *              output is waiting for the caller: some was handed over, NO input was taken (*in_used = 0) -- call again
 *   The device keeps 32 KiB of history plus the output not yet taken, not the whole output (delivered bytes are dropped in steps of
 *   64 MiB, env ZNG_B200_INFLATE_COMPACT), so streams of any length work in bounded memory; the check value is carried along.
 *   *in_used   bytes of this call's input that belong to the stream (< n only when the stream ended inside them, or 0 with status 2)
 *   *out_len   bytes written to h_out (at most cap; what does not fit stays pending: call again with n == 0) */
typedef struct zng_b200_inflate_stream zng_b200_inflate_stream;
int  zng_b200_inflate_stream_open(zng_b200_ctx *ctx, int window_bits, zng_b200_inflate_stream **out);
int  zng_b200_inflate_stream_feed(zng_b200_inflate_stream *st, const void *h_in, size_t n, void *h_out, size_t cap,
                                  size_t *in_used, size_t *out_len, int32_t *status, uint32_t *detail, uint32_t *check);
void zng_b200_inflate_stream_close(zng_b200_inflate_stream *st);
int zng_b200_crc32_host(zng_b200_ctx *ctx, const void *h_buf, size_t n, uint32_t init, uint32_t *result);
int zng_b200_adler32_host(zng_b200_ctx *ctx, const void *h_buf, size_t n, uint32_t init, uint32_t *result);

/* ---- operator surface: functable.h:26-42 and deflate.h:121-131, one device operator per call --------------
 * The per-position operators of the reference are __device__ functions inside K1 / K2 / K4 here; these entry points
 * run exactly that device code on caller-supplied device buffers so each can be checked against its counterpart.
 * (crc32 / adler32 of the table are zng_b200_crc32 / zng_b200_adler32 above.) */
/* functable.compare256 (compare256_c.c:12-43) for n_pairs operand pairs: d_out[i] = first differing byte (0..256) of the
 * 256 bytes at d_a + i*stride and d_b + i*stride.  Both buffers must be readable 16 bytes past the last operand. */
int zng_b200_op_compare256(zng_b200_ctx *ctx, const void *d_a, const void *d_b, size_t stride, uint32_t n_pairs,
                           uint32_t *d_out, void *stream);
/* functable.longest_match (match_tpl.h:26-280, level-2 parameters: chain 4, nice 8) for n_q independent queries on one
 * window: query i = (strstart d_pos[i], cur_match d_cand[i]); lookahead = n - d_pos[i]; d_prev = the 32768-entry prev[]
 * table.  d_len[i] = returned length when it is >= 4 else 0, d_start[i] = match_start.  The window must be readable
 * 280 bytes past n (what the reference's window slack provides). */
int zng_b200_op_longest_match(zng_b200_ctx *ctx, const void *d_window, uint32_t n, const uint16_t *d_prev,
                              const uint32_t *d_pos, const uint32_t *d_cand, uint32_t n_q, uint32_t *d_len,
                              uint32_t *d_start, void *stream);
/* insert_string(s, str, count) (insert_string_tpl.h:82-104) on d_head[65536] / d_prev[32768]. */
int zng_b200_op_insert_string(zng_b200_ctx *ctx, const void *d_window, uint16_t *d_head, uint16_t *d_prev, uint32_t str,
                              uint32_t count, void *stream);
/* functable.chunkmemset_safe (chunkset_tpl.h:112-283): d_out[pos + i] = d_out[pos + i - dist] for i in 0..len, byte serial. */
int zng_b200_op_chunkmemset(zng_b200_ctx *ctx, void *d_out, uint32_t pos, uint32_t dist, uint32_t len, void *stream);

/* functable.longest_match for the strategy of `level` 2..6 (configuration_table, deflate.c:142-168: deflate_fast {nice 8, chain 4},
 * deflate_medium {16,6} {32,24} {32,32} {128,128}); other arguments as zng_b200_op_longest_match (which is level 2).  d_len[i] is
 * the returned length when it is >= 4, else 0 (the strategies discard shorter results). */
int zng_b200_op_longest_match_level(zng_b200_ctx *ctx, const void *d_window, uint32_t n, const uint16_t *d_prev,
                                    const uint32_t *d_pos, const uint32_t *d_cand, uint32_t n_q, int level, uint32_t *d_len,
                                    uint32_t *d_start, void *stream);
/* quick_insert_string(s, str) (insert_string_tpl.h:58-75): inserts position str, *d_old_head = the head entry it replaced. */
int zng_b200_op_quick_insert_string(zng_b200_ctx *ctx, const void *d_window, uint16_t *d_head, uint16_t *d_prev, uint32_t str,
                                    uint32_t *d_old_head, void *stream);
/* functable.slide_hash (arch/generic/slide_hash_c.c:15-52): head[65536] and prev[wsize] rebased by wsize, smaller entries -> 0. */
int zng_b200_op_slide_hash(zng_b200_ctx *ctx, uint16_t *d_head, uint16_t *d_prev, uint32_t wsize, void *stream);
/* checksum-while-copy (functable.crc32_fold_copy / adler32_fold_copy; read_buf, deflate.c:1190-1212) for host buffers: the bytes
 * go to the device, are checksummed there and come back into h_dst. */
int zng_b200_crc32_copy_host(zng_b200_ctx *ctx, void *h_dst, const void *h_src, size_t n, uint32_t init, uint32_t *result);
int zng_b200_adler32_copy_host(zng_b200_ctx *ctx, void *h_dst, const void *h_src, size_t n, uint32_t init, uint32_t *result);

/* ---- the host-callable operator table: struct functable_s (functable.h:26-42), the same 15 slots in the same order, plus the
 * three hash callbacks of deflate_state (deflate.h:121-131).  Every slot runs on the GPU through the calling thread's context
 * (device scratch is owned by the context: no allocation per call); the state arguments are views of the fields the reference's
 * operators read and write, in HOST memory. */
struct zng_b200_crc32_fold {           /* crc32.h:8-14: the reference keeps 4 x 128 bits of folding state + value; here the   */
    uint8_t  fold[64];                 /* state IS the running CRC-32 (fold[] is unused and kept for layout)                  */
    uint32_t value;
};
struct zng_b200_match_state {          /* deflate_state as longest_match / insert_string / slide_hash see it (deflate.h:153-280) */
    const uint8_t *window;             /* s->window; window_len bytes of data (<= 65536), the rest reads as zeros            */
    uint32_t window_len;
    uint16_t *head;                    /* s->head, 65536 entries (HASH_SIZE, deflate.h:81-85)                                 */
    uint16_t *prev;                    /* s->prev, 32768 entries (w_size)                                                     */
    uint32_t strstart;                 /* s->strstart                                                                         */
    uint32_t lookahead;                /* s->lookahead (informative: longest_match uses window_len - strstart)                */
    uint32_t match_start;              /* out: s->match_start                                                                 */
    int32_t  level;                    /* 2..6: max_chain_length / nice_match of configuration_table (deflate.c:142-168)      */
};
struct zng_b200_functable {
    void     (*force_init)(void);
    uint32_t (*adler32)(uint32_t adler, const uint8_t *buf, size_t len);
    uint32_t (*adler32_fold_copy)(uint32_t adler, uint8_t *dst, const uint8_t *src, size_t len);
    uint8_t *(*chunkmemset_safe)(uint8_t *out, uint8_t *from, unsigned len, unsigned left);
    uint32_t (*chunksize)(void);
    uint32_t (*compare256)(const uint8_t *src0, const uint8_t *src1);
    uint32_t (*crc32)(uint32_t crc, const uint8_t *buf, size_t len);
    void     (*crc32_fold)(struct zng_b200_crc32_fold *crc, const uint8_t *src, size_t len, uint32_t init_crc);
    void     (*crc32_fold_copy)(struct zng_b200_crc32_fold *crc, uint8_t *dst, const uint8_t *src, size_t len);
    uint32_t (*crc32_fold_final)(struct zng_b200_crc32_fold *crc);
    uint32_t (*crc32_fold_reset)(struct zng_b200_crc32_fold *crc);
    /* inflate_fast(strm, start) (inffast_tpl.h:53): the reference enters it in the middle of a block with its decode tables in
     * inflate_state; here the slot decodes as far as strm's input and output allow -- one zng_inflate(strm, Z_SYNC_FLUSH) of the
     * host library (the batched device form is zng_b200_inflate_members).  strm is a zng_stream* (include/zlib-ng.h). */
    void     (*inflate_fast)(void *strm, uint32_t start);
    /* returns what the reference returns (match_tpl.h:268-279: best_len, 2 when no candidate improves, clipped to lookahead) for
     * levels 2..6; s->match_start is written when the result is >= 4 */
    uint32_t (*longest_match)(struct zng_b200_match_state *s, uint16_t cur_match);
    /* LONGEST_MATCH_SLOW serves deflate_slow (levels 7-9), outside the hot path: the slot is present, returns 0
     * and records "longest_match_slow: outside the hot path" in the context (Z_STREAM_ERROR semantics, no CPU fallback) */
    uint32_t (*longest_match_slow)(struct zng_b200_match_state *s, uint16_t cur_match);
    void     (*slide_hash)(struct zng_b200_match_state *s);
};
const struct zng_b200_functable *zng_b200_functable_get(void);
/* the hash callbacks (deflate.h:121-131; insert_string.c:11-19): HASH_CALC = (val * 2654435761) >> 16 */
uint32_t zng_b200_update_hash(uint32_t h, uint32_t val);
void     zng_b200_insert_string(struct zng_b200_match_state *s, uint32_t str, uint32_t count);
uint16_t zng_b200_quick_insert_string(struct zng_b200_match_state *s, uint32_t str);



#ifdef __cplusplus
}
#endif
#endif /* ZNG_B200_H */
