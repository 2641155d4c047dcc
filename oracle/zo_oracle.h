/* oracle/zo_oracle.h -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or
 * executed from the product path (zlib-ng_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Plain-C restatement (NOT a copy) of the reference algorithms on the hot path that
 * SURVEY.md section 8 scopes: zlib-ng 2.2.2 level-1 deflate_quick, level-2 deflate_fast with the
 * trees.c block writer, CRC-32 / Adler-32 and their combine functions, compare256, and
 * inflate of raw/zlib/gzip streams.  Every function cites the reference file:line it
 * restates.  Frozen parameters: windowBits 15, memLevel 8,
 * Z_DEFAULT_STRATEGY, default build flags, x86-64.
 *
 * PARITY PINNING: this restatement is checked (tests/test_oracle_*.py) against
 *   - the reference's own KATs (test/test_crc32.cc, test/test_adler32.cc, test/infcover.c,
 *     test/test_compare256.cc), extracted into tests/golden/ by tests/golden/make_golden.py;
 *   - outputs of the unmodified reference compiled here (oracle/_ref/libzng_ref.so), both
 *     live (when oracle/_ref is present) and through committed digests in tests/golden/.
 */
#ifndef ZO_ORACLE_H
#define ZO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* flush values (zlib-ng.h.in:165-171) */
#define ZO_NO_FLUSH   0
#define ZO_SYNC_FLUSH 2
#define ZO_FULL_FLUSH 3
#define ZO_FINISH     4

/* ---- checksums (zo_checksum.c) ---- */
uint32_t zo_crc32(uint32_t crc, const uint8_t *buf, size_t len);
uint32_t zo_adler32(uint32_t adler, const uint8_t *buf, size_t len);
uint32_t zo_crc32_combine(uint32_t crc1, uint32_t crc2, int64_t len2);
uint32_t zo_crc32_combine_gen(int64_t len2);
uint32_t zo_crc32_combine_op(uint32_t crc1, uint32_t crc2, uint32_t op);
uint32_t zo_adler32_combine(uint32_t adler1, uint32_t adler2, int64_t len2);
uint32_t zo_compare256(const uint8_t *a, const uint8_t *b);
size_t zo_frame_gzip_members(const uint8_t *bodies, size_t stride, const uint32_t *sizes, const uint32_t *crcs,
                             uint32_t raw_len, uint32_t last_raw_len, size_t n, int xfl, uint8_t *out, uint64_t *off);
size_t zo_compare_chunks(const uint8_t *a, size_t stride_a, const uint32_t *sizes_a,
                         const uint8_t *b, size_t stride_b, const uint32_t *sizes_b, size_t n, size_t *first_bad);

/* ---- deflate (zo_deflate.c) ---- */
/* Upper bound of one chunk's output for the frozen parameters (deflate.c:709-781 analogue,
 * conservative): n + n/8 + 64. */
size_t zo_deflate_bound(size_t n);

/* One chunk, one fresh raw-deflate stream state, one zng_deflate(flush) call.
 * level 1 = deflate_quick, level 2 = deflate_fast.  flush in {ZO_SYNC_FLUSH, ZO_FULL_FLUSH,
 * ZO_FINISH}.  tail/tail_len: the bytes the reference would see when it over-reads past the
 * end of the chunk (level 2 only; pass NULL/0 for "zeros" = fresh stream, short chunk; full
 * 65536-byte chunks ignore it and use the post-slide alias chunk[32768+k]).
 * Returns the number of bytes written, or (size_t)-1 if cap is too small / bad arguments. */
size_t zo_deflate_chunk(const uint8_t *in, uint32_t len, int level, int flush,
                        uint8_t *out, size_t cap);

/* Same signature/semantics as refdrv_deflate_chunks (oracle/ref_driver.c). */
int zo_deflate_chunks(const uint8_t *in, size_t n, uint32_t chunk, int level, int flush,
                      uint8_t *out, size_t out_stride, uint32_t *sizes,
                      uint32_t *crcs, uint32_t *adlers, int nthreads);

/* Token trace of the LZ77 parse for debugging mismatches: tokens[i] = literal byte, or
 * (1u<<31) | (len << 16) | dist.  Returns token count (cap = capacity of tokens). */
/* pigz's dependent mode: every chunk after the first is compressed by a FRESH stream primed with
 * zng_deflateSetDictionary(the 32768 stream bytes in front of it) (deflate.c:456-512).  Levels 1-6, chunk 65536. */
int zo_deflate_chunks_primed(const uint8_t *in, size_t n, uint32_t chunk, int level, int flush,
                             uint8_t *out, size_t out_stride, uint32_t *sizes,
                             uint32_t *crcs, uint32_t *adlers, int nthreads);

/* levels 2-6, every chunk on a fresh stream, through the window engine of zo_deflate.c (a second restatement that keeps the
 * reference's own window / head / prev state; used to cross-check the chunk-coordinate restatement) */
int zo_deflate_chunks_fresh_window(const uint8_t *in, size_t n, uint32_t chunk, int level, int flush,
                                   uint8_t *out, size_t out_stride, uint32_t *sizes,
                                   uint32_t *crcs, uint32_t *adlers, int nthreads);

size_t zo_deflate_tokens(const uint8_t *in, uint32_t len, int level, uint32_t *tokens, size_t cap);
/* the same for a primed chunk: in[-32768 .. 0) must be readable (the dictionary) */
size_t zo_deflate_tokens_primed(const uint8_t *in, uint32_t len, int level, uint32_t *tokens, size_t cap);

/* Single operators (operator-surface tests): functable.longest_match at level 2, insert_string. */
uint32_t zo_longest_match_l2(const uint8_t *window, uint32_t avail, uint32_t n, const uint16_t *prev, uint32_t pos, uint32_t cand, uint32_t *start);
void zo_insert_string(const uint8_t *window, uint32_t avail, uint16_t *head, uint16_t *prev, uint32_t str, uint32_t count);

/* ---- inflate (zo_inflate.c) ---- */
/* Return codes follow zlib-ng.h.in:180-188. */
#define ZO_OK            0
#define ZO_STREAM_END    1
#define ZO_NEED_DICT     2
#define ZO_DATA_ERROR  (-3)
#define ZO_BUF_ERROR   (-5)

/* Inflate one complete stream held in memory.  window_bits: -15 raw, 15 zlib, 31 gzip,
 * 47 auto-detect zlib/gzip.  Writes up to out_cap bytes; *out_len = bytes produced,
 * *in_used = bytes consumed, *check = crc32 (gzip) or adler32 (zlib) of the output.
 * msg (optional) receives a pointer to a static error string identical to the
 * reference's strm->msg for the same fault. */
int zo_inflate(const uint8_t *in, size_t in_len, int window_bits, uint8_t *out, size_t out_cap,
               size_t *out_len, size_t *in_used, uint32_t *check, const char **msg);

/* Same signature/semantics as refdrv_inflate_members. */
int zo_inflate_members(const uint8_t *in, const uint64_t *in_off, size_t n_members,
                       uint8_t *out, const uint64_t *out_off, uint32_t *sizes,
                       uint32_t *crcs, int32_t *status, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
