/* zlib-ng.h -- native (zng_) API of the B200 host library, libzng_b200.so.
 *
 * Written from scratch for this repository: it declares, with the same names, argument meaning,
 * struct layout and return codes, the subset of zlib-ng 2.2.2's public interface that lies on the
 * chunked-DEFLATE hot path (reference: zlib-ng.h.in:99-119 zng_stream, :165-215 constants, :249
 * zng_deflate, :401 zng_inflate, :540 zng_deflateInit2, :751 zng_deflateBound, :821 zng_inflateInit2,
 * :1692-1774 checksums; symbols as versioned in zlib-ng.map:1-77).  A program that uses only this
 * subset relinks against libzng_b200.so unchanged; everything runs on the GPU (no CPU fallback).
 *
 * Contract of zng_deflate here (the reference's own behaviour for the same
 * call sequence): input is compressed as independent 65536-byte chunks.  A call
 *   zng_deflate(strm, Z_FULL_FLUSH)  with avail_in = k*65536 (+ optional short tail)
 * produces exactly the bytes the reference produces for k (+1) calls zng_deflate(Z_FULL_FLUSH) that
 * feed one 65536-byte piece each, and zng_deflate(strm, Z_FINISH) finishes the stream the way the
 * reference does when the remaining input is fed the same way with Z_FINISH on the last piece.
 * Z_NO_FLUSH only buffers input.  Supported parameters: levels 1 (deflate_quick), 2 (deflate_fast), 3..6 (deflate_medium) and
 * Z_DEFAULT_COMPRESSION (= 6),
 * method Z_DEFLATED, windowBits -15 / 15 / 31, memLevel 8, Z_DEFAULT_STRATEGY; anything else is
 * Z_STREAM_ERROR.
 */
#ifndef ZNGLIB_B200_H
#define ZNGLIB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZLIBNG_VERSION "2.2.2-b200"
#define ZLIBNG_VERNUM 0x020202F0L

typedef void *(*alloc_func)(void *opaque, unsigned int items, unsigned int size);
typedef void  (*free_func)(void *opaque, void *address);

struct internal_state;

/* Caller-owned, zero-initialised stream descriptor; same layout as the reference's (104 bytes on LP64). */
typedef struct zng_stream_s {
    const uint8_t         *next_in;
    uint32_t               avail_in;
    size_t                 total_in;
    uint8_t               *next_out;
    uint32_t               avail_out;
    size_t                 total_out;
    const char            *msg;
    struct internal_state *state;
    alloc_func             zalloc;
    free_func              zfree;
    void                  *opaque;
    int                    data_type;
    uint32_t               adler;
    unsigned long          reserved;
} zng_stream;

typedef zng_stream *zng_streamp;

/* flush values */
#define Z_NO_FLUSH      0
#define Z_PARTIAL_FLUSH 1
#define Z_SYNC_FLUSH    2
#define Z_FULL_FLUSH    3
#define Z_FINISH        4
#define Z_BLOCK         5
#define Z_TREES         6

/* return codes */
#define Z_OK              0
#define Z_STREAM_END      1
#define Z_NEED_DICT       2
#define Z_ERRNO         (-1)
#define Z_STREAM_ERROR  (-2)
#define Z_DATA_ERROR    (-3)
#define Z_MEM_ERROR     (-4)
#define Z_BUF_ERROR     (-5)
#define Z_VERSION_ERROR (-6)

/* compression levels / strategies / data types / method */
#define Z_NO_COMPRESSION        0
#define Z_BEST_SPEED            1
#define Z_BEST_COMPRESSION      9
#define Z_DEFAULT_COMPRESSION (-1)
#define Z_FILTERED         1
#define Z_HUFFMAN_ONLY     2
#define Z_RLE              3
#define Z_FIXED            4
#define Z_DEFAULT_STRATEGY 0
#define Z_BINARY  0
#define Z_TEXT    1
#define Z_ASCII   Z_TEXT
#define Z_UNKNOWN 2
#define Z_DEFLATED 8
#define Z_NULL NULL

#define MAX_WBITS 15
#define DEF_MEM_LEVEL 8

typedef int64_t z_off64_t;

const char *zlibng_version(void);

/* ---- deflate ---- */
int32_t zng_deflateInit(zng_stream *strm, int32_t level);
int32_t zng_deflateInit2(zng_stream *strm, int32_t level, int32_t method, int32_t windowBits, int32_t memLevel, int32_t strategy);
int32_t zng_deflate(zng_stream *strm, int32_t flush);
int32_t zng_deflateReset(zng_stream *strm);
/* deflate.c:456-512; raw level-1 streams, >= 32768-byte dictionaries: switches the stream to pigz's dependent-chunk mode */
int32_t zng_deflateSetDictionary(zng_stream *strm, const uint8_t *dictionary, uint32_t dictLength);
int32_t zng_deflateEnd(zng_stream *strm);
unsigned long zng_deflateBound(zng_stream *strm, unsigned long sourceLen);

/* ---- inflate ---- */
int32_t zng_inflateInit(zng_stream *strm);
int32_t zng_inflateInit2(zng_stream *strm, int32_t windowBits);
int32_t zng_inflate(zng_stream *strm, int32_t flush);
int32_t zng_inflateReset(zng_stream *strm);
int32_t zng_inflateEnd(zng_stream *strm);

/* ---- one-shot utilities (compress.c / uncompr.c of the reference) ---- */
int32_t zng_compress2(uint8_t *dest, size_t *destLen, const uint8_t *source, size_t sourceLen, int32_t level);
int32_t zng_compress(uint8_t *dest, size_t *destLen, const uint8_t *source, size_t sourceLen);
size_t  zng_compressBound(size_t sourceLen);
int32_t zng_uncompress(uint8_t *dest, size_t *destLen, const uint8_t *source, size_t sourceLen);
int32_t zng_uncompress2(uint8_t *dest, size_t *destLen, const uint8_t *source, size_t *sourceLen);

/* ---- checksums ---- */
uint32_t zng_crc32(uint32_t crc, const uint8_t *buf, uint32_t len);
uint32_t zng_crc32_z(uint32_t crc, const uint8_t *buf, size_t len);
uint32_t zng_adler32(uint32_t adler, const uint8_t *buf, uint32_t len);
uint32_t zng_adler32_z(uint32_t adler, const uint8_t *buf, size_t len);
uint32_t zng_crc32_combine(uint32_t crc1, uint32_t crc2, z_off64_t len2);
uint32_t zng_crc32_combine_gen(z_off64_t len2);
uint32_t zng_crc32_combine_op(uint32_t crc1, uint32_t crc2, uint32_t op);
uint32_t zng_adler32_combine(uint32_t adler1, uint32_t adler2, z_off64_t len2);

#ifdef __cplusplus
}
#endif
#endif /* ZNGLIB_B200_H */
