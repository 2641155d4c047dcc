/* zng_host.h -- internals shared by the C11 host library files. */
#ifndef ZNG_HOST_H
#define ZNG_HOST_H
#include "zlib-ng.h"
#include "zng_b200.h"

#define ZNG_CHUNK 65536u

/* zng_ctx.c: the calling thread's GPU context (created on first use; NULL when there is no device) */
zng_b200_ctx *zng_b200_thread_ctx(void);

/* library-owned stream state (zng_stream.state); `kind` tells deflate ('D') from inflate ('I') states the
 * way the reference's deflateStateCheck / inflateStateCheck do (deflate.c:580-594, inflate.c:94-103) */
struct internal_state {
    zng_stream *strm;          /* back pointer */
    int kind;
    int status;
    int wrap;                  /* deflate: 0 raw, 1 zlib, 2 gzip;  inflate: the windowBits given to inflateInit2 */
    int level;
    int last_flush;
    /* input not yet processed (Z_NO_FLUSH buffering, tails < 64 KiB; inflate: the stream so far) */
    uint8_t *in_buf; size_t in_len, in_cap;
    /* produced bytes not yet handed to the caller */
    uint8_t *pend; size_t pend_pos, pend_len, pend_cap;
    uint32_t check;            /* running adler32 (zlib) / crc32 (gzip) of all consumed input */
    uint64_t check_len;
    int header_done, trailer_done, finished;
    /* dependent-chunk mode (zng_deflateSetDictionary on a raw level-1 stream): the 32768 bytes in front of the next piece */
    uint8_t *dict; int have_dict;
    /* inflate, incremental use: retained bytes at which the next decode attempt is due, and the size of the last piece fed */
    size_t next_try, last_piece;
    /* inflate, incremental use: the resumable device decoder (zng_b200_inflate_stream_*); no_resume: its header check sent the stream to the one-shot path */
    zng_b200_inflate_stream *res; int res_started, no_resume;
    int after_sync;            /* a Z_SYNC_FLUSH closed the (only) piece of a stream without dictionary: nothing may follow but Reset */
};

void *zng_host_default_alloc(void *opaque, unsigned items, unsigned size);
void  zng_host_default_free(void *opaque, void *p);

/* host GF(2) / modular arithmetic of the combine functions (zng_checksum.c) */
uint32_t zng_host_multmodp(uint32_t a, uint32_t b);
uint32_t zng_host_x2nmodp(int64_t n, unsigned k);

#endif
