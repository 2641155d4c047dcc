/* zng_deflate.c -- zng_deflateInit2 / zng_deflate / zng_deflateReset / zng_deflateEnd / zng_deflateBound
 * of the host library, routed to the GPU chunk compressor through the C-ABI.
 *
 * Reference interface and behaviour mirrored here:
 *   deflate.c:283-419   deflateInit2: argument validation order, wrap selection from windowBits
 *   deflate.c:819-1113  deflate(): argument / state checks (:823-863), zlib header (:868-900), gzip
 *                       header (:902-921), flush handling (:1061-1083), trailers (:1089-1112)
 *   deflate.c:534-578   deflateReset(Keep), :709-781 deflateBound, :1116-1126 deflateEnd
 *
 * What is different by design (include/zlib-ng.h states the contract): the strategy call of the
 * reference (deflate.c:1036-1043) is replaced -- like DFLTCC's DEFLATE_HOOK (deflate.c:1039) -- by
 * one GPU launch over all complete 65536-byte pieces that are available, each compressed as the
 * reference compresses a piece fed with Z_FULL_FLUSH (last piece with Z_FINISH when finishing).
 * Z_NO_FLUSH input is only buffered.  Output that does not fit avail_out is kept pending and handed
 * out by later calls, exactly like the reference's pending_buf.
 */
#include "zng_host.h"
#include <stdlib.h>
#include <string.h>

enum { ST_INIT = 1, ST_BUSY = 2, ST_FINISH = 3 };

void *zng_host_default_alloc(void *opaque, unsigned items, unsigned size) { (void)opaque; return malloc((size_t)items * size); }
void zng_host_default_free(void *opaque, void *p) { (void)opaque; free(p); }
#define default_alloc zng_host_default_alloc
#define default_free zng_host_default_free

static int state_check(zng_stream *strm, int kind) {
    if (strm == NULL || strm->zalloc == NULL || strm->zfree == NULL) return 1;
    struct internal_state *s = strm->state;
    if (s == NULL || s->strm != strm || s->kind != kind) return 1;
    return 0;
}

const char *zlibng_version(void) { return ZLIBNG_VERSION; }

int32_t zng_deflateInit2(zng_stream *strm, int32_t level, int32_t method, int32_t windowBits, int32_t memLevel, int32_t strategy) {
    int wrap = 1;
    if (strm == NULL) return Z_STREAM_ERROR;
    strm->msg = NULL;
    if (strm->zalloc == NULL) { strm->zalloc = default_alloc; strm->opaque = NULL; }
    if (strm->zfree == NULL) strm->zfree = default_free;
    if (level == Z_DEFAULT_COMPRESSION) level = 6;
    if (windowBits < 0) {
        wrap = 0;
        if (windowBits < -MAX_WBITS) return Z_STREAM_ERROR;
        windowBits = -windowBits;
    } else if (windowBits > MAX_WBITS) {
        wrap = 2;
        windowBits -= 16;
    }
    /* the reference's own range checks (deflate.c:316-321) ... */
    if (memLevel < 1 || memLevel > 9 || method != Z_DEFLATED || windowBits < 8 || windowBits > MAX_WBITS ||
        level < 0 || level > 9 || strategy < 0 || strategy > Z_FIXED || (windowBits == 8 && wrap != 1))
        return Z_STREAM_ERROR;
    /* ... then the frozen parameter set of the GPU path (no CPU fallback) */
    if (level < 1 || level > 6 || windowBits != 15 || memLevel != DEF_MEM_LEVEL || strategy != Z_DEFAULT_STRATEGY) {
        strm->msg = "unsupported parameters: levels 1-6, windowBits 15, memLevel 8, default strategy only";
        return Z_STREAM_ERROR;
    }
    if (zng_b200_thread_ctx() == NULL) { strm->msg = "no CUDA device"; return Z_MEM_ERROR; }
    struct internal_state *s = (struct internal_state *)strm->zalloc(strm->opaque, 1, sizeof(*s));
    if (s == NULL) return Z_MEM_ERROR;
    memset(s, 0, sizeof(*s));
    s->strm = strm; s->kind = 'D'; s->wrap = wrap; s->level = level;
    strm->state = s;
    return zng_deflateReset(strm);
}

int32_t zng_deflateInit(zng_stream *strm, int32_t level) {
    return zng_deflateInit2(strm, level, Z_DEFLATED, MAX_WBITS, DEF_MEM_LEVEL, Z_DEFAULT_STRATEGY);
}

int32_t zng_deflateReset(zng_stream *strm) {
    if (state_check(strm, 'D')) return Z_STREAM_ERROR;
    struct internal_state *s = strm->state;
    strm->total_in = strm->total_out = 0;
    strm->msg = NULL;
    strm->data_type = Z_UNKNOWN;
    s->in_len = 0; s->pend_pos = s->pend_len = 0;
    s->status = s->wrap ? ST_INIT : ST_BUSY;
    strm->adler = s->wrap == 2 ? 0u : 1u;                 /* deflate.c:556-561 */
    s->check = strm->adler; s->check_len = 0;
    s->last_flush = -2;
    s->header_done = s->trailer_done = s->finished = 0;
    s->have_dict = 0; s->after_sync = 0;
    return Z_OK;
}

int32_t zng_deflateEnd(zng_stream *strm) {
    if (state_check(strm, 'D')) return Z_STREAM_ERROR;
    struct internal_state *s = strm->state;
    int busy = (s->status == ST_BUSY && (s->in_len || s->pend_len > s->pend_pos));
    free(s->in_buf); free(s->pend); free(s->dict);
    strm->zfree(strm->opaque, s);
    strm->state = NULL;
    return busy ? Z_DATA_ERROR : Z_OK;                     /* deflate.c:1125 */
}

/* deflate.c:456-512.  Supported where the GPU path reproduces the reference bit for bit: a raw (windowBits -15) stream of
 * level 1..6 with no input pending and a dictionary of at least one window (the last 32768 bytes are used, :479-488).  The
 * stream then works in pigz's dependent mode: each 65536-byte piece is compressed exactly as a FRESH stream primed with
 * the 32768 bytes in front of it does (zng_deflateSetDictionary + one zng_deflate per piece), Z_SYNC_FLUSH allowed. */
int32_t zng_deflateSetDictionary(zng_stream *strm, const uint8_t *dictionary, uint32_t dictLength) {
    if (state_check(strm, 'D') || dictionary == NULL) return Z_STREAM_ERROR;
    struct internal_state *s = strm->state;
    if (s->wrap == 2 || (s->wrap == 1 && s->status != ST_INIT) || s->in_len) return Z_STREAM_ERROR;     /* deflate.c:468-469 */
    if (s->wrap != 0 || s->level < 1 || s->level > 6 || dictLength < 32768u) {
        strm->msg = "unsupported: preset dictionaries on raw streams of level 1..6, 32768 bytes or more"; return Z_STREAM_ERROR;
    }
    if (!s->dict && !(s->dict = (uint8_t *)malloc(32768))) return Z_MEM_ERROR;
    memcpy(s->dict, dictionary + (dictLength - 32768u), 32768);
    s->have_dict = 1;
    return Z_OK;
}

unsigned long zng_deflateBound(zng_stream *strm, unsigned long sourceLen) {
    /* chunked format: every 65536-byte piece costs at most 9 bits per byte + 3-bit header + 7-bit EOB +
     * the 5-byte flush marker (deflate.c:771-777 gives the same 9-bit bound for the quick strategy) */
    unsigned long pieces = sourceLen / ZNG_CHUNK + 1;
    unsigned long b = sourceLen + (sourceLen >> 3) + pieces * 8 + 16;
    if (strm && strm->state && strm->state->kind == 'D') b += strm->state->wrap == 2 ? 18 : (strm->state->wrap == 1 ? 6 : 0);
    else b += 18;
    return b;
}

static int pend_reserve(struct internal_state *s, size_t extra) {
    if (s->pend_pos == s->pend_len) s->pend_pos = s->pend_len = 0;
    if (s->pend_len + extra <= s->pend_cap) return 0;
    size_t cap = s->pend_cap ? s->pend_cap : 65536;
    while (cap < s->pend_len + extra) cap *= 2;
    uint8_t *p = (uint8_t *)realloc(s->pend, cap);
    if (!p) return -1;
    s->pend = p; s->pend_cap = cap;
    return 0;
}
static int pend_put(struct internal_state *s, const uint8_t *b, size_t n) {
    if (pend_reserve(s, n)) return -1;
    memcpy(s->pend + s->pend_len, b, n); s->pend_len += n;
    return 0;
}
static void pend_flush(zng_stream *strm) {                 /* flush_pending, deflate.c:789-807 */
    struct internal_state *s = strm->state;
    size_t n = s->pend_len - s->pend_pos;
    if (n > strm->avail_out) n = strm->avail_out;
    if (n == 0) return;
    memcpy(strm->next_out, s->pend + s->pend_pos, n);
    strm->next_out += n; strm->avail_out -= (uint32_t)n; strm->total_out += n; s->pend_pos += n;
}

static int in_append(struct internal_state *s, const uint8_t *b, size_t n) {
    if (s->in_len + n > s->in_cap) {
        size_t cap = s->in_cap ? s->in_cap : (size_t)1 << 20;
        while (cap < s->in_len + n) cap *= 2;
        uint8_t *p = (uint8_t *)realloc(s->in_buf, cap);
        if (!p) return -1;
        s->in_buf = p; s->in_cap = cap;
    }
    memcpy(s->in_buf + s->in_len, b, n); s->in_len += n;
    return 0;
}

/* compress [src, src+n) into the pending buffer; fin: Z_FINISH on the last piece.  When nothing is pending and the
 * caller's buffer can take the worst case, the bytes go straight to next_out (no staging copy; with a pinned next_out the
 * device writes into it directly). */
static int compress_into_pending(zng_stream *strm, const uint8_t *src, size_t n, int fin) {
    struct internal_state *s = strm->state;
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    if (!ctx) { strm->msg = "no CUDA device"; return Z_MEM_ERROR; }
    size_t cap = n + (n >> 3) + (n / ZNG_CHUNK + 1) * 8 + 16;   /* 9 bits per byte + 8 bytes per piece: what zng_deflateBound promises */
    size_t out_len = 0; uint32_t crc = 0, adler = 1;
    size_t queued = s->pend_len - s->pend_pos;
    int direct = (size_t)strm->avail_out >= queued + cap;
    int r;
    uint8_t *dst; size_t room;
    if (direct) {
        pend_flush(strm);                                   /* header bytes first */
        dst = strm->next_out; room = strm->avail_out;
    } else {
        if (pend_reserve(s, cap)) return Z_MEM_ERROR;
        dst = s->pend + s->pend_len; room = s->pend_cap - s->pend_len;
    }
    if (s->have_dict) {
        r = zng_b200_deflate_host_primed_level(ctx, s->dict, src, n, s->level, fin, dst, room, &out_len, &crc, &adler);
        if (n >= 32768) memcpy(s->dict, src + n - 32768, 32768);            /* the window in front of the next piece */
        else if (n) { memmove(s->dict, s->dict + n, 32768 - n); memcpy(s->dict + 32768 - n, src, n); }
    } else {
        r = zng_b200_deflate_host(ctx, src, n, ZNG_CHUNK, s->level, fin, dst, room, &out_len, &crc, &adler);
    }
    if (r != ZNG_B200_OK) { strm->msg = zng_b200_last_error(ctx); return r == ZNG_B200_BUF_ERROR ? Z_BUF_ERROR : (r == ZNG_B200_MEM_ERROR ? Z_MEM_ERROR : Z_STREAM_ERROR); }
    if (direct) { strm->next_out += out_len; strm->avail_out -= (uint32_t)out_len; strm->total_out += out_len; }
    else s->pend_len += out_len;
    if (s->wrap == 2) s->check = zng_crc32_combine(s->check, crc, (z_off64_t)n);
    else if (s->wrap == 1) s->check = zng_adler32_combine(s->check, adler, (z_off64_t)n);
    s->check_len += n;
    if (s->wrap) strm->adler = s->check;
    return Z_OK;
}

int32_t zng_deflate(zng_stream *strm, int32_t flush) {
    if (state_check(strm, 'D') || flush > Z_BLOCK || flush < 0) return Z_STREAM_ERROR;
    struct internal_state *s = strm->state;
    if (strm->next_out == NULL || (strm->avail_in != 0 && strm->next_in == NULL) || (s->status == ST_FINISH && flush != Z_FINISH)) {
        strm->msg = "stream error"; return Z_STREAM_ERROR;
    }
    if (strm->avail_out == 0) { strm->msg = "buffer error"; return Z_BUF_ERROR; }
    if (s->after_sync && strm->avail_in) {
        strm->msg = "unsupported: input after Z_SYNC_FLUSH continues the reference's window; call zng_deflateReset (+ zng_deflateSetDictionary)";
        return Z_STREAM_ERROR;
    }
    if (flush == Z_SYNC_FLUSH && s->have_dict) flush = Z_FULL_FLUSH;   /* dependent mode: the pieces are joined by sync-flush markers */
    else if (flush == Z_SYNC_FLUSH && strm->total_in == 0 && s->in_len == 0 && strm->avail_in <= ZNG_CHUNK) {
        /* pigz's first chunk: one piece on a fresh stream; the bytes of a sync flush and of a full flush are the same
         * (deflate.c:1061-1083), only what may follow differs */
        flush = Z_FULL_FLUSH; s->after_sync = 1;
    }
    if (flush != Z_NO_FLUSH && flush != Z_FULL_FLUSH && flush != Z_FINISH) {
        /* Z_SYNC/PARTIAL_FLUSH/Z_BLOCK keep the window across the flush in the reference: not a chunk boundary */
        strm->msg = "unsupported flush mode: Z_NO_FLUSH, Z_FULL_FLUSH, Z_FINISH only"; return Z_STREAM_ERROR;
    }
    int old_flush = s->last_flush;
    s->last_flush = flush;

    if (s->pend_len > s->pend_pos) {                        /* deflate.c:838-850 */
        pend_flush(strm);
        if (strm->avail_out == 0) { s->last_flush = -1; return Z_OK; }
    } else if (strm->avail_in == 0 && flush <= old_flush && flush != Z_FINISH) {
        strm->msg = "buffer error"; return Z_BUF_ERROR;    /* deflate.c:855-857 (RANK is monotonic on the modes kept) */
    }
    if (s->status == ST_FINISH && strm->avail_in != 0) { strm->msg = "buffer error"; return Z_BUF_ERROR; }

    /* header (deflate.c:868-921) */
    if (!s->header_done) {
        if (s->wrap == 1) {
            unsigned header = (Z_DEFLATED + ((15 - 8) << 4)) << 8;
            unsigned level_flags = s->level < 2 ? 0 : (s->level < 6 ? 1 : 2);   /* deflate.c:873-880 */
            header |= level_flags << 6;
            header += 31 - (header % 31);
            uint8_t h[2] = {(uint8_t)(header >> 8), (uint8_t)header};
            if (pend_put(s, h, 2)) return Z_MEM_ERROR;
        } else if (s->wrap == 2) {
            uint8_t h[10] = {31, 139, 8, 0, 0, 0, 0, 0, (uint8_t)(s->level < 2 ? 4 : 0), 3 /* OS_CODE unix */};
            if (pend_put(s, h, 10)) return Z_MEM_ERROR;
        }
        s->header_done = 1; s->status = ST_BUSY;
    }

    if (!s->finished) {
        const uint8_t *src = strm->next_in; size_t n = strm->avail_in;
        if (flush == Z_NO_FLUSH) {
            /* buffer only; complete pieces are compressed once a flush point or enough data arrives */
            if (n) { if (in_append(s, src, n)) return Z_MEM_ERROR; strm->next_in += n; strm->total_in += n; strm->avail_in = 0; }
            size_t whole = s->in_len - (s->in_len % ZNG_CHUNK);
            if (whole >= ((size_t)256 << 20)) {
                int r = compress_into_pending(strm, s->in_buf, whole, 0);
                if (r != Z_OK) return r;
                memmove(s->in_buf, s->in_buf + whole, s->in_len - whole); s->in_len -= whole;
            }
        } else {
            int fin = (flush == Z_FINISH);
            const size_t kept = s->in_len;                  /* buffered before this call */
            if (s->in_len) {                                /* join buffered bytes with the new input */
                if (n) { if (in_append(s, src, n)) return Z_MEM_ERROR; }
                src = s->in_buf; n = s->in_len;
            }
            size_t took = strm->avail_in;
            int r = Z_OK;
            if (n || fin) r = compress_into_pending(strm, src, n, fin);
            else {                                          /* a flush with nothing to compress: the bare marker (trees.c:592-609) */
                static const uint8_t marker[5] = {0, 0, 0, 0xff, 0xff};
                if (pend_put(s, marker, 5)) return Z_MEM_ERROR;
            }
            if (r != Z_OK) {                                /* nothing was consumed: a retry must not see this call's input twice */
                s->in_len = kept;
                return r;
            }
            strm->next_in += took; strm->total_in += took; strm->avail_in = 0; s->in_len = 0;
            if (fin) {
                s->finished = 1; s->status = ST_FINISH;
                if (s->wrap == 1) {                         /* deflate.c:1099-1100 */
                    uint8_t t[4] = {(uint8_t)(s->check >> 24), (uint8_t)(s->check >> 16), (uint8_t)(s->check >> 8), (uint8_t)s->check};
                    if (pend_put(s, t, 4)) return Z_MEM_ERROR;
                } else if (s->wrap == 2) {                  /* deflate.c:1091-1096 */
                    uint32_t c = s->check, l = (uint32_t)s->check_len;
                    uint8_t t[8] = {(uint8_t)c, (uint8_t)(c >> 8), (uint8_t)(c >> 16), (uint8_t)(c >> 24), (uint8_t)l, (uint8_t)(l >> 8), (uint8_t)(l >> 16), (uint8_t)(l >> 24)};
                    if (pend_put(s, t, 8)) return Z_MEM_ERROR;
                }
            }
        }
    }
    pend_flush(strm);
    if (s->pend_len > s->pend_pos) { s->last_flush = -1; return Z_OK; }
    if (flush == Z_FINISH && s->finished) return Z_STREAM_END;
    return Z_OK;
}

/* ---- one-shot wrappers (compress.c:28-98 of the reference) ---- */
size_t zng_compressBound(size_t sourceLen) { return (size_t)zng_deflateBound(NULL, (unsigned long)sourceLen); }

int32_t zng_compress2(uint8_t *dest, size_t *destLen, const uint8_t *source, size_t sourceLen, int32_t level) {
    zng_stream strm;
    memset(&strm, 0, sizeof(strm));
    if (level == Z_DEFAULT_COMPRESSION) level = 6;
    int err = zng_deflateInit(&strm, level);
    if (err != Z_OK) return err;
    size_t left = *destLen; *destLen = 0;
    strm.next_out = dest; strm.next_in = source;
    do {                                                    /* uInt-sized windows like compress.c:52-63 */
        if (strm.avail_out == 0) { strm.avail_out = left > 0xffffffffu ? 0xffffffffu : (uint32_t)left; left -= strm.avail_out; }
        if (strm.avail_in == 0) { strm.avail_in = sourceLen > 0x40000000u ? 0x40000000u : (uint32_t)sourceLen; sourceLen -= strm.avail_in; }
        err = zng_deflate(&strm, sourceLen ? Z_NO_FLUSH : Z_FINISH);
    } while (err == Z_OK);
    *destLen = strm.total_out;
    zng_deflateEnd(&strm);
    return err == Z_STREAM_END ? Z_OK : err;
}

int32_t zng_compress(uint8_t *dest, size_t *destLen, const uint8_t *source, size_t sourceLen) {
    return zng_compress2(dest, destLen, source, sourceLen, Z_DEFAULT_COMPRESSION);
}
