/* oracle/ref_ops.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference's OPERATOR surface, one operator per call, for the parity tests of include/zng_b200.h's operator table:
 * every function here sets up a real deflate_state of the unmodified reference (zng_deflateInit2) and calls the reference's own
 * dispatched variant through its functable (functable.h:26-42, FUNCTABLE_CALL) or its hash callbacks (deflate.h:121-131).
 * Compiled with the reference's internal headers where they lie under /root/reference (oracle/Makefile, same flags as the
 * reference objects) into oracle/_ref/libzng_ref.so; nothing of the reference is copied.
 */
#include "zbuild.h"
#include "deflate.h"
#include "functable.h"
#include "crc32.h"
#include <string.h>

#define REFOPS_EXPORT __attribute__((visibility("default")))

static deflate_state *make_state(zng_stream *strm, int level) {
    memset(strm, 0, sizeof(*strm));
    if (zng_deflateInit2(strm, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return NULL;
    return (deflate_state *)strm->state;
}

/* functable.longest_match (match_tpl.h:26-280) with configuration_table[level]'s parameters: window[0..n) + zeros behind it,
 * prev[32768], strstart, cur_match; lookahead = n - strstart.  Returns the length; *match_start = s->match_start. */
REFOPS_EXPORT uint32_t refops_longest_match(const uint8_t *window, uint32_t n, const uint16_t *prev, uint32_t strstart,
                                            uint32_t cur_match, int level, uint32_t *match_start) {
    zng_stream strm;
    deflate_state *s = make_state(&strm, level);
    if (!s || n > 2u * s->w_size) return 0xffffffffu;
    memset(s->window, 0, 2u * s->w_size);
    memcpy(s->window, window, n);
    memcpy(s->prev, prev, s->w_size * sizeof(Pos));
    s->strstart = strstart; s->lookahead = n - strstart; s->prev_length = 0; s->match_start = 0;
    uint32_t r = FUNCTABLE_CALL(longest_match)(s, (Pos)cur_match);
    if (match_start) *match_start = s->match_start;
    zng_deflateEnd(&strm);
    return r;
}

/* insert_string(s, str, count) / quick_insert_string(s, str) (insert_string_tpl.h:58-104) on caller-supplied head[65536] /
 * prev[32768], updated in place.  count == 0: quick_insert_string, whose return value (the replaced head) is returned. */
REFOPS_EXPORT uint32_t refops_insert_string(const uint8_t *window, uint32_t n, uint16_t *head, uint16_t *prev, uint32_t str, uint32_t count) {
    zng_stream strm;
    deflate_state *s = make_state(&strm, 2);
    if (!s || n > 2u * s->w_size) return 0xffffffffu;
    memset(s->window, 0, 2u * s->w_size);
    memcpy(s->window, window, n);
    memcpy(s->head, head, HASH_SIZE * sizeof(Pos));
    memcpy(s->prev, prev, s->w_size * sizeof(Pos));
    uint32_t r = 0;
    if (count) insert_string(s, str, count); else r = quick_insert_string(s, str);
    memcpy(head, s->head, HASH_SIZE * sizeof(Pos));
    memcpy(prev, s->prev, s->w_size * sizeof(Pos));
    zng_deflateEnd(&strm);
    return r;
}

/* functable.slide_hash (slide_hash_c.c:15-52 and its SIMD twins) */
REFOPS_EXPORT int refops_slide_hash(uint16_t *head, uint16_t *prev) {
    zng_stream strm;
    deflate_state *s = make_state(&strm, 2);
    if (!s) return -1;
    memcpy(s->head, head, HASH_SIZE * sizeof(Pos));
    memcpy(s->prev, prev, s->w_size * sizeof(Pos));
    FUNCTABLE_CALL(slide_hash)(s);
    memcpy(head, s->head, HASH_SIZE * sizeof(Pos));
    memcpy(prev, s->prev, s->w_size * sizeof(Pos));
    zng_deflateEnd(&strm);
    return 0;
}

REFOPS_EXPORT uint32_t refops_update_hash(uint32_t h, uint32_t val) { return update_hash(h, val); }
REFOPS_EXPORT uint32_t refops_compare256(const uint8_t *a, const uint8_t *b) { FUNCTABLE_INIT; return FUNCTABLE_CALL(compare256)(a, b); }
REFOPS_EXPORT uint32_t refops_chunksize(void) { FUNCTABLE_INIT; return FUNCTABLE_CALL(chunksize)(); }

/* functable.chunkmemset_safe(out, from = out - dist, len, left) on buf (pos = offset of out); returns the advance of out */
REFOPS_EXPORT uint32_t refops_chunkmemset_safe(uint8_t *buf, uint32_t pos, uint32_t dist, uint32_t len, uint32_t left) {
    FUNCTABLE_INIT;
    uint8_t *r = FUNCTABLE_CALL(chunkmemset_safe)(buf + pos, buf + pos - dist, len, left);
    return (uint32_t)(r - (buf + pos));
}

/* functable.crc32_fold_reset / crc32_fold(_copy) over `pieces` equal pieces / crc32_fold_final; dst may be NULL (no copy) */
REFOPS_EXPORT uint32_t refops_crc32_fold(const uint8_t *src, size_t n, size_t piece, uint8_t *dst) {
    FUNCTABLE_INIT;
    crc32_fold f;
    memset(&f, 0, sizeof(f));
    FUNCTABLE_CALL(crc32_fold_reset)(&f);
    for (size_t o = 0; o < n; o += piece) {
        size_t k = n - o < piece ? n - o : piece;
        if (dst) FUNCTABLE_CALL(crc32_fold_copy)(&f, dst + o, src + o, k);
        else FUNCTABLE_CALL(crc32_fold)(&f, src + o, k, 0);
    }
    return FUNCTABLE_CALL(crc32_fold_final)(&f);
}

REFOPS_EXPORT uint32_t refops_adler32_fold_copy(uint32_t adler, uint8_t *dst, const uint8_t *src, size_t n) {
    FUNCTABLE_INIT;
    return FUNCTABLE_CALL(adler32_fold_copy)(adler, dst, src, n);
}
