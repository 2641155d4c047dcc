/* zng_inflate.c -- zng_inflateInit2 / zng_inflate / zng_inflateReset / zng_inflateEnd and zng_uncompress(2) of
 * the host library, routed to the GPU member decoder (K4, csrc/inflate.cu) through the C-ABI.
 *
 * Reference interface and behaviour mirrored here:
 *   inflate.c:219-255    inflateInit2: windowBits decoding and range checks
 *   inflate.c:476-490    inflate(): state / argument checks
 *   inflate.c:1176-1200  return value: Z_STREAM_END, Z_DATA_ERROR (+ strm->msg), Z_NEED_DICT, Z_BUF_ERROR when a
 *                        Z_FINISH call cannot finish or a call makes no progress
 *   uncompr.c:20-80      zng_uncompress2 result mapping
 *
 * The GPU decodes whole members, so this layer keeps the bytes of the stream until the member is complete:
 * one zng_inflate(Z_FINISH) call on a complete stream -- the call the batched path replaces -- returns exactly
 * what the reference returns (code, msg, output, total_in, total_out, adler).  Incremental use (Z_NO_FLUSH with
 * partial input) is accepted: input is retained, output is delivered once the member has been decoded, and
 * next_in/avail_in are wound back to the end of the member so that concatenated members can follow.
 * Piecewise use goes through the resumable device decoder (zng_b200_inflate_stream_feed): every call decodes from the last deflate
 * block boundary reached so far to the last one inside the input that has arrived, hands the decoded bytes over at once and
 * drops the input behind them -- linear work, piecewise output, like the reference's resumable inflate() at block granularity.
 * Streams whose header that decoder does not parse itself (FDICT, FHCRC, malformed) keep the retained-member path, where a
 * decode attempt costs the whole retained input and is therefore repeated only when the input has grown by half, when the piece
 * fed is shorter than the one before, on Z_FINISH, or on a call without new input.
 */
#include "zng_host.h"
#include <stdlib.h>
#include <string.h>

enum { IN_RUN = 1, IN_DONE = 2, IN_BAD = 3, IN_DICT = 4 };

static int istate_check(zng_stream *strm) {
    if (strm == NULL || strm->zalloc == NULL || strm->zfree == NULL) return 1;
    struct internal_state *s = strm->state;
    if (s == NULL || s->strm != strm || s->kind != 'I') return 1;
    return 0;
}

int32_t zng_inflateReset(zng_stream *strm) {
    if (istate_check(strm)) return Z_STREAM_ERROR;
    struct internal_state *s = strm->state;
    strm->total_in = strm->total_out = 0;
    strm->msg = NULL;
    strm->adler = (uint32_t)(s->level & 1);                 /* inflate.c:123-124: wrap & 1 */
    s->status = IN_RUN;
    s->in_len = 0; s->pend_pos = s->pend_len = 0;
    s->finished = 0;
    s->check_len = 0;                                       /* bytes of the member's output already handed out */
    s->next_try = 0; s->last_piece = 0;
    if (s->res) { zng_b200_inflate_stream_close(s->res); s->res = NULL; }
    s->res_started = 0; s->no_resume = 0;
    return Z_OK;
}

int32_t zng_inflateInit2(zng_stream *strm, int32_t windowBits) {
    if (strm == NULL) return Z_STREAM_ERROR;
    strm->msg = NULL;
    if (strm->zalloc == NULL) { strm->zalloc = zng_host_default_alloc; strm->opaque = NULL; }
    if (strm->zfree == NULL) strm->zfree = zng_host_default_free;
    int wrap, wb = windowBits;                              /* inflate.c:232-251 */
    if (wb < 0) { if (wb < -MAX_WBITS) return Z_STREAM_ERROR; wrap = 0; wb = -wb; }
    else { wrap = (wb >> 4) + 5; if (wb < 48) wb &= MAX_WBITS; }
    if (wb && (wb < 8 || wb > MAX_WBITS)) return Z_STREAM_ERROR;
    if (zng_b200_thread_ctx() == NULL) { strm->msg = "no CUDA device"; return Z_MEM_ERROR; }
    struct internal_state *s = (struct internal_state *)strm->zalloc(strm->opaque, 1, sizeof(*s));
    if (s == NULL) return Z_MEM_ERROR;
    memset(s, 0, sizeof(*s));
    s->strm = strm; s->kind = 'I'; s->wrap = windowBits; s->level = wrap;
    strm->state = s;
    return zng_inflateReset(strm);
}

int32_t zng_inflateInit(zng_stream *strm) { return zng_inflateInit2(strm, MAX_WBITS); }

int32_t zng_inflateEnd(zng_stream *strm) {
    if (istate_check(strm)) return Z_STREAM_ERROR;
    struct internal_state *s = strm->state;
    free(s->in_buf); free(s->pend);
    if (s->res) zng_b200_inflate_stream_close(s->res);
    strm->zfree(strm->opaque, s);
    strm->state = NULL;
    return Z_OK;
}

static size_t deliver(zng_stream *strm) {
    struct internal_state *s = strm->state;
    size_t n = s->pend_len - s->pend_pos;
    if (n > strm->avail_out) n = strm->avail_out;
    if (n) {
        memcpy(strm->next_out, s->pend + s->pend_pos, n);
        strm->next_out += n; strm->avail_out -= (uint32_t)n; strm->total_out += n; s->pend_pos += n;
    }
    return n;
}

/* one stream through the GPU: returns the zng_inflate code of a Z_FINISH call with `cap` bytes of output room */
static int gpu_member(zng_stream *strm, const uint8_t *in, size_t n, uint8_t *out, size_t cap, uint32_t *out_len,
                      uint32_t *in_used, uint32_t *check, uint32_t *detail) {
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    if (!ctx) { strm->msg = "no CUDA device"; return Z_MEM_ERROR; }
    if (n > 0xffffffffu) n = 0xffffffffu;
    if (cap > 0xfffffff0u) cap = 0xfffffff0u;
    int32_t status = Z_BUF_ERROR;
    size_t ol = 0, iu = 0;
    uint8_t dummy_in = 0, dummy_out = 0;
    int r = zng_b200_inflate_stream_host(ctx, n ? in : &dummy_in, n, strm->state->wrap, cap ? out : &dummy_out, cap, &ol, &iu, check, &status, detail);
    if (r != ZNG_B200_OK) { strm->msg = zng_b200_last_error(ctx); return r == ZNG_B200_MEM_ERROR ? Z_MEM_ERROR : Z_STREAM_ERROR; }
    *out_len = (uint32_t)ol; *in_used = (uint32_t)iu;
    return status;
}

int32_t zng_inflate(zng_stream *strm, int32_t flush) {
    if (istate_check(strm) || strm->next_out == NULL || (strm->next_in == NULL && strm->avail_in != 0)) return Z_STREAM_ERROR;
    struct internal_state *s = strm->state;
    if (s->status == IN_BAD) return Z_DATA_ERROR;
    if (s->status == IN_DICT) return Z_NEED_DICT;           /* zng_inflateSetDictionary is outside the hot path */
    const size_t in0 = strm->avail_in, out0 = strm->avail_out;
    if (s->status == IN_DONE) {
        if (s->res_started) {                               /* output still pending in the resumable decoder */
            size_t iu = 0, ol = 0; int32_t st = 0;
            if (zng_b200_inflate_stream_feed(s->res, NULL, 0, strm->next_out, strm->avail_out, &iu, &ol, &st, NULL, NULL) != ZNG_B200_OK) return Z_MEM_ERROR;
            strm->next_out += ol; strm->avail_out -= (uint32_t)ol; strm->total_out += ol;
            if (st == 1) return Z_STREAM_END;
            return flush == Z_FINISH ? Z_BUF_ERROR : Z_OK;
        }
        deliver(strm);
        if (s->pend_pos == s->pend_len) return Z_STREAM_END;
        return flush == Z_FINISH ? Z_BUF_ERROR : Z_OK;
    }

    /* the member so far = retained bytes + this call's input */
    const uint8_t *src = strm->next_in; size_t n = strm->avail_in;
    if (s->in_len) {
        if (s->in_len + n > s->in_cap) {
            size_t cap = s->in_cap ? s->in_cap : 65536;
            while (cap < s->in_len + n) cap *= 2;
            uint8_t *p = (uint8_t *)realloc(s->in_buf, cap);
            if (!p) return Z_MEM_ERROR;
            s->in_buf = p; s->in_cap = cap;
        }
        if (n) memcpy(s->in_buf + s->in_len, src, n);
        src = s->in_buf; n += s->in_len;
    }
    const size_t kept = s->in_len;

    /* ---- piecewise use: the resumable decoder (block-boundary checkpoints on the device): linear work, output as it is decoded.
     * The one-shot call (a complete stream with Z_FINISH, nothing retained) keeps the path below, whose codes, messages and
     * total_in are the reference's to the byte. */
    if (!s->no_resume && (s->res_started || !(kept == 0 && flush == Z_FINISH))) {
        zng_b200_ctx *ctx = zng_b200_thread_ctx();
        if (!ctx) { strm->msg = "no CUDA device"; return Z_MEM_ERROR; }
        if (!s->res && zng_b200_inflate_stream_open(ctx, s->wrap, &s->res) != ZNG_B200_OK) return Z_MEM_ERROR;
        size_t iu = 0, ol = 0; int32_t st = 0; uint32_t det = 0, chk = 0;
        if (zng_b200_inflate_stream_feed(s->res, src, n, strm->next_out, strm->avail_out, &iu, &ol, &st, &det, &chk) != ZNG_B200_OK) {
            strm->msg = zng_b200_last_error(ctx); return Z_MEM_ERROR;
        }
        if (st == 2) {                                                        /* output is piling up: some was handed over, no input taken */
            strm->next_out += ol; strm->avail_out -= (uint32_t)ol; strm->total_out += ol;
            if (flush == Z_FINISH) return Z_BUF_ERROR;
            return ol ? Z_OK : Z_BUF_ERROR;
        }
        if (st == ZNG_B200_NOT_RESUMABLE) s->no_resume = 1;                   /* FDICT / FHCRC / header errors: the exact path below */
        else if (iu == 0 && n != 0 && st == 0 && !s->res_started) {
            /* the header is not complete yet: keep the bytes here and pass them again with the next piece */
            if (kept == 0) {
                if (n > s->in_cap) { uint8_t *p = (uint8_t *)realloc(s->in_buf, n); if (!p) return Z_MEM_ERROR; s->in_buf = p; s->in_cap = n; }
                memcpy(s->in_buf, src, n);
            }
            s->in_len = n;
            strm->next_in += in0; strm->total_in += in0; strm->avail_in = 0;
            return flush == Z_FINISH ? Z_BUF_ERROR : (in0 ? Z_OK : Z_BUF_ERROR);
        } else {
            s->res_started = 1; s->in_len = 0;
            strm->next_out += ol; strm->avail_out -= (uint32_t)ol; strm->total_out += ol;
            const size_t used_now = iu > kept ? iu - kept : 0;                  /* of this call's input */
            strm->next_in += used_now; strm->avail_in -= (uint32_t)used_now; strm->total_in += used_now;
            if (st == ZNG_B200_DATA_ERROR) {
                strm->msg = zng_b200_inflate_msg(det);
                strm->next_in += strm->avail_in; strm->total_in += strm->avail_in; strm->avail_in = 0;
                s->status = IN_BAD;
                return Z_DATA_ERROR;
            }
            if (st == 1 || (s->res && iu < n)) {                                /* the stream ended (maybe with output still pending) */
                strm->adler = chk;
                s->status = IN_DONE;
                if (st == 1) return Z_STREAM_END;
                return flush == Z_FINISH ? Z_BUF_ERROR : Z_OK;
            }
            if (flush == Z_FINISH) return Z_BUF_ERROR;                          /* incomplete stream, or no room for its output */
            return (used_now || ol) ? Z_OK : Z_BUF_ERROR;                       /* inflate.c:1197-1199: no progress */
        }
    }

    if (flush != Z_FINISH && in0 != 0 && s->next_try && n < s->next_try && in0 >= s->last_piece) {
        /* not due yet (see the header): keep the bytes, consume them, no device work */
        if (kept == 0) {
            if (n > s->in_cap) { uint8_t *p = (uint8_t *)realloc(s->in_buf, n); if (!p) return Z_MEM_ERROR; s->in_buf = p; s->in_cap = n; }
            memcpy(s->in_buf, src, n);
        }
        s->in_len = n; s->last_piece = in0;
        strm->next_in += in0; strm->total_in += in0; strm->avail_in = 0;
        return Z_OK;
    }

    uint32_t out_len = 0, in_used = 0, check = 0, detail = 0;
    int r;
    if (kept == 0 && flush == Z_FINISH) {
        /* the replaced call: decode straight into the caller's buffer */
        r = gpu_member(strm, src, n, strm->next_out, strm->avail_out, &out_len, &in_used, &check, &detail);
        if (r == Z_MEM_ERROR || r == Z_STREAM_ERROR) return r;
        if (r != Z_BUF_ERROR || !(detail & 0x100u)) {
            strm->next_out += out_len; strm->avail_out -= out_len; strm->total_out += out_len;
            s->check_len += out_len;
            goto finish;
        }
        /* output room ran out: fall through and decode the member into the library's own buffer */
    }
    {
        size_t cap = n * 4 + 65536;
        if (cap < strm->avail_out) cap = strm->avail_out;
        for (;;) {
            if (cap > s->pend_cap) {
                uint8_t *p = (uint8_t *)realloc(s->pend, cap);
                if (!p) return Z_MEM_ERROR;
                s->pend = p; s->pend_cap = cap;
            }
            r = gpu_member(strm, src, n, s->pend, s->pend_cap, &out_len, &in_used, &check, &detail);
            if (r == Z_MEM_ERROR || r == Z_STREAM_ERROR) return r;
            if (r == Z_BUF_ERROR && (detail & 0x100u) && cap < 0xfffffff0u) {
                if (detail & 0x400u) cap = (size_t)out_len + 64;                /* the decoder reported the size it needs */
                else cap = cap * 4 > 0xfffffff0u ? 0xfffffff0u : cap * 4;
                continue;
            }
            break;
        }
        s->pend_pos = 0; s->pend_len = 0;
        if (r == Z_BUF_ERROR && flush != Z_FINISH) {
            /* incomplete member: keep the input, nothing is delivered yet */
            if (kept == 0 && n) {
                if (n > s->in_cap) { uint8_t *p = (uint8_t *)realloc(s->in_buf, n); if (!p) return Z_MEM_ERROR; s->in_buf = p; s->in_cap = n; }
                memcpy(s->in_buf, src, n);
            }
            s->in_len = n;
            s->next_try = n + n / 2 + 1; s->last_piece = in0;
            strm->next_in += in0; strm->total_in += in0; strm->avail_in = 0;
            return in0 ? Z_OK : Z_BUF_ERROR;                /* inflate.c:1197-1199: no progress */
        }
        s->pend_len = out_len;
        s->pend_pos = s->check_len < out_len ? (size_t)s->check_len : out_len;   /* a retried member: skip what went out before */
        s->check_len += deliver(strm);
    }
finish:
    if (r == Z_STREAM_END || r == Z_NEED_DICT) {
        /* wind next_in back to the end of the member */
        const size_t used_now = in_used > kept ? in_used - kept : 0;
        strm->next_in += used_now; strm->avail_in -= (uint32_t)used_now; strm->total_in += used_now;
        s->in_len = 0; s->next_try = 0; s->last_piece = 0;
        strm->adler = check;
        if (r == Z_NEED_DICT) { s->status = IN_DICT; return Z_NEED_DICT; }
        s->status = IN_DONE;
        if (s->pend_pos == s->pend_len) return Z_STREAM_END;
        return flush == Z_FINISH ? Z_BUF_ERROR : Z_OK;
    }
    if (r == Z_DATA_ERROR) {
        strm->msg = zng_b200_inflate_msg(detail);
        strm->next_in += in0; strm->total_in += in0; strm->avail_in = 0;
        s->status = IN_BAD;
        return Z_DATA_ERROR;
    }
    /* Z_BUF_ERROR under Z_FINISH: the member is incomplete, or the caller's buffer is full */
    if (kept == 0 && n) {
        if (n > s->in_cap) { uint8_t *p = (uint8_t *)realloc(s->in_buf, n); if (!p) return Z_MEM_ERROR; s->in_buf = p; s->in_cap = n; }
        memcpy(s->in_buf, src, n);
    }
    s->in_len = n;
    strm->next_in += in0; strm->total_in += in0; strm->avail_in = 0;
    (void)out0;
    return Z_BUF_ERROR;
}

/* ---- one-shot wrappers (uncompr.c:20-80 of the reference) ---- */
int32_t zng_uncompress2(uint8_t *dest, size_t *destLen, const uint8_t *source, size_t *sourceLen) {
    zng_stream strm;
    uint8_t one[1] = {0};
    size_t len = *sourceLen, left;
    memset(&strm, 0, sizeof(strm));
    if (*destLen) { left = *destLen; *destLen = 0; } else { left = 1; dest = one; }   /* uncompr.c:33-39 */
    int err = zng_inflateInit(&strm);
    if (err != Z_OK) return err;
    strm.next_in = source; strm.avail_in = len > 0xffffffffu ? 0xffffffffu : (uint32_t)len;
    strm.next_out = dest; strm.avail_out = left > 0xffffffffu ? 0xffffffffu : (uint32_t)left;
    len -= strm.avail_in; left -= strm.avail_out;
    err = zng_inflate(&strm, Z_FINISH);
    *sourceLen -= len + strm.avail_in;
    if (dest != one) *destLen = strm.total_out;
    else if (strm.total_out && err == Z_BUF_ERROR) left = 1;
    zng_inflateEnd(&strm);
    return err == Z_STREAM_END ? Z_OK : err == Z_NEED_DICT ? Z_DATA_ERROR : (err == Z_BUF_ERROR && left + strm.avail_out) ? Z_DATA_ERROR : err;
}

int32_t zng_uncompress(uint8_t *dest, size_t *destLen, const uint8_t *source, size_t sourceLen) {
    return zng_uncompress2(dest, destLen, source, &sourceLen);
}
