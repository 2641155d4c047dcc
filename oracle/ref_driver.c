/* oracle/ref_driver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A pthread driver around the UNMODIFIED reference's public API (zlib-ng 2.2.2,
 * native zng_ names).  It is compiled together with the reference sources into
 * oracle/_ref/libzng_ref.so by oracle/Makefile and is used (a) to pin the C
 * restatement in oracle/zo_*.c, (b) to generate tests/golden/ fixtures and
 * (c) as the timed CPU baseline ("kind": "reference") in bench.py.
 *
 * Call sequences are the ones SURVEY.md section 8(c) fixes:
 *   deflate chunk : zng_deflateInit2(level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) once per
 *                   worker; per chunk zng_deflateReset + one zng_deflate(Z_FULL_FLUSH)
 *   gzip member   : zng_deflateInit2(level, Z_DEFLATED, 31, 8, 0) + one zng_deflate(Z_FINISH)
 *   inflate member: zng_inflateInit2(31) once per worker; zng_inflateReset + zng_inflate(Z_FINISH)
 *   primed chunk  : (pigz's default mode) a FRESH zng_deflateInit2(level, Z_DEFLATED, -15, 8, 0) per chunk,
 *                   zng_deflateSetDictionary(the 32768 stream bytes in front of the chunk), one zng_deflate(flush)
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "zlib-ng.h"

#define REFDRV_EXPORT __attribute__((visibility("default")))

typedef struct {
    /* common */
    int kind;               /* 0 deflate chunks, 1 checksum chunks, 2 inflate members, 3 gzip members, 4 primed chunks */
    atomic_size_t next;     /* next unit to claim */
    size_t n_units;
    atomic_int err;
    /* deflate / checksum */
    const uint8_t *in;
    size_t n;
    uint32_t chunk;
    int level;
    int flush;
    uint8_t *out;
    size_t out_stride;
    uint32_t *sizes;
    uint32_t *crcs;
    uint32_t *adlers;
    /* members */
    const uint64_t *in_off;   /* n_units+1 offsets into in */
    const uint64_t *out_off;  /* n_units+1 offsets into out (capacity per unit) */
    int32_t *status;
} job_t;

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    zng_stream s;
    memset(&s, 0, sizeof(s));
    int inited = 0;
    for (;;) {
        size_t u = atomic_fetch_add(&j->next, 1);
        if (u >= j->n_units) break;
        if (j->kind == 0) {
            size_t off = u * (size_t)j->chunk;
            uint32_t len = (uint32_t)((j->n - off < j->chunk) ? (j->n - off) : j->chunk);
            if (!inited) {
                if (zng_deflateInit2(&s, j->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { atomic_store(&j->err, 1); break; }
                inited = 1;
            } else {
                zng_deflateReset(&s);
            }
            s.next_in = j->in + off; s.avail_in = len;
            s.next_out = j->out + u * j->out_stride; s.avail_out = (uint32_t)j->out_stride;
            int r = zng_deflate(&s, j->flush);
            if ((j->flush == Z_FINISH ? r != Z_STREAM_END : r != Z_OK) || s.avail_in != 0) atomic_store(&j->err, 2);
            j->sizes[u] = (uint32_t)s.total_out;
            if (j->crcs) j->crcs[u] = zng_crc32(0, j->in + off, len);
            if (j->adlers) j->adlers[u] = zng_adler32(1, j->in + off, len);
        } else if (j->kind == 4) {
            size_t off = u * (size_t)j->chunk;
            uint32_t len = (uint32_t)((j->n - off < j->chunk) ? (j->n - off) : j->chunk);
            zng_stream d;
            memset(&d, 0, sizeof(d));
            if (zng_deflateInit2(&d, j->level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { atomic_store(&j->err, 1); break; }
            if (off >= 32768 && zng_deflateSetDictionary(&d, j->in + off - 32768, 32768) != Z_OK) atomic_store(&j->err, 3);
            d.next_in = j->in + off; d.avail_in = len;
            d.next_out = j->out + u * j->out_stride; d.avail_out = (uint32_t)j->out_stride;
            int r = zng_deflate(&d, j->flush);
            if ((j->flush == Z_FINISH ? r != Z_STREAM_END : r != Z_OK) || d.avail_in != 0) atomic_store(&j->err, 2);
            j->sizes[u] = (uint32_t)d.total_out;
            if (j->crcs) j->crcs[u] = zng_crc32(0, j->in + off, len);
            if (j->adlers) j->adlers[u] = zng_adler32(1, j->in + off, len);
            zng_deflateEnd(&d);
        } else if (j->kind == 1) {
            size_t off = u * (size_t)j->chunk;
            uint32_t len = (uint32_t)((j->n - off < j->chunk) ? (j->n - off) : j->chunk);
            if (j->crcs) j->crcs[u] = zng_crc32(0, j->in + off, len);
            if (j->adlers) j->adlers[u] = zng_adler32(1, j->in + off, len);
        } else if (j->kind == 2) {
            if (!inited) {
                if (zng_inflateInit2(&s, 31) != Z_OK) { atomic_store(&j->err, 1); break; }
                inited = 1;
            } else {
                zng_inflateReset(&s);
            }
            s.next_in = j->in + j->in_off[u]; s.avail_in = (uint32_t)(j->in_off[u + 1] - j->in_off[u]);
            s.next_out = j->out + j->out_off[u]; s.avail_out = (uint32_t)(j->out_off[u + 1] - j->out_off[u]);
            int r = zng_inflate(&s, Z_FINISH);
            j->status[u] = r;
            j->sizes[u] = (uint32_t)s.total_out;
            if (j->crcs) j->crcs[u] = s.adler;   /* gzip: running crc32 of the output */
        } else if (j->kind == 3) {
            zng_stream d;
            memset(&d, 0, sizeof(d));
            if (zng_deflateInit2(&d, j->level, Z_DEFLATED, 31, 8, Z_DEFAULT_STRATEGY) != Z_OK) { atomic_store(&j->err, 1); break; }
            d.next_in = j->in + j->in_off[u]; d.avail_in = (uint32_t)(j->in_off[u + 1] - j->in_off[u]);
            d.next_out = j->out + j->out_off[u]; d.avail_out = (uint32_t)(j->out_off[u + 1] - j->out_off[u]);
            int r = zng_deflate(&d, Z_FINISH);
            if (r != Z_STREAM_END) atomic_store(&j->err, 2);
            j->sizes[u] = (uint32_t)d.total_out;
            zng_deflateEnd(&d);
        }
    }
    if (inited) { if (j->kind == 0) zng_deflateEnd(&s); else zng_inflateEnd(&s); }
    return NULL;
}

static int run_job(job_t *j, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    atomic_store(&j->next, 0);
    atomic_store(&j->err, 0);
    if (nthreads == 1) { worker(j); return atomic_load(&j->err); }
    pthread_t *t = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int i = 0; i < nthreads; i++) { if (pthread_create(&t[i], NULL, worker, j) == 0) started++; else break; }
    if (started == 0) worker(j);
    for (int i = 0; i < started; i++) pthread_join(t[i], NULL);
    free(t);
    return atomic_load(&j->err);
}

/* Compress n bytes as ceil(n/chunk) independent raw-deflate chunks, chunk i written at
 * out + i*out_stride, its size in sizes[i]; crcs/adlers (optional) get zng_crc32 / zng_adler32
 * of the chunk's INPUT.  flush = Z_FULL_FLUSH (3) or Z_SYNC_FLUSH (2) or Z_FINISH (4). */
REFDRV_EXPORT int refdrv_deflate_chunks(const uint8_t *in, size_t n, uint32_t chunk, int level, int flush,
                                        uint8_t *out, size_t out_stride, uint32_t *sizes,
                                        uint32_t *crcs, uint32_t *adlers, int nthreads) {
    job_t j; memset(&j, 0, sizeof(j));
    j.kind = 0; j.in = in; j.n = n; j.chunk = chunk; j.level = level; j.flush = flush;
    j.out = out; j.out_stride = out_stride; j.sizes = sizes; j.crcs = crcs; j.adlers = adlers;
    j.n_units = chunk ? (n + chunk - 1) / chunk : 0;
    return run_job(&j, nthreads);
}

/* As refdrv_deflate_chunks, but every chunk after the first is primed with the 32768 stream bytes in front of it
 * (deflateSetDictionary on a fresh raw stream): pigz's default, dependent-chunk mode.  chunk must be >= 32768. */
REFDRV_EXPORT int refdrv_deflate_chunks_primed(const uint8_t *in, size_t n, uint32_t chunk, int level, int flush,
                                               uint8_t *out, size_t out_stride, uint32_t *sizes,
                                               uint32_t *crcs, uint32_t *adlers, int nthreads) {
    job_t j; memset(&j, 0, sizeof(j));
    if (chunk < 32768) return 4;
    j.kind = 4; j.in = in; j.n = n; j.chunk = chunk; j.level = level; j.flush = flush;
    j.out = out; j.out_stride = out_stride; j.sizes = sizes; j.crcs = crcs; j.adlers = adlers;
    j.n_units = (n + chunk - 1) / chunk;
    return run_job(&j, nthreads);
}

/* Per-chunk zng_crc32 / zng_adler32 (either pointer may be NULL). */
REFDRV_EXPORT int refdrv_checksum_chunks(const uint8_t *in, size_t n, uint32_t chunk,
                                         uint32_t *crcs, uint32_t *adlers, int nthreads) {
    job_t j; memset(&j, 0, sizeof(j));
    j.kind = 1; j.in = in; j.n = n; j.chunk = chunk; j.crcs = crcs; j.adlers = adlers;
    j.n_units = chunk ? (n + chunk - 1) / chunk : 0;
    return run_job(&j, nthreads);
}

/* Chunk-parallel whole-buffer checksums folded with zng_crc32_combine / zng_adler32_combine. */
REFDRV_EXPORT int refdrv_checksum_flat(const uint8_t *in, size_t n, uint32_t chunk, int nthreads,
                                       uint32_t *crc_out, uint32_t *adler_out) {
    size_t units = chunk ? (n + chunk - 1) / chunk : 0;
    uint32_t *c = (uint32_t *)malloc(sizeof(uint32_t) * (units ? units : 1));
    uint32_t *a = (uint32_t *)malloc(sizeof(uint32_t) * (units ? units : 1));
    int r = refdrv_checksum_chunks(in, n, chunk, c, a, nthreads);
    uint32_t crc = 0, ad = 1;
    for (size_t u = 0; u < units; u++) {
        size_t off = u * (size_t)chunk;
        size_t len = (n - off < chunk) ? (n - off) : chunk;
        crc = zng_crc32_combine(crc, c[u], (z_off64_t)len);
        ad = zng_adler32_combine(ad, a[u], (z_off64_t)len);
    }
    free(c); free(a);
    if (crc_out) *crc_out = crc;
    if (adler_out) *adler_out = ad;
    return r;
}

/* Inflate n_members independent gzip members.  Member i is in[in_off[i]..in_off[i+1]) and is
 * written to out[out_off[i]..out_off[i+1]).  status[i] = zng_inflate return code. */
REFDRV_EXPORT int refdrv_inflate_members(const uint8_t *in, const uint64_t *in_off, size_t n_members,
                                         uint8_t *out, const uint64_t *out_off, uint32_t *sizes,
                                         uint32_t *crcs, int32_t *status, int nthreads) {
    job_t j; memset(&j, 0, sizeof(j));
    j.kind = 2; j.in = in; j.in_off = in_off; j.out = out; j.out_off = out_off;
    j.sizes = sizes; j.crcs = crcs; j.status = status; j.n_units = n_members;
    return run_job(&j, nthreads);
}

/* Build n_members gzip members (what minigzip -<level> writes for a small input:
 * gzwrite.c:46 deflateInit2(level, Z_DEFLATED, MAX_WBITS+16, DEF_MEM_LEVEL, strategy)). */
REFDRV_EXPORT int refdrv_gzip_members(const uint8_t *in, const uint64_t *in_off, size_t n_members, int level,
                                      uint8_t *out, const uint64_t *out_off, uint32_t *sizes, int nthreads) {
    job_t j; memset(&j, 0, sizeof(j));
    j.kind = 3; j.in = in; j.in_off = in_off; j.out = out; j.out_off = out_off;
    j.sizes = sizes; j.level = level; j.n_units = n_members;
    return run_job(&j, nthreads);
}

/* Streaming inflate of one (possibly multi-GiB) gzip/zlib/raw stream; returns the zng_inflate
 * code of the last call; *total_out, *crc get the totals.  If expect != NULL the output is
 * compared on the fly (mismatch -> returns -100). */
REFDRV_EXPORT int refdrv_inflate_stream(const uint8_t *in, size_t n, int window_bits, const uint8_t *expect,
                                        size_t expect_len, uint64_t *total_out, uint32_t *crc) {
    zng_stream s; memset(&s, 0, sizeof(s));
    if (zng_inflateInit2(&s, window_bits) != Z_OK) return -99;
    enum { BUF = 1 << 20 };
    uint8_t *buf = (uint8_t *)malloc(BUF);
    size_t ipos = 0; uint64_t opos = 0; int r = Z_OK; int bad = 0;
    do {
        if (s.avail_in == 0 && ipos < n) {
            size_t take = n - ipos; if (take > (1u << 30)) take = 1u << 30;
            s.next_in = in + ipos; s.avail_in = (uint32_t)take; ipos += take;
        }
        s.next_out = buf; s.avail_out = BUF;
        r = zng_inflate(&s, Z_NO_FLUSH);
        size_t got = BUF - s.avail_out;
        if (expect) {
            if (opos + got > expect_len || memcmp(expect + opos, buf, got) != 0) bad = 1;
        }
        opos += got;
        if (r != Z_OK) break;
        if (got == 0 && s.avail_in == 0 && ipos >= n) { r = Z_BUF_ERROR; break; }
    } while (1);
    if (total_out) *total_out = opos;
    if (crc) *crc = s.adler;
    zng_inflateEnd(&s);
    free(buf);
    if (bad || (expect && opos != expect_len)) return -100;
    return r;
}

/* One zng_inflate(Z_FINISH) call on a fresh stream: what zo_inflate restates.  msg (>= 64 bytes) receives
 * strm->msg or an empty string.  Returns the zng_inflate code, or the zng_inflateInit2 code if that failed. */
REFDRV_EXPORT int refdrv_inflate_oneshot(const uint8_t *in, size_t n, int window_bits, uint8_t *out, size_t cap,
                                         size_t *out_len, size_t *in_used, uint32_t *check, char *msg) {
    zng_stream s; memset(&s, 0, sizeof(s));
    if (msg) msg[0] = 0;
    int r = zng_inflateInit2(&s, window_bits);
    if (r != Z_OK) return r;
    s.next_in = in; s.avail_in = (uint32_t)n; s.next_out = out; s.avail_out = (uint32_t)cap;
    r = zng_inflate(&s, Z_FINISH);
    if (out_len) *out_len = s.total_out;
    if (in_used) *in_used = s.total_in;
    if (check) *check = s.adler;
    if (msg && s.msg) { strncpy(msg, s.msg, 63); msg[63] = 0; }
    zng_inflateEnd(&s);
    return r;
}

/* One whole gzip/zlib/raw stream through the reference's deflate with the pigz-style call sequence the
 * host library promises to reproduce: Z_FULL_FLUSH after every `piece` bytes, Z_FINISH on the last piece. */
REFDRV_EXPORT int refdrv_deflate_stream(const uint8_t *in, size_t n, uint32_t piece, int level, int window_bits,
                                        uint8_t *out, size_t cap, size_t *out_len) {
    zng_stream s; memset(&s, 0, sizeof(s));
    int r = zng_deflateInit2(&s, level, Z_DEFLATED, window_bits, 8, Z_DEFAULT_STRATEGY);
    if (r != Z_OK) return r;
    s.next_out = out; s.avail_out = (uint32_t)(cap > 0xffffffffu ? 0xffffffffu : cap);
    size_t off = 0;
    do {
        size_t take = n - off < piece ? n - off : piece;
        int fin = (off + take == n);
        s.next_in = in + off; s.avail_in = (uint32_t)take;
        r = zng_deflate(&s, fin ? Z_FINISH : Z_FULL_FLUSH);
        off += take;
        if (fin) break;
        if (r != Z_OK) break;
    } while (1);
    if (out_len) *out_len = s.total_out;
    zng_deflateEnd(&s);
    return r;
}

REFDRV_EXPORT double refdrv_now(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
