/* zng_functable.c -- the host-callable operator table (struct zng_b200_functable, include/zng_b200.h): the 15 slots of the
 * reference's struct functable_s (functable.h:26-42) in the same order, and the three hash callbacks of deflate_state
 * (deflate.h:121-131).  Every entry runs the device operator K1 / K2 / K3 / K4 use (csrc/lz_ops.cuh, csrc/checksum.cu) through
 * the calling thread's context; operands travel through a scratch area the context owns (zng_b200_ctx_arena: no allocation per
 * call).  There is no CPU implementation behind any of them: without a device they return 0 / do nothing and the context (or
 * zng_b200_last_error) says why.  These calls cost a host<->device round trip each -- they exist so that every operator can be
 * checked against the reference's own variant, not to be called per position; the batched forms are the zng_b200_op_* entries. */
#include "zng_host.h"
#include <cuda_runtime_api.h>
#include <string.h>

#define WIN_PAD 512u                         /* readable slack behind the window (compare256 over-read), zero filled */
#define HEAD_BYTES (65536u * 2u)
#define PREV_BYTES (32768u * 2u)

void *zng_b200_ctx_arena(zng_b200_ctx *ctx, size_t bytes);

static void ft_force_init(void) { (void)zng_b200_thread_ctx(); }            /* functable.c:45-266: here = create the thread's context */
static uint32_t ft_adler32(uint32_t adler, const uint8_t *buf, size_t len) { return zng_adler32_z(adler, buf, len); }
static uint32_t ft_crc32(uint32_t crc, const uint8_t *buf, size_t len) { return zng_crc32_z(crc, buf, len); }
static uint32_t ft_chunksize(void) { return 32; }                            /* bytes moved per wave of the device copy: one warp */

/* arch/generic/adler32_fold_c.c:11-15 */
static uint32_t ft_adler32_fold_copy(uint32_t adler, uint8_t *dst, const uint8_t *src, size_t len) {
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    uint32_t r = adler;
    if (!src) return 1;                                                      /* adler32.c:19 semantics of a NULL buffer */
    if (!ctx || zng_b200_adler32_copy_host(ctx, dst, src, len, adler, &r) != ZNG_B200_OK) return 0;
    return r;
}

/* arch/generic/crc32_fold_c.c:10-30 */
static uint32_t ft_crc32_fold_reset(struct zng_b200_crc32_fold *crc) { crc->value = 0; return crc->value; }
static void ft_crc32_fold(struct zng_b200_crc32_fold *crc, const uint8_t *src, size_t len, uint32_t init_crc) {
    (void)init_crc;                                                          /* unused by the reference's generic variant as well */
    crc->value = zng_crc32_z(crc->value, src, len);
}
static void ft_crc32_fold_copy(struct zng_b200_crc32_fold *crc, uint8_t *dst, const uint8_t *src, size_t len) {
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    uint32_t r = crc->value;
    if (ctx && zng_b200_crc32_copy_host(ctx, dst, src, len, crc->value, &r) == ZNG_B200_OK) crc->value = r;
}
static uint32_t ft_crc32_fold_final(struct zng_b200_crc32_fold *crc) { return crc->value; }

/* functable.compare256(src0, src1): both operands are 256 readable bytes (compare256_c.c:12) */
static uint32_t ft_compare256(const uint8_t *src0, const uint8_t *src1) {
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    uint8_t *d = ctx ? (uint8_t *)zng_b200_ctx_arena(ctx, 1024) : NULL;
    uint32_t r = 0;
    if (!d) return 0;
    cudaMemset(d, 0, 1024);
    cudaMemcpy(d, src0, 256, cudaMemcpyHostToDevice);
    cudaMemcpy(d + 288, src1, 256, cudaMemcpyHostToDevice);
    if (zng_b200_op_compare256(ctx, d, d + 288, 0, 1, (uint32_t *)(d + 576), NULL) == ZNG_B200_OK)
        cudaMemcpy(&r, d + 576, 4, cudaMemcpyDeviceToHost);
    return r;
}

/* functable.chunkmemset_safe(out, from, len, left): copy len bytes from `from` (= out - dist) to out, never writing past
 * out + left; returns out + len (chunkset_tpl.h:229-283) */
static uint8_t *ft_chunkmemset_safe(uint8_t *out, uint8_t *from, unsigned len, unsigned left) {
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    if (len > left) len = left;
    if (!ctx || len == 0 || from >= out) return out + len;
    size_t dist = (size_t)(out - from);
    uint8_t *d = (uint8_t *)zng_b200_ctx_arena(ctx, dist + len + 64);
    if (!d) return out + len;
    cudaMemcpy(d, from, dist, cudaMemcpyHostToDevice);
    if (zng_b200_op_chunkmemset(ctx, d, (uint32_t)dist, (uint32_t)dist, len, NULL) == ZNG_B200_OK)
        cudaMemcpy(out, d + dist, len, cudaMemcpyDeviceToHost);
    return out + len;
}

/* inffast_tpl.h:53: see the slot's comment in include/zng_b200.h */
static void ft_inflate_fast(void *strm, uint32_t start) {
    (void)start;
    if (strm) (void)zng_inflate((zng_stream *)strm, Z_SYNC_FLUSH);
}

/* arena layout of the match-state operators: window (65536 + pad) | head | prev | 64 bytes of scalars */
static uint8_t *stage_state(zng_b200_ctx *ctx, const struct zng_b200_match_state *s, int want_head) {
    if (!ctx || !s || !s->window || !s->prev || s->window_len > 65536u || (want_head && !s->head)) return NULL;
    uint8_t *d = (uint8_t *)zng_b200_ctx_arena(ctx, 65536u + WIN_PAD + HEAD_BYTES + PREV_BYTES + 64u);
    if (!d) return NULL;
    cudaMemset(d, 0, 65536u + WIN_PAD);
    if (s->window_len) cudaMemcpy(d, s->window, s->window_len, cudaMemcpyHostToDevice);
    if (want_head) cudaMemcpy(d + 65536u + WIN_PAD, s->head, HEAD_BYTES, cudaMemcpyHostToDevice);
    cudaMemcpy(d + 65536u + WIN_PAD + HEAD_BYTES, s->prev, PREV_BYTES, cudaMemcpyHostToDevice);
    return d;
}

/* match_tpl.h:26-280 (non-SLOW instantiation) */
static uint32_t ft_longest_match(struct zng_b200_match_state *s, uint16_t cur_match) {
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    uint8_t *d = stage_state(ctx, s, 0);
    if (!d) return 0;
    uint32_t *sc = (uint32_t *)(d + 65536u + WIN_PAD + HEAD_BYTES + PREV_BYTES);
    uint32_t q[2] = {s->strstart, cur_match}, r[2] = {0, 0};
    cudaMemcpy(sc, q, 8, cudaMemcpyHostToDevice);
    if (zng_b200_op_longest_match_level(ctx, d, s->window_len, (const uint16_t *)(d + 65536u + WIN_PAD + HEAD_BYTES), sc, sc + 1, 1,
                                        s->level | 0x100, sc + 2, sc + 3, NULL) != ZNG_B200_OK) return 0;
    cudaMemcpy(r, sc + 2, 8, cudaMemcpyDeviceToHost);
    if (r[0] > 2u) s->match_start = r[1];
    return r[0];
}

static uint32_t ft_longest_match_slow(struct zng_b200_match_state *s, uint16_t cur_match) {
    (void)s; (void)cur_match;
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    if (ctx) (void)zng_b200_op_longest_match_level(ctx, NULL, 0, NULL, NULL, NULL, 0, 9, NULL, NULL, NULL);   /* records the reason */
    return 0;
}

/* arch/generic/slide_hash_c.c:47-52 */
static void ft_slide_hash(struct zng_b200_match_state *s) {
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    if (!ctx || !s || !s->head || !s->prev) return;
    uint8_t *d = (uint8_t *)zng_b200_ctx_arena(ctx, HEAD_BYTES + PREV_BYTES);
    if (!d) return;
    cudaMemcpy(d, s->head, HEAD_BYTES, cudaMemcpyHostToDevice);
    cudaMemcpy(d + HEAD_BYTES, s->prev, PREV_BYTES, cudaMemcpyHostToDevice);
    if (zng_b200_op_slide_hash(ctx, (uint16_t *)d, (uint16_t *)(d + HEAD_BYTES), 32768u, NULL) != ZNG_B200_OK) return;
    cudaMemcpy(s->head, d, HEAD_BYTES, cudaMemcpyDeviceToHost);
    cudaMemcpy(s->prev, d + HEAD_BYTES, PREV_BYTES, cudaMemcpyDeviceToHost);
}

static const struct zng_b200_functable table = {
    ft_force_init, ft_adler32, ft_adler32_fold_copy, ft_chunkmemset_safe, ft_chunksize, ft_compare256, ft_crc32,
    ft_crc32_fold, ft_crc32_fold_copy, ft_crc32_fold_final, ft_crc32_fold_reset, ft_inflate_fast,
    ft_longest_match, ft_longest_match_slow, ft_slide_hash};
const struct zng_b200_functable *zng_b200_functable_get(void) { return &table; }

/* ---- the hash callbacks (deflate.h:121-131) ---- */
/* insert_string.c:11-13 HASH_CALC: pure 32-bit arithmetic on the caller's value, nothing to offload */
uint32_t zng_b200_update_hash(uint32_t h, uint32_t val) { (void)h; return (val * 2654435761u) >> 16; }

void zng_b200_insert_string(struct zng_b200_match_state *s, uint32_t str, uint32_t count) {
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    uint8_t *d = stage_state(ctx, s, 1);
    if (!d || count == 0) return;
    uint16_t *dh = (uint16_t *)(d + 65536u + WIN_PAD), *dp = (uint16_t *)(d + 65536u + WIN_PAD + HEAD_BYTES);
    if (zng_b200_op_insert_string(ctx, d, dh, dp, str, count, NULL) != ZNG_B200_OK) return;
    cudaMemcpy(s->head, dh, HEAD_BYTES, cudaMemcpyDeviceToHost);
    cudaMemcpy(s->prev, dp, PREV_BYTES, cudaMemcpyDeviceToHost);
}

uint16_t zng_b200_quick_insert_string(struct zng_b200_match_state *s, uint32_t str) {
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    uint8_t *d = stage_state(ctx, s, 1);
    if (!d) return 0;
    uint16_t *dh = (uint16_t *)(d + 65536u + WIN_PAD), *dp = (uint16_t *)(d + 65536u + WIN_PAD + HEAD_BYTES);
    uint32_t *sc = (uint32_t *)(d + 65536u + WIN_PAD + HEAD_BYTES + PREV_BYTES), old = 0;
    if (zng_b200_op_quick_insert_string(ctx, d, dh, dp, str, sc, NULL) != ZNG_B200_OK) return 0;
    cudaMemcpy(&old, sc, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(s->head, dh, HEAD_BYTES, cudaMemcpyDeviceToHost);
    cudaMemcpy(s->prev, dp, PREV_BYTES, cudaMemcpyDeviceToHost);
    return (uint16_t)old;
}
