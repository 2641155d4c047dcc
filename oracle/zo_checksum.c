/* oracle/zo_checksum.c -- TEST INFRASTRUCTURE ONLY (see zo_oracle.h).
 *
 * CPU restatement of the reference's checksum operators:
 *   CRC-32   : crc32.c:27-41 (NULL -> 0, pre/post inversion), arch/generic/crc32_braid_c.c:62-216
 *              (reflected polynomial 0xEDB88320; the braid is an optimisation of the plain
 *              table-driven CRC, restated here as slicing-by-8 with tables generated at start-up
 *              the way tools/makecrct.c generates crc32_braid_tbl.h)
 *   combine  : crc32_braid_comb_p.h:8-40 (multmodp, x2nmodp), crc32_braid_comb.c:16-24
 *   Adler-32 : arch/generic/adler32_c.c:11-54, adler32_p.h:11-68 (BASE 65521, NMAX 5552)
 *   combine  : adler32.c:32-54
 *   compare256: arch/generic/compare256_c.c:12-43
 */
#include "zo_oracle.h"
#include <pthread.h>

#define ZO_POLY 0xedb88320u   /* crc32_braid_p.h: reflected CRC-32 polynomial */

static uint32_t zo_tbl[8][256];
static uint32_t zo_x2n[32];
static pthread_once_t zo_once = PTHREAD_ONCE_INIT;

/* crc32_braid_comb_p.h:8-23 -- a(x)*b(x) mod p(x), bit 31 is x^0.  Requires a != 0. */
static uint32_t zo_multmodp(uint32_t a, uint32_t b) {
    uint32_t acc = 0;
    for (uint32_t bit = 0x80000000u;; bit >>= 1) {
        if (a & bit) {
            acc ^= b;
            if ((a & (bit - 1)) == 0) return acc;
        }
        b = (b >> 1) ^ ((b & 1) ? ZO_POLY : 0);
    }
}

static void zo_init_tables(void) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1) ? ZO_POLY : 0);
        zo_tbl[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int t = 1; t < 8; t++)
            zo_tbl[t][i] = (zo_tbl[t - 1][i] >> 8) ^ zo_tbl[0][zo_tbl[t - 1][i] & 0xff];
    /* x2n_table[k] = x^(2^k) mod p (crc32_braid_tbl.h:9437-9444 holds the same 32 values) */
    zo_x2n[0] = 0x40000000u;
    for (int k = 1; k < 32; k++) zo_x2n[k] = zo_multmodp(zo_x2n[k - 1], zo_x2n[k - 1]);
}

/* crc32_braid_comb_p.h:29-40 -- x^(n * 2^k) mod p */
static uint32_t zo_x2nmodp(int64_t n, unsigned k) {
    pthread_once(&zo_once, zo_init_tables);
    uint32_t p = 0x80000000u;
    while (n) {
        if (n & 1) p = zo_multmodp(zo_x2n[k & 31], p);
        n >>= 1;
        k++;
    }
    return p;
}

uint32_t zo_crc32(uint32_t crc, const uint8_t *buf, size_t len) {
    if (buf == NULL) return 0;                      /* crc32.c:28 */
    pthread_once(&zo_once, zo_init_tables);
    uint32_t c = ~crc;                              /* crc32_braid_c.c:66 */
    while (len && ((uintptr_t)buf & 7)) { c = (c >> 8) ^ zo_tbl[0][(c ^ *buf++) & 0xff]; len--; }
    while (len >= 8) {
        uint32_t lo = (uint32_t)buf[0] | ((uint32_t)buf[1] << 8) | ((uint32_t)buf[2] << 16) | ((uint32_t)buf[3] << 24);
        uint32_t hi = (uint32_t)buf[4] | ((uint32_t)buf[5] << 8) | ((uint32_t)buf[6] << 16) | ((uint32_t)buf[7] << 24);
        lo ^= c;
        c = zo_tbl[7][lo & 0xff] ^ zo_tbl[6][(lo >> 8) & 0xff] ^ zo_tbl[5][(lo >> 16) & 0xff] ^ zo_tbl[4][lo >> 24] ^
            zo_tbl[3][hi & 0xff] ^ zo_tbl[2][(hi >> 8) & 0xff] ^ zo_tbl[1][(hi >> 16) & 0xff] ^ zo_tbl[0][hi >> 24];
        buf += 8; len -= 8;
    }
    while (len--) c = (c >> 8) ^ zo_tbl[0][(c ^ *buf++) & 0xff];
    return ~c;
}

/* crc32_braid_comb.c:16-24 */
uint32_t zo_crc32_combine(uint32_t crc1, uint32_t crc2, int64_t len2) {
    return zo_multmodp(zo_x2nmodp(len2, 3), crc1) ^ crc2;
}
uint32_t zo_crc32_combine_gen(int64_t len2) { return zo_x2nmodp(len2, 3); }
uint32_t zo_crc32_combine_op(uint32_t crc1, uint32_t crc2, uint32_t op) { return zo_multmodp(op, crc1) ^ crc2; }

#define ZO_BASE 65521u   /* adler32_p.h:11 */
#define ZO_NMAX 5552u    /* adler32_p.h:12 */

uint32_t zo_adler32(uint32_t adler, const uint8_t *buf, size_t len) {
    uint32_t s2 = (adler >> 16) & 0xffff, s1 = adler & 0xffff;
    if (len == 1) {                                 /* adler32_c.c:20-21, adler32_p.h:21-29: reads buf[0] */
        s1 += buf[0]; if (s1 >= ZO_BASE) s1 -= ZO_BASE;
        s2 += s1;     if (s2 >= ZO_BASE) s2 -= ZO_BASE;
        return s1 | (s2 << 16);
    }
    if (buf == NULL) return 1;                      /* adler32_c.c:24-25 */
    while (len) {
        size_t n = len < ZO_NMAX ? len : ZO_NMAX;  /* one modulo per <= NMAX bytes, adler32_c.c:32-50 */
        len -= n;
        while (n--) { s1 += *buf++; s2 += s1; }
        s1 %= ZO_BASE; s2 %= ZO_BASE;
    }
    return s1 | (s2 << 16);
}

/* adler32.c:32-54 */
uint32_t zo_adler32_combine(uint32_t adler1, uint32_t adler2, int64_t len2) {
    if (len2 < 0) return 0xffffffffu;
    uint32_t rem = (uint32_t)(len2 % ZO_BASE);
    uint32_t s1 = adler1 & 0xffff;
    uint32_t s2 = (rem * s1) % ZO_BASE;
    s1 += (adler2 & 0xffff) + ZO_BASE - 1;
    s2 += ((adler1 >> 16) & 0xffff) + ((adler2 >> 16) & 0xffff) + ZO_BASE - rem;
    if (s1 >= ZO_BASE) s1 -= ZO_BASE;
    if (s1 >= ZO_BASE) s1 -= ZO_BASE;
    if (s2 >= (ZO_BASE << 1)) s2 -= (ZO_BASE << 1);
    if (s2 >= ZO_BASE) s2 -= ZO_BASE;
    return s1 | (s2 << 16);
}

/* arch/generic/compare256_c.c:12-43 -- index of the first differing byte, 256 if none */
uint32_t zo_compare256(const uint8_t *a, const uint8_t *b) {
    uint32_t i = 0;
    while (i < 256 && a[i] == b[i]) i++;
    return i;
}

/* Test helper (no reference counterpart): compare n chunk outputs byte for byte.  `a` holds chunk i at a + i*stride_a
 * (stride_a != 0: slots) or packed back to back in chunk order (stride_a == 0); `b` likewise.  Returns the number of
 * chunks whose size or bytes differ; *first_bad = index of the first one (n if none). */
#include <string.h>
size_t zo_compare_chunks(const uint8_t *a, size_t stride_a, const uint32_t *sizes_a,
                         const uint8_t *b, size_t stride_b, const uint32_t *sizes_b, size_t n, size_t *first_bad) {
    size_t bad = 0, oa = 0, ob = 0, first = n;
    for (size_t i = 0; i < n; i++) {
        const uint8_t *pa = stride_a ? a + i * stride_a : a + oa;
        const uint8_t *pb = stride_b ? b + i * stride_b : b + ob;
        if (sizes_a[i] != sizes_b[i] || memcmp(pa, pb, sizes_a[i]) != 0) { if (!bad) first = i; bad++; }
        oa += sizes_a[i]; ob += sizes_b[i];
    }
    if (first_bad) *first_bad = first;
    return bad;
}

/* Test helper: wrap n raw-deflate bodies (body i at bodies + i*stride, sizes[i] bytes, CRC-32 crcs[i] of its raw_len[i]
 * uncompressed bytes) into gzip members as minigzip -1 frames them (deflate.c:902-921 header 1f 8b 08 00 00000000 XFL OS with
 * XFL 4 for level 1, OS 3; deflate.c:1091-1096 trailer CRC-32, ISIZE), packed back to back.  off[n+1] receives the member
 * offsets.  Returns the total length (call with out == NULL to size the buffer). */
size_t zo_frame_gzip_members(const uint8_t *bodies, size_t stride, const uint32_t *sizes, const uint32_t *crcs,
                             uint32_t raw_len, uint32_t last_raw_len, size_t n, int xfl, uint8_t *out, uint64_t *off) {
    size_t o = 0;
    for (size_t i = 0; i < n; i++) {
        if (off) off[i] = o;
        if (out) {
            const uint8_t hdr[10] = {0x1f, 0x8b, 8, 0, 0, 0, 0, 0, (uint8_t)xfl, 3};
            memcpy(out + o, hdr, 10);
            memcpy(out + o + 10, bodies + i * stride, sizes[i]);
            uint32_t t[2] = {crcs[i], (i + 1 == n) ? last_raw_len : raw_len};
            memcpy(out + o + 10 + sizes[i], t, 8);          /* little-endian host (x86-64, the frozen oracle platform) */
        }
        o += 18u + sizes[i];
    }
    if (off) off[n] = o;
    return o;
}
