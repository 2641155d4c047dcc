/* zng_checksum.c -- zng_crc32 / zng_adler32 (+ _z) and the combine functions of the host library.
 *
 * Reference interface: crc32.c:27-41 (zng_crc32_z: NULL buffer -> 0), adler32.c:15-28 +
 * arch/generic/adler32_c.c:20-25 (len == 1 is evaluated before the NULL test; NULL -> 1),
 * crc32_braid_comb.c:16-53 (combine, combine_gen, combine_op), adler32.c:32-68 (adler32_combine).
 * The data-touching functions run on the GPU (K3, csrc/checksum.cu) through the C-ABI; the combine
 * functions are 32-bit scalar arithmetic on two words and stay on the host, as in the reference.
 */
#include "zng_host.h"
#include <stdio.h>
#include <stdlib.h>

#define POLY 0xedb88320u
#define BASE 65521u

/* crc32_braid_comb_p.h:8-23  a(x)*b(x) mod p(x); bit 31 is the x^0 coefficient */
uint32_t zng_host_multmodp(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (uint32_t m = 0x80000000u; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b >> 1) ^ ((b & 1u) ? POLY : 0u);
    }
    return p;
}

/* crc32_braid_comb_p.h:29-40  x^(n * 2^k) mod p(x); x2n_table[i] = x^(2^i) mod p is rebuilt by squaring */
uint32_t zng_host_x2nmodp(int64_t n, unsigned k) {
    static uint32_t x2n[32];
    static int ready;
    if (!ready) {            /* idempotent initialisation: racing threads write the same values */
        uint32_t v = 0x40000000u;
        uint32_t t[32];
        for (int i = 0; i < 32; i++) { t[i] = v; v = zng_host_multmodp(v, v); }
        for (int i = 0; i < 32; i++) x2n[i] = t[i];
        __atomic_store_n(&ready, 1, __ATOMIC_RELEASE);
    }
    uint32_t p = 0x80000000u;
    while (n) {
        if (n & 1) p = zng_host_multmodp(x2n[k & 31], p);
        n >>= 1;
        k++;
    }
    return p;
}

uint32_t zng_crc32_combine(uint32_t crc1, uint32_t crc2, z_off64_t len2) {
    return zng_host_multmodp(zng_host_x2nmodp(len2, 3), crc1) ^ crc2;
}
uint32_t zng_crc32_combine_gen(z_off64_t len2) { return zng_host_x2nmodp(len2, 3); }
uint32_t zng_crc32_combine_op(uint32_t crc1, uint32_t crc2, uint32_t op) { return zng_host_multmodp(op, crc1) ^ crc2; }

uint32_t zng_adler32_combine(uint32_t adler1, uint32_t adler2, z_off64_t len2) {
    if (len2 < 0) return 0xffffffffu;                      /* adler32.c:37-38 */
    uint32_t rem = (uint32_t)(len2 % BASE);
    uint32_t sum1 = adler1 & 0xffff;
    uint32_t sum2 = (rem * sum1) % BASE;
    sum1 += (adler2 & 0xffff) + BASE - 1;
    sum2 += ((adler1 >> 16) & 0xffff) + ((adler2 >> 16) & 0xffff) + BASE - rem;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum1 >= BASE) sum1 -= BASE;
    if (sum2 >= (BASE << 1)) sum2 -= (BASE << 1);
    if (sum2 >= BASE) sum2 -= BASE;
    return sum1 | (sum2 << 16);
}

/* zng_crc32 / zng_adler32 cannot report an error (crc32.c:27-41, adler32.c:15-28 return the value), and any value returned after a
 * failed GPU call would look like a valid checksum -- a caller writing a gzip trailer with it would emit a corrupt stream without
 * knowing.  There is no CPU implementation to fall back to (north star), so the failure is made impossible to miss: the reason goes
 * to stderr and the process aborts.  (zng_deflate / zng_inflate report device failures through their return codes instead.) */
static void checksum_failed(const char *fn, zng_b200_ctx *ctx) {
    fprintf(stderr, "libzng_b200: %s: %s -- no CPU fallback, aborting\n", fn, ctx ? zng_b200_last_error(ctx) : "no CUDA device / context");
    abort();
}

uint32_t zng_crc32_z(uint32_t crc, const uint8_t *buf, size_t len) {
    if (buf == NULL) return 0;                              /* crc32.c:28 */
    if (len == 0) return crc;
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    uint32_t r = 0;
    if (!ctx || zng_b200_crc32_host(ctx, buf, len, crc, &r) != ZNG_B200_OK) checksum_failed("zng_crc32", ctx);
    return r;
}
uint32_t zng_crc32(uint32_t crc, const uint8_t *buf, uint32_t len) { return zng_crc32_z(crc, buf, len); }

uint32_t zng_adler32_z(uint32_t adler, const uint8_t *buf, size_t len) {
    if (len != 1 && buf == NULL) return 1;                  /* adler32_c.c:20-25: len == 1 reads buf[0] first */
    if (len == 0) return adler;
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    uint32_t r = 1;
    if (!ctx || zng_b200_adler32_host(ctx, buf, len, adler, &r) != ZNG_B200_OK) checksum_failed("zng_adler32", ctx);
    return r;
}
uint32_t zng_adler32(uint32_t adler, const uint8_t *buf, uint32_t len) { return zng_adler32_z(adler, buf, len); }
