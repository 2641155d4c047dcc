/* synth.c -- deterministic synthetic "mixed text/binary" input for the chunked-DEFLATE path.
 *
 * BASELINE.json names the workload "synthetic mixed text/binary buffer split into independent
 * 64 KiB chunks"; SURVEY.md section 8(d) fixes the mix: per 10 consecutive 64 KiB units, 4 are
 * text (skewed word frequencies from a 4096-word dictionary), 3 are structured binary (32-byte
 * records), 2 are incompressible (raw PRNG) and 1 is highly repetitive (runs and short-period
 * patterns).  Every 64 KiB unit is generated from (seed, unit index) alone, so each rank of a
 * multi-GPU run can fill its own shard and all ranks agree on the bytes.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY: built into tests/libzng_synth.so (oracle/Makefile target `synth`), loaded by
 * tests/synthdata.py.  Not part of the product library and not part of the reference's API.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SYNTH_UNIT 65536u
#define DICT_WORDS 4096u
#define WORD_MAX   12u

static inline uint64_t mix64(uint64_t *st) {      /* splitmix64 */
    uint64_t z = (*st += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

typedef struct { uint8_t len[DICT_WORDS]; uint8_t txt[DICT_WORDS][WORD_MAX]; } dict_t;

static void build_dict(dict_t *d, uint64_t seed) {
    static const char freq_letters[] = "eeeeeeetttttaaaaooooiiiinnnnsssshhhrrrddllcumwfgypbvk";
    uint64_t st = seed ^ 0xD1C7D1C7D1C7D1C7ull;
    for (uint32_t w = 0; w < DICT_WORDS; w++) {
        uint64_t r = mix64(&st);
        uint32_t len = 2u + (uint32_t)(r % 5u) + (uint32_t)((r >> 8) % 5u) + (w > 512u ? (uint32_t)((r >> 16) & 1u) : 0u);
        if (len > WORD_MAX) len = WORD_MAX;
        d->len[w] = (uint8_t)len;
        for (uint32_t k = 0; k < len; k++) {
            uint64_t q = mix64(&st);
            d->txt[w][k] = (uint8_t)freq_letters[q % (sizeof(freq_letters) - 1)];
        }
        if ((r >> 40) % 11u == 0) d->txt[w][0] = (uint8_t)(d->txt[w][0] - 32);   /* capitalised */
    }
}

static void fill_text(uint8_t *p, uint32_t n, uint64_t st, const dict_t *d) {
    uint32_t o = 0, col = 0;
    while (o < n) {
        uint64_t r = mix64(&st);
        /* skewed rank: cube of a uniform 10-bit fraction spreads over 4096 ranks, head-heavy */
        uint32_t u = (uint32_t)(r & 0x3ffu);
        uint32_t rank = (uint32_t)(((uint64_t)u * u * u) >> 18);     /* 0 .. 4095 */
        if (rank >= DICT_WORDS) rank = DICT_WORDS - 1;
        uint32_t len = d->len[rank];
        for (uint32_t k = 0; k < len && o < n; k++) p[o++] = d->txt[rank][k];
        col += len + 1;
        uint32_t t = (uint32_t)(r >> 32) % 23u;
        if (o < n && t == 0) p[o++] = ',';
        if (o < n && t == 1) { p[o++] = '.'; }
        if (o < n) {
            if (col > 72u) { p[o++] = '\n'; col = 0; }
            else p[o++] = ' ';
        }
        if (t == 2 && o + 6 < n) {                                    /* a number now and then */
            uint32_t v = (uint32_t)(r >> 20) % 100000u;
            char tmp[8]; int k = 0;
            do { tmp[k++] = (char)('0' + v % 10u); v /= 10u; } while (v);
            while (k) p[o++] = (uint8_t)tmp[--k];
            p[o++] = ' ';
        }
    }
}

static void fill_records(uint8_t *p, uint32_t n, uint64_t st, uint64_t unit) {
    uint32_t counter = (uint32_t)(unit * 2048u);
    uint32_t o = 0;
    uint8_t rec[32];
    while (o < n) {
        uint64_t r = mix64(&st);
        memset(rec, 0, sizeof(rec));
        rec[0] = (uint8_t)counter; rec[1] = (uint8_t)(counter >> 8); rec[2] = (uint8_t)(counter >> 16); rec[3] = (uint8_t)(counter >> 24);
        rec[4] = (uint8_t)(r % 7u);                          /* small enum */
        rec[5] = (uint8_t)((r >> 8) % 3u);
        /* 6..11 zero padding */
        uint64_t v = mix64(&st);
        memcpy(rec + 12, &v, 8);                              /* 8 random bytes */
        uint32_t ts = 1700000000u + counter * 3u + (uint32_t)((r >> 16) & 3u);
        memcpy(rec + 20, &ts, 4);                             /* slowly increasing timestamp */
        rec[24] = (uint8_t)((r >> 24) & 1u);
        /* 25..31 zero */
        uint32_t take = n - o < 32u ? n - o : 32u;
        memcpy(p + o, rec, take);
        o += take; counter++;
    }
}

static void fill_random(uint8_t *p, uint32_t n, uint64_t st) {
    uint32_t o = 0;
    while (o + 8 <= n) { uint64_t r = mix64(&st); memcpy(p + o, &r, 8); o += 8; }
    if (o < n) { uint64_t r = mix64(&st); memcpy(p + o, &r, n - o); }
}

static void fill_repetitive(uint8_t *p, uint32_t n, uint64_t st) {
    uint32_t o = 0;
    while (o < n) {
        uint64_t r = mix64(&st);
        uint32_t kind = (uint32_t)(r % 4u);
        uint32_t span = 16u + (uint32_t)((r >> 8) % 3000u);
        if (span > n - o) span = n - o;
        if (kind == 0) {                                     /* run of one byte (dist-1 matches, 258 clipping) */
            memset(p + o, (int)((r >> 24) & 0xff), span);
        } else if (kind == 1) {                              /* short period pattern */
            uint32_t period = 2u + (uint32_t)((r >> 32) % 62u);
            uint64_t s2 = r;
            for (uint32_t k = 0; k < span; k++) {
                if (k < period) p[o + k] = (uint8_t)mix64(&s2);
                else p[o + k] = p[o + k - period];
            }
        } else if (kind == 2 && o > 4096u) {                 /* far copy of earlier bytes of this unit */
            uint32_t back = 1u + (uint32_t)((r >> 32) % (o < 40000u ? o : 40000u));
            for (uint32_t k = 0; k < span; k++) p[o + k] = p[o + k - back];
        } else {                                             /* zeros: aliases the empty hash slot (position 0) */
            memset(p + o, 0, span);
        }
        o += span;
    }
}

/* Fill buf[0..n) with the bytes of the synthetic stream starting at absolute byte offset
 * `offset` (must be a multiple of 65536).  Deterministic in (seed, offset, n). */
__attribute__((visibility("default"))) int zng_synth_fill(void *buf, size_t n, uint64_t seed, uint64_t offset) {
    if (offset % SYNTH_UNIT) return -2;
    dict_t *d = (dict_t *)malloc(sizeof(dict_t));
    if (!d) return -4;
    build_dict(d, seed);
    uint8_t *p = (uint8_t *)buf;
    uint64_t unit = offset / SYNTH_UNIT;
    size_t o = 0;
    while (o < n) {
        uint32_t len = (uint32_t)(n - o < SYNTH_UNIT ? n - o : SYNTH_UNIT);
        uint64_t st = seed ^ (unit * 0xA24BAED4963EE407ull + 0x9FB21C651E98DF25ull);
        (void)mix64(&st);
        switch (unit % 10u) {
            case 0: case 2: case 5: case 7: fill_text(p + o, len, st, d); break;
            case 1: case 4: case 8:         fill_records(p + o, len, st, unit); break;
            case 3: case 9:                 fill_random(p + o, len, st); break;
            default:                        fill_repetitive(p + o, len, st); break;   /* 6 */
        }
        o += len; unit++;
    }
    free(d);
    return 0;
}
