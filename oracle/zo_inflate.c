/* oracle/zo_inflate.c -- TEST INFRASTRUCTURE ONLY (see zo_oracle.h).
 *
 * CPU restatement of what ONE call zng_inflate(strm, Z_FINISH) of the reference returns for a stream
 * held completely in memory (fresh state, no preset dictionary), for raw / zlib / gzip wrappers:
 *
 *   inflate.c:219-255    inflateInit2 windowBits decoding (raw < 0, +16 gzip, +32 auto)
 *   inflate.c:509-703    HEAD .. HCRC: zlib header check, gzip header fields, optional header CRC
 *   inflate.c:726-799    TYPEDO / STORED / COPY
 *   inflate.c:801-922    TABLE / LENLENS / CODELENS: dynamic block header and its six error strings
 *   inftrees.c:30-295    zng_inflate_table: over-subscribed / incomplete code detection (:107-130)
 *   inflate.c:928-1107, inffast_tpl.h:151-298   symbol decode, length/distance extras, match copy,
 *                        "invalid literal/length code", "invalid distance code", "invalid distance too far back"
 *   inflate.c:1109-1151  CHECK / LENGTH trailers
 *   inflate.c:1176-1200  return value: Z_STREAM_END, Z_DATA_ERROR, Z_NEED_DICT, else Z_BUF_ERROR (Z_FINISH)
 *
 * The reference decodes through two-level lookup tables; this restatement decodes the same canonical
 * code bit-serially from per-length counts (the tables are an optimisation of that), which makes the
 * validity rules explicit.  Error strings are the reference's, byte for byte.
 */
#include "zo_oracle.h"
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

typedef struct { const uint8_t *p; size_t n, pos; uint64_t hold; unsigned bits; } bitr;

/* make sure `need` bits are buffered; 0 if the input ran out */
static int need_bits(bitr *b, unsigned need) {
    while (b->bits < need) {
        if (b->pos >= b->n) return 0;
        b->hold |= (uint64_t)b->p[b->pos++] << b->bits; b->bits += 8;
    }
    return 1;
}
static unsigned peek(const bitr *b, unsigned k) { return (unsigned)(b->hold & ((1ull << k) - 1)); }
static void drop(bitr *b, unsigned k) { b->hold >>= k; b->bits -= k; }
static void byte_align(bitr *b) { drop(b, b->bits & 7); }

/* canonical Huffman code described by per-length counts and symbols sorted by (length, symbol) */
typedef struct { uint16_t count[16]; uint16_t sym[320]; int maxlen; } huff;

/* inftrees.c:107-130: returns 0 ok, -1 invalid.  kind 0 = CODES, 1 = LENS, 2 = DISTS */
static int build(huff *h, const uint16_t *lens, int n, int kind) {
    int left = 1, max = 0;
    uint16_t offs[16];
    memset(h->count, 0, sizeof(h->count));
    for (int i = 0; i < n; i++) h->count[lens[i]]++;
    for (int l = 15; l >= 1; l--) if (h->count[l]) { max = l; break; }
    h->maxlen = max;
    if (max == 0) return 0;                       /* no codes: a table of invalid entries, not an error */
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= h->count[l]; if (left < 0) return -1; }   /* over-subscribed */
    if (left > 0 && (kind == 0 || max != 1)) return -1;                                        /* incomplete */
    offs[1] = 0;
    for (int l = 1; l < 15; l++) offs[l + 1] = (uint16_t)(offs[l] + h->count[l]);
    for (int i = 0; i < n; i++) if (lens[i]) h->sym[offs[lens[i]]++] = (uint16_t)i;
    return 0;
}

/* decode one symbol: >= 0 symbol, -1 need more input, -2 invalid code (unused codeword of an incomplete
 * or empty code; the reference's op == 64 table entries) */
static int decode(bitr *b, const huff *h) {
    int code = 0, first = 0, index = 0;
    if (h->maxlen == 0) {                          /* empty table: every entry is {op 64, bits 1} (inftrees.c:122-130) */
        if (!need_bits(b, 1)) return -1;
        return -2;
    }
    for (int len = 1; len <= h->maxlen; len++) {
        if (!need_bits(b, (unsigned)len)) {
            /* zero-extended partial bits could only match a code of length <= bits available: none did */
            return -1;
        }
        code |= (int)((b->hold >> (len - 1)) & 1);
        int cnt = h->count[len];
        if (code - cnt < first) { drop(b, (unsigned)len); return h->sym[index + (code - first)]; }
        index += cnt; first += cnt; first <<= 1; code <<= 1;
    }
    /* incomplete single-code set: the other 1-bit codeword is invalid, consuming its 1 bit */
    return -2;
}

static const uint16_t L_BASE[29] = {3,4,5,6,7,8,9,10,11,13,15,17,19,23,27,31,35,43,51,59,67,83,99,115,131,163,195,227,258};
static const uint8_t  L_EXT[29]  = {0,0,0,0,0,0,0,0,1,1,1,1,2,2,2,2,3,3,3,3,4,4,4,4,5,5,5,5,0};
static const uint16_t D_BASE[30] = {1,2,3,4,5,7,9,13,17,25,33,49,65,97,129,193,257,385,513,769,1025,1537,2049,3073,4097,6145,8193,12289,16385,24577};
static const uint8_t  D_EXT[30]  = {0,0,0,0,1,1,2,2,3,3,4,4,5,5,6,6,7,7,8,8,9,9,10,10,11,11,12,12,13,13};
static const uint8_t  CL_ORDER[19] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};

static huff fixed_l, fixed_d;
static pthread_once_t fixed_once = PTHREAD_ONCE_INIT;
static void make_fixed(void) {                     /* inflate.c:304-309 fixedtables / RFC 1951 3.2.6 */
    uint16_t lens[288];
    for (int i = 0; i < 144; i++) lens[i] = 8;
    for (int i = 144; i < 256; i++) lens[i] = 9;
    for (int i = 256; i < 280; i++) lens[i] = 7;
    for (int i = 280; i < 288; i++) lens[i] = 8;
    build(&fixed_l, lens, 288, 1);
    for (int i = 0; i < 32; i++) lens[i] = 5;
    build(&fixed_d, lens, 32, 2);
}

#define BAD(m) do { *msg = (m); ret = ZO_DATA_ERROR; goto done; } while (0)
#define MORE() do { goto done; } while (0)         /* ran out of input or output: not an error by itself */

int zo_inflate(const uint8_t *in, size_t in_len, int window_bits, uint8_t *out, size_t out_cap,
               size_t *out_len, size_t *in_used, uint32_t *check, const char **msgp) {
    const char *dummy; const char **msg = msgp ? msgp : &dummy;
    *msg = NULL;
    pthread_once(&fixed_once, make_fixed);
    int wrap, wbits = window_bits, ret = ZO_BUF_ERROR;   /* anything that does not reach the end under Z_FINISH */
    size_t o = 0;
    uint32_t chk = 0; int gz = 0, have_trailer = 0;
    bitr b = {in, in_len, 0, 0, 0};
    /* inflate.c:232-251 */
    if (wbits < 0) { if (wbits < -15) return -2; wrap = 0; wbits = -wbits; }
    else { wrap = (wbits >> 4) + 5; if (wbits < 48) wbits &= 15; }
    if (wbits && (wbits < 8 || wbits > 15)) return -2;

    if (wrap) {                                    /* HEAD */
        if (!need_bits(&b, 16)) MORE();
        if ((wrap & 2) && peek(&b, 16) == 0x8b1f) {            /* gzip */
            uint32_t hcrc = zo_crc32(0, in, 2);
            drop(&b, 16);
            if (!need_bits(&b, 16)) MORE();
            unsigned flags = peek(&b, 16);
            if ((flags & 0xff) != 8) BAD("unknown compression method");
            if (flags & 0xe000) BAD("unknown header flags set");
            drop(&b, 16);
            if (!need_bits(&b, 32)) MORE();
            drop(&b, 32);                                       /* TIME */
            if (!need_bits(&b, 16)) MORE();
            drop(&b, 16);                                       /* XFL, OS */
            if (flags & 0x0400) {                               /* EXLEN, EXTRA */
                if (!need_bits(&b, 16)) MORE();
                unsigned xlen = peek(&b, 16); drop(&b, 16);
                if (b.n - b.pos < xlen) { b.pos = b.n; MORE(); }
                b.pos += xlen;
            }
            if (flags & 0x0800) {                               /* NAME */
                for (;;) { if (b.pos >= b.n) MORE(); if (b.p[b.pos++] == 0) break; }
            }
            if (flags & 0x1000) {                               /* COMMENT */
                for (;;) { if (b.pos >= b.n) MORE(); if (b.p[b.pos++] == 0) break; }
            }
            if (flags & 0x0200) {                               /* HCRC over every header byte so far */
                hcrc = zo_crc32(0, in, b.pos);
                if (!need_bits(&b, 16)) MORE();
                if (peek(&b, 16) != (hcrc & 0xffff)) BAD("header crc mismatch");
                drop(&b, 16);
            }
            (void)hcrc;
            gz = 1; chk = 0;
        } else {                                                /* zlib */
            unsigned h = peek(&b, 16);
            if (!(wrap & 1) || (((h & 0xff) << 8) + (h >> 8)) % 31) BAD("incorrect header check");
            if ((h & 0xf) != 8) BAD("unknown compression method");
            unsigned len = ((h >> 4) & 0xf) + 8;
            if (wbits == 0) wbits = (int)len;
            if (len > 15 || len > (unsigned)wbits) BAD("invalid window size");
            chk = 1;
            drop(&b, 16);
            if (h & 0x2000) {                                   /* FDICT (hold & 0x200 before the byte swap view) */
                if (!need_bits(&b, 32)) MORE();
                drop(&b, 32);
                ret = ZO_NEED_DICT; goto done;
            }
        }
    }

    for (int last = 0; !last;) {                   /* TYPEDO */
        if (!need_bits(&b, 3)) MORE();
        last = (int)peek(&b, 1); drop(&b, 1);
        unsigned type = peek(&b, 2); drop(&b, 2);
        if (type == 3) BAD("invalid block type");
        if (type == 0) {                           /* STORED */
            byte_align(&b);
            if (!need_bits(&b, 32)) MORE();
            unsigned v = peek(&b, 16), nv = (unsigned)((b.hold >> 16) & 0xffff);
            if (v != (nv ^ 0xffff)) BAD("invalid stored block lengths");
            drop(&b, 32);
            b.pos -= b.bits >> 3; b.hold = 0; b.bits = 0;          /* (bits is 0 here; keep the byte copy honest) */
            size_t can = v;
            if (can > b.n - b.pos) can = b.n - b.pos;
            if (can > out_cap - o) can = out_cap - o;
            memcpy(out + o, b.p + b.pos, can); o += can; b.pos += can;
            if (can < v) MORE();
            continue;
        }
        huff dl, dd; const huff *hl, *hd;
        if (type == 1) { hl = &fixed_l; hd = &fixed_d; }
        else {                                     /* TABLE */
            if (!need_bits(&b, 14)) MORE();
            unsigned nlen = peek(&b, 5) + 257; drop(&b, 5);
            unsigned ndist = peek(&b, 5) + 1; drop(&b, 5);
            unsigned ncode = peek(&b, 4) + 4; drop(&b, 4);
            if (nlen > 286 || ndist > 30) BAD("too many length or distance symbols");
            uint16_t lens[320]; huff hc;
            memset(lens, 0, sizeof(lens));
            for (unsigned i = 0; i < ncode; i++) { if (!need_bits(&b, 3)) MORE(); lens[CL_ORDER[i]] = (uint16_t)peek(&b, 3); drop(&b, 3); }
            if (build(&hc, lens, 19, 0)) BAD("invalid code lengths set");
            unsigned have = 0;
            uint16_t ll[320];
            while (have < nlen + ndist) {
                int sym;
                if (hc.maxlen == 0) {              /* empty CODES table: entries {bits 1, val 0} read as "length 0" (inflate.c:843-848) */
                    if (!need_bits(&b, 1)) MORE();
                    drop(&b, 1); sym = 0;
                } else {
                    /* the repeat codes need their extra bits before anything is consumed (NEEDBITS(here.bits + k)) */
                    bitr save = b;
                    sym = decode(&b, &hc);
                    if (sym == -1) MORE();
                    if (sym >= 16) {
                        unsigned xb = sym == 16 ? 2 : (sym == 17 ? 3 : 7);
                        if (!need_bits(&b, xb)) { b = save; b.pos = b.n; MORE(); }
                    }
                }
                if (sym < 16) { ll[have++] = (uint16_t)sym; continue; }
                unsigned len = 0, copy;
                if (sym == 16) {
                    if (have == 0) BAD("invalid bit length repeat");
                    len = ll[have - 1]; copy = 3 + peek(&b, 2); drop(&b, 2);
                } else if (sym == 17) { copy = 3 + peek(&b, 3); drop(&b, 3); }
                else { copy = 11 + peek(&b, 7); drop(&b, 7); }
                if (have + copy > nlen + ndist) BAD("invalid bit length repeat");
                while (copy--) ll[have++] = (uint16_t)len;
            }
            if (ll[256] == 0) BAD("invalid code -- missing end-of-block");
            if (build(&dl, ll, (int)nlen, 1)) BAD("invalid literal/lengths set");
            if (build(&dd, ll + nlen, (int)ndist, 2)) BAD("invalid distances set");
            hl = &dl; hd = &dd;
        }
        for (;;) {                                 /* LEN .. MATCH */
            bitr save = b;
            int sym = decode(&b, hl);
            if (sym == -1) MORE();
            if (sym == -2 || sym > 285) BAD("invalid literal/length code");
            if (sym < 256) {
                if (o >= out_cap) { b = save; MORE(); }
                out[o++] = (uint8_t)sym;
                continue;
            }
            if (sym == 256) break;
            unsigned li = (unsigned)sym - 257, len = L_BASE[li];
            if (L_EXT[li]) { if (!need_bits(&b, L_EXT[li])) MORE(); len += peek(&b, L_EXT[li]); drop(&b, L_EXT[li]); }
            int ds = decode(&b, hd);
            if (ds == -1) MORE();
            if (ds == -2 || ds > 29) BAD("invalid distance code");
            unsigned dist = D_BASE[ds];
            if (D_EXT[ds]) { if (!need_bits(&b, D_EXT[ds])) MORE(); dist += peek(&b, D_EXT[ds]); drop(&b, D_EXT[ds]); }
            if (dist > o) BAD("invalid distance too far back");       /* nothing before this call's output (no window yet) */
            for (unsigned k = 0; k < len; k++) {   /* chunkset_tpl.h semantics: byte-serial copy from out - dist */
                if (o >= out_cap) MORE();
                out[o] = out[o - dist]; o++;
            }
        }
    }

    /* CHECK / LENGTH (inflate.c:1109-1151) */
    if (wrap) {
        byte_align(&b);
        if (!need_bits(&b, 32)) MORE();
        uint32_t got = (uint32_t)(b.hold & 0xffffffffu);
        if (gz) { chk = zo_crc32(0, out, o); if (got != chk) BAD("incorrect data check"); }
        else {
            chk = zo_adler32(1, out, o);
            uint32_t be = ((got & 0xff) << 24) | ((got & 0xff00) << 8) | ((got >> 8) & 0xff00) | (got >> 24);
            if (be != chk) BAD("incorrect data check");
        }
        drop(&b, 32);
        if (gz) {
            if (!need_bits(&b, 32)) MORE();
            if ((uint32_t)(b.hold & 0xffffffffu) != (uint32_t)o) BAD("incorrect length check");
            drop(&b, 32);
        }
        have_trailer = 1;
    }
    (void)have_trailer;
    ret = ZO_STREAM_END;
done:
    if (!have_trailer && wrap && ret != ZO_STREAM_END) { chk = gz ? zo_crc32(0, out, o) : zo_adler32(1, out, o); }
    if (out_len) *out_len = o;
    if (in_used) *in_used = b.pos - (b.bits >> 3);            /* unused whole bytes go back (inflate.c RESTORE) */
    if (check) *check = chk;
    return ret;
}

/* ---- member batch with a small pthread pool (mirrors refdrv_inflate_members: zng_inflateInit2(31) streams) ---- */
typedef struct {
    const uint8_t *in; const uint64_t *in_off; size_t n; uint8_t *out; const uint64_t *out_off;
    uint32_t *sizes, *crcs; int32_t *status; atomic_size_t next;
} mjob;

static void *mworker(void *arg) {
    mjob *j = (mjob *)arg;
    for (;;) {
        size_t u = atomic_fetch_add(&j->next, 1);
        if (u >= j->n) break;
        size_t ol = 0; uint32_t c = 0;
        int r = zo_inflate(j->in + j->in_off[u], (size_t)(j->in_off[u + 1] - j->in_off[u]), 31, j->out + j->out_off[u],
                           (size_t)(j->out_off[u + 1] - j->out_off[u]), &ol, NULL, &c, NULL);
        j->status[u] = r; j->sizes[u] = (uint32_t)ol;
        if (j->crcs) j->crcs[u] = c;
    }
    return NULL;
}

int zo_inflate_members(const uint8_t *in, const uint64_t *in_off, size_t n_members, uint8_t *out, const uint64_t *out_off,
                       uint32_t *sizes, uint32_t *crcs, int32_t *status, int nthreads) {
    mjob j = {in, in_off, n_members, out, out_off, sizes, crcs, status, 0};
    atomic_store(&j.next, 0);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads == 1) { mworker(&j); return 0; }
    pthread_t t[256]; int started = 0;
    for (int i = 0; i < nthreads; i++) { if (pthread_create(&t[i], NULL, mworker, &j) == 0) started++; else break; }
    if (!started) mworker(&j);
    for (int i = 0; i < started; i++) pthread_join(t[i], NULL);
    return 0;
}
