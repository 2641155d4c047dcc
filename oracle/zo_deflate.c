/* oracle/zo_deflate.c -- TEST INFRASTRUCTURE ONLY (see zo_oracle.h).
 *
 * CPU restatement of the reference's chunk compressor for the frozen parameters
 * (windowBits 15, memLevel 8, Z_DEFAULT_STRATEGY, one zng_deflate(flush) call per <= 65536-byte
 * chunk on a stream whose hash state is empty):
 *
 *   level 1  deflate_quick            deflate_quick.c:47-130
 *   level 2  deflate_fast             deflate_fast.c:19-104
 *   levels 3-6  deflate_medium        deflate_medium.c:22-278
 *            longest_match            match_tpl.h:26-280 (non-SLOW; nice 8, chain 4: deflate.c:142-168)
 *   hashing  quick_insert_string      insert_string_tpl.h:58-75, insert_string.c:13
 *   emission zng_emit_lit/_dist       trees_emit.h:102-164
 *   blocks   zng_tr_flush_block &co   trees.c:106-120,151-173,185-270,280-312,322-405,411-587,625-741
 *   framing  deflate() flush handling deflate.c:1061-1083, zng_tr_stored_block trees.c:592-609
 *
 * The window is NOT slid: positions are kept in original chunk coordinates, which
 * SURVEY.md section 8(a8) shows is equivalent for match finding.  What the slide changes is
 * modelled explicitly: (i) bytes read past the end of the data ("virtual bytes")
 * come from the 64 KiB window image W (W[pos] below 65536, W[pos-32768] above), and
 * (ii) a block that started below 32768 when the slide happened cannot be stored
 * (deflate_p.h:104-112 buf == NULL, trees.c:673).
 *
 * The static code tables are generated from RFC 1951 section 3.2.5/3.2.6 (they are what
 * tools/maketrees.c writes to trees_tbl.h).
 */
#include "zo_oracle.h"
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>

#define ZO_WSIZE     32768u
#define ZO_MAX_DIST  (ZO_WSIZE - 262u)        /* deflate.h:410-415: w_size - MIN_LOOKAHEAD */
#define ZO_MAX_MATCH 258u
#define ZO_WANT_MIN  4u                        /* WANT_MIN_MATCH */
#define ZO_CHUNK_MAX 65536u
#define ZO_SYM_END   16383u                    /* deflate.c:400-403 with memLevel 8, LIT_MEM */
#define ZO_SLIDE_AT  (ZO_WSIZE + ZO_MAX_DIST) /* deflate.c:1285 */

#define L_CODES 286
#define D_CODES 30
#define BL_CODES 19
#define HEAP_SZ (2 * L_CODES + 1)
#define MAX_BITS 15

/* ------------------------------------------------------------------ static tables */
static uint16_t fx_lcode[288]; static uint8_t fx_llen[288];
static uint16_t fx_dcode[30];
static uint8_t  len_sym[256];              /* len-3 -> length code 0..28 (zng_length_code) */
static uint16_t len_base[29]; static uint8_t len_xbits[29];
static uint16_t dst_base[30]; static uint8_t dst_xbits[30];
static uint8_t  dst_sym_lo[256], dst_sym_hi[256];   /* zng_dist_code split in its two halves */
static const uint8_t bl_order_[BL_CODES] = {16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15};  /* RFC 1951 3.2.7 */
static const uint8_t bl_xbits[BL_CODES] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,3,7};
static pthread_once_t tbl_once = PTHREAD_ONCE_INIT;

static unsigned rev_bits(unsigned v, int n) {          /* trees.c:811-818 bi_reverse */
    unsigned r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

static void build_static_tables(void) {
    for (int v = 0; v < 288; v++) {                    /* RFC 1951 3.2.6 */
        unsigned code; int n;
        if (v < 144)      { n = 8; code = 0x30 + v; }
        else if (v < 256) { n = 9; code = 0x190 + (v - 144); }
        else if (v < 280) { n = 7; code = v - 256; }
        else              { n = 8; code = 0xC0 + (v - 280); }
        fx_llen[v] = (uint8_t)n; fx_lcode[v] = (uint16_t)rev_bits(code, n);
    }
    for (int d = 0; d < 30; d++) fx_dcode[d] = (uint16_t)rev_bits((unsigned)d, 5);
    /* length codes: 8 codes without extra bits, then groups of 4 with 1..5 extra bits, then 258 */
    unsigned l = 0;
    for (int c = 0; c < 28; c++) {
        int xb = c < 8 ? 0 : (c - 4) / 4;
        len_base[c] = (uint16_t)l; len_xbits[c] = (uint8_t)xb;
        for (unsigned k = 0; k < (1u << xb); k++) len_sym[l++] = (uint8_t)c;
    }
    len_sym[255] = 28; len_base[28] = 0; len_xbits[28] = 0;     /* trees.c tr_static_init analogue */
    /* distance codes: 4 without extra bits, then pairs with 1..13 extra bits */
    unsigned d = 0;
    for (int c = 0; c < 30; c++) {
        int xb = c < 4 ? 0 : (c - 2) / 2;
        dst_base[c] = (uint16_t)d; dst_xbits[c] = (uint8_t)xb;
        for (unsigned k = 0; k < (1u << xb); k++, d++) {
            if (d < 256) dst_sym_lo[d] = (uint8_t)c;
            if ((d & 127) == 0 && d >= 256) dst_sym_hi[d >> 7] = (uint8_t)c;
        }
    }
}

static inline unsigned dist_sym(unsigned dm1) {        /* deflate.h:436 d_code */
    return dm1 < 256 ? dst_sym_lo[dm1] : dst_sym_hi[dm1 >> 7];
}

/* ------------------------------------------------------------------ bit writer */
typedef struct { uint8_t *p; size_t cap, n; uint64_t acc; unsigned cnt; int ovf; } bitw;

static void bw_put(bitw *b, uint64_t v, unsigned nb) { /* LSB-first, trees_emit.h:42-61 semantics */
    b->acc |= v << b->cnt; b->cnt += nb;
    while (b->cnt >= 8) {
        if (b->n < b->cap) b->p[b->n] = (uint8_t)b->acc; else b->ovf = 1;
        b->n++; b->acc >>= 8; b->cnt -= 8;
    }
}
static void bw_align(bitw *b) { if (b->cnt) bw_put(b, 0, 8 - b->cnt); }   /* bi_windup */
static void bw_raw(bitw *b, const uint8_t *s, size_t k) {
    if (b->n + k <= b->cap) memcpy(b->p + b->n, s, k); else b->ovf = 1;
    b->n += k;
}

/* one literal/length/distance symbol with the given code tables (trees_emit.h:102-164) */
static void emit_lit(bitw *b, const uint16_t *lc, const uint8_t *ll, unsigned c) { bw_put(b, lc[c], ll[c]); }
static void emit_match(bitw *b, const uint16_t *lc, const uint8_t *ll, const uint16_t *dc, const uint8_t *dl,
                       unsigned len, unsigned dist) {
    unsigned lm3 = len - 3, ls = len_sym[lm3];
    uint64_t v = lc[257 + ls]; unsigned nb = ll[257 + ls];
    if (len_xbits[ls]) { v |= (uint64_t)(lm3 - len_base[ls]) << nb; nb += len_xbits[ls]; }
    unsigned dm1 = dist - 1, ds = dist_sym(dm1);
    v |= (uint64_t)dc[ds] << nb; nb += dl[ds];
    if (dst_xbits[ds]) { v |= (uint64_t)(dm1 - dst_base[ds]) << nb; nb += dst_xbits[ds]; }
    bw_put(b, v, nb);
}
static const uint8_t fx_dlen[30] = {5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5};

/* ------------------------------------------------------------------ window model */
typedef struct {
    uint8_t  W[ZO_CHUNK_MAX + 8];   /* window image: data, then zeros or the previous chunk's stale bytes */
    uint32_t len;
    uint16_t head[65536];
    uint16_t prev[ZO_WSIZE];
} lzstate;

static inline uint8_t vb(const lzstate *s, uint32_t pos) {  /* virtual byte, SURVEY 0.6 / deflate.c:1286 */
    return pos < ZO_CHUNK_MAX ? s->W[pos] : s->W[pos - ZO_WSIZE];
}
static inline uint32_t hash4(const lzstate *s, uint32_t pos) {   /* insert_string.c:13 */
    uint32_t v = (uint32_t)vb(s, pos) | ((uint32_t)vb(s, pos + 1) << 8) | ((uint32_t)vb(s, pos + 2) << 16) | ((uint32_t)vb(s, pos + 3) << 24);
    return (v * 2654435761u) >> 16;
}
static inline uint32_t quick_insert(lzstate *s, uint32_t pos) { /* insert_string_tpl.h:58-75 */
    uint32_t h = hash4(s, pos), old = s->head[h];
    if (old != (uint16_t)pos) { s->prev[pos & (ZO_WSIZE - 1)] = (uint16_t)old; s->head[h] = (uint16_t)pos; }
    return old;
}
static inline uint32_t match_run(const lzstate *s, uint32_t a, uint32_t b) { /* compare256(a+2,b+2)+2 */
    uint32_t k = 0;
    while (k < 256 && vb(s, a + 2 + k) == vb(s, b + 2 + k)) k++;
    return k + 2;
}

static void lz_reset(lzstate *s, const uint8_t *in, uint32_t len, const uint8_t *stale) {
    memcpy(s->W, in, len);
    if (len < ZO_CHUNK_MAX) {
        if (stale) { for (uint32_t p = len; p < ZO_CHUNK_MAX; p++) s->W[p] = stale[p & (ZO_WSIZE - 1)]; }
        else memset(s->W + len, 0, ZO_CHUNK_MAX - len);
    }
    memset(s->W + ZO_CHUNK_MAX, 0, 8);
    s->len = len;
    memset(s->head, 0, sizeof(s->head));     /* deflate.c:182-184 CLEAR_HASH: empty slot == position 0 */
    memset(s->prev, 0, sizeof(s->prev));
}

/* ------------------------------------------------------------------ level 1 */
typedef void (*tok_sink)(void *ctx, uint32_t tok);

static void quick_parse(lzstate *s, bitw *b, tok_sink sink, void *ctx) {   /* deflate_quick.c:65-120 */
    uint32_t pos = 0, n = s->len; int slid = 0;
    while (pos < n) {
        uint32_t left = n - pos;
        if (left < 262 && pos >= ZO_SLIDE_AT) slid = 1;       /* fill_window slides (deflate.c:1285-1299) */
        if (left >= ZO_WANT_MIN) {
            uint32_t cand = quick_insert(s, pos);
            /* slide_hash turned every entry below 32768 into 0, and deflate_quick has no hash_head != 0 test: such a
             * slot now names window position 0 = original position 32768 (in reach only for strstart 65274) */
            if (slid && cand < ZO_WSIZE) cand = ZO_WSIZE;
            uint32_t dist = pos - cand;
            if (dist > 0 && dist <= ZO_MAX_DIST && cand < pos &&
                s->W[pos] == s->W[cand] && s->W[pos + 1] == s->W[cand + 1]) {
                uint32_t ml = match_run(s, pos, cand);
                if (ml >= ZO_WANT_MIN) {
                    if (ml > left) ml = left;
                    if (ml > ZO_MAX_MATCH) ml = ZO_MAX_MATCH;
                    if (b) emit_match(b, fx_lcode, fx_llen, fx_dcode, fx_dlen, ml, dist);
                    if (sink) sink(ctx, 0x80000000u | (ml << 16) | dist);
                    pos += ml;
                    continue;
                }
            }
        }
        if (b) emit_lit(b, fx_lcode, fx_llen, s->W[pos]);
        if (sink) sink(ctx, s->W[pos]);
        pos++;
    }
}

/* ------------------------------------------------------------------ level 2: trees.c restated */
typedef struct {
    uint16_t freq[HEAP_SZ], code[HEAP_SZ], len[HEAP_SZ], dad[HEAP_SZ];
    int elems, max_len, max_code;
    const uint8_t *stat_len;      /* static lengths or NULL */
    const uint8_t *xbits; int xbase;
} hufftree;

typedef struct {
    hufftree lt, dt, bt;
    int heap[HEAP_SZ]; int heap_len, heap_max; uint8_t depth[HEAP_SZ];
    uint16_t bl_count[MAX_BITS + 1];
    uint64_t opt_len, static_len;
    uint16_t d_buf[ZO_SYM_END + 1]; uint8_t l_buf[ZO_SYM_END + 1]; uint32_t sym_next;
} blockstate;

#define NODE_LESS(t, n, m, dep) ((t)->freq[n] < (t)->freq[m] || ((t)->freq[n] == (t)->freq[m] && (dep)[n] <= (dep)[m]))

static void sift_down(blockstate *bs, hufftree *t, int k) {      /* trees.c:151-173 */
    int v = bs->heap[k], j = k << 1;
    while (j <= bs->heap_len) {
        if (j < bs->heap_len && NODE_LESS(t, bs->heap[j + 1], bs->heap[j], bs->depth)) j++;
        if (NODE_LESS(t, v, bs->heap[j], bs->depth)) break;
        bs->heap[k] = bs->heap[j]; k = j; j <<= 1;
    }
    bs->heap[k] = v;
}

static void assign_lengths(blockstate *bs, hufftree *t) {        /* trees.c:185-270 gen_bitlen */
    int overflow = 0;
    for (int i = 0; i <= MAX_BITS; i++) bs->bl_count[i] = 0;
    t->len[bs->heap[bs->heap_max]] = 0;
    int h;
    for (h = bs->heap_max + 1; h < HEAP_SZ; h++) {
        int n = bs->heap[h];
        unsigned bits = t->len[t->dad[n]] + 1u;
        if (bits > (unsigned)t->max_len) { bits = (unsigned)t->max_len; overflow++; }
        t->len[n] = (uint16_t)bits;
        if (n > t->max_code) continue;
        bs->bl_count[bits]++;
        int xb = (n >= t->xbase) ? t->xbits[n - t->xbase] : 0;
        bs->opt_len += (uint64_t)t->freq[n] * (bits + (unsigned)xb);
        if (t->stat_len) bs->static_len += (uint64_t)t->freq[n] * ((unsigned)t->stat_len[n] + (unsigned)xb);
    }
    if (!overflow) return;
    do {
        unsigned bits = (unsigned)t->max_len - 1;
        while (bs->bl_count[bits] == 0) bits--;
        bs->bl_count[bits]--; bs->bl_count[bits + 1] += 2; bs->bl_count[t->max_len]--;
        overflow -= 2;
    } while (overflow > 0);
    for (unsigned bits = (unsigned)t->max_len; bits != 0; bits--) {
        int n = bs->bl_count[bits];
        while (n != 0) {
            int m = bs->heap[--h];
            if (m > t->max_code) continue;
            if (t->len[m] != bits) {
                bs->opt_len += (uint64_t)bits * t->freq[m];
                bs->opt_len -= (uint64_t)t->len[m] * t->freq[m];
                t->len[m] = (uint16_t)bits;
            }
            n--;
        }
    }
}

static void assign_codes(blockstate *bs, hufftree *t) {          /* trees.c:280-312 gen_codes */
    uint16_t next[MAX_BITS + 1]; unsigned c = 0;
    for (int bits = 1; bits <= MAX_BITS; bits++) { c = (c + bs->bl_count[bits - 1]) << 1; next[bits] = (uint16_t)c; }
    for (int n = 0; n <= t->max_code; n++) {
        int l = t->len[n];
        if (l) t->code[n] = (uint16_t)rev_bits(next[l]++, l);
    }
}

static void build_huffman(blockstate *bs, hufftree *t) {         /* trees.c:322-405 build_tree */
    int max_code = -1, node;
    bs->heap_len = 0; bs->heap_max = HEAP_SZ;
    for (int n = 0; n < t->elems; n++) {
        if (t->freq[n]) { bs->heap[++bs->heap_len] = max_code = n; bs->depth[n] = 0; }
        else t->len[n] = 0;
    }
    while (bs->heap_len < 2) {
        node = bs->heap[++bs->heap_len] = (max_code < 2 ? ++max_code : 0);
        t->freq[node] = 1; bs->depth[node] = 0;
        bs->opt_len--;
        if (t->stat_len) bs->static_len -= t->stat_len[node];
    }
    t->max_code = max_code;
    for (int n = bs->heap_len / 2; n >= 1; n--) sift_down(bs, t, n);
    node = t->elems;
    do {
        int n = bs->heap[1];
        bs->heap[1] = bs->heap[bs->heap_len--];
        sift_down(bs, t, 1);
        int m = bs->heap[1];
        bs->heap[--bs->heap_max] = n;
        bs->heap[--bs->heap_max] = m;
        t->freq[node] = (uint16_t)(t->freq[n] + t->freq[m]);
        bs->depth[node] = (uint8_t)((bs->depth[n] >= bs->depth[m] ? bs->depth[n] : bs->depth[m]) + 1);
        t->dad[n] = t->dad[m] = (uint16_t)node;
        bs->heap[1] = node++;
        sift_down(bs, t, 1);
    } while (bs->heap_len >= 2);
    bs->heap[--bs->heap_max] = bs->heap[1];
    assign_lengths(bs, t);
    assign_codes(bs, t);
}

/* trees.c:411-519: run-length walk over a code-length array; emit==NULL counts into bt.freq */
static void rle_lengths(blockstate *bs, hufftree *t, int max_code, bitw *emit) {
    int prevlen = -1, nextlen = t->len[0], count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    if (!emit) t->len[max_code + 1] = 0xffff;       /* guard (scan_tree) */
    for (int n = 0; n <= max_code; n++) {
        int curlen = nextlen; nextlen = t->len[n + 1];
        if (++count < max_count && curlen == nextlen) continue;
        if (count < min_count) {
            if (emit) { do bw_put(emit, bs->bt.code[curlen], bs->bt.len[curlen]); while (--count); }
            else bs->bt.freq[curlen] = (uint16_t)(bs->bt.freq[curlen] + count);
        } else if (curlen != 0) {
            if (curlen != prevlen) {
                if (emit) { bw_put(emit, bs->bt.code[curlen], bs->bt.len[curlen]); count--; }
                else bs->bt.freq[curlen]++;
            }
            if (emit) { bw_put(emit, bs->bt.code[16], bs->bt.len[16]); bw_put(emit, (uint64_t)(count - 3), 2); }
            else bs->bt.freq[16]++;
        } else if (count <= 10) {
            if (emit) { bw_put(emit, bs->bt.code[17], bs->bt.len[17]); bw_put(emit, (uint64_t)(count - 3), 3); }
            else bs->bt.freq[17]++;
        } else {
            if (emit) { bw_put(emit, bs->bt.code[18], bs->bt.len[18]); bw_put(emit, (uint64_t)(count - 11), 7); }
            else bs->bt.freq[18]++;
        }
        count = 0; prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

static void block_reset(blockstate *bs) {                        /* trees.c:106-120 init_block */
    memset(bs->lt.freq, 0, sizeof(uint16_t) * L_CODES);
    memset(bs->dt.freq, 0, sizeof(uint16_t) * D_CODES);
    memset(bs->bt.freq, 0, sizeof(uint16_t) * BL_CODES);
    bs->lt.freq[256] = 1;
    bs->opt_len = bs->static_len = 0; bs->sym_next = 0;
}

static void block_init(blockstate *bs) {
    memset(bs, 0, sizeof(*bs));
    bs->lt.elems = L_CODES; bs->lt.max_len = 15; bs->lt.stat_len = fx_llen; bs->lt.xbits = len_xbits; bs->lt.xbase = 257;
    bs->dt.elems = D_CODES; bs->dt.max_len = 15; bs->dt.stat_len = fx_dlen; bs->dt.xbits = dst_xbits; bs->dt.xbase = 0;
    bs->bt.elems = BL_CODES; bs->bt.max_len = 7; bs->bt.stat_len = NULL; bs->bt.xbits = bl_xbits; bs->bt.xbase = 0;
    block_reset(bs);
}

static void write_symbols(blockstate *bs, bitw *b, const uint16_t *lc, const uint8_t *ll,
                          const uint16_t *dc, const uint8_t *dl) {        /* trees.c:708-741 compress_block */
    for (uint32_t i = 0; i < bs->sym_next; i++) {
        if (bs->d_buf[i] == 0) emit_lit(b, lc, ll, bs->l_buf[i]);
        else emit_match(b, lc, ll, dc, dl, bs->l_buf[i] + 3u, bs->d_buf[i]);
    }
    bw_put(b, lc[256], ll[256]);
}

/* trees.c:625-703 zng_tr_flush_block.  raw == NULL means "buf == NULL" (block too old to store). */
static void flush_block(blockstate *bs, bitw *b, const uint8_t *raw, uint32_t stored_len, int last) {
    uint64_t opt_lenb, static_lenb; int max_blindex = 0;
    if (bs->sym_next == 0) {
        opt_lenb = static_lenb = 0; bs->static_len = 7;
    } else {
        build_huffman(bs, &bs->lt);
        build_huffman(bs, &bs->dt);
        rle_lengths(bs, &bs->lt, bs->lt.max_code, NULL);          /* trees.c:525-552 build_bl_tree */
        rle_lengths(bs, &bs->dt, bs->dt.max_code, NULL);
        build_huffman(bs, &bs->bt);
        for (max_blindex = BL_CODES - 1; max_blindex >= 3; max_blindex--)
            if (bs->bt.len[bl_order_[max_blindex]] != 0) break;
        bs->opt_len += 3 * ((uint64_t)max_blindex + 1) + 5 + 5 + 4;
        opt_lenb = (bs->opt_len + 3 + 7) >> 3;
        static_lenb = (bs->static_len + 3 + 7) >> 3;
        if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
    }
    if ((uint64_t)stored_len + 4 <= opt_lenb && raw != NULL) {
        bw_put(b, (uint64_t)(0 << 1) + (unsigned)last, 3);         /* trees.c:592-609 */
        bw_align(b);
        bw_put(b, stored_len & 0xffff, 16); bw_put(b, (~stored_len) & 0xffff, 16);
        bw_raw(b, raw, stored_len);
    } else if (static_lenb == opt_lenb) {
        bw_put(b, (uint64_t)(1 << 1) + (unsigned)last, 3);
        write_symbols(bs, b, fx_lcode, fx_llen, fx_dcode, fx_dlen);
    } else {
        uint8_t ll8[L_CODES], dl8[D_CODES];
        bw_put(b, (uint64_t)(2 << 1) + (unsigned)last, 3);
        int lcodes = bs->lt.max_code + 1, dcodes = bs->dt.max_code + 1, blcodes = max_blindex + 1;
        bw_put(b, (uint64_t)(lcodes - 257), 5); bw_put(b, (uint64_t)(dcodes - 1), 5); bw_put(b, (uint64_t)(blcodes - 4), 4);
        for (int r = 0; r < blcodes; r++) bw_put(b, bs->bt.len[bl_order_[r]], 3);
        rle_lengths(bs, &bs->lt, lcodes - 1, b);
        rle_lengths(bs, &bs->dt, dcodes - 1, b);
        for (int i = 0; i < L_CODES; i++) ll8[i] = (uint8_t)bs->lt.len[i];
        for (int i = 0; i < D_CODES; i++) dl8[i] = (uint8_t)bs->dt.len[i];
        write_symbols(bs, b, bs->lt.code, ll8, bs->dt.code, dl8);
    }
    block_reset(bs);
    if (last) bw_align(b);
}

/* match_tpl.h:26-280 (non-SLOW) restated.  best_len starts at 2 (prev_length is 0 on this path, so good_match never
 * shortens the chain).  With OPTIMAL_CMP 64 the pre-filter (:141-165) compares the first 2 / 4 / 8 bytes and the 2 / 4 / 8
 * bytes that END at index best_len; for best_len 2..15 the two ranges touch or overlap, so the filter is "bytes
 * 0..best_len equal" (SURVEY 8(a6)); from best_len 16 on it leaves the gap [8, best_len-7), and a candidate that passes
 * the filter without beating best_len ends the search on levels below 5 (early_exit, :127,261-266). */
static uint32_t longest_match_lv(const lzstate *s, uint32_t pos, uint32_t cand, uint32_t lookahead, uint32_t chain, uint32_t nice,
                                 int level, uint32_t *mstart) {
    uint32_t best = 2;
    uint32_t limit = pos > ZO_MAX_DIST ? pos - ZO_MAX_DIST : 0;
    for (;;) {
        if (cand >= pos) break;
        uint32_t w = best < 4 ? 2 : (best < 8 ? 4 : 8), off = best + 1 - w;
        int pass = 1;
        for (uint32_t k = 0; k < w && pass; k++)
            if (vb(s, cand + k) != vb(s, pos + k) || vb(s, cand + off + k) != vb(s, pos + off + k)) pass = 0;
        if (pass) {
            uint32_t len = match_run(s, pos, cand);
            if (len > best) {
                *mstart = cand;
                if (len > lookahead) return lookahead;
                best = len;
                if (best >= nice) return best;
            } else if (level < 5) break;
        }
        if (--chain == 0) break;
        cand = s->prev[cand & (ZO_WSIZE - 1)];
        if (cand <= limit) break;
    }
    return best;
}

static uint32_t longest_match_l2(const lzstate *s, uint32_t pos, uint32_t cand, uint32_t lookahead, uint32_t *mstart) {
    return longest_match_lv(s, pos, cand, lookahead, 4, 8, 2, mstart);
}

typedef struct { lzstate lz; blockstate bs; } deflater;

static void tally_lit(blockstate *bs, uint8_t c, tok_sink sink, void *ctx) {
    bs->d_buf[bs->sym_next] = 0; bs->l_buf[bs->sym_next++] = c;
    bs->lt.freq[c]++;
    if (sink) sink(ctx, c);
}
static void tally_match(blockstate *bs, uint32_t dist, uint32_t ml, tok_sink sink, void *ctx) {
    bs->d_buf[bs->sym_next] = (uint16_t)dist; bs->l_buf[bs->sym_next++] = (uint8_t)(ml - 3);
    bs->lt.freq[257 + len_sym[ml - 3]]++; bs->dt.freq[dist_sym(dist - 1)]++;
    if (sink) sink(ctx, 0x80000000u | (ml << 16) | dist);
}

/* level 2: deflate_fast.c:25-103 */
static void fast_parse(deflater *d, bitw *b, int last, tok_sink sink, void *ctx) {
    lzstate *s = &d->lz; blockstate *bs = &d->bs;
    uint32_t pos = 0, n = s->len, block_start = 0; int slid = 0;
    block_init(bs);
    for (;;) {
        if (!slid && pos >= ZO_SLIDE_AT && n - (pos < n ? pos : n) < 262) slid = 1;   /* fill_window runs at lookahead < 262 and slides at strstart >= 65274 (deflate.c:1285) */
        if (pos >= n) break;
        uint32_t left = n - pos, ml = 0, mstart = 0;
        if (left >= ZO_WANT_MIN) {
            uint32_t cand = quick_insert(s, pos);
            uint32_t dist = pos - cand;
            if (dist > 0 && dist <= ZO_MAX_DIST && cand != 0 && cand < pos)
                ml = longest_match_l2(s, pos, cand, left, &mstart);
        }
        if (ml >= ZO_WANT_MIN) {
            tally_match(bs, pos - mstart, ml, sink, ctx);
            left -= ml;
            if (ml <= 4 && left >= ZO_WANT_MIN) {        /* max_insert_length = max_lazy = 4 at level 2 */
                for (uint32_t k = 1; k < ml; k++) quick_insert(s, pos + k);
                pos += ml;
            } else {
                pos += ml;
                quick_insert(s, pos - 1);               /* deflate_fast.c:80 (may hash virtual bytes) */
            }
        } else {
            tally_lit(bs, s->W[pos], sink, ctx);
            pos++;
        }
        if (bs->sym_next == ZO_SYM_END) {
            if (b) flush_block(bs, b, (slid && block_start < ZO_WSIZE) ? NULL : s->W + block_start, pos - block_start, 0);
            else block_reset(bs);
            block_start = pos;
        }
    }
    if (!b) return;
    const uint8_t *raw = (slid && block_start < ZO_WSIZE) ? NULL : s->W + block_start;
    if (last) flush_block(bs, b, raw, pos - block_start, 1);
    else if (bs->sym_next) flush_block(bs, b, raw, pos - block_start, 0);
}

/* levels 3-6: deflate_medium.c.  {nice, chain} = {16,6} {32,24} {32,32} {128,128} (deflate.c:160-163); levels 3-4 run
 * without the look-ahead-one branch (early_exit, deflate_medium.c:151,234), i.e. as a greedy parse. */
typedef struct { uint32_t from, len, at, org; } mmatch;      /* match_start, match_length, strstart, orgstart (:15-20) */
static const uint16_t medium_nice[7]  = {0, 0, 0, 16, 32, 32, 128};
static const uint16_t medium_chain[7] = {0, 0, 0, 6, 24, 32, 128};

static void medium_insert(lzstate *s, mmatch m, uint32_t lookahead) {       /* insert_match, :44-82 */
    if (lookahead <= m.len + ZO_WANT_MIN) return;
    m.at++; m.len--;                                     /* the string at strstart is in the table already */
    if (m.len < ZO_WANT_MIN - 1) {
        if (m.len > 0 && m.at >= m.org) {
            uint32_t cnt = (m.at + m.len - 1 >= m.org) ? m.len : m.org - m.at + 1;
            for (uint32_t k = 0; k < cnt; k++) quick_insert(s, m.at + k);
        }
        return;
    }
    if (m.at >= m.org) {
        uint32_t cnt = (m.at + m.len - 1 >= m.org) ? m.len : m.org - m.at + 1;
        for (uint32_t k = 0; k < cnt; k++) quick_insert(s, m.at + k);
    } else if (m.org < m.at + m.len) {
        for (uint32_t q = m.org; q < m.at + m.len; q++) quick_insert(s, q);
    }
}

static void medium_fizzle(const lzstate *s, mmatch *cur, mmatch *nxt) {     /* fizzle_matches, :84-144 */
    if (cur->len <= 1) return;
    if (cur->len > 1 + nxt->from || cur->len > 1 + nxt->at) return;
    if (s->W[nxt->from + 1 - cur->len] != s->W[nxt->at + 1 - cur->len]) return;   /* the quick exit check */
    mmatch c = *cur, n = *nxt;
    uint32_t limit = nxt->at > ZO_MAX_DIST ? nxt->at - ZO_MAX_DIST : 0;
    int moved = 0;
    while (s->W[n.from - 1] == s->W[n.at - 1]) {         /* pull the next match to the left, shortening the current one */
        if (c.len < 1 || n.at <= limit || n.len >= 256 || n.from <= 1) break;
        n.at--; n.from--; n.len++; c.len--; moved++;
    }
    if (!moved) return;
    if (c.len <= 1 && n.len != 2) { n.org++; *cur = c; *nxt = n; }
}

static void medium_find(lzstate *s, uint32_t pos, uint32_t cand, uint32_t lookahead, int level, mmatch *m) {   /* :191-215 */
    uint32_t dist = pos - cand;
    m->at = m->org = pos;
    if (dist <= ZO_MAX_DIST && dist > 0 && cand != 0 && cand < pos) {
        uint32_t from = 0;
        m->len = longest_match_lv(s, pos, cand, lookahead, medium_chain[level], medium_nice[level], level, &from);
        m->from = from;
        if (m->len < ZO_WANT_MIN || m->from >= pos) m->len = 1;
    } else { m->from = 0; m->len = 1; }
}

static void medium_parse(deflater *d, bitw *b, int last, tok_sink sink, void *ctx, int level) {
    lzstate *s = &d->lz; blockstate *bs = &d->bs;
    uint32_t pos = 0, n = s->len, block_start = 0; int slid = 0;
    const int greedy = level < 5;
    mmatch cur = {0, 0, 0, 0}, nxt = {0, 0, 0, 0};
    block_init(bs);
    for (;;) {
        uint32_t left = n - pos;
        if (left < 262) {                                /* :164-172 fill_window; it slides at strstart >= 65274 */
            if (!slid && pos >= ZO_SLIDE_AT) slid = 1;
            if (left == 0) break;
            nxt.len = 0;
        }
        if (!greedy && nxt.len > 0) { cur = nxt; nxt.len = 0; }
        else medium_find(s, pos, left >= ZO_WANT_MIN ? quick_insert(s, pos) : 0, left, level, &cur);
        medium_insert(s, cur, left);
        if (!greedy && left > 262 && cur.at + cur.len < ZO_CHUNK_MAX - 262) {   /* :234 look ahead one */
            uint32_t np = cur.at + cur.len;
            medium_find(s, np, quick_insert(s, np), left /* s->lookahead is not advanced yet */, level, &nxt);
            if (nxt.len >= ZO_WANT_MIN) medium_fizzle(s, &cur, &nxt);
        } else nxt.len = 0;
        if (cur.len < ZO_WANT_MIN) { for (uint32_t k = 0; k < cur.len; k++) tally_lit(bs, s->W[cur.at + k], sink, ctx); }   /* emit_match, :22-42 */
        else tally_match(bs, cur.at - cur.from, cur.len, sink, ctx);
        pos += cur.len;
        if (bs->sym_next == ZO_SYM_END) {
            if (b) flush_block(bs, b, (slid && block_start < ZO_WSIZE) ? NULL : s->W + block_start, pos - block_start, 0);
            else block_reset(bs);
            block_start = pos;
        }
    }
    if (!b) return;
    const uint8_t *raw = (slid && block_start < ZO_WSIZE) ? NULL : s->W + block_start;
    if (last) flush_block(bs, b, raw, pos - block_start, 1);
    else if (bs->sym_next) flush_block(bs, b, raw, pos - block_start, 0);
}

/* ------------------------------------------------------------------ primed chunks (pigz's dependent mode), level 1
 * Call sequence restated: fresh zng_deflateInit2(1, -15) ; zng_deflateSetDictionary(the 32768 bytes in front of the chunk,
 * deflate.c:456-512) ; one zng_deflate(flush) (deflate_quick.c:47-130 with fill_window, deflate.c:1272-1376).
 * Positions are ABSOLUTE from the start of the dictionary (0 .. 32768 + len), the head table holds 32-bit absolute
 * positions and is never slid: an entry the reference's slide_hash would have zeroed lies more than MAX_DIST behind
 * strstart, and so does the window position 0 an empty slot stands for, so the distance test (:90-92) rejects the same
 * candidates.  What the slides and refills DO change is restated: (i) deflateSetDictionary hashes position 32765 with the
 * still-zero byte behind the dictionary, (ii) the first fill_window of zng_deflate re-inserts 32765 and inserts the two
 * pending strings 32766, 32767 (s->insert), (iii) every later refill that reads input inserts strstart - 1
 * (deflate.c:1321-1336). */
typedef struct { const uint8_t *B; uint32_t N, R; uint32_t *head; } pstate;   /* B[0..N): dictionary + chunk; R: bytes loaded */

static inline uint32_t p_hash(const pstate *s, uint32_t pos) {
    uint32_t v = 0;
    for (int k = 0; k < 4; k++) v |= (uint32_t)(pos + k < s->R ? s->B[pos + k] : 0) << (8 * k);   /* not yet loaded: zeroed (high_water) */
    return (v * 2654435761u) >> 16;
}
static inline uint32_t p_insert(pstate *s, uint32_t pos) {
    uint32_t h = p_hash(s, pos), old = s->head[h];
    if (old != pos) s->head[h] = pos;
    return old;
}

static void quick_parse_primed(pstate *s, uint32_t D, bitw *b, tok_sink sink, void *ctx) {
    uint32_t pos = D, base = 0, pend = 2;
    s->R = D;
    for (uint32_t q = 0; q + 2 < D; q++) p_insert(s, q);          /* insert_string(s, 0, D - 2) */
    for (;;) {
        if (s->R - pos < 262) {                                   /* fill_window */
            do {
                if (pos - base >= ZO_SLIDE_AT) base += ZO_WSIZE;
                if (s->R == s->N) break;
                uint32_t room = base + ZO_CHUNK_MAX - s->R;
                s->R = s->N - s->R < room ? s->N : s->R + room;
                if (s->R - pos + pend >= 3) {
                    uint32_t str = pos - pend;
                    if (str - base >= 1) p_insert(s, str - 1);
                    uint32_t cnt = pend;
                    if (s->R - pos == 1) cnt--;
                    for (uint32_t k = 0; k < cnt; k++) p_insert(s, str + k);
                    pend -= cnt;
                }
            } while (s->R - pos < 262 && s->R != s->N);
            if (s->R == pos) break;
        }
        uint32_t left = s->R - pos;
        if (left >= ZO_WANT_MIN) {
            uint32_t cand = p_insert(s, pos);
            if (cand < base) cand = base;                       /* empty or slid-out slot: window index 0 (no hash_head != 0 test) */
            uint32_t dist = pos - cand;
            if (cand < pos && dist <= ZO_MAX_DIST && s->B[pos] == s->B[cand] && s->B[pos + 1] == s->B[cand + 1]) {
                uint32_t ml = 2;
                while (ml < 258 && pos + ml < s->N && s->B[pos + ml] == s->B[cand + ml]) ml++;
                if (ml >= ZO_WANT_MIN) {
                    if (ml > left) ml = left;
                    if (b) emit_match(b, fx_lcode, fx_llen, fx_dcode, fx_dlen, ml, dist);
                    if (sink) sink(ctx, 0x80000000u | (ml << 16) | dist);
                    pos += ml;
                    continue;
                }
            }
        }
        if (b) emit_lit(b, fx_lcode, fx_llen, s->B[pos]);
        if (sink) sink(ctx, s->B[pos]);
        pos++;
    }
}

/* one primed chunk: in[-dict .. 0) is the dictionary (dict = 32768), in[0 .. len) the chunk */
static size_t deflate_one_primed(const uint8_t *in, uint32_t len, int flush, uint8_t *out, size_t cap) {
    pthread_once(&tbl_once, build_static_tables);
    int last = (flush == ZO_FINISH);
    bitw b = {out, cap, 0, 0, 0, 0};
    pstate s = {in - ZO_WSIZE, ZO_WSIZE + len, 0, (uint32_t *)calloc(65536, sizeof(uint32_t))};
    if (!s.head) return (size_t)-1;
    if (len > 0 || last) {
        bw_put(&b, (uint64_t)(1 << 1) + (unsigned)last, 3);
        quick_parse_primed(&s, ZO_WSIZE, &b, NULL, NULL);
        bw_put(&b, fx_lcode[256], fx_llen[256]);
        if (last) bw_align(&b);
    }
    free(s.head);
    if (!last) { bw_put(&b, 0, 3); bw_align(&b); bw_put(&b, 0x0000, 16); bw_put(&b, 0xffff, 16); }
    return b.ovf ? (size_t)-1 : b.n;
}

/* ------------------------------------------------------------------ window engine: levels 2-6 with a REAL window
 * The restatements above keep positions in chunk coordinates and model the one tail slide of a <= 64 KiB chunk.  A primed
 * chunk (32 KiB dictionary + up to 64 KiB) slides twice and refills in between, and at levels 2-6 the slides are visible
 * (prev[] is indexed modulo 32768, head/prev values are cut to 0, deflate_medium drops its next_match at every refill).  So
 * this engine keeps the reference's own state -- window[], head[], prev[], strstart, lookahead, block_start, insert,
 * high_water -- and restates fill_window (deflate.c:1272-1376), slide_hash (arch/generic/slide_hash_c.c:15-52),
 * deflateSetDictionary (deflate.c:456-512), deflate_fast (deflate_fast.c:19-104) and deflate_medium
 * (deflate_medium.c:22-278) on it.  It is also a second, independent restatement of the un-primed levels 2-6
 * (tests/test_oracle_primed.py compares the two). */
typedef struct {
    uint8_t  win[ZO_CHUNK_MAX + 512];
    uint16_t head[65536], prev[ZO_WSIZE];
    uint32_t strstart, lookahead, insert, high_water, match_start;
    int32_t  block_start;
    const uint8_t *next_in; uint32_t avail_in;
    int level;
} wstate;

static inline uint32_t w_hash(const wstate *s, uint32_t pos) {
    uint32_t v = (uint32_t)s->win[pos] | ((uint32_t)s->win[pos + 1] << 8) | ((uint32_t)s->win[pos + 2] << 16) | ((uint32_t)s->win[pos + 3] << 24);
    return (v * 2654435761u) >> 16;
}
static inline uint32_t w_insert(wstate *s, uint32_t pos) {                 /* insert_string_tpl.h:58-75 */
    uint32_t h = w_hash(s, pos), old = s->head[h];
    if (old != (uint16_t)pos) { s->prev[pos & (ZO_WSIZE - 1)] = (uint16_t)old; s->head[h] = (uint16_t)pos; }
    return old;
}
static void w_fill(wstate *s) {                                             /* fill_window */
    do {
        uint32_t more = ZO_CHUNK_MAX - s->lookahead - s->strstart;
        if (s->strstart >= ZO_SLIDE_AT) {
            memcpy(s->win, s->win + ZO_WSIZE, ZO_WSIZE);
            if (s->match_start >= ZO_WSIZE) s->match_start -= ZO_WSIZE; else s->match_start = 0;
            s->strstart -= ZO_WSIZE;
            s->block_start -= (int32_t)ZO_WSIZE;
            if (s->insert > s->strstart) s->insert = s->strstart;
            for (uint32_t i = 0; i < 65536; i++) s->head[i] = (uint16_t)(s->head[i] >= ZO_WSIZE ? s->head[i] - ZO_WSIZE : 0);
            for (uint32_t i = 0; i < ZO_WSIZE; i++) s->prev[i] = (uint16_t)(s->prev[i] >= ZO_WSIZE ? s->prev[i] - ZO_WSIZE : 0);
            more += ZO_WSIZE;
        }
        if (s->avail_in == 0) break;
        uint32_t n = s->avail_in < more ? s->avail_in : more;
        memcpy(s->win + s->strstart + s->lookahead, s->next_in, n);
        s->next_in += n; s->avail_in -= n; s->lookahead += n;
        if (s->lookahead + s->insert >= 3) {                                /* deflate.c:1321-1336 */
            uint32_t str = s->strstart - s->insert;
            if (str >= 1) w_insert(s, str - 1);
            uint32_t count = s->insert;
            if (s->lookahead == 1) count--;
            for (uint32_t k = 0; k < count; k++) w_insert(s, str + k);
            s->insert -= count;
        }
    } while (s->lookahead < 262 && s->avail_in != 0);
    if (s->high_water < ZO_CHUNK_MAX) {                                     /* deflate.c:1345-1372 */
        uint32_t curr = s->strstart + s->lookahead, init;
        if (s->high_water < curr) {
            init = ZO_CHUNK_MAX - curr; if (init > 258) init = 258;
            memset(s->win + curr, 0, init); s->high_water = curr + init;
        } else if (s->high_water < curr + 258) {
            init = curr + 258 - s->high_water;
            if (init > ZO_CHUNK_MAX - s->high_water) init = ZO_CHUNK_MAX - s->high_water;
            memset(s->win + s->high_water, 0, init); s->high_water += init;
        }
    }
}
static uint32_t w_longest_match(wstate *s, uint32_t cand) {                 /* match_tpl.h:26-280, non-SLOW, prev_length 0 */
    static const uint16_t nice_[7] = {0, 0, 8, 16, 32, 32, 128}, chain_[7] = {0, 0, 4, 6, 24, 32, 128};
    const uint32_t pos = s->strstart, nice = nice_[s->level];
    uint32_t best = 2, chain = chain_[s->level];
    const uint32_t limit = pos > ZO_MAX_DIST ? pos - ZO_MAX_DIST : 0;
    for (;;) {
        if (cand >= pos) break;
        uint32_t w = best < 4 ? 2 : (best < 8 ? 4 : 8), off = best + 1 - w;
        if (memcmp(s->win + cand, s->win + pos, w) == 0 && memcmp(s->win + cand + off, s->win + pos + off, w) == 0) {
            uint32_t len = 2;
            while (len < 258 && s->win[pos + len] == s->win[cand + len]) len++;
            if (len > best) {
                s->match_start = cand;
                if (len > s->lookahead) return s->lookahead;
                best = len;
                if (best >= nice) return best;
            } else if (s->level < 5) break;
        }
        if (--chain == 0) break;
        cand = s->prev[cand & (ZO_WSIZE - 1)];
        if (cand <= limit) break;
    }
    return best;
}
typedef struct { wstate *s; blockstate *bs; bitw *b; tok_sink sink; void *ctx; } wenv;
static void w_flush(wenv *e, int last) {                                   /* FLUSH_BLOCK_ONLY, deflate_p.h:104-112 */
    wstate *s = e->s;
    flush_block(e->bs, e->b, s->block_start >= 0 ? s->win + s->block_start : NULL, (uint32_t)((int32_t)s->strstart - s->block_start), last);
    s->block_start = (int32_t)s->strstart;
}
static void w_deflate_fast(wenv *e, int last) {
    wstate *s = e->s; blockstate *bs = e->bs;
    uint32_t match_len = 0;
    for (;;) {
        if (s->lookahead < 262) { w_fill(s); if (s->lookahead == 0) break; }
        if (s->lookahead >= ZO_WANT_MIN) {
            uint32_t hh = w_insert(s, s->strstart);
            int64_t dist = (int64_t)s->strstart - hh;
            if (dist <= (int64_t)ZO_MAX_DIST && dist > 0 && hh != 0) match_len = w_longest_match(s, hh);
        }
        if (match_len >= ZO_WANT_MIN) {
            tally_match(bs, s->strstart - s->match_start, match_len, e->sink, e->ctx);
            s->lookahead -= match_len;
            if (match_len <= 4 && s->lookahead >= ZO_WANT_MIN) {
                match_len--; s->strstart++;
                for (uint32_t k = 0; k < match_len; k++) w_insert(s, s->strstart + k);
                s->strstart += match_len;
            } else { s->strstart += match_len; w_insert(s, s->strstart - 1); }
            match_len = 0;
        } else { tally_lit(bs, s->win[s->strstart], e->sink, e->ctx); s->lookahead--; s->strstart++; }
        if (bs->sym_next == ZO_SYM_END) w_flush(e, 0);
    }
    if (last) w_flush(e, 1); else if (bs->sym_next) w_flush(e, 0);
}
static void w_find(wstate *s, uint32_t cand, mmatch *m) {                   /* deflate_medium.c:191-215 / :243-262 */
    int64_t dist = (int64_t)s->strstart - cand;
    m->at = m->org = s->strstart;
    if (dist <= (int64_t)ZO_MAX_DIST && dist > 0 && cand != 0) {
        m->len = w_longest_match(s, cand); m->from = s->match_start;
        if (m->len < ZO_WANT_MIN || m->from >= m->at) m->len = 1;
    } else { m->from = 0; m->len = 1; }
}
static void w_insert_match(wstate *s, mmatch m) {                           /* :44-82 */
    if (s->lookahead <= m.len + ZO_WANT_MIN) return;
    m.at++; m.len--;
    if (m.len < ZO_WANT_MIN - 1) {
        if (m.len > 0 && m.at >= m.org) {
            uint32_t cnt = (m.at + m.len - 1 >= m.org) ? m.len : m.org - m.at + 1;
            for (uint32_t k = 0; k < cnt; k++) w_insert(s, m.at + k);
        }
        return;
    }
    if (m.at >= m.org) {
        uint32_t cnt = (m.at + m.len - 1 >= m.org) ? m.len : m.org - m.at + 1;
        for (uint32_t k = 0; k < cnt; k++) w_insert(s, m.at + k);
    } else if (m.org < m.at + m.len) {
        for (uint32_t q = m.org; q < m.at + m.len; q++) w_insert(s, q);
    }
}
static void w_fizzle(const wstate *s, mmatch *cur, mmatch *nxt) {           /* :84-144 */
    if (cur->len <= 1) return;
    if (cur->len > 1 + nxt->from || cur->len > 1 + nxt->at) return;
    if (s->win[nxt->from + 1 - cur->len] != s->win[nxt->at + 1 - cur->len]) return;
    mmatch c = *cur, n = *nxt;
    uint32_t limit = nxt->at > ZO_MAX_DIST ? nxt->at - ZO_MAX_DIST : 0;
    int moved = 0;
    while (s->win[n.from - 1] == s->win[n.at - 1]) {
        if (c.len < 1 || n.at <= limit || n.len >= 256 || n.from <= 1) break;
        n.at--; n.from--; n.len++; c.len--; moved++;
    }
    if (!moved) return;
    if (c.len <= 1 && n.len != 2) { n.org++; *cur = c; *nxt = n; }
}
static void w_deflate_medium(wenv *e, int last) {
    wstate *s = e->s; blockstate *bs = e->bs;
    const int greedy = s->level < 5;
    mmatch cur = {0, 0, 0, 0}, nxt = {0, 0, 0, 0};
    for (;;) {
        if (s->lookahead < 262) { w_fill(s); if (s->lookahead == 0) break; nxt.len = 0; }
        if (!greedy && nxt.len > 0) { cur = nxt; nxt.len = 0; }
        else w_find(s, s->lookahead >= ZO_WANT_MIN ? w_insert(s, s->strstart) : 0, &cur);
        w_insert_match(s, cur);
        if (!greedy && s->lookahead > 262 && cur.at + cur.len < ZO_CHUNK_MAX - 262) {
            s->strstart = cur.at + cur.len;
            w_find(s, w_insert(s, s->strstart), &nxt);
            if (nxt.len >= ZO_WANT_MIN) w_fizzle(s, &cur, &nxt);
            s->strstart = cur.at;
        } else nxt.len = 0;
        if (cur.len < ZO_WANT_MIN) { for (uint32_t k = 0; k < cur.len; k++) { tally_lit(bs, s->win[cur.at + k], e->sink, e->ctx); s->lookahead--; } }
        else { tally_match(bs, cur.at - cur.from, cur.len, e->sink, e->ctx); s->lookahead -= cur.len; }
        s->strstart += cur.len;
        if (bs->sym_next == ZO_SYM_END) w_flush(e, 0);
    }
    if (last) w_flush(e, 1); else if (bs->sym_next) w_flush(e, 0);
}
/* one chunk on a fresh stream, optionally primed with `dict` (32768 bytes); levels 2-6.  sink: the LZ77 tokens as they are tallied */
static size_t window_deflate_one_sink(const uint8_t *dict, const uint8_t *in, uint32_t len, int level, int flush, uint8_t *out, size_t cap,
                                      tok_sink sink, void *ctx) {
    pthread_once(&tbl_once, build_static_tables);
    if (level < 2 || level > 6 || len > ZO_CHUNK_MAX) return (size_t)-1;
    wstate *s = (wstate *)calloc(1, sizeof(wstate));
    blockstate *bs = (blockstate *)malloc(sizeof(blockstate));
    if (!s || !bs) { free(s); free(bs); return (size_t)-1; }
    s->level = level;
    if (dict) {                                                             /* deflateSetDictionary, dictLength == w_size */
        s->next_in = dict; s->avail_in = ZO_WSIZE;
        w_fill(s);
        while (s->lookahead >= 3) {
            uint32_t str = s->strstart, n = s->lookahead - 2;
            for (uint32_t k = 0; k < n; k++) w_insert(s, str + k);
            s->strstart = str + n; s->lookahead = 2;
            w_fill(s);
        }
        s->strstart += s->lookahead; s->block_start = (int32_t)s->strstart; s->insert = s->lookahead; s->lookahead = 0;
    }
    s->next_in = in; s->avail_in = len;
    int last = (flush == ZO_FINISH);
    bitw b = {out, cap, 0, 0, 0, 0};
    block_init(bs);
    wenv e = {s, bs, &b, sink, ctx};
    if (level == 2) w_deflate_fast(&e, last); else w_deflate_medium(&e, last);
    if (!last) { bw_put(&b, 0, 3); bw_align(&b); bw_put(&b, 0x0000, 16); bw_put(&b, 0xffff, 16); }
    free(s); free(bs);
    return b.ovf ? (size_t)-1 : b.n;
}

static size_t window_deflate_one(const uint8_t *dict, const uint8_t *in, uint32_t len, int level, int flush, uint8_t *out, size_t cap) {
    return window_deflate_one_sink(dict, in, len, level, flush, out, cap, NULL, NULL);
}

/* ------------------------------------------------------------------ public */
size_t zo_deflate_bound(size_t n) { return n + (n >> 3) + 64; }

static size_t deflate_one(deflater *d, const uint8_t *in, uint32_t len, int level, int flush,
                          const uint8_t *stale, uint8_t *out, size_t cap) {
    pthread_once(&tbl_once, build_static_tables);
    if (len > ZO_CHUNK_MAX || level < 1 || level > 6) return (size_t)-1;
    if (flush != ZO_SYNC_FLUSH && flush != ZO_FULL_FLUSH && flush != ZO_FINISH) return (size_t)-1;
    int last = (flush == ZO_FINISH);
    bitw b = {out, cap, 0, 0, 0, 0};
    lz_reset(&d->lz, in, len, stale);
    if (level == 1) {
        /* deflate_quick.c:53-63,123-128: a block is opened only if there is input, or for Z_FINISH */
        if (len > 0 || last) {
            bw_put(&b, (uint64_t)(1 << 1) + (unsigned)last, 3);
            quick_parse(&d->lz, &b, NULL, NULL);
            bw_put(&b, fx_lcode[256], fx_llen[256]);
            if (last) bw_align(&b);
        }
    } else {
        if (level == 2) fast_parse(d, &b, last, NULL, NULL); else medium_parse(d, &b, last, NULL, NULL, level);
    }
    if (!last) {                         /* deflate.c:1064-1065: empty stored block for SYNC/FULL flush */
        bw_put(&b, 0, 3); bw_align(&b); bw_put(&b, 0x0000, 16); bw_put(&b, 0xffff, 16);
    }
    return b.ovf ? (size_t)-1 : b.n;
}

size_t zo_deflate_chunk(const uint8_t *in, uint32_t len, int level, int flush, uint8_t *out, size_t cap) {
    deflater *d = (deflater *)malloc(sizeof(deflater));
    if (!d) return (size_t)-1;
    size_t r = deflate_one(d, in, len, level, flush, NULL, out, cap);
    free(d);
    return r;
}

typedef struct { uint32_t *t; size_t cap, n; } tokbuf;
static void tok_push(void *ctx, uint32_t tok) { tokbuf *tb = (tokbuf *)ctx; if (tb->n < tb->cap) tb->t[tb->n] = tok; tb->n++; }

size_t zo_deflate_tokens(const uint8_t *in, uint32_t len, int level, uint32_t *tokens, size_t cap) {
    pthread_once(&tbl_once, build_static_tables);
    if (len > ZO_CHUNK_MAX || level < 1 || level > 6) return (size_t)-1;
    deflater *d = (deflater *)malloc(sizeof(deflater));
    if (!d) return (size_t)-1;
    tokbuf tb = {tokens, cap, 0};
    lz_reset(&d->lz, in, len, NULL);
    if (level == 1) quick_parse(&d->lz, NULL, tok_push, &tb);
    else if (level == 2) fast_parse(d, NULL, 0, tok_push, &tb);
    else medium_parse(d, NULL, 0, tok_push, &tb, level);
    free(d);
    return tb.n;
}

/* the LZ77 tokens of a PRIMED chunk (in[-32768 .. 0) is its dictionary): level 1 through quick_parse_primed, levels 2-6
 * through the window engine -- the trace a kernel under construction is diffed against */
size_t zo_deflate_tokens_primed(const uint8_t *in, uint32_t len, int level, uint32_t *tokens, size_t cap) {
    pthread_once(&tbl_once, build_static_tables);
    if (len > ZO_CHUNK_MAX || level < 1 || level > 6) return (size_t)-1;
    tokbuf tb = {tokens, cap, 0};
    if (level == 1) {
        pstate s = {in - ZO_WSIZE, ZO_WSIZE + len, 0, (uint32_t *)calloc(65536, sizeof(uint32_t))};
        if (!s.head) return (size_t)-1;
        quick_parse_primed(&s, ZO_WSIZE, NULL, tok_push, &tb);
        free(s.head);
    } else {
        size_t bound = zo_deflate_bound(len) + 64;
        uint8_t *scratch = (uint8_t *)malloc(bound);
        if (!scratch) return (size_t)-1;
        size_t r = window_deflate_one_sink(in - ZO_WSIZE, in, len, level, ZO_SYNC_FLUSH, scratch, bound, tok_push, &tb);
        free(scratch);
        if (r == (size_t)-1) return (size_t)-1;
    }
    return tb.n;
}

/* ---- single operators, for the operator-surface tests (functable.longest_match, insert_string) ---- */
/* window: `avail` readable bytes of which the first n are data (avail <= 65536); prev: 32768 entries.
 * Returns what longest_match returns for strstart = pos, cur_match = cand, lookahead = n - pos (match_tpl.h:26-280 with
 * the level-2 parameters); *start = match_start. */
uint32_t zo_longest_match_l2(const uint8_t *window, uint32_t avail, uint32_t n, const uint16_t *prev, uint32_t pos, uint32_t cand, uint32_t *start) {
    lzstate *s = (lzstate *)calloc(1, sizeof(lzstate));
    if (!s || avail > ZO_CHUNK_MAX) { free(s); return 0; }
    memcpy(s->W, window, avail); s->len = n;
    memcpy(s->prev, prev, sizeof(s->prev));
    uint32_t st = 0, r = longest_match_l2(s, pos, cand, n - pos, &st);
    if (start) *start = st;
    free(s);
    return r;
}

/* insert_string(s, str, count) (insert_string_tpl.h:82-104) on caller-owned head[65536] / prev[32768]. */
void zo_insert_string(const uint8_t *window, uint32_t avail, uint16_t *head, uint16_t *prev, uint32_t str, uint32_t count) {
    lzstate *s = (lzstate *)calloc(1, sizeof(lzstate));
    if (!s || avail > ZO_CHUNK_MAX) { free(s); return; }
    memcpy(s->W, window, avail); s->len = avail;
    memcpy(s->head, head, sizeof(s->head)); memcpy(s->prev, prev, sizeof(s->prev));
    for (uint32_t k = 0; k < count; k++) quick_insert(s, str + k);
    memcpy(head, s->head, sizeof(s->head)); memcpy(prev, s->prev, sizeof(s->prev));
    free(s);
}

/* ---- chunk batch with a tiny pthread pool (mirrors refdrv_deflate_chunks) ---- */
typedef struct {
    const uint8_t *in; size_t n; uint32_t chunk; int level, flush;
    uint8_t *out; size_t out_stride; uint32_t *sizes, *crcs, *adlers;
    size_t units; atomic_size_t next; atomic_int err; int primed;
} zjob;

static void *zworker(void *arg) {
    zjob *j = (zjob *)arg;
    deflater *d = (deflater *)malloc(sizeof(deflater));
    if (!d) { atomic_store(&j->err, 1); return NULL; }
    for (;;) {
        size_t u = atomic_fetch_add(&j->next, 1);
        if (u >= j->units) break;
        size_t off = u * (size_t)j->chunk;
        uint32_t len = (uint32_t)((j->n - off < j->chunk) ? (j->n - off) : j->chunk);
        /* a short chunk that follows a full one sees the previous chunk's stale upper half */
        const uint8_t *stale = (len < ZO_CHUNK_MAX && u > 0 && j->chunk == ZO_CHUNK_MAX) ? j->in + off - ZO_WSIZE : NULL;
        size_t r = (j->primed == 2) ? window_deflate_one(NULL, j->in + off, len, j->level, j->flush, j->out + u * j->out_stride, j->out_stride)
                 : (j->primed && j->level >= 2) ? window_deflate_one(u > 0 ? j->in + off - ZO_WSIZE : NULL, j->in + off, len, j->level, j->flush, j->out + u * j->out_stride, j->out_stride)
                 : (j->primed && u > 0) ? deflate_one_primed(j->in + off, len, j->flush, j->out + u * j->out_stride, j->out_stride)
                                        : deflate_one(d, j->in + off, len, j->level, j->flush, j->primed ? NULL : stale, j->out + u * j->out_stride, j->out_stride);
        if (r == (size_t)-1) { atomic_store(&j->err, 2); r = 0; }
        j->sizes[u] = (uint32_t)r;
        if (j->crcs) j->crcs[u] = zo_crc32(0, j->in + off, len);
        if (j->adlers) j->adlers[u] = zo_adler32(1, j->in + off, len);
    }
    free(d);
    return NULL;
}

static int run_zjob_impl(zjob *j, int nthreads);
static int run_zjob(zjob *j, int nthreads) { return run_zjob_impl(j, nthreads); }

int zo_deflate_chunks(const uint8_t *in, size_t n, uint32_t chunk, int level, int flush,
                      uint8_t *out, size_t out_stride, uint32_t *sizes,
                      uint32_t *crcs, uint32_t *adlers, int nthreads) {
    if (chunk == 0 || chunk > ZO_CHUNK_MAX) return -2;
    zjob j = {in, n, chunk, level, flush, out, out_stride, sizes, crcs, adlers, (n + chunk - 1) / chunk, 0, 0, 0};
    return run_zjob(&j, nthreads);
}

/* pigz's dependent mode: chunk u > 0 is compressed by a fresh stream primed (deflateSetDictionary) with the 32768 stream
 * bytes in front of it.  Levels 1-6 (2-6 through the window engine), chunk = 65536 (every dictionary is then a full window). */
int zo_deflate_chunks_primed(const uint8_t *in, size_t n, uint32_t chunk, int level, int flush,
                             uint8_t *out, size_t out_stride, uint32_t *sizes,
                             uint32_t *crcs, uint32_t *adlers, int nthreads) {
    if (chunk != ZO_CHUNK_MAX || level < 1 || level > 6) return -2;
    zjob j = {in, n, chunk, level, flush, out, out_stride, sizes, crcs, adlers, (n + chunk - 1) / chunk, 0, 0, 1};
    return run_zjob(&j, nthreads);
}

/* levels 2-6, every chunk on a FRESH stream, through the window engine (the second restatement; cross-check only) */
int zo_deflate_chunks_fresh_window(const uint8_t *in, size_t n, uint32_t chunk, int level, int flush,
                                   uint8_t *out, size_t out_stride, uint32_t *sizes,
                                   uint32_t *crcs, uint32_t *adlers, int nthreads) {
    if (chunk == 0 || chunk > ZO_CHUNK_MAX || level < 2 || level > 6) return -2;
    zjob j = {in, n, chunk, level, flush, out, out_stride, sizes, crcs, adlers, (n + chunk - 1) / chunk, 0, 0, 2};
    return run_zjob(&j, nthreads);
}

static int run_zjob_impl(zjob *j, int nthreads) {
    atomic_store(&j->next, 0); atomic_store(&j->err, 0);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads == 1) { zworker(j); return atomic_load(&j->err); }
    pthread_t t[256]; int started = 0;
    for (int i = 0; i < nthreads; i++) { if (pthread_create(&t[i], NULL, zworker, j) == 0) started++; else break; }
    if (!started) zworker(j);
    for (int i = 0; i < started; i++) pthread_join(t[i], NULL);
    return atomic_load(&j->err);
}
