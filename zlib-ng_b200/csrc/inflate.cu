// inflate.cu -- K4: batched inflate of independent members (raw / zlib / gzip), one warp per member.
//
// Reference semantics reproduced (files under /root/reference) -- what ONE zng_inflate(strm, Z_FINISH)
// call returns for a member held completely in memory, fresh state, no preset dictionary:
//   inflate.c:219-255      inflateInit2 windowBits decoding (raw < 0, +16 gzip, +32 auto-detect)
//   inflate.c:509-703      HEAD..HCRC: zlib header check, gzip header fields, optional header CRC
//   inflate.c:726-799      TYPEDO / STORED / COPY
//   inflate.c:801-922      TABLE / LENLENS / CODELENS and their error strings
//   inftrees.c:30-295      zng_inflate_table: over-subscribed / incomplete sets (:107-130)
//   inffast_tpl.h:151-298, inflate.c:928-1107   symbol decode, extra bits, match copy, error strings
//   chunkset_tpl.h:112-283 chunkmemset_safe == byte-serial out[i] = out[i - dist]
//   inflate.c:1109-1151    CHECK / LENGTH trailers;  :1176-1200 return value under Z_FINISH
//
// B200 mapping.  Members are independent, symbol decode inside a member is a serial chain, so the unit
// of parallelism is the member = one warp, ~40 members in flight per SM.  The warp runs the block state
// machine uniformly (all lanes hold the same bit buffer; input words are broadcast loads) and uses its
// lanes where the reference uses tables and SIMD:
//   * Huffman decode WITHOUT lookup tables: lane l (1..15) owns code length l (first canonical code,
//     count, symbol offset in registers) and tests whether the next l bits are a length-l codeword; one
//     ballot + ffs + shuffle replaces inflate_table's root/sub-table walk, and building a code for a
//     dynamic block is a 15-step scan instead of filling up to 1924 table entries (inftrees.c:137-291);
//   * match copy = 32 lanes x 1 byte per step with the (k mod dist) rule for overlapping runs;
//   * members whose output fits 4 KiB are assembled in shared memory and leave with coalesced 32-bit
//     stores; their CRC-32 / Adler-32 is computed from shared memory (32 slices + x^(8*after) combine).
#include "common.cuh"
#include "kernels.h"
#include "lz_ops.cuh"

namespace zb {

constexpr int      kInfWarps = 8;          // members in flight per CTA
constexpr uint32_t kInfStage = 4096;       // output capacity up to which a member is assembled in shared memory
constexpr uint32_t kRootBits = 9;          // literal/length codes up to this length decode with one table lookup
constexpr uint32_t kDRootBits = 7;         // same for distance codes
constexpr uint32_t kLitLimit = 256u << 4;  // root entries below this are literals (symbol << 4 | length)
constexpr uint16_t kSlowEntry = 0xffffu;   // root-table entry of a longer (or unused) codeword: ballot decode decides

// error strings: ids are what d_detail carries in its low byte (inflate.c SET_BAD sites)
__host__ const char* inflate_msg(uint32_t id) {
    static const char* const msgs[] = {
        nullptr, "incorrect header check", "unknown compression method", "invalid window size", "unknown header flags set",
        "header crc mismatch", "invalid block type", "invalid stored block lengths", "too many length or distance symbols",
        "invalid code lengths set", "invalid bit length repeat", "invalid code -- missing end-of-block",
        "invalid literal/lengths set", "invalid distances set", "invalid literal/length code", "invalid distance code",
        "invalid distance too far back", "incorrect data check", "incorrect length check"};
    return id < sizeof(msgs) / sizeof(msgs[0]) ? msgs[id] : nullptr;
}
enum : uint32_t {
    E_HEADER = 1, E_METHOD, E_WINDOW, E_FLAGS, E_HCRC, E_BTYPE, E_STORED, E_TOOMANY, E_CODELENS, E_REPEAT, E_NOEOB,
    E_LITLENS, E_DISTS, E_LITCODE, E_DISTCODE, E_FAR, E_CHECK, E_LENGTH
};
constexpr uint32_t kDetailOutFull = 0x100u;     // stopped because the output capacity was reached
constexpr uint32_t kDetailInEnd   = 0x200u;     // stopped because the input ran out

struct InfWarp {
    uint32_t cnt[16];
    uint32_t offs[16];
    uint16_t lsym[288];
    uint16_t dsym[32];
    uint16_t csym[32];
    uint8_t  lens[320];
    uint32_t tfirst[16], tcount[16], toff[16];   // the code under table construction, per length
    uint16_t ltab[1u << kRootBits];              // literal/length root table of the current dynamic block
    uint16_t dtab[1u << kDRootBits];             // distance root table
    alignas(16) uint8_t stage[kInfStage];
};
struct InfShared {
    uint32_t crctab[256];
    uint32_t x2n[32];
    uint16_t fix_lsym[288];
    uint16_t fix_dsym[32];
    uint16_t fix_ltab[1u << kRootBits];
    uint16_t fix_dtab[1u << kDRootBits];
    InfWarp w[kInfWarps];
};

// A canonical Huffman code, spread over the warp: lane l (1..15) keeps the data of code length l.
struct Code { uint32_t first, count, off, maxlen; };

// inftrees.c:107-130.  kind 0 = CODES, 1 = LENS, 2 = DISTS.  Returns 0 ok / -1 invalid set.
__device__ int build_code(const uint8_t* lens, uint32_t n, int kind, uint16_t* sym, InfWarp& P, Code& c, unsigned lane) {
    if (lane < 16u) P.cnt[lane] = 0u;
    __syncwarp();
    for (uint32_t i = lane; i < n; i += 32u) atomicAdd(&P.cnt[lens[i]], 1u);
    __syncwarp();
    const uint32_t my = (lane >= 1u && lane < 16u) ? P.cnt[lane] : 0u;
    const unsigned nz = __ballot_sync(ZB_FULL, my != 0u);
    c.maxlen = nz ? 31u - (uint32_t)__clz(nz) : 0u;
    c.first = 0; c.count = my; c.off = 0;
    if (c.maxlen == 0u) return 0;                           // no codes: every lookup is "invalid code", not an error here
    int left = 1; bool over = false;
    uint32_t code = 0, o = 0;
#pragma unroll
    for (unsigned k = 1; k < 16u; k++) {
        const uint32_t ck = __shfl_sync(ZB_FULL, my, k);
        left = 2 * left - (int)ck;
        over |= left < 0;
        if (lane == k) { c.first = code; c.off = o; }
        code = (code + ck) << 1;
        o += ck;
    }
    if (over) return -1;                                    // over-subscribed
    if (left > 0 && (kind == 0 || c.maxlen != 1u)) return -1;   // incomplete (a single 1-bit distance/length code is allowed)
    if (lane < 16u) P.offs[lane] = c.off;
    __syncwarp();
    const unsigned lt = (1u << lane) - 1u;
    for (uint32_t base = 0; base < n; base += 32u) {        // symbols in (length, symbol) order
        const uint32_t i = base + lane;
        const uint32_t len = i < n ? lens[i] : 0u;
        const unsigned peers = __match_any_sync(ZB_FULL, len);
        const unsigned rank = __popc(peers & lt);
        if (len) sym[P.offs[len] + rank] = (uint16_t)i;
        __syncwarp();
        if (len && rank == 0u) P.offs[len] += __popc(peers);
        __syncwarp();
    }
    return 0;
}

// Next symbol of code c in the low bits of `hold` (`bits` of them are real).  Returns the symbol and its
// code length in `len` (nothing is consumed), -1 if more input is needed first, -2 for an unused codeword.
__device__ __forceinline__ int decode_sym(uint64_t hold, uint32_t bits, const Code& c, const uint16_t* sym, unsigned lane, uint32_t& len) {
    if (c.maxlen == 0u) { len = 1; return bits >= 1u ? -2 : -1; }
    const uint32_t rev = __brev((uint32_t)hold);
    const uint32_t code = lane ? rev >> (32u - lane) : 0u;
    const uint32_t d = code - c.first;
    const unsigned m = __ballot_sync(ZB_FULL, d < c.count);
    if (m == 0u) { len = c.maxlen; return bits >= c.maxlen ? -2 : -1; }
    const unsigned l = (unsigned)__ffs(m) - 1u;
    const uint32_t idx = __shfl_sync(ZB_FULL, c.off + d, l);
    if (l > bits) return -1;
    len = l;
    return (int)sym[idx];
}

// Root table of a literal/length code (the shortcut inflate_table's root table gives the reference, inftrees.c:137-291):
// entry e = the kRootBits next bits of the stream, LSB first; value = symbol << 4 | code length when a codeword of
// <= kRootBits bits starts there, kSlowEntry otherwise.  Every lane resolves 16 entries with the canonical rule.
__device__ void build_root_table(const Code& c, const uint16_t* sym, InfWarp& P, uint16_t* tab, uint32_t root, unsigned lane) {
    if (lane >= 1u && lane < 16u) { P.tfirst[lane] = c.first; P.tcount[lane] = c.count; P.toff[lane] = c.off; }
    __syncwarp();
    const uint32_t top = min(c.maxlen, root);
    for (uint32_t e = lane; e < (1u << root); e += 32u) {
        const uint32_t rev = __brev(e);
        uint16_t ent = kSlowEntry;
        for (uint32_t l = 1; l <= top; l++) {
            const uint32_t d = (rev >> (32u - l)) - P.tfirst[l];
            if (d < P.tcount[l]) { ent = (uint16_t)((sym[P.toff[l] + d] << 4) | l); break; }
        }
        tab[e] = ent;
    }
    __syncwarp();
}

// Bit reader over 4-byte-aligned words; only bytes below `e` (byte offset from w) count as input.
struct Reader {
    const uint32_t* w; uint32_t wi, e; uint64_t hold; uint32_t bits;
    __device__ __forceinline__ void init(const uint32_t* w_, uint32_t bp, uint32_t e_) {
        w = w_; e = e_; hold = 0; bits = 0; wi = (e + 3u) >> 2;     // bp >= e: nothing left, pos() == e
        if (bp < e) {
            wi = bp >> 2;
            const uint32_t sk = bp & 3u;
            uint32_t valid = e - 4u * wi; if (valid > 4u) valid = 4u;
            uint32_t x = __ldg(w + wi);
            if (valid < 4u) x &= (1u << (8u * valid)) - 1u;
            hold = x >> (8u * sk); bits = 8u * (valid - sk); wi++;
        }
    }
    __device__ __forceinline__ void refill() {               // after this: bits >= 32, or every input byte is in `hold`
        if (bits < 32u) {
            if (4u * wi + 4u <= e) { hold |= (uint64_t)__ldg(w + wi) << bits; bits += 32u; wi++; }
            else if (4u * wi < e) {                          // the last, partial word
                const uint32_t valid = e - 4u * wi;
                const uint32_t x = __ldg(w + wi) & ((1u << (8u * valid)) - 1u);
                hold |= (uint64_t)x << bits; bits += 8u * valid; wi++;
            }
        }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t k) const { return (uint32_t)hold & ((1u << k) - 1u); }
    __device__ __forceinline__ void drop(uint32_t k) { hold >>= k; bits -= k; }
    __device__ __forceinline__ uint32_t pos() const { return min(4u * wi, e) - (bits >> 3); }   // first byte not consumed
};

__device__ __forceinline__ uint32_t in_byte(const uint32_t* w, uint32_t bp) { return (__ldg(w + (bp >> 2)) >> (8u * (bp & 3u))) & 0xffu; }

// length / distance bases and extra bits (inftrees.c:52-65 lbase/lext/dbase/dext; RFC 1951 3.2.5)
__device__ __forceinline__ void len_base(uint32_t li, uint32_t& base, uint32_t& ext) {
    if (li < 8u) { base = 3u + li; ext = 0; }
    else if (li == 28u) { base = 258u; ext = 0; }
    else { ext = (li >> 2) - 1u; base = 3u + ((4u + (li & 3u)) << ext); }
}
__device__ __forceinline__ void dist_base(uint32_t ds, uint32_t& base, uint32_t& ext) {
    if (ds < 4u) { base = 1u + ds; ext = 0; }
    else { ext = (ds >> 1) - 1u; base = 1u + ((2u + (ds & 1u)) << ext); }
}

struct InfResult { int32_t ret; uint32_t detail, out_len, in_used, check, bb_byte, bb_bit, bb_out; };

// CRC-32 (gz != 0) or Adler-32 of ob[0..o): 32 contiguous slices, one per lane, then the combine algebra.
__device__ uint32_t warp_check(const uint8_t* ob, uint32_t o, bool gz, const InfShared& S, unsigned lane) {
    // 32 slices of `per` bytes.  per is an ODD number of words: with the output staged in shared memory, a power-of-two slice
    // (4 KiB member: 128 bytes = 32 words) puts the 32 lanes' reads on ONE bank -- a 32-way conflict per byte, 1.1 G conflicts per
    // GiB in the round-1 profile; stride 33 words is conflict free.
    const uint32_t per = (((o + 127u) >> 7) | 1u) << 2;
    const uint32_t beg = min(lane * per, o), end = min(beg + per, o), after = o - end;
    if (gz) {
        uint32_t c = lane == 0u ? 0xffffffffu : 0u;
        for (uint32_t i = beg; i < end; i++) c = S.crctab[(c ^ ob[i]) & 0xffu] ^ (c >> 8);
        if (after && c) c = multmodp(x2nmodp(S.x2n, after, 3), c);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c ^= __shfl_xor_sync(ZB_FULL, c, d);
        return ~c;
    }
    unsigned long long a = 0, b = 0;
    for (uint32_t i = beg; i < end; i++) { const uint32_t by = ob[i]; a += by; b += (unsigned long long)(end - i) * by; }
    // reduce before the position weight: a * after passes 2^64 from ~380 MiB of output on (a <= 255 * per, after < 2^32)
    a %= kAdlerBase; b %= kAdlerBase;
    b += a * (unsigned long long)(after % kAdlerBase);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(ZB_FULL, a, d); b += __shfl_xor_sync(ZB_FULL, b, d); }
    const uint32_t s1 = (uint32_t)((1ull + a) % kAdlerBase), s2 = (uint32_t)((o + b) % kAdlerBase);
    return s1 | (s2 << 16);
}

#define INF_BAD(id) do { R.ret = -3; R.detail = (id); goto done; } while (0)
#define INF_MORE_IN() do { R.detail |= kDetailInEnd; goto done; } while (0)
#define INF_MORE_OUT() do { R.detail |= kDetailOutFull; goto done; } while (0)

// mode bit 0 (segment): the input is one flush-delimited piece of a longer raw stream; running out of input exactly at
// a block boundary is a clean end (ret 0 = Z_OK) instead of Z_BUF_ERROR.  mode bit 1 (count): decode the symbols and add
// up the lengths only -- nothing is stored, the output capacity is unlimited (the size pass of the parallel stream inflate).
// mode bit 3 (resume): the input starts at a block boundary inside a raw deflate stream, `start_bit` bits into its first byte, and
// gout[0 .. out_start) holds the stream's output so far (the history back-references reach into).  No header, no trailer, no staging.
// Besides the usual results the decoder reports the last block boundary it passed -- (bb_byte, bb_bit) into the input, bb_out = output
// length there -- so that a caller fed piecewise can drop what is decoded for good and resume there with more input
// (zng_b200_inflate_stream_feed, what zng_inflate(Z_NO_FLUSH) calls).
__device__ void inflate_member(const uint8_t* in, uint32_t n, int window_bits, uint8_t* gout, uint32_t cap,
                               InfShared& S, InfWarp& P, const Code& FL, const Code& FD, InfResult& R, unsigned lane, int mode,
                               uint32_t start_bit = 0, uint32_t out_start = 0) {
    const bool segment = mode & 1, count = mode & 2, resume = mode & 8;
    if (count) cap = 0xffffffffu;
    const uint32_t sk0 = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 3u);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in - sk0);
    const uint32_t e = sk0 + n;
    const bool staged = cap <= kInfStage && !count && !resume;
    uint8_t* ob = staged ? P.stage : gout;
    uint32_t o = resume ? out_start : 0u, bp = sk0;
    R.bb_byte = 0; R.bb_bit = start_bit; R.bb_out = o;
    bool gz = false;
    Reader b;
    R.ret = -5; R.detail = 0; R.check = 0;                  // anything that does not reach the end under Z_FINISH: Z_BUF_ERROR
    int wrap, wbits = window_bits;
    if (wbits < 0) { wrap = 0; wbits = -wbits; } else { wrap = (wbits >> 4) + 5; if (wbits < 48) wbits &= 15; }
    if (resume) wrap = 0;
    b.init(w, e, e);

    if (wrap) {                                             // HEAD
        if (n < 2u) INF_MORE_IN();
        const uint32_t h = in_byte(w, bp) | (in_byte(w, bp + 1u) << 8);
        if ((wrap & 2) && h == 0x8b1fu) {                   // gzip (inflate.c:521-703)
            if (n < 4u) INF_MORE_IN();
            const uint32_t flags = in_byte(w, bp + 2u) | (in_byte(w, bp + 3u) << 8);
            if ((flags & 0xffu) != 8u) INF_BAD(E_METHOD);
            if (flags & 0xe000u) INF_BAD(E_FLAGS);
            if (n < 10u) INF_MORE_IN();
            bp += 10u;                                      // TIME, XFL, OS
            if (flags & 0x0400u) {
                if (e - bp < 2u) INF_MORE_IN();
                const uint32_t xlen = in_byte(w, bp) | (in_byte(w, bp + 1u) << 8);
                bp += 2u;
                if (e - bp < xlen) INF_MORE_IN();
                bp += xlen;
            }
            if (flags & 0x0800u) { for (;;) { if (bp >= e) INF_MORE_IN(); if (in_byte(w, bp++) == 0u) break; } }
            if (flags & 0x1000u) { for (;;) { if (bp >= e) INF_MORE_IN(); if (in_byte(w, bp++) == 0u) break; } }
            if (flags & 0x0200u) {
                uint32_t c = 0xffffffffu;
                for (uint32_t i = sk0; i < bp; i++) c = S.crctab[(c ^ in_byte(w, i)) & 0xffu] ^ (c >> 8);
                c = ~c;
                if (e - bp < 2u) INF_MORE_IN();
                if ((in_byte(w, bp) | (in_byte(w, bp + 1u) << 8)) != (c & 0xffffu)) INF_BAD(E_HCRC);
                bp += 2u;
            }
            gz = true;
        } else {                                            // zlib (inflate.c:535-569)
            if (!(wrap & 1) || (((h & 0xffu) << 8) + (h >> 8)) % 31u) INF_BAD(E_HEADER);
            if ((h & 0xfu) != 8u) INF_BAD(E_METHOD);
            const uint32_t len = ((h >> 4) & 0xfu) + 8u;
            if (wbits == 0) wbits = (int)len;
            if (len > 15u || len > (uint32_t)wbits) INF_BAD(E_WINDOW);
            R.check = 1u;
            bp += 2u;
            if (h & 0x2000u) {                              // FDICT
                if (e - bp < 4u) INF_MORE_IN();
                bp += 4u;
                R.ret = 2; b.init(w, bp, e); goto done;
            }
        }
    }
    b.init(w, bp, e);
    if (resume && start_bit) { b.refill(); if (b.bits < start_bit) INF_MORE_IN(); b.drop(start_bit); }

    for (bool last = false; !last;) {                       // TYPEDO
        if (resume) {                                       // a block boundary: everything before it is decoded for good
            const uint32_t bits_used = 8u * (min(4u * b.wi, b.e) - sk0) - b.bits;
            R.bb_byte = bits_used >> 3; R.bb_bit = bits_used & 7u; R.bb_out = o;
        }
        b.refill();
        if (segment && b.bits == 0u) { R.ret = 0; goto done; }   // every byte used, at a block boundary
        if (b.bits < 3u) INF_MORE_IN();
        last = b.peek(1) != 0u; b.drop(1);
        const uint32_t type = b.peek(2); b.drop(2);
        if (type == 3u) INF_BAD(E_BTYPE);
        if (type == 0u) {                                   // STORED
            b.drop(b.bits & 7u);
            b.refill();
            if (b.bits < 32u) INF_MORE_IN();
            const uint32_t v = b.peek(16), nv = (uint32_t)(b.hold >> 16) & 0xffffu;
            if (v != (nv ^ 0xffffu)) INF_BAD(E_STORED);
            b.drop(32);
            const uint32_t src = b.pos();
            uint32_t can = min(v, e - src);
            const bool in_short = can < v;
            const bool out_short = can > cap - o;
            can = min(can, cap - o);
            __syncwarp();
            if (!count) for (uint32_t k = lane; k < can; k += 32u) ob[o + k] = (uint8_t)in_byte(w, src + k);
            o += can;
            b.init(w, src + can, e);
            if (out_short) INF_MORE_OUT();
            if (in_short) INF_MORE_IN();
            continue;
        }
        Code cl, cd; const uint16_t* lsym; const uint16_t* dsym; const uint16_t* ltab; const uint16_t* dtab;
        if (type == 1u) { cl = FL; cd = FD; lsym = S.fix_lsym; dsym = S.fix_dsym; ltab = S.fix_ltab; dtab = S.fix_dtab; }
        else {                                              // TABLE
            b.refill();
            if (b.bits < 14u) INF_MORE_IN();
            const uint32_t nlen = b.peek(5) + 257u; b.drop(5);
            const uint32_t ndist = b.peek(5) + 1u; b.drop(5);
            const uint32_t ncode = b.peek(4) + 4u; b.drop(4);
            if (nlen > 286u || ndist > 30u) INF_BAD(E_TOOMANY);
            __syncwarp();
            if (lane < 19u) P.lens[lane] = 0;
            __syncwarp();
            for (uint32_t i = 0; i < ncode; i++) {          // LENLENS, order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15
                b.refill();
                if (b.bits < 3u) INF_MORE_IN();
                // i >= 4: 4->8, 5->7, 6->9, 7->6, 8->10, 9->5, ... 16->14, 17->1, 18->15   (inflate.c:817 order[])
                const uint32_t ord = i < 3u ? 16u + i : (i == 3u ? 0u : ((i & 1u) ? 7u - ((i - 5u) >> 1) : 8u + ((i - 4u) >> 1)));
                if (lane == 0u) P.lens[ord] = (uint8_t)b.peek(3);
                b.drop(3);
            }
            __syncwarp();
            Code cc;
            if (build_code(P.lens, 19u, 0, P.csym, P, cc, lane)) INF_BAD(E_CODELENS);
            const uint32_t total = nlen + ndist;
            uint32_t have = 0, prev = 0;
            __syncwarp();
            while (have < total) {                          // CODELENS
                b.refill();
                uint32_t sym, cb = 0;
                if (cc.maxlen == 0u) {                      // empty CODES table: {bits 1, val 0} entries (inflate.c:843-848)
                    if (b.bits < 1u) INF_MORE_IN();
                    b.drop(1); sym = 0;
                } else {
                    const int s = decode_sym(b.hold, b.bits, cc, P.csym, lane, cb);
                    if (s < 0) INF_MORE_IN();               // a complete code always matches once the bits are there
                    sym = (uint32_t)s;
                    if (sym >= 16u) {
                        const uint32_t xb = sym == 16u ? 2u : (sym == 17u ? 3u : 7u);
                        if (b.bits < cb + xb) INF_MORE_IN();
                    }
                    b.drop(cb);
                }
                if (sym < 16u) { if (lane == 0u) P.lens[have] = (uint8_t)sym; prev = sym; have++; continue; }
                uint32_t len = 0, copy;
                if (sym == 16u) {
                    if (have == 0u) INF_BAD(E_REPEAT);
                    len = prev; copy = 3u + b.peek(2); b.drop(2);
                } else if (sym == 17u) { copy = 3u + b.peek(3); b.drop(3); }
                else { copy = 11u + b.peek(7); b.drop(7); }
                if (have + copy > total) INF_BAD(E_REPEAT);
                for (uint32_t k = lane; k < copy; k += 32u) P.lens[have + k] = (uint8_t)len;
                have += copy; prev = len;
            }
            __syncwarp();
            if (P.lens[256] == 0) INF_BAD(E_NOEOB);
            if (build_code(P.lens, nlen, 1, P.lsym, P, cl, lane)) INF_BAD(E_LITLENS);
            if (build_code(P.lens + nlen, ndist, 2, P.dsym, P, cd, lane)) INF_BAD(E_DISTS);
            __syncwarp();
            lsym = P.lsym; dsym = P.dsym; ltab = P.ltab; dtab = P.dtab;
            build_root_table(cl, lsym, P, P.ltab, kRootBits, lane);
            build_root_table(cd, dsym, P, P.dtab, kDRootBits, lane);
        }
        for (;;) {                                          // LEN .. MATCH
            b.refill();
            uint32_t cb = 0;
            int sym;
            uint32_t ent = ltab[(uint32_t)b.hold & ((1u << kRootBits) - 1u)];
            if (ent < kLitLimit && b.bits >= 3u * kRootBits && o + 3u <= cap) {
                // literal run: up to three literals from the 32 buffered bits, stored by lanes 0..2 in one instruction
                uint32_t h = (uint32_t)b.hold >> (ent & 15u), used = ent & 15u, nlit = 1, lit = ent >> 4;
                const uint32_t e1 = ltab[h & ((1u << kRootBits) - 1u)];
                if (e1 < kLitLimit) {
                    h >>= (e1 & 15u); used += e1 & 15u; nlit = 2; if (lane == 1u) lit = e1 >> 4;
                    const uint32_t e2 = ltab[h & ((1u << kRootBits) - 1u)];
                    if (e2 < kLitLimit) { used += e2 & 15u; nlit = 3; if (lane == 2u) lit = e2 >> 4; }
                }
                if (lane < nlit && !count) ob[o + lane] = (uint8_t)lit;
                b.drop(used);
                o += nlit;
                continue;
            }
            if (ent != kSlowEntry && (ent & 15u) <= b.bits) { sym = (int)(ent >> 4); cb = ent & 15u; }   // one lookup
            else sym = decode_sym(b.hold, b.bits, cl, lsym, lane, cb);
            if (sym == -1) INF_MORE_IN();
            if (sym == -2 || sym > 285) INF_BAD(E_LITCODE);
            if (sym < 256) {
                if (o >= cap) INF_MORE_OUT();
                b.drop(cb);
                if (lane == 0u && !count) ob[o] = (uint8_t)sym;
                o++;
                continue;
            }
            b.drop(cb);
            if (sym == 256) break;
            uint32_t len, lx;
            len_base((uint32_t)sym - 257u, len, lx);
            if (lx) { if (b.bits < lx) INF_MORE_IN(); len += b.peek(lx); b.drop(lx); }
            b.refill();
            int ds;
            ent = dtab[(uint32_t)b.hold & ((1u << kDRootBits) - 1u)];
            if (ent != kSlowEntry && (ent & 15u) <= b.bits) { ds = (int)(ent >> 4); cb = ent & 15u; }
            else ds = decode_sym(b.hold, b.bits, cd, dsym, lane, cb);
            if (ds == -1) INF_MORE_IN();
            if (ds == -2 || ds > 29) INF_BAD(E_DISTCODE);
            b.drop(cb);
            uint32_t dist, dx;
            dist_base((uint32_t)ds, dist, dx);
            if (dx) { if (b.bits < dx) INF_MORE_IN(); dist += b.peek(dx); b.drop(dx); }
            if (dist > o) INF_BAD(E_FAR);                   // nothing precedes this call's output (no window yet)
            const uint32_t can = min(len, cap - o);
            // out[i] = out[i - dist] byte-serially (chunkset_tpl.h): waves of min(32, D) bytes from D back, where D is a
            // multiple of dist that doubles while it is below 32 (every copied wave is one more period of the run)
            if (!count) wave_copy(ob, o, dist, can, lane);  // lz_ops.cuh: the one chunkmemset of this library
            o += can;
            if (can < len) INF_MORE_OUT();
        }
    }

    if (wrap) {                                             // CHECK / LENGTH (inflate.c:1109-1151)
        b.drop(b.bits & 7u);
        b.refill();
        if (b.bits < 32u) INF_MORE_IN();
        const uint32_t got = (uint32_t)b.hold;
        __syncwarp();
        R.check = warp_check(ob, o, gz, S, lane);
        if (gz) { if (got != R.check) INF_BAD(E_CHECK); }
        else if (__byte_perm(got, 0, 0x0123) != R.check) INF_BAD(E_CHECK);
        b.drop(32);
        if (gz) {
            b.refill();
            if (b.bits < 32u) INF_MORE_IN();
            if ((uint32_t)b.hold != o) INF_BAD(E_LENGTH);
            b.drop(32);
        }
    }
    R.ret = 1;
done:
    __syncwarp();
    if (staged && o && !count) {                            // shared-memory image -> the caller's buffer
        if ((reinterpret_cast<uintptr_t>(gout) & 3u) == 0u) {
            const uint32_t* s4 = reinterpret_cast<const uint32_t*>(P.stage);
            uint32_t* g4 = reinterpret_cast<uint32_t*>(gout);
            for (uint32_t k = lane; k < (o >> 2); k += 32u) g4[k] = s4[k];
            for (uint32_t k = (o & ~3u) + lane; k < o; k += 32u) gout[k] = P.stage[k];
        } else {
            for (uint32_t k = lane; k < o; k += 32u) gout[k] = P.stage[k];
        }
    }
    R.out_len = o;
    R.in_used = b.pos() - sk0;
    __syncwarp();
}

__global__ void __launch_bounds__(kInfWarps * 32)
inflate_members_kernel(const uint8_t* __restrict__ in, const uint64_t* __restrict__ in_off, uint32_t n_members, int window_bits,
                       uint8_t* __restrict__ out, const uint64_t* __restrict__ out_off, uint32_t* __restrict__ sizes,
                       uint32_t* __restrict__ checks, int32_t* __restrict__ status, uint32_t* __restrict__ in_used,
                       uint32_t* __restrict__ detail, uint32_t* __restrict__ counter, int mode, uint32_t* __restrict__ resume_io) {
    extern __shared__ __align__(16) unsigned char inf_smem[];
    InfShared& S = *reinterpret_cast<InfShared*>(inf_smem);
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    InfWarp& P = S.w[warp];
    for (uint32_t i = threadIdx.x; i < 256u; i += blockDim.x) S.crctab[i] = crc_table_entry(i, 0);
    if (threadIdx.x == 0) build_x2n(S.x2n);
    // the fixed code (inflate.c:304-309 fixedtables; RFC 1951 3.2.6), built by every warp with the generic builder
    Code FL, FD;
    for (uint32_t i = lane; i < 288u; i += 32u) P.lens[i] = i < 144u ? 8 : (i < 256u ? 9 : (i < 280u ? 7 : 8));
    __syncwarp();
    build_code(P.lens, 288u, 1, warp == 0 ? S.fix_lsym : P.lsym, P, FL, lane);
    __syncwarp();
    if (warp == 0) build_root_table(FL, S.fix_lsym, P, S.fix_ltab, kRootBits, lane);
    __syncwarp();
    P.lens[lane] = 5;
    __syncwarp();
    build_code(P.lens, 32u, 2, warp == 0 ? S.fix_dsym : P.dsym, P, FD, lane);
    __syncwarp();
    if (warp == 0) build_root_table(FD, S.fix_dsym, P, S.fix_dtab, kDRootBits, lane);
    __syncthreads();

    for (;;) {
        uint32_t m = 0;
        if (lane == 0u) m = atomicAdd(counter, 1u);
        m = __shfl_sync(ZB_FULL, m, 0);
        if (m >= n_members) break;
        // mode bit 2: the offset arrays hold (begin, end) pairs per member instead of n + 1 boundaries
        const size_t mi = (mode & 4) ? 2u * (size_t)m : (size_t)m;
        const uint64_t i0 = in_off[mi], i1 = in_off[mi + 1u], o0 = out_off[mi], o1 = out_off[mi + 1u];
        InfResult R;
        // resume mode: resume_io[4m .. 4m+4) = {start_bit, out_start} in, {bb_byte, bb_bit | bb_out handled below} out
        const uint32_t sb = (mode & 8) ? resume_io[4u * m] : 0u, os = (mode & 8) ? resume_io[4u * m + 1u] : 0u;
        inflate_member(in + i0, (uint32_t)(i1 - i0), window_bits, out + o0, (uint32_t)(o1 - o0), S, P, FL, FD, R, lane, mode, sb, os);
        if (lane == 0u) {
            if (mode & 8) { resume_io[4u * m] = R.bb_byte; resume_io[4u * m + 1u] = R.bb_bit; resume_io[4u * m + 2u] = R.bb_out; }
            status[m] = R.ret; sizes[m] = R.out_len;
            if (checks) checks[m] = R.check;
            if (in_used) in_used[m] = R.in_used;
            if (detail) detail[m] = R.detail;
        }
    }
}

cudaError_t launch_inflate_members(const uint8_t* in, const uint64_t* in_off, uint32_t n_members, int window_bits,
                                   uint8_t* out, const uint64_t* out_off, uint32_t* sizes, uint32_t* checks, int32_t* status,
                                   uint32_t* in_used, uint32_t* detail, uint32_t* counter, int num_sms, cudaStream_t stream, int mode,
                                   uint32_t* resume_io) {
    if (n_members == 0) return cudaSuccess;
    if ((mode & 8) && !resume_io) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(inflate_members_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(InfShared));
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    uint32_t grid = (uint32_t)num_sms * 5u;                 // 5 CTAs x 8 warps fit the shared memory of one SM
    const uint32_t need = (n_members + kInfWarps - 1u) / kInfWarps;
    if (grid > need) grid = need;
    inflate_members_kernel<<<grid, kInfWarps * 32, sizeof(InfShared), stream>>>(in, in_off, n_members, window_bits, out, out_off,
                                                                                sizes, checks, status, in_used, detail, counter, mode, resume_io);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- flush-marker scan (inflate.c:1290-1306 syncsearch on the GPU)
// Every byte offset p in [from, n] with in[p-4..p) == 00 00 FF FF (the empty stored block Z_SYNC/FULL_FLUSH leaves,
// deflate.c:1064) is appended to pos[] (unordered, at most cap entries; *count keeps counting).
__global__ void marker_scan_kernel(const uint8_t* __restrict__ in, size_t n, size_t from, unsigned long long* __restrict__ pos,
                                   uint32_t cap, uint32_t* __restrict__ count) {
    const size_t nwords = (n + 3) / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(in);
        const uint32_t w0 = __ldg(w + i), w1 = (i + 1 < nwords) ? __ldg(w + i + 1) : 0u;
#pragma unroll
        for (uint32_t k = 0; k < 4u; k++) {
            const size_t p = i * 4 + k + 4;                 // offset just past the 4 bytes under test
            if (__funnelshift_r(w0, w1, 8u * k) == 0xffff0000u && p <= n && p - 4 >= from) {
                const uint32_t slot = atomicAdd(count, 1u);
                if (slot < cap) pos[slot] = p;
            }
        }
    }
}

cudaError_t launch_marker_scan(const uint8_t* in, size_t n, size_t from, unsigned long long* pos, uint32_t cap, uint32_t* count,
                               int num_sms, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    if (n < 4) return cudaSuccess;
    marker_scan_kernel<<<num_sms * 8, 256, 0, stream>>>(in, n, from, pos, cap, count);
    return cudaGetLastError();
}

}  // namespace zb
