// lz_ops.cuh -- the device-side operators of the match finder, shared by K2 (deflate_fast.cu) and the operator test
// kernels (ops.cu).  Each is the sm_100a form of one entry of the reference's operator surface:
//   vwarp_compare256        functable.compare256      arch/generic/compare256_c.c:12-43   (32 lanes x 8 bytes, ballot/ffs)
//   longest_match_lane<L>   functable.longest_match   match_tpl.h:26-280 with the level-2 / level-3 parameters
//   insert step (in the parsers)  quick_insert_string / insert_string   insert_string_tpl.h:48-104
#pragma once
#include "common.cuh"

namespace zb {

struct VWindow {
    const uint32_t* w;       // 4-byte aligned base of the chunk
    uint32_t skew;           // chunk byte 0 is byte `skew` of w[0]
    const uint32_t* tail;    // image of the words from tw0 on (real bytes below len, then virtual bytes)
    uint32_t tw0;
    __device__ __forceinline__ uint32_t word(uint32_t i) const { return i >= tw0 ? __ldcg(tail + (i - tw0)) : __ldg(w + i); }
};

__device__ __forceinline__ void load12(const VWindow& W, uint32_t pos, uint32_t& v, uint64_t& x) {
    const uint32_t qb = pos + W.skew, i = qb >> 2, sh = (qb & 3u) << 3;
    const uint32_t a0 = W.word(i), a1 = W.word(i + 1), a2 = W.word(i + 2), a3 = W.word(i + 3);
    v = __funnelshift_r(a0, a1, sh);
    x = (uint64_t)__funnelshift_r(a1, a2, sh) | ((uint64_t)__funnelshift_r(a2, a3, sh) << 32);
}

// functable.compare256 (arch/generic/compare256_c.c:12-43), warp wide: first mismatch of the bytes at byte offsets a / b (skew
// included), at most `limit` (<= 256) of them -- lane l compares the 8 bytes at +8l, ballot / ffs finds the first differing
// lane.  Returns the number of equal leading bytes; 256 when all `limit` are equal.  ONE definition for every user: K1's two
// parsers (deflate_quick.cu, deflate_quick_cta.cu), K2 (deflate_fast.cu, lazy_parse.cuh) and the operator kernel (ops.cu).
template <class Win>
__device__ __forceinline__ uint32_t warp_compare_bytes(const Win& W, uint32_t a, uint32_t b, uint32_t limit, unsigned lane) {
    uint64_t x = 0ull;
    if (8u * lane < limit) {
        a += 8u * lane; b += 8u * lane;
        const uint32_t ia = a >> 2, sa = (a & 3u) << 3, ib = b >> 2, sb = (b & 3u) << 3;
        const uint32_t a0 = W.word(ia), a1 = W.word(ia + 1), a2 = W.word(ia + 2);
        const uint32_t b0 = W.word(ib), b1 = W.word(ib + 1), b2 = W.word(ib + 2);
        x = ((uint64_t)__funnelshift_r(a0, a1, sa) | ((uint64_t)__funnelshift_r(a1, a2, sa) << 32)) ^
            ((uint64_t)__funnelshift_r(b0, b1, sb) | ((uint64_t)__funnelshift_r(b1, b2, sb) << 32));
    }
    const unsigned diff = __ballot_sync(ZB_FULL, x != 0ull);
    if (diff == 0u) return 256u;
    const unsigned f = __ffs(diff) - 1u;
    unsigned byte = (unsigned)(__ffsll((long long)x) - 1) >> 3;
    byte = __shfl_sync(ZB_FULL, byte, f);
    return 8u * f + byte;
}
template <class Win>
__device__ __forceinline__ uint32_t vwarp_compare256(const Win& W, uint32_t a, uint32_t b, unsigned lane) {
    return warp_compare_bytes(W, a, b, 256u, lane);
}

// insert_string / quick_insert_string (insert_string_tpl.h:58-104) for the lanes of mask I, lane l holding position q = p + l with
// hash h; `table_head` = head[h] as it stood before this group.  The serial insert order is reproduced: a position's prev[] link
// is the nearest lower inserted lane with its hash (else the table's entry), and the highest inserted lane of a hash owns head[].
// peers = __match_any_sync over the hashes.  Used by K2's parsers and by the insert_string operator kernel.
__device__ __forceinline__ void insert_lanes(uint16_t* head, uint16_t* prev, uint32_t h, uint32_t q, uint32_t p, uint32_t table_head,
                                             unsigned peers, unsigned I, unsigned lane) {
    const unsigned lt = (1u << lane) - 1u;
    if ((I >> lane) & 1u) {
        const unsigned prior = peers & I & lt;
        const uint32_t old = prior ? p + (31u - (uint32_t)__clz(prior)) : table_head;
        if (old != (q & 0xffffu)) __stcg(prev + (q & (kWSize - 1u)), (uint16_t)old);      // insert_string_tpl.h:70-73 head != str
        if ((peers & I & ~lt & ~(1u << lane)) == 0u) __stcg(head + h, (uint16_t)q);
    }
}

// functable.chunkmemset_safe / CHUNKCOPY (chunkset_tpl.h:112-283): out[o + i] = out[o + i - dist] for i in 0..len, byte serial,
// by one warp: waves of min(32, D) bytes from D back, where D is a multiple of dist that doubles while it is below 32 (every
// copied wave is one more period of the run).  Used by K4's match copy and by the chunkmemset operator kernel.
__device__ __forceinline__ void wave_copy(uint8_t* out, uint32_t o, uint32_t dist, uint32_t len, unsigned lane) {
    uint32_t D = dist, rem = len;
    while (rem) {
        __syncwarp();                                       // earlier stores of other lanes -> visible
        const uint32_t wave = min(min(D, 32u), rem);
        if (lane < wave) out[o + lane] = out[o - D + lane];
        o += wave; rem -= wave;
        if (D < 32u) D += D;
    }
}


// Parameters of the hash-chain strategies on this path (configuration_table, deflate.c:142-168): level 2 = deflate_fast
// {nice 8, chain 4}; levels 3-6 = deflate_medium {16,6} {32,24} {32,32} {128,128}.  kCmp = bytes a lane compares by itself;
// a common prefix of kCmp bytes means "kCmp or more", which is >= nice_match and therefore final.
template <int LEVEL> struct LmParams;
template <> struct LmParams<2> { static constexpr uint32_t kNice = 8, kChain = 4, kCmp = 12; };
template <> struct LmParams<3> { static constexpr uint32_t kNice = 16, kChain = 6, kCmp = 16; };
template <> struct LmParams<4> { static constexpr uint32_t kNice = 32, kChain = 24, kCmp = 32; };
template <> struct LmParams<5> { static constexpr uint32_t kNice = 32, kChain = 32, kCmp = 32; };
template <> struct LmParams<6> { static constexpr uint32_t kNice = 128, kChain = 128, kCmp = 128; };

// the 4 bytes at chunk position `pos` (any alignment)
__device__ __forceinline__ uint32_t load32(const VWindow& W, uint32_t pos) {
    const uint32_t b = pos + W.skew, i = b >> 2;
    return __funnelshift_r(W.word(i), W.word(i + 1), (b & 3u) << 3);
}
__device__ __forceinline__ uint32_t load8(const VWindow& W, uint32_t pos) {
    const uint32_t b = pos + W.skew;
    return (W.word(b >> 2) >> ((b & 3u) << 3)) & 0xffu;
}

// common prefix (0..kCmp) of the bytes at the scan position q (the first 16 of them in v, x, z) and at candidate `cand`
template <uint32_t kCmp>
__device__ __forceinline__ uint32_t prefix_len(const VWindow& W, uint32_t q, uint32_t cand, uint32_t v, uint64_t x, uint32_t z) {
    const uint32_t cb = cand + W.skew, i = cb >> 2, sh = (cb & 3u) << 3;
    const uint32_t b0 = W.word(i), b1 = W.word(i + 1), b2 = W.word(i + 2), b3 = W.word(i + 3);
    const uint32_t d0 = v ^ __funnelshift_r(b0, b1, sh);
    if (d0) return (uint32_t)(__ffs((int)d0) - 1) >> 3;
    const uint64_t d = x ^ ((uint64_t)__funnelshift_r(b1, b2, sh) | ((uint64_t)__funnelshift_r(b2, b3, sh) << 32));
    if (d) return 4u + ((uint32_t)(__ffsll((long long)d) - 1) >> 3);
    if (kCmp == 12u) return 12u;
    uint32_t bl = W.word(i + 4);
    const uint32_t dz = z ^ __funnelshift_r(b3, bl, sh);
    if (dz) return 12u + ((uint32_t)(__ffs((int)dz) - 1) >> 3);
    if (kCmp == 16u) return 16u;
    const uint32_t qb = q + W.skew, qi = qb >> 2, qsh = (qb & 3u) << 3;
    uint32_t al = W.word(qi + 4);
    for (uint32_t k = 4; k < kCmp / 4u; k++) {               // words 4.. of both strings, one aligned word of each per step
        const uint32_t an = W.word(qi + k + 1), bn = W.word(i + k + 1);
        const uint32_t dd = __funnelshift_r(al, an, qsh) ^ __funnelshift_r(bl, bn, sh);
        if (dd) return 4u * k + ((uint32_t)(__ffs((int)dd) - 1) >> 3);
        al = an; bl = bn;
    }
    return kCmp;
}

// longest_match for one position q (one lane): v / x / z = the bytes at q, cand0 = hash head (already range-checked),
// look = lookahead (bytes left from q).  Walks <= kChain candidates through prev[].  best_len starts at 2; with
// OPTIMAL_CMP 64 the pre-filter (match_tpl.h:141-165) compares the first 2 / 4 / 8 bytes and the 2 / 4 / 8 bytes that end
// at index best_len, which for best_len 2..15 is exactly "bytes 0..best_len equal": a candidate improves iff its common
// prefix exceeds best_len.  From best_len 16 on (levels 4+) the filter leaves the gap [8, best_len-7): a candidate that
// passes it without improving ends the search below level 5 (early_exit, :127,261-266) and is skipped from level 5 on.
// Returns 0 (no match >= 4), 4..kCmp-1 (exact length, clipped to look), or kCmp = "kCmp or more: measure with
// vwarp_compare256 and clip"; mcand = match_start of the returned match.
// kShared: `prev` is K2c's copy of the 32 Ki-entry prev[] ring in SHARED memory (same index, same links).
__device__ __forceinline__ uint32_t lds_u16(const uint16_t* base, uint32_t i) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"((uint32_t)__cvta_generic_to_shared(base) + 2u * i));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(const uint32_t* base, uint32_t i) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(base) + 4u * i));
    return v;
}
template <int LEVEL, bool kShared = false>
__device__ __forceinline__ uint32_t longest_match_lane(const VWindow& W, uint32_t q, uint32_t v, uint64_t x, uint32_t z, uint32_t cand0,
                                                       uint32_t look, const uint16_t* prev, uint32_t& mcand, uint32_t* raw_best = nullptr) {
    using P = LmParams<LEVEL>;
    constexpr bool kDeep = P::kChain > 6u;                  // worth a 4-byte pre-check at the end of the current best
    uint32_t best = 2, chain = P::kChain, cand = cand0, endw = 0;
    const uint32_t limit = q > kMaxDist ? q - kMaxDist : 0u;
    for (;;) {
        if (kDeep && cand >= q) break;                     // match_tpl.h:131-132 (a re-inserted string can link forward)
        const uint32_t link = kShared ? lds_u16(prev, cand & (kWSize - 1u)) : (uint32_t)__ldcg(prev + (cand & (kWSize - 1u)));   // issued before the compare: the two latencies overlap
        if (!kDeep || best < 3u || load32(W, cand + best - 3u) == endw) {      // bytes best-3..best must match to improve
            const uint32_t cl = prefix_len<P::kCmp>(W, q, cand, v, x, z);
            if (cl > best) {
                mcand = cand;
                if (cl == P::kCmp) { best = P::kCmp; break; }   // >= nice_match: final, measured by the warp later
                if (cl > look) { best = look; break; }         // match_tpl.h:177-183 len > lookahead: return lookahead
                best = cl;
                if (best >= P::kNice) break;                   // nice_match
                if (kDeep) endw = load32(W, q + best - 3u);
            } else if (LEVEL == 4 && best >= 16u && cl >= 8u &&
                       load32(W, cand + best - 7u) == load32(W, q + best - 7u)) break;   // filter passed, no gain: early_exit
        }
        if (--chain == 0u) break;
        cand = link;
        if (cand <= limit) break;                          // match_tpl.h:48-51
    }
    if (raw_best) *raw_best = best;                        // what the reference returns (2 or 3 when nothing useful was found)
    return best >= kWantMin ? best : 0u;
}

}  // namespace zb
