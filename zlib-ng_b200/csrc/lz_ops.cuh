// lz_ops.cuh -- the device-side operators of the match finder, shared by K2 (deflate_fast.cu) and the operator test
// kernels (ops.cu).  Each is the sm_100a form of one entry of the reference's operator surface:
//   vwarp_compare256        functable.compare256      arch/generic/compare256_c.c:12-43   (32 lanes x 8 bytes, ballot/ffs)
//   longest_match_l2_lane   functable.longest_match   match_tpl.h:26-280 at level 2 {good 4, lazy 4, nice 8, chain 4}
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

// warp-wide compare of 256 bytes at byte offsets a / b (skew included): number of equal leading bytes
__device__ __forceinline__ uint32_t vwarp_compare256(const VWindow& W, uint32_t a, uint32_t b, unsigned lane) {
    a += 8u * lane; b += 8u * lane;
    const uint32_t ia = a >> 2, sa = (a & 3u) << 3, ib = b >> 2, sb = (b & 3u) << 3;
    const uint32_t a0 = W.word(ia), a1 = W.word(ia + 1), a2 = W.word(ia + 2);
    const uint32_t b0 = W.word(ib), b1 = W.word(ib + 1), b2 = W.word(ib + 2);
    const uint64_t x = ((uint64_t)__funnelshift_r(a0, a1, sa) | ((uint64_t)__funnelshift_r(a1, a2, sa) << 32)) ^
                       ((uint64_t)__funnelshift_r(b0, b1, sb) | ((uint64_t)__funnelshift_r(b1, b2, sb) << 32));
    const unsigned diff = __ballot_sync(ZB_FULL, x != 0ull);
    if (diff == 0u) return 256u;
    const unsigned f = __ffs(diff) - 1u;
    unsigned byte = (unsigned)(__ffsll((long long)x) - 1) >> 3;
    byte = __shfl_sync(ZB_FULL, byte, f);
    return 8u * f + byte;
}


// longest_match for one position q (one lane): v / x = the 12 bytes at q, cand0 = hash head (already range-checked),
// look = lookahead (bytes left from q).  Walks <= 4 candidates through prev[].  best_len starts at 2 and the
// pre-filter at best_len 2..7 is "bytes 0..best_len equal", so a candidate improves iff its common prefix exceeds
// best_len.  Returns 0 (no match >= 4), 4..11 (exact length, clipped to look), or 12 = "12 or more: measure with
// vwarp_compare256 and clip"; mcand = match_start of the returned match.
__device__ __forceinline__ uint32_t longest_match_l2_lane(const VWindow& W, uint32_t q, uint32_t v, uint64_t x, uint32_t cand0,
                                                          uint32_t look, const uint16_t* prev, uint32_t& mcand) {
    uint32_t best = 2, chain = 4, cand = cand0;
    const uint32_t limit = q > kMaxDist ? q - kMaxDist : 0u;
    for (;;) {
        uint32_t cv; uint64_t cx;
        load12(W, cand, cv, cx);
        const uint32_t d0 = v ^ cv;
        uint32_t cl;
        if (d0) cl = (uint32_t)(__ffs((int)d0) - 1) >> 3;
        else { const uint64_t d = x ^ cx; cl = d ? 4u + ((uint32_t)(__ffsll((long long)d) - 1) >> 3) : 12u; }
        if (cl > best) {
            mcand = cand;
            if (cl == 12u) { best = 12u; break; }     // >= nice_match: final, measured by the warp later
            if (cl > look) { best = look; break; }     // match_tpl.h:177-183 len > lookahead: return lookahead
            best = cl;
            if (best >= 8u) break;                    // nice_match
        }
        if (--chain == 0u) break;
        cand = (uint32_t)__ldcg(prev + (cand & (kWSize - 1u)));
        if (cand <= limit) break;                     // match_tpl.h:48-51
    }
    return best >= kWantMin ? best : 0u;
}

}  // namespace zb
