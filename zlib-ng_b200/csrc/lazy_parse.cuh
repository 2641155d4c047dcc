// lazy_parse.cuh -- K2a for levels 5 and 6: deflate_medium WITH its look-ahead-one branch (deflate_medium.c:146-278).
//
// What the reference does per iteration: take the match found for the current position (`current_match`, normally the
// `next_match` of the iteration before), insert every position it covers (insert_match, :44-82), look up the longest match
// at the position right behind it (`next_match`, :234-268), let fizzle_matches (:84-144) move the boundary between the two
// to the left when the next match can swallow all but <= 1 byte of the current one, then emit the current match.
//
// Two facts make this parallel on a warp:
//  (1) while the look-ahead runs (lookahead > 262), EVERY position below the one being looked up is in the hash table
//      (literals are inserted when they are looked up, matches insert all they cover), so the match found at a position
//      does not depend on the parse: 32 lanes look up 32 consecutive positions, a lane whose hash also occurs at a lower
//      lane of the window cuts the window there (it would have seen that lane as its head);
//  (2) fizzle_matches keeps at + len of the next match, so the chain of looked-up positions is the plain greedy chain
//      q -> q + len(q) (walk 1 below); only the emitted tokens change, pair by pair (walk 3).
// The last ~550 bytes of a chunk, where the reference stops looking ahead, drops a pending next_match and re-inserts
// strings that are already in the table (:164-172), are parsed by serial_medium(): the same loop, one step at a time, on
// the real head[] / prev[] -- entered at an iteration boundary with exactly the reference's state.
#pragma once
#include "lz_ops.cuh"

namespace zb {

struct MMatch { uint32_t from, len, at, org; };              // match_start, match_length, strstart, orgstart (:15-20)

// fizzle_matches (:84-144) for current = (clen) and next = (nat, nfrom, nlen): how many bytes the next match moves
// left (0 = unchanged).  Warp-cooperative: lane i tests the (base + i + 1)-th step.
__device__ __forceinline__ uint32_t fizzle_moved(const VWindow& W, uint32_t clen, uint32_t nat, uint32_t nfrom, uint32_t nlen, unsigned lane) {
    if (clen <= 1u || clen > 1u + nfrom || clen > 1u + nat) return 0u;
    if (load8(W, nfrom + 1u - clen) != load8(W, nat + 1u - clen)) return 0u;       // the quick exit check
    const uint32_t limit = nat > kMaxDist ? nat - kMaxDist : 0u;
    uint32_t moved = 0;
    for (;;) {
        const uint32_t t = moved + lane + 1u;                 // this lane's step: t - 1 moves are done
        bool ok = clen >= t && nat - (t - 1u) > limit && nlen + (t - 1u) < 256u && nfrom >= t + 1u;
        if (ok) ok = load8(W, nfrom - t) == load8(W, nat - t);
        const unsigned bad = __ballot_sync(ZB_FULL, !ok);
        if (bad) { moved += (uint32_t)__ffs(bad) - 1u; break; }
        moved += 32u;
    }
    if (moved == 0u) return 0u;
    return (clen - moved <= 1u && nlen + moved != 2u) ? moved : 0u;
}

// quick_insert_string (insert_string_tpl.h:58-75) by the whole warp on the real table; returns the old head
__device__ __forceinline__ uint32_t serial_insert(const VWindow& W, uint16_t* head, uint16_t* prev, uint32_t pos, unsigned lane) {
    const uint32_t h = hash4(load32(W, pos));
    const uint32_t old = (uint32_t)__ldcg(head + h);
    if (old != pos && lane == 0) { __stcg(prev + (pos & (kWSize - 1u)), (uint16_t)old); __stcg(head + h, (uint16_t)pos); }
    __syncwarp();
    return old;
}

// :191-215 / :243-262: the match at `pos` given the old head `cand`; every lane computes the same thing
template <int LEVEL>
__device__ __forceinline__ void serial_find(const VWindow& W, const uint16_t* prev, uint32_t pos, uint32_t cand, uint32_t look, unsigned lane, MMatch& m) {
    using P = LmParams<LEVEL>;
    m.at = m.org = pos; m.from = 0; m.len = 1;
    if (cand != 0u && cand < pos && pos - cand <= kMaxDist) {
        uint32_t v, z; uint64_t x;
        {
            const uint32_t qb = pos + W.skew, i = qb >> 2, sh = (qb & 3u) << 3;
            const uint32_t a0 = W.word(i), a1 = W.word(i + 1), a2 = W.word(i + 2), a3 = W.word(i + 3);
            v = __funnelshift_r(a0, a1, sh);
            x = (uint64_t)__funnelshift_r(a1, a2, sh) | ((uint64_t)__funnelshift_r(a2, a3, sh) << 32);
            z = __funnelshift_r(a3, W.word(i + 4), sh);
        }
        uint32_t from = 0;
        uint32_t len = longest_match_lane<LEVEL>(W, pos, v, x, z, cand, look, prev, from);
        if (len >= P::kCmp) {
            len = P::kCmp + vwarp_compare256(W, pos + P::kCmp + W.skew, from + P::kCmp + W.skew, lane);
            len = min(min(len, kMaxMatch), look);
        }
        if (len >= kWantMin && from < pos) { m.len = len; m.from = from; }
    }
}

// deflate_medium.c:146-278 one iteration at a time, from an iteration boundary: strstart = pos, next_match = nxt when
// have_nxt.  Tokens are appended at tok[wr..]; returns the new count.
template <int LEVEL>
__device__ uint32_t serial_medium(const VWindow& W, uint32_t n, uint16_t* head, uint16_t* prev, uint32_t* __restrict__ tok,
                                  uint32_t wr, uint32_t pos, MMatch nxt, bool have_nxt, unsigned lane) {
    MMatch cur{0, 0, 0, 0};
    for (;;) {
        const uint32_t left = n - pos;
        if (left < 262u) {                                   // :164-172 fill_window finds no more input
            if (left == 0u) break;
            have_nxt = false;
        }
        if (have_nxt) { cur = nxt; have_nxt = false; }
        else serial_find<LEVEL>(W, prev, pos, left >= kWantMin ? serial_insert(W, head, prev, pos, lane) : 0u, left, lane, cur);
        if (left > cur.len + kWantMin && cur.len >= 2u) {     // insert_match (:44-82): cur.len is 1 or >= 4 here
            uint32_t lo = cur.at + 1u;
            const uint32_t hi = cur.at + cur.len;             // the string at cur.at is in the table already
            if (lo < cur.org) lo = cur.org;
            for (uint32_t r = lo; r < hi; r++) serial_insert(W, head, prev, r, lane);
        }
        if (left > 262u && cur.at + cur.len < kChunkMax - 262u) {          // :234 look ahead one
            const uint32_t np = cur.at + cur.len;
            serial_find<LEVEL>(W, prev, np, serial_insert(W, head, prev, np, lane), left, lane, nxt);
            have_nxt = true;
            if (nxt.len >= kWantMin) {
                const uint32_t moved = fizzle_moved(W, cur.len, nxt.at, nxt.from, nxt.len, lane);
                if (moved) { cur.len -= moved; nxt.at -= moved; nxt.from -= moved; nxt.len += moved; nxt.org++; }
            }
        }
        if (cur.len != 0u) {                                 // emit_match (:22-42)
            if (lane == 0) __stcs(tok + wr, cur.len >= kWantMin ? (kTokMatch | (cur.len << 16) | (cur.at - cur.from)) : load8(W, cur.at));
            wr++;
        }
        pos += cur.len;
    }
    return wr;
}

// K2c (deflate_fast.cu: medium_cta_kernel): the windows below `F` were inserted by the CTA's inserter warp and searched by its searcher
// warps; this warp only replays the decisions.  ring[q % kCoopRing] = match length | match start << 16 of position q, valid once
// ready[w % kCoopReady] == w + 1 (w = q / 32); res_win tells the searchers which ring slots are free again.
constexpr uint32_t kLazyGuard = 32u + kMaxMatch + 262u + 2u;   // a window starting here or later may leave the look-ahead regime
constexpr uint32_t kCoopRing = 512, kCoopReady = kCoopRing / 32u;   // 16 windows; at most 8 are in flight (kCoopAhead)
constexpr uint32_t kCoopAhead = 7;
struct CoopView { const uint32_t* ring; volatile uint32_t* ready; volatile uint32_t* res_win; volatile uint32_t* ins_upto; uint32_t F; };
__device__ __forceinline__ void coop_wait(volatile uint32_t* word, uint32_t at_least, uint32_t ns = 400) {   // bounded: a protocol bug traps, it does not hang
    for (uint32_t tries = 0; *word < at_least; tries++) { __nanosleep(ns); if (tries > (1u << 24)) __trap(); }   // (long naps: a spinning warp costs its neighbours issue slots)
}

// Parse one chunk at level 5 or 6; tokens to tok[0..count) + end marker.  Returns the token count.
template <int LEVEL, bool kCoop = false>
__device__ uint32_t lazy_parse_warp(const VWindow W, uint32_t n, uint16_t* head, uint16_t* prev, uint32_t* __restrict__ tok, const CoopView* cv = nullptr) {
    using P = LmParams<LEVEL>;
    constexpr uint32_t kCmp = P::kCmp;
    constexpr uint32_t kGuard = kLazyGuard;
    const unsigned lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t kNoClip = 0x7fffffffu;                     // s->lookahead during a look-ahead search is > 262: it never clips
    uint32_t wr = 0, p = 0, skip = 0;
    bool pm_valid = false; int pm_lane = -1;                  // the match whose successor is not known yet (current_match)
    uint32_t pm_at = 0, pm_from = 0, pm_len = 0;
    if (n < kGuard) {
        wr = serial_medium<LEVEL>(W, n, head, prev, tok, 0u, 0u, MMatch{0, 0, 0, 0}, false, lane);
        if (lane == 0) __stcs(tok + wr, kTokEnd);
        return wr;
    }
    for (;;) {
        const bool endzone = p + kGuard > n;
        const bool pre = kCoop && !endzone;                   // K2c: this window is inserted and searched already (p is a multiple of 32, p < F)
        if (kCoop) {
            if (pre) {          // (also for a window a long match covers entirely: its searcher must be done before its ring slots are given away)
                if (lane == 0) { const uint32_t w = p >> 5; for (uint32_t tries = 0; cv->ready[w & (kCoopReady - 1u)] != w + 1u; tries++) { __nanosleep(64); if (tries > (1u << 26)) __trap(); } }
            } else if (!pre) { if (lane == 0) coop_wait(cv->ins_upto, cv->F); }
            __syncwarp();
            __threadfence_block();
        }
        const uint32_t q = p + lane;
        const bool inb = q + kWantMin <= n;                   // only the end zone has lanes past the data; they are never used
        uint32_t v, z; uint64_t x;
        {
            const uint32_t qb = q + W.skew, i = qb >> 2, sh = (qb & 3u) << 3;
            const uint32_t a0 = W.word(i), a1 = W.word(i + 1), a2 = W.word(i + 2), a3 = W.word(i + 3);
            v = __funnelshift_r(a0, a1, sh);
            x = (uint64_t)__funnelshift_r(a1, a2, sh) | ((uint64_t)__funnelshift_r(a2, a3, sh) << 32);
            z = __funnelshift_r(a3, W.word(i + 4), sh);
        }
        const uint32_t h = hash4(v);
        const uint32_t cand0 = (inb && !pre) ? (uint32_t)__ldcg(head + h) : 0u;
        uint32_t mlen = 0, mcand = 0;                        // mlen: 0 none, 4..kCmp-1 exact, kCmp = "kCmp or more"
        if (pre) {
            if (lane >= skip) { const uint32_t r = lds_u32(cv->ring, q & (kCoopRing - 1u)); mlen = r & 0xffffu; mcand = r >> 16; }
        } else if (inb && lane >= skip && (!endzone || lane == skip) && cand0 != 0u && (q - cand0 - 1u) < kMaxDist)
            mlen = longest_match_lane<LEVEL>(W, q, v, x, z, cand0, kNoClip, prev, mcand);
        const unsigned peers = pre ? 0u : __match_any_sync(ZB_FULL, inb ? h : (0x10000u + lane));   // (pre: the links are exact, nothing is stale)
        const unsigned M = __ballot_sync(ZB_FULL, mlen != 0u);
        // ---- walk 1: the greedy chain of looked-up positions
        unsigned cur = skip, covered = 0;
        const unsigned wend = endzone ? min(skip + 1u, 32u) : 32u;      // the end zone stops at its first looked-up position
        while (cur < wend) {
            const unsigned rest = M & ~lane_range(0, cur);
            if (rest == 0u) { cur = wend; break; }
            const unsigned k = (unsigned)(__ffs(rest) - 1);
            if (k >= wend) { cur = wend; break; }
            uint32_t len = __shfl_sync(ZB_FULL, mlen, k);
            if (len >= kCmp) {
                const uint32_t ck = __shfl_sync(ZB_FULL, mcand, k);
                len = min(kCmp + vwarp_compare256(W, p + k + kCmp + W.skew, ck + kCmp + W.skew, lane), kMaxMatch);
                if (lane == k) mlen = len;
            }
            covered |= lane_range(k + 1u, k + len);
            cur = k + len;
        }
        unsigned V = lane_range(skip, min(cur, wend)) & ~covered;
        unsigned I = lane_range(0, min(cur, wend));            // every position up to the last looked-up one is inserted
        // ---- walk 2: a looked-up lane whose hash was inserted by a lower lane of this window saw a stale head
        const unsigned S = __ballot_sync(ZB_FULL, ((V >> lane) & 1u) && (peers & I & lt) != 0u);
        uint32_t next_p, next_skip;
        if (S) {
            const unsigned j = __ffs(S) - 1u;
            V &= lane_range(0, j); I &= lane_range(0, j);
            next_p = p + j; next_skip = 0;
        } else if (cur > wend) { next_p = p + wend; next_skip = cur - wend; }
        else { next_p = p + wend; next_skip = 0; }
        // ---- walk 3: fizzle_matches between each match and the match right behind it; the tokens of this window
        unsigned E = V;                                       // lanes that emit a token now
        uint32_t moved_me = 0, flen = mlen;                   // this lane's match: pulled left by moved_me, final length flen
        uint32_t carry_tok = 0; bool carry_emit = false;
        const unsigned VM = V & M;
        unsigned scan_from = 0;
        for (;;) {
            if (!pm_valid) {
                const unsigned rest = VM & ~lane_range(0, scan_from);
                if (rest == 0u) break;
                const unsigned k = (unsigned)(__ffs(rest) - 1);
                pm_valid = true; pm_lane = (int)k; pm_at = p + k;
                pm_from = __shfl_sync(ZB_FULL, mcand, k); pm_len = __shfl_sync(ZB_FULL, mlen, k);
            }
            const uint32_t np = pm_at + pm_len;               // the position looked up behind it (fizzle keeps at + len)
            if (np < p + 32u && ((V >> (np - p)) & 1u)) {
                const unsigned r = np - p;
                uint32_t new_len = pm_len, moved = 0, nfrom = 0, nlen = 0;
                const bool next_is_match = (M >> r) & 1u;
                if (next_is_match) {
                    nfrom = __shfl_sync(ZB_FULL, mcand, r); nlen = __shfl_sync(ZB_FULL, mlen, r);
                    moved = fizzle_moved(W, pm_len, np, nfrom, nlen, lane);
                    new_len = pm_len - moved;
                }
                if (pm_lane >= 0) { if ((int)lane == pm_lane) flen = new_len; if (new_len == 0u) E &= ~(1u << pm_lane); }
                else if (new_len != 0u) { carry_emit = true; carry_tok = new_len >= kWantMin ? (kTokMatch | (new_len << 16) | (pm_at - pm_from)) : load8(W, pm_at); }
                if (next_is_match) {
                    if (lane == r) moved_me = moved;
                    pm_lane = (int)r; pm_at = np - moved; pm_from = nfrom - moved; pm_len = nlen + moved;
                } else { pm_valid = false; scan_from = r + 1u; }
            } else {                                          // its successor is not looked up in this window: keep it
                if (pm_lane >= 0) E &= ~(1u << pm_lane);
                pm_lane = -1;
                break;
            }
        }
        unsigned first_v = V ? (unsigned)(__ffs(V) - 1) : 32u;
        if (endzone && V) { E = 0u; I = lane_range(0, first_v + 1u); }     // hand over at this iteration boundary (below)
        if (carry_emit) { if (lane == 0) __stcs(tok + wr, carry_tok); wr++; }
        if ((E >> lane) & 1u) {
            uint32_t t;
            if (mlen == 0u) t = v & 0xffu;
            else if (flen >= kWantMin) t = kTokMatch | (flen << 16) | (q - mcand);
            else t = load8(W, q - moved_me);                  // shrunk to one literal at its (moved) start
            __stcs(tok + wr + __popc(E & lt), t);
        }
        wr += __popc(E);
        if (!pre && ((I >> lane) & 1u)) {                    // insert_string_tpl.h:58-75 for every inserted position
            const unsigned prior = peers & I & lt;
            const uint32_t old = prior ? p + (31u - (uint32_t)__clz(prior)) : cand0;
            __stcg(prev + (q & (kWSize - 1u)), (uint16_t)old);
            if ((peers & I & ~lt & ~(1u << lane)) == 0u) __stcg(head + h, (uint16_t)q);
        }
        __syncwarp();
        if (endzone && V) {
            // iteration boundary of the reference: strstart = the (moved) start of lane first_v's match, next_match = that
            // match, table = every position <= p + first_v.  pm, if any, is this very match (not emitted yet).
            MMatch nx;
            const uint32_t fl = __shfl_sync(ZB_FULL, mlen, first_v), fc = __shfl_sync(ZB_FULL, mcand, first_v);
            const uint32_t fm = __shfl_sync(ZB_FULL, moved_me, first_v);
            if (fl) { nx.at = p + first_v - fm; nx.from = fc - fm; nx.len = fl + fm; nx.org = p + first_v + (fm ? 1u : 0u); }
            else { nx.at = nx.org = p + first_v; nx.from = 0; nx.len = 1; }
            wr = serial_medium<LEVEL>(W, n, head, prev, tok, wr, nx.at, nx, true, lane);
            break;
        }
        if (pre && lane == 0) *cv->res_win = (p >> 5) + 1u;   // the searchers may reuse this window's ring slots
        p = next_p; skip = next_skip;
    }
    if (lane == 0) __stcs(tok + wr, kTokEnd);
    return wr;
}

}  // namespace zb
