// deflate_window.cu -- K2w: dictionary-primed chunks (pigz's dependent mode) at levels 2-6.
//
// Reference call sequence per chunk: a FRESH zng_deflateInit2(level, Z_DEFLATED, -15, 8, 0),
// zng_deflateSetDictionary(the 32768 stream bytes in front of the chunk) (deflate.c:456-512), one zng_deflate(flush).  Such a
// chunk has a 96 KiB history: the reference slides its 64 KiB window twice and refills it in between, and at levels 2-6 the
// slides are visible -- prev[] is indexed modulo 32768, slide_hash cuts head / prev entries to 0, block_start goes negative
// (no stored block for a block that began before the slide, deflate_p.h:104-112), deflate_medium drops its next_match at every
// refill, and the refill re-inserts the pending strings (deflate.c:1321-1336).  Instead of re-deriving each of these effects in
// chunk coordinates (as K2's speculative parser does for the single tail slide of an un-primed chunk), this kernel keeps THE
// REFERENCE'S OWN STATE per chain -- a 64 KiB window buffer that is really slid (memcpy of the upper half), head[65536] and
// prev[32768] with relative positions that are really rebased (functable.slide_hash), strstart, lookahead, block_start, insert,
// high_water -- and runs fill_window (deflate.c:1272-1376), deflate_fast (deflate_fast.c:19-104) and deflate_medium
// (deflate_medium.c:22-278) on it, statement by statement, the same restatement as oracle/zo_deflate.c's window engine (which is
// pinned to the unmodified reference on 60 primed cases + live).
//
// B200 mapping: one warp per chunk chain, up to 32 chains per SM.  The parse itself is the serial reference loop, executed
// warp-uniformly (every lane holds the same scalar state, loads are broadcasts, lane 0 stores); the lanes split the bulk steps:
// the window slide and refill copies, slide_hash (8 entries per lane and step, saturating 16-bit subtract), the dictionary's
// 32 766 inserts (32 per step with the nearest-lower-peer rule of lz_ops.cuh:insert_lanes) and the zero fills.  Output: the same
// LZ77 token lists K2's block writer consumes, plus one flag per block: "began before a slide" (buf == NULL).  This is the
// complete-and-exact path for primed chunks, not a fast one: every step waits for its own loads (3-5 GB/s per B200 at level 2,
// against ~1.5 GB/s for the reference's same call sequence on 16 cores); independent chunks keep the speculative K2.
#include "common.cuh"
#include "kernels.h"
#include "lz_ops.cuh"

namespace zb {

namespace {
constexpr uint32_t kWinBuf   = kChunkMax + 512u;      // window + readable slack (the reference's WINDOW_PAD), zero filled
constexpr uint32_t kSlideAtW = kWSize + kMaxDist;     // deflate.c:1285
constexpr uint32_t kSymEndW  = 16383u;                // lit_bufsize - 1 (deflate.c:403)
constexpr int      kWinWarps = 4;

struct WinState {
    uint8_t* win; uint16_t* head; uint16_t* prev;
    uint32_t strstart, lookahead, insert, high_water, match_start;
    int32_t  block_start;
    const uint8_t* next_in; uint32_t avail_in;
};
struct MMatch { uint32_t len, at, from, org; };       // deflate_medium.c:22-27 struct match

// all window / table reads go to L2 (written by other lanes of this warp, ordered by __syncwarp)
__device__ __forceinline__ uint32_t wb(const WinState& s, uint32_t pos) { return (uint32_t)__ldcg(s.win + pos); }
__device__ __forceinline__ uint32_t w32(const WinState& s, uint32_t pos) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s.win) + (pos >> 2);
    return __funnelshift_r(__ldcg(w), __ldcg(w + 1), (pos & 3u) << 3);
}
__device__ __forceinline__ uint64_t w64(const WinState& s, uint32_t pos) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(s.win) + (pos >> 2);
    const uint32_t a = __ldcg(w), b = __ldcg(w + 1), c = __ldcg(w + 2), sh = (pos & 3u) << 3;
    return (uint64_t)__funnelshift_r(a, b, sh) | ((uint64_t)__funnelshift_r(b, c, sh) << 32);
}

// quick_insert_string / one step of insert_string (insert_string_tpl.h:58-75); returns the head it found
__device__ __forceinline__ uint32_t w_insert(WinState& s, uint32_t pos, unsigned lane) {
    const uint32_t h = hash4(w32(s, pos));
    const uint32_t old = (uint32_t)__ldcg(s.head + h);
    if (old != (pos & 0xffffu) && lane == 0) {
        __stcg(s.prev + (pos & (kWSize - 1u)), (uint16_t)old);
        __stcg(s.head + h, (uint16_t)pos);
    }
    __syncwarp();
    return old;
}

// fill_window (deflate.c:1272-1376)
__device__ void w_fill(WinState& s, unsigned lane) {
    do {
        uint32_t more = kChunkMax - s.lookahead - s.strstart;
        if (s.strstart >= kSlideAtW) {
            __syncwarp();
            {   // memcpy(window, window + wsize, wsize)
                const uint4* up = reinterpret_cast<const uint4*>(s.win + kWSize);
                uint4* lo = reinterpret_cast<uint4*>(s.win);
                for (uint32_t i = lane; i < kWSize / 16u; i += 32u) __stcg(lo + i, __ldcg(up + i));
            }
            s.match_start = s.match_start >= kWSize ? s.match_start - kWSize : 0u;
            s.strstart -= kWSize;
            s.block_start -= (int32_t)kWSize;
            if (s.insert > s.strstart) s.insert = s.strstart;
            {   // functable.slide_hash (slide_hash_c.c:15-52)
                const uint32_t w2 = kWSize | (kWSize << 16);
                uint4* h4 = reinterpret_cast<uint4*>(s.head);
                for (uint32_t i = lane; i < 65536u / 8u; i += 32u) {
                    uint4 v = __ldcg(h4 + i);
                    v.x = __vsubus2(v.x, w2); v.y = __vsubus2(v.y, w2); v.z = __vsubus2(v.z, w2); v.w = __vsubus2(v.w, w2);
                    __stcg(h4 + i, v);
                }
                uint4* p4 = reinterpret_cast<uint4*>(s.prev);
                for (uint32_t i = lane; i < kWSize / 8u; i += 32u) {
                    uint4 v = __ldcg(p4 + i);
                    v.x = __vsubus2(v.x, w2); v.y = __vsubus2(v.y, w2); v.z = __vsubus2(v.z, w2); v.w = __vsubus2(v.w, w2);
                    __stcg(p4 + i, v);
                }
            }
            __syncwarp();
            more += kWSize;
        }
        if (s.avail_in == 0u) break;
        const uint32_t n = s.avail_in < more ? s.avail_in : more;
        {   // read_buf: memcpy(window + strstart + lookahead, next_in, n)
            uint8_t* dst = s.win + s.strstart + s.lookahead;
            for (uint32_t i = lane; i < n; i += 32u) __stcg(dst + i, s.next_in[i]);
        }
        s.next_in += n; s.avail_in -= n; s.lookahead += n;
        __syncwarp();
        if (s.lookahead + s.insert >= 3u) {                     // deflate.c:1321-1336
            const uint32_t str = s.strstart - s.insert;
            if (str >= 1u) w_insert(s, str - 1u, lane);
            uint32_t count = s.insert;
            if (s.lookahead == 1u) count--;
            for (uint32_t k = 0; k < count; k++) w_insert(s, str + k, lane);
            s.insert -= count;
        }
    } while (s.lookahead < 262u && s.avail_in != 0u);
    if (s.high_water < kChunkMax) {                             // deflate.c:1345-1372
        const uint32_t curr = s.strstart + s.lookahead;
        uint32_t z0 = 0, zn = 0;
        if (s.high_water < curr) {
            zn = kChunkMax - curr; if (zn > 258u) zn = 258u;
            z0 = curr; s.high_water = curr + zn;
        } else if (s.high_water < curr + 258u) {
            zn = curr + 258u - s.high_water;
            if (zn > kChunkMax - s.high_water) zn = kChunkMax - s.high_water;
            z0 = s.high_water; s.high_water += zn;
        }
        for (uint32_t i = lane; i < zn; i += 32u) __stcg(s.win + z0 + i, (uint8_t)0);
        __syncwarp();
    }
}

// longest_match (match_tpl.h:26-280, non-SLOW instantiation, prev_length 0) with configuration_table[LEVEL] (deflate.c:142-168)
template <int LEVEL>
__device__ uint32_t w_longest_match(WinState& s, uint32_t cand) {
    using P = LmParams<LEVEL>;
    const uint32_t pos = s.strstart;
    uint32_t best = 2, chain = P::kChain;
    const uint32_t limit = pos > kMaxDist ? pos - kMaxDist : 0u;
    for (;;) {
        if (cand >= pos) break;
        const uint32_t w = best < 4u ? 2u : (best < 8u ? 4u : 8u), off = best + 1u - w;
        const uint64_t mask = w == 8u ? ~0ull : ((1ull << (8u * w)) - 1ull);
        if (((w64(s, cand) ^ w64(s, pos)) & mask) == 0ull && ((w64(s, cand + off) ^ w64(s, pos + off)) & mask) == 0ull) {
            uint32_t len = 0;                                   // common prefix, at most 258 (compare256 + 2)
            for (;;) {
                const uint64_t d = w64(s, pos + len) ^ w64(s, cand + len);
                if (d) { len += (uint32_t)(__ffsll((long long)d) - 1) >> 3; break; }
                len += 8u;
                if (len >= kMaxMatch) break;
            }
            if (len > kMaxMatch) len = kMaxMatch;
            if (len > best) {
                s.match_start = cand;
                if (len > s.lookahead) return s.lookahead;
                best = len;
                if (best >= P::kNice) return best;
            } else if (LEVEL < 5) break;                        // early_exit (match_tpl.h:261-266)
        }
        if (--chain == 0u) break;
        cand = (uint32_t)__ldcg(s.prev + (cand & (kWSize - 1u)));
        if (cand <= limit) break;
    }
    return best;
}

// Every step of the serial loops waits for its own loads, and the tables of all resident chains (256 KiB each) are far larger
// than L2, so most of those loads go to DRAM.  This runs ahead of the loop and only warms L2 -- it changes no state and no result:
// once per 32 positions lane l reads head[hash] of position strstart + kAheadBy + l (the value is used one batch later, so nobody
// waits for it) and prefetches the window bytes and the prev[] link of the candidate it found in the previous batch.
constexpr uint32_t kAheadBy = 8;
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
struct Ahead {
    uint32_t at = 0, hv = 0xffffffffu;
    __device__ __forceinline__ void step(const WinState& s, unsigned lane) {
        if (s.strstart < at && at - s.strstart <= 64u) return;    // (a slide moves strstart back by 32768: start over)
        if (hv != 0xffffffffu && hv != 0u) { prefetch_l2(s.win + hv); prefetch_l2(s.prev + (hv & (kWSize - 1u))); }
        const uint32_t q = s.strstart + kAheadBy + lane;
        hv = 0xffffffffu;
        if (q + 4u <= s.strstart + s.lookahead) hv = (uint32_t)__ldcg(s.head + hash4(w32(s, q)));
        at = s.strstart + 32u;
    }
};

struct Sink {                                                   // zng_tr_tally_* (deflate_p.h:61-98) + FLUSH_BLOCK bookkeeping
    uint32_t* tok; uint8_t* flags; uint32_t wr, sym, blk;
    __device__ __forceinline__ void put(uint32_t t, unsigned lane) { if (lane == 0) __stcs(tok + wr, t); wr++; sym++; }
    // FLUSH_BLOCK_ONLY (deflate_p.h:104-112): buf == NULL when the block began before a slide
    __device__ __forceinline__ void flush(WinState& s, unsigned lane) {
        if (lane == 0 && blk < 8u) flags[blk] = s.block_start < 0 ? 1u : 0u;
        blk++; sym = 0;
        s.block_start = (int32_t)s.strstart;
    }
};

// deflate_fast (deflate_fast.c:19-104)
__device__ void w_deflate_fast(WinState& s, Sink& k, int last, unsigned lane) {
    uint32_t match_len = 0;
    Ahead ah;
    for (;;) {
        if (s.lookahead < 262u) { w_fill(s, lane); if (s.lookahead == 0u) break; }
        ah.step(s, lane);
        if (s.lookahead >= kWantMin) {
            const uint32_t hh = w_insert(s, s.strstart, lane);
            const int64_t dist = (int64_t)s.strstart - (int64_t)hh;
            if (dist <= (int64_t)kMaxDist && dist > 0 && hh != 0u) match_len = w_longest_match<2>(s, hh);
        }
        if (match_len >= kWantMin) {
            k.put(kTokMatch | (match_len << 16) | (s.strstart - s.match_start), lane);
            s.lookahead -= match_len;
            if (match_len <= 4u && s.lookahead >= kWantMin) {   // max_insert_length 4 at level 2
                match_len--; s.strstart++;
                for (uint32_t i = 0; i < match_len; i++) w_insert(s, s.strstart + i, lane);
                s.strstart += match_len;
            } else { s.strstart += match_len; w_insert(s, s.strstart - 1u, lane); }
            match_len = 0;
        } else { k.put(wb(s, s.strstart), lane); s.lookahead--; s.strstart++; }
        if (k.sym == kSymEndW) k.flush(s, lane);
    }
    if (last || k.sym) k.flush(s, lane);
}

template <int LEVEL>
__device__ __forceinline__ void w_find(WinState& s, uint32_t cand, MMatch& m) {   // deflate_medium.c:191-215 / :243-262
    const int64_t dist = (int64_t)s.strstart - (int64_t)cand;
    m.at = m.org = s.strstart;
    if (dist <= (int64_t)kMaxDist && dist > 0 && cand != 0u) {
        m.len = w_longest_match<LEVEL>(s, cand); m.from = s.match_start;
        if (m.len < kWantMin || m.from >= m.at) m.len = 1u;
    } else { m.from = 0u; m.len = 1u; }
}
__device__ void w_insert_match(WinState& s, MMatch m, unsigned lane) {             // deflate_medium.c:44-82
    if (s.lookahead <= m.len + kWantMin) return;
    m.at++; m.len--;
    if (m.len < kWantMin - 1u) {
        if (m.len > 0u && m.at >= m.org) {
            const uint32_t cnt = (m.at + m.len - 1u >= m.org) ? m.len : m.org - m.at + 1u;
            for (uint32_t i = 0; i < cnt; i++) w_insert(s, m.at + i, lane);
        }
        return;
    }
    if (m.at >= m.org) {
        const uint32_t cnt = (m.at + m.len - 1u >= m.org) ? m.len : m.org - m.at + 1u;
        for (uint32_t i = 0; i < cnt; i++) w_insert(s, m.at + i, lane);
    } else if (m.org < m.at + m.len) {
        for (uint32_t q = m.org; q < m.at + m.len; q++) w_insert(s, q, lane);
    }
}
__device__ void w_fizzle(const WinState& s, MMatch& cur, MMatch& nxt) {             // deflate_medium.c:84-144
    if (cur.len <= 1u) return;
    if (cur.len > 1u + nxt.from || cur.len > 1u + nxt.at) return;
    if (wb(s, nxt.from + 1u - cur.len) != wb(s, nxt.at + 1u - cur.len)) return;
    MMatch c = cur, n = nxt;
    const uint32_t limit = nxt.at > kMaxDist ? nxt.at - kMaxDist : 0u;
    int moved = 0;
    while (wb(s, n.from - 1u) == wb(s, n.at - 1u)) {
        if (c.len < 1u || n.at <= limit || n.len >= 256u || n.from <= 1u) break;
        n.at--; n.from--; n.len++; c.len--; moved++;
    }
    if (!moved) return;
    if (c.len <= 1u && n.len != 2u) { n.org++; cur = c; nxt = n; }
}
// deflate_medium (deflate_medium.c:146-278); below level 5 without its look-ahead branch (:151,234)
template <int LEVEL>
__device__ void w_deflate_medium(WinState& s, Sink& k, int last, unsigned lane) {
    constexpr bool greedy = LEVEL < 5;
    MMatch cur = {0, 0, 0, 0}, nxt = {0, 0, 0, 0};
    Ahead ah;
    for (;;) {
        if (s.lookahead < 262u) { w_fill(s, lane); if (s.lookahead == 0u) break; nxt.len = 0; }
        ah.step(s, lane);
        if (!greedy && nxt.len > 0u) { cur = nxt; nxt.len = 0; }
        else w_find<LEVEL>(s, s.lookahead >= kWantMin ? w_insert(s, s.strstart, lane) : 0u, cur);
        w_insert_match(s, cur, lane);
        if (!greedy && s.lookahead > 262u && cur.at + cur.len < kChunkMax - 262u) {
            s.strstart = cur.at + cur.len;
            w_find<LEVEL>(s, w_insert(s, s.strstart, lane), nxt);
            if (nxt.len >= kWantMin) w_fizzle(s, cur, nxt);
            s.strstart = cur.at;
        } else nxt.len = 0;
        if (cur.len < kWantMin) { for (uint32_t i = 0; i < cur.len; i++) { k.put(wb(s, cur.at + i), lane); s.lookahead--; } }
        else { k.put(kTokMatch | (cur.len << 16) | (cur.at - cur.from), lane); s.lookahead -= cur.len; }
        s.strstart += cur.len;
        if (k.sym == kSymEndW) k.flush(s, lane);
    }
    if (last || k.sym) k.flush(s, lane);
}

template <int LEVEL>
__global__ void __launch_bounds__(kWinWarps * 32, 12)
window_parse_kernel(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks, uint32_t first, int last,
                    uint32_t* __restrict__ tokens, uint32_t tok_stride, uint32_t* __restrict__ ntok, uint8_t* __restrict__ blkflags,
                    uint32_t* __restrict__ counter, uint16_t* heads, uint16_t* prevs, uint8_t* wins, unsigned long long* sm_slots) {
    const unsigned lane = lane_id();
    const uint32_t sm = smid();
    uint32_t slot = 0;
    if (lane == 0) slot = slot_acquire(sm_slots + sm);
    slot = __shfl_sync(ZB_FULL, slot, 0);
    const size_t slab = (size_t)sm * 64u + slot;
    WinState s;
    s.head = heads + slab * 65536u; s.prev = prevs + slab * kWSize; s.win = wins + slab * kWinBuf;
    for (;;) {
        uint32_t ci = 0;
        if (lane == 0) ci = atomicAdd(counter, 1u);
        ci = __shfl_sync(ZB_FULL, ci, 0);
        if (ci >= nchunks) break;
        {   // a fresh deflate_state: window (zng_zcalloc'd in effect: high_water 0), head and prev all zero (deflate.c:182-184,261)
            uint4* p = reinterpret_cast<uint4*>(s.win);
            for (uint32_t i = lane; i < kWinBuf / 16u; i += 32u) __stcg(p + i, make_uint4(0, 0, 0, 0));
            p = reinterpret_cast<uint4*>(s.head);
            for (uint32_t i = lane; i < 65536u / 8u; i += 32u) __stcg(p + i, make_uint4(0, 0, 0, 0));
            p = reinterpret_cast<uint4*>(s.prev);
            for (uint32_t i = lane; i < kWSize / 8u; i += 32u) __stcg(p + i, make_uint4(0, 0, 0, 0));
        }
        __syncwarp();
        const size_t off = (size_t)ci * chunk;
        const uint32_t len = (uint32_t)min((size_t)chunk, n - off);
        const uint8_t* src = in + off;
        s.strstart = s.lookahead = s.insert = s.high_water = s.match_start = 0u; s.block_start = 0;
        if (first + ci > 0u) {                                  // deflateSetDictionary(src - 32768, 32768) (deflate.c:456-512)
            s.next_in = src - kWSize; s.avail_in = kWSize;
            w_fill(s, lane);
            while (s.lookahead >= 3u) {
                const uint32_t str = s.strstart, cnt = s.lookahead - 2u;
                for (uint32_t p0 = 0; p0 < cnt; p0 += 32u) {    // insert_string(s, str, cnt), 32 positions per step
                    const uint32_t q = str + p0 + lane;
                    const bool on = p0 + lane < cnt;
                    const uint32_t h = on ? hash4(w32(s, q)) : 0u;
                    const unsigned peers = __match_any_sync(ZB_FULL, on ? h : (0x10000u + lane));
                    const unsigned I = __ballot_sync(ZB_FULL, on);
                    insert_lanes(s.head, s.prev, h, q, str + p0, on ? (uint32_t)__ldcg(s.head + h) : 0u, peers, I, lane);
                    __syncwarp();
                }
                s.strstart = str + cnt; s.lookahead = 2u;
                w_fill(s, lane);
            }
            s.strstart += s.lookahead; s.block_start = (int32_t)s.strstart; s.insert = s.lookahead; s.lookahead = 0u;
        }
        s.next_in = src; s.avail_in = len;
        Sink k{tokens + (size_t)ci * tok_stride, blkflags + (size_t)ci * 8u, 0u, 0u, 0u};
        if (lane < 8u) blkflags[(size_t)ci * 8u + lane] = 0u;
        __syncwarp();
        if constexpr (LEVEL == 2) w_deflate_fast(s, k, last, lane); else w_deflate_medium<LEVEL>(s, k, last, lane);
        if (lane == 0) { __stcs(k.tok + k.wr, kTokEnd); ntok[ci] = k.wr; }
    }
    __syncwarp();
    if (lane == 0) atomicAnd(sm_slots + sm, ~(1ull << slot));
}
}  // namespace

size_t deflate_window_win_bytes(uint32_t nsmid) { return (size_t)nsmid * 64u * kWinBuf; }

// `in` points at chunk `first` of the stream (first > 0: the 32768 bytes in front of it are readable: its dictionary).
// heads / prevs: the K2 slab pools (u16 entries); wins: deflate_window_win_bytes(); blkflags: 8 bytes per chunk.
cudaError_t launch_window_parse(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks, uint32_t first, int last, uint32_t* tokens,
                                uint32_t tok_stride, uint32_t* ntok, uint8_t* blkflags, uint32_t* counter, uint16_t* heads, uint16_t* prevs,
                                uint8_t* wins, unsigned long long* sm_slots, int num_sms, int chains_per_sm, int level, cudaStream_t stream) {
    if (nchunks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    uint32_t ctas_per_sm = ((uint32_t)chains_per_sm + kWinWarps - 1u) / kWinWarps;
    uint32_t grid = (uint32_t)num_sms * ctas_per_sm;
    const uint64_t need = ((uint64_t)nchunks + kWinWarps - 1u) / kWinWarps;
    if (need < grid) grid = (uint32_t)need;
#define ZB_LAUNCH_WIN(L) window_parse_kernel<L><<<grid, kWinWarps * 32, 0, stream>>>(in, n, chunk, nchunks, first, last, tokens, tok_stride, ntok, \
                                                                                  blkflags, counter, heads, prevs, wins, sm_slots)
    switch (level) {
        case 2: ZB_LAUNCH_WIN(2); break;
        case 3: ZB_LAUNCH_WIN(3); break;
        case 4: ZB_LAUNCH_WIN(4); break;
        case 5: ZB_LAUNCH_WIN(5); break;
        case 6: ZB_LAUNCH_WIN(6); break;
        default: return cudaErrorInvalidValue;
    }
#undef ZB_LAUNCH_WIN
    return cudaGetLastError();
}

}  // namespace zb
