// deflate_fast.cu -- K2: levels 2 (deflate_fast) and 3-6 (deflate_medium; 5-6 in lazy_parse.cuh) of independent <= 64 KiB chunks.
//
// Reference semantics reproduced bit-exactly (files under /root/reference):
//   deflate_fast.c:19-104         greedy parse: hash_head != 0, longest_match, insert_string for short matches
//   deflate_medium.c:22-278       levels 3-6: insert_match, (5-6) the look-ahead-one branch and fizzle_matches
//   match_tpl.h:26-280            longest_match, non-SLOW, {nice, chain} of configuration_table (deflate.c:142-168) per level
//   insert_string_tpl.h:48-104    quick_insert_string / insert_string: head[] + prev[] hash chains
//   deflate_p.h:61-112            zng_tr_tally_lit/_dist, FLUSH_BLOCK (block full at sym_next == 16383)
//   trees.c:106-120,151-173,185-270,280-312,322-405   init_block, pqdownheap, gen_bitlen, gen_codes, build_tree
//   trees.c:411-587,592-609,625-741                   scan_tree/send_tree, build_bl_tree, send_all_trees,
//                                                     zng_tr_stored_block, zng_tr_flush_block, compress_block
//   deflate.c:1285-1299           the tail slide: what over-reads past the data see and
//                                 deflate_p.h:104-112 buf == NULL for a block that began below 32768
//   deflate.c:1061-1083           empty stored block appended for Z_SYNC_FLUSH / Z_FULL_FLUSH
//
// B200 mapping: same split as K1 (deflate_quick.cu).
//   K2a fast_parse_kernel   one warp per chain.  head[] (128 KiB) and prev[] (64 KiB) of the chain live in
//     per-warp slabs of global memory (L2 resident).  The warp speculates on 32 consecutive positions: every
//     lane looks up its hash head against the table state at the window start and walks its own hash chain
//     (<= 4 candidates, 12-byte prefix compare per candidate; longer matches are measured by the whole warp);
//     a ballot/ffs walk replays the reference's decisions, a lane whose hash equals that of an earlier
//     INSERTED lane of the window saw a stale head and cuts the window there.  Inserted-but-covered
//     positions (the inside of a length-4 match, the last byte of a longer one) take their prev[] link from
//     the nearest earlier inserted lane with the same hash, which is what the serial insert order produces.
//     Bytes past the end of the data come from a small per-chain image of the reference's stale window.
//   K2b block_emit_kernel   one warp per chunk: tokens -> deflate blocks of <= 16383 symbols.  Frequencies
//     with shared-memory atomics, the three Huffman trees serially on lane 0 exactly as trees.c orders its
//     heap, stored / static / dynamic choice by the reference's byte counts, then warp-parallel bit packing
//     (code lookup, prefix scan of lengths, OR into a staging ring, 128-bit stores).
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "lz_ops.cuh"
#include "lazy_parse.cuh"

namespace zb {

constexpr int      kFastWarps  = 4;
constexpr uint32_t kTailWords  = 96;              // image of [len, len + ~380): what over-reads past the data see
constexpr uint32_t kSymEnd     = 16383;           // deflate.c:400-403 sym_end with memLevel 8 (LIT_MEM)
constexpr uint32_t kSlideAt    = kWSize + kMaxDist;   // deflate.c:1285

// ---------------------------------------------------------------- K2a parser (operators: lz_ops.cuh)
// Parse one chunk; tokens to tok[0..count) + end marker.  Returns the token count.
// LEVEL 2 = deflate_fast (deflate_fast.c:19-104); LEVEL 3 = deflate_medium below level 5 (deflate_medium.c:146-278 with
// early_exit: no look-ahead-one branch, so a greedy parse; a match shorter than 4 becomes one literal; insert_match
// (:44-82) inserts EVERY position inside the match unless lookahead <= match_length + WANT_MIN_MATCH).
template <int LEVEL>
__device__ uint32_t fast_parse_warp(const VWindow W, uint32_t n, uint16_t* head, uint16_t* prev, uint32_t* __restrict__ tok) {
    using P = LmParams<LEVEL>;
    constexpr uint32_t kCmp = P::kCmp;
    enum : unsigned { kNone = 0, kAll = 1, kLast = 2 };       // what is still to be inserted behind a match that leaves the window
    const unsigned lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    uint32_t wr = 0, p = 0, skip = 0;                         // lanes below `skip` are insert-only (already covered by a match)
    while (p < n) {
        const uint32_t q = p + lane;
        const bool inb = q < n;
        const bool act = q + kWantMin <= n;                  // deflate_fast.c:43,72 / deflate_medium.c:181 lookahead >= WANT_MIN_MATCH
        uint32_t v, z = 0; uint64_t x;
        {
            const uint32_t qb = q + W.skew, i = qb >> 2, sh = (qb & 3u) << 3;
            const uint32_t a0 = W.word(i), a1 = W.word(i + 1), a2 = W.word(i + 2), a3 = W.word(i + 3);
            v = __funnelshift_r(a0, a1, sh);
            x = (uint64_t)__funnelshift_r(a1, a2, sh) | ((uint64_t)__funnelshift_r(a2, a3, sh) << 32);
            if (kCmp > 12u) z = __funnelshift_r(a3, W.word(i + 4), sh);
        }
        const uint32_t h = hash4(v);
        const uint32_t cand0 = act ? (uint32_t)__ldcg(head + h) : 0u;
        uint32_t mlen = 0, mcand = 0;                        // mlen: 0 none, 4..kCmp-1 exact (clipped), kCmp = "kCmp or more"
        if (act && lane >= skip && cand0 != 0u && (q - cand0 - 1u) < kMaxDist)      // deflate_fast.c:48-50, deflate_medium.c:194
            mlen = longest_match_lane<LEVEL>(W, q, v, x, z, cand0, n - q, prev, mcand);
        const unsigned actm = __ballot_sync(ZB_FULL, act);
        const unsigned peers = __match_any_sync(ZB_FULL, act ? h : (0x10000u + lane));
        const unsigned M = __ballot_sync(ZB_FULL, mlen != 0u);
        const unsigned nl = min(32u, n - p);
        // ---- walk 1: replay the greedy decisions
        unsigned cur = skip, covered = 0, icov = 0;
        unsigned pending = skip >= 32u ? kAll : kNone;       // a window of insert-only lanes continues an insert-all match
        while (cur < nl) {
            const unsigned rest = M & ~lane_range(0, cur);
            if (rest == 0u) { cur = nl; break; }
            const unsigned k = (unsigned)(__ffs(rest) - 1);
            uint32_t len = __shfl_sync(ZB_FULL, mlen, k);
            const uint32_t qk = p + k;
            if (len >= kCmp) {
                const uint32_t ck = __shfl_sync(ZB_FULL, mcand, k);
                len = kCmp + vwarp_compare256(W, qk + kCmp + W.skew, ck + kCmp + W.skew, lane);
                len = min(len, kMaxMatch);                   // compare256 + 2 never exceeds 258
                len = min(len, n - qk);                      // len > lookahead: return lookahead
                if (lane == k) mlen = len;
            }
            covered |= lane_range(k + 1u, k + len);
            cur = k + len;
            if (LEVEL == 2) {
                // deflate_fast.c:72-85: short matches insert every covered position, longer ones only the last
                if (len <= 4u && n - (qk + len) >= kWantMin) { icov |= lane_range(k + 1u, cur); pending = kAll; }
                else { if (cur - 1u < 32u) icov |= 1u << (cur - 1u); pending = kLast; }
            } else {
                // deflate_medium.c:44-82 insert_match: all of strstart+1 .. strstart+len-1, or nothing near the end
                if (n - qk > len + kWantMin) { icov |= lane_range(k + 1u, cur); pending = kAll; }
                else pending = kNone;
            }
        }
        unsigned V = lane_range(skip, min(cur, nl)) & ~covered;
        unsigned I = (V | icov | lane_range(0, skip)) & actm;
        // ---- walk 2: a visited lane whose hash was inserted earlier in this window looked up a stale head
        const unsigned S = __ballot_sync(ZB_FULL, ((V >> lane) & 1u) && (peers & I & lt) != 0u);
        uint32_t next_p, next_skip;
        if (S) {
            const unsigned j = __ffs(S) - 1u;
            V &= lane_range(0, j); I &= lane_range(0, j);
            next_p = p + j; next_skip = 0;
        } else if (cur > 32u) {                              // the last match runs past the window
            if (pending == kAll) { next_p = p + 32u; next_skip = cur - 32u; }          // its covered positions: insert-only lanes
            else if (pending == kLast && cur - 1u >= 32u) { next_p = p + cur - 1u; next_skip = 1u; }
            else { next_p = p + cur; next_skip = 0; }
        } else { next_p = p + cur; next_skip = 0; }
        const bool vis = (V >> lane) & 1u;
        if (vis && inb) {
            const uint32_t t = mlen ? (kTokMatch | (mlen << 16) | (q - mcand)) : (v & 0xffu);
            __stcs(tok + wr + __popc(V & lt), t);
        }
        wr += __popc(V);
        insert_lanes(head, prev, h, q, p, cand0, peers, I, lane);    // insert_string_tpl.h:58-75 for every inserted position
        p = next_p; skip = next_skip;
        __syncwarp();
    }
    if (lane == 0) __stcs(tok + wr, kTokEnd);
    return wr;
}

// Levels 4-6 walk hash chains of up to 24 / 32 / 128 dependent prev[] links per position and are bound by that latency.  Trading
// registers for occupancy (40 registers, 48 chains per SM) was measured and gains 0-5 % (profiles/r2_sweep_k2_levels.txt): not kept.
template <int LEVEL>
__global__ void __launch_bounds__(kFastWarps * 32, 8)
fast_parse_kernel(const uint8_t* __restrict__ in, size_t n, uint32_t chunk, uint32_t nchunks,
                  uint32_t* __restrict__ tokens, uint32_t tok_stride, uint32_t* __restrict__ ntok,
                  uint32_t* __restrict__ counter, uint16_t* __restrict__ heads, uint16_t* __restrict__ prevs,
                  uint32_t* __restrict__ tails, unsigned long long* __restrict__ sm_slots, int have_prev, StreamSync sy) {
    extern __shared__ uint32_t carve_out_only[];              // never touched: see launch_fast_parse (streamed launches)
    const unsigned lane = lane_id();
    const uint32_t sm = smid();
    uint32_t slot = 0;
    if (lane == 0) slot = slot_acquire(sm_slots + sm);
    slot = __shfl_sync(ZB_FULL, slot, 0);
    const size_t slab = (size_t)sm * 64u + slot;
    uint16_t* head = heads + slab * 65536u;
    uint16_t* prev = prevs + slab * kWSize;
    uint32_t* tail = tails + slab * kTailWords;
    for (;;) {
        uint32_t ci = 0;
        if (lane == 0) ci = atomicAdd(counter, 1u);
        ci = __shfl_sync(ZB_FULL, ci, 0);
        if (ci >= nchunks) break;
        if (sy.ready) {                                        // streamed input (see quick_parse_kernel)
            if (lane == 0) {
                const long long t0 = clock64();
                while (*(volatile const uint32_t*)sy.ready <= ci) {
                    __nanosleep(1000);
                    if (clock64() - t0 > sy.patience) { atomicExch(sy.failed, 1u); break; }
                }
            }
            __syncwarp();
        }
        {   // CLEAR_HASH (deflate.c:182-184); prev[] needs no clearing: only links of inserted positions are ever followed
            uint4* h4 = reinterpret_cast<uint4*>(head);
#pragma unroll 8
            for (uint32_t i = lane; i < 65536u * 2u / 16u; i += 32u) h4[i] = make_uint4(0, 0, 0, 0);
        }
        const size_t off = (size_t)ci * chunk;
        const uint32_t len = (uint32_t)min((size_t)chunk, n - off);
        const uint8_t* src = in + off;
        VWindow W;
        W.skew = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
        W.w = reinterpret_cast<const uint32_t*>(src - W.skew);
        W.tail = tail;
        W.tw0 = (W.skew + len) >> 2;
        // Window image past the data: a full chunk was slid, so offset 65536+k reads chunk[32768+k];
        // a short chunk on a reused stream sees the previous chunk's bytes (window position p holds stream byte
        // off - 32768 + (p & 32767)), on a fresh stream zeros (deflate.c:1348-1372).
        const bool stale = chunk == kChunkMax && (ci > 0u || have_prev) && len < kChunkMax;
        for (uint32_t t = lane; t < kTailWords; t += 32u) {
            uint32_t wv = 0;
#pragma unroll
            for (uint32_t b = 0; b < 4u; b++) {
                const int pos = (int)(4u * (W.tw0 + t) + b) - (int)W.skew;
                uint32_t by = 0;
                if (pos >= 0) {
                    uint32_t pp = (uint32_t)pos;
                    if (pp >= kChunkMax) pp -= kWSize;
                    if (pp < len) by = src[pp];
                    else if (stale) by = *(src - (size_t)kWSize + (pp & (kWSize - 1u)));
                }
                wv |= by << (8u * b);
            }
            __stcg(tail + t, wv);
        }
        __syncwarp();
        uint32_t cnt;
        if constexpr (LEVEL >= 5) cnt = lazy_parse_warp<LEVEL>(W, len, head, prev, tokens + (size_t)ci * tok_stride);
        else cnt = fast_parse_warp<LEVEL>(W, len, head, prev, tokens + (size_t)ci * tok_stride);
        if (lane == 0) ntok[ci] = cnt;
        if (sy.done) {
            __syncwarp();
            __threadfence();
            if (lane == 0) {
                const uint32_t k = atomicAdd(sy.done + (ci >> sy.done_shift), 1u) + 1u;
                if (sy.host_done) {
                    const uint32_t slab = ci >> sy.done_shift, first = slab << sy.done_shift;
                    const uint32_t want = min(nchunks - first, 1u << sy.done_shift);
                    if (k == want) { __threadfence_system(); *(volatile uint32_t*)(sy.host_done + slab) = want; }
                }
            }
        }
    }
    __syncwarp();
    if (lane == 0) atomicAnd(sm_slots + sm, ~(1ull << slot));
}

// ---------------------------------------------------------------- K2c: one CTA per chunk for the deep-chain levels (5, 6)
// The warp-per-chain parser above runs at the random-access rate of L2 + DRAM: 256 KiB of tables per chain, thousands of chains, and
// at level 6 up to 128 dependent prev[] links per position.  Here three chunks share an SM and the hot part of their state is on
// chip: each CTA keeps its chunk's prev[] ring in shared memory (64 KiB), head[] and the chunks themselves stay L2 / L1 resident
// (444 chunks in flight instead of 4 736).  While the look-ahead regime lasts every position is inserted in order and the match at a
// position does not depend on the parse (lazy_parse.cuh), so the work splits three ways:
//   warp 1       inserter: insert_string for ALL positions below F, 32 per step, in order (also into the global prev[] ring, which
//                the end zone's serial code continues from); at most kCoopAhead windows ahead of the resolver, because entering
//                position t overwrites the ring slot of t - 32768, which a searcher at s may still follow while t > s + 262
//   warps 2..    searchers: claim 32-position windows, wait for their links, walk the chains (longest_match_lane on shared memory),
//                leave (length, start) per position in a small ring
//   warp 0       resolver: lazy_parse_warp's walks on the ring's results; from F on (the last ~550 bytes) the unchanged serial path
constexpr int kCtaWarps = 10;
struct alignas(16) CtaSmem {
    uint16_t prevf[kWSize];
    uint32_t ring[kCoopRing];
    uint32_t ready[kCoopReady];
    uint32_t ins_upto, next_win, res_win, ci, slot;
};

// the image of the window past the data (see fast_parse_kernel), by `nthreads` threads
__device__ __forceinline__ void build_tail_image(const VWindow& W, const uint8_t* src, uint32_t len, bool stale, uint32_t* tail, unsigned t0, unsigned nthreads) {
    for (uint32_t t = t0; t < kTailWords; t += nthreads) {
        uint32_t wv = 0;
#pragma unroll
        for (uint32_t b = 0; b < 4u; b++) {
            const int pos = (int)(4u * (W.tw0 + t) + b) - (int)W.skew;
            uint32_t by = 0;
            if (pos >= 0) {
                uint32_t pp = (uint32_t)pos;
                if (pp >= kChunkMax) pp -= kWSize;
                if (pp < len) by = src[pp];
                else if (stale) by = *(src - (size_t)kWSize + (pp & (kWSize - 1u)));
            }
            wv |= by << (8u * b);
        }
        __stcg(tail + t, wv);
    }
}

template <int LEVEL>
__global__ void __launch_bounds__(kCtaWarps * 32, 3)
medium_cta_kernel(const uint8_t* __restrict__ in, size_t n, uint32_t chunk, uint32_t nchunks,
                  uint32_t* __restrict__ tokens, uint32_t tok_stride, uint32_t* __restrict__ ntok,
                  uint32_t* __restrict__ counter, uint16_t* __restrict__ heads, uint16_t* __restrict__ prevs,
                  uint32_t* __restrict__ tails, unsigned long long* __restrict__ sm_slots, int have_prev) {
    extern __shared__ __align__(16) unsigned char cta_smem[];
    CtaSmem& S = *reinterpret_cast<CtaSmem*>(cta_smem);
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5, lt = (1u << lane) - 1u;
    const uint32_t sm = smid();
    if (threadIdx.x == 0) S.slot = slot_acquire(sm_slots + sm);
    __syncthreads();
    const size_t slab = (size_t)sm * 64u + S.slot;
    uint16_t* head = heads + slab * 65536u;
    uint16_t* prev = prevs + slab * kWSize;
    uint32_t* tail = tails + slab * kTailWords;
    for (;;) {
        if (threadIdx.x == 0) S.ci = atomicAdd(counter, 1u);
        __syncthreads();
        const uint32_t ci = S.ci;
        if (ci >= nchunks) break;
        const size_t off = (size_t)ci * chunk;
        const uint32_t len = (uint32_t)min((size_t)chunk, n - off);
        const uint8_t* src = in + off;
        VWindow W;
        W.skew = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
        W.w = reinterpret_cast<const uint32_t*>(src - W.skew);
        W.tail = tail;
        W.tw0 = (W.skew + len) >> 2;
        {   // CLEAR_HASH (deflate.c:182-184) by the whole CTA; prev[] needs no clearing
            uint4* h4 = reinterpret_cast<uint4*>(head);
            for (uint32_t i = threadIdx.x; i < 65536u * 2u / 16u; i += blockDim.x) __stcg(h4 + i, make_uint4(0, 0, 0, 0));
        }
        const bool stale = chunk == kChunkMax && (ci > 0u || have_prev) && len < kChunkMax;
        build_tail_image(W, src, len, stale, tail, threadIdx.x, blockDim.x);
        // F: the first window start (a multiple of 32) that lies in the end zone of lazy_parse_warp; everything below it is the CTA's
        const uint32_t F = len < kLazyGuard ? 0u : ((len - kLazyGuard) / 32u + 1u) * 32u;
        if (threadIdx.x < kCoopReady) S.ready[threadIdx.x] = 0u;
        if (threadIdx.x == 0) { S.ins_upto = 0u; S.next_win = 0u; S.res_win = 0u; }
        __syncthreads();
        volatile uint32_t* ins_upto = &S.ins_upto; volatile uint32_t* res_win = &S.res_win; volatile uint32_t* ready = S.ready;
        if (warp == 0) {
            CoopView cv{S.ring, ready, res_win, ins_upto, F};
            const uint32_t cnt = lazy_parse_warp<LEVEL, true>(W, len, head, prev, tokens + (size_t)ci * tok_stride, &cv);
            if (lane == 0) ntok[ci] = cnt;
        } else if (warp == 1) {
            uint32_t vn = F ? load32(W, lane) : 0u;
            for (uint32_t base = 0; base < F; base += 32u) {
                const uint32_t q = base + lane, v = vn;
                if (lane == 0) for (uint32_t tries = 0; (base >> 5) > *res_win + kCoopAhead; tries++) { __nanosleep(64); if (tries > (1u << 26)) __trap(); }
                __syncwarp();
                if (base + 32u < F) vn = load32(W, q + 32u);             // the next window's bytes are on their way while this one waits for head[]
                const uint32_t h = hash4(v);
                const uint32_t table_head = (uint32_t)__ldcg(head + h);
                const unsigned peers = __match_any_sync(ZB_FULL, h);
                const unsigned prior = peers & lt;
                const uint32_t old = prior ? base + (31u - (uint32_t)__clz(prior)) : table_head;   // insert_string_tpl.h:58-75 in serial order
                S.prevf[q & (kWSize - 1u)] = (uint16_t)old;
                __stcg(prev + (q & (kWSize - 1u)), (uint16_t)old);
                if ((peers & ~lt & ~(1u << lane)) == 0u) __stcg(head + h, (uint16_t)q);
                __syncwarp();
                __threadfence_block();
                if (lane == 0) *ins_upto = base + 32u;
            }
        } else {
            for (;;) {
                uint32_t w = 0;
                if (lane == 0) w = atomicAdd(&S.next_win, 1u);
                w = __shfl_sync(ZB_FULL, w, 0);
                if (32u * w >= F) break;
                if (lane == 0) {
                    coop_wait(ins_upto, 32u * (w + 1u));
                    for (uint32_t tries = 0; w >= *res_win + kCoopReady; tries++) { __nanosleep(400); if (tries > (1u << 24)) __trap(); }
                }
                __syncwarp();
                __threadfence_block();
                const uint32_t q = 32u * w + lane;
                uint32_t v, z; uint64_t x;
                {
                    const uint32_t qb = q + W.skew, i = qb >> 2, sh = (qb & 3u) << 3;
                    const uint32_t a0 = W.word(i), a1 = W.word(i + 1), a2 = W.word(i + 2), a3 = W.word(i + 3);
                    v = __funnelshift_r(a0, a1, sh);
                    x = (uint64_t)__funnelshift_r(a1, a2, sh) | ((uint64_t)__funnelshift_r(a2, a3, sh) << 32);
                    z = __funnelshift_r(a3, W.word(i + 4), sh);
                }
                const uint32_t cand0 = lds_u16(S.prevf, q & (kWSize - 1u));
                uint32_t mlen = 0, mcand = 0;
                if (cand0 != 0u && (q - cand0 - 1u) < kMaxDist)
                    mlen = longest_match_lane<LEVEL, true>(W, q, v, x, z, cand0, 0x7fffffffu, S.prevf, mcand);
                S.ring[q & (kCoopRing - 1u)] = mlen | (mcand << 16);
                __syncwarp();
                __threadfence_block();
                if (lane == 0) ready[w & (kCoopReady - 1u)] = w + 1u;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAnd(sm_slots + sm, ~(1ull << S.slot));
}

// ---------------------------------------------------------------- K2b block writer
constexpr int      kBlkWarps   = 4;
constexpr uint32_t kBStageWords = 512;             // 2 KiB staging ring per warp (a block header is <= 141 words)
constexpr uint32_t kBStageSeg   = 128;
constexpr int L_CODES = 286, D_CODES = 30, BL_CODES = 19, HEAP_SZ = 2 * L_CODES + 1, MAX_BITS = 15;

struct alignas(16) BlockWs {
    uint32_t stage[kBStageWords];
    uint32_t lfreq[L_CODES + 2], dfreq[D_CODES + 2], bfreq[BL_CODES + 1];
    uint16_t wfreq[L_CODES + 2], wdad[HEAP_SZ], wlen[HEAP_SZ];   // the tree under construction (trees.c ct_data); freq: leaves only
    // the priority queue of build_tree, 1-based, never more than L_CODES entries (one per leaf): sort keys and node ids in
    // separate arrays, so that the two children 2k, 2k+1 of a node come with ONE 8-byte load and compare as plain integers
    alignas(8) uint32_t hkey[L_CODES + 4];
    uint16_t hid[L_CODES + 4];
    int16_t  heap[HEAP_SZ];                                   // nodes in the order they left the queue
    uint16_t bl_count[MAX_BITS + 1];
    uint16_t lcode[L_CODES + 2], dcode[D_CODES + 2], bcode[BL_CODES + 1];
    uint16_t llen[L_CODES + 2], dlen[D_CODES + 2], blen[BL_CODES + 1];     // +1: the scan_tree guard entry
};

__device__ __forceinline__ uint32_t len_xbits(uint32_t c) { return (c < 8u || c == 28u) ? 0u : (c - 4u) >> 2; }
__device__ __forceinline__ uint32_t dst_xbits(uint32_t c) { return c < 4u ? 0u : (c - 2u) >> 1; }
__device__ __forceinline__ uint32_t fx_llen(uint32_t n) { return n < 144u ? 8u : (n < 256u ? 9u : (n < 280u ? 7u : 8u)); }
__device__ __forceinline__ uint32_t bl_xbits(uint32_t c) { return c == 16u ? 2u : (c == 17u ? 3u : (c == 18u ? 7u : 0u)); }
__device__ __forceinline__ uint32_t bl_order(uint32_t i) { return i < 3u ? 16u + i : (i == 3u ? 0u : ((i & 1u) ? 7u - ((i - 5u) >> 1) : 8u + ((i - 4u) >> 1))); }

// zng_length_code[len-3] / d_code(dist-1) with their extra bits (trees_tbl.h, deflate.h:436), computed
__device__ __forceinline__ void len_symbol(uint32_t lc, uint32_t& ls, uint32_t& lx, uint32_t& lxb) {
    if (lc < 8u) { ls = lc; lx = 0; lxb = 0; }
    else if (lc == 255u) { ls = 28; lx = 0; lxb = 0; }
    else { const uint32_t nb = 31u - __clz(lc); lxb = nb - 2u; ls = 4u * (nb - 1u) + ((lc >> lxb) & 3u); lx = lc & ((1u << lxb) - 1u); }
}
__device__ __forceinline__ void dist_symbol(uint32_t d, uint32_t& ds, uint32_t& dx, uint32_t& dxb) {
    if (d < 4u) { ds = d; dx = 0; dxb = 0; }
    else { const uint32_t nb = 31u - __clz(d); dxb = nb - 1u; ds = 2u * nb + ((d >> dxb) & 1u); dx = d & ((1u << dxb) - 1u); }
}

struct Emitter {
    uint32_t* stage; uint8_t* out; uint32_t bitpos, flushed;
    __device__ __forceinline__ void or_bits(uint32_t at, uint64_t bits, uint32_t nbits) {
        if (nbits == 0u) return;
        const uint32_t wd = at >> 5, sh = at & 31u;
        const uint32_t lo = (uint32_t)bits << sh;
        const uint64_t rest = sh ? (bits >> (32u - sh)) : (bits >> 32);      // what does not fit the first word
        if (lo) atomicOr(&stage[wd & (kBStageWords - 1u)], lo);
        const uint32_t mid = (uint32_t)rest, hi = (uint32_t)(rest >> 32);
        if (mid) atomicOr(&stage[(wd + 1u) & (kBStageWords - 1u)], mid);
        if (hi) atomicOr(&stage[(wd + 2u) & (kBStageWords - 1u)], hi);
    }
    // store [flushed, upto_words) of the ring to global memory and re-zero it (warp-collective)
    __device__ __forceinline__ void flush(uint32_t upto_words, unsigned lane) {
        __syncwarp();
        while (flushed < upto_words) {
            const uint32_t cnt = min(upto_words - flushed, kBStageSeg);
            const uint32_t base = flushed & (kBStageWords - 1u);
            if ((cnt & 3u) == 0u && (flushed & 3u) == 0u) {
                uint4* g = reinterpret_cast<uint4*>(out + (size_t)flushed * 4u);
                uint4* sm = reinterpret_cast<uint4*>(&stage[base]);
                for (uint32_t i = lane; i < cnt / 4u; i += 32u) { __stcs(g + i, sm[i]); sm[i] = make_uint4(0, 0, 0, 0); }
            } else {
                uint32_t* g = reinterpret_cast<uint32_t*>(out + (size_t)flushed * 4u);
                for (uint32_t i = lane; i < cnt; i += 32u) { g[i] = stage[base + i]; stage[base + i] = 0u; }
            }
            flushed += cnt;
        }
        __syncwarp();
    }
    // every lane contributes (bits, nbits <= 48) in lane order (warp-collective)
    __device__ __forceinline__ void put_warp(uint64_t bits, uint32_t nbits, unsigned lane) {
        uint32_t incl = nbits;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(ZB_FULL, incl, d); if ((int)lane >= d) incl += u; }
        or_bits(bitpos + incl - nbits, bits, nbits);
        bitpos += __shfl_sync(ZB_FULL, incl, 31);
        if ((bitpos >> 5) - flushed >= kBStageSeg) flush(((bitpos >> 5) / kBStageSeg) * kBStageSeg, lane);
    }
    // lane 0 only, no flush (callers bound the amount: block headers, tree descriptions)
    __device__ __forceinline__ void put_serial(uint32_t bits, uint32_t nbits) { or_bits(bitpos, bits, nbits); bitpos += nbits; }
};

// ---- trees.c restated for one lane --------------------------------------------------------------------
struct TreeKind { int elems, max_len, kind; };          // kind 0 literal/length, 1 distance, 2 bit-length

// Heap entries carry their own sort key so that a comparison needs no second lookup: key = frequency << 8 | depth.
// smaller(n, m) of trees.c:141-143 -- freq[n] < freq[m] || (freq equal && depth[n] <= depth[m]) -- is then key_n <= key_m.
__device__ __forceinline__ uint32_t heap_key(uint32_t freq, uint32_t depth) { return (freq << 8) | depth; }
__device__ __forceinline__ void sift_down(uint32_t* key, uint16_t* id, int k, int heap_len) {   // trees.c:151-173 pqdownheap
    const uint32_t vk = key[k];
    const uint16_t vi = id[k];
    int j = k << 1;
    while (j <= heap_len) {
        const uint2 kk = *reinterpret_cast<const uint2*>(key + j);          // key[j], key[j+1]: j is even
        uint32_t kj = kk.x;
        if (j < heap_len && kk.y <= kj) { kj = kk.y; j++; }
        if (vk <= kj) break;
        key[k] = kj; id[k] = id[j]; k = j; j <<= 1;
    }
    key[k] = vk; id[k] = vi;
}

// build_tree + gen_bitlen + gen_codes (trees.c:185-405).  freq_in -> code_out/len_out; returns max_code.
// One copy of this code (not one per tree and call site): the kernel was 15 000 instructions = 240 KiB, and a fifth of the warps' stall
// samples were instruction fetch (no_instruction 2.25 per issue slot in profiles/r2_block_emit_ncu_summary.txt).
__device__ __noinline__ int build_huffman(BlockWs& T, const uint32_t* freq_in, TreeKind tk, uint16_t* code_out, uint16_t* len_out,
                             uint32_t& opt_len, uint32_t& static_len) {
    uint32_t* key = T.hkey; uint16_t* id = T.hid;
    int heap_len = 0, heap_max = HEAP_SZ, max_code = -1, node;
    for (int n = 0; n < tk.elems; n++) {
        T.wfreq[n] = (uint16_t)freq_in[n];
        if (freq_in[n]) { ++heap_len; key[heap_len] = heap_key(freq_in[n], 0); id[heap_len] = (uint16_t)(max_code = n); }
        else T.wlen[n] = 0;
    }
    while (heap_len < 2) {                                  // trees.c:352-360: force at least two codes
        node = (max_code < 2 ? ++max_code : 0);
        ++heap_len; key[heap_len] = heap_key(1, 0); id[heap_len] = (uint16_t)node;
        T.wfreq[node] = 1;
        opt_len--;
        if (tk.kind == 0) static_len -= fx_llen((uint32_t)node); else if (tk.kind == 1) static_len -= 5u;
    }
    for (int n = heap_len / 2; n >= 1; n--) sift_down(key, id, n, heap_len);
    node = tk.elems;
    do {
        const uint32_t kn = key[1]; const int n = id[1];
        key[1] = key[heap_len]; id[1] = id[heap_len]; heap_len--;
        sift_down(key, id, 1, heap_len);
        const uint32_t km = key[1]; const int m = id[1];
        T.heap[--heap_max] = (int16_t)n;                    // the sorted node list gen_bitlen walks
        T.heap[--heap_max] = (int16_t)m;
        const uint32_t dn = kn & 0xffu, dm = km & 0xffu;
        T.wdad[n] = T.wdad[m] = (uint16_t)node;
        key[1] = heap_key((kn >> 8) + (km >> 8), ((dn >= dm ? dn : dm) + 1u) & 0xffu); id[1] = (uint16_t)node;
        node++;
        sift_down(key, id, 1, heap_len);
    } while (heap_len >= 2);
    T.heap[--heap_max] = (int16_t)id[1];
    // gen_bitlen (trees.c:185-270)
    int overflow = 0, h;
    for (int i = 0; i <= MAX_BITS; i++) T.bl_count[i] = 0;
    T.wlen[T.heap[heap_max]] = 0;
    for (h = heap_max + 1; h < HEAP_SZ; h++) {
        const int n = T.heap[h];
        uint32_t bits = T.wlen[T.wdad[n]] + 1u;
        if (bits > (uint32_t)tk.max_len) { bits = (uint32_t)tk.max_len; overflow++; }
        T.wlen[n] = (uint16_t)bits;
        if (n > max_code) continue;
        T.bl_count[bits]++;
        uint32_t xb = 0, sl = 0;
        if (tk.kind == 0) { xb = n >= 257 ? len_xbits((uint32_t)n - 257u) : 0u; sl = fx_llen((uint32_t)n); }
        else if (tk.kind == 1) { xb = dst_xbits((uint32_t)n); sl = 5u; }
        else xb = bl_xbits((uint32_t)n);
        opt_len += (uint32_t)T.wfreq[n] * (bits + xb);
        if (tk.kind != 2) static_len += (uint32_t)T.wfreq[n] * (sl + xb);
    }
    if (overflow) {
        do {
            uint32_t bits = (uint32_t)tk.max_len - 1u;
            while (T.bl_count[bits] == 0) bits--;
            T.bl_count[bits]--; T.bl_count[bits + 1u] += 2; T.bl_count[tk.max_len]--;
            overflow -= 2;
        } while (overflow > 0);
        for (uint32_t bits = (uint32_t)tk.max_len; bits != 0u; bits--) {
            int n = T.bl_count[bits];
            while (n != 0) {
                const int m = T.heap[--h];
                if (m > max_code) continue;
                if (T.wlen[m] != bits) {
                    opt_len += bits * T.wfreq[m];
                    opt_len -= (uint32_t)T.wlen[m] * T.wfreq[m];
                    T.wlen[m] = (uint16_t)bits;
                }
                n--;
            }
        }
    }
    // gen_codes (trees.c:280-312)
    uint16_t next[MAX_BITS + 1];
    uint32_t c = 0;
    for (int bits = 1; bits <= MAX_BITS; bits++) { c = (c + T.bl_count[bits - 1]) << 1; next[bits] = (uint16_t)c; }
    for (int n = 0; n < tk.elems; n++) {
        const uint32_t l = n <= max_code ? T.wlen[n] : 0u;
        len_out[n] = (uint16_t)l;
        code_out[n] = l ? (uint16_t)(__brev((uint32_t)next[l]++) >> (32u - l)) : 0;
    }
    return max_code;
}

// scan_tree (emit == nullptr: count into bfreq) / send_tree (trees.c:411-519) over lens[0..max_code]
__device__ void rle_lengths(BlockWs& T, uint16_t* lens, int max_code, Emitter* emit) {
    int prevlen = -1, nextlen = lens[0], count = 0, max_count = 7, min_count = 4;
    if (nextlen == 0) { max_count = 138; min_count = 3; }
    if (!emit) lens[max_code + 1] = 0xffff;                 // guard; send_tree still sees it (trees.c:419 never resets it)
    for (int n = 0; n <= max_code; n++) {
        const int curlen = nextlen; nextlen = lens[n + 1];
        if (++count < max_count && curlen == nextlen) continue;
        if (count < min_count) {
            if (emit) { do emit->put_serial(T.bcode[curlen], T.blen[curlen]); while (--count); }
            else T.bfreq[curlen] += (uint32_t)count;
        } else if (curlen != 0) {
            if (curlen != prevlen) {
                if (emit) { emit->put_serial(T.bcode[curlen], T.blen[curlen]); count--; }
                else T.bfreq[curlen]++;
            }
            if (emit) { emit->put_serial(T.bcode[16], T.blen[16]); emit->put_serial((uint32_t)(count - 3), 2); }
            else T.bfreq[16]++;
        } else if (count <= 10) {
            if (emit) { emit->put_serial(T.bcode[17], T.blen[17]); emit->put_serial((uint32_t)(count - 3), 3); }
            else T.bfreq[17]++;
        } else {
            if (emit) { emit->put_serial(T.bcode[18], T.blen[18]); emit->put_serial((uint32_t)(count - 11), 7); }
            else T.bfreq[18]++;
        }
        count = 0; prevlen = curlen;
        if (nextlen == 0) { max_count = 138; min_count = 3; }
        else if (curlen == nextlen) { max_count = 6; min_count = 3; }
        else { max_count = 7; min_count = 4; }
    }
}

// compress_block (trees.c:708-741) for tokens [t0, t1): dynamic codes from T, or the fixed code
template <bool kDynamic>
__device__ void emit_tokens(Emitter& E, const BlockWs& T, const uint32_t* __restrict__ tok, uint32_t t0, uint32_t t1, unsigned lane) {
    uint32_t next = t0 + lane < t1 ? __ldcs(tok + t0 + lane) : 0u;
    for (uint32_t base = t0; base < t1; base += 32u) {
        const uint32_t t = next;
        const uint32_t nidx = base + 32u + lane;
        next = nidx < t1 ? __ldcs(tok + nidx) : 0u;
        uint64_t bits = 0; uint32_t nb = 0;
        if (base + lane < t1) {
            if (!kDynamic) { uint32_t b32; fixed_code_token(t, b32, nb); bits = b32; }
            else if (!(t & kTokMatch)) { const uint32_t c = t & 0xffu; bits = T.lcode[c]; nb = T.llen[c]; }
            else {
                uint32_t ls, lx, lxb, ds, dx, dxb;
                len_symbol(((t >> 16) & 0x1ffu) - 3u, ls, lx, lxb);
                dist_symbol((t & 0xffffu) - 1u, ds, dx, dxb);
                bits = T.lcode[257u + ls]; nb = T.llen[257u + ls];
                bits |= (uint64_t)lx << nb; nb += lxb;
                bits |= (uint64_t)T.dcode[ds] << nb; nb += T.dlen[ds];
                bits |= (uint64_t)dx << nb; nb += dxb;
            }
        }
        E.put_warp(bits, nb, lane);
    }
}

// zng_tr_flush_block (trees.c:625-703) for tokens [t0, t1), which cover the chunk bytes from bstart on.  Returns the
// number of bytes the block covers.  `rest` != 0: this is the flush at the end of the input (the block covers the rest of
// the chunk, len - bstart bytes, and fill_window has run at lookahead 0); otherwise the block was flushed because it is full.
__device__ uint32_t flush_block(Emitter& E, BlockWs& T, const uint32_t* __restrict__ tok, uint32_t t0, uint32_t t1,
                                const uint8_t* src, uint32_t bstart, uint32_t len, bool rest, int last, unsigned lane, int nostore = -1) {
    uint32_t opt_lenb = 0, static_lenb = 0, span = 0;
    int max_blindex = 0, lmax = 0, dmax = 0;
    const uint32_t nsym = t1 - t0;
    if (nsym != 0u) {
        for (uint32_t i = lane; i < (uint32_t)L_CODES + 2u; i += 32u) T.lfreq[i] = 0;
        if (lane < (uint32_t)D_CODES + 2u) T.dfreq[lane] = 0;
        if (lane < (uint32_t)BL_CODES + 1u) T.bfreq[lane] = 0;
        __syncwarp();
        // zng_tr_tally_lit / _dist (deflate_p.h:61-98), four tokens per lane in flight; the byte span rides along
        for (uint32_t base = t0; base < t1; base += 128u) {
            uint32_t tk[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const uint32_t i = base + 32u * u + lane; tk[u] = i < t1 ? __ldcs(tok + i) : kTokEnd; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t t = tk[u];
                if (t == kTokEnd) continue;
                if (!(t & kTokMatch)) { atomicAdd(&T.lfreq[t & 0xffu], 1u); span += 1u; }
                else {
                    uint32_t ls, lx, lxb, ds, dx, dxb;
                    len_symbol(((t >> 16) & 0x1ffu) - 3u, ls, lx, lxb);
                    dist_symbol((t & 0xffffu) - 1u, ds, dx, dxb);
                    atomicAdd(&T.lfreq[257u + ls], 1u);
                    atomicAdd(&T.dfreq[ds], 1u);
                    span += (t >> 16) & 0x1ffu;
                }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) span += __shfl_xor_sync(ZB_FULL, span, d);
        __syncwarp();
        if (lane == 0) {
            T.lfreq[256] = 1;                               // init_block: END_BLOCK
            uint32_t opt_len = 0, static_len = 0;
            lmax = build_huffman(T, T.lfreq, TreeKind{L_CODES, 15, 0}, T.lcode, T.llen, opt_len, static_len);
            dmax = build_huffman(T, T.dfreq, TreeKind{D_CODES, 15, 1}, T.dcode, T.dlen, opt_len, static_len);
            rle_lengths(T, T.llen, lmax, nullptr);           // build_bl_tree (trees.c:525-552)
            rle_lengths(T, T.dlen, dmax, nullptr);
            uint32_t dummy = 0;
            build_huffman(T, T.bfreq, TreeKind{BL_CODES, 7, 2}, T.bcode, T.blen, opt_len, dummy);
            for (max_blindex = BL_CODES - 1; max_blindex >= 3; max_blindex--)
                if (T.blen[bl_order((uint32_t)max_blindex)] != 0) break;
            opt_len += 3u * ((uint32_t)max_blindex + 1u) + 5u + 5u + 4u;
            opt_lenb = (opt_len + 3u + 7u) >> 3;
            static_lenb = (static_len + 3u + 7u) >> 3;
            if (static_lenb <= opt_lenb) opt_lenb = static_lenb;
        }
        opt_lenb = __shfl_sync(ZB_FULL, opt_lenb, 0);
        static_lenb = __shfl_sync(ZB_FULL, static_lenb, 0);
    }
    __syncwarp();
    // what the block covers, and whether it may still be stored: after the tail slide (deflate.c:1285-1299) a block that
    // began below 32768 has block_start < 0, FLUSH_BLOCK passes buf == NULL (deflate_p.h:104-112) and trees.c:673 refuses
    uint32_t stored_len; bool slid;
    if (rest) { stored_len = len - bstart; slid = len >= kSlideAt; }
    else {
        const uint32_t tl = __ldcs(tok + t1 - 1u);
        const uint32_t s_last = bstart + span - ((tl & kTokMatch) ? ((tl >> 16) & 0x1ffu) : 1u);   // parser position before the last symbol
        stored_len = span; slid = s_last >= kSlideAt && len - s_last < 262u;
    }
    // nostore >= 0: the parser kept the reference's own block_start through its slides (primed chunks, deflate_window.cu) and says
    // whether this block began before one
    const bool can_store = nostore >= 0 ? nostore == 0 : !(slid && bstart < kWSize);
    const uint8_t* raw = src + bstart;
    if (stored_len + 4u <= opt_lenb && can_store) {         // zng_tr_stored_block (trees.c:592-609)
        if (lane == 0) {
            E.put_serial((uint32_t)last, 3);
            E.bitpos = (E.bitpos + 7u) & ~7u;
            E.put_serial(stored_len & 0xffffu, 16); E.put_serial((~stored_len) & 0xffffu, 16);
        }
        E.bitpos = __shfl_sync(ZB_FULL, E.bitpos, 0);
        for (uint32_t base = 0; base < stored_len; base += 128u) {
            const uint32_t i = base + 4u * lane;
            uint32_t wv = 0, nby = 0;
            if (i + 8u <= stored_len) {                        // interior: two aligned words, funnel-shifted (no read past the block)
                const uintptr_t a = reinterpret_cast<uintptr_t>(raw + i);
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
                wv = __funnelshift_r(__ldg(wp), __ldg(wp + 1), (uint32_t)(a & 3u) << 3);
                nby = 4u;
            } else if (i < stored_len) {
                nby = min(4u, stored_len - i);
                for (uint32_t b = 0; b < nby; b++) wv |= (uint32_t)raw[i + b] << (8u * b);
            }
            E.put_warp(wv, 8u * nby, lane);
        }
    } else if (static_lenb == opt_lenb) {                   // static trees (also the empty block: sym_next == 0)
        if (lane == 0) E.put_serial((1u << 1) + (uint32_t)last, 3);
        E.bitpos = __shfl_sync(ZB_FULL, E.bitpos, 0);
        emit_tokens<false>(E, T, tok, t0, t1, lane);
        E.put_warp(0, lane == 0 ? 7u : 0u, lane);           // END_BLOCK under the fixed code: 7 zero bits
    } else {                                                // dynamic trees: send_all_trees (trees.c:559-587)
        if (lane == 0) {
            E.put_serial((2u << 1) + (uint32_t)last, 3);
            const int blcodes = max_blindex + 1;
            E.put_serial((uint32_t)(lmax + 1 - 257), 5); E.put_serial((uint32_t)(dmax + 1 - 1), 5); E.put_serial((uint32_t)(blcodes - 4), 4);
            for (int r = 0; r < blcodes; r++) E.put_serial(T.blen[bl_order((uint32_t)r)], 3);
            rle_lengths(T, T.llen, lmax, &E);
            rle_lengths(T, T.dlen, dmax, &E);
        }
        E.bitpos = __shfl_sync(ZB_FULL, E.bitpos, 0);
        __syncwarp();
        if ((E.bitpos >> 5) - E.flushed >= kBStageSeg) E.flush(((E.bitpos >> 5) / kBStageSeg) * kBStageSeg, lane);
        emit_tokens<true>(E, T, tok, t0, t1, lane);
        E.put_warp(T.lcode[256], lane == 0 ? (uint32_t)T.llen[256] : 0u, lane);
    }
    if (last) { E.bitpos = (E.bitpos + 7u) & ~7u; }          // bi_windup
    __syncwarp();
    return stored_len;
}

__global__ void __launch_bounds__(kBlkWarps * 32)
block_emit_kernel(const uint8_t* __restrict__ in, const uint32_t* __restrict__ tokens, uint32_t tok_stride,
                  const uint32_t* __restrict__ ntok, size_t n, uint32_t chunk, uint32_t nchunks, int last,
                  uint8_t* __restrict__ out, size_t out_stride, uint32_t* __restrict__ sizes, const uint8_t* __restrict__ blkflags,
                  uint32_t* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char blk_smem[];
    BlockWs& T = reinterpret_cast<BlockWs*>(blk_smem)[threadIdx.x >> 5];
    const unsigned lane = lane_id();
    for (uint32_t i = lane; i < kBStageWords; i += 32u) T.stage[i] = 0u;
    __syncwarp();
    // chunks are claimed from a counter (stored blocks cost a fraction of dynamic ones: a fixed assignment leaves warps idle at the end)
    for (uint32_t ci = blockIdx.x * kBlkWarps + (threadIdx.x >> 5);; ci += gridDim.x * kBlkWarps) {
        if (counter) { if (lane == 0) ci = atomicAdd(counter, 1u); ci = __shfl_sync(ZB_FULL, ci, 0); }
        if (ci >= nchunks) break;
        const size_t off = (size_t)ci * chunk;
        const uint32_t len = (uint32_t)min((size_t)chunk, n - off);
        const uint8_t* src = in + off;
        const uint32_t* tok = tokens + (size_t)ci * tok_stride;
        const uint32_t nt = ntok[ci];
        Emitter E{T.stage, out + (size_t)ci * out_stride, 0u, 0u};
        uint32_t t0 = 0, bstart = 0, blk = 0;
        const uint8_t* bf = blkflags ? blkflags + (size_t)ci * 8u : nullptr;
        // deflate_fast.c:93-94: a block is flushed as soon as it holds 16383 symbols; deflate_fast.c:96-103: then the rest (also an
        // empty last block for Z_FINISH).  One call site: the block writer is inlined once.
        for (;;) {
            const bool full = nt - t0 >= kSymEnd;
            if (!full && !(last || nt > t0)) break;
            bstart += flush_block(E, T, tok, t0, full ? t0 + kSymEnd : nt, src, bstart, len, !full, full ? 0 : last, lane, bf ? (int)bf[min(blk, 7u)] : -1);
            if (!full) break;
            t0 += kSymEnd; blk++;
        }
        if (!last) {                                        // deflate.c:1064-1065 zng_tr_stored_block(NULL, 0, 0)
            if (lane == 0) {
                E.bitpos += 3;
                E.bitpos = (E.bitpos + 7u) & ~7u;
                E.or_bits(E.bitpos, 0xffff0000u, 32);
                E.bitpos += 32;
            }
            E.bitpos = __shfl_sync(ZB_FULL, E.bitpos, 0);
        }
        E.flush((E.bitpos + 31u) >> 5, lane);
        if (lane == 0) sizes[ci] = E.bitpos >> 3;
    }
}

size_t deflate_fast_prev_bytes(uint32_t nsmid) { return (size_t)nsmid * 64u * kWSize * sizeof(uint16_t); }
// + 64 bytes: the warp-wide compare behind a 128-byte prefix (level 6) may read a few words past the last image
size_t deflate_fast_tail_bytes(uint32_t nsmid) { return (size_t)nsmid * 64u * kTailWords * sizeof(uint32_t) + 64u; }

cudaError_t launch_fast_parse(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks, uint32_t* tokens, uint32_t tok_stride,
                              uint32_t* ntok, uint32_t* counter, uint16_t* heads, uint16_t* prevs, uint32_t* tails,
                              unsigned long long* sm_slots, int num_sms, int chains_per_sm, int have_prev, int level, cudaStream_t stream,
                              const StreamSync* sync, int dyn_smem) {
    if (nchunks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    // levels 5-6, device-resident and slab calls: K2c, one CTA per chunk -- opt-in (env ZNG_B200_K2C=1): bit-exact, but at 2.98 / 5.5 GB/s
    // (levels 6 / 5) still behind the warp-per-chain parser's 3.24 / 6.8 (profiles/r2_k2c.md says what holds it back)
    static const int k2c = [] { const char* e = getenv("ZNG_B200_K2C"); return e ? atoi(e) : 0; }();
    if (level >= 5 && !sync && k2c) {
        const int smem = (int)sizeof(CtaSmem);
        static const int per_sm = [] { const char* e = getenv("ZNG_B200_K2C_CTAS"); const int v = e ? atoi(e) : 3; return v >= 1 && v <= 3 ? v : 3; }();
        const int carve = per_sm == 3 ? 100 : (per_sm == 2 ? 64 : 32);            // what is not shared memory is L1 for the chunks' bytes
        const uint32_t g = nchunks < (uint32_t)(per_sm * num_sms) ? nchunks : (uint32_t)(per_sm * num_sms);
        if (level == 5) {
            e = cudaFuncSetAttribute(medium_cta_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            cudaFuncSetAttribute(medium_cta_kernel<5>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            medium_cta_kernel<5><<<g, kCtaWarps * 32, smem, stream>>>(in, n, chunk, nchunks, tokens, tok_stride, ntok, counter, heads, prevs, tails, sm_slots, have_prev);
        } else {
            e = cudaFuncSetAttribute(medium_cta_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            cudaFuncSetAttribute(medium_cta_kernel<6>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
            medium_cta_kernel<6><<<g, kCtaWarps * 32, smem, stream>>>(in, n, chunk, nchunks, tokens, tok_stride, ntok, counter, heads, prevs, tails, sm_slots, have_prev);
        }
        return cudaGetLastError();
    }
    uint32_t ctas_per_sm = ((uint32_t)chains_per_sm + kFastWarps - 1u) / kFastWarps;
    uint32_t grid = (uint32_t)num_sms * ctas_per_sm;
    const uint64_t need = ((uint64_t)nchunks + kFastWarps - 1u) / kFastWarps;
    if (need < grid) grid = (uint32_t)need;
    const StreamSync sy = sync ? *sync : StreamSync{};
#define ZB_LAUNCH_PARSE(L) fast_parse_kernel<L><<<grid, kFastWarps * 32, dyn_smem, stream>>>(in, n, chunk, nchunks, tokens, tok_stride, ntok, \
                                                                                            counter, heads, prevs, tails, sm_slots, have_prev, sy)
    switch (level) {
        case 2: ZB_LAUNCH_PARSE(2); break;
        case 3: ZB_LAUNCH_PARSE(3); break;
        case 4: ZB_LAUNCH_PARSE(4); break;
        case 5: ZB_LAUNCH_PARSE(5); break;
        case 6: ZB_LAUNCH_PARSE(6); break;
        default: return cudaErrorInvalidValue;
    }
#undef ZB_LAUNCH_PARSE
    return cudaGetLastError();
}

cudaError_t launch_block_emit(const uint8_t* in, const uint32_t* tokens, uint32_t tok_stride, const uint32_t* ntok, size_t n,
                              uint32_t chunk, uint32_t nchunks, int last, uint8_t* out, size_t out_stride, uint32_t* sizes,
                              int num_sms, cudaStream_t stream, int co_carve, const uint8_t* blkflags, uint32_t* counter) {
    if (nchunks == 0) return cudaSuccess;
    if (counter) { cudaError_t e0 = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream); if (e0 != cudaSuccess) return e0; }
    const int smem = (int)(sizeof(BlockWs) * kBlkWarps);
    cudaError_t e = cudaFuncSetAttribute(block_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    // co_carve < 0 (the kernel has the SMs to itself): ask for all of the shared memory.  The driver's default carve-out stops at
    // 132 KiB = 3 CTAs (12 warps) per SM (sm__warps_active 18 % in profiles/r2_block_emit_ncu_summary.txt); 5 CTAs fit.
    if (co_carve < 0) {
        static const int own = [] { const char* e = getenv("ZNG_B200_EMIT_CARVE"); return e ? atoi(e) : 100; }();   // -1 = the driver's default
        co_carve = own;
    }
    cudaFuncSetAttribute(block_emit_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co_carve);   // see launch_static_emit
    uint32_t grid = (uint32_t)num_sms * (uint32_t)((227 * 1024) / (smem + 1024));     // as many CTAs as the shared memory of an SM holds
    const uint32_t need = (nchunks + kBlkWarps - 1u) / kBlkWarps;
    if (grid > need) grid = need;
    block_emit_kernel<<<grid, kBlkWarps * 32, smem, stream>>>(in, tokens, tok_stride, ntok, n, chunk, nchunks, last, out, out_stride, sizes, blkflags, counter);
    return cudaGetLastError();
}

}  // namespace zb
