// deflate_quick.cu -- K1: level-1 (deflate_quick) compression of independent <= 64 KiB chunks.
//
// Reference semantics reproduced bit-exactly (files under /root/reference):
//   deflate_quick.c:47-130        greedy single-probe parse, static-Huffman emission
//   insert_string_tpl.h:58-75     quick_insert_string (head[] update only at visited positions)
//   arch/generic/compare256_c.c   first-mismatch compare (here: 32 lanes x 8 bytes + ballot/ffs)
//   trees_emit.h:102-225          zng_tr_emit_lit / zng_tr_emit_dist / emit_tree / emit_end_block
//   deflate.c:1061-1083           empty stored block appended for Z_SYNC_FLUSH / Z_FULL_FLUSH
//
// B200 mapping (v3, "many chains per SM, parse and emit split").  The parse of one chunk is a
// serial dependency chain (which positions get hashed depends on every earlier match decision);
// ncu on v1 (one chunk per SM, window + head table in shared memory) showed the SM 93% idle: one
// warp advances one chain at ~0.1 IPC.  So the unit of parallelism is the CHAIN = one warp, and the
// SM runs as many as its register file allows (up to 48):
//   K1a quick_parse_kernel   one warp per chain, no shared memory, no block-level synchronisation.
//     * the 65536-entry u16 hash-head table of deflate_state (deflate.h:232, 128 KiB) lives in a
//       per-warp slab of global memory (L2 / HBM resident), accessed with ld/st.global.cg and
//       cleared per chunk (CLEAR_HASH, deflate.c:182); its latency is hidden by the other chains;
//     * the window is the input itself, read through L1 (ld.global.nc): no staging copy;
//     * output: the LZ77 token list of the chunk (u32 per token) in a global scratch buffer.
//   K1b static_emit_kernel   one warp per chunk: tokens -> fixed-Huffman bit stream.  Per 32 tokens:
//       code bits, warp prefix scan of the bit lengths, OR into a shared staging ring, 1 KiB segments
//       stored to global memory as uint4.  Embarrassingly parallel and short (a few % of K1a).
// PARSER.  The warp speculates: 32 lanes hash 32 consecutive positions against the table state at
//   the window start, test their candidates and measure short matches (< 12 bytes) in parallel; a
//   ballot/ffs walk then replays the reference's decisions (literal runs are accepted wholesale, a
//   match costs one shuffle, or one warp-wide compare when it is 12 bytes or longer).  A lane whose
//   hash equals that of an earlier VISITED lane of the same window saw a stale candidate; the window
//   is cut at the first such lane and restarted there, which keeps the result exact.
// Per-chunk CRC-32 / Adler-32 come from the K3 tile kernel (checksum.cu) launched on the same stream.
#include "common.cuh"
#include "kernels.h"
#include "lz_ops.cuh"
#include <cstdlib>

namespace zb {

constexpr int      kParseWarps   = 4;             // chains per CTA of the parse kernel
constexpr int      kEmitWarps    = 4;             // chunks in flight per CTA of the emit kernel
constexpr uint32_t kStageWords   = 1024;          // u32 words (4 KiB) per emitting warp
constexpr uint32_t kStageSeg     = 256;           // words per flush segment (1 KiB)

// The chunk as the parser sees it: a word-aligned base in global memory plus a byte skew.
struct Window {
    const uint32_t* w;             // 4-byte aligned
    uint32_t skew;                 // 0..3: chunk byte 0 is byte `skew` of w[0]
    // An ordinary cached load (L1 + L2), NOT the non-coherent path: in the streamed host pipeline the copy engine is still writing
    // other parts of this allocation while the kernel runs (each chunk is released to the parser only after it and its successor
    // have landed), which is outside the read-only contract of ld.global.nc / __ldg.
    __device__ __forceinline__ uint32_t word(uint32_t i) const {
        uint32_t r;
        asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(r) : "l"(w + i));    // volatile: stays behind the wait for the chunk's delivery
        return r;
    }
};

// ---------------------------------------------------------------- parser
// compare256 (compare256_c.c:12-43) is lz_ops.cuh:warp_compare_bytes -- the one warp-wide first-mismatch of this library.

// Parse one chunk; tokens go to tok[0..count) in global memory (coalesced: the visited lanes of a
// window write consecutive slots), followed by the kTokEnd marker.  Returns the token count.

__device__ uint32_t quick_parse_warp(const Window W, uint32_t n, uint16_t* head, uint32_t* __restrict__ tok) {
    const unsigned lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    uint32_t wr = 0;
    uint32_t p = 0;
    while (p < n) {
        const uint32_t q = p + lane;
        const bool inb = q < n;
        const bool act = q + kWantMin <= n;                  // deflate_quick.c:88 lookahead >= WANT_MIN_MATCH
        // 12 bytes at q: v = bytes 0..3 (hashed), x = bytes 4..11
        uint32_t v; uint64_t x;
        {
            const uint32_t qb = q + W.skew, i = qb >> 2, sh = (qb & 3u) << 3;
            const uint32_t a0 = W.word(i), a1 = W.word(i + 1), a2 = W.word(i + 2), a3 = W.word(i + 3);
            v = __funnelshift_r(a0, a1, sh);
            x = (uint64_t)__funnelshift_r(a1, a2, sh) | ((uint64_t)__funnelshift_r(a2, a3, sh) << 32);
        }
        const uint32_t h = hash4(v);
        uint32_t cand = 0u;
        if (act) cand = (uint32_t)__ldcg(head + h);
        // A chunk of 65275..65535 bytes is slid when the parser stands at 65274 (deflate.c:1285-1299): slide_hash zeroes every
        // entry below 32768, and as deflate_quick has no hash_head != 0 test such a slot then names window position 0 =
        // original position 32768, at distance MAX_DIST exactly.  (Later positions are out of its reach.)
        if (q == kWSize + kMaxDist && n < kChunkMax && cand < kWSize) cand = kWSize;
        // "touch" loads: real loads whose values are never needed; the fill they trigger is the 512 bytes of window the
        // next steps will read (measured +5 %; prefetch instructions and head-entry prefetches gained nothing)
        uint32_t sink = 0;
        if (lane < 16u && p + 160u + 512u <= n) sink = W.word((((p + W.skew + 160u) & ~31u) >> 2) + 8u * lane);
        // deflate_quick.c:90-92: 0 < dist <= MAX_DIST;  :96-99: 2 bytes equal and compare256+2 >= 4  <=>  4 bytes equal
        uint32_t slen = 0;                                    // 0 none, 4..11 exact, 12 = "12 or more"
        if (act && (q - cand - 1u) < kMaxDist) {
            const uint32_t cb = cand + W.skew, i = cb >> 2, sh = (cb & 3u) << 3;
            const uint32_t b0 = W.word(i), b1 = W.word(i + 1), b2 = W.word(i + 2), b3 = W.word(i + 3);
            if (__funnelshift_r(b0, b1, sh) == v) {
                const uint64_t y = (uint64_t)__funnelshift_r(b1, b2, sh) | ((uint64_t)__funnelshift_r(b2, b3, sh) << 32);
                const uint64_t d = x ^ y;
                slen = d ? 4u + ((unsigned)(__ffsll((long long)d) - 1) >> 3) : 12u;
                if (slen < 12u) slen = min(slen, n - q);     // deflate_quick.c:100-101 clip to lookahead (>= 4 here)
            }
        }
        const unsigned peers = __match_any_sync(ZB_FULL, act ? h : (0x10000u + lane));
        const unsigned low = peers & lt;                     // earlier lanes of this window with the same hash
        const unsigned M = __ballot_sync(ZB_FULL, slen != 0u);
        const unsigned nl = min(32u, n - p);                 // lanes that hold a byte
        uint32_t mytok = v & 0xffu;
        // ---- walk 1: replay the greedy decisions on the speculative data -> covered lanes
        unsigned cur = 0, covered = 0;
        while (cur < nl) {
            const unsigned rest = M & ~lane_range(0, cur);
            if (rest == 0u) { cur = nl; break; }
            const unsigned k = (unsigned)(__ffs(rest) - 1);  // next lane claiming a match
            uint32_t len = __shfl_sync(ZB_FULL, slen, k);
            if (len >= 12u) {
                // a long match is measured by the whole warp; it is only valid if lane k is not stale
                // (checked below) -- a stale long match is cut away before it is used.
                const uint32_t ck = __shfl_sync(ZB_FULL, cand, k);
                const uint32_t qk = p + k;
                len = 12u + warp_compare_bytes(W, qk + 12u + W.skew, ck + 12u + W.skew, 256u, lane);
                len = min(len, n - qk);
                len = min(len, kMaxMatch);                   // deflate_quick.c:102-103
                if (lane == k) slen = len;
            }
            covered |= lane_range(k + 1u, k + len);
            cur = k + len;
        }
        // ---- walk 2: visited lanes, first stale one (if any) cuts the window
        unsigned V = lane_range(0, min(cur, nl)) & ~covered;
        const unsigned S = __ballot_sync(ZB_FULL, ((V >> lane) & 1u) && (low & V) != 0u);
        if (S) {
            const unsigned j = __ffs(S) - 1u;
            V &= lane_range(0, j);
            cur = j;
        }
        const bool vis = (V >> lane) & 1u;
        if (vis && slen) mytok = kTokMatch | (slen << 16) | (q - cand);
        if (vis && act) {                                    // insert_string_tpl.h:70-73 (visited positions only)
            __stcg(head + h, (uint16_t)q);
        }
        if (vis && inb) __stcs(tok + wr + __popc(V & lt), mytok);
        wr += __popc(V);
        asm volatile("{ .reg .pred pp; setp.eq.u32 pp, %0, %1; @pp nanosleep.u32 1; }" :: "r"(sink), "r"(0x5a5a5a5au));   // keeps the touch loads alive
        p += cur;
        __syncwarp();                                        // orders this window's head stores before the next lookups
    }
    if (lane == 0) __stcs(tok + wr, kTokEnd);
    return wr;
}

// ---------------------------------------------------------------- primed parser (pigz's dependent-chunk mode)
// Reference call sequence per chunk: a fresh zng_deflateInit2(1, -15), zng_deflateSetDictionary(the 32768 stream bytes in
// front of the chunk, deflate.c:456-512), one zng_deflate(flush).  Positions are ABSOLUTE from the start of the dictionary
// (0 .. D + len, D = 32768 or 0 for the first chunk) and the head table holds 32-bit absolute positions, so the two window
// slides of such a chunk (deflate.c:1285-1299) need no pass over the table: an entry slide_hash would have zeroed lies
// more than MAX_DIST behind strstart.  What the slides and refills do change is restated:
//   * `base` = absolute position of window index 0; a slot that is empty or below `base` names window index 0
//     (deflate_quick.c:88-92 has no hash_head != 0 test) -- in reach only right after a slide at strstart 65274;
//   * deflateSetDictionary hashes position D-3 with the still-zero byte behind the dictionary; the first fill_window of
//     zng_deflate re-inserts D-3 and inserts the pending strings D-2, D-1 (s->insert); every later refill that reads
//     input inserts strstart-1 (deflate.c:1321-1336);
//   * R = bytes loaded so far: a position is only parsed once 262 bytes of lookahead are loaded (or the input has ended).
__device__ __forceinline__ uint32_t load4_upto(const Window& W, uint32_t pos, uint32_t limit) {    // bytes at pos.., zero from `limit` on
    const uint32_t qb = pos + W.skew, i = qb >> 2, sh = (qb & 3u) << 3;
    uint32_t v = __funnelshift_r(W.word(i), W.word(i + 1), sh);
    if (pos + 4u > limit) v &= pos >= limit ? 0u : (0xffffffffu >> (8u * (pos + 4u - limit)));
    return v;
}
__device__ __forceinline__ void primed_insert(const Window& W, uint32_t* head, uint32_t pos, uint32_t limit, unsigned lane) {
    const uint32_t h = hash4(load4_upto(W, pos, limit));
    if (lane == 0) __stcg(head + h, pos);
    __syncwarp();
}

__device__ uint32_t primed_parse_warp(const Window W, uint32_t D, uint32_t N, uint32_t* head, uint32_t* __restrict__ tok) {
    const unsigned lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    for (uint32_t p0 = 0; p0 + 2u < D; p0 += 32u) {           // insert_string(s, 0, D - 2): the highest position of a hash wins
        const uint32_t q = p0 + lane;
        const bool ins = q + 2u < D;
        const uint32_t h = hash4(load4_upto(W, q, D));
        const unsigned peers = __match_any_sync(ZB_FULL, ins ? h : (0x10000u + lane));
        if (ins && (peers & ~lt & ~(1u << lane)) == 0u) __stcg(head + h, q);
    }
    __syncwarp();
    uint32_t wr = 0, p = D, base = 0, R = D, pend = D ? 2u : 0u;
    for (;;) {
        if (R - p < 262u) {                                   // fill_window at an iteration boundary (deflate.c:1272-1340)
            do {
                if (p - base >= kWSize + kMaxDist) base += kWSize;
                if (R == N) break;
                const uint32_t room = base + kChunkMax - R;
                R = (N - R < room) ? N : R + room;
                if (R - p + pend >= 3u) {
                    const uint32_t str = p - pend;
                    if (str - base >= 1u) primed_insert(W, head, str - 1u, R, lane);
                    uint32_t cnt = pend;
                    if (R - p == 1u) cnt--;
                    for (uint32_t k = 0; k < cnt; k++) primed_insert(W, head, str + k, R, lane);
                    pend -= cnt;
                }
            } while (R - p < 262u && R != N);
            if (R == p) break;
        }
        const unsigned nl = min(32u, R == N ? N - p : R - 261u - p);      // lanes whose own iteration needs no refill first
        const uint32_t q = p + lane;
        const bool inb = lane < nl;
        const bool act = inb && q + kWantMin <= R;             // deflate_quick.c:88 lookahead >= WANT_MIN_MATCH
        uint32_t v; uint64_t x;
        {
            const uint32_t qb = q + W.skew, i = qb >> 2, sh = (qb & 3u) << 3;
            const uint32_t a0 = W.word(i), a1 = W.word(i + 1), a2 = W.word(i + 2), a3 = W.word(i + 3);
            v = __funnelshift_r(a0, a1, sh);
            x = (uint64_t)__funnelshift_r(a1, a2, sh) | ((uint64_t)__funnelshift_r(a2, a3, sh) << 32);
        }
        const uint32_t h = hash4(v);
        uint32_t cand = 0u;
        if (act) {
            cand = __ldcg(head + h);
            const uint32_t bq = (R == N && q + 262u > N && q - base >= kWSize + kMaxDist) ? base + kWSize : base;   // slid at or before q
            if (cand < bq) cand = bq;
        }
        uint32_t slen = 0;                                    // 0 none, 4..11 exact, 12 = "12 or more"
        if (act && (q - cand - 1u) < kMaxDist) {
            const uint32_t cb = cand + W.skew, i = cb >> 2, sh = (cb & 3u) << 3;
            const uint32_t b0 = W.word(i), b1 = W.word(i + 1), b2 = W.word(i + 2), b3 = W.word(i + 3);
            if (__funnelshift_r(b0, b1, sh) == v) {
                const uint64_t d = x ^ ((uint64_t)__funnelshift_r(b1, b2, sh) | ((uint64_t)__funnelshift_r(b2, b3, sh) << 32));
                slen = d ? 4u + ((unsigned)(__ffsll((long long)d) - 1) >> 3) : 12u;
                if (slen < 12u) slen = min(slen, R - q);
            }
        }
        const unsigned peers = __match_any_sync(ZB_FULL, act ? h : (0x10000u + lane));
        const unsigned low = peers & lt;
        const unsigned M = __ballot_sync(ZB_FULL, slen != 0u);
        uint32_t mytok = v & 0xffu;
        unsigned cur = 0, covered = 0;
        while (cur < nl) {
            const unsigned rest = M & ~lane_range(0, cur);
            if (rest == 0u) { cur = nl; break; }
            const unsigned k = (unsigned)(__ffs(rest) - 1);
            uint32_t len = __shfl_sync(ZB_FULL, slen, k);
            if (len >= 12u) {
                const uint32_t ck = __shfl_sync(ZB_FULL, cand, k);
                const uint32_t qk = p + k;
                len = 12u + warp_compare_bytes(W, qk + 12u + W.skew, ck + 12u + W.skew, 256u, lane);
                len = min(min(len, R - qk), kMaxMatch);
                if (lane == k) slen = len;
            }
            covered |= lane_range(k + 1u, k + len);
            cur = k + len;
        }
        unsigned V = lane_range(0, min(cur, nl)) & ~covered;
        const unsigned S = __ballot_sync(ZB_FULL, ((V >> lane) & 1u) && (low & V) != 0u);
        if (S) {
            const unsigned j = __ffs(S) - 1u;
            V &= lane_range(0, j);
            cur = j;
        }
        const bool vis = (V >> lane) & 1u;
        if (vis && slen) mytok = kTokMatch | (slen << 16) | (q - cand);
        if (vis && act) __stcg(head + h, q);
        if (vis) __stcs(tok + wr + __popc(V & lt), mytok);
        wr += __popc(V);
        p += cur;
        __syncwarp();
    }
    if (lane == 0) __stcs(tok + wr, kTokEnd);
    return wr;
}

// ---------------------------------------------------------------- emitter (K1b, one warp per chunk)
struct EmitSmem { uint32_t stage[kEmitWarps][kStageWords]; };

__device__ __forceinline__ void stage_flush(uint32_t* stage, uint8_t* out, uint32_t& flushed, uint32_t upto_words, unsigned lane) {
    // store [flushed, upto_words) words of the staging ring to global memory and re-zero them
    __syncwarp();
    while (flushed < upto_words) {
        uint32_t cnt = min(upto_words - flushed, kStageSeg);
        uint32_t base = flushed & (kStageWords - 1u);
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

__device__ __forceinline__ void stage_or(uint32_t* stage, uint32_t bitpos, uint32_t bits, uint32_t nbits) {
    if (nbits == 0u) return;
    uint32_t w = bitpos >> 5, sh = bitpos & 31u;
    uint64_t v = (uint64_t)bits << sh;
    atomicOr(&stage[w & (kStageWords - 1u)], (uint32_t)v);
    uint32_t hi = (uint32_t)(v >> 32);
    if (hi) atomicOr(&stage[(w + 1u) & (kStageWords - 1u)], hi);
}

// returns the chunk's compressed size in bytes
__device__ uint32_t static_emit_warp(uint32_t* stage, const uint32_t* __restrict__ tok, uint32_t ntok, uint8_t* out, int last, bool open_block) {
    const unsigned lane = lane_id();
    uint32_t bitpos = 0, flushed = 0;
    if (open_block) {                                       // trees_emit.h:198-207: (STATIC_TREES<<1)+last, 3 bits
        if (lane == 0) stage_or(stage, 0, (1u << 1) + (uint32_t)last, 3);
        bitpos = 3;
    }
    uint32_t next = lane < ntok ? __ldcs(tok + lane) : 0u;  // software pipeline: one batch of 32 tokens ahead
    for (uint32_t base = 0; base < ntok; base += 32u) {
        const uint32_t t = next;
        const uint32_t nb_idx = base + 32u + lane;
        next = nb_idx < ntok ? __ldcs(tok + nb_idx) : 0u;
        uint32_t bits = 0, nb = 0;
        if (base + lane < ntok) fixed_code_token(t, bits, nb);
        uint32_t incl = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t u = __shfl_up_sync(ZB_FULL, incl, d); if ((int)lane >= d) incl += u; }
        stage_or(stage, bitpos + incl - nb, bits, nb);
        bitpos += __shfl_sync(ZB_FULL, incl, 31);
        if ((bitpos >> 5) - flushed >= kStageSeg) stage_flush(stage, out, flushed, ((bitpos >> 5) / kStageSeg) * kStageSeg, lane);
    }
    __syncwarp();
    // block end + flush marker
    if (lane == 0) {
        if (open_block) { bitpos += 7; }                    // END_BLOCK under the fixed code: 7 zero bits
        if (last) {
            bitpos = (bitpos + 7u) & ~7u;                   // bi_windup
        } else {                                            // deflate.c:1064-1065 zng_tr_stored_block(NULL,0,0)
            bitpos += 3;                                    // (STORED_BLOCK<<1)+0
            bitpos = (bitpos + 7u) & ~7u;
            stage_or(stage, bitpos, 0xffff0000u, 32);       // LEN=0x0000, NLEN=0xffff
            bitpos += 32;
        }
    }
    bitpos = __shfl_sync(ZB_FULL, bitpos, 0);
    stage_flush(stage, out, flushed, (bitpos + 31u) >> 5, lane);
    return bitpos >> 3;
}

// ---------------------------------------------------------------- kernels
__global__ void nsmid_kernel(uint32_t* out) { uint32_t r; asm volatile("mov.u32 %0, %%nsmid;" : "=r"(r)); *out = r; }

// heads: the slab pool.  tail: a zero-padded private copy of the chunks
// from `tail_first` on -- those whose read-ahead (<= kWinPad bytes past the chunk) could leave the
// caller's allocation; every other chunk reads ahead into its successors, whose bytes cannot
// influence the result (lengths are clipped to the chunk).
__global__ void __launch_bounds__(kParseWarps * 32, 12)
quick_parse_kernel(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks,
                   uint32_t* __restrict__ tokens, uint32_t tok_stride, uint32_t* __restrict__ ntok,
                   uint32_t* __restrict__ counter, uint16_t* __restrict__ heads, unsigned long long* __restrict__ sm_slots,
                   const uint8_t* tail, uint32_t tail_first, StreamSync sy) {
    extern __shared__ uint32_t carve_out_only[];              // never touched: see launch_quick_parse (streamed launches)
    const unsigned lane = lane_id();
    const uint32_t sm = smid();
    uint32_t slot = 0;
    if (lane == 0) slot = slot_acquire(sm_slots + sm);
    slot = __shfl_sync(ZB_FULL, slot, 0);
    uint16_t* head = heads + ((size_t)sm * 64u + slot) * 65536u;
    for (;;) {
        uint32_t ci = 0;
        if (lane == 0) ci = atomicAdd(counter, 1u);
        ci = __shfl_sync(ZB_FULL, ci, 0);
        if (ci >= nchunks) break;
        if (sy.ready) {                                        // streamed input: chunk ci is parsed once the copy engine has delivered it
            if (lane == 0) {
                const long long t0 = clock64();
                while (*(volatile const uint32_t*)sy.ready <= ci) {
                    __nanosleep(1000);
                    if (clock64() - t0 > sy.patience) { atomicExch(sy.failed, 1u); break; }
                }
            }
            __syncwarp();
        }
        // CLEAR_HASH (deflate.c:182-184): 128 KiB of zeros into this chain's slab
        {
            uint4* h4 = reinterpret_cast<uint4*>(head);
#pragma unroll 8
            for (uint32_t i = lane; i < 65536u * 2u / 16u; i += 32u) h4[i] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        const size_t off = (size_t)ci * chunk;
        const uint32_t len = (uint32_t)min((size_t)chunk, n - off);
        const uint8_t* src = (ci >= tail_first) ? tail + (size_t)(ci - tail_first) * chunk : in + off;
        Window W;
        W.skew = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
        W.w = reinterpret_cast<const uint32_t*>(src - W.skew);
        const uint32_t cnt = quick_parse_warp(W, len, head, tokens + (size_t)ci * tok_stride);
        if (lane == 0) ntok[ci] = cnt;
        if (sy.done) {                                         // per output slab: how many of its chunks are parsed
            __syncwarp();
            __threadfence();
            if (lane == 0) {
                const uint32_t k = atomicAdd(sy.done + (ci >> sy.done_shift), 1u) + 1u;
                if (sy.host_done) {                            // tell the host (mapped pinned memory) when a slab is complete
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


// Primed chunks: chunk g (global index first + ci) reads its dictionary from the 32768 bytes in front of it; g == 0 has none.
// heads: 65536 x u32 per chain.  tail: a zero-padded private copy of [tail_first * chunk - 32768, n) for the chunks whose
// read-ahead could leave the caller's allocation (tail_dict = dictionary bytes present in front of that copy).
__global__ void __launch_bounds__(kParseWarps * 32, 12)
primed_parse_kernel(const uint8_t* __restrict__ in, size_t n, uint32_t chunk, uint32_t nchunks, uint32_t first,
                    uint32_t* __restrict__ tokens, uint32_t tok_stride, uint32_t* __restrict__ ntok,
                    uint32_t* __restrict__ counter, uint32_t* __restrict__ heads, unsigned long long* __restrict__ sm_slots,
                    const uint8_t* __restrict__ tail, uint32_t tail_first, uint32_t tail_dict) {
    const unsigned lane = lane_id();
    const uint32_t sm = smid();
    uint32_t slot = 0;
    if (lane == 0) slot = slot_acquire(sm_slots + sm);
    slot = __shfl_sync(ZB_FULL, slot, 0);
    uint32_t* head = heads + ((size_t)sm * 64u + slot) * 65536u;
    for (;;) {
        uint32_t ci = 0;
        if (lane == 0) ci = atomicAdd(counter, 1u);
        ci = __shfl_sync(ZB_FULL, ci, 0);
        if (ci >= nchunks) break;
        {
            uint4* h4 = reinterpret_cast<uint4*>(head);
#pragma unroll 8
            for (uint32_t i = lane; i < 65536u * 4u / 16u; i += 32u) h4[i] = make_uint4(0, 0, 0, 0);
        }
        __syncwarp();
        const size_t off = (size_t)ci * chunk;
        const uint32_t len = (uint32_t)min((size_t)chunk, n - off);
        const uint32_t D = (first + ci) > 0u ? kWSize : 0u;
        const uint8_t* src = (ci >= tail_first) ? tail + tail_dict + (size_t)(ci - tail_first) * chunk : in + off;
        src -= D;
        Window W;
        W.skew = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
        W.w = reinterpret_cast<const uint32_t*>(src - W.skew);
        const uint32_t cnt = primed_parse_warp(W, D, D + len, head, tokens + (size_t)ci * tok_stride);
        if (lane == 0) ntok[ci] = cnt;
    }
    __syncwarp();
    if (lane == 0) atomicAnd(sm_slots + sm, ~(1ull << slot));
}

__global__ void __launch_bounds__(kEmitWarps * 32)
static_emit_kernel(const uint32_t* __restrict__ tokens, uint32_t tok_stride, const uint32_t* __restrict__ ntok,
                   size_t n, uint32_t chunk, uint32_t nchunks, int last,
                   uint8_t* __restrict__ out, size_t out_stride, uint32_t* __restrict__ sizes) {
    __shared__ __align__(16) EmitSmem s;
    const unsigned warp = threadIdx.x >> 5;
    uint32_t* stage = s.stage[warp];
    for (uint32_t i = lane_id(); i < kStageWords; i += 32u) stage[i] = 0u;
    __syncwarp();
    for (uint32_t ci = blockIdx.x * kEmitWarps + warp; ci < nchunks; ci += gridDim.x * kEmitWarps) {
        const size_t off = (size_t)ci * chunk;
        const uint32_t len = (uint32_t)min((size_t)chunk, n - off);
        const bool open_block = (len > 0) || last;          // deflate_quick.c:53-63
        const uint32_t sz = static_emit_warp(stage, tokens + (size_t)ci * tok_stride, ntok[ci], out + (size_t)ci * out_stride, last, open_block);
        if (lane_id() == 0) sizes[ci] = sz;
    }
}

size_t deflate_quick_head_bytes(uint32_t nsmid) { return (size_t)nsmid * 64u * 65536u * sizeof(uint16_t); }

cudaError_t query_nsmid(uint32_t* d_scratch, uint32_t* nsmid) {
    nsmid_kernel<<<1, 1>>>(d_scratch);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaMemcpy(nsmid, d_scratch, sizeof(uint32_t), cudaMemcpyDeviceToHost);
}
size_t deflate_quick_tail_bytes() { return 2u * kChunkMax + 4u * kWinPad; }

// grid (CTAs of kParseWarps chains) for `chains_per_sm` chains on each SM
uint32_t deflate_quick_grid(uint32_t nchunks, int num_sms, int chains_per_sm) {
    uint32_t ctas_per_sm = ((uint32_t)chains_per_sm + kParseWarps - 1u) / kParseWarps;
    uint32_t grid = (uint32_t)num_sms * ctas_per_sm;
    uint64_t need = ((uint64_t)nchunks + kParseWarps - 1u) / kParseWarps;
    return need < grid ? (uint32_t)need : grid;
}

cudaError_t launch_quick_parse(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks,
                               uint32_t* tokens, uint32_t tok_stride, uint32_t* ntok, uint32_t* counter,
                               uint16_t* heads, unsigned long long* sm_slots, uint32_t grid, uint8_t* tail, cudaStream_t stream, const StreamSync* sync) {
    if (grid == 0 || nchunks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    if (sync) {            // streamed input: the caller's buffer is padded by kWinPad readable bytes, no private copy of the last chunks
        // This launch owns the SMs for the whole call and needs no shared memory itself.  The L1 / shared-memory split of an SM
        // can only change while the SM is idle, and for a kernel without shared memory it is all-L1: the emit / checksum /
        // gather kernels of the finished slabs (17 KiB of shared memory per CTA) could then not start before this kernel ends.
        // A few KiB per CTA that are never touched make the driver pick a middle split at no cost to the parser (a carve-out
        // preference alone is ignored below 50 %, and from 50 % on the parser loses its L1: 31 -> 23 GB/s); the kernels that
        // must run next to it ask for the same split (launch_static_emit's co_carve).  Measured: profiles/r1_e2e_pipeline.md.
        static int dyn = [] { const char* e = getenv("ZNG_B200_STREAM_DYNSMEM"); return e ? atoi(e) : 4096; }();
        quick_parse_kernel<<<grid, kParseWarps * 32, dyn, stream>>>(in, n, chunk, nchunks, tokens, tok_stride, ntok, counter, heads, sm_slots, in, nchunks, *sync);
        return cudaGetLastError();
    }
    // chunk i may read kWinPad bytes past its end: safe iff (i+1)*chunk + kWinPad <= n
    const uint32_t tail_first = n >= kWinPad ? (uint32_t)((n - kWinPad) / chunk) : 0u;
    const size_t tail_off = (size_t)tail_first * chunk, tail_bytes = n - tail_off;       // <= 2*chunk + kWinPad
    e = cudaMemcpyAsync(tail, in + tail_off, tail_bytes, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(tail + tail_bytes, 0, 2u * kWinPad, stream);
    if (e != cudaSuccess) return e;
    quick_parse_kernel<<<grid, kParseWarps * 32, 0, stream>>>(in, n, chunk, nchunks, tokens, tok_stride, ntok, counter, heads, sm_slots, tail, tail_first, StreamSync{});
    return cudaGetLastError();
}

size_t deflate_primed_head_bytes(uint32_t nsmid) { return (size_t)nsmid * 64u * 65536u * sizeof(uint32_t); }
size_t deflate_primed_tail_bytes() { return kWSize + 2u * kChunkMax + 4u * kWinPad; }

// `in` points at chunk `first` of the stream (first > 0: the 32768 bytes in front of it are readable: its dictionary)
cudaError_t launch_primed_parse(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks, uint32_t first,
                                uint32_t* tokens, uint32_t tok_stride, uint32_t* ntok, uint32_t* counter,
                                uint32_t* heads, unsigned long long* sm_slots, uint32_t grid, uint8_t* tail, cudaStream_t stream) {
    if (grid == 0 || nchunks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    const uint32_t tail_first = n >= kWinPad ? (uint32_t)((n - kWinPad) / chunk) : 0u;
    const size_t tail_off = (size_t)tail_first * chunk, tail_bytes = n - tail_off;
    const uint32_t tail_dict = (first + tail_first) > 0u ? kWSize : 0u;
    e = cudaMemcpyAsync(tail, in + tail_off - tail_dict, tail_dict + tail_bytes, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(tail + tail_dict + tail_bytes, 0, 2u * kWinPad, stream);
    if (e != cudaSuccess) return e;
    primed_parse_kernel<<<grid, kParseWarps * 32, 0, stream>>>(in, n, chunk, nchunks, first, tokens, tok_stride, ntok, counter, heads, sm_slots,
                                                              tail, tail_first, tail_dict);
    return cudaGetLastError();
}

cudaError_t launch_static_emit(const uint32_t* tokens, uint32_t tok_stride, const uint32_t* ntok, size_t n, uint32_t chunk,
                               uint32_t nchunks, int last, uint8_t* out, size_t out_stride, uint32_t* sizes,
                               int num_sms, cudaStream_t stream, int co_carve) {
    if (nchunks == 0) return cudaSuccess;
    uint32_t grid = (uint32_t)num_sms * 12u;
    uint32_t need = (nchunks + kEmitWarps - 1u) / kEmitWarps;
    if (grid > need) grid = need;
    // co_carve >= 0: this launch has to start NEXT TO a running (streamed) parse kernel.  Two kernels only share an SM when
    // they ask for the same L1 / shared-memory split; left alone the driver would give this kernel the all-shared split.
    cudaFuncSetAttribute(static_emit_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co_carve);
    static_emit_kernel<<<grid, kEmitWarps * 32, 0, stream>>>(tokens, tok_stride, ntok, n, chunk, nchunks, last, out, out_stride, sizes);
    return cudaGetLastError();
}

}  // namespace zb
