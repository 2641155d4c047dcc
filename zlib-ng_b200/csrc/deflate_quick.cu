// deflate_quick.cu -- K1: level-1 (deflate_quick) compression of independent <= 64 KiB chunks.
//
// Reference semantics reproduced bit-exactly (files under /root/reference):
//   deflate_quick.c:47-130        greedy single-probe parse, static-Huffman emission
//   insert_string_tpl.h:58-75     quick_insert_string (head[] update only at visited positions)
//   arch/generic/compare256_c.c   first-mismatch compare (here: 32 lanes x 8 bytes + ballot/ffs)
//   trees_emit.h:102-225          zng_tr_emit_lit / zng_tr_emit_dist / emit_tree / emit_end_block
//   deflate.c:1061-1083           empty stored block appended for Z_SYNC_FLUSH / Z_FULL_FLUSH
//
// B200 mapping (v2, "many chains per SM").  The parse of one chunk is a serial dependency chain
// (which positions get hashed depends on every earlier match decision), and ncu on v1 (one chunk
// per SM, window + head table in shared memory) showed the SM 93% idle: one warp advances one
// chain at ~0.1 IPC.  So the unit of parallelism is the CHAIN, and the SM runs several:
//   * a CTA is one chain: warp 0 PARSER, warp 1 EMITTER, 12 KiB of shared memory (token ring +
//     bit staging ring); the grid is (chains per SM) x 148 persistent CTAs pulling chunk indices
//     from an atomic counter;
//   * the 65536-entry u16 hash-head table of deflate_state (deflate.h:232, 128 KiB) lives in a
//     per-CTA slab of global memory that stays L2-resident (4 x 148 x 128 KiB = 74 MiB of the
//     126 MB L2), accessed with ld/st.global.cg; it is cleared per chunk (CLEAR_HASH, deflate.c:182);
//   * the window is the input itself, read through L1 (ld.global.nc): no staging copy at all.
// PARSER.  The warp speculates: 32 lanes hash 32 consecutive positions against the table state at
//   the window start, test their candidates and measure short matches (< 12 bytes) in parallel; a
//   ballot/ffs walk then replays the reference's decisions (literal runs are accepted wholesale, a
//   match costs one shuffle, or one warp-wide compare when it is 12 bytes or longer).  A lane whose
//   hash equals that of an earlier VISITED lane of the same window saw a stale candidate; the window
//   is cut at the first such lane and restarted there, which keeps the result exact.
// EMITTER.  Per 32 tokens: fixed-code bits, warp prefix scan of the bit lengths, OR into the
//   staging ring, 2 KiB segments stored to global memory as uint4.
// Per-chunk CRC-32 / Adler-32 come from the K3 tile kernel (checksum.cu) launched on the same
// stream; the chunk it re-reads is L2-resident.
#include "common.cuh"
#include "kernels.h"

namespace zb {

constexpr int      kQThreads     = 64;
constexpr uint32_t kTokRing      = 1024;          // tokens (u32)
constexpr uint32_t kStageWords   = 2048;          // u32 words (8 KiB)
constexpr uint32_t kStageSeg     = 512;           // words per flush segment (2 KiB)
constexpr uint32_t kSpinLimit    = 1u << 24;      // ring-wait watchdog: trap instead of hanging

struct QuickSmem {
    uint32_t tok[kTokRing];
    uint32_t stage[kStageWords];
    volatile uint32_t tok_wr;      // tokens produced (monotonic)
    volatile uint32_t tok_rd;      // tokens consumed (monotonic)
    uint32_t chunk_idx;
};

// The chunk as the parser sees it: a word-aligned base in global memory plus a byte skew.
struct Window {
    const uint32_t* w;             // 4-byte aligned
    uint32_t skew;                 // 0..3: chunk byte 0 is byte `skew` of w[0]
    __device__ __forceinline__ uint32_t word(uint32_t i) const { return __ldg(w + i); }
};

// ---------------------------------------------------------------- parser (warp 0)
// Warp-wide first-mismatch over 256 bytes: lane l compares the 8 bytes at a+8l / b+8l (byte
// offsets already include the skew).  Returns the number of equal leading bytes (0..256).
// (compare256 analogue, compare256_c.c:12-43)
__device__ __forceinline__ uint32_t warp_compare256(const Window& W, uint32_t a, uint32_t b, unsigned lane) {
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

__device__ __forceinline__ void tok_push(QuickSmem& s, bool mine, uint32_t rank, uint32_t tok, uint32_t cnt,
                                         uint32_t& wr, uint32_t& rd_seen, unsigned lane, uint32_t* dbg) {
    // wait for ring space (the emitter advances tok_rd); re-poll only when the cached value says "full"
    if (wr + cnt - rd_seen > kTokRing) {
        for (uint32_t spins = 0;; spins++) {
            uint32_t rd = 0;
            if (lane == 0) rd = s.tok_rd;
            rd_seen = __shfl_sync(ZB_FULL, rd, 0);
            if (wr + cnt - rd_seen <= kTokRing) break;
            if (spins > kSpinLimit) __trap();               // a stuck ring is a bug: fail, never hang the GPU
            __nanosleep(64);
        }
    }
    if (mine) {
        s.tok[(wr + rank) & (kTokRing - 1u)] = tok;
        if (dbg) dbg[wr + rank] = tok;
    }
    wr += cnt;
    __threadfence_block();
    __syncwarp();
    if (lane == 0) s.tok_wr = wr;
}

__device__ void quick_parse_warp(QuickSmem& s, const Window W, uint32_t n, uint16_t* __restrict__ head, uint32_t* dbg) {
    const unsigned lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    uint32_t wr = 0;
    if (lane == 0) wr = s.tok_wr;
    wr = __shfl_sync(ZB_FULL, wr, 0);
    uint32_t rd_seen = wr;                                    // the emitter had drained the ring at the chunk boundary
    if (dbg) dbg -= wr;                                       // dbg[wr + rank] indexes from 0 for this chunk
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
        const uint32_t cand = act ? (uint32_t)__ldcg(head + h) : 0u;
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
                len = 12u + warp_compare256(W, qk + 12u + W.skew, ck + 12u + W.skew, lane);
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
        if (vis && act) __stcg(head + h, (uint16_t)q);       // insert_string_tpl.h:70-73 (visited positions only)
        tok_push(s, vis && inb, __popc(V & lt), mytok, __popc(V), wr, rd_seen, lane, dbg);
        p += cur;
    }
    // end marker
    tok_push(s, lane == 0, 0, kTokEnd, 1, wr, rd_seen, lane, dbg);
}

// ---------------------------------------------------------------- emitter (warp 1)
__device__ __forceinline__ void stage_flush(QuickSmem& s, uint8_t* out, uint32_t& flushed, uint32_t upto_words, unsigned lane) {
    // store [flushed, upto_words) words of the staging ring to global memory and re-zero them
    __syncwarp();
    while (flushed < upto_words) {
        uint32_t cnt = min(upto_words - flushed, kStageSeg);
        uint32_t base = flushed & (kStageWords - 1u);
        if ((cnt & 3u) == 0u && (flushed & 3u) == 0u) {
            uint4* g = reinterpret_cast<uint4*>(out + (size_t)flushed * 4u);
            uint4* sm = reinterpret_cast<uint4*>(&s.stage[base]);
            for (uint32_t i = lane; i < cnt / 4u; i += 32u) { __stcs(g + i, sm[i]); sm[i] = make_uint4(0, 0, 0, 0); }
        } else {
            uint32_t* g = reinterpret_cast<uint32_t*>(out + (size_t)flushed * 4u);
            for (uint32_t i = lane; i < cnt; i += 32u) { g[i] = s.stage[base + i]; s.stage[base + i] = 0u; }
        }
        flushed += cnt;
    }
    __syncwarp();
}

__device__ __forceinline__ void stage_or(QuickSmem& s, uint32_t bitpos, uint32_t bits, uint32_t nbits) {
    if (nbits == 0u) return;
    uint32_t w = bitpos >> 5, sh = bitpos & 31u;
    uint64_t v = (uint64_t)bits << sh;
    atomicOr(&s.stage[w & (kStageWords - 1u)], (uint32_t)v);
    uint32_t hi = (uint32_t)(v >> 32);
    if (hi) atomicOr(&s.stage[(w + 1u) & (kStageWords - 1u)], hi);
}

// returns the chunk's compressed size in bytes
__device__ uint32_t quick_emit_warp(QuickSmem& s, uint8_t* out, int last, bool open_block) {
    const unsigned lane = lane_id();
    uint32_t rd = 0;
    if (lane == 0) rd = s.tok_rd;
    rd = __shfl_sync(ZB_FULL, rd, 0);
    uint32_t bitpos = 0, flushed = 0;
    if (open_block) {                                       // trees_emit.h:198-207: (STATIC_TREES<<1)+last, 3 bits
        if (lane == 0) stage_or(s, 0, (1u << 1) + (uint32_t)last, 3);
        bitpos = 3;
    }
    bool end = false;
    while (!end) {
        uint32_t avail;
        for (uint32_t spins = 0;; spins++) {
            uint32_t w = 0;
            if (lane == 0) w = s.tok_wr;
            w = __shfl_sync(ZB_FULL, w, 0);
            avail = w - rd;
            if (avail) break;
            if (spins > kSpinLimit) __trap();
            __nanosleep(100);
        }
        __threadfence_block();
        const uint32_t m = min(avail, 32u);
        uint32_t tok = lane < m ? s.tok[(rd + lane) & (kTokRing - 1u)] : kTokEnd;
        uint32_t bits = 0, nb = 0;
        const unsigned E = __ballot_sync(ZB_FULL, lane < m && (tok & kTokEnd));
        uint32_t take = m;
        if (E) { take = __ffs(E) - 1u; end = true; }
        if (lane < take) fixed_code_token(tok, bits, nb);
        uint32_t incl = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(ZB_FULL, incl, d); if ((int)lane >= d) incl += t; }
        stage_or(s, bitpos + incl - nb, bits, nb);
        bitpos += __shfl_sync(ZB_FULL, incl, 31);
        rd += take + (end ? 1u : 0u);
        __syncwarp();
        if (lane == 0) s.tok_rd = rd;
        if ((bitpos >> 5) - flushed >= kStageSeg) stage_flush(s, out, flushed, ((bitpos >> 5) / kStageSeg) * kStageSeg, lane);
    }
    // block end + flush marker
    if (lane == 0) {
        if (open_block) { bitpos += 7; }                    // END_BLOCK under the fixed code: 7 zero bits
        if (last) {
            bitpos = (bitpos + 7u) & ~7u;                   // bi_windup
        } else {                                            // deflate.c:1064-1065 zng_tr_stored_block(NULL,0,0)
            bitpos += 3;                                    // (STORED_BLOCK<<1)+0
            bitpos = (bitpos + 7u) & ~7u;
            stage_or(s, bitpos, 0xffff0000u, 32);           // LEN=0x0000, NLEN=0xffff
            bitpos += 32;
        }
    }
    bitpos = __shfl_sync(ZB_FULL, bitpos, 0);
    stage_flush(s, out, flushed, (bitpos + 31u) >> 5, lane);
    return bitpos >> 3;
}

// ---------------------------------------------------------------- kernel
// heads: gridDim.x slabs of 65536 u16.  tail: a zero-padded private copy of the chunks from
// `tail_first` on -- those whose read-ahead (<= kWinPad bytes past the chunk) could leave the
// caller's allocation; every other chunk reads ahead into its successors, whose bytes cannot
// influence the result (lengths are clipped to the chunk).
__global__ void __launch_bounds__(kQThreads)
deflate_quick_kernel(const uint8_t* __restrict__ in, size_t n, uint32_t chunk, uint32_t nchunks, int last,
                     uint8_t* __restrict__ out, size_t out_stride, uint32_t* __restrict__ sizes,
                     uint32_t* __restrict__ counter, uint16_t* __restrict__ heads, const uint8_t* __restrict__ tail,
                     uint32_t tail_first, uint32_t* __restrict__ dbg_tokens, uint32_t dbg_stride) {
    __shared__ __align__(16) QuickSmem s;
    const unsigned tid = threadIdx.x, warp = tid >> 5;
    uint16_t* head = heads + (size_t)blockIdx.x * 65536u;

    if (tid == 0) { s.tok_wr = 0; s.tok_rd = 0; }
    for (uint32_t i = tid; i < kStageWords; i += kQThreads) s.stage[i] = 0u;

    for (;;) {
        __syncthreads();
        if (tid == 0) s.chunk_idx = atomicAdd(counter, 1u);
        // CLEAR_HASH (deflate.c:182-184): 128 KiB of zeros into the L2-resident slab
        {
            uint4* h4 = reinterpret_cast<uint4*>(head);
            for (uint32_t i = tid; i < 65536u * 2u / 16u; i += kQThreads) __stcg(h4 + i, make_uint4(0, 0, 0, 0));
        }
        __syncthreads();
        const uint32_t ci = s.chunk_idx;
        if (ci >= nchunks) break;
        const size_t off = (size_t)ci * chunk;
        const uint32_t len = (uint32_t)min((size_t)chunk, n - off);

        if (warp == 0) {
            const uint8_t* src = (ci >= tail_first) ? tail + (size_t)(ci - tail_first) * chunk : in + off;
            Window W;
            W.skew = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
            W.w = reinterpret_cast<const uint32_t*>(src - W.skew);
            quick_parse_warp(s, W, len, head, dbg_tokens ? dbg_tokens + (size_t)ci * dbg_stride : nullptr);
        } else {
            const bool open_block = (len > 0) || last;      // deflate_quick.c:53-63
            uint32_t sz = quick_emit_warp(s, out + (size_t)ci * out_stride, last, open_block);
            if ((tid & 31u) == 0) sizes[ci] = sz;
        }
    }
}

size_t deflate_quick_head_bytes(uint32_t grid) { return (size_t)grid * 65536u * sizeof(uint16_t); }
size_t deflate_quick_tail_bytes() { return 2u * kChunkMax + 4u * kWinPad; }

uint32_t deflate_quick_grid(uint32_t nchunks, int num_sms, int chains_per_sm) {
    uint32_t grid = (uint32_t)num_sms * (uint32_t)chains_per_sm;
    return nchunks < grid ? nchunks : grid;
}

cudaError_t launch_deflate_quick(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks, int last,
                                 uint8_t* out, size_t out_stride, uint32_t* sizes, uint32_t* counter,
                                 uint16_t* heads, uint32_t grid, uint8_t* tail, cudaStream_t stream,
                                 uint32_t* dbg_tokens, uint32_t dbg_stride) {
    if (grid == 0 || nchunks == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    // chunk i may read kWinPad bytes past its end: safe iff (i+1)*chunk + kWinPad <= n
    const uint32_t tail_first = n >= kWinPad ? (uint32_t)((n - kWinPad) / chunk) : 0u;
    const size_t tail_off = (size_t)tail_first * chunk, tail_bytes = n - tail_off;       // <= 2*chunk + kWinPad
    e = cudaMemcpyAsync(tail, in + tail_off, tail_bytes, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(tail + tail_bytes, 0, 2u * kWinPad, stream);
    if (e != cudaSuccess) return e;
    deflate_quick_kernel<<<grid, kQThreads, 0, stream>>>(in, n, chunk, nchunks, last, out, out_stride, sizes, counter,
                                                         heads, tail, tail_first, dbg_tokens, dbg_stride);
    return cudaGetLastError();
}

}  // namespace zb
