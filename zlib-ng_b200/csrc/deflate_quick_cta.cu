// deflate_quick_cta.cu -- K1a v6: level-1 (deflate_quick) parse, one CTA per chunk chain, producers and walker pipelined.
//
// Reference semantics (files under /root/reference), reproduced bit-exactly -- same rules as quick_parse_warp:
//   deflate_quick.c:47-130        greedy single-probe parse
//   insert_string_tpl.h:58-75     quick_insert_string: head[] is updated at VISITED positions only
//   arch/generic/compare256_c.c   first-mismatch compare
//
// Why.  Round 1's warp-per-chain parser pays, on the dependent chain of every 32-position step, two or three trips to L2 /
// DRAM (head lookup, candidate bytes) and one __match_any_sync, and hides that with 24 chains per SM whose 128 KiB head
// tables + 64 KiB chunks (680 MB) thrash the 126 MB L2 (ncu: 99 GB read + 17 GB written per GiB of input).  Here the work
// of one chain is split over the warps of a CTA and SOFTWARE-PIPELINED, so that the serial part never waits for memory:
//   * PRODUCER warps (PW of them, one aligned block of 32 positions each, window = 32 x PW positions) work one window
//     AHEAD of the walker: hash, head lookup in the table as it stands (global memory, L2 resident: few chains per SM),
//     candidate bytes, match length up to 64 bytes, and the distance d to the nearest earlier position of the block with the
//     same hash (one __match_any_sync).  They leave one 8-byte record per position in a double-buffered shared ring.
//   * The WALKER warp replays the greedy decisions of deflate_quick over the records, one aligned block per step, everything
//     from shared memory: ballot / ffs, and one shuffle per visited match to follow the orbit of the greedy jump function.
//   * Staleness.  A producer lookup for window w+1 runs while the walker inserts window w, so it may miss inserts of windows
//     w and w+1.  Every insert therefore also sets one bit (one bit per 16-bit hash value) in a bitmap that holds the inserts
//     of the previous and the current window, and leaves its position in a small direct-mapped table last[h & 2047].  A lane
//     whose hash bit is set takes its candidate from last[] (verified against the ring of recent hashes: the newest insert of that
//     hash) and measures it itself; if the slot was taken over by another hash the lane asks the table in global memory --
//     exact, because the walker's own inserts are ordered before it (rare: one L2 round trip).  A lane whose bit is clear had no
//     insert of its hash since the lookup, so the producer's candidate is the reference's.  A visited lane whose hash also
//     belongs to an earlier VISITED lane of its own step (seen from d) cuts the step there; the next step starts at that lane.
//   * Two CTA barriers per window: (A) walk w and production of w+1 are done -> every thread wipes the bitmap generation
//     that window w+1 will insert into (it still holds window w-1) -> (B) -> walk w+1 || production of w+2.
#include "common.cuh"
#include "kernels.h"
#include "lz_ops.cuh"

namespace zb {

// The chunk as the parser sees it: a word-aligned base plus a byte skew.  Loads are ordinary cached loads (ld.global.ca),
// not the non-coherent path: with streamed input the buffer is written by the copy engine during the kernel's lifetime
// (always before the chunk is released to the parser).
struct WindowCA {
    const uint32_t* w;
    uint32_t skew;
    __device__ __forceinline__ uint32_t word(uint32_t i) const {
        uint32_t r;
        asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(r) : "l"(w + i));
        return r;
    }
};

#ifndef ZB_PROD_CAP
#define ZB_PROD_CAP 20
#endif
constexpr uint32_t kProdCap = ZB_PROD_CAP; // producers measure a match up to this many bytes ("cap or more" beyond: the walk measures the rest)
constexpr uint32_t kWalkCap = 12u;        // the walker's own measurement of a recent candidate ("12 or more" beyond)

__device__ __forceinline__ void ca_load12(const WindowCA& W, uint32_t q, uint32_t& v, uint64_t& x) {
    const uint32_t qb = q + W.skew, i = qb >> 2, sh = (qb & 3u) << 3;
    const uint32_t a0 = W.word(i), a1 = W.word(i + 1), a2 = W.word(i + 2), a3 = W.word(i + 3);
    v = __funnelshift_r(a0, a1, sh);
    x = (uint64_t)__funnelshift_r(a1, a2, sh) | ((uint64_t)__funnelshift_r(a2, a3, sh) << 32);
}

__device__ __forceinline__ uint64_t ca_load8(const WindowCA& W, uint32_t pos) {
    const uint32_t qb = pos + W.skew, i = qb >> 2, sh = (qb & 3u) << 3;
    const uint32_t a0 = W.word(i), a1 = W.word(i + 1), a2 = W.word(i + 2);
    return (uint64_t)__funnelshift_r(a0, a1, sh) | ((uint64_t)__funnelshift_r(a1, a2, sh) << 32);
}

// deflate_quick.c:90-103 for one (position, candidate) pair, measured up to `cap` bytes (12 + 8k): 0 = no match (fewer than 4
// equal bytes), 4 .. cap-1 exact and already clipped to the lookahead, cap = "cap or more".  v / x: bytes 0..3 and 4..11 at q.
__device__ __forceinline__ uint32_t match_upto(const WindowCA& W, uint32_t q, uint32_t cand, uint32_t v, uint64_t x, uint32_t n, uint32_t cap) {
    const uint32_t cb = cand + W.skew, i = cb >> 2, sh = (cb & 3u) << 3;
    const uint32_t b0 = W.word(i), b1 = W.word(i + 1), b2 = W.word(i + 2), b3 = W.word(i + 3);
    if (__funnelshift_r(b0, b1, sh) != v) return 0u;
    const uint64_t y = (uint64_t)__funnelshift_r(b1, b2, sh) | ((uint64_t)__funnelshift_r(b2, b3, sh) << 32);
    uint64_t d = x ^ y;
    uint32_t len = 4u;
    while (d == 0ull) {
        len += 8u;
        if (len >= cap) return cap;
        d = ca_load8(W, q + len) ^ ca_load8(W, cand + len);
    }
    len += (unsigned)(__ffsll((long long)d) - 1) >> 3;
    if (len >= cap) return cap;
    return min(len, n - q);                                   // deflate_quick.c:100-101 clip to the lookahead
}

// mask of lanes >= s (s in 0..32 and beyond: clamped)
__device__ __forceinline__ uint32_t lanes_ge(uint32_t s) { return __funnelshift_lc(0u, 0xffffffffu, s); }

// ring record: x = hash << 16 | table candidate;  y = byte | slen << 8 (7 bits) | act << 15 | d << 16 (5 bits)
// Shared memory of one chain, ~16.5 KiB (so that 10+ chains fit one SM): rings are circular in the CHUNK POSITION.
constexpr uint32_t kRing = 256u;          // records: windows w (walked) and w+1 (produced) -> needs 2 x window <= 256
constexpr uint32_t kHsRing = 512u;        // hashes: windows w-2 .. w+1 -> needs 4 x window <= 512
constexpr uint32_t kLast = 2048u;
template <int PW, int BPW>
struct __align__(16) PipeSmem {
    static constexpr int kWin = PW * BPW * 32;
    static_assert(PW * BPW * 32 * 2 <= (int)kRing, "window too large for the record ring");
    uint32_t bits[2048];         // one bit per hash value: inserted by the walk of the previous or the current window
    uint16_t last[kLast];        // [hash & 2047] = position of the last insert that mapped here (verified against hs[])
    uint2    ring[kRing];        // record of position q in ring[q & 255]
    uint32_t hs[kHsRing];        // hash of position q in hs[q & 511] (inactive positions: values no hash can take)
    uint32_t vis[kHsRing / 32];  // visited mask of the block of position q in vis[(q >> 5) & 15]
    uint32_t ci;                 // chunk index of this round
    uint32_t slot;
};

struct WalkStats { uint32_t steps, hit_steps, bad_steps, cuts, longs; unsigned long long walk_clk; };

// One step of the walker: the aligned block `blk` of window w from lane s0 on.  Returns the chunk position where the next step
// starts; wr (token count) is advanced.
template <int PW, int BPW>
__device__ __forceinline__ uint32_t walk_step(const WindowCA& W, PipeSmem<PW, BPW>& sm, uint32_t w, uint32_t blk, uint32_t s0, uint32_t n,
                                              uint16_t* head, uint32_t* __restrict__ tok, uint32_t& wr, unsigned lane, WalkStats& ws) {
    constexpr uint32_t kWin = PW * BPW * 32;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t b0 = w * kWin + blk * 32u;                 // chunk position of lane 0
    ws.steps++;
    const uint32_t q = b0 + lane;
    const uint32_t nl = min(32u, n - b0);                     // lanes that hold a byte
    const uint2 rec = sm.ring[q & (kRing - 1u)];
    const uint32_t h = rec.x >> 16;
    uint32_t cand = rec.x & 0xffffu;
    uint32_t slen = (rec.y >> 8) & 127u;
    uint32_t mcap = kProdCap;                                 // slen == mcap: "mcap or more", the walk measures the rest
    const bool act = (rec.y >> 15) & 1u;
    const uint32_t d = (rec.y >> 16) & 31u;
    const bool live = act && lane >= s0;
    // ---- has the walk inserted my hash since the producers looked it up?  Then that position is the head entry now.
    const uint32_t bw = sm.bits[h >> 5];
    const bool hit = live && ((bw >> (h & 31u)) & 1u);
    if (__any_sync(ZB_FULL, hit)) {
        ws.hit_steps++;
        bool bad = false;
        if (hit) {
            const uint32_t j = sm.last[h & (kLast - 1u)];
            // written by a visited position with my hash; a later insert of my hash would have replaced it
            if (j < q && q - j <= 2u * kWin && sm.hs[j & (kHsRing - 1u)] == h) cand = j;
            else bad = true;                                  // another hash took the slot since
        }
        if (__any_sync(ZB_FULL, bad)) {
            ws.bad_steps++;
            if (bad) {                                        // the table itself: every earlier insert of this walker is ordered before
                cand = (uint32_t)__ldcg(head + h);
                if (q == kWSize + kMaxDist && n < kChunkMax && cand < kWSize) cand = kWSize;   // see quick_parse_warp (slide at 65274)
            }
        }
        if (hit) {
            uint32_t v; uint64_t x;
            ca_load12(W, q, v, x);
            slen = ((q - cand - 1u) < kMaxDist) ? match_upto(W, q, cand, v, x, n, kWalkCap) : 0u;
            mcap = kWalkCap;
        }
    }
    // ---- same hash at an earlier lane of THIS block?  j1 = nearest one; has2: j1 has one as well
    // (only lanes from s0 on belong to this step: the first lane of a step never waits for anyone, so every step advances)
    const bool has1 = live && d != 0u && d + s0 <= lane;
    const uint32_t j1 = (lane - d) & 31u;
    const uint32_t dj = __shfl_sync(ZB_FULL, d, j1);
    const bool has2 = has1 && dj != 0u && dj + s0 <= j1;
    const unsigned M = __ballot_sync(ZB_FULL, slen != 0u && live);
    const unsigned L = __ballot_sync(ZB_FULL, slen == mcap && live);
    // ---- walk 1: the orbit of the greedy jump function.  pack = end << 8 | next match lane at or after end (32: none)
    uint32_t pack;
    {
        const uint32_t end = lane + slen;
        const uint32_t r = M & lanes_ge(end);
        pack = (end << 8) | (r ? (uint32_t)(__ffs(r) - 1) : 32u);
    }
    unsigned K = 0;
    uint32_t k = M ? (uint32_t)(__ffs(M) - 1) : 32u, lastend = 0;
    while (k < 32u) {
        if ((L >> k) & 1u) {
            ws.longs++;
            // a long match is measured by the whole warp when the walk reaches it
            const uint32_t ck = __shfl_sync(ZB_FULL, cand, k);
            const uint32_t mk = __shfl_sync(ZB_FULL, mcap, k);
            const uint32_t qk = b0 + k;
            uint32_t len = mk + warp_compare_bytes(W, qk + mk + W.skew, ck + mk + W.skew, kMaxMatch - mk, lane);
            len = min(len, n - qk);
            len = min(len, kMaxMatch);                        // deflate_quick.c:102-103
            if (lane == k) {
                slen = len;
                const uint32_t end = lane + len;
                const uint32_t r = M & lanes_ge(end);
                pack = (end << 8) | (r ? (uint32_t)(__ffs(r) - 1) : 32u);
            }
        }
        K |= 1u << k;
        const uint32_t pk = __shfl_sync(ZB_FULL, pack, k);
        lastend = pk >> 8;
        k = pk & 63u;
    }
    uint32_t c = max(lastend, nl);
    // covered: behind the nearest taken match at or below me
    bool covered = false;
    {
        const unsigned m = K & (lt | (1u << lane));
        const uint32_t kk = m ? (31u - (uint32_t)__clz(m)) : lane;
        const uint32_t e = __shfl_sync(ZB_FULL, pack, kk) >> 8;
        covered = m != 0u && kk < lane && lane < e;
    }
    unsigned V = __ballot_sync(ZB_FULL, lane >= s0 && lane < nl && !covered);
    // ---- walk 2: the first visited lane that shares its hash with an earlier visited lane of this step cuts the step
    // (has2: the nearest one is not visited but has a predecessor of its own in the block -- cut as well, the next step decides)
    const unsigned S = __ballot_sync(ZB_FULL, ((V >> lane) & 1u) && has1 && (((V >> j1) & 1u) || has2));
    if (S) {
        const unsigned j = __ffs(S) - 1u;
        V &= ~lanes_ge(j);
        c = j;
        ws.cuts++;
    }
    const bool vis = (V >> lane) & 1u;
    uint32_t mytok = rec.y & 0xffu;
    if (vis && slen) mytok = kTokMatch | (slen << 16) | (q - cand);
    if (vis && act) {
        __stcg(head + h, (uint16_t)q);                        // insert_string_tpl.h:70-73
        atomicOr(&sm.bits[h >> 5], 1u << (h & 31u));          // ... and what later lookups of this and the next window must know
        sm.last[h & (kLast - 1u)] = (uint16_t)q;
    }
    if (lane == 0) sm.vis[(b0 >> 5) & (kHsRing / 32u - 1u)] |= V;             // for the re-seeding of the bitmap at the window's end
    if (vis) __stcs(tok + wr + __popc(V & lt), mytok);
    wr += __popc(V);
    __syncwarp();
    return b0 + c;
}

// One producer warp: the aligned block `blk` of window w, looked up in the table as it stands.
template <int PW, int BPW>
__device__ __forceinline__ void produce_block(const WindowCA& W, PipeSmem<PW, BPW>& sm, uint32_t w, uint32_t blk, uint32_t n,
                                              const uint16_t* head, unsigned lane) {
    constexpr uint32_t kWin = PW * BPW * 32;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t q = w * kWin + blk * 32u + lane;
    const bool inw = q < n;
    const bool act = q + kWantMin <= n;                       // deflate_quick.c:88 lookahead >= WANT_MIN_MATCH
    uint32_t v = 0; uint64_t x = 0;
    if (inw) ca_load12(W, q, v, x);
    const uint32_t h = hash4(v);
    const uint32_t myh = act ? h : (0x10000u + lane);
    sm.hs[q & (kHsRing - 1u)] = myh;
    if (lane == 0) sm.vis[(q >> 5) & (kHsRing / 32u - 1u)] = 0u;
    uint32_t cand = 0u;
    if (act) cand = (uint32_t)__ldcg(head + h);
    if (lane < 2u) {                                          // the bytes two windows ahead, on their way into L2 / L1
        const uint32_t pf = (w + 2u) * kWin + blk * 32u + 128u * lane;
        if ((blk & 7u) == 0u && pf + 128u <= n) asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const uint8_t*>(W.w) + W.skew + pf));
    }
    const unsigned peers = __match_any_sync(ZB_FULL, myh);
    const unsigned low = peers & lt;                          // earlier lanes of this block with my hash
    const uint32_t d = low ? lane - (31u - (uint32_t)__clz(low)) : 0u;
    uint32_t slen = 0u;
    if (act) {
        // the slide of a 65275..65535-byte chunk at strstart 65274 (deflate.c:1285-1299), see quick_parse_warp
        if (q == kWSize + kMaxDist && n < kChunkMax && cand < kWSize) cand = kWSize;
        if ((q - cand - 1u) < kMaxDist) slen = match_upto(W, q, cand, v, x, n, kProdCap);
    }
    sm.ring[q & (kRing - 1u)] = make_uint2((h << 16) | cand, (v & 0xffu) | (slen << 8) | ((uint32_t)act << 15) | (d << 16));
}

template <int PW, int BPW>
__global__ void __launch_bounds__((PW + 1) * 32, 1280 / ((PW + 1) * 32))
quick_parse_cta_kernel(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks,
                       uint32_t* __restrict__ tokens, uint32_t tok_stride, uint32_t* __restrict__ ntok,
                       uint32_t* __restrict__ counter, uint16_t* heads, unsigned long long* sm_slots,
                       const uint8_t* tail, uint32_t tail_first, StreamSync sy, unsigned long long* stats) {
    constexpr int kThreads = (PW + 1) * 32;
    constexpr uint32_t kWin = PW * BPW * 32;
    __shared__ PipeSmem<PW, BPW> sm;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const bool walker = warp == PW;                             // the last warp walks, warps 0..PW-1 produce block `warp` of each window
    if (tid == 0) sm.slot = slot_acquire(sm_slots + smid());
    __syncthreads();
    const uint32_t slot = sm.slot;
    uint16_t* head = heads + ((size_t)smid() * 64u + slot) * 65536u;
    for (;;) {
        if (tid == 0) {
            const uint32_t ci = atomicAdd(counter, 1u);
            sm.ci = ci;
            if (sy.ready && ci < nchunks) {                     // streamed input: chunk ci is parsed once chunk ci + 1 has been delivered
                const long long t0 = clock64();
                while (*(volatile const uint32_t*)sy.ready <= ci) {
                    __nanosleep(1000);
                    if (clock64() - t0 > sy.patience) { atomicExch(sy.failed, 1u); break; }
                }
                __threadfence();
            }
        }
        __syncthreads();
        const uint32_t ci = sm.ci;
        if (ci >= nchunks) break;
        {   // CLEAR_HASH (deflate.c:182-184), and both bitmap generations
            uint4* h4 = reinterpret_cast<uint4*>(head);
#pragma unroll 4
            for (uint32_t k = tid; k < 65536u * 2u / 16u; k += kThreads) __stcg(h4 + k, make_uint4(0, 0, 0, 0));
            uint4* g4 = reinterpret_cast<uint4*>(&sm.bits[0]);
            for (uint32_t k = tid; k < 2048u / 4u; k += kThreads) g4[k] = make_uint4(0, 0, 0, 0);
        }
        const size_t off = (size_t)ci * chunk;
        const uint32_t len = (uint32_t)min((size_t)chunk, n - off);
        const uint8_t* src = (ci >= tail_first) ? tail + (size_t)(ci - tail_first) * chunk : in + off;
        WindowCA W;
        W.skew = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
        W.w = reinterpret_cast<const uint32_t*>(src - W.skew);
        uint32_t* tok = tokens + (size_t)ci * tok_stride;
        const uint32_t nwin = (len + kWin - 1u) / kWin;
        uint32_t cur = 0, wr = 0;                               // walker: next position to parse, tokens written
        WalkStats ws{};                                         // debug counters (ZNG_B200_K1_STATS=1), lane-uniform
        const long long t_chunk = stats ? clock64() : 0;
        __syncthreads();                                        // the cleared table is what window 0's lookups see
        if (!walker && nwin) for (uint32_t b = warp; b < (uint32_t)(PW * BPW); b += PW) produce_block<PW, BPW>(W, sm, 0u, b, len, head, lane);
        __syncthreads();
        for (uint32_t w = 0; w < nwin; w++) {
            const long long t0 = stats ? clock64() : 0;
            if (walker) {
                const uint32_t wend = min(len, (w + 1u) * kWin);
                while (cur < wend) {
                    const uint32_t rel = cur - w * kWin;
                    cur = walk_step<PW, BPW>(W, sm, w, rel >> 5, rel & 31u, len, head, tok, wr, lane, ws);
                }
                if (stats) ws.walk_clk += (unsigned long long)(clock64() - t0);
            } else if (w + 1u < nwin) {
                for (uint32_t b = warp; b < (uint32_t)(PW * BPW); b += PW) produce_block<PW, BPW>(W, sm, w + 1u, b, len, head, lane);
                if (stats && warp == 0) ws.walk_clk += (unsigned long long)(clock64() - t0);      // producer warp 0: its production time
            }
            __syncthreads();                                    // (A) walk w and window w+1's records are complete
            if (w + 1u < nwin) {
                // window w+1 was looked up while window w was walked: its walk must know the inserts of windows w and w+1.  The
                // bitmap still holds window w-1 as well: wipe it, then set the bits of window w's visited positions again.
                uint4* g4 = reinterpret_cast<uint4*>(&sm.bits[0]);
                for (uint32_t k = tid; k < 2048u / 4u; k += kThreads) g4[k] = make_uint4(0, 0, 0, 0);
                __syncthreads();
                for (uint32_t t = tid; t < kWin; t += kThreads) {
                    const uint32_t q = w * kWin + t;
                    const uint32_t hh = sm.hs[q & (kHsRing - 1u)];
                    if (hh < 0x10000u && ((sm.vis[(q >> 5) & (kHsRing / 32u - 1u)] >> (q & 31u)) & 1u)) atomicOr(&sm.bits[hh >> 5], 1u << (hh & 31u));
                }
            }
            __syncthreads();                                    // (B)
        }
        if (stats && lane == 0 && (walker || warp == 0)) {
            if (walker) {
                atomicAdd(stats + 0, 1ull); atomicAdd(stats + 1, (unsigned long long)ws.steps); atomicAdd(stats + 2, ws.walk_clk);
                atomicAdd(stats + 3, (unsigned long long)(clock64() - t_chunk)); atomicAdd(stats + 4, (unsigned long long)ws.hit_steps);
                atomicAdd(stats + 5, (unsigned long long)ws.bad_steps); atomicAdd(stats + 6, (unsigned long long)ws.cuts);
                atomicAdd(stats + 7, (unsigned long long)ws.longs); atomicAdd(stats + 9, (unsigned long long)nwin);
            } else atomicAdd(stats + 8, ws.walk_clk);
        }
        if (walker) {
            if (lane == 0) { __stcs(tok + wr, kTokEnd); ntok[ci] = wr; }
            if (sy.done) {                                      // per output slab: how many of its chunks are parsed
                __threadfence();
                __syncwarp();
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
    }
    __syncthreads();
    if (tid == 0) atomicAnd(sm_slots + smid(), ~(1ull << slot));
}

template <int PW, int BPW>
static cudaError_t launch_cta(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks, uint32_t* tokens, uint32_t tok_stride,
                              uint32_t* ntok, uint32_t* counter, uint16_t* heads, unsigned long long* sm_slots, uint32_t grid,
                              const uint8_t* tail, uint32_t tail_first, cudaStream_t stream, const StreamSync& sy, int chains_per_sm,
                              unsigned long long* stats) {
    // shared memory per chain: ~39 KiB at PW = 8; ask for the split that just holds the chains of one SM (the rest stays L1)
    int carve = (int)((sizeof(PipeSmem<PW, BPW>) + 1024u) * (size_t)chains_per_sm * 100u / (228u * 1024u)) + 1;
    if (carve > 100) carve = 100;
    cudaFuncSetAttribute(quick_parse_cta_kernel<PW, BPW>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    quick_parse_cta_kernel<PW, BPW><<<grid, (PW + 1) * 32, 0, stream>>>(in, n, chunk, nchunks, tokens, tok_stride, ntok, counter, heads, sm_slots,
                                                                   tail, tail_first, sy, stats);
    return cudaGetLastError();
}

cudaError_t launch_quick_parse_cta(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks,
                                   uint32_t* tokens, uint32_t tok_stride, uint32_t* ntok, uint32_t* counter,
                                   uint16_t* heads, unsigned long long* sm_slots, int num_sms, int chains_per_sm, int warps, uint8_t* tail,
                                   cudaStream_t stream, const StreamSync* sync, unsigned long long* stats) {
    if (nchunks == 0) return cudaSuccess;
    uint32_t grid = (uint32_t)num_sms * (uint32_t)chains_per_sm;
    if (grid > nchunks) grid = nchunks;
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    const uint8_t* tl = in; uint32_t tail_first = nchunks; StreamSync sy{};
    if (sync) {            // streamed input: the caller's buffer is padded by kWinPad readable bytes, no private copy of the last chunks
        sy = *sync;
    } else {
        // chunk i may read kWinPad bytes past its end: safe iff (i+1)*chunk + kWinPad <= n
        tail_first = n >= kWinPad ? (uint32_t)((n - kWinPad) / chunk) : 0u;
        const size_t tail_off = (size_t)tail_first * chunk, tail_bytes = n - tail_off;       // <= 2*chunk + kWinPad
        e = cudaMemcpyAsync(tail, in + tail_off, tail_bytes, cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(tail + tail_bytes, 0, 2u * kWinPad, stream);
        if (e != cudaSuccess) return e;
        tl = tail;
    }
    // warps = 10 * (producer warps per chain) + (blocks per producer warp and window); window = 32 x their product <= 128 positions
#define ZB_CTA(P, B) case 10 * P + B: return launch_cta<P, B>(in, n, chunk, nchunks, tokens, tok_stride, ntok, counter, heads, sm_slots, grid, tl, tail_first, stream, sy, chains_per_sm, stats)
    switch (warps) {
        ZB_CTA(1, 2); ZB_CTA(1, 4); ZB_CTA(2, 1); ZB_CTA(2, 2); ZB_CTA(3, 1); ZB_CTA(4, 1);
        default: return cudaErrorInvalidValue;
    }
#undef ZB_CTA
}

}  // namespace zb
