// deflate_quick_cta.cu -- K1a v5: level-1 (deflate_quick) parse, one CTA per chunk chain.
//
// Reference semantics (files under /root/reference), reproduced bit-exactly -- same rules as quick_parse_warp:
//   deflate_quick.c:47-130        greedy single-probe parse
//   insert_string_tpl.h:58-75     quick_insert_string: head[] is updated at VISITED positions only
//   arch/generic/compare256_c.c   first-mismatch compare
//
// Why a second parser.  Round 1's warp-per-chain parser walks 32 positions per step and pays, on the dependent chain of every
// step, two or three trips to L2 / DRAM (head lookup, candidate bytes) and one __match_any_sync (measured on B200: 240..410
// clocks, profiles/r2_ubench_warp.txt).  It hides that latency with 24 chains per SM, whose 128 KiB head tables + 64 KiB
// chunks (680 MB) thrash the 126 MB L2: every speculative head lookup becomes a DRAM sector read (ncu: 99 GB read + 17 GB
// written per GiB of input).  This parser splits the work of one chain over the warps of a CTA:
//   * PRODUCER phase, all warps, one position per thread, W = 32 x warps positions per window: hash, head lookup in the table
//     as it stands at the window start, candidate bytes, 12-byte match measurement; plus, from a shared array of the
//     window's hashes, the distance d to the nearest earlier position of the window (<= 31 back) with the same hash and the
//     match against THAT position.  One trip to L2 for the lookups and one to L1 / L2 for the candidate bytes per W positions.
//   * WALK phase, one warp, 32 positions per step, everything from shared memory: ballot / ffs replay of the greedy
//     decisions; the orbit of the greedy jump function is followed with one shuffle per visited match (every lane
//     precomputes where a match taken at it leads).
//   * What the walk inserts is not visible to lookups made at the window start.  A small exact cache in shared memory
//     (2W sets x 4 ways, tag = the 16-bit hash + the window's epoch) holds every insert of the current window; a lane that
//     finds its hash there takes that (newer) candidate -- normally position - d, already measured by the producers.  A
//     visited lane whose hash also belongs to an earlier VISITED lane of its own step (seen from d, no match_any) cuts the
//     step there: the next step finds the entry in the cache.  If a cache set overflows the window ends at that point; at a
//     window start the table in global memory is exact.
#include "common.cuh"
#include "kernels.h"

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

__device__ __forceinline__ uint32_t warp_compare256_ca(const WindowCA& W, uint32_t a, uint32_t b, unsigned lane) {
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

// deflate_quick.c:90-103 for one (position, candidate) pair, up to 12 bytes: 0 = no match, 4..11 exact (already clipped to the
// lookahead), 12 = "12 or more".  v / x: bytes 0..3 and 4..11 at q.
__device__ __forceinline__ uint32_t match12(const WindowCA& W, uint32_t q, uint32_t cand, uint32_t v, uint64_t x, uint32_t n) {
    const uint32_t cb = cand + W.skew, i = cb >> 2, sh = (cb & 3u) << 3;
    const uint32_t b0 = W.word(i), b1 = W.word(i + 1), b2 = W.word(i + 2), b3 = W.word(i + 3);
    if (__funnelshift_r(b0, b1, sh) != v) return 0u;
    const uint64_t y = (uint64_t)__funnelshift_r(b1, b2, sh) | ((uint64_t)__funnelshift_r(b2, b3, sh) << 32);
    const uint64_t d = x ^ y;
    uint32_t slen = d ? 4u + ((unsigned)(__ffsll((long long)d) - 1) >> 3) : 12u;
    if (slen < 12u) slen = min(slen, n - q);
    return slen;
}

__device__ __forceinline__ void load12(const WindowCA& W, uint32_t q, uint32_t& v, uint64_t& x) {
    const uint32_t qb = q + W.skew, i = qb >> 2, sh = (qb & 3u) << 3;
    const uint32_t a0 = W.word(i), a1 = W.word(i + 1), a2 = W.word(i + 2), a3 = W.word(i + 3);
    v = __funnelshift_r(a0, a1, sh);
    x = (uint64_t)__funnelshift_r(a1, a2, sh) | ((uint64_t)__funnelshift_r(a2, a3, sh) << 32);
}

// mask of lanes >= s (s in 0..32 and beyond: clamped)
__device__ __forceinline__ uint32_t lanes_ge(uint32_t s) { return __funnelshift_lc(0u, 0xffffffffu, s); }

// ring record: x = hash << 16 | table candidate;  y = byte | slen << 8 | act << 12 | d << 16
template <int WARPS>
struct __align__(16) CtaSmem {
    static constexpr int kWin = WARPS * 32;
    uint32_t bitmap[2048];       // one bit per hash value: inserted by the walk of the current window
    uint8_t  table[4096];        // [hash & 4095] = window index of the last insert that mapped here (verified against hs[])
    uint2    ring[kWin];
    uint32_t hs[kWin + 32];      // the window's hashes behind 32 sentinels (never equal to a hash)
    uint32_t ci;                 // chunk index of this round
    uint32_t adv;                // positions consumed by the walk of the current window
    uint32_t slot;
};

// One 32-position step of the walker warp at window index `cur`.  Returns the positions consumed; wr (token count) is
// advanced.  ovf: the insert cache could not answer exactly -> the window must end here (the step is not executed, 0 is returned).
template <int WARPS>
__device__ __forceinline__ uint32_t walk_step(const WindowCA& W, CtaSmem<WARPS>& sm, uint32_t p, uint32_t cur, uint32_t nwin, uint32_t n,
                                              uint16_t* head, uint32_t* __restrict__ tok, uint32_t& wr, bool& ovf, unsigned lane) {
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t i = cur + lane;
    const bool inw = i < nwin;
    const uint32_t q = p + i;
    uint2 rec = make_uint2(0u, 0u);
    if (inw) rec = sm.ring[i];
    const uint32_t h = rec.x >> 16;
    uint32_t cand = rec.x & 0xffffu;
    uint32_t slen = (rec.y >> 8) & 15u;
    const bool act = (rec.y >> 12) & 1u;
    const uint32_t d = (rec.y >> 16) & 31u;
    // ---- has the walk of this window inserted my hash already?  Then that position is the head entry now.
    const uint32_t bw = sm.bitmap[h >> 5];
    const bool hit = act && ((bw >> (h & 31u)) & 1u);
    if (__any_sync(ZB_FULL, hit)) {
        bool bad = false;
        if (hit) {
            const uint32_t j = sm.table[h & 4095u];
            if (j < i && sm.hs[32u + j] == h) {       // written by a visited lane with my hash; a later insert of it would have replaced it
                uint32_t v; uint64_t x;
                load12(W, q, v, x);
                cand = p + j;
                slen = match12(W, q, cand, v, x, n);
            } else bad = true;                        // another hash took the slot since
        }
        if (__any_sync(ZB_FULL, bad)) { ovf = true; return 0u; }
    }
    // ---- same hash at an earlier lane of THIS step?  j1 = nearest one; has2: j1 has one as well
    const bool has1 = act && d != 0u && d <= lane;
    const uint32_t j1 = (lane - d) & 31u;
    const uint32_t dj = __shfl_sync(ZB_FULL, d, j1);
    const bool has2 = has1 && dj != 0u && dj <= j1;
    const unsigned M = __ballot_sync(ZB_FULL, slen != 0u);
    const unsigned L = __ballot_sync(ZB_FULL, slen == 12u);
    const unsigned nl = min(32u, nwin - cur);
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
            // a long match is measured by the whole warp when the walk reaches it
            const uint32_t ck = __shfl_sync(ZB_FULL, cand, k);
            const uint32_t qk = p + cur + k;
            uint32_t len = 12u + warp_compare256_ca(W, qk + 12u + W.skew, ck + 12u + W.skew, lane);
            len = min(len, n - qk);
            len = min(len, kMaxMatch);                   // deflate_quick.c:102-103
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
    unsigned V = __ballot_sync(ZB_FULL, lane < nl && !covered);
    // ---- walk 2: the first visited lane that shares its hash with an earlier visited lane of this step cuts the step
    // (has2: the nearest one is not visited but has a predecessor of its own in the step -- cut as well, the next step's lookup decides)
    const unsigned S = __ballot_sync(ZB_FULL, ((V >> lane) & 1u) && has1 && (((V >> j1) & 1u) || has2));
    if (S) {
        const unsigned j = __ffs(S) - 1u;
        V &= ~lanes_ge(j);
        c = j;
    }
    const bool vis = (V >> lane) & 1u;
    uint32_t mytok = rec.y & 0xffu;
    if (vis && slen) mytok = kTokMatch | (slen << 16) | (q - cand);
    if (vis && act) {
        __stcg(head + h, (uint16_t)q);                        // insert_string_tpl.h:70-73
        atomicOr(&sm.bitmap[h >> 5], 1u << (h & 31u));        // ... and what later steps of this window must know about it
        sm.table[h & 4095u] = (uint8_t)i;
    }
    if (vis) __stcs(tok + wr + __popc(V & lt), mytok);
    wr += __popc(V);
    __syncwarp();
    return c;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1024 / (WARPS * 32))
quick_parse_cta_kernel(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks,
                       uint32_t* __restrict__ tokens, uint32_t tok_stride, uint32_t* __restrict__ ntok,
                       uint32_t* __restrict__ counter, uint16_t* heads, unsigned long long* sm_slots,
                       const uint8_t* tail, uint32_t tail_first, StreamSync sy) {
    constexpr int kThreads = WARPS * 32;
    constexpr int kWin = kThreads;
    __shared__ CtaSmem<WARPS> sm;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) sm.slot = slot_acquire(sm_slots + smid());
    for (uint32_t k = tid; k < 2048u; k += kThreads) sm.bitmap[k] = 0u;
    if (tid < 32u) sm.hs[tid] = 0x20000u + tid;
    __syncthreads();
    const uint32_t slot = sm.slot;
    uint16_t* head = heads + ((size_t)smid() * 64u + slot) * 65536u;
    const unsigned ww = slot % WARPS;                           // the walker warp: chains of one SM spread over its schedulers
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
        {   // CLEAR_HASH (deflate.c:182-184)
            uint4* h4 = reinterpret_cast<uint4*>(head);
#pragma unroll 8
            for (uint32_t k = tid; k < 65536u * 2u / 16u; k += kThreads) __stcg(h4 + k, make_uint4(0, 0, 0, 0));
        }
        const size_t off = (size_t)ci * chunk;
        const uint32_t len = (uint32_t)min((size_t)chunk, n - off);
        const uint8_t* src = (ci >= tail_first) ? tail + (size_t)(ci - tail_first) * chunk : in + off;
        WindowCA W;
        W.skew = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
        W.w = reinterpret_cast<const uint32_t*>(src - W.skew);
        uint32_t* tok = tokens + (size_t)ci * tok_stride;
        uint32_t p = 0, wr = 0;
        uint32_t prev_h = 0;                                    // my position's hash in the previous window: its bitmap word is wiped
        __syncthreads();
        while (p < len) {
            // ---- producer phase: every thread looks one position up in the table as it stands at the window start
            sm.bitmap[prev_h >> 5] = 0u;                        // every insert of the last window belongs to some thread's position
            const uint32_t q = p + tid;
            const bool inw = q < len;
            const bool act = q + kWantMin <= len;               // deflate_quick.c:88 lookahead >= WANT_MIN_MATCH
            uint32_t v = 0; uint64_t x = 0;
            if (inw) load12(W, q, v, x);
            const uint32_t h = hash4(v);
            const uint32_t myh = act ? h : (0x10000u + tid);
            sm.hs[32u + tid] = myh;
            prev_h = h;
            uint32_t cand = 0u;
            if (act) cand = (uint32_t)__ldcg(head + h);
            if (warp == ((ww + 1u) % WARPS) && lane < 6u && p + kWin + 128u * lane + 640u <= len) {
                // the bytes of the next windows, on their way into L1 while this window is walked
                asm volatile("prefetch.global.L1 [%0];" :: "l"(reinterpret_cast<const uint8_t*>(W.w) + W.skew + p + kWin + 128u * lane));
            }
            __syncthreads();
            // nearest earlier position of this window (<= 31 back) with my hash; the sentinels in front never match
            uint32_t d = 0u;
#pragma unroll
            for (uint32_t k = 31u; k >= 1u; k--) if (sm.hs[32u + tid - k] == myh) d = k;
            if (inw) {
                uint32_t slen = 0u;
                if (act) {
                    // the slide of a 65275..65535-byte chunk at strstart 65274 (deflate.c:1285-1299), see quick_parse_warp
                    if (q == kWSize + kMaxDist && len < kChunkMax && cand < kWSize) cand = kWSize;
                    if ((q - cand - 1u) < kMaxDist) slen = match12(W, q, cand, v, x, len);
                }
                sm.ring[tid] = make_uint2((h << 16) | cand, (v & 0xffu) | (slen << 8) | ((uint32_t)act << 12) | (d << 16));
            }
            __syncthreads();
            if (warp == ww) {
                const uint32_t nwin = min((uint32_t)kWin, len - p);
                uint32_t cur = 0;
                bool ovf = false;
                while (cur < nwin) {
                    cur += walk_step<WARPS>(W, sm, p, cur, nwin, len, head, tok, wr, ovf, lane);
                    if (ovf) break;
                }
                if (lane == 0) sm.adv = cur;
            }
            __syncthreads();
            p += sm.adv;
        }
        {   // the last window's inserts
            sm.bitmap[prev_h >> 5] = 0u;
        }
        if (warp == ww) {
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

template <int WARPS>
static cudaError_t launch_cta(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks, uint32_t* tokens, uint32_t tok_stride,
                              uint32_t* ntok, uint32_t* counter, uint16_t* heads, unsigned long long* sm_slots, uint32_t grid,
                              const uint8_t* tail, uint32_t tail_first, cudaStream_t stream, const StreamSync& sy) {
    quick_parse_cta_kernel<WARPS><<<grid, WARPS * 32, 0, stream>>>(in, n, chunk, nchunks, tokens, tok_stride, ntok, counter, heads, sm_slots,
                                                                    tail, tail_first, sy);
    return cudaGetLastError();
}

cudaError_t launch_quick_parse_cta(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks,
                                   uint32_t* tokens, uint32_t tok_stride, uint32_t* ntok, uint32_t* counter,
                                   uint16_t* heads, unsigned long long* sm_slots, int num_sms, int chains_per_sm, int warps, uint8_t* tail,
                                   cudaStream_t stream, const StreamSync* sync) {
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
    switch (warps) {
        case 2:  return launch_cta<2>(in, n, chunk, nchunks, tokens, tok_stride, ntok, counter, heads, sm_slots, grid, tl, tail_first, stream, sy);
        case 4:  return launch_cta<4>(in, n, chunk, nchunks, tokens, tok_stride, ntok, counter, heads, sm_slots, grid, tl, tail_first, stream, sy);
        default: return launch_cta<8>(in, n, chunk, nchunks, tokens, tok_stride, ntok, counter, heads, sm_slots, grid, tl, tail_first, stream, sy);
    }
}

}  // namespace zb
