// deflate_quick.cu -- K1: level-1 (deflate_quick) compression of independent <= 64 KiB chunks.
//
// Reference semantics reproduced bit-exactly (files under /root/reference):
//   deflate_quick.c:47-130        greedy single-probe parse, static-Huffman emission
//   insert_string_tpl.h:58-75     quick_insert_string (head[] update only at visited positions)
//   arch/generic/compare256_c.c   first-mismatch compare (here: 32 lanes x 8 bytes + ballot/ffs)
//   trees_emit.h:102-225          zng_tr_emit_lit / zng_tr_emit_dist / emit_tree / emit_end_block
//   deflate.c:1061-1083           empty stored block appended for Z_SYNC_FLUSH / Z_FULL_FLUSH
//   crc32.c:27-41, adler32_c.c    per-chunk zng_crc32 / zng_adler32 of the input (fused: the chunk
//                                 is already staged in shared memory)
//
// B200 mapping.  One persistent CTA per SM (148), 8 warps, ~221 KiB of shared memory:
//   window  64 KiB + pad   the chunk, staged once with 128-bit loads
//   head    128 KiB        the 65536-entry u16 hash-head table of deflate_state (deflate.h:232)
//   tokens  8 KiB ring     parser -> emitter
//   stage   16 KiB ring    bit-packed output, flushed to HBM with coalesced 128-bit stores
// Warp roles per chunk:
//   warp 0  PARSER.  The parse is serial per chunk in the reference (which positions get hashed
//           depends on every earlier match decision), so the warp speculates: 32 lanes hash 32
//           consecutive positions against the table state at the window start, test their
//           candidates and measure short matches (< 12 bytes) in parallel; a ballot/ffs walk then
//           replays the reference's decisions (literal runs are accepted wholesale, a match costs
//           one shuffle, or one warp-wide compare when it is 12 bytes or longer).
//           A lane whose hash equals that of an earlier VISITED lane of the same window saw a stale
//           candidate; the window is cut there and restarted, which keeps the result exact.
//   warp 1  EMITTER.  Per 32 tokens: fixed-code bits, warp prefix scan of the bit lengths, OR into
//           the staging ring, 4 KiB segments stored to global memory as uint4.
//   warps 2-7 CHECKSUM.  CRC-32 (slicing-by-4 over 512-byte slices + GF(2) shift-combine) and
//           Adler-32 of the chunk.
#include "common.cuh"
#include "kernels.h"

namespace zb {

constexpr int      kQThreads     = 256;
constexpr uint32_t kTokRing      = 2048;          // tokens (u32)
constexpr uint32_t kStageWords   = 4096;          // u32 words (16 KiB)
constexpr uint32_t kStageSeg     = 1024;          // words per flush segment (4 KiB)
constexpr uint32_t kCsSlice      = 512;           // bytes per checksum slice
constexpr uint32_t kCsThreads    = 128;           // 65536 / 512
constexpr uint32_t kSpinLimit    = 1u << 24;      // ring-wait watchdog (~1 s): trap instead of hanging

struct QuickSmem {
    uint32_t win[(kChunkMax + kWinPad) / 4];
    uint16_t head[65536];
    uint32_t tok[kTokRing];
    uint32_t stage[kStageWords];
    uint32_t crctab[4][256];
    uint32_t x2n[32];
    uint32_t red_crc[4];
    uint32_t red_s1[4];
    unsigned long long red_s2[4];
    volatile uint32_t tok_wr;      // tokens produced (monotonic)
    volatile uint32_t tok_rd;      // tokens consumed (monotonic)
    uint32_t chunk_idx;
};

// ---------------------------------------------------------------- parser (warp 0)
// Warp-wide first-mismatch over 256 bytes: lane l compares the 8 bytes at a+8l / b+8l.
// Returns the number of equal leading bytes (0..256).  (compare256 analogue, compare256_c.c:12-43)
__device__ __forceinline__ uint32_t warp_compare256(const uint32_t* win, uint32_t a, uint32_t b, unsigned lane) {
    uint64_t x = ld64u(win, a + 8u * lane) ^ ld64u(win, b + 8u * lane);
    unsigned diff = __ballot_sync(ZB_FULL, x != 0ull);
    if (diff == 0u) return 256u;
    unsigned f = __ffs(diff) - 1u;
    unsigned byte = (unsigned)(__ffsll((long long)x) - 1) >> 3;
    byte = __shfl_sync(ZB_FULL, byte, f);
    return 8u * f + byte;
}

__device__ __forceinline__ void tok_push(QuickSmem& s, bool mine, uint32_t rank, uint32_t tok, uint32_t cnt,
                                         uint32_t& wr, unsigned lane, uint32_t* dbg) {
    // wait for ring space (the emitter advances tok_rd); lane 0 polls, the warp follows
    for (uint32_t spins = 0;; spins++) {
        uint32_t rd = 0;
        if (lane == 0) rd = s.tok_rd;
        rd = __shfl_sync(ZB_FULL, rd, 0);
        if (wr + cnt - rd <= kTokRing) break;
        if (spins > kSpinLimit) __trap();                   // a stuck ring is a bug: fail, never hang the GPU
        __nanosleep(64);
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

__device__ void quick_parse_warp(QuickSmem& s, uint32_t n, uint32_t* dbg) {
    const unsigned lane = lane_id();
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t* win = s.win;
    uint32_t wr = 0;
    if (lane == 0) wr = s.tok_wr;
    wr = __shfl_sync(ZB_FULL, wr, 0);
    if (dbg) dbg -= wr;                                       // dbg[wr + rank] indexes from 0 for this chunk
    uint32_t p = 0;
    while (p < n) {
        const uint32_t q = p + lane;
        const bool inb = q < n;
        const bool act = q + kWantMin <= n;                  // deflate_quick.c:88 lookahead >= WANT_MIN_MATCH
        // 12 bytes at q: v = bytes 0..3 (hashed), x = bytes 4..11
        uint32_t v; uint64_t x;
        {
            const uint32_t i = q >> 2, sh = (q & 3u) << 3;
            const uint32_t a0 = win[i], a1 = win[i + 1], a2 = win[i + 2], a3 = win[i + 3];
            v = __funnelshift_r(a0, a1, sh);
            x = (uint64_t)__funnelshift_r(a1, a2, sh) | ((uint64_t)__funnelshift_r(a2, a3, sh) << 32);
        }
        const uint32_t h = hash4(v);
        const uint32_t cand = act ? (uint32_t)s.head[h] : 0u;
        // deflate_quick.c:90-92: 0 < dist <= MAX_DIST;  :96-99: 2 bytes equal and compare256+2 >= 4  <=>  4 bytes equal
        uint32_t slen = 0;                                    // 0 none, 4..11 exact, 12 = "12 or more"
        if (act && (q - cand - 1u) < kMaxDist) {
            const uint32_t i = cand >> 2, sh = (cand & 3u) << 3;
            const uint32_t b0 = win[i], b1 = win[i + 1], b2 = win[i + 2], b3 = win[i + 3];
            if (__funnelshift_r(b0, b1, sh) == v) {
                const uint64_t y = (uint64_t)__funnelshift_r(b1, b2, sh) | ((uint64_t)__funnelshift_r(b2, b3, sh) << 32);
                const uint64_t d = x ^ y;
                slen = d ? 4u + ((unsigned)(__ffsll((long long)d) - 1) >> 3) : 12u;
            }
        }
        const unsigned peers = __match_any_sync(ZB_FULL, act ? h : (0x10000u + lane));
        const unsigned low = peers & lt;                     // earlier lanes of this window with the same hash
        const unsigned M = __ballot_sync(ZB_FULL, slen != 0u);
        const unsigned nl = min(32u, n - p);                 // lanes that hold a byte
        uint32_t mytok = v & 0xffu;
        unsigned cur = 0, V = 0;
        while (cur < nl) {
            const unsigned rest = M & ~lane_range(0, cur);
            const unsigned k = rest ? (unsigned)(__ffs(rest) - 1) : nl;   // next lane claiming a match
            const unsigned run = lane_range(cur, k);                       // literal run before it
            const unsigned scope = run | (k < nl ? (1u << k) : 0u);
            const bool stale = ((scope >> lane) & 1u) && (low & (V | (run & lt))) != 0u;
            const unsigned S = __ballot_sync(ZB_FULL, stale);
            if (S) {                                         // cut the window at the first stale lane
                const unsigned j = __ffs(S) - 1u;
                V |= lane_range(cur, j);
                cur = j;
                break;
            }
            V |= run;
            if (k >= nl) { cur = nl; break; }
            const uint32_t qk = p + k;
            uint32_t len = __shfl_sync(ZB_FULL, slen, k);
            if (len >= 12u) {
                const uint32_t ck = __shfl_sync(ZB_FULL, cand, k);
                len = 12u + warp_compare256(win, qk + 12u, ck + 12u, lane);
            }
            len = min(len, n - qk);                          // deflate_quick.c:100-101 clip to lookahead
            len = min(len, kMaxMatch);                       // deflate_quick.c:102-103
            if (lane == k) mytok = kTokMatch | (len << 16) | (q - cand);
            V |= 1u << k;
            cur = k + len;
        }
        const bool vis = (V >> lane) & 1u;
        if (vis && act) s.head[h] = (uint16_t)q;             // insert_string_tpl.h:70-73 (visited positions only)
        tok_push(s, vis && inb, __popc(V & lt), mytok, __popc(V), wr, lane, dbg);
        p += cur;
    }
    // end marker
    tok_push(s, lane == 0, 0, kTokEnd, 1, wr, lane, dbg);
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
            for (uint32_t i = lane; i < cnt / 4u; i += 32u) { g[i] = sm[i]; sm[i] = make_uint4(0, 0, 0, 0); }
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
            __nanosleep(32);
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

// ---------------------------------------------------------------- checksums (warps 2..7)
__device__ void chunk_checksums(QuickSmem& s, uint32_t n, uint32_t* crc_out, uint32_t* adler_out, bool want_adler) {
    const int t = (int)threadIdx.x - 64;                    // 0..191; slices handled by t < 128
    uint32_t crc = 0; uint32_t s1 = 0; unsigned long long s2 = 0;
    if (t < (int)kCsThreads) {
        const uint32_t beg = (uint32_t)t * kCsSlice;
        if (beg < n || t == 0) {                             // slice 0 always carries the initial values
            const uint32_t end = min(beg + kCsSlice, n);
            const uint8_t* wb = reinterpret_cast<const uint8_t*>(s.win);
            uint32_t c = (t == 0) ? 0xffffffffu : 0u;       // crc32_braid_c.c:66 pre-inversion, carried by slice 0
            uint32_t i = beg;
            for (; i + 4u <= end; i += 4u) {
                uint32_t w = s.win[i >> 2] ^ c;
                c = s.crctab[3][w & 0xffu] ^ s.crctab[2][(w >> 8) & 0xffu] ^ s.crctab[1][(w >> 16) & 0xffu] ^ s.crctab[0][w >> 24];
            }
            for (; i < end; i++) c = (c >> 8) ^ s.crctab[0][(c ^ wb[i]) & 0xffu];
            // shift the slice remainder to the end of the chunk: * x^(8*(n-end))
            uint32_t after = n - end;
            crc = after ? multmodp(x2nmodp(s.x2n, after, 3), c) : c;
            if (want_adler) {
                uint32_t a = 0; unsigned long long b = 0; uint32_t len = end - beg;
                for (uint32_t j = beg; j < end; j++) { uint32_t by = wb[j]; a += by; b += (unsigned long long)(len - (j - beg)) * by; }
                s1 = a;
                s2 = b + (unsigned long long)a * after;     // sum b_j * (n - j)
            }
        }
    }
    // reduce over the 6 warps
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        crc ^= __shfl_xor_sync(ZB_FULL, crc, d);
        s1 += __shfl_xor_sync(ZB_FULL, s1, d);
        s2 += __shfl_xor_sync(ZB_FULL, s2, d);
    }
    const int w = t >> 5;                                   // 0..5, only 0..3 carry data
    if ((threadIdx.x & 31) == 0 && w < 4) { s.red_crc[w] = crc; s.red_s1[w] = s1; s.red_s2[w] = s2; }
    asm volatile("bar.sync 1, 192;" ::: "memory");          // named barrier: the 6 checksum warps only
    if (t == 0) {
        uint32_t c = s.red_crc[0] ^ s.red_crc[1] ^ s.red_crc[2] ^ s.red_crc[3];
        *crc_out = ~c;                                      // crc32_braid_c.c:214 post-inversion
        if (want_adler) {
            unsigned long long a = 1ull + s.red_s1[0] + s.red_s1[1] + s.red_s1[2] + s.red_s1[3];
            unsigned long long b = (unsigned long long)n + s.red_s2[0] + s.red_s2[1] + s.red_s2[2] + s.red_s2[3];
            *adler_out = (uint32_t)(a % kAdlerBase) | ((uint32_t)(b % kAdlerBase) << 16);
        }
    }
    asm volatile("bar.sync 1, 192;" ::: "memory");          // red_* may be rewritten for the next chunk
}

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kQThreads, 1)
deflate_quick_kernel(const uint8_t* __restrict__ in, size_t n, uint32_t chunk, uint32_t nchunks, int last,
                     uint8_t* __restrict__ out, size_t out_stride, uint32_t* __restrict__ sizes,
                     uint32_t* __restrict__ crcs, uint32_t* __restrict__ adlers, uint32_t* __restrict__ counter,
                     uint32_t* __restrict__ dbg_tokens, uint32_t dbg_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    QuickSmem& s = *reinterpret_cast<QuickSmem*>(smem_raw);
    const unsigned tid = threadIdx.x, warp = tid >> 5;

    // one-time tables
    for (uint32_t i = tid; i < 1024u; i += kQThreads) s.crctab[i >> 8][i & 255u] = crc_table_entry(i & 255u, (int)(i >> 8));
    if (tid == 0) { build_x2n(s.x2n); s.tok_wr = 0; s.tok_rd = 0; }
    for (uint32_t i = tid; i < kStageWords; i += kQThreads) s.stage[i] = 0u;

    for (;;) {
        __syncthreads();
        if (tid == 0) s.chunk_idx = atomicAdd(counter, 1u);
        __syncthreads();
        const uint32_t ci = s.chunk_idx;
        if (ci >= nchunks) break;
        const size_t off = (size_t)ci * chunk;
        const uint32_t len = (uint32_t)min((size_t)chunk, n - off);

        // stage the chunk (128-bit loads when aligned), zero the pad, clear the head table
        {
            const uint8_t* src = in + off;
            uint8_t* wb = reinterpret_cast<uint8_t*>(s.win);
            uint4* w4 = reinterpret_cast<uint4*>(s.win);
            const uint32_t nvec = len >> 4;
            if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0u) {
                const uint4* g = reinterpret_cast<const uint4*>(src);
                for (uint32_t i = tid; i < nvec; i += kQThreads) w4[i] = __ldg(g + i);
            } else {
                for (uint32_t i = tid; i < (nvec << 4); i += kQThreads) wb[i] = src[i];
            }
            for (uint32_t i = (nvec << 4) + tid; i < len; i += kQThreads) wb[i] = src[i];
            const uint32_t padend = min(kChunkMax + kWinPad, ((len + kWinPad + 15u) & ~15u));
            for (uint32_t i = len + tid; i < padend; i += kQThreads) wb[i] = 0;
            uint4* h4 = reinterpret_cast<uint4*>(s.head);
            for (uint32_t i = tid; i < 65536u * 2u / 16u; i += kQThreads) h4[i] = make_uint4(0, 0, 0, 0);   // CLEAR_HASH
        }
        __syncthreads();

        if (warp == 0) {
            quick_parse_warp(s, len, dbg_tokens ? dbg_tokens + (size_t)ci * dbg_stride : nullptr);
        } else if (warp == 1) {
            const bool open_block = (len > 0) || last;      // deflate_quick.c:53-63
            uint32_t sz = quick_emit_warp(s, out + (size_t)ci * out_stride, last, open_block);
            if ((tid & 31u) == 0) sizes[ci] = sz;
        } else {
            uint32_t dummy;
            chunk_checksums(s, len, crcs ? &crcs[ci] : &dummy, adlers ? &adlers[ci] : &dummy, adlers != nullptr);
        }
    }
}

size_t deflate_quick_smem_bytes() { return sizeof(QuickSmem); }

cudaError_t launch_deflate_quick(const uint8_t* in, size_t n, uint32_t chunk, uint32_t nchunks, int last,
                                 uint8_t* out, size_t out_stride, uint32_t* sizes, uint32_t* crcs,
                                 uint32_t* adlers, uint32_t* counter, int num_sms, cudaStream_t stream,
                                 uint32_t* dbg_tokens, uint32_t dbg_stride) {
    cudaError_t e = cudaFuncSetAttribute(deflate_quick_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(QuickSmem));
    if (e != cudaSuccess) return e;
    uint32_t grid = nchunks < (uint32_t)num_sms ? nchunks : (uint32_t)num_sms;
    if (grid == 0) return cudaSuccess;
    e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    deflate_quick_kernel<<<grid, kQThreads, sizeof(QuickSmem), stream>>>(in, n, chunk, nchunks, last, out, out_stride,
                                                                          sizes, crcs, adlers, counter, dbg_tokens, dbg_stride);
    return cudaGetLastError();
}

}  // namespace zb
