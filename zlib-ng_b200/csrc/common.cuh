// common.cuh -- device helpers shared by the sm_100a kernels of the chunked-DEFLATE hot path.
//
// Everything here is integer/byte work (no tensor cores: nothing on this path is a dense
// contraction).  Reference semantics are cited per function as file:line under /root/reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ZB_FULL 0xffffffffu

namespace zb {

constexpr uint32_t kWSize      = 32768u;                 // windowBits 15
constexpr uint32_t kMaxDist    = kWSize - 262u;          // deflate.h:410-415  MAX_DIST = w_size - MIN_LOOKAHEAD
constexpr uint32_t kMaxMatch   = 258u;                   // STD_MAX_MATCH
constexpr uint32_t kWantMin    = 4u;                     // WANT_MIN_MATCH
constexpr uint32_t kChunkMax   = 65536u;                 // one zng_deflate(Z_FULL_FLUSH) unit
constexpr uint32_t kWinPad     = 320u;                   // readable slack after the data (compare256 over-read)
constexpr uint32_t kCrcPoly    = 0xedb88320u;            // crc32_braid_p.h reflected CRC-32 polynomial
constexpr uint32_t kAdlerBase  = 65521u;                 // adler32_p.h:11

// token produced by the LZ77 parse: literal byte, or kTokMatch | len<<16 | dist
constexpr uint32_t kTokMatch = 0x80000000u;
constexpr uint32_t kTokEnd   = 0x40000000u;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// mask of lanes [a, b)   (0 <= a <= b <= 32)
__device__ __forceinline__ uint32_t lane_range(uint32_t a, uint32_t b) {
    uint32_t hi = b >= 32u ? 0xffffffffu : ((1u << b) - 1u);
    uint32_t lo = a >= 32u ? 0xffffffffu : ((1u << a) - 1u);
    return hi & ~lo;
}

// insert_string.c:13  HASH_CALC: (le32 * 2654435761) >> 16
__device__ __forceinline__ uint32_t hash4(uint32_t le32) { return (le32 * 2654435761u) >> 16; }

// ---------------------------------------------------------------- GF(2) helpers for CRC-32
// crc32_braid_comb_p.h:8-23  a(x)*b(x) mod p(x), bit 31 = x^0 (no early exit: same value)
__host__ __device__ __forceinline__ uint32_t multmodp(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#pragma unroll 8
    for (int i = 0; i < 32; i++) {
        p ^= (a & (0x80000000u >> i)) ? b : 0u;
        b = (b >> 1) ^ ((b & 1u) ? kCrcPoly : 0u);
    }
    return p;
}

// x2n[k] = x^(2^k) mod p  (crc32_braid_tbl.h:9437-9444 holds the same 32 values)
__host__ __device__ __forceinline__ void build_x2n(uint32_t* x2n /*[32]*/) {
    uint32_t v = 0x40000000u;
    x2n[0] = v;
    for (int k = 1; k < 32; k++) { v = multmodp(v, v); x2n[k] = v; }
}

// crc32_braid_comb_p.h:29-40  x^(n * 2^k) mod p
__host__ __device__ __forceinline__ uint32_t x2nmodp(const uint32_t* x2n, uint64_t n, unsigned k) {
    uint32_t p = 0x80000000u;
    while (n) {
        if (n & 1u) p = multmodp(x2n[k & 31u], p);
        n >>= 1; k++;
    }
    return p;
}

// crc32_braid_comb.c:16-18  crc32_combine(crc1, crc2, len2)
__host__ __device__ __forceinline__ uint32_t crc32_combine_dev(const uint32_t* x2n, uint32_t crc1, uint32_t crc2, uint64_t len2) {
    return multmodp(x2nmodp(x2n, len2, 3), crc1) ^ crc2;
}

// adler32.c:32-54  adler32_combine_ (len2 >= 0)
__host__ __device__ __forceinline__ uint32_t adler32_combine_dev(uint32_t a1, uint32_t a2, uint64_t len2) {
    uint32_t rem = (uint32_t)(len2 % kAdlerBase);
    uint32_t s1 = a1 & 0xffffu;
    uint32_t s2 = (rem * s1) % kAdlerBase;
    s1 += (a2 & 0xffffu) + kAdlerBase - 1u;
    s2 += ((a1 >> 16) & 0xffffu) + ((a2 >> 16) & 0xffffu) + kAdlerBase - rem;
    if (s1 >= kAdlerBase) s1 -= kAdlerBase;
    if (s1 >= kAdlerBase) s1 -= kAdlerBase;
    if (s2 >= (kAdlerBase << 1)) s2 -= (kAdlerBase << 1);
    if (s2 >= kAdlerBase) s2 -= kAdlerBase;
    return s1 | (s2 << 16);
}

// byte-wise CRC table entry (tools/makecrct.c analogue), used to fill shared-memory tables
__device__ __forceinline__ uint32_t crc_table_entry(uint32_t i, int slice) {
    uint32_t c = i;
#pragma unroll
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1u) ? kCrcPoly : 0u);
    for (int s = 0; s < slice; s++) {                     // advance by one more zero byte per slice
#pragma unroll
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1u) ? kCrcPoly : 0u);
    }
    return c;
}

// ---------------------------------------------------------------- hash-table slab pool (K1 / K2 parse kernels)
// Hash-head slabs are handed out per SM: slab (smid, b) belongs to the warp that holds bit b of
// sm_slots[smid].  This lets any number of parse kernels (different streams, different batches) run
// concurrently on one pool, sized nsmid x 64 (a resident warp always finds a free bit).
__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

__device__ __forceinline__ uint32_t slot_acquire(unsigned long long* word) {
    // 64 slots per SM and at most 64 resident warps: a free bit always exists unless a slot leaked (a kernel that died holding one).
    // The spin is bounded so that such a leak ends in a trap (a loud CUDA error at the next synchronisation), never in a hang.
    for (uint32_t tries = 0; tries < (1u << 24); tries++) {
        const unsigned long long m = *(volatile unsigned long long*)word;
        const int b = __ffsll((long long)~m) - 1;
        if (b >= 0) {
            const unsigned long long bit = 1ull << b;
            if (!(atomicOr(word, bit) & bit)) return (uint32_t)b;
        } else __nanosleep(200);
    }
    __trap();
    return 0;
}


// ---------------------------------------------------------------- static Huffman symbol coding
// RFC 1951 3.2.5/3.2.6; the same values tools/maketrees.c writes to trees_tbl.h.  Returns the
// LSB-first bit string of one token under the fixed code (trees_emit.h:102-164 semantics).
__device__ __forceinline__ void fixed_code_token(uint32_t tok, uint32_t& bits, uint32_t& nbits) {
    if (!(tok & kTokMatch)) {
        uint32_t c = tok & 0xffu;
        if (c < 144u) { bits = __brev(0x30u + c) >> 24; nbits = 8; }
        else          { bits = __brev(0x190u + (c - 144u)) >> 23; nbits = 9; }
        return;
    }
    uint32_t len = (tok >> 16) & 0x1ffu, dist = tok & 0xffffu;
    uint32_t lc = len - 3u, ls, lx, lxb;                 // length symbol, extra value, extra bits
    if (lc < 8u) { ls = lc; lx = 0; lxb = 0; }
    else if (lc == 255u) { ls = 28; lx = 0; lxb = 0; }
    else {
        uint32_t n = 31u - __clz(lc);                    // 3..7
        lxb = n - 2u;
        ls = 4u * (n - 1u) + ((lc >> lxb) & 3u);
        lx = lc & ((1u << lxb) - 1u);
    }
    uint32_t b, nb;
    if (ls < 23u) { b = __brev(ls + 1u) >> 25; nb = 7; }            // symbols 257..279: 7 bits
    else          { b = __brev(0xC0u + (ls - 23u)) >> 24; nb = 8; } // symbols 280..285: 8 bits
    b |= lx << nb; nb += lxb;
    uint32_t d = dist - 1u, ds, dx, dxb;
    if (d < 4u) { ds = d; dx = 0; dxb = 0; }
    else {
        uint32_t n = 31u - __clz(d);
        dxb = n - 1u;
        ds = 2u * n + ((d >> dxb) & 1u);
        dx = d & ((1u << dxb) - 1u);
    }
    b |= (__brev(ds) >> 27) << nb; nb += 5u;
    b |= dx << nb; nb += dxb;                            // <= 8+5+5+13 = 31 bits
    bits = b; nbits = nb;
}

}  // namespace zb
