// checksum.cu -- K3: CRC-32 / Adler-32 over tiles, and the combine (fold) kernels.
//
// Reference semantics (files under /root/reference):
//   crc32.c:27-41, arch/generic/crc32_braid_c.c:62-216   zng_crc32: reflected 0xEDB88320, ~ in / ~ out
//   crc32_braid_comb.c:16-24, crc32_braid_comb_p.h:8-40   crc32_combine = multmodp(x^(8*len2), crc1) ^ crc2
//   arch/generic/adler32_c.c:11-54, adler32_p.h:11-12      zng_adler32: BASE 65521
//   adler32.c:32-54                                        adler32_combine
//
// Tile kernel (v2).  The SM has no carry-less multiply, so CRC-32 is table driven: one shared-memory lookup per
// input byte.  To run those lookups at the full 32 lanes / clock, the four byte tables are stored ONCE PER LANE
// (address = table + byte*128 + lane*4: lane l only ever touches bank l -- conflict free; 4 x 32 KiB), and the data
// never passes through shared memory at all: one warp owns one tile (<= 64 KiB) and streams it with coalesced
// 128-bit loads, 512 bytes per row.  Lane l keeps four accumulators, one per 32-bit word of its 16 bytes; the
// words of one accumulator are 512 bytes apart in the stream, so its tables advance the CRC state by 512 bytes
// instead of 4 (T_k[b] = (b << 8k) * x^(8*512) mod P) -- same number of lookups as slicing-by-4, but coalesced loads
// and 4 independent dependency chains per lane.  After the last row each accumulator is multiplied by
// x^(8 * bytes between its word and the end of the rows) and the 128 accumulators are XORed.  Unaligned heads (< 16
// bytes) and tails (< 512 bytes) go through a plain slicing-by-4 loop on lane 0.  Adler-32 rides along on the same
// registers with dp4a: per word sum(b) and sum(t*b), per row a running sum gives the position weights.
// Fold kernels: per-tile values -> one value; Horner over equal-length tiles, then crc32_combine's algebra.
#include "common.cuh"
#include "kernels.h"

namespace zb {

constexpr int      kCkWarps  = 32;           // tiles in flight per CTA (one per warp)
constexpr uint32_t kCkRow    = 512;          // bytes per warp row (32 lanes x 16)

constexpr uint32_t kCkBigEntries = 256 + 256 + 512 + 128;                      // fields of 8, 8, 9 and 7 bits
struct X2N { uint32_t v[32]; };
// shifts done as multiplies (FMA pipe): R = 32 (128-byte entries): w >> 8, w >> 16, w << 7; R = 1 (4-byte entries): w >> 13, w >> 21, w << 2, w >> 5
struct CkMul { uint32_t m[7]; };            // x^(2^k) mod P, k = 0..31 (crc32_braid_tbl.h:9437-9444 holds the same values)

template <int R>                             // R = 32: one copy of the tables per lane (conflict free, 128 KiB); R = 1: one shared copy (4 KiB)
struct CkShared {
    uint32_t big[kCkBigEntries][R];          // the advance-by-512-bytes tables of the four bit fields (see ZB_LK)
    uint32_t t4[4][256];                     // ordinary slicing-by-4 tables (heads / tails)
    uint32_t klast[128];                     // x^(8 * (512 - 16*lane - 4*k)): last-row accumulators -> end of the rows
    uint32_t x2n[32];
};

// c advanced over one 32-bit word of data (slicing-by-4, crc32_braid_c.c semantics for N = 1, W = 4)
template <int R>
__device__ __forceinline__ uint32_t crc_word(const CkShared<R>& s, uint32_t c, uint32_t w) {
    w ^= c;
    return s.t4[3][w & 0xffu] ^ s.t4[2][(w >> 8) & 0xffu] ^ s.t4[1][(w >> 16) & 0xffu] ^ s.t4[0][w >> 24];
}

// bytes [p, p+len) through lane-0 style serial code (any alignment)
template <int R>
__device__ uint32_t crc_serial(const CkShared<R>& s, uint32_t c, const uint8_t* p, uint32_t len) {
    while (len && (reinterpret_cast<uintptr_t>(p) & 3u)) { c = s.t4[0][(c ^ *p++) & 0xffu] ^ (c >> 8); len--; }
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
    for (; len >= 4u; len -= 4u) c = crc_word(s, c, __ldg(w++));
    p = reinterpret_cast<const uint8_t*>(w);
    while (len--) c = s.t4[0][(c ^ *p++) & 0xffu] ^ (c >> 8);
    return c;
}

template <bool kCrc, bool kAdler, int R = 32>
__global__ void __launch_bounds__(kCkWarps * 32, 1)
checksum_tiles_kernel(const uint8_t* __restrict__ in, size_t n, uint32_t tile_bytes, uint32_t ntiles,
                      uint32_t* __restrict__ crcs, uint32_t* __restrict__ adlers, const X2N x2n_host, const CkMul mul) {
    extern __shared__ __align__(16) unsigned char ck_smem[];
    CkShared<R>& s = *reinterpret_cast<CkShared<R>*>(ck_smem);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const unsigned cta_warps = blockDim.x >> 5;              // 32 when the kernel has the SM to itself, fewer next to a parse kernel
    if (kCrc) {
        if (tid < 32u) s.x2n[tid] = x2n_host.v[tid];
        for (unsigned e = tid; e < 1024u; e += blockDim.x) s.t4[e >> 8][e & 255u] = crc_table_entry(e & 255u, (int)(e >> 8));
        __syncthreads();
        const uint32_t adv = x2nmodp(s.x2n, kCkRow, 3);                         // x^(8*512)
        for (unsigned e = tid; e < kCkBigEntries; e += blockDim.x) {
            // entry e = the word with one field set, 512 bytes later: bits 14..7 | 22..15 | 31..23 | 6..0 (see ZB_LK)
            const uint32_t word = e < 256u ? e << 7 : (e < 512u ? (e - 256u) << 15 : (e < 1024u ? (e - 512u) << 23 : e - 1024u));
            const uint32_t v = multmodp(adv, word);
#pragma unroll 8
            for (int r = 0; r < R; r++) s.big[e][r] = v;
        }
        if (tid < 128u) s.klast[tid] = x2nmodp(s.x2n, kCkRow - 16u * (tid >> 2) - 4u * (tid & 3u), 3);
        __syncthreads();
    }
    // byte offsets into `big`: entry e at e*4*R, (R = 32) this lane's copy at lane*4
    const unsigned char* bigb = reinterpret_cast<const unsigned char*>(&s.big[0][0]) + (R == 32 ? lane * 4u : 0u);
    const uint32_t m1 = mul.m[R == 32 ? 0 : 3], m2 = mul.m[R == 32 ? 1 : 4], m3 = mul.m[R == 32 ? 2 : 5];

    for (uint32_t ti = blockIdx.x * cta_warps + warp; ti < ntiles; ti += gridDim.x * cta_warps) {
        const size_t off = (size_t)ti * tile_bytes;
        const uint32_t len = (uint32_t)min((size_t)tile_bytes, n - off);
        const uint8_t* src = in + off;
        const uint32_t head = min((uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u)) & 15u), len);
        const uint32_t rows = (len - head) / kCkRow;
        const uint32_t body = rows * kCkRow, tail = len - head - body;
        const uint4* rowp = reinterpret_cast<const uint4*>(src + head) + lane;

        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;                                 // CRC accumulators (words 0..3 of the lane's 16 bytes)
        uint32_t A = 0, U = 0, Q = 0;                                            // Adler: sum b, sum (offset in lane piece)*b, row-weighted sum
        if (kCrc && lane == 0) c0 = crc_serial(s, 0xffffffffu, src, head);       // pre-inversion + head bytes ride on the first word
        if (rows) {
            // three rows in flight per warp (32 warps x 3 x 512 B = 48 KiB per SM) to cover the HBM latency
            // The four lookups of a word need not cut it at byte boundaries (CRC is linear in the bits), and the loop is bound by the
            // ALU pipe (LOP3 / SHF: 84 % busy in profiles/r2_checksum_ncu_summary.txt) while the FMA pipe idles.  So: field 0 = bits
            // 14..7 sits where the table offset wants it (no shift); fields 1 and 2 = bits 22..15 and 31..23 come down with a
            // multiply-high, field 3 = bits 6..0 goes up with a multiply -- IMAD on the FMA pipe; the multipliers are kernel
            // parameters so that they stay multiplies.  One LOP3 per lookup (mask + this lane's table copy) is what the ALU keeps.
#define ZB_OFF(e0, x, bits) ((e0) * (4u * R) + ((x) & (((1u << (bits)) - 1u) << (R == 32 ? 7 : 2))))
#define ZB_LKAT(off) (*reinterpret_cast<const uint32_t*>(bigb + (off)))
#define ZB_CRC_WORD(w) (ZB_LKAT(ZB_OFF(0u, R == 32 ? (w) : __umulhi((w), mul.m[6]), 8)) ^ ZB_LKAT(ZB_OFF(256u, __umulhi((w), m1), 8)) ^ \
                        ZB_LKAT(ZB_OFF(512u, __umulhi((w), m2), 9)) ^ ZB_LKAT(ZB_OFF(1024u, (w) * m3, 7)))
#define ZB_FOLD_ROW(v)                                                                                              \
            do {                                                                                                    \
                if (kCrc) {                                                                                         \
                    const uint32_t w0 = (v).x ^ c0, w1 = (v).y ^ c1, w2 = (v).z ^ c2, w3 = (v).w ^ c3;              \
                    c0 = ZB_CRC_WORD(w0);                   \
                    c1 = ZB_CRC_WORD(w1);                   \
                    c2 = ZB_CRC_WORD(w2);                   \
                    c3 = ZB_CRC_WORD(w3);                   \
                }                                                                                                   \
                if (kAdler) {                                                                                       \
                    const uint32_t s0 = __dp4a((v).x, 0x01010101u, 0u), s1 = __dp4a((v).y, 0x01010101u, 0u);        \
                    const uint32_t s2 = __dp4a((v).z, 0x01010101u, 0u), s3 = __dp4a((v).w, 0x01010101u, 0u);        \
                    U = __dp4a((v).x, 0x03020100u, __dp4a((v).y, 0x03020100u, __dp4a((v).z, 0x03020100u, __dp4a((v).w, 0x03020100u, U)))); \
                    U += 4u * (s1 + 2u * s2 + 3u * s3);                                                             \
                    Q += A;                                                                                         \
                    A += s0 + s1 + s2 + s3;                                                                         \
                }                                                                                                   \
            } while (0)
            const uint32_t adv = rows - 1u;                                      // rows folded through the tables; the last one is multiplied
            // fold a row, THEN refill its registers (no copy of the row on the loop's back edge); the refill address is clamped to
            // the last row instead of predicated (a few redundant L1 hits at the end of a tile, no 64-bit address arithmetic per row)
            uint4 q0 = __ldg(rowp), q1 = __ldg(rowp + min(1u, adv) * 32u), q2 = __ldg(rowp + min(2u, adv) * 32u);
            uint32_t r = 0;
            for (; r + 3u <= adv; r += 3u) {
                ZB_FOLD_ROW(q0); q0 = __ldg(rowp + min(r + 3u, adv) * 32u);
                ZB_FOLD_ROW(q1); q1 = __ldg(rowp + min(r + 4u, adv) * 32u);
                ZB_FOLD_ROW(q2); q2 = __ldg(rowp + min(r + 5u, adv) * 32u);
            }
            const uint32_t rem = adv - r;                                        // 0..2 table rows left; q0, q1, q2 hold rows r, r+1, r+2 (clamped)
            if (rem >= 1u) ZB_FOLD_ROW(q0);
            if (rem >= 2u) ZB_FOLD_ROW(q1);
            const uint4 v = rem == 0u ? q0 : (rem == 1u ? q1 : q2);
            if (kCrc) {                                                          // last row: to the end of the rows by multiplication
                c0 = multmodp(s.klast[4u * lane + 0u], c0 ^ v.x);
                c1 = multmodp(s.klast[4u * lane + 1u], c1 ^ v.y);
                c2 = multmodp(s.klast[4u * lane + 2u], c2 ^ v.z);
                c3 = multmodp(s.klast[4u * lane + 3u], c3 ^ v.w);
            }
            if (kAdler) {
                const uint32_t s0 = __dp4a(v.x, 0x01010101u, 0u), s1 = __dp4a(v.y, 0x01010101u, 0u);
                const uint32_t s2 = __dp4a(v.z, 0x01010101u, 0u), s3 = __dp4a(v.w, 0x01010101u, 0u);
                U = __dp4a(v.x, 0x03020100u, __dp4a(v.y, 0x03020100u, __dp4a(v.z, 0x03020100u, __dp4a(v.w, 0x03020100u, U))));
                U += 4u * (s1 + 2u * s2 + 3u * s3);
                Q += A;
                A += s0 + s1 + s2 + s3;
            }
        }
        if (kCrc) {
            uint32_t c = c0 ^ c1 ^ c2 ^ c3;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) c ^= __shfl_xor_sync(ZB_FULL, c, d);
            if (lane == 0) {
                c = crc_serial(s, c, src + head + body, tail);
                if (crcs) crcs[ti] = ~c;
            }
        }
        if (kAdler) {
            // byte j of the tile weighs (len - j); a byte of the rows sits at j = head + 512 r + 16 lane + (offset in piece)
            unsigned long long s1 = A, s2 = 0;
            if (rows) {
                const unsigned long long rem_last = (unsigned long long)(len - head) - (unsigned long long)kCkRow * (rows - 1u);
                s2 = (rem_last - 16ull * lane) * A + (unsigned long long)kCkRow * Q - U;
            }
            for (uint32_t j = lane; j < head; j += 32u) { const uint32_t b = src[j]; s1 += b; s2 += (unsigned long long)(len - j) * b; }
            for (uint32_t j = head + body + lane; j < len; j += 32u) { const uint32_t b = src[j]; s1 += b; s2 += (unsigned long long)(len - j) * b; }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) { s1 += __shfl_xor_sync(ZB_FULL, s1, d); s2 += __shfl_xor_sync(ZB_FULL, s2, d); }
            if (lane == 0 && adlers) adlers[ti] = (uint32_t)((1ull + s1) % kAdlerBase) | ((uint32_t)((len + s2) % kAdlerBase) << 16);
        }
    }
}

// ---------------------------------------------------------------- folds (single CTA; 8 B per tile of input)
// crc = crc32_combine(...combine(combine(init, c_0, l_0), c_1, l_1)..., c_{k-1}, l_{k-1})            (crc32_braid_comb.c:16-18)
// Every thread folds a run of consecutive tiles by Horner (acc = acc * x^(8*tile_bytes) ^ c_i: one multmodp per tile),
// moves its partial to the end of the buffer with one x2nmodp, and the partials are XORed.
__global__ void __launch_bounds__(1024)
crc32_fold_kernel(const uint32_t* __restrict__ crcs, uint32_t ntiles, uint32_t tile_bytes, size_t n, uint32_t init,
                  uint32_t* __restrict__ result, const X2N x2n_host) {
    __shared__ uint32_t x2n[32];
    __shared__ uint32_t red[32];
    if (threadIdx.x < 32u) x2n[threadIdx.x] = x2n_host.v[threadIdx.x];
    __syncthreads();
    const uint32_t per = (ntiles + blockDim.x - 1u) / blockDim.x;
    const uint32_t t0 = min(threadIdx.x * per, ntiles), t1 = min(t0 + per, ntiles);
    uint32_t acc = 0;
    if (t0 < t1) {
        const uint32_t ktile = x2nmodp(x2n, tile_bytes, 3);
        for (uint32_t i = t0; i < t1; i++) {
            const size_t beg = (size_t)i * tile_bytes;
            const size_t li = min((size_t)tile_bytes, n - beg);
            acc = (acc ? multmodp(li == tile_bytes ? ktile : x2nmodp(x2n, li, 3), acc) : 0u) ^ crcs[i];
        }
        const size_t end = min((size_t)t1 * tile_bytes, n);
        if (n - end && acc) acc = multmodp(x2nmodp(x2n, n - end, 3), acc);
    }
    if (threadIdx.x == 0 && init) acc ^= n ? multmodp(x2nmodp(x2n, n, 3), init) : init;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc ^= __shfl_xor_sync(ZB_FULL, acc, d);
    if ((threadIdx.x & 31u) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = 0;
        for (unsigned w = 0; w < (blockDim.x + 31u) / 32u; w++) c ^= red[w];
        *result = c;
    }
}

// adler(init, data): s1 = s1_0 + sum b ; s2 = s2_0 + n*s1_0 + sum b_j*(n-j)      (adler32_c.c:11-54)
// per tile i (length L_i, `after_i` bytes follow): A_i = 1 + sum b,  B_i = L_i + sum b_j*(L_i - j)
__global__ void __launch_bounds__(1024)
adler32_fold_kernel(const uint32_t* __restrict__ adlers, uint32_t ntiles, uint32_t tile_bytes, size_t n, uint32_t init,
                    uint32_t* __restrict__ result) {
    __shared__ unsigned long long red1[32], red2[32];
    unsigned long long a = 0, b = 0;
    for (uint32_t i = threadIdx.x; i < ntiles; i += blockDim.x) {
        const size_t beg = (size_t)i * tile_bytes;
        const size_t end = min(beg + tile_bytes, n);
        const unsigned long long L = end - beg, after = n - end;
        const uint32_t v = adlers[i];
        const unsigned long long Ai = (v & 0xffffu) + kAdlerBase - 1u;                  // sum b  (mod BASE)
        const unsigned long long Bi = ((v >> 16) & 0xffffu) + kAdlerBase - (L % kAdlerBase);   // sum b_j*(L-j)
        a += Ai % kAdlerBase;
        b += (Bi + (Ai % kAdlerBase) * (after % kAdlerBase)) % kAdlerBase;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(ZB_FULL, a, d); b += __shfl_xor_sync(ZB_FULL, b, d); }
    if ((threadIdx.x & 31u) == 0) { red1[threadIdx.x >> 5] = a; red2[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long sa = 0, sb = 0;
        for (unsigned w = 0; w < (blockDim.x + 31u) / 32u; w++) { sa += red1[w] % kAdlerBase; sb += red2[w] % kAdlerBase; }
        const unsigned long long s10 = init & 0xffffu, s20 = (init >> 16) & 0xffffu;
        const unsigned long long s1 = (s10 + sa) % kAdlerBase;
        const unsigned long long s2 = (s20 + (unsigned long long)(n % kAdlerBase) * s10 + sb) % kAdlerBase;
        *result = (uint32_t)s1 | ((uint32_t)s2 << 16);
    }
}

static const X2N& host_x2n() {
    static const X2N t = [] { X2N x; build_x2n(x.v); return x; }();
    return t;
}

template <bool kCrc, bool kAdler, int R>
static cudaError_t launch_tiles(const uint8_t* in, size_t n, uint32_t tile_bytes, uint32_t ntiles, uint32_t* crcs, uint32_t* adlers,
                                int num_sms, cudaStream_t stream, int cta_warps) {
    const int smem = kCrc ? (int)sizeof(CkShared<R>) : 0;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(checksum_tiles_kernel<kCrc, kAdler, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    uint32_t grid = (uint32_t)num_sms * (kCrc ? 1u : 2u);
    const uint32_t need = (ntiles + (uint32_t)cta_warps - 1u) / (uint32_t)cta_warps;
    if (grid > need) grid = need;
    static const CkMul mul = {{1u << 24, 1u << 16, 1u << 7, 1u << 19, 1u << 11, 1u << 2, 1u << 27}};
    checksum_tiles_kernel<kCrc, kAdler, R><<<grid, cta_warps * 32, smem, stream>>>(in, n, tile_bytes, ntiles, crcs, adlers, host_x2n(), mul);
    return cudaGetLastError();
}

// cta_warps: 32 = the whole SM (device-resident calls); the host pipelines pass 8 so that the CTAs fit NEXT TO a running
// parse kernel: 1024 threads x 56 registers and 133 KiB of lane-private tables do not, and the checksums of a finished slab
// would wait for the parse to end
cudaError_t launch_checksum_tiles(const uint8_t* in, size_t n, uint32_t tile_bytes, uint32_t ntiles,
                                  uint32_t* crcs, uint32_t* adlers, int num_sms, cudaStream_t stream, int cta_warps) {
    if (ntiles == 0) return cudaSuccess;
    if (cta_warps < 1 || cta_warps > kCkWarps) cta_warps = kCkWarps;
    if (cta_warps < kCkWarps) {     // next to a parse kernel: 8 KiB of tables (bank conflicts cost ~3x on the lookups, the slab is small)
        if (crcs && adlers) return launch_tiles<true, true, 1>(in, n, tile_bytes, ntiles, crcs, adlers, num_sms, stream, cta_warps);
        if (crcs) return launch_tiles<true, false, 1>(in, n, tile_bytes, ntiles, crcs, nullptr, num_sms, stream, cta_warps);
    }
    if (crcs && adlers) return launch_tiles<true, true, 32>(in, n, tile_bytes, ntiles, crcs, adlers, num_sms, stream, cta_warps);
    if (crcs) return launch_tiles<true, false, 32>(in, n, tile_bytes, ntiles, crcs, nullptr, num_sms, stream, cta_warps);
    if (adlers) return launch_tiles<false, true, 32>(in, n, tile_bytes, ntiles, nullptr, adlers, num_sms, stream, cta_warps);
    return cudaSuccess;
}

cudaError_t launch_crc32_fold(const uint32_t* crcs, uint32_t ntiles, uint32_t tile_bytes, size_t n, uint32_t init,
                              uint32_t* result, cudaStream_t stream) {
    crc32_fold_kernel<<<1, 1024, 0, stream>>>(crcs, ntiles, tile_bytes, n, init, result, host_x2n());
    return cudaGetLastError();
}

cudaError_t launch_adler32_fold(const uint32_t* adlers, uint32_t ntiles, uint32_t tile_bytes, size_t n, uint32_t init,
                                uint32_t* result, cudaStream_t stream) {
    adler32_fold_kernel<<<1, 1024, 0, stream>>>(adlers, ntiles, tile_bytes, n, init, result);
    return cudaGetLastError();
}

}  // namespace zb
