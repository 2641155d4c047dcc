// checksum.cu -- K3: CRC-32 / Adler-32 over tiles, and the combine (fold) kernels.
//
// Reference semantics (files under /root/reference):
//   crc32.c:27-41, arch/generic/crc32_braid_c.c:62-216   zng_crc32: reflected 0xEDB88320, ~ in / ~ out
//   crc32_braid_comb.c:16-24, crc32_braid_comb_p.h:8-40   crc32_combine = multmodp(x^(8*len2), crc1) ^ crc2
//   arch/generic/adler32_c.c:11-54, adler32_p.h:11-12      zng_adler32: BASE 65521
//   adler32.c:32-54                                        adler32_combine
//
// Tile kernel: one CTA stages a tile (<= 64 KiB; a deflate chunk when called per chunk) in shared
// memory with coalesced 128-bit loads; 256 threads each run a table-driven CRC over a 256-byte
// slice (initial value 0) and multiply the slice remainder by x^(8 * bytes-after-slice) so that a
// plain XOR over the slices is the CRC of the tile; Adler-32 is the pair (sum b, sum b*(n-j)).
// Fold kernels: per-tile values -> one value, using crc32_combine's algebra (associative, so every
// tile is shifted by the byte count that follows it and XOR-reduced) and adler32_combine's sums.
#include "common.cuh"
#include "kernels.h"

namespace zb {

constexpr int kCkThreads = 256;
constexpr uint32_t kCkSlice = 256;      // bytes per thread

struct CkSmem {
    uint32_t tile[kChunkMax / 4];
    uint32_t crctab[4][256];
    uint32_t x2n[32];
    uint32_t red_crc[8];
    uint32_t red_s1[8];
    unsigned long long red_s2[8];
};

__global__ void __launch_bounds__(kCkThreads)
checksum_tiles_kernel(const uint8_t* __restrict__ in, size_t n, uint32_t tile_bytes, uint32_t ntiles,
                      uint32_t* __restrict__ crcs, uint32_t* __restrict__ adlers) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CkSmem& s = *reinterpret_cast<CkSmem*>(smem_raw);
    const unsigned tid = threadIdx.x;
    for (uint32_t i = tid; i < 1024u; i += kCkThreads) s.crctab[i >> 8][i & 255u] = crc_table_entry(i & 255u, (int)(i >> 8));
    if (tid == 0) build_x2n(s.x2n);
    const uint8_t* wb = reinterpret_cast<const uint8_t*>(s.tile);

    for (uint32_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        __syncthreads();
        const size_t off = (size_t)ti * tile_bytes;
        const uint32_t len = (uint32_t)min((size_t)tile_bytes, n - off);
        const uint8_t* src = in + off;
        {
            uint8_t* tb = reinterpret_cast<uint8_t*>(s.tile);
            const uint32_t nvec = len >> 4;
            if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0u) {
                const uint4* g = reinterpret_cast<const uint4*>(src);
                uint4* t4 = reinterpret_cast<uint4*>(s.tile);
                for (uint32_t i = tid; i < nvec; i += kCkThreads) t4[i] = __ldg(g + i);
            } else {
                for (uint32_t i = tid; i < (nvec << 4); i += kCkThreads) tb[i] = src[i];
            }
            for (uint32_t i = (nvec << 4) + tid; i < len; i += kCkThreads) tb[i] = src[i];
        }
        __syncthreads();

        uint32_t crc = 0, s1 = 0; unsigned long long s2 = 0;
        const uint32_t beg = tid * kCkSlice;
        if (beg < len || tid == 0) {
            const uint32_t end = min(beg + kCkSlice, len);
            uint32_t c = (tid == 0) ? 0xffffffffu : 0u;          // pre-inversion rides on slice 0
            uint32_t i = beg;
            for (; i + 4u <= end; i += 4u) {
                uint32_t w = s.tile[i >> 2] ^ c;
                c = s.crctab[3][w & 0xffu] ^ s.crctab[2][(w >> 8) & 0xffu] ^ s.crctab[1][(w >> 16) & 0xffu] ^ s.crctab[0][w >> 24];
            }
            for (; i < end; i++) c = (c >> 8) ^ s.crctab[0][(c ^ wb[i]) & 0xffu];
            const uint32_t after = len - end;
            crc = after ? multmodp(x2nmodp(s.x2n, after, 3), c) : c;
            if (adlers) {
                uint32_t a = 0, b = 0;                            // 256 bytes * 255 * 256 < 2^32
                for (uint32_t j = beg; j < end; j++) { uint32_t by = wb[j]; a += by; b += (end - j) * by; }
                s1 = a;
                s2 = (unsigned long long)b + (unsigned long long)a * after;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            crc ^= __shfl_xor_sync(ZB_FULL, crc, d);
            s1 += __shfl_xor_sync(ZB_FULL, s1, d);
            s2 += __shfl_xor_sync(ZB_FULL, s2, d);
        }
        if ((tid & 31u) == 0) { s.red_crc[tid >> 5] = crc; s.red_s1[tid >> 5] = s1; s.red_s2[tid >> 5] = s2; }
        __syncthreads();
        if (tid == 0) {
            uint32_t c = 0; unsigned long long a = 1ull, b = len;
            for (int w = 0; w < kCkThreads / 32; w++) { c ^= s.red_crc[w]; a += s.red_s1[w]; b += s.red_s2[w]; }
            if (crcs) crcs[ti] = ~c;
            if (adlers) adlers[ti] = (uint32_t)(a % kAdlerBase) | ((uint32_t)(b % kAdlerBase) << 16);
        }
    }
}

// ---------------------------------------------------------------- folds (single CTA; 8 B per tile of input)
// crc = crc32_combine(...combine(combine(init, c_0, l_0), c_1, l_1)..., c_{k-1}, l_{k-1})
//     = init * x^(8n)  ^  XOR_i c_i * x^(8 * bytes after tile i)          (crc32_braid_comb.c:16-18)
__global__ void __launch_bounds__(1024)
crc32_fold_kernel(const uint32_t* __restrict__ crcs, uint32_t ntiles, uint32_t tile_bytes, size_t n, uint32_t init,
                  uint32_t* __restrict__ result) {
    __shared__ uint32_t x2n[32];
    __shared__ uint32_t red[32];
    if (threadIdx.x == 0) build_x2n(x2n);
    __syncthreads();
    uint32_t acc = 0;
    for (uint32_t i = threadIdx.x; i < ntiles; i += blockDim.x) {
        const size_t end = min((size_t)(i + 1u) * tile_bytes, n);
        const size_t after = n - end;
        const uint32_t c = crcs[i];
        acc ^= (after && c) ? multmodp(x2nmodp(x2n, after, 3), c) : c;
    }
    if (threadIdx.x == 0 && init && n) acc ^= multmodp(x2nmodp(x2n, n, 3), init);
    if (threadIdx.x == 0 && init && !n) acc ^= init;
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

cudaError_t launch_checksum_tiles(const uint8_t* in, size_t n, uint32_t tile_bytes, uint32_t ntiles,
                                  uint32_t* crcs, uint32_t* adlers, int num_sms, cudaStream_t stream) {
    if (ntiles == 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(checksum_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CkSmem));
    if (e != cudaSuccess) return e;
    uint32_t grid = (uint32_t)num_sms * 3u;
    if (grid > ntiles) grid = ntiles;
    checksum_tiles_kernel<<<grid, kCkThreads, sizeof(CkSmem), stream>>>(in, n, tile_bytes, ntiles, crcs, adlers);
    return cudaGetLastError();
}

cudaError_t launch_crc32_fold(const uint32_t* crcs, uint32_t ntiles, uint32_t tile_bytes, size_t n, uint32_t init,
                              uint32_t* result, cudaStream_t stream) {
    crc32_fold_kernel<<<1, 1024, 0, stream>>>(crcs, ntiles, tile_bytes, n, init, result);
    return cudaGetLastError();
}

cudaError_t launch_adler32_fold(const uint32_t* adlers, uint32_t ntiles, uint32_t tile_bytes, size_t n, uint32_t init,
                                uint32_t* result, cudaStream_t stream) {
    adler32_fold_kernel<<<1, 1024, 0, stream>>>(adlers, ntiles, tile_bytes, n, init, result);
    return cudaGetLastError();
}

}  // namespace zb
