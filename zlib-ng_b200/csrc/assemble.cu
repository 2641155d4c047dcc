// assemble.cu -- turn per-chunk output slots into one contiguous raw-deflate stream.
//
// Every chunk's output ends on a byte boundary (it ends with the empty stored block that
// deflate() appends for Z_FULL_FLUSH, deflate.c:1064-1065, or with bi_windup for Z_FINISH), so
// the stream is the plain concatenation of the chunk outputs in order.  Two steps:
//   offsets_kernel   exclusive prefix sum of the u32 chunk sizes -> u64 byte offsets (+ total)
//   gather_kernel    chunk i's bytes -> dst + base + offsets[i], coalesced 4-byte stores on the
//                    destination's alignment, source words funnel-shifted
// In a multi-GPU run `base` and the sizes of the other ranks come from the allgather of
// (size, crc32) pairs; each rank gathers its own chunks only.
#include "common.cuh"
#include "kernels.h"

namespace zb {

__global__ void __launch_bounds__(1024)
offsets_kernel(const uint32_t* __restrict__ sizes, uint32_t n, uint64_t base, uint64_t* __restrict__ offsets) {
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t carry;
    const unsigned tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    if (tid == 0) carry = base;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n; b0 += 1024u) {
        const uint32_t i = b0 + tid;
        const uint64_t v = i < n ? (uint64_t)sizes[i] : 0ull;
        uint64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(ZB_FULL, incl, d); if ((int)lane >= d) incl += t; }
        if (lane == 31u) warp_tot[w] = incl;
        __syncthreads();
        if (w == 0) {
            uint64_t t = warp_tot[lane], ti = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint64_t u = __shfl_up_sync(ZB_FULL, ti, d); if ((int)lane >= d) ti += u; }
            warp_tot[lane] = ti - t;                     // exclusive over warps
        }
        __syncthreads();
        const uint64_t c = carry;
        if (i < n) offsets[i] = c + warp_tot[w] + incl - v;
        __syncthreads();
        if (tid == 1023u) carry = c + warp_tot[w] + incl;
        __syncthreads();
    }
    if (tid == 0) offsets[n] = carry;
}

__global__ void __launch_bounds__(256)
gather_kernel(const uint8_t* __restrict__ slots, size_t stride, const uint32_t* __restrict__ sizes,
              const uint64_t* __restrict__ offsets, uint32_t nchunks, uint8_t* __restrict__ dst) {
    for (uint32_t ci = blockIdx.x; ci < nchunks; ci += gridDim.x) {
        const uint8_t* src = slots + (size_t)ci * stride;           // 16-byte aligned
        const uint32_t size = sizes[ci];
        uint8_t* d = dst + offsets[ci];
        const uint32_t head = min(size, (uint32_t)((4u - (uint32_t)(reinterpret_cast<uintptr_t>(d) & 3u)) & 3u));
        const uint32_t nwords = (size - head) >> 2;
        const uint32_t tail0 = head + (nwords << 2);
        if (threadIdx.x < head) d[threadIdx.x] = src[threadIdx.x];
        if (threadIdx.x >= 32u && threadIdx.x - 32u < size - tail0) d[tail0 + threadIdx.x - 32u] = src[tail0 + threadIdx.x - 32u];
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(src);
        uint32_t* dw = reinterpret_cast<uint32_t*>(d + head);
        const uint32_t sh = head << 3;                               // head in 0..3
        for (uint32_t j = threadIdx.x; j < nwords; j += blockDim.x) {
            const uint32_t a = __ldg(sw + j);
            const uint32_t b = sh ? __ldg(sw + j + 1) : 0u;          // at most 4 bytes past `size`, inside the slot
            dw[j] = __funnelshift_r(a, b, sh);
        }
    }
}

cudaError_t launch_offsets(const uint32_t* sizes, uint32_t n, uint64_t base, uint64_t* offsets, cudaStream_t stream) {
    offsets_kernel<<<1, 1024, 0, stream>>>(sizes, n, base, offsets);
    return cudaGetLastError();
}

cudaError_t launch_gather(const uint8_t* slots, size_t stride, const uint32_t* sizes, const uint64_t* offsets,
                          uint32_t nchunks, uint8_t* dst, int num_sms, cudaStream_t stream) {
    if (nchunks == 0) return cudaSuccess;
    uint32_t grid = (uint32_t)num_sms * 8u;
    if (grid > nchunks) grid = nchunks;
    gather_kernel<<<grid, 256, 0, stream>>>(slots, stride, sizes, offsets, nchunks, dst);
    return cudaGetLastError();
}

}  // namespace zb
