// ops.cu -- the reference's operator surface (functable.h:26-42 + deflate.h:121-131) as batched device entry points.
//
// On a GPU the per-position operators of zlib-ng are __device__ functions inside K1 / K2 / K4; these kernels expose the
// very same device code one operator at a time, so that each can be checked against its reference counterpart the way
// the reference's own unit tests do (test/test_compare256.cc, test/test_adler32.cc, ...):
//   compare256         functable.compare256         arch/generic/compare256_c.c:12-43
//   longest_match      functable.longest_match      match_tpl.h:26-280 (level-2 parameters)
//   insert_string      insert_string / quick_insert_string   insert_string_tpl.h:48-104
//   chunkmemset_safe   functable.chunkmemset_safe   chunkset_tpl.h:112-283  (out[i] = out[i - dist], byte serial)
// crc32 / adler32 are K3 (checksum.cu).
#include "common.cuh"
#include "kernels.h"
#include "lz_ops.cuh"

namespace zb {

// pairs[i] = 256 bytes at a + i*stride vs 256 bytes at b + i*stride  ->  out[i] = first mismatch index (256 if none)
__global__ void op_compare256_kernel(const uint8_t* a, const uint8_t* b, size_t stride, uint32_t n_pairs, uint32_t* out) {
    const unsigned lane = lane_id();
    const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_pairs) return;
    // one window that spans both operands: word-aligned base below the lower pointer
    const uint8_t* pa = a + (size_t)i * stride; const uint8_t* pb = b + (size_t)i * stride;
    const uint8_t* lo = pa < pb ? pa : pb;
    VWindow W;
    W.skew = (uint32_t)(reinterpret_cast<uintptr_t>(lo) & 3u);
    W.w = reinterpret_cast<const uint32_t*>(lo - W.skew);
    W.tail = W.w; W.tw0 = 0xffffffffu;
    const uint32_t r = vwarp_compare256(W, (uint32_t)(pa - lo) + W.skew, (uint32_t)(pb - lo) + W.skew, lane);
    if (lane == 0) out[i] = r;
}

// queries q: position pos[q] with hash head cand[q] in a window of n bytes (+ >= 272 readable bytes after it) and its
// prev[] table -> len[q] (0 when longest_match returns less than 4 ... the reference's caller discards those), start[q]
template <int LEVEL>
__global__ void op_longest_match_kernel(const uint8_t* window, uint32_t n, const uint16_t* prev, const uint32_t* pos, const uint32_t* cand,
                                        uint32_t n_q, uint32_t* len, uint32_t* start, bool raw) {
    constexpr uint32_t kCmp = LmParams<LEVEL>::kCmp;
    const unsigned lane = lane_id();
    const uint32_t base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u;
    if (base >= n_q) return;
    VWindow W;
    W.skew = (uint32_t)(reinterpret_cast<uintptr_t>(window) & 3u);
    W.w = reinterpret_cast<const uint32_t*>(window - W.skew);
    W.tail = W.w; W.tw0 = 0xffffffffu;
    const uint32_t qi = base + lane;
    uint32_t q = 0, c0 = 0, ml = 0, mc = 0, rb = 2;
    if (qi < n_q) {
        q = pos[qi]; c0 = cand[qi];
        uint32_t v; uint64_t x;
        load12(W, q, v, x);
        const uint32_t z = kCmp > 12u ? load32(W, q + 12u) : 0u;
        if (q + kWantMin <= n && c0 != 0u && (q - c0 - 1u) < kMaxDist) ml = longest_match_lane<LEVEL>(W, q, v, x, z, c0, n - q, prev, mc, &rb);
    }
    unsigned L = __ballot_sync(ZB_FULL, ml >= kCmp);
    while (L) {                                              // long matches: measured by the whole warp, as in K2
        const unsigned j = (unsigned)(__ffs(L) - 1); L &= L - 1u;
        const uint32_t qj = __shfl_sync(ZB_FULL, q, j), cj = __shfl_sync(ZB_FULL, mc, j);
        uint32_t l = kCmp + vwarp_compare256(W, qj + kCmp + W.skew, cj + kCmp + W.skew, lane);
        l = min(min(l, kMaxMatch), n - qj);
        if (lane == j) ml = l;
    }
    if (qi < n_q) { len[qi] = raw ? (ml ? ml : rb) : ml; start[qi] = mc; }
}

// insert_string(str, count) on head[65536] / prev[32768] (one warp; the serial insert order is reproduced with the same
// nearest-lower-peer rule the parsers use)
__global__ void op_insert_string_kernel(const uint8_t* window, uint16_t* head, uint16_t* prev, uint32_t str, uint32_t count, uint32_t* old_head) {
    const unsigned lane = lane_id();
    for (uint32_t p = str; p < str + count; p += 32u) {
        const uint32_t q = p + lane;
        const bool on = q < str + count;
        uint32_t v = 0;
        if (on) v = (uint32_t)window[q] | ((uint32_t)window[q + 1] << 8) | ((uint32_t)window[q + 2] << 16) | ((uint32_t)window[q + 3] << 24);
        const uint32_t h = hash4(v);
        const unsigned peers = __match_any_sync(ZB_FULL, on ? h : (0x10000u + lane));
        const unsigned I = __ballot_sync(ZB_FULL, on);
        if (old_head && q == str) *old_head = (uint32_t)__ldcg(head + h);     // quick_insert_string returns the head it replaced
        insert_lanes(head, prev, h, q, p, on ? (uint32_t)__ldcg(head + h) : 0u, peers, I, lane);      // the device function K2 runs
        __syncwarp();
    }
}

// out[pos .. pos+len) = byte-serial copy from dist back (one warp; the doubling wave copy of K4)
__global__ void op_chunkmemset_kernel(uint8_t* out, uint32_t pos, uint32_t dist, uint32_t len) {
    wave_copy(out, pos, dist, len, lane_id());               // the device function K4 runs
}

cudaError_t launch_op_compare256(const uint8_t* a, const uint8_t* b, size_t stride, uint32_t n_pairs, uint32_t* out, cudaStream_t s) {
    if (!n_pairs) return cudaSuccess;
    op_compare256_kernel<<<(n_pairs + 3u) / 4u, 128, 0, s>>>(a, b, stride, n_pairs, out);
    return cudaGetLastError();
}
cudaError_t launch_op_longest_match(const uint8_t* window, uint32_t n, const uint16_t* prev, const uint32_t* pos, const uint32_t* cand,
                                    uint32_t n_q, uint32_t* len, uint32_t* start, cudaStream_t s, int level, bool raw) {
    if (!n_q) return cudaSuccess;
    const unsigned g = (n_q + 127u) / 128u;
    switch (level) {
        case 2: op_longest_match_kernel<2><<<g, 128, 0, s>>>(window, n, prev, pos, cand, n_q, len, start, raw); break;
        case 3: op_longest_match_kernel<3><<<g, 128, 0, s>>>(window, n, prev, pos, cand, n_q, len, start, raw); break;
        case 4: op_longest_match_kernel<4><<<g, 128, 0, s>>>(window, n, prev, pos, cand, n_q, len, start, raw); break;
        case 5: op_longest_match_kernel<5><<<g, 128, 0, s>>>(window, n, prev, pos, cand, n_q, len, start, raw); break;
        case 6: op_longest_match_kernel<6><<<g, 128, 0, s>>>(window, n, prev, pos, cand, n_q, len, start, raw); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
cudaError_t launch_op_insert_string(const uint8_t* window, uint16_t* head, uint16_t* prev, uint32_t str, uint32_t count, cudaStream_t s, uint32_t* old_head) {
    if (!count) return cudaSuccess;
    op_insert_string_kernel<<<1, 32, 0, s>>>(window, head, prev, str, count, old_head);
    return cudaGetLastError();
}

// functable.slide_hash (arch/generic/slide_hash_c.c:15-52): every entry of head[65536] and prev[wsize] is rebased by wsize,
// entries below it become 0 (NIL).  8 entries per thread as one 16-byte vector, saturating subtract per 16-bit lane.
__global__ void op_slide_hash_kernel(uint4* head, uint32_t head_vecs, uint4* prev, uint32_t prev_vecs, uint32_t wsize) {
    const uint32_t w2 = wsize | (wsize << 16);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < head_vecs + prev_vecs; i += gridDim.x * blockDim.x) {
        uint4* p = i < head_vecs ? head + i : prev + (i - head_vecs);
        uint4 v = *p;
        v.x = __vsubus2(v.x, w2); v.y = __vsubus2(v.y, w2); v.z = __vsubus2(v.z, w2); v.w = __vsubus2(v.w, w2);
        *p = v;
    }
}
cudaError_t launch_op_slide_hash(uint16_t* head, uint16_t* prev, uint32_t wsize, cudaStream_t s) {
    op_slide_hash_kernel<<<48, 256, 0, s>>>(reinterpret_cast<uint4*>(head), 65536u / 8u, reinterpret_cast<uint4*>(prev), wsize / 8u, wsize);
    return cudaGetLastError();
}
cudaError_t launch_op_chunkmemset(uint8_t* out, uint32_t pos, uint32_t dist, uint32_t len, cudaStream_t s) {
    if (!len) return cudaSuccess;
    op_chunkmemset_kernel<<<1, 32, 0, s>>>(out, pos, dist, len);
    return cudaGetLastError();
}

}  // namespace zb
