/* zng_functable.c -- the host-callable operator table (struct zng_b200_functable, include/zng_b200.h), shaped like the
 * reference's struct functable_s (functable.h:26-42) for the operators that have a host-buffer form.  Every entry runs
 * on the GPU through the calling thread's context; there is no CPU implementation behind any of them. */
#include "zng_host.h"
#include <cuda_runtime_api.h>
#include <string.h>

static uint32_t ft_adler32(uint32_t adler, const uint8_t *buf, size_t len) { return zng_adler32_z(adler, buf, len); }
static uint32_t ft_crc32(uint32_t crc, const uint8_t *buf, size_t len) { return zng_crc32_z(crc, buf, len); }
static uint32_t ft_chunksize(void) { return 32; }          /* bytes moved per step of the device copy: one warp */

/* functable.compare256(src0, src1): both operands are 256 readable bytes (compare256_c.c:12) */
static uint32_t ft_compare256(const uint8_t *src0, const uint8_t *src1) {
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    uint8_t *d = NULL; uint32_t r = 0;
    if (!ctx || cudaMalloc((void **)&d, 2 * 288 + 16) != cudaSuccess) return 0;
    cudaMemset(d, 0, 2 * 288 + 16);
    cudaMemcpy(d, src0, 256, cudaMemcpyHostToDevice);
    cudaMemcpy(d + 288, src1, 256, cudaMemcpyHostToDevice);
    if (zng_b200_op_compare256(ctx, d, d + 288, 0, 1, (uint32_t *)(d + 576), NULL) == ZNG_B200_OK)
        cudaMemcpy(&r, d + 576, 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return r;
}

/* functable.chunkmemset_safe(out, from, len, left): copy len bytes from `from` (= out - dist) to out, never writing past
 * out + left; returns out + len (chunkset_tpl.h:229-283) */
static uint8_t *ft_chunkmemset_safe(uint8_t *out, uint8_t *from, unsigned len, unsigned left) {
    zng_b200_ctx *ctx = zng_b200_thread_ctx();
    if (len > left) len = left;
    if (!ctx || len == 0 || from >= out) return out + len;
    size_t dist = (size_t)(out - from);
    uint8_t *d = NULL;
    if (cudaMalloc((void **)&d, dist + len) != cudaSuccess) return out + len;
    cudaMemcpy(d, from, dist, cudaMemcpyHostToDevice);
    if (zng_b200_op_chunkmemset(ctx, d, (uint32_t)dist, (uint32_t)dist, len, NULL) == ZNG_B200_OK)
        cudaMemcpy(out, d + dist, len, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return out + len;
}

static const struct zng_b200_functable table = {ft_adler32, ft_chunkmemset_safe, ft_chunksize, ft_compare256, ft_crc32};
const struct zng_b200_functable *zng_b200_functable_get(void) { return &table; }
