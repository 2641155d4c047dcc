#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 12) spin_kernel(long long clocks, int* sink) {
    extern __shared__ int dyn[];
    long long t0 = clock64();
    while (clock64() - t0 < clocks) __nanosleep(1000);
    if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}
__global__ void __launch_bounds__(128, 6) spin_kernel6(long long clocks, int* sink) {
    extern __shared__ int dyn[];
    long long t0 = clock64();
    while (clock64() - t0 < clocks) __nanosleep(1000);
    if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}
__global__ void small_kernel(int* out) {
    __shared__ int buf[4352];                    // 17408 bytes, like static_emit_kernel
    buf[threadIdx.x] = threadIdx.x; __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = buf[1];
}
int main() {
    int* d; cudaMalloc(&d, 1 << 20);
    cudaStream_t a, b; cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    const int dyns[] = {0, 1024, 4096, 8192, 12288, 16384, 20480, 32768};
    for (int which = 0; which < 2; which++)
    for (int dyn : dyns) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, a);
        if (which == 0) spin_kernel<<<148 * 6, 128, dyn, a>>>(40000000ll, d); else spin_kernel6<<<148 * 6, 128, dyn, a>>>(40000000ll, d);
        cudaEventRecord(e1, a);
        small_kernel<<<148 * 2, 128, 0, b>>>(d + 16);
        cudaEventRecord(e2, b);
        cudaDeviceSynchronize();
        float t1, t2; cudaEventElapsedTime(&t1, e0, e1); cudaEventElapsedTime(&t2, e0, e2);
        printf("min-blocks %d, spin kernel with %5d B dynamic smem per CTA (6 CTAs per SM): 17 KiB-smem kernel done at %.2f ms (spin %.2f)\n", which ? 6 : 12, dyn, t2, t1);
    }
    return 0;
}
