// which resource keeps a small kernel from starting next to a long-running one?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 12) spin_kernel(long long clocks, int* sink) {
    extern __shared__ int dyn[];
    long long t0 = clock64();
    while (clock64() - t0 < clocks) __nanosleep(1000);
    if (sink && threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}
template <int SM> __global__ void small_kernel(int* out) {
    __shared__ int buf[SM > 0 ? SM : 1];
    if (SM > 0) { buf[threadIdx.x % SM] = threadIdx.x; __syncthreads(); }
    if (threadIdx.x == 0) out[blockIdx.x] = SM > 0 ? buf[0] : 1;
}
int main() {
    int* d; cudaMalloc(&d, 1 << 20);
    cudaStream_t a, b; cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    for (int variant = 0; variant < 6; variant++) {
        int carve = -1, dyn = 0;
        if (variant == 1) carve = 50;
        if (variant == 2) { carve = 50; dyn = 1024; }
        if (variant == 3) dyn = 1024;
        if (variant == 4) carve = 100;
        if (variant == 5) { carve = 100; dyn = 4096; }
        cudaFuncSetAttribute(spin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaDeviceSynchronize();
        cudaEventRecord(e0, a);
        spin_kernel<<<148 * 6, 128, dyn, a>>>(40000000ll, d);      // ~20 ms
        cudaEventRecord(e1, a);
        small_kernel<4096><<<148 * 2, 128, 0, b>>>(d + 16);         // 16 KiB static shared memory
        cudaEventRecord(e2, b);
        small_kernel<0><<<148 * 2, 128, 0, b>>>(d + 4096);
        cudaEvent_t e3; cudaEventCreate(&e3); cudaEventRecord(e3, b);
        cudaDeviceSynchronize();
        float t1, t2, t3; cudaEventElapsedTime(&t1, e0, e1); cudaEventElapsedTime(&t2, e0, e2); cudaEventElapsedTime(&t3, e0, e3);
        printf("variant %d (carveout %d, dyn smem %d): spin done %.2f ms, 16 KiB-smem kernel done %.2f ms, no-smem kernel after it %.2f ms\n", variant, carve, dyn, t1, t2, t3);
    }
    // and the other order: a no-smem kernel first
    cudaFuncSetAttribute(spin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, -1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, a);
    spin_kernel<<<148 * 6, 128, 0, a>>>(40000000ll, d);
    cudaEventRecord(e1, a);
    small_kernel<0><<<148 * 2, 128, 0, b>>>(d + 4096);
    cudaEventRecord(e2, b);
    cudaDeviceSynchronize();
    float t1, t2; cudaEventElapsedTime(&t1, e0, e1); cudaEventElapsedTime(&t2, e0, e2);
    printf("no-smem small kernel next to the spin kernel: spin done %.2f ms, small done %.2f ms\n", t1, t2);
    return 0;
}
