// %globaltimer against the host clock and clock64 (development): is a nanosecond a nanosecond, and is it monotonic?
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void spin(unsigned long long* out, volatile int* stop) {
    unsigned long long t0, t, prev, maxjump = 0, back = 0, n = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    long long c0 = clock64();
    prev = t0;
    while (!*stop) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t < prev) back++;
        else if (t - prev > maxjump) maxjump = t - prev;
        prev = t;
        n++;
        __nanosleep(64);
    }
    out[0] = prev - t0; out[1] = clock64() - c0; out[2] = maxjump; out[3] = back; out[4] = n;
}
int main() {
    unsigned long long* d; int* stop; 
    cudaMalloc(&d, 64); cudaHostAlloc(&stop, 4, cudaHostAllocMapped); *stop = 0;
    int* dstop; cudaHostGetDevicePointer(&dstop, stop, 0);
    auto a = std::chrono::steady_clock::now();
    spin<<<1, 1>>>(d, dstop);
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count() < 0.5) {}
    *stop = 1;
    cudaDeviceSynchronize();
    double host = std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
    unsigned long long h[5]; cudaMemcpy(h, d, 40, cudaMemcpyDeviceToHost);
    printf("host %.3f s  globaltimer delta %llu  clock64 delta %llu  max jump %llu  backwards %llu  reads %llu\n", host, h[0], h[1], h[2], h[3], h[4]);
    return 0;
}
