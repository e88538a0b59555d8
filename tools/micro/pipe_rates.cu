// Microbenchmark: sustained lanes/clk/SM of the instructions the render inner loop can be built from.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define UNROLL 8

template <int OP>
__global__ void k(float* out, float seed, int n) {
    float f[UNROLL];
    int v[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { f[i] = seed + threadIdx.x * 0.37f + i; v[i] = threadIdx.x + i; }
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) { asm volatile("{.reg .s16 t; cvt.rzi.s16.f32 t, %1; cvt.s32.s16 %0, t;}" : "=r"(v[i]) : "f"(f[i])); f[i] = __int_as_float(__float_as_int(f[i]) ^ v[i]); }
            if (OP == 1) { asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(v[i]) : "f"(f[i])); f[i] = __int_as_float(__float_as_int(f[i]) ^ v[i]); }
            if (OP == 2) { asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(v[i]) : "f"(f[i])); f[i] = __int_as_float(__float_as_int(f[i]) ^ v[i]); }
            if (OP == 3) { asm volatile("{.reg .s16 t; mov.b32 {t, _}, %1; cvt.rn.f32.s16 %0, t;}" : "=f"(f[i]) : "r"(v[i])); v[i] ^= __float_as_int(f[i]); }
            if (OP == 4) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(v[i])); v[i] ^= __float_as_int(f[i]); }
            if (OP == 5) { float t; asm volatile("cvt.rzi.f32.f32 %0, %1;" : "=f"(t) : "f"(f[i])); f[i] = __int_as_float(__float_as_int(t) ^ v[i]); }
            if (OP == 6) { asm volatile("add.rz.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed)); }
            if (OP == 7) { asm volatile("prmt.b32 %0, %0, %1, 0x7610;" : "+r"(v[i]) : "r"(it)); }
            if (OP == 8) { asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed)); }
            if (OP == 9) { asm volatile("add.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(it)); }
            if (OP == 10) { asm volatile("lop3.b32 %0, %0, %1, 0x4b000000, 0xea;" : "+r"(v[i]) : "r"(it)); }
            if (OP == 11) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(seed)); }
            if (OP == 12) { asm volatile("{.reg .s16 t; cvt.rzi.sat.s16.f32 t, %1; cvt.s32.s16 %0, t;}" : "=r"(v[i]) : "f"(f[i])); f[i] = __int_as_float(__float_as_int(f[i]) ^ v[i]); }
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) acc += f[i] + v[i];
    if (acc == 12345.678f) out[0] = acc;
}

template <int OP>
void run(const char* name, int extra_ops) {
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int blocks = 148 * 8, threads = 256;
    k<OP><<<blocks, threads>>>(d, 1.5f, 16);
    cudaEventRecord(a);
    k<OP><<<blocks, threads>>>(d, 1.5f, ITERS);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)blocks * threads * ITERS * UNROLL;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-28s %8.3f ms  %7.2f lane-ops/clk/SM (nominal clk %d MHz; loop has %d companion op(s) per measured op)\n", name, ms, ops / cyc / 148.0, clk / 1000, extra_ops);
    cudaFree(d);
}

int main() {
    run<0>("F2I.S16.TRUNC (+xor)", 1);
    run<12>("F2I.S16.TRUNC.sat (+xor)", 1);
    run<1>("F2I.S32.TRUNC (+xor)", 1);
    run<2>("F2I.U32.TRUNC (+xor)", 1);
    run<3>("I2F.S16 (+xor)", 1);
    run<4>("I2F.S32 (+xor)", 1);
    run<5>("FRND.TRUNC (+xor)", 1);
    run<6>("FADD.RZ", 0);
    run<7>("PRMT", 0);
    run<8>("FMNMX", 0);
    run<9>("IADD", 0);
    run<10>("LOP3", 0);
    run<11>("FMUL", 0);
    return 0;
}
