// MUFU throughput on this GPU: tanh.approx.f32 / ex2.approx.f32 / rcp.approx.f32 / tanh.approx.f16x2 / tanh.approx.bf16x2,
// results per clock per SM (all SMs busy, 32 warps per SM, 8 independent chains per thread).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int OP>
__global__ void k(float* out, int iters) {
    float a[8];
    uint32_t h[8];
    for (int i = 0; i < 8; ++i) { a[i] = 0.001f * (threadIdx.x + i); h[i] = 0x3c003800u + threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
            if (OP == 4) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h[i]));
            if (OP == 5) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    const char* names[6] = {"tanh.approx.f32", "ex2.approx.f32", "rcp.approx.f32", "tanh.approx.f16x2", "tanh.approx.bf16x2", "ex2.approx.f16x2"};
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    for (int op = 0; op < 6; ++op) {
        const int iters = 20000;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            switch (op) {
                case 0: k<0><<<148, 1024>>>(out, iters); break;
                case 1: k<1><<<148, 1024>>>(out, iters); break;
                case 2: k<2><<<148, 1024>>>(out, iters); break;
                case 3: k<3><<<148, 1024>>>(out, iters); break;
                case 4: k<4><<<148, 1024>>>(out, iters); break;
                case 5: k<5><<<148, 1024>>>(out, iters); break;
            }
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double instr = 1024.0 * 8 * iters;  // thread-level instructions per SM
        const double clk = ms * 1e-3 * clk_khz * 1e3;
        printf("%-20s %8.3f ms  %6.2f thread-instr / clk / SM (%s results: x%d)\n", names[op], ms, instr / clk, op >= 3 ? "packed" : "scalar", op >= 3 ? 2 : 1);
    }
    printf("(clock rate attribute %d kHz; err %s)\n", clk_khz, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
