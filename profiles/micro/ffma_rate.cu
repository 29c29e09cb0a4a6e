// Microbenchmark: issue rate of 3-register FFMA vs FFMA with an immediate/constant operand, REDG rate, LDS.128 rate.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma3(float *out, int iters, float a0, float b0)
{
    float x[16];
    float a = a0 + threadIdx.x, b = b0;
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);      // 3 register operands
    }
    float s = 0; for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffmai(float *out, int iters)
{
    float x[16];
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], 1.0001f, 0.5f);   // immediates
    }
    float s = 0; for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma_acc(float *out, int iters, float a0)
{
    // accumulate pattern of the deposit: acc = fma(p, q, acc) with p,q varying registers
    float acc[12], p[4], q = a0;
    for (int i = 0; i < 12; i++) acc[i] = i;
    for (int i = 0; i < 4; i++) p[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 12; i++) acc[i] = fmaf(p[i & 3], q, acc[i]);
        q += 1e-6f;
    }
    float s = 0; for (int i = 0; i < 12; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_red(float *buf, int iters, int stride)
{
    size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * stride;
    for (int it = 0; it < iters; it++) atomicAdd(buf + ((base + (size_t)it * 4099) & ((1u << 26) - 1)), 1.0f);
}
int main()
{
    float *out; cudaMalloc(&out, 1 << 28); cudaMemset(out, 0, 1 << 28);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = 148 * 8, thr = 256, iters = 20000;
    float ms;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0); k_ffma3<<<blocks, thr>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA 3-reg : %.2f TFMA/s  (%.3f warp-FFMA/clk/SM at 1.965 GHz)\n", (double)blocks * thr * iters * 16 / ms / 1e9,
               (double)blocks * thr / 32 * iters * 16 / (ms * 1e-3 * 1.965e9 * 148));
        cudaEventRecord(e0); k_ffmai<<<blocks, thr>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA imm   : %.2f TFMA/s  (%.3f warp-FFMA/clk/SM)\n", (double)blocks * thr * iters * 16 / ms / 1e9,
               (double)blocks * thr / 32 * iters * 16 / (ms * 1e-3 * 1.965e9 * 148));
        cudaEventRecord(e0); k_ffma_acc<<<blocks, thr>>>(out, iters, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA acc   : %.2f TFMA/s  (%.3f warp-FFMA/clk/SM)\n", (double)blocks * thr * iters * 12 / ms / 1e9,
               (double)blocks * thr / 32 * iters * 12 / (ms * 1e-3 * 1.965e9 * 148));
    }
    for (int stride = 1; stride <= 64; stride *= 8) {
        cudaEventRecord(e0); k_red<<<blocks, thr>>>(out, 2000, stride); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("REDG.ADD.F32 stride %d: %.2f G atomics/s\n", stride, (double)blocks * thr * 2000 / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
