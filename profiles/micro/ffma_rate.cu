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
// sm_100 packed fp32: one FFMA2 instruction = two FMAs per lane (fma.rn.f32x2)
__global__ void k_ffma2(float *out, int iters, float a0, float b0)
{
    float2 x[8];
    const float2 a = make_float2(a0 + threadIdx.x, a0 - threadIdx.x), b = make_float2(b0, -b0);
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __ffma2_rn(x[i], a, b);
    }
    float s = 0; for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// issue-slot sharing: NF FMA-class instructions interleaved with NI integer instructions per iteration
template <bool PACKED>
__global__ void k_mix(float *out, int iters, float a0, float b0, int m0)
{
    float2 x[8]; int y[8];
    const float2 a = make_float2(a0 + threadIdx.x, a0 - threadIdx.x), b = make_float2(b0, -b0);
    for (int i = 0; i < 8; i++) { x[i] = make_float2(threadIdx.x + i, threadIdx.x - i); y[i] = threadIdx.x * 7 + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (PACKED) x[i] = __ffma2_rn(x[i], a, b);
            else { x[i].x = fmaf(x[i].x, a.x, b.x); x[i].y = fmaf(x[i].y, a.y, b.y); }
            y[i] = (y[i] ^ m0) + (y[i] >> 3);            // LOP3 + SHF/IADD3: two integer-pipe instructions
        }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += x[i].x + x[i].y + y[i];
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
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0); k_ffma2<<<blocks, thr>>>(out, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA2      : %.2f TFMA/s  (%.3f warp-FFMA2/clk/SM at 1.965 GHz)\n", (double)blocks * thr * iters * 16 / ms / 1e9,
               (double)blocks * thr / 32 * iters * 8 / (ms * 1e-3 * 1.965e9 * 148));
        cudaEventRecord(e0); k_mix<false><<<blocks, thr>>>(out, iters, 1.0001f, 0.5f, 12345); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("mix scalar : 16 FFMA + 16 int per iter: %.3f ms  (%.3f warp-inst/clk/SM)\n", ms,
               (double)blocks * thr / 32 * iters * 32 / (ms * 1e-3 * 1.965e9 * 148));
        cudaEventRecord(e0); k_mix<true><<<blocks, thr>>>(out, iters, 1.0001f, 0.5f, 12345); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("mix packed : 8 FFMA2 + 16 int per iter: %.3f ms  (%.3f warp-inst/clk/SM)\n", ms,
               (double)blocks * thr / 32 * iters * 24 / (ms * 1e-3 * 1.965e9 * 148));
    }
    for (int stride = 1; stride <= 64; stride *= 8) {
        cudaEventRecord(e0); k_red<<<blocks, thr>>>(out, 2000, stride); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("REDG.ADD.F32 stride %d: %.2f G atomics/s\n", stride, (double)blocks * thr * 2000 / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
