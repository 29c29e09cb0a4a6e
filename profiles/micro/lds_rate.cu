// Microbenchmark: shared-memory load cost on B200 as a function of width and of how many distinct addresses a warp asks for.
// The cell-run deposit (cellrun.cu, phase 2) reads per-particle factors with broadcast LDS.128; this measures what such a
// load costs the SM's LSU data path, to decide how many lanes should share a footprint.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_rate lds_rate.cu ;  run: ./lds_rate
// Output: cycles per warp-level load instruction per SM (32 warps per SM keep every SMSP busy; loads are independent).
#include <cstdio>
#include <cuda_runtime.h>

// group = lane / GROUP_LANES reads its own 16-byte (or 8 / 4) slot; slots of different groups are `gstride` floats apart
template <int WIDTH>                                     // floats per lane per load: 1, 2, 4
__global__ void __launch_bounds__(1024) k_lds(float *out, int iters, const int *lane_off, long long *cyc, int zero)
{
    __shared__ __align__(16) float sm[8][1500];
    const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 7;
    for (int i = threadIdx.x; i < 8 * 1500; i += blockDim.x) (&sm[0][0])[i] = i * 0.001f;
    __syncthreads();
    const float *p = &sm[warp][lane_off[lane]];
    int acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        p += acc[it & 7] & zero;                             // opaque dependence on earlier loads, once per 8 loads
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const unsigned q = (unsigned)__cvta_generic_to_shared(p + u * 16);
            if (WIDTH == 4) { int a, b, c, d; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(q) : "memory"); acc[u] ^= a ^ b; acc[u] ^= c ^ d; }
            else if (WIDTH == 2) { int a, b; asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(q) : "memory"); acc[u] ^= a ^ b; }
            else { int a; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a) : "r"(q) : "memory"); acc[u] ^= a; }
        }
    }
    long long t1 = clock64();
    int s = 0; for (int u = 0; u < 8; u++) s ^= acc[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int WIDTH, class F>
static void run(const char *what, F off, float *out, long long *cyc)
{
    const int iters = 1 << 14;
    int h[32], *d;
    for (int l = 0; l < 32; l++) h[l] = off(l);
    cudaMalloc(&d, sizeof h); cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_lds<WIDTH><<<148, 1024>>>(out, 16, d, cyc, 0);
    cudaEventRecord(e0);
    k_lds<WIDTH><<<148, 1024>>>(out, iters, d, cyc, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return; }
    long long c = -1; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    // 8 warps x iters x 8 loads per SM
    printf("%-78s %6.2f cycles per warp-load per SM (clock64), %6.2f (events at 1.965 GHz)\n", what, (double)c / (32.0 * iters * 8),
           ms * 1e-3 * 1.965e9 / (32.0 * iters * 8));
}

int main()
{
    float *out; cudaMalloc(&out, 148 * 1024 * 4);
    long long *cyc; cudaMalloc(&cyc, 8);
    const int H = 720;                                   // half-warp staging areas are 720 floats (16 banks) apart
    run<4>("LDS.128  1 address per warp", [](int l) { return 0; }, out, cyc);
    run<4>("LDS.128  2 addresses (per half-warp), 16 banks apart", [=](int l) { return (l >> 4) * H; }, out, cyc);
    run<4>("LDS.128  2 addresses (per half-warp), same banks", [](int l) { return (l >> 4) * 704; }, out, cyc);
    run<4>("LDS.128  4 addresses (per quarter-warp), 8 banks apart", [](int l) { return (l >> 3) * 360; }, out, cyc);
    run<4>("LDS.128  8 addresses (per group of 4 lanes), 4 banks apart", [](int l) { return (l >> 2) * 100; }, out, cyc);
    run<4>("LDS.128  row j = lane&3 of the half's particle (4 x 16 B contiguous per half)", [=](int l) { return (l >> 4) * H + (l & 3) * 4; }, out, cyc);
    run<4>("LDS.128  row k = (lane>>2)&3 of the half's particle", [=](int l) { return (l >> 4) * H + ((l >> 2) & 3) * 4; }, out, cyc);
    run<4>("LDS.128  row j of the quarter's particle (4 particles per warp)", [](int l) { return (l >> 3) * 360 + (l & 3) * 4; }, out, cyc);
    run<4>("LDS.128  32 addresses contiguous (512 B)", [](int l) { return l * 4; }, out, cyc);
    run<4>("LDS.128  32 addresses, stride 44 floats (phase-1 staging stores pattern)", [](int l) { return l * 44; }, out, cyc);
    run<2>("LDS.64   1 address per warp", [](int l) { return 0; }, out, cyc);
    run<2>("LDS.64   2 addresses (per half-warp)", [=](int l) { return (l >> 4) * H; }, out, cyc);
    run<2>("LDS.64   4 addresses (per quarter-warp)", [](int l) { return (l >> 3) * 360; }, out, cyc);
    run<2>("LDS.64   32 addresses contiguous", [](int l) { return l * 2; }, out, cyc);
    run<1>("LDS.32   1 address per warp", [](int l) { return 0; }, out, cyc);
    run<1>("LDS.32   2 addresses (per half-warp)", [=](int l) { return (l >> 4) * H; }, out, cyc);
    run<1>("LDS.32   8 addresses (4 per half)", [=](int l) { return (l >> 4) * H + (l & 3) * 4; }, out, cyc);
    run<1>("LDS.32   32 addresses contiguous", [](int l) { return l; }, out, cyc);
    run<1>("LDS.32   32 addresses stride 44 floats (4-way conflict)", [](int l) { return l * 44; }, out, cyc);
    return 0;
}
