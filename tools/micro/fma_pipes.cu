// fma_pipes.cu -- microbenchmark: fp32 throughput of scalar FFMA, packed FFMA2 and mixes of both on one B200 SM
// partition (how many issue slots / pipe cycles does a packed op cost?).  Build: nvcc -arch=sm_100a -O3 -o fma_pipes fma_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>   // 0: scalar FFMA x 16 chains; 1: FFMA2 x 8 chains (same flops); 2: half packed, half scalar; 3: FFMA2 x 16 chains (2x flops)
__global__ void k(float *out, int iters, float a, float b)
{
    float2 v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) { if (i < 8) { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b); } }
            else if (MODE == 1) { if (i < 8) v[i] = __ffma2_rn(v[i], A, B); }
            else if (MODE == 2) { if (i < 4) v[i] = __ffma2_rn(v[i], A, B); else if (i < 8) { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b); } }
            else v[i] = __ffma2_rn(v[i], A, B);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int flops_per_iter_per_thread)
{
    float *d;
    cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<MODE><<<148, warps * 32>>>(d, iters, 1.0001f, 0.5f);
        cudaEventRecord(e0);
        k<MODE><<<148, warps * 32>>>(d, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = (double)148 * warps * 32 * iters * flops_per_iter_per_thread;
        printf("%-34s %2d warps/SM: %7.2f TFLOP/s\n", name, warps, fl / (ms * 1e-3) / 1e12);
    }
}

int main()
{
    run<0>("scalar FFMA (16 fma/iter)", 32);
    run<1>("packed FFMA2 (8 ffma2/iter)", 32);
    run<2>("4 ffma2 + 8 scalar fma per iter", 32);
    run<3>("packed FFMA2 (16 ffma2/iter)", 64);
    return 0;
}
