// exchange_shfl.cu -- microbenchmark behind DESIGN.md 4.6 ("warp-shuffle FFT: not used"): what does ONE exchange of an FFT
// pass cost when every thread hands on 16 complex values,
//   (a) through padded shared memory (16 STS.64, barrier, 16 LDS.64: a conflict-free 16 x 128 transpose -- what the fused
//       kernels do; reaches all 128 threads of the group),
//   (b) through warp shuffles (a 16 x 16 transpose inside each half-warp: four butterfly stages of lane-xor exchanges --
//       the cheapest shuffle pattern there is, and it reaches only 16 lanes: the real pass-1 -> pass-2 exchange of window
//       2048 is a 16 x 128 transpose and would still need shared memory for the other three warps)?
// Both variants keep the values live in registers and apply one packed FMA per value between exchanges so that the compiler
// cannot drop anything.  Same launch shape as the corrected kernel: 128 threads per CTA, 4 CTAs per SM resident.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_shfl exchange_shfl.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int T = 128, R = 16, LD = 129;

__global__ void __launch_bounds__(T, 4) smem_exchange(float2 *out, int iters, float a)
{
    extern __shared__ float2 buf[];                      // 51 KB per CTA like the real kernel (4 CTAs per SM)
    float2 v[R];
#pragma unroll
    for (int i = 0; i < R; i++) v[i] = make_float2(threadIdx.x + i, i * 0.5f);
    const int k = threadIdx.x % 16, c = threadIdx.x / 16;              // reader: row k, columns n2 * 8 + c
    const float2 A = make_float2(a, a);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < R; r++) buf[r * LD + threadIdx.x] = v[r];  // a 16 x 128 transpose, rows padded to 129: conflict-free
        __syncthreads();
#pragma unroll
        for (int n2 = 0; n2 < R; n2++) v[n2] = __ffma2_rn(buf[k * LD + n2 * 8 + c], A, v[n2]);
        __syncthreads();
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < R; i++) { s.x += v[i].x; s.y += v[i].y; }
    out[blockIdx.x * T + threadIdx.x] = s;
}

__global__ void __launch_bounds__(T, 4) shfl_exchange(float2 *out, int iters, float a)
{
    float2 v[R];
#pragma unroll
    for (int i = 0; i < R; i++) v[i] = make_float2(threadIdx.x + i, i * 0.5f);
    const int lane = threadIdx.x & 15;
    const float2 A = make_float2(a, a);
    for (int it = 0; it < iters; it++) {
        // 16 x 16 transpose over 16 lanes: stage s swaps v[i] (bit s of i differs from bit s of the lane) with lane ^ (1 << s)
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const bool up = (lane >> s) & 1;
#pragma unroll
            for (int i = 0; i < R; i++) {
                if (((i >> s) & 1) == 0) {
                    const int j = i | (1 << s);
                    // lanes with the bit clear send v[j] and keep v[i]; lanes with the bit set send v[i] and keep v[j]
                    float2 snd = up ? v[i] : v[j];
                    float2 rcv;
                    rcv.x = __shfl_xor_sync(0xffffffffu, snd.x, 1 << s);
                    rcv.y = __shfl_xor_sync(0xffffffffu, snd.y, 1 << s);
                    if (up) v[i] = rcv; else v[j] = rcv;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < R; i++) v[i] = __ffma2_rn(v[i], A, make_float2(1.f, 1.f));
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < R; i++) { s.x += v[i].x; s.y += v[i].y; }
    out[blockIdx.x * T + threadIdx.x] = s;
}

template <class K>
static float run(K kern, float2 *out, int grid, int iters, size_t smem)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<grid, T, smem>>>(out, iters, 1e-6f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<grid, T, smem>>>(out, iters, 1e-6f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 4, iters = 20000;
    float2 *out;
    cudaMalloc(&out, sizeof(float2) * grid * T);
    cudaFuncSetAttribute(smem_exchange, cudaFuncAttributeMaxDynamicSharedMemorySize, 51 * 1024);
    const float a = run(smem_exchange, out, grid, iters, 51 * 1024), b = run(shfl_exchange, out, grid, iters, 0);
    if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); return 1; }
    // one "exchange" = 16 complex values per thread handed on; per SM 4 CTAs x 128 threads run concurrently
    printf("Both loops also apply 16 FFMA2 per thread and exchange.\n\n| exchange of 16 complex values per thread (4 CTAs x 128 threads per SM, %d SMs) | ns per exchange per CTA | reach |\n|---|---|---|\n", sms);
    printf("| shared memory: 16 STS.64 + barrier + 16 LDS.64 + barrier | %.1f | all 128 threads of the group |\n", a * 1e6 / iters);
    printf("| warp shuffles: 16 x 16 transpose, 4 stages x (16 SHFL + 32 selects) | %.1f | 16 lanes |\n", b * 1e6 / iters);
    return 0;
}
