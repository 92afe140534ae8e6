// pv_fft_batch.cu -- stand-alone batched complex FFT.
//
// The reference carries three hand-written GPU FFTs (karnel/hpfft.cu: Stockham radix-2 with one launch per
// stage :145-203, a shared-memory variant :104-143, an O(N^2) DFT :35-68) and a cuFFT wrapper
// (karnel/cufft_.cu:19-26) which its milestone deck times against each other on single transforms of
// 32..1024 points.  This is the sm_100a counterpart: ONE launch per batch, Stockham autosort with radix-16
// butterflies in registers (pv_fft_regs.cuh), 16 points per thread, the first pass reading global memory and
// the last pass writing it, so a transform crosses HBM exactly once each way (16*n bytes) and shared memory
// ceil(log16 n)-1 times, in ONE padded buffer (middle passes run in place: load, barrier, store).
// The transform length is a template parameter: every shared-memory offset is an immediate.  Unnormalised in
// both directions, forward kernel e^{-j...} (cuFFT's convention, which karnel/kernel.cu:324-326,363-368 rely on).
// Scalar fp32 butterflies in this translation unit: the stand-alone transforms are HBM-bound, not issue-bound: packed FADD2/FFMA2 only add latency here (n = 256 batch: 96 us packed, 83 us scalar = 99 % of the HBM peak; one transform in a graph 1.83 vs 1.58 us).
#ifndef PV_FFT_PACKED
#define PV_NO_PACKED 1
#endif
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "pv_fft_smem.cuh"
#include "pv_internal.h"

namespace {

using namespace pvsmem;

// grid: ceil(batch / G) CTAs, G transforms each; blockDim = a multiple of 32 >= G*n/16
template <int LG_N, int DIR>
__global__ void __launch_bounds__(512)
fft_batch_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int G, long long batch,
                 const float2 *__restrict__ twp)
{
    extern __shared__ float2 sm[];
    const long long t0 = (long long)blockIdx.x * G;
    const int g_here = (int)(batch - t0 < G ? batch - t0 : G);
    const int tot16 = (g_here << LG_N) >> 4;
    const FullTw tw{twp};
    const auto tw2 = load_reg_tw<LG_N, SecondRadix<LG_N>::v, DIR>(tw);
    NoTw none;
    stockham_pass<LG_N, 16, 0, DIR, true, false, false>(in + (t0 << LG_N), sm, tot16, tw, none);
    __syncthreads();
    remaining_passes<LG_N, DIR>(sm, out + (t0 << LG_N), tot16, tw, tw2);
}

// n <= 16: a single butterfly per transform, registers only
template <int DIR>
__global__ void fft_tiny_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int lg_n, long long batch)
{
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= batch) return;
    const float2 *s = in + (g << lg_n);
    float2 *d = out + (g << lg_n);
    if (lg_n == 4) { float2 v[16]; for (int r = 0; r < 16; r++) v[r] = s[r]; dft<16, DIR>(v); for (int r = 0; r < 16; r++) d[r] = v[r]; }
    else if (lg_n == 3) { float2 v[8]; for (int r = 0; r < 8; r++) v[r] = s[r]; dft<8, DIR>(v); for (int r = 0; r < 8; r++) d[r] = v[r]; }
    else if (lg_n == 2) { float2 v[4]; for (int r = 0; r < 4; r++) v[r] = s[r]; dft<4, DIR>(v); for (int r = 0; r < 4; r++) d[r] = v[r]; }
    else if (lg_n == 1) { float2 v[2] = {s[0], s[1]}; dft<2, DIR>(v); d[0] = v[0]; d[1] = v[1]; }
    else d[0] = s[0];
}

// ---- bulk asynchronous copy of one tile (cp.async.bulk + mbarrier, the TMA engine's 1-D form) ----
__device__ __forceinline__ void mbar_init(unsigned long long *mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, unsigned long long *mbar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst), m = (unsigned)__cvta_generic_to_shared(mbar);
    // the staging buffer was last read through the generic proxy: order those reads before the async write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes),
                 "r"(m) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *mbar, unsigned parity)
{
    const unsigned m = (unsigned)__cvta_generic_to_shared(mbar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(m), "r"(parity) : "memory");
}

// Persistent variant for large batches of n >= 512 (three or more passes; measured faster there, slower below): every CTA walks
// tiles of G transforms and keeps the NEXT tile's input in flight while it transforms the current one, so HBM reads never wait
// for the butterflies.  The staging buffer is free again as soon as the first pass has moved the tile into registers.
// Two measured choices per transform length (same-box A/B, 256 MB in + 256 MB out, us per launch;
// "bulk" = ONE cp.async.bulk per tile completed on an mbarrier instead of a 16-byte cp.async loop over the threads,
// "ct" = full tiles run passes with compile-time butterfly counts, loops unrolled):
//      n      neither   bulk    ct     both    cuFFT
//     512      91.7     92.6   101.6  102.5     84
//    1024      91.2     94.1   101.6  102.6     84
//    2048      92.2     97.0    97.0  102.8     93
//    4096     113.1     93.9   109.0   97.2     98
//    8192     146.2    143.9   134.9  128.7    110
// A 16 KB tile (n <= 2048) is too small for the bulk engine to beat 128 threads issuing 16-byte copies, and unrolling the last
// pass of several small butterflies per thread only lengthens the dependency chains; 32 / 64 KB tiles are the other way round.
template <int LG_N, int DIR, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
fft_batch_pipelined_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int G, long long batch,
                           const float2 *__restrict__ twp)
{
    constexpr bool BULK = LG_N >= 12, CT = LG_N >= 13;
    extern __shared__ float2 sm[];
    const int tile_pts = G << LG_N;
    float2 *stage = sm, *work = sm + tile_pts;
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(work + ((size_t)tile_pts * 17 / 16 + 2));
    const long long n_tiles = (batch + G - 1) / G;
    const FullTw tw{twp};
    const auto tw2 = load_reg_tw<LG_N, SecondRadix<LG_N>::v, DIR>(tw);
    if (BULK && threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    auto prefetch = [&](long long t) {
        const long long left = batch - t * G;
        const int pts = (int)(left < G ? left : G) << LG_N;
        if constexpr (BULK) {
            if (threadIdx.x == 0) bulk_load(stage, in + t * tile_pts, (unsigned)pts * (unsigned)sizeof(float2), mbar);
        } else {
            const float2 *src = in + t * tile_pts;
            for (int i = 2 * threadIdx.x; i < pts; i += 2 * blockDim.x) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(stage + i);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(src + i) : "memory");
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
    };
    long long tile = blockIdx.x;
    unsigned phase = 0;
    if (tile < n_tiles) prefetch(tile);
    for (; tile < n_tiles; tile += gridDim.x) {
        const long long left = batch - tile * G;
        const int g_here = (int)(left < G ? left : G);
        const int tot16 = (g_here << LG_N) >> 4;
        if constexpr (BULK) {
            mbar_wait(mbar, phase);                   // the tile has landed (every waiting thread sees it)
            phase ^= 1u;
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncthreads();                              // the previous tile's last pass is done with `work`
        NoTw none;
        if (CT && g_here == G) {
            stockham_pass_ct<LG_N, 16, 0, DIR, true, false, false, BLOCK, BLOCK>(stage, work, tot16, tw, none);
            __syncthreads();
            if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x);
            remaining_passes_ct<LG_N, DIR, BLOCK, BLOCK>(work, out + tile * tile_pts, tot16, tw, tw2);
        } else {                                      // the batch's last, partial tile
            stockham_pass<LG_N, 16, 0, DIR, true, false, false>(stage, work, tot16, tw, none);
            __syncthreads();
            if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x);
            remaining_passes<LG_N, DIR>(work, out + tile * tile_pts, tot16, tw, tw2);
        }
    }
}

int g_sms = 0;

template <int LG_N, int DIR>
cudaError_t launch_n(const float2 *in, float2 *out, int64_t batch, const float2 *tw, cudaStream_t st)
{
    constexpr int n = 1 << LG_N;
    int G = n >= 2048 ? 1 : 2048 / n, threads = n >= 2048 ? n / 16 : 128;
    if (batch < G) {                                  // a small batch: do not launch idle threads
        G = (int)batch;
        threads = std::max(32, ((G * n) / 16 + 31) & ~31);
    }
    const size_t work = sizeof(float2) * (size_t)(((size_t)G * n) * 17 / 16 + 2);
    const int64_t n_tiles = (batch + G - 1) / G;
    const bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    if (!g_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    cudaError_t e;
    const char *force = getenv("PV_FFT_PIPELINE");    // tuning knob: "0" never, "1" whenever aligned
    constexpr int PIPE_BLOCK = n >= 2048 ? n / 16 : 128;
    const bool pipelined = aligned && threads == PIPE_BLOCK && (force ? force[0] == '1' : (LG_N >= 9 && n_tiles > 4 * (int64_t)g_sms));
    if (pipelined) {
        constexpr int BLOCK = PIPE_BLOCK;                       // == threads: one radix-16 butterfly per thread and full tile
        const size_t smem = work + sizeof(float2) * (size_t)G * n + 16;       // + the tile's mbarrier
        // attribute and occupancy are properties of this instantiation and launch shape: query once, not per call
        static int per_sm = 0, per_sm_threads = 0;
        static size_t per_sm_smem = 0;
        if (per_sm == 0 || per_sm_threads != threads || per_sm_smem != smem) {
            e = cudaFuncSetAttribute(fft_batch_pipelined_kernel<LG_N, DIR, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_batch_pipelined_kernel<LG_N, DIR, BLOCK>, threads, smem);
            if (e != cudaSuccess) return e;
            per_sm_threads = threads;
            per_sm_smem = smem;
        }
        const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, (int64_t)g_sms * std::max(1, per_sm));
        fft_batch_pipelined_kernel<LG_N, DIR, BLOCK><<<grid, threads, smem, st>>>(in, out, G, batch, tw);
        return cudaGetLastError();
    }
    static bool attr_set = false;                     // once per instantiation (single-transform calls are latency-bound)
    if (!attr_set) {
        e = cudaFuncSetAttribute(fft_batch_kernel<LG_N, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    fft_batch_kernel<LG_N, DIR><<<(unsigned)n_tiles, threads, work, st>>>(in, out, G, batch, tw);
    return cudaGetLastError();
}

template <int DIR>
cudaError_t launch_dir(const float2 *in, float2 *out, int lg_n, int64_t batch, const float2 *tw, cudaStream_t st)
{
    switch (lg_n) {
        case 5: return launch_n<5, DIR>(in, out, batch, tw, st);
        case 6: return launch_n<6, DIR>(in, out, batch, tw, st);
        case 7: return launch_n<7, DIR>(in, out, batch, tw, st);
        case 8: return launch_n<8, DIR>(in, out, batch, tw, st);
        case 9: return launch_n<9, DIR>(in, out, batch, tw, st);
        case 10: return launch_n<10, DIR>(in, out, batch, tw, st);
        case 11: return launch_n<11, DIR>(in, out, batch, tw, st);
        case 12: return launch_n<12, DIR>(in, out, batch, tw, st);
        case 13: return launch_n<13, DIR>(in, out, batch, tw, st);
        default: break;
    }
    if (lg_n < 0 || lg_n > 4) return cudaErrorInvalidValue;
    fft_tiny_kernel<DIR><<<(unsigned)((batch + 127) / 128), 128, 0, st>>>(in, out, lg_n, batch);
    return cudaGetLastError();
}

}  // namespace

cudaError_t pv_launch_fft_batch(const float2 *in, float2 *out, int lg_n, int64_t batch, int dir, const float2 *tw,
                                cudaStream_t st)
{
    if (batch <= 0) return cudaSuccess;
    return dir < 0 ? launch_dir<-1>(in, out, lg_n, batch, tw, st) : launch_dir<1>(in, out, lg_n, batch, tw, st);
}
