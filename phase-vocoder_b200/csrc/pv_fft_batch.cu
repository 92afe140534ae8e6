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
#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "pv_fft_regs.cuh"
#include "pv_internal.h"

namespace {

using namespace pvfft;

// one float2 of padding every 16 keeps the stride-16 writes of the first pass conflict free
__device__ __forceinline__ int pad(int i) { return i + (i >> 4); }

template <int R>
struct Log2 { static constexpr int v = R == 2 ? 1 : R == 4 ? 2 : R == 8 ? 3 : 4; };

// Twiddles of the pass with Ns = 16 (always the second one): k = idx mod 16 = threadIdx mod 16 for every butterfly
// of a thread (blockDim is a multiple of 16), so w^r is loop invariant and lives in registers.
template <int R>
struct RegTw { float2 w[R]; };

template <int LG_N, int R, int DIR>
__device__ __forceinline__ RegTw<R> load_reg_tw(const float2 *__restrict__ tw)
{
    constexpr int n = 1 << LG_N, LR = Log2<R>::v;
    RegTw<R> t;
    const int kb = (threadIdx.x & 15) << (LG_N - 4 - LR);
    t.w[0] = make_float2(1.f, 0.f);
#pragma unroll
    for (int r = 1; r < R; r++) {
        t.w[r] = __ldg(tw + ((r * kb) & (n - 1)));
        if (DIR > 0) t.w[r].y = -t.w[r].y;
    }
    return t;
}

struct NoTw { float2 w[1]; };

// One Stockham pass of radix R over `total` = G*n/R butterflies (G transforms of n points side by side).
// Butterfly idx of a transform reads points idx + r*n/R and, with k = idx mod Ns, writes (idx-k)*R + k + r*Ns.
// SRC_LIN / DST_LIN: linear indexing (global memory or the unpadded staging buffer), else the padded work buffer.
// INPLACE: src == dst; every thread owns at most one butterfly (total <= blockDim), so one barrier between its
// loads and its stores is all the ordering the autosort permutation needs.
template <int LG_N, int R, int LG_NS, int DIR, bool SRC_LIN, bool DST_LIN, bool INPLACE, class TW>
__device__ __forceinline__ void stockham_pass(const float2 *src, float2 *dst, int total, const float2 *__restrict__ tw,
                                              const TW &rtw)
{
    constexpr int LR = Log2<R>::v, LG_PER = LG_N - LR, per = 1 << LG_PER, Ns = 1 << LG_NS;
    static_assert(SRC_LIN || per >= 16, "r*per must stay a multiple of the padding period");
    if (INPLACE && (int)threadIdx.x >= total) __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int g = i >> LG_PER, idx = i & (per - 1), k = idx & (Ns - 1);
        const int base = g << LG_N;
        float2 v[R];
        if (SRC_LIN) {
            const float2 *s = src + base + idx;
#pragma unroll
            for (int r = 0; r < R; r++) v[r] = s[r * per];
        } else {                              // r*per is a multiple of 16: the padding of the offset is a constant
            const float2 *s = src + pad(base + idx);
#pragma unroll
            for (int r = 0; r < R; r++) v[r] = s[r * per + r * per / 16];
        }
        if constexpr (LG_NS == 4) {
#pragma unroll
            for (int r = 1; r < R; r++) v[r] = cmul(v[r], rtw.w[r]);
        } else if constexpr (LG_NS > 0) {
            // w^r, w = exp(DIR * 2 pi i k / (Ns R)); up to four table loads, the other powers are products
            const int kb = k << (LG_N - LG_NS - LR);
            float2 w[R];
            w[1] = __ldg(tw + kb);
            if (R > 2) w[2] = __ldg(tw + 2 * kb);
            if (R > 4) w[4] = __ldg(tw + 4 * kb);
            if (R > 8) w[8] = __ldg(tw + 8 * kb);
            if (DIR > 0) {
                w[1].y = -w[1].y;
                if (R > 2) w[2].y = -w[2].y;
                if (R > 4) w[4].y = -w[4].y;
                if (R > 8) w[8].y = -w[8].y;
            }
            if (R > 2) w[3] = cmul(w[2], w[1]);
            if (R > 4) { w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]); }
            if (R > 8) {
#pragma unroll
                for (int r = 9; r < R; r++) w[r] = cmul(w[8], w[r - 8]);
            }
#pragma unroll
            for (int r = 1; r < R; r++) v[r] = cmul(v[r], w[r]);
        }
        dft<R, DIR>(v);
        if (INPLACE) __syncthreads();
        const int j0 = base + ((idx - k) << LR) + k;
        if (DST_LIN) {
            float2 *d = dst + j0;
#pragma unroll
            for (int r = 0; r < R; r++) d[r * Ns] = v[r];
        } else if (Ns >= 16) {
            float2 *d = dst + pad(j0);
#pragma unroll
            for (int r = 0; r < R; r++) d[r * Ns + r * Ns / 16] = v[r];
        } else {                              // Ns == 1: j0 is a multiple of 16, the R outputs are one padded row
            float2 *d = dst + pad(j0);
#pragma unroll
            for (int r = 0; r < R; r++) d[r] = v[r];
        }
    }
}

template <int LG_N>
struct SecondRadix { static constexpr int v = (LG_N - 4 >= 4) ? 16 : (1 << (LG_N - 4)); };

// Passes after the first one, for n = 2^LG_N >= 32: radix 16 with Ns = 16, 256 in place while more than 16 points
// per butterfly column remain, then one pass of radix n / Ns straight to global memory.
template <int LG_N, int DIR, class TW2>
__device__ __forceinline__ void remaining_passes(float2 *work, float2 *dst, int tot16, const float2 *tw, const TW2 &tw2)
{
    NoTw none;
    constexpr int REST = LG_N - 4;            // log2 of what is left after the first pass
    if constexpr (REST <= 4) {
        stockham_pass<LG_N, (1 << REST), 4, DIR, false, true, false>(work, dst, tot16 * (16 >> REST), tw, tw2);
    } else {
        stockham_pass<LG_N, 16, 4, DIR, false, false, true>(work, work, tot16, tw, tw2);
        __syncthreads();
        if constexpr (REST <= 8) {
            stockham_pass<LG_N, (1 << (REST - 4)), 8, DIR, false, true, false>(work, dst, tot16 * (16 >> (REST - 4)), tw, none);
        } else {
            stockham_pass<LG_N, 16, 8, DIR, false, false, true>(work, work, tot16, tw, none);
            __syncthreads();
            stockham_pass<LG_N, (1 << (REST - 8)), 12, DIR, false, true, false>(work, dst, tot16 * (16 >> (REST - 8)), tw, none);
        }
    }
}

// grid: ceil(batch / G) CTAs, G transforms each; blockDim = a multiple of 32 >= G*n/16
template <int LG_N, int DIR>
__global__ void __launch_bounds__(512)
fft_batch_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int G, long long batch,
                 const float2 *__restrict__ tw)
{
    extern __shared__ float2 sm[];
    const long long t0 = (long long)blockIdx.x * G;
    const int g_here = (int)(batch - t0 < G ? batch - t0 : G);
    const int tot16 = (g_here << LG_N) >> 4;
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

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}

// Persistent variant for large batches of n >= 512 (three or more passes; measured faster there, slower below): every CTA walks tiles of G transforms and keeps the NEXT tile's input in
// flight (16-byte cp.async into a staging buffer) while it transforms the current one, so HBM reads never wait for
// the butterflies.  The staging buffer is free again as soon as the first pass has moved the tile into registers.
template <int LG_N, int DIR>
__global__ void __launch_bounds__(512)
fft_batch_pipelined_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int G, long long batch,
                           const float2 *__restrict__ tw)
{
    extern __shared__ float2 sm[];
    const int tile_pts = G << LG_N;
    float2 *stage = sm, *work = sm + tile_pts;
    const long long n_tiles = (batch + G - 1) / G;
    const auto tw2 = load_reg_tw<LG_N, SecondRadix<LG_N>::v, DIR>(tw);
    auto prefetch = [&](long long t) {
        const long long left = batch - t * G;
        const int pts = (int)(left < G ? left : G) << LG_N;
        const float2 *src = in + t * tile_pts;
        for (int i = 2 * threadIdx.x; i < pts; i += 2 * blockDim.x) cp_async16(stage + i, src + i);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    long long tile = blockIdx.x;
    if (tile < n_tiles) prefetch(tile);
    for (; tile < n_tiles; tile += gridDim.x) {
        const long long left = batch - tile * G;
        const int g_here = (int)(left < G ? left : G);
        const int tot16 = (g_here << LG_N) >> 4;
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();                              // the tile has landed; the previous tile's last pass is done with `work`
        NoTw none;
        stockham_pass<LG_N, 16, 0, DIR, true, false, false>(stage, work, tot16, tw, none);
        __syncthreads();
        if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x);
        remaining_passes<LG_N, DIR>(work, out + tile * tile_pts, tot16, tw, tw2);
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
    const bool pipelined = aligned && (force ? force[0] == '1' : (LG_N >= 9 && n_tiles > 4 * (int64_t)g_sms));
    if (pipelined) {
        const size_t smem = work + sizeof(float2) * (size_t)G * n;
        e = cudaFuncSetAttribute(fft_batch_pipelined_kernel<LG_N, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_batch_pipelined_kernel<LG_N, DIR>, threads, smem);
        if (e != cudaSuccess) return e;
        const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, (int64_t)g_sms * std::max(1, per_sm));
        fft_batch_pipelined_kernel<LG_N, DIR><<<grid, threads, smem, st>>>(in, out, G, batch, tw);
        return cudaGetLastError();
    }
    e = cudaFuncSetAttribute(fft_batch_kernel<LG_N, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return e;
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
