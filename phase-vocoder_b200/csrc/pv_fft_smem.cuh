// pv_fft_smem.cuh -- Stockham autosort passes with radix-16 register butterflies over shared memory.
//
// Shared by the stand-alone batched FFT (pv_fft_batch.cu) and the shape-generic stream kernels
// (pv_generic_kernels.cu).  The twiddle source is a functor tw(k) = exp(-2 pi i k / n), 0 <= k < n, so that each
// user brings its own table layout.
#pragma once
#include "pv_fft_regs.cuh"

namespace pvsmem {

using namespace pvfft;

// n-th roots of unity stored as a plain table of n entries
struct FullTw {
    const float2 *p;
    __device__ __forceinline__ float2 operator()(int k) const { return __ldg(p + k); }
};

// one float2 of padding every 16 keeps the stride-16 writes of the first pass conflict free
__device__ __forceinline__ int pad(int i) { return i + (i >> 4); }

template <int R>
struct Log2 { static constexpr int v = R == 2 ? 1 : R == 4 ? 2 : R == 8 ? 3 : 4; };

// Twiddles of the pass with Ns = 16 (always the second one): k = idx mod 16 = threadIdx mod 16 for every butterfly
// of a thread (blockDim is a multiple of 16), so w^r is loop invariant and lives in registers.
template <int R>
struct RegTw { float2 w[R]; };

template <int LG_N, int R, int DIR, class TWF>
__device__ __forceinline__ RegTw<R> load_reg_tw(const TWF &tw)
{
    constexpr int n = 1 << LG_N, LR = Log2<R>::v;
    RegTw<R> t;
    const int kb = (threadIdx.x & 15) << (LG_N - 4 - LR);
    t.w[0] = make_float2(1.f, 0.f);
#pragma unroll
    for (int r = 1; r < R; r++) {
        t.w[r] = tw((r * kb) & (n - 1));
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
// CT_TOTAL / CT_BLOCK > 0: `total` and blockDim.x are these compile-time constants (a full tile of the persistent batched FFT,
// CT_TOTAL a multiple of CT_BLOCK): the butterfly loop unrolls and its index arithmetic folds into immediates.
template <int LG_N, int R, int LG_NS, int DIR, bool SRC_LIN, bool DST_LIN, bool INPLACE, int CT_TOTAL, int CT_BLOCK, class TWF, class TW>
__device__ __forceinline__ void stockham_pass_ct(const float2 *src, float2 *dst, int total, const TWF &tw, const TW &rtw)
{
    constexpr int LR = Log2<R>::v, LG_PER = LG_N - LR, per = 1 << LG_PER, Ns = 1 << LG_NS;
    static_assert(SRC_LIN || per >= 16, "r*per must stay a multiple of the padding period");
    constexpr bool CT = CT_TOTAL > 0;
    static_assert(!CT || (CT_BLOCK > 0 && CT_TOTAL % CT_BLOCK == 0), "full tiles only");
    if (!CT && INPLACE && (int)threadIdx.x >= total) __syncthreads();
    const int step = CT ? CT_BLOCK : (int)blockDim.x;
    constexpr int TRIPS = CT ? CT_TOTAL / (CT_BLOCK > 0 ? CT_BLOCK : 1) : 0;
#pragma unroll(CT ? 16 : 1)
    for (int it = 0; CT ? it < TRIPS : (int)threadIdx.x + it * step < total; it++) {
        const int i = (int)threadIdx.x + it * step;
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
            w[1] = tw(kb);
            if (R > 2) w[2] = tw(2 * kb);
            if (R > 4) w[4] = tw(4 * kb);
            if (R > 8) w[8] = tw(8 * kb);
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

template <int LG_N, int R, int LG_NS, int DIR, bool SRC_LIN, bool DST_LIN, bool INPLACE, class TWF, class TW>
__device__ __forceinline__ void stockham_pass(const float2 *src, float2 *dst, int total, const TWF &tw, const TW &rtw)
{
    stockham_pass_ct<LG_N, R, LG_NS, DIR, SRC_LIN, DST_LIN, INPLACE, 0, 0>(src, dst, total, tw, rtw);
}

// A pass that hands its outputs to `sink(n, value)` (n = index within the transform, natural order after the last
// pass) instead of storing them: lets the consumer of a transform work straight from the butterfly's registers.
// Source: the padded work buffer.  Twiddles by table lookup (any Ns > 1).
template <int LG_N, int R, int LG_NS, int DIR, class TWF, class SINK>
__device__ __forceinline__ void stockham_pass_sink(const float2 *src, int total, const TWF &tw, SINK sink)
{
    constexpr int LR = Log2<R>::v, LG_PER = LG_N - LR, per = 1 << LG_PER, Ns = 1 << LG_NS;
    static_assert(per >= 16 && LG_NS > 0 && LG_NS != 4, "table-twiddle pass on the padded buffer");
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int idx = i & (per - 1), k = idx & (Ns - 1);
        float2 v[R];
        const float2 *s = src + pad(idx);
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = s[r * per + r * per / 16];
        const int kb = k << (LG_N - LG_NS - LR);
        float2 w[R];
        w[1] = tw(kb);
        if (R > 2) w[2] = tw(2 * kb);
        if (R > 4) w[4] = tw(4 * kb);
        if (R > 8) w[8] = tw(8 * kb);
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
        dft<R, DIR>(v);
        const int j0 = ((idx - k) << LR) + k;
#pragma unroll
        for (int r = 0; r < R; r++) sink(j0 + (r << LG_NS), v[r]);
    }
}

template <int LG_N>
struct SecondRadix { static constexpr int v = (LG_N - 4 >= 4) ? 16 : (1 << (LG_N - 4)); };

// Passes after the first one, for n = 2^LG_N >= 32: radix 16 with Ns = 16, 256 in place while more than 16 points
// per butterfly column remain, then one pass of radix n / Ns straight to global memory.
// TOT16 / BLOCK > 0: a full tile with a compile-time butterfly count and block size (stockham_pass_ct).
template <int LG_N, int DIR, int TOT16, int BLOCK, class TWF, class TW2>
__device__ __forceinline__ void remaining_passes_ct(float2 *work, float2 *dst, int tot16, const TWF &tw, const TW2 &tw2)
{
    NoTw none;
    constexpr int REST = LG_N - 4;            // log2 of what is left after the first pass
    if constexpr (REST <= 4) {
        stockham_pass_ct<LG_N, (1 << REST), 4, DIR, false, true, false, TOT16 * (16 >> REST), BLOCK>(work, dst, tot16 * (16 >> REST), tw, tw2);
    } else {
        stockham_pass_ct<LG_N, 16, 4, DIR, false, false, true, TOT16, BLOCK>(work, work, tot16, tw, tw2);
        __syncthreads();
        if constexpr (REST <= 8) {
            stockham_pass_ct<LG_N, (1 << (REST - 4)), 8, DIR, false, true, false, TOT16 * (16 >> (REST - 4)), BLOCK>(
                work, dst, tot16 * (16 >> (REST - 4)), tw, none);
        } else {
            stockham_pass_ct<LG_N, 16, 8, DIR, false, false, true, TOT16, BLOCK>(work, work, tot16, tw, none);
            __syncthreads();
            stockham_pass_ct<LG_N, (1 << (REST - 8)), 12, DIR, false, true, false, TOT16 * (16 >> (REST - 8)), BLOCK>(
                work, dst, tot16 * (16 >> (REST - 8)), tw, none);
        }
    }
}

template <int LG_N, int DIR, class TWF, class TW2>
__device__ __forceinline__ void remaining_passes(float2 *work, float2 *dst, int tot16, const TWF &tw, const TW2 &tw2)
{
    remaining_passes_ct<LG_N, DIR, 0, 0>(work, dst, tot16, tw, tw2);
}

}  // namespace pvsmem
