// pv_generic_kernels.cu -- shape-generic sm_100a kernels.
//
// These cover (a) the reference's per-frame entry points (analysis / resynthesis /
// test_overlap_add), which by contract read and write the full 2N-bin {mag, phase} buffers
// and are therefore HBM-bound on those buffers, and (b) a fused compat-mode stream kernel
// that accepts ANY power-of-two window and any hop pair.  The tuned fused kernels for the
// benchmarked shapes live in pv_fused_kernels.cu; pv_capi.cu picks between them.
//
// Transform layout (both kernel families): a real sequence of length 2M is packed into M
// complex points, c[n] = x[2n] + j x[2n+1]; one M-point complex FFT plus a split step gives the
// non-redundant half of the real spectrum.  Reference steps:
//   A window              karnel/kernel.cu:68-74
//   B zero-phase + pad    karnel/kernel.cu:25-32     (folded into the load index)
//   C 2N-point DFT        karnel/kernel.cu:324-326   (cuFFT in the reference; in-kernel here)
//   D {mag, atanf}        karnel/kernel.cu:101-109
//   E polar->rect (D2)    karnel/kernel.cu:121-129
//   F N-point C2R         karnel/kernel.cu:363-366
//   G /N, half swap, win  karnel/kernel.cu:130-138, 51-59, 75-81
//   H overlap-add         karnel/kernel.cu:111-119
//   I emit hop            src/main.cpp:281-295
// Scalar fp32 butterflies in this translation unit: the per-frame contract kernels move 32 KB of {mag, phase} per frame through HBM and the shape-generic stream kernels are fallbacks: scalar butterflies, as measured for the HBM-bound stand-alone FFT (pv_fft_batch.cu).
#define PV_NO_PACKED 1
#include <algorithm>
#include <cstdlib>

#include "pv_fft_smem.cuh"
#include "pv_fused_corrected.cuh"
#include "pv_internal.h"

namespace {

constexpr int kGenericThreads = 256;      // launch bound; small windows launch 128

inline int generic_threads(int N) { return N >= 2048 ? 256 : 128; }

__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Stockham radix-2 autosort FFT of M points held in shared memory (ping-pong x <-> y).
// e^{-2 pi i k / M} = tw[k * tws] with tw the 2N-th roots table.  Returns the result buffer.
__device__ float2 *fft_stockham(float2 *x, float2 *y, int M, int tws, const float2 *__restrict__ tw,
                                bool inverse)
{
    int lg_s = 0;
    for (int n = M; n > 1; n >>= 1, ++lg_s) {
        const int m = n >> 1, s = 1 << lg_s;
        for (int idx = threadIdx.x; idx < (M >> 1); idx += blockDim.x) {
            const int p = idx >> lg_s, q = idx & (s - 1);
            float2 w = tw[(p << lg_s) * tws];
            if (inverse) w.y = -w.y;
            const float2 a = x[q + s * p], b = x[q + s * (p + m)];
            y[q + s * (2 * p)] = make_float2(a.x + b.x, a.y + b.y);
            y[q + s * (2 * p + 1)] = cmul(make_float2(a.x - b.x, a.y - b.y), w);
        }
        __syncthreads();
        float2 *t = x; x = y; y = t;
    }
    return x;
}

// tw[] covers exp(-j*pi*k/N) for k < N (half a circle); k in [N, 2N) is the negated first half.
__device__ __forceinline__ float2 tw_at(const float2 *__restrict__ tw, int k, int N)
{
    float2 w = tw[k & (N - 1)];
    if (k & N) { w.x = -w.x; w.y = -w.y; }
    return w;
}

// Stockham radix-4 autosort FFT (one radix-2 stage first when log2 M is odd): half the passes, barriers and
// shared-memory traffic of the radix-2 version.  Same contract as fft_stockham.
__device__ float2 *fft_stockham4(float2 *x, float2 *y, int M, int tws, const float2 *__restrict__ tw, int N,
                                 bool inverse)
{
    int n = M, lg_s = 0;
    int lgM = 0;
    while ((1 << lgM) < M) ++lgM;
    if (lgM & 1) {                                   // one radix-2 stage
        const int m = n >> 1;
        for (int idx = threadIdx.x; idx < m; idx += blockDim.x) {
            float2 w = tw_at(tw, idx * tws, N);
            if (inverse) w.y = -w.y;
            const float2 a = x[idx], b = x[idx + m];
            y[2 * idx] = make_float2(a.x + b.x, a.y + b.y);
            y[2 * idx + 1] = cmul(make_float2(a.x - b.x, a.y - b.y), w);
        }
        __syncthreads();
        float2 *t = x; x = y; y = t;
        n = m;
        lg_s = 1;
    }
    for (; n > 1; n >>= 2, lg_s += 2) {
        const int m = n >> 2, s = 1 << lg_s;
        for (int idx = threadIdx.x; idx < (M >> 2); idx += blockDim.x) {
            const int p = idx >> lg_s, q = idx & (s - 1);
            const float2 a0 = x[q + s * p], a1 = x[q + s * (p + m)], a2 = x[q + s * (p + 2 * m)], a3 = x[q + s * (p + 3 * m)];
            const float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y), t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
            const float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y);
            const float2 d = make_float2(a1.x - a3.x, a1.y - a3.y);
            // forward: multiply (a1 - a3) by -j, inverse: by +j
            const float2 t3 = inverse ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
            const int kb = (p << lg_s) * tws;                     // W_n^p = exp(-2 pi i p s / M)
            float2 w1 = tw_at(tw, kb, N), w2 = tw_at(tw, 2 * kb, N), w3 = tw_at(tw, 3 * kb, N);
            if (inverse) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
            float2 *o = y + q + s * (4 * p);
            o[0] = make_float2(t0.x + t2.x, t0.y + t2.y);
            o[s] = cmul(make_float2(t1.x + t3.x, t1.y + t3.y), w1);
            o[2 * s] = cmul(make_float2(t0.x - t2.x, t0.y - t2.y), w2);
            o[3 * s] = cmul(make_float2(t1.x - t3.x, t1.y - t3.y), w3);
        }
        __syncthreads();
        float2 *t = x; x = y; y = t;
    }
    return x;
}

// ---- radix-16 Stockham (pv_fft_smem.cuh) for the stream kernels ----
// exp(-2 pi i k / M) out of the handle's half circle of 2N-th roots: tw[j] = exp(-j pi j / N), j < N.
struct HalfTw {
    const float2 *p;
    int stride, half;                                // stride = 2N / M, half = M / 2
    __device__ __forceinline__ float2 operator()(int k) const
    {
        const bool neg = k >= half;
        float2 w = p[(neg ? k - half : k) * stride];
        if (neg) { w.x = -w.x; w.y = -w.y; }
        return w;
    }
};

template <int LG, int DIR>
__device__ __forceinline__ void fft16_fixed(float2 *x, float2 *work, const HalfTw &tw)
{
    using namespace pvsmem;
    constexpr int tot16 = 1 << (LG - 4);
    const auto tw2 = load_reg_tw<LG, SecondRadix<LG>::v, DIR>(tw);
    NoTw none;
    stockham_pass<LG, 16, 0, DIR, true, false, false>(x, work, tot16, tw, none);
    __syncthreads();
    remaining_passes<LG, DIR>(work, x, tot16, tw, tw2);
    __syncthreads();
}

// One compiled copy per direction for all the kernels of this file (the transform works through shared memory
// only, so an out-of-line call costs nothing that matters).
template <int DIR>
__device__ __noinline__ void fft16_dispatch(float2 *x, float2 *work, int M, HalfTw tw)
{
    switch (M) {
        case 64: fft16_fixed<6, DIR>(x, work, tw); break;
        case 128: fft16_fixed<7, DIR>(x, work, tw); break;
        case 256: fft16_fixed<8, DIR>(x, work, tw); break;
        case 512: fft16_fixed<9, DIR>(x, work, tw); break;
        case 1024: fft16_fixed<10, DIR>(x, work, tw); break;
        case 2048: fft16_fixed<11, DIR>(x, work, tw); break;
        default: fft16_fixed<12, DIR>(x, work, tw); break;
    }
}

// Work-buffer elements fft_auto needs behind an M-point transform (one float2 of padding every 16).
__host__ __device__ constexpr int fft_work_elems(int M) { return M + M / 16 + 2; }

// M-point transform of x (linear) with y (>= fft_work_elems(M)) as scratch; returns the buffer holding the result.
// 64 <= M <= 4096 with at least M/16 threads: radix-16 passes, result back in x.  Else the radix-4 ping-pong.
__device__ float2 *fft_auto(float2 *x, float2 *y, int M, int tws, const float2 *__restrict__ tw, int N, bool inverse)
{
    if (M >= 64 && M <= 4096 && (M >> 4) <= (int)blockDim.x && (blockDim.x & 15) == 0) {
        const HalfTw t{tw, tws, M >> 1};
        if (inverse) fft16_dispatch<1>(x, y, M, t);
        else fft16_dispatch<-1>(x, y, M, t);
        return x;
    }
    return fft_stockham4(x, y, M, tws, tw, N, inverse);
}

// Steps A+B: window, zero-phase shift, zero pad to 2N, packed as N complex points.
__device__ void load_frame_compat(float2 *c, const float *__restrict__ in, int64_t base, int64_t n_in,
                                  const PvDev &d)
{
    const int N = d.N, q = N >> 2;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float2 v = make_float2(0.f, 0.f);
        int i = -1;
        if (n < q) i = (N >> 1) + 2 * n;
        else if (n >= 3 * q) i = 2 * (n - 3 * q);
        if (i >= 0) {
            const int64_t g = base + i;
            const float x0 = (g < n_in) ? in[g] : 0.f;
            const float x1 = (g + 1 < n_in) ? in[g + 1] : 0.f;
            v = make_float2(x0 * d.win[i], x1 * d.win[i + 1]);
        }
        c[n] = v;
    }
}

// Bin k (0 <= k < N) of the 2N-point real spectrum from the packed N-point transform C.
__device__ __forceinline__ float2 split_bin(const float2 *C, int k, const PvDev &d)
{
    const int N = d.N;
    const float2 a = C[k];
    float2 b = C[(N - k) & (N - 1)];
    b.y = -b.y;
    const float2 e = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));
    const float2 dd = make_float2(a.x - b.x, a.y - b.y);
    const float2 o = make_float2(0.5f * dd.y, -0.5f * dd.x);      // (a-b)/(2j)
    const float2 t = cmul(d.tw[k], o);
    return make_float2(e.x + t.x, e.y + t.y);
}

// Step D exactly as cudaMagFreq (kernel.cu:101-109).
__device__ __forceinline__ float2 mag_phase(float2 X, int flags)
{
    const float mag = sqrtf(X.x * X.x + X.y * X.y);
    float ph;
    if (X.x == 0.f && X.y == 0.f && !(flags & PV_FLAG_NAN_COMPAT)) ph = 0.f;
    else ph = atanf(X.y / X.x);
    return make_float2(mag, ph);
}

// Step E exactly as cudaTimeScale with timeScale = 1 (kernel.cu:121-129, defect D2).
__device__ __forceinline__ float2 polar_to_rect_d2(float2 mp)
{
    const float xr = mp.x * cosf(mp.y);
    const float xi = xr * sinf(mp.y);
    return make_float2(xr, xi);
}

// Pre-twiddle of the packed Hermitian inverse (N real outputs from N/2 complex points):
// Z[k] = (Yk + conj(Ym)) + j * w * (Yk - conj(Ym)),  w = e^{+2 pi i k / N}, Ym = Y[N/2-k].
__device__ __forceinline__ float2 herm_pack(float2 yk, float2 ym, float2 w)
{
    const float2 s = make_float2(yk.x + ym.x, yk.y - ym.y);
    const float2 df = make_float2(yk.x - ym.x, yk.y + ym.y);
    const float2 t = cmul(w, df);
    return make_float2(s.x - t.y, s.y + t.x);
}

// Steps F (pre-twiddle) for all bins: reads Y via functor, writes Z[0..N/2) into z.
template <class YF>
__device__ void build_inverse_input(float2 *z, YF Y, const PvDev &d)
{
    const int N = d.N, h = N >> 1, q = N >> 2;
    for (int k = threadIdx.x; k <= q; k += blockDim.x) {
        float2 yk = Y(k), ym = Y(h - k);
        if (k == 0) { yk.y = 0.f; ym.y = 0.f; }    // C2R ignores Im of bins 0 and N/2
        const float2 t = d.tw[2 * k];               // e^{-2 pi i k / N}
        z[k] = herm_pack(yk, ym, make_float2(t.x, -t.y));
        if (k != 0 && k != q) z[h - k] = herm_pack(ym, yk, make_float2(-t.x, -t.y));
    }
}

// Steps G+H for one frame: r = unnormalised packed inverse output, acc = OLA ring of N floats.
__device__ void ola_accumulate(float *acc, const float2 *r, int pos0, bool zero_frame, const PvDev &d,
                               float scale = 0.f)
{
    const int N = d.N, h = N >> 1, keep = N - d.Hs;
    if (scale == 0.f) scale = 1.0f / (float)N;      // x/N == x*(1/N) exactly for a power of two
    for (int n = threadIdx.x; n < h; n += blockDim.x) {
        const float2 v = zero_frame ? make_float2(0.f, 0.f) : r[n];
        const int i = (2 * n + h) & (N - 1);        // half swap: y'[i] = y[(i+N/2) mod N]
        const float y0 = (v.x * scale) * d.win[i];
        const float y1 = (v.y * scale) * d.win[i + 1];
        const int p0 = (pos0 + i) & (N - 1), p1 = (pos0 + i + 1) & (N - 1);
        acc[p0] = (i < keep ? acc[p0] : 0.f) + y0;
        acc[p1] = (i + 1 < keep ? acc[p1] : 0.f) + y1;
    }
}

// ---------------------------------------------------------------------------------------------
// per-frame API kernels
// ---------------------------------------------------------------------------------------------

// One CTA per frame: steps A-D, all 2N bins written as {mag, phase}.
__global__ void __launch_bounds__(kGenericThreads)
analysis_batch_kernel(PvDev d, const float *__restrict__ in, int64_t n_in, float2 *__restrict__ out)
{
    extern __shared__ float2 sm[];
    const int N = d.N;
    float2 *a = sm, *b = sm + N;
    const int64_t k = blockIdx.x;
    load_frame_compat(a, in, k * (int64_t)d.Ha, n_in, d);
    __syncthreads();
    const float2 *C = fft_stockham4(a, b, N, 2, d.tw, d.N, false);
    float2 *o = out + k * 2 * (int64_t)N;
    for (int kk = threadIdx.x; kk <= N; kk += blockDim.x) {
        float2 X;
        if (kk == N) X = make_float2(C[0].x - C[0].y, 0.f);
        else X = split_bin(C, kk, d);
        o[kk] = mag_phase(X, d.flags);
        if (kk > 0 && kk < N) o[2 * N - kk] = mag_phase(make_float2(X.x, -X.y), d.flags);
    }
}

// One CTA walks the frames of one spectra batch sequentially: steps E-I.
// full_frame_out != nullptr: additionally dump the whole accumulated frame of the LAST frame
// (the per-frame API returns all N samples, kernel.cu:352).
__global__ void __launch_bounds__(kGenericThreads)
resynthesis_batch_kernel(PvDev d, const float2 *__restrict__ spectra, int64_t n_frames,
                         const float *__restrict__ back_in, float *__restrict__ back_out,
                         float *__restrict__ out, float *__restrict__ full_frame_out)
{
    extern __shared__ float2 sm[];
    const int N = d.N, h = N >> 1;
    float2 *a = sm, *b = sm + h;
    float *acc = reinterpret_cast<float *>(sm + N);
    for (int i = threadIdx.x; i < N; i += blockDim.x)
        acc[i] = (back_in != nullptr && i + d.Hs < N) ? back_in[i + d.Hs] : 0.f;
    __syncthreads();
    int pos0 = 0;
    for (int64_t k = 0; k < n_frames; ++k) {
        const float2 *sp = spectra + k * 2 * (int64_t)N;
        build_inverse_input(a, [&](int kk) { return polar_to_rect_d2(sp[kk]); }, d);
        __syncthreads();
        const float2 *r = fft_stockham4(a, b, h, 4, d.tw, d.N, true);
        ola_accumulate(acc, r, pos0, false, d);
        __syncthreads();
        if (out != nullptr)
            for (int j = threadIdx.x; j < d.Hs; j += blockDim.x)
                out[k * (int64_t)d.Hs + j] = acc[(pos0 + j) & (N - 1)];
        if (k + 1 == n_frames) {
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                const float v = acc[(pos0 + i) & (N - 1)];
                if (back_out != nullptr) back_out[i] = v;
                if (full_frame_out != nullptr) full_frame_out[i] = v;
            }
        }
        __syncthreads();
        pos0 = (pos0 + d.Hs) & (N - 1);
    }
}

// test_overlap_add (kernel.cu:289-298): the two half swaps cancel.
__global__ void test_overlap_add_kernel(PvDev d, const float *__restrict__ in,
                                        const float *__restrict__ back, float *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.N) return;
    float v = (in[i] * d.win[i]) * d.win[i];
    if (i + d.Hs < d.N) v += back[i + d.Hs];
    out[i] = v;
}

// ---------------------------------------------------------------------------------------------
// generic fused compat stream kernel: one CTA per frame-range segment
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGenericThreads)
compat_generic_kernel(PvDev d, PvProcessArgs a)
{
    extern __shared__ float2 sm[];
    const int N = d.N, h = N >> 1;
    float2 *bufA = sm, *bufB = sm + N;
    float *acc = reinterpret_cast<float *>(sm + N + fft_work_elems(N));
    const PvSegment seg = a.segs[blockIdx.x];
    const float *in = a.in + seg.stream * a.in_stride;
    float *out = a.out + seg.stream * a.out_stream_stride;
    float *state = a.state ? reinterpret_cast<float *>(a.state + seg.stream * a.state_stride) : nullptr;

    for (int i = threadIdx.x; i < N; i += blockDim.x)
        acc[i] = (seg.carry_in && state && i + d.Hs < N) ? state[i + d.Hs] : 0.f;
    __syncthreads();

    int pos0 = 0;
    for (int64_t k = seg.k_begin; k < seg.k_end; ++k) {
        const bool analysed = k < a.n_analysed;
        const float2 *r = nullptr;
        if (analysed) {
            load_frame_compat(bufA, in, k * (int64_t)d.Ha, a.n_in, d);
            __syncthreads();
            float2 *C = fft_auto(bufA, bufB, N, 2, d.tw, d.N, false);
            float2 *Z = (C == bufA) ? bufB : bufA;
            build_inverse_input(Z, [&](int kk) {
                return polar_to_rect_d2(mag_phase(split_bin(C, kk, d), d.flags)); }, d);
            __syncthreads();
            r = fft_auto(Z, C, h, 4, d.tw, d.N, true);
        }
        ola_accumulate(acc, r, pos0, !analysed, d);
        __syncthreads();
        if (k >= seg.k_emit)
            for (int j = threadIdx.x; j < d.Hs; j += blockDim.x)
                out[k * (int64_t)d.Hs + j] = acc[(pos0 + j) & (N - 1)];
        if (seg.carry_out && state && k + 1 == seg.k_end)
            for (int i = threadIdx.x; i < N; i += blockDim.x)
                state[i] = acc[(pos0 + i) & (N - 1)];
        __syncthreads();
        pos0 = (pos0 + d.Hs) & (N - 1);
    }
}

// ---------------------------------------------------------------------------------------------
// Large-window compat stream kernel (window 4096 = 16^3 packed points): the whole frame lives in ONE padded
// shared-memory buffer.  Pass 1 reads the windowed, zero-phase, zero-padded frame straight from global memory
// (the zero half prunes the butterfly), every later pass of both transforms runs in place (load, barrier,
// store), the real-FFT split / collapsed steps D+E / Hermitian pack go through registers between two barriers,
// and the overlap-add ring sits next to the buffer: 51 KB per CTA of N/16 threads instead of 84 KB, three CTAs
// per SM by registers.  Same segment contract as compat_generic_kernel.
// ---------------------------------------------------------------------------------------------
// Pulls the NEW hop of the next frame (samples [base + N, base + N + Ha)) into L2 while the current frame is being
// transformed: the in-place kernels read their input with plain loads in pass 1, and each line is used once.
__device__ __forceinline__ void prefetch_next_hop(const float *__restrict__ in, long long base, int N, int Ha, long long n_in)
{
    for (long long s = base + N + 32ll * threadIdx.x; s < base + N + Ha && s < n_in; s += 32ll * blockDim.x)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(in + s));
}

__device__ __forceinline__ float2 compat_map_fast(float2 X, bool nan_compat)
{
    // steps D+E collapsed (re' = |Re X|, im' = Re X Im X / |X|), as in the tuned kernels (pv_fused_core.cuh)
    const float m2 = X.x * X.x + X.y * X.y;
    if (m2 == 0.f) {
        const float z = nan_compat ? __builtin_nanf("") : 0.f;
        return make_float2(z, z);
    }
    if (m2 < 1.17549435e-38f) return make_float2(fabsf(X.x), 0.f);
    return make_float2(fabsf(X.x), X.x * X.y * rsqrtf(m2));
}

template <int LG_N>
__global__ void __launch_bounds__(1 << (LG_N - 4), 3)
compat_inplace_kernel(PvDev d, PvProcessArgs a)
{
    using namespace pvsmem;
    constexpr int N = 1 << LG_N, T = N / 16, H = N / 2, Q = N / 4, J = Q / T + 1;
    static_assert(LG_N == 12, "the in-place schedule below is written for 16^3 packed points");
    extern __shared__ float2 sm[];
    float2 *W = sm;
    float *acc = reinterpret_cast<float *>(sm + fft_work_elems(N));
    const int tid = threadIdx.x;
    const PvSegment seg = a.segs[blockIdx.x];
    const float *in = a.in + seg.stream * a.in_stride;
    float *out = a.out + seg.stream * a.out_stream_stride;
    float *state = a.state ? reinterpret_cast<float *>(a.state + seg.stream * a.state_stride) : nullptr;
    const bool nan_compat = (d.flags & PV_FLAG_NAN_COMPAT) != 0;
    const int Hs = d.Hs, keep = N - Hs;
    const float scale = 1.0f / (float)N;
    const HalfTw twF{d.tw, 2, N / 2}, twI{d.tw, 4, H / 2};
    NoTw none;

    for (int i = tid; i < N; i += T) acc[i] = (seg.carry_in && state && i + Hs < N) ? state[i + Hs] : 0.f;
    __syncthreads();

    int pos0 = 0;
    for (int64_t k = seg.k_begin; k < seg.k_end; ++k) {
        const bool analysed = k < a.n_analysed;
        const int64_t base0 = k * (int64_t)d.Ha;
        if (analysed) {
            {   // forward pass 1 (Ns = 1): c[n] = f[N/2 + 2n] for n < N/4, f[2(n - 3N/4)] for n >= 3N/4, else 0
                const int64_t base = base0;
                float2 v[16];
#pragma unroll
                for (int r = 0; r < 16; r++) {
                    if (r >= 4 && r < 12) { v[r] = make_float2(0.f, 0.f); continue; }
                    const int n = tid + r * T;
                    const int i = r < 4 ? H + 2 * n : 2 * (n - 3 * Q);
                    const int64_t g = base + i;
                    const float x0 = g < a.n_in ? in[g] : 0.f, x1 = g + 1 < a.n_in ? in[g + 1] : 0.f;
                    v[r] = make_float2(x0 * d.win[i], x1 * d.win[i + 1]);
                }
                dft_pruned_fwd<16>(v);
                float2 *dst = W + pad(tid * 16);
#pragma unroll
                for (int r = 0; r < 16; r++) dst[r] = v[r];
            }
            __syncthreads();
            if (k + 1 < seg.k_end) prefetch_next_hop(in, base0, N, d.Ha, a.n_in);
            {
                const auto tw2 = load_reg_tw<LG_N, 16, -1>(twF);
                stockham_pass<LG_N, 16, 4, -1, false, false, true>(W, W, T, twF, tw2);
            }
            __syncthreads();
            stockham_pass<LG_N, 16, 8, -1, false, false, true>(W, W, T, twF, none);
            __syncthreads();
            // split + steps D/E + Hermitian pack for kk and H - kk, through registers
            float2 z0[J], z1[J];
#pragma unroll
            for (int j = 0; j < J; j++) {
                const int kk = tid + j * T;
                if (kk > Q) break;
                auto bin = [&](int b) -> float2 {      // bin b <= N/2 of the 2N-point real spectrum
                    const float2 aa = W[pad(b)];
                    float2 bb = W[pad((N - b) & (N - 1))];
                    bb.y = -bb.y;
                    const float2 e = make_float2(0.5f * (aa.x + bb.x), 0.5f * (aa.y + bb.y));
                    const float2 dd = make_float2(aa.x - bb.x, aa.y - bb.y);
                    const float2 o = make_float2(0.5f * dd.y, -0.5f * dd.x);
                    const float2 t = cmul(d.tw[b], o);
                    return make_float2(e.x + t.x, e.y + t.y);
                };
                float2 yk = compat_map_fast(bin(kk), nan_compat), ym = compat_map_fast(bin(H - kk), nan_compat);
                if (kk == 0) { yk.y = 0.f; ym.y = 0.f; }            // C2R ignores Im of bins 0 and N/2
                const float2 t = d.tw[2 * kk];                       // e^{-2 pi i kk / N}
                z0[j] = herm_pack(yk, ym, make_float2(t.x, -t.y));
                z1[j] = herm_pack(ym, yk, make_float2(-t.x, -t.y));
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < J; j++) {
                const int kk = tid + j * T;
                if (kk > Q) break;
                W[pad(kk)] = z0[j];
                if (kk != 0 && kk != Q) W[pad(H - kk)] = z1[j];
            }
            __syncthreads();
            // inverse: H = 16 * 16 * 8 packed points, all in place
            stockham_pass<LG_N - 1, 16, 0, 1, false, false, true>(W, W, H / 16, twI, none);
            __syncthreads();
            {
                const auto tw2 = load_reg_tw<LG_N - 1, 16, 1>(twI);
                stockham_pass<LG_N - 1, 16, 4, 1, false, false, true>(W, W, H / 16, twI, tw2);
            }
            __syncthreads();
        }
        // steps G+H: /N, half swap, window, overlap-add into the ring (the fresh tail replaces what was emitted),
        // fed straight from the registers of the last inverse pass
        auto ola = [&](int n, float2 v) {
            const int i = (2 * n + H) & (N - 1);
            const float y0 = (v.x * scale) * d.win[i], y1 = (v.y * scale) * d.win[i + 1];
            const int p0 = (pos0 + i) & (N - 1), p1 = (pos0 + i + 1) & (N - 1);
            acc[p0] = (i < keep ? acc[p0] : 0.f) + y0;
            acc[p1] = (i + 1 < keep ? acc[p1] : 0.f) + y1;
        };
        if (analysed) stockham_pass_sink<LG_N - 1, 8, 8, 1>(W, H / 8, twI, ola);
        else
            for (int n = tid; n < H; n += T) ola(n, make_float2(0.f, 0.f));
        __syncthreads();
        if (k >= seg.k_emit)
            for (int j = tid; j < Hs; j += T) out[k * (int64_t)Hs + j] = acc[(pos0 + j) & (N - 1)];
        if (seg.carry_out && state && k + 1 == seg.k_end)
            for (int i = tid; i < N; i += T) state[i] = acc[(pos0 + i) & (N - 1)];
        __syncthreads();
        pos0 = (pos0 + Hs) & (N - 1);
    }
}


// ---------------------------------------------------------------------------------------------
// Large-window corrected stream kernels (window 4096: 2048 = 16 * 16 * 8 packed points each way), same in-place
// scheme as compat_inplace_kernel, stream state (previous phase, accumulators, OLA rings) in shared memory.
// The forward half -- transform, real-FFT split, magnitude and phase -- is ONE out-of-line function shared by the
// processing kernel and the phase-carry aggregate: both run the very same instructions, so their phases agree
// bit for bit, which is what makes frame-range splitting and sharding exact.
// ---------------------------------------------------------------------------------------------
template <int LG_N>
// Outputs per bin: magnitude, and the unwrapped phase difference D against the previous frame's phase (the plain
// phase when there is no previous frame); Pprev is advanced to this frame's phase.  Every bin has one owner thread.
__device__ __noinline__ void analysis_inplace(float2 *W, const float *__restrict__ in, long long base, long long n_in,
                                              const float *__restrict__ win, const float2 *__restrict__ tw, float *magS,
                                              int32_t *dS, uint32_t *Pprev, bool have_prev, int Ha)
{
    using namespace pvsmem;
    constexpr int N = 1 << LG_N, M = N / 2, T = N / 16, LG_M = LG_N - 1;
    const int tid = threadIdx.x;
    const HalfTw twF{tw, 4, M / 2};
    NoTw none;
    if (tid < M / 16) {     // forward pass 1 (Ns = 1) straight from global memory: c[m] = windowed pair (M + 2m) mod N
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int m = tid + r * (M / 16);
            const int i = (M + 2 * m) & (N - 1);
            const long long g = base + i;
            const float x0 = g < n_in ? in[g] : 0.f, x1 = g + 1 < n_in ? in[g + 1] : 0.f;
            v[r] = make_float2(x0 * win[i], x1 * win[i + 1]);
        }
        dft<16, -1>(v);
        float2 *dst = W + pad(tid * 16);
#pragma unroll
        for (int r = 0; r < 16; r++) dst[r] = v[r];
    }
    __syncthreads();
    {
        const auto tw2 = load_reg_tw<LG_M, 16, -1>(twF);
        stockham_pass<LG_M, 16, 4, -1, false, false, true>(W, W, M / 16, twF, tw2);
    }
    __syncthreads();
    stockham_pass<LG_M, 8, 8, -1, false, false, true>(W, W, M / 8, twF, none);
    __syncthreads();
    for (int kk = tid; kk <= M / 2; kk += T) {
        float2 xk, xm;
        if (kk == 0) {
            const float2 c0 = W[0];
            xk = make_float2(c0.x + c0.y, 0.f);
            xm = make_float2(c0.x - c0.y, 0.f);
        } else {
            const float2 aa = W[pad(kk)], bb = W[pad(M - kk)];
            const float2 e = make_float2(0.5f * (aa.x + bb.x), 0.5f * (aa.y - bb.y));
            const float2 o = make_float2(0.5f * (aa.y + bb.y), -0.5f * (aa.x - bb.x));
            const float2 t = cmul(tw[2 * kk], o);          // W_N^k = exp(-j*pi*2k/N)
            xk = make_float2(e.x + t.x, e.y + t.y);
            xm = make_float2(e.x - t.x, -(e.y - t.y));
        }
        auto put = [&](int b, float2 x) {
            magS[b] = sqrtf(x.x * x.x + x.y * x.y);
            const uint32_t P = pvfused::phase_turns32(x.x, x.y);
            const uint32_t nomA = ((uint32_t)b * (uint32_t)Ha) << (32 - LG_N);      // integer wrap-around does the unwrap
            dS[b] = have_prev ? (int32_t)(P - Pprev[b] - nomA) : (int32_t)P;
            Pprev[b] = P;
        };
        put(kk, xk);
        if (kk != M - kk) put(M - kk, xm);
    }
}

template <int LG_N>
struct InplaceLayout {
    static constexpr int N = 1 << LG_N, M = N / 2, NB = M + 1, NBP = (NB + 3) & ~3;
    // W | mag | phase/D | previous phase | psi[V] | acc[V]
    static size_t bytes(int V, bool with_state)
    {
        size_t b = sizeof(float2) * (size_t)fft_work_elems(M) + 3 * sizeof(float) * NBP;
        if (with_state) b += (size_t)V * NBP * 8 + (size_t)V * N * 4;
        return b;
    }
};

template <int LG_N>
__global__ void __launch_bounds__(1 << (LG_N - 4), 3)
corrected_inplace_kernel(PvDev d, PvProcessArgs a)
{
    using namespace pvsmem;
    using L = InplaceLayout<LG_N>;
    constexpr int N = L::N, M = L::M, NB = L::NB, NBP = L::NBP, T = N / 16, LG_M = LG_N - 1;
    extern __shared__ float2 sm[];
    float2 *W = sm;
    float *magS = reinterpret_cast<float *>(sm + fft_work_elems(M));
    int32_t *dS = reinterpret_cast<int32_t *>(magS + NBP);
    uint32_t *Pprev = reinterpret_cast<uint32_t *>(dS + NBP);
    unsigned long long *psi = reinterpret_cast<unsigned long long *>(Pprev + NBP);
    const int V = d.V, Hs = d.Hs, tid = threadIdx.x, keep = N - Hs, lsh = 32 - d.lgN;
    float *acc = reinterpret_cast<float *>(psi + (size_t)V * NBP);
    const PvSegment seg = a.segs[blockIdx.x];
    const float *in = a.in + seg.stream * a.in_stride;
    float *out = a.out + seg.stream * a.out_stream_stride;
    unsigned char *state = a.state + (long long)seg.state_idx * a.state_stride;   // always present (caller or scratch)
    uint32_t *hdr = reinterpret_cast<uint32_t *>(state);
    uint32_t *gPprev = hdr + 2;
    unsigned long long *gpsi = reinterpret_cast<unsigned long long *>(state + 8 + ((NB * 4 + 7) / 8) * 8);
    float *gacc = reinterpret_cast<float *>(gpsi + (size_t)V * NB);
    const float scale = d.gain / (float)N;
    const HalfTw twI{d.tw, 4, M / 2};
    NoTw none;

    bool have_prev = seg.carry_in ? hdr[0] != 0 : false;
    int pos0 = seg.carry_in ? (Hs & (N - 1)) : 0;       // the state keeps each ring linear, last frame at index 0
    for (int i = tid; i < NB; i += T) Pprev[i] = seg.carry_in ? gPprev[i] : 0u;
    for (int v = 0; v < V; v++) {
        for (int i = tid; i < NB; i += T) psi[(size_t)v * NBP + i] = seg.carry_in ? gpsi[(size_t)v * NB + i] : 0ull;
        for (int i = tid; i < N; i += T) acc[(size_t)v * N + i] = seg.carry_in ? gacc[(size_t)v * N + i] : 0.f;
    }
    __syncthreads();

    for (long long k = seg.k_begin; k < seg.k_end; ++k) {
        analysis_inplace<LG_N>(W, in, k * (long long)d.Ha, a.n_in, d.win, d.tw, magS, dS, Pprev, have_prev, d.Ha);
        if (k + 1 < seg.k_end) prefetch_next_hop(in, k * (long long)d.Ha, N, d.Ha, a.n_in);
        __syncthreads();
        for (int v = 0; v < V; v++) {
            unsigned long long *ps = psi + (size_t)v * NBP;
            const unsigned long long Rq = d.Rq[v];
            const unsigned long long bqs = (d.beta_q[v] * (unsigned long long)Hs) << lsh;    // nomS[s] = a_hi * bqs mod 2^64
            const uint32_t *gt = d.gather_nat + (size_t)v * NB;                                 // a_lo | a_hi << 16, NB = no source bin
            auto synth = [&](int s) -> float2 {
                const uint32_t ge = __ldg(gt + s);
                const uint32_t lo = ge & 0xffffu, hi = ge >> 16;
                if (lo == (uint32_t)NB) return make_float2(0.f, 0.f);
                float m = magS[lo];
                for (uint32_t b = lo + 1; b <= hi; b++) m += magS[b];
                const int32_t dd = dS[hi];
                unsigned long long p;
                if (!have_prev) p = (unsigned long long)(uint32_t)dd << 32;
                else p = pvfused::mad_s32_u64(dd, Rq, pvfused::mad_u32_u64(hi, bqs, ps[s]));
                ps[s] = p;
                const float2 cs = pvfused::cis_turns64(p);
                return make_float2(m * cs.x, m * cs.y);
            };
            // Hermitian pack of the N-point inverse: Z[kk], Z[M - kk] from Y[kk], Y[M - kk]; every pair has one owner
            for (int kk = tid; kk <= M / 2; kk += T) {
                float2 yk = synth(kk);
                float2 ym = (kk == M - kk) ? yk : synth(M - kk);
                if (kk == 0) { yk.y = 0.f; ym.y = 0.f; }            // the inverse ignores Im of bins 0 and N/2
                const float2 t = d.tw[2 * kk];
                W[pad(kk)] = herm_pack(yk, ym, make_float2(t.x, -t.y));
                if (kk != 0 && kk != M / 2) W[pad(M - kk)] = herm_pack(ym, yk, make_float2(-t.x, -t.y));
            }
            __syncthreads();
            stockham_pass<LG_M, 16, 0, 1, false, false, true>(W, W, M / 16, twI, none);
            __syncthreads();
            {
                const auto tw2 = load_reg_tw<LG_M, 16, 1>(twI);
                stockham_pass<LG_M, 16, 4, 1, false, false, true>(W, W, M / 16, twI, tw2);
            }
            __syncthreads();
            float *ac = acc + (size_t)v * N;
            // last inverse pass, its outputs scaled, windowed and overlap-added straight from the registers
            stockham_pass_sink<LG_M, 8, 8, 1>(W, M / 8, twI, [&](int n, float2 r) {
                const int i = (2 * n + M) & (N - 1);
                const float y0 = (r.x * scale) * d.win[i], y1 = (r.y * scale) * d.win[i + 1];
                const int p0 = (pos0 + i) & (N - 1), p1 = (pos0 + i + 1) & (N - 1);
                ac[p0] = (i < keep ? ac[p0] : 0.f) + y0;
                ac[p1] = (i + 1 < keep ? ac[p1] : 0.f) + y1;
            });
            __syncthreads();
            if (k >= seg.k_emit) {
                float *o = out + v * a.out_voice_stride + k * (long long)Hs;
                for (int j = tid; j < Hs; j += T) o[j] = ac[(pos0 + j) & (N - 1)];
            }
        }
        __syncthreads();
        have_prev = true;
        pos0 = (pos0 + Hs) & (N - 1);
    }
    // leave the state in its linear form: accumulated frame after the last frame at index 0
    const int plast = (pos0 - Hs) & (N - 1);
    if (tid == 0) { hdr[0] = 1u; hdr[1] = 0u; }
    for (int i = tid; i < NB; i += T) gPprev[i] = Pprev[i];
    for (int v = 0; v < V; v++) {
        for (int i = tid; i < NB; i += T) gpsi[(size_t)v * NB + i] = psi[(size_t)v * NBP + i];
        for (int i = tid; i < N; i += T) gacc[(size_t)v * N + i] = acc[(size_t)v * N + ((plast + i) & (N - 1))];
    }
}

// phase-carry aggregate on the same forward half (PvAggArgs), one CTA per frame-range segment
template <int LG_N>
__global__ void __launch_bounds__(1 << (LG_N - 4), 3)
aggregate_inplace_kernel(PvDev d, PvAggArgs a)
{
    using L = InplaceLayout<LG_N>;
    constexpr int N = L::N, M = L::M, NB = L::NB, NBP = L::NBP, T = N / 16;
    extern __shared__ float2 sm[];
    float2 *W = sm;
    float *magS = reinterpret_cast<float *>(sm + fft_work_elems(M));
    int32_t *dS = reinterpret_cast<int32_t *>(magS + NBP);
    uint32_t *Pp = reinterpret_cast<uint32_t *>(dS + NBP);
    const int tid = threadIdx.x;
    const long long sg = blockIdx.x;
    const PvSegment seg = a.segs[sg];
    const float *in = a.in + seg.stream * a.in_stride;
    const bool carried = seg.carry_in && a.P_prev != nullptr &&
                         (!a.P_prev_in_state || a.P_prev[(long long)seg.stream * a.P_prev_stride - 2] != 0u);
    bool have_prev = carried;
    for (int b = tid; b < NB; b += T) {
        Pp[b] = carried ? a.P_prev[(long long)seg.stream * a.P_prev_stride + b] : 0u;
        a.S[sg * NB + b] = 0;
        if (a.H) a.H[sg * NB + b] = 0;
        if (a.P_first) a.P_first[sg * NB + b] = 0u;
    }
    __syncthreads();
    for (long long k = seg.k_begin; k < seg.k_end; ++k) {
        long long *dst = reinterpret_cast<long long *>((k < seg.k_emit) ? a.H : a.S);
        analysis_inplace<LG_N>(W, in, k * (long long)d.Ha, a.n_in, d.win, d.tw, magS, dS, Pp, have_prev, d.Ha);
        __syncthreads();
        for (int b = tid; b < NB; b += T) {
            if (have_prev) { if (dst) dst[sg * NB + b] += (long long)dS[b]; }
            else if (a.P_first) a.P_first[sg * NB + b] = (uint32_t)dS[b];
        }
        have_prev = true;
        __syncthreads();
    }
    if (a.P_last)
        for (int b = tid; b < NB; b += T) a.P_last[sg * NB + b] = Pp[b];
}

// ---------------------------------------------------------------------------------------------
// generic fused CORRECTED stream kernel: one CTA per stream, any power-of-two window.
// Stream state (previous phase, phase accumulators, OLA rings) lives in the per-stream state buffer in
// global memory (L2 resident); the arithmetic is the specification of DESIGN.md "corrected mode".
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t g_phase_turns32(float re, float im)
{
    const float t = atan2f(im, re) * 0.15915494309189535f;
    return (uint32_t)__float2ll_rn(t * 4294967296.0f);
}

// SMEM_STATE: the stream state is copied into shared memory for the segment (when it fits next to the transform
// buffers with two CTAs per SM) and written back at the end; otherwise every frame works on it in global memory.
template <bool SMEM_STATE>
__global__ void __launch_bounds__(kGenericThreads)
corrected_generic_kernel(PvDev d, PvProcessArgs a)
{
    extern __shared__ float2 sm[];
    const int N = d.N, M = N >> 1, NB = M + 1, V = d.V, Hs = d.Hs;
    float2 *bufA = sm, *bufB = sm + NB, *Ys = bufB + fft_work_elems(M);
    float *magS = reinterpret_cast<float *>(Ys + NB);
    int32_t *dS = reinterpret_cast<int32_t *>(magS + NB);
    const PvSegment seg = a.segs[blockIdx.x];
    const float *in = a.in + seg.stream * a.in_stride;
    float *out = a.out + seg.stream * a.out_stream_stride;
    unsigned char *state = a.state + (long long)seg.state_idx * a.state_stride;   // always present (caller or scratch)
    uint32_t *hdr = reinterpret_cast<uint32_t *>(state);
    uint32_t *gPprev = hdr + 2;
    unsigned long long *gpsi = reinterpret_cast<unsigned long long *>(state + 8 + ((NB * 4 + 7) / 8) * 8);
    float *gacc = reinterpret_cast<float *>(gpsi + (size_t)V * NB);
    unsigned long long *psi = SMEM_STATE ? reinterpret_cast<unsigned long long *>(dS + NB) : gpsi;   // 8*NB bytes of mag + D: aligned
    float *acc = SMEM_STATE ? reinterpret_cast<float *>(psi + (size_t)V * NB) : gacc;
    uint32_t *Pprev = SMEM_STATE ? reinterpret_cast<uint32_t *>(acc + (size_t)V * N) : gPprev;
    const float scale = d.gain / (float)N;
    const int lsh = 32 - d.lgN;
    if (SMEM_STATE && seg.carry_in) {
        for (int i = threadIdx.x; i < NB; i += blockDim.x) Pprev[i] = gPprev[i];
        for (int i = threadIdx.x; i < V * NB; i += blockDim.x) psi[i] = gpsi[i];
        for (int i = threadIdx.x; i < V * N; i += blockDim.x) acc[i] = gacc[i];
    } else if (SMEM_STATE) {
        // accumulators of synthesis bins without a source bin are never written: start them at zero so that the
        // state written back does not carry stale shared memory
        for (int i = threadIdx.x; i < NB; i += blockDim.x) Pprev[i] = 0u;
        for (int i = threadIdx.x; i < V * NB; i += blockDim.x) psi[i] = 0ull;
    }

    // The state stores each voice's accumulated frame linearly (index 0 = first sample of the LAST frame).
    // It is used in place as a ring: the next frame starts at linear index Hs, and its last Hs samples
    // overwrite the already-emitted entries [0, Hs) (ola_accumulate's "fresh tail" rule).
    const bool have_prev0 = seg.carry_in ? hdr[0] != 0 : false;
    bool have_prev = have_prev0;
    int pos0 = seg.carry_in ? (Hs & (N - 1)) : 0;
    if (!seg.carry_in)
        for (int i = threadIdx.x; i < V * N; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();

    for (long long k = seg.k_begin; k < seg.k_end; ++k) {
        // ---- forward: N-point real FFT as N/2 complex points ----
        const long long base = k * (long long)d.Ha;
        for (int m = threadIdx.x; m < M; m += blockDim.x) {
            const int i = (M + 2 * m) & (N - 1);
            const long long g = base + i;
            const float x0 = g < a.n_in ? in[g] : 0.f, x1 = g + 1 < a.n_in ? in[g + 1] : 0.f;
            bufA[m] = make_float2(x0 * d.win[i], x1 * d.win[i + 1]);
        }
        __syncthreads();
        const float2 *C = fft_auto(bufA, bufB, M, 4, d.tw, d.N, false);
        for (int kk = threadIdx.x; kk <= M / 2; kk += blockDim.x) {
            float2 xk, xm;
            if (kk == 0) {
                xk = make_float2(C[0].x + C[0].y, 0.f);
                xm = make_float2(C[0].x - C[0].y, 0.f);
            } else {
                const float2 aa = C[kk];
                float2 bb = C[M - kk];
                const float2 e = make_float2(0.5f * (aa.x + bb.x), 0.5f * (aa.y - bb.y));
                const float2 o = make_float2(0.5f * (aa.y + bb.y), -0.5f * (aa.x - bb.x));
                const float2 t = cmul(d.tw[2 * kk], o);          // W_N^k = exp(-j*pi*2k/N)
                xk = make_float2(e.x + t.x, e.y + t.y);
                xm = make_float2(e.x - t.x, -(e.y - t.y));
            }
#pragma unroll
            for (int side = 0; side < 2; side++) {
                const int bin = side ? M - kk : kk;
                if (side && bin == kk) break;                     // kk == M/2: self-mirrored
                const float2 x = side ? xm : xk;
                const uint32_t Pc = g_phase_turns32(x.x, x.y);
                magS[bin] = sqrtf(x.x * x.x + x.y * x.y);
                const uint32_t nomA = ((uint32_t)bin * (uint32_t)d.Ha) << lsh;
                dS[bin] = have_prev ? (int32_t)(Pc - Pprev[bin] - nomA) : (int32_t)Pc;
                Pprev[bin] = Pc;
            }
        }
        __syncthreads();
        // ---- synthesis per voice ----
        for (int v = 0; v < V; v++) {
            for (int s = threadIdx.x; s < NB; s += blockDim.x) {
                const int lo = d.a_lo[v * NB + s], hi = d.a_hi[v * NB + s];
                float2 y = make_float2(0.f, 0.f);
                if (lo <= hi) {
                    float m = 0.f;
                    for (int b = lo; b <= hi; b++) m += magS[b];
                    const int32_t dd = dS[hi];
                    unsigned long long p;
                    if (!have_prev) p = (unsigned long long)(uint32_t)dd << 32;
                    else p = psi[(size_t)v * NB + s] + d.nomS[v * NB + s] +
                             (unsigned long long)((long long)dd * (long long)d.Rq[v]);
                    psi[(size_t)v * NB + s] = p;
                    const float t = (float)(int32_t)(p >> 32) * (1.0f / 4294967296.0f);
                    float sn, cs;
                    sincospif(2.0f * t, &sn, &cs);
                    y = make_float2(m * cs, m * sn);
                }
                Ys[s] = y;
            }
            __syncthreads();
            float2 *Z = bufA;
            build_inverse_input(Z, [&](int kk) { return Ys[kk]; }, d);
            __syncthreads();
            const float2 *r = fft_auto(Z, bufB, M, 4, d.tw, d.N, true);
            float *ac = acc + (size_t)v * N;
            ola_accumulate(ac, r, pos0, false, d, scale);
            __syncthreads();
            if (k >= seg.k_emit) {
                float *o = out + v * a.out_voice_stride + k * (long long)Hs;
                for (int j = threadIdx.x; j < Hs; j += blockDim.x) o[j] = ac[(pos0 + j) & (N - 1)];
            }
            __syncthreads();
        }
        have_prev = true;
        pos0 = (pos0 + Hs) & (N - 1);
    }
    // leave the state in its linear form: accumulated frame after the last frame at index 0
    const int plast = (pos0 - Hs) & (N - 1);
    if (threadIdx.x == 0) { hdr[0] = 1u; hdr[1] = 0u; }
    if (SMEM_STATE) {
        __syncthreads();
        for (int i = threadIdx.x; i < NB; i += blockDim.x) gPprev[i] = Pprev[i];
        for (int i = threadIdx.x; i < V * NB; i += blockDim.x) gpsi[i] = psi[i];
        for (int i = threadIdx.x; i < V * N; i += blockDim.x) gacc[i] = acc[(i & ~(N - 1)) + ((plast + i) & (N - 1))];
        return;
    }
    for (int v = 0; v < V; v++) {
        float *ac = acc + (size_t)v * N;
        // rotate the ring by plast through shared memory
        float *tmp = reinterpret_cast<float *>(sm);
        for (int i = threadIdx.x; i < N; i += blockDim.x) tmp[i] = ac[(plast + i) & (N - 1)];
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x) ac[i] = tmp[i];
        __syncthreads();
    }
}

// generic phase-carry aggregate (analysis only), one CTA per frame-range segment (PvAggArgs)
__global__ void __launch_bounds__(kGenericThreads)
aggregate_generic_kernel(PvDev d, PvAggArgs a)
{
    extern __shared__ float2 sm[];
    const int N = d.N, M = N >> 1, NB = M + 1, lsh = 32 - d.lgN;
    float2 *bufA = sm, *bufB = sm + NB;
    uint32_t *Pp = reinterpret_cast<uint32_t *>(bufB + fft_work_elems(M));
    const long long sg = blockIdx.x;
    const PvSegment seg = a.segs[sg];
    const float *in = a.in + seg.stream * a.in_stride;
    const bool carried = seg.carry_in && a.P_prev != nullptr &&
                         (!a.P_prev_in_state || a.P_prev[(long long)seg.stream * a.P_prev_stride - 2] != 0u);
    bool have_prev = carried;
    for (int b = threadIdx.x; b < NB; b += blockDim.x) {
        Pp[b] = carried ? a.P_prev[(long long)seg.stream * a.P_prev_stride + b] : 0u;
        a.S[sg * NB + b] = 0;
        if (a.H) a.H[sg * NB + b] = 0;
        if (a.P_first) a.P_first[sg * NB + b] = 0u;
    }
    __syncthreads();
    for (long long k = seg.k_begin; k < seg.k_end; ++k) {
        const long long base = k * (long long)d.Ha;
        long long *dst = reinterpret_cast<long long *>((k < seg.k_emit) ? a.H : a.S);
        for (int m = threadIdx.x; m < M; m += blockDim.x) {
            const int i = (M + 2 * m) & (N - 1);
            const long long g = base + i;
            const float x0 = g < a.n_in ? in[g] : 0.f, x1 = g + 1 < a.n_in ? in[g + 1] : 0.f;
            bufA[m] = make_float2(x0 * d.win[i], x1 * d.win[i + 1]);
        }
        __syncthreads();
        const float2 *C = fft_auto(bufA, bufB, M, 4, d.tw, d.N, false);
        for (int kk = threadIdx.x; kk <= M / 2; kk += blockDim.x) {
            float2 xk, xm;
            if (kk == 0) {
                xk = make_float2(C[0].x + C[0].y, 0.f);
                xm = make_float2(C[0].x - C[0].y, 0.f);
            } else {
                const float2 aa = C[kk], bb = C[M - kk];
                const float2 e = make_float2(0.5f * (aa.x + bb.x), 0.5f * (aa.y - bb.y));
                const float2 o = make_float2(0.5f * (aa.y + bb.y), -0.5f * (aa.x - bb.x));
                const float2 t = cmul(d.tw[2 * kk], o);
                xk = make_float2(e.x + t.x, e.y + t.y);
                xm = make_float2(e.x - t.x, -(e.y - t.y));
            }
#pragma unroll
            for (int side = 0; side < 2; side++) {
                const int bin = side ? M - kk : kk;
                if (side && bin == kk) break;
                const float2 x = side ? xm : xk;
                const uint32_t Pc = g_phase_turns32(x.x, x.y);
                const uint32_t nomA = ((uint32_t)bin * (uint32_t)d.Ha) << lsh;
                if (have_prev) { if (dst) dst[sg * NB + bin] += (long long)(int32_t)(Pc - Pp[bin] - nomA); }
                else if (a.P_first) a.P_first[sg * NB + bin] = Pc;
                Pp[bin] = Pc;
            }
        }
        have_prev = true;
        __syncthreads();
    }
    if (a.P_last)
        for (int b = threadIdx.x; b < NB; b += blockDim.x) a.P_last[sg * NB + b] = Pp[b];
}

}  // namespace

// window 4096 with the state of all voices next to the transform buffer and at least two CTAs per SM
static bool use_inplace_corrected(const PvDev &d)
{
    return d.N == 4096 && InplaceLayout<12>::bytes(d.V, true) <= 200 * 1024 && !getenv("PV_NO_INPLACE");
}

cudaError_t pv_launch_aggregate_generic(const PvDev &d, const PvAggArgs &a, cudaStream_t st)
{
    if (a.n_segs <= 0) return cudaSuccess;
    if (use_inplace_corrected(d)) {         // must be the aggregate that shares analysis_inplace with the processing kernel
        const size_t smem = InplaceLayout<12>::bytes(d.V, false);
        cudaError_t e = pv_max_smem_once<aggregate_inplace_kernel<12>, false>();
        if (e != cudaSuccess) return e;
        aggregate_inplace_kernel<12><<<a.n_segs, 256, smem, st>>>(d, a);
        return cudaGetLastError();
    }
    const size_t NB = d.N / 2 + 1;
    const size_t smem = sizeof(float2) * (NB + fft_work_elems(d.N / 2)) + sizeof(uint32_t) * NB;
    cudaError_t e = pv_max_smem_once<aggregate_generic_kernel, false>();
    if (e != cudaSuccess) return e;
    aggregate_generic_kernel<<<(unsigned)a.n_segs, generic_threads(d.N), smem, st>>>(d, a);
    return cudaGetLastError();
}

// ---- 16-bit PCM conversions (AudioFile rules) on a column range of a row-major batch, 4 samples per thread ----
__global__ void pcm16_to_float_kernel(const int16_t *__restrict__ in, float *__restrict__ out, long long pitch,
                                      long long c0, long long c1, long long n_valid)
{
    const long long r = blockIdx.y;
    const long long c = c0 + 4 * (blockIdx.x * (long long)blockDim.x + threadIdx.x);   // c0 and pitch: multiples of 4
    if (c >= c1) return;
    const int16_t *src = in + r * pitch + c;
    float4 v;                                                    // sixteenBitIntToSample, AudioFile.h:1038-1042
    v.x = c + 0 < n_valid ? (float)src[0] / 32768.0f : 0.f;
    v.y = c + 1 < n_valid ? (float)src[1] / 32768.0f : 0.f;
    v.z = c + 2 < n_valid ? (float)src[2] / 32768.0f : 0.f;
    v.w = c + 3 < n_valid ? (float)src[3] / 32768.0f : 0.f;
    *reinterpret_cast<float4 *>(out + r * pitch + c) = v;
}

__device__ __forceinline__ int16_t to_pcm16(float x)
{
    x = fminf(x, 1.0f);                                          // clamp: NaN -> 1 like std::min(value, max)
    x = fmaxf(x, -1.0f);
    // sampleToSixteenBitInt, AudioFile.h:1045-1049: trunc((double)x * 32767.0).  |x*32767| < 2^15 and a float
    // holds 24 bits, so rounding the product toward zero keeps its integer part: same result without FP64.
    return (int16_t)(int)__fmul_rz(x, 32767.0f);
}

__global__ void float_to_pcm16_kernel(const float *__restrict__ in, int16_t *__restrict__ out, long long pitch,
                                      long long c0, long long c1, int vec)
{
    const long long r = blockIdx.y;
    const long long c = c0 + 4 * (blockIdx.x * (long long)blockDim.x + threadIdx.x);
    if (c >= c1) return;
    const float *src = in + r * pitch + c;
    int16_t *dst = out + r * pitch + c;
    if (vec && c + 3 < c1) {
        const float4 v = *reinterpret_cast<const float4 *>(src);
        short4 o;
        o.x = to_pcm16(v.x); o.y = to_pcm16(v.y); o.z = to_pcm16(v.z); o.w = to_pcm16(v.w);
        *reinterpret_cast<short4 *>(dst) = o;
    } else {
        for (int j = 0; j < 4 && c + j < c1; j++) dst[j] = to_pcm16(src[j]);
    }
}

// in/out: `rows` rows of `pitch` samples (pitch and c0 multiples of 4); converts columns [c0, c1), zero-filling
// columns >= n_valid up to the next multiple of 4
cudaError_t pv_launch_pcm16_to_float(const int16_t *in, float *out, int64_t rows, int64_t pitch, int64_t c0, int64_t c1,
                                     int64_t n_valid, cudaStream_t st)
{
    if (rows <= 0 || c1 <= c0) return cudaSuccess;
    if ((pitch & 3) || (c0 & 3)) return cudaErrorInvalidValue;
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {          // rows ride in gridDim.y: slabs of at most 65535
        const int64_t nr = std::min<int64_t>(65535, rows - r0);
        const dim3 grid((unsigned)(((c1 - c0 + 3) / 4 + 255) / 256), (unsigned)nr);
        pcm16_to_float_kernel<<<grid, 256, 0, st>>>(in + r0 * pitch, out + r0 * pitch, pitch, c0, c1, n_valid);
    }
    return cudaGetLastError();
}

// in/out: `rows` rows of `pitch` samples; converts columns [c0, c1)
cudaError_t pv_launch_float_to_pcm16(const float *in, int16_t *out, int64_t rows, int64_t pitch, int64_t c0, int64_t c1,
                                     cudaStream_t st)
{
    if (rows <= 0 || c1 <= c0) return cudaSuccess;
    const int vec = !((pitch | c0) & 3);
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        const int64_t nr = std::min<int64_t>(65535, rows - r0);
        const dim3 grid((unsigned)(((c1 - c0 + 3) / 4 + 255) / 256), (unsigned)nr);
        float_to_pcm16_kernel<<<grid, 256, 0, st>>>(in + r0 * pitch, out + r0 * pitch, pitch, c0, c1, vec);
    }
    return cudaGetLastError();
}

// ---- packed 24-bit PCM (three little-endian bytes per sample), AudioFile rules: sign-extend, / 8388608 in
// (src/AudioFile.h:508-518); (int32)(x * 8388608), low three bytes out (:755-766: no clamp there -- the conversion
// saturates here where the reference's is undefined, |x| >= 256) ----
__global__ void pcm24_to_float_kernel(const uint8_t *__restrict__ in, float *__restrict__ out, long long pitch,
                                      long long c0, long long c1, long long n_valid)
{
    const long long r = blockIdx.y;
    const long long c = c0 + 4 * (blockIdx.x * (long long)blockDim.x + threadIdx.x);   // c0 and pitch: multiples of 4
    if (c >= c1) return;
    // four samples = 12 bytes = three aligned words (rows are `pitch` samples = 3 * pitch bytes, pitch % 4 == 0)
    const uint32_t *src = reinterpret_cast<const uint32_t *>(in + (r * pitch + c) * 3);
    const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];
    const int32_t s0 = (int32_t)(w0 << 8) >> 8;
    const int32_t s1 = (int32_t)(((w0 >> 24) | (w1 << 8)) << 8) >> 8;
    const int32_t s2 = (int32_t)(((w1 >> 16) | (w2 << 16)) << 8) >> 8;
    const int32_t s3 = (int32_t)w2 >> 8;
    float4 v;
    v.x = c + 0 < n_valid ? (float)s0 / 8388608.0f : 0.f;
    v.y = c + 1 < n_valid ? (float)s1 / 8388608.0f : 0.f;
    v.z = c + 2 < n_valid ? (float)s2 / 8388608.0f : 0.f;
    v.w = c + 3 < n_valid ? (float)s3 / 8388608.0f : 0.f;
    *reinterpret_cast<float4 *>(out + r * pitch + c) = v;
}

__device__ __forceinline__ uint32_t to_pcm24(float x)
{
    // x * 2^23 is exact in fp32 (a power of two), the cast truncates toward zero like the reference's
    return (uint32_t)__float2int_rz(x * 8388608.0f) & 0xffffffu;
}

__global__ void float_to_pcm24_kernel(const float *__restrict__ in, uint8_t *__restrict__ out, long long pitch,
                                      long long c0, long long c1, int vec)
{
    const long long r = blockIdx.y;
    const long long c = c0 + 4 * (blockIdx.x * (long long)blockDim.x + threadIdx.x);
    if (c >= c1) return;
    const float *src = in + r * pitch + c;
    uint8_t *dst = out + (r * pitch + c) * 3;
    if (vec && c + 3 < c1) {
        const float4 v = *reinterpret_cast<const float4 *>(src);
        const uint32_t a = to_pcm24(v.x), b = to_pcm24(v.y), cc = to_pcm24(v.z), d = to_pcm24(v.w);
        uint32_t *w = reinterpret_cast<uint32_t *>(dst);
        w[0] = a | (b << 24);
        w[1] = (b >> 8) | (cc << 16);
        w[2] = (cc >> 16) | (d << 8);
    } else {
        for (int j = 0; j < 4 && c + j < c1; j++) {
            const uint32_t a = to_pcm24(src[j]);
            dst[3 * j] = (uint8_t)a; dst[3 * j + 1] = (uint8_t)(a >> 8); dst[3 * j + 2] = (uint8_t)(a >> 16);
        }
    }
}

// same contracts as the 16-bit launchers; `in` / `out` are byte pointers, pitch and columns count SAMPLES
cudaError_t pv_launch_pcm24_to_float(const uint8_t *in, float *out, int64_t rows, int64_t pitch, int64_t c0, int64_t c1,
                                     int64_t n_valid, cudaStream_t st)
{
    if (rows <= 0 || c1 <= c0) return cudaSuccess;
    if ((pitch & 3) || (c0 & 3) || (reinterpret_cast<uintptr_t>(in) & 3)) return cudaErrorInvalidValue;
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        const int64_t nr = std::min<int64_t>(65535, rows - r0);
        const dim3 grid((unsigned)(((c1 - c0 + 3) / 4 + 255) / 256), (unsigned)nr);
        pcm24_to_float_kernel<<<grid, 256, 0, st>>>(in + r0 * pitch * 3, out + r0 * pitch, pitch, c0, c1, n_valid);
    }
    return cudaGetLastError();
}

cudaError_t pv_launch_float_to_pcm24(const float *in, uint8_t *out, int64_t rows, int64_t pitch, int64_t c0, int64_t c1,
                                     cudaStream_t st)
{
    if (rows <= 0 || c1 <= c0) return cudaSuccess;
    const int vec = !((pitch | c0) & 3) && !(reinterpret_cast<uintptr_t>(out) & 3);
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        const int64_t nr = std::min<int64_t>(65535, rows - r0);
        const dim3 grid((unsigned)(((c1 - c0 + 3) / 4 + 255) / 256), (unsigned)nr);
        float_to_pcm24_kernel<<<grid, 256, 0, st>>>(in + r0 * pitch, out + r0 * pitch * 3, pitch, c0, c1, vec);
    }
    return cudaGetLastError();
}

cudaError_t pv_launch_corrected_generic(const PvDev &d, const PvProcessArgs &a, cudaStream_t st)
{
    if (a.n_segs <= 0) return cudaSuccess;
    if (use_inplace_corrected(d)) {
        const size_t smem = InplaceLayout<12>::bytes(d.V, true);
        cudaError_t e = pv_max_smem_once<corrected_inplace_kernel<12>, false>();
        if (e != cudaSuccess) return e;
        corrected_inplace_kernel<12><<<a.n_segs, 256, smem, st>>>(d, a);
        return cudaGetLastError();
    }
    const size_t NB = d.N / 2 + 1;
    const size_t smem = std::max(sizeof(float2) * (2 * NB + fft_work_elems(d.N / 2)) + sizeof(float) * 2 * NB, sizeof(float) * (size_t)d.N);
    const size_t with_state = smem + (size_t)d.V * NB * 8 + (size_t)d.V * d.N * 4 + NB * 4;
    if (with_state <= 113 * 1024) {         // two CTAs per SM still fit
        cudaError_t e = pv_max_smem_once<corrected_generic_kernel<true>, false>();
        if (e != cudaSuccess) return e;
        corrected_generic_kernel<true><<<a.n_segs, generic_threads(d.N), with_state, st>>>(d, a);
        return cudaGetLastError();
    }
    cudaError_t e = pv_max_smem_once<corrected_generic_kernel<false>, false>();
    if (e != cudaSuccess) return e;
    corrected_generic_kernel<false><<<a.n_segs, generic_threads(d.N), smem, st>>>(d, a);
    return cudaGetLastError();
}

cudaError_t pv_launch_analysis_batch(const PvDev &d, const float *in, int64_t n_in, int64_t n_frames,
                                     float *out_magphase, cudaStream_t st)
{
    if (n_frames <= 0) return cudaSuccess;
    const size_t smem = sizeof(float2) * 2 * (size_t)d.N;
    cudaError_t e = pv_max_smem_once<analysis_batch_kernel, false>();
    if (e != cudaSuccess) return e;
    analysis_batch_kernel<<<(unsigned)n_frames, generic_threads(d.N), smem, st>>>(
        d, in, n_in, reinterpret_cast<float2 *>(out_magphase));
    return cudaGetLastError();
}

static cudaError_t launch_resynth(const PvDev &d, const float *spectra, int64_t n_frames, const float *back_in,
                                  float *back_out, float *out, float *full, cudaStream_t st)
{
    const size_t smem = sizeof(float2) * (size_t)d.N + sizeof(float) * (size_t)d.N;
    cudaError_t e = pv_max_smem_once<resynthesis_batch_kernel, false>();
    if (e != cudaSuccess) return e;
    resynthesis_batch_kernel<<<1, generic_threads(d.N), smem, st>>>(d, reinterpret_cast<const float2 *>(spectra), n_frames,
                                                             back_in, back_out, out, full);
    return cudaGetLastError();
}

cudaError_t pv_launch_resynthesis_batch(const PvDev &d, const float *spectra, int64_t n_frames, float *back,
                                        float *out, cudaStream_t st)
{
    if (n_frames <= 0) return cudaSuccess;
    return launch_resynth(d, spectra, n_frames, back, back, out, nullptr, st);
}

cudaError_t pv_launch_resynthesis_frame(const PvDev &d, const float *back, const float *front, float *out,
                                        cudaStream_t st)
{
    return launch_resynth(d, front, 1, back, nullptr, nullptr, out, st);
}

cudaError_t pv_launch_test_overlap_add(const PvDev &d, const float *in, const float *back, float *out,
                                       cudaStream_t st)
{
    test_overlap_add_kernel<<<(d.N + 255) / 256, 256, 0, st>>>(d, in, back, out);
    return cudaGetLastError();
}

cudaError_t pv_launch_compat_generic(const PvDev &d, const PvProcessArgs &a, cudaStream_t st)
{
    if (a.n_segs <= 0) return cudaSuccess;
    if (d.N == 4096 && !getenv("PV_NO_INPLACE")) {
        const size_t smem = sizeof(float2) * (size_t)fft_work_elems(d.N) + sizeof(float) * (size_t)d.N;
        cudaError_t e = pv_max_smem_once<compat_inplace_kernel<12>, false>();
        if (e != cudaSuccess) return e;
        compat_inplace_kernel<12><<<a.n_segs, 256, smem, st>>>(d, a);
        return cudaGetLastError();
    }
    const size_t smem = sizeof(float2) * (size_t)(d.N + fft_work_elems(d.N)) + sizeof(float) * (size_t)d.N;
    cudaError_t e = pv_max_smem_once<compat_generic_kernel, false>();
    if (e != cudaSuccess) return e;
    compat_generic_kernel<<<a.n_segs, generic_threads(d.N), smem, st>>>(d, a);
    return cudaGetLastError();
}
