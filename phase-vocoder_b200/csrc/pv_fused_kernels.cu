// pv_fused_kernels.cu -- the tuned fused stream kernels (compat mode) for windows 256..2048.
//
// Work unit: a frame-range segment of one stream (PvSegment).  A group of T = N/16 threads owns
// a segment and walks its frames sequentially, keeping the overlap-add accumulator in shared
// memory; G groups share a CTA of 128 threads.  The per-frame body is pv_fused_core.cuh.
//
// Per-frame shared-memory traffic is two exchanges per transform direction plus the OLA ring;
// the real-FFT split, the reference's mag/phase + polar->rect steps and the Hermitian pack of
// the inverse run in registers.  HBM traffic is the compulsory 4*Ha + 4*Hs bytes per frame
// (the 75 % overlap re-reads of the input hit L1/L2).
#include "pv_fused_core.cuh"
#include "pv_internal.h"

namespace {

using namespace pvfused;

template <int LOG2N>
struct Launch {
    using S = Shape<LOG2N>;
    static constexpr int T = S::T;
    static constexpr int G = (T >= 128) ? 1 : 128 / T;       // groups per CTA
    static constexpr int THREADS = T * G;
    static constexpr int GROUP_F2 = S::BUF_A + S::BUF_B;     // float2 per group
    static constexpr size_t SMEM = (size_t)G * (GROUP_F2 * sizeof(float2) + S::N * sizeof(float));
};

template <int T, int G>
struct GroupSync {
    int g;
    unsigned mask;
    __device__ __forceinline__ void operator()() const
    {
        if constexpr (G == 1) __syncthreads();
        else if constexpr (T >= 32) asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(T) : "memory");
        else __syncwarp(mask);
    }
};

template <int LOG2N, int MINB>
__global__ void __launch_bounds__(Launch<LOG2N>::THREADS, MINB)
compat_fused_kernel(PvDev d, Tables tb, PvProcessArgs a, int vec_in_ok, int vec_out_ok)
{
    using S = Shape<LOG2N>;
    using L = Launch<LOG2N>;
    constexpr int N = S::N, T = S::T, G = L::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = threadIdx.x / T, tid = threadIdx.x % T;
    const int seg_idx = blockIdx.x * G + g;
    if (seg_idx >= a.n_segs) return;

    float2 *bufA = reinterpret_cast<float2 *>(smem_raw) + (size_t)g * L::GROUP_F2;
    float2 *bufB = bufA + S::BUF_A;
    float *acc = reinterpret_cast<float *>(reinterpret_cast<float2 *>(smem_raw) + (size_t)G * L::GROUP_F2) + (size_t)g * N;

    GroupSync<T, G> sync{g, T >= 32 ? 0xffffffffu : (((1u << (T & 31)) - 1u) << ((threadIdx.x & 31) / T * T))};

    const PvSegment seg = a.segs[seg_idx];
    const float *in = a.in + seg.stream * a.in_stride;
    float *out = a.out + seg.stream * a.out_stream_stride;
    float *state = a.state ? a.state + seg.stream * a.state_stride : nullptr;
    const int Hs = d.Hs;
    const bool nan_compat = (d.flags & PV_FLAG_NAN_COMPAT) != 0;

    for (int i = tid; i < N; i += T) acc[i] = (seg.carry_in && state && i + Hs < N) ? state[i + Hs] : 0.f;
    sync();

    int pos0 = 0;
    for (long long k = seg.k_begin; k < seg.k_end; ++k) {
        FrameIO io{in, a.n_in, k * (long long)d.Ha, k < a.n_analysed, vec_in_ok != 0};
        frame_compat<LOG2N>(tid, io, tb, nan_compat, bufA, bufB, acc, pos0, Hs, sync);
        sync();
        if (k >= seg.k_emit) {
            float *o = out + k * (long long)Hs;
            if (vec_out_ok) {
                for (int j = 4 * tid; j < Hs; j += 4 * T)
                    *reinterpret_cast<float4 *>(o + j) = *reinterpret_cast<const float4 *>(acc + ((pos0 + j) & (N - 1)));
            } else {
                for (int j = tid; j < Hs; j += T) o[j] = acc[(pos0 + j) & (N - 1)];
            }
        }
        if (seg.carry_out && state && k + 1 == seg.k_end)
            for (int i = tid; i < N; i += T) state[i] = acc[(pos0 + i) & (N - 1)];
        sync();
        pos0 = (pos0 + Hs) & (N - 1);
    }
}

template <int LOG2N, int MINB>
cudaError_t launch(const PvDev &d, const Tables &tb, const PvProcessArgs &a, int vec_in_ok, int vec_out_ok,
                   cudaStream_t st)
{
    using L = Launch<LOG2N>;
    auto kern = compat_fused_kernel<LOG2N, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM);
    if (e != cudaSuccess) return e;
    const int grid = (a.n_segs + L::G - 1) / L::G;
    kern<<<grid, L::THREADS, L::SMEM, st>>>(d, tb, a, vec_in_ok, vec_out_ok);
    return cudaGetLastError();
}

}  // namespace

bool pv_fused_compat_supported(int N, int Hs) { return (N == 256 || N == 512 || N == 1024 || N == 2048) && (Hs % 2) == 0; }

template <int LOG2N, int MINB>
static int capacity(int sm_count)
{
    using L = Launch<LOG2N>;
    auto kern = compat_fused_kernel<LOG2N, MINB>;
    int nb = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, L::THREADS, L::SMEM) != cudaSuccess || nb < 1)
        nb = 1;
    return nb * sm_count * L::G;
}

int pv_fused_compat_capacity(int N, int sm_count)
{
    switch (N) {
        case 256: return capacity<8, 4>(sm_count);
        case 512: return capacity<9, 4>(sm_count);
        case 1024: return capacity<10, 4>(sm_count);
        case 2048: return capacity<11, 4>(sm_count);
        default: return sm_count * 8;
    }
}

cudaError_t pv_launch_compat_fused(const PvDev &d, const PvFusedTables &t, const PvProcessArgs &a, cudaStream_t st)
{
    if (a.n_segs <= 0) return cudaSuccess;
    Tables tb{t.tw1, t.tw2, t.tw2n, t.itw1, t.itw2, d.win};
    const bool in_ok = (d.Ha % 2 == 0) && (a.in_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 7) == 0);
    const bool out_ok = (d.Hs % 4 == 0) && (a.out_stream_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
    switch (d.N) {
        case 256: return launch<8, 4>(d, tb, a, in_ok, out_ok, st);
        case 512: return launch<9, 4>(d, tb, a, in_ok, out_ok, st);
        case 1024: return launch<10, 4>(d, tb, a, in_ok, out_ok, st);
        case 2048: return launch<11, 4>(d, tb, a, in_ok, out_ok, st);
        default: return cudaErrorInvalidValue;
    }
}
