// pv_fused_kernels.cu -- the tuned fused stream kernels (compat mode) for windows 256..4096.
//
// Work unit: a frame-range segment of one stream (PvSegment).  A group of T = N/16 threads owns
// a segment and walks its frames sequentially, keeping the overlap-add accumulator in shared
// memory; G groups share a CTA of 128 threads.  The per-frame body is pv_fused_core.cuh.
//
// Per-frame shared-memory traffic is two exchanges per transform direction plus the OLA ring;
// the real-FFT split, the reference's mag/phase + polar->rect steps and the Hermitian pack of
// the inverse run in registers.  HBM traffic is the compulsory 4*Ha + 4*Hs bytes per frame
// (the 75 % overlap re-reads of the input hit L1/L2).
#include <cstdlib>

#include "pv_fused_core.cuh"
#include "pv_internal.h"

// every emitted hop is zeroed as it is written out: it becomes the fresh tail of the next frame, so the overlap-add is a
// plain accumulate (pv_fused_core.cuh, inverse_23_ola)
#define PV_ZERO_ON_EMIT true

namespace {

using namespace pvfused;

template <int LOG2N>
struct Launch {
    using S = Shape<LOG2N>;
    static constexpr int T = S::T;
    static constexpr int G = (T >= 128) ? 1 : 128 / T;       // groups per CTA
    static constexpr int THREADS = T * G;
    static constexpr int GROUP_F2 = S::BUF_A + S::BUF_B;     // float2 per group
    // resident CTAs per SM with the 16-byte ring: 40.7 KB and 96 registers at window 2048 -> 5; window 4096 (256 threads,
    // 81 KB) -> 2
    static constexpr int MINB_RING = LOG2N >= 12 ? 2 : 5;
    // exchange buffers + OLA accumulator (+ private input ring)
    static constexpr size_t smem(bool ring) { return (size_t)G * (GROUP_F2 * sizeof(float2) + (ring ? 2 : 1) * S::N * sizeof(float) + 16); }   // + mbarrier
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

// RING: 0 = inputs straight from global memory, 1 = private 8-byte cp.async ring (each thread copies exactly
// the samples it consumes, no barrier), 2 = cooperative 16-byte cp.async.cg ring (L1 bypass), 3 = the new hop of every frame
// as ONE bulk asynchronous copy (cp.async.bulk + mbarrier, the TMA engine's 1-D form) issued by one elected thread
template <int LOG2N, int MINB, int RING>
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
    float *fbase = reinterpret_cast<float *>(reinterpret_cast<float2 *>(smem_raw) + (size_t)G * L::GROUP_F2);
    float *acc = fbase + (size_t)g * N;
    float *ring = RING ? fbase + (size_t)(G + g) * N : nullptr;
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(fbase + (size_t)(RING ? 2 : 1) * G * N) + 2 * g;
    unsigned mbar_phase = 0;
    bool bulk_pending = false;

    GroupSync<T, G> sync{g, T >= 32 ? 0xffffffffu : (((1u << (T & 31)) - 1u) << ((threadIdx.x & 31) / T * T))};

    const PvSegment seg = a.segs[seg_idx];
    const float *in = a.in + seg.stream * a.in_stride;
    float *out = a.out + seg.stream * a.out_stream_stride;
    float *state = a.state ? reinterpret_cast<float *>(a.state + seg.stream * a.state_stride) : nullptr;
    const int Hs = d.Hs;
    const bool nan_compat = (d.flags & PV_FLAG_NAN_COMPAT) != 0;

    for (int i = tid; i < N; i += T) acc[i] = (seg.carry_in && state && i + Hs < N) ? state[i + Hs] : 0.f;
    sync();

    // emits the output hop of frame kk whose ring position is pp (src/main.cpp:281-295)
    // `zero`: clear the hop after reading it -- it becomes the fresh tail of the next frame, which lets the
    // overlap-add be a plain accumulate (no per-sample "first writer" test)
    auto emit = [&](long long kk, int pp, bool zero) {
        const bool wr = kk >= seg.k_emit;
        float *o = out + kk * (long long)Hs;
        if (vec_out_ok && (Hs & 3) == 0) {
            for (int j = 4 * tid; j < Hs; j += 4 * T) {
                float4 *sl = reinterpret_cast<float4 *>(acc + ((pp + j) & (N - 1)));
                if (wr) *reinterpret_cast<float4 *>(o + j) = *sl;
                if (zero) *sl = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int j = tid; j < Hs; j += T) {
                float *sl = acc + ((pp + j) & (N - 1));
                if (wr) o[j] = *sl;
                if (zero) *sl = 0.f;
            }
        }
    };

    const long long k_an_end = seg.k_end < a.n_analysed ? seg.k_end : a.n_analysed;   // frames >= this are zero spectra
    if (RING == 3 && tid == 0) mbar_init(mbar, 1);
    if (RING && seg.k_begin < k_an_end) {
        FrameIO io0{in, a.n_in, seg.k_begin * (long long)d.Ha, true, true};
        if (RING == 1) ring_prefetch<LOG2N>(tid, io0, ring, 0);
        else {
            ring_prefetch_coop16<N, T>(tid, io0, ring, 0);
            cp_async_wait_all();
            sync();
        }
    }
    // per-thread twiddle bases live in registers for the whole segment
    constexpr bool TWREG = (S::S1 == S::T) && (S::R1 == 16);
    const ThreadTw tt = load_thread_tw<LOG2N>(tid, tb);
    int pos0 = 0;
    for (long long k = seg.k_begin; k < seg.k_end; ++k) {
        FrameIO io{in, a.n_in, k * (long long)d.Ha, k < a.n_analysed, vec_in_ok != 0};
        // runs after the first barrier of the frame: prefetch the next frame's new samples into this
        // thread's private ring slots, then write out the previous frame's hop
        auto hook = [&]() {
            if (RING && k + 1 < k_an_end) {
                FrameIO nx{in, a.n_in, (k + 1) * (long long)d.Ha, true, true};
                const long long g0 = nx.base + (N - d.Ha);            // first new sample of the next frame
                if (RING == 1) ring_prefetch<LOG2N>(tid, nx, ring, N - d.Ha);
                else if (RING == 3 && g0 + d.Ha <= a.n_in) {           // whole hop inside the stream: one bulk copy
                    if (tid == 0) bulk_load_hop(ring + (int)(g0 & (N - 1)), in + g0, (unsigned)d.Ha * 4u, mbar);
                    bulk_pending = true;
                } else ring_prefetch_coop16<N, T>(tid, nx, ring, N - d.Ha);
            }
            if (k > seg.k_begin) emit(k - 1, (pos0 - Hs) & (N - 1), PV_ZERO_ON_EMIT);
        };
        if (RING == 1) cp_async_wait_all();
        frame_compat<LOG2N, TWREG>(tid, io, tb, tt, nan_compat, ring, bufA, bufB, acc, pos0, Hs, sync, hook,
                                   [&]() { if (RING >= 2) cp_async_wait_all(); });
        if (RING == 3 && bulk_pending) {      // every consumer observes the completion itself
            mbar_wait(mbar, mbar_phase);
            mbar_phase ^= 1u;
            bulk_pending = false;
        }
        pos0 = (pos0 + Hs) & (N - 1);
    }
    sync();
    const int plast = (pos0 - Hs) & (N - 1);
    emit(seg.k_end - 1, plast, false);
    if (seg.carry_out && state)
        for (int i = tid; i < N; i += T) state[i] = acc[(plast + i) & (N - 1)];
}

// The private ring needs every sample to be consumed by one thread (Ha multiple of 2*S1) and
// 8-byte aligned rows for cp.async.
template <int LOG2N>
bool ring_ok(const PvDev &d, bool vec_in_ok) { return vec_in_ok && d.Ha <= d.N && (d.Ha % (2 * Shape<LOG2N>::S1)) == 0; }

template <int LOG2N, int MINB, int RING>
cudaError_t launch2(const PvDev &d, const Tables &tb, const PvProcessArgs &a, int vec_in_ok, int vec_out_ok,
                    cudaStream_t st)
{
    using L = Launch<LOG2N>;
    auto kern = compat_fused_kernel<LOG2N, MINB, RING>;
    const size_t smem = L::smem(RING != 0);
    cudaError_t e = pv_max_smem_once<compat_fused_kernel<LOG2N, MINB, RING>>();
    if (e != cudaSuccess) return e;
    const int grid = (a.n_segs + L::G - 1) / L::G;
    kern<<<grid, L::THREADS, smem, st>>>(d, tb, a, vec_in_ok, vec_out_ok);
    return cudaGetLastError();
}

template <int LOG2N, int MINB>
cudaError_t launch(const PvDev &d, const Tables &tb, const PvProcessArgs &a, int vec_in_ok, int vec_out_ok,
                   cudaStream_t st)
{
    // 16-byte cooperative ring needs 16-byte aligned rows and hops
    const bool al16 = vec_in_ok && d.Ha <= d.N && (d.Ha % 4 == 0) && (a.in_stride % 4 == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.in) & 15) == 0);
    // 5 CTAs/SM: with exchange 2 in place a group needs 40.7 KB of shared memory at N = 2048 and 96 registers
    // Bulk asynchronous hop copies (ring mode 3) where they measured faster: window 4096 (+1.3 %).  At windows <= 2048 this
    // kernel runs five CTAs per SM on 96 registers and the mbarrier bookkeeping spills (-3 %), so the 16-byte cp.async ring
    // stays; PV_RING_BULK / PV_RING_LDGSTS force either for the A/B (read once).
    static const bool ldgsts = getenv("PV_RING_LDGSTS") != nullptr, force_bulk = getenv("PV_RING_BULK") != nullptr;
    if (al16 && !ldgsts && (LOG2N >= 12 || force_bulk) && d.N % d.Ha == 0) return launch2<LOG2N, Launch<LOG2N>::MINB_RING, 3>(d, tb, a, vec_in_ok, vec_out_ok, st);
    if (al16) return launch2<LOG2N, Launch<LOG2N>::MINB_RING, 2>(d, tb, a, vec_in_ok, vec_out_ok, st);
    if (ring_ok<LOG2N>(d, vec_in_ok != 0)) return launch2<LOG2N, MINB, 1>(d, tb, a, vec_in_ok, vec_out_ok, st);
    return launch2<LOG2N, MINB, 0>(d, tb, a, vec_in_ok, vec_out_ok, st);
}

}  // namespace

bool pv_fused_compat_supported(int N, int Hs) { return (N == 256 || N == 512 || N == 1024 || N == 2048 || N == 4096) && (Hs % 2) == 0; }

template <int LOG2N, int MINB>
static int capacity(int sm_count)
{
    using L = Launch<LOG2N>;
    auto kern = compat_fused_kernel<LOG2N, L::MINB_RING, 2>;
    const size_t smem = L::smem(true);
    int nb = 0;
    if (pv_max_smem_once<compat_fused_kernel<LOG2N, L::MINB_RING, 2>>() != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, L::THREADS, smem) != cudaSuccess || nb < 1)
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
        case 4096: return capacity<12, 2>(sm_count);
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
        case 4096: return launch<12, 2>(d, tb, a, in_ok, out_ok, st);
        default: return cudaErrorInvalidValue;
    }
}
