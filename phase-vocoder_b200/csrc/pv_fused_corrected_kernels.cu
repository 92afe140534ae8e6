// pv_fused_corrected_kernels.cu -- fused CORRECTED-mode stream kernel (windows 256..4096).
//
// Same work decomposition as the compat kernel: a group of T = N/16 threads walks the frames of
// one stream segment; the stream state (previous analysis phase in registers, 64-bit phase
// accumulators and the per-voice overlap-add rings in shared memory) never leaves the SM between
// frames.  HBM traffic per frame: 4*Ha bytes in, 4*V*Hs bytes out.
#include <algorithm>
#include <cstdlib>

#include "pv_fused_corrected.cuh"
#include "pv_internal.h"

// every emitted hop is zeroed as it is written out: it becomes the fresh tail of the next frame, so the overlap-add is a
// plain accumulate (pv_fused_core.cuh, inverse_23_ola)
#define PV_ZERO_ON_EMIT true

namespace {

using namespace pvfused;

template <int LOG2N>
struct CLaunch {
    using C = CShape<LOG2N>;
    static constexpr int T = C::T;
    static constexpr int G = (T >= 128) ? 1 : 128 / T;
    static constexpr int THREADS = T * G;
    // per group: exchange buffers | psi (u64) | mag | D | acc[V] | ring
    static size_t group_bytes(int V)
    {
        return (size_t)(C::BUF_A + C::BUF_B) * sizeof(float2) + (size_t)V * ((C::NB + 1) & ~1) * 8 +
               (size_t)((C::NB + 3) & ~3) * 4 * 2 + (size_t)V * C::N * 4 + (size_t)C::N * 4 + 16 + 128;
        // ... + the mbarrier + room behind the ring: the stored-analysis mode keeps a second {|X|, D} buffer of N/2 + 2 float2 =
        // 4 N + 16 bytes there.  128 and not 16: two groups of a small window share a warp, and the bank phase between their
        // bases (the group size mod 128) is part of the tuned layout -- +16 cost the two- and four-voice window-256 runs 5 %.
    }
};

template <int T, int G>
struct CGroupSync {
    int g;
    unsigned mask;
    __device__ __forceinline__ void operator()() const
    {
        if constexpr (G == 1) __syncthreads();
        else if constexpr (T >= 32) asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(T) : "memory");
        else __syncwarp(mask);
    }
};

// MODE (frame_corrected): 1 = the stored-analysis processing pass (PvProcessArgs::md), 2 = the analysis pass that stores
// (PvAggArgs::md); both are instantiations of their own: the normal one (processing with the forward transform, and the
// analysis-only mode that shares its compiled copy) sits at the 128-register limit and lost 2 - 5 % when the extra modes were
// run-time branches in it.
template <int LOG2N, int MINB, int MODE = 0>
__global__ void __launch_bounds__(CLaunch<LOG2N>::THREADS, MINB)
corrected_fused_kernel(PvDev d, CTables tb, PvProcessArgs a, int vec_in_ok, int vec_out_ok, int use_ring,
                       unsigned group_bytes, PvAggArgs ag)
{
    using C = CShape<LOG2N>;
    using L = CLaunch<LOG2N>;
    constexpr int N = C::N, T = C::T, G = L::G, NB = C::NB, B3 = C::B3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = threadIdx.x / T, tid = threadIdx.x % T;
    const int seg_idx = blockIdx.x * G + g;
    if (seg_idx >= a.n_segs) return;
    const int V = tb.V;

    unsigned char *base = smem_raw + (size_t)g * group_bytes;
    float2 *bufA = reinterpret_cast<float2 *>(base);
    float2 *bufB = bufA + C::BUF_A;
    unsigned long long *psi = reinterpret_cast<unsigned long long *>(bufB + C::BUF_B);
    float *magS = reinterpret_cast<float *>(psi + (size_t)V * ((NB + 1) & ~1));
    float2 *mdS = reinterpret_cast<float2 *>(magS);            // {|X|, D} per bin, NB + 1 entries (the last one: the dummy bin)
    float *acc = magS + 2 * ((NB + 3) & ~3);
    float *ring = use_ring ? acc + (size_t)V * N : nullptr;
    // ring mode 3: the new hop of every frame arrives as ONE bulk asynchronous copy (cp.async.bulk + mbarrier)
    unsigned long long *mbar = reinterpret_cast<unsigned long long *>(base + group_bytes - 16);
    unsigned mbar_phase = 0;
    bool bulk_pending = false;
    // Stored-analysis mode (PvProcessArgs::md): the frames' {|X|, D} come from global memory, double-buffered in the array's
    // own region and the (then unused) input ring, one bulk copy per frame on the ring's mbarrier
    constexpr int NBP = NB + 1;                                 // floats2 per stored frame (even: rows stay 16-byte aligned)
    // (nothing of this mode may stay live across the frame loop of the normal mode: the kernel sits at the 128-register limit;
    // buffer parity = frame parity within the segment, pointers are rebuilt where they are used)
    constexpr bool md_load = MODE == 1;
    unsigned long long *psi_v0 = psi;      // voice stride in psi is NB (kernel body) -> keep packed
    (void)psi_v0;

    CGroupSync<T, G> sync{g, T >= 32 ? 0xffffffffu : (((1u << (T & 31)) - 1u) << ((threadIdx.x & 31) / T * T))};

    const PvSegment seg = a.segs[seg_idx];
    const float *in = a.in + seg.stream * a.in_stride;
    const int Hs = d.Hs;

    // Analysis-only mode (phase-carry aggregate, PvAggArgs): the frame loop below is shared with the processing
    // mode -- ONE call site of frame_corrected, hence one compiled copy of the forward transform, so both modes
    // compute bit-identical phases.  The per-bin sums live in shared memory (psi / mag+D regions, unused then).
    const bool agg_mode = ag.S != nullptr;
    long long *sumS = reinterpret_cast<long long *>(psi);      // NB x 8 B
    long long *sumH = reinterpret_cast<long long *>(magS);     // mag + D regions: 2 x NB x 4 B
    uint32_t *agg_pf = (agg_mode && ag.P_first) ? ag.P_first + (long long)seg_idx * NB : nullptr;

    float *out = agg_mode ? nullptr : a.out + seg.stream * a.out_stream_stride;
    unsigned char *state = (a.state && !agg_mode) ? a.state + (long long)seg.state_idx * a.state_stride : nullptr;

    // state layout: [have_prev u32][pad][P_prev u32 x NB (8-byte padded)][psi u64 x V*NB][acc f32 x V*N]
    uint32_t *st_hdr = reinterpret_cast<uint32_t *>(state);
    uint32_t *st_P = st_hdr ? st_hdr + 2 : nullptr;
    // psi and the OLA rings of ALL the handle's voices; this launch works on voices [voice0, voice0 + V)
    unsigned long long *st_psi_all = state ? reinterpret_cast<unsigned long long *>(state + 8 + ((NB * 4 + 7) / 8) * 8) : nullptr;
    unsigned long long *st_psi = st_psi_all ? st_psi_all + (size_t)tb.voice0 * NB : nullptr;
    float *st_acc = st_psi_all ? reinterpret_cast<float *>(st_psi_all + (size_t)tb.V_total * NB) + (size_t)tb.voice0 * N : nullptr;

    const CThreadTw tt = load_cthread_tw<LOG2N>(tid, tb);
    CState st;
    st.have_prev = 0;
#pragma unroll
    for (int sl = 0; sl < 9; sl++) st.Pexp[sl] = 0;
    const bool cin = seg.carry_in && state;
    if (cin) {
        st.have_prev = (int)st_hdr[0];
#pragma unroll
        for (int sl = 0; sl < 9; sl++)
            if ((sl < 8 || tid == 0) && st.have_prev) st.Pexp[sl] = st_P[slot_bin<B3>(tid, sl)] + slot_nomA<LOG2N>(tid, sl, d.Ha);
    }
    if (agg_mode) {
        const bool carried = seg.carry_in && ag.P_prev != nullptr &&
                             (!ag.P_prev_in_state || ag.P_prev[(long long)seg.stream * ag.P_prev_stride - 2] != 0u);
        st.have_prev = carried;
#pragma unroll
        for (int sl = 0; sl < 9; sl++) {
            if (sl == 8 && tid != 0) break;
            const int bin = slot_bin<B3>(tid, sl);
            st.Pexp[sl] = carried ? ag.P_prev[(long long)seg.stream * ag.P_prev_stride + bin] + slot_nomA<LOG2N>(tid, sl, d.Ha) : 0u;
            sumS[bin] = 0;
            sumH[bin] = 0;
            if (agg_pf) agg_pf[bin] = 0u;
        }
    } else {
        if (tid == 0 && !md_load) mdS[NB] = make_float2(0.f, 0.f);     // the dummy bin behind empty gather entries (pv_fused_tables.h); stored rows carry it
        // zero until the stream's first frame: that frame's update relies on it (psi_step)
        for (int i = tid; i < V * NB; i += T) psi[i] = (cin && st.have_prev) ? st_psi[i] : 0ull;
        for (int i = tid; i < V * N; i += T) {
            const int ii = i & (N - 1);
            acc[i] = (cin && ii + Hs < N) ? st_acc[i + Hs] : 0.f;
        }
    }
    auto md_buf = [&](long long k) { return ((k - seg.k_begin) & 1) ? reinterpret_cast<float2 *>(acc + (size_t)V * N) : mdS; };
    auto md_fetch = [&](long long k) {           // thread 0: one bulk copy of frame k's stored row into its buffer
        const float2 *src = a.md + seg.stream * a.md_stream_stride + k * NBP;
        bulk_load_hop(reinterpret_cast<float *>(md_buf(k)), reinterpret_cast<const float *>(src), (unsigned)(NBP * sizeof(float2)), mbar);
    };
    if ((use_ring == 3 || md_load) && tid == 0) mbar_init(mbar, 1);
    if (md_load && tid == 0) md_fetch(seg.k_begin);
    if (use_ring) {
        FrameIO io0{in, a.n_in, seg.k_begin * (long long)d.Ha, true, true};
        if (use_ring >= 2) ring_prefetch_coop16<N, T>(tid, io0, ring, 0);
        else ring_prefetch_coop<N, T>(tid, io0, ring, 0);
        cp_async_wait_all();
    }
    sync();

    auto emit = [&](long long kk, int pp, bool zero) {
        const bool wr = kk >= seg.k_emit;
        for (int v = 0; v < V; v++) {
            float *o = out + v * a.out_voice_stride + kk * (long long)Hs;
            float *ac = acc + (size_t)v * N;
            if (vec_out_ok && (Hs & 3) == 0) {
                for (int j = 4 * tid; j < Hs; j += 4 * T) {
                    float4 *sl = reinterpret_cast<float4 *>(ac + ((pp + j) & (N - 1)));
                    if (wr) *reinterpret_cast<float4 *>(o + j) = *sl;
                    if (zero) *sl = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                for (int j = tid; j < Hs; j += T) {
                    float *sl = ac + ((pp + j) & (N - 1));
                    if (wr) o[j] = *sl;
                    if (zero) *sl = 0.f;
                }
            }
        }
    };

    int pos0 = 0;
    for (long long k = seg.k_begin; k < seg.k_end; ++k) {
        FrameIO io{in, a.n_in, k * (long long)d.Ha, true, vec_in_ok != 0};
        auto hook = [&]() {
            if (md_load) {
                // this frame's copy was issued a frame ago; only then the next one (one copy in flight per mbarrier phase),
                // into the buffer whose last readers -- the previous frame's slot loops -- are behind the frame's first barrier
                mbar_wait(mbar, mbar_phase);
                mbar_phase ^= 1u;
                if (k + 1 < seg.k_end && tid == 0) md_fetch(k + 1);
            }
            if (use_ring && k + 1 < seg.k_end) {
                FrameIO nx{in, a.n_in, (k + 1) * (long long)d.Ha, true, true};
                const long long g0 = nx.base + (N - d.Ha);                // first new sample of the next frame
                if (use_ring == 3 && g0 + d.Ha <= a.n_in) {              // whole hop inside the stream: one bulk copy
                    if (tid == 0) bulk_load_hop(ring + (int)(g0 & (N - 1)), in + g0, (unsigned)d.Ha * 4u, mbar);
                    bulk_pending = true;
                } else if (use_ring >= 2) ring_prefetch_coop16<N, T>(tid, nx, ring, N - d.Ha);
                else ring_prefetch_coop<N, T>(tid, nx, ring, N - d.Ha);
            }
            if (!agg_mode && k > seg.k_begin) emit(k - 1, (pos0 - Hs) & (N - 1), PV_ZERO_ON_EMIT);
        };
        const AggCtx ac{agg_mode, k < seg.k_emit ? sumH : sumS, agg_pf,
                        (MODE == 2 && k >= seg.k_emit) ? ag.md + seg.stream * ag.md_stream_stride + k * NBP : nullptr};
        frame_corrected<LOG2N, MODE>(tid, io, tb, tt, ring, bufA, bufB, md_load ? md_buf(k) : mdS, psi, acc, st, pos0, Hs, sync, hook,
                                       [&]() { if (use_ring) cp_async_wait_all(); }, ac);
        if (agg_mode && use_ring) {        // analysis only: no inverse passes whose last barrier would complete the refill
            cp_async_wait_all();
            sync();
        }
        if (bulk_pending) {                // every consumer observes the completion itself: no barrier needed for visibility
            mbar_wait(mbar, mbar_phase);
            mbar_phase ^= 1u;
            bulk_pending = false;
        }
        pos0 = (pos0 + Hs) & (N - 1);
    }
    if (agg_mode) {
#pragma unroll
        for (int sl = 0; sl < 9; sl++) {
            if (sl == 8 && tid != 0) break;
            const int bin = slot_bin<B3>(tid, sl);
            const long long o = (long long)seg_idx * NB + bin;
            ag.S[o] = sumS[bin];
            if (ag.H) ag.H[o] = sumH[bin];
            if (ag.P_last) ag.P_last[o] = st.have_prev ? st.Pexp[sl] - slot_nomA<LOG2N>(tid, sl, d.Ha) : 0u;
        }
        return;
    }
    sync();
    const int plast = (pos0 - Hs) & (N - 1);
    emit(seg.k_end - 1, plast, false);
    if (seg.carry_out && state) {
        if (tb.last_group) {
            if (tid == 0) { st_hdr[0] = (uint32_t)st.have_prev; st_hdr[1] = 0; }
#pragma unroll
            for (int sl = 0; sl < 9; sl++)
                if (sl < 8 || tid == 0) {
                    const int bin = slot_bin<B3>(tid, sl);
                    // stored analysis: the phases never passed through this launch; the analysis pass left the last one
                    st_P[bin] = md_load ? a.P_last[(long long)seg_idx * NB + bin]
                                        : (st.have_prev ? st.Pexp[sl] - slot_nomA<LOG2N>(tid, sl, d.Ha) : 0u);
                }
        }
        for (int i = tid; i < V * NB; i += T) st_psi[i] = psi[i];
        for (int i = tid; i < V * N; i += T) st_acc[i] = acc[(i & ~(N - 1)) + ((plast + i) & (N - 1))];
    }
}

// psi[v][s] = (P_first[a] << 32) + (n_before - 1) * nomS[v][s] + Rq[v] * sumD[a],  a = a_hi[v][s]
__global__ void state_from_carry_kernel(PvDev d, CTables tb, long long n_streams, const uint32_t *P_first,
                                        const long long *sumD, long long n_before, const uint32_t *P_prev,
                                        unsigned char *state, long long state_stride)
{
    const int NB = d.N / 2 + 1, V = tb.V, N = d.N;
    const long long s = blockIdx.x;
    if (s >= n_streams) return;
    unsigned char *st = state + s * state_stride;
    uint32_t *hdr = reinterpret_cast<uint32_t *>(st);
    uint32_t *stP = hdr + 2;
    unsigned long long *psi = reinterpret_cast<unsigned long long *>(st + 8 + ((NB * 4 + 7) / 8) * 8);
    float *acc = reinterpret_cast<float *>(psi + (size_t)V * NB);
    if (threadIdx.x == 0) { hdr[0] = n_before > 0 ? 1u : 0u; hdr[1] = 0u; }
    for (int b = threadIdx.x; b < NB; b += blockDim.x) stP[b] = (n_before > 0 && P_prev) ? P_prev[s * NB + b] : 0u;
    for (int i = threadIdx.x; i < V * NB; i += blockDim.x) {
        const int v = i / NB, sb = i - v * NB;
        const int lo = tb.a_lo[i], hi = tb.a_hi[i];
        unsigned long long p = 0;
        if (n_before > 0 && lo <= hi)
            p = ((unsigned long long)P_first[s * NB + hi] << 32) + (unsigned long long)(n_before - 1) * tb.nomS[i] +
                (unsigned long long)(sumD[s * NB + hi] * (long long)tb.Rq[v]);
        psi[i] = p;
        (void)sb;
    }
    for (int i = threadIdx.x; i < V * N; i += blockDim.x) acc[i] = 0.f;
}

template <int LOG2N, int MINB>
cudaError_t claunch(const PvDev &d, const CTables &tb, const PvProcessArgs &a, cudaStream_t st, bool no_ring = false)
{
    using L = CLaunch<LOG2N>;
    const bool in_ok = (d.Ha % 2 == 0) && (a.in_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 7) == 0);
    const bool out_ok = (d.Hs % 4 == 0) && (a.out_stream_stride % 4 == 0) && (a.out_voice_stride % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0);
    // 2: 16-byte cp.async.cg pieces (L1 bypass) when rows and hops are 16-byte aligned, 1: 8-byte pieces
    const bool al16 = in_ok && (d.Ha % 4 == 0) && (a.in_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.in) & 15) == 0);
    // 3: bulk asynchronous hop copies (needs the hop to divide the window so that a hop never wraps in the ring)
    static const bool ldgsts = getenv("PV_RING_LDGSTS") != nullptr;       // A/B switch, read once
    int ring = (in_ok && d.Ha <= d.N && !no_ring) ? (al16 ? 2 : 1) : 0;
    // (window 256: a bulk request per 256-byte hop and 16-thread group does not pay -- measured 908 M frames/s with the bulk ring,
    // 924 M with the 16-byte cp.async ring, also with a retry loop the two groups of a warp reconverge behind; the stored
    // analysis, 1 KB per request, loses there as well, pv_capi.cu)
    if (ring == 2 && !ldgsts && d.N % d.Ha == 0 && LOG2N >= 9) ring = 3;
    if (a.md) ring = 0;                                    // stored analysis: no input is read; the ring's region holds a buffer
    const size_t gb = L::group_bytes(tb.V) - ((ring || a.md) ? 0 : (size_t)d.N * 4);
    const size_t smem = gb * L::G;
    const int grid = (a.n_segs + L::G - 1) / L::G;
    cudaError_t e;
    if (a.md) {
        e = pv_max_smem_once<corrected_fused_kernel<LOG2N, MINB, 1>>();
        if (e != cudaSuccess) return e;
        corrected_fused_kernel<LOG2N, MINB, 1><<<grid, L::THREADS, smem, st>>>(d, tb, a, in_ok, out_ok, ring, (unsigned)gb, PvAggArgs{});
    } else {
        e = pv_max_smem_once<corrected_fused_kernel<LOG2N, MINB>>();
        if (e != cudaSuccess) return e;
        corrected_fused_kernel<LOG2N, MINB><<<grid, L::THREADS, smem, st>>>(d, tb, a, in_ok, out_ok, ring, (unsigned)gb, PvAggArgs{});
    }
    return cudaGetLastError();
}

template <int LOG2N, int MINB>
int ccapacity(int V, int sm_count)
{
    using L = CLaunch<LOG2N>;
    auto kern = corrected_fused_kernel<LOG2N, MINB>;
    const size_t smem = L::group_bytes(V) * L::G;
    int nb = 0;
    if (pv_max_smem_once<corrected_fused_kernel<LOG2N, MINB>>() != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, L::THREADS, smem) != cudaSuccess || nb < 1)
        nb = 1;
    return nb * sm_count * L::G;
}

}  // namespace

bool pv_fused_corrected_supported(int N, int Ha, int Hs)
{
    (void)Ha;
    return (N == 256 || N == 512 || N == 1024 || N == 2048 || N == 4096) && (Hs % 2) == 0;
}

int pv_fused_corrected_capacity(int N, int V, int sm_count)
{
    if (V >= 3 && !getenv("PV_VOICES_ONE_LAUNCH")) V = 2;        // voices_per_launch()
    switch (N) {
        case 256: return ccapacity<8, 4>(V, sm_count);
        case 512: return ccapacity<9, 4>(V, sm_count);
        case 1024: return ccapacity<10, 4>(V, sm_count);
        case 2048: return ccapacity<11, 4>(V, sm_count);
        case 4096: return ccapacity<12, 2>(V, sm_count);
        default: return sm_count;
    }
}

// Tables for voices [v0, v0 + nv) of the handle (nv = 0: all of them).
static CTables make_ctables(const PvDev &d, const PvFusedTables &t, int v0 = 0, int nv = 0)
{
    if (nv <= 0) nv = d.V - v0;
    const size_t nb = (size_t)d.N / 2 + 1;
    CTables tb{};
    tb.ctw1 = t.ctw1; tb.ctw2 = t.ctw2; tb.tw2n = t.tw2n; tb.itw1 = t.itw1; tb.itw2 = t.itw2;
    tb.win = d.win; tb.nomA = d.nomA;
    tb.a_lo = d.a_lo ? d.a_lo + (size_t)v0 * nb : nullptr;
    tb.a_hi = d.a_hi ? d.a_hi + (size_t)v0 * nb : nullptr;
    tb.nomS = d.nomS ? reinterpret_cast<const unsigned long long *>(d.nomS) + (size_t)v0 * nb : nullptr;
    tb.scale = d.gain / (float)d.N;
    tb.V = nv;
    tb.V_total = d.V;
    tb.voice0 = v0;
    tb.last_group = (v0 + nv == d.V) ? 1 : 0;
    tb.Ha = d.Ha;
    tb.gather = d.gather ? d.gather + (size_t)v0 * (size_t)(d.N / 16) * 9 : nullptr;
    for (int v = 0; v < nv; v++) {
        tb.Rq[v] = d.Rq[v0 + v];
        tb.beta_q[v] = d.beta_q[v0 + v];
        tb.bqs[v] = (d.beta_q[v0 + v] * (unsigned long long)d.Hs) << (32 - d.lgN);
        tb.multi[v] = d.multi[v0 + v];
    }
    return tb;
}

// The aggregate runs in the SAME kernel instantiation as the processing pass (see AggCtx).
template <int LOG2N, int MINB>
static cudaError_t agg_launch(const PvDev &d, const CTables &tb, const PvAggArgs &ag, cudaStream_t st)
{
    using L = CLaunch<LOG2N>;
    auto kern = corrected_fused_kernel<LOG2N, MINB>;
    const bool in_ok = (d.Ha % 2 == 0) && (ag.in_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(ag.in) & 7) == 0);
    // the analysis pass streams its input through the same shared-memory ring as the processing pass (each hop is
    // fetched once, one frame ahead); without it every frame re-reads its whole window from L2 with the latency exposed
    const bool al16 = in_ok && (d.Ha % 4 == 0) && (ag.in_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(ag.in) & 15) == 0);
    const int ring = (in_ok && d.Ha <= d.N) ? (al16 ? 2 : 1) : 0;
    const size_t gb = L::group_bytes(tb.V) - (ring ? 0 : (size_t)d.N * 4);
    const size_t smem = gb * L::G;
    cudaError_t e = ag.md ? pv_max_smem_once<corrected_fused_kernel<LOG2N, MINB, 2>>() : pv_max_smem_once<corrected_fused_kernel<LOG2N, MINB>>();
    if (e != cudaSuccess) return e;
    PvProcessArgs a{};
    a.in = ag.in;
    a.in_stride = ag.in_stride;
    a.n_in = ag.n_in;
    a.segs = ag.segs;
    a.n_segs = ag.n_segs;
    const int grid = (ag.n_segs + L::G - 1) / L::G;
    if (ag.md) corrected_fused_kernel<LOG2N, MINB, 2><<<grid, L::THREADS, smem, st>>>(d, tb, a, in_ok, 0, ring, (unsigned)gb, ag);
    else kern<<<grid, L::THREADS, smem, st>>>(d, tb, a, in_ok, 0, ring, (unsigned)gb, ag);
    return cudaGetLastError();
}

cudaError_t pv_launch_corrected_aggregate(const PvDev &d, const PvFusedTables &t, const PvAggArgs &a, cudaStream_t st)
{
    if (a.n_segs <= 0) return cudaSuccess;
    const CTables tb = make_ctables(d, t);
    switch (d.N) {
        case 256: return agg_launch<8, 4>(d, tb, a, st);
        case 512: return agg_launch<9, 4>(d, tb, a, st);
        case 1024: return agg_launch<10, 4>(d, tb, a, st);
        case 2048: return agg_launch<11, 4>(d, tb, a, st);
        case 4096: return agg_launch<12, 2>(d, tb, a, st);
        default: return cudaErrorInvalidValue;
    }
}

// Per-part carried states of split streams.  grid = (stream, tile of 256 entries): thread i of a stream owns entry i of
// every array it fills -- accumulator (voice, synthesis bin) i < V*NB with its own running sum of S at its source bin
// a_hi, previous-phase bin i < NB, words i of the copied caller state -- and walks the parts in order.  (Round 1 used ONE
// block per stream: 1.7 ms for the two channels of C5, 10 % of the whole call.)
__global__ void split_states_kernel(PvDev d, CTables tb, int parts, const PvSegment *segs, const long long *S,
                                    const long long *H, const uint32_t *P_first, unsigned char *slots, long long slot_stride,
                                    const unsigned char *state_in)
{
    const int NB = d.N / 2 + 1, V = tb.V, N = d.N;
    const long long s = blockIdx.x;
    const int i = blockIdx.y * blockDim.x + threadIdx.x;
    const int stride = gridDim.y * blockDim.x;
    const uint32_t *P0 = P_first + (s * parts) * NB;
    // carried-in stream: psi continues from the caller's accumulators and frame 0 has a phase difference too
    const unsigned char *sin = state_in ? state_in + s * slot_stride : nullptr;
    const bool cont = sin && reinterpret_cast<const uint32_t *>(sin)[0] != 0;
    const unsigned long long *psi_in = sin ? reinterpret_cast<const unsigned long long *>(sin + 8 + ((NB * 4 + 7) / 8) * 8) : nullptr;
    if (sin) {      // part 0 starts from the caller's state as it is
        unsigned char *dst = slots + (long long)segs[s * parts].state_idx * slot_stride;
        for (long long w = i; w < slot_stride / 4; w += stride)
            reinterpret_cast<uint32_t *>(dst)[w] = reinterpret_cast<const uint32_t *>(sin)[w];
    }
    const bool owns_psi = i < V * NB;
    const int v = owns_psi ? i / NB : 0;
    const int lo = owns_psi ? tb.a_lo[i] : 1, hi = owns_psi ? tb.a_hi[i] : 0;
    const bool src = owns_psi && lo <= hi;
    const unsigned long long base = src ? (cont ? psi_in[i] : ((unsigned long long)P0[hi] << 32)) : 0ull;
    const unsigned long long nomS = src ? tb.nomS[i] : 0ull;
    const long long Rq = (long long)tb.Rq[v];
    long long run = 0;                       // sum_{q<p} S_q[hi]
    for (int p = 0; p < parts; p++) {
        const long long gidx = s * parts + p;
        const PvSegment seg = segs[gidx];
        if (seg.carry_in && p > 0) {
            unsigned char *st = slots + (long long)seg.state_idx * slot_stride;
            uint32_t *hdr = reinterpret_cast<uint32_t *>(st);
            uint32_t *stP = hdr + 2;
            unsigned long long *psi = reinterpret_cast<unsigned long long *>(st + 8 + ((NB * 4 + 7) / 8) * 8);
            float *acc = reinterpret_cast<float *>(psi + (size_t)V * NB);
            if (i == 0) { hdr[0] = 1u; hdr[1] = 0u; }
            if (i < NB) stP[i] = P_first[gidx * NB + i];
            if (owns_psi) {
                const unsigned long long nb4 = (unsigned long long)(cont ? seg.k_begin : seg.k_begin - 1);
                psi[i] = src ? base + nb4 * nomS + (unsigned long long)((run - H[gidx * NB + hi]) * Rq) : 0ull;
            }
            for (int w = i; w < V * N; w += stride) acc[w] = 0.f;
        }
        if (src) run += S[gidx * NB + hi];
    }
}

__global__ void reduce_parts_kernel(int nb, int parts, const long long *S, const uint32_t *Pf, const uint32_t *Pl,
                                    long long *sumD, uint32_t *P_first, uint32_t *P_last)
{
    const long long s = blockIdx.x;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        long long acc = 0;
        for (int p = 0; p < parts; p++) acc += S[(s * parts + p) * nb + b];
        sumD[s * nb + b] = acc;
        if (P_first) P_first[s * nb + b] = Pf[(s * parts) * nb + b];
        if (P_last) P_last[s * nb + b] = Pl[(s * parts + parts - 1) * nb + b];
    }
}

cudaError_t pv_launch_reduce_parts(int nb, int64_t n_streams, int32_t parts, const int64_t *S, const uint32_t *Pf,
                                   const uint32_t *Pl, int64_t *sumD, uint32_t *P_first, uint32_t *P_last, cudaStream_t st)
{
    if (n_streams <= 0) return cudaSuccess;
    reduce_parts_kernel<<<(unsigned)n_streams, 256, 0, st>>>(nb, parts, reinterpret_cast<const long long *>(S), Pf, Pl,
                                                            reinterpret_cast<long long *>(sumD), P_first, P_last);
    return cudaGetLastError();
}

cudaError_t pv_launch_split_states(const PvDev &d, int64_t n_streams, int32_t parts, const PvSegment *proc_segs,
                                   const int64_t *S, const int64_t *H, const uint32_t *P_first, unsigned char *slots,
                                   int64_t slot_stride, const unsigned char *state_in, cudaStream_t st)
{
    if (n_streams <= 0) return cudaSuccess;
    PvFusedTables none;
    const CTables tb = make_ctables(d, none);
    const int entries = std::max(d.V * (d.N / 2 + 1), 1);
    const dim3 grid((unsigned)n_streams, (unsigned)((entries + 255) / 256));
    split_states_kernel<<<grid, 256, 0, st>>>(d, tb, parts, proc_segs, reinterpret_cast<const long long *>(S),
                                                               reinterpret_cast<const long long *>(H), P_first, slots, slot_stride,
                                                               state_in);
    return cudaGetLastError();
}

cudaError_t pv_launch_state_from_carry(const PvDev &d, const PvFusedTables &t, int64_t n_streams, const uint32_t *P_first,
                                       const int64_t *sumD, int64_t n_before, const uint32_t *P_prev, void *state,
                                       int64_t state_stride, cudaStream_t st)
{
    if (n_streams <= 0) return cudaSuccess;
    const CTables tb = make_ctables(d, t);
    state_from_carry_kernel<<<(unsigned)n_streams, 256, 0, st>>>(d, tb, n_streams, P_first,
                                                                reinterpret_cast<const long long *>(sumD), n_before, P_prev,
                                                                reinterpret_cast<unsigned char *>(state), state_stride);
    return cudaGetLastError();
}

// ---- frame-range sharding across ranks (pv_shard_begin / pv_shard_finish): the carry record of one stream is
// [nb int64 sums | nb uint32 phase of frame 0 (rank 0 only), packed into (nb+1)/2 int64] ----
__global__ void shard_pack_kernel(int nb, int elems, const long long *total, const long long *minus, const uint32_t *P0,
                                  long long *carry)
{
    const long long s = blockIdx.x;
    long long *rec = carry + s * elems;
    uint32_t *recP = reinterpret_cast<uint32_t *>(rec + nb);
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        rec[b] = (total ? total[s * nb + b] : 0) - (minus ? minus[s * nb + b] : 0);
        recP[b] = P0 ? P0[s * nb + b] : 0u;
    }
    if (threadIdx.x == 0 && (nb & 1)) recP[nb] = 0u;
}

// prefix[s][b] = sum_{r < rank} carry_all[r][s].sums[b] - minus[s][b];  P0[s][b] = carry_all[0][s].P0[b]
__global__ void shard_prefix_kernel(int nb, int elems, int rank, long long n_streams, const long long *carry_all,
                                    const long long *minus, long long *prefix, uint32_t *P0)
{
    const long long s = blockIdx.x;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        long long acc = minus ? -minus[s * nb + b] : 0;
        for (int r = 0; r < rank; r++) acc += carry_all[((long long)r * n_streams + s) * elems + b];
        prefix[s * nb + b] = acc;
        P0[s * nb + b] = reinterpret_cast<const uint32_t *>(carry_all + s * elems + nb)[b];
    }
}

cudaError_t pv_launch_shard_pack(int nb, int elems, int64_t n_streams, const int64_t *total, const int64_t *minus,
                                 const uint32_t *P0, int64_t *carry, cudaStream_t st)
{
    if (n_streams <= 0) return cudaSuccess;
    shard_pack_kernel<<<(unsigned)n_streams, 256, 0, st>>>(nb, elems, reinterpret_cast<const long long *>(total),
                                                          reinterpret_cast<const long long *>(minus), P0,
                                                          reinterpret_cast<long long *>(carry));
    return cudaGetLastError();
}

cudaError_t pv_launch_shard_prefix(int nb, int elems, int rank, int64_t n_streams, const int64_t *carry_all,
                                   const int64_t *minus, int64_t *prefix, uint32_t *P0, cudaStream_t st)
{
    if (n_streams <= 0) return cudaSuccess;
    shard_prefix_kernel<<<(unsigned)n_streams, 256, 0, st>>>(nb, elems, rank, (long long)n_streams,
                                                            reinterpret_cast<const long long *>(carry_all),
                                                            reinterpret_cast<const long long *>(minus),
                                                            reinterpret_cast<long long *>(prefix), P0);
    return cudaGetLastError();
}

// Voices per launch.  Every voice adds a phase-accumulator array and an overlap-add ring to the group's shared memory, and
// at three or four voices that costs resident CTAs (window 256: 4 -> 2 per SM).  Two launches of two voices each -- the
// forward transform is computed twice -- measured faster than one launch of four (tools/voices_probe.py).
static int voices_per_launch(const PvDev &d)
{
    static const bool no_split = getenv("PV_VOICES_ONE_LAUNCH") != nullptr;      // A/B switch, read once
    return (d.V >= 3 && !no_split) ? 2 : d.V;
}

cudaError_t pv_launch_corrected_fused(const PvDev &d, const PvFusedTables &t, const PvProcessArgs &a0, cudaStream_t st)
{
    if (a0.n_segs <= 0) return cudaSuccess;
    // a call of a few frames (a real-time block) is launch-bound: one launch for all voices then (C4 step: 43 vs 49 us)
    const int per = a0.n_frames >= 16 ? voices_per_launch(d) : d.V;
    for (int v0 = 0; v0 < d.V; v0 += per) {
        const CTables tb = make_ctables(d, t, v0, std::min(per, d.V - v0));
        PvProcessArgs a = a0;
        a.out = a0.out + (long long)v0 * a0.out_voice_stride;
        cudaError_t e;
        switch (d.N) {
            case 256: e = claunch<8, 4>(d, tb, a, st); break;
            case 512: e = claunch<9, 4>(d, tb, a, st); break;
            case 1024: e = claunch<10, 4>(d, tb, a, st); break;
            case 2048: e = claunch<11, 4>(d, tb, a, st); break;
            case 4096: e = claunch<12, 2>(d, tb, a, st); break;
            default: return cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
