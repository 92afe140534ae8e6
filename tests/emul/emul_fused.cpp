// emul_fused.cpp -- CPU execution of the fused kernel body (pv_fused_core.cuh) with one
// std::thread per CUDA thread and std::barrier as __syncthreads.  TEST HARNESS ONLY: it lets
// the index algebra of the register-blocked kernel be checked against the oracle without a GPU.
#define PV_HOST_EMUL 1
#include <barrier>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "../../phase-vocoder_b200/csrc/pv_fused_core.cuh"
#include "../../phase-vocoder_b200/csrc/pv_fused_tables.h"

using namespace pvfused;

// Race-check self test (tsan_main.cpp): every thread skips its g_drop_barrier-th barrier (all threads skip the same
// one, so the barrier phases stay aligned).  -1 = run normally.  ThreadSanitizer must then report a race.
static int g_drop_barrier = -1;
extern "C" void emul_drop_barrier(int k) { g_drop_barrier = k; }
// Stored-analysis split (frame_corrected MODE 2 then MODE 1): emul_corrected runs the analysis-only pass over the whole stream,
// keeping {|X|, D} of every frame, and then the processing pass that synthesises from the stored rows (double-buffered, the next
// row copied by thread 0 where the kernel issues its bulk copy).  Same output as the normal frame, bit for bit.
static int g_stored = 0;
extern "C" void emul_stored_analysis(int on) { g_stored = on; }

template <int LOG2N>
static int run(const float *x, long n_in, int Ha, int Hs, const float *win, long n_analysed, long n_frames,
               int nan_compat, float *out)
{
    using S = Shape<LOG2N>;
    constexpr int N = S::N, T = S::T;
    HostTables ht;
    build_tables(LOG2N, ht);
    Tables tb{ht.tw1.data(), ht.tw2.data(), ht.tw2n.data(), ht.itw1.data(), ht.itw2.data(), win};
    std::vector<float2> bufA(S::BUF_A), bufB(S::BUF_B);
    std::vector<float> acc(N, 0.f), ringbuf(N, 0.f);
    std::barrier bar(T);
    const bool use_ring = (Ha % (2 * S::S1)) == 0;
    float *ring = use_ring ? ringbuf.data() : nullptr;
    auto body = [&](int tid) {
        int nbar = 0;
        auto sync = [&]() { if (nbar++ != g_drop_barrier) bar.arrive_and_wait(); };
        int pos0 = 0;
        if (use_ring && n_analysed > 0) {
            FrameIO io0{x, n_in, 0, true, (Ha % 2) == 0};
            ring_prefetch<LOG2N>(tid, io0, ring, 0);
        }
        for (long k = 0; k < n_frames; k++) {
            FrameIO io{x, n_in, k * (long long)Ha, k < n_analysed, (Ha % 2) == 0};
            auto hook = [&]() {
                if (use_ring && k + 1 < n_analysed) {
                    FrameIO nx{x, n_in, (k + 1) * (long long)Ha, true, true};
                    ring_prefetch<LOG2N>(tid, nx, ring, N - Ha);
                }
                if (k > 0) {
                    const int pp = (pos0 - Hs) & (N - 1);
                    for (int j = tid; j < Hs; j += T) {
                        out[(k - 1) * (long)Hs + j] = acc[(pp + j) & (N - 1)];
                        acc[(pp + j) & (N - 1)] = 0.f;         // emitted hop becomes the fresh tail
                    }
                }
            };
            cp_async_wait_all();
            constexpr bool TWREG = (S::S1 == S::T) && (S::R1 == 16);
            const ThreadTw tt = load_thread_tw<LOG2N>(tid, tb);
            frame_compat<LOG2N, TWREG>(tid, io, tb, tt, nan_compat != 0, ring, bufA.data(), bufB.data(), acc.data(),
                                       pos0, Hs, sync, hook, []() {});
            pos0 = (pos0 + Hs) & (N - 1);
        }
        sync();
        const int pp = (pos0 - Hs) & (N - 1);
        for (int j = tid; j < Hs; j += T) out[(n_frames - 1) * (long)Hs + j] = acc[(pp + j) & (N - 1)];
    };
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(body, t);
    for (auto &t : th) t.join();
    return 0;
}

extern "C" int emul_compat(int log2n, const float *x, long n_in, int Ha, int Hs, const float *win, long n_analysed,
                           long n_frames, int nan_compat, float *out)
{
    switch (log2n) {
        case 8: return run<8>(x, n_in, Ha, Hs, win, n_analysed, n_frames, nan_compat, out);
        case 9: return run<9>(x, n_in, Ha, Hs, win, n_analysed, n_frames, nan_compat, out);
        case 10: return run<10>(x, n_in, Ha, Hs, win, n_analysed, n_frames, nan_compat, out);
        case 11: return run<11>(x, n_in, Ha, Hs, win, n_analysed, n_frames, nan_compat, out);
        case 12: return run<12>(x, n_in, Ha, Hs, win, n_analysed, n_frames, nan_compat, out);
        default: return -1;
    }
}

// ---------------- corrected mode ----------------
#include "../../phase-vocoder_b200/csrc/pv_fused_corrected.cuh"

template <int LOG2N>
static int run_corrected(const float *x, long n_in, int Ha, int Hs, const float *win, int V, const uint32_t *nomA,
                         const int32_t *a_lo, const int32_t *a_hi, const unsigned long long *nomS,
                         const unsigned long long *Rq, const unsigned long long *beta_q, float gain, long n_frames,
                         float *out, long out_stride)
{
    using C = CShape<LOG2N>;
    constexpr int N = C::N, T = C::T, NB = C::NB;
    HostTables ht;
    build_tables(LOG2N, ht);
    CTables tb{};
    tb.ctw1 = ht.ctw1.data(); tb.ctw2 = ht.ctw2.data(); tb.tw2n = ht.tw2n.data();
    tb.itw1 = ht.itw1.data(); tb.itw2 = ht.itw2.data(); tb.win = win;

    tb.nomA = nomA; tb.a_lo = a_lo; tb.a_hi = a_hi; tb.nomS = nomS;
    for (int v = 0; v < V; v++) tb.Rq[v] = Rq[v];
    tb.scale = gain / (float)N;
    tb.V = V;
    tb.Ha = Ha;
    std::vector<uint32_t> gath;
    int32_t multi[8] = {};
    build_gather_table(N, V, a_lo, a_hi, gath, multi);
    tb.gather = gath.data();
    for (int v = 0; v < V; v++) {
        tb.beta_q[v] = beta_q[v];
        tb.bqs[v] = (beta_q[v] * (unsigned long long)Hs) << (32 - LOG2N);
        tb.multi[v] = multi[v];
    }
    std::vector<float2> bufA(C::BUF_A), bufB(C::BUF_B);
    std::vector<float> acc((size_t)V * N, 0.f), ringbuf(N, 0.f);
    std::vector<float2> mdS(NB + 3, make_float2(0.f, 0.f));
    std::vector<unsigned long long> psi((size_t)V * NB, 0ull);
    std::barrier bar(T);
    const bool use_ring = (Ha % 2) == 0 && Ha <= N;
    float *ring = use_ring ? ringbuf.data() : nullptr;
    constexpr int NBP = NB + 1;
    std::vector<float2> md_all(g_stored ? (size_t)n_frames * NBP : 0), md_alt(NBP);
    std::vector<long long> sums(NB, 0);
    auto body_stored = [&](int tid) {
        int nbar = 0;
        auto sync = [&]() { if (nbar++ != g_drop_barrier) bar.arrive_and_wait(); };
        const CThreadTw tt = load_cthread_tw<LOG2N>(tid, tb);
        {   // pass 1: analysis only, storing
            CState st{};
            if (use_ring) {
                FrameIO io0{x, n_in, 0, true, true};
                ring_prefetch_coop<N, T>(tid, io0, ring, 0);
                cp_async_wait_all();
            }
            sync();
            for (long k = 0; k < n_frames; k++) {
                FrameIO io{x, n_in, k * (long long)Ha, true, (Ha % 2) == 0};
                auto hook = [&]() {
                    if (use_ring && k + 1 < n_frames) {
                        FrameIO nx{x, n_in, (k + 1) * (long long)Ha, true, true};
                        ring_prefetch_coop<N, T>(tid, nx, ring, N - Ha);
                    }
                };
                const AggCtx ac{true, sums.data(), nullptr, md_all.data() + (size_t)k * NBP};
                frame_corrected<LOG2N, 2>(tid, io, tb, tt, ring, bufA.data(), bufB.data(), mdS.data(), psi.data(), acc.data(), st, 0,
                                          Hs, sync, hook, [&]() { cp_async_wait_all(); }, ac);
                if (use_ring) { cp_async_wait_all(); sync(); }
            }
            sync();
        }
        // pass 2: processing from the stored rows
        CState st{};
        int pos0 = 0;
        auto md_buf = [&](long k) { return (k & 1) ? md_alt.data() : mdS.data(); };
        if (tid == 0) std::memcpy(md_buf(0), md_all.data(), sizeof(float2) * NBP);
        sync();
        for (long k = 0; k < n_frames; k++) {
            FrameIO io{x, n_in, k * (long long)Ha, true, (Ha % 2) == 0};
            auto hook = [&]() {
                if (k + 1 < n_frames && tid == 0) std::memcpy(md_buf(k + 1), md_all.data() + (size_t)(k + 1) * NBP, sizeof(float2) * NBP);
                if (k > 0) {
                    const int pp = (pos0 - Hs) & (N - 1);
                    for (int v = 0; v < V; v++)
                        for (int j = tid; j < Hs; j += T) {
                            out[v * out_stride + (k - 1) * (long)Hs + j] = acc[(size_t)v * N + ((pp + j) & (N - 1))];
                            acc[(size_t)v * N + ((pp + j) & (N - 1))] = 0.f;
                        }
                }
            };
            frame_corrected<LOG2N, 1>(tid, io, tb, tt, nullptr, bufA.data(), bufB.data(), md_buf(k), psi.data(), acc.data(), st, pos0, Hs,
                                      sync, hook, []() {});
            pos0 = (pos0 + Hs) & (N - 1);
        }
        sync();
        const int pp = (pos0 - Hs) & (N - 1);
        for (int v = 0; v < V; v++)
            for (int j = tid; j < Hs; j += T)
                out[v * out_stride + (n_frames - 1) * (long)Hs + j] = acc[(size_t)v * N + ((pp + j) & (N - 1))];
    };
    auto body = [&](int tid) {
        int nbar = 0;
        auto sync = [&]() { if (nbar++ != g_drop_barrier) bar.arrive_and_wait(); };
        CState st{};
        const CThreadTw tt = load_cthread_tw<LOG2N>(tid, tb);
        int pos0 = 0;
        if (use_ring) {
            FrameIO io0{x, n_in, 0, true, true};
            ring_prefetch_coop<N, T>(tid, io0, ring, 0);
            cp_async_wait_all();
        }
        sync();
        for (long k = 0; k < n_frames; k++) {
            FrameIO io{x, n_in, k * (long long)Ha, true, (Ha % 2) == 0};
            auto hook = [&]() {
                if (use_ring && k + 1 < n_frames) {
                    FrameIO nx{x, n_in, (k + 1) * (long long)Ha, true, true};
                    ring_prefetch_coop<N, T>(tid, nx, ring, N - Ha);
                }
                if (k > 0) {
                    const int pp = (pos0 - Hs) & (N - 1);
                    for (int v = 0; v < V; v++)
                        for (int j = tid; j < Hs; j += T) {
                            out[v * out_stride + (k - 1) * (long)Hs + j] = acc[(size_t)v * N + ((pp + j) & (N - 1))];
                            acc[(size_t)v * N + ((pp + j) & (N - 1))] = 0.f;
                        }
                }
            };
            frame_corrected<LOG2N>(tid, io, tb, tt, ring, bufA.data(), bufB.data(), mdS.data(), psi.data(),
                                   acc.data(), st, pos0, Hs, sync, hook, [&]() { cp_async_wait_all(); });
            pos0 = (pos0 + Hs) & (N - 1);
        }
        sync();
        const int pp = (pos0 - Hs) & (N - 1);
        for (int v = 0; v < V; v++)
            for (int j = tid; j < Hs; j += T)
                out[v * out_stride + (n_frames - 1) * (long)Hs + j] = acc[(size_t)v * N + ((pp + j) & (N - 1))];
    };
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) {
        if (g_stored) th.emplace_back(body_stored, t);
        else th.emplace_back(body, t);
    }
    for (auto &t : th) t.join();
    return 0;
}

extern "C" int emul_corrected(int log2n, const float *x, long n_in, int Ha, int Hs, const float *win, int V,
                              const uint32_t *nomA, const int32_t *a_lo, const int32_t *a_hi,
                              const unsigned long long *nomS, const unsigned long long *Rq,
                              const unsigned long long *beta_q, float gain, long n_frames, float *out, long out_stride)
{
#define RC(L) case L: return run_corrected<L>(x, n_in, Ha, Hs, win, V, nomA, a_lo, a_hi, nomS, Rq, beta_q, gain, n_frames, out, out_stride)
    switch (log2n) {
        RC(8); RC(9); RC(10); RC(11); RC(12);
        default: return -1;
    }
#undef RC
}
