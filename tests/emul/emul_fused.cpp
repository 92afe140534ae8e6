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
        auto sync = [&]() { bar.arrive_and_wait(); };
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
                    for (int j = tid; j < Hs; j += T) out[(k - 1) * (long)Hs + j] = acc[(pp + j) & (N - 1)];
                }
            };
            cp_async_wait_all();
            constexpr bool TWREG = (S::S1 == S::T) && (S::R1 == 16);
            const ThreadTw tt = load_thread_tw<LOG2N>(tid, tb);
            frame_compat<LOG2N, TWREG>(tid, io, tb, tt, nan_compat != 0, ring, bufA.data(), bufB.data(), acc.data(),
                                       pos0, Hs, sync, hook);
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
        default: return -1;
    }
}
