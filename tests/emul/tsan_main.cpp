// tsan_main.cpp -- race check of the fused kernel bodies WITHOUT a GPU tool (compute-sanitizer is closed on this
// pool).  The emulation (emul_fused.cpp) runs pv_fused_core.cuh / pv_fused_corrected.cuh with one std::thread per CUDA
// thread and std::barrier as the group barrier; built with -fsanitize=thread, every pair of conflicting shared-buffer
// accesses (exchange buffers incl. the in-place exchange with its elided barriers, the input ring refilled one frame
// ahead, the overlap-add ring with the deferred emit, mag / D / psi arrays) that is not ordered by a barrier is
// reported as a data race.  Exit code 0 and no "WARNING: ThreadSanitizer" = no missing barrier in the frame bodies.
// TEST INFRASTRUCTURE ONLY (links the oracle for the corrected-mode tables).
#include <cmath>
#include <cstdio>
#include <vector>

#include "emul_fused.cpp"
extern "C" {
#include "../../oracle/pv_oracle.h"
}

int main(int argc, char **argv)
{
    const int only = argc > 1 ? atoi(argv[1]) : 0;
    if (argc > 2) emul_drop_barrier(atoi(argv[2]));      // self test: drop one barrier, ThreadSanitizer must complain
    int rc = 0;
    for (int lg : {8, 9, 10, 11}) {
        if (only && lg != only) continue;
        const int N = 1 << lg, H = N / 4, nf = lg >= 10 ? 5 : 8;
        const long n_in = N + (long)(nf - 1) * H - 3;
        std::vector<float> x((size_t)n_in), win((size_t)N), out((size_t)nf * H * 2);
        for (long i = 0; i < n_in; i++) x[(size_t)i] = 0.3f * sinf(0.05f * (float)i) + 0.1f * sinf(0.31f * (float)i + 1.f);
        pvo_window(PVO_WIN_HAMMING, N, win.data());
        rc |= emul_compat(lg, x.data(), n_in, H, H, win.data(), nf - 1, nf, 0, out.data());
        const int nb = N / 2 + 1, V = 2;
        const double beta[2] = {1.5, 0.8};
        std::vector<int32_t> alo((size_t)V * nb), ahi((size_t)V * nb);
        std::vector<uint64_t> nomS((size_t)V * nb);
        std::vector<uint32_t> nomA((size_t)nb);
        uint64_t bq[2], Rq[2];
        pvo_window(PVO_WIN_HANN_PERIODIC, N, win.data());
        for (int v = 0; v < V; v++)
            pvo_corrected_tables(N, H, H, beta[v], &bq[v], &Rq[v], &alo[(size_t)v * nb], &ahi[(size_t)v * nb], &nomS[(size_t)v * nb],
                                 v == 0 ? nomA.data() : nullptr);
        rc |= emul_corrected(lg, x.data(), n_in, H, H, win.data(), V, nomA.data(), alo.data(), ahi.data(),
                             (const unsigned long long *)nomS.data(), (const unsigned long long *)Rq,
                             (const unsigned long long *)bq, pvo_corrected_gain(win.data(), N, H), nf, out.data(), (long)nf * H);
        // the stored-analysis split: analysis pass that stores, then the processing pass with its own barrier structure
        emul_stored_analysis(1);
        rc |= emul_corrected(lg, x.data(), n_in, H, H, win.data(), V, nomA.data(), alo.data(), ahi.data(),
                             (const unsigned long long *)nomS.data(), (const unsigned long long *)Rq,
                             (const unsigned long long *)bq, pvo_corrected_gain(win.data(), N, H), nf, out.data(), (long)nf * H);
        emul_stored_analysis(0);
        printf("window %d: compat + corrected (2 voices) + stored-analysis split x %d frames emulated with %d threads, rc=%d\n", N, nf, N / 16, rc);
    }
    return rc;
}
