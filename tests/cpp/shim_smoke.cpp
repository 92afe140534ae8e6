// Compiles the C++ shim the way the reference's driver uses it (src/main.cpp:84,234,271) and, on a
// GPU box, runs one frame through it.  Without a device the constructor must exit(EXIT_FAILURE)
// with the reference's error line (no CPU fallback).
#include <cmath>
#include <cstring>
#include <vector>

#include "../../phase-vocoder_b200/host/phaseVocoder.h"

int main(int argc, char** argv)
{
    const bool run = argc > 1 && !strcmp(argv[1], "run");
    PhaseVocoder* phase = new PhaseVocoder(256, TIME_SHIFT, 1, 2);      // src/main.cpp:84
    if (phase->nSamps != 256 || phase->hopSize != 128 || phase->outHopSize != 128) return 2;
    if (std::fabs(phase->imp[0] - 0.08f) > 1e-6f) return 3;              // Hamming end point
    // the reference's public fields and stream getters (src/phaseVocoder.h:16-32, 118-126) exist with the same meaning
    if (phase->plan != 0 || phase->ifft != 0 || phase->curr_input != nullptr || phase->R != 1) return 5;
    cudaStream_t* s0 = phase->getStream();
    if (s0 != &phase->streams[0] || phase->getPrevStream() != &phase->streams[0] || phase->getStream() != &phase->streams[1] ||
        phase->getStream() != &phase->streams[2] || phase->getStream() != &phase->streams[0]) return 6;
    if (run) {
        std::vector<float> x(256 * 40), y(256 * 40, 0.f);
        for (size_t i = 0; i < x.size(); i++) x[i] = 0.25f * std::sin(0.05f * (float)i);
        const long n = phase->process(x.data(), (long)x.size(), y.data());
        double e = 0;
        for (long i = 0; i < n; i++) e += (double)y[i] * y[i];
        printf("shim ok: %ld samples, energy %.6f\n", n, e);
        if (!(e > 0)) return 4;
    }
    delete phase;
    return 0;
}
