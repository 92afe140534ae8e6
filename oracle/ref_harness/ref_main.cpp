// ref_main.cpp -- drives the UNMODIFIED reference pipeline (karnel/kernel.cu, karnel/hpfft.cu,
// karnel/common.cu, src/phaseVocoder.cpp, src/io.cpp compiled where they lie under /root/reference)
// the way src/main.cpp:204-297 does, minus AudioFile / RtAudio / hard-coded paths.
// TEST INFRASTRUCTURE: pins oracle/pv_oracle.c to the reference's own float output on a GPU and
// gives the "reference cuFFT build" timing.  Built only into oracle/_ref/ (git-ignored).
//
//   pv_ref_harness in.f32 out.f32 N hop_divisor [spectra.f32 n_spectra]
//
// in.f32 : raw float32 samples of one channel; out.f32: numSamples/outHop * outHop samples;
// spectra.f32: the first n_spectra analysis buffers, float2[2N] {mag, phase} each.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "phaseVocoder.h"

static void ck(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); exit(2); }
}

int main(int argc, char **argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s in.f32 out.f32 N hop_divisor [spectra.f32 n]\n", argv[0]); return 1; }
    const int N = atoi(argv[3]), hopdiv = atoi(argv[4]);
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) { perror(argv[1]); return 1; }
    fseek(fi, 0, SEEK_END);
    const long numSamples = ftell(fi) / 4;
    fseek(fi, 0, SEEK_SET);
    std::vector<float> x(numSamples);
    if (fread(x.data(), 4, numSamples, fi) != (size_t)numSamples) return 1;
    fclose(fi);

    PhaseVocoder *phase = new PhaseVocoder(N, TIME_SHIFT, 1, hopdiv);            // src/main.cpp:84
    const int hop = phase->hopSize, outHop = phase->outHopSize;

    // src/main.cpp:145-155: managed copy of the channel, padded so that the last frames' reads past the
    // end see zeros (the reference reads whatever follows its allocation)
    float *d_input;
    cudaMallocManaged((void **)&d_input, (numSamples + N) * sizeof(float), cudaMemAttachHost);
    ck("input");
    for (long j = 0; j < numSamples + N; j++) d_input[j] = j < numSamples ? x[j] : 0.f;

    // src/main.cpp:204-219: one zeroed float2[2N] buffer per frame
    const long nbuf = (numSamples + hop - 1) / hop;
    std::vector<float2 *> d_output(nbuf);
    for (long j = 0; j < nbuf; j++) {
        cudaMalloc((void **)&d_output[j], sizeof(float2) * 2 * N);
        cudaMemset(d_output[j], 0, sizeof(float2) * 2 * N);
    }
    ck("frames");
    float *intermediary;
    cudaMalloc((void **)&intermediary, sizeof(float) * N);
    float2 *fft;
    cudaMallocManaged((void **)&fft, sizeof(float2) * 2 * N, cudaMemAttachGlobal);
    ck("scratch");

    // src/main.cpp:228-250
    cudaStreamAttachMemAsync(NULL, d_input, 0, cudaMemAttachGlobal);
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    long nA = 0;
    for (long i = 0; i < numSamples - hop; i += hop, nA++) {
        cudaStreamSynchronize(NULL);
        phase->analysis_CUFFT(&d_input[i], d_output[i / hop], fft, intermediary);
    }
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();

    if (argc >= 7) {
        const int ns = atoi(argv[6]);
        std::vector<float2> h(2 * N);
        FILE *fs = fopen(argv[5], "wb");
        for (int k = 0; k < ns && k < nA; k++) {
            cudaMemcpy(h.data(), d_output[k], sizeof(float2) * 2 * N, cudaMemcpyDeviceToHost);
            fwrite(h.data(), sizeof(float2), 2 * N, fs);
        }
        fclose(fs);
    }

    // src/main.cpp:253-297
    float *backFrame;
    cudaMallocManaged((void **)&backFrame, sizeof(float) * N, cudaMemAttachHost);
    for (int i = 0; i < N; i++) backFrame[i] = 0;
    const long nS = numSamples / outHop;
    std::vector<float> out(nS * outHop, 0.f);
    auto t2 = std::chrono::steady_clock::now();
    for (long i = 0; i < nS; i++) {
        float *final_output;
        cudaMallocManaged((void **)&final_output, sizeof(float) * N, cudaMemAttachHost);
        phase->resynthesis_CUFFT(backFrame, d_output[i], final_output);
        cudaMemcpy(backFrame, final_output, sizeof(float) * N, cudaMemcpyHostToHost);
        cudaFree(final_output);
        for (int j = 0; j < outHop; j++) out[i * outHop + j] = backFrame[j];
    }
    cudaDeviceSynchronize();
    auto t3 = std::chrono::steady_clock::now();
    ck("resynthesis");

    FILE *fo = fopen(argv[2], "wb");
    fwrite(out.data(), 4, out.size(), fo);
    fclose(fo);
    const double ta = std::chrono::duration<double>(t1 - t0).count(), ts = std::chrono::duration<double>(t3 - t2).count();
    printf("{\"N\": %d, \"hop\": %d, \"frames_analysed\": %ld, \"frames_synth\": %ld, \"analysis_s\": %.6f, "
           "\"resynthesis_s\": %.6f, \"frames_per_s\": %.1f}\n", N, hop, nA, nS, ta, ts, (double)nS / (ta + ts));
    return 0;
}
