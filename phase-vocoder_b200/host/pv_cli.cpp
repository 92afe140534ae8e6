// pv_cli.cpp -- offline driver with the semantics of the reference's src/main.cpp (":60-81, 128-143,
// 204-309"), on top of the C++ shim / C ABI:
//
//   pv_cli <in.wav> [t|p] [out.wav] [--window N] [--hop-div D] [--scale S]
//
//   * positional arguments as in main.cpp:65-81 (wav, effect char, output name), minus the hard-coded
//     /home/davis prefixes; defaults window 256, hop divisor 2, scale 1 (main.cpp:84)
//   * channel 0 is processed (main.cpp:288-290 `if(numChannels = 1)` collapses the loop to one channel)
//     and duplicated into channel 1; output is always 2 channels, 44.1 kHz, 16 bit, timeScale*numSamples
//     samples long, samples after the last full hop stay 0 (main.cpp:140-143)
//   * 't': compat pipeline with outHopSize = scale*hopSize; 'p': corrected-mode pitch shift by `scale`
//
// WAV in and out: by default the built-in codec (pv_wav.h, checked byte for byte against the reference's AudioFile).  With
// -DPV_USE_AUDIOFILE -I<reference>/src the driver uses the reference's OWN, unmodified src/AudioFile.h exactly as
// src/main.cpp does (:128-143 load / setAudioBufferSize / setBitDepth / setSampleRate, :309 save): the third-party header
// is not vendored here; oracle/ref_harness/Makefile builds that variant into oracle/_ref/pv_cli_audiofile and
// tests/test_wav_and_cli.py checks that both variants write identical files.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "phaseVocoder.h"
#ifdef PV_USE_AUDIOFILE
#include "AudioFile.h"
#else
#include "pv_wav.h"
#endif

int main(int argc, char **argv)
{
    std::string in, out = "out.wav";
    Effect effect = TIME_SHIFT;
    int window = 256, hopdiv = 2, pos = 0;
    float scale = 1.f;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--window") && i + 1 < argc) window = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--hop-div") && i + 1 < argc) hopdiv = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--scale") && i + 1 < argc) scale = (float)atof(argv[++i]);
        else if (pos == 0) { in = argv[i]; pos++; }
        else if (pos == 1) { effect = static_cast<Effect>(*argv[i]); pos++; }      // main.cpp:75
        else if (pos == 2) { out = argv[i]; pos++; }
    }
    if (in.empty()) { fprintf(stderr, "usage: %s in.wav [t|p] [out.wav] [--window N] [--hop-div D] [--scale S]\n", argv[0]); return 1; }
    printf("Offline Vocoding\n");                                                    // main.cpp:126
#ifdef PV_USE_AUDIOFILE
    AudioFile<float> audioFile;                                                      // main.cpp:128
    if (!audioFile.load(in)) { printf("err: wav failed to load\n"); return 1; }      // main.cpp:130-134
    const long numSamples = audioFile.getNumSamplesPerChannel();
    PhaseVocoder phase(window, effect, scale, hopdiv);
    AudioFile<float> outFile;
    outFile.setAudioBufferSize(2, (int)(phase.timeScale * numSamples));              // main.cpp:140
    outFile.setBitDepth(16);                                                         // main.cpp:141
    outFile.setSampleRate(44100);                                                    // main.cpp:142
    printf("analysis...\nresynthesis...\n");
    std::vector<float> y((size_t)numSamples + (size_t)phase.outHopSize, 0.f);
    const long n = phase.process(audioFile.samples[0].data(), numSamples, y.data());
    for (long i = 0; i < n && i < (long)outFile.samples[0].size(); i++) outFile.samples[0][i] = outFile.samples[1][i] = y[i];
    printf("writing to file\n");                                                     // main.cpp:308
    if (!outFile.save(out)) return 1;                                                // main.cpp:309
#else
    pvwav::Audio a;
    std::string err;
    if (!pvwav::load(in, a, err)) { printf("err: wav failed to load\n"); fprintf(stderr, "%s\n", err.c_str()); return 1; }   // main.cpp:130-134
    const long numSamples = (long)a.samples[0].size();
    PhaseVocoder phase(window, effect, scale, hopdiv);
    pvwav::Audio o;
    o.sample_rate = 44100;                                                           // main.cpp:142
    o.samples.assign(2, std::vector<float>((size_t)(phase.timeScale * numSamples), 0.f));   // main.cpp:140
    printf("analysis...\nresynthesis...\n");
    std::vector<float> y((size_t)numSamples + (size_t)phase.outHopSize, 0.f);
    const long n = phase.process(a.samples[0].data(), numSamples, y.data());
    for (long i = 0; i < n && i < (long)o.samples[0].size(); i++) o.samples[0][i] = o.samples[1][i] = y[i];
    printf("writing to file\n");                                                     // main.cpp:308
    if (!pvwav::save16(out, o, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
#endif
    return 0;
}
