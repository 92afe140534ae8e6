// audiofile_tool.cpp -- TEST INFRASTRUCTURE ONLY.  Drives the reference's OWN, unmodified WAV codec
// (src/AudioFile.h, included from where it lies under the reference checkout: -I $(REF)/src) so that the
// product's codec (phase-vocoder_b200/host/pv_wav.h) can be compared with it byte for byte
// (tests/test_wav_and_cli.py).  Built into git-ignored oracle/_ref/ by oracle/ref_harness/Makefile.
//
//   audiofile_tool decode in.wav out.f32       AudioFile<float>::load (AudioFile.h:382-416, decodeWaveFile :418-530);
//                                              prints "channels samples rate bits"; exit 3 when load() fails
//   audiofile_tool encode in.f32 ch out.wav    setAudioBuffer + 16 bit / 44.1 kHz + save (saveToWaveFile :703-785),
//                                              the calls of src/main.cpp:140-143,309
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "AudioFile.h"

int main(int argc, char **argv)
{
    if (argc == 4 && !strcmp(argv[1], "decode")) {
        AudioFile<float> a;
        if (!a.load(argv[2])) return 3;
        FILE *f = fopen(argv[3], "wb");
        if (!f) return 2;
        for (int c = 0; c < a.getNumChannels(); c++) fwrite(a.samples[c].data(), 4, a.samples[c].size(), f);
        fclose(f);
        printf("%d %d %u %d\n", a.getNumChannels(), a.getNumSamplesPerChannel(), a.getSampleRate(), a.getBitDepth());
        return 0;
    }
    if (argc == 5 && !strcmp(argv[1], "encode")) {
        const int ch = atoi(argv[3]);
        FILE *f = fopen(argv[2], "rb");
        if (!f || ch < 1) return 2;
        fseek(f, 0, SEEK_END);
        const size_t n = (size_t)ftell(f) / 4 / (size_t)ch;
        fseek(f, 0, SEEK_SET);
        AudioFile<float>::AudioBuffer buf((size_t)ch, std::vector<float>(n));
        for (auto &c : buf)
            if (fread(c.data(), 4, n, f) != n) return 2;
        fclose(f);
        AudioFile<float> a;
        a.setAudioBuffer(buf);
        a.setBitDepth(16);
        a.setSampleRate(44100);
        return a.save(argv[4]) ? 0 : 3;
    }
    fprintf(stderr, "usage: audiofile_tool decode in.wav out.f32 | encode in.f32 channels out.wav\n");
    return 1;
}
