// Test helper for phase-vocoder_b200/host/pv_wav.h: decode a WAV to raw float32 (channel-major) or
// encode raw float32 to a 16-bit WAV, so that the Python tests can compare the C++ codec with the
// numpy restatement of AudioFile's rules (oracle/wav_oracle.py).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../phase-vocoder_b200/host/pv_wav.h"

int main(int argc, char **argv)
{
    std::string err;
    if (argc == 4 && !strcmp(argv[1], "decode")) {
        pvwav::Audio a;
        if (!pvwav::load(argv[2], a, err)) { fprintf(stderr, "%s\n", err.c_str()); return 3; }
        FILE *f = fopen(argv[3], "wb");
        for (auto &c : a.samples) fwrite(c.data(), 4, c.size(), f);
        fclose(f);
        printf("%zu %zu %u %d\n", a.samples.size(), a.samples[0].size(), a.sample_rate, a.bit_depth);
        return 0;
    }
    if (argc == 5 && !strcmp(argv[1], "encode")) {
        const int ch = atoi(argv[3]);
        FILE *f = fopen(argv[2], "rb");
        fseek(f, 0, SEEK_END);
        const size_t n = (size_t)ftell(f) / 4 / ch;
        fseek(f, 0, SEEK_SET);
        pvwav::Audio a;
        a.samples.assign(ch, std::vector<float>(n));
        for (auto &c : a.samples)
            if (fread(c.data(), 4, n, f) != n) return 2;
        fclose(f);
        return pvwav::save16(argv[4], a, err) ? 0 : 3;
    }
    fprintf(stderr, "usage: wav_tool decode in.wav out.f32 | encode in.f32 channels out.wav\n");
    return 1;
}
