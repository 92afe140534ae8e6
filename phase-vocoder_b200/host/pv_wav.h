// pv_wav.h -- minimal WAV reader/writer with the sample conversion rules of the reference's
// AudioFile<float> (src/AudioFile.h), written from scratch:
//   decode  chunk discovery by the FIRST occurrence of "data"/"fmt" (getIndexOfString :1017-1034),
//           PCM only (:454), mono/stereo only (:461), 8/16/24 bit; 16-bit: s/32768 (:1038-1042),
//           24-bit: sign-extend, /8388608 (:508-518), 8-bit: (s-128)/128 (:1065-1068)
//   encode  44-byte header (:703-745), 16-bit: (int16) trunc(clamp(x,-1,1)*32767) (:1045-1049)
// A file shorter than its header claims is zero-filled (the reference reads past its buffer).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace pvwav {

struct Audio {
    std::vector<std::vector<float>> samples;   // [channel][n]
    uint32_t sample_rate = 44100;
    int bit_depth = 16;
};

inline long find(const std::vector<uint8_t> &d, const char *tag)
{
    const size_t n = strlen(tag);
    for (size_t i = 0; i + n <= d.size(); i++)
        if (!memcmp(&d[i], tag, n)) return (long)i;
    return -1;
}
// Little-endian field readers; the callers check i + width <= d.size() first.
inline uint32_t rdu32(const std::vector<uint8_t> &d, size_t i) { return (uint32_t)d[i] | ((uint32_t)d[i + 1] << 8) | ((uint32_t)d[i + 2] << 16) | ((uint32_t)d[i + 3] << 24); }
inline int16_t rd16(const std::vector<uint8_t> &d, size_t i) { return (int16_t)(uint16_t)(d[i] | (d[i + 1] << 8)); }

// Decodes a whole WAV image.  The file is NOT trusted: every header field is bounds-checked before it is
// read, the data size is an unsigned 32-bit field (streaming writers store 0xFFFFFFFF), and the sample count
// is clamped to what the file can hold plus a bounded zero fill (kZeroFillMax bytes: the reference's own test
// tones are 2 bytes shorter than their header claims, testtones/440sine.wav), so a small file cannot make the
// loader allocate gigabytes.
constexpr size_t kZeroFillMax = 4096;

inline bool decode(const std::vector<uint8_t> &d, Audio &a, std::string &err)
{
    if (d.size() < 44 || memcmp(&d[0], "RIFF", 4) || memcmp(&d[8], "WAVE", 4)) { err = "not a RIFF/WAVE file"; return false; }
    const long dl = find(d, "data"), fl = find(d, "fmt");
    if (dl < 0 || fl < 0) { err = "missing fmt/data chunk"; return false; }
    const size_t di = (size_t)dl, fi = (size_t)fl;
    if (fi + 24 > d.size() || di + 8 > d.size()) { err = "truncated fmt/data chunk header"; return false; }
    const int fmt = rd16(d, fi + 8), ch = rd16(d, fi + 10);
    a.sample_rate = rdu32(d, fi + 12);
    const uint32_t bps = rdu32(d, fi + 16);
    const int block = rd16(d, fi + 20);
    a.bit_depth = rd16(d, fi + 22);
    if (fmt != 1) { err = "not PCM"; return false; }
    if (ch < 1 || ch > 2) { err = "neither mono nor stereo"; return false; }
    if (a.bit_depth != 8 && a.bit_depth != 16 && a.bit_depth != 24) { err = "unsupported bit depth"; return false; }
    const int nbytes = a.bit_depth / 8;
    if ((uint64_t)bps != (uint64_t)ch * a.sample_rate * (uint64_t)a.bit_depth / 8 || block != ch * nbytes) { err = "inconsistent header"; return false; }
    const size_t start = di + 8;
    const size_t avail = d.size() - start;                       // payload bytes actually in the file
    size_t count = (size_t)rdu32(d, di + 4) / (size_t)block;     // frames the header claims
    const size_t max_count = (avail + kZeroFillMax) / (size_t)block;
    if (count > max_count) { err = "data chunk larger than the file"; return false; }
    a.samples.assign((size_t)ch, std::vector<float>(count));
    auto byte_at = [&](size_t p) -> uint32_t { return p < d.size() ? d[p] : 0u; };   // zero fill past the end
    for (size_t i = 0; i < count; i++)
        for (int c = 0; c < ch; c++) {
            const size_t p = start + i * (size_t)block + (size_t)c * (size_t)nbytes;
            float v;
            if (a.bit_depth == 8) v = (float)((int)byte_at(p) - 128) / 128.f;
            else if (a.bit_depth == 16) v = (float)(int16_t)(uint16_t)(byte_at(p) | (byte_at(p + 1) << 8)) / 32768.f;
            else {
                int32_t s = (int32_t)(byte_at(p) | (byte_at(p + 1) << 8) | (byte_at(p + 2) << 16));
                if (s & 0x800000) s |= ~0xFFFFFF;
                v = (float)s / 8388608.f;
            }
            a.samples[(size_t)c][i] = v;
        }
    return true;
}

inline bool load(const std::string &path, Audio &a, std::string &err)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    std::vector<uint8_t> d;
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + n);
    fclose(f);
    return decode(d, a, err);
}

inline int16_t to_s16(float x)
{
    x = x < 1.f ? x : 1.f;
    x = x > -1.f ? x : -1.f;
    return (int16_t)(x * 32767.);       // double product, truncation toward zero
}

inline bool save16(const std::string &path, const Audio &a, std::string &err)
{
    const int ch = (int)a.samples.size();
    const size_t n = ch ? a.samples[0].size() : 0;
    std::vector<uint8_t> d;
    auto s = [&](const char *t) { d.insert(d.end(), t, t + 4); };
    auto u32 = [&](uint32_t v) { for (int i = 0; i < 4; i++) d.push_back((uint8_t)(v >> (8 * i))); };
    auto u16 = [&](uint16_t v) { d.push_back((uint8_t)v); d.push_back((uint8_t)(v >> 8)); };
    const uint32_t bytes = (uint32_t)(n * ch * 2);
    s("RIFF"); u32(4 + 24 + 8 + bytes); s("WAVE"); s("fmt "); u32(16); u16(1); u16((uint16_t)ch); u32(a.sample_rate);
    u32(ch * a.sample_rate * 16 / 8); u16((uint16_t)(ch * 2)); u16(16); s("data"); u32(bytes);
    for (size_t i = 0; i < n; i++)
        for (int c = 0; c < ch; c++) u16((uint16_t)to_s16(a.samples[c][i]));
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    const bool ok = fwrite(d.data(), 1, d.size(), f) == d.size();
    fclose(f);
    if (!ok) err = "short write";
    return ok;
}

}  // namespace pvwav
