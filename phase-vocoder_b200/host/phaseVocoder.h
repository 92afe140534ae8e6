// phaseVocoder.h -- drop-in C++ mirror of the reference's `class PhaseVocoder`
// (src/phaseVocoder.h:9-139) on top of the C ABI (include/pv_b200.h).
//
// Same constructor arguments, method names, argument order and public fields as the reference, so
// that its driver code (src/main.cpp:84, :234, :271) compiles against this header unchanged:
//
//     PhaseVocoder* phase = new PhaseVocoder(256, effect, 1, 2);
//     phase->analysis_CUFFT(&d_input[ch][i], d_output[ch][i / phase->hopSize], fft, intermediary);
//     phase->resynthesis_CUFFT(backFrame, d_output[ch][i], final_output);
//
// Error behaviour follows checkCUDAError_ (src/io.cpp:115-124): print to stderr, exit(EXIT_FAILURE).
// New code should use process() -- both host loops of src/main.cpp:228-297 in one call.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/pv_b200.h"

#if defined(__CUDACC__) || __has_include(<vector_types.h>)
#include <vector_types.h>
#else
struct float2 { float x, y; };
#endif

// The reference's public stream fields (src/phaseVocoder.h:4, :31-32): plain g++ callers get the opaque runtime types.
// Compiled by nvcc (as the reference's driver is) or with -DPV_SHIM_WITH_CUDART the three streams are really created;
// a plain g++ build that does not link the CUDA runtime keeps them null.
#if defined(__CUDACC__) || defined(PV_SHIM_WITH_CUDART)
#include <cuda_runtime_api.h>
#define PV_SHIM_HAS_CUDART 1
#elif __has_include(<driver_types.h>)
#include <driver_types.h>
#else
typedef struct CUstream_st* cudaStream_t;
#endif
#ifndef _CUFFT_H_
typedef int cufftHandle;                                    // cufft.h: `typedef int cufftHandle;` -- no cuFFT is linked here
#endif
#define NUM_STREAMS 3                                       // src/phaseVocoder.h:4

enum Effect { TIME_SHIFT = 't', PITCH_SHIFT = 'p' };       // src/phaseVocoder.h:5-8

class PhaseVocoder {
    pv_handle* h_ = nullptr;
    std::vector<float> win_;

    static void die(const char* msg, int line)
    {
        if (line >= -1) fprintf(stderr, "Line %d: ", line);
        fprintf(stderr, "Cuda error: %s: %s.\n", msg, pv_last_error());
        exit(EXIT_FAILURE);
    }
    void create(int samples, int hop_in, int hop_out, int mode, int window_type, float pitch)
    {
        pv_params p{};
        p.window = samples;
        p.hop_in = hop_in;
        p.hop_out = hop_out;
        p.mode = mode;
        p.window_type = window_type;
        p.n_voices = 1;
        p.pitch[0] = pitch;
        p.device = -1;
        if (pv_create(&p, &h_) != PV_OK) die("PhaseVocoder", __LINE__);
        win_.resize(samples);
        pv_window_table(h_, win_.data());
        imp = imp1 = win_.data();
        for (int i = 0; i < NUM_STREAMS; i++) {              // src/phaseVocoder.h:113-115
            streams[i] = nullptr;
#ifdef PV_SHIM_HAS_CUDART
            cudaStreamCreate(&streams[i]);
#endif
        }
    }

public:
    float* imp = nullptr;       // window table (host copy; the device copy is owned by the engine)
    float* imp1 = nullptr;
    // per-frame scratch of the reference's unfinished real-time path (src/phaseVocoder.h:18-22): never allocated by the
    // four-argument constructor there either; kept so that code naming them compiles
    float* curr_input = nullptr;
    float* prev_input = nullptr;
    float* prev_output = nullptr;
    float2* prev_mag_phase = nullptr;
    float2* curr_mag_phase = nullptr;
    // src/phaseVocoder.h:23-24: the reference creates two cuFFT plans here and never uses them (kernel.cu makes its own
    // per frame, :324, :363).  This engine has no cuFFT at all; the handles stay 0.
    cufftHandle plan = 0;
    cufftHandle ifft = 0;
    int hopSize;
    int nSamps;
    int R = 1;
    int N = 0;
    float timeScale = 1.f;
    int outHopSize;
    int stream = 0;                                      // src/phaseVocoder.h:31
    cudaStream_t streams[NUM_STREAMS];                   // :32, created in the constructor like the reference's (:113-115)

    // src/phaseVocoder.h:118-126: round-robin stream getters (the reference never calls them; same arithmetic)
    cudaStream_t* getStream()
    {
        cudaStream_t* out = &streams[stream++];
        stream %= NUM_STREAMS;
        return out;
    }
    cudaStream_t* getPrevStream() { return &streams[(stream + NUM_STREAMS - 1) % NUM_STREAMS]; }

    // src/phaseVocoder.h:46-78: periodic Hann, hop = samples/2
    explicit PhaseVocoder(int samples) : hopSize(samples / 2), nSamps(samples), outHopSize(samples / 2)
    {
        create(samples, hopSize, outHopSize, PV_MODE_COMPAT, PV_WIN_HANN_PERIODIC, 1.f);
    }
    // src/phaseVocoder.h:79-116: Hamming, hopSize = samples/hop, TIME_SHIFT: outHopSize = scale*hopSize.
    // PITCH_SHIFT leaves timeScale/outHopSize uninitialised in the reference; here it selects the
    // corrected mode with pitch ratio `scaleFactor` and outHopSize = hopSize.
    PhaseVocoder(int samples, Effect e, float scaleFactor, int hop) : hopSize(samples / hop), nSamps(samples)
    {
        if (e == PITCH_SHIFT) {
            outHopSize = hopSize;
            create(samples, hopSize, outHopSize, PV_MODE_CORRECTED, PV_WIN_HANN_PERIODIC, scaleFactor);
        } else {
            timeScale = scaleFactor;
            outHopSize = (int)(scaleFactor * hopSize);
            create(samples, hopSize, outHopSize, PV_MODE_COMPAT, PV_WIN_HAMMING, 1.f);
        }
    }
    ~PhaseVocoder()
    {
#ifdef PV_SHIM_HAS_CUDART
        for (int i = 0; i < NUM_STREAMS; i++)
            if (streams[i]) cudaStreamDestroy(streams[i]);
#endif
        pv_destroy(h_);
    }
    PhaseVocoder(const PhaseVocoder&) = delete;
    PhaseVocoder& operator=(const PhaseVocoder&) = delete;

    pv_handle* handle() const { return h_; }

    // src/phaseVocoder.cpp:25-33.  fft / intermediary are the reference's scratch buffers: unused.
    void analysis_CUFFT(float* input, float2* output, float2* /*fft*/, float* /*intermediary*/)
    {
        if (pv_analysis(h_, input, reinterpret_cast<float*>(output)) != PV_OK) die("pv_analysis ", __LINE__);
    }
    void analysis(float* input, float2* output, float2* fft, float* intermediary)
    {
        analysis_CUFFT(input, output, fft, intermediary);      // src/phaseVocoder.cpp:34-42
    }
    // src/phaseVocoder.cpp:60-76
    void resynthesis_CUFFT(float* backFrame, float2* frontFrame, float* output)
    {
        if (pv_resynthesis(h_, backFrame, reinterpret_cast<const float*>(frontFrame), output) != PV_OK)
            die("resynthesis", __LINE__);
    }
    void resynthesis(float* backFrame, float2* frontFrame, float2* /*intermediary*/, float* output)
    {
        resynthesis_CUFFT(backFrame, frontFrame, output);       // src/phaseVocoder.cpp:44-58
    }
    void resynthesis(float*, float2*, float*, void (*)()) {}     // empty in the reference too (:77-78)
    // src/phaseVocoder.cpp:20-23
    void test_overlap_add(float* input, float* output, float* /*intermediary*/, float* backFrame, int /*N*/)
    {
        if (pv_test_overlap_add(h_, input, backFrame, output) != PV_OK) die("test_overlap_add", __LINE__);
    }

    // Both host loops of src/main.cpp:228-297 in one call, HOST buffers: x[n_samples] -> out.
    // Returns the number of output samples written (n_synth * outHopSize).
    long process(const float* x, long n_samples, float* out)
    {
        int64_t na = 0, ns = 0;
        pv_reference_schedule(h_, n_samples, &na, &ns);
        if (pv_process_host(h_, x, 1, n_samples, n_samples, na, ns, out, ns * outHopSize, ns * outHopSize, nullptr, 0) != PV_OK)
            die("pv_process", __LINE__);
        return (long)(ns * outHopSize);
    }
};
