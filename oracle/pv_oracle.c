/*
 * pv_oracle.c -- CPU restatement of the davispolito/Phase-Vocoder hot path.
 * TEST INFRASTRUCTURE ONLY (see pv_oracle.h).  Plain C11, no dependencies.
 *
 * compat mode  : PINNED against output/testout.wav and output/1000hzout.wav.
 * corrected    : PARITY UNPINNED (no reference implementation exists).
 */
#define _GNU_SOURCE
#include "pv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---------------- window tables: src/phaseVocoder.h:62-69, 84-94 ---------------- */

void pvo_window(int type, int N, float *w)
{
    /* The reference computes in float on the host: `float omega = 2.f * M_PI / (samples-1)`
     * then `0.54f - 0.46f * cos(omega*(i))` (float overload).                          */
    if (type == PVO_WIN_HANN_PERIODIC) {
        for (int i = 0; i < N; i++)
            w[i] = 0.5f * (1.f - cosf((float)(2.f * M_PI * i / N)));      /* phaseVocoder.h:65 */
        return;
    }
    const float omega = (float)(2.f * M_PI / (N - 1));                    /* phaseVocoder.h:85 */
    for (int i = 0; i < N; i++) {
        if (type == PVO_WIN_HAMMING)
            w[i] = 0.54f - 0.46f * cosf(omega * (float)i);                /* phaseVocoder.h:88 */
        else
            w[i] = 0.5f * (1.f - cosf(omega * (float)i));                 /* phaseVocoder.h:87 */
    }
}

/* ---------------- frame schedule: src/main.cpp:231, 266 ---------------- */

void pvo_reference_schedule(long num_samples, int Ha, int Hs, long *n_analysed, long *n_synth)
{
    long na = 0;
    for (long i = 0; i < num_samples - Ha; i += Ha) na++;                 /* main.cpp:231 */
    if (n_analysed) *n_analysed = na;
    if (n_synth) *n_synth = num_samples / Hs;                             /* main.cpp:266 */
}

/* ---------------- corrected-mode integer tables ---------------- */

void pvo_corrected_tables(int N, int Ha, int Hs, double beta,
                          uint64_t *beta_q_out, uint64_t *Rq_out, int32_t *a_lo, int32_t *a_hi,
                          uint64_t *nomS, uint32_t *nomA)
{
    const int h = N / 2, nb = h + 1;
    int lg = 0;
    while ((1 << lg) < N) lg++;
    const uint64_t beta_q = (uint64_t)llround(beta * 4294967296.0);
    const uint64_t Rq = (beta_q * (uint64_t)Hs + (uint64_t)(Ha / 2)) / (uint64_t)Ha;
    for (int s = 0; s < nb; s++) { a_lo[s] = 1; a_hi[s] = 0; }
    for (int a = 0; a < nb; a++) {
        uint64_t s = ((uint64_t)a * beta_q + 0x80000000ull) >> 32;
        if (s > (uint64_t)h) break;
        if (a_lo[s] > a_hi[s]) a_lo[s] = a;
        a_hi[s] = a;
    }
    for (int s = 0; s < nb; s++) {
        if (a_lo[s] > a_hi[s]) { nomS[s] = 0; continue; }
        nomS[s] = (beta_q * (uint64_t)a_hi[s] * (uint64_t)Hs) << (32 - lg);
    }
    if (nomA)
        for (int b = 0; b < nb; b++)
            nomA[b] = (uint32_t)(((uint64_t)b * (uint64_t)Ha) << (32 - lg));
    if (beta_q_out) *beta_q_out = beta_q;
    if (Rq_out) *Rq_out = Rq;
}

float pvo_corrected_gain(const float *win, int N, int Hs)
{
    double s = 0;
    for (int i = 0; i < N; i++) s += (double)win[i] * (double)win[i];
    return (float)((double)Hs / s);
}

uint32_t pvo_phase_turns32(double re, double im)
{
    double t = atan2(im, re) * (0.5 / M_PI);
    return (uint32_t)(int64_t)llrint(t * 4294967296.0);
}

/* ---------------- the two arithmetic variants ---------------- */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL double
#define SFX(n) CAT(n, _f64)
#define R_SIN sin
#define R_COS cos
#define R_SQRT sqrt
#define R_ATAN atan
#define R_ATAN2 atan2
#include "pv_oracle_impl.inc"
#undef REAL
#undef SFX
#undef R_SIN
#undef R_COS
#undef R_SQRT
#undef R_ATAN
#undef R_ATAN2

#define REAL float
#define SFX(n) CAT(n, _f32)
#define R_SIN sinf
#define R_COS cosf
#define R_SQRT sqrtf
#define R_ATAN atanf
#define R_ATAN2 atan2f
#include "pv_oracle_impl.inc"
#undef REAL
#undef SFX

void pvo_fft_f64(double *re, double *im, int n, int dir) { fft_f64(re, im, n, dir); }

int pvo_process_corrected_traced(const float *x, long n_in, int N, int Ha, int Hs, const float *win,
                                 int n_voices, const double *beta, long n_frames,
                                 pvo_corrected_state *state, int precision,
                                 double *out, long out_stride, const pvo_corrected_trace *trace)
{
    if (N < 4 || (N & (N - 1)) || Ha < 1 || Hs < 1 || Hs > N || n_voices < 1 ||
        n_voices > PVO_MAX_VOICES)
        return -1;
    if (precision == 32)
        return corrected_process_f32(x, n_in, N, Ha, Hs, win, n_voices, beta, n_frames, state, out,
                                     out_stride, trace);
    return corrected_process_f64(x, n_in, N, Ha, Hs, win, n_voices, beta, n_frames, state, out,
                                 out_stride, trace);
}

int pvo_process_corrected(const float *x, long n_in, int N, int Ha, int Hs, const float *win,
                          int n_voices, const double *beta, long n_frames,
                          pvo_corrected_state *state, int precision,
                          double *out, long out_stride)
{
    return pvo_process_corrected_traced(x, n_in, N, Ha, Hs, win, n_voices, beta, n_frames, state, precision, out,
                                        out_stride, NULL);
}

int pvo_corrected_aggregate(const float *x, long n_in, int N, int Ha, const float *win,
                            long n_frames, int have_prev, const uint32_t *P_prev_in,
                            int precision, int64_t *sumD, uint32_t *P_last)
{
    if (N < 4 || (N & (N - 1)) || Ha < 1) return -1;
    if (precision == 32)
        return corrected_aggregate_f32(x, n_in, N, Ha, win, n_frames, have_prev, P_prev_in, sumD,
                                       P_last);
    return corrected_aggregate_f64(x, n_in, N, Ha, win, n_frames, have_prev, P_prev_in, sumD,
                                   P_last);
}
