/*
 * pv_oracle.h -- CPU restatement of the davispolito/Phase-Vocoder hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and there only as the checker or the timed CPU
 * baseline.  The product path (phase-vocoder_b200/csrc) never links it.
 *
 * Two arithmetic variants are built from the same template (pv_oracle_impl.inc):
 *   *_f64  double precision -- the parity "truth" (golden WAVs are matched with it)
 *   *_f32  single precision -- the reference's own arithmetic type; used as the
 *          CPU baseline that bench.py times on the host cores.
 *
 * Parity status:
 *   compat mode    PINNED by output/testout.wav and output/1000hzout.wav of the
 *                  reference (tests/test_oracle_golden.py, tests/golden/).
 *   corrected mode PARITY UNPINNED -- the reference implements neither phase
 *                  unwrapping nor pitch shift (src/phaseVocoder.h:107-111,
 *                  src/main.cpp:301-303); this oracle is the specification
 *                  (DESIGN.md "corrected mode"), checked by analytic properties.
 *
 * All "file:line" citations are relative to the reference checkout.
 */
#ifndef PV_ORACLE_H
#define PV_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* window tables (src/phaseVocoder.h:62-69, 84-94) */
enum {
    PVO_WIN_HAMMING = 0,       /* 0.54-0.46cos(2*pi*i/(N-1))  -- HEAD, phaseVocoder.h:85-89   */
    PVO_WIN_HANN_SYM = 1,      /* 0.5(1-cos(2*pi*i/(N-1)))    -- commented line phaseVocoder.h:87 */
    PVO_WIN_HANN_PERIODIC = 2  /* 0.5(1-cos(2*pi*i/N))        -- 1-arg ctor phaseVocoder.h:64-66, kernel.cu:85-91 */
};

/* flags */
enum {
    PVO_FLAG_NAN_COMPAT = 1    /* propagate atanf(0/0)=NaN like kernel.cu:108; default maps 0/0 -> phase 0 */
};

#define PVO_MAX_VOICES 8

/* Builds the float window table exactly as the reference host code does (float arithmetic). */
void pvo_window(int type, int N, float *w);

/* Reference frame schedule (src/main.cpp:231 and :266):
 *   analysed    k = 0 .. ceil((numSamples-Ha)/Ha)-1     (loop i=0; i<numSamples-Ha; i+=Ha)
 *   synthesised k = 0 .. numSamples/Hs - 1              (integer division)               */
void pvo_reference_schedule(long num_samples, int Ha, int Hs, long *n_analysed, long *n_synth);

/* ---- compat mode, per-frame entry points (the reference's own granularity) ---- */

/* Steps A-D of SURVEY 3.2 = CudaPhase::pv_analysis_CUFFT (karnel/kernel.cu:299-348):
 * window (:301, :68-74), zero-phase shift + zero pad to 2N (:314, :25-32), forward
 * 2N-point DFT (:324-326), {mag, atanf(im/re)} (:337, :101-109).
 * in: N samples, win: N, out: 2N interleaved {mag, phase}.                              */
void pvo_analysis_frame_f64(const float *in, const float *win, int N, int flags, double *out_magphase);
void pvo_analysis_frame_f32(const float *in, const float *win, int N, int flags, float *out_magphase);

/* Steps E-H = CudaPhase::resynthesis_CUFFT (karnel/kernel.cu:352-432):
 * polar->rect with the overwritten-x defect (:354, :121-129), size-N C2R on bins 0..N/2 of
 * the 2N spectrum (:363-366), /N (:380), half swap (:393), window (:406), overlap-add of
 * the previous accumulated frame (:419, :111-119).
 * front: 2N interleaved {mag, phase}; back: N (read only); out: N.                      */
void pvo_resynthesis_frame_f64(const double *back, const double *front_magphase, const float *win,
                               int N, int Hs, double *out);
void pvo_resynthesis_frame_f32(const float *back, const float *front_magphase, const float *win,
                               int N, int Hs, float *out);

/* ---- compat mode, whole stream = the two host loops of src/main.cpp:228-250, 264-297 ----
 * x[n_in] (zero beyond n_in), frames first_frame .. first_frame+n_frames-1 are
 * synthesised; frame k is analysed from x[k*Ha ..] if k < n_analysed, otherwise its spectrum
 * is all-zero (an un-analysed, pre-zeroed d_output buffer, main.cpp:216).
 * back[N] is the carried accumulated frame (main.cpp:253-258, 279): in/out.
 * out receives n_frames*Hs samples (main.cpp:281-295).  Returns 0 or -1 on bad params.  */
int pvo_process_compat_f64(const float *x, long n_in, int N, int Ha, int Hs, const float *win,
                           long n_analysed, long first_frame, long n_frames, int flags,
                           double *back, double *out);
int pvo_process_compat_f32(const float *x, long n_in, int N, int Ha, int Hs, const float *win,
                           long n_analysed, long first_frame, long n_frames, int flags,
                           float *back, float *out);

/* ---- corrected mode (specification; parity unpinned, see header comment) ----
 * Integer helper tables shared by definition with the product:
 *   beta_q  = llround(beta * 2^32)                                  (Q32.32)
 *   s(a)    = (a*beta_q + 2^31) >> 32          analysis bin a -> synthesis bin
 *   a_lo/a_hi[s] = contiguous range of a in [0,N/2] mapping to s (a_lo>a_hi: empty)
 *   nomA[b] = ((b*Ha) << (32-lgN)) mod 2^32    expected analysis phase advance, turns*2^32
 *   nomS[s] = ((beta_q*a_hi[s]*Hs) << (32-lgN)) mod 2^64   expected synthesis advance, turns*2^64
 *   Rq      = (beta_q*Hs + Ha/2) / Ha          deviation scale beta*Hs/Ha (Q32.32)        */
typedef struct pvo_voice_tables {
    uint64_t beta_q;
    uint64_t Rq;
    int32_t *a_lo;      /* N/2+1 */
    int32_t *a_hi;      /* N/2+1 */
    uint64_t *nomS;     /* N/2+1 */
} pvo_voice_tables;

/* Fills caller-allocated arrays. */
void pvo_corrected_tables(int N, int Ha, int Hs, double beta,
                          uint64_t *beta_q, uint64_t *Rq, int32_t *a_lo, int32_t *a_hi,
                          uint64_t *nomS, uint32_t *nomA /* may be NULL */);

/* WOLA gain Hs / sum(w^2) as float (computed in double from the float table). */
float pvo_corrected_gain(const float *win, int N, int Hs);

/* Stream state carried between calls / segments / GPUs. */
typedef struct pvo_corrected_state {
    int32_t have_prev;              /* 0: next frame is the first of the stream                */
    uint32_t *P_prev;               /* N/2+1 : analysis phase of the previous frame, turns*2^32 */
    uint64_t *psi;                  /* V*(N/2+1): synthesis phase accumulators, turns*2^64      */
    double *tail;                   /* V*N : OLA accumulator ring, linearised: tail[v][0..N-Hs) */
} pvo_corrected_state;

/* Processes frames 0..n_frames-1 of x (frame k reads x[k*Ha .. k*Ha+N), zero beyond n_in),
 * n_voices pitch ratios; out[v*out_stride + k*Hs + j].  state may be all-zero (fresh stream).
 * precision: 64 -> double FFT/trig, 32 -> float FFT/trig (integer phase path identical).  */
int pvo_process_corrected(const float *x, long n_in, int N, int Ha, int Hs, const float *win,
                          int n_voices, const double *beta, long n_frames,
                          pvo_corrected_state *state, int precision,
                          double *out, long out_stride);

/* Decision trace of the corrected mode, for DECISION-ALIGNED parity checks (DESIGN.md "conditioning of the
 * phase unwrap").  The unwrap D = (int32)(P_k - P_{k-1} - nomA) is discontinuous at +-1/2 turn, so two correct
 * implementations with different rounding may pick neighbouring aliases for a bin whose phase difference lies
 * within their phase error of the boundary.  A checker records the oracle's own decisions (D_out, with mag_out to
 * judge the bin's phase uncertainty), compares them with the implementation's, verifies that every disagreement is
 * such a boundary case, and re-runs the oracle with those few decisions moved (unwrap_adjust = +-1 turn):
 * everything else must then agree to the full tolerance.  All arrays are [n_frames][N/2+1]; any may be NULL. */
typedef struct pvo_corrected_trace {
    int32_t *D_out;               /* the oracle's unwrapped phase difference, turns*2^32 (0 for a first frame) */
    double *mag_out;              /* |X| per bin */
    const int8_t *unwrap_adjust;  /* added to D in whole turns before the accumulation */
} pvo_corrected_trace;

int pvo_process_corrected_traced(const float *x, long n_in, int N, int Ha, int Hs, const float *win,
                                 int n_voices, const double *beta, long n_frames,
                                 pvo_corrected_state *state, int precision,
                                 double *out, long out_stride, const pvo_corrected_trace *trace);

/* Analysis only: per-bin sum of D_k = (int32)(P_k - P_{k-1} - nomA) over frames [0,n_frames) as
 * int64, plus P of the last frame -- the "segment aggregate" of the frame-range scan. */
int pvo_corrected_aggregate(const float *x, long n_in, int N, int Ha, const float *win,
                            long n_frames, int have_prev, const uint32_t *P_prev_in,
                            int precision, int64_t *sumD, uint32_t *P_last);

/* phase helper: turns*2^32 of atan2(im,re) */
uint32_t pvo_phase_turns32(double re, double im);

/* plain complex DFT used by the oracle (exposed for unit tests): dir=-1 forward, +1 inverse
 * (unnormalised), n power of two. */
void pvo_fft_f64(double *re, double *im, int n, int dir);

#ifdef __cplusplus
}
#endif
#endif
