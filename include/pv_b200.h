/*
 * pv_b200.h -- C ABI of the B200-native phase-vocoder engine (libpv_b200.so).
 *
 * This is the drop-in boundary for the analysis -> processing -> resynthesis path of
 * davispolito/Phase-Vocoder.  Every entry point names the reference interface it replaces
 * (paths relative to the reference checkout).  Plain pointers and sizes only; no C++ or
 * torch types.  All functions return PV_OK (0) or a negative pv_status and record a
 * message retrievable with pv_last_error() (the reference prints and exit()s instead:
 * checkCUDAError_, src/io.cpp:115-124 -- the C++ shim host/phaseVocoder.h restores that).
 *
 * A pv_handle is NOT thread-safe (it caches its segment plan and staging buffers; the reference's
 * class is not re-entrant either, karnel/kernel.cu:170-174): use one handle per host thread.
 *
 * There is no CPU fallback: every call needs a CUDA device of compute capability 10.x and
 * fails with PV_ERR_CUDA otherwise.
 *
 * Modes
 *   PV_MODE_COMPAT     bit-for-intent reproduction of the reference pipeline, including its
 *                      four defects (atanf quadrant loss, overwritten-x polar->rect, size-N
 *                      C2R on a 2N spectrum, no OLA gain normalisation) -- SURVEY 3.2.
 *   PV_MODE_CORRECTED  true phase vocoder: atan2, phase-difference unwrapping to true bin
 *                      frequency, pitch / time scaling, fixed-point phase accumulation,
 *                      WOLA gain.  Not implemented by the reference (parity unpinned).
 */
#ifndef PV_B200_H
#define PV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PV_MAX_VOICES 8
#define PV_MIN_WINDOW 64
#define PV_MAX_WINDOW 4096

typedef enum pv_status {
    PV_OK = 0,
    PV_ERR_PARAM = -1,       /* invalid argument                                   */
    PV_ERR_CUDA = -2,        /* CUDA runtime / launch failure or no usable device  */
    PV_ERR_ALLOC = -3,       /* allocation failure                                 */
    PV_ERR_UNSUPPORTED = -4  /* valid but not available in this build              */
} pv_status;

typedef enum pv_mode { PV_MODE_COMPAT = 0, PV_MODE_CORRECTED = 1 } pv_mode;

/* Window tables -- PhaseVocoder ctor, src/phaseVocoder.h:62-69 and :84-94. */
typedef enum pv_window_type {
    PV_WIN_HAMMING = 0,       /* 0.54-0.46cos(2 pi i/(N-1))   HEAD, phaseVocoder.h:85-89        */
    PV_WIN_HANN_SYM = 1,      /* 0.5(1-cos(2 pi i/(N-1)))     commented line phaseVocoder.h:87  */
    PV_WIN_HANN_PERIODIC = 2  /* 0.5(1-cos(2 pi i/N))         1-arg ctor :64-66, kernel.cu:85-91 */
} pv_window_type;

enum { PV_FLAG_NAN_COMPAT = 1 /* keep atanf(0/0)=NaN of kernel.cu:108 (default: phase 0) */ };

/* Parameter carrier = the fields of class PhaseVocoder (src/phaseVocoder.h:25-30) plus the
 * knobs the reference hard-codes.                                                         */
typedef struct pv_params {
    int32_t window;                 /* N  = PhaseVocoder::nSamps, power of two 64..4096       */
    int32_t hop_in;                 /* Ha = PhaseVocoder::hopSize (= N / hop divisor, :79)    */
    int32_t hop_out;                /* Hs = PhaseVocoder::outHopSize (= scale*hopSize, :104)  */
    int32_t mode;                   /* pv_mode                                                */
    int32_t window_type;            /* pv_window_type                                         */
    int32_t n_voices;               /* corrected mode: pitch voices per stream (1..8); compat: 1 */
    float pitch[PV_MAX_VOICES];     /* corrected mode: pitch ratio per voice (1.0 = none)     */
    int32_t flags;
    int32_t device;                 /* CUDA ordinal, -1 = current device                      */
} pv_params;

typedef struct pv_handle pv_handle;

/* Last error message of the calling thread ("" if none). */
const char *pv_last_error(void);

/* Library / build description: "pv_b200 <version> sm_100a ..." */
const char *pv_version(void);

/* Replaces PhaseVocoder::PhaseVocoder(int,Effect,float,int) (src/phaseVocoder.h:79-116):
 * validates parameters, builds the window table on the host exactly as the reference does
 * (float arithmetic) and uploads it with the twiddle tables.                               */
int pv_create(const pv_params *params, pv_handle **out);

/* Replaces PhaseVocoder::~PhaseVocoder (src/phaseVocoder.h:128-130). */
void pv_destroy(pv_handle *h);

int pv_get_params(const pv_handle *h, pv_params *out);

/* Copies the N-entry window table (PhaseVocoder::imp, src/phaseVocoder.h:16,85-89) to host. */
int pv_window_table(const pv_handle *h, float *host_out);

/* Reference frame schedule for a stream of num_samples (src/main.cpp:231 and :266).        */
int pv_reference_schedule(const pv_handle *h, int64_t num_samples, int64_t *n_analysed,
                          int64_t *n_synth);

/* ------------------------------------------------------------------------------------------
 * Per-frame entry points: the reference's own granularity.  Pointers must be device
 * accessible (device or managed memory), exactly as the reference requires.  Synchronous.
 * ---------------------------------------------------------------------------------------- */

/* Replaces PhaseVocoder::analysis_CUFFT (src/phaseVocoder.cpp:25-33) ->
 * CudaPhase::pv_analysis_CUFFT (karnel/kernel.cu:299-348): in[N] -> out float2[2N] {mag,phase}.
 * Unlike the reference the output buffer need not be pre-zeroed and no scratch is needed.  */
int pv_analysis(pv_handle *h, const float *in, float *out_magphase);

/* Replaces PhaseVocoder::resynthesis_CUFFT (src/phaseVocoder.cpp:60-76) ->
 * CudaPhase::resynthesis_CUFFT (karnel/kernel.cu:352-432): back[N], front float2[2N] -> out[N].
 * front is NOT modified (the reference rewrites it in place, kernel.cu:354).               */
int pv_resynthesis(pv_handle *h, const float *back, const float *front_magphase, float *out);

/* Replaces PhaseVocoder::test_overlap_add (src/phaseVocoder.cpp:20-23, kernel.cu:289-298):
 * window -> half swap -> half swap -> window -> overlap-add, no FFT.                        */
int pv_test_overlap_add(pv_handle *h, const float *in, const float *back, float *out);

/* ------------------------------------------------------------------------------------------
 * Batched entry points (device pointers, asynchronous on `cuda_stream`, a cudaStream_t
 * passed as void*; NULL = default stream).
 * ---------------------------------------------------------------------------------------- */

/* The analysis loop of src/main.cpp:228-250 in one launch: frame k (0 <= k < n_frames) reads
 * in[k*Ha .. k*Ha+N) (samples at index >= n_in read as 0) and writes
 * out_magphase[k][2N] {mag, phase}.                                                         */
int pv_analysis_batch(pv_handle *h, const float *in, int64_t n_in, int64_t n_frames,
                      float *out_magphase, void *cuda_stream);

/* The resynthesis loop of src/main.cpp:264-297 in one launch: consumes spectra[k][2N],
 * carries back[N] (in/out, main.cpp:253-258,279) and writes out[k*Hs .. (k+1)*Hs).           */
int pv_resynthesis_batch(pv_handle *h, const float *spectra, int64_t n_frames, float *back,
                         float *out, void *cuda_stream);

/* Bytes of per-stream carried state (compat: the accumulated frame `back`, N floats;
 * corrected: previous analysis phase, phase accumulators and OLA accumulators per voice).    */
size_t pv_state_bytes(const pv_handle *h);

/* THE HOT PATH: fused analysis -> processing -> resynthesis/overlap-add over a batch of
 * independent streams (channels), replacing both host loops of src/main.cpp:228-297 and every
 * kernel / cuFFT call under them.
 *
 *   in          n_streams rows of n_in samples, row pitch in_stride (floats)
 *   n_analysed  compat: frames k >= n_analysed are zero spectra (main.cpp:216,231); pass
 *               n_frames for "all".  corrected: ignored.
 *   n_frames    frames synthesised per stream; frame k reads in[k*Ha .. k*Ha+N), zero past n_in
 *   out         [stream][voice][n_frames*Hs]: out + s*out_stream_stride + v*out_voice_stride
 *   state       NULL, or n_streams * pv_state_bytes(h) bytes: read as the carry-in when
 *               (flags & PV_PROCESS_CARRY_IN), written as the carry-out when
 *               (flags & PV_PROCESS_CARRY_OUT).  An all-zero state is a fresh start, so a block-by-block
 *               caller may zero it once and pass both flags on every call.
 * Long streams are split into frame-range segments processed concurrently (the (N-Hs) OLA
 * halo is recomputed, so the result does not depend on the split).  Corrected mode: a split runs an
 * analysis pass first (per-bin phase carries); for windows >= 512 that pass keeps magnitude and
 * phase difference of every frame in device scratch owned by the handle -- (N/2 + 2) * 8 bytes per
 * frame and stream, at most 16 GB, freed by pv_destroy -- and the processing pass synthesises from it
 * instead of repeating the forward transform; a call whose scratch would not fit (or cannot be
 * allocated) recomputes.  Either way the result is bit-identical to the unsplit run.              */
enum { PV_PROCESS_CARRY_IN = 1, PV_PROCESS_CARRY_OUT = 2, PV_PROCESS_REUSE_AGGREGATE = 4 };

int pv_process_device(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride,
                      int64_t n_in, int64_t n_analysed, int64_t n_frames, float *out,
                      int64_t out_stream_stride, int64_t out_voice_stride, void *state,
                      int32_t flags, void *cuda_stream);

/* pv_process_device with `skip_frames` leading frames computed but NOT written: frame k of the call
 * (k >= skip_frames) lands at out[(k - skip_frames)*Hs ..].  This is how a frame range [k0, k1) of a
 * longer stream is produced on its own (another GPU, another call): pass the input from frame
 * k0 - halo, halo = (N-1)/Hs, and skip the halo.  Compat frames are independent, so nothing else is
 * needed; corrected mode additionally takes the phase carry as the carried-in state (below).
 * Replaces the frame bookkeeping of src/main.cpp:231,266 for a sharded stream.                      */
int pv_process_device_ex(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride,
                         int64_t n_in, int64_t n_analysed, int64_t n_frames, int64_t skip_frames,
                         float *out, int64_t out_stream_stride, int64_t out_voice_stride,
                         void *state, int32_t flags, void *cuda_stream);

/* Corrected mode, frame-range scan support (the per-bin phase carry of SURVEY 8e).  Analysis-only
 * pass over frames 0..n_frames-1 of every stream:
 *   sumD[stream][bin]   int64  sum of the unwrapped phase differences D_k (k >= 1; k >= 0 if P_prev given)
 *   P_first[stream][bin] u32   phase of frame 0 (turns*2^32); may be NULL
 *   P_last[stream][bin]  u32   phase of the last frame; may be NULL
 * Integer sums are associative, so per-range aggregates combine exactly (any order, any sharding). */
int pv_corrected_aggregate(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride,
                           int64_t n_in, int64_t n_frames, const uint32_t *P_prev, int64_t *sumD,
                           uint32_t *P_first, uint32_t *P_last, void *cuda_stream);

/* Builds the carried state at frame boundary n_before (frames 0..n_before-1 already accounted for):
 * psi = (P_first[a] << 32) + (n_before-1)*nomS + Rq * sumD[a], P_prev as given, empty OLA
 * accumulators (run the halo frames with skip_frames to fill them).  All pointers are device memory. */
int pv_corrected_state_from_carry(pv_handle *h, int64_t n_streams, const uint32_t *P_first,
                                  const int64_t *sumD, int64_t n_before, const uint32_t *P_prev,
                                  void *state, void *cuda_stream);

/* The analysis pass of pv_process_device_ex(h, in, ..., n_frames, skip_frames, ..., state, flags, stream) on its own:
 * sumD[stream][bin] = sum of the unwrapped phase differences over ALL frames of that call (halo / skipped frames
 * included; frame 0 counts only when a state is carried in, which supplies the previous phase).  When the call is
 * one that the library cuts into frame-range parts (few streams), the per-part sums stay in the handle, and the
 * matching pv_process_device_ex call -- same arguments, issued next on the same stream, with
 * PV_PROCESS_REUSE_AGGREGATE added to its flags -- skips its own analysis pass.  This is what lets a rank of a
 * sharded long stream learn its phase-carry contribution, exchange it, and then finish with one pass instead of
 * two; the flag is ignored whenever nothing reusable is there, so it is always safe to pass.                     */
int pv_corrected_split_aggregate(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride,
                                 int64_t n_in, int64_t n_frames, int64_t skip_frames, const void *state,
                                 int32_t flags, int64_t *sumD, void *cuda_stream);

/* ------------------------------------------------------------------------------------------
 * Frame-range sharding of long streams over the GPUs of a box (BASELINE.json config 5): rank r of
 * `world` (one handle per rank / GPU) produces frames [k0, k1) of every stream.  The reference has no such path
 * (one host loop over all frames, src/main.cpp:228-297); this replaces that loop for a sharded file.
 *
 * What crosses ranks:  compat -- nothing (frames are independent; the (N-1)/Hs frames in front of a range are
 * recomputed from the input halo).  corrected -- the per-bin phase carry: ONE all-gather of
 * pv_shard_carry_elems() int64 per stream and rank (the rank's sum of unwrapped phase differences per bin, plus
 * rank 0's phase of frame 0).  Integer sums are associative, so the sharded output is bit-identical to the
 * single-GPU one.  The library does everything except the collective itself, which the caller issues between the
 * two calls with whatever it has (ncclAllGather, MPI_Allgather, torch.distributed, cudaMemcpyPeerAsync):
 *
 *     pv_shard_begin (h, in, ..., world, rank, carry_send, stream);          // analysis of the rank's range
 *     ncclAllGather(carry_send, carry_all, n_streams * pv_shard_carry_elems(h), ncclInt64, comm, stream);
 *     pv_shard_finish(h, in, ..., world, rank, carry_all, out, ..., stream); // state from the carry + processing
 *
 * `in` points at sample in_first_frame*Ha of stream 0 (in_first_frame <= max(0, ks-1), so a rank only needs its
 * own range plus the halo resident; rank 0 needs frame 0), n_in counts the valid samples from there (zero beyond).
 * carry_send: device, [n_streams][elems]; carry_all: device, [world][n_streams][elems].  The range is analysed
 * once: pv_shard_finish reuses the per-part sums pv_shard_begin left in the handle (PV_PROCESS_REUSE_AGGREGATE),
 * provided nothing else ran on the handle in between (otherwise it recomputes them; the result is the same).
 * out receives frames [k0, k1): out[stream][voice][(k - k0)*Hs ..].
 * ---------------------------------------------------------------------------------------- */
typedef struct pv_shard_plan {
    int64_t k0, k1;   /* frames this rank writes */
    int64_t ks;       /* first frame it computes: k0 - halo, clipped at 0 */
    int64_t halo;     /* (N-1)/Hs */
} pv_shard_plan;

int pv_shard_plan_frames(const pv_handle *h, int64_t n_frames, int32_t world, int32_t rank, pv_shard_plan *out);
size_t pv_shard_carry_elems(const pv_handle *h);
int pv_shard_begin(pv_handle *h, const float *in, int64_t in_first_frame, int64_t n_streams, int64_t in_stride,
                   int64_t n_in, int64_t n_frames, int32_t world, int32_t rank, int64_t *carry_send,
                   void *cuda_stream);
int pv_shard_finish(pv_handle *h, const float *in, int64_t in_first_frame, int64_t n_streams, int64_t in_stride,
                    int64_t n_in, int64_t n_analysed, int64_t n_frames, int32_t world, int32_t rank,
                    const int64_t *carry_all, float *out, int64_t out_stream_stride, int64_t out_voice_stride,
                    void *cuda_stream);

/* Same with HOST buffers (pinned or pageable): H2D, kernel, D2H, synchronise.  This is the
 * call the C++ PhaseVocoder shim and the CLI make, and what bench.py times as `e2e`.         */
int pv_process_host(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride,
                    int64_t n_in, int64_t n_analysed, int64_t n_frames, float *out,
                    int64_t out_stream_stride, int64_t out_voice_stride, void *state,
                    int32_t flags);

/* Same with 16-bit PCM host buffers: the sample conversions of the reference's AudioFile<float> -- s/32768 on
 * load (src/AudioFile.h:1038-1042) and (int16) trunc(clamp(x,-1,1)*32767) on save (:1045-1049) -- run on the
 * device, so a WAV file's samples cross PCIe as they are stored (half the bytes of the float call).         */
int pv_process_host_pcm16(pv_handle *h, const int16_t *in, int64_t n_streams, int64_t in_stride,
                          int64_t n_in, int64_t n_analysed, int64_t n_frames, int16_t *out,
                          int64_t out_stream_stride, int64_t out_voice_stride, void *state,
                          int32_t flags);

/* Same with packed 24-bit PCM host buffers (three little-endian bytes per sample, as in the `data` chunk of the
 * reference's testtones/MAT_ZO_24_bit.wav): sign-extend and /8388608 on load (src/AudioFile.h:508-518), the low
 * three bytes of (int32)(x*8388608) on save (:755-766; AudioFile does not clamp there and neither does this).
 * Strides and counts are in SAMPLES.  Three quarters of the bytes of the float call cross PCIe.              */
int pv_process_host_pcm24(pv_handle *h, const uint8_t *in, int64_t n_streams, int64_t in_stride,
                          int64_t n_in, int64_t n_analysed, int64_t n_frames, uint8_t *out,
                          int64_t out_stream_stride, int64_t out_voice_stride, void *state,
                          int32_t flags);

/* ------------------------------------------------------------------------------------------
 * Real-time block server: the reference's unfinished RtAudio path (src/main.cpp:45-59 `callback`,
 * README.md:44-50, img/Realtime.png).  Its contract: every time the audio buffer is full the
 * callback receives nBufferFrames new samples and must return nBufferFrames processed ones,
 * "looking backwards at the previous input through a ring buffer to maintain continuity".
 * Here one server handles n_streams such callbacks at once.  A block is block_frames hops:
 * block_frames*Ha new samples per stream in, block_frames*Hs samples per stream and voice out.
 * The last N-Ha input samples stay on the device (the ring), the phase accumulators and the
 * overlap-add tail are carried from block to block, and one block is ONE CUDA-graph launch
 * (H2D of the block, the fused kernel, the ring advance, D2H of the result).
 *
 * The output equals the offline result of pv_process_* on the input delayed by the latency
 * N-Ha (the ring starts out silent), bit for bit.
 * ---------------------------------------------------------------------------------------- */
typedef struct pv_rt pv_rt;

int pv_rt_open(pv_handle *h, int64_t n_streams, int32_t block_frames, pv_rt **out);
void pv_rt_close(pv_rt *rt);
/* Silence the ring and the carried state (a new take). */
int pv_rt_reset(pv_rt *rt);
/* Input-to-output delay in samples at the input rate: N - Ha. */
int64_t pv_rt_latency_samples(const pv_rt *rt);
/* Page-locked staging owned by the server: input [n_streams][block_frames*Ha], output
 * [n_streams][voices][block_frames*Hs].  Fill the input, call pv_rt_step, read the output:
 * the zero-copy variant of the callback for callers that can record straight into it.       */
float *pv_rt_input(pv_rt *rt);
float *pv_rt_output(pv_rt *rt);
int pv_rt_step(pv_rt *rt);
/* The reference's callback signature for n_streams planar channels (RTAUDIO_NONINTERLEAVED
 * layout): in [n_streams][nBufferFrames], out [n_streams][voices][nBufferFrames*Hs/Ha];
 * nBufferFrames must be block_frames*Ha.  Returns 0 like the reference's callback.          */
int pv_rt_callback(pv_rt *rt, float *outputBuffer, const float *inputBuffer, uint32_t nBufferFrames);

/* ------------------------------------------------------------------------------------------
 * Stand-alone batched complex FFT: the counterpart of the reference's FFT back-ends and of the
 * micro-benchmark that times them (karnel/hpfft.cu:145-203 `GPU_FFT`, :104-143 `FFTShMem`,
 * karnel/cufft_.cu:19-26 `computeCuFFT`).  in/out: batch x n interleaved complex floats on the
 * device (in == out allowed), n a power of two <= 8192, direction -1 forward / +1 inverse,
 * unnormalised both ways (cuFFT's convention).  One launch per call.                          */
int pv_fft_batch(pv_handle *h, const float *in, float *out, int32_t n, int64_t batch,
                 int32_t direction, void *cuda_stream);

/* Number of kernel launches issued by this handle so far (bench.py's gpu_launches). */
int64_t pv_launch_count(const pv_handle *h);

/* Average device time in ms of the fused kernel launches issued since the last call of
 * pv_timing_reset (CUDA events recorded on the launching stream); 0 if none.               */
int pv_timing_enable(pv_handle *h, int32_t on);
int pv_timing_read(pv_handle *h, double *total_ms, int64_t *launches);

#ifdef __cplusplus
}
#endif
#endif /* PV_B200_H */
