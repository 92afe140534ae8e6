// pv_internal.h -- shared declarations of the engine (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pv_b200.h"

// Immutable per-handle tables, passed to kernels by value.
struct PvDev {
    const float *win;     // N        window (PhaseVocoder::imp)
    const float2 *tw;     // N        tw[k] = exp(-j*pi*k/N): the 2N-th roots of unity, k < N
    int N;                // window length
    int lgN;
    int Ha, Hs;
    int flags;
    float inv_N;
    // corrected mode
    int V;
    float gain;           // Hs / sum(w^2)
    const uint32_t *nomA; // nb
    const int32_t *a_lo;  // V*nb
    const int32_t *a_hi;  // V*nb
    const uint64_t *nomS; // V*nb
    const uint32_t *gather; // V*T*9: per-thread packed a_lo | a_hi << 16 (pv_fused_tables.h)
    const uint32_t *gather_nat; // V*nb: the same in natural bin order (window 4096 in-place kernel)
    uint64_t beta_q[PV_MAX_VOICES];
    uint64_t Rq[PV_MAX_VOICES];
    int32_t multi[PV_MAX_VOICES];   // voice sums several analysis bins into some synthesis bin (pitch ratio < 1)
};

// A frame-range segment of one stream: frames [k_begin, k_end) are computed, frames
// [k_emit, k_end) are written to the output; k_begin < k_emit are the recomputed OLA halo.
struct PvSegment {
    int32_t stream;
    int32_t carry_in;     // 1: k_begin == 0 and the accumulator starts from the stream state
    int32_t carry_out;    // 1: this segment holds the last frame and writes the stream state
    int32_t state_idx;    // state slot (stream index, or a per-part slot when a corrected stream is split)
    int64_t k_begin, k_emit, k_end;
};

struct PvProcessArgs {
    const float *in;
    int64_t in_stride, n_in, n_analysed, n_frames;
    float *out;
    int64_t out_stream_stride, out_voice_stride;
    unsigned char *state; // per-stream carried state, pv_state_bytes() each
    int64_t state_stride; // bytes
    const PvSegment *segs;
    int32_t n_segs;
    // Stored analysis (fused corrected kernel; null = compute it): {|X|, D} per frame and bin as PvAggArgs::md left them; the
    // frames are then synthesised from it -- no input samples are read.  P_last [n_segs][N/2 + 1]: the phase of each segment's
    // last frame (from the same analysis pass), which a segment with carry_out writes into the state.
    const float2 *md;
    int64_t md_stream_stride;
    const uint32_t *P_last;
};

// Device copies of the fused kernels' twiddle tables (pv_fused_tables.h).
struct PvFusedTables {
    const float2 *tw1 = nullptr, *tw2 = nullptr, *tw2n = nullptr, *itw1 = nullptr, *itw2 = nullptr;
    const float2 *ctw1 = nullptr, *ctw2 = nullptr;
};

// Opts a kernel in to the device's full dynamic shared memory ONCE per kernel and device.  The attribute is process-wide per
// function: setting it to each launch's own size (round 1) let two host threads with different voice counts or ring modes
// shrink it under each other's feet ("too many resources requested for launch").
template <auto Kern, bool CARVEOUT = true>
inline cudaError_t pv_max_smem_once()
{
    static bool done[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !done[dev]) {
        e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        if (CARVEOUT) e = cudaFuncSetAttribute(Kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    return cudaSuccess;
}

// ---- launchers (defined in the .cu files) ----
bool pv_fused_compat_supported(int N, int Hs);
// number of segment groups that can be resident on the device at once
int pv_fused_compat_capacity(int N, int sm_count);
cudaError_t pv_launch_compat_fused(const PvDev &d, const PvFusedTables &t, const PvProcessArgs &a, cudaStream_t st);
bool pv_fused_corrected_supported(int N, int Ha, int Hs);
int pv_fused_corrected_capacity(int N, int V, int sm_count);
cudaError_t pv_launch_corrected_fused(const PvDev &d, const PvFusedTables &t, const PvProcessArgs &a, cudaStream_t st);
// Analysis-only pass over a table of frame ranges.  For segment g: frames [k_begin, k_end) of stream
// segs[g].stream are analysed; the unwrapped phase differences D_k (defined when a previous phase exists:
// k > k_begin, or k == k_begin with carry_in and P_prev[stream]) are summed into H[g] for k < k_emit and
// into S[g] for k >= k_emit.  P_first[g] / P_last[g] = phase of the first / last analysed frame.
// H, P_first, P_last may be null.  Outputs are indexed by segment.
struct PvAggArgs {
    const float *in;
    int64_t in_stride, n_in;
    const PvSegment *segs;
    int32_t n_segs;
    const uint32_t *P_prev;   // per stream, used by segments with carry_in: P_prev + stream*P_prev_stride
    int64_t P_prev_stride;    // in uint32 units (nb for a dense array, state_bytes/4 inside state blobs)
    int64_t *S, *H;
    uint32_t *P_first, *P_last;
    int32_t P_prev_in_state;  // P_prev rows sit inside state records: the have_prev word (P_prev[-2]) gates the carry
    // Analysis store (fused kernels; null = off): the pass also writes {|X|, D} of every frame a segment OWNS (k >= k_emit) to
    // md + stream * md_stream_stride + frame * (N/2 + 2), float2 units, so that the processing pass of a frame-range split
    // reads them back instead of repeating the forward transform (PvProcessArgs::md).
    float2 *md;
    int64_t md_stream_stride;
};
cudaError_t pv_launch_corrected_aggregate(const PvDev &d, const PvFusedTables &t, const PvAggArgs &a, cudaStream_t st);
cudaError_t pv_launch_state_from_carry(const PvDev &d, const PvFusedTables &t, int64_t n_streams, const uint32_t *P_first,
                                       const int64_t *sumD, int64_t n_before, const uint32_t *P_prev, void *state,
                                       int64_t state_stride, cudaStream_t st);
cudaError_t pv_launch_analysis_batch(const PvDev &d, const float *in, int64_t n_in, int64_t n_frames,
                                     float *out_magphase, cudaStream_t st);
cudaError_t pv_launch_resynthesis_batch(const PvDev &d, const float *spectra, int64_t n_frames,
                                        float *back, float *out, cudaStream_t st);
cudaError_t pv_launch_resynthesis_frame(const PvDev &d, const float *back, const float *front, float *out,
                                        cudaStream_t st);
cudaError_t pv_launch_test_overlap_add(const PvDev &d, const float *in, const float *back, float *out,
                                       cudaStream_t st);
// generic (any N / hop) fused compat path
cudaError_t pv_launch_compat_generic(const PvDev &d, const PvProcessArgs &a, cudaStream_t st);
cudaError_t pv_launch_aggregate_generic(const PvDev &d, const PvAggArgs &a, cudaStream_t st);
// Turns the per-part aggregates of split corrected streams into the carried state of every part:
// part p of stream s (parts_per_stream each, table index s*parts + p) starts computing at frame ks[p];
// its state slot gets psi = (P0[a] << 32) + (ks-1)*nomS + Rq*(sum_{q<p} S_q - H_p)[a], P_prev = P_first[p].
// state_in (or null): the caller's carried-in state per stream; then psi starts from it and frame 0 counts.
cudaError_t pv_launch_split_states(const PvDev &d, int64_t n_streams, int32_t parts, const PvSegment *proc_segs,
                                   const int64_t *S, const int64_t *H, const uint32_t *P_first, unsigned char *slots,
                                   int64_t slot_stride, const unsigned char *state_in, cudaStream_t st);
// sumD[s] = sum over the parts of stream s of S; P_first from part 0; P_last from the last part.
cudaError_t pv_launch_reduce_parts(int nb, int64_t n_streams, int32_t parts, const int64_t *S, const uint32_t *Pf,
                                   const uint32_t *Pl, int64_t *sumD, uint32_t *P_first, uint32_t *P_last, cudaStream_t st);
// frame-range sharding across ranks: pack / unpack of the per-stream carry record (pv_shard_begin / pv_shard_finish)
cudaError_t pv_launch_shard_pack(int nb, int elems, int64_t n_streams, const int64_t *total, const int64_t *minus,
                                 const uint32_t *P0, int64_t *carry, cudaStream_t st);
cudaError_t pv_launch_shard_prefix(int nb, int elems, int rank, int64_t n_streams, const int64_t *carry_all,
                                   const int64_t *minus, int64_t *prefix, uint32_t *P0, cudaStream_t st);
// generic (any window) fused corrected path; a.state must be non-null (caller's or library scratch)
cudaError_t pv_launch_corrected_generic(const PvDev &d, const PvProcessArgs &a, cudaStream_t st);

// 16-bit PCM <-> float with the reference's AudioFile rules (src/AudioFile.h:1038-1049)
cudaError_t pv_launch_fft_batch(const float2 *in, float2 *out, int lg_n, int64_t batch, int dir, const float2 *tw,
                                cudaStream_t st);
cudaError_t pv_launch_pcm16_to_float(const int16_t *in, float *out, int64_t rows, int64_t pitch, int64_t c0, int64_t c1,
                                     int64_t n_valid, cudaStream_t st);
cudaError_t pv_launch_float_to_pcm16(const float *in, int16_t *out, int64_t rows, int64_t pitch, int64_t c0, int64_t c1,
                                     cudaStream_t st);
// packed 24-bit PCM (3 bytes per sample); pitch and columns in samples
cudaError_t pv_launch_pcm24_to_float(const uint8_t *in, float *out, int64_t rows, int64_t pitch, int64_t c0, int64_t c1,
                                     int64_t n_valid, cudaStream_t st);
cudaError_t pv_launch_float_to_pcm24(const float *in, uint8_t *out, int64_t rows, int64_t pitch, int64_t c0, int64_t c1,
                                     cudaStream_t st);
