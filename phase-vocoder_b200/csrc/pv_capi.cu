// pv_capi.cu -- implementation of the C ABI declared in include/pv_b200.h.
// Host-side only: parameter validation, table construction, segment planning, launches.
#include <cmath>
#include <cstdarg>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "pv_internal.h"
#include "pv_fused_tables.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct pv_handle {
    pv_params p;
    int device;
    int sm_count;
    PvDev dev;
    // owned device memory
    float *d_win = nullptr;
    float2 *d_tw = nullptr;
    uint32_t *d_nomA = nullptr;
    int32_t *d_alo = nullptr, *d_ahi = nullptr;
    uint64_t *d_nomS = nullptr;
    uint32_t *d_gather = nullptr, *d_gather_nat = nullptr;
    std::vector<float> h_win;
    // fused-kernel tables
    bool fused = false;
    int capacity = 0;
    PvFusedTables ft;
    float2 *d_ft[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // segment plan cache
    // split of corrected streams into frame-range parts (intra-GPU phase-carry scan)
    uint32_t *d_Pl = nullptr;
    int64_t *d_S = nullptr, *d_H = nullptr;
    uint32_t *d_Pf = nullptr;
    size_t carry_cap = 0;
    unsigned char *d_slots = nullptr;
    size_t slots_cap = 0;
    // Stored analysis of a frame-range split (PvAggArgs::md): {|X|, D} of every frame, [stream][frame][N/2 + 2] float2.  The
    // analysis pass writes it, the processing pass reads it back instead of repeating the forward transform.
    float2 *d_md = nullptr;
    size_t md_cap = 0;                      // float2 elements
    const void *md_valid_for = nullptr;     // plan whose analysis is in d_md (set and checked together with agg_valid_for)
    bool env_no_md = false;                 // PV_NO_MD_STORE: always recompute (test / A-B knob, read once at pv_create)
    int md_min_window = 512;                // PV_MD_MIN_WINDOW overrides (A-B knob): smallest window that stores its analysis
    // Segment tables, cached by shape.  Every entry owns its device tables, so a plan that queued launches
    // still read is never overwritten by the next shape (the pipelined host path alternates between plans).
    struct Plan {
        int kind = -1;                 // -1 free, 0 plan_segments, 1 corrected split
        int64_t n_streams = 0, n_frames = 0, skip = 0;
        int32_t flags = 0, parts = 0;
        PvSegment *d_segs = nullptr, *d_agg_segs = nullptr;
        size_t cap = 0, agg_cap = 0;
        int32_t n_segs = 0;
        uint64_t stamp = 0;
    };
    Plan plans[32];
    uint64_t plan_clock = 0;
    // test / tuning knobs, read ONCE at pv_create (not on every call): PV_NO_SPLIT, PV_FORCE_GENERIC
    bool env_no_split = false, env_force_generic = false;
    const Plan *agg_valid_for = nullptr;   // plan whose per-part sums pv_corrected_split_aggregate left in d_S / d_H / d_Pf
    float2 *d_fft_tw[14] = {};      // stand-alone FFT: n-th roots of unity per log2 n, built on first use
    // staging for the host-pointer entry point
    float *d_in = nullptr, *d_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    float *d_in16 = nullptr, *d_out16 = nullptr;      // 16-bit PCM staging (sized in floats)
    size_t in16_cap = 0, out16_cap = 0;
    void *d_state = nullptr;
    size_t state_cap = 0;
    void *d_scratch_state = nullptr;
    size_t scratch_cap = 0;
    // frame-range sharding across ranks (pv_shard_begin / pv_shard_finish): per-stream scratch of the rank
    int64_t *d_sh_total = nullptr, *d_sh_minus = nullptr;
    uint32_t *d_sh_Pm1 = nullptr, *d_sh_P0 = nullptr;
    unsigned char *d_sh_state = nullptr;
    size_t sh_cap = 0;
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};        // host path: H2D, kernels, D2H
    std::vector<cudaEvent_t> pipe_events;
    // accounting
    int64_t launches = 0;
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
};

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define PV_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(PV_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int ilog2(int n)
{
    int l = 0;
    while ((1 << l) < n) ++l;
    return l;
}

// Window tables, float arithmetic on the host as the reference does (src/phaseVocoder.h:62-69,
// 84-94): omega is a float, the cosine is the float overload.
void make_window(int type, int N, std::vector<float> &w)
{
    w.resize(N);
    if (type == PV_WIN_HANN_PERIODIC) {
        for (int i = 0; i < N; i++) w[i] = 0.5f * (1.f - cosf((float)(2.f * M_PI * i / N)));
        return;
    }
    const float omega = (float)(2.f * M_PI / (N - 1));
    for (int i = 0; i < N; i++) {
        const float c = cosf(omega * (float)i);
        w[i] = (type == PV_WIN_HAMMING) ? (0.54f - 0.46f * c) : (0.5f * (1.f - c));
    }
}

template <class T>
int upload(T **dst, const std::vector<T> &src)
{
    PV_CUDA(cudaMalloc((void **)dst, sizeof(T) * src.size()));
    PV_CUDA(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
    return PV_OK;
}

// Integer tables of the corrected mode (DESIGN.md "corrected mode"; same definitions as
// oracle/pv_oracle.h, implemented independently).
void corrected_tables(int N, int Ha, int Hs, double beta, uint64_t *Rq, int32_t *a_lo, int32_t *a_hi,
                      uint64_t *nomS)
{
    const int h = N / 2, nb = h + 1, lg = ilog2(N);
    const uint64_t bq = (uint64_t)llround(beta * 4294967296.0);
    *Rq = (bq * (uint64_t)Hs + (uint64_t)(Ha / 2)) / (uint64_t)Ha;
    for (int s = 0; s < nb; s++) { a_lo[s] = 1; a_hi[s] = 0; }
    for (int a = 0; a < nb; a++) {
        const uint64_t s = ((uint64_t)a * bq + 0x80000000ull) >> 32;
        if (s > (uint64_t)h) break;
        if (a_lo[s] > a_hi[s]) a_lo[s] = a;
        a_hi[s] = a;
    }
    for (int s = 0; s < nb; s++)
        nomS[s] = (a_lo[s] > a_hi[s]) ? 0 : ((bq * (uint64_t)a_hi[s] * (uint64_t)Hs) << (32 - lg));
}

// Looks a plan up by shape; on a miss hands out the least recently used entry (after the device has
// drained, so that no queued launch still reads the tables about to be rewritten).
pv_handle::Plan *find_plan(pv_handle *h, int kind, int64_t n_streams, int64_t n_frames, int64_t skip, int32_t flags,
                           int32_t parts, bool *hit)
{
    pv_handle::Plan *lru = &h->plans[0];
    for (auto &pl : h->plans) {
        if (pl.kind == kind && pl.n_streams == n_streams && pl.n_frames == n_frames && pl.skip == skip &&
            pl.flags == flags && pl.parts == parts) {
            pl.stamp = ++h->plan_clock;
            *hit = true;
            return &pl;
        }
        if (pl.stamp < lru->stamp) lru = &pl;
    }
    if (lru->kind >= 0) cudaDeviceSynchronize();
    lru->kind = -1;                  // valid only once the caller has filled it
    lru->n_streams = n_streams;
    lru->n_frames = n_frames;
    lru->skip = skip;
    lru->flags = flags;
    lru->parts = parts;
    lru->stamp = ++h->plan_clock;
    *hit = false;
    return lru;
}

int upload_segments(PvSegment **d, size_t *cap, const std::vector<PvSegment> &segs)
{
    if (segs.size() > *cap) {
        cudaFree(*d);
        *d = nullptr;
        *cap = 0;
        PV_CUDA(cudaMalloc((void **)d, sizeof(PvSegment) * segs.size()));
        *cap = segs.size();
    }
    // One blocking upload per new shape; every later launch of the same shape (any stream) reuses it.
    PV_CUDA(cudaMemcpy(*d, segs.data(), sizeof(PvSegment) * segs.size(), cudaMemcpyHostToDevice));
    return PV_OK;
}

// Splits every stream into frame-range segments so that the grid fills the machine.
// The (R-1)-frame OLA halo in front of each segment is recomputed (compat frames are
// independent), so the output does not depend on the split.
int plan_segments(pv_handle *h, int64_t n_streams, int64_t n_frames, int64_t skip, int32_t flags, pv_handle::Plan **out)
{
    bool hit = false;
    pv_handle::Plan *pl = find_plan(h, 0, n_streams, n_frames, skip, flags, 0, &hit);
    *out = pl;
    if (hit) return PV_OK;
    const int N = h->p.window, Hs = h->p.hop_out;
    const int64_t halo = (N - 1) / Hs;                     // frames k' < k that still overlap frame k
    // aim at ~8 waves of resident groups so that the tail wave is small, but keep the halo
    // recompute below ~3 % of a segment
    const int64_t target = (int64_t)h->capacity * 8;
    int64_t per_stream = (target + n_streams - 1) / n_streams;
    if (per_stream < 1) per_stream = 1;
    // corrected mode carries the phase accumulators from frame to frame: a stream is one segment
    // (frame-range splitting of a corrected stream goes through the phase-carry path instead)
    if (h->p.mode == PV_MODE_CORRECTED) per_stream = 1;
    int64_t seg_len = (n_frames + per_stream - 1) / per_stream;
    const int64_t min_len = std::max<int64_t>(32 * halo, 32);
    if (seg_len < min_len) seg_len = min_len;
    std::vector<PvSegment> segs;
    for (int64_t s = 0; s < n_streams; s++) {
        for (int64_t k0 = skip; k0 < n_frames; k0 += seg_len) {
            PvSegment g{};
            g.stream = (int32_t)s;
            g.state_idx = (int32_t)s;
            g.k_emit = k0;
            g.k_end = std::min(n_frames, k0 + seg_len);
            // the first segment also computes the caller's skipped (halo) frames from frame 0
            g.k_begin = (k0 == skip) ? 0 : std::max<int64_t>(0, k0 - halo);
            g.carry_in = (k0 == skip && (flags & PV_PROCESS_CARRY_IN)) ? 1 : 0;
            g.carry_out = (g.k_end == n_frames && (flags & PV_PROCESS_CARRY_OUT)) ? 1 : 0;
            segs.push_back(g);
        }
    }
    int rc = upload_segments(&pl->d_segs, &pl->cap, segs);
    if (rc != PV_OK) return rc;
    pl->n_segs = (int32_t)segs.size();
    pl->kind = 0;
    return PV_OK;
}

// Corrected mode with few streams: cut every stream into `parts` frame ranges so that the grid fills the
// machine.  Two tables (index = stream*parts + part): the analysis ranges for the phase-carry aggregate
// and the processing ranges (halo + owned frames) that start from the rebuilt state of each part.
int plan_corrected_split(pv_handle *h, int64_t n_streams, int64_t n_frames, int32_t parts, bool user_carry,
                         int64_t skip, pv_handle::Plan **out)
{
    bool hit = false;
    pv_handle::Plan *pl = find_plan(h, 1, n_streams, n_frames, skip, user_carry ? 1 : 0, parts, &hit);
    *out = pl;
    if (hit) return PV_OK;
    const int N = h->p.window, Hs = h->p.hop_out;
    const int64_t halo = (N - 1) / Hs;
    const int64_t L = (n_frames + parts - 1) / parts;
    std::vector<PvSegment> agg((size_t)(n_streams * parts)), proc((size_t)(n_streams * parts));
    for (int64_t s = 0; s < n_streams; s++)
        for (int32_t p = 0; p < parts; p++) {
            const int64_t k0 = std::min<int64_t>(n_frames, p * L), k1 = std::min<int64_t>(n_frames, k0 + L);
            const int64_t ks = std::max<int64_t>(0, k0 - halo);
            const size_t i = (size_t)(s * parts + p);
            PvSegment a{}, q{};
            a.stream = q.stream = (int32_t)s;
            a.state_idx = q.state_idx = (int32_t)i;
            a.k_begin = std::max<int64_t>(0, ks - 1);       // the frame before ks supplies P_prev
            a.k_emit = k0;
            a.k_end = k1;
            q.k_begin = ks;
            q.k_emit = std::min<int64_t>(k1, std::max<int64_t>(k0, skip));   // the caller's skipped frames are computed, not written
            q.k_end = k1;
            q.carry_in = (ks >= 1 || user_carry) ? 1 : 0;   // parts that reach frame 0 start fresh (or from the caller's state)
            a.carry_in = (p == 0) ? 1 : 0;                  // effective only when the launch supplies P_prev
            q.carry_out = (p == parts - 1) ? 1 : 0;         // the last part leaves the stream's final state in its slot
            agg[i] = a;
            proc[i] = q;
        }
    const size_t n = agg.size();
    int rc = upload_segments(&pl->d_agg_segs, &pl->agg_cap, agg);
    if (rc == PV_OK) rc = upload_segments(&pl->d_segs, &pl->cap, proc);
    if (rc != PV_OK) return rc;
    const size_t nb = (size_t)N / 2 + 1;
    if (n * nb > h->carry_cap) {
        cudaDeviceSynchronize();
        cudaFree(h->d_S); cudaFree(h->d_H); cudaFree(h->d_Pf); cudaFree(h->d_Pl);
        h->d_S = h->d_H = nullptr; h->d_Pf = h->d_Pl = nullptr; h->carry_cap = 0;
        PV_CUDA(cudaMalloc((void **)&h->d_S, sizeof(int64_t) * n * nb));
        PV_CUDA(cudaMalloc((void **)&h->d_H, sizeof(int64_t) * n * nb));
        PV_CUDA(cudaMalloc((void **)&h->d_Pf, sizeof(uint32_t) * n * nb));
        PV_CUDA(cudaMalloc((void **)&h->d_Pl, sizeof(uint32_t) * n * nb));
        h->carry_cap = n * nb;
    }
    const size_t sb = pv_state_bytes(h);
    if (n * sb > h->slots_cap) {
        cudaDeviceSynchronize();
        cudaFree(h->d_slots);
        h->d_slots = nullptr;
        h->slots_cap = 0;
        PV_CUDA(cudaMalloc((void **)&h->d_slots, n * sb));
        h->slots_cap = n * sb;
    }
    pl->n_segs = (int32_t)n;
    pl->kind = 1;
    return PV_OK;
}

// Cost model of a corrected run of n streams x F frames cut into p frame-range parts per stream, in units of the time
// one resident group needs for one full frame.  Segments run in waves of `capacity` resident groups; a split adds the
// analysis-only aggregate pass (measured ~0.45 of a full frame per frame), the recomputed overlap-add halo of every part
// and two more launches.  p = 1 is the plain one-segment-per-stream launch.
// Can a split of n streams x F frames keep its analysis ({|X|, D} per frame and bin) in device memory?  Up to 16 GB of scratch.
bool md_fits(const pv_handle *h, int64_t n, int64_t F)
{
    if (h->env_no_md || !h->fused || h->env_force_generic) return false;
    // measured (tools/md_ab.py, split runs, stored / recomputing): window 256 3.3 / 3.1 ms at one voice and 10.9 / 8.5 ms at four
    // (a 1 KB row per frame and 16-thread group: the copies cost more than the forward transform they save); 512: 1.7 / 2.2;
    // 1024: 0.78 / 0.94; 2048: 2.8 / 3.6; 4096: 8.7 / 11.6
    if (h->p.window < h->md_min_window) return false;
    return (double)n * (double)F * (double)(h->p.window / 2 + 2) * 8.0 <= 16.0 * 1073741824.0;
}

double split_cost(const pv_handle *h, int64_t n, int64_t F, int64_t p)
{
    const int64_t cap = std::max(1, h->capacity), halo = (h->p.window - 1) / h->p.hop_out;
    const int64_t L = (F + p - 1) / p;
    const double waves = (double)((n * p + cap - 1) / cap);
    if (p <= 1) return waves * (double)F;
    // One voice: the forward half (transform + analysis) is ~0.45 of a frame, a voice's synthesis half ~0.55.  V voices run in
    // launches of at most two voices (pv_fused_corrected_kernels.cu), each with its own forward half.  With the analysis stored
    // the processing pass is the synthesis halves only (+ ~0.05 per launch for reading the stored frame) and the analysis pass
    // also stores (+ ~0.05); without, the processing pass is a full frame again.
    const int V = std::max(1, h->p.n_voices), groups = V >= 3 ? (V + 1) / 2 : 1;
    const double full = 0.45 * groups + 0.55 * V;
    double proc = 1.0, agg = 0.45 / full;
    if (md_fits(h, n, F)) {
        proc = (0.55 * V + 0.05 * groups) / full;
        agg = 0.50 / full;
    }
    return waves * (proc * (double)(L + halo) + agg * (double)(L + 1)) + 3.0;
}

// number of frame-range parts per stream for a corrected run (1 = do not split): the cheapest under split_cost.
// (Round 1 split whenever fewer than two waves of streams were given, which made 300..1180 streams SLOWER than not
// splitting: profiles/r02_stream_sweep.md.)
int64_t corrected_parts(const pv_handle *h, int64_t n_streams, int64_t n_frames)
{
    if (h->env_no_split) return 1;
    const int64_t halo = (h->p.window - 1) / h->p.hop_out;
    const int64_t min_len = std::max<int64_t>(16 * (halo + 1), 32);     // keep the extra analysis + halo small
    const int64_t pmax = std::min<int64_t>(n_frames / min_len, (1 << 20) / std::max<int64_t>(1, n_streams));
    int64_t best = 1;
    double best_cost = split_cost(h, n_streams, n_frames, 1);
    for (int64_t p = 2; p <= pmax; p++) {
        const double c = split_cost(h, n_streams, n_frames, p);
        if (c < 0.92 * best_cost) {            // a split must pay clearly: it costs scratch memory and launches
            best = p;
            best_cost = c;
        }
    }
    if (best < 2) return 1;
    const int64_t L = (n_frames + best - 1) / best;
    return (n_frames + L - 1) / L;                                       // no empty trailing part
}

int ensure(float **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return PV_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc((void **)buf, need * sizeof(float));
    if (e != cudaSuccess) return fail(PV_ERR_ALLOC, "cudaMalloc(%zu floats) failed: %s", need, cudaGetErrorString(e));
    *cap = need;
    return PV_OK;
}

}  // namespace

extern "C" {

const char *pv_last_error(void) { return g_err.c_str(); }

const char *pv_version(void) { return "pv_b200 0.1 (sm_100a; in-kernel FFT, no cuFFT, no CPU fallback)"; }

int pv_create(const pv_params *params, pv_handle **out)
{
    if (!params || !out) return fail(PV_ERR_PARAM, "pv_create: null argument");
    *out = nullptr;
    const pv_params &p = *params;
    const int N = p.window;
    if (N < PV_MIN_WINDOW || N > PV_MAX_WINDOW || (N & (N - 1)))
        return fail(PV_ERR_PARAM, "window must be a power of two in [%d, %d], got %d", PV_MIN_WINDOW, PV_MAX_WINDOW, N);
    if (p.hop_in < 1) return fail(PV_ERR_PARAM, "hop_in must be >= 1, got %d", p.hop_in);
    if (p.hop_out < 1 || p.hop_out > N)
        return fail(PV_ERR_PARAM, "hop_out must be in [1, window], got %d", p.hop_out);
    if (p.mode != PV_MODE_COMPAT && p.mode != PV_MODE_CORRECTED) return fail(PV_ERR_PARAM, "bad mode %d", p.mode);
    if (p.window_type < PV_WIN_HAMMING || p.window_type > PV_WIN_HANN_PERIODIC)
        return fail(PV_ERR_PARAM, "bad window_type %d", p.window_type);
    int V = p.n_voices;
    if (p.mode == PV_MODE_COMPAT) {
        if (V != 0 && V != 1) return fail(PV_ERR_PARAM, "compat mode has exactly one voice");
        V = 1;
    } else {
        if (V < 1 || V > PV_MAX_VOICES) return fail(PV_ERR_PARAM, "n_voices must be in [1, %d]", PV_MAX_VOICES);
        for (int v = 0; v < V; v++)
            if (!(p.pitch[v] >= 0.25f && p.pitch[v] <= 4.0f))
                return fail(PV_ERR_PARAM, "pitch[%d] = %g outside [0.25, 4]", v, (double)p.pitch[v]);
    }
    int device = p.device;
    if (device < 0) PV_CUDA(cudaGetDevice(&device));
    cudaDeviceProp prop;
    PV_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(PV_ERR_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major,
                    prop.minor);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(PV_ERR_CUDA, "cannot select device %d", device);

    pv_handle *h = new pv_handle();
    h->p = p;
    h->p.n_voices = V;
    h->p.device = device;
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->env_no_split = getenv("PV_NO_SPLIT") != nullptr;
    h->env_force_generic = getenv("PV_FORCE_GENERIC") != nullptr;
    h->env_no_md = getenv("PV_NO_MD_STORE") != nullptr;
    if (const char *e = getenv("PV_MD_MIN_WINDOW")) h->md_min_window = std::max(256, atoi(e));
    make_window(p.window_type, N, h->h_win);
    std::vector<float2> tw(N);
    for (int k = 0; k < N; k++) {
        const double a = -M_PI * (double)k / (double)N;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    int rc = upload(&h->d_win, h->h_win);
    if (rc == PV_OK) rc = upload(&h->d_tw, tw);
    PvDev &d = h->dev;
    memset(&d, 0, sizeof d);
    d.N = N;
    d.lgN = ilog2(N);
    d.Ha = p.hop_in;
    d.Hs = p.hop_out;
    d.flags = p.flags;
    d.inv_N = 1.0f / (float)N;
    d.V = V;
    if (rc == PV_OK && p.mode == PV_MODE_CORRECTED) {
        const int nb = N / 2 + 1, lg = d.lgN;
        std::vector<uint32_t> nomA(nb);
        for (int b = 0; b < nb; b++) nomA[b] = (uint32_t)(((uint64_t)b * (uint64_t)p.hop_in) << (32 - lg));
        std::vector<int32_t> alo((size_t)V * nb), ahi((size_t)V * nb);
        std::vector<uint64_t> nomS((size_t)V * nb);
        for (int v = 0; v < V; v++) {
            corrected_tables(N, p.hop_in, p.hop_out, (double)p.pitch[v], &d.Rq[v], &alo[(size_t)v * nb],
                             &ahi[(size_t)v * nb], &nomS[(size_t)v * nb]);
            d.beta_q[v] = (uint64_t)llround((double)p.pitch[v] * 4294967296.0);
        }
        double s2 = 0;
        for (int i = 0; i < N; i++) s2 += (double)h->h_win[i] * (double)h->h_win[i];
        d.gain = (float)((double)p.hop_out / s2);
        rc = upload(&h->d_nomA, nomA);
        if (rc == PV_OK) rc = upload(&h->d_alo, alo);
        if (rc == PV_OK) rc = upload(&h->d_ahi, ahi);
        if (rc == PV_OK) rc = upload(&h->d_nomS, nomS);
        if (rc == PV_OK && N >= 256 && N <= 4096) {       // per-thread slot order of the fused kernels
            std::vector<uint32_t> gath;
            build_gather_table(N, V, alo.data(), ahi.data(), gath, d.multi);
            rc = upload(&h->d_gather, gath);
        }
        if (rc == PV_OK && N == 4096) {                   // natural bin order for the in-place large-window kernel (fallback path)
            std::vector<uint32_t> gath;
            build_gather_natural(N, V, alo.data(), ahi.data(), gath, d.multi);
            rc = upload(&h->d_gather_nat, gath);
        }
    }
    h->capacity = h->sm_count * 8;
    const bool want_fused = (p.mode == PV_MODE_COMPAT) ? pv_fused_compat_supported(N, p.hop_out)
                                                       : pv_fused_corrected_supported(N, p.hop_in, p.hop_out);
    if (rc == PV_OK && want_fused) {
        HostTables ht;
        build_tables(d.lgN, ht);
        const std::vector<float2> *src[7] = {&ht.tw1, &ht.tw2, &ht.tw2n, &ht.itw1, &ht.itw2, &ht.ctw1, &ht.ctw2};
        for (int i = 0; i < 7 && rc == PV_OK; i++) rc = upload(&h->d_ft[i], *src[i]);
        h->ft.tw1 = h->d_ft[0];
        h->ft.tw2 = h->d_ft[1];
        h->ft.tw2n = h->d_ft[2];
        h->ft.itw1 = h->d_ft[3];
        h->ft.itw2 = h->d_ft[4];
        h->ft.ctw1 = h->d_ft[5];
        h->ft.ctw2 = h->d_ft[6];
        h->fused = rc == PV_OK;
        if (h->fused)
            h->capacity = (p.mode == PV_MODE_COMPAT) ? pv_fused_compat_capacity(N, h->sm_count)
                                                     : pv_fused_corrected_capacity(N, V, h->sm_count);
    }
    if (rc != PV_OK) {
        pv_destroy(h);
        return rc;
    }
    d.win = h->d_win;
    d.tw = h->d_tw;
    d.nomA = h->d_nomA;
    d.a_lo = h->d_alo;
    d.a_hi = h->d_ahi;
    d.nomS = h->d_nomS;
    d.gather = h->d_gather;
    d.gather_nat = h->d_gather_nat;
    *out = h;
    return PV_OK;
}

void pv_destroy(pv_handle *h)
{
    if (!h) return;
    DeviceGuard guard(h->device);
    cudaFree(h->d_win);
    cudaFree(h->d_tw);
    cudaFree(h->d_nomA);
    cudaFree(h->d_alo);
    cudaFree(h->d_ahi);
    cudaFree(h->d_nomS);
    cudaFree(h->d_gather);
    cudaFree(h->d_gather_nat);
    for (auto p : h->d_ft) cudaFree(p);
    for (auto &pl : h->plans) {
        cudaFree(pl.d_segs);
        cudaFree(pl.d_agg_segs);
    }
    for (auto e : h->pipe_events) cudaEventDestroy(e);
    for (auto p : h->d_fft_tw) cudaFree(p);
    cudaFree(h->d_md);
    cudaFree(h->d_S);
    cudaFree(h->d_H);
    cudaFree(h->d_Pf);
    cudaFree(h->d_Pl);
    cudaFree(h->d_slots);
    cudaFree(h->d_in16);
    cudaFree(h->d_out16);
    cudaFree(h->d_in);
    cudaFree(h->d_out);
    cudaFree(h->d_state);
    cudaFree(h->d_scratch_state);
    cudaFree(h->d_sh_total);
    cudaFree(h->d_sh_minus);
    cudaFree(h->d_sh_Pm1);
    cudaFree(h->d_sh_P0);
    cudaFree(h->d_sh_state);
    for (auto ps : h->pipe)
        if (ps) cudaStreamDestroy(ps);
    for (auto &e : h->events) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    delete h;
}

int pv_get_params(const pv_handle *h, pv_params *out)
{
    if (!h || !out) return fail(PV_ERR_PARAM, "null argument");
    *out = h->p;
    return PV_OK;
}

int pv_window_table(const pv_handle *h, float *host_out)
{
    if (!h || !host_out) return fail(PV_ERR_PARAM, "null argument");
    memcpy(host_out, h->h_win.data(), sizeof(float) * h->h_win.size());
    return PV_OK;
}

int pv_reference_schedule(const pv_handle *h, int64_t num_samples, int64_t *n_analysed, int64_t *n_synth)
{
    if (!h) return fail(PV_ERR_PARAM, "null handle");
    const int64_t Ha = h->p.hop_in, Hs = h->p.hop_out;
    // for (i = 0; i < numSamples - hop; i += hop)   src/main.cpp:231
    int64_t na = 0;
    if (num_samples - Ha > 0) na = (num_samples - Ha + Ha - 1) / Ha;
    if (n_analysed) *n_analysed = na;
    if (n_synth) *n_synth = num_samples / Hs;     // src/main.cpp:266
    return PV_OK;
}

size_t pv_state_bytes(const pv_handle *h)
{
    if (!h) return 0;
    const size_t N = (size_t)h->p.window, nb = N / 2 + 1, V = (size_t)h->p.n_voices;
    if (h->p.mode == PV_MODE_COMPAT) return N * sizeof(float);
    // [have_prev u32 + pad][P_prev u32 nb (padded to 8)][psi u64 V*nb][acc f32 V*N]
    return 8 + ((nb * 4 + 7) / 8) * 8 + V * nb * 8 + V * N * 4;
}

#define PV_COMPAT_ONLY(h, what)                                                                 \
    if ((h)->p.mode != PV_MODE_COMPAT)                                                          \
        return fail(PV_ERR_PARAM, what " is the reference's per-frame contract and exists in compat mode only")

int pv_analysis(pv_handle *h, const float *in, float *out_magphase)
{
    if (!h || !in || !out_magphase) return fail(PV_ERR_PARAM, "pv_analysis: null argument");
    PV_COMPAT_ONLY(h, "pv_analysis");
    DeviceGuard guard(h->device);
    PV_CUDA(pv_launch_analysis_batch(h->dev, in, h->p.window, 1, out_magphase, nullptr));
    h->launches++;
    PV_CUDA(cudaStreamSynchronize(nullptr));     // phaseVocoder.cpp:30
    return PV_OK;
}

int pv_resynthesis(pv_handle *h, const float *back, const float *front_magphase, float *out)
{
    if (!h || !back || !front_magphase || !out) return fail(PV_ERR_PARAM, "pv_resynthesis: null argument");
    PV_COMPAT_ONLY(h, "pv_resynthesis");
    DeviceGuard guard(h->device);
    PV_CUDA(pv_launch_resynthesis_frame(h->dev, back, front_magphase, out, nullptr));
    h->launches++;
    PV_CUDA(cudaStreamSynchronize(nullptr));     // phaseVocoder.cpp:75
    return PV_OK;
}

int pv_test_overlap_add(pv_handle *h, const float *in, const float *back, float *out)
{
    if (!h || !in || !back || !out) return fail(PV_ERR_PARAM, "pv_test_overlap_add: null argument");
    DeviceGuard guard(h->device);
    PV_CUDA(pv_launch_test_overlap_add(h->dev, in, back, out, nullptr));
    h->launches++;
    PV_CUDA(cudaStreamSynchronize(nullptr));     // phaseVocoder.cpp:22
    return PV_OK;
}

int pv_analysis_batch(pv_handle *h, const float *in, int64_t n_in, int64_t n_frames, float *out_magphase,
                      void *cuda_stream)
{
    if (!h || !in || !out_magphase || n_in < 0 || n_frames < 0)
        return fail(PV_ERR_PARAM, "pv_analysis_batch: bad argument");
    PV_COMPAT_ONLY(h, "pv_analysis_batch");
    if (n_frames > 0x7fffffffLL) return fail(PV_ERR_PARAM, "too many frames for one launch");
    DeviceGuard guard(h->device);
    PV_CUDA(pv_launch_analysis_batch(h->dev, in, n_in, n_frames, out_magphase, (cudaStream_t)cuda_stream));
    if (n_frames) h->launches++;
    return PV_OK;
}

int pv_resynthesis_batch(pv_handle *h, const float *spectra, int64_t n_frames, float *back, float *out,
                         void *cuda_stream)
{
    if (!h || !spectra || !back || !out || n_frames < 0) return fail(PV_ERR_PARAM, "pv_resynthesis_batch: bad argument");
    PV_COMPAT_ONLY(h, "pv_resynthesis_batch");
    DeviceGuard guard(h->device);
    PV_CUDA(pv_launch_resynthesis_batch(h->dev, spectra, n_frames, back, out, (cudaStream_t)cuda_stream));
    if (n_frames) h->launches++;
    return PV_OK;
}

int pv_process_device(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride, int64_t n_in,
                      int64_t n_analysed, int64_t n_frames, float *out, int64_t out_stream_stride,
                      int64_t out_voice_stride, void *state, int32_t flags, void *cuda_stream)
{
    return pv_process_device_ex(h, in, n_streams, in_stride, n_in, n_analysed, n_frames, 0, out, out_stream_stride,
                                out_voice_stride, state, flags, cuda_stream);
}

static int process_impl(pv_handle *h, const float *in, int64_t n_streams, int64_t plan_streams, int64_t in_stride,
                        int64_t n_in, int64_t n_analysed, int64_t n_frames, int64_t skip_frames, float *out,
                        int64_t out_stream_stride, int64_t out_voice_stride, void *state, int32_t flags,
                        void *cuda_stream, int64_t *agg_only = nullptr);
static int aggregate_plain(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride, int64_t n_in, int64_t n_frames,
                           const uint32_t *P_prev, int64_t P_prev_stride, int32_t in_state, int64_t *sumD, uint32_t *P_first,
                           uint32_t *P_last, void *cuda_stream);
static int launch_segments(pv_handle *h, PvProcessArgs &a, int64_t n_streams, int32_t flags, cudaStream_t st);

int pv_corrected_aggregate(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride, int64_t n_in,
                           int64_t n_frames, const uint32_t *P_prev, int64_t *sumD, uint32_t *P_first,
                           uint32_t *P_last, void *cuda_stream)
{
    if (!h || !in || !sumD || n_streams < 0 || n_frames < 0) return fail(PV_ERR_PARAM, "pv_corrected_aggregate: bad argument");
    if (h->p.mode != PV_MODE_CORRECTED) return fail(PV_ERR_PARAM, "pv_corrected_aggregate needs a corrected-mode handle");
    if (n_streams == 0) return PV_OK;
    DeviceGuard guard(h->device);
    {   // few long streams: aggregate frame-range parts concurrently, then add the per-part sums
        const int64_t parts = corrected_parts(h, n_streams, n_frames);
        if (parts >= 2) {
            pv_handle::Plan *pl = nullptr;
            int rc0 = plan_corrected_split(h, n_streams, n_frames, (int32_t)parts, false, 0, &pl);
            if (rc0 != PV_OK) return rc0;
            h->agg_valid_for = nullptr;          // the shared per-part sums are about to be overwritten
            h->md_valid_for = nullptr;           // ... and the last phases that go with a stored analysis
            const int nb = h->p.window / 2 + 1;
            PvAggArgs ag{in, in_stride, n_in, pl->d_agg_segs, pl->n_segs, P_prev, (int64_t)nb, h->d_S, nullptr, h->d_Pf, h->d_Pl, 0, nullptr, 0};
            if (h->fused && !h->env_force_generic) PV_CUDA(pv_launch_corrected_aggregate(h->dev, h->ft, ag, (cudaStream_t)cuda_stream));
            else PV_CUDA(pv_launch_aggregate_generic(h->dev, ag, (cudaStream_t)cuda_stream));
            PV_CUDA(pv_launch_reduce_parts(nb, n_streams, (int32_t)parts, h->d_S, h->d_Pf, h->d_Pl, sumD, P_first, P_last,
                                           (cudaStream_t)cuda_stream));
            h->launches += 2;
            return PV_OK;
        }
    }
    return aggregate_plain(h, in, n_streams, in_stride, n_in, n_frames, P_prev, (int64_t)(h->p.window / 2 + 1), 0, sumD, P_first,
                           P_last, cuda_stream);
}

// One analysis segment per stream (no frame-range split).  P_prev rows: a dense [stream][bin] array, or the previous
// phase inside state records (in_state: the record's have_prev word gates the carry).
static int aggregate_plain(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride, int64_t n_in, int64_t n_frames,
                           const uint32_t *P_prev, int64_t P_prev_stride, int32_t in_state, int64_t *sumD, uint32_t *P_first,
                           uint32_t *P_last, void *cuda_stream)
{
    // one table per (streams, frames, carry) shape in the plan cache: a queued launch never sees its table rewritten
    bool hit = false;
    pv_handle::Plan *pl = find_plan(h, 2, n_streams, n_frames, 0, P_prev != nullptr ? 1 : 0, 0, &hit);
    if (!hit) {
        std::vector<PvSegment> segs((size_t)n_streams);
        for (int64_t s = 0; s < n_streams; s++) {
            PvSegment g{};
            g.stream = (int32_t)s;
            g.state_idx = (int32_t)s;
            g.carry_in = P_prev != nullptr;
            g.k_begin = 0;
            g.k_emit = 0;
            g.k_end = n_frames;
            segs[(size_t)s] = g;
        }
        int rc = upload_segments(&pl->d_segs, &pl->cap, segs);
        if (rc != PV_OK) return rc;
        pl->n_segs = (int32_t)n_streams;
        pl->kind = 2;
    }
    PvAggArgs a{in, in_stride, n_in, pl->d_segs, (int32_t)n_streams, P_prev, P_prev_stride, sumD, nullptr, P_first, P_last, in_state, nullptr, 0};
    if (h->fused && !h->env_force_generic) PV_CUDA(pv_launch_corrected_aggregate(h->dev, h->ft, a, (cudaStream_t)cuda_stream));
    else PV_CUDA(pv_launch_aggregate_generic(h->dev, a, (cudaStream_t)cuda_stream));
    h->launches++;
    return PV_OK;
}

int pv_corrected_split_aggregate(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride, int64_t n_in,
                                 int64_t n_frames, int64_t skip_frames, const void *state, int32_t flags, int64_t *sumD,
                                 void *cuda_stream)
{
    if (!h || !in || !sumD) return fail(PV_ERR_PARAM, "pv_corrected_split_aggregate: null argument");
    if (h->p.mode != PV_MODE_CORRECTED) return fail(PV_ERR_PARAM, "pv_corrected_split_aggregate needs a corrected-mode handle");
    return process_impl(h, in, n_streams, n_streams, in_stride, n_in, n_frames, n_frames, skip_frames, nullptr, 0, 0,
                        const_cast<void *>(state), flags & PV_PROCESS_CARRY_IN, cuda_stream, sumD);
}

int pv_corrected_state_from_carry(pv_handle *h, int64_t n_streams, const uint32_t *P_first, const int64_t *sumD,
                                  int64_t n_before, const uint32_t *P_prev, void *state, void *cuda_stream)
{
    if (!h || !state || n_streams < 0 || n_before < 0 || (n_before > 0 && (!P_first || !sumD || !P_prev)))
        return fail(PV_ERR_PARAM, "pv_corrected_state_from_carry: bad argument");
    if (h->p.mode != PV_MODE_CORRECTED)
        return fail(PV_ERR_PARAM, "pv_corrected_state_from_carry needs a corrected-mode handle");
    DeviceGuard guard(h->device);
    PV_CUDA(pv_launch_state_from_carry(h->dev, h->ft, n_streams, P_first, sumD, n_before, P_prev, state,
                                       (int64_t)pv_state_bytes(h), (cudaStream_t)cuda_stream));
    h->launches++;
    return PV_OK;
}


int pv_process_device_ex(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride, int64_t n_in,
                         int64_t n_analysed, int64_t n_frames, int64_t skip_frames, float *out,
                         int64_t out_stream_stride, int64_t out_voice_stride, void *state, int32_t flags,
                         void *cuda_stream)
{
    return process_impl(h, in, n_streams, n_streams, in_stride, n_in, n_analysed, n_frames, skip_frames, out,
                        out_stream_stride, out_voice_stride, state, flags, cuda_stream);
}

// Launches the stream kernel of the handle's mode over the segment table in `a` (timing events, launch count).
// Shared by pv_process_* (tables from the plan cache) and the real-time server (a table it owns).
static int launch_segments(pv_handle *h, PvProcessArgs &a, int64_t n_streams, int32_t flags, cudaStream_t st)
{
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->timing) {
        PV_CUDA(cudaEventCreate(&e0));
        PV_CUDA(cudaEventCreate(&e1));
        PV_CUDA(cudaEventRecord(e0, st));
    }
    const bool force_generic = h->env_force_generic;
    if (h->p.mode == PV_MODE_CORRECTED && (!h->fused || force_generic)) {
        // shape-generic kernels work on the per-stream state in global memory and always update it.  Without a
        // caller state, or with a state that is carried IN only (the contract: written only under CARRY_OUT), they
        // run on a library-owned scratch copy, so a caller can re-run from a saved state
        const bool in_only = a.state && (flags & PV_PROCESS_CARRY_IN) && !(flags & PV_PROCESS_CARRY_OUT);
        if (!a.state || in_only) {
            const size_t need = (size_t)n_streams * pv_state_bytes(h);
            if (need > h->scratch_cap) {
                if (h->scratch_cap) cudaDeviceSynchronize();     // queued launches may still use the old scratch
                cudaFree(h->d_scratch_state);
                h->d_scratch_state = nullptr;
                h->scratch_cap = 0;
                PV_CUDA(cudaMalloc(&h->d_scratch_state, need));
                h->scratch_cap = need;
            }
            if (in_only) PV_CUDA(cudaMemcpyAsync(h->d_scratch_state, a.state, need, cudaMemcpyDeviceToDevice, st));
            a.state = (unsigned char *)h->d_scratch_state;
        }
        PV_CUDA(pv_launch_corrected_generic(h->dev, a, st));
    } else if (h->p.mode == PV_MODE_CORRECTED) PV_CUDA(pv_launch_corrected_fused(h->dev, h->ft, a, st));
    else if (h->fused && !force_generic) PV_CUDA(pv_launch_compat_fused(h->dev, h->ft, a, st));
    else PV_CUDA(pv_launch_compat_generic(h->dev, a, st));
    h->launches++;
    if (h->timing) {
        PV_CUDA(cudaEventRecord(e1, st));
        h->events.emplace_back(e0, e1);
    }
    return PV_OK;
}

// `plan_streams` >= n_streams: the segment table is planned (and cached) for plan_streams streams; a call
// with fewer streams uses its stream-major prefix.  Lets the pipelined host path reuse one plan for all chunks.
static int process_impl(pv_handle *h, const float *in, int64_t n_streams, int64_t plan_streams, int64_t in_stride,
                        int64_t n_in, int64_t n_analysed, int64_t n_frames, int64_t skip_frames, float *out,
                        int64_t out_stream_stride, int64_t out_voice_stride, void *state, int32_t flags,
                        void *cuda_stream, int64_t *agg_only)
{
    if (!h || !in || (!out && !agg_only)) return fail(PV_ERR_PARAM, "pv_process: null argument");
    if (skip_frames < 0 || skip_frames > n_frames) return fail(PV_ERR_PARAM, "pv_process: bad skip_frames");
    if (n_streams < 0 || n_in < 0 || n_frames < 0 || in_stride < n_in)
        return fail(PV_ERR_PARAM, "pv_process: bad sizes (n_streams=%lld n_in=%lld n_frames=%lld in_stride=%lld)",
                    (long long)n_streams, (long long)n_in, (long long)n_frames, (long long)in_stride);
    if (n_streams > 0x7fffffffLL) return fail(PV_ERR_PARAM, "too many streams");
    if (!agg_only && out_stream_stride < (n_frames - skip_frames) * h->p.hop_out * (int64_t)h->p.n_voices && n_streams > 1)
        return fail(PV_ERR_PARAM, "pv_process: out_stream_stride too small");
    if ((flags & (PV_PROCESS_CARRY_IN | PV_PROCESS_CARRY_OUT)) && !state)
        return fail(PV_ERR_PARAM, "pv_process: carry requested without a state buffer");
    if (n_streams == 0 || n_frames == skip_frames) return PV_OK;
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    // ---- corrected mode, more streams than fit at once and a ragged last wave: run the full waves as they are and
    // the remaining streams as a second call, which is free to cut them into frame-range parts.  1300 streams on 592
    // resident groups used to take three waves for 2.2 waves of work. ----
    if (h->p.mode == PV_MODE_CORRECTED && plan_streams == n_streams && !agg_only && !h->env_no_split && h->capacity > 0 &&
        n_streams > h->capacity && n_streams % h->capacity != 0) {
        const int64_t full = n_streams / h->capacity * h->capacity, rest = n_streams - full;
        const int64_t p = corrected_parts(h, rest, n_frames);
        if (p >= 2 && split_cost(h, rest, n_frames, p) < 0.85 * (double)n_frames) {
            const int64_t sb = (int64_t)pv_state_bytes(h);
            int rc1 = process_impl(h, in, full, full, in_stride, n_in, n_analysed, n_frames, skip_frames, out, out_stream_stride,
                                   out_voice_stride, state, flags, cuda_stream, nullptr);
            if (rc1 != PV_OK) return rc1;
            return process_impl(h, in + full * in_stride, rest, rest, in_stride, n_in, n_analysed, n_frames, skip_frames,
                                out + full * out_stream_stride, out_stream_stride, out_voice_stride,
                                state ? (unsigned char *)state + full * sb : nullptr, flags, cuda_stream, nullptr);
        }
    }
    // ---- corrected mode, few streams: split into frame-range parts with an on-device phase-carry scan ----
    const bool force_generic0 = h->env_force_generic;
    if (h->p.mode == PV_MODE_CORRECTED && plan_streams == n_streams) {
        const bool user_carry = (flags & PV_PROCESS_CARRY_IN) != 0;
        const int64_t parts = corrected_parts(h, n_streams, n_frames);
        if (parts >= 2) {
            pv_handle::Plan *pl = nullptr;
            int rc0 = plan_corrected_split(h, n_streams, n_frames, (int32_t)parts, user_carry, skip_frames, &pl);
            if (rc0 != PV_OK) return rc0;
            const bool fused_ok = h->fused && !force_generic0;
            const int64_t sb = (int64_t)pv_state_bytes(h);
            PvAggArgs ag{in, in_stride, n_in, pl->d_agg_segs, pl->n_segs,
                         user_carry ? reinterpret_cast<const uint32_t *>((const unsigned char *)state + 8) : nullptr, sb / 4,
                         h->d_S, h->d_H, h->d_Pf, nullptr, 1, nullptr, 0};
            // PV_PROCESS_REUSE_AGGREGATE: pv_corrected_split_aggregate has just left the per-part sums of exactly this
            // call in the handle (same plan, same input), so the analysis pass is not repeated
            const bool reuse = (flags & PV_PROCESS_REUSE_AGGREGATE) && h->agg_valid_for == pl;
            h->agg_valid_for = nullptr;
            // Stored analysis: the analysis pass keeps {|X|, D} of every frame, the processing pass synthesises from it
            bool use_md = reuse ? h->md_valid_for == pl : (fused_ok && md_fits(h, n_streams, n_frames));
            h->md_valid_for = nullptr;
            const size_t nbp = (size_t)h->p.window / 2 + 2;
            if (use_md && !reuse) {
                const size_t need = (size_t)n_streams * (size_t)n_frames * nbp;
                if (need > h->md_cap) {
                    cudaDeviceSynchronize();
                    cudaFree(h->d_md);
                    h->d_md = nullptr;
                    h->md_cap = 0;
                    if (cudaMalloc((void **)&h->d_md, need * sizeof(float2)) == cudaSuccess) h->md_cap = need;
                    else {
                        cudaGetLastError();          // no room for the scratch: recompute instead
                        h->d_md = nullptr;
                        use_md = false;
                    }
                }
            }
            if (use_md) {
                ag.md = h->d_md;
                ag.md_stream_stride = (int64_t)((size_t)n_frames * nbp);
                ag.P_last = h->d_Pl;
            }
            if (!reuse) {
                if (fused_ok) PV_CUDA(pv_launch_corrected_aggregate(h->dev, h->ft, ag, st));
                else PV_CUDA(pv_launch_aggregate_generic(h->dev, ag, st));
                h->launches++;
            }
            if (agg_only) {
                // totals over all frames of the call (halo and skipped frames included): sum of the per-part sums
                PV_CUDA(pv_launch_reduce_parts(h->p.window / 2 + 1, n_streams, (int32_t)parts, h->d_S, nullptr, nullptr, agg_only,
                                               nullptr, nullptr, st));
                h->launches++;
                h->agg_valid_for = pl;
                h->md_valid_for = use_md ? pl : nullptr;
                return PV_OK;
            }
            PV_CUDA(pv_launch_split_states(h->dev, n_streams, (int32_t)parts, pl->d_segs, h->d_S, h->d_H, h->d_Pf,
                                           h->d_slots, sb, user_carry ? (const unsigned char *)state : nullptr, st));
            PvProcessArgs a{};
            a.in = in;
            a.in_stride = in_stride;
            a.n_in = n_in;
            a.n_analysed = n_analysed;
            a.n_frames = n_frames;
            a.out = out - skip_frames * (int64_t)h->p.hop_out;     // kernels index the output by frame number
            a.out_stream_stride = out_stream_stride;
            a.out_voice_stride = out_voice_stride;
            a.state = h->d_slots;
            a.state_stride = sb;
            a.segs = pl->d_segs;
            a.n_segs = pl->n_segs;
            if (use_md) {
                a.md = h->d_md;
                a.md_stream_stride = (int64_t)((size_t)n_frames * nbp);
                a.P_last = h->d_Pl;
            }
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (h->timing) {
                PV_CUDA(cudaEventCreate(&e0));
                PV_CUDA(cudaEventCreate(&e1));
                PV_CUDA(cudaEventRecord(e0, st));
            }
            if (fused_ok) PV_CUDA(pv_launch_corrected_fused(h->dev, h->ft, a, st));
            else PV_CUDA(pv_launch_corrected_generic(h->dev, a, st));
            if (h->timing) {
                PV_CUDA(cudaEventRecord(e1, st));
                h->events.emplace_back(e0, e1);
            }
            h->launches += 2;
            if ((flags & PV_PROCESS_CARRY_OUT) && state) {
                // the last part of every stream holds the stream's final state (the generic kernel always
                // leaves it in the slot; the fused kernel writes it when carry_out is set -> set below)
                const size_t last = (size_t)(parts - 1);
                PV_CUDA(cudaMemcpy2DAsync(state, (size_t)sb, h->d_slots + last * (size_t)sb, (size_t)sb * (size_t)parts, (size_t)sb,
                                          (size_t)n_streams, cudaMemcpyDeviceToDevice, st));
            }
            return PV_OK;
        }
    }
    if (agg_only) {
        // no frame-range split for this shape: the totals are one plain analysis pass (nothing is kept for reuse)
        const uint32_t *P_prev = (flags & PV_PROCESS_CARRY_IN) && state
                                     ? reinterpret_cast<const uint32_t *>((const unsigned char *)state + 8) : nullptr;
        return aggregate_plain(h, in, n_streams, in_stride, n_in, n_frames, P_prev, (int64_t)pv_state_bytes(h) / 4, 1, agg_only,
                               nullptr, nullptr, cuda_stream);
    }
    pv_handle::Plan *pl = nullptr;
    int rc = plan_segments(h, plan_streams, n_frames, skip_frames, flags, &pl);
    if (rc != PV_OK) return rc;
    PvProcessArgs a{};
    a.in = in;
    a.in_stride = in_stride;
    a.n_in = n_in;
    a.n_analysed = n_analysed;
    a.n_frames = n_frames;
    a.out = out - skip_frames * (int64_t)h->p.hop_out;     // kernels index the output by frame number
    a.out_stream_stride = out_stream_stride;
    a.out_voice_stride = out_voice_stride;
    a.state = (unsigned char *)state;
    a.state_stride = (int64_t)pv_state_bytes(h);
    a.segs = pl->d_segs;
    a.n_segs = (int32_t)((int64_t)pl->n_segs / plan_streams * n_streams);      // stream-major table
    return launch_segments(h, a, n_streams, flags, st);
}


// ---------------------------------------------------------------------------------------------------
// Frame-range sharding across ranks (include/pv_b200.h, "Frame-range sharding of long streams")
// ---------------------------------------------------------------------------------------------------
int pv_shard_plan_frames(const pv_handle *h, int64_t n_frames, int32_t world, int32_t rank, pv_shard_plan *out)
{
    if (!h || !out || n_frames < 0 || world < 1 || rank < 0 || rank >= world) return fail(PV_ERR_PARAM, "pv_shard_plan_frames: bad argument");
    const int64_t per = (n_frames + world - 1) / world;
    out->k0 = std::min<int64_t>(n_frames, (int64_t)rank * per);
    out->k1 = std::min<int64_t>(n_frames, out->k0 + per);
    out->halo = (h->p.window - 1) / h->p.hop_out;
    out->ks = std::max<int64_t>(0, out->k0 - out->halo);
    return PV_OK;
}

size_t pv_shard_carry_elems(const pv_handle *h)
{
    if (!h) return 0;
    const size_t nb = (size_t)h->p.window / 2 + 1;
    return nb + (nb + 1) / 2;
}

static int shard_scratch(pv_handle *h, int64_t n_streams)
{
    const size_t nb = (size_t)h->p.window / 2 + 1, need = (size_t)n_streams * nb;
    if (need <= h->sh_cap) return PV_OK;
    if (h->sh_cap) cudaDeviceSynchronize();
    cudaFree(h->d_sh_total); cudaFree(h->d_sh_minus); cudaFree(h->d_sh_Pm1); cudaFree(h->d_sh_P0); cudaFree(h->d_sh_state);
    h->d_sh_total = h->d_sh_minus = nullptr; h->d_sh_Pm1 = h->d_sh_P0 = nullptr; h->d_sh_state = nullptr; h->sh_cap = 0;
    PV_CUDA(cudaMalloc((void **)&h->d_sh_total, sizeof(int64_t) * need));
    PV_CUDA(cudaMalloc((void **)&h->d_sh_minus, sizeof(int64_t) * need));
    PV_CUDA(cudaMalloc((void **)&h->d_sh_Pm1, sizeof(uint32_t) * need));
    PV_CUDA(cudaMalloc((void **)&h->d_sh_P0, sizeof(uint32_t) * need));
    PV_CUDA(cudaMalloc((void **)&h->d_sh_state, (size_t)n_streams * pv_state_bytes(h)));
    h->sh_cap = need;
    return PV_OK;
}

int pv_shard_begin(pv_handle *h, const float *in, int64_t in_first_frame, int64_t n_streams, int64_t in_stride,
                   int64_t n_in, int64_t n_frames, int32_t world, int32_t rank, int64_t *carry_send, void *cuda_stream)
{
    if (!h || !in || !carry_send || n_streams < 0 || in_first_frame < 0) return fail(PV_ERR_PARAM, "pv_shard_begin: bad argument");
    pv_shard_plan p;
    int rc = pv_shard_plan_frames(h, n_frames, world, rank, &p);
    if (rc != PV_OK) return rc;
    if (n_streams == 0) return PV_OK;
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int nb = h->p.window / 2 + 1, elems = (int)pv_shard_carry_elems(h);
    const int64_t Ha = h->p.hop_in;
    if (h->p.mode != PV_MODE_CORRECTED || p.k1 <= p.k0) {        // nothing to contribute: compat mode, or an empty range
        PV_CUDA(cudaMemsetAsync(carry_send, 0, sizeof(int64_t) * (size_t)n_streams * (size_t)elems, st));
        return PV_OK;
    }
    if (in_first_frame > std::max<int64_t>(0, p.ks - 1))
        return fail(PV_ERR_PARAM, "pv_shard_begin: rank %d needs the input from frame %lld on, got it from frame %lld", rank,
                    (long long)std::max<int64_t>(0, p.ks - 1), (long long)in_first_frame);
    rc = shard_scratch(h, n_streams);
    if (rc != PV_OK) return rc;
    auto at = [&](int64_t frame) { return in + (frame - in_first_frame) * Ha; };
    auto left = [&](int64_t frame) { return std::max<int64_t>(0, n_in - (frame - in_first_frame) * Ha); };
    if (p.ks == 0) {
        // the range reaches frame 0: fresh start.  D over [1, k1) from the analysis pass of the processing call itself,
        // minus what lies before k0 (another rank reports that)
        rc = process_impl(h, at(0), n_streams, n_streams, in_stride, left(0), p.k1, p.k1, p.k0, nullptr, 0, 0, nullptr, 0,
                          cuda_stream, h->d_sh_total);
        if (rc != PV_OK) return rc;
        const pv_handle::Plan *keep = h->agg_valid_for;
        const bool before = p.k0 > 1;
        if (before) rc = pv_corrected_aggregate(h, at(0), n_streams, in_stride, left(0), p.k0, nullptr, h->d_sh_minus, nullptr, nullptr, cuda_stream);
        if (rc == PV_OK && p.k0 == 0)
            rc = pv_corrected_aggregate(h, at(0), n_streams, in_stride, left(0), 1, nullptr, h->d_sh_minus, h->d_sh_P0, nullptr, cuda_stream);
        if (rc != PV_OK) return rc;
        h->agg_valid_for = keep;       // the small aggregates above ran unsplit: the per-part sums of the range are still there
        PV_CUDA(pv_launch_shard_pack(nb, elems, n_streams, h->d_sh_total, before ? h->d_sh_minus : nullptr,
                                     p.k0 == 0 ? h->d_sh_P0 : nullptr, carry_send, st));
        h->launches++;
        return PV_OK;
    }
    // D over the halo [ks, k0) and the phase of frame ks-1, from a handful of frames
    rc = pv_corrected_aggregate(h, at(p.ks - 1), n_streams, in_stride, left(p.ks - 1), p.k0 - p.ks + 1, nullptr, h->d_sh_minus,
                                h->d_sh_Pm1, nullptr, cuda_stream);
    if (rc != PV_OK) return rc;
    // a state that holds only the previous phase (accumulators unknown yet), so that frame ks has a phase difference
    PV_CUDA(cudaMemsetAsync(h->d_sh_total, 0, sizeof(int64_t) * (size_t)n_streams * nb, st));
    PV_CUDA(cudaMemsetAsync(h->d_sh_P0, 0, sizeof(uint32_t) * (size_t)n_streams * nb, st));
    PV_CUDA(pv_launch_state_from_carry(h->dev, h->ft, n_streams, h->d_sh_P0, h->d_sh_total, 1, h->d_sh_Pm1, h->d_sh_state,
                                       (int64_t)pv_state_bytes(h), st));
    h->launches++;
    rc = process_impl(h, at(p.ks), n_streams, n_streams, in_stride, left(p.ks), p.k1 - p.ks, p.k1 - p.ks, p.k0 - p.ks, nullptr, 0, 0,
                      h->d_sh_state, PV_PROCESS_CARRY_IN, cuda_stream, h->d_sh_total);       // D over [ks, k1)
    if (rc != PV_OK) return rc;
    PV_CUDA(pv_launch_shard_pack(nb, elems, n_streams, h->d_sh_total, h->d_sh_minus, nullptr, carry_send, st));
    h->launches++;
    return PV_OK;
}

int pv_shard_finish(pv_handle *h, const float *in, int64_t in_first_frame, int64_t n_streams, int64_t in_stride,
                    int64_t n_in, int64_t n_analysed, int64_t n_frames, int32_t world, int32_t rank,
                    const int64_t *carry_all, float *out, int64_t out_stream_stride, int64_t out_voice_stride,
                    void *cuda_stream)
{
    if (!h || !in || !out || n_streams < 0 || in_first_frame < 0) return fail(PV_ERR_PARAM, "pv_shard_finish: bad argument");
    pv_shard_plan p;
    int rc = pv_shard_plan_frames(h, n_frames, world, rank, &p);
    if (rc != PV_OK) return rc;
    if (n_streams == 0 || p.k1 <= p.k0) return PV_OK;
    if (in_first_frame > std::max<int64_t>(0, p.ks - (h->p.mode == PV_MODE_CORRECTED ? 1 : 0)))
        return fail(PV_ERR_PARAM, "pv_shard_finish: the input must start at or before frame %lld", (long long)std::max<int64_t>(0, p.ks - 1));
    DeviceGuard guard(h->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int64_t Ha = h->p.hop_in;
    const float *x = in + (p.ks - in_first_frame) * Ha;
    const int64_t n_left = std::max<int64_t>(0, n_in - (p.ks - in_first_frame) * Ha);
    const int64_t nf = p.k1 - p.ks, skip = p.k0 - p.ks;
    const int64_t an = std::max<int64_t>(0, std::min<int64_t>(nf, n_analysed - p.ks));
    if (h->p.mode != PV_MODE_CORRECTED)        // compat: frames are independent, the halo is recomputed
        return process_impl(h, x, n_streams, n_streams, in_stride, n_left, an, nf, skip, out, out_stream_stride, out_voice_stride,
                            nullptr, 0, cuda_stream);
    if (p.ks == 0)
        return process_impl(h, x, n_streams, n_streams, in_stride, n_left, nf, nf, skip, out, out_stream_stride, out_voice_stride,
                            nullptr, PV_PROCESS_REUSE_AGGREGATE, cuda_stream);
    if (!carry_all) return fail(PV_ERR_PARAM, "pv_shard_finish: corrected mode needs the gathered carry");
    const int nb = h->p.window / 2 + 1, elems = (int)pv_shard_carry_elems(h);
    if ((size_t)n_streams * nb > h->sh_cap) return fail(PV_ERR_PARAM, "pv_shard_finish without a matching pv_shard_begin");
    // prefix up to ks = everything the ranks before this one own, minus the D of the halo frames [ks, k0)
    PV_CUDA(pv_launch_shard_prefix(nb, elems, rank, n_streams, carry_all, h->d_sh_minus, h->d_sh_total, h->d_sh_P0, st));
    PV_CUDA(pv_launch_state_from_carry(h->dev, h->ft, n_streams, h->d_sh_P0, h->d_sh_total, p.ks, h->d_sh_Pm1, h->d_sh_state,
                                       (int64_t)pv_state_bytes(h), st));
    h->launches += 2;
    return process_impl(h, x, n_streams, n_streams, in_stride, n_left, nf, nf, skip, out, out_stream_stride, out_voice_stride,
                        h->d_sh_state, PV_PROCESS_CARRY_IN | PV_PROCESS_REUSE_AGGREGATE, cuda_stream);
}

}  // extern "C"

// Shared body of the host entry points.  PCM = 2 / 3: the host buffers hold 16-bit / packed 24-bit PCM and the conversions
// of the reference's AudioFile (16 bit: s/32768 in, trunc(clamp(x,-1,1)*32767) out, src/AudioFile.h:1038-1049; 24 bit:
// sign-extend, /8388608 in, (int32)(x*8388608) out, :508-518, :755-766) run on the device, which halves / cuts by a quarter
// the bytes crossing PCIe in both directions.  PCM = 0: float samples.
template <int PCM>
static int process_host_impl(pv_handle *h, const void *in_v, int64_t n_streams, int64_t in_stride, int64_t n_in,
                             int64_t n_analysed, int64_t n_frames, void *out_v, int64_t out_stream_stride,
                             int64_t out_voice_stride, void *state, int32_t flags)
{
    if (!h || !in_v || !out_v) return fail(PV_ERR_PARAM, "pv_process_host: null argument");
    if (n_streams <= 0 || n_frames <= 0) return PV_OK;
    {   // validate the layout before any copy is queued
        const int64_t need_out = n_frames * h->p.hop_out;
        if (n_in < 0 || in_stride < n_in)
            return fail(PV_ERR_PARAM, "pv_process_host: bad input layout (n_in=%lld in_stride=%lld)", (long long)n_in, (long long)in_stride);
        if (h->p.n_voices > 1 && out_voice_stride < need_out)
            return fail(PV_ERR_PARAM, "pv_process_host: out_voice_stride %lld < n_frames*hop_out = %lld", (long long)out_voice_stride,
                        (long long)need_out);
        if (n_streams > 1 && out_stream_stride < need_out + (int64_t)(h->p.n_voices - 1) * out_voice_stride)
            return fail(PV_ERR_PARAM, "pv_process_host: out_stream_stride %lld too small", (long long)out_stream_stride);
    }
    if ((flags & (PV_PROCESS_CARRY_IN | PV_PROCESS_CARRY_OUT)) && !state)
        return fail(PV_ERR_PARAM, "pv_process_host: carry requested without a state buffer");
    DeviceGuard guard(h->device);
    const int64_t V = h->p.n_voices, n_out = n_frames * h->p.hop_out;
    constexpr bool PCM16 = PCM != 0;          // any integer sample type: staged in d_in16 / d_out16 and converted on the device
    const size_t esz = PCM ? (size_t)PCM : 4;
    // device rows are padded to 4 samples so that the kernels keep their 16-byte aligned fast paths
    const int64_t n_in_p = (n_in + 3) & ~int64_t(3);
    int rc = ensure(&h->d_in, &h->in_cap, std::max<size_t>(4, (size_t)(n_streams * n_in_p)));   // n_in == 0: all-zero input, not a null pointer
    if (rc == PV_OK) rc = ensure(&h->d_out, &h->out_cap, (size_t)(n_streams * V * n_out));
    if (rc == PV_OK && PCM16) rc = ensure(&h->d_in16, &h->in16_cap, ((size_t)(n_streams * n_in_p) * esz + 3) / 4);
    if (rc == PV_OK && PCM16) rc = ensure(&h->d_out16, &h->out16_cap, ((size_t)(n_streams * V * n_out) * esz + 3) / 4);
    if (rc != PV_OK) return rc;
    for (auto &ps : h->pipe)
        if (!ps) PV_CUDA(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    cudaStream_t s_in = h->pipe[0], s_k = h->pipe[1], s_out = h->pipe[2];
    // Software pipeline over chunks of FRAMES (all streams each): the H2D copy of chunk c+1, the kernels of
    // chunk c and the D2H copy of chunk c-1 overlap (PCIe is full duplex), so a large batch costs about
    // max(H2D, D2H, kernel) instead of their sum.  Every chunk keeps the whole batch's parallelism, and the
    // per-stream state (phase accumulators + overlap-add tail) is carried from chunk to chunk on the device,
    // which is the same carry a caller streaming block by block uses -- the result is bit-identical to one
    // unchunked call.  Pinned host buffers are needed for real overlap.
    const int N = h->p.window, Ha = h->p.hop_in, Hs = h->p.hop_out;
    const int64_t bytes_in = n_streams * n_in * (int64_t)esz;
    // ~32 MB per chunk and direction, at most 64 chunks.  Measured on the headline batch (tools/host_chunks_probe.py,
    // tools/pcm16_probe.py): the copy engines of this box have slow spots by copy size -- float: 64 chunks of 33 MB take 45 ms
    // where 32 chunks of 65 MB take 54 ms; 16-bit PCM: 32 chunks of 33 MB take 22.3 ms, 43 chunks of 24 MB 26.6 ms, 64 chunks of
    // 16 MB 23.0 ms -- so both sample types aim at the ~33 MB chunk.
    int64_t n_chunks = std::min<int64_t>(64, std::max<int64_t>(1, bytes_in / (32ll << 20)));
    if (const char *e = getenv("PV_HOST_CHUNKS")) n_chunks = std::max(1, atoi(e));      // test / tuning knob
    const int64_t min_fc = std::max<int64_t>(8, 4 * ((N + Ha - 1) / Ha));
    int64_t fc = std::max(min_fc, (n_frames + n_chunks - 1) / n_chunks);
    n_chunks = (n_frames + fc - 1) / fc;
    const bool use_state = n_chunks > 1 || state != nullptr;
    const size_t sb = pv_state_bytes(h);
    if (use_state && (size_t)n_streams * sb > h->state_cap) {
        cudaFree(h->d_state);
        h->d_state = nullptr;
        h->state_cap = 0;
        PV_CUDA(cudaMalloc(&h->d_state, (size_t)n_streams * sb));
        h->state_cap = (size_t)n_streams * sb;
    }
    while ((int64_t)h->pipe_events.size() < 2 * n_chunks) {
        cudaEvent_t e = nullptr;
        PV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->pipe_events.push_back(e);
    }
    int32_t cflags = 0;
    if (use_state) {
        if (flags & PV_PROCESS_CARRY_IN)
            PV_CUDA(cudaMemcpyAsync(h->d_state, state, (size_t)n_streams * sb, cudaMemcpyHostToDevice, s_k));
        else if (n_chunks > 1)          // an all-zero state is a fresh start
            PV_CUDA(cudaMemsetAsync(h->d_state, 0, (size_t)n_streams * sb, s_k));
        cflags = n_chunks > 1 ? (PV_PROCESS_CARRY_IN | PV_PROCESS_CARRY_OUT) : flags;
    }
    const unsigned char *in = (const unsigned char *)in_v;
    unsigned char *out = (unsigned char *)out_v;
    unsigned char *d16 = reinterpret_cast<unsigned char *>(h->d_in16), *o16 = reinterpret_cast<unsigned char *>(h->d_out16);
    int64_t e0 = 0;                                       // input samples already on the device
    // inside the pipeline a failed call must not return before the streams have drained: the caller's buffers are
    // still the source / target of copies in flight
#define PIPE_CUDA(call)                                                                                         \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess) {                                                                                \
            rc = fail(PV_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            goto drain;                                                                                         \
        }                                                                                                       \
    } while (0)
    for (int64_t c = 0; c < n_chunks && rc == PV_OK; ++c) {
        const int64_t k0 = c * fc, k1 = std::min(n_frames, k0 + fc);
        const int64_t e1 = std::max(e0, std::min(n_in, (k1 - 1) * Ha + N));
        cudaEvent_t ev_in = h->pipe_events[(size_t)(2 * c)], ev_k = h->pipe_events[(size_t)(2 * c + 1)];
        if (e1 > e0) {
            if (PCM16)
                PIPE_CUDA(cudaMemcpy2DAsync(d16 + (size_t)e0 * esz, esz * n_in_p, in + (size_t)e0 * esz, esz * in_stride,
                                          esz * (size_t)(e1 - e0), (size_t)n_streams, cudaMemcpyHostToDevice, s_in));
            else
                PIPE_CUDA(cudaMemcpy2DAsync(h->d_in + e0, 4 * n_in_p, in + (size_t)e0 * 4, 4 * in_stride, 4 * (size_t)(e1 - e0),
                                          (size_t)n_streams, cudaMemcpyHostToDevice, s_in));
        }
        PIPE_CUDA(cudaEventRecord(ev_in, s_in));
        PIPE_CUDA(cudaStreamWaitEvent(s_k, ev_in, 0));
        if (PCM16 && e1 > e0) {
            if (PCM == 2)
                PIPE_CUDA(pv_launch_pcm16_to_float(reinterpret_cast<const int16_t *>(d16), h->d_in, n_streams, n_in_p,
                                                   e0 & ~int64_t(3), e1, n_in, s_k));
            else
                PIPE_CUDA(pv_launch_pcm24_to_float(d16, h->d_in, n_streams, n_in_p, e0 & ~int64_t(3), e1, n_in, s_k));
            h->launches++;
        }
        e0 = e1;
        const int64_t an = std::min(k1 - k0, std::max<int64_t>(0, n_analysed - k0));
        rc = process_impl(h, h->d_in + k0 * Ha, n_streams, n_streams, n_in_p, std::max<int64_t>(0, n_in - k0 * Ha), an, k1 - k0, 0,
                          h->d_out + k0 * Hs, V * n_out, n_out, use_state ? h->d_state : nullptr, cflags, s_k);
        if (rc != PV_OK) goto drain;
        if (PCM16) {
            if (PCM == 2)
                PIPE_CUDA(pv_launch_float_to_pcm16(h->d_out, reinterpret_cast<int16_t *>(o16), n_streams * V, n_out, k0 * Hs, k1 * Hs, s_k));
            else
                PIPE_CUDA(pv_launch_float_to_pcm24(h->d_out, o16, n_streams * V, n_out, k0 * Hs, k1 * Hs, s_k));
            h->launches++;
        }
        PIPE_CUDA(cudaEventRecord(ev_k, s_k));
        PIPE_CUDA(cudaStreamWaitEvent(s_out, ev_k, 0));
        const size_t w = (size_t)(k1 - k0) * Hs;
        for (int64_t v = 0; v < V; v++) {
            if (PCM16)
                PIPE_CUDA(cudaMemcpy2DAsync(out + ((size_t)v * out_voice_stride + (size_t)k0 * Hs) * esz, esz * out_stream_stride,
                                          o16 + (size_t)(v * n_out + k0 * Hs) * esz, esz * V * n_out, esz * w, (size_t)n_streams,
                                          cudaMemcpyDeviceToHost, s_out));
            else
                PIPE_CUDA(cudaMemcpy2DAsync(out + ((size_t)v * out_voice_stride + (size_t)k0 * Hs) * 4, 4 * out_stream_stride,
                                          h->d_out + v * n_out + k0 * Hs, 4 * V * n_out, 4 * w, (size_t)n_streams,
                                          cudaMemcpyDeviceToHost, s_out));
        }
    }
    if (rc == PV_OK && use_state && (flags & PV_PROCESS_CARRY_OUT))
        PIPE_CUDA(cudaMemcpyAsync(state, h->d_state, (size_t)n_streams * sb, cudaMemcpyDeviceToHost, s_k));
#undef PIPE_CUDA
drain:
    for (auto &ps : h->pipe) {
        cudaError_t e = cudaStreamSynchronize(ps);
        if (e != cudaSuccess && rc == PV_OK) rc = fail(PV_ERR_CUDA, "pipeline stream: %s", cudaGetErrorString(e));
    }
    return rc;
}

extern "C" {

int pv_process_host(pv_handle *h, const float *in, int64_t n_streams, int64_t in_stride, int64_t n_in,
                    int64_t n_analysed, int64_t n_frames, float *out, int64_t out_stream_stride,
                    int64_t out_voice_stride, void *state, int32_t flags)
{
    return process_host_impl<0>(h, in, n_streams, in_stride, n_in, n_analysed, n_frames, out, out_stream_stride,
                                    out_voice_stride, state, flags);
}

int pv_process_host_pcm16(pv_handle *h, const int16_t *in, int64_t n_streams, int64_t in_stride, int64_t n_in,
                          int64_t n_analysed, int64_t n_frames, int16_t *out, int64_t out_stream_stride,
                          int64_t out_voice_stride, void *state, int32_t flags)
{
    return process_host_impl<2>(h, in, n_streams, in_stride, n_in, n_analysed, n_frames, out, out_stream_stride,
                                   out_voice_stride, state, flags);
}

int pv_process_host_pcm24(pv_handle *h, const uint8_t *in, int64_t n_streams, int64_t in_stride, int64_t n_in,
                          int64_t n_analysed, int64_t n_frames, uint8_t *out, int64_t out_stream_stride,
                          int64_t out_voice_stride, void *state, int32_t flags)
{
    return process_host_impl<3>(h, in, n_streams, in_stride, n_in, n_analysed, n_frames, out, out_stream_stride,
                                out_voice_stride, state, flags);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// Real-time block server (include/pv_b200.h, "Real-time block server")
// ---------------------------------------------------------------------------------------------------
struct pv_rt {
    pv_handle *h = nullptr;
    int64_t S = 0, hist = 0, in_w = 0, out_w = 0, row = 0;
    int32_t B = 0;
    float *d_buf[2] = {nullptr, nullptr};     // [S][row]: ring history (N-Ha) followed by the new block
    float *d_out = nullptr;
    void *d_state = nullptr;
    // The server's OWN segment table (one segment per stream: frames [0, B), state carried in and out).  The
    // recorded graphs bake device pointers in, so nothing they reference may live in the handle's shared caches
    // (plan LRU, frame-range-split scratch): other shapes run on the same handle between two blocks would evict
    // or reallocate those and the next replay would read a foreign table or freed memory.
    PvSegment *d_segs = nullptr;
    float *h_in = nullptr, *h_out = nullptr;  // page-locked staging
    cudaStream_t st = nullptr;
    cudaGraphExec_t exec[2] = {nullptr, nullptr};
    int cur = 0;
    int kernels_per_step = 0;
};

namespace {

// One block on the server's stream: the four operations that the graph records.
int rt_enqueue(pv_rt *rt, int p)
{
    pv_handle *h = rt->h;
    const int64_t V = h->p.n_voices;
    PV_CUDA(cudaMemcpy2DAsync(rt->d_buf[p] + rt->hist, sizeof(float) * rt->row, rt->h_in, sizeof(float) * rt->in_w,
                              sizeof(float) * rt->in_w, (size_t)rt->S, cudaMemcpyHostToDevice, rt->st));
    PvProcessArgs a{};
    a.in = rt->d_buf[p];
    a.in_stride = rt->row;
    a.n_in = rt->hist + rt->in_w;
    a.n_analysed = rt->B;
    a.n_frames = rt->B;
    a.out = rt->d_out;
    a.out_stream_stride = V * rt->out_w;
    a.out_voice_stride = rt->out_w;
    a.state = (unsigned char *)rt->d_state;
    a.state_stride = (int64_t)pv_state_bytes(h);
    a.segs = rt->d_segs;
    a.n_segs = (int32_t)rt->S;
    int rc = launch_segments(h, a, rt->S, PV_PROCESS_CARRY_IN | PV_PROCESS_CARRY_OUT, rt->st);
    if (rc != PV_OK) return rc;
    if (rt->hist > 0)       // ring advance: the last N-Ha samples become the next block's history
        PV_CUDA(cudaMemcpy2DAsync(rt->d_buf[1 - p], sizeof(float) * rt->row, rt->d_buf[p] + rt->in_w, sizeof(float) * rt->row,
                                  sizeof(float) * rt->hist, (size_t)rt->S, cudaMemcpyDeviceToDevice, rt->st));
    PV_CUDA(cudaMemcpyAsync(rt->h_out, rt->d_out, sizeof(float) * (size_t)(rt->S * V * rt->out_w), cudaMemcpyDeviceToHost, rt->st));
    return PV_OK;
}

int rt_clear(pv_rt *rt)
{
    PV_CUDA(cudaMemsetAsync(rt->d_buf[0], 0, sizeof(float) * (size_t)(rt->S * rt->row), rt->st));
    PV_CUDA(cudaMemsetAsync(rt->d_buf[1], 0, sizeof(float) * (size_t)(rt->S * rt->row), rt->st));
    PV_CUDA(cudaMemsetAsync(rt->d_state, 0, (size_t)rt->S * pv_state_bytes(rt->h), rt->st));   // all-zero state = fresh start
    PV_CUDA(cudaStreamSynchronize(rt->st));
    rt->cur = 0;
    return PV_OK;
}

}  // namespace

extern "C" {

void pv_rt_close(pv_rt *rt)
{
    if (!rt) return;
    DeviceGuard guard(rt->h->device);
    if (rt->st) cudaStreamSynchronize(rt->st);
    for (auto &e : rt->exec)
        if (e) cudaGraphExecDestroy(e);
    cudaFree(rt->d_buf[0]);
    cudaFree(rt->d_buf[1]);
    cudaFree(rt->d_out);
    cudaFree(rt->d_state);
    cudaFree(rt->d_segs);
    cudaFreeHost(rt->h_in);
    cudaFreeHost(rt->h_out);
    if (rt->st) cudaStreamDestroy(rt->st);
    delete rt;
}

int pv_rt_open(pv_handle *h, int64_t n_streams, int32_t block_frames, pv_rt **out)
{
    if (!h || !out || n_streams <= 0 || n_streams > 0x7fffffffLL || block_frames <= 0)
        return fail(PV_ERR_PARAM, "pv_rt_open: bad argument");
    *out = nullptr;
    DeviceGuard guard(h->device);
    pv_rt *rt = new pv_rt;
    rt->h = h;
    rt->S = n_streams;
    rt->B = block_frames;
    const int64_t V = h->p.n_voices;
    rt->hist = h->p.window - h->p.hop_in;
    rt->in_w = (int64_t)block_frames * h->p.hop_in;
    rt->out_w = (int64_t)block_frames * h->p.hop_out;
    rt->row = (rt->hist + rt->in_w + 3) & ~int64_t(3);
    int rc = PV_OK;
    auto ck = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == PV_OK) rc = fail(PV_ERR_CUDA, "pv_rt_open: %s: %s", what, cudaGetErrorString(e));
    };
    ck(cudaStreamCreateWithFlags(&rt->st, cudaStreamNonBlocking), "stream");
    ck(cudaMalloc((void **)&rt->d_buf[0], sizeof(float) * (size_t)(rt->S * rt->row)), "ring");
    ck(cudaMalloc((void **)&rt->d_buf[1], sizeof(float) * (size_t)(rt->S * rt->row)), "ring");
    ck(cudaMalloc((void **)&rt->d_out, sizeof(float) * (size_t)(rt->S * V * rt->out_w)), "output");
    ck(cudaMalloc(&rt->d_state, (size_t)rt->S * pv_state_bytes(h)), "state");
    {
        std::vector<PvSegment> segs((size_t)rt->S);
        for (int64_t s = 0; s < rt->S; s++) {
            PvSegment g{};
            g.stream = (int32_t)s;
            g.state_idx = (int32_t)s;
            g.k_begin = 0;
            g.k_emit = 0;
            g.k_end = block_frames;
            g.carry_in = 1;
            g.carry_out = 1;
            segs[(size_t)s] = g;
        }
        ck(cudaMalloc((void **)&rt->d_segs, sizeof(PvSegment) * segs.size()), "segment table");
        if (rc == PV_OK) ck(cudaMemcpy(rt->d_segs, segs.data(), sizeof(PvSegment) * segs.size(), cudaMemcpyHostToDevice), "segment table");
    }
    ck(cudaHostAlloc((void **)&rt->h_in, sizeof(float) * (size_t)(rt->S * rt->in_w), cudaHostAllocDefault), "pinned input");
    ck(cudaHostAlloc((void **)&rt->h_out, sizeof(float) * (size_t)(rt->S * V * rt->out_w), cudaHostAllocDefault), "pinned output");
    if (rc == PV_OK) {
        memset(rt->h_in, 0, sizeof(float) * (size_t)(rt->S * rt->in_w));
        rc = rt_clear(rt);
    }
    // one eager block: sets the kernel attributes (cudaFuncSetAttribute cannot be recorded)
    const int64_t l0 = h->launches;
    if (rc == PV_OK) rc = rt_enqueue(rt, 0);
    if (rc == PV_OK) ck(cudaStreamSynchronize(rt->st), "first block");
    rt->kernels_per_step = (int)(h->launches - l0);
    for (int p = 0; p < 2 && rc == PV_OK; p++) {
        cudaGraph_t g = nullptr;
        ck(cudaStreamBeginCapture(rt->st, cudaStreamCaptureModeThreadLocal), "begin capture");
        if (rc != PV_OK) break;
        int rc2 = rt_enqueue(rt, p);
        cudaError_t e = cudaStreamEndCapture(rt->st, &g);
        if (rc2 != PV_OK) rc = rc2;
        else ck(e, "end capture");
        if (rc == PV_OK) ck(cudaGraphInstantiate(&rt->exec[p], g, 0), "instantiate");
        if (g) cudaGraphDestroy(g);
    }
    h->launches = l0;            // the recording issued nothing; pv_rt_step counts the replays
    if (rc == PV_OK) rc = rt_clear(rt);
    if (rc != PV_OK) {
        pv_rt_close(rt);
        return rc;
    }
    *out = rt;
    return PV_OK;
}

int pv_rt_reset(pv_rt *rt)
{
    if (!rt) return fail(PV_ERR_PARAM, "pv_rt_reset: null server");
    DeviceGuard guard(rt->h->device);
    return rt_clear(rt);
}

int64_t pv_rt_latency_samples(const pv_rt *rt) { return rt ? rt->hist : 0; }
float *pv_rt_input(pv_rt *rt) { return rt ? rt->h_in : nullptr; }
float *pv_rt_output(pv_rt *rt) { return rt ? rt->h_out : nullptr; }

int pv_rt_step(pv_rt *rt)
{
    if (!rt) return fail(PV_ERR_PARAM, "pv_rt_step: null server");
    DeviceGuard guard(rt->h->device);
    PV_CUDA(cudaGraphLaunch(rt->exec[rt->cur], rt->st));
    PV_CUDA(cudaStreamSynchronize(rt->st));
    rt->cur ^= (rt->hist > 0);
    rt->h->launches += rt->kernels_per_step;
    return PV_OK;
}

int pv_rt_callback(pv_rt *rt, float *outputBuffer, const float *inputBuffer, uint32_t nBufferFrames)
{
    if (!rt || !outputBuffer || !inputBuffer) return fail(PV_ERR_PARAM, "pv_rt_callback: null argument");
    if ((int64_t)nBufferFrames != rt->in_w)
        return fail(PV_ERR_PARAM, "pv_rt_callback: nBufferFrames=%u, the server was opened for blocks of %lld samples",
                    nBufferFrames, (long long)rt->in_w);
    memcpy(rt->h_in, inputBuffer, sizeof(float) * (size_t)(rt->S * rt->in_w));
    int rc = pv_rt_step(rt);
    if (rc != PV_OK) return rc;
    memcpy(outputBuffer, rt->h_out, sizeof(float) * (size_t)(rt->S * rt->h->p.n_voices * rt->out_w));
    return 0;
}

int pv_fft_batch(pv_handle *h, const float *in, float *out, int32_t n, int64_t batch, int32_t direction, void *cuda_stream)
{
    if (!h || !in || !out || batch < 0 || (direction != 1 && direction != -1))
        return fail(PV_ERR_PARAM, "pv_fft_batch: bad argument");
    if (n < 1 || n > 8192 || (n & (n - 1))) return fail(PV_ERR_PARAM, "pv_fft_batch: n=%d is not a power of two in 1..8192", n);
    if (batch == 0) return PV_OK;
    DeviceGuard guard(h->device);
    const int lg = ilog2(n);
    if (!h->d_fft_tw[lg]) {
        std::vector<float2> tw((size_t)n);
        for (int k = 0; k < n; k++) {
            const double a = -2.0 * M_PI * (double)k / (double)n;
            tw[(size_t)k] = make_float2((float)cos(a), (float)sin(a));
        }
        int rc = upload(&h->d_fft_tw[lg], tw);
        if (rc != PV_OK) return rc;
    }
    PV_CUDA(pv_launch_fft_batch(reinterpret_cast<const float2 *>(in), reinterpret_cast<float2 *>(out), lg, batch, direction,
                                h->d_fft_tw[lg], (cudaStream_t)cuda_stream));
    h->launches++;
    return PV_OK;
}

int64_t pv_launch_count(const pv_handle *h) { return h ? h->launches : 0; }

int pv_timing_enable(pv_handle *h, int32_t on)
{
    if (!h) return fail(PV_ERR_PARAM, "null handle");
    h->timing = on != 0;
    return PV_OK;
}

int pv_timing_read(pv_handle *h, double *total_ms, int64_t *launches)
{
    if (!h) return fail(PV_ERR_PARAM, "null handle");
    DeviceGuard guard(h->device);
    double tot = 0;
    int64_t n = 0;
    for (auto &e : h->events) {
        PV_CUDA(cudaEventSynchronize(e.second));
        float ms = 0;
        PV_CUDA(cudaEventElapsedTime(&ms, e.first, e.second));
        tot += ms;
        n++;
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    h->events.clear();
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return PV_OK;
}

}  // extern "C"
