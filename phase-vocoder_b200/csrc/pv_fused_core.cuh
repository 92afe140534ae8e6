// pv_fused_core.cuh -- per-frame body of the fused kernels (compat mode), register blocked.
//
// One thread group of T = N/16 threads walks the frames of a segment.  Per frame:
//
//   fwd pass 1   radix-R1 over n1        inputs straight from global memory (window, zero-phase
//                                        shift and zero pad folded into the load index)
//   exchange 1   shared memory (padded, conflict free)
//   fwd pass 2   radix-R2 over n2
//   exchange 2
//   fwd pass 3   radix-8 over n3, butterflies t3 = u and t3 = B3-u in the SAME thread, so that the
//                conjugate partners C[k], C[N-k] sit in one thread's registers
//   middle       real-FFT split, step D+E of the reference (collapsed: re'=|Re X|, im'=Re X Im X/|X|),
//                Hermitian pack of the N-point C2R -- all in registers, no shared memory
//   inv pass 1   radix-4 over the four packed values per butterfly, in registers
//   exchange 3
//   inv pass 2   radix-R1
//   exchange 4
//   inv pass 3   radix-R2 -> time samples; /N, half swap, window, overlap-add into the ring
//
// Index algebra (M = N complex points forward, M' = N/2 inverse, B3 = N/8 = 2T):
//   forward  n = n1*S1 + n2*8 + n3          k = k1 + R1*k2 + B3*k3        S1 = N/R1 = R2*8
//   inverse  kappa = n1*B3 + n2*R2 + n3     n = m1 + 4*m2 + 4*R1*m3       (B3 = R1*R2)
#pragma once
#include <type_traits>

#include "pv_fft_regs.cuh"

namespace pvfused {
using namespace pvfft;

template <int LOG2N>
struct Shape {
    static constexpr int N = 1 << LOG2N;
    static constexpr int T = N / 16;              // threads per frame group
    static constexpr int B3 = N / 8;              // butterflies of the last forward pass (= 2T)
    static constexpr int R1 = (LOG2N >= 10) ? 16 : 8;
    static constexpr int R2 = B3 / R1;            // 2048:16 1024:8 512:8 256:4
    static constexpr int S1 = N / R1;             // = R2*8
    static constexpr int PAD1 = 16 / R1;          // exchange-1 row padding (float2 units)
    static constexpr int LD1 = S1 + PAD1;
    // exchange 2 is done IN PLACE in the exchange-1 buffer: a pass-2 butterfly (k1, n3) reads the 16
    // addresses k1*LD1 + n2*8 + n3 and writes its outputs k2 back to the same addresses (k2 in the place
    // of n2); no other thread touches them, so no barrier and no second buffer are needed
    static constexpr int EX1 = R1 * LD1;          // float2 elements
    // inverse: M' = 4*B3, B3 = R1*R2
    static constexpr int ILD1 = B3 + (R2 < 16 ? R2 : 0);   // exchange-3 rows (m1): m1*ILD1 + t1
    static constexpr int IEX1 = 4 * ILD1;
    static constexpr int ILD2 = R2 + 1;           // exchange-4: (m1 + 4*m2)*(R2+1) + n3
    static constexpr int IEX2 = 4 * R1 * ILD2;
    // compat kernel: A = exchanges 1, 2 (in place) and 4, B = exchange 3
    static constexpr int BUF_A = (EX1 > IEX2 ? EX1 : IEX2);
    static constexpr int BUF_B = IEX1;
    static_assert(R1 * R2 == B3 && R2 >= 4, "shape");
};

// Twiddle tables in global memory (built on the host in double precision).
struct Tables {
    const float2 *tw1;    // [(R1-1)][S1]   W_N^{k1*t1}
    const float2 *tw2;    // [(R2-1)][8]    W_S1^{k2*n3}
    const float2 *tw2n;   // [N]            exp(-j*pi*k/N) (2N-th roots) for the split / pack steps
    const float2 *itw1;   // [3][B3]        exp(+2 pi i m1 t1 / (N/2))
    const float2 *itw2;   // [(R1-1)][R2]   exp(+2 pi i m2 n3 / B3)
    const float *win;     // [N]   analysis window
};

// Loop-invariant per-thread twiddles kept in registers (used when every thread owns exactly one
// pass-1 butterfly, S1 == T, so that t1 == u == tid for the whole segment).  All other twiddles of
// the thread are products of these with compile-time constants, which removes ~37 L1 loads and
// their address arithmetic per thread and frame.
// Thread 0 owns the two self-paired columns t3 = 0 and t3 = B3/2.  Its "p side" behaves like a regular
// thread with u = 0 and its "q side" like a regular thread with u = B3/2 (B3 - B3/2 = B3/2), so giving every
// thread separate p-side and q-side bases (equal for tid != 0) lets ONE code path serve all threads: warp 0
// no longer executes a divergent copy of the middle section on the critical path of every frame.
struct ThreadTw {
    float2 w1, w2, w4, w8;   // W_N^{u}, ^2u, ^4u, ^8u  for u = tid  (pass-1 twiddles, p side)
    float2 hp, hq;           // exp(-j*pi*u/N) for u = tid and for u_q = tid ? tid : B3/2
    float2 q1, q2, q4;       // W_N^{u_q}, ^2u_q, ^4u_q
};

template <int LOG2N>
PV_DEV ThreadTw load_thread_tw(int u, const Tables &tb)
{
    using S = Shape<LOG2N>;
    ThreadTw t;
    t.w1 = PV_LDG(tb.tw1 + 0 * S::S1 + u);
    t.w2 = PV_LDG(tb.tw1 + 1 * S::S1 + u);
    t.w4 = PV_LDG(tb.tw1 + 3 * S::S1 + u);
    t.w8 = (S::R1 > 8) ? PV_LDG(tb.tw1 + 7 * S::S1 + u) : make_float2(1.f, 0.f);
    const int uq = u ? u : S::B3 / 2;
    t.hp = PV_LDG(tb.tw2n + u);
    t.hq = PV_LDG(tb.tw2n + uq);
    t.q1 = PV_LDG(tb.tw2n + 2 * uq);
    t.q2 = PV_LDG(tb.tw2n + 4 * uq);
    t.q4 = cmul(t.q2, t.q2);
    return t;
}

struct FrameIO {
    const float *in;      // stream base
    long long n_in;       // valid samples in the stream
    long long base;       // first sample of this frame (k*Ha)
    bool analysed;        // false: zero spectrum (never analysed frame)
    bool vec_ok;          // 8-byte aligned float2 loads are legal for this stream
};

// Loads x[base+i], x[base+i+1] (zero beyond n_in) times the window.
PV_DEV float2 load_pair(const FrameIO &io, const float *win, int i)
{
    const long long g = io.base + i;
    float x0, x1;
    if (io.vec_ok && g + 1 < io.n_in) {
        const float2 v = PV_LDG(reinterpret_cast<const float2 *>(io.in + g));
        x0 = v.x; x1 = v.y;
    } else {
        x0 = (g < io.n_in) ? PV_LDG(io.in + g) : 0.f;
        x1 = (g + 1 < io.n_in) ? PV_LDG(io.in + g + 1) : 0.f;
    }
    const float2 w = PV_LDG(reinterpret_cast<const float2 *>(win + i));
    return f2mul(make_float2(x0, x1), w);
}

// ---- private input ring ----------------------------------------------------------------------
// When Ha is a multiple of 2*S1 every sample is consumed by the SAME thread in all the frames that
// overlap it (frame coordinate i = 2*t1 mod 2*S1 is invariant under a shift by Ha).  Each thread
// then owns a disjoint set of slots of a shared-memory ring of N floats: it copies its own new
// samples of frame k+1 with cp.async while frame k is being processed and reads them back with no
// barrier in between.  ring slot of absolute sample s: s mod N.
#if defined(PV_HOST_EMUL)
PV_DEV void cp_async8(float *dst, const float *src, int src_bytes)
{
    dst[0] = src_bytes >= 4 ? src[0] : 0.f;
    dst[1] = src_bytes >= 8 ? src[1] : 0.f;
}
PV_DEV void cp_async_wait_all() {}
#else
PV_DEV void cp_async8(float *dst, const float *src, int src_bytes)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
PV_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
#endif

#if defined(PV_HOST_EMUL)
PV_DEV void cp_async16(float *dst, const float *src, int src_bytes)
{
    for (int j = 0; j < 4; j++) dst[j] = src_bytes >= 4 * (j + 1) ? src[j] : 0.f;
}
#else
// 16-byte copy through L2 only (.cg): the input stream is consumed once per SM and must not evict the
// window / twiddle tables from the ~24 KB of L1 that remain next to 204 KB of shared memory
PV_DEV void cp_async16(float *dst, const float *src, int src_bytes)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
#endif

// ---- bulk asynchronous copy of one input hop (the TMA engine's 1-D form: cp.async.bulk + mbarrier) ----
// ONE elected thread moves the whole new hop of the next frame into the ring; completion is counted in bytes on an
// mbarrier in shared memory that every consumer waits on (phase parity = frame parity).  Replaces T x 16-byte
// cp.async (LDGSTS) copies and their per-thread address / bounds arithmetic.
#if !defined(PV_HOST_EMUL)
PV_DEV void mbar_init(unsigned long long *mbar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
PV_DEV void bulk_load_hop(float *dst, const float *src, unsigned bytes, unsigned long long *mbar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst), m = (unsigned)__cvta_generic_to_shared(mbar);
    // the slots about to be overwritten were last read through the generic proxy: order those reads before the async write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(m), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes),
                 "r"(m) : "memory");
}
PV_DEV void mbar_wait(unsigned long long *mbar, unsigned parity)
{
    const unsigned m = (unsigned)__cvta_generic_to_shared(mbar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(m), "r"(parity) : "memory");
}
#else
PV_DEV void mbar_init(unsigned long long *, unsigned) {}
PV_DEV void bulk_load_hop(float *dst, const float *src, unsigned bytes, unsigned long long *) { for (unsigned j = 0; j < bytes / 4; j++) dst[j] = src[j]; }
PV_DEV void mbar_wait(unsigned long long *, unsigned) {}
#endif

// Cooperative refill with 16-byte pieces: samples [lo, N) (frame coordinates, multiples of 4) of the frame at
// io.base, spread over T threads.  Issue after a barrier that follows the last read of the replaced samples;
// complete with cp_async_wait_all() + a barrier before the first read of the new ones.
template <int N, int T>
PV_DEV void ring_prefetch_coop16(int tid, const FrameIO &io, float *ring, int lo)
{
    for (int i = lo + 4 * tid; i < N; i += 4 * T) {
        const long long g = io.base + i;
        const long long left = (io.n_in - g) * 4;
        const int bytes = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
        cp_async16(ring + ((int)(g & (N - 1))), io.in + (bytes > 0 ? g : 0), bytes);
    }
}

// Copies samples [i, i+2) (frame coordinates) of the frame starting at io.base into the ring.
template <int N>
PV_DEV void ring_fetch(const FrameIO &io, float *ring, int i)
{
    const long long g = io.base + i;
    long long left = (io.n_in - g) * 4;
    const int bytes = left >= 8 ? 8 : (left > 0 ? (int)left : 0);
    const float *src = io.in + (bytes > 0 ? g : 0);           // keep the address valid when nothing is read
    cp_async8(ring + ((int)(g & (N - 1))), src, bytes);
}

// Issues the copies of frame coordinates [lo, N) that belong to thread `tid` (lo = 0: whole frame).
template <int LOG2N>
PV_DEV void ring_prefetch(int tid, const FrameIO &io, float *ring, int lo)
{
    using S = Shape<LOG2N>;
    constexpr int N = S::N, T = S::T, S1 = S::S1;
#pragma unroll
    for (int t1 = tid; t1 < S1; t1 += T)
        for (int i = lo + 2 * t1; i < N; i += 2 * S1) ring_fetch<N>(io, ring, i);
}

// ---- forward passes 1 and 2 (results left in bufB for pass 3) ----
// ring == nullptr: inputs come straight from global memory.
template <int LOG2N, bool TWREG, class Sync, class Hook>
PV_DEV void forward_12(int tid, const FrameIO &io, const Tables &tb, const ThreadTw &tt, const float *ring,
                       float2 *bufA, float2 *bufB, Sync sync, Hook after_exchange1)
{
    using S = Shape<LOG2N>;
    constexpr int N = S::N, T = S::T, R1 = S::R1, R2 = S::R2, S1 = S::S1;
    const int rbase = (int)(io.base & (N - 1));
    // c[n], n = n1*S1 + t1: n < N/4 -> f[N/2 + 2n]; n >= 3N/4 -> f[2(n - 3N/4)]; else 0
    auto src_index = [&](int n1, int t1) -> int {
        return n1 < R1 / 4 ? N / 2 + 2 * (n1 * S1 + t1) : 2 * ((n1 - 3 * R1 / 4) * S1 + t1);
    };
    // pass 1: butterflies t1 in [0, S1)
#pragma unroll
    for (int t1 = tid; t1 < S1; t1 += T) {
        float2 v[R1];
        // The ring / global choice is made once per butterfly, not per sample pair: the hot (ring) path stays one
        // straight run of instructions instead of short runs separated by the bounds-checked global fallback,
        // which the instruction cache pays for.
        if (ring != nullptr) {
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) {
                if (n1 >= R1 / 4 && n1 < 3 * R1 / 4) { v[n1] = make_float2(0.f, 0.f); continue; }
                const int i = src_index(n1, t1);
                const float2 x = *reinterpret_cast<const float2 *>(ring + ((rbase + i) & (N - 1)));
                const float2 w = PV_LDG(reinterpret_cast<const float2 *>(tb.win + i));
                v[n1] = f2mul(x, w);
            }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) {
                if (n1 >= R1 / 4 && n1 < 3 * R1 / 4) { v[n1] = make_float2(0.f, 0.f); continue; }
                v[n1] = load_pair(io, tb.win, src_index(n1, t1));
            }
        }
        dft_pruned_fwd<R1>(v);
        bufA[t1] = v[0];
        if constexpr (TWREG && R1 == 16) {
            static_assert(!TWREG || S1 == T, "register twiddles need one pass-1 butterfly per thread");
            float2 w[16];
            w[1] = tt.w1; w[2] = tt.w2; w[4] = tt.w4; w[8] = tt.w8;
            w[3] = cmul(w[1], w[2]); w[5] = cmul(w[1], w[4]); w[6] = cmul(w[2], w[4]); w[7] = cmul(w[3], w[4]);
#pragma unroll
            for (int k1 = 9; k1 < 16; k1++) w[k1] = cmul(w[k1 - 8], w[8]);
#pragma unroll
            for (int k1 = 1; k1 < R1; k1++) bufA[k1 * S::LD1 + t1] = cmul(v[k1], w[k1]);
        } else {
#pragma unroll
            for (int k1 = 1; k1 < R1; k1++)
                bufA[k1 * S::LD1 + t1] = cmul(v[k1], PV_LDG(tb.tw1 + (k1 - 1) * S1 + t1));
        }
    }
    sync();
    after_exchange1();
    // pass 2: butterflies (k1, n3), k1 fastest across threads
    if constexpr (R2 == 32) {
        // window 4096: 128 radix-32 butterflies for 256 threads -> two threads per butterfly (even / odd outputs, the
        // half is warp-uniform).  Still in place: both threads read all 32 inputs, folded to 16 values while loading,
        // and a barrier separates the loads from the stores
        static_assert(2 * R1 * 8 == T, "pass 2 cover");
        const int b = tid % (R1 * 8), half = tid / (R1 * 8);
        const int k1 = b % R1, n3 = b / R1;
        float2 o[16];
        if (half == 0) {
#pragma unroll
            for (int n2 = 0; n2 < 16; n2++) o[n2] = dft32_fold<false>(bufA[k1 * S::LD1 + n2 * 8 + n3], bufA[k1 * S::LD1 + (n2 + 16) * 8 + n3]);
        } else {
#pragma unroll
            for (int n2 = 0; n2 < 16; n2++) o[n2] = dft32_fold<true>(bufA[k1 * S::LD1 + n2 * 8 + n3], bufA[k1 * S::LD1 + (n2 + 16) * 8 + n3]);
        }
        sync();
        if (half == 0) dft32_half_finish<-1, false>(o);
        else dft32_half_finish<-1, true>(o);
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const int k2 = 2 * q + half;
            float2 r = o[q];
            if (k2 != 0) r = cmul(r, PV_LDG(tb.tw2 + (k2 - 1) * 8 + n3));
            bufA[k1 * S::LD1 + k2 * 8 + n3] = r;
        }
    } else {
#pragma unroll
        for (int b = tid; b < R1 * 8; b += T) {
            const int k1 = b % R1, n3 = b / R1;
            float2 v[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) v[n2] = bufA[k1 * S::LD1 + n2 * 8 + n3];
            dft<R2, -1>(v);
            bufA[k1 * S::LD1 + n3] = v[0];
#pragma unroll
            for (int k2 = 1; k2 < R2; k2++)
                bufA[k1 * S::LD1 + k2 * 8 + n3] = cmul(v[k2], PV_LDG(tb.tw2 + (k2 - 1) * 8 + n3));
        }
    }
    sync();
}

// ---- forward pass 3 for the butterfly pair of thread u: P[j] = C[tP + B3*j], Q[j] = C[tQ + B3*j] ----
template <int LOG2N>
PV_DEV void forward_3(int u, const float2 *bufA, float2 (&P)[8], float2 (&Q)[8])
{
    using S = Shape<LOG2N>;
    const int tP = u, tQ = (u == 0) ? S::B3 / 2 : S::B3 - u;        // t3 = k1 + R1*k2
    const int oP = (tP % S::R1) * S::LD1 + (tP / S::R1) * 8, oQ = (tQ % S::R1) * S::LD1 + (tQ / S::R1) * 8;
#pragma unroll
    for (int n3 = 0; n3 < 8; n3++) {
        P[n3] = bufA[oP + n3];
        Q[n3] = bufA[oQ + n3];
    }
    dft<8, -1>(P);
    dft<8, -1>(Q);
}

// 2*X[k] of the 2N-point real spectrum from a = C[k], b = C[N-k], w = exp(-j*pi*k/N).  The factor 2 is
// carried through the (degree-1 homogeneous) compat map and the linear Hermitian pack and removed by the
// pre-scaled synthesis window w/(2N): exact, power of two.
PV_DEV float2 split(float2 a, float2 b, float2 w)
{
    const float2 e = f2add(a, cconj(b));
    const float2 o = f2add(mul_mj(a), f2swap(b));                 // (a - conj b)/j = (a.y + b.y, b.x - a.x)
    return cadd(e, cmul(w, o));
}

// steps D+E of the reference collapsed algebraically: re' = |Re X|, im' = Re X Im X/|X|
PV_DEV float2 compat_map(float2 X, bool nan_compat)
{
    // Branch-free on purpose: a thread maps nine independent bins, and only straight-line code lets the scheduler
    // interleave their rsqrt (MUFU) latencies.
    const float m2 = X.x * X.x + X.y * X.y;
#ifdef PV_HOST_EMUL
    const float r = m2 > 0.f ? 1.0f / sqrtf(m2) : 0.f;
#else
    float r;     // single MUFU.RSQ (2 ulp); a zero or denormal |X|^2 flushes to 0 -> r = inf, replaced below
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(m2));
#endif
    const bool tiny = m2 < 1.17549435e-38f;                         // zero or denormal |X|^2: Im' = 0
    const bool zero_nan = nan_compat && m2 == 0.f;                  // atanf(0/0) of kernel.cu:108 when asked for
    const float z = __builtin_nanf("");
    const float re = zero_nan ? z : fabsf(X.x);
    const float im = zero_nan ? z : (tiny ? 0.f : X.x * X.y * r);
    return make_float2(re, im);
}

// Z[k] = (Yk + conj(Ym)) + j*w*(Yk - conj(Ym)), w = exp(+2 pi i k/N), Ym = Y[N/2 - k]
PV_DEV float2 herm_pack(float2 yk, float2 ym, float2 w)
{
    const float2 s = f2add(yk, cconj(ym));
    const float2 d = f2add(yk, make_float2(-ym.x, ym.y));
    const float2 t = cmul(w, d);
    return f2add(s, mul_pj(t));                                   // (s.x - t.y, s.y + t.x)
}

// ---- middle: P,Q (16 transform outputs) -> Zp[4], Zq[4] = packed inverse inputs at
//      kappa = tP + B3*n1 and tQ + B3*n1 ----
// exp(-j*pi*(base + B3*J)/N) from exp(-j*pi*base/N): B3/N = 1/8 -> a multiple of the 16th root
template <int J>
PV_DEV float2 rot16(float2 a) { return twid16<J, -1>(a); }

template <int LOG2N, bool TWREG>
PV_DEV void middle_compat(int u, const Tables &tb, const ThreadTw &tt, bool nan_compat, const float2 (&P)[8],
                          const float2 (&Q)[8], float2 (&Zp)[4], float2 (&Zq)[4])
{
    using S = Shape<LOG2N>;
    constexpr int N = S::N, B3 = S::B3;
    float2 Yp[5], Yq[4];
    {
        // one path for all threads (see ThreadTw): selects instead of a divergent copy for thread 0
        const bool u0 = (u == 0);
        auto sel = [&](float2 a, float2 b) { return u0 ? a : b; };
        // partners: regular thread (P[j], Q[7-j]) and (Q[j], P[7-j]); thread 0 (P[j], P[(8-j)&7]) and (Q[j], Q[7-j])
        const float2 b1_7 = sel(P[0], Q[7]), b1_6 = sel(P[7], Q[6]), b1_5 = sel(P[6], Q[5]), b1_4 = sel(P[5], Q[4]);
        const float2 b2_7 = sel(Q[7], P[7]), b2_6 = sel(Q[6], P[6]), b2_5 = sel(Q[5], P[5]), b2_4 = sel(Q[4], P[4]);
        const float2 hc = cconj(tt.hq), q1c = cconj(tt.q1);
        Yp[0] = compat_map(split(P[0], b1_7, tt.hp), nan_compat);
        Yp[1] = compat_map(split(P[1], b1_6, rot16<1>(tt.hp)), nan_compat);
        Yp[2] = compat_map(split(P[2], b1_5, rot16<2>(tt.hp)), nan_compat);
        Yp[3] = compat_map(split(P[3], b1_4, rot16<3>(tt.hp)), nan_compat);
        Yq[0] = compat_map(split(Q[0], b2_7, rot16<1>(hc)), nan_compat);
        Yq[1] = compat_map(split(Q[1], b2_6, rot16<2>(hc)), nan_compat);
        Yq[2] = compat_map(split(Q[2], b2_5, rot16<3>(hc)), nan_compat);
        Yq[3] = compat_map(split(Q[3], b2_4, rot16<4>(hc)), nan_compat);
        Yp[4] = make_float2(0.f, 0.f);
        if (u0) {               // bin N/2 and the real-only DC / Nyquist of the C2R (cuFFT, kernel.cu:366)
            Yp[4] = compat_map(split(P[4], P[4], make_float2(0.f, -1.f)), nan_compat);
            Yp[0].y = 0.f;
            Yp[4].y = 0.f;
        }
        // exp(+2 pi i k/N) = conj(W_N^k): W_N^(u + B3 j) = w1 * W16^(2j); W_N^(B3 - u_q + B3 j) = conj(q1) * W16^(2j+2)
        Zp[0] = herm_pack(Yp[0], sel(Yp[4], Yq[3]), cconj(tt.w1));
        Zp[1] = herm_pack(Yp[1], sel(Yp[3], Yq[2]), cconj(rot16<2>(tt.w1)));
        Zp[2] = herm_pack(Yp[2], sel(Yp[2], Yq[1]), cconj(rot16<4>(tt.w1)));
        Zp[3] = herm_pack(Yp[3], sel(Yp[1], Yq[0]), cconj(rot16<6>(tt.w1)));
        Zq[0] = herm_pack(Yq[0], sel(Yq[3], Yp[3]), cconj(rot16<2>(q1c)));
        Zq[1] = herm_pack(Yq[1], sel(Yq[2], Yp[2]), cconj(rot16<4>(q1c)));
        Zq[2] = herm_pack(Yq[2], sel(Yq[1], Yp[1]), cconj(rot16<6>(q1c)));
        Zq[3] = herm_pack(Yq[3], sel(Yq[0], Yp[0]), cconj(rot16<8>(q1c)));
    }
    (void)tb;
    (void)B3;
    (void)N;
}

// ---- inverse pass 1 (radix 4 in registers) for one butterfly column t1, writes exchange 3 ----
template <int LOG2N>
PV_DEV void inverse_1(int t1, const Tables &tb, float2 (&Z)[4], float2 *bufA)
{
    using S = Shape<LOG2N>;
    dft<4, +1>(Z);
    bufA[t1] = Z[0];
#pragma unroll
    for (int m1 = 1; m1 < 4; m1++)
        bufA[m1 * S::ILD1 + t1] = cmul(Z[m1], PV_LDG(tb.itw1 + (m1 - 1) * S::B3 + t1));
}

// same with the three twiddles exp(+2 pi i m1 t1/(N/2)), m1 = 1..3, supplied by the caller
template <int LOG2N>
PV_DEV void inverse_1_tw(int t1, float2 a1, float2 a2, float2 a3, float2 (&Z)[4], float2 *bufA)
{
    using S = Shape<LOG2N>;
    dft<4, +1>(Z);
    bufA[t1] = Z[0];
    bufA[1 * S::ILD1 + t1] = cmul(Z[1], a1);
    bufA[2 * S::ILD1 + t1] = cmul(Z[2], a2);
    bufA[3 * S::ILD1 + t1] = cmul(Z[3], a3);
}

// ---- inverse passes 2 and 3 + steps G/H (scale, half swap, window, overlap-add) ----
// acc: OLA ring of N floats, pos0: ring position of sample 0 of this frame.
// `scale` multiplies the unnormalised inverse (compat: 1/N; corrected: gain/N); `pre_last_sync` runs
// just before the last barrier of the frame (used to complete asynchronous ring copies).
struct TableTw2 {      // default source of the inverse pass-2 twiddles: the global table
    const float2 *itw2;
    int R2;
    PV_DEV float2 operator()(int /*q*/, int m2, int n3) const { return PV_LDG(itw2 + (m2 - 1) * R2 + n3); }
};

// UNIT_SCALE: the caller has already scaled the spectrum (the corrected kernel folds gain / N into the 1/2 of its real-FFT
// split, where it costs nothing); `scale` is ignored then.
template <int LOG2N, class Sync, class PreLast, class Tw2 = TableTw2, bool UNIT_SCALE = false>
PV_DEV void inverse_23_ola(int tid, const Tables &tb, float2 *bufA, float2 *bufB, float *acc, int pos0, int Hs,
                           bool zero_frame, float scale, Sync sync, PreLast pre_last_sync,
                           Tw2 tw2 = TableTw2{nullptr, 0})
{
    using S = Shape<LOG2N>;
    constexpr int N = S::N, T = S::T, R1 = S::R1, R2 = S::R2, B3 = S::B3;
    constexpr int C2 = 4 * R2;       // pass-2 butterflies (m1, n3), radix R1
    constexpr int C3 = 4 * R1;       // pass-3 butterflies (m1, m2), radix R2
    static_assert(C2 >= T || 2 * C2 == T, "pass 2 cover");
    constexpr bool QUAD3 = (4 * C3 == T) && (R2 == 32);    // window 4096: four threads per radix-32 butterfly
    static_assert(C3 >= T || 2 * C3 == T || QUAD3, "pass 3 cover");
    constexpr bool SPLIT2 = (2 * C2 == T) && (R1 == 16);   // two threads per radix-16 butterfly
    constexpr bool SPLIT3 = (2 * C3 == T) && (R2 == 16);
    // steps G+H for complex output n (samples 2n, 2n+1).  The last Hs slots of the frame are "fresh": the
    // caller zeroes every hop right after emitting it, so the overlap-add is a plain accumulate.
    auto ola = [&](int n, float2 v) {
        const int i = (2 * n + N / 2) & (N - 1);       // half swap (kernel.cu:51-59)
        // cudaDivVec kernel.cu:130-138 (x/N == x*(1/N) exactly, N a power of two) and cudaWindow :75-81.
        // One window table serves analysis and synthesis: with 4 CTAs x 51 KB of shared memory only ~24 KB
        // of L1 remain per SM, and a second (pre-scaled) 8 KB table measurably thrashes it.
        float2 w = PV_LDG(reinterpret_cast<const float2 *>(tb.win + i));
        if constexpr (!UNIT_SCALE) w = f2mul(w, f2bc(scale));
        float2 *slot = reinterpret_cast<float2 *>(acc + ((pos0 + i) & (N - 1)));
        *slot = f2fma(v, w, *slot);                    // cudaOverlapAdd kernel.cu:111-119
    };
    (void)Hs;
    if (!zero_frame) {
        sync();
        if constexpr (SPLIT2) {
            const int b = tid % C2;
            const int n3 = b % R2, m1 = b / R2;
            float2 v[16], o[8];
#pragma unroll
            for (int n2 = 0; n2 < 16; n2++) v[n2] = bufA[m1 * S::ILD1 + n2 * R2 + n3];
            const int half = tid / C2;                  // warp-uniform
            if (half == 0) dft16_half<+1, false>(v, o);
            else dft16_half<+1, true>(v, o);
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int m2 = 2 * q + half;
                float2 r = o[q];
                if (m2 != 0) {
                    if constexpr (std::is_same<Tw2, TableTw2>::value) r = cmul(r, PV_LDG(tb.itw2 + (m2 - 1) * R2 + n3));
                    else r = cmul(r, tw2(q, m2, n3));
                }
                bufB[(m1 + 4 * m2) * S::ILD2 + n3] = r;
            }
        } else {
#pragma unroll
            for (int b = tid; b < C2; b += T) {
                const int n3 = b % R2, m1 = b / R2;         // n3 fastest: conflict-free on both sides
                float2 v[R1];
#pragma unroll
                for (int n2 = 0; n2 < R1; n2++) v[n2] = bufA[m1 * S::ILD1 + n2 * R2 + n3];
                dft<R1, +1>(v);
                bufB[m1 * S::ILD2 + n3] = v[0];
#pragma unroll
                for (int m2 = 1; m2 < R1; m2++)
                    bufB[(m1 + 4 * m2) * S::ILD2 + n3] = cmul(v[m2], PV_LDG(tb.itw2 + (m2 - 1) * R2 + n3));
            }
        }
        pre_last_sync();
        sync();
    }
    // pass 3 -> time samples
    if constexpr (QUAD3) {
        const int b = tid % C3, quarter = tid / C3;         // warp-uniform quarter: outputs m3 = quarter + 4 j
        float2 o[8];
        if (!zero_frame) {
            float2 v[32];
#pragma unroll
            for (int n3 = 0; n3 < 32; n3++) v[n3] = bufB[b * S::ILD2 + n3];
            if (quarter == 0) dft32_quarter<+1, 0>(v, o);
            else if (quarter == 1) dft32_quarter<+1, 1>(v, o);
            else if (quarter == 2) dft32_quarter<+1, 2>(v, o);
            else dft32_quarter<+1, 3>(v, o);
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) o[j] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) ola(b + C3 * (quarter + 4 * j), o[j]);
    } else if constexpr (SPLIT3) {
        const int b = tid % C3, half = tid / C3;
        float2 v[16], o[8];
        if (!zero_frame) {
#pragma unroll
            for (int n3 = 0; n3 < 16; n3++) v[n3] = bufB[b * S::ILD2 + n3];
            if (half == 0) dft16_half<+1, false>(v, o);
            else dft16_half<+1, true>(v, o);
        } else {
#pragma unroll
            for (int q = 0; q < 8; q++) o[q] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) ola(b + C3 * (2 * q + half), o[q]);
    } else if constexpr (R2 <= 16) {
#pragma unroll
        for (int b = tid; b < C3; b += T) {
            float2 v[R2];
            if (!zero_frame) {
#pragma unroll
                for (int n3 = 0; n3 < R2; n3++) v[n3] = bufB[b * S::ILD2 + n3];
                dft<R2, +1>(v);
            } else {
#pragma unroll
                for (int n3 = 0; n3 < R2; n3++) v[n3] = make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int m3 = 0; m3 < R2; m3++) ola(b + C3 * m3, v[m3]);
        }
    }
    (void)B3;
}

// ---- one whole frame ----
// `hook` runs once per frame at a point where (a) every thread has finished the overlap-add of the
// PREVIOUS frame and (b) the next write to the accumulator is at least one barrier away: the caller
// uses it to emit the previous frame's output hop without dedicated barriers.
template <int LOG2N, bool TWREG, class Sync, class Hook, class PreLast>
PV_DEV void frame_compat(int tid, const FrameIO &io, const Tables &tb, const ThreadTw &tt, bool nan_compat,
                         const float *ring, float2 *bufA, float2 *bufB, float *acc, int pos0, int Hs, Sync sync,
                         Hook hook, PreLast pre_last_sync)
{
    using S = Shape<LOG2N>;
    if (io.analysed) {
        forward_12<LOG2N, TWREG>(tid, io, tb, tt, ring, bufA, bufB, sync, hook);
        float2 P[8], Q[8], Zp[4], Zq[4];
        forward_3<LOG2N>(tid, bufA, P, Q);
        middle_compat<LOG2N, TWREG>(tid, tb, tt, nan_compat, P, Q, Zp, Zq);
        // Exchange 3 goes to B (other threads may still be reading A in pass 3); exchange 4 goes back to A.
        // The caller puts one barrier at the end of the frame so that the next frame's pass 1 cannot
        // overwrite A while the last inverse pass still reads it.
        {
            // exp(+2 pi i m1 u/(N/2)) = conj(W_N^{2 m1 u}); for t1 = B3-u_q: j^m1 * W_N^{2 m1 u_q}
            const float2 w6 = cmul(tt.w2, tt.w4), q6 = cmul(tt.q2, tt.q4);
            inverse_1_tw<LOG2N>(tid, cconj(tt.w2), cconj(tt.w4), cconj(w6), Zp, bufB);
            inverse_1_tw<LOG2N>(tid == 0 ? S::B3 / 2 : S::B3 - tid, mul_pj(tt.q2), make_float2(-tt.q4.x, -tt.q4.y),
                                mul_mj(q6), Zq, bufB);
        }
    } else {
        sync();
        hook();
        sync();
    }
    // 1/(2N): the split step above works with 2*X (see split())
    inverse_23_ola<LOG2N>(tid, tb, bufB, bufA, acc, pos0, Hs, !io.analysed, 0.5f / (float)S::N, sync, pre_last_sync);
    sync();
}

}  // namespace pvfused
