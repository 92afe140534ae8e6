// pv_fused_corrected.cuh -- per-frame body of the fused CORRECTED-mode kernel.
//
// The reference never implemented this stage (PITCH_SHIFT is an empty case, src/phaseVocoder.h:
// 107-111, src/main.cpp:301-303); the arithmetic is specified in DESIGN.md "corrected mode" and
// restated by oracle/pv_oracle_impl.inc (corrected_process).  Per frame and stream:
//
//   forward   N-point real FFT of the windowed, zero-phase frame as an N/2-point complex FFT:
//             radix-R1, exchange, radix-R2, exchange, radix-4 with the butterfly pair (u, B3-u) in
//             one thread, so the real-FFT split is register-local
//   analysis  mag = |X|, P = atan2 in turns*2^32 (uint32), D = (int32)(P - P_prev - nomA): the
//             phase-difference unwrap is integer wrap-around; P_prev stays in registers
//   exchange  mag[] and D[] through shared memory (the pitch map gathers across bins)
//   per voice synthesis bin s gathers analysis bins a_lo..a_hi (pitch ratio beta), accumulates the
//             phase psi[s] += nomS[s] + D[a_hi]*Rq in 64-bit fixed point (turns*2^64: an associative
//             scan across frames, wrapped mod 2 pi for free), Y = m * exp(2 pi j psi); Hermitian
//             pack, inverse N-point real FFT (same passes as the compat kernel), gain/N, half
//             swap, window, overlap-add
#pragma once
#include "pv_fused_core.cuh"

#ifdef PV_HOST_EMUL
#include <cstring>
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
#endif

namespace pvfused {

template <int LOG2N>
struct CShape {
    using S = Shape<LOG2N>;
    static constexpr int N = S::N, T = S::T, B3 = S::B3, R1 = S::R1, R2 = S::R2;
    static constexpr int M = N / 2;                 // complex points of the forward transform
    static constexpr int S1 = M / R1;               // = 4*R2
    static constexpr int LD1 = S1 + 16 / R1;
    static constexpr int LD2 = 5;                   // exchange 2: t3*5 + n3, n3 < 4
    static constexpr int EX1 = R1 * LD1, EX2 = B3 * LD2;
    static constexpr int BUF_A = EX1 > S::IEX1 ? EX1 : S::IEX1;
    static constexpr int BUF_B = EX2 > S::IEX2 ? EX2 : S::IEX2;
    static constexpr int NB = N / 2 + 1;
    static constexpr int C1 = S1;                   // pass-1 butterflies (radix R1)
    static constexpr int C2 = R1 * 4;               // pass-2 butterflies (radix R2)
    static_assert(C1 >= T || (2 * C1 == T && R1 == 16), "pass 1 cover");
    static_assert(C2 >= T || (2 * C2 == T && R2 == 16) || (4 * C2 == T && R2 == 32), "pass 2 cover");
};

struct CTables {
    const float2 *ctw1;     // [(R1-1)][S1]  W_{N/2}^{k1 t1}
    const float2 *ctw2;     // [(R2-1)][4]   W_{S1}^{k2 n3}
    const float2 *tw2n;     // [N]           exp(-j pi k / N)
    const float2 *itw1;     // inverse tables, shared with the compat kernel
    const float2 *itw2;
    const float *win;
    const uint32_t *nomA;   // [NB]
    const int32_t *a_lo;    // [V][NB]
    const int32_t *a_hi;    // [V][NB]
    const unsigned long long *nomS;   // [V][NB]
    const uint32_t *gather; // [V][T][9] per-thread source bins in slot order, pv_fused_tables.h (4.6 KB per voice at N = 2048:
                            // small on purpose -- only ~24 KB of L1 remain next to the shared-memory carve-out)
    unsigned long long Rq[8];
    unsigned long long beta_q[8];   // pitch ratio, Q32.32
    unsigned long long bqs[8];      // (beta_q * Hs) << (32 - lgN): nomS[s] = a_hi * bqs mod 2^64 is recomputed per slot
    int multi[8];                   // voice has synthesis bins that sum several analysis bins
    float scale;            // gain / N
    int V;                  // voices of THIS launch (tables above start at its first voice)
    int V_total, voice0;    // voices of the handle and first voice of this launch: the carried state holds all V_total
    int last_group;         // this launch holds the handle's last voice: it writes the state's common part (previous phase);
                            // earlier voice groups must leave it as carried in for the launches that follow
    int Ha;
};

// Cooperative ring refill: pairs [lo, N) of the frame at io.base, spread over the T threads.
// Must be issued after a barrier that follows the last read of the replaced samples, and be
// completed (cp_async_wait_all + barrier) before the first read of the new ones.
template <int N, int T>
PV_DEV void ring_prefetch_coop(int tid, const FrameIO &io, float *ring, int lo)
{
    for (int i = lo + 2 * tid; i < N; i += 2 * T) ring_fetch<N>(io, ring, i);
}

// atan2(im, re) in turns, scaled by 2^32 (wraps mod 2^32).  Octant reduction + a degree-15 odd minimax
// polynomial for atan(t)/(2 pi) on [0, 1] (max error 2.6e-8 turns in fp32, i.e. the accuracy of atan2f):
// branch-free, no special-case handling (atan2(0,0) = 0), ~1/3 of the instructions of atan2f.
// The three steps are separate functions so that TWO bins can share the polynomial as packed pairs (phase_turns32x2): each
// lane of a packed instruction rounds like the scalar one, so a bin's phase does not depend on whether it was paired.
PV_DEV float phase_ratio(float re, float im)            // min(|re|, |im|) / max(|re|, |im|) in [0, 1]
{
    const float ax = fabsf(re), ay = fabsf(im);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
#ifdef PV_HOST_EMUL
    return mx > 1e-30f ? mn / mx : 0.f;
#else
    // one MUFU.RCP + one multiply: __fdividef wraps the same reciprocal in a denormal-divisor rescue (two compares, predicate
    // logic, two scalings: ~10 instructions per bin); a bin below 1e-30 (-600 dB) has no phase worth rescuing
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(mx));
    return mx > 1e-30f ? mn * rc : 0.f;
#endif
}

PV_DEV float2 phase_poly2(float2 t)                      // atan(t) / (2 pi) for two ratios: [0, 1/8] turn
{
    const float2 s = f2mul(t, t);
    float2 q = f2bc(-0.0006453014793805778f);
    q = f2fma(q, s, f2bc(0.0034795869141817093f));
    q = f2fma(q, s, f2bc(-0.00889870710670948f));
    q = f2fma(q, s, f2bc(0.015346021391451359f));
    q = f2fma(q, s, f2bc(-0.02213626727461815f));
    q = f2fma(q, s, f2bc(0.031745944172143936f));
    q = f2fma(q, s, f2bc(-0.05304612219333649f));
    q = f2fma(q, s, f2bc(0.15915483236312866f));
    return f2mul(q, t);
}

PV_DEV uint32_t phase_finish(float r, float re, float im)
{
    r = fabsf(im) > fabsf(re) ? 0.25f - r : r;
    r = re < 0.f ? 0.5f - r : r;
    r = im < 0.f ? -r : r;
    // r*2^32 reaches +-2^31 (r = +-1/2 turn), outside the int32 range: convert r*2^31 (32-bit F2I instead of the slow 64-bit one)
    // and double it.  A float r has 24 mantissa bits, so for |r| >= 2^-7 turn the product is an integer and nothing is lost; below
    // that the phase is quantised to 2^-31 turn (1.5e-9 rad)
#ifdef PV_HOST_EMUL
    return (uint32_t)((int32_t)lrintf(r * 2147483648.0f)) << 1;
#else
    return (uint32_t)__float2int_rn(r * 2147483648.0f) << 1;
#endif
}

PV_DEV uint32_t phase_turns32(float re, float im)
{
    const float t = phase_ratio(re, im);
    return phase_finish(phase_poly2(make_float2(t, t)).x, re, im);
}

// two bins at once: the polynomial (ten of the ~25 instructions of a bin) runs once on a packed pair
PV_DEV void phase_turns32x2(float2 xa, float2 xb, uint32_t &Pa, uint32_t &Pb)
{
    const float2 r = phase_poly2(make_float2(phase_ratio(xa.x, xa.y), phase_ratio(xb.x, xb.y)));
    Pa = phase_finish(r.x, xa.x, xa.y);
    Pb = phase_finish(r.y, xb.x, xb.y);
}

// |X|: one MUFU (sqrt.approx.ftz, 1 ulp) instead of the IEEE sqrtf sequence
PV_DEV float fast_sqrt(float v)
{
#ifdef PV_HOST_EMUL
    return sqrtf(v);
#else
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
#endif
}

// a * b + c and (signed) d * b + c modulo 2^64 with a 32-bit multiplier: one wide multiply-add for the low word of
// b and 32-bit multiply-adds into the high word (the compiler's generic 64 x 64 product costs twice as much)
PV_DEV unsigned long long mad_u32_u64(uint32_t a, unsigned long long b, unsigned long long c)
{
    const unsigned long long r = (unsigned long long)a * (uint32_t)b + c;
    const uint32_t hi = (uint32_t)(r >> 32) + a * (uint32_t)(b >> 32);
    return ((unsigned long long)hi << 32) | (uint32_t)r;
}
PV_DEV unsigned long long mad_s32_u64(int32_t d, unsigned long long b, unsigned long long c)
{
    const unsigned long long r = mad_u32_u64((uint32_t)d, b, c);
    // (uint32)d = d + 2^32 for negative d: take b << 32 back out
    const uint32_t hi = (uint32_t)(r >> 32) + (uint32_t)(d >> 31) * (uint32_t)b;
    return ((unsigned long long)hi << 32) | (uint32_t)r;
}

// The accumulator update of one synthesis bin, psi += a_hi * bqs + D * Rq (mod 2^64), in four multiply-adds.
//   e8 = 8 * a_hi (the byte offset of the source bin's {|X|, D} slot, which the slot loop has anyway) times bq = bqs / 8;
//   Rq = rq_hi * 2^32 + rq_lo with rq_lo SIGNED: D * rq_lo is then one signed 32 x 32 -> 64-bit multiply-add, and only the low
//   word of D * rq_hi matters -- no sign fix-up for negative D.
// A stream's first frame sets psi = D << 32 (D = the analysis phase itself then): the same update with the multipliers
// (bq, Rq) = (0, 2^32) on the zero-initialised accumulator, chosen once per frame instead of two selects per bin.
struct PsiMul { uint32_t bq_lo, bq_hi; int32_t rq_lo; uint32_t rq_hi; };

PV_DEV PsiMul psi_mul(unsigned long long bqs, unsigned long long Rq, bool first)
{
    const unsigned long long bq = bqs >> 3;          // bqs = x << (32 - lgN): a multiple of 2^20
    PsiMul m;
    m.bq_lo = first ? 0u : (uint32_t)bq;
    m.bq_hi = first ? 0u : (uint32_t)(bq >> 32);
    m.rq_lo = first ? 0 : (int32_t)(uint32_t)Rq;
    m.rq_hi = first ? 1u : (uint32_t)(Rq >> 32) + ((uint32_t)Rq >> 31);
    return m;
}

PV_DEV unsigned long long psi_step(unsigned long long ps, uint32_t e8, int32_t d, const PsiMul &m)
{
    unsigned long long t = (unsigned long long)e8 * m.bq_lo + ps;
    t = (unsigned long long)((long long)d * (long long)m.rq_lo + (long long)t);
    uint32_t hi = (uint32_t)(t >> 32);
    hi += e8 * m.bq_hi;
    hi += (uint32_t)d * m.rq_hi;
    return ((unsigned long long)hi << 32) | (uint32_t)t;
}

PV_DEV float2 cis_turns64(unsigned long long psi)
{
    // top 32 bits as signed turns in [-0.5, 0.5), converted to radians in ONE multiply (2 pi * 2^-32)
    const float ang = (float)(int32_t)(psi >> 32) * 1.4629180792671596e-9f;
    float s, c;
#ifdef PV_HOST_EMUL
    s = sinf(ang); c = cosf(ang);
#else
    // the argument is already reduced to [-pi, pi): the SFU approximations are accurate to 2^-21 absolute
    // there (4e-7 of the bin magnitude, -128 dB), and cost 2 MUFU instead of ~25 instructions
    __sincosf(ang, &s, &c);
#endif
    return make_float2(c, s);
}

// X[k] and X[M-k] of the N-point real spectrum (M = N/2) from a = C[k], b = C[M-k], w = W_N^k
// h = 1/2 times whatever the caller wants the spectrum scaled by (the output gain: linear all the way to the overlap-add)
PV_DEV void split_both(float2 a, float2 b, float2 w, float h, float2 &xk, float2 &xm)
{
    const float2 e = f2mul(f2add(a, cconj(b)), f2bc(h));
    const float2 o = f2mul(f2add(mul_mj(a), f2swap(b)), f2bc(h));          // (a.y + b.y, b.x - a.x) * h
    const float2 t = cmul(w, o);
    xk = f2add(e, t);
    xm = f2add(cconj(e), make_float2(-t.x, t.y));                          // conj(e - t)
}

// Loop-invariant per-thread twiddle bases (N = 2048 layout: both forward passes and inverse pass 2 use
// split radix-16 butterflies, thread = (butterfly b = tid % 64, half = tid / 64), outputs 2q + half).
// The 8 twiddles of a pass are o * g1^q0 * g2^q1 * g4^q2 for q = q0 + 2 q1 + 4 q2.
struct TwBase4 { float2 o, g1, g2, g4; };

PV_DEV void tw_expand(const TwBase4 &b, float2 (&e)[8])
{
    e[0] = b.o;
    e[1] = cmul(b.o, b.g1);
    e[2] = cmul(b.o, b.g2);
    e[3] = cmul(e[1], b.g2);
    e[4] = cmul(b.o, b.g4);
    e[5] = cmul(e[1], b.g4);
    e[6] = cmul(e[2], b.g4);
    e[7] = cmul(e[3], b.g4);
}

struct CThreadTw {
    TwBase4 p1, p2;          // forward pass 1, forward pass 2 (the inverse pass-2 twiddles come from the L1-resident table:
                             // with packed arithmetic eight more live registers cost more in spills than four loads per frame)
    float2 wN, w2, w4;       // W_N^u, W_N^2u, W_N^4u  (split / pack / inverse pass 1), p side: u = tid
    float2 q1, q2, q4;       // same for the q side: u_q = tid ? tid : B3/2 (thread 0 owns columns 0 and B3/2, see
                             // ThreadTw in pv_fused_core.cuh: one code path for all threads)
};

// exact table values for the bases (cis of integer fractions, rounded once)
template <int LOG2N>
PV_DEV CThreadTw load_cthread_tw(int tid, const CTables &tb)
{
    using C = CShape<LOG2N>;
    constexpr int S1 = C::S1, R2 = C::R2;
    CThreadTw t;
    auto one = make_float2(1.f, 0.f);
    t.p1 = t.p2 = TwBase4{one, one, one, one};
    if constexpr (2 * C::C1 == C::T) {
        // forward pass 1 (split radix 16): W_M^{k1 t1}, k1 = 2q + half, t1 = tid % C1; table row k1-1
        const int half = tid / C::C1;
        const int t1 = tid % C::C1;
        t.p1.o = half ? PV_LDG(tb.ctw1 + 0 * S1 + t1) : one;
        t.p1.g1 = PV_LDG(tb.ctw1 + 1 * S1 + t1);
        t.p1.g2 = PV_LDG(tb.ctw1 + 3 * S1 + t1);
        t.p1.g4 = PV_LDG(tb.ctw1 + 7 * S1 + t1);
    }
    if constexpr (2 * C::C2 == C::T) {
        // forward pass 2 (split radix 16): W_S1^{k2 n3}, n3 = (tid % C2) / R1
        const int half = tid / C::C2;
        const int n3 = (tid % C::C2) / C::R1;
        t.p2.o = half ? PV_LDG(tb.ctw2 + 0 * 4 + n3) : one;
        t.p2.g1 = PV_LDG(tb.ctw2 + 1 * 4 + n3);
        t.p2.g2 = PV_LDG(tb.ctw2 + 3 * 4 + n3);
        t.p2.g4 = PV_LDG(tb.ctw2 + 7 * 4 + n3);
    }
    t.wN = PV_LDG(tb.tw2n + 2 * tid);
    t.w2 = PV_LDG(tb.tw2n + 4 * tid);
    t.w4 = cmul(t.w2, t.w2);
    const int uq = tid ? tid : C::B3 / 2;
    t.q1 = PV_LDG(tb.tw2n + 2 * uq);
    t.q2 = PV_LDG(tb.tw2n + 4 * uq);
    t.q4 = cmul(t.q2, t.q2);
    return t;
}

struct CState {             // per-thread registers carried across the frames of a stream
    // EXPECTED phase of the next frame per slot: previous phase + nomA[bin], nomA[bin] = (bin * Ha * 2^32 / N) mod 2^32
    // (see pv_capi.cu); 0 before the first frame.  The phase difference of a frame is then one subtraction, D = P - Pexp, and
    // the first frame's D = P (the phase itself, pv_oracle.c) needs no special case.
    uint32_t Pexp[9];
    int have_prev;
};

// bin owned by slot sl of thread u: slots 0..3 = p side (u + B3 j), 4..7 = q side (B3 - u + B3 j),
// slot 8 = bin N/2 (thread 0 only)
template <int B3>
PV_DEV int slot_bin(int u, int sl)
{
    if (sl == 8) return 4 * B3;
    const int j = sl & 3;
    if (sl < 4) return (u == 0 ? 0 : u) + B3 * j;
    return (u == 0 ? B3 / 2 : B3 - u) + B3 * j;
}

// nomA of the bin of a slot (cold paths: carried state in and out)
template <int LOG2N>
PV_DEV uint32_t slot_nomA(int u, int sl, int Ha)
{
    return ((uint32_t)slot_bin<CShape<LOG2N>::B3>(u, sl) * (uint32_t)Ha) << (32 - LOG2N);
}

// The same for all slots of a thread from two products: bins u + B3 j and (B3 - u) + B3 j step by (B3 * Ha) << (32 - LOG2N)
// = Ha << 29 (B3 = N / 8)
struct SlotNomA {
    uint32_t p0, q0, step;
    PV_DEV uint32_t operator()(int sl) const { return sl == 8 ? step << 2 : (sl < 4 ? p0 : q0) + (uint32_t)(sl & 3) * step; }
};
template <int LOG2N>
PV_DEV SlotNomA slot_nomA_all(int u, int Ha)
{
    SlotNomA n;
    n.step = (uint32_t)Ha << 29;
    n.p0 = ((uint32_t)u * (uint32_t)Ha) << (32 - LOG2N);
    n.q0 = u ? n.step - n.p0 : (uint32_t)Ha << 28;       // bin B3 - u; thread 0: bin B3 / 2
    return n;
}

// Forward transform of one frame: X[slot] = spectrum at the bins owned by this thread, wp[j] = W_N^bin of
// the p-side slots (reused by the Hermitian pack).  Two barriers; `hook` runs after the first.
template <int LOG2N, bool TWREG, class Sync, class Hook>
PV_DEV void cforward(int tid, const FrameIO &io, const CTables &tb, const CThreadTw &tt, const float *ring,
                     float2 *bufA, float2 *bufB, Sync sync, Hook hook, float2 (&X)[9], float2 (&wp)[4])
{
    using C = CShape<LOG2N>;
    using S = Shape<LOG2N>;
    constexpr int N = C::N, T = C::T, B3 = C::B3, R1 = C::R1, R2 = C::R2, M = C::M, S1 = C::S1, NB = C::NB;
    const int rbase = (int)(io.base & (N - 1));
    // Loads the R windowed sample pairs of one pass-1 butterfly.  The ring / global choice is made ONCE per
    // butterfly, not per pair: the hot (ring) path stays one straight run of instructions instead of sixteen
    // short ones separated by the bounds-checked global fallback, which the instruction cache pays for.
    auto win_pair = [&](int n1, int t1) {
        return PV_LDG(reinterpret_cast<const float2 *>(tb.win + ((N / 2 + 2 * (n1 * S1 + t1)) & (N - 1))));
    };
    auto ld_block = [&](auto &v, int t1, auto windowed) {
        constexpr int R = (int)(sizeof(v) / sizeof(v[0]));
        if (ring != nullptr) {
#pragma unroll
            for (int n1 = 0; n1 < R; n1++) {
                const int i = (N / 2 + 2 * (n1 * S1 + t1)) & (N - 1);
                v[n1] = *reinterpret_cast<const float2 *>(ring + ((rbase + i) & (N - 1)));
            }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < R; n1++) {
                const int i = (N / 2 + 2 * (n1 * S1 + t1)) & (N - 1);
                const long long g = io.base + i;
                float2 x;
                if (io.vec_ok && g + 1 < io.n_in) x = PV_LDG(reinterpret_cast<const float2 *>(io.in + g));
                else x = make_float2(g < io.n_in ? PV_LDG(io.in + g) : 0.f, g + 1 < io.n_in ? PV_LDG(io.in + g + 1) : 0.f);
                v[n1] = x;
            }
        }
        if constexpr (decltype(windowed)::value) {
#pragma unroll
            for (int n1 = 0; n1 < R; n1++) v[n1] = f2mul(v[n1], win_pair(n1, t1));
        }
    };

    // ---- forward pass 1: c[m] = (f[(N/2 + 2m) mod N], f[.. + 1]), m = n1*S1 + t1 ----
    if constexpr (2 * C::C1 == T) {
        const int t1 = tid % C::C1, half = tid / C::C1;
        float2 v[16], o[8];
        ld_block(v, t1, std::false_type{});                     // the window goes into the butterfly's first stage
        auto w = [&](int n1) { return win_pair(n1, t1); };
        if (half == 0) dft16_half_win<-1, false>(v, w, o);
        else dft16_half_win<-1, true>(v, w, o);
        float2 e[8];
        constexpr bool TWREG1 = TWREG || (2 * C::C1 == T);      // window 4096: pass 1 is a split radix 16 too
        if constexpr (TWREG1) tw_expand(tt.p1, e);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int k1 = 2 * q + half;
            float2 r = o[q];
            if (k1 != 0) r = cmul(r, TWREG1 ? e[q] : PV_LDG(tb.ctw1 + (k1 - 1) * S1 + t1));
            bufA[k1 * C::LD1 + t1] = r;
        }
    } else {
#pragma unroll
        for (int t1 = tid; t1 < C::C1; t1 += T) {
            float2 v[R1];
            ld_block(v, t1, std::true_type{});
            dft<R1, -1>(v);
            bufA[t1] = v[0];
#pragma unroll
            for (int k1 = 1; k1 < R1; k1++) bufA[k1 * C::LD1 + t1] = cmul(v[k1], PV_LDG(tb.ctw1 + (k1 - 1) * S1 + t1));
        }
    }
    sync();
    hook();
    // ---- forward pass 2: butterflies (k1, n3), radix R2 over n2 ----
    if constexpr (2 * C::C2 == T) {
        const int b = tid % C::C2, half = tid / C::C2;
        const int k1 = b % R1, n3 = b / R1;
        float2 v[16], o[8];
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) v[n2] = bufA[k1 * C::LD1 + n2 * 4 + n3];
        if (half == 0) dft16_half<-1, false>(v, o);
        else dft16_half<-1, true>(v, o);
        float2 e[8];
        if constexpr (TWREG) tw_expand(tt.p2, e);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int k2 = 2 * q + half;
            float2 r = o[q];
            if (k2 != 0) r = cmul(r, TWREG ? e[q] : PV_LDG(tb.ctw2 + (k2 - 1) * 4 + n3));
            bufB[(k1 + R1 * k2) * C::LD2 + n3] = r;
        }
    } else if constexpr (4 * C::C2 == T && R2 == 32) {
        // window 4096: 64 radix-32 butterflies for 256 threads -> four threads per butterfly (outputs k2 = quarter + 4 j)
        const int b = tid % C::C2, quarter = tid / C::C2;       // warp-uniform
        const int k1 = b % R1, n3 = b / R1;
        float2 v[32], o[8];
#pragma unroll
        for (int n2 = 0; n2 < 32; n2++) v[n2] = bufA[k1 * C::LD1 + n2 * 4 + n3];
        if (quarter == 0) dft32_quarter<-1, 0>(v, o);
        else if (quarter == 1) dft32_quarter<-1, 1>(v, o);
        else if (quarter == 2) dft32_quarter<-1, 2>(v, o);
        else dft32_quarter<-1, 3>(v, o);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k2 = quarter + 4 * j;
            float2 r = o[j];
            if (k2 != 0) r = cmul(r, PV_LDG(tb.ctw2 + (k2 - 1) * 4 + n3));
            bufB[(k1 + R1 * k2) * C::LD2 + n3] = r;
        }
    } else if constexpr (R2 <= 16) {
#pragma unroll
        for (int b = tid; b < C::C2; b += T) {
            const int k1 = b % R1, n3 = b / R1;
            float2 v[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) v[n2] = bufA[k1 * C::LD1 + n2 * 4 + n3];
            dft<R2, -1>(v);
            bufB[k1 * C::LD2 + n3] = v[0];
#pragma unroll
            for (int k2 = 1; k2 < R2; k2++)
                bufB[(k1 + R1 * k2) * C::LD2 + n3] = cmul(v[k2], PV_LDG(tb.ctw2 + (k2 - 1) * 4 + n3));
        }
    }
    sync();
    // ---- forward pass 3 (radix 4, paired) + real-FFT split ----
    const int u = tid;
    const int tP = u, tQ = (u == 0) ? B3 / 2 : B3 - u;
    float2 P[4], Q[4];
#pragma unroll
    for (int n3 = 0; n3 < 4; n3++) {
        P[n3] = bufB[tP * C::LD2 + n3];
        Q[n3] = bufB[tQ * C::LD2 + n3];
    }
    dft<4, -1>(P);
    dft<4, -1>(Q);
    {
        // One path for all threads.  Regular thread: pairs (P[j], Q[3-j]) -> X[p slot j], X[q slot 3-j].
        // Thread 0 (columns 0 and B3/2): (P0,P0) -> bins 0, N/2; (P1,P3); (P2,P2); (Q0,Q3); (Q1,Q2).
        const bool u0 = (u == 0);
        auto sel = [&](float2 a, float2 b) { return u0 ? a : b; };
        // the spectrum leaves the split scaled by gain / N: magnitudes carry the output scale from here on (phases do not
        // care), and the overlap-add multiplies by the bare window
        const float h = 0.5f * tb.scale;
        // W_N^{u + B3 j} = W_N^u * exp(-2 pi i j/8)
        wp[0] = tt.wN; wp[1] = twid16<2, -1>(tt.wN); wp[2] = twid16<4, -1>(tt.wN); wp[3] = twid16<6, -1>(tt.wN);
        float2 xk0, xm0, xk1, xm1, xk2, xm2, xk3, xm3;
        split_both(P[0], sel(P[0], Q[3]), wp[0], h, xk0, xm0);
        split_both(P[1], sel(P[3], Q[2]), wp[1], h, xk1, xm1);
        split_both(P[2], sel(P[2], Q[1]), wp[2], h, xk2, xm2);
        split_both(sel(Q[0], P[3]), sel(Q[3], Q[0]), sel(tt.q1, wp[3]), h, xk3, xm3);
        float2 xke = make_float2(0.f, 0.f), xme = xke;
        if (u0) split_both(Q[1], Q[2], twid16<2, -1>(tt.q1), h, xke, xme);       // W_N^{3 B3/2}
        X[0] = xk0; X[1] = xk1; X[2] = xk2;
        X[3] = sel(xm1, xk3);
        X[4] = sel(xk3, xm3);
        X[5] = sel(xke, xm2);
        X[6] = sel(xme, xm1);
        X[7] = sel(xm3, xm0);
        X[8] = sel(xm0, make_float2(0.f, 0.f));
    }
    (void)M;
}

// Analysis-only ("phase-carry aggregate") mode of the SAME frame body.  The aggregate and the processing pass
// must produce bit-identical phases P (the split / sharded result is bit-identical to the sequential one only
// then), so they share one kernel instantiation and therefore one compiled copy of the forward transform: two
// separately compiled copies may contract multiply-adds differently.
struct AggCtx {
    bool on;                 // analysis only: no synthesis, no barriers after the forward transform
    long long *sum;          // [NB] per-bin running sum of D for this frame (shared memory, thread-private slots)
    uint32_t *P_first;       // global [NB] or null: phase of the first analysed frame of the segment
    float2 *md_row;          // global [NB + 1] or null: {|X|, D} of this frame, kept for the processing pass (analysis store)
};

// MODE 0: the normal frame (and the analysis-only mode that shares its compiled forward transform).  The two halves of the
// stored-analysis split are kernel instantiations of their own, so that the normal one -- at the 128-register limit -- carries
// none of them: MODE 2 = analysis-only that also STORES {|X|, D} (AggCtx::md_row), MODE 1 = processing FROM the stored values.
template <int LOG2N, int MODE = 0, class Sync, class Hook, class PreLast>
PV_DEV void frame_corrected(int tid, const FrameIO &io, const CTables &tb, const CThreadTw &tt, const float *ring,
                            float2 *bufA, float2 *bufB, float2 *mdS, unsigned long long *psi, float *acc,
                            CState &st, int pos0, int Hs, Sync sync, Hook hook, PreLast pre_last_sync,
                            const AggCtx agg = AggCtx{false, nullptr, nullptr, nullptr})
{
    using C = CShape<LOG2N>;
    using S = Shape<LOG2N>;
    constexpr int N = C::N, B3 = C::B3, M = C::M, NB = C::NB;
    constexpr bool TWREG = (2 * C::C1 == C::T) && (2 * C::C2 == C::T);
    const int u = tid;
    const int tP = u, tQ = (u == 0) ? B3 / 2 : B3 - u;
    float2 X[9], wp[4];
    const bool first = st.have_prev == 0;
    if constexpr (MODE == 1) {
        // Stored analysis: mdS already holds {|X|, D} of this frame (the caller's `hook` waits for its copy), so the forward
        // transform, the analysis and two of the five barriers of a frame go.  The barrier orders the previous frame's
        // overlap-add before the hook's emit, as the first barrier of the forward transform does otherwise.
        sync();
        hook();
        wp[0] = tt.wN; wp[1] = twid16<2, -1>(tt.wN); wp[2] = twid16<4, -1>(tt.wN); wp[3] = twid16<6, -1>(tt.wN);
    } else {
    cforward<LOG2N, TWREG>(tid, io, tb, tt, ring, bufA, bufB, sync, hook, X, wp);
    // ---- analysis: magnitude, phase (turns*2^32), unwrapped phase difference ----
    const SlotNomA nomA = slot_nomA_all<LOG2N>(u, tb.Ha);
    if (agg.on) {
        uint32_t Pn[9];
#pragma unroll
        for (int sl = 0; sl < 8; sl += 2) phase_turns32x2(X[sl], X[sl + 1], Pn[sl], Pn[sl + 1]);
        Pn[8] = 0u;
        if (u == 0) Pn[8] = phase_turns32(X[8].x, X[8].y);
#pragma unroll
        for (int sl = 0; sl < 9; sl++) {
            if (sl == 8 && u != 0) break;
            const int bin = slot_bin<B3>(u, sl);
            const uint32_t Pc = Pn[sl];
            const int32_t dd = (int32_t)(Pc - st.Pexp[sl]);          // first frame: Pexp = 0, D = the phase itself
            // branch-free in the common case (see the synthesis slot loop): a first frame adds 0
            agg.sum[bin] += first ? 0ll : (long long)dd;
            st.Pexp[sl] = Pc + nomA(sl);
            if constexpr (MODE == 2) {
                if (agg.md_row) {        // uniform: the same two values the processing pass would put in shared memory
                    const float2 x = X[sl];
                    agg.md_row[bin] = make_float2(fast_sqrt(x.x * x.x + x.y * x.y), __int_as_float(dd));
                }
            }
        }
        if constexpr (MODE == 2) {
            if (agg.md_row && u == 0) agg.md_row[NB] = make_float2(0.f, 0.f);      // the dummy bin
        }
        if (first && agg.P_first) {          // once per segment, outside the slot loop
#pragma unroll
            for (int sl = 0; sl < 9; sl++) {
                if (sl == 8 && u != 0) break;
                agg.P_first[slot_bin<B3>(u, sl)] = Pn[sl];
            }
        }
        st.have_prev = 1;
        return;          // the next frame's first barrier orders the reuse of the exchange buffers
    }
    uint32_t Pn[9];
#pragma unroll
    for (int sl = 0; sl < 8; sl += 2) phase_turns32x2(X[sl], X[sl + 1], Pn[sl], Pn[sl + 1]);
    Pn[8] = 0u;
    if (u == 0) Pn[8] = phase_turns32(X[8].x, X[8].y);
#pragma unroll
    for (int sl = 0; sl < 9; sl++) {
        if (sl == 8 && u != 0) break;
        const int bin = slot_bin<B3>(u, sl);
        const float2 x = X[sl];
        const uint32_t Pc = Pn[sl];
        const int32_t dd = (int32_t)(Pc - st.Pexp[sl]);
        // {|X|, D} of a bin side by side: one 8-byte store here, and ONE 8-byte load per synthesis bin in the common case of a
        // single source bin (pitch ratio >= 1)
        mdS[bin] = make_float2(fast_sqrt(x.x * x.x + x.y * x.y), __int_as_float(dd));
        st.Pexp[sl] = Pc + nomA(sl);
    }
    sync();
    }
    st.have_prev = 1;
    // ---- synthesis, one voice at a time ----
    for (int v = 0; v < tb.V; v++) {
        const uint32_t *gt = tb.gather + ((size_t)v * C::T + u) * 9;
        unsigned long long *ps = psi + (size_t)v * NB;
        const PsiMul pm = psi_mul(tb.bqs[v], tb.Rq[v], first);
        const bool multi = tb.multi[v] != 0;
        float2 Y[9];
        // The slot loop is branch-free on the common path (pitch ratio >= 1: every synthesis bin has at most one source
        // bin): the nine slots' chains -- gather, four multiply-adds, int -> float, two MUFU, one packed multiply -- are
        // independent, and only straight-line code lets the scheduler interleave them.  The multi-source loop of ratios < 1
        // and the first-frame initialisation used to sit inside every slot as branches.
        auto slots = [&](auto multi_tag) {
            constexpr bool MULTI = decltype(multi_tag)::value;
#pragma unroll
            for (int sl = 0; sl < 9; sl++) {
                Y[sl] = make_float2(0.f, 0.f);
                if (sl == 8 && u != 0) break;
                const int s = slot_bin<B3>(u, sl);
                const uint32_t ge = PV_LDG(gt + sl);         // no source bin: the dummy bin NB ({0, 0}: the slot stays zero)
                uint32_t e8;                                 // 8 * a_hi
                float2 md;
                float m;
                if constexpr (MULTI) {                       // entry = a_lo | a_hi << 16
                    const uint32_t lo = ge & 0xffffu, hi = ge >> 16;
                    md = mdS[lo];
                    m = md.x;
#pragma unroll 1
                    for (uint32_t a = lo + 1; a <= hi; a++) { md = mdS[a]; m += md.x; }   // ascending, as the specification sums
                    e8 = hi << 3;
                } else {                                     // entry = 8 * a_hi: the load address and the multiplier as they are
                    e8 = ge;
                    md = *reinterpret_cast<const float2 *>(reinterpret_cast<const char *>(mdS) + e8);
                    m = md.x;
                }
                const int32_t d = __float_as_int(md.y);      // D of the LAST source bin (a_hi)
                const unsigned long long p = psi_step(ps[s], e8, d, pm);
                if (e8 != (uint32_t)(8 * NB)) ps[s] = p;     // an empty range leaves the accumulator untouched
                const float2 cs = cis_turns64(p);
                Y[sl] = f2mul(cs, f2bc(m));
            }
        };
        if (multi) slots(std::true_type{});                  // pitch ratio < 1 only (uniform per voice)
        else slots(std::false_type{});
        // Hermitian pack (same register pattern as the compat kernel); exp(+2 pi i k/N) = conj(W_N^k)
        float2 Zp[4], Zq[4];
        {
            const bool u0 = (u == 0);
            auto sel = [&](float2 a, float2 b) { return u0 ? a : b; };
            if (u0) { Y[0].y = 0.f; Y[8].y = 0.f; }          // the Hermitian inverse ignores Im of bins 0 and N/2
            const float2 q1c = cconj(tt.q1);
            // p side: Z[u + B3 j] pairs with bin M - (u + B3 j): q slot 3-j (thread 0: p slot 4-j, slot "4" = bin N/2)
            Zp[0] = herm_pack(Y[0], sel(Y[8], Y[7]), cconj(wp[0]));
            Zp[1] = herm_pack(Y[1], sel(Y[3], Y[6]), cconj(wp[1]));
            Zp[2] = herm_pack(Y[2], sel(Y[2], Y[5]), cconj(wp[2]));
            Zp[3] = herm_pack(Y[3], sel(Y[1], Y[4]), cconj(wp[3]));
            // q side: bin (B3 - u_q) + B3 j, W_N^bin = conj(W_N^{u_q}) * exp(-2 pi i (j+1)/8); partner p slot 3-j
            // (thread 0: q slot 3-j)
            Zq[0] = herm_pack(Y[4], sel(Y[7], Y[3]), cconj(twid16<2, -1>(q1c)));
            Zq[1] = herm_pack(Y[5], sel(Y[6], Y[2]), cconj(twid16<4, -1>(q1c)));
            Zq[2] = herm_pack(Y[6], sel(Y[5], Y[1]), cconj(twid16<6, -1>(q1c)));
            Zq[3] = herm_pack(Y[7], sel(Y[4], Y[0]), cconj(twid16<8, -1>(q1c)));
        }
        Tables itb{nullptr, nullptr, tb.tw2n, tb.itw1, tb.itw2, tb.win};
        auto last = [&]() { if (v + 1 == tb.V) pre_last_sync(); };
        {
            // exp(+2 pi i m1 u/(N/2)) = conj(W_N^{2 m1 u}); for t1 = B3-u_q: j^m1 * W_N^{2 m1 u_q}
            const float2 w6 = cmul(tt.w2, tt.w4), q6 = cmul(tt.q2, tt.q4);
            inverse_1_tw<LOG2N>(tP, cconj(tt.w2), cconj(tt.w4), cconj(w6), Zp, bufA);
            inverse_1_tw<LOG2N>(tQ, mul_pj(tt.q2), make_float2(-tt.q4.x, -tt.q4.y), mul_mj(q6), Zq, bufA);
        }
        inverse_23_ola<LOG2N, Sync, decltype(last), TableTw2, true>(tid, itb, bufA, bufB, acc + (size_t)v * N, pos0, Hs, false, 1.f,
                                                                    sync, last);
        (void)S::T;
    }
    (void)M;
}

}  // namespace pvfused
