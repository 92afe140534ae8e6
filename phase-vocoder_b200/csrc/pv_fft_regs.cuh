// pv_fft_regs.cuh -- register-resident radix-2/4/8/16 DFTs and complex helpers.
//
// Compiles as CUDA device code and, with PV_HOST_EMUL defined, as plain C++ so that the
// fused kernel's index algebra can be executed on the CPU by tests/emul (one std::thread per
// CUDA thread, std::barrier for __syncthreads).  The emulation is a TEST harness: the product
// only ever runs the CUDA build.
#pragma once

#ifdef PV_HOST_EMUL
#include <cmath>
#include <cstdint>
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
#define PV_DEV inline
#define PV_HD
#define PV_LDG(p) (*(p))
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define PV_DEV __device__ __forceinline__
#define PV_HD __host__ __device__
#define PV_LDG(p) __ldg(p)
#endif

namespace pvfft {

// ---- packed fp32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 work on a 64-bit register pair, one issue slot for two
// lanes; each lane rounds like the scalar instruction).  A complex number IS such a pair.  The swapped / negated
// operands written with make_float2 below cost nothing: SASS takes per-operand lane swap (.LO_HI), whole-pair and
// per-lane negation (.NP / .PN) and scalar broadcast (.F32) as modifiers.  The fused kernels are bound by
// instruction issue, so every butterfly is written on pairs.
#if defined(PV_HOST_EMUL) || defined(PV_NO_PACKED)
PV_DEV float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
PV_DEV float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
PV_DEV float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#else
PV_DEV float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
PV_DEV float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
PV_DEV float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#endif
PV_DEV float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
PV_DEV float2 f2swap(float2 a) { return make_float2(a.y, a.x); }
PV_DEV float2 f2bc(float s) { return make_float2(s, s); }

PV_DEV float2 cadd(float2 a, float2 b) { return f2add(a, b); }
PV_DEV float2 csub(float2 a, float2 b) { return f2add(a, f2neg(b)); }
// (a.x b.x - a.y b.y, a.y b.x + a.x b.y): two packed instructions
PV_DEV float2 cmul(float2 a, float2 b) { return f2fma(a, f2bc(b.x), f2mul(f2swap(a), make_float2(-b.y, b.y))); }
PV_DEV float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
// a * (+j) and a * (-j)
PV_DEV float2 mul_pj(float2 a) { return make_float2(-a.y, a.x); }
PV_DEV float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }

// a * exp(DIR * j * 2*pi * K / 16), K and DIR compile-time (DIR = -1 forward, +1 inverse)
template <int K, int DIR>
PV_DEV float2 twid16(float2 a)
{
    constexpr int k = ((K % 16) + 16) % 16;
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
    if constexpr (k == 0) return a;
    else if constexpr (k == 4) return DIR > 0 ? mul_pj(a) : mul_mj(a);
    else if constexpr (k == 8) return make_float2(-a.x, -a.y);
    else if constexpr (k == 12) return DIR > 0 ? mul_mj(a) : mul_pj(a);
    else {
        // general: (c + j*DIR*s) with c = cos(2 pi k/16), s = sin(2 pi k/16)
        constexpr float c = (k == 1 || k == 15) ? C1 : (k == 2 || k == 14) ? H : (k == 3 || k == 13) ? S1
                          : (k == 5 || k == 11) ? -S1 : (k == 6 || k == 10) ? -H : -C1;   // k == 7, 9
        constexpr float sa = (k == 1 || k == 7) ? S1 : (k == 2 || k == 6) ? H : (k == 3 || k == 5) ? C1
                           : (k == 9 || k == 15) ? -S1 : (k == 10 || k == 14) ? -H : -C1;  // k == 11, 13
        constexpr float s = DIR > 0 ? sa : -sa;
        return f2fma(a, f2bc(c), f2mul(f2swap(a), make_float2(-s, s)));
    }
}

template <int DIR>
PV_DEV void dft2(float2 &a, float2 &b)
{
    const float2 t = csub(a, b);
    a = cadd(a, b);
    b = t;
}

// natural order in, natural order out
template <int DIR>
PV_DEV void dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3)
{
    const float2 t0 = cadd(v0, v2), t1 = csub(v0, v2), t2 = cadd(v1, v3);
    const float2 d = csub(v1, v3);
    const float2 t3 = DIR > 0 ? mul_pj(d) : mul_mj(d);
    v0 = cadd(t0, t2);
    v2 = csub(t0, t2);
    v1 = cadd(t1, t3);
    v3 = csub(t1, t3);
}

template <int DIR>
PV_DEV void dft8(float2 &v0, float2 &v1, float2 &v2, float2 &v3, float2 &v4, float2 &v5, float2 &v6, float2 &v7)
{
    dft4<DIR>(v0, v2, v4, v6);      // even samples -> e0..e3 in (v0, v2, v4, v6)
    dft4<DIR>(v1, v3, v5, v7);      // odd samples  -> o0..o3 in (v1, v3, v5, v7)
    const float2 o1 = twid16<2, DIR>(v3), o2 = twid16<4, DIR>(v5), o3 = twid16<6, DIR>(v7);
    const float2 e0 = v0, e1 = v2, e2 = v4, e3 = v6, o0 = v1;
    v0 = cadd(e0, o0); v4 = csub(e0, o0);
    v1 = cadd(e1, o1); v5 = csub(e1, o1);
    v2 = cadd(e2, o2); v6 = csub(e2, o2);
    v3 = cadd(e3, o3); v7 = csub(e3, o3);
}

template <int R, int DIR>
PV_DEV void dft(float2 (&v)[R])
{
    static_assert(R == 2 || R == 4 || R == 8 || R == 16, "radix");
    if constexpr (R == 2) dft2<DIR>(v[0], v[1]);
    else if constexpr (R == 4) dft4<DIR>(v[0], v[1], v[2], v[3]);
    else if constexpr (R == 8) dft8<DIR>(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
    else {
        dft8<DIR>(v[0], v[2], v[4], v[6], v[8], v[10], v[12], v[14]);    // e_k in v[2k]
        dft8<DIR>(v[1], v[3], v[5], v[7], v[9], v[11], v[13], v[15]);    // o_k in v[2k+1]
        float2 e[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { e[k] = v[2 * k]; o[k] = v[2 * k + 1]; }
        o[1] = twid16<1, DIR>(o[1]); o[2] = twid16<2, DIR>(o[2]); o[3] = twid16<3, DIR>(o[3]);
        o[4] = twid16<4, DIR>(o[4]); o[5] = twid16<5, DIR>(o[5]); o[6] = twid16<6, DIR>(o[6]);
        o[7] = twid16<7, DIR>(o[7]);
#pragma unroll
        for (int k = 0; k < 8; k++) { v[k] = cadd(e[k], o[k]); v[k + 8] = csub(e[k], o[k]); }
    }
}

// One half of a 16-point DFT: outputs k = 2q + ODD, q = 0..7.  Two threads share a butterfly when
// a pass has fewer radix-16 butterflies than threads; the halves add up to exactly one dft<16>.
template <int DIR, bool ODD>
PV_DEV void dft16_half(const float2 (&a)[16], float2 (&o)[8])
{
#pragma unroll
    for (int n = 0; n < 8; n++) o[n] = ODD ? csub(a[n], a[n + 8]) : cadd(a[n], a[n + 8]);
    if constexpr (ODD) {
        o[1] = twid16<1, DIR>(o[1]); o[2] = twid16<2, DIR>(o[2]); o[3] = twid16<3, DIR>(o[3]);
        o[4] = twid16<4, DIR>(o[4]); o[5] = twid16<5, DIR>(o[5]); o[6] = twid16<6, DIR>(o[6]);
        o[7] = twid16<7, DIR>(o[7]);
    }
    dft8<DIR>(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
}

// The same with the analysis window folded into the first stage: inputs x[n] * w(n), the products of the upper half
// contracted into the fold (one FMUL2 + one FFMA2 per output instead of two FMUL2 + one FADD2).
template <int DIR, bool ODD, class W>
PV_DEV void dft16_half_win(const float2 (&x)[16], W w, float2 (&o)[8])
{
#pragma unroll
    for (int n = 0; n < 8; n++) o[n] = f2fma(x[n + 8], ODD ? f2neg(w(n + 8)) : w(n + 8), f2mul(x[n], w(n)));
    if constexpr (ODD) {
        o[1] = twid16<1, DIR>(o[1]); o[2] = twid16<2, DIR>(o[2]); o[3] = twid16<3, DIR>(o[3]);
        o[4] = twid16<4, DIR>(o[4]); o[5] = twid16<5, DIR>(o[5]); o[6] = twid16<6, DIR>(o[6]);
        o[7] = twid16<7, DIR>(o[7]);
    }
    dft8<DIR>(o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7]);
}

// ---- radix-32 pieces for window 4096 (pv_fused_core.cuh, Shape<12>: R2 = 32) ----
// cos(2 pi k / 32) for any integer k, compile time
PV_HD constexpr float pv_cos32(int k)
{
    constexpr float C[9] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f, 0.70710678118654752f,
                            0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f, 0.0f};
    k = ((k % 32) + 32) % 32;
    if (k > 16) k = 32 - k;
    return k > 8 ? -C[16 - k] : C[k];
}

// a * exp(DIR * j * 2*pi * K / 32)
template <int K, int DIR>
PV_DEV float2 twid32(float2 a)
{
    constexpr int k = ((K % 32) + 32) % 32;
    if constexpr (k % 2 == 0) return twid16<k / 2, DIR>(a);
    else {
        constexpr float c = pv_cos32(k), s = (DIR > 0 ? 1.f : -1.f) * pv_cos32(k - 8);      // sin(x) = cos(x - pi/2)
        return f2fma(a, f2bc(c), f2mul(f2swap(a), make_float2(-s, s)));
    }
}

// One half of a 32-point DFT: outputs k = 2q + ODD, q = 0..15 (two threads share a butterfly).  Two steps, so that a
// caller working in place can put a barrier between its loads and its stores with only 16 values live:
//   o[n] = a[n] +- a[n + 16]                      (dft32_fold, while loading)
//   dft32_half_finish: twiddle W32^n (odd half), 16-point DFT
template <bool ODD>
PV_DEV float2 dft32_fold(float2 lo, float2 hi) { return ODD ? csub(lo, hi) : cadd(lo, hi); }

template <int DIR, bool ODD>
PV_DEV void dft32_half_finish(float2 (&o)[16])
{
    if constexpr (ODD) {
        o[1] = twid32<1, DIR>(o[1]); o[2] = twid32<2, DIR>(o[2]); o[3] = twid32<3, DIR>(o[3]);
        o[4] = twid32<4, DIR>(o[4]); o[5] = twid32<5, DIR>(o[5]); o[6] = twid32<6, DIR>(o[6]);
        o[7] = twid32<7, DIR>(o[7]); o[8] = twid32<8, DIR>(o[8]); o[9] = twid32<9, DIR>(o[9]);
        o[10] = twid32<10, DIR>(o[10]); o[11] = twid32<11, DIR>(o[11]); o[12] = twid32<12, DIR>(o[12]);
        o[13] = twid32<13, DIR>(o[13]); o[14] = twid32<14, DIR>(o[14]); o[15] = twid32<15, DIR>(o[15]);
    }
    dft<16, DIR>(o);
}

template <int DIR, bool ODD>
PV_DEV void dft32_half(const float2 (&a)[32], float2 (&o)[16])
{
#pragma unroll
    for (int n = 0; n < 16; n++) o[n] = dft32_fold<ODD>(a[n], a[n + 16]);
    dft32_half_finish<DIR, ODD>(o);
}

// One quarter of a 32-point DFT: outputs k = Q + 4j, j = 0..7 (four threads share a butterfly).
// n = n0 + 8 n1:  X[Q + 4j] = sum_{n0} W8^{n0 j} * [ W32^{n0 Q} * sum_{n1} a[n0 + 8 n1] W4^{n1 Q} ]
template <int DIR, int Q>
PV_DEV void dft32_quarter(const float2 (&a)[32], float2 (&o)[8])
{
    static_assert(Q >= 0 && Q < 4, "quarter");
#pragma unroll
    for (int n0 = 0; n0 < 8; n0++) {
        const float2 a0 = a[n0], a1 = a[n0 + 8], a2 = a[n0 + 16], a3 = a[n0 + 24];
        if constexpr (Q == 0) o[n0] = cadd(cadd(a0, a2), cadd(a1, a3));
        else if constexpr (Q == 2) o[n0] = csub(cadd(a0, a2), cadd(a1, a3));
        else {
            const float2 d = csub(a1, a3);
            const float2 jd = DIR > 0 ? mul_pj(d) : mul_mj(d);             // (DIR*j) * (a1 - a3)
            o[n0] = (Q == 1) ? cadd(csub(a0, a2), jd) : csub(csub(a0, a2), jd);
        }
    }
    if constexpr (Q != 0) {
        o[1] = twid32<1 * Q, DIR>(o[1]); o[2] = twid32<2 * Q, DIR>(o[2]); o[3] = twid32<3 * Q, DIR>(o[3]);
        o[4] = twid32<4 * Q, DIR>(o[4]); o[5] = twid32<5 * Q, DIR>(o[5]); o[6] = twid32<6 * Q, DIR>(o[6]);
        o[7] = twid32<7 * Q, DIR>(o[7]);
    }
    dft<8, DIR>(o);
}

// Forward R-point DFT of a sequence whose middle half is zero (only v[0..R/4) and v[3R/4..R) are
// set): the zero-padded, zero-phase frame layout of the compat path.  With m = n for the low
// quarter and m = n - R for the high quarter, X[2q] = DFT_{R/2}(u)[q] and X[2q+1] = DFT_{R/2}(u')[q],
// u[m mod R/2] = v[m], u'[m mod R/2] = v[m] W_R^m.  ~20 % fewer flops than the full butterfly.
template <int R>
PV_DEV void dft_pruned_fwd(float2 (&v)[R])
{
    static_assert(R == 8 || R == 16, "radix");
    constexpr int H = R / 2, Q = R / 4, STEP = 16 / R;
    float2 e[H], o[H];
#pragma unroll
    for (int i = 0; i < Q; i++) { e[i] = v[i]; e[Q + i] = v[3 * Q + i]; }
    if constexpr (R == 16) {
        o[0] = e[0]; o[1] = twid16<1, -1>(e[1]); o[2] = twid16<2, -1>(e[2]); o[3] = twid16<3, -1>(e[3]);
        o[4] = twid16<12, -1>(e[4]); o[5] = twid16<13, -1>(e[5]); o[6] = twid16<14, -1>(e[6]); o[7] = twid16<15, -1>(e[7]);
    } else {
        o[0] = e[0]; o[1] = twid16<2, -1>(e[1]); o[2] = twid16<12, -1>(e[2]); o[3] = twid16<14, -1>(e[3]);
    }
    dft<H, -1>(e);
    dft<H, -1>(o);
#pragma unroll
    for (int q = 0; q < H; q++) { v[2 * q] = e[q]; v[2 * q + 1] = o[q]; }
    (void)STEP;
}

}  // namespace pvfft
