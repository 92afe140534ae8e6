// pv_fused_tables.h -- host-side construction of the twiddle tables of the fused kernels
// (double precision, rounded once to float).  Plain C++: shared by pv_capi.cu and the CPU
// emulation harness in tests/emul.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#ifndef PV_HOST_EMUL
#include <cuda_runtime.h>
#endif

struct HostTables {
    std::vector<float2> tw1, tw2, tw2n, itw1, itw2;
    std::vector<float2> ctw1, ctw2;          // corrected-mode forward (N/2 complex points)
};

inline float2 pv_cis(double turns)   // exp(j * 2*pi * turns)
{
    const double a = 2.0 * 3.14159265358979323846 * turns;
    float2 r;
    r.x = (float)cos(a);
    r.y = (float)sin(a);
    return r;
}

// Shapes must match pvfused::Shape<LOG2N> (pv_fused_core.cuh).
inline void build_tables(int log2n, HostTables &t)
{
    const int N = 1 << log2n, B3 = N / 8, R1 = (log2n >= 10) ? 16 : 8, R2 = B3 / R1, S1 = N / R1;
    t.tw1.resize((size_t)(R1 - 1) * S1);
    for (int k1 = 1; k1 < R1; k1++)
        for (int t1 = 0; t1 < S1; t1++) t.tw1[(size_t)(k1 - 1) * S1 + t1] = pv_cis(-(double)((long)k1 * t1 % N) / N);
    t.tw2.resize((size_t)(R2 - 1) * 8);
    for (int k2 = 1; k2 < R2; k2++)
        for (int n3 = 0; n3 < 8; n3++) t.tw2[(size_t)(k2 - 1) * 8 + n3] = pv_cis(-(double)(k2 * n3) / S1);
    t.tw2n.resize(N);
    for (int k = 0; k < N; k++) t.tw2n[k] = pv_cis(-(double)k / (2.0 * N));
    t.itw1.resize((size_t)3 * B3);
    for (int m1 = 1; m1 < 4; m1++)
        for (int t1 = 0; t1 < B3; t1++) t.itw1[(size_t)(m1 - 1) * B3 + t1] = pv_cis((double)(m1 * t1) / (N / 2));
    // corrected forward: M = N/2 = R1*R2*4, S1c = M/R1 = 4*R2
    const int M = N / 2, S1c = M / R1;
    t.ctw1.resize((size_t)(R1 - 1) * S1c);
    for (int k1 = 1; k1 < R1; k1++)
        for (int t1 = 0; t1 < S1c; t1++) t.ctw1[(size_t)(k1 - 1) * S1c + t1] = pv_cis(-(double)((long)k1 * t1 % M) / M);
    t.ctw2.resize((size_t)(R2 - 1) * 4);
    for (int k2 = 1; k2 < R2; k2++)
        for (int n3 = 0; n3 < 4; n3++) t.ctw2[(size_t)(k2 - 1) * 4 + n3] = pv_cis(-(double)(k2 * n3) / S1c);
    t.itw2.resize((size_t)(R1 - 1) * R2);
    for (int m2 = 1; m2 < R1; m2++)
        for (int n3 = 0; n3 < R2; n3++) t.itw2[(size_t)(m2 - 1) * R2 + n3] = pv_cis((double)(m2 * n3) / B3);
}

// The same packing in natural bin order, [v][NB], for the large-window in-place kernel (pv_generic_kernels.cu).
inline void build_gather_natural(int N, int V, const int32_t *a_lo, const int32_t *a_hi, std::vector<uint32_t> &out,
                                 int32_t *multi = nullptr)
{
    const int NB = N / 2 + 1;
    const uint32_t dummy = (uint32_t)NB | ((uint32_t)NB << 16);
    out.assign((size_t)V * NB, dummy);
    for (int v = 0; v < V; v++) {
        if (multi) multi[v] = 0;
        for (int s = 0; s < NB; s++) {
            const size_t i = (size_t)v * NB + s;
            if (a_lo[i] > a_hi[i]) continue;
            if (multi && a_hi[i] > a_lo[i]) multi[v] = 1;
            out[i] = (uint32_t)a_lo[i] | ((uint32_t)a_hi[i] << 16);
        }
    }
}

// Per-thread gather table of the corrected kernel, entry [v][u][slot] for the synthesis bin owned by slot `slot` of thread
// `u` (pvfused::slot_bin).  A voice whose synthesis bins have at most ONE source bin each (pitch ratio >= 1) stores 8 * a_hi:
// the byte offset of the source bin's {|X|, D} slot and, at the same time, the multiplier of the accumulator update
// (pvfused::psi_step) -- the slot loop needs no unpacking.  A voice that sums several analysis bins into some synthesis bin
// (pitch ratio < 1, multi[v] = 1) stores a_lo | a_hi << 16.  A synthesis bin that no analysis bin maps to points at the DUMMY
// bin NB: the kernel keeps magnitude 0 and phase difference 0 there, so the slot needs no special case.
inline void build_gather_table(int N, int V, const int32_t *a_lo, const int32_t *a_hi, std::vector<uint32_t> &out,
                               int32_t *multi = nullptr)
{
    const int T = N / 16, B3 = N / 8, NB = N / 2 + 1;
    out.assign((size_t)V * T * 9, 0u);
    for (int v = 0; v < V; v++) {
        bool mv = false;
        for (int s = 0; s < NB; s++) mv = mv || a_hi[(size_t)v * NB + s] > a_lo[(size_t)v * NB + s];
        if (multi) multi[v] = mv ? 1 : 0;
        const uint32_t dummy = mv ? (uint32_t)NB | ((uint32_t)NB << 16) : 8u * (uint32_t)NB;
        for (int u = 0; u < T; u++)
            for (int sl = 0; sl < 9; sl++) {
                uint32_t &e = out[((size_t)v * T + u) * 9 + sl];
                e = dummy;
                if (sl == 8 && u != 0) continue;                                   // slot not used
                const int j = sl & 3;
                const int bin = (sl == 8) ? 4 * B3 : (sl < 4 ? u + B3 * j : (u == 0 ? B3 / 2 : B3 - u) + B3 * j);
                const size_t i = (size_t)v * NB + bin;
                if (a_lo[i] > a_hi[i]) continue;                                   // empty range
                e = mv ? (uint32_t)a_lo[i] | ((uint32_t)a_hi[i] << 16) : 8u * (uint32_t)a_hi[i];
            }
    }
}
