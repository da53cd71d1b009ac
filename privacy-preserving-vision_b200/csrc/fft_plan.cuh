// b200cam: two-pass (four-step) N-point complex FFT shared between LANES threads.
//
// N = R1*R2.  Two data distributions over the lanes of one FFT:
//   P ("natural/space" side): lane a < R2 holds element  R2*i + a,  i < R1   (R1 registers)
//   Q ("frequency" side)    : lane b < R1 holds element  b + R1*i,  i < R2   (R2 registers)
//
//   forward  (DIR=-1):  P --stepA--> smem E --stepB--> Q
//   inverse  (DIR=+1):  Q --stepC--> smem E --stepD--> P          (unnormalised)
//
// so a forward transform immediately followed by an inverse one (the convolution pattern)
// needs no reordering in between: point-wise work happens on the Q registers.
// Each step is one barrier phase; E is a padded R1 x (R2+1) (resp. R2 x (R1+1)) float2 array.
//   X[k1 + R1*k2] = sum_{n2} w_R2^{n2 k2} * w_N^{n2 k1} * sum_{n1} x[R2*n1 + n2] w_R1^{n1 k1}
// `tw` is the table tw[j] = exp(-2*pi*i*j/N), j < N.
#pragma once

#include "compat.cuh"
#include "fft_radix.cuh"

namespace b200cam {

template <int N> struct Factor;
template <> struct Factor<64> { static constexpr int R1 = 8, R2 = 8; };
template <> struct Factor<128> { static constexpr int R1 = 8, R2 = 16; };
template <> struct Factor<256> { static constexpr int R1 = 16, R2 = 16; };
template <> struct Factor<512> { static constexpr int R1 = 32, R2 = 16; };   // R2 sizes the per-thread state of the point-wise stage
template <> struct Factor<1024> { static constexpr int R1 = 32, R2 = 32; };

template <int N>
struct Plan {
    static constexpr int R1 = Factor<N>::R1;
    static constexpr int R2 = Factor<N>::R2;
    static constexpr int LANES = R1 > R2 ? R1 : R2;
    static constexpr int PITCH_AB = R2 + 1;           // E as R1 rows of R2 (+1 pad)
    static constexpr int PITCH_CD = R1 + 1;           // E as R2 rows of R1 (+1 pad)
    static constexpr int E_SIZE = N + LANES;          // float2 elements, covers both shapes
    static_assert(R1 * R2 == N, "bad factorisation");

    // P -> E.  lane a < R2, v[i] = x[R2*i + a]
    static B200_HD void stepA(float2 (&v)[R1], int a, float2* E, const float2* tw) {
        RegFFT<R1, -1>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            float2 t = v[k1];
            if (k1 > 0) t = cmul(t, ld_ro(tw + a * k1));
            E[k1 * PITCH_AB + a] = t;
        }
    }
    // The same two steps with the lane's twiddles w[k] = tw[lane*k] held in REGISTERS (persistent kernels load them once:
    // the 15 table look-ups per step are a quarter of the shared/L1 pipe traffic of a column pass).  Needs R1 == R2.
    static constexpr bool REG_TW = (R1 == R2) && (R1 <= 16);
    static B200_HD void load_tw(float2 (&w)[R1], const float2* tw, int lane) {
#pragma unroll
        for (int k = 0; k < R1; ++k) w[k] = ld_ro(tw + lane * k);
    }
    static B200_HD void stepA(float2 (&v)[R1], int a, float2* E, const float2 (&w)[R1]) {
        RegFFT<R1, -1>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) E[k1 * PITCH_AB + a] = k1 > 0 ? cmul(v[k1], w[k1]) : v[k1];
    }
    static B200_HD void stepC(float2 (&v)[R2], int b, float2* E, const float2 (&w)[R1]) {
        static_assert(R1 == R2, "register twiddles need R1 == R2");
        RegFFT<R2, +1>::run(v);
#pragma unroll
        for (int m2 = 0; m2 < R2; ++m2) E[m2 * PITCH_CD + b] = m2 > 0 ? cmulc(v[m2], w[m2]) : v[m2];
    }
    // E -> Q.  lane b < R1, result v[i] = X[b + R1*i]
    static B200_HD void stepB(float2 (&v)[R2], int b, const float2* E) {
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) v[n2] = E[b * PITCH_AB + n2];
        RegFFT<R2, -1>::run(v);
    }
    // Q -> E.  lane b < R1, v[i] = X[b + R1*i]
    static B200_HD void stepC(float2 (&v)[R2], int b, float2* E, const float2* tw) {
        RegFFT<R2, +1>::run(v);
#pragma unroll
        for (int m2 = 0; m2 < R2; ++m2) {
            float2 t = v[m2];
            if (m2 > 0) t = cmulc(t, ld_ro(tw + m2 * b));
            E[m2 * PITCH_CD + b] = t;
        }
    }
    // E -> P.  lane a < R2, result v[i] = x[R2*i + a]
    static B200_HD void stepD(float2 (&v)[R1], int a, const float2* E) {
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) v[k1] = E[a * PITCH_CD + k1];
        RegFFT<R1, +1>::run(v);
    }
};

}  // namespace b200cam
