// b200cam: register FFTs on packed fp32 pairs (sm_100a FADD2 / FMUL2 / FFMA2).
//
// A complex number is one 64-bit register pair; Blackwell's packed fp32 instructions operate on both
// halves at once and take the half-swap and per-half negation as operand modifiers, so
//     a + b, a - b, a +- i*b      : 1 instruction     (2 .. 4 scalar ones)
//     a * w  (general twiddle)    : 2 instructions    (4 scalar ones)
// The fp32 pipes retire the same number of flops per clock either way (measured, tools/microbench/
// fp_rate.cu: 3.8 scalar vs 1.96 packed warp-instructions / clk / SM), but the packed form needs half
// the issue slots, which leaves the other half for the shared-memory / global traffic of the FFT.
// Device-only (the generic kernels in kernels.cuh keep the scalar, host-emulable RegFFT).
#pragma once

#include <cuda_runtime.h>

namespace b200cam {
namespace pk {

typedef float2 c32;

__device__ __forceinline__ c32 mk(float x, float y) { return make_float2(x, y); }
__device__ __forceinline__ c32 add(c32 a, c32 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ c32 sub(c32 a, c32 b) { return __ffma2_rn(b, mk(-1.f, -1.f), a); }
__device__ __forceinline__ c32 swp(c32 a) { return mk(a.y, a.x); }
// t + i*d  and  t - i*d
__device__ __forceinline__ c32 add_i(c32 t, c32 d) { return __ffma2_rn(swp(d), mk(-1.f, 1.f), t); }
__device__ __forceinline__ c32 sub_i(c32 t, c32 d) { return __ffma2_rn(swp(d), mk(1.f, -1.f), t); }
__device__ __forceinline__ c32 scale(c32 a, float s) { return __fmul2_rn(a, mk(s, s)); }
__device__ __forceinline__ c32 conj(c32 a) { return mk(a.x, -a.y); }
// a * w
__device__ __forceinline__ c32 mul(c32 a, c32 w) {
    return __ffma2_rn(swp(a), mk(-w.y, w.y), __fmul2_rn(a, mk(w.x, w.x)));
}
// a * conj(w)
__device__ __forceinline__ c32 mulc(c32 a, c32 w) {
    return __ffma2_rn(swp(a), mk(w.y, -w.y), __fmul2_rn(a, mk(w.x, w.x)));
}
// acc + a * w
__device__ __forceinline__ c32 fma(c32 a, c32 w, c32 acc) {
    return __ffma2_rn(swp(a), mk(-w.y, w.y), __ffma2_rn(a, mk(w.x, w.x), acc));
}
// a * exp(DIR * 2*pi*i * j / 32), j a compile-time constant after unrolling
__device__ __forceinline__ constexpr float c32cos(int j) {
    j &= 31;
    if (j > 16) j = 32 - j;
    return j == 0 ? 1.0f : j == 1 ? 0.98078528040323044913f : j == 2 ? 0.92387953251128675613f :
           j == 3 ? 0.83146961230254523708f : j == 4 ? 0.70710678118654752440f : j == 5 ? 0.55557023301960222474f :
           j == 6 ? 0.38268343236508977173f : j == 7 ? 0.19509032201612826785f : j == 8 ? 0.0f :
           j == 9 ? -0.19509032201612826785f : j == 10 ? -0.38268343236508977173f : j == 11 ? -0.55557023301960222474f :
           j == 12 ? -0.70710678118654752440f : j == 13 ? -0.83146961230254523708f : j == 14 ? -0.92387953251128675613f :
           j == 15 ? -0.98078528040323044913f : -1.0f;
}
__device__ __forceinline__ constexpr float c32sin(int j) { return c32cos(j - 8); }

template <int DIR>
__device__ __forceinline__ c32 mul_w32(c32 a, int j) {
    j &= 31;
    if (j == 0) return a;
    if (j == 16) return mk(-a.x, -a.y);
    if (j == 8) return DIR > 0 ? mk(-a.y, a.x) : mk(a.y, -a.x);
    if (j == 24) return DIR > 0 ? mk(a.y, -a.x) : mk(-a.y, a.x);
    const float c = c32cos(j);
    const float s = DIR > 0 ? c32sin(j) : -c32sin(j);
    return mul(a, mk(c, s));
}

template <int R, int DIR>
struct Fft;

template <int DIR>
struct Fft<2, DIR> {
    static __device__ __forceinline__ void run(c32 (&v)[2]) {
        const c32 a = v[0], b = v[1];
        v[0] = add(a, b);
        v[1] = sub(a, b);
    }
};

template <int DIR>
struct Fft<4, DIR> {
    static __device__ __forceinline__ void run(c32 (&v)[4]) {
        const c32 t0 = add(v[0], v[2]);
        const c32 t1 = sub(v[0], v[2]);
        const c32 t2 = add(v[1], v[3]);
        const c32 d = sub(v[1], v[3]);
        v[0] = add(t0, t2);
        v[2] = sub(t0, t2);
        v[1] = DIR > 0 ? add_i(t1, d) : sub_i(t1, d);
        v[3] = DIR > 0 ? sub_i(t1, d) : add_i(t1, d);
    }
};

// R = 4 * RB: n = RB*na + nb, k = ka + 4*kb; natural order in and out
template <int R, int DIR>
struct Fft {
    static_assert(R == 8 || R == 16 || R == 32, "packed register radices: 2, 4, 8, 16, 32");
    static constexpr int RB = R / 4;
    static __device__ __forceinline__ void run(c32 (&v)[R]) {
#pragma unroll
        for (int nb = 0; nb < RB; ++nb) {
            c32 t[4];
#pragma unroll
            for (int na = 0; na < 4; ++na) t[na] = v[RB * na + nb];
            Fft<4, DIR>::run(t);
#pragma unroll
            for (int ka = 0; ka < 4; ++ka) v[RB * ka + nb] = mul_w32<DIR>(t[ka], (32 / R) * nb * ka);
        }
        c32 out[R];
#pragma unroll
        for (int ka = 0; ka < 4; ++ka) {
            c32 s[RB];
#pragma unroll
            for (int nb = 0; nb < RB; ++nb) s[nb] = v[RB * ka + nb];
            Fft<RB, DIR>::run(s);
#pragma unroll
            for (int kb = 0; kb < RB; ++kb) out[ka + 4 * kb] = s[kb];
        }
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = out[i];
    }
};

}  // namespace pk
}  // namespace b200cam
