// b200cam: small FFTs held entirely in registers (radix 2/4/8/16/32), natural order in and out.
//
// RegFFT<R, DIR>::run(v) transforms float2 v[R] in place; DIR = -1 is the forward transform
// (kernel exp(-2*pi*i*n*k/R)), DIR = +1 the unnormalised inverse.  Everything is fully
// unrolled with compile-time indices so `v` lives in registers and the twiddle constants
// (angles that are multiples of 2*pi/32) fold into immediates.
#pragma once

#include "compat.cuh"
#if defined(__CUDACC__)
#include "pkfft.cuh"
#endif

namespace b200cam {

// cos(2*pi*j/32)
B200_HD constexpr float cos32(int j) {
    switch (j & 31) {
        case 0: return 1.0f;
        case 1: case 31: return 0.98078528040323044913f;
        case 2: case 30: return 0.92387953251128675613f;
        case 3: case 29: return 0.83146961230254523708f;
        case 4: case 28: return 0.70710678118654752440f;
        case 5: case 27: return 0.55557023301960222474f;
        case 6: case 26: return 0.38268343236508977173f;
        case 7: case 25: return 0.19509032201612826785f;
        case 8: case 24: return 0.0f;
        case 9: case 23: return -0.19509032201612826785f;
        case 10: case 22: return -0.38268343236508977173f;
        case 11: case 21: return -0.55557023301960222474f;
        case 12: case 20: return -0.70710678118654752440f;
        case 13: case 19: return -0.83146961230254523708f;
        case 14: case 18: return -0.92387953251128675613f;
        case 15: case 17: return -0.98078528040323044913f;
        default: return -1.0f;  // 16
    }
}
B200_HD constexpr float sin32(int j) { return cos32(j - 8); }

// a * exp(DIR * 2*pi*i * idx/32)
template <int DIR>
B200_HD float2 mul_w32(float2 a, int idx) {
    idx &= 31;
    if (idx == 0) return a;
    if (idx == 16) return make_float2(-a.x, -a.y);
    if (idx == 8) return DIR > 0 ? cmul_i(a) : cmul_mi(a);
    if (idx == 24) return DIR > 0 ? cmul_mi(a) : cmul_i(a);
    const float c = cos32(idx);
    const float s = DIR > 0 ? sin32(idx) : -sin32(idx);
    return make_float2(a.x * c - a.y * s, a.x * s + a.y * c);
}

template <int R, int DIR>
struct RegFFT;

template <int DIR>
struct RegFFT<1, DIR> {
    static B200_HD void run(float2 (&)[1]) {}
};

template <int DIR>
struct RegFFT<2, DIR> {
    static B200_HD void run(float2 (&v)[2]) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <int DIR>
struct RegFFT<4, DIR> {
    static B200_HD void run(float2 (&v)[4]) {
#if defined(__CUDA_ARCH__)
        pk::Fft<4, DIR>::run(v);         // packed fp32 (FADD2 / FFMA2) on the device
        return;
#endif
        const float2 t0 = cadd(v[0], v[2]);
        const float2 t1 = csub(v[0], v[2]);
        const float2 t2 = cadd(v[1], v[3]);
        const float2 d = csub(v[1], v[3]);
        const float2 t3 = DIR > 0 ? cmul_i(d) : cmul_mi(d);
        v[0] = cadd(t0, t2);
        v[1] = cadd(t1, t3);
        v[2] = csub(t0, t2);
        v[3] = csub(t1, t3);
    }
};

// R = RA * RB (RA = 4): n = RB*na + nb, k = ka + RA*kb
template <int R, int DIR>
struct RegFFT {
    static_assert(R == 8 || R == 16 || R == 32, "supported register radices: 2,4,8,16,32");
    static constexpr int RA = 4;
    static constexpr int RB = R / 4;
    static B200_HD void run(float2 (&v)[R]) {
#if defined(__CUDA_ARCH__)
        pk::Fft<R, DIR>::run(v);         // packed fp32 (FADD2 / FFMA2) on the device
        return;
#endif
#pragma unroll
        for (int nb = 0; nb < RB; ++nb) {
            float2 t[RA];
#pragma unroll
            for (int na = 0; na < RA; ++na) t[na] = v[RB * na + nb];
            RegFFT<RA, DIR>::run(t);
#pragma unroll
            for (int ka = 0; ka < RA; ++ka) v[RB * ka + nb] = mul_w32<DIR>(t[ka], (32 / R) * nb * ka);
        }
        float2 out[R];
#pragma unroll
        for (int ka = 0; ka < RA; ++ka) {
            float2 s[RB];
#pragma unroll
            for (int nb = 0; nb < RB; ++nb) s[nb] = v[RB * ka + nb];
            RegFFT<RB, DIR>::run(s);
#pragma unroll
            for (int kb = 0; kb < RB; ++kb) out[ka + RA * kb] = s[kb];
        }
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = out[i];
    }
};

}  // namespace b200cam
