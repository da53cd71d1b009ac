#include <cuda_runtime.h>
__global__ void k_probe(float2* p, float2 w) {
    float2 a = p[threadIdx.x], b = p[threadIdx.x + 32];
    float2 s = __fadd2_rn(a, b);
    float2 d = __ffma2_rn(b, make_float2(-1.f, -1.f), a);   // a - b
    float2 dsw = make_float2(d.y, d.x);
    float2 r = __ffma2_rn(dsw, make_float2(1.f, -1.f), s);  // s + (-i)*d  = (s.x + d.y, s.y - d.x)
    // twiddle: r*w = r*(c,c) + r_sw*(-s, s)
    float2 rsw = make_float2(r.y, r.x);
    float2 t = __ffma2_rn(rsw, make_float2(-w.y, w.y), __fmul2_rn(r, make_float2(w.x, w.x)));
    p[threadIdx.x] = t;
}
