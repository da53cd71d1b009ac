// b200cam: fused N=256 sensor kernels - a plane's spectrum never leaves the SM.
//
// Replaces, for 256x256 planes, the row / column / row kernel sequence of kernels.cuh (conv2D,
// Face-DeId/Camera/Utils.py:7-12, and the per-image max of Optics.py:128) by ONE persistent kernel per
// direction.  One CTA (17 warps) owns one image plane at a time:
//
//   a 256x256 real plane = two 128x256 sub-planes E (even rows) and O (odd rows).  The half spectrum of
//   a sub-plane is 128 x 129 complex = 129 KB and lives in shared memory ("S", pitch 129 so rows AND
//   columns are bank-conflict free).  The plane's spectrum is the radix-2 combination
//        X[v]     = E^[v] + w256^v O^[v]          X[v+128] = E^[v] - w256^v O^[v]        (v < 128)
//   so E^ is produced first and PARKED IN TENSOR MEMORY (tcgen05.st, 129 KB of the SM's 256 KB TMEM:
//   every thread re-reads exactly the values it wrote, which is what TMEM's lane-private addressing
//   wants), then O^ is produced in the same shared memory, combined with E^ in registers, multiplied
//   by the OTF, split again for the inverse (S' = Y[v]+Y[v+128], D' = (Y[v]-Y[v+128]) conj(w^v)), and
//   the two sub-planes are inverse transformed one after the other (D' waits in TMEM meanwhile).
//
// All FFTs are 128-point complex transforms over 8 lanes x 16 registers (radix 16 x 8), packed fp32
// arithmetic (pkfft.cuh), twiddles from a 2 KB shared table, exchange IN PLACE in the row / column of S
// that the transform owns (XOR-swizzled so that every 64-bit shared access is conflict free).
// Real rows use the "256 reals = 128 complex" trick, so global loads / stores are 8-byte vectors.
//
// HBM traffic per plane: read x once, write conv once (+ write the spectrum once when the backward
// will need it); the OTF (792 KB for 3 channels) stays in L2.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

#include "pkfft.cuh"

namespace b200cam {
namespace f256 {

using pk::c32;

constexpr int N = 256;
constexpr int M = 128;                 // rows of a sub-plane
constexpr int NC = 129;                // spectral columns u = 0..128 (global layouts)
constexpr int THREADS = 512;           // 16 warps = 64 groups of 8 lanes
constexpr int PITCH = 129;             // float2 pitch of S (odd: rows and columns conflict free)
constexpr int S_ELEMS = M * PITCH;     // 16512 float2; on chip columns 0 and 128 share slot 0 (re = DC, im = Nyquist)
// per-lane twiddle rows, read as 128-bit words (the four groups of a warp read the same addresses: broadcast);
// row strides 208 B / 144 B put the eight lanes of a group on distinct 16-byte bank groups
constexpr int TA_STRIDE = 26;          // row a: [k1 < 16] w128^(a k1), [16 + aa, aa < 8] w128^(aa (a+8))
constexpr int TA_OFF = S_ELEMS;
constexpr int TQ_STRIDE = 18;          // row b: [m < 16] w256^(b + 8m)
constexpr int TQ_OFF = TA_OFF + 8 * TA_STRIDE;
constexpr int RED_OFF = TQ_OFF + 8 * TQ_STRIDE;   // 32 floats of reduction scratch
constexpr int TM_OFF = RED_OFF + 16;   // TMEM base address
constexpr int SMEM_FLOAT2 = TM_OFF + 2;
constexpr int SMEM_BYTES = SMEM_FLOAT2 * 8;      // 133,808 B: with the 1 KB system reserve still inside a 132 KB carve-out
constexpr int TMEM_COLS = 256;         // 16 warps x 2 rounds x 32 words / 4 warps per lane quadrant
constexpr int SPEC_PLANE = NC * N;     // float2 elements of one saved spectrum plane [u][v]

// ---- thread geometry ---------------------------------------------------------------------------
// group g = tid/8 (8 lanes per FFT), lane b = tid%8.  Groups 2i and 2i+1 share a half warp; their
// rows / columns differ by 8 so that the two groups fall on disjoint banks.
__device__ __forceinline__ int slotmap(int g) { return 16 * (g >> 4) + ((g >> 1) & 7) + 8 * (g & 1); }
__device__ __forceinline__ int swz(int k1, int p) { return ((k1 & 7) >> 1) | (p << 2); }
__device__ __forceinline__ void pf_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(ptr)); }

// ---- tensor memory ------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(smem_slot))),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "r"(TMEM_COLS) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// 16 complex values (32 words) of this thread's TMEM lane, columns [col, col+32)
__device__ __forceinline__ void tmem_st(uint32_t taddr, const c32 (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
        "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y),
        "f"(v[4].x), "f"(v[4].y), "f"(v[5].x), "f"(v[5].y), "f"(v[6].x), "f"(v[6].y), "f"(v[7].x), "f"(v[7].y),
        "f"(v[8].x), "f"(v[8].y), "f"(v[9].x), "f"(v[9].y), "f"(v[10].x), "f"(v[10].y), "f"(v[11].x), "f"(v[11].y),
        "f"(v[12].x), "f"(v[12].y), "f"(v[13].x), "f"(v[13].y), "f"(v[14].x), "f"(v[14].y), "f"(v[15].x), "f"(v[15].y)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_ld(uint32_t taddr, c32 (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y),
          "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y), "=f"(v[7].x), "=f"(v[7].y),
          "=f"(v[8].x), "=f"(v[8].y), "=f"(v[9].x), "=f"(v[9].y), "=f"(v[10].x), "=f"(v[10].y), "=f"(v[11].x),
          "=f"(v[11].y), "=f"(v[12].x), "=f"(v[12].y), "=f"(v[13].x), "=f"(v[13].y), "=f"(v[14].x), "=f"(v[14].y),
          "=f"(v[15].x), "=f"(v[15].y)
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// ---- 128-point FFT over the 8 lanes of a group ------------------------------------------------------
// "Line" L = the 128 float2 slots, stride ST, that this transform owns in S.
//   P distribution: lane a holds x[8*i + a], i < 16          (space side)
//   Q distribution: lane b holds X[b + 8*m], m < 16          (frequency side)
// Element (k1, a) of the exchange sits at slot 8*k1 + (a ^ swz(k1, p)).
// Callers guarantee (syncwarp) that no lane of the group still reads L when the exchange is written.
// ta = this lane's row of the TA twiddle table.
template <int ST>
__device__ __forceinline__ void fwd128(c32 (&v)[16], c32* L, int a, int p, const float4* ta) {
    pk::Fft<16, -1>::run(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; k1 += 2) {
        const float4 w = ta[k1 >> 1];
        const c32 t0 = (k1 == 0) ? v[0] : pk::mul(v[k1], pk::mk(w.x, w.y));
        const c32 t1 = pk::mul(v[k1 + 1], pk::mk(w.z, w.w));
        L[(8 * k1 + (a ^ swz(k1, p))) * ST] = t0;
        L[(8 * (k1 + 1) + (a ^ swz(k1 + 1, p))) * ST] = t1;
    }
    __syncwarp();
    c32 w0[8], w1[8];
    const int f = swz(a, p);
#pragma unroll
    for (int aa = 0; aa < 8; ++aa) {
        w0[aa] = L[(8 * a + (aa ^ f)) * ST];
        w1[aa] = L[(8 * (a + 8) + (aa ^ f)) * ST];
    }
    pk::Fft<8, -1>::run(w0);
    pk::Fft<8, -1>::run(w1);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
        v[2 * k2] = w0[k2];          // X[a + 16*k2]
        v[2 * k2 + 1] = w1[k2];      // X[a + 8 + 16*k2]
    }
}

// unnormalised inverse: Q distribution in, P distribution out
template <int ST>
__device__ __forceinline__ void inv128(c32 (&v)[16], c32* L, int b, int p, const float4* ta) {
    c32 w0[8], w1[8];
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
        w0[k2] = v[2 * k2];
        w1[k2] = v[2 * k2 + 1];
    }
    pk::Fft<8, +1>::run(w0);
    pk::Fft<8, +1>::run(w1);
    const int f = swz(b, p);
#pragma unroll
    for (int aa = 0; aa < 8; aa += 2) {
        const float4 wa = ta[aa >> 1];            // w128^(aa b), w128^((aa+1) b)
        const float4 wc = ta[8 + (aa >> 1)];      // w128^(aa (b+8)), w128^((aa+1) (b+8))
        const c32 t0 = (aa == 0) ? w0[0] : pk::mulc(w0[aa], pk::mk(wa.x, wa.y));
        const c32 t1 = (aa == 0) ? w1[0] : pk::mulc(w1[aa], pk::mk(wc.x, wc.y));
        L[(8 * b + (aa ^ f)) * ST] = t0;
        L[(8 * (b + 8) + (aa ^ f)) * ST] = t1;
        L[(8 * b + ((aa + 1) ^ f)) * ST] = pk::mulc(w0[aa + 1], pk::mk(wa.z, wa.w));
        L[(8 * (b + 8) + ((aa + 1) ^ f)) * ST] = pk::mulc(w1[aa + 1], pk::mk(wc.z, wc.w));
    }
    __syncwarp();
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) v[k1] = L[(8 * k1 + (b ^ swz(k1, p))) * ST];
    pk::Fft<16, +1>::run(v);
}

// Lane b holds Z[b + 8m]; returns Z[(128 - b - 8m) mod 128] for every m.  The partner lives in lane (8-b)%8
// at register 15-m (lane 0 is its own partner, at register (16-m)%16).
__device__ __forceinline__ void mirror(const c32 (&v)[16], c32 (&q)[16], int b) {
    const int src = (8 - b) & 7;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const c32 give = (b == 0) ? v[(16 - m) & 15] : v[15 - m];
        q[m].x = __shfl_sync(0xffffffffu, give.x, src, 8);
        q[m].y = __shfl_sync(0xffffffffu, give.y, src, 8);
    }
}

// ---- row passes -------------------------------------------------------------------------------------
__device__ __forceinline__ const float* row_ptr(const float* plane, int q, int grp, int r) {
    return plane + static_cast<size_t>(2 * (64 * r + slotmap(grp)) + q) * N;
}
// each lane pulls 128 B of its group's next row into L1 ahead of use
__device__ __forceinline__ void rows_prefetch(const float* plane, int q, int grp, int a, int r) {
    pf_l1(row_ptr(plane, q, grp, r) + 32 * a);
}

// forward: real rows y = 2j+q of `plane` -> S[j][u] = 2 * rfft(row)[u]; slot 0 = (DC, Nyquist)
//   z[n] = x[2n] + i x[2n+1];  Z = FFT128(z);  X[u] = (Z[u] + conj Z[-u]) - i w256^u (Z[u] - conj Z[-u])
__device__ __forceinline__ void rows_fwd(const float* __restrict__ plane, int q, c32* S, const float4* ta,
                                         const float4* tq, int grp, int a) {
    const int p = grp & 1;
    rows_prefetch(plane, q, grp, a, 1);
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        const int j = 64 * r + slotmap(grp);
        const float2* src = reinterpret_cast<const float2*>(row_ptr(plane, q, grp, r));
        c32* L = S + j * PITCH;
        c32 v[16], zq[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __ldg(src + 8 * i + a);
        fwd128<1>(v, L, a, p, ta);
        mirror(v, zq, a);
        __syncwarp();                                                    // exchange reads done before L is rewritten
#pragma unroll
        for (int m = 0; m < 16; m += 2) {
            const float4 w = tq[m >> 1];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const c32 P = v[m + h], Z2 = zq[m + h];
                const c32 wu = h == 0 ? pk::mk(w.x, w.y) : pk::mk(w.z, w.w);
                const c32 e = __ffma2_rn(Z2, pk::mk(1.f, -1.f), P);      // P + conj(Z2)
                const c32 d = __ffma2_rn(Z2, pk::mk(-1.f, 1.f), P);      // P - conj(Z2)
                c32 x = pk::sub_i(e, pk::mul(d, wu));                    // e - i w^u d
                if (m + h == 0 && a == 0) x = pk::mk(x.x, e.x - d.y);    // slot 0 = (X[0], X[128]): both real
                L[a + 8 * (m + h)] = x;
            }
        }
    }
}

// inverse: S[j][u] (Hermitian half rows) -> real rows y = 2j+q of `out`; returns the running max
//   Z[u] = (X[u] + conj X[128-u]) + i conj(w256^u) (X[u] - conj X[128-u]);  z = IFFT128(Z)
__device__ __forceinline__ float rows_inv(float* __restrict__ out, int q, c32* S, const float4* ta, const float4* tq,
                                          int grp, int b, float mx) {
    const int p = grp & 1;
#pragma unroll 1
    for (int r = 0; r < 2; ++r) {
        const int j = 64 * r + slotmap(grp);
        c32* L = S + j * PITCH;
        c32 x[16], xq[16], v[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) x[m] = L[b + 8 * m];
        mirror(x, xq, b);
        if (b == 0) {                                                    // slot 0 = (DC, Nyquist); Im ignored (irfft)
            xq[0] = pk::mk(x[0].y, 0.f);
            x[0].y = 0.f;
        }
#pragma unroll
        for (int m = 0; m < 16; m += 2) {
            const float4 w = tq[m >> 1];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const c32 A = x[m + h], Bv = xq[m + h];
                const c32 wu = h == 0 ? pk::mk(w.x, w.y) : pk::mk(w.z, w.w);
                const c32 e = __ffma2_rn(Bv, pk::mk(1.f, -1.f), A);      // A + conj(B)
                const c32 d = __ffma2_rn(Bv, pk::mk(-1.f, 1.f), A);      // A - conj(B)
                v[m + h] = pk::add_i(e, pk::mulc(d, wu));                // e + i conj(w^u) d
            }
        }
        __syncwarp();                                                    // all lanes have read the row
        inv128<1>(v, L, b, p, ta);
        float2* dst = reinterpret_cast<float2*>(out + static_cast<size_t>(2 * j + q) * N);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            dst[8 * i + b] = v[i];
            mx = fmaxf(mx, fmaxf(v[i].x, v[i].y));
        }
    }
    return mx;
}

// ---- column helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ int column_of(int grp, int round) { return 64 * round + slotmap(grp); }
__device__ __forceinline__ uint32_t tmem_slot(uint32_t base, int warp, int round) {
    return base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(((warp >> 2) * 2 + round) * 32);
}

__device__ __forceinline__ void col_load(c32 (&v)[16], const c32* S, int u, int a) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = S[(8 * i + a) * PITCH + u];
    __syncwarp();
}
__device__ __forceinline__ void col_store(const c32 (&v)[16], c32* S, int u, int a) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) S[(8 * i + a) * PITCH + u] = v[i];
}

// forward column transforms of the sub-plane in S; results parked in tensor memory
__device__ __forceinline__ void cols_fwd_park(c32* S, const float4* ta, uint32_t tbase, int warp, int grp, int a) {
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        const int u = column_of(grp, round);
        c32 v[16];
        col_load(v, S, u, a);
        fwd128<PITCH>(v, S + u, a, grp & 1, ta);
        tmem_st(tmem_slot(tbase, warp, round), v);
    }
}

// inverse column transforms of the values parked in tensor memory -> S
__device__ __forceinline__ void cols_inv_unpark(c32* S, const float4* ta, uint32_t tbase, int warp, int grp, int b) {
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        const int u = column_of(grp, round);
        c32 v[16];
        tmem_ld(tmem_slot(tbase, warp, round), v);
        inv128<PITCH>(v, S + u, b, grp & 1, ta);
        col_store(v, S, u, b);
    }
}

// The packed column (slot 0) holds C[v] = A[v] + i B[v] with A = column u=0 and B = column u=128, both Hermitian
// in v.  Un-packing needs C[(128 - v) mod 128], which lives in another lane: group 0 spills its two register
// columns to S (its own column 0 and the otherwise unused slot 128 of every row) and reads the partners back.
__device__ __forceinline__ void packed_publish(c32* S, const c32 (&vo)[16], const c32 (&ve)[16], int b) {
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        S[(b + 8 * m) * PITCH] = vo[m];
        S[(b + 8 * m) * PITCH + 128] = ve[m];
    }
    __syncwarp(0xffu);
}
// A = (C + conj Cp)/2, B = -i (C - conj Cp)/2
__device__ __forceinline__ void unpack2(c32 C, c32 Cp, c32& A, c32& Bq) {
    A = pk::mk(0.5f * (C.x + Cp.x), 0.5f * (C.y - Cp.y));
    Bq = pk::mk(0.5f * (C.y + Cp.y), -0.5f * (C.x - Cp.x));
}

// =====================================================================================================
//  forward
// =====================================================================================================
struct FwdParams {
    const float* img;       // [planes][256][256]
    float* out;             // [planes][256][256]  un-normalised convolution; nullptr: only X is produced
    const float2* K;        // [3][129][256]  (-1)^(u+v) rfft2(psf)[v][u] / (2 N^2), column major
    float2* X;              // nullable: [planes][129][256]  2 * rfft2(img)[v][u], column major
    const float2* tw256;    // exp(-2 pi i j / 256)
    float* plane_max;       // [planes]
    int* tie_count;         // [planes/3], zeroed here for the normalise kernel that follows
    int planes;
};

// one spectral pair (v, v+128) of one column: E^/O^ -> X -> x OTF -> (S', D')
__device__ __forceinline__ void pw_fwd(c32 eE, c32 eO, c32 twv, const float2* __restrict__ kcol, float2* __restrict__ xcol,
                                       int off, c32& Sp, c32& Dp) {
    const c32 wo = pk::mul(eO, twv);
    const c32 x0 = pk::add(eE, wo);
    const c32 x1 = pk::sub(eE, wo);
    const c32 k0 = __ldg(kcol + off);
    const c32 k1 = __ldg(kcol + off + 128);
    if (xcol != nullptr) {
        xcol[off] = x0;
        xcol[off + 128] = x1;
    }
    const c32 y0 = pk::mul(x0, k0);
    const c32 y1 = pk::mul(x1, k1);
    Sp = pk::add(y0, y1);
    Dp = pk::mulc(pk::sub(y0, y1), twv);
}

// columns of O^ (in S) + parked E^ -> spectrum, x OTF, split; S' inverse transformed into S, D' parked
__device__ __forceinline__ void cols_pointwise_fwd(c32* S, const float4* ta, const float4* tq, uint32_t tbase, int warp,
                                                   int grp, int b, const float2* __restrict__ Kc, float2* __restrict__ Xp) {
    const c32* tqs = reinterpret_cast<const c32*>(tq);
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        const int u = column_of(grp, round);
        const float2* kcol = Kc + u * N + b;
        float2* xcol = Xp != nullptr ? Xp + u * N + b : nullptr;
        // this lane's share of the OTF column(s): 2 KB per column = 16 lines, two per lane
        pf_l1(Kc + u * N + 16 * b);
        pf_l1(Kc + u * N + 128 + 16 * b);
        if (u == 0) { pf_l1(Kc + 128 * N + 16 * b); pf_l1(Kc + 128 * N + 128 + 16 * b); }
        c32 v[16], e[16];
        col_load(v, S, u, b);
        fwd128<PITCH>(v, S + u, b, grp & 1, ta);
        tmem_ld(tmem_slot(tbase, warp, round), e);
        if (u != 0) {
#pragma unroll
            for (int m = 0; m < 16; ++m) pw_fwd(e[m], v[m], tqs[m], kcol, xcol, 8 * m, v[m], e[m]);
        } else {
            packed_publish(S, v, e, b);
            const float2* kcol2 = kcol + 128 * N;
            float2* xcol2 = xcol != nullptr ? xcol + 128 * N : nullptr;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                c32 aE, bE, aO, bO, sA, dA, sB, dB;
                const int pidx = ((128 - b - 8 * m) & 127) * PITCH;
                unpack2(e[m], S[pidx + 128], aE, bE);
                unpack2(v[m], S[pidx], aO, bO);
                const c32 twv = tqs[m];
                pw_fwd(aE, aO, twv, kcol, xcol, 8 * m, sA, dA);
                pw_fwd(bE, bO, twv, kcol2, xcol2, 8 * m, sB, dB);
                v[m] = pk::add_i(sA, sB);       // S'_A + i S'_B
                e[m] = pk::add_i(dA, dB);
            }
        }
        tmem_st(tmem_slot(tbase, warp, round), e);
        __syncwarp();
        inv128<PITCH>(v, S + u, b, grp & 1, ta);
        col_store(v, S, u, b);
    }
}

__device__ __forceinline__ float block_max(float mx, float* red, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red[tid >> 5] = mx;
    __syncthreads();
    float r = red[0];
    for (int w = 1; w < THREADS / 32; ++w) r = fmaxf(r, red[w]);
    return r;
}

__device__ __forceinline__ void prefetch_plane_l2(const float* plane, int tid) {
    if (tid < 32)      // 32 x 8 KB = one 256 KB plane
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(plane + tid * 2048), "r"(8192) : "memory");
}

__device__ __forceinline__ uint32_t kernel_prologue(float2* smem, const float2* tw256, int tid, int warp) {
    uint32_t* tm = reinterpret_cast<uint32_t*>(smem + TM_OFF);
    for (int i = tid; i < 8 * 24; i += THREADS) {
        const int a = i / 24, e = i % 24;
        const int idx = e < 16 ? a * e : (e - 16) * (a + 8);            // power of w128
        smem[TA_OFF + a * TA_STRIDE + e] = tw256[(2 * idx) & 255];
    }
    for (int i = tid; i < 8 * 16; i += THREADS) smem[TQ_OFF + (i >> 4) * TQ_STRIDE + (i & 15)] = tw256[(i >> 4) + 8 * (i & 15)];
    if (warp == 0) tmem_alloc(tm);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    return tm[0];
}

__device__ __forceinline__ void fwd_kernel_body(const FwdParams& p, float2* smem) {
    c32* S = smem;
    const float4* ta = reinterpret_cast<const float4*>(smem + TA_OFF + TA_STRIDE * (threadIdx.x & 7));
    const float4* tq = reinterpret_cast<const float4*>(smem + TQ_OFF + TQ_STRIDE * (threadIdx.x & 7));
    float* red = reinterpret_cast<float*>(smem + RED_OFF);
    const int tid = threadIdx.x, warp = tid >> 5, grp = tid >> 3, b = tid & 7;
    if (static_cast<int>(blockIdx.x) < p.planes) prefetch_plane_l2(p.img + static_cast<size_t>(blockIdx.x) * N * N, tid);
    const uint32_t tbase = kernel_prologue(smem, p.tw256, tid, warp);

    for (int plane = blockIdx.x; plane < p.planes; plane += gridDim.x) {
        const float* xp = p.img + static_cast<size_t>(plane) * N * N;
        float* op = p.out + static_cast<size_t>(plane) * N * N;
        const int ch = plane % 3;
        if (plane + static_cast<int>(gridDim.x) < p.planes)
            prefetch_plane_l2(p.img + static_cast<size_t>(plane + gridDim.x) * N * N, tid);
        if (tid == 0 && ch == 0 && p.tie_count != nullptr) p.tie_count[plane / 3] = 0;
        float mx = __int_as_float(0xff800000);

        rows_fwd(xp, 0, S, ta, tq, grp, b);
        __syncthreads();
        rows_prefetch(xp, 1, grp, b, 0);
        cols_fwd_park(S, ta, tbase, warp, grp, b);
        __syncthreads();
        rows_fwd(xp, 1, S, ta, tq, grp, b);
        __syncthreads();
        cols_pointwise_fwd(S, ta, tq, tbase, warp, grp, b, p.K + static_cast<size_t>(ch) * SPEC_PLANE,
                           p.X != nullptr ? p.X + static_cast<size_t>(plane) * SPEC_PLANE : nullptr);
        __syncthreads();
        if (p.out == nullptr) continue;          // spectrum-only run (backward without a saved spectrum)
        mx = rows_inv(op, 0, S, ta, tq, grp, b, mx);
        __syncthreads();
        cols_inv_unpark(S, ta, tbase, warp, grp, b);
        __syncthreads();
        mx = rows_inv(op, 1, S, ta, tq, grp, b, mx);
        const float m = block_max(mx, red, tid);
        if (tid == 0) p.plane_max[plane] = m;
        __syncthreads();
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase);
}

// =====================================================================================================
//  normalise: y = conv / max over the image's three planes (Optics.py:128) + arg-max positions
// =====================================================================================================
struct NormParams {
    float* y;                 // [B][3][256][256] in place
    const float* plane_max;   // [3B]
    float* img_max;           // [B] out
    int* tie_count;           // [B] (zero on entry)
    int* tie_pos;             // [B][max_ties]
    long long n4;             // float4 elements in total
    int max_ties;
};

__device__ __forceinline__ void norm_kernel_body(const NormParams& p) {
    constexpr int PER_IMAGE4 = 3 * N * N / 4;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.n4; i += stride) {
        const int img = static_cast<int>(i / PER_IMAGE4);
        const int r = static_cast<int>(i % PER_IMAGE4);
        const float m = fmaxf(fmaxf(__ldg(p.plane_max + 3 * img), __ldg(p.plane_max + 3 * img + 1)),
                              __ldg(p.plane_max + 3 * img + 2));
        if (r == 0) p.img_max[img] = m;
        float4 v = reinterpret_cast<float4*>(p.y)[i];
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
            if (e[qq] == m) {
                const int slot = atomicAdd(p.tie_count + img, 1);
                if (slot < p.max_ties) p.tie_pos[img * p.max_ties + slot] = r * 4 + qq;
            }
        }
        v.x = e[0] / m; v.y = e[1] / m; v.z = e[2] / m; v.w = e[3] / m;
        reinterpret_cast<float4*>(p.y)[i] = v;
    }
}

// =====================================================================================================
//  backward: acc[c][u][v] = sum_b conj(X_b) G_b / max_b  (closed form of autograd through conv2D wrt the
//  kernel) and the per-image dot products of the amax term, in the frequency domain.
//  CTA k owns channel k % 3 and a private accumulator plane (global, L2 resident, read-modify-write by
//  the thread that owns the element - no atomics, fixed order).
// =====================================================================================================
struct BwdParams {
    const float* g;          // [planes][256][256]   dL/dsensor
    const float2* X;         // [planes][129][256]   saved by the forward
    const float2* K;         // [3][129][256]
    const float2* tw256;
    const float* img_max;    // [B]
    float2* acc;             // [grid][129][256]
    float* sdot;             // [planes]   sum_f wt_u Re(K~ conj(t)),  t = conj(X~) G~  ->  sum(g*conv) = sdot / 2
    int B;
};

// one spectral pair of one column: G -> t = conj(X) G; acc += t / max; s2 += K (.) t  (element-wise)
__device__ __forceinline__ void pw_bwd(c32 eE, c32 eO, c32 twv, const float2* __restrict__ kcol,
                                       const float2* __restrict__ xcol, float2* __restrict__ acol, int off, float inv_m,
                                       bool first, c32& s2) {
    const c32 wo = pk::mul(eO, twv);
    const c32 g0 = pk::add(eE, wo);
    const c32 g1 = pk::sub(eE, wo);
    const c32 x0 = __ldg(xcol + off), x1 = __ldg(xcol + off + 128);
    const c32 k0 = __ldg(kcol + off), k1 = __ldg(kcol + off + 128);
    c32 a0 = pk::mk(0.f, 0.f), a1 = a0;
    if (!first) { a0 = acol[off]; a1 = acol[off + 128]; }
    const c32 t0 = pk::mulc(g0, x0);
    const c32 t1 = pk::mulc(g1, x1);
    s2 = __ffma2_rn(k0, t0, s2);
    s2 = __ffma2_rn(k1, t1, s2);
    acol[off] = __ffma2_rn(t0, pk::mk(inv_m, inv_m), a0);
    acol[off + 128] = __ffma2_rn(t1, pk::mk(inv_m, inv_m), a1);
}

__device__ __forceinline__ float cols_pointwise_bwd(c32* S, const float4* ta, const float4* tq, uint32_t tbase, int warp,
                                                    int grp, int b, const float2* __restrict__ Kc,
                                                    const float2* __restrict__ Xp, float2* __restrict__ acc, float inv_m,
                                                    bool first) {
    const c32* tqs = reinterpret_cast<const c32*>(tq);
    float sd = 0.f;
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        const int u = column_of(grp, round);
        const float2* kcol = Kc + u * N + b;
        const float2* xcol = Xp + u * N + b;
        float2* acol = acc + u * N + b;
        pf_l1(Xp + u * N + 16 * b);
        pf_l1(Xp + u * N + 128 + 16 * b);
        pf_l1(Kc + u * N + 16 * b);
        pf_l1(Kc + u * N + 128 + 16 * b);
        c32 v[16], e[16];
        col_load(v, S, u, b);
        fwd128<PITCH>(v, S + u, b, grp & 1, ta);
        tmem_ld(tmem_slot(tbase, warp, round), e);
        c32 s2 = pk::mk(0.f, 0.f);
        if (u != 0) {
#pragma unroll
            for (int m = 0; m < 16; ++m)
                pw_bwd(e[m], v[m], tqs[m], kcol, xcol, acol, 8 * m, inv_m, first, s2);
            sd += 2.0f * (s2.x + s2.y);
        } else {
            packed_publish(S, v, e, b);
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                c32 aE, bE, aO, bO;
                const int pidx = ((128 - b - 8 * m) & 127) * PITCH;
                unpack2(e[m], S[pidx + 128], aE, bE);
                unpack2(v[m], S[pidx], aO, bO);
                const c32 twv = tqs[m];
                pw_bwd(aE, aO, twv, kcol, xcol, acol, 8 * m, inv_m, first, s2);
                pw_bwd(bE, bO, twv, kcol + 128 * N, xcol + 128 * N, acol + 128 * N, 8 * m, inv_m, first, s2);
            }
            sd += s2.x + s2.y;          // columns 0 and 128 count once in the Hermitian sum
        }
        __syncwarp();
    }
    return sd;
}

__device__ __forceinline__ void bwd_kernel_body(const BwdParams& p, float2* smem) {
    c32* S = smem;
    const float4* ta = reinterpret_cast<const float4*>(smem + TA_OFF + TA_STRIDE * (threadIdx.x & 7));
    const float4* tq = reinterpret_cast<const float4*>(smem + TQ_OFF + TQ_STRIDE * (threadIdx.x & 7));
    float* red = reinterpret_cast<float*>(smem + RED_OFF);
    const int tid = threadIdx.x, warp = tid >> 5, grp = tid >> 3, b = tid & 7;
    const int ch = blockIdx.x % 3;
    const int per_ch = (static_cast<int>(gridDim.x) - ch + 2) / 3;      // CTAs that own this channel
    if (static_cast<int>(blockIdx.x) / 3 < p.B)
        prefetch_plane_l2(p.g + static_cast<size_t>((blockIdx.x / 3) * 3 + ch) * N * N, tid);
    const uint32_t tbase = kernel_prologue(smem, p.tw256, tid, warp);

    float2* acc = p.acc + static_cast<size_t>(blockIdx.x) * SPEC_PLANE;
    bool first = true;
    for (int img = blockIdx.x / 3; img < p.B; img += per_ch) {
        const int plane = img * 3 + ch;
        const float* gp = p.g + static_cast<size_t>(plane) * N * N;
        if (img + per_ch < p.B) prefetch_plane_l2(p.g + static_cast<size_t>((img + per_ch) * 3 + ch) * N * N, tid);
        rows_fwd(gp, 0, S, ta, tq, grp, b);
        __syncthreads();
        rows_prefetch(gp, 1, grp, b, 0);
        cols_fwd_park(S, ta, tbase, warp, grp, b);
        __syncthreads();
        rows_fwd(gp, 1, S, ta, tq, grp, b);
        __syncthreads();
        const float inv_m = 1.0f / __ldg(p.img_max + img);
        float s = cols_pointwise_bwd(S, ta, tq, tbase, warp, grp, b, p.K + static_cast<size_t>(ch) * SPEC_PLANE,
                                     p.X + static_cast<size_t>(plane) * SPEC_PLANE, acc, inv_m, first);
        first = false;
        // deterministic block sum
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((tid & 31) == 0) red[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            float tot = 0.f;
            for (int w = 0; w < THREADS / 32; ++w) tot += red[w];
            p.sdot[plane] = tot;
        }
        __syncthreads();
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase);
}

}  // namespace f256
}  // namespace b200cam
