// b200cam: fused N=256 sensor kernels - the spectrum of a plane never leaves the SM.
//
// A 256x256 fp32 plane has a 256 KB half-spectrum, more than the 227 KB of shared memory of one
// SM.  The plane is therefore split by one vertical radix-2 decimation-in-frequency step:
//      a[y] = x[y] + x[y+128]                       -> even spectrum rows v = 2v'
//      b[y] = (x[y] - x[y+128]) * w256^y            -> odd  spectrum rows v = 2v'+1
// Both halves are REAL 128x256 images, so each is transformed with the two-rows-per-complex-FFT
// trick and its half spectrum is 128 rows x 128 packed columns (DC and Nyquist columns share one
// complex column: re = X[.,0], im = X[.,128]) = 128 KB, which fits.  Per half ("pass"):
//      R  rows:     64 complex 256-FFTs (row pairs)            -> SP[y'][u]   (shared memory)
//      C  columns: 128 complex 128-FFTs, x OTF, 128 inverse FFTs, in place    (registers + exchange)
//      I  rows:     64 inverse 256-FFTs                         -> e[y][x] (pass 0) / o[y][x] (pass 1)
// and the output is  out[y] = e + o,  out[y+128] = e - o.  The pass-0 result e is "parked"
// (thread-private scratch: global/L2 in this version) until pass 1 produces o in the same registers.
// HBM traffic per plane: read x once (+ once more from L2), write conv once; the forward column
// spectra can be saved (coalesced, in register order) for the backward pass.
//
// Work distribution: persistent CTAs of 512 threads, one plane at a time; the CTA that finishes the
// last of an image's three planes rescales that image by its maximum (Optics.py:128) out of L2.
#pragma once

#include "compat.cuh"
#include "exec.cuh"
#include "fft_plan.cuh"

namespace b200cam {
namespace f256 {

constexpr int N = 256;
constexpr int NH = 128;            // rows of a half image / spectrum rows per pass
constexpr int THREADS = 512;
constexpr int RGROUPS = 32;        // row-FFT groups of 16 lanes
constexpr int CGROUPS = 64;        // column-FFT groups of 8 lanes
constexpr int NCOL = 128;          // packed spectral columns
constexpr int SPITCH = 130;        // float2 pitch of SP rows: 2*130 = 4 (mod 32) words -> conflict-free columns
constexpr int RE_SIZE = 272;       // exchange entries of one 256-point row FFT (16 x 17)
constexpr int CE_SIZE = 144;       // exchange entries of one 128-point column FFT (16 x 9 / 8 x 17)
constexpr int QN = 16;             // complex registers per lane in every distribution

// shared memory map (float2 units)
constexpr int SP_OFF = 0;
constexpr int E_OFF = SP_OFF + NH * SPITCH;                       // 16640
constexpr int E_SIZE = (RGROUPS * RE_SIZE > CGROUPS * CE_SIZE) ? RGROUPS * RE_SIZE : CGROUPS * CE_SIZE;   // 9216
constexpr int TWR_OFF = E_OFF + E_SIZE;                           // [16][16]  w256^(a*k)
constexpr int TWC_OFF = TWR_OFF + 256;                            // [16][8]   w128^(a*k), a<8
constexpr int TWCT_OFF = TWC_OFF + 128;                           // [8][16]   transposed copy
constexpr int TWV_OFF = TWCT_OFF + 128;                           // [128]     w256^y (vertical DIF twiddle)
constexpr int RED_OFF = TWV_OFF + 128;                            // 512 floats
constexpr int FLAG_OFF = RED_OFF + 256;                           // a few ints
constexpr int SMEM_FLOAT2 = FLAG_OFF + 8;
constexpr int SMEM_BYTES = SMEM_FLOAT2 * 8;

// ---- 128-point FFT over 8 lanes: 128 = 16 (registers) x 8 (lanes) ---------------------------------
//   P: lane a<8 holds y' = 8*i + a, i<16.      Q: lane b<8 holds v' = (b + 8*s) + 16*k2, index q = 8*s + k2.
struct Col128 {
    static B200_HD void stepA(float2 (&v)[16], int a, float2* E, const float2* twc) {
        RegFFT<16, -1>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            float2 t = v[k1];
            if (k1 > 0) t = cmul(t, twc[k1 * 8 + a]);
            E[k1 * 9 + a] = t;
        }
    }
    static B200_HD void stepB(float2 (&q)[16], int b, const float2* E) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            float2 t[8];
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) t[n2] = E[(b + 8 * s) * 9 + n2];
            RegFFT<8, -1>::run(t);
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) q[8 * s + k2] = t[k2];
        }
    }
    static B200_HD void stepC(float2 (&q)[16], int b, float2* E, const float2* twct) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            float2 t[8];
#pragma unroll
            for (int k2 = 0; k2 < 8; ++k2) t[k2] = q[8 * s + k2];
            RegFFT<8, +1>::run(t);
#pragma unroll
            for (int m2 = 0; m2 < 8; ++m2) {
                float2 w = t[m2];
                if (m2 > 0) w = cmulc(w, twct[m2 * 16 + b + 8 * s]);
                E[m2 * 17 + b + 8 * s] = w;
            }
        }
    }
    static B200_HD void stepD(float2 (&v)[16], int a, const float2* E) {
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) v[k1] = E[a * 17 + k1];
        RegFFT<16, +1>::run(v);
    }
    // spectral row index v' held in register q of lane b
    static B200_HD int vprime(int b, int q) { return (b + 8 * (q >> 3)) + 16 * (q & 7); }
};

// ---- 256-point row FFT over 16 lanes with the twiddles taken from the shared [16][16] table ----------
struct Row256 {
    static B200_HD void stepA(float2 (&v)[16], int a, float2* E, const float2* twr) {
        RegFFT<16, -1>::run(v);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            float2 t = v[k1];
            if (k1 > 0) t = cmul(t, twr[k1 * 16 + a]);
            E[k1 * 17 + a] = t;
        }
    }
    static B200_HD void stepB(float2 (&v)[16], int b, const float2* E) {
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) v[n2] = E[b * 17 + n2];
        RegFFT<16, -1>::run(v);
    }
    static B200_HD void stepC(float2 (&v)[16], int b, float2* E, const float2* twr) {
        RegFFT<16, +1>::run(v);
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) {
            float2 t = v[m2];
            if (m2 > 0) t = cmulc(t, twr[m2 * 16 + b]);
            E[m2 * 17 + b] = t;
        }
    }
    static B200_HD void stepD(float2 (&v)[16], int a, const float2* E) {
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) v[k1] = E[a * 17 + k1];
        RegFFT<16, +1>::run(v);
    }
};

// ---- layouts of the tables this kernel family reads/writes in HBM -----------------------------------
// fused OTF      kf[c][h][q][u][l]   (complex)  value K[c][v = 2*v'(l,q)+h][u] / N^2, u = 1..127;
//                                    u = 0 slot holds (K0 + K128)/2 of the packed column
//                kq[c][h][q][l]      (complex)  (K0 - K128)/2 of the packed column
// saved spectrum xs[plane][h][q][u][l] (complex) forward column spectra in register order (packed column as is)
B200_HD size_t kf_index(int c, int h, int q, int u, int l) {
    return ((((static_cast<size_t>(c) * 2 + h) * QN + q) * NCOL + u) * 8 + l);
}
B200_HD size_t kq_index(int c, int h, int q, int l) { return (((static_cast<size_t>(c) * 2 + h) * QN + q) * 8 + l); }
constexpr size_t KF_ELEMS = 3 * 2 * QN * NCOL * 8;    // 98304 complex = 768 KB
constexpr size_t KQ_ELEMS = 3 * 2 * QN * 8;
constexpr size_t XS_PLANE = 2 * QN * NCOL * 8;        // 32768 complex = 256 KB per plane

// partner of spectral row v' under conjugation: even rows v -> -v, odd rows 2v'+1 -> -(2v'+1)
B200_HD int partner(int vp, int h) { return h == 0 ? ((NH - vp) & (NH - 1)) : (NH - 1 - vp); }

// ---------------------------------------------------------------------------------------------------
// prep: standard OTF layout otf[c][u<=128][v<256] (b200cam_sensor_fwd's `otf`, already /N^2 and
// sign-twisted) -> fused layout.  grid-stride, 1 thread per (c,h,q,u,l).
// ---------------------------------------------------------------------------------------------------
struct PrepParams {
    const float2* otf;     // [3][129][256]
    float2* kf;            // KF_ELEMS
    float2* kq;            // KQ_ELEMS
    int* done;             // [B] per-image finished-plane counters, reset here
    float* img_max;        // [B] reset to -inf
    int* tie_count;        // [B] reset to 0
    int B;
};

template <class Exec>
B200_HD void prep_body(Exec& ex, const PrepParams& p, int grid_x) {
    ex.phase([&](int tid) {
        const int stride = grid_x * ex.nthreads();
        const int total = static_cast<int>(KF_ELEMS);
        for (int idx = ex.bx() * ex.nthreads() + tid; idx < total; idx += stride) {
            const int l = idx & 7, u = (idx >> 3) & (NCOL - 1), q = (idx >> 10) & (QN - 1), h = (idx >> 14) & 1, c = idx >> 15;
            const int v = 2 * Col128::vprime(l, q) + h;
            if (u > 0) {
                p.kf[idx] = p.otf[(static_cast<size_t>(c) * 129 + u) * N + v];
            } else {
                const float2 k0 = p.otf[(static_cast<size_t>(c) * 129 + 0) * N + v];
                const float2 k128 = p.otf[(static_cast<size_t>(c) * 129 + 128) * N + v];
                p.kf[idx] = make_float2(0.5f * (k0.x + k128.x), 0.5f * (k0.y + k128.y));
                p.kq[kq_index(c, h, q, l)] = make_float2(0.5f * (k0.x - k128.x), 0.5f * (k0.y - k128.y));
            }
        }
        for (int b = ex.bx() * ex.nthreads() + tid; b < p.B; b += stride) {
            p.done[b] = 0;
            p.img_max[b] = neg_inf();
            p.tie_count[b] = 0;
        }
    });
}

// ---------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------
struct FwdParams {
    const float* x;        // [planes][256][256]
    float* y;              // [planes][256][256]  conv, then rescaled in place to the sensor image
    const float2* kf;      // fused OTF
    const float2* kq;
    const float2* tw;      // tw[j] = exp(-2 pi i j / 256)
    float2* xs;            // nullable: saved forward spectra [planes] x XS_PLANE
    float* park;           // [grid][64][512] floats: thread-private parking of the pass-0 image
    float* img_max;        // [B]
    int* done;             // [B]
    int* tie_count;        // [B]
    int* tie_pos;          // [B][MAX_TIES]
    int planes;            // 3*B
    int max_ties;
};

template <class Exec>
B200_HD void load_tables(Exec& ex, const float2* tw, float2* smem) {
    ex.phase([&](int tid) {
        for (int i = tid; i < 256; i += THREADS) {
            const int k = i >> 4, a = i & 15;
            smem[TWR_OFF + i] = tw[(a * k) & 255];
        }
        for (int i = tid; i < 128; i += THREADS) {
            const int k = i >> 3, a = i & 7;
            smem[TWC_OFF + i] = tw[(2 * a * k) & 255];
            const int m2 = i >> 4, k1 = i & 15;
            smem[TWCT_OFF + i] = tw[(2 * m2 * k1) & 255];
            smem[TWV_OFF + i] = tw[i];
        }
    });
}

// per-thread registers that live across (warp-level) phases
struct FState {
    float2 v[16];
};

// R phase of one pass: rows of the half image -> SP.  `sign` = +1 (pass 0: x[y]+x[y+128]) or -1.
// `acc_rows(tid, row, x, top0, top1, bot0, bot1)` sees every loaded element (the backward uses it
// to accumulate sum(g*y)).
template <class Exec, class Acc>
B200_HD void rows_forward(Exec& ex, const float* plane, float sign, float2* smem, FState* st, Acc&& acc_rows) {
    float2* SP = smem + SP_OFF;
    float2* E = smem + E_OFF;
    const float2* twr = smem + TWR_OFF;
    for (int r = 0; r < 2; ++r) {
        ex.warp_phase([&](int tid) {
            const int g = tid >> 4, a = tid & 15, j = g + RGROUPS * r;
            const float* r0 = plane + static_cast<size_t>(2 * j) * N;      // rows 2j, 2j+1 and the same +128
            float2 (&v)[16] = st[ex.slot(tid)].v;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int xx = 16 * i + a;
                const float t0 = ld_ro(r0 + xx), t1 = ld_ro(r0 + N + xx);
                const float b0 = ld_ro(r0 + NH * N + xx), b1 = ld_ro(r0 + NH * N + N + xx);
                acc_rows(tid, 2 * j, xx, t0, t1, b0, b1);
                v[i] = make_float2(t0 + sign * b0, t1 + sign * b1);
            }
            Row256::stepA(v, a, E + g * RE_SIZE, twr);
        });
        ex.warp_phase([&](int tid) {
            const int g = tid >> 4, b = tid & 15;
            Row256::stepB(st[ex.slot(tid)].v, b, E + g * RE_SIZE);
        });
        ex.warp_phase([&](int tid) {      // pair spectrum Z[k], natural order, back into the exchange buffer
            const int g = tid >> 4, b = tid & 15;
            const float2 (&v)[16] = st[ex.slot(tid)].v;
#pragma unroll
            for (int i = 0; i < 16; ++i) E[g * RE_SIZE + b + 16 * i] = v[i];
        });
        ex.warp_phase([&](int tid) {      // un-mix the two real rows (Hermitian symmetry) -> SP rows 2j, 2j+1
            const int g = tid >> 4, a = tid & 15, j = g + RGROUPS * r;
            const float2* Z = E + g * RE_SIZE;
            float2* s0 = SP + (2 * j) * SPITCH;
            float2* s1 = s0 + SPITCH;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int u = a + 16 * i;
                if (u == 0) {
                    const float2 z0 = Z[0], zn = Z[128];
                    s0[0] = make_float2(z0.x, zn.x);       // packed column: (X[.,0], X[.,128]), both real
                    s1[0] = make_float2(z0.y, zn.y);
                } else {
                    const float2 z1 = Z[u], z2 = Z[N - u];
                    s0[u] = make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
                    s1[u] = make_float2(0.5f * (z1.y + z2.y), -0.5f * (z1.x - z2.x));
                }
            }
        });
    }
}

// I phase of one pass: SP -> real rows in registers; emit(tid, r, i, row, x, even_row_value, odd_row_value)
template <class Exec, class Emit>
B200_HD void rows_inverse(Exec& ex, float2* smem, FState* st, Emit&& emit) {
    float2* SP = smem + SP_OFF;
    float2* E = smem + E_OFF;
    const float2* twr = smem + TWR_OFF;
    for (int r = 0; r < 2; ++r) {
        ex.warp_phase([&](int tid) {
            const int g = tid >> 4, b = tid & 15, j = g + RGROUPS * r;
            const float2* s0 = SP + (2 * j) * SPITCH;
            const float2* s1 = s0 + SPITCH;
            float2 (&v)[16] = st[ex.slot(tid)].v;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = b + 16 * i;
                if (k == 0) v[i] = make_float2(s0[0].x, s1[0].x);
                else if (k == NH) v[i] = make_float2(s0[0].y, s1[0].y);
                else if (k < NH) {
                    const float2 e = s0[k], o = s1[k];
                    v[i] = make_float2(e.x - o.y, e.y + o.x);                 // X_even + i X_odd
                } else {
                    const float2 e = s0[N - k], o = s1[N - k];
                    v[i] = make_float2(e.x + o.y, o.x - e.y);                 // conj(X_even) + i conj(X_odd)
                }
            }
            Row256::stepC(v, b, E + g * RE_SIZE, twr);
        });
        ex.warp_phase([&](int tid) {
            const int g = tid >> 4, a = tid & 15, j = g + RGROUPS * r;
            float2 (&v)[16] = st[ex.slot(tid)].v;
            Row256::stepD(v, a, E + g * RE_SIZE);
#pragma unroll
            for (int i = 0; i < 16; ++i) emit(tid, r, i, 2 * j, 16 * i + a, v[i].x, v[i].y);
        });
    }
}

// C phase, forward transform of the columns of SP: afterwards lane b of column group gc holds the
// spectrum of column u in st.v (Q order).  `after(tid, r, u, b)` runs on those registers and must
// leave in st.v what is to be inverse-transformed back into SP (or return false to skip the inverse).
template <class Exec>
B200_HD void cols_forward_round(Exec& ex, int r, int h, float2* smem, FState* st) {
    float2* SP = smem + SP_OFF;
    float2* E = smem + E_OFF;
    const float2* twc = smem + TWC_OFF;
    const float2* twv = smem + TWV_OFF;
    ex.warp_phase([&](int tid) {
        const int gc = tid >> 3, a = tid & 7, u = gc + CGROUPS * r;
        float2 (&v)[16] = st[ex.slot(tid)].v;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int yp = 8 * i + a;
            float2 t = SP[yp * SPITCH + u];
            if (h == 1) t = cmul(t, twv[yp]);          // vertical DIF twiddle of the odd half
            v[i] = t;
        }
        Col128::stepA(v, a, E + gc * CE_SIZE, twc);
    });
    ex.warp_phase([&](int tid) {
        const int gc = tid >> 3, b = tid & 7;
        Col128::stepB(st[ex.slot(tid)].v, b, E + gc * CE_SIZE);
    });
}

template <class Exec>
B200_HD void cols_inverse_round(Exec& ex, int r, int h, float2* smem, FState* st) {
    float2* SP = smem + SP_OFF;
    float2* E = smem + E_OFF;
    const float2* twct = smem + TWCT_OFF;
    const float2* twv = smem + TWV_OFF;
    ex.warp_phase([&](int tid) {
        const int gc = tid >> 3, b = tid & 7;
        Col128::stepC(st[ex.slot(tid)].v, b, E + gc * CE_SIZE, twct);
    });
    ex.warp_phase([&](int tid) {
        const int gc = tid >> 3, a = tid & 7, u = gc + CGROUPS * r;
        float2 (&v)[16] = st[ex.slot(tid)].v;
        Col128::stepD(v, a, E + gc * CE_SIZE);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int yp = 8 * i + a;
            float2 t = v[i];
            if (h == 1) t = cmulc(t, twv[yp]);
            SP[yp * SPITCH + u] = t;
        }
    });
}

// the packed column (u = 0) needs the value at the conjugate-partner row: stage it in natural order
template <class Exec>
B200_HD void stage_packed_column(Exec& ex, int r, float2* smem, FState* st) {
    float2* E = smem + E_OFF;
    ex.warp_phase([&](int tid) {
        const int gc = tid >> 3, b = tid & 7;
        if (gc + CGROUPS * r == 0) {
            const float2 (&q)[16] = st[ex.slot(tid)].v;
#pragma unroll
            for (int i = 0; i < 16; ++i) E[Col128::vprime(b, i)] = q[i];
        }
    });
}

// C phase of the forward kernel: column FFT, optional save, x OTF, inverse column FFT, back into SP.
template <class Exec>
B200_HD void cols_convolve(Exec& ex, const FwdParams& p, int plane, int h, float2* smem, FState* st) {
    const float2* E = smem + E_OFF;
    const int c = plane % 3;
    for (int r = 0; r < 2; ++r) {
        cols_forward_round(ex, r, h, smem, st);
        stage_packed_column(ex, r, smem, st);
        ex.warp_phase([&](int tid) {
            const int gc = tid >> 3, b = tid & 7, u = gc + CGROUPS * r;
            float2 (&q)[16] = st[ex.slot(tid)].v;
            if (p.xs != nullptr) {
                float2* dst = p.xs + (static_cast<size_t>(plane) * 2 + h) * (QN * NCOL * 8);
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[(static_cast<size_t>(i) * NCOL + u) * 8 + b] = q[i];
            }
            if (u == 0) {
                // C'[v] = P[v]*(K0+K128)/2 + conj(P[-v])*(K0-K128)/2
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float2 pc = E[partner(Col128::vprime(b, i), h)];
                    const float2 kp = ld_ro(p.kf + kf_index(c, h, i, 0, b));
                    const float2 kq = ld_ro(p.kq + kq_index(c, h, i, b));
                    q[i] = cadd(cmul(q[i], kp), cmul(cconj(pc), kq));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) q[i] = cmul(q[i], ld_ro(p.kf + kf_index(c, h, i, u, b)));
            }
        });
        cols_inverse_round(ex, r, h, smem, st);
    }
}

template <class Exec>
B200_HD void fwd_body(Exec& ex, const FwdParams& p, float2* smem, int grid_x, FState* st) {
    load_tables(ex, p.tw, smem);
    float* red = reinterpret_cast<float*>(smem + RED_OFF);
    int* flag = reinterpret_cast<int*>(smem + FLAG_OFF);
    float* park = p.park + static_cast<size_t>(ex.bx()) * 64 * THREADS;
    for (int plane = ex.bx(); plane < p.planes; plane += grid_x) {
        const float* xin = p.x + static_cast<size_t>(plane) * N * N;
        float* yout = p.y + static_cast<size_t>(plane) * N * N;
        ex.phase([&](int tid) { red[tid] = neg_inf(); });
        for (int h = 0; h < 2; ++h) {
            rows_forward(ex, xin, h == 0 ? 1.0f : -1.0f, smem, st, [](int, int, int, float, float, float, float) {});
            ex.barrier();
            cols_convolve(ex, p, plane, h, smem, st);
            ex.barrier();
            if (h == 0) {
                rows_inverse(ex, smem, st, [&](int tid, int r, int i, int, int, float e0, float e1) {
                    park[((r * 16 + i) * 2 + 0) * THREADS + tid] = e0;
                    park[((r * 16 + i) * 2 + 1) * THREADS + tid] = e1;
                });
            } else {
                rows_inverse(ex, smem, st, [&](int tid, int r, int i, int row, int xx, float o0, float o1) {
                    const float e0 = park[((r * 16 + i) * 2 + 0) * THREADS + tid];
                    const float e1 = park[((r * 16 + i) * 2 + 1) * THREADS + tid];
                    const float a0 = e0 + o0, a1 = e1 + o1, b0 = e0 - o0, b1 = e1 - o1;
                    yout[static_cast<size_t>(row) * N + xx] = a0;
                    yout[static_cast<size_t>(row + 1) * N + xx] = a1;
                    yout[static_cast<size_t>(row + NH) * N + xx] = b0;
                    yout[static_cast<size_t>(row + NH + 1) * N + xx] = b1;
                    red[tid] = fmaxf(red[tid], fmaxf(fmaxf(a0, a1), fmaxf(b0, b1)));
                });
            }
            ex.barrier();
        }
        // ---- per-image maximum; the CTA finishing an image's last plane rescales the image -------------
        const int img = plane / 3;
        ex.phase([&](int tid) {
            if (tid == 0) {
                float mx = red[0];
                for (int t = 1; t < THREADS; ++t) mx = fmaxf(mx, red[t]);
                atomic_max_float(p.img_max + img, mx);
                ex.threadfence();
                flag[0] = atomic_add_int(p.done + img, 1);
            }
        });
        if (flag[0] == 2) {
            ex.phase([&](int tid) {
                ex.threadfence();
                const float m = ex.load_cg(p.img_max + img);
                float4* base = reinterpret_cast<float4*>(p.y + static_cast<size_t>(img) * 3 * N * N);
                for (int i = tid; i < 3 * N * N / 4; i += THREADS) {
                    float4 v = ex.load_cg4(base + i);
                    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (e[k] == m) {
                            const int slot = atomic_add_int(p.tie_count + img, 1);
                            if (slot < p.max_ties) p.tie_pos[img * p.max_ties + slot] = 4 * i + k;
                        }
                    }
                    v.x = e[0] / m; v.y = e[1] / m; v.z = e[2] / m; v.w = e[3] / m;
                    base[i] = v;
                }
            });
        }
        ex.barrier();
    }
}

}  // namespace f256
}  // namespace b200cam
