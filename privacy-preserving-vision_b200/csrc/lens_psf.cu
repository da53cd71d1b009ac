// b200cam: PSF synthesis of the Image_Caption camera (OpticsZernike, Image_Caption/Camera/Lens.py:176-274) as CUDA kernels.
//
//   h (+ tolerance noise)  --phase plate-->  field = A * exp(i delta_l h)           Utils.py:192-205, 396-410; Lens.py:191-213
//   field --zero-pad R -> n = 3R/2, FFT2, x H, IFFT2, crop-->  U                    Utils.py:329-378 (Fresnel propagation)
//   |U|^2 --nearest up-sample + up x up average-->  raw (P x P)                      Utils.py:208, 216-248
//   raw / sum(raw) per channel, optional disc masks and energy loss                  Lens.py:239, 269-274
// and the adjoint chain back to dL/dh.
//
// n is not a power of two for the shipped geometry (R = 896 -> n = 1344 = 2^6 * 3 * 7), so the transforms are a
// mixed-radix Stockham FFT held in shared memory: radices 2/4/8 (register butterflies of fft_radix.cuh), 3/5/7
// (direct), any other prime <= 31 through a generic loop.  The zero padding is never materialised: the row passes
// transform the R live rows only, the column pass reads R of n rows, multiplies by the transfer function (rebuilt from
// two 1-D fp64 tables, H(fx, fy) = E(fx) E(fy)) and writes back only the R rows that survive the crop.
// All arithmetic is complex64 / fp32 except the phase (fp64 sincos, as the reference) and the transfer function.
// Batch independent: 3 planes per step.
#include <cuda_runtime.h>

#include <cstdlib>

#include "../../include/b200cam.h"
#include "compat.cuh"
#include "fft_radix.cuh"

namespace b200cam {
void note_launches(int n);      // b200cam.cu: the library's launch counter
namespace lens {

constexpr int MAX_STAGES = 12;
constexpr int MAX_GENERIC = 31;
constexpr int ROW_LINES = 2;       // image rows per CTA in the row passes
constexpr int COL_LINES = 2;       // spectral columns per CTA in the column pass (16-byte segments per row; 4 CTAs per SM)
constexpr int COL_PAD = 16 / COL_LINES;   // line pitch n + COL_PAD: the transposing loads / stores of a half warp hit 32 distinct banks
constexpr int THREADS = 256;

struct FftPlan {
    int n;
    int nstages;
    int radix[MAX_STAGES];
};

// odd radices first (the first stages write with stride Ns * r: small odd strides are bank-conflict free), then 2 / 4, 8s last
static bool make_plan(int n, FftPlan* pl) {
    pl->n = n;
    pl->nstages = 0;
    int m = n;
    for (int p = 3; p <= MAX_GENERIC; p += 2)
        while (m % p == 0) {
            if (pl->nstages == MAX_STAGES) return false;
            pl->radix[pl->nstages++] = p;
            m /= p;
        }
    int twos = 0;
    while (m % 2 == 0) { m /= 2; ++twos; }
    if (m != 1) return false;
    if (twos % 3 == 1) { if (pl->nstages == MAX_STAGES) return false; pl->radix[pl->nstages++] = 2; }
    if (twos % 3 == 2) { if (pl->nstages == MAX_STAGES) return false; pl->radix[pl->nstages++] = 4; }
    for (int i = 0; i < twos / 3; ++i) { if (pl->nstages == MAX_STAGES) return false; pl->radix[pl->nstages++] = 8; }
    return true;
}

// ---- small DFTs ----------------------------------------------------------------------------------------------
template <int R> struct OddTw;
template <> struct OddTw<3> {
    static __device__ __forceinline__ float c(int k) { return -0.5f; }
    static __device__ __forceinline__ float s(int k) { return k == 1 ? 0.86602540378443864676f : -0.86602540378443864676f; }
};
template <> struct OddTw<5> {
    static __device__ __forceinline__ float c(int k) { return (k == 1 || k == 4) ? 0.30901699437494742410f : -0.80901699437494742410f; }
    static __device__ __forceinline__ float s(int k) {
        return k == 1 ? 0.95105651629515357212f : k == 2 ? 0.58778525229247312917f : k == 3 ? -0.58778525229247312917f : -0.95105651629515357212f;
    }
};
template <> struct OddTw<7> {
    static __device__ __forceinline__ float c(int k) {
        return (k == 1 || k == 6) ? 0.62348980185873353053f : (k == 2 || k == 5) ? -0.22252093395631440429f : -0.90096886790241912624f;
    }
    static __device__ __forceinline__ float s(int k) {
        const float v = (k == 1 || k == 6) ? 0.78183148246802980871f : (k == 2 || k == 5) ? 0.97492791218182360702f : 0.43388373911755812048f;
        return k <= 3 ? v : -v;
    }
};

// v <- DFT_R(v), kernel exp(DIR * 2 pi i a b / R)
// Odd R: conjugate-pair form.  With S_b = v[b] + v[R-b], D_b = v[b] - v[R-b] (b = 1 .. (R-1)/2):
//   out[a]   = v0 + sum_b cos(t_ab) S_b + i DIR sum_b sin(t_ab) D_b,     out[R-a] = the same with the sine part negated
// (R-1)^2 / 2 real FMAs per half instead of (R-1)^2 complex multiplies: 2.2x fewer operations for the radix-7 stage, which is
// half of the butterfly arithmetic of a 1344-point line.
template <int R, int DIR>
__device__ __forceinline__ void small_dft(float2 (&v)[R]) {
    if constexpr (R == 2 || R == 4 || R == 8 || R == 16) {
        RegFFT<R, DIR>::run(v);
    } else {
        constexpr int H = (R - 1) / 2;
        float2 S[H], D[H];
#pragma unroll
        for (int b = 1; b <= H; ++b) {
            S[b - 1] = make_float2(v[b].x + v[R - b].x, v[b].y + v[R - b].y);
            D[b - 1] = make_float2(v[b].x - v[R - b].x, v[b].y - v[R - b].y);
        }
        float2 o[R];
        o[0] = v[0];
#pragma unroll
        for (int b = 0; b < H; ++b) {
            o[0].x += S[b].x;
            o[0].y += S[b].y;
        }
#pragma unroll
        for (int a = 1; a <= H; ++a) {
            float2 A = v[0], B = make_float2(0.f, 0.f);
#pragma unroll
            for (int b = 1; b <= H; ++b) {
                const int k = (a * b) % R;
                const float c = OddTw<R>::c(k), sn = OddTw<R>::s(k);
                A.x = fmaf(c, S[b - 1].x, A.x);
                A.y = fmaf(c, S[b - 1].y, A.y);
                B.x = fmaf(sn, D[b - 1].x, B.x);
                B.y = fmaf(sn, D[b - 1].y, B.y);
            }
            // i * DIR * B
            const float ix = DIR > 0 ? -B.y : B.y, iy = DIR > 0 ? B.x : -B.x;
            o[a] = make_float2(A.x + ix, A.y + iy);
            o[R - a] = make_float2(A.x - ix, A.y - iy);
        }
#pragma unroll
        for (int a = 0; a < R; ++a) v[a] = o[a];
    }
}

// One Stockham stage of radix R over `lines` lines of n points each (line l at in + l * pitch), out of place.
//   out[(j / Ns) * Ns * R + j % Ns + t * Ns] = DFT_R over t' of in[j + t' * n / R] * w^{t' (j % Ns)},  w = exp(DIR 2 pi i / (Ns R))
// tw[q] = exp(-2 pi i q / n).
// CN / CNS > 0: line length and sub-transform size known at compile time (the shipped 1344-point plan): every index
// multiplier, the j % Ns and the twiddle stride fold into constants - the generic stage spends most of its instructions on
// that integer arithmetic, not on the butterflies.
template <int R, int DIR, int CN = 0, int CNS = 0>
__device__ __forceinline__ void stage(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw, int n_rt, int Ns_rt,
                                      int lines, int pitch) {
    const int n = CN > 0 ? CN : n_rt;
    const int Ns = CNS > 0 ? CNS : Ns_rt;
    const int nb = n / R;
    const int twstep = n / (Ns * R);
    const float inv_ns = 1.0f / static_cast<float>(Ns);
    const int per = blockDim.x / lines;                       // threads per line (lines divides the block size)
    const int line = threadIdx.x / per, t0 = threadIdx.x - line * per;
    const float2* src = in + line * pitch;
    float2* dst = out + line * pitch;
    for (int j = t0; j < nb; j += per) {
        int k;
        if constexpr (CNS > 0) k = j % CNS;                                                     // constant divisor
        else k = j - __float2int_rz((static_cast<float>(j) + 0.5f) * inv_ns) * Ns;             // j % Ns (j < 2^20: exact)
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) v[t] = src[j + t * nb];
        if (Ns > 1) {
#pragma unroll
            for (int t = 1; t < R; ++t) {
                const float2 wv = tw[t * k * twstep];
                v[t] = DIR < 0 ? cmul(v[t], wv) : cmulc(v[t], wv);
            }
        }
        small_dft<R, DIR>(v);
        const int base = (j - k) * R + k;
#pragma unroll
        for (int t = 0; t < R; ++t) dst[base + t * Ns] = v[t];
    }
}

// any prime radix r <= MAX_GENERIC (local-memory arrays: slow, only for odd geometries such as the constructor default 736 -> 1104 = 2^4 3 23)
template <int DIR>
__device__ void stage_generic(int r, const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw, int n, int Ns,
                              int lines, int pitch) {
    const int nb = n / r;
    const int twstep = n / (Ns * r);
    const int rstep = n / r;
    for (int w = threadIdx.x; w < lines * nb; w += blockDim.x) {
        const int line = w / nb, j = w - line * nb;
        const float2* src = in + line * pitch;
        float2* dst = out + line * pitch;
        const int k = j % Ns;
        float2 v[MAX_GENERIC];
        for (int t = 0; t < r; ++t) {
            float2 x = src[j + t * nb];
            if (Ns > 1 && t > 0) {
                const float2 wv = tw[t * k * twstep];
                x = DIR < 0 ? cmul(x, wv) : cmulc(x, wv);
            }
            v[t] = x;
        }
        const int base = (j - k) * r + k;
        for (int a = 0; a < r; ++a) {
            float2 acc = v[0];
            for (int b = 1; b < r; ++b) {
                const float2 wv = tw[((a * b) % r) * rstep];
                const float2 p = DIR < 0 ? cmul(v[b], wv) : cmulc(v[b], wv);
                acc.x += p.x;
                acc.y += p.y;
            }
            dst[base + a * Ns] = acc;
        }
    }
}

// full transform of `lines` lines; returns the buffer that holds the result (a or b).  Block barriers inside.
template <int DIR>
__device__ float2* fft_lines(float2* a, float2* b, const float2* tw, const FftPlan& pl, int lines, int pitch) {
    if (pl.n == 1344) {             // make_plan(1344) = {3, 7, 8, 8}: the shipped geometry (wave resolution 896), fully constant
        stage<3, DIR, 1344, 1>(a, b, tw, 1344, 1, lines, pitch);
        __syncthreads();
        stage<7, DIR, 1344, 3>(b, a, tw, 1344, 3, lines, pitch);
        __syncthreads();
        stage<8, DIR, 1344, 21>(a, b, tw, 1344, 21, lines, pitch);
        __syncthreads();
        stage<8, DIR, 1344, 168>(b, a, tw, 1344, 168, lines, pitch);
        __syncthreads();
        return a;
    }
    int Ns = 1;
    for (int s = 0; s < pl.nstages; ++s) {
        const int r = pl.radix[s];
        switch (r) {
            case 2: stage<2, DIR>(a, b, tw, pl.n, Ns, lines, pitch); break;
            case 3: stage<3, DIR>(a, b, tw, pl.n, Ns, lines, pitch); break;
            case 4: stage<4, DIR>(a, b, tw, pl.n, Ns, lines, pitch); break;
            case 5: stage<5, DIR>(a, b, tw, pl.n, Ns, lines, pitch); break;
            case 7: stage<7, DIR>(a, b, tw, pl.n, Ns, lines, pitch); break;
            case 8: stage<8, DIR>(a, b, tw, pl.n, Ns, lines, pitch); break;
            default: stage_generic<DIR>(r, a, b, tw, pl.n, Ns, lines, pitch); break;
        }
        __syncthreads();
        Ns *= r;
        float2* t = a;
        a = b;
        b = t;
    }
    return a;
}

extern __shared__ __align__(16) unsigned char lens_smem[];

// geometry shared by all kernels
struct Geom {
    int R;        // wave resolution (live rows / columns)
    int n;        // padded size R + 2 * pad
    int pad;      // R / 4
    int P;        // patch size (PSF side)
    int up;       // pooling window; the intensity is nearest-resized to up * P first (Utils.py:216-248)
    FftPlan plan;
    double delta[3];   // 2 pi / lambda * (n_lambda - 1)   (Utils.py:192-197)
};
// 32-bit arithmetic: k < up * P <= 10 * R and R <= 4096 (make_geom), so k * R < 2^31
__device__ __forceinline__ int src_index(const Geom& g, int k) {
    return static_cast<int>((static_cast<unsigned>(k) * static_cast<unsigned>(g.R)) / static_cast<unsigned>(g.up * g.P));
}
// first k of the up-sampled axis that maps to source index y:  ceil(y * U / R)
__device__ __forceinline__ int first_k(const Geom& g, int y) {
    const unsigned U = static_cast<unsigned>(g.up * g.P);
    return static_cast<int>((static_cast<unsigned>(y) * U + static_cast<unsigned>(g.R) - 1u) / static_cast<unsigned>(g.R));
}

// ---- forward ---------------------------------------------------------------------------------------------------
// K1: phase plate, wavefront / aperture, zero-padded forward row transform.  grid (R / ROW_LINES, 3)
struct RowsFwdParams {
    const float* h;        // [R][R]
    const float* noise;    // nullable [R][R]: U(-tol, tol) drawn by the caller (PhasePlate._build, Utils.py:396-406)
    const float2* A;       // [3][R][R] aperture * spherical wavefront (Lens.py:191-213, Utils.py:88-97)
    float2* field;         // out [3][R][R]
    float2* W;             // out [3][R][n] row spectra of the live rows
    const float2* tw;
};
__global__ void __launch_bounds__(THREADS) k_lens_rows_fwd(Geom g, RowsFwdParams p) {
    float2* a = reinterpret_cast<float2*>(lens_smem);
    float2* b = a + ROW_LINES * g.n;
    float2* stw = b + ROW_LINES * g.n;                               // the twiddle table, in shared memory (random 8-byte look-ups)
    for (int i = threadIdx.x; i < g.n; i += blockDim.x) stw[i] = __ldg(p.tw + i);
    const int lam = blockIdx.y;
    const int y0 = blockIdx.x * ROW_LINES;
    for (int i = threadIdx.x; i < ROW_LINES * g.n; i += blockDim.x) a[i] = make_float2(0.f, 0.f);
    __syncthreads();
    for (int i = threadIdx.x; i < ROW_LINES * g.R; i += blockDim.x) {
        const int l = i / g.R, x = i - l * g.R;
        const int y = y0 + l;
        if (y < g.R) {
            float hh = __ldg(p.h + static_cast<size_t>(y) * g.R + x);
            if (p.noise != nullptr) hh += __ldg(p.noise + static_cast<size_t>(y) * g.R + x);
            double s, c;
            sincos(g.delta[lam] * static_cast<double>(hh), &s, &c);          // fp64 phase, cast to complex64 (Utils.py:80-85)
            const float2 sh = make_float2(static_cast<float>(c), static_cast<float>(s));
            const float2 A = __ldg(p.A + (static_cast<size_t>(lam) * g.R + y) * g.R + x);
            const float2 f = make_float2(sh.x * A.x - sh.y * A.y, sh.x * A.y + sh.y * A.x);
            p.field[(static_cast<size_t>(lam) * g.R + y) * g.R + x] = f;
            a[l * g.n + g.pad + x] = f;
        }
    }
    __syncthreads();
    const float2* res = fft_lines<-1>(a, b, stw, g.plan, ROW_LINES, g.n);
    for (int i = threadIdx.x; i < ROW_LINES * g.n; i += blockDim.x) {
        const int l = i / g.n, u = i - l * g.n;
        const int y = y0 + l;
        if (y < g.R) p.W[(static_cast<size_t>(lam) * g.R + y) * g.n + u] = res[i];
    }
}

// K2: column transform, x transfer function (or its conjugate: adjoint), inverse column transform, in place on the live rows.
// grid (n / COL_LINES, 3)
struct ColsParams {
    float2* W;             // [3][R][n]
    const double2* Hx;     // [3][n]  exp(-i pi lambda z f_k^2), f in FFT order (Utils.py:365-373), fp64
    const float2* tw;
    int conj_h;
};
__global__ void __launch_bounds__(THREADS) k_lens_cols(Geom g, ColsParams p) {
    float2* a = reinterpret_cast<float2*>(lens_smem);
    const int pitch = g.n + COL_PAD;
    float2* b = a + COL_LINES * pitch;
    float2* stw = b + COL_LINES * pitch;                               // the twiddle table, in shared memory (random 8-byte look-ups)
    for (int i = threadIdx.x; i < g.n; i += blockDim.x) stw[i] = __ldg(p.tw + i);
    const int lam = blockIdx.y;
    const int u0 = blockIdx.x * COL_LINES;
    for (int i = threadIdx.x; i < COL_LINES * pitch; i += blockDim.x) a[i] = make_float2(0.f, 0.f);
    __syncthreads();
    for (int i = threadIdx.x; i < COL_LINES * g.R; i += blockDim.x) {
        const int y = i / COL_LINES, c = i - y * COL_LINES;
        if (u0 + c < g.n) a[c * pitch + g.pad + y] = p.W[(static_cast<size_t>(lam) * g.R + y) * g.n + u0 + c];
    }
    __syncthreads();
    float2* res = fft_lines<-1>(a, b, stw, g.plan, COL_LINES, pitch);
    float2* other = res == a ? b : a;
    for (int i = threadIdx.x; i < COL_LINES * g.n; i += blockDim.x) {
        const int c = i / g.n, v = i - c * g.n;
        if (u0 + c < g.n) {
            const double2 hu = __ldg(p.Hx + lam * g.n + u0 + c);
            const double2 hv = __ldg(p.Hx + lam * g.n + v);
            const float hr = static_cast<float>(hu.x * hv.x - hu.y * hv.y);
            float hi = static_cast<float>(hu.x * hv.y + hu.y * hv.x);
            if (p.conj_h) hi = -hi;
            const float2 z = res[c * pitch + v];
            res[c * pitch + v] = make_float2(z.x * hr - z.y * hi, z.x * hi + z.y * hr);
        }
    }
    __syncthreads();
    const float2* out = fft_lines<+1>(res, other, stw, g.plan, COL_LINES, pitch);
    for (int i = threadIdx.x; i < COL_LINES * g.R; i += blockDim.x) {
        const int y = i / COL_LINES, c = i - y * COL_LINES;
        if (u0 + c < g.n) p.W[(static_cast<size_t>(lam) * g.R + y) * g.n + u0 + c] = out[c * pitch + g.pad + y];
    }
}

// K3: inverse row transform, crop, 1/n^2, intensity, horizontal half of the area down-sampling.  grid (R / ROW_LINES, 3)
struct RowsInvParams {
    const float2* W;       // [3][R][n]
    float2* U;             // out [3][R][R] propagated field (kept for the backward)
    float* Ih;             // out [3][R][P] horizontally pooled intensity (sum, not mean)
    const float2* tw;
};
__global__ void __launch_bounds__(THREADS) k_lens_rows_inv(Geom g, RowsInvParams p) {
    float2* a = reinterpret_cast<float2*>(lens_smem);
    float2* b = a + ROW_LINES * g.n;
    float2* stw = b + ROW_LINES * g.n;                               // the twiddle table, in shared memory (random 8-byte look-ups)
    for (int i = threadIdx.x; i < g.n; i += blockDim.x) stw[i] = __ldg(p.tw + i);
    const int lam = blockIdx.y;
    const int y0 = blockIdx.x * ROW_LINES;
    for (int i = threadIdx.x; i < ROW_LINES * g.n; i += blockDim.x) {
        const int l = i / g.n, u = i - l * g.n;
        const int y = y0 + l;
        a[i] = y < g.R ? p.W[(static_cast<size_t>(lam) * g.R + y) * g.n + u] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    float2* res = fft_lines<+1>(a, b, stw, g.plan, ROW_LINES, g.n);
    float* inten = reinterpret_cast<float*>(res == a ? b : a);      // [ROW_LINES][R]
    const float sc = 1.0f / (static_cast<float>(g.n) * static_cast<float>(g.n));
    for (int i = threadIdx.x; i < ROW_LINES * g.R; i += blockDim.x) {
        const int l = i / g.R, x = i - l * g.R;
        const int y = y0 + l;
        float2 u = res[l * g.n + g.pad + x];
        u.x *= sc;
        u.y *= sc;
        inten[i] = u.x * u.x + u.y * u.y;                            // get_intensities, Utils.py:208
        if (y < g.R) p.U[(static_cast<size_t>(lam) * g.R + y) * g.R + x] = u;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ROW_LINES * g.P; i += blockDim.x) {
        const int l = i / g.P, j = i - l * g.P;
        const int y = y0 + l;
        float acc = 0.f;
        for (int q = 0; q < g.up; ++q) acc += inten[l * g.R + src_index(g, g.up * j + q)];
        if (y < g.R) p.Ih[(static_cast<size_t>(lam) * g.R + y) * g.P + j] = acc;
    }
}

__device__ __forceinline__ float block_sum(float v, float* red) {     // fixed-order tree: deterministic
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
    return t;
}
__device__ __forceinline__ double block_sum_d(double v, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
    return t;
}

// K4: vertical half of the down-sampling and the row sums.  grid (P, 3)
struct PoolParams {
    const float* Ih;       // [3][R][P]
    float* raw;            // out [3][P][P]
    float* rowsum;         // out [3][P]
};
__global__ void __launch_bounds__(THREADS) k_lens_pool(Geom g, PoolParams p) {
    __shared__ float red[32];
    const int lam = blockIdx.y, i = blockIdx.x;
    const float inv = 1.0f / (static_cast<float>(g.up) * static_cast<float>(g.up));
    float part = 0.f;
    for (int j = threadIdx.x; j < g.P; j += blockDim.x) {
        float acc = 0.f;
        for (int q = 0; q < g.up; ++q) acc += __ldg(p.Ih + (static_cast<size_t>(lam) * g.R + src_index(g, g.up * i + q)) * g.P + j);
        acc *= inv;
        p.raw[(static_cast<size_t>(lam) * g.P + i) * g.P + j] = acc;
        part += acc;
    }
    const float t = block_sum(part, red);
    if (threadIdx.x == 0) p.rowsum[lam * g.P + i] = t;
}

// K5: per-channel normalisation (Lens.py:239), optional disc masks and energy loss (Lens.py:269-274).  grid (P, 3)
struct NormParams {
    const float* raw;      // [3][P][P]
    const float* rowsum;   // [3][P]
    float* psf;            // out [P][P][3] fp32, normalised, unmasked (kept for the backward)
    float* chan_sum;       // out [3]
    const double* mask1;   // [P][P][3] fp64 or null
    const double* mask2;
    double* psf_out;       // out [P][P][3] fp64: psf (* mask2 with flag 2)  - the reference returns fp64 (fp64 masks promote)
    double* loss_part;     // out [3 * P] partial sums of (psf * mask1 - psf)^2 (flag 1)
    int flags;
};
__global__ void __launch_bounds__(THREADS) k_lens_norm(Geom g, NormParams p) {
    __shared__ float red[32];
    __shared__ double redd[32];
    const int lam = blockIdx.y, i = blockIdx.x;
    float part = 0.f;
    for (int q = threadIdx.x; q < g.P; q += blockDim.x) part += __ldg(p.rowsum + lam * g.P + q);
    const float S = block_sum(part, red);
    if (i == 0 && threadIdx.x == 0) p.chan_sum[lam] = S;
    double lp = 0.0;
    for (int j = threadIdx.x; j < g.P; j += blockDim.x) {
        const float v = __ldg(p.raw + (static_cast<size_t>(lam) * g.P + i) * g.P + j) / S;
        const size_t o = (static_cast<size_t>(i) * g.P + j) * 3 + lam;
        p.psf[o] = v;
        double out = static_cast<double>(v);
        if (p.flags & 1) {
            const double d = out * __ldg(p.mask1 + o) - out;
            lp += d * d;
        }
        if (p.flags & 2) out *= __ldg(p.mask2 + o);
        p.psf_out[o] = out;
    }
    if (p.flags & 1) {
        const double t = block_sum_d(lp, redd);
        if (threadIdx.x == 0) p.loss_part[lam * g.P + i] = t;
    }
}
// loss = sqrt(sum of the partials): torch.norm (Lens.py:270).  one CTA
__global__ void __launch_bounds__(THREADS) k_lens_loss(const double* part, int count, double* loss) {
    __shared__ double redd[32];
    double v = 0.0;
    for (int q = threadIdx.x; q < count; q += blockDim.x) v += part[q];
    const double t = block_sum_d(v, redd);
    if (threadIdx.x == 0) *loss = sqrt(t);
}

// ---- backward --------------------------------------------------------------------------------------------------
// Kb1: adjoint of masks / loss; row partials of sum(gn * psf).  grid (P, 3)
struct NormBwdParams {
    const double* gout;    // [P][P][3] dL/dpsf_out (fp64) or null
    const double* gloss;   // device scalar dL/dloss or null
    const double* loss;    // device scalar (forward value)
    const float* psf;      // [P][P][3]
    const double* mask1;
    const double* mask2;
    float* gn;             // out [3][P][P] dL/dpsf (normalised, unmasked)
    float* rowdot;         // out [3][P]
    int flags;
};
__global__ void __launch_bounds__(THREADS) k_lens_norm_bwd(Geom g, NormBwdParams p) {
    __shared__ float red[32];
    const int lam = blockIdx.y, i = blockIdx.x;
    double gl = 0.0;
    if ((p.flags & 1) && p.gloss != nullptr) {
        const double L = *p.loss;
        gl = L > 0.0 ? *p.gloss / L : 0.0;
    }
    float part = 0.f;
    for (int j = threadIdx.x; j < g.P; j += blockDim.x) {
        const size_t o = (static_cast<size_t>(i) * g.P + j) * 3 + lam;
        const float ps = __ldg(p.psf + o);
        double gv = p.gout != nullptr ? __ldg(p.gout + o) : 0.0;
        if (p.flags & 2) gv *= __ldg(p.mask2 + o);
        if (p.flags & 1) {
            const double m = __ldg(p.mask1 + o) - 1.0;
            gv += gl * static_cast<double>(ps) * m * m;
        }
        const float gf = static_cast<float>(gv);
        p.gn[(static_cast<size_t>(lam) * g.P + i) * g.P + j] = gf;
        part += gf * ps;
    }
    const float t = block_sum(part, red);
    if (threadIdx.x == 0) p.rowdot[lam * g.P + i] = t;
}

// Kb3: adjoint of normalisation + down-sampling + intensity, zero-padded forward row transform.  grid (R / ROW_LINES, 3)
struct RowsBwdParams {
    const float* gn;       // [3][P][P]
    const float* rowdot;   // [3][P]
    const float* chan_sum; // [3]
    const float2* U;       // [3][R][R]
    float2* W;             // out [3][R][n]
    const float2* tw;
};
__global__ void __launch_bounds__(THREADS) k_lens_rows_bwd(Geom g, RowsBwdParams p) {
    __shared__ float red[32];
    float2* a = reinterpret_cast<float2*>(lens_smem);
    float2* b = a + ROW_LINES * g.n;
    float2* stw = b + ROW_LINES * g.n;                               // the twiddle table, in shared memory (random 8-byte look-ups)
    for (int i = threadIdx.x; i < g.n; i += blockDim.x) stw[i] = __ldg(p.tw + i);
    const int lam = blockIdx.y;
    const int y0 = blockIdx.x * ROW_LINES;
    float part = 0.f;
    for (int q = threadIdx.x; q < g.P; q += blockDim.x) part += __ldg(p.rowdot + lam * g.P + q);
    const float dot = block_sum(part, red);
    const float S = __ldg(p.chan_sum + lam);
    const float inv = 1.0f / (S * static_cast<float>(g.up) * static_cast<float>(g.up));
    for (int i = threadIdx.x; i < ROW_LINES * g.n; i += blockDim.x) a[i] = make_float2(0.f, 0.f);
    __syncthreads();
    for (int i = threadIdx.x; i < ROW_LINES * g.R; i += blockDim.x) {
        const int l = i / g.R, x = i - l * g.R;
        const int y = y0 + l;
        if (y < g.R) {
            const int ky0 = first_k(g, y), ky1 = first_k(g, y + 1);
            const int kx0 = first_k(g, x), kx1 = first_k(g, x + 1);
            float acc = 0.f;
            for (int ky = ky0; ky < ky1; ++ky)
                for (int kx = kx0; kx < kx1; ++kx) acc += __ldg(p.gn + (static_cast<size_t>(lam) * g.P + ky / g.up) * g.P + kx / g.up);
            const float cnt = static_cast<float>((ky1 - ky0) * (kx1 - kx0));
            const float gI = (acc - cnt * dot) * inv;
            const float2 u = __ldg(p.U + (static_cast<size_t>(lam) * g.R + y) * g.R + x);
            a[l * g.n + g.pad + x] = make_float2(2.f * gI * u.x, 2.f * gI * u.y);
        }
    }
    __syncthreads();
    const float2* res = fft_lines<-1>(a, b, stw, g.plan, ROW_LINES, g.n);
    for (int i = threadIdx.x; i < ROW_LINES * g.n; i += blockDim.x) {
        const int l = i / g.n, u = i - l * g.n;
        const int y = y0 + l;
        if (y < g.R) p.W[(static_cast<size_t>(lam) * g.R + y) * g.n + u] = res[i];
    }
}

// Kb5: inverse row transform of the three planes, crop, dL/dphi = Im(conj(field) dfield), dL/dh = sum_l delta_l dphi_l.
// grid (R / ROW_LINES)
struct HGradParams {
    const float2* W;       // [3][R][n]
    const float2* field;   // [3][R][R]
    float* grad_h;         // out [R][R]
    const float2* tw;
};
__global__ void __launch_bounds__(THREADS) k_lens_hgrad(Geom g, HGradParams p) {
    float2* a = reinterpret_cast<float2*>(lens_smem);
    float2* b = a + ROW_LINES * g.n;
    float2* stw = b + ROW_LINES * g.n;                               // the twiddle table, in shared memory (random 8-byte look-ups)
    for (int i = threadIdx.x; i < g.n; i += blockDim.x) stw[i] = __ldg(p.tw + i);
    double* acc = reinterpret_cast<double*>(stw + g.n);      // [ROW_LINES][R]
    const int y0 = blockIdx.x * ROW_LINES;
    const double sc = 1.0 / (static_cast<double>(g.n) * static_cast<double>(g.n));
    for (int i = threadIdx.x; i < ROW_LINES * g.R; i += blockDim.x) acc[i] = 0.0;
    for (int lam = 0; lam < 3; ++lam) {
        __syncthreads();
        for (int i = threadIdx.x; i < ROW_LINES * g.n; i += blockDim.x) {
            const int l = i / g.n, u = i - l * g.n;
            const int y = y0 + l;
            a[i] = y < g.R ? p.W[(static_cast<size_t>(lam) * g.R + y) * g.n + u] : make_float2(0.f, 0.f);
        }
        __syncthreads();
        const float2* res = fft_lines<+1>(a, b, stw, g.plan, ROW_LINES, g.n);
        for (int i = threadIdx.x; i < ROW_LINES * g.R; i += blockDim.x) {
            const int l = i / g.R, x = i - l * g.R;
            const int y = y0 + l;
            if (y < g.R) {
                const float2 gf = res[l * g.n + g.pad + x];
                const float2 f = __ldg(p.field + (static_cast<size_t>(lam) * g.R + y) * g.R + x);
                const float dphi = f.x * gf.y - f.y * gf.x;
                acc[i] += g.delta[lam] * sc * static_cast<double>(dphi);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ROW_LINES * g.R; i += blockDim.x) {
        const int l = i / g.R, x = i - l * g.R;
        const int y = y0 + l;
        if (y < g.R) p.grad_h[static_cast<size_t>(y) * g.R + x] = static_cast<float>(acc[i]);
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
static int pool_factor(int R, int P) {      // area_downsampling_tf, Utils.py:216-248
    if (R % P == 0) return R / P;
    long long a = R, b = P;
    while (b) { const long long t = a % b; a = b; b = t; }
    const long long lcm_over_p = static_cast<long long>(R) / a;      // lcm(R, P) / P
    return lcm_over_p > 10 ? 10 : static_cast<int>(lcm_over_p);
}

static bool make_geom(int R, int P, const double* delta, Geom* g) {
    if (R < 8 || R > 4096 || R % 4 != 0 || P < 1 || P > R) return false;
    g->R = R;
    g->pad = R / 4;
    g->n = R + 2 * g->pad;
    g->P = P;
    g->up = pool_factor(R, P);
    if (!make_plan(g->n, &g->plan)) return false;
    for (int i = 0; i < 3; ++i) g->delta[i] = delta != nullptr ? delta[i] : 0.0;
    return true;
}

struct Ws {
    float2* W;
    float* Ih;
    float* raw;
    float* rowsum;
    float* gn;
    float* rowdot;
    double* loss_part;
    size_t bytes;
    Ws(void* base, const Geom& g) {
        size_t off = 0;
        auto take = [&](size_t b) {
            void* p = base ? static_cast<char*>(base) + off : nullptr;
            off += (b + 255) / 256 * 256;
            return p;
        };
        W = static_cast<float2*>(take(sizeof(float2) * 3 * g.R * g.n));
        Ih = static_cast<float*>(take(sizeof(float) * 3 * g.R * g.P));
        raw = static_cast<float*>(take(sizeof(float) * 3 * g.P * g.P));
        rowsum = static_cast<float*>(take(sizeof(float) * 3 * g.P));
        gn = static_cast<float*>(take(sizeof(float) * 3 * g.P * g.P));
        rowdot = static_cast<float*>(take(sizeof(float) * 3 * g.P));
        loss_part = static_cast<double*>(take(sizeof(double) * 3 * g.P));
        bytes = off;
    }
};

static size_t rows_smem(const Geom& g) { return sizeof(float2) * (2 * ROW_LINES + 1) * g.n; }
static size_t hgrad_smem(const Geom& g) { return rows_smem(g) + sizeof(double) * ROW_LINES * g.R; }
static size_t cols_smem(const Geom& g) { return sizeof(float2) * (2 * COL_LINES * (g.n + COL_PAD) + g.n); }

template <class K>
static cudaError_t optin(K kernel, size_t bytes) {
    return bytes > 48 * 1024 ? cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)) : cudaSuccess;
}

#define LCK(expr)                                               \
    do {                                                        \
        cudaError_t e_ = (expr);                                \
        if (e_ != cudaSuccess) return static_cast<int>(e_);     \
    } while (0)
#define LLAUNCH()                 \
    do {                          \
        LCK(cudaGetLastError());  \
        note_launches(1);         \
    } while (0)

}  // namespace lens
}  // namespace b200cam

using namespace b200cam;
using namespace b200cam::lens;

extern "C" {

int b200cam_lens_psf_supported(int R, int P) {
    Geom g;
    if (!make_geom(R, P, nullptr, &g)) return 0;
    return cols_smem(g) <= 200 * 1024 && hgrad_smem(g) <= 200 * 1024 ? 1 : 0;
}

int b200cam_lens_psf_padded(int R) { return R + 2 * (R / 4); }

size_t b200cam_lens_psf_workspace_bytes(int R, int P) {
    Geom g;
    if (!make_geom(R, P, nullptr, &g)) return 0;
    return Ws(nullptr, g).bytes;
}

int b200cam_lens_psf_fwd(const float* h, const float* noise, const float* A, const double* delta, const double* Hx, const float* tw,
                         float* field, float* U, float* psf, float* chan_sum, const double* mask1, const double* mask2, int flags,
                         double* psf_out, double* loss, void* workspace, size_t workspace_bytes, int R, int P, void* stream) {
    Geom g;
    if (delta == nullptr || !make_geom(R, P, delta, &g) || !b200cam_lens_psf_supported(R, P)) return B200CAM_E_BAD_SIZE;
    if (!h || !A || !Hx || !tw || !field || !U || !psf || !chan_sum || !psf_out || !workspace) return B200CAM_E_NULL;
    if (((flags & 1) && (!mask1 || !loss)) || ((flags & 2) && !mask2)) return B200CAM_E_NULL;
    Ws ws(workspace, g);
    if (workspace_bytes < ws.bytes) return B200CAM_E_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const float2* twp = reinterpret_cast<const float2*>(tw);
    const int rgrid = (R + ROW_LINES - 1) / ROW_LINES, cgrid = (g.n + COL_LINES - 1) / COL_LINES;
    LCK(optin(k_lens_rows_fwd, rows_smem(g)));
    LCK(optin(k_lens_cols, cols_smem(g)));
    LCK(optin(k_lens_rows_inv, rows_smem(g)));
    k_lens_rows_fwd<<<dim3(rgrid, 3), THREADS, rows_smem(g), s>>>(
        g, RowsFwdParams{h, noise, reinterpret_cast<const float2*>(A), reinterpret_cast<float2*>(field), ws.W, twp});
    LLAUNCH();
    k_lens_cols<<<dim3(cgrid, 3), THREADS, cols_smem(g), s>>>(g, ColsParams{ws.W, reinterpret_cast<const double2*>(Hx), twp, 0});
    LLAUNCH();
    k_lens_rows_inv<<<dim3(rgrid, 3), THREADS, rows_smem(g), s>>>(g, RowsInvParams{ws.W, reinterpret_cast<float2*>(U), ws.Ih, twp});
    LLAUNCH();
    k_lens_pool<<<dim3(P, 3), THREADS, 0, s>>>(g, PoolParams{ws.Ih, ws.raw, ws.rowsum});
    LLAUNCH();
    k_lens_norm<<<dim3(P, 3), THREADS, 0, s>>>(g, NormParams{ws.raw, ws.rowsum, psf, chan_sum, mask1, mask2, psf_out, ws.loss_part, flags});
    LLAUNCH();
    if (flags & 1) {
        k_lens_loss<<<1, THREADS, 0, s>>>(ws.loss_part, 3 * P, loss);
        LLAUNCH();
    }
    return 0;
}

int b200cam_lens_psf_bwd(const double* grad_psf_out, const double* grad_loss, const double* loss, const float* psf, const float* chan_sum,
                         const float* field, const float* U, const double* delta, const double* Hx, const float* tw, const double* mask1,
                         const double* mask2, int flags, float* grad_h, void* workspace, size_t workspace_bytes, int R, int P,
                         void* stream) {
    Geom g;
    if (delta == nullptr || !make_geom(R, P, delta, &g) || !b200cam_lens_psf_supported(R, P)) return B200CAM_E_BAD_SIZE;
    if (!psf || !chan_sum || !field || !U || !Hx || !tw || !grad_h || !workspace) return B200CAM_E_NULL;
    if (((flags & 1) && (!mask1 || !loss)) || ((flags & 2) && !mask2)) return B200CAM_E_NULL;
    Ws ws(workspace, g);
    if (workspace_bytes < ws.bytes) return B200CAM_E_WORKSPACE;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const float2* twp = reinterpret_cast<const float2*>(tw);
    const int rgrid = (R + ROW_LINES - 1) / ROW_LINES, cgrid = (g.n + COL_LINES - 1) / COL_LINES;
    LCK(optin(k_lens_rows_bwd, rows_smem(g)));
    LCK(optin(k_lens_cols, cols_smem(g)));
    LCK(optin(k_lens_hgrad, hgrad_smem(g)));
    k_lens_norm_bwd<<<dim3(P, 3), THREADS, 0, s>>>(g, NormBwdParams{grad_psf_out, grad_loss, loss, psf, mask1, mask2, ws.gn, ws.rowdot, flags});
    LLAUNCH();
    k_lens_rows_bwd<<<dim3(rgrid, 3), THREADS, rows_smem(g), s>>>(
        g, RowsBwdParams{ws.gn, ws.rowdot, chan_sum, reinterpret_cast<const float2*>(U), ws.W, twp});
    LLAUNCH();
    k_lens_cols<<<dim3(cgrid, 3), THREADS, cols_smem(g), s>>>(g, ColsParams{ws.W, reinterpret_cast<const double2*>(Hx), twp, 1});
    LLAUNCH();
    k_lens_hgrad<<<rgrid, THREADS, hgrad_smem(g), s>>>(g, HGradParams{ws.W, reinterpret_cast<const float2*>(field), grad_h, twp});
    LLAUNCH();
    return 0;
}

}  // extern "C"
