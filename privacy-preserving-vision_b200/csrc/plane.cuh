// b200cam: plane-resident sensor kernels for N = 256 (round 2).
//
// conv2D (Face-DeId/Camera/Utils.py:7-12) + the per-image amax normalisation (Optics.py:128) and their adjoint, as
// three kernels instead of six passes:
//
//   k_prow   rows:   x -> real row FFTs -> half spectra "A" (runs beside the PSF chain; needs no PSF)
//   k_pconv  cluster of 8 CTAs = one plane:  A -> column FFT -> [X^ kept in place for the backward] -> x OTF ->
//            inverse columns -> crossing through an L2-resident scratch -> cluster barrier -> inverse rows ->
//            per-image max taken ACROSS the three clusters of the image (they are co-resident and run in lock step) ->
//            y = conv / max written once.  No un-normalised image, no normalise pass.
//   k_pacc   cluster of 8 CTAs = one plane:  g -> row FFTs -> crossing -> cluster barrier -> column FFT ->
//            acc += conj(X^) G^ / max  (registers, the cluster always owns the same channel) + Parseval partials of
//            sum(g * conv) -> one partial plane per cluster triple at the end.
//
// Why the crossing goes through L2 and not through distributed shared memory: measured on B200
// (tools/microbench/dsmem_probe.cu, dsmem_bulk_probe.cu; profiles/r02_dsmem_probe.md) remote st.shared::cluster
// sustains 6-10 B/clk/SM (<= 2.9 TB/s chip wide, bulk copies no better) while an L2-resident buffer is read+written at
// 15-21 TB/s.  The scratch of a cluster is 2 x 256 KB, reused for every plane it processes, so it never leaves L2; the
// cluster supplies what the exchange needs besides bandwidth: co-scheduling and a ~500-cycle hardware barrier.
//
// Real rows are transformed in pairs (P, P+128) as one complex FFT and un-mixed by Hermitian symmetry; rows P and
// P+128 fall on the same lane (P mod 16) of the column transform, so one 16-byte element of the crossing layouts is
// produced by one thread and consumed by one thread (no shuffles).  Spectral column 0 carries DC and Nyquist packed as
// one complex column (both are transforms of real sequences); it is un-mixed where the OTF is applied.
//
// Layouts (float4 = two complex values, "rows P and P+128"):
//   A   [plane][k < 128][P < 128]   (S_P[k], S_{P+128}[k]),  S_y = 2 * rfft(x_y);  k = 0: (S[0].re, S[128].re) per row
//       after k_pconv the same memory holds X^ = column transform of A:  [plane][u][s < 128] = (X^[u][s], X^[u][s+128])
//   Bs  [cluster][3][P < 128][k < 128]  (Y_P[k], Y_{P+128}[k])  inverse-column output, packed at k = 0 like A
//   As  [cluster][3][k][P]          A of the upstream gradient (k_pacc)
// Every body is a sequence of phases over an execution policy (device: one CUDA thread; host: the test-only emulator
// loops over the threads of a whole cluster), see exec.cuh for the idea.
#pragma once

#include "compat.cuh"
#include "fft_plan.cuh"

namespace b200cam {
namespace plane {

constexpr int N = 256;
constexpr int NP = 128;                       // row pairs = packed spectral columns
constexpr int C = 8;                          // CTAs per cluster (one plane per cluster)
constexpr int THREADS = 256;                  // 16 groups of 16 lanes
constexpr int GROUPS = 16;
constexpr int E_GROUP = 272;                  // float2 per group: 16 x 17 exchange block of the two-pass FFT
constexpr int RED_OFF = GROUPS * E_GROUP;     // float2 units
constexpr int SMEM_FLOAT2 = RED_OFF + 144;    // + 288 floats of reduction scratch
constexpr int SMEM_BYTES = SMEM_FLOAT2 * 8;   // 35,968 B
constexpr int PLANE_F4 = NP * NP;             // float4 elements of one plane in the crossing layouts (256 KB)
constexpr int NC = 129;
constexpr int MAXT = 8;                       // recorded arg-max positions per image (B200CAM_MAX_TIES)
using P = Plan<256>;

struct Thread {
    float2 v[16];
    float2 w[16];       // w[k] = exp(-2 pi i lane k / 256)
    float2 u[16];       // scratch of the packed-column branch / backward accumulators
    float2 k[16];       // k_pconv: this thread's OTF values (the cluster always owns the same channel and columns)
    unsigned key;       // order-preserving key of a maximum
    float dot;
};

// ---- global-memory helpers (L2 only: the crossing buffers are written by other CTAs of the cluster) -------------
B200_HD float4 ld_cg4(const float4* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
B200_HD void st_cg4(float4* p, float4 v) {
#if defined(__CUDA_ARCH__)
    __stcg(p, v);
#else
    *p = v;
#endif
}
B200_HD float4 f4(float2 a, float2 b) { return make_float4(a.x, a.y, b.x, b.y); }

// ---- float <-> order-preserving unsigned key (max through integer atomics, zero = below everything) ----------
B200_HD unsigned float_key(float f) {
    unsigned b;
    std::memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
B200_HD float key_float(unsigned k) {
    const unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    float f;
    std::memcpy(&f, &b, 4);
    return f;
}

// Z[256 - k] for k = lane + 16 i, held by lane (16 - lane) mod 16 in register 15 - i (lane 0: own register (16 - i) mod 16)
// `half`: only this half warp (one group) takes part in the exchange (the packed-column branch)
template <class Ctx>
B200_HD float2 mirror_v(Ctx& c, int i, bool half = false) {
    const int b = c.tid & 15;
    float2 m = c.shfl_v(15 - i, (16 - b) & 15, half);
    if (b == 0) m = c.t.v[(16 - i) & 15];
    return m;
}
template <class Ctx>
B200_HD float2 mirror_u(Ctx& c, int i, bool half = false) {
    const int b = c.tid & 15;
    float2 m = c.shfl_u(15 - i, (16 - b) & 15, half);
    if (b == 0) m = c.t.u[(16 - i) & 15];
    return m;
}

template <class Ctx>
B200_HD void load_twiddles(Ctx& c, const float2* tw) {
    const int lane = c.tid & 15;
#pragma unroll
    for (int k = 0; k < 16; ++k) c.t.w[k] = ld_ro(tw + lane * k);
}

// =====================================================================================================================
// Row phase: two real rows (P, P+128) of `src` -> un-mixed half spectra -> dst[k][P] (float4 pitch NP per k)
//   phases: rows_load (global -> registers, first pass, exchange block) | warp sync | rows_unmix_store
// =====================================================================================================================
template <class Ctx>
B200_HD void rows_load(Ctx& c, const float* plane_src, int Pidx) {
    const int g = c.tid >> 4, a = c.tid & 15;
    const float* r0 = plane_src + Pidx * N;
    const float* r1 = r0 + NP * N;
#pragma unroll
    for (int i = 0; i < 16; ++i) c.t.v[i] = make_float2(ld_ro(r0 + 16 * i + a), ld_ro(r1 + 16 * i + a));
    P::stepA(c.t.v, a, c.smem + g * E_GROUP, c.t.w);
}
template <class Ctx>
B200_HD void rows_second(Ctx& c) {
    const int g = c.tid >> 4, b = c.tid & 15;
    P::stepB(c.t.v, b, c.smem + g * E_GROUP);
}
// after rows_second: v[i] = Z[b + 16 i], Z = FFT(row_P + i row_{P+128}).  Writes 2*rfft of both rows for k = b + 16 i, i < 8.
template <class Ctx>
B200_HD void rows_unmix_store(Ctx& c, float4* dst, int Pidx) {
    const int b = c.tid & 15;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 zk = c.t.v[i];
        const float2 zm = mirror_v(c, i);
        float4 o = make_float4(zk.x + zm.x, zk.y - zm.y, zk.y + zm.y, zm.x - zk.x);
        if (i == 0 && b == 0) {
            const float2 z0 = c.t.v[0], zn = c.t.v[8];                 // DC and Nyquist: both real for real rows
            o = make_float4(2.f * z0.x, 2.f * zn.x, 2.f * z0.y, 2.f * zn.y);
        }
        st_cg4(dst + static_cast<size_t>(b + 16 * i) * NP + Pidx, o);
    }
}

// =====================================================================================================================
// Column phase, forward: column u of src[u][P] -> registers v[i] = V[u][b + 16 i]
// =====================================================================================================================
template <class Ctx>
B200_HD void cols_load(Ctx& c, const float4* src, int u) {
    const int g = c.tid >> 4, a = c.tid & 15;
    const float4* col = src + static_cast<size_t>(u) * NP;
    float4 q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = ld_cg4(col + 16 * i + a);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        c.t.v[i] = make_float2(q[i].x, q[i].y);
        c.t.v[i + 8] = make_float2(q[i].z, q[i].w);
    }
    P::stepA(c.t.v, a, c.smem + g * E_GROUP, c.t.w);
}
template <class Ctx>
B200_HD void cols_second(Ctx& c) {
    const int g = c.tid >> 4, b = c.tid & 15;
    P::stepB(c.t.v, b, c.smem + g * E_GROUP);
}

// =====================================================================================================================
// k_prow: persistent grid over (plane, block of 16 row pairs)
// =====================================================================================================================
struct RowParams {
    const float* x;        // [planes][256][256]
    float4* A;             // [planes][128][128]
    const float2* tw;
    int planes;
};

template <class X>
B200_HD void prow_body(X& x, const RowParams& p, int first_item, int item_stride) {
    x.each([&](auto& c) { load_twiddles(c, p.tw); });
    const int items = p.planes * 8;
    for (int it = first_item; it < items; it += item_stride) {
        const int plane = it >> 3, blk = it & 7;
        x.each([&](auto& c) { rows_load(c, p.x + static_cast<size_t>(plane) * N * N, 16 * blk + (c.tid >> 4)); });
        x.sync_warp();
        x.each([&](auto& c) { rows_second(c); });
        x.each([&](auto& c) { rows_unmix_store(c, p.A + static_cast<size_t>(plane) * PLANE_F4, 16 * blk + (c.tid >> 4)); });
        x.sync_warp();                         // the exchange block is rewritten by the next item
    }
}

// =====================================================================================================================
// k_pconv: columns -> x OTF -> inverse columns -> crossing -> inverse rows -> image max -> normalised store
//
// Software pipeline of one CTA over the planes t = 0 .. T-1 of its cluster (everything a plane waits for is started one
// step earlier and collected one step later, so no wait is exposed):
//
//   step t:  H1(t)      column half of plane t: its 16 columns were fetched by ONE bulk copy (cp.async.bulk + mbarrier)
//                        issued during step t-1; the copy for plane t+1 is issued as soon as the stage has been read
//            F(t-2)      finish plane t-2: the image maximum published during step t-1 has long arrived - scale the
//                        stashed outputs, write y
//            wait(t-1)   cluster barrier of plane t-1 (armed at the end of step t-1), gather its rows from the crossing scratch
//            arrive(t)   releases the scatter of H1(t); completes while H2(t-1) and H1(t+1) compute
//            H2(t-1)     inverse rows of plane t-1, CTA max -> global (max, arrival) without waiting, outputs -> stash
//
// Crossing scratch is triple buffered (a CTA may write plane t+1 while a peer still gathers plane t-1).
// =====================================================================================================================
constexpr int STAGE_F4 = 16 * NP;             // 16 columns (or 16 row pairs) x 128 float4 = 32 KB
constexpr int CONV_STAGE_OFF = (RED_OFF + 16 + 15) / 16 * 16;          // float2 units, 128-byte aligned; 16 float2 of reduction words
constexpr int CONV_STASH_OFF = CONV_STAGE_OFF + 2 * STAGE_F4;
constexpr int CONV_BAR_OFF = CONV_STASH_OFF + 16 * THREADS;
constexpr int CONV_SMEM_FLOAT2 = CONV_BAR_OFF + 2;
constexpr int CONV_SMEM_BYTES = CONV_SMEM_FLOAT2 * 8;                  // ~100 KB: two CTAs per SM
constexpr int NBUF = 3;

struct ConvParams {
    float4* A;              // [planes][128][128] in: row spectra; out (save != 0): X^ in place
    const float2* otf;      // [3][129][256], already / N^2, (-1)^(u+v) twist of the centred PSF included
    float4* Bs;             // [nclusters][NBUF][128][128] crossing scratch
    float* y;               // [planes][256][256] out: sensor image (nullptr: only X^ is produced)
    const float2* tw;
    unsigned* sync;         // [B][2] {max key, arrivals}, zero on entry
    float* img_max;         // [B] out: per-image maximum of conv (Optics.py:128)
    int* tie_count;         // [B] zero on entry
    int* tie_pos;           // [B][MAXT]
    int B;
    int G3;                 // cluster triples in the grid: cluster j owns channel j % 3 of images j / 3 + G3 * t
    int save;
    int normalise;          // 1: y = conv / max (Face-DeId); 0: y = conv (plain convolution, Image_Caption)
};

B200_HD int planes_of_cluster(int cluster, int B, int G3) {
    const int q = cluster / 3;
    return q < B ? (B - q + G3 - 1) / G3 : 0;
}

// multiply this thread's column values by the OTF held in k[]; column 0 (DC + Nyquist packed) is un-mixed, multiplied and
// re-packed into u[] (the caller copies it back once every lane has read its partner's v)
template <class Ctx>
B200_HD void conv_multiply(Ctx& c, const ConvParams& p, int ch, int u) {
    const int b = c.tid & 15;
    if (u != 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) c.t.v[i] = cmul(c.t.v[i], c.t.k[i]);
    } else {
        const float2* kny = p.otf + (static_cast<size_t>(ch) * NC + NP) * N;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float2 vk = c.t.v[i];
            const float2 vm = mirror_v(c, i, true);
            const float2 c0 = make_float2(0.5f * (vk.x + vm.x), 0.5f * (vk.y - vm.y));     // column u = 0
            const float2 cn = make_float2(0.5f * (vk.y + vm.y), 0.5f * (vm.x - vk.x));     // column u = 128
            const float2 y0 = cmul(c0, c.t.k[i]);
            const float2 yn = cmul(cn, ld_ro(kny + b + 16 * i));
            c.t.u[i] = make_float2(y0.x - yn.y, y0.y + yn.x);                              // y0 + i yn
        }
    }
}

template <class X>
B200_HD void pconv_init(X& x, const ConvParams& p, int cluster) {
    const int T = planes_of_cluster(cluster, p.B, p.G3);
    if (T == 0) return;
    const int ch = cluster % 3, q = cluster / 3;
    auto stage_src = [&](int t, int rank) { return p.A + (static_cast<size_t>((q + p.G3 * t) * 3 + ch) * NP + 16 * rank) * NP; };
    x.each([&](auto& c) {
        load_twiddles(c, p.tw);
        const int b = c.tid & 15, u = 16 * c.rank + (c.tid >> 4);
        const float2* kcol = p.otf + (static_cast<size_t>(ch) * NC + u) * N;
#pragma unroll
        for (int i = 0; i < 16; ++i) c.t.k[i] = ld_ro(kcol + b + 16 * i);
        if (c.tid == 0) {
            unsigned long long* bar = reinterpret_cast<unsigned long long*>(c.smem + CONV_BAR_OFF);
            c.bulk_init(bar);
            c.bulk_fence_init();
            c.bulk_expect(bar, STAGE_F4 * 16);
            c.bulk_load(c.smem + CONV_STAGE_OFF, stage_src(0, c.rank), STAGE_F4 * 16, bar);
        }
    });
    x.sync_cta();
}

// one step of the pipeline; the kernel runs t = 0 .. T + 1
template <class X>
B200_HD void pconv_step(X& x, const ConvParams& p, int cluster, int t) {
    const int T = planes_of_cluster(cluster, p.B, p.G3);
    const int ch = cluster % 3, q = cluster / 3;
    const bool inverse = p.y != nullptr;
    const bool norm = inverse && p.normalise;
    auto plane_of = [&](int tt) { return (q + p.G3 * tt) * 3 + ch; };
    auto img_of = [&](int tt) { return q + p.G3 * tt; };
    auto stage_src = [&](int tt, int rank) { return p.A + (static_cast<size_t>(plane_of(tt)) * NP + 16 * rank) * NP; };
    {
        // ---- H1(t): column half ------------------------------------------------------------------------------------
        if (t < T) {
            float4* Ap = p.A + static_cast<size_t>(plane_of(t)) * PLANE_F4;
            float4* Bp = p.Bs + (static_cast<size_t>(cluster) * NBUF + t % NBUF) * PLANE_F4;
            x.each([&](auto& c) {
                const int g = c.tid >> 4, a = c.tid & 15;
                c.bulk_wait(reinterpret_cast<unsigned long long*>(c.smem + CONV_BAR_OFF), t & 1);
                const float4* col = reinterpret_cast<const float4*>(c.smem + CONV_STAGE_OFF) + g * NP;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v4 = col[16 * i + a];
                    c.t.v[i] = make_float2(v4.x, v4.y);
                    c.t.v[i + 8] = make_float2(v4.z, v4.w);
                }
            });
            x.sync_cta();                           // the stage has been read: refill it with the next plane's columns
            x.each([&](auto& c) {
                if (c.tid == 0 && t + 1 < T) {
                    unsigned long long* bar = reinterpret_cast<unsigned long long*>(c.smem + CONV_BAR_OFF);
                    c.bulk_expect(bar, STAGE_F4 * 16);
                    c.bulk_load(c.smem + CONV_STAGE_OFF, stage_src(t + 1, c.rank), STAGE_F4 * 16, bar);
                }
                P::stepA(c.t.v, c.tid & 15, c.smem + (c.tid >> 4) * E_GROUP, c.t.w);
            });
            x.sync_warp();
            x.each([&](auto& c) { cols_second(c); });
            x.each([&](auto& c) {
                const int b = c.tid & 15, u = 16 * c.rank + (c.tid >> 4);
                if (p.save) {
                    float4* col = Ap + static_cast<size_t>(u) * NP;
#pragma unroll
                    for (int i = 0; i < 8; ++i) st_cg4(col + b + 16 * i, f4(c.t.v[i], c.t.v[i + 8]));
                }
                if (inverse) conv_multiply(c, p, ch, u);
            });
            x.sync_warp();                          // every lane has read the exchange block and its partner's v
            if (inverse) {
                x.each([&](auto& c) {
                    const int g = c.tid >> 4, b = c.tid & 15;
                    if (16 * c.rank + g == 0) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) c.t.v[i] = c.t.u[i];
                    }
                    P::stepC(c.t.v, b, c.smem + g * E_GROUP, c.t.w);
                });
                x.sync_warp();
                x.each([&](auto& c) {
                    const int g = c.tid >> 4, a = c.tid & 15, u = 16 * c.rank + g;
                    P::stepD(c.t.v, a, c.smem + g * E_GROUP);       // v[i] = Ycol[u][y = 16 i + a]
#pragma unroll
                    for (int i = 0; i < 8; ++i) st_cg4(Bp + static_cast<size_t>(16 * i + a) * NP + u, f4(c.t.v[i], c.t.v[i + 8]));
                });
            }
        }
        if (!inverse) return;
        // ---- F(t-2): the image maximum of plane t-2 was published a whole step ago ----------------------------------
        if (norm && t >= 2) {
            const int s = t - 2, img = img_of(s), plane = plane_of(s);
            x.each([&](auto& c) {
                if ((c.tid & 31) == 0) c.t.key = c.wait_max(p.sync + 2 * img, 3 * C);
            });
            x.each([&](auto& c) {
                const int g = c.tid >> 4, a = c.tid & 15, Pidx = 16 * c.rank + g;
                const float m = key_float(c.lane0_key());                                  // = 2 * max(conv)
                const float inv = 1.0f / m;
                const float2* stash = c.smem + CONV_STASH_OFF;
                float* out0 = p.y + (static_cast<size_t>(plane) * N + Pidx) * N;
                float* out1 = out0 + NP * N;
                if (ch == 0 && c.rank == 0 && c.tid == 0) p.img_max[img] = 0.5f * m;
                unsigned eq = 0u;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float2 e = stash[i * THREADS + c.tid];
                    eq |= (e.x == m ? 1u : 0u) << (2 * i);
                    eq |= (e.y == m ? 1u : 0u) << (2 * i + 1);
                    out0[16 * i + a] = e.x == m ? 1.0f : e.x * inv;                        // arg-max: exactly 1 (torch: x / x)
                    out1[16 * i + a] = e.y == m ? 1.0f : e.y * inv;
                }
                while (eq != 0u) {                                                          // rare: record the arg-max position(s)
                    int bit = 0;
                    while (((eq >> bit) & 1u) == 0u) ++bit;
                    eq &= eq - 1u;
                    const int slot = atomic_add_int(p.tie_count + img, 1);
                    if (slot < MAXT) p.tie_pos[img * MAXT + slot] = ch * N * N + (Pidx + (bit & 1) * NP) * N + 16 * (bit >> 1) + a;
                }
            });
        }
        // ---- wait(t-1), gather, arrive(t) ---------------------------------------------------------------------------
        const bool rows = t >= 1 && t - 1 < T;
        if (rows) {
            const float4* Bp = p.Bs + (static_cast<size_t>(cluster) * NBUF + (t - 1) % NBUF) * PLANE_F4;
            x.cluster_wait();
            x.each([&](auto& c) {
                const int g = c.tid >> 4, b = c.tid & 15;
                const float4* row = Bp + static_cast<size_t>(16 * c.rank + g) * NP;
                float4 qv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int k = b + 16 * i;
                    qv[i] = ld_cg4(row + (k <= NP ? (k & (NP - 1)) : N - k));
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int k = b + 16 * i;
                    float2 z;
                    if (k == 0) z = make_float2(qv[i].x, qv[i].z);
                    else if (k == NP) z = make_float2(qv[i].y, qv[i].w);
                    else if (k < NP) z = make_float2(qv[i].x - qv[i].w, qv[i].y + qv[i].z);   // Y_P[k] + i Y_{P+128}[k]
                    else z = make_float2(qv[i].x + qv[i].w, qv[i].z - qv[i].y);               // conj(Y_P[256-k]) + i conj(Y_{P+128}[256-k])
                    c.t.v[i] = z;
                }
            });
        }
        if (t < T) x.cluster_arrive();
        // ---- H2(t-1): inverse rows, CTA max, stash ------------------------------------------------------------------
        if (rows) {
            const int s = t - 1, img = img_of(s), plane = plane_of(s);
            x.sync_warp();                          // H1(t) has finished with the exchange block
            x.each([&](auto& c) { P::stepC(c.t.v, c.tid & 15, c.smem + (c.tid >> 4) * E_GROUP, c.t.w); });
            x.sync_warp();
            x.each([&](auto& c) {
                const int g = c.tid >> 4, a = c.tid & 15;
                P::stepD(c.t.v, a, c.smem + g * E_GROUP);           // v[i] = (conv_P[16 i + a], conv_{P+128}[16 i + a]) * 2
                if (!norm) {
                    const int Pidx = 16 * c.rank + g;
                    float* out0 = p.y + (static_cast<size_t>(plane) * N + Pidx) * N;
                    float* out1 = out0 + NP * N;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        out0[16 * i + a] = 0.5f * c.t.v[i].x;
                        out1[16 * i + a] = 0.5f * c.t.v[i].y;
                    }
                } else {
                    float m = c.t.v[0].x;
#pragma unroll
                    for (int i = 0; i < 16; ++i) m = fmaxf(m, fmaxf(c.t.v[i].x, c.t.v[i].y));
                    c.t.key = float_key(m);
                    float2* stash = c.smem + CONV_STASH_OFF;
#pragma unroll
                    for (int i = 0; i < 16; ++i) stash[i * THREADS + c.tid] = c.t.v[i];
                }
            });
            if (norm) {
                x.each([&](auto& c) {
                    const unsigned wk = c.warp_max_key();
                    if ((c.tid & 31) == 0) reinterpret_cast<unsigned*>(c.smem + RED_OFF)[c.tid >> 5] = wk;
                });
                x.sync_cta();
                x.each([&](auto& c) {
                    if (c.tid == 0) {
                        const unsigned* red = reinterpret_cast<const unsigned*>(c.smem + RED_OFF);
                        unsigned k = red[0];
                        for (int i = 1; i < THREADS / 32; ++i) k = red[i] > k ? red[i] : k;
                        c.publish_max(p.sync + 2 * img, k);             // atomicMax(key) ; fence ; atomicAdd(arrivals): no wait here
                    }
                });
            }
        }
    }
}

// =====================================================================================================================
// k_pacc: backward accumulation  acc[c][u][v] += conj(X^) G^ / max   and the Parseval partials of sum(g * conv)
//   step t:  R(t)      rows of the upstream gradient of plane t (both 16-row slabs fetched by bulk copies one step
//                       earlier) -> un-mixed -> crossing scratch
//            wait(t-1), gather, arrive(t)
//            C(t-1)     column transform of plane t-1, X^ of that plane from its own bulk-copied stage, accumulate
// =====================================================================================================================
constexpr int ACC_GSTAGE_OFF = CONV_STAGE_OFF;
constexpr int ACC_XSTAGE_OFF = ACC_GSTAGE_OFF + 2 * STAGE_F4;
constexpr int ACC_BAR_OFF = ACC_XSTAGE_OFF + 2 * STAGE_F4;
constexpr int ACC_SMEM_FLOAT2 = ACC_BAR_OFF + 2;
constexpr int ACC_SMEM_BYTES = ACC_SMEM_FLOAT2 * 8;
constexpr int DOT_PER_PLANE = C * (THREADS / 32);       // one partial per warp

struct AccParams {
    const float* g;         // [planes][256][256] upstream gradient dL/dsensor
    const float4* Xh;       // [planes][128][128] X^ left by k_pconv
    const float2* otf;      // [3][129][256]
    float4* As;             // [nclusters][NBUF][128][128] crossing scratch
    const float2* tw;
    const float* img_max;   // [B]; nullptr: no per-image scale and no Parseval partials (plain convolution adjoint)
    float2* partial;        // [G3][3][129][256] out: sum over the images of a cluster, natural spectrum layout
    float* dotp;            // [planes][DOT_PER_PLANE] out: per-warp partials of 4 * sum(g * conv)
    int B;
    int G3;
};

template <class X>
B200_HD void pacc_body(X& x, const AccParams& p, int cluster) {
    const int T = planes_of_cluster(cluster, p.B, p.G3);
    const int ch = cluster % 3, q = cluster / 3;
    auto plane_of = [&](int t) { return (q + p.G3 * t) * 3 + ch; };
    auto img_of = [&](int t) { return q + p.G3 * t; };
    // thread 0: the two 16-row slabs (rows 16 r .. and 128 + 16 r ..) of plane t -> g stage
    auto fetch_g = [&](auto& c, int t) {
        unsigned long long* bar = reinterpret_cast<unsigned long long*>(c.smem + ACC_BAR_OFF);
        const float* src = p.g + static_cast<size_t>(plane_of(t)) * N * N + 16 * c.rank * N;
        float* dst = reinterpret_cast<float*>(c.smem + ACC_GSTAGE_OFF);
        c.bulk_expect(bar, 2 * 16 * N * 4);
        c.bulk_load(dst, src, 16 * N * 4, bar);
        c.bulk_load(dst + 16 * N, src + NP * N, 16 * N * 4, bar);
    };
    auto fetch_x = [&](auto& c, int t) {
        unsigned long long* bar = reinterpret_cast<unsigned long long*>(c.smem + ACC_BAR_OFF) + 1;
        c.bulk_expect(bar, STAGE_F4 * 16);
        c.bulk_load(c.smem + ACC_XSTAGE_OFF, p.Xh + (static_cast<size_t>(plane_of(t)) * NP + 16 * c.rank) * NP, STAGE_F4 * 16, bar);
    };
    x.each([&](auto& c) {
        load_twiddles(c, p.tw);
#pragma unroll
        for (int i = 0; i < 16; ++i) c.t.u[i] = make_float2(0.f, 0.f);
        if (c.tid == 0 && T > 0) {
            unsigned long long* bar = reinterpret_cast<unsigned long long*>(c.smem + ACC_BAR_OFF);
            c.bulk_init(bar);
            c.bulk_init(bar + 1);
            c.bulk_fence_init();
            fetch_g(c, 0);
            fetch_x(c, 0);
        }
    });
    x.sync_cta();
    for (int t = 0; t <= T; ++t) {
        if (t < T) {
            float4* Ap = p.As + (static_cast<size_t>(cluster) * NBUF + t % NBUF) * PLANE_F4;
            x.each([&](auto& c) {
                const int g = c.tid >> 4, a = c.tid & 15;
                c.bulk_wait(reinterpret_cast<unsigned long long*>(c.smem + ACC_BAR_OFF), t & 1);
                const float* r0 = reinterpret_cast<const float*>(c.smem + ACC_GSTAGE_OFF) + g * N;
                const float* r1 = r0 + 16 * N;
#pragma unroll
                for (int i = 0; i < 16; ++i) c.t.v[i] = make_float2(r0[16 * i + a], r1[16 * i + a]);
            });
            x.sync_cta();
            x.each([&](auto& c) {
                if (c.tid == 0 && t + 1 < T) fetch_g(c, t + 1);
                P::stepA(c.t.v, c.tid & 15, c.smem + (c.tid >> 4) * E_GROUP, c.t.w);
            });
            x.sync_warp();
            x.each([&](auto& c) { rows_second(c); });
            x.each([&](auto& c) { rows_unmix_store(c, Ap, 16 * c.rank + (c.tid >> 4)); });
        }
        const bool cols = t >= 1;
        if (cols) {
            const float4* Ap = p.As + (static_cast<size_t>(cluster) * NBUF + (t - 1) % NBUF) * PLANE_F4;
            x.cluster_wait();
            x.sync_warp();                          // R(t) has finished with the exchange block
            x.each([&](auto& c) { cols_load(c, Ap, 16 * c.rank + (c.tid >> 4)); });
        }
        if (t < T) x.cluster_arrive();
        if (cols) {
            const int s = t - 1, img = img_of(s), plane = plane_of(s);
            x.sync_warp();
            x.each([&](auto& c) { cols_second(c); });
            x.each([&](auto& c) {
                const int b = c.tid & 15, u = 16 * c.rank + (c.tid >> 4);
                c.bulk_wait(reinterpret_cast<unsigned long long*>(c.smem + ACC_BAR_OFF) + 1, s & 1);
                const float inv_m = p.img_max != nullptr ? 1.0f / ld_ro(p.img_max + img) : 1.0f;
                const float4* xcol = reinterpret_cast<const float4*>(c.smem + ACC_XSTAGE_OFF) + (c.tid >> 4) * NP;
                const float2* kcol = p.otf + (static_cast<size_t>(ch) * NC + u) * N;
                float2 d2 = make_float2(0.f, 0.f);
                if (u != 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 qv = xcol[b + 16 * i];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int r = i + 8 * h;
                            const float2 xv = h ? make_float2(qv.z, qv.w) : make_float2(qv.x, qv.y);
                            const float2 tt = cmulc(c.t.v[r], xv);                    // G conj(X)
                            c.t.u[r].x += tt.x * inv_m;
                            c.t.u[r].y += tt.y * inv_m;
                            if (p.img_max != nullptr) {
                                const float2 k = ld_ro(kcol + b + 16 * r);
                                d2.x += tt.x * k.x;
                                d2.y += tt.y * k.y;
                            }
                        }
                    }
                    c.t.dot = 2.0f * (d2.x + d2.y);
                } else {
                    const float2* kny = p.otf + (static_cast<size_t>(ch) * NC + NP) * N;
                    const float2* xc2 = reinterpret_cast<const float2*>(xcol);      // element s: float2 index 2*(s & 127) + (s >> 7)
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int e = b + 16 * i, em = (N - e) & (N - 1);
                        const float2 gk = c.t.v[i];
                        const float2 gm = mirror_v(c, i, true);
                        const float2 xk = xc2[2 * (e & (NP - 1)) + (e >> 7)];
                        const float2 xm = xc2[2 * (em & (NP - 1)) + (em >> 7)];
                        const float2 g0 = make_float2(0.5f * (gk.x + gm.x), 0.5f * (gk.y - gm.y));
                        const float2 gn = make_float2(0.5f * (gk.y + gm.y), 0.5f * (gm.x - gk.x));
                        const float2 x0 = make_float2(0.5f * (xk.x + xm.x), 0.5f * (xk.y - xm.y));
                        const float2 xn = make_float2(0.5f * (xk.y + xm.y), 0.5f * (xm.x - xk.x));
                        const float2 t0 = cmulc(g0, x0), tn = cmulc(gn, xn);
                        c.t.u[i].x += (t0.x - tn.y) * inv_m;                        // t0 + i tn: packed like the column itself
                        c.t.u[i].y += (t0.y + tn.x) * inv_m;
                        if (p.img_max != nullptr) {
                            const float2 k0 = ld_ro(kcol + e), kn = ld_ro(kny + e);
                            d2.x += t0.x * k0.x + tn.x * kn.x;
                            d2.y += t0.y * k0.y + tn.y * kn.y;
                        }
                    }
                    c.t.dot = d2.x + d2.y;
                }
            });
            if (p.img_max != nullptr) {
                x.each([&](auto& c) {
                    const float ws = c.warp_sum_dot();
                    if ((c.tid & 31) == 0) p.dotp[static_cast<size_t>(plane) * DOT_PER_PLANE + c.rank * (THREADS / 32) + (c.tid >> 5)] = ws;
                });
            }
            x.sync_cta();                           // the X^ stage has been read: fetch the next plane's
            x.each([&](auto& c) {
                if (c.tid == 0 && s + 1 < T) fetch_x(c, s + 1);
            });
        }
    }
    // the cluster's accumulators -> partial[cluster / 3][cluster % 3][u][v]; column 0 is un-mixed into u = 0 and u = 128
    x.each([&](auto& c) {
        const int b = c.tid & 15, u = 16 * c.rank + (c.tid >> 4);
        float2* base = p.partial + (static_cast<size_t>(cluster / 3) * 3 + cluster % 3) * NC * N;
        if (u != 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) base[static_cast<size_t>(u) * N + b + 16 * i] = c.t.u[i];
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float2 ak = c.t.u[i];
                const float2 am = mirror_u(c, i, true);
                base[b + 16 * i] = make_float2(0.5f * (ak.x + am.x), 0.5f * (ak.y - am.y));
                base[static_cast<size_t>(NP) * N + b + 16 * i] = make_float2(0.5f * (ak.y + am.y), 0.5f * (am.x - ak.x));
            }
        }
    });
}

}  // namespace plane
}  // namespace b200cam
