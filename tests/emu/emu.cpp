// TEST INFRASTRUCTURE ONLY - CPU emulator of the b200cam kernel bodies.
//
// Compiles privacy-preserving-vision_b200/csrc/kernels.cuh with g++ (no CUDA) and runs every
// kernel body under HostExec: each barrier phase becomes a loop over the block's thread ids, the
// grid becomes a loop over blocks.  This checks the index arithmetic, layouts and launch
// sequencing on a machine without a GPU (tests/test_emulator.py compares against the oracle).
// It is never linked into libb200cam.so and never used by the product.
#include <cstring>
#include <vector>

#include "../../privacy-preserving-vision_b200/csrc/kernels.cuh"

using namespace b200cam;

namespace {

std::vector<float2> make_twiddle(int N) {
    std::vector<float2> tw(N);
    for (int j = 0; j < N; ++j) {
        const double a = -2.0 * M_PI * j / N;
        tw[j] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
    }
    return tw;
}

template <class F>
void grid2(int gx, int gy, int threads, F&& f) {
    for (int y = 0; y < gy; ++y)
        for (int x = 0; x < gx; ++x) {
            HostExec ex{x, y, threads};
            f(ex);
        }
}

constexpr int EW_GRID = 7;   // any grid works for the grid-stride bodies; a small odd one on purpose

int accum_chunks(int N, int B) {          // the library sizes this from the device's occupancy; any value works
    const int colgroups = (3 * (N / 2 + 1) + 7) / 8;
    int n = 444 / colgroups;
    if (n > 16) n = 16;
    if (n > B) n = B;
    if (n < 1) n = 1;
    return n;
}

template <int N>
void otf_impl(const float* psf, float2* otf, const float2* tw) {
    using T = Tile<N>;
    std::vector<float2> smem(RowsR2CSmem<N>::FLOAT2S > ColsSmem<N>::FLOAT2S ? RowsR2CSmem<N>::FLOAT2S : ColsSmem<N>::FLOAT2S);
    grid2(N / T::ROWS, 3, RowsR2CSmem<N>::THREADS, [&](HostExec& ex) {
        rows_r2c_body<N>(ex, RowsR2CParams{psf, otf, tw, nullptr, nullptr}, smem.data());
    });
    const int total = 3 * T::NC;
    grid2((total + T::COLS - 1) / T::COLS, 1, ColsSmem<N>::THREADS, [&](HostExec& ex) {
        cols_fwd_body<N>(ex, ColsFwdParams{otf, tw, total, 1, 1.0f / (static_cast<float>(N) * N), nullptr, 0}, smem.data());
    });
}

int conv_chunks(int N, int B) {
    const int colgroups = (3 * (N / 2 + 1) + 7) / 8;
    int n = 592 / colgroups;
    if (n > 16) n = 16;
    if (n > B) n = B;
    if (n < 1) n = 1;
    return n;
}

template <int N>
int sensor_fwd_impl(int B, const float* img, const float* psf, float* sensor, float* img_max, int* tie_count,
                    int* tie_pos, float2* otf, float2* spectrum) {
    using T = Tile<N>;
    auto tw = make_twiddle(N);
    otf_impl<N>(psf, otf, tw.data());
    const int planes = 3 * B;
    std::vector<float2> stx(static_cast<size_t>(planes) * T::NC * N), st2(static_cast<size_t>(planes) * T::NC * N);
    float2* srow = spectrum != nullptr ? spectrum : stx.data();
    std::vector<float2> smem(RowsR2CSmem<N>::FLOAT2S > ColsSmem<N>::FLOAT2S ? RowsR2CSmem<N>::FLOAT2S : ColsSmem<N>::FLOAT2S);
    {   // the library runs the image rows as a persistent, TMA-staged grid: same here with a small odd grid
        const int total = (N / T::ROWS) * planes, nctas = 5;
        std::vector<float2> ssmem(RowsStreamSmem<N>::FLOAT2S);
        grid2(nctas, 1, RowsStreamSmem<N>::THREADS, [&](HostExec& ex) {
            rows_r2c_stream_body<N>(ex, RowsR2CParams{img, srow, tw.data(), img_max, tie_count}, ssmem.data(), total, nctas);
        });
    }
    const int colgroups = (3 * T::NC + T::COLS - 1) / T::COLS;
    const int nchunks = conv_chunks(N, B);
    std::vector<ConvState<N>> cst(ColsSmem<N>::THREADS);
    grid2(colgroups, nchunks, ColsSmem<N>::THREADS, [&](HostExec& ex) {
        cols_conv_body<N>(ex, ColsConvParams{srow, st2.data(), otf, tw.data(), nullptr, B, nchunks, 0, 1.0f}, smem.data(), cst.data());
    });
    {
        const int total = (N / T::ROWS) * planes, nctas = 7;
        std::vector<float2> ssmem(RowsC2RStreamSmem<N>::FLOAT2S);
        grid2(nctas, 1, RowsC2RStreamSmem<N>::THREADS, [&](HostExec& ex) {
            rows_c2r_stream_body<N>(ex, RowsC2RParams{st2.data(), sensor, tw.data(), img_max, 1.0f, nullptr, nullptr, 0},
                                    ssmem.data(), total, nctas);
        });
    }
    const long long n4 = static_cast<long long>(planes) * N * N / 4;
    grid2(EW_GRID, 1, EW_THREADS, [&](HostExec& ex) {
        normalise_body(ex, NormaliseParams{sensor, img_max, tie_count, tie_pos, n4, 3 * N * N / 4}, EW_GRID);
    });
    return 0;
}

template <int N>
int sensor_bwd_impl(int B, const float* g, const float* img, const float* sensor, const float* img_max,
                    const int* tie_count, const int* tie_pos, const float* psf, const float2* otf,
                    const float2* spectrum, float* grad_psf, float* grad_img) {
    using T = Tile<N>;
    auto tw = make_twiddle(N);
    const int planes = 3 * B, tiles = N / T::ROWS;
    const size_t plane_sz = static_cast<size_t>(T::NC) * N;
    std::vector<float2> stx(planes * plane_sz), stg(planes * plane_sz), stp(3 * plane_sz);
    const int nchunks = accum_chunks(N, B);
    const int used_chunks = nchunks;
    std::vector<float2> partial(static_cast<size_t>(nchunks) * 3 * plane_sz);
    std::vector<float> dot_lanes(static_cast<size_t>(B) * 3 * T::NC * 32), coef(B);
    (void)sensor;
    std::vector<float2> smem(RowsR2CSmem<N>::FLOAT2S > ColsSmem<N>::FLOAT2S ? RowsR2CSmem<N>::FLOAT2S : ColsSmem<N>::FLOAT2S);
    const float2* srow = spectrum;
    if (srow == nullptr) {
        grid2(tiles, planes, RowsR2CSmem<N>::THREADS, [&](HostExec& ex) {
            rows_r2c_body<N>(ex, RowsR2CParams{img, stx.data(), tw.data(), nullptr, nullptr}, smem.data());
        });
        srow = stx.data();
    }
    grid2(tiles, planes, RowsR2CSmem<N>::THREADS, [&](HostExec& ex) {
        rows_r2c_body<N>(ex, RowsR2CParams{g, stg.data(), tw.data(), nullptr, nullptr}, smem.data());
    });
    const int colgroups = (3 * T::NC + T::COLS - 1) / T::COLS;
    std::vector<AccumState<N>> states(ColsSmem<N>::THREADS);
    // the arg-max term of the amax backward is folded into cols_reduce_inv (spectral form, the library's default): the warp
    // partials of the Parseval dot product replace the per-thread ones, the spatial tie_term kernel is not run
    const int dot_count = colgroups * ColsSmem<N>::WARPS;
    std::vector<float> dot_warps(static_cast<size_t>(B) * dot_count);
    grid2(colgroups, used_chunks, ColsSmem<N>::THREADS, [&](HostExec& ex) {
        cols_accum_body<N>(ex, ColsAccumParams{srow, stg.data(), partial.data(), tw.data(), img_max, otf,
                                               dot_lanes.data(), B, nchunks, 0, dot_warps.data()}, smem.data(), states.data());
    });
    std::vector<float2> rsmem(ReduceInvSmem<N>::FLOAT2S);
    const bool side_job = grad_img != nullptr;
    grid2(3 * T::NC + (side_job ? B : 0), 1, ReduceInvSmem<N>::THREADS, [&](HostExec& ex) {
        cols_reduce_inv_body<N>(ex, ColsReduceInvParams{partial.data(), stp.data(), tw.data(), used_chunks,
                                                        1.0f / (static_cast<float>(N) * N), side_job ? dot_lanes.data() : nullptr, img_max,
                                                        tie_count, coef.data(), B, nullptr, tie_pos, srow, dot_warps.data(), dot_count},
                                rsmem.data());
    });
    grid2(tiles, 3, RowsR2CSmem<N>::THREADS, [&](HostExec& ex) {
        rows_c2r_body<N>(ex, RowsC2RParams{stp.data(), grad_psf, tw.data(), nullptr, 1.0f}, smem.data());
    });
    (void)img;
    if (grad_img != nullptr) {
        const int cchunks = conv_chunks(N, B);
        std::vector<ConvState<N>> cst(ColsSmem<N>::THREADS);
        grid2(colgroups, cchunks, ColsSmem<N>::THREADS, [&](HostExec& ex) {
            cols_conv_body<N>(ex, ColsConvParams{stg.data(), stg.data(), otf, tw.data(), img_max, B, cchunks, 1, 1.0f}, smem.data(),
                              cst.data());
        });
        grid2(tiles, planes, RowsR2CSmem<N>::THREADS, [&](HostExec& ex) {
            rows_c2r_body<N>(ex, RowsC2RParams{stg.data(), grad_img, tw.data(), nullptr, 1.0f}, smem.data());
        });
        grid2(EW_GRID, 1, EW_THREADS, [&](HostExec& ex) {
            tie_term_img_body(ex, TieTermImgParams{grad_img, psf, tie_count, tie_pos, coef.data(), B, N}, EW_GRID);
        });
    }
    return 0;
}

struct PsfBuffers {
    std::vector<float2> st;
    std::vector<float> I, gtot, part_rows, part_ew;
    int arrive = 12345;     // deliberately dirty: the pipeline must zero it itself
    explicit PsfBuffers(int N) : st(3 * N * N), I(3 * N * N), gtot(3 * N * N), part_rows(3 * N), part_ew(3 * 1024) {}
};

template <int N>
int psf_fwd_impl(const float* h, const float2* A, const float2* Ht, const float* rho, const float* kappa, float* psf,
                 float2* field, float* stats) {
    using T = Tile<N>;
    auto tw = make_twiddle(N);
    PsfBuffers ws(N);
    std::vector<float2> smem(CRowsSmem<N>::FLOAT2S > CColsSmem<N>::FLOAT2S ? CRowsSmem<N>::FLOAT2S : CColsSmem<N>::FLOAT2S);
    std::vector<float> red(3 * EW_THREADS + 2);
    PupilLoad load{A, h, {kappa[0], kappa[1], kappa[2]}, N};
    grid2(N / T::CROWS, 3, CRowsSmem<N>::THREADS, [&](HostExec& ex) {
        crows_fwd_body<N>(ex, CRowsFwdParams{ws.st.data(), tw.data()}, load, smem.data());
    });
    grid2((N + CColsSmem<N>::CC - 1) / CColsSmem<N>::CC, 1, CColsSmem<N>::THREADS, [&](HostExec& ex) {
        ccols_mix_body<N>(ex, CColsMixParams{ws.st.data(), Ht, tw.data(), 0, 1.0f / (3.0f * N * N)}, smem.data());
    });
    IntensityEpilogue epi{field, ws.I.data(), ws.part_rows.data(), &ws.arrive, N};
    grid2(N / T::CROWS, 3, CRowsSmem<N>::THREADS, [&](HostExec& ex) {
        crows_inv_body<N>(ex, CRowsInvParams{ws.st.data(), tw.data()}, epi, smem.data());
    });
    PsfFinaliseParams fin{ws.I.data(), rho, stats, psf, ws.part_ew.data(), ws.part_rows.data(), &ws.arrive,
                          3 * (N / T::CROWS), N};
    grid2(EW_GRID, 1, EW_THREADS, [&](HostExec& ex) { psf_finalise_body(ex, fin, EW_GRID, red.data()); });
    return ws.arrive == 0 ? 0 : -7;
}

template <int N>
int psf_bwd_impl(const float* gpsf, const float* g_rad, const float* g_cen, const float* h, const float2* A, const float2* Ht,
                 const float* rho, const float* kappa, const float* psf, const float2* field, float* stats,
                 float* grad_h) {
    using T = Tile<N>;
    auto tw = make_twiddle(N);
    PsfBuffers ws(N);
    std::vector<float2> smem(CRowsSmem<N>::FLOAT2S > CColsSmem<N>::FLOAT2S ? CRowsSmem<N>::FLOAT2S : CColsSmem<N>::FLOAT2S);
    std::vector<float> red(3 * EW_THREADS + 2);
    grid2(EW_GRID, 1, EW_THREADS, [&](HostExec& ex) {
        psf_grad_prepare_body(ex, PsfGradPrepParams{gpsf, g_rad, g_cen, psf, rho, stats, ws.gtot.data(), ws.part_ew.data(), N},
                              EW_GRID, red.data());
    });
    GradFieldLoad load{field, ws.gtot.data(), stats, ws.part_ew.data(), nullptr, EW_GRID, N};
    grid2(N / T::CROWS, 3, CRowsSmem<N>::THREADS, [&](HostExec& ex) {
        crows_fwd_body<N>(ex, CRowsFwdParams{ws.st.data(), tw.data()}, load, smem.data());
    });
    grid2((N + CColsSmem<N>::CC - 1) / CColsSmem<N>::CC, 1, CColsSmem<N>::THREADS, [&](HostExec& ex) {
        ccols_mix_body<N>(ex, CColsMixParams{ws.st.data(), Ht, tw.data(), 1, 1.0f / (3.0f * N * N)}, smem.data());
    });
    PupilLoad pupil{A, h, {kappa[0], kappa[1], kappa[2]}, N};
    std::vector<float2> hsmem(HGradSmem<N>::FLOAT2S);
    grid2(N / T::CROWS, 1, HGradSmem<N>::THREADS, [&](HostExec& ex) {
        crows_inv_hgrad_body<N>(ex, CRowsInvParams{ws.st.data(), tw.data()}, pupil, grad_h, hsmem.data());
    });
    return 0;
}

}  // namespace

#define DISPATCH_N(N_, CALL)                                   \
    switch (N_) {                                              \
        case 64: { constexpr int NN_ = 64; return CALL; }      \
        case 128: { constexpr int NN_ = 128; return CALL; }    \
        case 256: { constexpr int NN_ = 256; return CALL; }    \
        case 512: { constexpr int NN_ = 512; return CALL; }    \
        case 1024: { constexpr int NN_ = 1024; return CALL; }  \
        default: return -1;                                    \
    }

extern "C" {

int emu_fft(int N, int inverse, const float* in, float* out) {
    // one N-point FFT through stepA..D with 'LANES' emulated lanes; exercises Plan<N> in isolation
    auto run = [&](auto tag) -> int {
        constexpr int NN_ = decltype(tag)::value;
        using P = Plan<NN_>;
        auto tw = make_twiddle(NN_);
        std::vector<float2> E(P::E_SIZE);
        const float2* x = reinterpret_cast<const float2*>(in);
        float2* y = reinterpret_cast<float2*>(out);
        if (!inverse) {
            for (int a = 0; a < P::R2; ++a) {
                float2 v[P::R1];
                for (int i = 0; i < P::R1; ++i) v[i] = x[P::R2 * i + a];
                P::stepA(v, a, E.data(), tw.data());
            }
            for (int b = 0; b < P::R1; ++b) {
                float2 v[P::R2];
                P::stepB(v, b, E.data());
                for (int i = 0; i < P::R2; ++i) y[b + P::R1 * i] = v[i];
            }
        } else {
            for (int b = 0; b < P::R1; ++b) {
                float2 v[P::R2];
                for (int i = 0; i < P::R2; ++i) v[i] = x[b + P::R1 * i];
                P::stepC(v, b, E.data(), tw.data());
            }
            for (int a = 0; a < P::R2; ++a) {
                float2 v[P::R1];
                P::stepD(v, a, E.data());
                for (int i = 0; i < P::R1; ++i) y[P::R2 * i + a] = v[i];
            }
        }
        return 0;
    };
    switch (N) {
        case 64: return run(std::integral_constant<int, 64>{});
        case 128: return run(std::integral_constant<int, 128>{});
        case 256: return run(std::integral_constant<int, 256>{});
        case 512: return run(std::integral_constant<int, 512>{});
        case 1024: return run(std::integral_constant<int, 1024>{});
        default: return -1;
    }
}

int emu_psf_fwd(int N, const float* h, const float* A, const float* Ht, const float* rho, const float* kappa,
                float* psf, float* field, float* stats) {
    DISPATCH_N(N, (psf_fwd_impl<NN_>(h, reinterpret_cast<const float2*>(A), reinterpret_cast<const float2*>(Ht), rho,
                                     kappa, psf, reinterpret_cast<float2*>(field), stats)));
}

int emu_psf_bwd(int N, const float* gpsf, const float* g_rad, const float* g_cen, const float* h, const float* A, const float* Ht,
                const float* rho, const float* kappa, const float* psf, const float* field, float* stats,
                float* grad_h) {
    DISPATCH_N(N, (psf_bwd_impl<NN_>(gpsf, g_rad, g_cen, h, reinterpret_cast<const float2*>(A),
                                     reinterpret_cast<const float2*>(Ht), rho, kappa, psf,
                                     reinterpret_cast<const float2*>(field), stats, grad_h)));
}

int emu_sensor_fwd(int N, int B, const float* img, const float* psf, float* sensor, float* img_max, int* tie_count,
                   int* tie_pos, float* otf, float* spectrum) {
    DISPATCH_N(N, (sensor_fwd_impl<NN_>(B, img, psf, sensor, img_max, tie_count, tie_pos, reinterpret_cast<float2*>(otf),
                                        reinterpret_cast<float2*>(spectrum))));
}

int emu_sensor_bwd(int N, int B, const float* g, const float* img, const float* sensor, const float* img_max,
                   const int* tie_count, const int* tie_pos, const float* psf, const float* otf, const float* spectrum,
                   float* grad_psf, float* grad_img) {
    DISPATCH_N(N, (sensor_bwd_impl<NN_>(B, g, img, sensor, img_max, tie_count, tie_pos, psf,
                                        reinterpret_cast<const float2*>(otf), reinterpret_cast<const float2*>(spectrum),
                                        grad_psf, grad_img)));
}

// Zernike projection bodies: forward with an odd K split, adjoint
int emu_zernike(int T, int NN, const float* coef, const float* Z, float* h, const float* gh, float* gcoef) {
    const int NN4 = NN / 4, xblocks = (NN4 + EW_THREADS - 1) / EW_THREADS, KS = T < 3 ? 1 : 3;
    std::vector<float4> partial(static_cast<size_t>(KS) * NN4);
    std::vector<int> arrive(xblocks, 0);
    int flag = 0;
    grid2(xblocks, KS, EW_THREADS, [&](HostExec& ex) {
        zernike_fwd_body(ex, ZernikeFwdParams{coef, reinterpret_cast<const float4*>(Z), partial.data(),
                                              reinterpret_cast<float4*>(h), arrive.data(), T, NN4, KS}, &flag);
    });
    for (int a : arrive) if (a != 0) return -9;      // counters must be left at zero
    std::vector<float> red(EW_THREADS);
    grid2(T, 1, EW_THREADS, [&](HostExec& ex) {
        zernike_bwd_body(ex, ZernikeBwdParams{reinterpret_cast<const float4*>(gh), reinterpret_cast<const float4*>(Z), gcoef, NN4},
                         red.data());
    });
    return 0;
}

}  // extern "C"
