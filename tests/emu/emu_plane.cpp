// TEST INFRASTRUCTURE ONLY - CPU emulator of the plane-resident N=256 sensor kernels (csrc/plane.cuh).
//
// Compiles the very same phase bodies with g++ and runs them with a host execution policy: a "cluster" is 8 emulated
// CTAs x 256 emulated threads, every phase is a loop over all of them, barriers are phase boundaries, warp shuffles read
// the partner thread's state, the cross-cluster image-max exchange is emulated by running the three clusters of an
// image up to their publish step before any of them continues.  Checks index arithmetic / layouts without a GPU
// (tests/test_emulator_plane.py compares against torch.fft).  Never linked into libb200cam.so.
#include <algorithm>
#include <cassert>
#include <cstring>
#include <cmath>
#include <vector>

#include "../../privacy-preserving-vision_b200/csrc/plane.cuh"

using namespace b200cam;
using namespace b200cam::plane;

namespace {

std::vector<float2> make_twiddle(int n) {
    std::vector<float2> tw(n);
    for (int j = 0; j < n; ++j) {
        const double a = -2.0 * M_PI * j / n;
        tw[j] = make_float2(static_cast<float>(cos(a)), static_cast<float>(sin(a)));
    }
    return tw;
}

struct HostCtx {
    int rank, tid;
    Thread& t;
    float2* smem;
    Thread* cta;          // the 256 thread states of this CTA
    float2 shfl_v(int idx, int src, bool) { return cta[(tid & ~15) | src].v[idx]; }
    float2 shfl_u(int idx, int src, bool) { return cta[(tid & ~15) | src].u[idx]; }
    unsigned lane0_key() { return cta[tid & ~31].key; }
    unsigned warp_max_key() {
        unsigned k = 0;
        for (int l = 0; l < 32; ++l) k = cta[(tid & ~31) + l].key > k ? cta[(tid & ~31) + l].key : k;
        return k;
    }
    float warp_sum_dot() {
        float s = 0.f;
        for (int l = 0; l < 32; ++l) s += cta[(tid & ~31) + l].dot;
        return s;
    }
    // bulk copies happen at issue time; barriers are phase boundaries of the emulator
    void bulk_init(unsigned long long*) {}
    void bulk_fence_init() {}
    void bulk_expect(unsigned long long*, unsigned) {}
    void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long*) { std::memcpy(dst, src, bytes); }
    void bulk_wait(unsigned long long*, unsigned) {}
    void publish_max(unsigned* s, unsigned key) {
        if (key > s[0]) s[0] = key;
        s[1] += 1;
    }
    unsigned wait_max(unsigned* s, unsigned target) {
        assert(s[1] == target);
        return s[0];
    }
};

struct HostCluster {
    int nranks;
    std::vector<Thread> threads;
    std::vector<float2> smem;
    int smem_f2;
    explicit HostCluster(int n, int f2 = SMEM_FLOAT2) : nranks(n), threads(static_cast<size_t>(n) * THREADS),
                                                        smem(static_cast<size_t>(n) * f2), smem_f2(f2) {}
    template <class F>
    void each(F&& f) {
        for (int r = 0; r < nranks; ++r)
            for (int tid = 0; tid < THREADS; ++tid) {
                HostCtx c{r, tid, threads[static_cast<size_t>(r) * THREADS + tid], smem.data() + static_cast<size_t>(r) * smem_f2,
                          threads.data() + static_cast<size_t>(r) * THREADS};
                f(c);
            }
    }
    void sync_warp() {}
    void sync_cta() {}
    void sync_cluster() {}
    void cluster_arrive() {}
    void cluster_wait() {}
};

}  // namespace

extern "C" {

int emu_prow(int planes, const float* x, float* A, int nctas) {
    auto tw = make_twiddle(N);
    RowParams p{x, reinterpret_cast<float4*>(A), tw.data(), planes};
    for (int cta = 0; cta < nctas; ++cta) {
        HostCluster ex(1);
        prow_body(ex, p, cta, nctas);
    }
    return 0;
}

// G3 cluster triples; runs the plane loop exactly like the device kernel
int emu_pconv(int B, int G3, float* A, const float* otf, float* y, float* img_max, int* tie_count, int* tie_pos, int save,
              int normalise) {
    auto tw = make_twiddle(N);
    const int G = 3 * G3;
    std::vector<float4> Bs(static_cast<size_t>(G) * NBUF * PLANE_F4);
    std::vector<unsigned> sync(static_cast<size_t>(2) * B, 0u);
    for (int b = 0; b < B; ++b) tie_count[b] = 0;
    ConvParams p{reinterpret_cast<float4*>(A), reinterpret_cast<const float2*>(otf), Bs.data(), y, tw.data(), sync.data(),
                 img_max, tie_count, tie_pos, B, G3, save, normalise};
    std::vector<HostCluster> cl;
    for (int j = 0; j < G; ++j) cl.emplace_back(C, CONV_SMEM_FLOAT2);
    for (int j = 0; j < G; ++j) pconv_init(cl[j], p, j);
    // the device runs every cluster through steps t = 0 .. T + 1; clusters of one image are in lock step, which the
    // emulator reproduces by advancing all clusters one step at a time
    int Tmax = 0;
    for (int j = 0; j < G; ++j) Tmax = std::max(Tmax, planes_of_cluster(j, B, G3));
    for (int t = 0; t <= Tmax + 1; ++t)
        for (int j = 0; j < G; ++j)
            if (planes_of_cluster(j, B, G3) > 0 && t <= planes_of_cluster(j, B, G3) + 1) pconv_step(cl[j], p, j, t);
    return 0;
}

int emu_pacc(int B, int G3, const float* g, const float* Xh, const float* otf, const float* img_max, float* partial, float* dotp) {
    auto tw = make_twiddle(N);
    const int G = 3 * G3;
    std::vector<float4> As(static_cast<size_t>(G) * NBUF * PLANE_F4);
    AccParams p{g, reinterpret_cast<const float4*>(Xh), reinterpret_cast<const float2*>(otf), As.data(), tw.data(), img_max,
                reinterpret_cast<float2*>(partial), dotp, B, G3};
    for (int j = 0; j < G; ++j) {
        HostCluster ex(C, ACC_SMEM_FLOAT2);
        pacc_body(ex, p, j);
    }
    return 0;
}

}  // extern "C"
