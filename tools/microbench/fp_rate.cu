// Microbenchmark: issue rate of scalar vs packed (f32x2) FP32 instructions on sm_100a, alone and mixed with LDS.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp_rate fp_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int ITERS = 2048;
constexpr int ACC = 8;

template <int MODE>
__global__ void __launch_bounds__(512) k_rate(float2* out, float2 seed, int iters) {
    __shared__ float2 sm[1024];
    sm[threadIdx.x] = seed; sm[threadIdx.x + 512] = seed;
    __syncthreads();
    float2 a[ACC];
#pragma unroll
    for (int i = 0; i < ACC; ++i) a[i] = make_float2(seed.x + i, seed.y - i + threadIdx.x);
    const float2 w = make_float2(seed.y, seed.x);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < ACC; ++i) {
                if (MODE == 0) { a[i].x += w.x; a[i].y += w.y; }                       // 2 scalar FADD
                if (MODE == 1) { a[i] = __fadd2_rn(a[i], w); }                         // 1 FADD2
                if (MODE == 2) { a[i].x = fmaf(a[i].x, w.x, w.y); a[i].y = fmaf(a[i].y, w.x, w.y); }   // 2 FFMA
                if (MODE == 3) { a[i] = __ffma2_rn(a[i], w, w); }                      // 1 FFMA2
                if (MODE == 4) { a[i] = __ffma2_rn(make_float2(a[i].y, a[i].x), w, w); }  // FFMA2 with swap modifier
                if (MODE == 5) { a[i] = __fadd2_rn(a[i], w); if ((i & 1) == 0) a[i].x += sm[(threadIdx.x + it * 4 + r + i) & 1023].x; }  // FADD2 + LDS per 2
                if (MODE == 6) { a[i].x += w.x; a[i].y += w.y; if ((i & 1) == 0) a[i].x += sm[(threadIdx.x + it * 4 + r + i) & 1023].x; }
                if (MODE == 7) { a[i].x = fmaf(a[i].x, w.x, w.y); a[i].y += w.y; }     // FFMA + FADD mix
            }
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < ACC; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int ops_per_inner, float2* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 2;
    k_rate<MODE><<<grid, 512>>>(out, make_float2(1.0001f, 0.9999f), 16);
    cudaEventRecord(e0);
    k_rate<MODE><<<grid, 512>>>(out, make_float2(1.0001f, 0.9999f), ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = double(grid) * 16 * ITERS * 4 * ACC * ops_per_inner;   // FP warp-instructions
    const double per_clk_sm = warp_instr / (ms * 1e-3) / 1.965e9 / 148;
    printf("%-28s %8.3f ms  %6.3f FP warp-instr/clk/SM (at 1965 MHz)  %s\n", name, ms, per_clk_sm, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float2* out; cudaMalloc(&out, sizeof(float2) * 148 * 2 * 512);
    run<0>("scalar FADD x2", 2, out);
    run<1>("FADD2", 1, out);
    run<2>("scalar FFMA x2", 2, out);
    run<3>("FFMA2", 1, out);
    run<4>("FFMA2 swap", 1, out);
    run<5>("FADD2 + LDS/2", 1, out);
    run<6>("FADD x2 + LDS/2", 2, out);
    run<7>("FFMA + FADD", 2, out);
    return 0;
}
