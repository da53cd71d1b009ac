// Microbenchmark: distributed-shared-memory transposition cost on sm_100a, as used by the plane-resident sensor kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_probe dsmem_probe.cu
//
// A cluster of C CTAs (256 threads each) repeats: every thread issues NV vector stores of W bytes to the shared memory of
// cluster CTAs (i % C), then a cluster barrier.  Reports bytes / clk / SM through DSMEM and the cost of the bare barrier.
// Also prints the co-residency the occupancy API reports for the shapes the kernels want and whether a cooperative
// launch can be combined with cluster dimensions.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(unsigned addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(unsigned addr, float2 v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};\n" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }

extern __shared__ __align__(16) unsigned char smem[];

// MODE 0: barrier only; 1: v4 stores (8 per thread per round, one per destination i % C); 2: v2 stores (16 per thread per round)
// 3: the same v4 stores but all to the CTA's OWN shared memory through the cluster window (local reference)
template <int MODE>
__global__ void __launch_bounds__(256) k_probe(float* out, int rounds, int C) {
    const unsigned rank = cg::this_cluster().block_rank();
    const int tid = threadIdx.x, g = tid >> 4, b = tid & 15;
    float* buf = reinterpret_cast<float*>(smem);
    buf[tid] = tid;
    __syncthreads();
    cluster_arrive(); cluster_wait();
    // column buffer: 16 columns x pitch 548 floats; thread (g,b) writes float4 at column b, row pair (32*rank + 2g)
    const unsigned base = smem_u32(buf) + 4096;
    float acc = 0.f;
    for (int r = 0; r < rounds; ++r) {
        if (MODE == 1 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned dst = (MODE == 3) ? rank : static_cast<unsigned>(i % C);
                const unsigned a = mapa(base + (b * 548 + (32 * ((rank + i) & 7) + 2 * g) * 2) * 4, dst);
                st_cluster_v4(a, make_float4(r, i, tid, acc));
            }
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const unsigned dst = static_cast<unsigned>((i >> 1) % C);
                const int pair = (b >> 1) + 8 * (i & 1);
                const unsigned a = mapa(base + (pair * 560 + ((g ^ (pair & 7)) + 16 * (rank & 7)) * 4 + (b & 1) * 2) * 4, dst);
                st_cluster_v2(a, make_float2(r, acc));
            }
        }
        cluster_arrive();
        cluster_wait();
        acc += buf[1024 + ((tid * 7 + r) & 1023)];
    }
    if (acc == 12345.678f) out[blockIdx.x * 256 + tid] = acc;
}

template <int MODE>
static void run(const char* name, int C, int smem_bytes, int ctas_per_sm, float* out, bool coop = false) {
    cudaFuncSetAttribute(k_probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    cudaFuncSetAttribute(k_probe<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = coop ? 2 : 1;
    cfg.gridDim = dim3(C);
    int nclusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_probe<MODE>, &cfg);
    if (e != cudaSuccess) { printf("%-34s C=%2d smem=%6d: occupancy query failed: %s\n", name, C, smem_bytes, cudaGetErrorString(e)); (void)cudaGetLastError(); return; }
    int want = 148 * ctas_per_sm / C;
    if (want > nclusters) want = nclusters;
    cfg.gridDim = dim3(want * C);
    const int rounds = 400;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    e = cudaLaunchKernelEx(&cfg, k_probe<MODE>, out, 8, C);
    if (e != cudaSuccess) { printf("%-34s C=%2d: launch failed: %s\n", name, C, cudaGetErrorString(e)); (void)cudaGetLastError(); return; }
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    e = cudaLaunchKernelEx(&cfg, k_probe<MODE>, out, rounds, C);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    e = cudaGetLastError();
    const double bytes_per_cta_round = (MODE == 0) ? 0.0 : 256.0 * 128.0;
    const double ctas = double(want) * C;
    const double clk = ms * 1e-3 * 1.965e9;
    printf("%-34s C=%2d smem=%6d coop=%d max_clusters=%4d launched=%4d (%.2f CTA/SM): %8.3f ms, %7.1f clk/round, %6.2f B/clk/SM  %s\n",
           name, C, smem_bytes, coop ? 1 : 0, nclusters, want, ctas / 148.0, ms, clk / rounds,
           bytes_per_cta_round * ctas * rounds / clk / 148.0, cudaGetErrorString(e));
}

int main() {
    float* out; cudaMalloc(&out, sizeof(float) * 148 * 8 * 256);
    const int SM70 = 72 * 1024, SM105 = 108 * 1024, SM36 = 40 * 1024;
    for (int C : {2, 4, 8, 16}) {
        run<0>("barrier only", C, SM70, 3, out);
        run<1>("v4 scatter (8 x 16 B / thread)", C, SM70, 3, out);
        run<2>("v2 scatter (16 x 8 B / thread)", C, SM70, 3, out);
    }
    run<3>("v4 to own smem via cluster window", 8, SM70, 3, out);
    run<1>("v4 scatter, 1 CTA/SM", 8, SM70, 1, out);
    run<1>("v4 scatter, 2 CTA/SM", 8, SM105, 2, out);
    run<2>("v2 scatter, 2 CTA/SM", 8, SM105, 2, out);
    run<1>("v4 scatter, 5 CTA/SM", 8, SM36, 5, out);
    run<1>("v4 scatter coop+cluster", 8, SM70, 3, out, true);
    run<1>("v4 scatter coop+cluster 2/SM", 8, SM105, 2, out, true);
    return 0;
}
