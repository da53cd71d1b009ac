// Microbenchmark 2: (a) bulk (TMA engine) shared -> peer-shared copies inside a cluster, (b) L2-resident read / read+write
// bandwidth seen by plain kernels.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bulk_probe dsmem_bulk_probe.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ unsigned mapa(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }

extern __shared__ __align__(128) unsigned char smem[];

// every CTA owns [C chunks of CHUNK bytes] of outgoing data and the same of incoming data; per round it sends chunk d to CTA d
// (slot `rank` there) with ONE bulk copy per destination, then waits on its own mbarrier for the C-1 incoming chunks.
template <int CHUNK>
__global__ void __launch_bounds__(256) k_bulk(float* out, int rounds, int C) {
    const unsigned rank = cg::this_cluster().block_rank();
    const int tid = threadIdx.x;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem);
    unsigned char* outgoing = smem + 128;
    unsigned char* incoming = outgoing + 16 * CHUNK;
    for (int i = tid; i < 16 * CHUNK / 4; i += 256) reinterpret_cast<float*>(outgoing)[i] = i + rank;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    cluster_arrive(); cluster_wait();
    float acc = 0.f;
    for (int r = 0; r < rounds; ++r) {
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"((C - 1) * CHUNK) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        if (tid < C && tid != static_cast<int>(rank)) {
            const unsigned dst = mapa(smem_u32(incoming + rank * CHUNK), tid);
            const unsigned rbar = mapa(smem_u32(bar), tid);
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                         "r"(smem_u32(outgoing + tid * CHUNK)), "r"(CHUNK), "r"(rbar)
                         : "memory");
        }
        unsigned ok;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok) : "r"(smem_u32(bar)), "r"(r & 1) : "memory");
        } while (!ok);
        acc += reinterpret_cast<float*>(incoming)[(tid * 13 + r) & (CHUNK / 4 - 1)];
        cluster_arrive(); cluster_wait();     // peers have consumed `incoming` before the next round overwrites it
    }
    if (acc == 12345.678f) out[blockIdx.x * 256 + tid] = acc;
}

template <int CHUNK>
static void run_bulk(int C, int ctas_per_sm, float* out) {
    const int smem_bytes = 128 + 32 * CHUNK;
    cudaFuncSetAttribute(k_bulk<CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    cudaFuncSetAttribute(k_bulk<CHUNK>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(C);
    int nclusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_bulk<CHUNK>, &cfg);
    if (e != cudaSuccess) { printf("bulk C=%d: occupancy query failed: %s\n", C, cudaGetErrorString(e)); (void)cudaGetLastError(); return; }
    int want = 148 * ctas_per_sm / C;
    if (want > nclusters) want = nclusters;
    cfg.gridDim = dim3(want * C);
    const int rounds = 400;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, k_bulk<CHUNK>, out, 8, C);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, k_bulk<CHUNK>, out, rounds, C);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    e = cudaGetLastError();
    const double clk = ms * 1e-3 * 1.965e9;
    const double remote = double(want) * C * (C - 1) * CHUNK * rounds;
    printf("bulk copy C=%2d chunk=%5d B smem=%6d max_clusters=%4d launched=%4d (%.2f CTA/SM): %8.3f ms, %7.1f clk/round, remote %6.2f B/clk/SM  %s\n",
           C, CHUNK, smem_bytes, nclusters, want, double(want) * C / 148.0, ms, clk / rounds, remote / clk / 148.0, cudaGetErrorString(e));
}

// ---- L2 bandwidth -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_l2_read(const float4* buf, size_t n4, int iters, float* out) {
    float4 acc = make_float4(0, 0, 0, 0);
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (int it = 0; it < iters; ++it) {
        for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < n4; i += 4 * stride) {
            const float4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
            acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y; acc.z += a.z + b.z; acc.w += c.w + d.w;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}
__global__ void __launch_bounds__(256) k_l2_rw(float4* buf, size_t n4, int iters) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (int it = 0; it < iters; ++it) {
        for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < n4; i += 4 * stride) {
            float4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
            a.x += 1.f; b.x += 1.f; c.x += 1.f; d.x += 1.f;
            buf[i] = a; buf[i + stride] = b; buf[i + 2 * stride] = c; buf[i + 3 * stride] = d;
        }
    }
}

static void run_l2(size_t mbytes, float* out) {
    const size_t bytes = mbytes << 20, n4 = bytes / 16;
    float4* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 50, grid = 148 * 8;
    float ms;
    k_l2_read<<<grid, 256>>>(buf, n4, 2, out);
    cudaEventRecord(e0); k_l2_read<<<grid, 256>>>(buf, n4, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("L2 read      %4zu MB x %d: %8.3f ms  %7.2f TB/s\n", mbytes, iters, ms, double(bytes) * iters / (ms * 1e-3) / 1e12);
    k_l2_rw<<<grid, 256>>>(buf, n4, 2);
    cudaEventRecord(e0); k_l2_rw<<<grid, 256>>>(buf, n4, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("L2 read+write%4zu MB x %d: %8.3f ms  %7.2f TB/s (read + written bytes)\n", mbytes, iters, ms, 2.0 * double(bytes) * iters / (ms * 1e-3) / 1e12);
    cudaFree(buf);
}

int main() {
    float* out; cudaMalloc(&out, sizeof(float) * 148 * 8 * 256);
    run_bulk<4096>(8, 3, out);
    run_bulk<4096>(8, 1, out);
    run_bulk<4096>(4, 3, out);
    run_bulk<4096>(2, 3, out);
    run_bulk<2048>(8, 3, out);
    run_bulk<2048>(16, 3, out);
    run_bulk<1024>(8, 3, out);
    for (size_t mb : {16, 32, 64, 96, 256, 1024}) run_l2(mb, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
