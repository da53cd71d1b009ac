// b200cam: execution policies for phase-structured kernel bodies.
//
// A kernel body is a function template `body(Exec& ex, const Params& p, char* smem, State* st)`
// that calls `ex.phase(nthreads_active, lambda(tid))` repeatedly.  A phase boundary is a block
// barrier.  Per-thread values that must survive a barrier live in `st[ex.slot(tid)]`.
//
//   DeviceExec : one CUDA thread runs each lambda once, `__syncthreads()` after every phase.
//   HostExec   : (test-only emulator) each phase is a loop over all thread ids of the block.
#pragma once

#include "compat.cuh"

namespace b200cam {

#if defined(__CUDACC__)
struct DeviceExec {
    static constexpr bool IS_HOST = false;
    __device__ __forceinline__ int bx() const { return blockIdx.x; }
    __device__ __forceinline__ int by() const { return blockIdx.y; }
    __device__ __forceinline__ int nthreads() const { return blockDim.x; }
    __device__ __forceinline__ int slot(int) const { return 0; }
    template <class F>
    __device__ __forceinline__ void phase(F&& f) {
        f(static_cast<int>(threadIdx.x));
        __syncthreads();
    }
    // a phase whose data exchange stays inside one warp (FFT lane groups never straddle warps)
    template <class F>
    __device__ __forceinline__ void warp_phase(F&& f) {
        f(static_cast<int>(threadIdx.x));
        __syncwarp();
    }
    __device__ __forceinline__ void barrier() { __syncthreads(); }
    __device__ __forceinline__ void threadfence() { __threadfence(); }
    __device__ __forceinline__ float load_cg(const float* p) { return __ldcg(p); }
    __device__ __forceinline__ float4 load_cg4(const float4* p) { return __ldcg(p); }
};
// A "virtual block" of a cooperative kernel: the CTA plays block (vbx, vby) of a body written for
// `nthr` threads; surplus threads only take part in the barriers.
struct VirtualExec {
    static constexpr bool IS_HOST = false;
    int vbx, vby, nthr;
    __device__ __forceinline__ int bx() const { return vbx; }
    __device__ __forceinline__ int by() const { return vby; }
    __device__ __forceinline__ int nthreads() const { return nthr; }
    __device__ __forceinline__ int slot(int) const { return 0; }
    template <class F>
    __device__ __forceinline__ void phase(F&& f) {
        if (static_cast<int>(threadIdx.x) < nthr) f(static_cast<int>(threadIdx.x));
        __syncthreads();
    }
    template <class F>
    __device__ __forceinline__ void warp_phase(F&& f) {
        if (static_cast<int>(threadIdx.x) < nthr) f(static_cast<int>(threadIdx.x));
        __syncwarp();
    }
    __device__ __forceinline__ void barrier() { __syncthreads(); }
    __device__ __forceinline__ void threadfence() { __threadfence(); }
    __device__ __forceinline__ float load_cg(const float* p) { return __ldcg(p); }
    __device__ __forceinline__ float4 load_cg4(const float4* p) { return __ldcg(p); }
};
#endif

struct HostExec {
    static constexpr bool IS_HOST = true;     // per-thread state that crosses a phase needs one slot per thread
    int bx_, by_, nthreads_;
    int bx() const { return bx_; }
    int by() const { return by_; }
    int nthreads() const { return nthreads_; }
    int slot(int tid) const { return tid; }
    template <class F>
    void phase(F&& f) {
        for (int t = 0; t < nthreads_; ++t) f(t);
    }
    template <class F>
    void warp_phase(F&& f) {
        for (int t = 0; t < nthreads_; ++t) f(t);
    }
    void barrier() {}
    void threadfence() {}
    float load_cg(const float* p) { return *p; }
    float4 load_cg4(const float4* p) { return *p; }
};

}  // namespace b200cam
