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
// ---- TMA bulk copies (cp.async.bulk, 1-D) global -> shared with mbarrier completion --------------------------------
// One thread arms the barrier with the byte count and issues the copy; every consumer waits on the barrier's phase
// parity.  The copy engine writes shared memory directly: no registers, no LSU instructions, fully coalesced.
struct BulkOps {
    static __device__ __forceinline__ unsigned saddr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
    static __device__ __forceinline__ void init(unsigned long long* bar) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(saddr(bar)) : "memory");
    }
    static __device__ __forceinline__ void fence_init() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
    // arm the barrier for `bytes` of copies (one arrival: the caller is the barrier's single participant)
    static __device__ __forceinline__ void expect(unsigned long long* bar, unsigned bytes) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(saddr(bar)), "r"(bytes) : "memory");
    }
    // one copy; any thread may issue it, before or after the barrier was armed
    static __device__ __forceinline__ void copy(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");       // earlier generic reads of dst are done
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(saddr(dst)),
                     "l"(src), "r"(bytes), "r"(saddr(bar))
                     : "memory");
    }
    static __device__ __forceinline__ void wait(unsigned long long* bar, unsigned parity) {
        unsigned ok;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok)
                         : "r"(saddr(bar)), "r"(parity)
                         : "memory");
        } while (!ok);
    }
};
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
    __device__ __forceinline__ void bulk_init(unsigned long long* bar) { BulkOps::init(bar); }
    __device__ __forceinline__ void bulk_fence_init() { BulkOps::fence_init(); }
    __device__ __forceinline__ void bulk_expect(unsigned long long* bar, unsigned bytes) { BulkOps::expect(bar, bytes); }
    __device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) { BulkOps::copy(dst, src, bytes, bar); }
    __device__ __forceinline__ void bulk_wait(unsigned long long* bar, unsigned parity) { BulkOps::wait(bar, parity); }
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
    __device__ __forceinline__ void bulk_init(unsigned long long* bar) { BulkOps::init(bar); }
    __device__ __forceinline__ void bulk_fence_init() { BulkOps::fence_init(); }
    __device__ __forceinline__ void bulk_expect(unsigned long long* bar, unsigned bytes) { BulkOps::expect(bar, bytes); }
    __device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) { BulkOps::copy(dst, src, bytes, bar); }
    __device__ __forceinline__ void bulk_wait(unsigned long long* bar, unsigned parity) { BulkOps::wait(bar, parity); }
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
    // emulator: the bulk copy happens at issue time, waiting is a no-op
    void bulk_init(unsigned long long*) {}
    void bulk_fence_init() {}
    void bulk_expect(unsigned long long*, unsigned) {}
    void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long*) { std::memcpy(dst, src, bytes); }
    void bulk_wait(unsigned long long*, unsigned) {}
};

}  // namespace b200cam
