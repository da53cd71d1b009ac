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

// ---- tensor memory as a thread-private stash -------------------------------------------------------------------
// TMEM (256 KB per SM) is addressed [lane][column]; a warp reaches the 32 lanes of its own quadrant, so "32 columns of
// my lane" is 128 bytes of storage private to one thread that costs neither registers nor shared memory.  The fused
// inverse-row + normalise kernel parks a tile's outputs there until the image maximum is known.
struct Tmem {
    static __device__ __forceinline__ void alloc(unsigned* smem_slot, unsigned ncols) {          // one warp, converged
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                         static_cast<unsigned>(__cvta_generic_to_shared(smem_slot))), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    static __device__ __forceinline__ void dealloc(unsigned base, unsigned ncols) {               // the same warp
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(base), "r"(ncols) : "memory");
    }
    static __device__ __forceinline__ void st32(unsigned taddr, const float2 (&v)[16]) {
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
            "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),
            "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y),
            "f"(v[4].x), "f"(v[4].y), "f"(v[5].x), "f"(v[5].y), "f"(v[6].x), "f"(v[6].y), "f"(v[7].x), "f"(v[7].y),
            "f"(v[8].x), "f"(v[8].y), "f"(v[9].x), "f"(v[9].y), "f"(v[10].x), "f"(v[10].y), "f"(v[11].x), "f"(v[11].y),
            "f"(v[12].x), "f"(v[12].y), "f"(v[13].x), "f"(v[13].y), "f"(v[14].x), "f"(v[14].y), "f"(v[15].x), "f"(v[15].y)
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
    }
    static __device__ __forceinline__ void ld32(unsigned taddr, float2 (&v)[16]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
            "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y),
              "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y), "=f"(v[7].x), "=f"(v[7].y),
              "=f"(v[8].x), "=f"(v[8].y), "=f"(v[9].x), "=f"(v[9].y), "=f"(v[10].x), "=f"(v[10].y), "=f"(v[11].x),
              "=f"(v[11].y), "=f"(v[12].x), "=f"(v[12].y), "=f"(v[13].x), "=f"(v[13].y), "=f"(v[14].x), "=f"(v[14].y),
              "=f"(v[15].x), "=f"(v[15].y)
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    }
};
// cross-CTA hand-off without acquire fences (an acquire at gpu scope makes ptxas emit CCTL.IVALL - the SM's whole L1 is
// invalidated): release on the writer, L2-coherent relaxed accesses on the reader
struct HandOff {
    static __device__ __forceinline__ void arrive(int* counter) {          // earlier writes of this thread, then +1
        asm volatile("red.release.gpu.global.add.s32 [%0], 1;\n" ::"l"(counter) : "memory");
    }
    // lane 0 of the warp polls, every lane gets the answer; false = gave up after ~2 s
    static __device__ __forceinline__ bool wait_count(const int* counter, int target) {
        int ok = 1;
        if ((threadIdx.x & 31) == 0) {
            int n;
            long long t0 = 0;
            for (int spin = 0;; ++spin) {
                asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(n) : "l"(counter) : "memory");
                if (n >= target) break;
                if (spin == 64) t0 = clock64();
                if (spin > 64 && clock64() - t0 > 4000000000LL) { ok = 0; break; }
            }
        }
        return __shfl_sync(0xffffffffu, ok, 0) != 0;
    }
    static __device__ __forceinline__ int peek(const int* counter) {       // one relaxed L2 read, no waiting
        int n;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(n) : "l"(counter) : "memory");
        return n;
    }
    // "the value is its own flag": a 4-byte word that is zero until its producer has stored a non-zero key - no fence, no
    // atomic, no separate counter (the low-latency protocol of the peer all-reduce, used here between CTAs)
    static __device__ __forceinline__ void store_key(unsigned* p, unsigned key) {
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(key) : "memory");
    }
    static __device__ __forceinline__ unsigned load_key(const unsigned* p) {
        unsigned v;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
        return v;
    }
    // order-preserving, never-zero key of a float (not NaN) and back
    static __device__ __forceinline__ unsigned float_key(float v) {
        const unsigned b = __float_as_uint(v);
        return (b & 0x80000000u) ? ~b : (b | 0x80000000u);      // negative: ~b in [0x007fffff, 0x7fffffff]; never 0 (b = ~0 is a NaN)
    }
    static __device__ __forceinline__ float key_float(unsigned k) {
        return __uint_as_float((k & 0x80000000u) ? k & 0x7fffffffu : ~k);
    }
    // all lanes: poll the slot row until every live word is non-zero, return the decoded maximum; false = gave up after ~2 s
    static __device__ __forceinline__ bool wait_keys(const unsigned* row, int lane, unsigned k0, unsigned k1, float* out) {
        bool ok = true;
        long long t0 = 0;
        for (int spin = 0;; ++spin) {
            if (__all_sync(0xffffffffu, k0 != 0u && k1 != 0u)) break;
            if (k0 == 0u) k0 = load_key(row + lane);
            if (k1 == 0u) k1 = load_key(row + 32 + lane);
            if (spin == 64) t0 = clock64();
            if (spin > 64 && clock64() - t0 > 4000000000LL) { ok = false; break; }
        }
        unsigned k = k0 > k1 ? k0 : k1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned other = __shfl_xor_sync(0xffffffffu, k, o);
            k = other > k ? other : k;
        }
        *out = key_float(k);
        return ok;
    }
    static __device__ __forceinline__ float load_float(const float* p) {
        float v;
        asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];\n" : "=f"(v) : "l"(p) : "memory");
        return v;
    }
};
struct DeviceExec {
    static constexpr bool IS_HOST = false;
    unsigned tmem = 0;       // base address of this CTA's TMEM allocation (kernels that stash), else unused
    unsigned* err = nullptr; // device error word
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
    // thread-private stash of up to 16 complex values (see Tmem)
    template <int R>
    __device__ __forceinline__ void stash_put(const float2 (&v)[R]) {
        static_assert(R <= 16, "stash holds 16 complex values per thread");
        float2 w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = i < R ? v[i] : make_float2(0.f, 0.f);
        Tmem::st32(tmem + (((threadIdx.x >> 5) & 3u) << 21), w);
    }
    template <int R>
    __device__ __forceinline__ void stash_get(float2 (&v)[R]) {
        float2 w[16];
        Tmem::ld32(tmem + (((threadIdx.x >> 5) & 3u) << 21), w);
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = w[i];
    }
    __device__ __forceinline__ void arrive(int* counter) { HandOff::arrive(counter); }
    __device__ __forceinline__ bool wait_count(const int* counter, int target) { return HandOff::wait_count(counter, target); }
    __device__ __forceinline__ float load_coherent(const float* p) { return HandOff::load_float(p); }
    __device__ __forceinline__ int peek_count(const int* counter) { return HandOff::peek(counter); }
    __device__ __forceinline__ void key_store(unsigned* p, float v) { HandOff::store_key(p, HandOff::float_key(v)); }
    __device__ __forceinline__ unsigned key_load(const unsigned* p) { return HandOff::load_key(p); }
    __device__ __forceinline__ bool key_wait(const unsigned* row, int lane, unsigned k0, unsigned k1, float* out) {
        return HandOff::wait_keys(row, lane, k0, k1, out);
    }
    // block maximum, first half: the warp's maximum lands in red[warp] (the emulator keeps one slot per thread)
    __device__ __forceinline__ void stage_max(float v, float* red, int tid) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((tid & 31) == 0) red[tid >> 5] = v;
    }
    __device__ __forceinline__ int staged(int nthreads) const { return nthreads / 32; }
    // sum over the warp (fixed tree), lane 0 stores; every lane of the warp must call it
    __device__ __forceinline__ void warp_sum_store(float v, float* dst, int tid) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) *dst = v;
    }
    __device__ __forceinline__ void report(unsigned code) { if (err != nullptr) atomicExch(err, code); }
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
    template <int R> __device__ __forceinline__ void stash_put(const float2 (&)[R]) {}
    template <int R> __device__ __forceinline__ void stash_get(float2 (&)[R]) {}
    __device__ __forceinline__ void arrive(int* counter) { HandOff::arrive(counter); }
    __device__ __forceinline__ bool wait_count(const int* counter, int target) { return HandOff::wait_count(counter, target); }
    __device__ __forceinline__ float load_coherent(const float* p) { return HandOff::load_float(p); }
    __device__ __forceinline__ void report(unsigned) {}
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
    // the fused inverse-row + normalise mode exchanges maxima between concurrently running CTAs: device only
    template <int R> void stash_put(const float2 (&)[R]) {}
    template <int R> void stash_get(float2 (&)[R]) {}
    void arrive(int* counter) { *counter += 1; }
    bool wait_count(const int*, int) { return true; }
    float load_coherent(const float* p) { return *p; }
    int peek_count(const int* counter) { return *counter; }
    void key_store(unsigned*, float) {}                      // the one-pass exchange between concurrent CTAs is device only
    unsigned key_load(const unsigned*) { return 1u; }
    bool key_wait(const unsigned*, int, unsigned, unsigned, float* out) { *out = 1.f; return true; }
    void stage_max(float v, float* red, int tid) { red[tid] = v; }
    int staged(int nthreads) const { return nthreads; }
    void warp_sum_store(float v, float* dst, int tid) {       // the phase loop visits the lanes of a warp in order
        if ((tid & 31) == 0) *dst = 0.f;
        *dst += v;
    }
    void report(unsigned) {}
};

}  // namespace b200cam
