// mbarrier / bulk-copy (TMA unit) / cp.async PTX wrappers shared by the streaming kernels (tp_stream.cu, env_stream.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

// ---- PTX wrappers ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TPS_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TPS_DONE;\n\t"
        "bra TPS_WAIT;\n\t"
        "TPS_DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// Producer-side wait: the producer runs stages ahead, so a slow poll costs nothing -- but a tight try_wait loop on a
// warp that is blocked most of the time steals issue slots from the consumer warps of the same sub-partition.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        __nanosleep(256);
    }
}
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, uint32_t ns) {
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        __nanosleep(ns);
    }
}
// try_wait with a suspend-time hint: the thread may stay suspended up to `ns` before the instruction returns false
__device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(ns)
            : "memory");
        if (ok) return;
    }
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (TMA unit, no descriptor)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
// arrive on the mbarrier once all cp.async issued so far by this thread have landed (count pre-accounted)
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }


// First centre of CTA b's range: the edge stream is cut every E/G edges, snapped forward to the next centre
// boundary.  Two dependent loads (ctr[t], row_ptr[c]) instead of a binary search over row_ptr.
__device__ __forceinline__ int64_t cut_centre(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ ctr, int64_t N, int64_t E,
                                              int64_t b, int64_t G) {
    if (b <= 0) return 0;
    if (b >= G) return N;
    const int64_t t = b * E / G;
    if (t >= E) return N;
    const int64_t c = ctr[t];
    return row_ptr[c] == t ? c : c + 1;
}


// Reduce N (<= 8) per-lane values across the warp: P = next power of two, P/2 + P/4 + .. + 1 shuffles for the
// value-halving steps, then plain butterflies.  On return the lane holds the total of value `idx_of(lane)`.
template <int N>
struct MultiSum {
    static constexpr int P = N <= 1 ? 1 : N <= 2 ? 2 : N <= 4 ? 4 : 8;
    static constexpr int STEPS = P == 1 ? 0 : P == 2 ? 1 : P == 4 ? 2 : 3;
    template <typename T>
    static __device__ __forceinline__ T run(const T (&v)[N], int lane) {
        T a[P];
#pragma unroll
        for (int t = 0; t < P; ++t) a[t] = t < N ? v[t] : T(0);
        int off = 16;
#pragma unroll
        for (int cnt = P; cnt > 1; cnt >>= 1) {
            const int half = cnt >> 1;
            const bool up = lane & off;
#pragma unroll
            for (int t = 0; t < half; ++t) {
                const T send = up ? a[t] : a[t + half];
                const T got = __shfl_xor_sync(0xffffffffu, send, off);
                a[t] = (up ? a[t + half] : a[t]) + got;
            }
            off >>= 1;
        }
        T r = a[0];
#pragma unroll
        for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
        return r;
    }
    // value index held by `lane` (every lane of a group of 32/P lanes holds the same total)
    static __device__ __forceinline__ int idx_of(int lane) {
        int idx = 0;
        if constexpr (STEPS >= 1) idx |= ((lane >> 4) & 1) << (STEPS - 1);
        if constexpr (STEPS >= 2) idx |= ((lane >> 3) & 1) << (STEPS - 2);
        if constexpr (STEPS >= 3) idx |= ((lane >> 2) & 1) << (STEPS - 3);
        return idx;
    }
    static __device__ __forceinline__ bool is_writer(int lane) { return (lane & ((32 >> STEPS) - 1)) == 0; }
};

}  // namespace
