// Warp-autonomous edge streaming with bulk-async (TMA 1-D) row staging -- shared by K2 / K3 / K4.
//
// A warp owns up to 32 consecutive segments (aggregation rows, columns, or one hub task). Their edges form
// one stream n = 0..T-1. For every edge the warp's lane 0 issues cp.async.bulk copies of the gathered table
// rows (P2~[j], P3~[k], ... 16-byte aligned, contiguous) into a per-warp ring of shared-memory slots, running
// SLOTS edges ahead of consumption; completion is tracked by one mbarrier per slot. The depth of the gather
// pipeline therefore costs shared memory, not registers, and does not drain at segment boundaries.
#pragma once
#include "spk_common.cuh"

namespace spk {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void sbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "SW_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra SW_DONE;\n\t"
        "bra SW_WAIT;\n\t"
        "SW_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
}

// Segment table of a warp: lane l holds (beg, deg, exclusive prefix) of its segment.
struct SegTable {
    int beg, deg, pre, total;
    // entry index (position in the CSR / CSC arrays) of stream element m (m < total); all lanes must call
    __device__ __forceinline__ int entry_of(int m, int& seg) const {
        int lo = 0, hi = 31;
#pragma unroll
        for (int it = 0; it < 5; ++it) {
            const int mid = (lo + hi + 1) >> 1;
            const int pm = __shfl_sync(0xffffffffu, pre, mid);
            if (pm <= m) lo = mid; else hi = mid - 1;
        }
        seg = lo;
        return __shfl_sync(0xffffffffu, beg, lo) + (m - __shfl_sync(0xffffffffu, pre, lo));
    }
};

}  // namespace spk
