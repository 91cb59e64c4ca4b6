// Shared device/host helpers for libspkbgat (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#define SPK_MAX_HEADS 4          // heads fused per edge-kernel launch (q columns live in one float4)
#define SPK_WARPS_PER_CTA 8
#define SPK_CTA_THREADS (SPK_WARPS_PER_CTA * 32)

namespace spk {

void set_error(const char* fmt, ...);
int check_launch(const char* what);   // returns 0 or error code after a kernel launch

// Raises a kernel's dynamic shared-memory limit. The attribute is per device, so the cache is keyed by the current
// device id; entries only grow and the driver call is idempotent, so concurrent callers at worst repeat it.
struct SmemLimit {
    size_t set[64] = {};
    template <class K>
    cudaError_t ensure(K kernel, size_t bytes) {
        int dev = 0;
        const bool keyed = cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64;
        if (keyed && set[dev] >= bytes) return cudaSuccess;
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess && keyed) set[dev] = bytes;
        return e;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum each of the N per-lane values over the 32 lanes with N (not 5N) shuffles: at every stage a lane keeps half
// of its values and hands the other half to its partner. Lane L ends with the total of value (L * N) >> 5.
template <int N>
__device__ __forceinline__ float transposed_warp_sum(float (&p)[N], int lane) {
    int off = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
        const int half = n >> 1;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const float keep = up ? p[k + half] : p[k];
            const float send = up ? p[k] : p[k + half];
            p[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    float r = p[0];
    for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}

// log1p(x) on (-1, 0] as log(u) * x / (u - 1), u = fl(1 + x): the rounding of u cancels to first order (|rel err| ~ 2e-7)
__device__ __forceinline__ float fast_log1p(float x) {
    const float u = 1.0f + x;
    const float d = u - 1.0f;
    return d == 0.f ? x : __logf(u) * __fdividef(x, d);
}

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming (read-once) 128-bit load: do not allocate in L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 f4add(float4 a, float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ void f4fma(float4& acc, float w, float4 v) {
    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
}
__device__ __forceinline__ float f4dot(float4 a, float4 b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float f4get(const float4& v, int k) {
    return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}
__device__ __forceinline__ float sel4(int h, float a, float b, float c, float d) {
    return h == 0 ? a : (h == 1 ? b : (h == 2 ? c : d));
}

template <int HT>
__device__ __forceinline__ float selh(int h, const float (&v)[HT]) {
    float r = v[0];
#pragma unroll
    for (int i = 1; i < HT; ++i) r = (h == i) ? v[i] : r;
    return r;
}

// CTA-wide (NW warps, default 8) deterministic sum of the partial rows [t0, t1) of `part` (row stride ld, `len`
// floats used). Warp w adds rows t0+w, t0+w+8, ... in ascending order with 4 rows of loads in flight,
// then the 8 warp sums are added in warp order. Result in red[0 .. len). NREG >= ceil(len / 32).
template <int NREG, int NW = SPK_WARPS_PER_CTA>
__device__ __forceinline__ void cta_sum_partials(const float* __restrict__ part, long ld, int t0, int t1, int len,
                                                 float* red, int red_ld) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float acc[NREG];
#pragma unroll
    for (int k = 0; k < NREG; ++k) acc[k] = 0.f;
    int t = t0 + wid;
    for (; t + 3 * NW < t1; t += 4 * NW) {
        float v[4][NREG];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < NREG; ++k) {
                const int c = lane + 32 * k;
                v[u][k] = c < len ? part[(long)(t + u * NW) * ld + c] : 0.f;
            }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < NREG; ++k) acc[k] += v[u][k];
    }
    for (; t < t1; t += NW) {
#pragma unroll
        for (int k = 0; k < NREG; ++k) {
            const int c = lane + 32 * k;
            if (c < len) acc[k] += part[(long)t * ld + c];
        }
    }
#pragma unroll
    for (int k = 0; k < NREG; ++k) {
        const int c = lane + 32 * k;
        if (c < len) red[wid * red_ld + c] = acc[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < len; c += (NW * 32)) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += red[w * red_ld + c];
        red[c] = s;
    }
    __syncthreads();
}

// Geometry of one fused attention-layer group. Rows of the projected tables are
//   [ H*Dp floats of projection | H score scalars q | zero pad ]  (width Wd, multiple of 8)
// Dp = D rounded up to 4 so a head never straddles a float4 chunk.
struct LayerGeom {
    int H;        // heads in this launch (<= SPK_MAX_HEADS)
    int D;        // out_features per head
    int Dp4;      // Dp / 4
    int Dt4;      // H * Dp / 4 : number of projection chunks; chunk Dt4 holds the q scalars
    int Wd4;      // Wd / 4 : chunks per table row
};

}  // namespace spk
