// Assembly of the extended weight matrices from the reference-shaped attention parameters, forward and backward.
//   a_h [D, 2F+Rd] = [A1 | A2 | A3],  a_2,h [1, D]   (GAT/layers.py:100-105)
//   mode 0 (projected tables):        Wn [F, 2Wd] = [A1^T | A1^T a_2^T | 0 || A2^T | A2^T a_2^T | 0],  Wr [Rd, Wd] = [A3^T | A3^T a_2^T | 0]
//   mode 1 (aggregate-then-project):  Wa [H, LZ, D] = a_h^T with zero rows at the pad positions of Zn_h,
//                                     V [F, 4] = (q2_0, q2_1, q1_0, q1_1) score vectors of X~,  V3 [Rd, 4] = (q3_0, q3_1, 0, 0)
// These were ~100 tiny eager launches per step (slice assignments + their autograd); here: a zero fill and one kernel
// forward, two kernels backward. One warp per (input column c, head h); sums over lanes in a fixed order: deterministic.
#include "spk_common.cuh"
#include "spk_rowops.cuh"

namespace spk {
namespace {

__device__ __forceinline__ void w_locate(const AttnWeightsArgs& w, int h, int c, float*& vec, float*& q) {
    if (w.mode == 0) {
        const int lo = h * w.Dp;
        if (c < w.F) { float* r = w.W0 + (long)c * w.ld0; vec = r + lo; q = r + w.Dt + h; }
        else if (c < 2 * w.F) { float* r = w.W0 + (long)(c - w.F) * w.ld0 + w.Wd; vec = r + lo; q = r + w.Dt + h; }
        else { float* r = w.W1 + (long)(c - 2 * w.F) * w.ld1; vec = r + lo; q = r + w.Dt + h; }
    } else {
        float* base = w.W0 + (long)h * w.LZ * w.D;
        if (c < w.F) { vec = base + (long)c * w.D; q = w.W1 + c * 4 + 2 + h; }
        else if (c < 2 * w.F) { const int r = c - w.F; vec = base + (long)(w.Fp + r) * w.D; q = w.W1 + r * 4 + h; }
        else { const int r = c - 2 * w.F; vec = base + (long)(2 * w.Fp + r) * w.D; q = w.W2 + r * 4 + h; }
    }
}

__global__ void __launch_bounds__(256)
attn_weights_fwd_kernel(const AttnWeightsArgs w) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int C = 2 * w.F + w.Rd;
    if (item >= C * w.H) return;
    const int h = item / C, c = item - h * C;
    float *vec, *q;
    w_locate(w, h, c, vec, q);
    const float* a = w.a[h];
    const float* a2 = w.a2[h];
    float acc = 0.f;
    for (int d = lane; d < w.D; d += 32) {
        const float v = __ldg(a + (long)d * C + c);
        vec[d] = v;
        acc = fmaf(v, __ldg(a2 + d), acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) *q = acc;
}

// da[d, c] = dvec(c)[d] + a_2[d] * dq(c)
__global__ void __launch_bounds__(256)
attn_weights_bwd_a_kernel(const AttnWeightsArgs w) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int C = 2 * w.F + w.Rd;
    if (item >= C * w.H) return;
    const int h = item / C, c = item - h * C;
    float *vec, *q;
    w_locate(w, h, c, vec, q);
    const float gq = *q;
    const float* a2 = w.a2[h];
    float* da = w.da[h];
    for (int d = lane; d < w.D; d += 32) da[(long)d * C + c] = fmaf(__ldg(a2 + d), gq, vec[d]);
}

// da_2[d] = sum_c a[d, c] * dq(c)
__global__ void __launch_bounds__(256)
attn_weights_bwd_a2_kernel(const AttnWeightsArgs w) {
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (item >= w.D * w.H) return;
    const int h = item / w.D, d = item - h * w.D;
    const int C = 2 * w.F + w.Rd;
    const float* a = w.a[h] + (long)d * C;
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
        float *vec, *q;
        w_locate(w, h, c, vec, q);
        acc = fmaf(__ldg(a + c), *q, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) w.da2[h][d] = acc;
}

}  // namespace

int launch_attn_weights_fwd(const AttnWeightsArgs& w, cudaStream_t s) {
    // zero fill (pad columns / rows), then the scatter
    if (w.mode == 0) {
        if (cudaMemsetAsync(w.W0, 0, sizeof(float) * (size_t)w.F * (size_t)w.ld0, s) != cudaSuccess ||
            cudaMemsetAsync(w.W1, 0, sizeof(float) * (size_t)w.Rd * (size_t)w.ld1, s) != cudaSuccess) return check_launch("attn_weights_zero");
    } else {
        if (cudaMemsetAsync(w.W0, 0, sizeof(float) * (size_t)w.H * w.LZ * w.D, s) != cudaSuccess ||
            cudaMemsetAsync(w.W1, 0, sizeof(float) * (size_t)w.F * 4, s) != cudaSuccess ||
            cudaMemsetAsync(w.W2, 0, sizeof(float) * (size_t)w.Rd * 4, s) != cudaSuccess) return check_launch("attn_weights_zero");
    }
    const int items = (2 * w.F + w.Rd) * w.H;
    attn_weights_fwd_kernel<<<(items + 7) / 8, 256, 0, s>>>(w);
    return check_launch("attn_weights_fwd");
}

int launch_attn_weights_bwd(const AttnWeightsArgs& w, cudaStream_t s) {
    const int items = (2 * w.F + w.Rd) * w.H;
    attn_weights_bwd_a_kernel<<<(items + 7) / 8, 256, 0, s>>>(w);
    if (int rc = check_launch("attn_weights_bwd_a")) return rc;
    attn_weights_bwd_a2_kernel<<<(w.D * w.H + 7) / 8, 256, 0, s>>>(w);
    return check_launch("attn_weights_bwd_a2");
}

}  // namespace spk
