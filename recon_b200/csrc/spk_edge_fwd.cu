// K2: fused forward of one KBGAT attention-layer group (all heads of the group in one pass).
//
// Replaces GAT/layers.py:124-175 of the reference (gather/cat edge_h, a.mm(edge_h), a_2.mm,
// LeakyReLU, exp, two SpecialSpmmFinal segmented sums, divide, ELU) for every head at once:
//   m_e = P1[i] + P2[j] + P3[k] (+P3[k2]),  s_e = q1[i] + q2[j] + q3[k] (+q3[k2])
//   ee_e = exp(-LeakyReLU(s_e)), den_i = sum ee_e, num_i = sum msk_e ee_e m_e, out = ELU(num/den)
// One warp owns one CSR segment (aggregation row); lanes own float4 chunks of the row, every
// edge is one 128-bit-per-lane gather of P2[j] (+ L2/L1-resident P3[k]). No atomics: hub rows
// are cut into fixed chunks whose partials are added in a fixed order by a finalize kernel.
// Templated on NCH (float4 chunks per lane), HT (compile-time bound on heads) and HAS2 (2-hop
// edges present) so the per-head state lives in registers without dead slots.
#include "spk_edge.cuh"

namespace spk {

template <int NCH, int HT>
struct FwdAcc {
    float4 acc[NCH];
    float den[HT];
    float sw[HT];
};

template <int NCH, int HT, bool HAS2>
__device__ __forceinline__ void fwd_accumulate(const EdgeFwdArgs& a, int row, int beg, int end, int lane,
                                               const int (&hc)[NCH], FwdAcc<NCH, HT>& st, bool& bad) {
    constexpr int U = (NCH <= 2) ? 4 : 2;
    const int H = a.g.H, Wd4 = a.g.Wd4;
    const int qlane = a.g.Dt4 & 31, qci = a.g.Dt4 >> 5;
    const float4 q1v = ldg4(a.P1 + (long)row * a.ld1 + (long)a.g.Dt4 * 4);
    float q1[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) q1[h] = f4get(q1v, h);
    const bool has_mask = a.mask != nullptr;

    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int my_col = 0, my_t1 = 0, my_t2 = -1;
        float my_m[HT];
#pragma unroll
        for (int h = 0; h < HT; ++h) my_m[h] = 1.f;
        if (lane < n) {
            my_col = __ldg(a.col + base + lane);
            my_t1 = __ldg(a.t1 + base + lane);
            if (HAS2) my_t2 = __ldg(a.t2 + base + lane);
            if (has_mask) {
#pragma unroll
                for (int h = 0; h < HT; ++h)
                    if (h < H) my_m[h] = __ldg(a.mask + (long)h * a.mask_stride + base + lane);
            }
        }
        for (int u0 = 0; u0 < n; u0 += U) {
            float4 v[U][NCH];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = u0 + u;
                const int j = __shfl_sync(0xffffffffu, my_col, idx & 31);
                const int k1 = __shfl_sync(0xffffffffu, my_t1, idx & 31);
                int k2 = -1;
                if (HAS2) k2 = __shfl_sync(0xffffffffu, my_t2, idx & 31);
                const float* p2 = a.P2 + (long)j * a.ld2;
                const float* p3 = a.P3 + (long)k1 * a.ld3;
                const float* p3b = a.P3 + (long)(k2 < 0 ? 0 : k2) * a.ld3;
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    const int c4 = lane + 32 * ci;
                    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < n && c4 < Wd4) {
                        x = f4add(ldg4(p2 + c4 * 4), ldg4(p3 + c4 * 4));
                        if (HAS2 && k2 >= 0) x = f4add(x, ldg4(p3b + c4 * 4));
                    }
                    v[u][ci] = x;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = u0 + u;
                if (idx >= n) break;
                float4 qv = v[u][0];
#pragma unroll
                for (int ci = 1; ci < NCH; ++ci)
                    if (qci == ci) qv = v[u][ci];
                float w[HT];
#pragma unroll
                for (int h = 0; h < HT; ++h) {
                    w[h] = 0.f;
                    if (h < H) {
                        const float s = q1[h] + __shfl_sync(0xffffffffu, f4get(qv, h), qlane);
                        const float ee = expf(-(s > 0.f ? s : a.alpha * s));
                        bad |= (ee != ee);
                        st.den[h] += ee;
                        float m = 1.f;
                        if (has_mask) m = __shfl_sync(0xffffffffu, my_m[h], idx & 31);
                        w[h] = ee * m;
                        st.sw[h] += w[h];
                    }
                }
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) f4fma(st.acc[ci], selh<HT>(hc[ci], w), v[u][ci]);
            }
        }
    }
}

template <int NCH, int HT>
__device__ __forceinline__ void fwd_init(const LayerGeom& g, int lane, int (&hc)[NCH], FwdAcc<NCH, HT>& st) {
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        hc[ci] = (HT > 1 && c4 < g.Dt4) ? c4 / g.Dp4 : 0;
        st.acc[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int h = 0; h < HT; ++h) st.den[h] = st.sw[h] = 0.f;
}

// num = sw*P1[i] + acc ; h = num/den ; ELU ; store (head-concat layout [n_rows, H*D]).
template <int NCH, int HT>
__device__ __forceinline__ void fwd_finalize(const EdgeFwdArgs& a, int row, int lane, const int (&hc)[NCH],
                                             FwdAcc<NCH, HT>& st, bool& bad) {
    const LayerGeom g = a.g;
#pragma unroll
    for (int h = 0; h < HT; ++h)
        if (st.den[h] == 0.f) st.den[h] = 1e-12f;           // layers.py:152
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        if (c4 >= g.Dt4) continue;
        const int h = hc[ci];
        const float d = selh<HT>(h, st.den);
        const float s = selh<HT>(h, st.sw);
        const float4 p1 = ldg4(a.P1 + (long)row * a.ld1 + c4 * 4);
        float o[4] = {fmaf(s, p1.x, st.acc[ci].x), fmaf(s, p1.y, st.acc[ci].y),
                      fmaf(s, p1.z, st.acc[ci].z), fmaf(s, p1.w, st.acc[ci].w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bad |= (o[k] != o[k]);                            // layers.py:167
            o[k] = o[k] / d;                                  // layers.py:169
            bad |= (o[k] != o[k]);                            // layers.py:172
            if (a.apply_elu) o[k] = o[k] > 0.f ? o[k] : expm1f(o[k]);
        }
        const int off = (c4 - h * g.Dp4) * 4;
        float* dst = a.out + (long)row * a.ldo + (long)h * g.D + off;
        if (a.out_vec) {
            *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (off + k < g.D) dst[k] = o[k];
        }
    }
    if (lane < g.H) {
        a.den[(long)row * g.H + lane] = selh<HT>(lane, st.den);
        a.sw[(long)row * g.H + lane] = selh<HT>(lane, st.sw);
    }
}

template <int NCH, int HT, bool HAS2>
__global__ void __launch_bounds__(SPK_CTA_THREADS, (NCH <= 2) ? 3 : 2)
edge_fwd_rows_kernel(const EdgeFwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (row >= a.n_rows) return;
    const int beg = __ldg(a.segptr + row), end = __ldg(a.segptr + row + 1);
    if (end - beg > a.hub.hub_thresh) return;                 // hub: handled by the task kernels
    int hc[NCH];
    FwdAcc<NCH, HT> st;
    bool bad = false;
    fwd_init<NCH, HT>(a.g, lane, hc, st);
    fwd_accumulate<NCH, HT, HAS2>(a, row, beg, end, lane, hc, st, bad);
    fwd_finalize<NCH, HT>(a, row, lane, hc, st, bad);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.nanflag, 1);
}

template <int NCH, int HT, bool HAS2>
__global__ void __launch_bounds__(SPK_CTA_THREADS, (NCH <= 2) ? 3 : 2)
edge_fwd_tasks_kernel(const EdgeFwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int task = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (task >= a.hub.n_tasks) return;
    const int row = __ldg(a.hub.task_seg + task);
    int hc[NCH];
    FwdAcc<NCH, HT> st;
    bool bad = false;
    fwd_init<NCH, HT>(a.g, lane, hc, st);
    fwd_accumulate<NCH, HT, HAS2>(a, row, __ldg(a.hub.task_beg + task), __ldg(a.hub.task_end + task), lane, hc, st, bad);
    float* part = a.hub.partial + (long)task * a.hub.ldpart;
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        if (c4 < a.g.Wd4) *reinterpret_cast<float4*>(part + c4 * 4) = st.acc[ci];
    }
    if (lane < SPK_MAX_HEADS) {
        part[a.g.Wd4 * 4 + lane] = lane < HT ? selh<HT>(lane, st.den) : 0.f;
        part[a.g.Wd4 * 4 + SPK_MAX_HEADS + lane] = lane < HT ? selh<HT>(lane, st.sw) : 0.f;
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.nanflag, 1);
}

// One CTA per hub row: the 8 warps add disjoint, interleaved subsets of the task partials (4 loads in
// flight each), then the 8 warp sums are added in warp order -> fixed summation tree, run-to-run identical.
template <int NCH, int HT>
__global__ void __launch_bounds__(SPK_CTA_THREADS)
edge_fwd_hub_finalize_kernel(const EdgeFwdArgs a) {
    __shared__ __align__(16) float red[SPK_WARPS_PER_CTA][NCH * 128 + 2 * SPK_MAX_HEADS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int hub = blockIdx.x;
    const int row = __ldg(a.hub.hub_seg + hub);
    const int t0 = __ldg(a.hub.hub_task_ptr + hub), t1 = __ldg(a.hub.hub_task_ptr + hub + 1);
    const int len = a.g.Wd4 * 4 + 2 * SPK_MAX_HEADS;         // floats used per partial row
    cta_sum_partials<NCH * 4 + 1>(a.hub.partial, a.hub.ldpart, t0, t1, len, &red[0][0], NCH * 128 + 2 * SPK_MAX_HEADS);
    if (wid != 0) return;
    int hc[NCH];
    FwdAcc<NCH, HT> st;
    bool bad = false;
    fwd_init<NCH, HT>(a.g, lane, hc, st);
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        if (c4 < a.g.Wd4) st.acc[ci] = *reinterpret_cast<const float4*>(&red[0][c4 * 4]);
    }
#pragma unroll
    for (int h = 0; h < HT; ++h) {
        st.den[h] = red[0][a.g.Wd4 * 4 + h];
        st.sw[h] = red[0][a.g.Wd4 * 4 + SPK_MAX_HEADS + h];
    }
    fwd_finalize<NCH, HT>(a, row, lane, hc, st, bad);
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.nanflag, 1);
}

template <int NCH, int HT, bool HAS2>
static int launch_fwd_t(const EdgeFwdArgs& a, cudaStream_t s) {
    if (a.n_rows > 0) {
        const unsigned grid = (a.n_rows + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
        edge_fwd_rows_kernel<NCH, HT, HAS2><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("edge_fwd_rows")) return rc;
    }
    if (a.hub.n_tasks > 0) {
        const unsigned grid = (a.hub.n_tasks + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
        edge_fwd_tasks_kernel<NCH, HT, HAS2><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("edge_fwd_tasks")) return rc;
        edge_fwd_hub_finalize_kernel<NCH, HT><<<a.hub.n_hubs, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("edge_fwd_hub_finalize")) return rc;
    }
    return 0;
}

template <int NCH>
static int launch_fwd_n(const EdgeFwdArgs& a, cudaStream_t s) {
    const bool has2 = a.t2 != nullptr;
    if (a.g.H == 1) return has2 ? launch_fwd_t<NCH, 1, true>(a, s) : launch_fwd_t<NCH, 1, false>(a, s);
    if (a.g.H == 2) return has2 ? launch_fwd_t<NCH, 2, true>(a, s) : launch_fwd_t<NCH, 2, false>(a, s);
    return has2 ? launch_fwd_t<NCH, 4, true>(a, s) : launch_fwd_t<NCH, 4, false>(a, s);
}

int launch_edge_fwd(const EdgeFwdArgs& a, cudaStream_t s) {
    switch ((a.g.Wd4 + 31) / 32) {
        case 1: return launch_fwd_n<1>(a, s);
        case 2: return launch_fwd_n<2>(a, s);
        case 3: return launch_fwd_n<3>(a, s);
        case 4: return launch_fwd_n<4>(a, s);
        default:
            set_error("edge_fwd: row width %d floats exceeds the supported 512", a.g.Wd4 * 4);
            return 2;
    }
}

}  // namespace spk
