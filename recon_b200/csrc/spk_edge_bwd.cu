// K3 / K4: backward of the fused attention-layer group (closed form of SURVEY.md 8 a-5; the
// reference gets it from autograd over GAT/layers.py:124-175 + SpecialSpmmFunctionFinal.backward
// layers.py:67-79).
//
// K3 (rows, CSR order; one warp per aggregation row i):
//   dh = dOut * ELU'(h),  dnum = dh/den,  dden = -(dh.h)/den
//   per edge: t = dnum . m_e,  w = msk*ee,  ds = -(msk*t + dden) * ee * LeakyReLU'(s)
//   emits rec[e] = (w, ds) per head, G[i] = dnum, dP1~[i] = [ (sum w) * dnum | sum ds | 0 ]
// K4 (segments keyed on edge[1] = CSC, or on relation id): out[seg] = [ sum w*G[row_e] | sum ds | 0 ]
//   which is dP2~ (per gathered node) resp. dP3~ (per relation).
// Both are deterministic: fixed edge order inside a segment, hub segments split into chunks whose
// partials are added in chunk order.
#include <stdlib.h>
#include "spk_edge.cuh"
#include "spk_edge_bwd.cuh"

namespace spk {

template <int NCH, int HT, bool HAS2>
__device__ __forceinline__ void bwd_row_edges(const EdgeBwdRowsArgs& a, int beg, int end, int lane,
                                              const RowCtx<NCH, HT>& rc, float (&usum)[HT],
                                              float (&swsum)[HT]) {
    constexpr int U = (NCH <= 2) ? 4 : 2;
    const LayerGeom g = a.g;
    const int qlane = g.Dt4 & 31, qci = g.Dt4 >> 5;
    const bool has_mask = a.mask != nullptr;
    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int my_col = 0, my_t1 = 0, my_t2 = -1;
        float my_m[HT], my_w[HT], my_ds[HT];
#pragma unroll
        for (int h = 0; h < HT; ++h) { my_m[h] = 1.f; my_w[h] = 0.f; my_ds[h] = 0.f; }
        if (lane < n) {
            my_col = __ldg(a.col + base + lane);
            my_t1 = __ldg(a.t1 + base + lane);
            if (HAS2) my_t2 = __ldg(a.t2 + base + lane);
            if (has_mask) {
#pragma unroll
                for (int h = 0; h < HT; ++h)
                    if (h < g.H) my_m[h] = __ldg(a.mask + (long)h * a.mask_stride + base + lane);
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
                    if (idx < n && c4 < g.Wd4) {
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
                float pd[HT];
#pragma unroll
                for (int h = 0; h < HT; ++h) pd[h] = 0.f;
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    const float d = f4dot(rc.dnum[ci], v[u][ci]);     // dnum is 0 on q / pad chunks
#pragma unroll
                    for (int h = 0; h < HT; ++h) pd[h] += (HT == 1 || rc.hc[ci] == h) ? d : 0.f;
                }
#pragma unroll
                for (int h = 0; h < HT; ++h) {
                    if (h < g.H) {
                        const float t = rc.c1[h] + warp_sum(pd[h]);
                        const float s = rc.q1[h] + __shfl_sync(0xffffffffu, f4get(qv, h), qlane);
                        const float slope = s > 0.f ? 1.f : a.alpha;
                        const float ee = expf(-(s * slope));
                        float m = 1.f;
                        if (has_mask) m = __shfl_sync(0xffffffffu, my_m[h], idx & 31);
                        const float w = ee * m;
                        const float ds = -(m * t + rc.dden[h]) * ee * slope;
                        usum[h] += ds;
                        swsum[h] += w;
                        if (lane == idx) { my_w[h] = w; my_ds[h] = ds; }
                    }
                }
            }
        }
        if (lane < n) {
            float* r = a.rec + (long)(base + lane) * (2 * g.H);
#pragma unroll
            for (int h = 0; h < HT; ++h)
                if (h < g.H) *reinterpret_cast<float2*>(r + 2 * h) = make_float2(my_w[h], my_ds[h]);
        }
    }
}

template <int NCH, int HT, bool HAS2>
__global__ void __launch_bounds__(SPK_CTA_THREADS, (NCH <= 2) ? 3 : 2)
edge_bwd_rows_kernel(const EdgeBwdRowsArgs a) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (row >= a.n_rows) return;
    const int beg = __ldg(a.segptr + row), end = __ldg(a.segptr + row + 1);
    if (end - beg > a.hub.hub_thresh) return;
    RowCtx<NCH, HT> rc;
    bwd_row_prologue<NCH, HT>(a, row, lane, rc);
    float usum[HT], swsum[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) usum[h] = swsum[h] = 0.f;
    bwd_row_edges<NCH, HT, HAS2>(a, beg, end, lane, rc, usum, swsum);
    bwd_row_store<NCH, HT>(a, row, lane, rc, usum, swsum, true, true);
}

template <int NCH, int HT, bool HAS2>
__global__ void __launch_bounds__(SPK_CTA_THREADS, (NCH <= 2) ? 3 : 2)
edge_bwd_rows_tasks_kernel(const EdgeBwdRowsArgs a) {
    const int lane = threadIdx.x & 31;
    const int task = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (task >= a.hub.n_tasks) return;
    const int row = __ldg(a.hub.task_seg + task);
    RowCtx<NCH, HT> rc;
    bwd_row_prologue<NCH, HT>(a, row, lane, rc);
    float usum[HT], swsum[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) usum[h] = swsum[h] = 0.f;
    bwd_row_edges<NCH, HT, HAS2>(a, __ldg(a.hub.task_beg + task), __ldg(a.hub.task_end + task), lane, rc, usum, swsum);
    if (lane < SPK_MAX_HEADS) {
        float* part = a.hub.partial + (long)task * a.hub.ldpart;
        part[lane] = lane < HT ? selh<HT>(lane, usum) : 0.f;
        part[SPK_MAX_HEADS + lane] = lane < HT ? selh<HT>(lane, swsum) : 0.f;
    }
}

// one CTA per hub row; see edge_fwd_hub_finalize_kernel for the summation order
template <int NCH, int HT>
__global__ void __launch_bounds__(SPK_CTA_THREADS)
edge_bwd_rows_hub_finalize_kernel(const EdgeBwdRowsArgs a) {
    __shared__ float red[SPK_WARPS_PER_CTA][2 * SPK_MAX_HEADS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int hub = blockIdx.x;
    const int row = __ldg(a.hub.hub_seg + hub);
    const int t0 = __ldg(a.hub.hub_task_ptr + hub), t1 = __ldg(a.hub.hub_task_ptr + hub + 1);
    cta_sum_partials<1>(a.hub.partial, a.hub.ldpart, t0, t1, 2 * SPK_MAX_HEADS, &red[0][0], 2 * SPK_MAX_HEADS);
    if (wid != 0) return;
    RowCtx<NCH, HT> rc;
    bwd_row_prologue<NCH, HT>(a, row, lane, rc);
    float usum[HT], swsum[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) { usum[h] = red[0][h]; swsum[h] = red[0][SPK_MAX_HEADS + h]; }
    bwd_row_store<NCH, HT>(a, row, lane, rc, usum, swsum, true, true);
}

template <int NCH, int HT, bool HAS2>
static int launch_bwd_rows_t(const EdgeBwdRowsArgs& a, cudaStream_t s) {
    if (a.n_rows > 0) {
        const unsigned grid = (a.n_rows + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
        edge_bwd_rows_kernel<NCH, HT, HAS2><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("edge_bwd_rows")) return rc;
    }
    if (a.hub.n_tasks > 0) {
        const unsigned grid = (a.hub.n_tasks + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
        edge_bwd_rows_tasks_kernel<NCH, HT, HAS2><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("edge_bwd_rows_tasks")) return rc;
        edge_bwd_rows_hub_finalize_kernel<NCH, HT><<<a.hub.n_hubs, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("edge_bwd_rows_hub_finalize")) return rc;
    }
    return 0;
}

template <int NCH>
static int launch_bwd_rows_n(const EdgeBwdRowsArgs& a, cudaStream_t s) {
    const bool has2 = a.t2 != nullptr;
    if (a.g.H == 1) return has2 ? launch_bwd_rows_t<NCH, 1, true>(a, s) : launch_bwd_rows_t<NCH, 1, false>(a, s);
    if (a.g.H == 2) return has2 ? launch_bwd_rows_t<NCH, 2, true>(a, s) : launch_bwd_rows_t<NCH, 2, false>(a, s);
    return has2 ? launch_bwd_rows_t<NCH, 4, true>(a, s) : launch_bwd_rows_t<NCH, 4, false>(a, s);
}

template <int NCH>
static int launch_bwd_fin_n(const EdgeBwdRowsArgs& a, cudaStream_t s) {
    if (a.g.H == 1) edge_bwd_rows_hub_finalize_kernel<NCH, 1><<<a.hub.n_hubs, SPK_CTA_THREADS, 0, s>>>(a);
    else if (a.g.H == 2) edge_bwd_rows_hub_finalize_kernel<NCH, 2><<<a.hub.n_hubs, SPK_CTA_THREADS, 0, s>>>(a);
    else edge_bwd_rows_hub_finalize_kernel<NCH, 4><<<a.hub.n_hubs, SPK_CTA_THREADS, 0, s>>>(a);
    return check_launch("edge_bwd_rows_hub_finalize");
}

int launch_edge_bwd_rows_hub_finalize(const EdgeBwdRowsArgs& a, cudaStream_t s) {
    switch ((a.g.Wd4 + 31) / 32) {
        case 1: return launch_bwd_fin_n<1>(a, s);
        case 2: return launch_bwd_fin_n<2>(a, s);
        case 3: return launch_bwd_fin_n<3>(a, s);
        default: return launch_bwd_fin_n<4>(a, s);
    }
}

int launch_edge_bwd_rows(const EdgeBwdRowsArgs& a, cudaStream_t s) {
    const int rc = launch_edge_bwd_rows_stream(a, s);
    if (rc >= 0) return rc;
    switch ((a.g.Wd4 + 31) / 32) {
        case 1: return launch_bwd_rows_n<1>(a, s);
        case 2: return launch_bwd_rows_n<2>(a, s);
        case 3: return launch_bwd_rows_n<3>(a, s);
        case 4: return launch_bwd_rows_n<4>(a, s);
        default: set_error("edge_bwd_rows: row width %d floats exceeds the supported 512", a.g.Wd4 * 4); return 2;
    }
}

// ------------------------------------------------------------------------------------------------
// K4: out[seg] = [ sum_e w_e,h * G[src_e] | sum_e ds_e,h | 0 ]
// ------------------------------------------------------------------------------------------------
template <int NCH>
struct SegAcc {
    float4 acc[NCH];
    float vs[SPK_MAX_HEADS];     // lane-local partial sums of ds (reduced across the warp at the end)
};

template <int NCH, int U>
__device__ __forceinline__ void seg_accumulate(const SegGatherArgs& a, int beg, int end, int lane,
                                               const int (&hc)[NCH], SegAcc<NCH>& st) {
    const LayerGeom g = a.g;
    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int my_src = 0;
        float my_w[SPK_MAX_HEADS] = {0.f, 0.f, 0.f, 0.f};
        if (lane < n) {
            my_src = __ldg(a.src + base + lane);
            const int p = __ldg(a.pos + base + lane);
            const float* r = a.rec + (long)p * (2 * g.H);
#pragma unroll
            for (int h = 0; h < SPK_MAX_HEADS; ++h) {
                if (h < g.H) {
                    const float2 wd = __ldg(reinterpret_cast<const float2*>(r + 2 * h));
                    my_w[h] = wd.x;
                    st.vs[h] += wd.y;
                }
            }
        }
        for (int u0 = 0; u0 < n; u0 += U) {
            float4 v[U][NCH];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = u0 + u;
                const int i = __shfl_sync(0xffffffffu, my_src, idx & 31);
                const float* gp = a.G + (long)i * a.ldg;
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    const int c4 = lane + 32 * ci;
                    v[u][ci] = (idx < n && c4 < g.Dt4) ? ldg4(gp + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = u0 + u;
                if (idx >= n) break;
                float w[SPK_MAX_HEADS];
#pragma unroll
                for (int h = 0; h < SPK_MAX_HEADS; ++h)
                    w[h] = h < g.H ? __shfl_sync(0xffffffffu, my_w[h], idx & 31) : 0.f;
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci)
                    f4fma(st.acc[ci], sel4(hc[ci], w[0], w[1], w[2], w[3]), v[u][ci]);
            }
        }
    }
}

template <int NCH>
__device__ __forceinline__ void seg_init(const LayerGeom& g, int lane, int (&hc)[NCH], SegAcc<NCH>& st) {
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        hc[ci] = c4 < g.Dt4 ? c4 / g.Dp4 : 0;
        st.acc[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int h = 0; h < SPK_MAX_HEADS; ++h) st.vs[h] = 0.f;
}

// vs must already be warp-reduced (identical in all lanes)
template <int NCH>
__device__ __forceinline__ void seg_store(float* dst, const LayerGeom& g, int lane, const SegAcc<NCH>& st) {
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        if (c4 >= g.Wd4) continue;
        float4 o = st.acc[ci];
        if (c4 == g.Dt4) o = make_float4(st.vs[0], st.vs[1], st.vs[2], st.vs[3]);
        else if (c4 > g.Dt4) o = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(dst + c4 * 4) = o;
    }
}

template <int NCH, int U, int MINB>
__global__ void __launch_bounds__(SPK_CTA_THREADS, MINB)
seg_gather_kernel(const SegGatherArgs a) {
    const int lane = threadIdx.x & 31;
    const int seg = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (seg >= a.n_seg) return;
    const int beg = __ldg(a.segptr + seg), end = __ldg(a.segptr + seg + 1);
    if (end - beg > a.hub.hub_thresh) return;
    int hc[NCH];
    SegAcc<NCH> st;
    seg_init<NCH>(a.g, lane, hc, st);
    seg_accumulate<NCH, U>(a, beg, end, lane, hc, st);
#pragma unroll
    for (int h = 0; h < SPK_MAX_HEADS; ++h) st.vs[h] = warp_sum(st.vs[h]);
    seg_store<NCH>(a.outp + (long)seg * a.ldout, a.g, lane, st);
}

template <int NCH, int U, int MINB>
__global__ void __launch_bounds__(SPK_CTA_THREADS, MINB)
seg_gather_tasks_kernel(const SegGatherArgs a) {
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (slot >= a.hub.n_tasks) return;
    const int task = hub_task_of_slot(a.hub, slot);
    int hc[NCH];
    SegAcc<NCH> st;
    seg_init<NCH>(a.g, lane, hc, st);
    seg_accumulate<NCH, U>(a, __ldg(a.hub.task_beg + task), __ldg(a.hub.task_end + task), lane, hc, st);
#pragma unroll
    for (int h = 0; h < SPK_MAX_HEADS; ++h) st.vs[h] = warp_sum(st.vs[h]);
    seg_store<NCH>(a.hub.partial + (long)task * a.hub.ldpart, a.g, lane, st);
}

template <int NCH>
__global__ void __launch_bounds__(SPK_CTA_THREADS)
seg_gather_hub_finalize_kernel(const SegGatherArgs a) {
    __shared__ __align__(16) float red[SPK_WARPS_PER_CTA][NCH * 128];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int hub = blockIdx.x;
    const int seg = __ldg(a.hub.hub_seg + hub);
    const int t0 = __ldg(a.hub.hub_task_ptr + hub), t1 = __ldg(a.hub.hub_task_ptr + hub + 1);
    cta_sum_partials<NCH * 4>(a.hub.partial, a.hub.ldpart, t0, t1, a.g.Wd4 * 4, &red[0][0], NCH * 128);
    if (wid != 0) return;
    float* dst = a.outp + (long)seg * a.ldout;             // partial rows already have the output format
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        if (c4 < a.g.Wd4) *reinterpret_cast<float4*>(dst + c4 * 4) = *reinterpret_cast<const float4*>(&red[0][c4 * 4]);
    }
}

// tuning variant (unroll depth U, min CTAs per SM): SPK_SEG_VARIANT=0..3, default chosen from B200 measurements
static int seg_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SPK_SEG_VARIANT"); v = e ? atoi(e) : 1; if (v < 0 || v > 6) v = 1; }
    return v;
}

template <int NCH>
static int launch_seg_t(const SegGatherArgs& a, cudaStream_t s) {
    constexpr int U0 = (NCH <= 2) ? 4 : 2;
    const int var = seg_variant();
    if (a.n_seg > 0) {
        const unsigned grid = (a.n_seg + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
        if (var == 1 && NCH == 1) seg_gather_kernel<NCH, U0, 6><<<grid, SPK_CTA_THREADS, 0, s>>>(a);   // narrow rows: 6 CTAs/SM measured best
        else if (var == 1) seg_gather_kernel<NCH, U0, 4><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else if (var == 4) seg_gather_kernel<NCH, U0, 5><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else if (var == 5) seg_gather_kernel<NCH, U0, 6><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else if (var == 6) seg_gather_kernel<NCH, (U0 > 2 ? U0 / 2 : U0), 6><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else if (var == 2) seg_gather_kernel<NCH, 2 * U0, 2><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else if (var == 3) seg_gather_kernel<NCH, 2 * U0, 3><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else seg_gather_kernel<NCH, U0, 3><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("seg_gather")) return rc;
    }
    if (a.hub.n_tasks > 0) {
        const unsigned grid = (a.hub.n_tasks + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
        if (var == 1) seg_gather_tasks_kernel<NCH, U0, 4><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else if (var == 4) seg_gather_tasks_kernel<NCH, U0, 5><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else if (var == 5) seg_gather_tasks_kernel<NCH, U0, 6><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else if (var == 6) seg_gather_tasks_kernel<NCH, (U0 > 2 ? U0 / 2 : U0), 6><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else if (var == 2) seg_gather_tasks_kernel<NCH, 2 * U0, 2><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else if (var == 3) seg_gather_tasks_kernel<NCH, 2 * U0, 3><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        else seg_gather_tasks_kernel<NCH, U0, 3><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("seg_gather_tasks")) return rc;
        seg_gather_hub_finalize_kernel<NCH><<<a.hub.n_hubs, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("seg_gather_hub_finalize")) return rc;
    }
    return 0;
}

int launch_seg_gather_hub_finalize(const SegGatherArgs& a, cudaStream_t s) {
    switch ((a.g.Wd4 + 31) / 32) {
        case 1: seg_gather_hub_finalize_kernel<1><<<a.hub.n_hubs, SPK_CTA_THREADS, 0, s>>>(a); break;
        case 2: seg_gather_hub_finalize_kernel<2><<<a.hub.n_hubs, SPK_CTA_THREADS, 0, s>>>(a); break;
        case 3: seg_gather_hub_finalize_kernel<3><<<a.hub.n_hubs, SPK_CTA_THREADS, 0, s>>>(a); break;
        default: seg_gather_hub_finalize_kernel<4><<<a.hub.n_hubs, SPK_CTA_THREADS, 0, s>>>(a); break;
    }
    return check_launch("seg_gather_hub_finalize");
}

int launch_seg_gather_stream(const SegGatherArgs& a, cudaStream_t s);   // spk_seg_stream.cu; -1 = not supported

int launch_seg_gather(const SegGatherArgs& a, cudaStream_t s) {
    const int src = launch_seg_gather_stream(a, s);
    if (src >= 0) return src;
    const int nch = (a.g.Wd4 + 31) / 32;
    switch (nch) {
        case 1: return launch_seg_t<1>(a, s);
        case 2: return launch_seg_t<2>(a, s);
        case 3: return launch_seg_t<3>(a, s);
        case 4: return launch_seg_t<4>(a, s);
        default: set_error("seg_gather: row width %d floats exceeds the supported 512", a.g.Wd4 * 4); return 2;
    }
}

}  // namespace spk
