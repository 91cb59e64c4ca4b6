// K3'/K4': backward of one projected attention-layer group with the per-edge work done ONCE, in column (CSC) order.
//
// Same math as spk_edge_bwd.cu (closed form of the autograd of GAT/layers.py:124-175 and of
// SpecialSpmmFunctionFinal.backward, layers.py:67-79), different schedule. The row-major pass K3
// gathers P2~[j] for every edge only to form t_e = dnum_i . m_e, and the column pass K4 then gathers dnum_i for the
// same edge again. Here the edge is visited once, from its gathered node j:
//   node pass   (warp per row i)     : dnum_i = dOut*ELU'(h)/den -> G[i];  dden_i, c1_i = dnum_i . P1[i], q1_i -> rowsc[i];
//                                      dP1~[i] = [ sw_i * dnum_i | . | 0 ]
//   column pass (warp per column j)  : P2~[j] stays in registers; per edge gather G[i] (+ the L2-resident P3~[k]):
//                                      t = c1_i + G[i] . (P2[j] + P3[k]),  ee = exp(-LeakyReLU(q1_i + q2_j + q3_k)),
//                                      w = msk*ee,  ds = -(msk*t + dden_i) * ee * LeakyReLU'(s);  rec[e] = (w, ds);
//                                      dP2~[j] = [ sum w*G[i] | sum ds | 0 ]
//   row sums    (warp per row i)     : u_i = sum of ds over the row's (contiguous, CSR-ordered) records -> dP1~[i] q slot
// The relation pass (dP3~) is the unchanged K4 over the relation segments. One 4*Dt-byte row gather per edge is gone
// together with the whole edge loop of K3. Deterministic: fixed edge order inside a column, hub columns as fixed
// chunks added in chunk order, row sums by a fixed lane assignment.
#include <stdlib.h>
#include "spk_edge.cuh"
#include "spk_edge_bwd.cuh"
#include "spk_edge_bwd_fused.cuh"

namespace spk {
namespace {

constexpr unsigned FULLM = 0xffffffffu;
__device__ __forceinline__ float fexp(float x) { return exp2f(x * 1.4426950408889634f); }

// ---- node pass -------------------------------------------------------------------------------------
template <int NCH, int HT>
__global__ void __launch_bounds__(SPK_CTA_THREADS)
bwd_node_kernel(const EdgeBwdRowsArgs a, const float* __restrict__ sw, float* __restrict__ rowsc) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (row >= a.n_rows) return;
    if (__ldg(a.segptr + row) == __ldg(a.segptr + row + 1)) {      // no edges: nothing depends on this row (num = 0)
        for (int c4 = lane; c4 < a.g.Wd4; c4 += 32)
            *reinterpret_cast<float4*>(a.dP1 + (long)row * a.ldd1 + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;                                                     // G / rowsc of the row are never gathered
    }
    RowCtx<NCH, HT> rc;
    bwd_row_prologue<NCH, HT>(a, row, lane, rc);
    float usum[HT], swsum[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) {
        usum[h] = 0.f;
        swsum[h] = h < a.g.H ? __ldg(sw + (long)row * a.g.H + h) : 0.f;
    }
    bwd_row_store<NCH, HT>(a, row, lane, rc, usum, swsum, true, true);
    if (lane < a.g.H) {
        const float4 o = make_float4(selh<HT>(lane, rc.q1), selh<HT>(lane, rc.c1), selh<HT>(lane, rc.dden), 0.f);
        *reinterpret_cast<float4*>(rowsc + ((long)row * a.g.H + lane) * 4) = o;
    }
}

// ---- column pass -----------------------------------------------------------------------------------
template <int NCH, int HT>
struct ColCtx {
    float4 p2[NCH];
    int hc[NCH];
    float q2[HT];
};

template <int NCH, int HT>
__device__ __forceinline__ void col_ctx_load(const BwdFusedArgs& a, int colj, int lane, ColCtx<NCH, HT>& cc) {
    const LayerGeom g = a.g;
    const float* p = a.P2 + (long)colj * a.ld2;
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        cc.hc[ci] = (HT > 1 && c4 < g.Dt4) ? c4 / g.Dp4 : 0;
        cc.p2[ci] = c4 < g.Dt4 ? ldg4(p + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 qv = ldg4(p + (long)g.Dt4 * 4);
#pragma unroll
    for (int h = 0; h < HT; ++h) cc.q2[h] = f4get(qv, h);
}

template <int NCH, int HT>
struct ColAcc {
    float4 acc[NCH];
    float vs[HT];        // lane-local partial sums of ds
};

template <int NCH, int HT, bool HAS2, int U>
__device__ __forceinline__ void col_accumulate(const BwdFusedArgs& a, int beg, int end, int lane,
                                               const ColCtx<NCH, HT>& cc, ColAcc<NCH, HT>& st) {
    const LayerGeom g = a.g;
    const int H = g.H;
    const bool has_mask = a.mask != nullptr;
    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int my_row = 0, my_pos = 0, my_k1 = 0, my_k2 = -1;
        float my_w[HT], my_ca[HT], my_cb[HT], my_t[HT];
#pragma unroll
        for (int h = 0; h < HT; ++h) { my_w[h] = 0.f; my_ca[h] = 0.f; my_cb[h] = 0.f; my_t[h] = 0.f; }
        if (lane < n) {                                            // every per-edge scalar, one edge per lane
            my_row = __ldg(a.csc_row + base + lane);
            my_pos = __ldg(a.csc_pos + base + lane);
            my_k1 = __ldg(a.csc_t1 + base + lane);
            if (HAS2) my_k2 = __ldg(a.csc_t2 + base + lane);
            float4 q3 = ldg4(a.P3 + (long)my_k1 * a.ld3 + (long)g.Dt4 * 4);
            if (HAS2 && my_k2 >= 0) q3 = f4add(q3, ldg4(a.P3 + (long)my_k2 * a.ld3 + (long)g.Dt4 * 4));
#pragma unroll
            for (int h = 0; h < HT; ++h) {
                if (h < H) {
                    const float4 rs = ldg4(a.rowsc + ((long)my_row * H + h) * 4);     // q1, c1, dden
                    const float m = has_mask ? __ldg(a.mask + (long)h * a.mask_stride + my_pos) : 1.f;
                    const float s = rs.x + cc.q2[h] + f4get(q3, h);
                    const float slope = s > 0.f ? 1.f : a.alpha;
                    const float ee = fexp(-(s * slope));                                // layers.py:143-146
                    my_w[h] = ee * m;                                                   // layers.py:158
                    my_ca[h] = my_w[h] * slope;                                         // ds = -(ca * dot + cb)
                    my_cb[h] = fmaf(m, rs.y, rs.z) * ee * slope;
                }
            }
        }
        for (int u0 = 0; u0 < n; u0 += U) {
            float4 v[U][NCH], r[U][NCH];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = u0 + u;
                const int i = __shfl_sync(FULLM, my_row, idx & 31);
                const int k1 = __shfl_sync(FULLM, my_k1, idx & 31);
                int k2 = -1;
                if (HAS2) k2 = __shfl_sync(FULLM, my_k2, idx & 31);
                const float* gp = a.G + (long)i * a.ldg;
                const float* p3 = a.P3 + (long)k1 * a.ld3;
                const float* p3b = a.P3 + (long)(k2 < 0 ? 0 : k2) * a.ld3;
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    const int c4 = lane + 32 * ci;
                    const bool ok = idx < n && c4 < g.Dt4;
                    v[u][ci] = ok ? ldg4(gp + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 x = ok ? ldg4(p3 + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (HAS2 && ok && k2 >= 0) x = f4add(x, ldg4(p3b + c4 * 4));
                    r[u][ci] = x;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = u0 + u;
                if (idx >= n) break;
                float pd[HT];
#pragma unroll
                for (int h = 0; h < HT; ++h) pd[h] = 0.f;
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) {
                    const float d = f4dot(v[u][ci], f4add(cc.p2[ci], r[u][ci]));
#pragma unroll
                    for (int h = 0; h < HT; ++h) pd[h] += (HT == 1 || cc.hc[ci] == h) ? d : 0.f;
                }
                float w[HT];
#pragma unroll
                for (int h = 0; h < HT; ++h) {
                    const float t = warp_sum(pd[h]);
                    if (lane == idx) my_t[h] = t;
                    w[h] = __shfl_sync(FULLM, my_w[h], idx & 31);
                }
#pragma unroll
                for (int ci = 0; ci < NCH; ++ci) f4fma(st.acc[ci], selh<HT>(cc.hc[ci], w), v[u][ci]);
            }
        }
        if (lane < n) {
            float* rp = a.rec + (long)my_pos * (2 * H);
#pragma unroll
            for (int h = 0; h < HT; ++h) {
                if (h < H) {
                    const float ds = -fmaf(my_ca[h], my_t[h], my_cb[h]);
                    st.vs[h] += ds;
                    *reinterpret_cast<float2*>(rp + 2 * h) = make_float2(my_w[h], ds);
                }
            }
        }
    }
}

template <int NCH, int HT>
__device__ __forceinline__ void col_store(float* dst, const LayerGeom& g, int lane, const ColAcc<NCH, HT>& st) {
    float vs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int h = 0; h < HT; ++h) vs[h] = warp_sum(st.vs[h]);
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        if (c4 >= g.Wd4) continue;
        float4 o = st.acc[ci];
        if (c4 == g.Dt4) o = make_float4(vs[0], vs[1], vs[2], vs[3]);
        else if (c4 > g.Dt4) o = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(dst + c4 * 4) = o;
    }
}

template <int NCH, int HT>
__device__ __forceinline__ void col_acc_init(ColAcc<NCH, HT>& st) {
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) st.acc[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int h = 0; h < HT; ++h) st.vs[h] = 0.f;
}

template <int NCH, int HT, bool HAS2, int U, int MINB>
__global__ void __launch_bounds__(SPK_CTA_THREADS, MINB)
bwd_cols_fused_kernel(const BwdFusedArgs a) {
    const int lane = threadIdx.x & 31;
    const int colj = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (colj >= a.n_cols) return;
    const int beg = __ldg(a.colptr + colj), end = __ldg(a.colptr + colj + 1);
    if (end - beg > a.col_hub.hub_thresh) return;
    ColAcc<NCH, HT> st;
    col_acc_init<NCH, HT>(st);
    if (end > beg) {
        ColCtx<NCH, HT> cc;
        col_ctx_load<NCH, HT>(a, colj, lane, cc);
        col_accumulate<NCH, HT, HAS2, U>(a, beg, end, lane, cc, st);
    }
    col_store<NCH, HT>(a.dP2 + (long)colj * a.ldd2, a.g, lane, st);
}

template <int NCH, int HT, bool HAS2, int U, int MINB>
__global__ void __launch_bounds__(SPK_CTA_THREADS, MINB)
bwd_cols_fused_tasks_kernel(const BwdFusedArgs a) {
    const int lane = threadIdx.x & 31;
    const int task = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (task >= a.col_hub.n_tasks) return;
    const int colj = __ldg(a.col_hub.task_seg + task);
    ColAcc<NCH, HT> st;
    col_acc_init<NCH, HT>(st);
    ColCtx<NCH, HT> cc;
    col_ctx_load<NCH, HT>(a, colj, lane, cc);
    col_accumulate<NCH, HT, HAS2, U>(a, __ldg(a.col_hub.task_beg + task), __ldg(a.col_hub.task_end + task), lane, cc, st);
    col_store<NCH, HT>(a.col_hub.partial + (long)task * a.col_hub.ldpart, a.g, lane, st);
}

// ---- row sums of ds --------------------------------------------------------------------------------
// rec is [E, 2H] = (w, ds) per head in CSR order, so a row's records are contiguous. Warp per row, lane-strided.
template <int HT>
__global__ void __launch_bounds__(SPK_CTA_THREADS)
bwd_rowsum_kernel(const int* __restrict__ rowptr, const float* __restrict__ rec, int H, int n_rows, int hub_thresh,
                  float* __restrict__ dP1, long ldd1, int qoff) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    if (end - beg > hub_thresh) return;
    float u[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) u[h] = 0.f;
    for (int e = beg + lane; e < end; e += 32) {
#pragma unroll
        for (int h = 0; h < HT; ++h)
            if (h < H) u[h] += __ldg(rec + (long)e * (2 * H) + 2 * h + 1);
    }
#pragma unroll
    for (int h = 0; h < HT; ++h) u[h] = warp_sum(u[h]);
    if (lane < H) dP1[(long)row * ldd1 + qoff + lane] = selh<HT>(lane, u);
}

// one CTA per hub row: thread-strided sums in ascending order, then a fixed-order tree over the CTA
template <int HT>
__global__ void __launch_bounds__(512)
bwd_rowsum_hub_kernel(const int* __restrict__ rowptr, const int* __restrict__ hub_seg, const float* __restrict__ rec, int H,
                      float* __restrict__ dP1, long ldd1, int qoff) {
    __shared__ float red[HT][512];
    const int row = __ldg(hub_seg + blockIdx.x);
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    float u[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) u[h] = 0.f;
    for (int e = beg + threadIdx.x; e < end; e += 512) {
#pragma unroll
        for (int h = 0; h < HT; ++h)
            if (h < H) u[h] += __ldg(rec + (long)e * (2 * H) + 2 * h + 1);
    }
#pragma unroll
    for (int h = 0; h < HT; ++h) red[h][threadIdx.x] = u[h];
    __syncthreads();
    for (int s = 256; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
#pragma unroll
            for (int h = 0; h < HT; ++h) red[h][threadIdx.x] += red[h][threadIdx.x + s];
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < H) dP1[(long)row * ldd1 + qoff + threadIdx.x] = red[threadIdx.x < HT ? threadIdx.x : 0][0];
}

// tuning variant of the column pass: SPK_FUSED_VARIANT=0 (U=4 gathers in flight per warp, 2 CTAs/SM) or 1 (U=2, 3 CTAs/SM)
static int fused_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SPK_FUSED_VARIANT"); v = e ? atoi(e) : 0; if (v < 0 || v > 1) v = 0; }
    return v;
}

template <int NCH, int HT, bool HAS2, int U, int MINB>
int launch_cols_v(const BwdFusedArgs& a, cudaStream_t s) {
    if (a.n_cols > 0) {
        const unsigned grid = (a.n_cols + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
        bwd_cols_fused_kernel<NCH, HT, HAS2, U, MINB><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("bwd_cols_fused")) return rc;
    }
    if (a.col_hub.n_tasks > 0) {
        const unsigned grid = (a.col_hub.n_tasks + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
        bwd_cols_fused_tasks_kernel<NCH, HT, HAS2, U, MINB><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("bwd_cols_fused_tasks")) return rc;
    }
    return 0;
}

template <int NCH, int HT, bool HAS2>
int launch_fused_t(const BwdFusedArgs& a, cudaStream_t s) {
    constexpr int U = (NCH <= 2) ? 4 : 2;
    constexpr int MINB = (NCH <= 2) ? 2 : 1;
    bool done = false;
    if constexpr (NCH <= 2) {
        if (fused_variant() == 1) { if (int rc = launch_cols_v<NCH, HT, HAS2, 2, 3>(a, s)) return rc; done = true; }
    }
    if (!done) { if (int rc = launch_cols_v<NCH, HT, HAS2, U, MINB>(a, s)) return rc; }
    if (a.col_hub.n_tasks > 0) {
        SegGatherArgs fa;
        fa.segptr = a.colptr; fa.src = nullptr; fa.pos = nullptr; fa.G = a.G; fa.ldg = a.ldg; fa.rec = a.rec;
        fa.outp = a.dP2; fa.ldout = a.ldd2; fa.n_seg = a.n_cols; fa.prefer_stream = 0; fa.g = a.g; fa.hub = a.col_hub;
        if (int rc = launch_seg_gather_hub_finalize(fa, s)) return rc;
    }
    if (a.n_rows > 0) {
        const unsigned grid = (a.n_rows + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
        bwd_rowsum_kernel<HT><<<grid, SPK_CTA_THREADS, 0, s>>>(a.rowptr, a.rec, a.g.H, a.n_rows, a.row_hub.hub_thresh,
                                                               a.dP1, a.ldd1, a.g.Dt4 * 4);
        if (int rc = check_launch("bwd_rowsum")) return rc;
        if (a.row_hub.n_hubs > 0) {
            bwd_rowsum_hub_kernel<HT><<<a.row_hub.n_hubs, 512, 0, s>>>(a.rowptr, a.row_hub.hub_seg, a.rec, a.g.H, a.dP1,
                                                                       a.ldd1, a.g.Dt4 * 4);
            if (int rc = check_launch("bwd_rowsum_hub")) return rc;
        }
    }
    return 0;
}

template <int NCH>
int launch_fused_n(const BwdFusedArgs& a, cudaStream_t s) {
    const bool has2 = a.csc_t2 != nullptr;
    if (a.g.H == 1) return has2 ? launch_fused_t<NCH, 1, true>(a, s) : launch_fused_t<NCH, 1, false>(a, s);
    if (a.g.H == 2) return has2 ? launch_fused_t<NCH, 2, true>(a, s) : launch_fused_t<NCH, 2, false>(a, s);
    return has2 ? launch_fused_t<NCH, 4, true>(a, s) : launch_fused_t<NCH, 4, false>(a, s);
}


template <int NCH>
int launch_node_n(const EdgeBwdRowsArgs& ra, const BwdFusedArgs& a, cudaStream_t s) {
    const unsigned grid = (a.n_rows + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
    if (a.g.H == 1) bwd_node_kernel<NCH, 1><<<grid, SPK_CTA_THREADS, 0, s>>>(ra, a.sw, a.rowsc);
    else if (a.g.H == 2) bwd_node_kernel<NCH, 2><<<grid, SPK_CTA_THREADS, 0, s>>>(ra, a.sw, a.rowsc);
    else bwd_node_kernel<NCH, 4><<<grid, SPK_CTA_THREADS, 0, s>>>(ra, a.sw, a.rowsc);
    return check_launch("bwd_node");
}

}  // namespace

int launch_edge_bwd_node(const BwdFusedArgs& a, cudaStream_t s) {
    if (a.n_rows <= 0) return 0;
    EdgeBwdRowsArgs ra;                                            // the node pass reuses the K3 row prologue / store
    ra.segptr = a.rowptr; ra.col = nullptr; ra.t1 = nullptr; ra.t2 = nullptr;
    ra.P1 = a.P1; ra.ld1 = a.ld1; ra.P2 = a.P2; ra.ld2 = a.ld2; ra.P3 = a.P3; ra.ld3 = a.ld3;
    ra.mask = nullptr; ra.mask_stride = 0;
    ra.out = a.out; ra.dout = a.dout; ra.ldo = a.ldo; ra.den = a.den;
    ra.G = a.G; ra.ldg = a.ldg; ra.dP1 = a.dP1; ra.ldd1 = a.ldd1; ra.rec = a.rec;
    ra.n_rows = a.n_rows; ra.g = a.g; ra.alpha = a.alpha; ra.apply_elu = a.apply_elu; ra.out_vec = a.out_vec;
    ra.hub = a.row_hub;
    switch ((a.g.Wd4 + 31) / 32) {
        case 1: return launch_node_n<1>(ra, a, s);
        case 2: return launch_node_n<2>(ra, a, s);
        case 3: return launch_node_n<3>(ra, a, s);
        case 4: return launch_node_n<4>(ra, a, s);
        default: set_error("edge_bwd_node: row width %d floats exceeds the supported 512", a.g.Wd4 * 4); return 2;
    }
}

int launch_edge_bwd_fused(const BwdFusedArgs& a, cudaStream_t s) {
    if (int rc = launch_edge_bwd_node(a, s)) return rc;
    switch ((a.g.Wd4 + 31) / 32) {
        case 1: return launch_fused_n<1>(a, s);
        case 2: return launch_fused_n<2>(a, s);
        case 3: return launch_fused_n<3>(a, s);
        case 4: return launch_fused_n<4>(a, s);
        default: set_error("edge_bwd_fused: row width %d floats exceeds the supported 512", a.g.Wd4 * 4); return 2;
    }
}

}  // namespace spk
