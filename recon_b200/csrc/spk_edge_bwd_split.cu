// K3"/K4": backward of one projected attention-layer group without the row-major edge pass and without any per-edge
// gather of projected rows: the dot t_e = dnum_i . m_e = c1_i + dnum_i . P2[j] + dnum_i . P3[k] (closed form of the
// autograd of GAT/layers.py:124-175) is split over the two passes that gather dnum_i anyway.
//   node pass     (spk_edge_bwd_fused.cu) : dnum -> G, (q1, c1, dden) -> rowsc, dP1~ = [sw*dnum | . | 0]
//   column pass   (warp per column j, P2~[j] in registers): per edge gather G[i]:  t2 = G[i] . P2[j],
//                 ee = exp(-LeakyReLU(q1_i + q2_j + q3_k)),  w = msk*ee,  A = w*slope,
//                 B = A*(c1_i + t2) + dden_i*ee*slope   ->  rec4[e] = (w, A, B, 0);   dP2~[j] += w * G[i]
//   relation pass (warp per relation chunk, P3~[k] in registers): per edge gather G[i]:  t3 = G[i] . P3[k],
//                 ds = -(A*t3 + B)  ->  dsv[e];   dP3~[k] = [ sum w*G[i] | sum ds | 0 ]
//   sums          : dP1~[i] q slot = sum of ds over the row (contiguous records), dP2~[j] q slot = sum over the column
// Every edge costs two Dt-wide gathers of G (one per pass) instead of three (K3: P2[j]; K4 cols, K4 rels: G[i]) plus the
// L2 traffic of the P3 rows. Only for graphs without 2-hop edges (a 2-hop edge has two relation rows). Deterministic.
#include <stdlib.h>
#include "spk_edge.cuh"
#include "spk_edge_bwd.cuh"
#include "spk_edge_bwd_fused.cuh"

namespace spk {
namespace {

constexpr unsigned FULLM = 0xffffffffu;
__device__ __forceinline__ float sexp(float x) { return exp2f(x * 1.4426950408889634f); }

template <int NCH, int HT>
struct SegCtx {
    float4 p[NCH];       // the segment's own projected row (P2~[j] or P3~[k]) in registers
    int hc[NCH];
    float q[HT];
};

template <int NCH, int HT>
__device__ __forceinline__ void seg_ctx_load(const float* __restrict__ row, const LayerGeom& g, int lane, SegCtx<NCH, HT>& cc,
                                             int dup = 0) {
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        cc.hc[ci] = (HT > 1 && c4 < g.Dt4) ? c4 / g.Dp4 : 0;
        // dup: the table row holds ONE copy of the vector (Dp floats) that every head uses
        const int src4 = dup ? c4 % g.Dp4 : c4;
        cc.p[ci] = c4 < g.Dt4 ? ldg4(row + src4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 qv = ldg4(row + (long)(dup ? g.Dp4 : g.Dt4) * 4);
#pragma unroll
    for (int h = 0; h < HT; ++h) cc.q[h] = f4get(qv, h);
}

template <int NCH, int HT>
struct SplitAcc {
    float4 acc[NCH];
    float vs[HT];
};

template <int NCH, int HT>
__device__ __forceinline__ void split_acc_init(SplitAcc<NCH, HT>& st) {
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) st.acc[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int h = 0; h < HT; ++h) st.vs[h] = 0.f;
}

template <int NCH, int HT>
__device__ __forceinline__ void split_store(float* dst, const LayerGeom& g, int lane, const SplitAcc<NCH, HT>& st) {
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

// (row gathers stay on ld.global.nc with L1 allocation: with L2-only loads (ld.global.cg / L1::no_allocate) the column pass
//  gained 2 % but the windowed relation pass went 2.25 -> 3.9-4.2 ms -- the G rows of its window are also reused out of L1)
// gathers of the U edges starting at batch position u0 (rows broadcast from the lanes that hold them)
template <int NCH, int U>
__device__ __forceinline__ void gather_rows(const float* __restrict__ G, long ldg, int Dt4, int my_row, int u0, int n, int lane,
                                            float4 (&v)[U][NCH]) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int idx = u0 + u;
        const int i = __shfl_sync(FULLM, my_row, idx & 31);
        const float* gp = G + (long)i * ldg;
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
            const int c4 = lane + 32 * ci;
            v[u][ci] = (idx < n && c4 < Dt4) ? ldg4(gp + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// dot of every gathered row with the segment context per head -> my_t of the lane that owns the edge; acc += w * row
// (reducing the U * HT partial dots together with transposed_warp_sum was measured: no gain on the 26-float4 rows of layer 1,
// 8 % slower on the 50-float4 rows of layer 2 -- these passes wait on the gathers, not on the shuffles)
template <int NCH, int HT, int U>
__device__ __forceinline__ void consume_rows(const SegCtx<NCH, HT>& cc, const float4 (&v)[U][NCH], int u0, int n, int lane,
                                             const float (&my_w)[HT], float (&my_t)[HT], SplitAcc<NCH, HT>& st) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int idx = u0 + u;
        if (idx >= n) break;
        float pd[HT];
#pragma unroll
        for (int h = 0; h < HT; ++h) pd[h] = 0.f;
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
            const float d = f4dot(v[u][ci], cc.p[ci]);
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

// ---- column pass -----------------------------------------------------------------------------------
template <int NCH, int HT, int U>
__device__ __forceinline__ void split_cols_accumulate(const BwdSplitArgs& a, int beg, int end, int lane,
                                                      const SegCtx<NCH, HT>& cc, SplitAcc<NCH, HT>& st) {
    const LayerGeom g = a.g;
    const int H = g.H;
    const bool has_mask = a.mask != nullptr;
    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int my_row = 0, my_pos = 0, my_k1 = 0;
        if (lane < n) {
            my_row = __ldg(a.csc_row + base + lane);
            my_pos = __ldg(a.csc_pos + base + lane);
            my_k1 = __ldg(a.csc_t1 + base + lane);
        }
        float4 v[U][NCH];
        gather_rows<NCH, U>(a.G, a.ldg, g.Dt4, my_row, 0, n, lane, v);      // in flight while the scalars are formed
        float my_w[HT], my_A[HT], my_B[HT], my_t[HT];
#pragma unroll
        for (int h = 0; h < HT; ++h) { my_w[h] = 0.f; my_A[h] = 0.f; my_B[h] = 0.f; my_t[h] = 0.f; }
        if (lane < n) {
            const float4 q3 = ldg4(a.P3 + (long)my_k1 * a.ld3 + (long)(a.dup ? g.Dp4 : g.Dt4) * 4);
#pragma unroll
            for (int h = 0; h < HT; ++h) {
                if (h < H) {
                    const float4 rs = ldg4(a.rowsc + ((long)my_row * H + h) * 4);     // q1, c1, dden
                    const float m = has_mask ? __ldg(a.mask + (long)h * a.mask_stride + my_pos) : 1.f;
                    const float s = rs.x + cc.q[h] + f4get(q3, h);
                    const float slope = s > 0.f ? 1.f : a.alpha;
                    const float ee = sexp(-(s * slope));                                // layers.py:143-146
                    my_w[h] = ee * m;                                                   // layers.py:158
                    my_A[h] = my_w[h] * slope;
                    my_B[h] = fmaf(my_A[h], rs.y, rs.z * ee * slope);                   // + A * t2 below
                }
            }
        }
        for (int u0 = 0;;) {
            consume_rows<NCH, HT, U>(cc, v, u0, n, lane, my_w, my_t, st);
            u0 += U;
            if (u0 >= n) break;
            gather_rows<NCH, U>(a.G, a.ldg, g.Dt4, my_row, u0, n, lane, v);
        }
        if (lane < n) {
#pragma unroll
            for (int h = 0; h < HT; ++h)
                if (h < H)
                    *reinterpret_cast<float4*>(a.rec4 + ((long)my_pos * H + h) * 4) =
                        make_float4(my_w[h], my_A[h], fmaf(my_A[h], my_t[h], my_B[h]), 0.f);
        }
    }
}

template <int NCH, int HT, int U, int MINB>
__global__ void __launch_bounds__(SPK_CTA_THREADS, MINB)
split_cols_kernel(const BwdSplitArgs a) {
    // persistent warps (grid = resident CTAs): warp w walks columns w, w + W, w + 2W, ... with the pointers of its next column
    // requested one column ahead, so neighbouring warps still work on neighbouring columns and a warp's dependent chain
    // (pointers -> indices -> row gathers) overlaps across its columns; no CTA turnover for 10-edge columns
    const int lane = threadIdx.x & 31;
    const int nw = gridDim.x * SPK_WARPS_PER_CTA;
    int colj = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (colj >= a.n_cols) return;
    // (also prefetching the next column's first index batch cost more in spills than it hid: cols 4.2 -> 4.8 ms)
    int beg_n = __ldg(a.colptr + colj), end_n = __ldg(a.colptr + colj + 1);
    for (; colj < a.n_cols; colj += nw) {
        const int beg = beg_n, end = end_n;
        if (colj + nw < a.n_cols) { beg_n = __ldg(a.colptr + colj + nw); end_n = __ldg(a.colptr + colj + nw + 1); }
        if (end - beg > a.col_hub.hub_thresh) continue;
        SplitAcc<NCH, HT> st;
        split_acc_init<NCH, HT>(st);
        if (end > beg) {
            SegCtx<NCH, HT> cc;
            seg_ctx_load<NCH, HT>(a.P2 + (long)colj * a.ld2, a.g, lane, cc, a.dup);
            split_cols_accumulate<NCH, HT, U>(a, beg, end, lane, cc, st);
        }
        split_store<NCH, HT>(a.dP2 + (long)colj * a.ldd2, a.g, lane, st);
    }
}

template <int NCH, int HT, int U, int MINB>
__global__ void __launch_bounds__(SPK_CTA_THREADS, MINB)
split_cols_tasks_kernel(const BwdSplitArgs a) {
    const int lane = threadIdx.x & 31;
    const int task = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (task >= a.col_hub.n_tasks) return;
    const int colj = __ldg(a.col_hub.task_seg + task);
    SplitAcc<NCH, HT> st;
    split_acc_init<NCH, HT>(st);
    SegCtx<NCH, HT> cc;
    seg_ctx_load<NCH, HT>(a.P2 + (long)colj * a.ld2, a.g, lane, cc, a.dup);
    split_cols_accumulate<NCH, HT, U>(a, __ldg(a.col_hub.task_beg + task), __ldg(a.col_hub.task_end + task), lane, cc, st);
    split_store<NCH, HT>(a.col_hub.partial + (long)task * a.col_hub.ldpart, a.g, lane, st);
}

// ---- relation pass ---------------------------------------------------------------------------------
template <int NCH, int HT, int U>
__device__ __forceinline__ void split_rels_accumulate(const BwdSplitArgs& a, int beg, int end, int lane,
                                                      const SegCtx<NCH, HT>& cc, SplitAcc<NCH, HT>& st) {
    const LayerGeom g = a.g;
    const int H = g.H;
    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int my_row = 0, my_pos = 0;
        if (lane < n) {
            my_row = __ldg(a.rel_row + base + lane);
            my_pos = __ldg(a.rel_pos + base + lane);
        }
        float4 v[U][NCH];
        gather_rows<NCH, U>(a.G_rel, a.ldg_rel, g.Dt4, my_row, 0, n, lane, v);
        float my_w[HT], my_A[HT], my_B[HT], my_t[HT];
#pragma unroll
        for (int h = 0; h < HT; ++h) { my_w[h] = 0.f; my_A[h] = 0.f; my_B[h] = 0.f; my_t[h] = 0.f; }
        if (lane < n) {
#pragma unroll
            for (int h = 0; h < HT; ++h) {
                if (h < H) {
                    const float4 r4 = ldg4(a.rec4 + ((long)my_pos * H + h) * 4);
                    my_w[h] = r4.x; my_A[h] = r4.y; my_B[h] = r4.z;
                }
            }
        }
        for (int u0 = 0;;) {
            consume_rows<NCH, HT, U>(cc, v, u0, n, lane, my_w, my_t, st);
            u0 += U;
            if (u0 >= n) break;
            gather_rows<NCH, U>(a.G_rel, a.ldg_rel, g.Dt4, my_row, u0, n, lane, v);
        }
        if (lane < n) {
#pragma unroll
            for (int h = 0; h < HT; ++h) {
                if (h < H) {
                    const float ds = -fmaf(my_A[h], my_t[h], my_B[h]);
                    st.vs[h] += ds;
                    a.dsv[(long)my_pos * H + h] = ds;
                }
            }
        }
    }
}

template <int NCH, int HT, int U, int MINB>
__global__ void __launch_bounds__(SPK_CTA_THREADS, MINB)
split_rels_kernel(const BwdSplitArgs a) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (k >= a.n_rel) return;
    const int beg = __ldg(a.relptr + k), end = __ldg(a.relptr + k + 1);
    if (end - beg > a.rel_hub.hub_thresh) return;
    SplitAcc<NCH, HT> st;
    split_acc_init<NCH, HT>(st);
    if (end > beg) {
        SegCtx<NCH, HT> cc;
        seg_ctx_load<NCH, HT>(a.P3 + (long)k * a.ld3, a.g, lane, cc, a.dup);
        split_rels_accumulate<NCH, HT, U>(a, beg, end, lane, cc, st);
    }
    split_store<NCH, HT>(a.dP3 + (long)k * a.ldd3, a.g, lane, st);
}

template <int NCH, int HT, int U, int MINB>
__global__ void __launch_bounds__(SPK_CTA_THREADS, MINB)
split_rels_tasks_kernel(const BwdSplitArgs a) {
    // (one warp per task: persistent warps walking slots w, w + W, ... were measured 2-20 % slower here -- they stretch the row
    //  window that the launch order keeps L2-resident)
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (slot >= a.rel_hub.n_tasks) return;
    const int task = hub_task_of_slot(a.rel_hub, slot);      // row-window order: concurrently running tasks share G rows in L2
    const int k = __ldg(a.rel_hub.task_seg + task);
    SplitAcc<NCH, HT> st;
    split_acc_init<NCH, HT>(st);
    SegCtx<NCH, HT> cc;
    seg_ctx_load<NCH, HT>(a.P3 + (long)k * a.ld3, a.g, lane, cc, a.dup);
    split_rels_accumulate<NCH, HT, U>(a, __ldg(a.rel_hub.task_beg + task), __ldg(a.rel_hub.task_end + task), lane, cc, st);
    split_store<NCH, HT>(a.rel_hub.partial + (long)task * a.rel_hub.ldpart, a.g, lane, st);
}

// ---- sums of ds ------------------------------------------------------------------------------------
// dst[seg, qoff + h] = sum over the segment's entries e of dsv[idx(e), h]; idx = identity (rows: records are contiguous)
// or csc_pos (columns). Warp per segment, lane-strided in ascending order + butterfly: fixed summation order.
// SUM_G lanes per segment (segments average ~10 entries: a whole warp per segment spent its time on two dependent round
// trips and a butterfly for 10 numbers; 2M warps of that cost 0.4-0.5 ms per launch, four launches per step)
constexpr int SUM_G = 4;
template <int HT>
__global__ void __launch_bounds__(SPK_CTA_THREADS)
split_sum_kernel(const int* __restrict__ segptr, const int* __restrict__ idx, const float* __restrict__ dsv, int H, int n_seg,
                 int hub_thresh, float* __restrict__ dst, long ldd, int qoff) {
    const int sub = threadIdx.x % SUM_G;
    const long seg_l = ((long)blockIdx.x * SPK_CTA_THREADS + threadIdx.x) / SUM_G;
    const bool live = seg_l < n_seg;
    const int seg = live ? (int)seg_l : n_seg - 1;
    const int beg = __ldg(segptr + seg);
    int end = __ldg(segptr + seg + 1);
    const bool skip = !live || end - beg > hub_thresh;
    if (skip) end = beg;
    float u[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) u[h] = 0.f;
#pragma unroll 2
    for (int e = beg + sub; e < end; e += SUM_G) {
        const long p = idx ? (long)__ldg(idx + e) : (long)e;
        if (HT == 2 && H == 2) {
            const float2 d = __ldg(reinterpret_cast<const float2*>(dsv) + p);
            u[0] += d.x; u[HT > 1 ? 1 : 0] += d.y;
        } else {
#pragma unroll
            for (int h = 0; h < HT; ++h)
                if (h < H) u[h] += __ldg(dsv + p * H + h);
        }
    }
#pragma unroll
    for (int h = 0; h < HT; ++h) {
#pragma unroll
        for (int o = SUM_G / 2; o > 0; o >>= 1) u[h] += __shfl_xor_sync(0xffffffffu, u[h], o);
    }
    if (!skip && sub < H) dst[(long)seg * ldd + qoff + sub] = selh<HT>(sub, u);
}

// hub segments: one warp per 256-entry task (the hub task table of the segment ordering) writes a partial, then one CTA per
// hub adds the partials: thread-strided in ascending order, then a fixed tree over the CTA
template <int HT>
__global__ void __launch_bounds__(SPK_CTA_THREADS)
split_sum_tasks_kernel(const HubTasks hub, const int* __restrict__ idx, const float* __restrict__ dsv, int H) {
    const int lane = threadIdx.x & 31;
    const int task = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (task >= hub.n_tasks) return;
    const int beg = __ldg(hub.task_beg + task), end = __ldg(hub.task_end + task);
    float u[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) u[h] = 0.f;
    for (int e = beg + lane; e < end; e += 32) {
        const long p = idx ? (long)__ldg(idx + e) : (long)e;
#pragma unroll
        for (int h = 0; h < HT; ++h)
            if (h < H) u[h] += __ldg(dsv + p * H + h);
    }
#pragma unroll
    for (int h = 0; h < HT; ++h) u[h] = warp_sum(u[h]);
    if (lane < HT) hub.partial[(long)task * hub.ldpart + lane] = selh<HT>(lane, u);
}

template <int HT>
__global__ void __launch_bounds__(256)
split_sum_hub_kernel(const HubTasks hub, int H, float* __restrict__ dst, long ldd, int qoff) {
    __shared__ float red[HT][256];
    const int seg = __ldg(hub.hub_seg + blockIdx.x);
    const int t0 = __ldg(hub.hub_task_ptr + blockIdx.x), t1 = __ldg(hub.hub_task_ptr + blockIdx.x + 1);
    float u[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) u[h] = 0.f;
    for (int t = t0 + threadIdx.x; t < t1; t += 256) {
#pragma unroll
        for (int h = 0; h < HT; ++h) u[h] += hub.partial[(long)t * hub.ldpart + h];
    }
#pragma unroll
    for (int h = 0; h < HT; ++h) red[h][threadIdx.x] = u[h];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
#pragma unroll
            for (int h = 0; h < HT; ++h) red[h][threadIdx.x] += red[h][threadIdx.x + s];
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < H) dst[(long)seg * ldd + qoff + threadIdx.x] = red[threadIdx.x < HT ? threadIdx.x : 0][0];
}

template <int HT>
int launch_sums(const int* segptr, const int* idx, const float* dsv, int H, int n_seg, const HubTasks& hub, float* dst,
                long ldd, int qoff, cudaStream_t s) {
    if (n_seg <= 0) return 0;
    static_assert(HT <= SUM_G, "a segment's lanes store one head each");
    const unsigned grid = (unsigned)(((long)n_seg * SUM_G + SPK_CTA_THREADS - 1) / SPK_CTA_THREADS);
    split_sum_kernel<HT><<<grid, SPK_CTA_THREADS, 0, s>>>(segptr, idx, dsv, H, n_seg, hub.hub_thresh, dst, ldd, qoff);
    if (int rc = check_launch("split_sum")) return rc;
    if (hub.n_tasks > 0) {
        split_sum_tasks_kernel<HT><<<(hub.n_tasks + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA, SPK_CTA_THREADS, 0, s>>>(hub, idx, dsv, H);
        if (int rc = check_launch("split_sum_tasks")) return rc;
        split_sum_hub_kernel<HT><<<hub.n_hubs, 256, 0, s>>>(hub, H, dst, ldd, qoff);
        if (int rc = check_launch("split_sum_hub")) return rc;
    }
    return 0;
}

// tuning variants of the two gather passes (SPK_SPLIT_VARIANT): unroll depth U (row gathers in flight per warp) and CTAs/SM
static int split_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SPK_SPLIT_VARIANT"); v = e ? atoi(e) : 0; if (v < 0 || v > 3) v = 0; }
    return v;
}

static int sm_count() {
    static int n[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (n[dev] == 0) { int v = 148; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); n[dev] = v > 0 ? v : 148; }
    return n[dev];
}
// SPK_SPLIT_PERSIST = waves of resident CTAs in the column-pass grid (0: one warp per column, the round-1 launch)
static int split_persist() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SPK_SPLIT_PERSIST"); v = e ? atoi(e) : 8; if (v < 0 || v > 64) v = 8; }
    return v;
}

template <int NCH, int HT, int U, int MINB>
int launch_split_passes(const BwdSplitArgs& a, cudaStream_t s, int which) {
    if (which == 0) {
        if (a.n_cols > 0) {
            const unsigned full = (a.n_cols + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
            const unsigned resident = (unsigned)(sm_count() * MINB) * (unsigned)split_persist();
            const unsigned grid = (split_persist() > 0 && full > resident) ? resident : full;
            split_cols_kernel<NCH, HT, U, MINB><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
            if (int rc = check_launch("split_cols")) return rc;
        }
        if (a.col_hub.n_tasks > 0) {
            const unsigned grid = (a.col_hub.n_tasks + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
            split_cols_tasks_kernel<NCH, HT, U, MINB><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
            if (int rc = check_launch("split_cols_tasks")) return rc;
        }
    } else {
        if (a.n_rel > 0) {
            const unsigned grid = (a.n_rel + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
            split_rels_kernel<NCH, HT, U, MINB><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
            if (int rc = check_launch("split_rels")) return rc;
        }
        if (a.rel_hub.n_tasks > 0) {
            const unsigned grid = (a.rel_hub.n_tasks + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
            split_rels_tasks_kernel<NCH, HT, U, MINB><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
            if (int rc = check_launch("split_rels_tasks")) return rc;
        }
    }
    return 0;
}

// narrow rows (one float4 per lane, the aggregate-then-project layer): a whole short segment in flight at once
static int split_variant_narrow() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SPK_SPLIT_VARIANT1"); v = e ? atoi(e) : 0; if (v < 0 || v > 4) v = 0; }
    return v;
}

template <int NCH, int HT>
int launch_split_pass(const BwdSplitArgs& a, cudaStream_t s, int which) {
    if constexpr (NCH == 1) {
        switch (split_variant_narrow()) {
            // measured at C2 (cols / rels ms): <4,4> 3.79 / 1.94, <8,3> 4.34 / 2.09, <8,4> 4.27 / 2.01, <16,3> 5.54 / 2.66,
            // <16,2> 6.41 / 2.63: fewer warps lose more than deeper gathers win (short columns: latency chains, not bytes)
            case 1: return launch_split_passes<NCH, HT, 2, 6>(a, s, which);
            case 2: return launch_split_passes<NCH, HT, 2, 8>(a, s, which);
            case 3: return launch_split_passes<NCH, HT, 4, 5>(a, s, which);
            case 4: return launch_split_passes<NCH, HT, 4, 6>(a, s, which);
            default: return launch_split_passes<NCH, HT, 4, 4>(a, s, which);
        }
    } else if constexpr (NCH <= 2) {
        switch (split_variant()) {
            case 1: return launch_split_passes<NCH, HT, 4, 3>(a, s, which);
            case 2: return launch_split_passes<NCH, HT, 2, 4>(a, s, which);
            case 3: return launch_split_passes<NCH, HT, 8, 2>(a, s, which);
            default: return launch_split_passes<NCH, HT, 4, 4>(a, s, which);   // 32 warps/SM measured best (10.4 vs 11.0 / 11.2 / 14.2 ms)
        }
    } else {
        return launch_split_passes<NCH, HT, 2, 2>(a, s, which);
    }
}

template <int NCH, int HT>
int launch_split_t(const BwdSplitArgs& a, cudaStream_t s) {
    if (a.phases & 1) {
        if (int rc = launch_edge_bwd_node(a.f, s)) return rc;
    }
    if (a.phases & 2) {
        if (int rc = launch_split_pass<NCH, HT>(a, s, 0)) return rc;
        if (a.col_hub.n_tasks > 0) {
            SegGatherArgs fa;
            fa.segptr = a.colptr; fa.src = nullptr; fa.pos = nullptr; fa.G = a.G; fa.ldg = a.ldg; fa.rec = nullptr;
            fa.outp = a.dP2; fa.ldout = a.ldd2; fa.n_seg = a.n_cols; fa.prefer_stream = 0; fa.g = a.g; fa.hub = a.col_hub;
            if (int rc = launch_seg_gather_hub_finalize(fa, s)) return rc;
        }
    }
    if (a.phases & 4) {
        if (int rc = launch_split_pass<NCH, HT>(a, s, 1)) return rc;
        if (a.rel_hub.n_tasks > 0) {
            SegGatherArgs fa;
            fa.segptr = a.relptr; fa.src = nullptr; fa.pos = nullptr; fa.G = a.G; fa.ldg = a.ldg; fa.rec = nullptr;
            fa.outp = a.dP3; fa.ldout = a.ldd3; fa.n_seg = a.n_rel; fa.prefer_stream = 0; fa.g = a.g; fa.hub = a.rel_hub;
            if (int rc = launch_seg_gather_hub_finalize(fa, s)) return rc;
        }
        // q slot of dP1~ (rows: contiguous records), or the caller's separate row-sum array
        if (a.rowsum) {
            if (int rc = launch_sums<HT>(a.f.rowptr, nullptr, a.dsv, a.g.H, a.f.n_rows, a.f.row_hub, a.rowsum, a.ld_rowsum, 0, s)) return rc;
        } else {
            if (int rc = launch_sums<HT>(a.f.rowptr, nullptr, a.dsv, a.g.H, a.f.n_rows, a.f.row_hub, a.f.dP1, a.f.ldd1, a.g.Dt4 * 4, s)) return rc;
        }
    }
    if (a.phases & 8) {
        // q slot of dP2~ (columns: records through csc_pos), or the caller's separate column-sum array
        if (a.colsum) return launch_sums<HT>(a.colptr, a.csc_pos, a.dsv, a.g.H, a.n_cols, a.col_hub, a.colsum, a.ld_colsum, 0, s);
        return launch_sums<HT>(a.colptr, a.csc_pos, a.dsv, a.g.H, a.n_cols, a.col_hub, a.dP2, a.ldd2, a.g.Dt4 * 4, s);
    }
    return 0;
}

template <int NCH>
int launch_split_n(const BwdSplitArgs& a, cudaStream_t s) {
    if (a.g.H == 1) return launch_split_t<NCH, 1>(a, s);
    if (a.g.H == 2) return launch_split_t<NCH, 2>(a, s);
    return launch_split_t<NCH, 4>(a, s);
}

}  // namespace

int launch_edge_bwd_split(const BwdSplitArgs& a, cudaStream_t s) {
    switch ((a.g.Wd4 + 31) / 32) {
        case 1: return launch_split_n<1>(a, s);
        case 2: return launch_split_n<2>(a, s);
        case 3: return launch_split_n<3>(a, s);
        case 4: return launch_split_n<4>(a, s);
        default: set_error("edge_bwd_split: row width %d floats exceeds the supported 512", a.g.Wd4 * 4); return 2;
    }
}

}  // namespace spk
