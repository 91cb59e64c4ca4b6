// Aggregate-then-project edge kernels (see spk_agg.cuh): forward, backward over rows, and the small row-wise
// helpers around them. Replaces, for input-narrow layer groups, the same reference lines as K2/K3
// (GAT/layers.py:124-175 and their autograd, SpecialSpmmFunctionFinal.backward layers.py:67-79).
//
// Per edge the warp moves one X~[j] row (lanes 0-15) and one Rel~[k] row (lanes 16-31) as a single
// 128-bit load per lane, eight edges in flight; every per-edge scalar (score, LeakyReLU, exp, dropout
// multiplier, and in the backward ds) is computed lane-parallel for 32 edges at a time. No atomics: a row is
// summed in edge order by one warp, hub rows as fixed 256-edge tasks added in task order.
#include <stdlib.h>
#include "spk_agg.cuh"
#include "spk_stream.cuh"

namespace spk {
namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float agg_exp(float x) { return exp2f(x * 1.4426950408889634f); }
__device__ __forceinline__ float4 f4scale(float4 v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }
__device__ __forceinline__ bool f4nan(float4 v) { return (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---- table builder: T[i] = [ x_i | 0 | x_i . V[:,0..3] | 0 ] ------------------------------------
__global__ void __launch_bounds__(256)
agg_table_kernel(const float* __restrict__ X, long ldx, const float* __restrict__ V, float* __restrict__ T, long ldt,
                 long n, int F, int F4) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* x = X + row * ldx;
    float xv[2];
    float q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int f = lane + 32 * r;
        xv[r] = f < F ? __ldg(x + f) : 0.f;
        if (f < F) {
            const float4 v = ldg4(V + f * 4);
            q[0] = fmaf(xv[r], v.x, q[0]); q[1] = fmaf(xv[r], v.y, q[1]);
            q[2] = fmaf(xv[r], v.z, q[2]); q[3] = fmaf(xv[r], v.w, q[3]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) q[c] = warp_sum(q[c]);
    float* t = T + row * ldt;
    const int qoff = 4 * F4;
    for (int c = lane; c < ldt; c += 32) {
        float o = 0.f;
        if (c < F) o = c < 32 ? xv[0] : xv[1];
        else if (c >= qoff && c < qoff + 4) o = f4get(make_float4(q[0], q[1], q[2], q[3]), c - qoff);
        t[c] = o;
    }
}

// even F <= 64, 8-byte aligned rows, ldt <= 128: a lane produces float2 columns (2 lane, 2 lane + 1) and (+64) of the table
// row; a warp handles AT_R consecutive rows with all row loads issued up front
constexpr int AT_R = 4;
__global__ void __launch_bounds__(256)
agg_table_vec2_kernel(const float* __restrict__ X, long ldx, const float* __restrict__ V, float* __restrict__ T, long ldt,
                      long n, int F, int F4) {
    const int lane = threadIdx.x & 31;
    const long row0 = ((long)blockIdx.x * 8 + (threadIdx.x >> 5)) * AT_R;
    const int f = 2 * lane;
    float2 xv[AT_R];
#pragma unroll
    for (int r = 0; r < AT_R; ++r)
        xv[r] = (row0 + r < n && f < F) ? reinterpret_cast<const float2*>(X + (row0 + r) * ldx)[lane] : make_float2(0.f, 0.f);
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (f < F) { v0 = ldg4(V + f * 4); v1 = ldg4(V + f * 4 + 4); }
    const int qoff = 4 * F4, l2 = (int)(ldt >> 1);
#pragma unroll
    for (int r = 0; r < AT_R; ++r) {
        if (row0 + r >= n) break;
        float q[4];
        q[0] = warp_sum(fmaf(xv[r].x, v0.x, xv[r].y * v1.x)); q[1] = warp_sum(fmaf(xv[r].x, v0.y, xv[r].y * v1.y));
        q[2] = warp_sum(fmaf(xv[r].x, v0.z, xv[r].y * v1.z)); q[3] = warp_sum(fmaf(xv[r].x, v0.w, xv[r].y * v1.w));
        float2* t = reinterpret_cast<float2*>(T + (row0 + r) * ldt);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int c2 = lane + 32 * k, c = 2 * c2;
            if (c2 >= l2) break;
            float2 o = make_float2(0.f, 0.f);
            if (c < F) o = xv[r];
            else if (c == qoff) o = make_float2(q[0], q[1]);
            else if (c == qoff + 2) o = make_float2(q[2], q[3]);
            t[c2] = o;
        }
    }
}

// ---- shared pieces of the streaming edge kernels ---------------------------------------------------
// A warp owns 32 consecutive aggregation rows (or one hub task); their edges form one stream n = 0..T-1 (spk_stream.cuh).
// Three index batches of 32 stream positions are in flight: `cur` (scores, exp, weights ready), `b1` (indices loaded,
// score scalars requested) and `b2` (indices requested), so neither of the two dependent load round trips is exposed.
// The gathered X~[j] / Rel~[k] rows are staged by cp.async.bulk into a per-warp ring of shared-memory slots
// (NG groups of G slots, one mbarrier per group), lanes 0..G-1 issuing one edge each, 16 edges ahead of consumption.
constexpr int AGS_WARPS = 8;

template <int HT>
struct AggRaw { int col, t1, t2, seg, ent; float m[HT]; };
template <int HT>
struct AggScr { float sx[HT], sr[HT]; };

template <int HT, bool HAS2>
__device__ __forceinline__ void ags_load_idx(const int* __restrict__ col, const int* __restrict__ t1, const int* __restrict__ t2,
                                             const float* __restrict__ mask, long mask_stride, int H, const SegTable& st,
                                             int T, int nb0, int lane, AggRaw<HT>& b) {
    b.col = 0; b.t1 = 0; b.t2 = -1; b.seg = 0; b.ent = 0;
#pragma unroll
    for (int h = 0; h < HT; ++h) b.m[h] = 1.f;
    if (nb0 >= T) return;                                       // warp-uniform
    const int m = nb0 + lane;
    int seg;
    const int e = st.entry_of(m < T ? m : T - 1, seg);
    b.seg = seg; b.ent = e;
    if (m < T) {
        b.col = __ldg(col + e);
        b.t1 = __ldg(t1 + e);
        if (HAS2) b.t2 = __ldg(t2 + e);
        if (mask != nullptr) {
#pragma unroll
            for (int h = 0; h < HT; ++h)
                if (h < H) b.m[h] = __ldg(mask + (long)h * mask_stride + e);
        }
    }
}

template <int HT, bool HAS2>
__device__ __forceinline__ void ags_load_scr(const float* __restrict__ Xcol, long ldxc, const float* __restrict__ Rt, long ldr,
                                             int qx, int qr, int T, int nb0, int lane, const AggRaw<HT>& b, AggScr<HT>& s) {
#pragma unroll
    for (int h = 0; h < HT; ++h) { s.sx[h] = 0.f; s.sr[h] = 0.f; }
    if (nb0 + lane < T) {
        const float4 sx = ldg4(Xcol + (long)b.col * ldxc + qx);
        float4 sr = ldg4(Rt + (long)b.t1 * ldr + qr);
        if (HAS2 && b.t2 >= 0) sr = f4add(sr, ldg4(Rt + (long)b.t2 * ldr + qr));
#pragma unroll
        for (int h = 0; h < HT; ++h) { s.sx[h] = f4get(sx, h); s.sr[h] = f4get(sr, h); }
    }
}

// copies of the group of G stream positions starting at n0 (all lanes call; lanes 0..G-1 issue one edge each)
template <int HT, bool HAS2, int G, int NG>
__device__ __forceinline__ void ags_issue_group(const float* __restrict__ Xcol, long ldxc, const float* __restrict__ Rt, long ldr,
                                                uint32_t rbx, uint32_t rbr, uint32_t slot_bytes, uint32_t wbase, uint32_t bars,
                                                int T, int nb, int n0, int lane, const AggRaw<HT>& cur, const AggRaw<HT>& b1) {
    const int n = n0 + lane;
    const int d = n - nb, src = d & 31;
    const int jc = __shfl_sync(FULL, cur.col, src), jn = __shfl_sync(FULL, b1.col, src);
    const int kc = __shfl_sync(FULL, cur.t1, src), kn = __shfl_sync(FULL, b1.t1, src);
    int k2 = -1;
    if (HAS2) {
        const int k2c = __shfl_sync(FULL, cur.t2, src), k2n = __shfl_sync(FULL, b1.t2, src);
        k2 = d < 32 ? k2c : k2n;
    }
    const bool mine = lane < G && n < T;
    const uint32_t bar = bars + 8u * ((n0 / G) % NG);
    unsigned n3 = 0;
    if (HAS2) n3 = __popc(__ballot_sync(FULL, mine && k2 >= 0));
    if (lane == 0) {
        const int cnt = T - n0 < G ? T - n0 : G;
        sbar_expect(bar, (uint32_t)cnt * (rbx + rbr) + n3 * rbr);
    }
    if (mine) {
        const int j = d < 32 ? jc : jn, k1 = d < 32 ? kc : kn;
        const uint32_t slot = wbase + (uint32_t)(n % (G * NG)) * slot_bytes;
        bulk_g2s(slot, Xcol + (long)j * ldxc, rbx, bar);
        bulk_g2s(slot + rbx, Rt + (long)k1 * ldr, rbr, bar);
        if (HAS2 && k2 >= 0) bulk_g2s(slot + rbx + rbr, Rt + (long)k2 * ldr, rbr, bar);
    }
}

template <bool HAS2, int G, int NG>
__host__ __device__ constexpr uint32_t ags_warp_bytes_c(uint32_t rbx, uint32_t rbr) {
    return (uint32_t)(G * NG) * (rbx + rbr * (HAS2 ? 2u : 1u)) + 64u;
}

// ---- forward ---------------------------------------------------------------------------------------
template <int HT>
struct AggAcc {
    float4 acc[HT];
    float den[HT];
    float sw[HT];
};

template <int HT>
__device__ __forceinline__ void agg_acc_init(AggAcc<HT>& st) {
#pragma unroll
    for (int h = 0; h < HT; ++h) { st.acc[h] = make_float4(0.f, 0.f, 0.f, 0.f); st.den[h] = 0.f; st.sw[h] = 0.f; }
}

// den / sw hold the row totals in every lane. Zn_h = [ sw x_i | acc_x | acc_r ] / den  (layers.py:152,169)
template <int HT>
__device__ __forceinline__ void agg_fwd_finalize(const AggFwdArgs& a, int row, int lane, const AggAcc<HT>& st, float4 xi, bool& bad) {
    const int sub = lane & 15;
    const bool isx = lane < 16;
    const bool active = sub < (isx ? a.g.Fx4 : a.g.Fr4);
#pragma unroll
    for (int h = 0; h < HT; ++h) {
        if (h >= a.g.H) break;
        float d = st.den[h];
        if (d == 0.f) d = 1e-12f;
        const float rd = 1.0f / d;
        float* z = a.Z + (long)row * a.ldz + (long)h * a.g.LZ;
        if (lane < a.g.Fx4) {
            const float4 o = f4scale(xi, st.sw[h] * rd);
            bad |= f4nan(o);
            st4(z + lane * 4, o);
        }
        if (active) {
            const float4 o = f4scale(st.acc[h], rd);
            bad |= f4nan(o);
            st4(z + (isx ? 4 : 8) * a.g.Fx4 + sub * 4, o);
        }
        if (lane == 0) { a.den[(long)row * a.g.H + h] = d; a.sw[(long)row * a.g.H + h] = st.sw[h]; }
    }
}

// row without edges: Zn = 0, den -> 1e-12 (layers.py:152), sw = 0
__device__ __forceinline__ void agg_fwd_zero_row(const AggFwdArgs& a, int row, int lane) {
    float* z = a.Z + (long)row * a.ldz;
    const int nz4 = a.g.H * a.g.LZ / 4;
    for (int c = lane; c < nz4; c += 32) st4(z + c * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    if (lane < a.g.H) { a.den[(long)row * a.g.H + lane] = 1e-12f; a.sw[(long)row * a.g.H + lane] = 0.f; }
}

template <int HT, bool HAS2, bool TASKS>
__global__ void __launch_bounds__(AGS_WARPS * 32, 3)
agg_fwd_stream_kernel(const AggFwdArgs a) {
    constexpr int G = 4, NG = HAS2 ? 3 : 4, S = G * NG;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int H = a.g.H;
    const uint32_t rbx = (uint32_t)(a.g.Fx4 + 1) * 16u, rbr = (uint32_t)(a.g.Fr4 + 1) * 16u;
    const uint32_t slot_bytes = rbx + rbr * (HAS2 ? 2u : 1u);
    const uint32_t wbase = smem_addr(smem_raw) + (uint32_t)wid * ags_warp_bytes_c<HAS2, G, NG>(rbx, rbr);
    const uint32_t bars = wbase + S * slot_bytes;
    if (lane == 0) {
        for (int i = 0; i < NG; ++i) sbar_init(bars + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    // ---- segment table ----
    SegTable st;
    st.beg = 0; st.deg = 0;
    bool is_hub = false;
    long row0 = 0;
    int task_row = 0;
    if (!TASKS) {
        row0 = ((long)blockIdx.x * AGS_WARPS + wid) * 32;
        const long r = row0 + lane;
        if (r < a.n_rows) {
            const int b = __ldg(a.segptr + r), e = __ldg(a.segptr + r + 1);
            st.beg = b;
            if (e - b > a.hub.hub_thresh) is_hub = true; else st.deg = e - b;
        }
    } else {
        const int task = blockIdx.x * AGS_WARPS + wid;
        if (task >= a.hub.n_tasks) return;
        task_row = __ldg(a.hub.task_seg + task);
        if (lane == 0) { st.beg = __ldg(a.hub.task_beg + task); st.deg = __ldg(a.hub.task_end + task) - st.beg; }
    }
    st.pre = warp_excl_scan(st.deg, lane, st.total);
    const int T = st.total;
    const int qx = a.g.Fx4 * 4, qr = a.g.Fr4 * 4;
    const bool has_mask = a.mask != nullptr;

    // q1 of this lane's row (X~ scalars are q2_0 q2_1 q1_0 q1_1)
    float q1r[HT];
    {
        float4 xs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (TASKS) xs = ldg4(a.Xrow + (long)task_row * a.ldxr + qx);
        else if (st.deg > 0) xs = ldg4(a.Xrow + (row0 + lane) * a.ldxr + qx);
#pragma unroll
        for (int h = 0; h < HT; ++h) q1r[h] = f4get(xs, 2 + h);
    }

    if (!TASKS) {                                                  // empty rows are written up front
        unsigned empt = __ballot_sync(FULL, st.deg == 0 && !is_hub && row0 + lane < a.n_rows);
        while (empt) { const int r = __ffs(empt) - 1; empt &= empt - 1; agg_fwd_zero_row(a, (int)(row0 + r), lane); }
    }
    if (T == 0) return;

    AggRaw<HT> cur, b1, b2;
    AggScr<HT> s1;
    float w[HT], ee[HT];
    bool bad = false;
    auto compute = [&](int nb0, const AggRaw<HT>& b, const AggScr<HT>& s) {      // cur weights from (b, s); all lanes
#pragma unroll
        for (int h = 0; h < HT; ++h) {
            const float q1 = __shfl_sync(FULL, q1r[h], b.seg);
            const float sc = q1 + (s.sx[h] + s.sr[h]);
            float e = agg_exp(-(sc > 0.f ? sc : a.alpha * sc));                  // layers.py:143-146
            if (nb0 + lane >= T || h >= H) e = 0.f;
            bad |= (e != e);
            ee[h] = e;
            w[h] = e * b.m[h];                                                   // layers.py:158
        }
    };
    ags_load_idx<HT, HAS2>(a.col, a.t1, a.t2, a.mask, a.mask_stride, H, st, T, 0, lane, cur);
    ags_load_idx<HT, HAS2>(a.col, a.t1, a.t2, a.mask, a.mask_stride, H, st, T, 32, lane, b1);
    ags_load_idx<HT, HAS2>(a.col, a.t1, a.t2, a.mask, a.mask_stride, H, st, T, 64, lane, b2);
    {
        AggScr<HT> s0;
        ags_load_scr<HT, HAS2>(a.Xcol, a.ldxc, a.Rt, a.ldr, qx, qr, T, 0, lane, cur, s0);
        ags_load_scr<HT, HAS2>(a.Xcol, a.ldxc, a.Rt, a.ldr, qx, qr, T, 32, lane, b1, s1);
        compute(0, cur, s0);
    }
    int nb = 0;
    for (int gi = 0; gi < NG && gi * G < T; ++gi)
        ags_issue_group<HT, HAS2, G, NG>(a.Xcol, a.ldxc, a.Rt, a.ldr, rbx, rbr, slot_bytes, wbase, bars, T, nb, gi * G, lane, cur, b1);

    // ---- row tracking ----
    unsigned act = TASKS ? 1u : __ballot_sync(FULL, st.deg > 0);
    int r = __ffs(act) - 1;                                        // current segment (lane index); T > 0 -> valid
    act &= act - 1;
    int row_end = __shfl_sync(FULL, st.pre + st.deg, r);
    auto load_xi = [&](int seg) {
        const long row = TASKS ? (long)task_row : row0 + seg;
        return lane < a.g.Fx4 ? ldg4(a.Xrow + row * a.ldxr + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 xi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!TASKS) xi = load_xi(r);

    const int sub = lane & 15;
    const bool isx = lane < 16;
    const bool active = sub < (isx ? a.g.Fx4 : a.g.Fr4);
    const uint32_t loff = (isx ? 0u : rbx) + (uint32_t)sub * 16u;
    AggAcc<HT> acc;
    agg_acc_init<HT>(acc);

    for (int n0 = 0; n0 < T; n0 += G) {
        if (n0 - nb == 32) {                                       // rotate the index batches
            nb += 32;
            cur = b1;
            compute(nb, b1, s1);
            b1 = b2;
            ags_load_scr<HT, HAS2>(a.Xcol, a.ldxc, a.Rt, a.ldr, qx, qr, T, nb + 32, lane, b1, s1);
            ags_load_idx<HT, HAS2>(a.col, a.t1, a.t2, a.mask, a.mask_stride, H, st, T, nb + 64, lane, b2);
        }
        const int grp = (n0 / G) % NG;
        sbar_wait(bars + 8u * grp, (uint32_t)(n0 / S) & 1u);
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const int n = n0 + u;
            if (n >= T) break;
            const uint32_t slot = wbase + (uint32_t)(grp * G + u) * slot_bytes;
            const int src = n - nb;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) v = lds4(slot + loff);
            if (HAS2) {
                const int k2 = __shfl_sync(FULL, cur.t2, src);
                if (active && !isx && k2 >= 0) v = f4add(v, lds4(slot + loff + rbr));
            }
#pragma unroll
            for (int h = 0; h < HT; ++h) {
                const float wh = __shfl_sync(FULL, w[h], src);
                f4fma(acc.acc[h], wh, v);
                acc.sw[h] += wh;
                acc.den[h] += has_mask ? __shfl_sync(FULL, ee[h], src) : wh;
            }
            if (!TASKS && n + 1 == row_end) {                      // row complete -> finalize, move to the next active row
                agg_fwd_finalize<HT>(a, (int)(row0 + r), lane, acc, xi, bad);
                agg_acc_init<HT>(acc);
                if (act) {
                    r = __ffs(act) - 1;
                    act &= act - 1;
                    row_end = __shfl_sync(FULL, st.pre + st.deg, r);
                    xi = load_xi(r);
                }
            }
        }
        __syncwarp();                                              // whole group consumed -> refill its slots
        if (n0 + S < T)
            ags_issue_group<HT, HAS2, G, NG>(a.Xcol, a.ldxc, a.Rt, a.ldr, rbx, rbr, slot_bytes, wbase, bars, T, nb, n0 + S, lane, cur, b1);
    }
    if (TASKS) {
        const int task = blockIdx.x * AGS_WARPS + wid;
        float* part = a.hub.partial + (long)task * a.hub.ldpart;
#pragma unroll
        for (int h = 0; h < 2; ++h) st4(part + h * 128 + lane * 4, h < HT ? acc.acc[h < HT ? h : 0] : make_float4(0.f, 0.f, 0.f, 0.f));
        if (lane < 4) {
            part[256 + lane] = lane < HT ? selh<HT>(lane, acc.den) : 0.f;
            part[260 + lane] = lane < HT ? selh<HT>(lane, acc.sw) : 0.f;
        }
    }
    if (__any_sync(FULL, bad) && lane == 0) atomicOr(a.nanflag, 1);
}


// ---- register-gather forward (round 2) --------------------------------------------------------------
// Same work as agg_fwd_stream_kernel, without the shared-memory ring. ncu's instruction mix of the ring version showed
// 73 warp instructions per edge of which 8.5 % were FFMA: issuing two cp.async.bulk copies per edge costs an ELECT loop
// with four R2UR each (UBLKCP takes uniform registers) plus the mbarrier bookkeeping. A gathered row here is ONE float4
// per lane (X~[j] chunk in lanes 0-15, Rel~[k] chunk in lanes 16-31), so a plain LDG.128 per lane per edge with AG_U edges
// in flight costs AG_U registers per lane, not the 16+ the 832-byte rows of K2 needed: the latency depth fits registers.
constexpr int AGR_WARPS = 8;

template <int HT, bool HAS2, bool TASKS, int AG_U, int MINB>
__global__ void __launch_bounds__(AGR_WARPS * 32, MINB)
agg_fwd_reg_kernel(const AggFwdArgs a) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int H = a.g.H;
    SegTable st;
    st.beg = 0; st.deg = 0;
    bool is_hub = false;
    long row0 = 0;
    int task_row = 0;
    if (!TASKS) {
        row0 = ((long)blockIdx.x * AGR_WARPS + wid) * 32;
        const long r = row0 + lane;
        if (r < a.n_rows) {
            const int b = __ldg(a.segptr + r), e = __ldg(a.segptr + r + 1);
            st.beg = b;
            if (e - b > a.hub.hub_thresh) is_hub = true; else st.deg = e - b;
        }
    } else {
        const int task = blockIdx.x * AGR_WARPS + wid;
        if (task >= a.hub.n_tasks) return;
        task_row = __ldg(a.hub.task_seg + task);
        if (lane == 0) { st.beg = __ldg(a.hub.task_beg + task); st.deg = __ldg(a.hub.task_end + task) - st.beg; }
    }
    st.pre = warp_excl_scan(st.deg, lane, st.total);
    const int T = st.total;
    const int qx = a.g.Fx4 * 4, qr = a.g.Fr4 * 4;
    const bool has_mask = a.mask != nullptr;

    float q1r[HT];
    {
        float4 xs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (TASKS) xs = ldg4(a.Xrow + (long)task_row * a.ldxr + qx);
        else if (st.deg > 0) xs = ldg4(a.Xrow + (row0 + lane) * a.ldxr + qx);
#pragma unroll
        for (int h = 0; h < HT; ++h) q1r[h] = f4get(xs, 2 + h);
    }
    if (!TASKS) {
        unsigned empt = __ballot_sync(FULL, st.deg == 0 && !is_hub && row0 + lane < a.n_rows);
        while (empt) { const int r = __ffs(empt) - 1; empt &= empt - 1; agg_fwd_zero_row(a, (int)(row0 + r), lane); }
    }
    if (T == 0) return;

    AggRaw<HT> cur, nxt;
    AggScr<HT> scur, snxt;
    float w[HT], ee[HT];
    bool bad = false;
    auto compute = [&](int nb0, const AggRaw<HT>& b, const AggScr<HT>& s) {
#pragma unroll
        for (int h = 0; h < HT; ++h) {
            const float q1 = __shfl_sync(FULL, q1r[h], b.seg);
            const float sc = q1 + (s.sx[h] + s.sr[h]);
            float e = agg_exp(-(sc > 0.f ? sc : a.alpha * sc));                  // layers.py:143-146
            if (nb0 + lane >= T || h >= H) e = 0.f;
            bad |= (e != e);
            ee[h] = e;
            w[h] = e * b.m[h];                                                   // layers.py:158
        }
    };
    ags_load_idx<HT, HAS2>(a.col, a.t1, a.t2, a.mask, a.mask_stride, H, st, T, 0, lane, cur);
    ags_load_scr<HT, HAS2>(a.Xcol, a.ldxc, a.Rt, a.ldr, qx, qr, T, 0, lane, cur, scur);
    ags_load_idx<HT, HAS2>(a.col, a.t1, a.t2, a.mask, a.mask_stride, H, st, T, 32, lane, nxt);

    unsigned act = TASKS ? 1u : __ballot_sync(FULL, st.deg > 0);
    int r = __ffs(act) - 1;
    act &= act - 1;
    int row_end = __shfl_sync(FULL, st.pre + st.deg, r);
    auto load_xi = [&](int seg) {
        const long row = TASKS ? (long)task_row : row0 + seg;
        return lane < a.g.Fx4 ? ldg4(a.Xrow + row * a.ldxr + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 xi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!TASKS) xi = load_xi(r);

    const int sub = lane & 15;
    const bool isx = lane < 16;
    const bool active = sub < (isx ? a.g.Fx4 : a.g.Fr4);
    // this lane's table and chunk offset: X~ rows for lanes 0-15, Rel~ rows for lanes 16-31
    const float* tab = (isx ? a.Xcol : a.Rt) + sub * 4;
    const long ldt = isx ? a.ldxc : a.ldr;
    AggAcc<HT> acc;
    agg_acc_init<HT>(acc);

    for (int nb = 0; nb < T; nb += 32) {
        // score scalars of the NEXT batch are requested now and consumed after this batch's rows (one round trip hidden)
        ags_load_scr<HT, HAS2>(a.Xcol, a.ldxc, a.Rt, a.ldr, qx, qr, T, nb + 32, lane, nxt, snxt);
        compute(nb, cur, scur);
        const int nvalid = T - nb < 32 ? T - nb : 32;
        for (int u0 = 0; u0 < nvalid; u0 += AG_U) {
            float4 v[AG_U];
#pragma unroll
            for (int k = 0; k < AG_U; ++k) {
                const int src = u0 + k;
                const int j = __shfl_sync(FULL, cur.col, src & 31), k1 = __shfl_sync(FULL, cur.t1, src & 31);
                const int idx = isx ? j : k1;
                v[k] = (src < nvalid && active) ? ldg4(tab + (long)idx * ldt) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (HAS2) {
                    const int k2 = __shfl_sync(FULL, cur.t2, src & 31);
                    if (src < nvalid && active && !isx && k2 >= 0) v[k] = f4add(v[k], ldg4(tab + (long)k2 * ldt));
                }
            }
#pragma unroll
            for (int k = 0; k < AG_U; ++k) {
                const int src = u0 + k;
                if (src >= nvalid) break;
                const int n = nb + src;
#pragma unroll
                for (int h = 0; h < HT; ++h) {
                    const float wh = __shfl_sync(FULL, w[h], src);
                    f4fma(acc.acc[h], wh, v[k]);
                    acc.sw[h] += wh;
                    acc.den[h] += has_mask ? __shfl_sync(FULL, ee[h], src) : wh;
                }
                if (!TASKS && n + 1 == row_end) {
                    agg_fwd_finalize<HT>(a, (int)(row0 + r), lane, acc, xi, bad);
                    agg_acc_init<HT>(acc);
                    if (act) {
                        r = __ffs(act) - 1;
                        act &= act - 1;
                        row_end = __shfl_sync(FULL, st.pre + st.deg, r);
                        xi = load_xi(r);
                    }
                }
            }
        }
        cur = nxt; scur = snxt;
        ags_load_idx<HT, HAS2>(a.col, a.t1, a.t2, a.mask, a.mask_stride, H, st, T, nb + 64, lane, nxt);
    }
    if (TASKS) {
        const int task = blockIdx.x * AGR_WARPS + wid;
        float* part = a.hub.partial + (long)task * a.hub.ldpart;
#pragma unroll
        for (int h = 0; h < 2; ++h) st4(part + h * 128 + lane * 4, h < HT ? acc.acc[h < HT ? h : 0] : make_float4(0.f, 0.f, 0.f, 0.f));
        if (lane < 4) {
            part[256 + lane] = lane < HT ? selh<HT>(lane, acc.den) : 0.f;
            part[260 + lane] = lane < HT ? selh<HT>(lane, acc.sw) : 0.f;
        }
    }
    if (__any_sync(FULL, bad) && lane == 0) atomicOr(a.nanflag, 1);
}

// one CTA per hub row: partials added in the fixed order of cta_sum_partials, then the row epilogue
template <int HT>
__global__ void __launch_bounds__(1024)
agg_fwd_hub_finalize_kernel(const AggFwdArgs a) {
    __shared__ __align__(16) float red[32][AGG_LDPART];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int hub = blockIdx.x;
    const int row = __ldg(a.hub.hub_seg + hub);
    const int t0 = __ldg(a.hub.hub_task_ptr + hub), t1 = __ldg(a.hub.hub_task_ptr + hub + 1);
    cta_sum_partials<(AGG_LDPART + 31) / 32, 32>(a.hub.partial, a.hub.ldpart, t0, t1, AGG_LDPART, &red[0][0], AGG_LDPART);
    if (wid != 0) return;
    AggAcc<HT> st;
#pragma unroll
    for (int h = 0; h < HT; ++h) {
        st.acc[h] = *reinterpret_cast<const float4*>(&red[0][h * 128 + lane * 4]);
        st.den[h] = red[0][256 + h];
        st.sw[h] = red[0][260 + h];
    }
    bool bad = false;
    float4 xi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < a.g.Fx4) xi = ldg4(a.Xrow + (long)row * a.ldxr + lane * 4);
    agg_fwd_finalize<HT>(a, row, lane, st, xi, bad);
    if (__any_sync(FULL, bad) && lane == 0) atomicOr(a.nanflag, 1);
}

// SPK_AGG_FWD=ring selects the shared-memory ring kernels (round 1); default: register gathers
static bool agg_fwd_ring() {
    const char* e = getenv("SPK_AGG_FWD");
    return e && e[0] == 'r';
}

static int agg_fwd_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SPK_AGG_FWD_VARIANT"); v = e ? atoi(e) : 0; if (v < 0 || v > 2) v = 0; }
    return v;
}

template <int HT, bool HAS2>
int launch_agg_fwd_t(const AggFwdArgs& a, cudaStream_t s) {
    if (!agg_fwd_ring()) {
        if (a.n_rows > 0) {
            const unsigned grid = (unsigned)((a.n_rows + 32L * AGR_WARPS - 1) / (32L * AGR_WARPS));
            // (gathers in flight per warp, CTAs of 8 warps per SM), measured at C2: <4,3> 2.82 ms, <8,2> 3.01, <4,4> 3.16,
            // <8,3> 3.36, <16,2> 3.63 -- neither more warps nor deeper gathers is the limiter here
            switch (agg_fwd_variant()) {
                case 1: agg_fwd_reg_kernel<HT, HAS2, false, 8, 2><<<grid, AGR_WARPS * 32, 0, s>>>(a); break;
                case 2: agg_fwd_reg_kernel<HT, HAS2, false, 4, 4><<<grid, AGR_WARPS * 32, 0, s>>>(a); break;
                default: agg_fwd_reg_kernel<HT, HAS2, false, 4, 3><<<grid, AGR_WARPS * 32, 0, s>>>(a); break;
            }
            if (int rc = check_launch("agg_fwd_reg_rows")) return rc;
        }
        if (a.hub.n_tasks > 0) {
            const unsigned grid = (a.hub.n_tasks + AGR_WARPS - 1) / AGR_WARPS;
            agg_fwd_reg_kernel<HT, HAS2, true, 8, 2><<<grid, AGR_WARPS * 32, 0, s>>>(a);
            if (int rc = check_launch("agg_fwd_reg_tasks")) return rc;
            agg_fwd_hub_finalize_kernel<HT><<<a.hub.n_hubs, 1024, 0, s>>>(a);
            if (int rc = check_launch("agg_fwd_hub_finalize")) return rc;
        }
        return 0;
    }
    constexpr int G = 4, NG = HAS2 ? 3 : 4;
    const uint32_t rbx = (uint32_t)(a.g.Fx4 + 1) * 16u, rbr = (uint32_t)(a.g.Fr4 + 1) * 16u;
    const size_t smem = (size_t)AGS_WARPS * ags_warp_bytes_c<HAS2, G, NG>(rbx, rbr);
    static SmemLimit lim_rows, lim_tasks;
    if (a.n_rows > 0) {
        lim_rows.ensure(agg_fwd_stream_kernel<HT, HAS2, false>, smem);
        const unsigned grid = (unsigned)((a.n_rows + 32L * AGS_WARPS - 1) / (32L * AGS_WARPS));
        agg_fwd_stream_kernel<HT, HAS2, false><<<grid, AGS_WARPS * 32, smem, s>>>(a);
        if (int rc = check_launch("agg_fwd_stream_rows")) return rc;
    }
    if (a.hub.n_tasks > 0) {
        lim_tasks.ensure(agg_fwd_stream_kernel<HT, HAS2, true>, smem);
        const unsigned grid = (a.hub.n_tasks + AGS_WARPS - 1) / AGS_WARPS;
        agg_fwd_stream_kernel<HT, HAS2, true><<<grid, AGS_WARPS * 32, smem, s>>>(a);
        if (int rc = check_launch("agg_fwd_stream_tasks")) return rc;
        agg_fwd_hub_finalize_kernel<HT><<<a.hub.n_hubs, 1024, 0, s>>>(a);
        if (int rc = check_launch("agg_fwd_hub_finalize")) return rc;
    }
    return 0;
}

// ---- backward --------------------------------------------------------------------------------------
//   Y_h = dZn_h / den_h = [Ya | Yb | Yc];  c_h = Ya.x_i;  per edge t = c + Yb.x_j + Yc.r_k
//   ds = -(msk t + dden) ee LeakyReLU'(s);  rec = (w, ds);  dq1 = sum ds;  dX_i(row part) = sum_h sw_h Ya_h
//
// Row-context kernel (one warp per row, pure streaming): from dZn, den, sw, dden, X~ it emits
//   Gx[i] = Yb per head, Gr[i] = Yc per head (the rows the column / relation passes gather, and the per-row context of
//   the edge kernel), rowout[i, :4*Fx4] = sum_h sw_h Ya_h, rowsc[i] = (q1_0, q1_1, c_0, c_1, dden_0, dden_1, 0, 0).
template <bool SPLIT_LAYOUT>
__global__ void __launch_bounds__(SPK_CTA_THREADS)
agg_bwd_ctx_kernel(const AggBwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * SPK_WARPS_PER_CTA + (threadIdx.x >> 5);
    if (row >= a.n_rows) return;
    const int sub = lane & 15;
    const bool isx = lane < 16;
    const bool active = sub < (isx ? a.g.Fx4 : a.g.Fr4);
    const int H = a.g.H;
    // every load of the row is issued before the first use (one DRAM round trip per row)
    const float4 xs = ldg4(a.Xrow + (long)row * a.ldxr + a.g.Fx4 * 4);
    float4 xi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < a.g.Fx4) xi = ldg4(a.Xrow + (long)row * a.ldxr + lane * 4);
    float4 ya[2], yy[2];
    float den[2] = {1.f, 1.f}, swh[2] = {0.f, 0.f}, dd[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        ya[h] = make_float4(0.f, 0.f, 0.f, 0.f);
        yy[h] = ya[h];
        if (h < H) {
            const float* dz = a.dZ + (long)row * a.ldz + (long)h * a.g.LZ;
            den[h] = __ldg(a.den + (long)row * H + h);
            swh[h] = __ldg(a.sw + (long)row * H + h);
            dd[h] = __ldg(a.dden + (long)row * H + h);
            if (lane < a.g.Fx4) ya[h] = ldg4(dz + lane * 4);
            if (active) yy[h] = ldg4(dz + (isx ? 4 : 8) * a.g.Fx4 + sub * 4);
        }
    }
    float4 dxr = make_float4(0.f, 0.f, 0.f, 0.f);
    float c[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (h >= H) break;
        const float rd = 1.0f / den[h];
        const float4 y0 = f4scale(ya[h], rd);
        c[h] = warp_sum(f4dot(y0, xi));
        f4fma(dxr, swh[h], y0);
        if (active) {
            const float4 y = f4scale(yy[h], rd);
            if (isx) st4(a.Gx + (long)row * a.ldgx + (long)h * 4 * a.g.Fx4 + sub * 4, y);
            else st4(a.Gr + (long)row * a.ldgr + (long)h * 4 * a.g.Fr4 + sub * 4, y);
        }
    }
    if (lane < a.g.Fx4) st4(a.rowout + (long)row * a.ldro + lane * 4, dxr);
    if (lane == 0) {
        if (SPLIT_LAYOUT) {                        // [n_rows, H, 4] = (q1, c, dden, 0): what the split-dot passes read per row
            st4(a.rowsc + (long)row * 4 * H, make_float4(xs.z, c[0], dd[0], 0.f));
            if (H > 1) st4(a.rowsc + (long)row * 4 * H + 4, make_float4(xs.w, c[1], dd[1], 0.f));
        } else {
            st4(a.rowsc + (long)row * 8, make_float4(xs.z, xs.w, c[0], c[1]));
            st4(a.rowsc + (long)row * 8 + 4, make_float4(dd[0], dd[1], 0.f, 0.f));
        }
    }
}

template <int HT, bool HAS2, bool TASKS>
__global__ void __launch_bounds__(AGS_WARPS * 32, 2)
agg_bwd_stream_kernel(const AggBwdArgs a) {
    constexpr int G = 8, NG = 2, S = G * NG, NV = G * HT;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int H = a.g.H;
    const uint32_t rbx = (uint32_t)(a.g.Fx4 + 1) * 16u, rbr = (uint32_t)(a.g.Fr4 + 1) * 16u;
    const uint32_t slot_bytes = rbx + rbr * (HAS2 ? 2u : 1u);
    const uint32_t wbase = smem_addr(smem_raw) + (uint32_t)wid * ags_warp_bytes_c<HAS2, G, NG>(rbx, rbr);
    const uint32_t bars = wbase + S * slot_bytes;
    if (lane == 0) {
        for (int i = 0; i < NG; ++i) sbar_init(bars + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    SegTable st;
    st.beg = 0; st.deg = 0;
    bool is_hub = false;
    long row0 = 0;
    int task_row = 0;
    if (!TASKS) {
        row0 = ((long)blockIdx.x * AGS_WARPS + wid) * 32;
        const long r = row0 + lane;
        if (r < a.n_rows) {
            const int b = __ldg(a.segptr + r), e = __ldg(a.segptr + r + 1);
            st.beg = b;
            if (e - b > a.hub.hub_thresh) is_hub = true; else st.deg = e - b;
        }
    } else {
        const int task = blockIdx.x * AGS_WARPS + wid;
        if (task >= a.hub.n_tasks) return;
        task_row = __ldg(a.hub.task_seg + task);
        if (lane == 0) { st.beg = __ldg(a.hub.task_beg + task); st.deg = __ldg(a.hub.task_end + task) - st.beg; }
    }
    st.pre = warp_excl_scan(st.deg, lane, st.total);
    const int T = st.total;
    const int qx = a.g.Fx4 * 4, qr = a.g.Fr4 * 4;
    const bool row_valid = !TASKS && !is_hub && row0 + lane < a.n_rows;

    float usum[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) usum[h] = 0.f;
    if (T > 0) {
        // per-row scalars of this lane's row
        float q1r[HT], cr[HT], ddr[HT];
        {
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1v = s0;
            const long srow = TASKS ? (long)task_row : row0 + lane;
            if (TASKS || st.deg > 0) { s0 = ldg4(a.rowsc + srow * 8); s1v = ldg4(a.rowsc + srow * 8 + 4); }
#pragma unroll
            for (int h = 0; h < HT; ++h) { q1r[h] = f4get(s0, h); cr[h] = f4get(s0, 2 + h); ddr[h] = f4get(s1v, h); }
        }
        AggRaw<HT> cur, b1, b2;
        AggScr<HT> s1;
        float w[HT], ca[HT], cb[HT], tt[HT];
        auto compute = [&](int nb0, const AggRaw<HT>& b, const AggScr<HT>& s) {
#pragma unroll
            for (int h = 0; h < HT; ++h) {
                const float q1 = __shfl_sync(FULL, q1r[h], b.seg);
                const float ce = __shfl_sync(FULL, cr[h], b.seg);
                const float dde = __shfl_sync(FULL, ddr[h], b.seg);
                const float sc = q1 + (s.sx[h] + s.sr[h]);
                const float slope = sc > 0.f ? 1.f : a.alpha;
                float e = agg_exp(-(sc * slope));
                if (nb0 + lane >= T || h >= H) e = 0.f;
                w[h] = e * b.m[h];
                ca[h] = w[h] * slope;                               // ds = -(ca * (c + dot) + cb) = -(ca * dot + (ca * c + cb))
                cb[h] = fmaf(ca[h], ce, e * slope * dde);
                tt[h] = 0.f;
            }
        };
        ags_load_idx<HT, HAS2>(a.col, a.t1, a.t2, a.mask, a.mask_stride, H, st, T, 0, lane, cur);
        ags_load_idx<HT, HAS2>(a.col, a.t1, a.t2, a.mask, a.mask_stride, H, st, T, 32, lane, b1);
        ags_load_idx<HT, HAS2>(a.col, a.t1, a.t2, a.mask, a.mask_stride, H, st, T, 64, lane, b2);
        {
            AggScr<HT> s0;
            ags_load_scr<HT, HAS2>(a.Xcol, a.ldxc, a.Rt, a.ldr, qx, qr, T, 0, lane, cur, s0);
            ags_load_scr<HT, HAS2>(a.Xcol, a.ldxc, a.Rt, a.ldr, qx, qr, T, 32, lane, b1, s1);
            compute(0, cur, s0);
        }
        int nb = 0;
        for (int gi = 0; gi < NG && gi * G < T; ++gi)
            ags_issue_group<HT, HAS2, G, NG>(a.Xcol, a.ldxc, a.Rt, a.ldr, rbx, rbr, slot_bytes, wbase, bars, T, nb, gi * G, lane, cur, b1);

        // ---- row tracking; the context Y of the next active row is prefetched one row ahead ----
        const int sub = lane & 15;
        const bool isx = lane < 16;
        const bool active = sub < (isx ? a.g.Fx4 : a.g.Fr4);
        const uint32_t loff = (isx ? 0u : rbx) + (uint32_t)sub * 16u;
        const float* ybase = isx ? a.Gx : a.Gr;
        const long yld = isx ? a.ldgx : a.ldgr;
        const int ych = 4 * (isx ? a.g.Fx4 : a.g.Fr4);
        auto load_y = [&](int seg, float4 (&y)[HT]) {
            const long row = TASKS ? (long)task_row : row0 + seg;
#pragma unroll
            for (int h = 0; h < HT; ++h)
                y[h] = (active && h < H) ? ldg4(ybase + row * yld + (long)h * ych + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        unsigned act = TASKS ? 1u : __ballot_sync(FULL, st.deg > 0);
        int r = __ffs(act) - 1;
        act &= act - 1;
        int row_end = __shfl_sync(FULL, st.pre + st.deg, r);
        float4 Y[HT], Yn[HT];
        load_y(r, Y);
        int rn = -1;
        if (act) { rn = __ffs(act) - 1; act &= act - 1; load_y(rn, Yn); }

        auto batch_epilogue = [&]() {                              // ds, records and the per-row sums of the batch `cur`
            float ds[HT];
#pragma unroll
            for (int h = 0; h < HT; ++h) ds[h] = -fmaf(ca[h], tt[h], cb[h]);
            if (nb + lane < T) {
                float* rp = a.rec + (long)cur.ent * (2 * H);
                if (HT == 2) st4(rp, make_float4(w[0], ds[0], w[HT - 1], ds[HT - 1]));
                else *reinterpret_cast<float2*>(rp) = make_float2(w[0], ds[0]);
            } else {
#pragma unroll
                for (int h = 0; h < HT; ++h) ds[h] = 0.f;
            }
            // segmented inclusive scan keyed on the (non-decreasing) segment index, then each row owner picks its run's total
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int ps = __shfl_up_sync(FULL, cur.seg, off);
                const bool take = lane >= off && ps == cur.seg;
#pragma unroll
                for (int h = 0; h < HT; ++h) {
                    const float t = __shfl_up_sync(FULL, ds[h], off);
                    if (take) ds[h] += t;
                }
            }
            const int last = st.pre + st.deg - 1 - nb;
            const bool hit = st.deg > 0 && st.pre < nb + 32 && last >= 0;
            const int pick = hit ? (last < 31 ? last : 31) : 0;
#pragma unroll
            for (int h = 0; h < HT; ++h) {
                const float t = __shfl_sync(FULL, ds[h], pick);
                if (hit) usum[h] += t;
            }
        };

        for (int n0 = 0; n0 < T; n0 += G) {
            if (n0 - nb == 32) {
                batch_epilogue();
                nb += 32;
                cur = b1;
                compute(nb, b1, s1);
                b1 = b2;
                ags_load_scr<HT, HAS2>(a.Xcol, a.ldxc, a.Rt, a.ldr, qx, qr, T, nb + 32, lane, b1, s1);
                ags_load_idx<HT, HAS2>(a.col, a.t1, a.t2, a.mask, a.mask_stride, H, st, T, nb + 64, lane, b2);
            }
            const int grp = (n0 / G) % NG;
            sbar_wait(bars + 8u * grp, (uint32_t)(n0 / S) & 1u);
            float pd[NV];
#pragma unroll
            for (int u = 0; u < G; ++u) {
                const int n = n0 + u;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < T) {                                       // warp-uniform
                    const uint32_t slot = wbase + (uint32_t)(grp * G + u) * slot_bytes;
                    if (active) v = lds4(slot + loff);
                    if (HAS2) {
                        const int k2 = __shfl_sync(FULL, cur.t2, n - nb);
                        if (active && !isx && k2 >= 0) v = f4add(v, lds4(slot + loff + rbr));
                    }
                }
#pragma unroll
                for (int h = 0; h < HT; ++h) pd[u * HT + h] = f4dot(Y[h], v);
                if (!TASKS && n + 1 == row_end && rn >= 0) {       // row complete -> switch to the prefetched context
#pragma unroll
                    for (int h = 0; h < HT; ++h) Y[h] = Yn[h];
                    r = rn;
                    row_end = __shfl_sync(FULL, st.pre + st.deg, r);
                    rn = -1;
                    if (act) { rn = __ffs(act) - 1; act &= act - 1; load_y(rn, Yn); }
                }
            }
            const float tot = transposed_warp_sum<NV>(pd, lane);    // lane L: value index (L * NV) >> 5 = u * HT + h
            const bool mine = (lane >> 3) == ((n0 - nb) >> 3);     // lane-edge (n0 - nb) + (lane & 7)
#pragma unroll
            for (int h = 0; h < HT; ++h) {
                const float t = __shfl_sync(FULL, tot, (((lane & 7) * HT + h) * 32) / NV);
                if (mine) tt[h] = t;
            }
            __syncwarp();
            if (n0 + S < T)
                ags_issue_group<HT, HAS2, G, NG>(a.Xcol, a.ldxc, a.Rt, a.ldr, rbx, rbr, slot_bytes, wbase, bars, T, nb, n0 + S, lane, cur, b1);
        }
        batch_epilogue();
    }
    if (TASKS) {
        const int task = blockIdx.x * AGS_WARPS + wid;
        if (lane == 0) {
            float* part = a.hub.partial + (long)task * a.hub.ldpart;
#pragma unroll
            for (int h = 0; h < 4; ++h) part[h] = h < HT ? usum[h < HT ? h : 0] : 0.f;
        }
    } else if (row_valid) {
        st4(a.rowout + (row0 + lane) * a.ldro + 4 * a.g.Fx4, make_float4(usum[0], HT > 1 ? usum[HT - 1] : 0.f, 0.f, 0.f));
    }
}

__global__ void __launch_bounds__(SPK_CTA_THREADS)
agg_bwd_hub_finalize_kernel(const AggBwdArgs a) {
    __shared__ float red[SPK_WARPS_PER_CTA][8];
    const int hub = blockIdx.x;
    const int row = __ldg(a.hub.hub_seg + hub);
    const int t0 = __ldg(a.hub.hub_task_ptr + hub), t1 = __ldg(a.hub.hub_task_ptr + hub + 1);
    cta_sum_partials<1>(a.hub.partial, a.hub.ldpart, t0, t1, 4, &red[0][0], 8);
    if (threadIdx.x == 0)
        st4(a.rowout + (long)row * a.ldro + 4 * a.g.Fx4, make_float4(red[0][0], red[0][1], 0.f, 0.f));
}

template <int HT, bool HAS2>
int launch_agg_bwd_t(const AggBwdArgs& a, cudaStream_t s) {
    constexpr int G = 8, NG = 2;
    const uint32_t rbx = (uint32_t)(a.g.Fx4 + 1) * 16u, rbr = (uint32_t)(a.g.Fr4 + 1) * 16u;
    const size_t smem = (size_t)AGS_WARPS * ags_warp_bytes_c<HAS2, G, NG>(rbx, rbr);
    static SmemLimit lim_rows, lim_tasks;
    if (a.n_rows > 0) {
        agg_bwd_ctx_kernel<false><<<(a.n_rows + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("agg_bwd_ctx")) return rc;
        lim_rows.ensure(agg_bwd_stream_kernel<HT, HAS2, false>, smem);
        const unsigned grid = (unsigned)((a.n_rows + 32L * AGS_WARPS - 1) / (32L * AGS_WARPS));
        agg_bwd_stream_kernel<HT, HAS2, false><<<grid, AGS_WARPS * 32, smem, s>>>(a);
        if (int rc = check_launch("agg_bwd_stream_rows")) return rc;
    }
    if (a.hub.n_tasks > 0) {
        lim_tasks.ensure(agg_bwd_stream_kernel<HT, HAS2, true>, smem);
        const unsigned grid = (a.hub.n_tasks + AGS_WARPS - 1) / AGS_WARPS;
        agg_bwd_stream_kernel<HT, HAS2, true><<<grid, AGS_WARPS * 32, smem, s>>>(a);
        if (int rc = check_launch("agg_bwd_stream_tasks")) return rc;
        agg_bwd_hub_finalize_kernel<<<a.hub.n_hubs, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("agg_bwd_hub_finalize")) return rc;
    }
    return 0;
}

// ---- row-wise helpers ------------------------------------------------------------------------------
// dhn = dOut * ELU'(hn) from the saved output (hn > 0: out = hn; else out = e^hn - 1, ELU' = out + 1, hn = log1p(out));
// dden_h = -(dhn_h . hn_h) / den_h
__device__ __forceinline__ void bwd_pre_elem(float ov, float gv, int apply_elu, float& dh, float& p) {
    float hv = ov;
    dh = gv;
    if (apply_elu) {
        const bool pos = ov > 0.f;
        dh = pos ? gv : gv * (ov + 1.f);
        hv = pos ? ov : (ov > -1.f ? fast_log1p(ov) : 0.f);
    }
    p = fmaf(dh, hv, p);
}

// vector path: D % 4 == 0, 16-byte aligned rows, H <= 4; one warp per row, a lane owns float4 chunks c4 = lane + 32 k of the
// H*D-wide row (a chunk never straddles a head), the per-head dots are reduced with one butterfly per head
__global__ void __launch_bounds__(256)
agg_bwd_pre_vec_kernel(const float* __restrict__ out, const float* __restrict__ dout, long ldo, const float* __restrict__ den,
                       int H, int D4, int apply_elu, float* __restrict__ dhn, long ldd, float* __restrict__ dden, long n) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* o = out + row * ldo;
    const float* g = dout + row * ldo;
    float* d = dhn + row * ldd;
    const int n4 = H * D4;
    float p[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c4 = lane; c4 < n4; c4 += 32) {
        const float4 ov = ldg4_stream(o + c4 * 4), gv = ldg4_stream(g + c4 * 4);
        float4 dh;
        float acc = 0.f;
        bwd_pre_elem(ov.x, gv.x, apply_elu, dh.x, acc);
        bwd_pre_elem(ov.y, gv.y, apply_elu, dh.y, acc);
        bwd_pre_elem(ov.z, gv.z, apply_elu, dh.z, acc);
        bwd_pre_elem(ov.w, gv.w, apply_elu, dh.w, acc);
        st4(d + c4 * 4, dh);
        const int h = c4 / D4;
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] += (h == k) ? acc : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k >= H) break;
        const float t = warp_sum(p[k]);
        if (lane == 0) dden[row * H + k] = -t / __ldg(den + row * H + k);
    }
}

__global__ void __launch_bounds__(256)
agg_bwd_pre_kernel(const float* __restrict__ out, const float* __restrict__ dout, long ldo, const float* __restrict__ den,
                   int H, int D, int apply_elu, float* __restrict__ dhn, long ldd, float* __restrict__ dden, long n) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n) return;
    const float* o = out + row * ldo;
    const float* g = dout + row * ldo;
    float* d = dhn + row * ldd;
    for (int h = 0; h < H; ++h) {
        float p = 0.f;
        for (int c = lane; c < D; c += 32) {
            float dh;
            bwd_pre_elem(__ldg(o + h * D + c), __ldg(g + h * D + c), apply_elu, dh, p);
            d[h * D + c] = dh;
        }
        p = warp_sum(p);
        if (lane == 0) dden[row * H + h] = -p / __ldg(den + row * H + h);
    }
}

// dX[i,f] = rowout[i,f] + sum_h dxc[i, h*4F4 + f] + sum_c dq[i,c] V[f,c],  dq[i] = (dq2_0, dq2_1, dq1_0, dq1_1)
// a warp handles DX_R consecutive rows (F <= 64: lane f and f + 32); every load is issued before the first use
constexpr int DX_R = 2;
__global__ void __launch_bounds__(256)
agg_dx_kernel(const float* __restrict__ rowout, long ldro, const float* __restrict__ dxc, long ldc,
              const float* __restrict__ V, long n, int F, int F4, int H, float* __restrict__ dX, long lddx,
              float* __restrict__ dq) {
    const int lane = threadIdx.x & 31;
    const long i0 = ((long)blockIdx.x * 8 + (threadIdx.x >> 5)) * DX_R;
    float q[DX_R][4], r0[DX_R][2], c0[DX_R][2], c1[DX_R][2];
    float4 v[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) v[k] = (lane + 32 * k < F) ? ldg4(V + (lane + 32 * k) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < DX_R; ++r) {
        const long i = i0 + r < n ? i0 + r : n - 1;
        const float* ro = rowout + i * ldro;
        const float* xc = dxc + i * ldc;
        q[r][0] = __ldg(xc + H * 4 * F4); q[r][1] = H > 1 ? __ldg(xc + H * 4 * F4 + 1) : 0.f;
        q[r][2] = __ldg(ro + 4 * F4); q[r][3] = H > 1 ? __ldg(ro + 4 * F4 + 1) : 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int f = lane + 32 * k;
            r0[r][k] = f < F ? __ldg(ro + f) : 0.f;
            c0[r][k] = f < F ? __ldg(xc + f) : 0.f;
            c1[r][k] = (f < F && H > 1) ? __ldg(xc + 4 * F4 + f) : 0.f;
        }
    }
#pragma unroll
    for (int r = 0; r < DX_R; ++r) {
        const long i = i0 + r;
        if (i >= n) break;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int f = lane + 32 * k;
            const float val = fmaf(q[r][0], v[k].x, fmaf(q[r][1], v[k].y, fmaf(q[r][2], v[k].z, fmaf(q[r][3], v[k].w, r0[r][k] + c0[r][k] + c1[r][k]))));
            if (f < F) dX[i * lddx + f] = val;
        }
        if (lane == 0) st4(dq + i * 4, make_float4(q[r][0], q[r][1], q[r][2], q[r][3]));
    }
}

__device__ __forceinline__ float elu_exact(float x) {
    const float big = exp2f(x * 1.4426950408889634f) - 1.0f;
    const float small = x * (1.0f + x * (0.5f + x * (0.16666667f + x * (0.041666668f + x * 0.0083333338f))));
    const float neg = x > -0.125f ? small : big;
    return x > 0.f ? x : neg;
}

__global__ void __launch_bounds__(256)
elu_inplace_kernel(float* __restrict__ x, long ld, long n, int width) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * width) return;
    const long i = idx / width;
    float* p = x + i * ld + (idx - i * width);
    *p = elu_exact(*p);
}

}  // namespace

int launch_agg_table(const float* X, long ldx, const float* V, float* T, long ldt, long n, int F, int F4, cudaStream_t s) {
    if (n <= 0) return 0;
    const bool vec2 = (F % 2 == 0) && F <= 64 && (ldx % 2 == 0) && (ldt % 2 == 0) && ldt <= 128 &&
                      ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(T)) & 7) == 0 &&
                      (reinterpret_cast<uintptr_t>(V) & 15) == 0;
    if (vec2) agg_table_vec2_kernel<<<(unsigned)((n + 8 * AT_R - 1) / (8 * AT_R)), 256, 0, s>>>(X, ldx, V, T, ldt, n, F, F4);
    else agg_table_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(X, ldx, V, T, ldt, n, F, F4);
    return check_launch("agg_table");
}

int launch_agg_fwd(const AggFwdArgs& a, cudaStream_t s) {
    const bool has2 = a.t2 != nullptr;
    if (a.g.H == 1) return has2 ? launch_agg_fwd_t<1, true>(a, s) : launch_agg_fwd_t<1, false>(a, s);
    return has2 ? launch_agg_fwd_t<2, true>(a, s) : launch_agg_fwd_t<2, false>(a, s);
}

int launch_agg_bwd_rows(const AggBwdArgs& a, cudaStream_t s) {
    const bool has2 = a.t2 != nullptr;
    if (a.g.H == 1) return has2 ? launch_agg_bwd_t<1, true>(a, s) : launch_agg_bwd_t<1, false>(a, s);
    return has2 ? launch_agg_bwd_t<2, true>(a, s) : launch_agg_bwd_t<2, false>(a, s);
}

int launch_agg_bwd_ctx_split(const AggBwdArgs& a, cudaStream_t s) {
    if (a.n_rows <= 0) return 0;
    agg_bwd_ctx_kernel<true><<<(a.n_rows + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA, SPK_CTA_THREADS, 0, s>>>(a);
    return check_launch("agg_bwd_ctx");
}

int launch_agg_bwd_pre(const float* out, const float* dout, long ldo, const float* den, int H, int D, int apply_elu,
                       float* dhn, long ldd, float* dden, long n, cudaStream_t s) {
    if (n <= 0) return 0;
    const bool vec = (D % 4 == 0) && H <= 4 && (ldo % 4 == 0) && (ldd % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout) | reinterpret_cast<uintptr_t>(dhn)) & 15) == 0;
    if (vec)
        agg_bwd_pre_vec_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(out, dout, ldo, den, H, D / 4, apply_elu, dhn, ldd, dden, n);
    else
        agg_bwd_pre_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(out, dout, ldo, den, H, D, apply_elu, dhn, ldd, dden, n);
    return check_launch("agg_bwd_pre");
}

int launch_agg_dx(const float* rowout, long ldro, const float* dxc, long ldc, const float* V, long n, int F, int F4,
                  int H, float* dX, long lddx, float* dq, cudaStream_t s) {
    if (n <= 0) return 0;
    agg_dx_kernel<<<(unsigned)((n + 8 * DX_R - 1) / (8 * DX_R)), 256, 0, s>>>(rowout, ldro, dxc, ldc, V, n, F, F4, H, dX, lddx, dq);
    return check_launch("agg_dx");
}

int launch_elu_inplace(float* x, long ld, long n, int width, cudaStream_t s) {
    if (n <= 0 || width <= 0) return 0;
    const long total = n * width;
    elu_inplace_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(x, ld, n, width);
    return check_launch("elu_inplace");
}

}  // namespace spk
