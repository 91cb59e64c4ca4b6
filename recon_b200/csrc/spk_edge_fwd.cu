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
#include "spk_stream.cuh"
#include <stdlib.h>

namespace spk {

// ELU(x) = x (x > 0) else expm1(x): 2^(x log2e) - 1 away from 0, a degree-5 Taylor polynomial near 0 (|x| < 0.125,
// truncation < 3e-9 relative) where the subtraction would cancel.
__device__ __forceinline__ float elu1(float x) {
    const float big = exp2f(x * 1.4426950408889634f) - 1.0f;
    const float small = x * (1.0f + x * (0.5f + x * (0.16666667f + x * (0.041666668f + x * 0.0083333338f))));
    const float neg = x > -0.125f ? small : big;
    return x > 0.f ? x : neg;
}

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
                                             FwdAcc<NCH, HT>& st, bool& bad, uint32_t p1_smem = 0) {
    const LayerGeom g = a.g;
    float rden[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) {
        if (st.den[h] == 0.f) st.den[h] = 1e-12f;           // layers.py:152
        rden[h] = 1.0f / st.den[h];
    }
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        if (c4 >= g.Dt4) continue;
        const int h = hc[ci];
        const float d = selh<HT>(h, rden);
        const float s = selh<HT>(h, st.sw);
        const float4 p1 = p1_smem ? lds4(p1_smem + c4 * 16) : ldg4(a.P1 + (long)row * a.ld1 + c4 * 4);
        float o[4] = {fmaf(s, p1.x, st.acc[ci].x), fmaf(s, p1.y, st.acc[ci].y),
                      fmaf(s, p1.z, st.acc[ci].z), fmaf(s, p1.w, st.acc[ci].w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[k] = o[k] * d;                                  // layers.py:169 (NaN in num or 0*inf both surface here: 167, 172)
            bad |= (o[k] != o[k]);
            if (a.apply_elu && row < a.elu_rows) o[k] = elu1(o[k]);
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


// row without edges: num = 0, den -> 1e-12 (layers.py:152), out = ELU(0 / 1e-12) = 0, sw = 0
template <int NCH, int HT>
__device__ __forceinline__ void fwd_zero_row(const EdgeFwdArgs& a, int row, int lane) {
    const LayerGeom g = a.g;
    const int hd = g.H * g.D;
    float* dst = a.out + (long)row * a.ldo;
    for (int c = lane; c < hd; c += 32) dst[c] = 0.f;
    if (lane < g.H) { a.den[(long)row * g.H + lane] = 1e-12f; a.sw[(long)row * g.H + lane] = 0.f; }
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

// One CTA per hub row (32 warps so that a hub with thousands of task partials keeps enough loads in flight): the warps add disjoint, interleaved subsets of the task partials (4 loads in
// flight each), then the 8 warp sums are added in warp order -> fixed summation tree, run-to-run identical.
template <int NCH> struct FinCfg { static constexpr int NW = (NCH <= 2) ? 32 : 16; };   // warps of the finalize CTA

template <int NCH, int HT>
__global__ void __launch_bounds__(FinCfg<NCH>::NW * 32)
edge_fwd_hub_finalize_kernel(const EdgeFwdArgs a) {
    __shared__ __align__(16) float red[FinCfg<NCH>::NW][NCH * 128 + 2 * SPK_MAX_HEADS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int hub = blockIdx.x;
    const int row = __ldg(a.hub.hub_seg + hub);
    const int t0 = __ldg(a.hub.hub_task_ptr + hub), t1 = __ldg(a.hub.hub_task_ptr + hub + 1);
    const int len = a.g.Wd4 * 4 + 2 * SPK_MAX_HEADS;         // floats used per partial row
    cta_sum_partials<NCH * 4 + 1, FinCfg<NCH>::NW>(a.hub.partial, a.hub.ldpart, t0, t1, len, &red[0][0], NCH * 128 + 2 * SPK_MAX_HEADS);
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


// ------------------------------------------------------------------------------------------------
// Streaming variant (default): see spk_stream.cuh. Rows mode: a warp owns 32 consecutive rows; tasks mode: one
// hub chunk. Ring slot = [ P2~[j] | P3~[k] (| P3~[k2]) ]; the ring is cut into NG groups of G slots with one
// mbarrier per group, lanes 0..G-1 issue the copies of a whole group at once. All per-edge scalar work (entry
// lookup, indices, score q1+q2+q3, LeakyReLU, exp, dropout multiplier) is done lane-parallel for 32 edges at a
// time in the index batch, so the per-edge loop is: 4 LDS.128, 8 FADD, 8 FFMA and the broadcasts of ee / w.
// P1~ rows are prefetched two rows ahead.
// ------------------------------------------------------------------------------------------------
constexpr int STREAM_WARPS = 8;
constexpr int STREAM_G = 2;                                       // edges per group (divides 32)
template <int NCH, bool HAS2> struct StreamCfg { static constexpr int NG = (NCH <= 2) ? (HAS2 ? 2 : 3) : 2; };

template <int HT>
struct IdxBatch { int col, t1, t2; float m[HT]; };

// exp(x) = 2^(x*log2e): ex2.approx (2 ulp) + one rounding of the product; |rel err| <~ 1e-7 * (1 + |x|)
__device__ __forceinline__ float fast_exp(float x) { return exp2f(x * 1.4426950408889634f); }

// NGV / MINB: tuning variant (ring depth in groups, CTAs per SM); 0 = the StreamCfg default with 2 CTAs/SM
template <int NCH, int HT, bool HAS2, bool TASKS, int NGV = 0, int MINB = 2>
__global__ void __launch_bounds__(STREAM_WARPS * 32, MINB)
edge_fwd_stream_kernel(const EdgeFwdArgs a) {
    constexpr int G = STREAM_G, NG = NGV ? NGV : StreamCfg<NCH, HAS2>::NG, S = G * NG;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const LayerGeom g = a.g;
    const int H = g.H;
    const uint32_t row_bytes = (uint32_t)g.Wd4 * 16u;
    const uint32_t slot_bytes = row_bytes * (HAS2 ? 3u : 2u);
    const uint32_t warp_bytes = S * slot_bytes + 2u * row_bytes + 128u;
    const uint32_t wbase = smem_addr(smem_raw) + (uint32_t)wid * warp_bytes;
    const uint32_t p1buf = wbase + S * slot_bytes;
    const uint32_t bars = p1buf + 2u * row_bytes;                 // NG group barriers, then 2 P1 barriers
    if (lane == 0) {
        for (int i = 0; i < NG + 2; ++i) sbar_init(bars + 8u * i, 1);
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
        row0 = ((long)blockIdx.x * STREAM_WARPS + wid) * 32;
        const long r = row0 + lane;
        if (r < a.n_rows) {
            const int b = __ldg(a.segptr + r), e = __ldg(a.segptr + r + 1);
            st.beg = b;
            if (e - b > a.hub.hub_thresh) is_hub = true; else st.deg = e - b;
        }
    } else {
        const int task = blockIdx.x * STREAM_WARPS + wid;
        if (task >= a.hub.n_tasks) return;
        task_row = __ldg(a.hub.task_seg + task);
        if (lane == 0) { st.beg = __ldg(a.hub.task_beg + task); st.deg = __ldg(a.hub.task_end + task) - st.beg; }
    }
    st.pre = warp_excl_scan(st.deg, lane, st.total);
    const int T = st.total;
    const bool has_mask = a.mask != nullptr;
    const long qoff = (long)g.Dt4 * 4;

    bool bad = false;
    float4 q1v = make_float4(0.f, 0.f, 0.f, 0.f);                  // score scalars q1 of this lane's row
    if (!TASKS) { if (st.deg > 0) q1v = ldg4(a.P1 + (row0 + lane) * a.ld1 + qoff); }
    else q1v = ldg4(a.P1 + (long)task_row * a.ld1 + qoff);
    // ---- index batches (32 stream positions, lane-parallel), loaded one batch ahead; nothing here depends on a
    //      loaded value, so the loads stay in flight until the batch becomes current ----
    IdxBatch<HT> cur, nxt;
    auto load_batch = [&](int nb0, IdxBatch<HT>& b) {
        b.col = 0; b.t1 = 0; b.t2 = -1;
#pragma unroll
        for (int h = 0; h < HT; ++h) b.m[h] = 1.f;
        if (nb0 >= T) return;                                      // warp-uniform
        const int m = nb0 + lane;
        int seg;
        const int e = st.entry_of(m < T ? m : T - 1, seg);
        if (m < T) {
            b.col = __ldg(a.col + e);
            b.t1 = __ldg(a.t1 + e);
            if (HAS2) b.t2 = __ldg(a.t2 + e);
            if (has_mask) {
#pragma unroll
                for (int h = 0; h < HT; ++h)
                    if (h < H) b.m[h] = __ldg(a.mask + (long)h * a.mask_stride + e);
            }
        }
    };
    load_batch(0, cur);
    load_batch(32, nxt);
    int nb = 0;
    // issue the copies of the group starting at stream position n0 (lanes 0..G-1, one edge each)
    auto issue_group = [&](int n0) {
        const int n = n0 + lane;
        const int d = n - nb, src = d & 31;
        const int jc = __shfl_sync(0xffffffffu, cur.col, src), jn = __shfl_sync(0xffffffffu, nxt.col, src);
        const int kc = __shfl_sync(0xffffffffu, cur.t1, src), kn = __shfl_sync(0xffffffffu, nxt.t1, src);
        int k2 = -1;
        if (HAS2) {
            const int k2c = __shfl_sync(0xffffffffu, cur.t2, src), k2n = __shfl_sync(0xffffffffu, nxt.t2, src);
            k2 = d < 32 ? k2c : k2n;
        }
        const bool mine = lane < G && n < T;
        const uint32_t bar = bars + 8u * ((n0 / G) % NG);
        unsigned n3 = 0;
        if (HAS2) n3 = __popc(__ballot_sync(0xffffffffu, mine && k2 >= 0));
        if (lane == 0) {
            const int cnt = T - n0 < G ? T - n0 : G;
            sbar_expect(bar, (uint32_t)cnt * 2u * row_bytes + n3 * row_bytes);
        }
        if (mine) {
            const int j = d < 32 ? jc : jn, k1 = d < 32 ? kc : kn;
            const uint32_t slot = wbase + (uint32_t)(n % S) * slot_bytes;
            bulk_g2s(slot, a.P2 + (long)j * a.ld2, row_bytes, bar);
            bulk_g2s(slot + row_bytes, a.P3 + (long)k1 * a.ld3, row_bytes, bar);
            if (HAS2 && k2 >= 0) bulk_g2s(slot + 2u * row_bytes, a.P3 + (long)k2 * a.ld3, row_bytes, bar);
        }
    };
    for (int gi = 0; gi < NG && gi * G < T; ++gi) issue_group(gi * G);

    // ---- rows mode: empty rows are written up front; P1~ rows prefetched two active rows ahead ----
    unsigned act = TASKS ? 1u : __ballot_sync(0xffffffffu, st.deg > 0);
    if (!TASKS) {
        unsigned empt = __ballot_sync(0xffffffffu, st.deg == 0 && !is_hub && row0 + lane < a.n_rows);
        while (empt) { const int r = __ffs(empt) - 1; empt &= empt - 1; fwd_zero_row<NCH, HT>(a, (int)(row0 + r), lane); }
    }
    unsigned rem = TASKS ? 0u : act;
    int pk = 0;
    auto issue_p1 = [&]() {
        if (!rem) return;
        const int r = __ffs(rem) - 1;
        rem &= rem - 1;
        if (lane == 0) {
            const uint32_t bar = bars + 8u * (NG + (pk & 1));
            sbar_expect(bar, row_bytes);
            bulk_g2s(p1buf + (uint32_t)(pk & 1) * row_bytes, a.P1 + (row0 + r) * a.ld1, row_bytes, bar);
        }
        ++pk;
    };
    issue_p1();
    issue_p1();

    int hc[NCH];
    bool cvalid[NCH];
    uint32_t coff[NCH];
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) { cvalid[ci] = lane + 32 * ci < g.Wd4; coff[ci] = (uint32_t)(lane + 32 * ci) * 16u; }
    FwdAcc<NCH, HT> acc;
    fwd_init<NCH, HT>(g, lane, hc, acc);
    int ak = 0;
    int r = TASKS ? 0 : __ffs(act) - 1;                           // current segment (lane index)
    act &= act - 1;
    int row_end = T ? __shfl_sync(0xffffffffu, st.pre + st.deg, r < 0 ? 0 : r) : 0;
    float q1[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) q1[h] = TASKS ? f4get(q1v, h) : __shfl_sync(0xffffffffu, f4get(q1v, h), r < 0 ? 0 : r);
    const uint32_t qo = (uint32_t)g.Dt4 * 16u;                    // byte offset of the q scalars inside a staged row

    for (int n0 = 0; n0 < T; n0 += G) {
        if (n0 - nb == 32) { cur = nxt; nb += 32; load_batch(nb + 32, nxt); }
        const int grp = (n0 / G) % NG;
        sbar_wait(bars + 8u * grp, (uint32_t)(n0 / S) & 1u);
#pragma unroll
        for (int u = 0; u < G; ++u) {
            const int n = n0 + u;
            if (n >= T) break;
            const uint32_t slot = wbase + (uint32_t)(grp * G + u) * slot_bytes;
            const int src = n - nb;
            int k2 = -1;
            if (HAS2) k2 = __shfl_sync(0xffffffffu, cur.t2, src);
            float4 v[NCH];
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (cvalid[ci]) {
                    x = f4add(lds4(slot + coff[ci]), lds4(slot + row_bytes + coff[ci]));
                    if (HAS2 && k2 >= 0) x = f4add(x, lds4(slot + 2u * row_bytes + coff[ci]));
                }
                v[ci] = x;
            }
            // score from the staged rows (broadcast LDS): s = q1[i] + q2[j] + q3[k] (+ q3[k2])
            float4 qq = f4add(lds4(slot + qo), lds4(slot + row_bytes + qo));
            if (HAS2 && k2 >= 0) qq = f4add(qq, lds4(slot + 2u * row_bytes + qo));
            float w[HT];
#pragma unroll
            for (int h = 0; h < HT; ++h) {
                w[h] = 0.f;
                if (h < H) {
                    const float sc = q1[h] + f4get(qq, h);
                    const float ee = fast_exp(-(sc > 0.f ? sc : a.alpha * sc));     // layers.py:143-146
                    bad |= (ee != ee);
                    acc.den[h] += ee;
                    w[h] = has_mask ? ee * __shfl_sync(0xffffffffu, cur.m[h], src) : ee;   // layers.py:158
                    acc.sw[h] += w[h];
                }
            }
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) f4fma(acc.acc[ci], selh<HT>(hc[ci], w), v[ci]);
            if (!TASKS && n + 1 == row_end) {                      // row complete -> finalize, move to the next active row
                const uint32_t pb = p1buf + (uint32_t)(ak & 1) * row_bytes;
                sbar_wait(bars + 8u * (NG + (ak & 1)), (uint32_t)(ak >> 1) & 1u);
                fwd_finalize<NCH, HT>(a, (int)(row0 + r), lane, hc, acc, bad, pb);
                __syncwarp();
                issue_p1();                                        // pk == ak + 2: refills the buffer just read
                ++ak;
                fwd_init<NCH, HT>(g, lane, hc, acc);
                if (act) {
                    r = __ffs(act) - 1;
                    act &= act - 1;
                    row_end = __shfl_sync(0xffffffffu, st.pre + st.deg, r);
#pragma unroll
                    for (int h = 0; h < HT; ++h) q1[h] = __shfl_sync(0xffffffffu, f4get(q1v, h), r);
                }
            }
        }
        __syncwarp();                                              // whole group consumed -> refill its slots
        if (n0 + S < T) issue_group(n0 + S);
    }
    if (TASKS) {
        const int task = blockIdx.x * STREAM_WARPS + wid;
        float* part = a.hub.partial + (long)task * a.hub.ldpart;
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
            const int c4 = lane + 32 * ci;
            if (c4 < g.Wd4) *reinterpret_cast<float4*>(part + c4 * 4) = acc.acc[ci];
        }
        if (lane < SPK_MAX_HEADS) {
            part[g.Wd4 * 4 + lane] = lane < HT ? selh<HT>(lane, acc.den) : 0.f;
            part[g.Wd4 * 4 + SPK_MAX_HEADS + lane] = lane < HT ? selh<HT>(lane, acc.sw) : 0.f;
        }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(a.nanflag, 1);
}

template <int NCH, bool HAS2, int NGV = 0>
static size_t stream_smem_bytes(const LayerGeom& g) {
    const size_t row_bytes = (size_t)g.Wd4 * 16;
    return STREAM_WARPS * ((NGV ? NGV : StreamCfg<NCH, HAS2>::NG) * STREAM_G * row_bytes * (HAS2 ? 3 : 2) + 2 * row_bytes + 128);
}

// SPK_FWD_VARIANT=1 (default, measured 4.23 vs 4.64 ms): ring of 2 groups (4 slots per warp) and 3 CTAs/SM; 0: 3 groups, 2 CTAs/SM
static int fwd_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SPK_FWD_VARIANT"); v = e ? atoi(e) : 1; if (v < 0 || v > 1) v = 1; }
    return v;
}

template <int NCH, int HT, bool HAS2, int NGV, int MINB>
static int launch_fwd_stream_v(const EdgeFwdArgs& a, cudaStream_t s) {
    const size_t smem = stream_smem_bytes<NCH, HAS2, NGV>(a.g);
    static SmemLimit lim_rows, lim_tasks;
    if (a.n_rows > 0) {
        lim_rows.ensure(edge_fwd_stream_kernel<NCH, HT, HAS2, false, NGV, MINB>, smem);
        const unsigned grid = (unsigned)((a.n_rows + 32L * STREAM_WARPS - 1) / (32L * STREAM_WARPS));
        edge_fwd_stream_kernel<NCH, HT, HAS2, false, NGV, MINB><<<grid, STREAM_WARPS * 32, smem, s>>>(a);
        if (int rc = check_launch("edge_fwd_stream_rows")) return rc;
    }
    if (a.hub.n_tasks > 0) {
        lim_tasks.ensure(edge_fwd_stream_kernel<NCH, HT, HAS2, true, NGV, MINB>, smem);
        const unsigned grid = (a.hub.n_tasks + STREAM_WARPS - 1) / STREAM_WARPS;
        edge_fwd_stream_kernel<NCH, HT, HAS2, true, NGV, MINB><<<grid, STREAM_WARPS * 32, smem, s>>>(a);
        if (int rc = check_launch("edge_fwd_stream_tasks")) return rc;
        edge_fwd_hub_finalize_kernel<NCH, HT><<<a.hub.n_hubs, FinCfg<NCH>::NW * 32, 0, s>>>(a);
        if (int rc = check_launch("edge_fwd_hub_finalize")) return rc;
    }
    return 0;
}

static bool use_stream() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SPK_EDGE_STREAM"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

template <int NCH, int HT, bool HAS2>
static int launch_fwd_t(const EdgeFwdArgs& a, cudaStream_t s) {
    if (use_stream() && NCH <= 2 && !HAS2 && fwd_variant() == 1) return launch_fwd_stream_v<NCH, HT, HAS2, 2, 3>(a, s);
    if (use_stream()) {
        const size_t smem = stream_smem_bytes<NCH, HAS2>(a.g);
        static SmemLimit lim_rows, lim_tasks;
        lim_rows.ensure(edge_fwd_stream_kernel<NCH, HT, HAS2, false>, smem);
        if (a.n_rows > 0) {
            const unsigned grid = (unsigned)((a.n_rows + 32L * STREAM_WARPS - 1) / (32L * STREAM_WARPS));
            edge_fwd_stream_kernel<NCH, HT, HAS2, false><<<grid, STREAM_WARPS * 32, smem, s>>>(a);
            if (int rc = check_launch("edge_fwd_stream_rows")) return rc;
        }
        if (a.hub.n_tasks > 0) {
            lim_tasks.ensure(edge_fwd_stream_kernel<NCH, HT, HAS2, true>, smem);
            const unsigned grid = (a.hub.n_tasks + STREAM_WARPS - 1) / STREAM_WARPS;
            edge_fwd_stream_kernel<NCH, HT, HAS2, true><<<grid, STREAM_WARPS * 32, smem, s>>>(a);
            if (int rc = check_launch("edge_fwd_stream_tasks")) return rc;
            edge_fwd_hub_finalize_kernel<NCH, HT><<<a.hub.n_hubs, FinCfg<NCH>::NW * 32, 0, s>>>(a);
            if (int rc = check_launch("edge_fwd_hub_finalize")) return rc;
        }
        return 0;
    }
    if (a.n_rows > 0) {
        const unsigned grid = (a.n_rows + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
        edge_fwd_rows_kernel<NCH, HT, HAS2><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("edge_fwd_rows")) return rc;
    }
    if (a.hub.n_tasks > 0) {
        const unsigned grid = (a.hub.n_tasks + SPK_WARPS_PER_CTA - 1) / SPK_WARPS_PER_CTA;
        edge_fwd_tasks_kernel<NCH, HT, HAS2><<<grid, SPK_CTA_THREADS, 0, s>>>(a);
        if (int rc = check_launch("edge_fwd_tasks")) return rc;
        edge_fwd_hub_finalize_kernel<NCH, HT><<<a.hub.n_hubs, FinCfg<NCH>::NW * 32, 0, s>>>(a);
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
