// K3 streaming variant: backward over aggregation rows with bulk-async (TMA 1-D) row staging (spk_stream.cuh).
//
// A warp owns 32 consecutive rows (or one hub chunk). Gathered rows P2~[j] | P3~[k] (| P3~[k2]) are staged in a
// per-warp shared-memory ring with one mbarrier per slot; the row context (saved output, upstream gradient, P1~
// row) of the next row is prefetched while the current row's edges are processed. Edges of a row are consumed in
// pairs: the 2*H dot products t = dnum . m_e of a pair are reduced with one transposed butterfly (6 shuffles for
// H = 2 instead of 20), after which each lane group owns one (edge, head) scalar and computes
// ds = -(msk*t + dden) * ee * LeakyReLU'(s) for it, writes the (w, ds) record and keeps lane-local row sums.
#include <stdlib.h>
#include "spk_edge_bwd.cuh"
#include "spk_stream.cuh"

namespace spk {
namespace {

constexpr int BS_WARPS = 8;
template <int NCH, bool HAS2> struct BwdCfg { static constexpr int S = (NCH <= 2) ? (HAS2 ? 4 : 6) : (HAS2 ? 2 : 4); };

__device__ __forceinline__ float fast_exp_b(float x) { return exp2f(x * 1.4426950408889634f); }

template <int HT>
struct BBatch { int ent, col, t1, t2; float m[HT]; };

// Sum NV values (power of two) over the 32 lanes; lane L returns the total of value index L / (32 / NV).
template <int NV>
struct MultiReduce {
    static __device__ __forceinline__ float run(const float (&x)[NV], int lane, int off) {
        float y[NV / 2];
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < NV / 2; ++k) {
            const float keep = up ? x[k + NV / 2] : x[k];
            const float send = up ? x[k] : x[k + NV / 2];
            y[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        return MultiReduce<NV / 2>::run(y, lane, off >> 1);
    }
};
template <>
struct MultiReduce<1> {
    static __device__ __forceinline__ float run(const float (&x)[1], int lane, int off) {
        float v = x[0];
        for (int o = off; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
};

// Row prologue from the staged context: ctx = [ out row | dout row | P1~ row ] in shared memory.
template <int NCH, int HT>
__device__ __forceinline__ void prologue_smem(const EdgeBwdRowsArgs& a, long row, int lane, uint32_t ctx, uint32_t act_bytes,
                                              RowCtx<NCH, HT>& rc) {
    const LayerGeom g = a.g;
    float pdh[HT], pc1[HT], den[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) {
        pdh[h] = 0.f; pc1[h] = 0.f;
        den[h] = h < g.H ? __ldg(a.den + row * g.H + h) : 1.f;
    }
    const uint32_t p1 = ctx + 2u * act_bytes;
    const float4 q1v = lds4(p1 + (uint32_t)g.Dt4 * 16u);
#pragma unroll
    for (int h = 0; h < HT; ++h) rc.q1[h] = f4get(q1v, h);
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        rc.hc[ci] = 0;
        rc.dnum[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 >= g.Dt4) continue;
        const int h = HT > 1 ? c4 / g.Dp4 : 0;
        rc.hc[ci] = h;
        // stream path requires D % 4 == 0, so a padded chunk is a contiguous float4 of the unpadded row
        const uint32_t aoff = (uint32_t)(h * g.D + (c4 - h * g.Dp4) * 4) * 4u;
        const float4 o4 = lds4(ctx + aoff), g4 = lds4(ctx + act_bytes + aoff);
        const float o[4] = {o4.x, o4.y, o4.z, o4.w};
        const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
        float dh[4], hv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (a.apply_elu) {
                const bool pos = o[k] > 0.f;
                dh[k] = pos ? gg[k] : gg[k] * (o[k] + 1.f);
                hv[k] = pos ? o[k] : (o[k] > -1.f ? fast_log1p(o[k]) : 0.f);
            } else {
                dh[k] = gg[k];
                hv[k] = o[k];
            }
        }
        const float rd = 1.0f / selh<HT>(h, den);
        const float dot_h = fmaf(dh[0], hv[0], fmaf(dh[1], hv[1], fmaf(dh[2], hv[2], dh[3] * hv[3])));
        rc.dnum[ci] = make_float4(dh[0] * rd, dh[1] * rd, dh[2] * rd, dh[3] * rd);
        const float dot_c = f4dot(rc.dnum[ci], lds4(p1 + (uint32_t)c4 * 16u));
#pragma unroll
        for (int hh = 0; hh < HT; ++hh) {
            pdh[hh] += (h == hh) ? dot_h : 0.f;
            pc1[hh] += (h == hh) ? dot_c : 0.f;
        }
    }
#pragma unroll
    for (int h = 0; h < HT; ++h) {
        rc.dden[h] = 0.f; rc.c1[h] = 0.f;
        if (h < g.H) {
            rc.dden[h] = -warp_sum(pdh[h]) / den[h];
            rc.c1[h] = warp_sum(pc1[h]);
        }
    }
}

template <int NCH, int HT, bool HAS2, bool TASKS>
__global__ void __launch_bounds__(BS_WARPS * 32, 2)
edge_bwd_rows_stream_kernel(const EdgeBwdRowsArgs a) {
    constexpr int S = BwdCfg<NCH, HAS2>::S;
    constexpr int NV = 2 * HT;                                     // HT is 1, 2 or 4
    constexpr int LPG = 32 / NV;                                   // lanes per (edge, head) group after the reduction
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const LayerGeom g = a.g;
    const int H = g.H;
    const uint32_t row_bytes = (uint32_t)g.Wd4 * 16u;
    const uint32_t slot_bytes = row_bytes * (HAS2 ? 3u : 2u);
    const uint32_t act_bytes = ((uint32_t)(H * g.D) * 4u + 15u) & ~15u;
    const uint32_t ctx_bytes = 2u * act_bytes + row_bytes;
    const uint32_t warp_bytes = S * slot_bytes + ctx_bytes + 128u;
    const uint32_t wbase = smem_addr(smem_raw) + (uint32_t)wid * warp_bytes;
    const uint32_t ctx = wbase + S * slot_bytes;
    const uint32_t bars = ctx + ctx_bytes;                         // S slot barriers, then the context barrier
    if (lane == 0) {
        for (int i = 0; i < S + 1; ++i) sbar_init(bars + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    SegTable st;
    st.beg = 0; st.deg = 0;
    bool is_hub = false;
    long row0 = 0;
    int task_row = 0;
    if (!TASKS) {
        row0 = ((long)blockIdx.x * BS_WARPS + wid) * 32;
        const long r = row0 + lane;
        if (r < a.n_rows) {
            const int b = __ldg(a.segptr + r), e = __ldg(a.segptr + r + 1);
            st.beg = b;
            if (e - b > a.hub.hub_thresh) is_hub = true; else st.deg = e - b;
        }
    } else {
        const int task = blockIdx.x * BS_WARPS + wid;
        if (task >= a.hub.n_tasks) return;
        task_row = __ldg(a.hub.task_seg + task);
        if (lane == 0) { st.beg = __ldg(a.hub.task_beg + task); st.deg = __ldg(a.hub.task_end + task) - st.beg; }
    }
    st.pre = warp_excl_scan(st.deg, lane, st.total);
    const int T = st.total;
    const bool has_mask = a.mask != nullptr;
    const uint32_t qo = (uint32_t)g.Dt4 * 16u;

    BBatch<HT> cur, nxt;
    auto load_batch = [&](int nb0, BBatch<HT>& b) {
        b.ent = 0; b.col = 0; b.t1 = 0; b.t2 = -1;
#pragma unroll
        for (int h = 0; h < HT; ++h) b.m[h] = 1.f;
        if (nb0 >= T) return;
        const int m = nb0 + lane;
        int seg;
        const int e = st.entry_of(m < T ? m : T - 1, seg);
        if (m < T) {
            b.ent = e;
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
    // value of a batch field at stream position p (nb <= p < nb + 64); per-lane p allowed, all lanes must call
    auto at = [&](int fc, int fn, int p) {
        const int d = p - nb;
        const int vc = __shfl_sync(0xffffffffu, fc, d & 31), vn = __shfl_sync(0xffffffffu, fn, d & 31);
        return d < 32 ? vc : vn;
    };
    auto atf = [&](float fc, float fn, int p) {
        const int d = p - nb;
        const float vc = __shfl_sync(0xffffffffu, fc, d & 31), vn = __shfl_sync(0xffffffffu, fn, d & 31);
        return d < 32 ? vc : vn;
    };
    // lanes 0 / 1 issue the copies of stream positions p0 / p0+1 (second only if `two`), each arming its own slot barrier
    auto issue2 = [&](int p0, bool two) {
        const int p = p0 + (lane & 1);
        const int pc = p < T ? p : T - 1;
        const int j = at(cur.col, nxt.col, pc), k1 = at(cur.t1, nxt.t1, pc);
        int k2 = -1;
        if (HAS2) k2 = at(cur.t2, nxt.t2, pc);
        if (lane < 2 && p < T && (lane == 0 || two)) {
            const uint32_t slot = wbase + (uint32_t)(p % S) * slot_bytes, bar = bars + 8u * (p % S);
            sbar_expect(bar, (HAS2 && k2 >= 0) ? 3u * row_bytes : 2u * row_bytes);
            bulk_g2s(slot, a.P2 + (long)j * a.ld2, row_bytes, bar);
            bulk_g2s(slot + row_bytes, a.P3 + (long)k1 * a.ld3, row_bytes, bar);
            if (HAS2 && k2 >= 0) bulk_g2s(slot + 2u * row_bytes, a.P3 + (long)k2 * a.ld3, row_bytes, bar);
        }
    };
    if (T > 0)
        for (int p = 0; p < S && p < T; p += 2) issue2(p, p + 1 < S);

    unsigned act = TASKS ? 1u : __ballot_sync(0xffffffffu, st.deg > 0);
    if (!TASKS) {                                                  // rows without edges: dP1~ = 0 (G is never gathered for them)
        unsigned empt = __ballot_sync(0xffffffffu, st.deg == 0 && !is_hub && row0 + lane < a.n_rows);
        while (empt) {
            const int r = __ffs(empt) - 1;
            empt &= empt - 1;
            for (int c4 = lane; c4 < g.Wd4; c4 += 32)
                *reinterpret_cast<float4*>(a.dP1 + (row0 + r) * a.ldd1 + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    auto issue_ctx = [&](int r) {                                  // stage out / dout / P1~ rows of row r
        if (lane == 0) {
            const uint32_t bar = bars + 8u * S;
            sbar_expect(bar, ctx_bytes);
            bulk_g2s(ctx, a.out + (row0 + r) * a.ldo, act_bytes, bar);
            bulk_g2s(ctx + act_bytes, a.dout + (row0 + r) * a.ldo, act_bytes, bar);
            bulk_g2s(ctx + 2u * act_bytes, a.P1 + (row0 + r) * a.ld1, row_bytes, bar);
        }
    };
    if (!TASKS && act) issue_ctx(__ffs(act) - 1);

    // lane role after the transposed reduction
    const int my_idx = lane / LPG, my_u = my_idx / HT, my_h = my_idx % HT;
    const bool role_ok = my_h < H;

    RowCtx<NCH, HT> rc;
    int n = 0, ak = 0;
    float tot_u[HT], tot_sw[HT];
    while (act) {
        const int r = TASKS ? 0 : __ffs(act) - 1;
        act &= act - 1;
        const long row = TASKS ? task_row : row0 + r;
        const int d = __shfl_sync(0xffffffffu, st.deg, r);
        if (!TASKS) {
            sbar_wait(bars + 8u * S, (uint32_t)ak & 1u);
            prologue_smem<NCH, HT>(a, row, lane, ctx, act_bytes, rc);
            __syncwarp();
            if (act) issue_ctx(__ffs(act) - 1);                    // next active row lands while this row's edges run
            ++ak;
        } else {
            bwd_row_prologue<NCH, HT>(a, (int)row, lane, rc);
        }
        const float my_q1 = selh<HT>(my_h, rc.q1), my_dden = selh<HT>(my_h, rc.dden), my_c1 = selh<HT>(my_h, rc.c1);
        float my_usum = 0.f, my_sw = 0.f;
        for (int k = 0; k < d; k += 2, n += 2) {
            while (n - nb >= 32) { cur = nxt; nb += 32; load_batch(nb + 32, nxt); }
            const bool two = k + 1 < d;
            const uint32_t slot0 = wbase + (uint32_t)(n % S) * slot_bytes;
            const uint32_t slot1 = wbase + (uint32_t)((n + 1) % S) * slot_bytes;
            int k20 = -1, k21 = -1;
            if (HAS2) { k20 = at(cur.t2, nxt.t2, n); k21 = at(cur.t2, nxt.t2, two ? n + 1 : n); }
            sbar_wait(bars + 8u * (n % S), (uint32_t)(n / S) & 1u);
            if (two) sbar_wait(bars + 8u * ((n + 1) % S), (uint32_t)((n + 1) / S) & 1u);
            float pd[NV];
#pragma unroll
            for (int i = 0; i < NV; ++i) pd[i] = 0.f;
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c4 = lane + 32 * ci;
                if (c4 < g.Dt4) {                                  // dnum is zero beyond the projection chunks
                    const uint32_t co = (uint32_t)c4 * 16u;
                    float4 x0 = f4add(lds4(slot0 + co), lds4(slot0 + row_bytes + co));
                    if (HAS2 && k20 >= 0) x0 = f4add(x0, lds4(slot0 + 2u * row_bytes + co));
                    const float d0 = f4dot(rc.dnum[ci], x0);
                    float d1 = 0.f;
                    if (two) {
                        float4 x1 = f4add(lds4(slot1 + co), lds4(slot1 + row_bytes + co));
                        if (HAS2 && k21 >= 0) x1 = f4add(x1, lds4(slot1 + 2u * row_bytes + co));
                        d1 = f4dot(rc.dnum[ci], x1);
                    }
#pragma unroll
                    for (int h = 0; h < HT; ++h) {
                        pd[h] += (HT == 1 || rc.hc[ci] == h) ? d0 : 0.f;
                        pd[HT + h] += (HT == 1 || rc.hc[ci] == h) ? d1 : 0.f;
                    }
                }
            }
            // score scalars of my (edge, head) role, read before the slots are refilled
            const int p = n + my_u;
            const bool valid = role_ok && (my_u == 0 || two);
            const uint32_t slotp = my_u == 0 ? slot0 : slot1;
            const int k2p = my_u == 0 ? k20 : k21;
            float sc = my_q1;
            if (valid) {
                float q23;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q23) : "r"(slotp + qo + 4u * my_h));
                sc += q23;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q23) : "r"(slotp + row_bytes + qo + 4u * my_h));
                sc += q23;
                if (HAS2 && k2p >= 0) {
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(q23) : "r"(slotp + 2u * row_bytes + qo + 4u * my_h));
                    sc += q23;
                }
            }
            __syncwarp();                                          // both slots fully read -> refill them
            if (n + S < T) issue2(n + S, two);
            const float tot = MultiReduce<NV>::run(pd, lane, 16);
            float m = 1.f;
            if (has_mask) {                                        // the source lane holds all heads of its edge
                float mh[HT];
#pragma unroll
                for (int h = 0; h < HT; ++h) mh[h] = atf(cur.m[h], nxt.m[h], valid ? p : n);
                m = selh<HT>(my_h, mh);
            }
            const int ent = at(cur.ent, nxt.ent, (my_u == 0 || two) ? p : n);
            if (valid) {
                const float t = my_c1 + tot;
                const float slope = sc > 0.f ? 1.f : a.alpha;
                const float ee = fast_exp_b(-(sc * slope));
                const float w = ee * m;
                const float ds = -(m * t + my_dden) * ee * slope;
                my_usum += ds;
                my_sw += w;
                if ((lane & (LPG - 1)) == 0)
                    *reinterpret_cast<float2*>(a.rec + (long)ent * (2 * H) + 2 * my_h) = make_float2(w, ds);
            }
        }
        if ((d & 1) && d > 0) n -= 1;                              // the last pair of an odd row consumed one position
#pragma unroll
        for (int h = 0; h < HT; ++h) {
            tot_u[h] = __shfl_sync(0xffffffffu, my_usum, h * LPG) + __shfl_sync(0xffffffffu, my_usum, (HT + h) * LPG);
            tot_sw[h] = __shfl_sync(0xffffffffu, my_sw, h * LPG) + __shfl_sync(0xffffffffu, my_sw, (HT + h) * LPG);
        }
        if (!TASKS) bwd_row_store<NCH, HT>(a, (int)row, lane, rc, tot_u, tot_sw, true, true);
    }
    if (TASKS) {
        const int task = blockIdx.x * BS_WARPS + wid;
        if (lane < SPK_MAX_HEADS) {
            float* part = a.hub.partial + (long)task * a.hub.ldpart;
            part[lane] = lane < HT ? selh<HT>(lane, tot_u) : 0.f;
            part[SPK_MAX_HEADS + lane] = lane < HT ? selh<HT>(lane, tot_sw) : 0.f;
        }
    }
}

template <int NCH, bool HAS2>
size_t bwd_stream_smem(const LayerGeom& g) {
    const size_t row_bytes = (size_t)g.Wd4 * 16;
    const size_t act_bytes = ((size_t)g.H * g.D * 4 + 15) & ~(size_t)15;
    return BS_WARPS * (BwdCfg<NCH, HAS2>::S * row_bytes * (HAS2 ? 3 : 2) + 2 * act_bytes + row_bytes + 128);
}

// hub finalize for the stream path lives in spk_edge_bwd.cu (edge_bwd_rows_hub_finalize_kernel); declared here
template <int NCH, int HT, bool HAS2>
int launch_t(const EdgeBwdRowsArgs& a, cudaStream_t s, int (*finalize)(const EdgeBwdRowsArgs&, cudaStream_t)) {
    const size_t smem = bwd_stream_smem<NCH, HAS2>(a.g);
    static SmemLimit lim_rows, lim_tasks;
    if (a.n_rows > 0) {
        lim_rows.ensure(edge_bwd_rows_stream_kernel<NCH, HT, HAS2, false>, smem);
        const unsigned grid = (unsigned)((a.n_rows + 32L * BS_WARPS - 1) / (32L * BS_WARPS));
        edge_bwd_rows_stream_kernel<NCH, HT, HAS2, false><<<grid, BS_WARPS * 32, smem, s>>>(a);
        if (int rc = check_launch("edge_bwd_rows_stream")) return rc;
    }
    if (a.hub.n_tasks > 0) {
        lim_tasks.ensure(edge_bwd_rows_stream_kernel<NCH, HT, HAS2, true>, smem);
        const unsigned grid = (a.hub.n_tasks + BS_WARPS - 1) / BS_WARPS;
        edge_bwd_rows_stream_kernel<NCH, HT, HAS2, true><<<grid, BS_WARPS * 32, smem, s>>>(a);
        if (int rc = check_launch("edge_bwd_rows_stream_tasks")) return rc;
        return finalize(a, s);
    }
    return 0;
}

template <int NCH>
int launch_n(const EdgeBwdRowsArgs& a, cudaStream_t s, int (*fin)(const EdgeBwdRowsArgs&, cudaStream_t)) {
    const bool has2 = a.t2 != nullptr;
    if (a.g.H == 1) return has2 ? launch_t<NCH, 1, true>(a, s, fin) : launch_t<NCH, 1, false>(a, s, fin);
    if (a.g.H == 2) return has2 ? launch_t<NCH, 2, true>(a, s, fin) : launch_t<NCH, 2, false>(a, s, fin);
    return has2 ? launch_t<NCH, 4, true>(a, s, fin) : launch_t<NCH, 4, false>(a, s, fin);
}
}  // namespace

int launch_edge_bwd_rows_hub_finalize(const EdgeBwdRowsArgs& a, cudaStream_t s);   // spk_edge_bwd.cu

int launch_edge_bwd_rows_stream(const EdgeBwdRowsArgs& a, cudaStream_t s) {
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("SPK_EDGE_STREAM"); enabled = (e && e[0] == '0') ? 0 : 1; }
    // bulk copies of the activation rows need 16-byte aligned, 16-byte multiple rows
    const bool ok = enabled && a.out_vec && (a.g.D % 4 == 0) && ((a.g.H * a.g.D) % 4 == 0) && (a.ldo % 4 == 0) && (a.ld1 % 4 == 0);
    if (!ok) return -1;
    switch ((a.g.Wd4 + 31) / 32) {
        case 1: return launch_n<1>(a, s, launch_edge_bwd_rows_hub_finalize);
        case 2: return launch_n<2>(a, s, launch_edge_bwd_rows_hub_finalize);
        default: return -1;                                        // wide rows: register-gather kernels
    }
}

}  // namespace spk
