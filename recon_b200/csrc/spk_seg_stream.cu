// K4 streaming variant: out[seg] = [ sum_e w_e,h * G[src_e] | sum_e ds_e,h | 0 ] over CSC (dP2~) or relation (dP3~)
// segments, with bulk-async (TMA 1-D) staging of the gathered G rows (spk_stream.cuh).
//
// A warp owns 32 consecutive segments (or one hub chunk). The index side is a 3-stage software pipeline over
// batches of 32 entries so that no load is consumed in the rotation that issued it:
//     rotation b: (src, pos) of batch b+3 are requested; the records rec[pos] of batch b+2 are requested
//                 (its indices arrived meanwhile); batch b+1 becomes complete (its records arrived).
// When a batch completes, its ds values are summed per segment with a segmented warp scan and added to the
// accumulator of the segment's owner lane; the w values stay in the batch and are broadcast per entry.
#include <stdlib.h>
#include "spk_edge.cuh"
#include "spk_stream.cuh"

namespace spk {
int launch_seg_gather_hub_finalize(const SegGatherArgs& a, cudaStream_t s);   // spk_edge_bwd.cu

namespace {
constexpr int SS_WARPS = 8;
constexpr int SS_SLOTS = 8;

template <int HT> struct SBatch { int src; float w[HT]; };
template <int HT> struct SPendRec { int src, seg; float2 rec[HT]; };
struct SPendIdx { int src, pos, seg; };

template <int NCH, int HT, bool TASKS>
__global__ void __launch_bounds__(SS_WARPS * 32, 3)
seg_gather_stream_kernel(const SegGatherArgs a) {
    constexpr int S = SS_SLOTS;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const LayerGeom g = a.g;
    const int H = g.H;
    const uint32_t grow_bytes = (uint32_t)a.ldg * 4u;
    const uint32_t warp_bytes = S * grow_bytes + 128u;
    const uint32_t wbase = smem_addr(smem_raw) + (uint32_t)wid * warp_bytes;
    const uint32_t bars = wbase + S * grow_bytes;
    if (lane == 0) {
        for (int i = 0; i < S; ++i) sbar_init(bars + 8u * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    SegTable st;
    st.beg = 0; st.deg = 0;
    bool is_hub = false;
    long seg0 = 0;
    if (!TASKS) {
        seg0 = ((long)blockIdx.x * SS_WARPS + wid) * 32;
        const long sg = seg0 + lane;
        if (sg < a.n_seg) {
            const int b = __ldg(a.segptr + sg), e = __ldg(a.segptr + sg + 1);
            st.beg = b;
            if (e - b > a.hub.hub_thresh) is_hub = true; else st.deg = e - b;
        }
    } else {
        const int task = blockIdx.x * SS_WARPS + wid;            // (opt-in kernel: hub.task_order is not used here)
        if (task >= a.hub.n_tasks) return;
        if (lane == 0) { st.beg = __ldg(a.hub.task_beg + task); st.deg = __ldg(a.hub.task_end + task) - st.beg; }
    }
    st.pre = warp_excl_scan(st.deg, lane, st.total);
    const int T = st.total;

    float vs_acc[HT];                                              // sum of ds over the segment owned by this lane
#pragma unroll
    for (int h = 0; h < HT; ++h) vs_acc[h] = 0.f;

    // ---- 3-stage index pipeline ----
    SBatch<HT> cur, nxt;
    SPendRec<HT> pr;
    SPendIdx pi;
    auto req_idx = [&](int nb0, SPendIdx& o) {                      // stage 1: entry -> (src, pos)
        o.src = 0; o.pos = -1; o.seg = 32;
        if (nb0 >= T) return;
        const int m = nb0 + lane;
        int sg;
        const int e = st.entry_of(m < T ? m : T - 1, sg);
        if (m < T) { o.src = __ldg(a.src + e); o.pos = __ldg(a.pos + e); o.seg = sg; }
    };
    auto req_rec = [&](const SPendIdx& i, SPendRec<HT>& o) {        // stage 2: pos -> records
        o.src = i.src; o.seg = i.seg;
#pragma unroll
        for (int h = 0; h < HT; ++h) {
            o.rec[h] = make_float2(0.f, 0.f);
            if (h < H && i.pos >= 0) o.rec[h] = __ldg(reinterpret_cast<const float2*>(a.rec + (long)i.pos * (2 * H) + 2 * h));
        }
    };
    auto complete = [&](const SPendRec<HT>& i, int nb0, SBatch<HT>& o) {   // stage 3: w stays, ds goes to the owners
        o.src = i.src;
        const int tail = min(st.pre + st.deg, nb0 + 32) - 1 - nb0;          // my segment's last entry inside this batch
        const bool mine = st.deg > 0 && st.pre < nb0 + 32 && st.pre + st.deg > nb0;
#pragma unroll
        for (int h = 0; h < HT; ++h) {
            o.w[h] = i.rec[h].x;
            float v = i.rec[h].y;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {                        // segmented inclusive scan keyed on seg
                const float vu = __shfl_up_sync(0xffffffffu, v, off);
                const int su = __shfl_up_sync(0xffffffffu, i.seg, off);
                if (lane >= off && su == i.seg) v += vu;
            }
            const float run = __shfl_sync(0xffffffffu, v, mine ? tail : 0);
            if (mine) vs_acc[h] += run;
        }
    };
    {
        SPendIdx i0, i1, i2;
        req_idx(0, i0); req_idx(32, i1); req_idx(64, i2); req_idx(96, pi);
        SPendRec<HT> r0, r1;
        req_rec(i0, r0); req_rec(i1, r1); req_rec(i2, pr);
        complete(r0, 0, cur);
        complete(r1, 32, nxt);
    }
    int nb = 0;
    auto at = [&](int fc, int fn, int p) {
        const int d = p - nb;
        const int vc = __shfl_sync(0xffffffffu, fc, d & 31), vn = __shfl_sync(0xffffffffu, fn, d & 31);
        return d < 32 ? vc : vn;
    };
    auto issue4 = [&](int p0) {                                    // lanes 0..3 stage G rows of positions p0..p0+3
        const int p = p0 + (lane & 3);
        const int i = at(cur.src, nxt.src, p < T ? p : T - 1);
        if (lane < 4 && p < T) {
            const uint32_t bar = bars + 8u * (p % S);
            sbar_expect(bar, grow_bytes);
            bulk_g2s(wbase + (uint32_t)(p % S) * grow_bytes, a.G + (long)i * a.ldg, grow_bytes, bar);
        }
    };
    if (T > 0)
        for (int p = 0; p < S; p += 4) issue4(p);

    unsigned act = TASKS ? 1u : __ballot_sync(0xffffffffu, st.deg > 0);
    if (!TASKS) {                                                  // segments without entries: zero rows
        unsigned empt = __ballot_sync(0xffffffffu, st.deg == 0 && !is_hub && seg0 + lane < a.n_seg);
        while (empt) {
            const int r = __ffs(empt) - 1;
            empt &= empt - 1;
            for (int c4 = lane; c4 < g.Wd4; c4 += 32)
                *reinterpret_cast<float4*>(a.outp + (seg0 + r) * a.ldout + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    int hc[NCH];
    float4 acc[NCH];
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        hc[ci] = (HT > 1 && c4 < g.Dt4) ? c4 / g.Dp4 : 0;
        acc[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int r = TASKS ? 0 : (act ? __ffs(act) - 1 : 0);
    act &= act - 1;
    int seg_end = T ? __shfl_sync(0xffffffffu, st.pre + st.deg, r) : 0;

    for (int n = 0; n < T; ++n) {
        if (n - nb == 32) {                                        // rotate the index pipeline
            cur = nxt;
            nb += 32;
            complete(pr, nb + 32, nxt);
            req_rec(pi, pr);
            req_idx(nb + 96, pi);
        }
        const uint32_t slot = wbase + (uint32_t)(n % S) * grow_bytes;
        sbar_wait(bars + 8u * (n % S), (uint32_t)(n / S) & 1u);
        float w[HT];
#pragma unroll
        for (int h = 0; h < HT; ++h) w[h] = __shfl_sync(0xffffffffu, cur.w[h], n - nb);
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
            const int c4 = lane + 32 * ci;
            if (c4 < g.Dt4) f4fma(acc[ci], selh<HT>(hc[ci], w), lds4(slot + (uint32_t)c4 * 16u));
        }
        if ((n & 3) == 3 || n + 1 == T) {                          // a group of 4 slots is free again
            __syncwarp();
            const int p0 = (n & ~3) + S;
            if (p0 < T) issue4(p0);
        }
        if (n + 1 == seg_end) {                                    // segment complete -> store [acc | vs | 0]
            float vs[HT];
#pragma unroll
            for (int h = 0; h < HT; ++h) vs[h] = __shfl_sync(0xffffffffu, vs_acc[h], r);
            float* dst = TASKS ? a.hub.partial + (long)(blockIdx.x * SS_WARPS + wid) * a.hub.ldpart
                               : a.outp + (seg0 + r) * a.ldout;
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
                const int c4 = lane + 32 * ci;
                if (c4 >= g.Wd4) continue;
                float4 o = acc[ci];
                if (c4 == g.Dt4) o = make_float4(vs[0], HT > 1 ? vs[HT > 1 ? 1 : 0] : 0.f, HT > 2 ? vs[HT > 2 ? 2 : 0] : 0.f,
                                                 HT > 3 ? vs[HT > 3 ? 3 : 0] : 0.f);
                else if (c4 > g.Dt4) o = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(dst + c4 * 4) = o;
                acc[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (act) {
                r = __ffs(act) - 1;
                act &= act - 1;
                seg_end = __shfl_sync(0xffffffffu, st.pre + st.deg, r);
            }
        }
    }
}

template <int NCH, int HT>
int launch_t(const SegGatherArgs& a, cudaStream_t s) {
    const size_t smem = SS_WARPS * ((size_t)SS_SLOTS * a.ldg * 4 + 128);
    static SmemLimit lim_rows, lim_tasks;
    if (a.n_seg > 0) {
        lim_rows.ensure(seg_gather_stream_kernel<NCH, HT, false>, smem);
        const unsigned grid = (unsigned)((a.n_seg + 32L * SS_WARPS - 1) / (32L * SS_WARPS));
        seg_gather_stream_kernel<NCH, HT, false><<<grid, SS_WARPS * 32, smem, s>>>(a);
        if (int rc = check_launch("seg_gather_stream")) return rc;
    }
    if (a.hub.n_tasks > 0) {
        lim_tasks.ensure(seg_gather_stream_kernel<NCH, HT, true>, smem);
        const unsigned grid = (a.hub.n_tasks + SS_WARPS - 1) / SS_WARPS;
        seg_gather_stream_kernel<NCH, HT, true><<<grid, SS_WARPS * 32, smem, s>>>(a);
        if (int rc = check_launch("seg_gather_stream_tasks")) return rc;
        return launch_seg_gather_hub_finalize(a, s);
    }
    return 0;
}

template <int NCH>
int launch_n(const SegGatherArgs& a, cudaStream_t s) {
    if (a.g.H == 1) return launch_t<NCH, 1>(a, s);
    if (a.g.H == 2) return launch_t<NCH, 2>(a, s);
    return launch_t<NCH, 4>(a, s);
}
}  // namespace

int launch_seg_gather_stream(const SegGatherArgs& a, cudaStream_t s) {
    // Used when the caller flags many short segments (multi-GPU CSC: ~2.5 entries per gathered node, where a warp per
    // segment wastes the machine) or with SPK_SEG_STREAM=1. At C2 on one GPU (10 entries per segment) the register-gather
    // kernel measured faster (cols 4.9 vs 5.8 ms, relation segments 3.1 vs 6.1 ms per launch): it is light enough
    // (74 registers) not to be occupancy-bound.
    static int enabled = -1;
    if (enabled < 0) { const char* e = getenv("SPK_SEG_STREAM"); enabled = (e && e[0] == '1') ? 1 : 0; }
    if (!(enabled || a.prefer_stream) || (a.ldg % 4) != 0) return -1;
    switch ((a.g.Wd4 + 31) / 32) {
        case 1: return launch_n<1>(a, s);
        case 2: return launch_n<2>(a, s);
        default: return -1;
    }
}

}  // namespace spk
