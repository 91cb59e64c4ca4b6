// Pieces of the backward-rows pass (K3) shared by the register-gather and the streaming kernels.
#pragma once
#include "spk_edge.cuh"

namespace spk {

// ---- loads of the unpadded [n_rows, H*D] activations into padded float4 chunks ----------------
__device__ __forceinline__ float4 load_act_chunk(const float* base, int h, int off, int D, int vec) {
    const float* p = base + (long)h * D + off;
    if (vec) return ldg4(p);
    float4 r;
    r.x = off + 0 < D ? __ldg(p + 0) : 0.f;
    r.y = off + 1 < D ? __ldg(p + 1) : 0.f;
    r.z = off + 2 < D ? __ldg(p + 2) : 0.f;
    r.w = off + 3 < D ? __ldg(p + 3) : 0.f;
    return r;
}

template <int NCH, int HT>
struct RowCtx {
    float4 dnum[NCH];
    int hc[NCH];
    float q1[HT];
    float dden[HT];
    float c1[HT];       // dnum . P1[i] per head
};

template <int NCH, int HT>
__device__ __forceinline__ void bwd_row_prologue(const EdgeBwdRowsArgs& a, int row, int lane, RowCtx<NCH, HT>& rc) {
    const LayerGeom g = a.g;
    float pdh[HT], pc1[HT], den[HT];
#pragma unroll
    for (int h = 0; h < HT; ++h) {
        pdh[h] = 0.f; pc1[h] = 0.f;
        den[h] = h < g.H ? __ldg(a.den + (long)row * g.H + h) : 1.f;
    }
    const float4 q1v = ldg4(a.P1 + (long)row * a.ld1 + (long)g.Dt4 * 4);
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
        const int off = (c4 - h * g.Dp4) * 4;
        const float4 o4 = load_act_chunk(a.out + (long)row * a.ldo, h, off, g.D, a.out_vec);
        const float4 g4 = load_act_chunk(a.dout + (long)row * a.ldo, h, off, g.D, a.out_vec);
        const float o[4] = {o4.x, o4.y, o4.z, o4.w};
        const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
        float dh[4], hv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (a.apply_elu) {
                // out = ELU(h): h>0 -> out=h, d=1 ; else out=e^h-1, d=out+1, h=log1p(out)
                const bool pos = o[k] > 0.f;
                dh[k] = pos ? gg[k] : gg[k] * (o[k] + 1.f);
                hv[k] = pos ? o[k] : (o[k] > -1.f ? fast_log1p(o[k]) : 0.f);
            } else {
                dh[k] = gg[k];
                hv[k] = o[k];
            }
        }
        const float d = selh<HT>(h, den);
        const float dot_h = fmaf(dh[0], hv[0], fmaf(dh[1], hv[1], fmaf(dh[2], hv[2], dh[3] * hv[3])));
        rc.dnum[ci] = make_float4(dh[0] / d, dh[1] / d, dh[2] / d, dh[3] / d);
        const float4 p1 = ldg4(a.P1 + (long)row * a.ld1 + c4 * 4);
        const float dot_c = f4dot(rc.dnum[ci], p1);
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

template <int NCH, int HT>
__device__ __forceinline__ void bwd_row_store(const EdgeBwdRowsArgs& a, int row, int lane, const RowCtx<NCH, HT>& rc,
                                              const float (&usum)[HT], const float (&swsum)[HT],
                                              bool store_g, bool store_dp1) {
    const LayerGeom g = a.g;
#pragma unroll
    for (int ci = 0; ci < NCH; ++ci) {
        const int c4 = lane + 32 * ci;
        if (store_g && c4 * 4 < a.ldg)
            *reinterpret_cast<float4*>(a.G + (long)row * a.ldg + c4 * 4) = rc.dnum[ci];
        if (store_dp1 && c4 < g.Wd4) {
            float4 o;
            if (c4 < g.Dt4) {
                const float s = selh<HT>(rc.hc[ci], swsum);
                o = make_float4(s * rc.dnum[ci].x, s * rc.dnum[ci].y, s * rc.dnum[ci].z, s * rc.dnum[ci].w);
            } else if (c4 == g.Dt4) {
                o = make_float4(usum[0], HT > 1 ? usum[HT > 1 ? 1 : 0] : 0.f, HT > 2 ? usum[HT > 2 ? 2 : 0] : 0.f,
                                HT > 3 ? usum[HT > 3 ? 3 : 0] : 0.f);
            } else {
                o = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            *reinterpret_cast<float4*>(a.dP1 + (long)row * a.ldd1 + c4 * 4) = o;
        }
    }
}


int launch_edge_bwd_rows_stream(const EdgeBwdRowsArgs& a, cudaStream_t s);   // -1: shape not supported by the stream path

}  // namespace spk
