// "Aggregate-then-project" edge kernels for attention-layer groups whose INPUT is narrower than their
// projection (layer 1 of SpKBGAT: F = Rd = 50 against H*D = 200).
//
// The reference computes edge_m = a.[x_i | x_j | r_k] per edge and sums ee_e * edge_m over the row
// (GAT/layers.py:129-165). The sum is linear in the gathered vectors, so for head h
//     num_i = a_h . [ (sum_e w_e) x_i | sum_e w_e x_j | sum_e w_e r_k ]          (w_e = dropout(ee_e))
// and the per-edge work only touches the F- and Rd-wide INPUT rows; the projection by a_h happens once per
// row afterwards, as one tensor-core GEMM over the aggregated block Z. The attention score needs no wide
// row at all: s_e = q1[i] + q2[j] + q3[k] with q* = x . (A*^T a_2^T) carried as extra table columns.
//
// Table rows ("X~", "Rel~"):  [ x (F) | 0-pad to 4*Fx4 | q2_0 q2_1 q1_0 q1_1 | 0-pad to LX ]   (LX % 8 == 0)
//                             [ r (Rd)| 0-pad to 4*Fr4 | q3_0 q3_1 0    0    | 0-pad to LR ]
// A warp handles one aggregation row; lanes 0-15 own the float4 chunks of x_j, lanes 16-31 those of r_k
// (needs Fx4 <= 15, Fr4 <= 15, H <= 2). Aggregated block per row and head (LZ = 4*(2*Fx4+Fr4) floats):
//     Zn_h = [ sw_h x_i | sum w x_j | sum w r_k ] / den_h
#pragma once
#include "spk_edge.cuh"

namespace spk {

struct AggGeom {
    int H;        // heads (1 or 2)
    int Fx4;      // ceil(F / 4): float4 chunks of an input row; chunk Fx4 holds the score scalars
    int Fr4;      // ceil(Rd / 4)
    int LZ;       // 4 * (2*Fx4 + Fr4): floats of Zn per head
};

struct AggFwdArgs {
    const int* segptr; const int* col; const int* t1; const int* t2;
    const float* Xrow; long ldxr;     // [n_rows, >= 4*Fx4+4] X~ of the aggregation rows
    const float* Xcol; long ldxc;     // [n_cols, >= 4*Fx4+4] X~ of the gathered nodes
    const float* Rt; long ldr;        // [R, >= 4*Fr4+4]      Rel~
    const float* mask; long mask_stride;
    float* Z; long ldz;               // [n_rows, H*LZ]
    float* den; float* sw;            // [n_rows, H]
    int* nanflag;
    int n_rows;
    float alpha;
    AggGeom g;
    HubTasks hub;                     // partial: [n_tasks, 264] = acc[2][128], den[4], sw[4]
};

struct AggBwdArgs {
    const int* segptr; const int* col; const int* t1; const int* t2;
    const float* Xrow; long ldxr;
    const float* Xcol; long ldxc;
    const float* Rt; long ldr;
    const float* mask; long mask_stride;
    const float* dZ; long ldz;        // [n_rows, H*LZ] gradient w.r.t. Zn
    const float* den; const float* sw; const float* dden;    // [n_rows, H]
    float* Gx; long ldgx;             // [n_rows, H*4*Fx4]  dZ_h[x_j part] / den_h   (row context; gathered by the column pass)
    float* Gr; long ldgr;             // [n_rows, H*4*Fr4]  dZ_h[r_k part] / den_h   (row context; gathered by the relation pass)
    float* rowout; long ldro;         // [n_rows, 4*Fx4+4]  sum_h sw_h dZ_h[x_i part]/den_h | dq1_0 dq1_1 0 0
    float* rowsc;                     // [n_rows, 8] scratch: q1_0 q1_1 c_0 c_1 dden_0 dden_1 0 0
    float* rec;                       // [E, 2H] (w, ds) per head, CSR order
    int n_rows;
    float alpha;
    AggGeom g;
    HubTasks hub;                     // partial: [n_tasks, 8] = dq1[4], unused[4]
};

constexpr int AGG_LDPART = 264;

int launch_agg_table(const float* X, long ldx, const float* V, float* T, long ldt, long n, int F, int F4, cudaStream_t s);
int launch_agg_fwd(const AggFwdArgs& a, cudaStream_t s);
int launch_agg_bwd_pre(const float* out, const float* dout, long ldo, const float* den, int H, int D, int apply_elu,
                       float* dhn, long ldd, float* dden, long n, cudaStream_t s);
int launch_agg_bwd_rows(const AggBwdArgs& a, cudaStream_t s);
int launch_agg_bwd_ctx_split(const AggBwdArgs& a, cudaStream_t s);   // row-context kernel alone, rowsc as [n, H, 4]
int launch_agg_dx(const float* rowout, long ldro, const float* dxc, long ldc, const float* V, long n, int F, int F4,
                  int H, float* dX, long lddx, float* dq, cudaStream_t s);
int launch_elu_inplace(float* x, long ld, long n, int width, cudaStream_t s);

}  // namespace spk
