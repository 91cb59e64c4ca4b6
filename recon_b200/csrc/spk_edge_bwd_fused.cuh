// Argument block of the column-major fused backward (spk_edge_bwd_fused.cu).
#pragma once
#include "spk_edge.cuh"

namespace spk {

struct BwdFusedArgs {
    const int* rowptr;                    // [n_rows+1] CSR keyed on edge[0] (row sums of ds)
    const int* colptr;                    // [n_cols+1] CSC keyed on edge[1]
    const int* csc_row;                   // [E] aggregation row of the edge
    const int* csc_pos;                   // [E] CSR position of the edge (record / mask index)
    const int* csc_t1;                    // [E] relation id
    const int* csc_t2;                    // [E] second relation of a 2-hop edge (-1 for 1-hop); null if no 2-hop
    const float* P1; long ld1;
    const float* P2; long ld2;
    const float* P3; long ld3;
    const float* mask; long mask_stride;  // [H][E] dropout multipliers in CSR order, or null
    const float* out; const float* dout; long ldo;
    const float* den; const float* sw;    // [n_rows, H] saved by the forward
    float* G; long ldg;                   // [n_rows, ldg] dnum
    float* rowsc;                         // [n_rows, H, 4] (q1, c1, dden, 0)
    float* dP1; long ldd1;                // [n_rows, Wd]
    float* dP2; long ldd2;                // [n_cols, Wd]
    float* rec;                           // [E, 2H] (w, ds) in CSR order
    int n_rows, n_cols;
    LayerGeom g;
    float alpha;
    int apply_elu;
    int out_vec;
    HubTasks row_hub;                     // hub rows (hub_seg / n_hubs / hub_thresh used)
    HubTasks col_hub;                     // hub columns as tasks; partial: [n_tasks, >= Wd]
};

// column pass + relation pass with the dot t_e split between them (spk_edge_bwd_split.cu); graphs without 2-hop edges
struct BwdSplitArgs {
    BwdFusedArgs f;                       // node pass / row side: rowptr, P1, out, dout, den, sw, G, rowsc, dP1, row_hub, ...
    const int* colptr; const int* csc_row; const int* csc_pos; const int* csc_t1;
    const int* relptr; const int* rel_row; const int* rel_pos;
    const float* P2; long ld2;
    const float* P3; long ld3;
    const float* mask; long mask_stride;
    const float* G; long ldg;
    const float* rowsc;
    float* rec4;                          // [E, H, 4] (w, A, B, 0) in CSR order
    float* dsv;                           // [E, H] ds in CSR order
    float* dP2; long ldd2;                // [n_cols, Wd]
    float* dP3; long ldd3;                // [n_rel, Wd]
    int n_cols, n_rel;
    LayerGeom g;
    float alpha;
    HubTasks col_hub;                     // partial: [n_tasks, >= Wd]
    HubTasks rel_hub;                     // partial: [n_tasks, >= Wd]
    int phases;                           // bit 0: node pass, 1: column pass, 2: relation pass + row sums, 3: column sums
    float* colsum; long ld_colsum;        // optional destination of the column sums instead of the q slot of dP2~
    float* rowsum; long ld_rowsum;        // optional destination of the row sums instead of the q slot of dP1~
    const float* G_rel; long ldg_rel;     // rows gathered by the relation pass (default: G)
    int dup;                              // aggregate-then-project tables: P2 / P3 rows are [v (Dp) | H scalars | pad] and stand
                                          // for [v | v | ... | scalars]: the same input row under every head
};

int launch_edge_bwd_fused(const BwdFusedArgs& a, cudaStream_t s);
int launch_edge_bwd_node(const BwdFusedArgs& a, cudaStream_t s);      // node pass only
int launch_edge_bwd_split(const BwdSplitArgs& a, cudaStream_t s);
int launch_seg_gather_hub_finalize(const SegGatherArgs& a, cudaStream_t s);   // spk_edge_bwd.cu

}  // namespace spk
