// Argument blocks of the edge-phase kernels (K2 forward, K3 backward rows, K4 backward segments).
#pragma once
#include "spk_common.cuh"

namespace spk {

// Hub handling shared by all edge kernels: segments longer than hub_thresh are skipped by the
// warp-per-segment kernel and processed as fixed-size chunks ("tasks"), each writing a partial
// that a finalize kernel adds up in task order (deterministic, no atomics).
struct HubTasks {
    const int* task_seg;      // [n_tasks] segment (row / column / relation) of the task
    const int* task_beg;      // [n_tasks] first entry
    const int* task_end;      // [n_tasks] one past last entry
    const int* hub_seg;       // [n_hubs] segment id of each hub
    const int* hub_task_ptr;  // [n_hubs+1] tasks of hub h are [ptr[h], ptr[h+1])
    float* partial;           // [n_tasks, ldpart]
    long ldpart;
    int n_tasks;
    int n_hubs;
    int hub_thresh;
    const int* task_order;    // [n_tasks] launch slot -> task (null: identity); partials stay indexed by task
};

__device__ __forceinline__ int hub_task_of_slot(const HubTasks& h, int slot) {
    return h.task_order ? __ldg(h.task_order + slot) : slot;
}

struct EdgeFwdArgs {
    const int* segptr;        // [n_rows+1] CSR keyed on edge[0]
    const int* col;           // [E] edge[1] = gather index into P2
    const int* t1;            // [E] relation id
    const int* t2;            // [E] second relation of a 2-hop edge, -1 for 1-hop; null if no 2-hop
    const float* P1; long ld1;   // [n_rows, ld1]  X*A1^T | q1
    const float* P2; long ld2;   // [n_cols, ld2]  X*A2^T | q2
    const float* P3; long ld3;   // [R, ld3]       Rel*A3^T | q3
    const float* mask; long mask_stride;   // [H][E] dropout multipliers in CSR order, or null
    float* out; long ldo;     // [n_rows, H*D]
    float* den;               // [n_rows, H] row sums of exp (zeros replaced by 1e-12)
    float* sw;                // [n_rows, H] row sums of dropped exp
    int* nanflag;
    int n_rows;
    LayerGeom g;
    float alpha;
    int apply_elu;
    int elu_rows;             // with apply_elu: only rows < elu_rows get the ELU (rows behind them are partial sums that a
                              // caller combines first: the ghost rows of hub rows split across ranks)
    int out_vec;              // out rows may be stored as float4
    HubTasks hub;
};

struct EdgeBwdRowsArgs {
    const int* segptr; const int* col; const int* t1; const int* t2;
    const float* P1; long ld1;
    const float* P2; long ld2;
    const float* P3; long ld3;
    const float* mask; long mask_stride;
    const float* out; const float* dout; long ldo;   // saved forward output, upstream gradient
    const float* den;         // [n_rows, H]
    float* G; long ldg;       // [n_rows, ldg] dnum = dh / den
    float* dP1; long ldd1;    // [n_rows, Wd]  sw*dnum | u | 0
    float* rec;               // [E, 2H] (w, ds) per head in CSR order
    int n_rows;
    LayerGeom g;
    float alpha;
    int apply_elu;
    int out_vec;
    HubTasks hub;             // partial: [n_tasks, 8] = u[4], sw[4]
};

struct SegGatherArgs {
    const int* segptr;        // [n_seg+1]
    const int* src;           // [M] row of G to gather
    const int* pos;           // [M] CSR position of the edge (record index)
    const float* G; long ldg;
    const float* rec;         // [E, 2H]
    float* outp; long ldout;  // [n_seg, Wd]  sum w*G[src] | sum ds | 0
    int n_seg;
    int prefer_stream;        // many short segments (avg < ~5 entries): use the 32-segments-per-warp streaming kernel
    LayerGeom g;
    HubTasks hub;             // partial: [n_tasks, Wd + 4]
};

int launch_edge_fwd(const EdgeFwdArgs& a, cudaStream_t s);
int launch_edge_bwd_rows(const EdgeBwdRowsArgs& a, cudaStream_t s);
int launch_seg_gather(const SegGatherArgs& a, cudaStream_t s);

}  // namespace spk
