// K0: on-device edge construction -> CSR / CSC / per-relation segment layouts (integer work, HBM-bound).
//
// Replaces the host-side Python of the reference that feeds the layers: the 1-hop / 2-hop
// concatenation of GAT/layers.py:124-127 with GAT/models.py:141-148, and the implicit
// "coalesce by row" inside SpecialSpmmFunctionFinal (GAT/layers.py:56-58). Edges are kept in a
// STABLE order inside every segment (original order: 1-hop block, then 2-hop block), so results do
// not depend on the sort implementation.
//
//   spk::edges_concat     int64 API tensors -> int32 row/col/t1/t2 arrays of the combined edge list
//   spk::radix_sort_pairs stable LSD radix sort (8-bit digits) of (key, value) int32 pairs
//   spk::segment_ptr      ptr[r] = lower_bound(sorted_keys, r)
//   spk::gather_i32       out[p] = src[idx[p]]
#include "spk_common.cuh"
#include "spk_graph.cuh"

namespace spk {
namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;                       // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;     // keys per CTA
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_WARP_SPAN = 32 * RS_ITEMS;        // consecutive keys owned by one warp

__global__ void edges_concat_kernel(const long long* __restrict__ edge, long e1, const long long* __restrict__ etype,
                                    const long long* __restrict__ nhop, long e2,
                                    int* __restrict__ row, int* __restrict__ col, int* __restrict__ t1,
                                    int* __restrict__ t2, long n_nodes, long n_rel, int* __restrict__ err) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e1 + e2) return;
    long long r, c, a, b = -1;
    if (e < e1) {
        r = edge[e]; c = edge[e1 + e]; a = etype[e];
    } else {                                         // models.py:145-148: [s, r1, r2, t] -> row=t, col=s
        const long long* q = nhop + (e - e1) * 4;
        r = q[3]; c = q[0]; a = q[1]; b = q[2];
    }
    if (r < 0 || r >= n_nodes || c < 0 || c >= n_nodes || a < 0 || a >= n_rel || b >= n_rel || (e >= e1 && b < 0))
        atomicOr(err, 1);
    row[e] = (int)r; col[e] = (int)c; t1[e] = (int)a;
    if (t2) t2[e] = (int)b;
}

__global__ void iota_kernel(int* __restrict__ v, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int)i;
}

// hist[digit * nblk + blk]
__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const int* __restrict__ keys, long n, int shift, unsigned* __restrict__ hist, int nblk) {
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const long base = (long)blockIdx.x * RS_TILE;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        const long i = base + it * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&sh[((unsigned)keys[i] >> shift) & 255u], 1u);   // integer counts: order-independent
    }
    __syncthreads();
    hist[(long)threadIdx.x * nblk + blockIdx.x] = sh[threadIdx.x];
}

// exclusive scan of `len` unsigned counters, single CTA with a running carry
__global__ void __launch_bounds__(1024)
rs_scan_kernel(unsigned* __restrict__ data, long len) {
    __shared__ unsigned warp_tot[32];
    __shared__ unsigned carry_s, chunk_total;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long base = 0; base < len; base += 4096) {
        const long i0 = base + (long)threadIdx.x * 4;
        unsigned v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i0 + k < len) ? data[i0 + k] : 0u;
        const unsigned tsum = v[0] + v[1] + v[2] + v[3];
        unsigned inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            const unsigned w = warp_tot[lane];
            unsigned winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_tot[lane] = winc - w;              // exclusive offset of each warp
            if (lane == 31) chunk_total = winc;
        }
        __syncthreads();
        const unsigned carry = carry_s;
        unsigned excl = carry + warp_tot[wid] + (inc - tsum);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < len) data[i0 + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + chunk_total;
        __syncthreads();
    }
}

// Multi-CTA form of the same scan (the single-CTA loop over 256 * nblk counters took 0.39 ms per radix pass at 20M keys,
// half of the whole CSR / CSC / relation build): per-tile totals, the single-CTA scan of those few hundred totals, then
// every tile rescans itself on top of its base.
constexpr int SCN_TILE = 4096;                     // 1024 threads x 4 counters

__device__ __forceinline__ unsigned scn_tile_excl(unsigned tsum, unsigned* warp_tot, unsigned& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const unsigned w = warp_tot[lane];
        unsigned winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_tot[lane] = winc - w;
        if (lane == 31) warp_tot[32] = winc;
    }
    __syncthreads();
    total = warp_tot[32];
    return warp_tot[wid] + (inc - tsum);
}

__global__ void __launch_bounds__(1024)
rs_scan_tile_sums_kernel(const unsigned* __restrict__ data, long len, unsigned* __restrict__ sums) {
    __shared__ unsigned warp_tot[33];
    const long i0 = (long)blockIdx.x * SCN_TILE + (long)threadIdx.x * 4;
    unsigned tsum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) tsum += (i0 + k < len) ? data[i0 + k] : 0u;
    unsigned total;
    scn_tile_excl(tsum, warp_tot, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
rs_scan_apply_kernel(unsigned* __restrict__ data, long len, const unsigned* __restrict__ sums) {
    __shared__ unsigned warp_tot[33];
    const long i0 = (long)blockIdx.x * SCN_TILE + (long)threadIdx.x * 4;
    unsigned v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < len) ? data[i0 + k] : 0u;
    unsigned total;
    unsigned excl = scn_tile_excl(v[0] + v[1] + v[2] + v[3], warp_tot, total) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i0 + k < len) data[i0 + k] = excl;
        excl += v[k];
    }
}

// Stable scatter: warp w of a CTA owns keys [tile + w*RS_WARP_SPAN, +RS_WARP_SPAN) and walks them in order.
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const int* __restrict__ keys, const int* __restrict__ vals, long n, int shift,
                  const unsigned* __restrict__ offs, int nblk, int* __restrict__ keys_out, int* __restrict__ vals_out) {
    __shared__ unsigned cnt[RS_WARPS][256];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < RS_WARPS * 256; d += RS_THREADS) (&cnt[0][0])[d] = 0;
    __syncthreads();
    const long wbase = (long)blockIdx.x * RS_TILE + (long)wid * RS_WARP_SPAN;
    int k[RS_ITEMS];
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        const long i = wbase + it * 32 + lane;
        k[it] = (i < n) ? keys[i] : 0;
        if (i < n) atomicAdd(&cnt[wid][((unsigned)k[it] >> shift) & 255u], 1u);
    }
    __syncthreads();
    // per digit: exclusive prefix over warps + global base of this CTA
    {
        const int d = threadIdx.x;                  // RS_THREADS == 256 digits
        unsigned run = offs[(long)d * nblk + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { const unsigned c = cnt[w][d]; cnt[w][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < RS_ITEMS; ++it) {
        const long i = wbase + it * 32 + lane;
        const bool ok = i < n;
        const unsigned dg = ((unsigned)k[it] >> shift) & 255u;
        const unsigned act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const unsigned peers = __match_any_sync(act, dg);
            const unsigned rank = __popc(peers & ((1u << lane) - 1u));
            const unsigned dst = cnt[wid][dg] + rank;
            keys_out[dst] = k[it];
            vals_out[dst] = vals[i];
            __syncwarp(act);
            if (rank == 0) cnt[wid][dg] += __popc(peers);     // one lane per distinct digit advances the cursor
        }
        __syncwarp();
    }
}

__global__ void segment_ptr_kernel(const int* __restrict__ sorted_keys, long n, int n_seg, int* __restrict__ ptr) {
    const long r = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_seg) return;
    long lo = 0, hi = n;                            // first position with key >= r
    while (lo < hi) {
        const long mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < (int)r) lo = mid + 1; else hi = mid;
    }
    ptr[r] = (int)lo;
}

__global__ void gather_i32_kernel(const int* __restrict__ src, const int* __restrict__ idx, long n, int* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

// incidence list of relation ids over CSR positions: entry p < E -> (t1[p], p); entry E+p -> (t2[p], p) with
// 1-hop edges (t2 < 0) sent to the dummy segment n_rel, which the consumers never read.
__global__ void rel_incidence_kernel(const int* __restrict__ t1, const int* __restrict__ t2, long e, int n_rel,
                                     int* __restrict__ keys, int* __restrict__ vals) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < e) { keys[i] = t1[i]; vals[i] = (int)i; }
    else if (t2 && i < 2 * e) { const long p = i - e; const int k = t2[p]; keys[i] = k >= 0 ? k : n_rel; vals[i] = (int)p; }
}
}  // namespace

static inline unsigned blocks_for(long n, int per) { return (unsigned)((n + per - 1) / per); }

int edges_concat(const long long* edge, long e1, const long long* etype, const long long* nhop, long e2,
                 int* row, int* col, int* t1, int* t2, long n_nodes, long n_rel, int* err, cudaStream_t s) {
    const long e = e1 + e2;
    if (e <= 0) return 0;
    edges_concat_kernel<<<blocks_for(e, 256), 256, 0, s>>>(edge, e1, etype, nhop, e2, row, col, t1, t2, n_nodes, n_rel, err);
    return check_launch("edges_concat");
}

int iota_i32(int* v, long n, cudaStream_t s) {
    if (n <= 0) return 0;
    iota_kernel<<<blocks_for(n, 256), 256, 0, s>>>(v, n);
    return check_launch("iota");
}

long radix_sort_workspace_bytes(long n) {
    const long nblk = (n + RS_TILE - 1) / RS_TILE;
    return (256L * nblk + 16 + (256L * nblk + SCN_TILE - 1) / SCN_TILE + 16) * sizeof(unsigned);   // histogram + tile sums of its scan
}

// Sorts (keys, vals) by the low `key_bits` bits of the key; stable. Uses (keys_tmp, vals_tmp) as ping-pong
// buffers; returns in *result_in_tmp whether the sorted data ended up in the tmp buffers.
int radix_sort_pairs(int* keys, int* vals, int* keys_tmp, int* vals_tmp, long n, int key_bits,
                     void* workspace, int* result_in_tmp, cudaStream_t s) {
    *result_in_tmp = 0;
    if (n <= 0) return 0;
    const int nblk = (int)((n + RS_TILE - 1) / RS_TILE);
    unsigned* hist = reinterpret_cast<unsigned*>(workspace);
    int* kin = keys; int* vin = vals; int* kout = keys_tmp; int* vout = vals_tmp;
    const int passes = (key_bits + 7) / 8 < 1 ? 1 : (key_bits + 7) / 8;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        rs_hist_kernel<<<nblk, RS_THREADS, 0, s>>>(kin, n, shift, hist, nblk);
        if (int rc = check_launch("rs_hist")) return rc;
        const long len = 256L * nblk;
        const long ntiles = (len + SCN_TILE - 1) / SCN_TILE;
        if (ntiles <= 2) {
            rs_scan_kernel<<<1, 1024, 0, s>>>(hist, len);
            if (int rc = check_launch("rs_scan")) return rc;
        } else {
            unsigned* sums = hist + len + 16;
            rs_scan_tile_sums_kernel<<<(unsigned)ntiles, 1024, 0, s>>>(hist, len, sums);
            if (int rc = check_launch("rs_scan_tile_sums")) return rc;
            rs_scan_kernel<<<1, 1024, 0, s>>>(sums, ntiles);
            if (int rc = check_launch("rs_scan")) return rc;
            rs_scan_apply_kernel<<<(unsigned)ntiles, 1024, 0, s>>>(hist, len, sums);
            if (int rc = check_launch("rs_scan_apply")) return rc;
        }
        rs_scatter_kernel<<<nblk, RS_THREADS, 0, s>>>(kin, vin, n, shift, hist, nblk, kout, vout);
        if (int rc = check_launch("rs_scatter")) return rc;
        int* t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
        *result_in_tmp ^= 1;
    }
    return 0;
}

int segment_ptr(const int* sorted_keys, long n, int n_seg, int* ptr, cudaStream_t s) {
    segment_ptr_kernel<<<blocks_for((long)n_seg + 1, 256), 256, 0, s>>>(sorted_keys, n, n_seg, ptr);
    return check_launch("segment_ptr");
}

int gather_i32(const int* src, const int* idx, long n, int* out, cudaStream_t s) {
    if (n <= 0) return 0;
    gather_i32_kernel<<<blocks_for(n, 256), 256, 0, s>>>(src, idx, n, out);
    return check_launch("gather_i32");
}

int rel_incidence(const int* t1, const int* t2, long e, int n_rel, int* keys, int* vals, cudaStream_t s) {
    const long m = t2 ? 2 * e : e;
    if (m <= 0) return 0;
    rel_incidence_kernel<<<blocks_for(m, 256), 256, 0, s>>>(t1, t2, e, n_rel, keys, vals);
    return check_launch("rel_incidence");
}

}  // namespace spk
