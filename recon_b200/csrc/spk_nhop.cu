// K0b behind the C ABI: the batch adjacency and the 2-hop path rows of a batch of source entities, bit-exact (values AND
// order) with the reference's Python dict / queue code:
//   Corpus.get_batch_adj_data            GAT/create_batch.py:391-436   -> ([trgts; srcs], vals)
//   Corpus.bfs + get_further_neighbors   GAT/create_batch.py:788-869   FIFO BFS, first discoverer is the parent
//   Corpus.get_batch_nhop_neighbors_all  GAT/create_batch.py:871-895   -> rows [s, r(s->m)[0], r(m->t)[0], t]
// Input: the distinct-neighbour adjacency of the triple list in the reference's insertion order (recon_b200.nhop.TripleGraph:
// uptr / ut / ur0 / ugs / uge / rs). Instead of one BFS per source the whole batch is expanded level by level with
// prefix sums, and the BFS tie-breaks are recovered with STABLE radix sorts (spk_graph.cu): a 2-hop target t of source s is
// kept iff it is not s, not a 1-hop neighbour of s, and it is the FIRST candidate reaching t in (mid, tail) discovery order
// -- sort the records (blockers first, then candidates in discovery order) stably by (s, t) and keep the group heads that
// are candidates. Integer work, HBM-bound. Sizes are data dependent, so buffers come from the caller's allocator callback
// (PyTorch's caching allocator in recon_b200); the library itself never calls cudaMalloc.
#include "../../include/spkbgat.h"
#include "spk_common.cuh"
#include "spk_graph.cuh"

namespace spk {
namespace {

constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ int block_excl_scan(int v, int* sh, int& total) {      // 256 threads
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < SC_THREADS / 32 ? sh[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < SC_THREADS / 32) sh[lane] = winc - w;
        if (lane == SC_THREADS / 32 - 1) sh[SC_THREADS / 32] = winc;
    }
    __syncthreads();
    total = sh[SC_THREADS / 32];
    const int r = sh[wid] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SC_THREADS)
scan_tile_sums_kernel(const int* __restrict__ in, long n, int* __restrict__ sums) {
    __shared__ int sh[SC_THREADS / 32 + 1];
    const long base = (long)blockIdx.x * SC_TILE + (long)threadIdx.x * SC_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) s += base + k < n ? in[base + k] : 0;
    int total;
    block_excl_scan(s, sh, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the tile sums in place, grand total -> sums[ntiles]
__global__ void __launch_bounds__(SC_THREADS)
scan_sums_kernel(int* __restrict__ sums, int ntiles) {
    __shared__ int sh[SC_THREADS / 32 + 1];
    int carry = 0;
    for (int base = 0; base < ntiles; base += SC_THREADS) {
        const int i = base + threadIdx.x;
        const int v = i < ntiles ? sums[i] : 0;
        int total;
        const int ex = block_excl_scan(v, sh, total);
        if (i < ntiles) sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) sums[ntiles] = carry;
}

// out[i] = exclusive prefix, out[n] = total
__global__ void __launch_bounds__(SC_THREADS)
scan_apply_kernel(const int* __restrict__ in, long n, const int* __restrict__ sums, int ntiles, int* __restrict__ out) {
    __shared__ int sh[SC_THREADS / 32 + 1];
    const long base = (long)blockIdx.x * SC_TILE + (long)threadIdx.x * SC_ITEMS;
    int v[SC_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
    int total;
    int ex = block_excl_scan(s, sh, total) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[ntiles];
}

// position p of the expanded list -> (segment, offset inside it): off[seg] <= p < off[seg + 1]
__device__ __forceinline__ int seg_of(const int* __restrict__ off, int n, int p) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(off + mid) <= p) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void sources_kernel(const long long* __restrict__ src, long nb, int n_nodes, const int* __restrict__ uptr,
                               int* __restrict__ s32, int* __restrict__ cnt1, int* __restrict__ err) {
    const long b = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    long long s = src[b];
    if (s < 0 || s >= n_nodes) { atomicOr(err, 1); s = 0; }
    s32[b] = (int)s;
    cnt1[b] = uptr[s + 1] - uptr[s];
}

// level 1 (create_batch.py:806-816 at depth 1): distinct out-neighbours of every source in first-seen order; a self loop is
// "already visited" and dropped
__global__ void level1_kernel(const int* __restrict__ off1, int nb, int L1, const int* __restrict__ s32,
                              const int* __restrict__ uptr, const int* __restrict__ ut,
                              int* __restrict__ b_out, int* __restrict__ a_out, int* __restrict__ m_out, int* __restrict__ keep) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= L1) return;
    const int b = seg_of(off1, nb, p);
    const int s = s32[b];
    const int a = uptr[s] + (p - off1[b]);
    const int m = ut[a];
    b_out[p] = b; a_out[p] = a; m_out[p] = m; keep[p] = m != s;
}

__global__ void compact3_kernel(const int* __restrict__ keep, const int* __restrict__ pos, int n, const int* __restrict__ x0,
                                const int* __restrict__ x1, const int* __restrict__ x2, int* __restrict__ y0,
                                int* __restrict__ y1, int* __restrict__ y2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const int q = pos[i];
    y0[q] = x0[i];
    if (x1) y1[q] = x1[i];
    if (x2) y2[q] = x2[i];
}

__global__ void pair_counts_kernel(const int* __restrict__ l1_a, const int* __restrict__ l1_m, int k1,
                                   const int* __restrict__ ugs, const int* __restrict__ uge, const int* __restrict__ uptr,
                                   int* __restrict__ ce, int* __restrict__ cnt2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k1) return;
    const int a = l1_a[i];
    ce[i] = uge[a] - ugs[a];
    if (cnt2) { const int m = l1_m[i]; cnt2[i] = uptr[m + 1] - uptr[m]; }
}

// batch adjacency (create_batch.py:413-436): every parallel relation of every kept (source, neighbour) pair
__global__ void adjacency_kernel(const int* __restrict__ offe, int k1, int E1, const int* __restrict__ l1_b,
                                 const int* __restrict__ l1_a, const int* __restrict__ l1_m, const int* __restrict__ s32,
                                 const int* __restrict__ ugs, const int* __restrict__ rs,
                                 long long* __restrict__ adj_idx, long long* __restrict__ adj_val) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E1) return;
    const int i = seg_of(offe, k1, p);
    adj_idx[p] = l1_m[i];                               // trgts
    adj_idx[(long)E1 + p] = s32[l1_b[i]];               // srcs
    adj_val[p] = rs[ugs[l1_a[i]] + (p - offe[i])];
}

// level-2 candidates in discovery order (source, mid in level-1 order, tail in mid's order) appended behind the blockers
// (the source itself, its level-1 nodes): record = (batch slot, node)
__global__ void records_kernel(const int* __restrict__ off2, int nb, int k1, int Cn, const int* __restrict__ s32,
                               const int* __restrict__ l1_b, const int* __restrict__ l1_m, const int* __restrict__ uptr,
                               const int* __restrict__ ut, int* __restrict__ rb, int* __restrict__ rt,
                               int* __restrict__ c_l1, int* __restrict__ c_c) {
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nr = (long)nb + k1 + Cn;
    if (p >= nr) return;
    if (p < nb) { rb[p] = (int)p; rt[p] = s32[p]; return; }
    if (p < nb + k1) { const int i = (int)(p - nb); rb[p] = l1_b[i]; rt[p] = l1_m[i]; return; }
    const int c = (int)(p - nb - k1);
    const int i = seg_of(off2, k1, c);
    const int cc = uptr[l1_m[i]] + (c - off2[i]);
    c_l1[c] = i; c_c[c] = cc;
    rb[p] = l1_b[i]; rt[p] = ut[cc];
}

// after the stable sort by (slot, node): the head of every group wins; it is an accepted 2-hop target iff it is a candidate
__global__ void accept_kernel(const int* __restrict__ ob, const int* __restrict__ ot, const int* __restrict__ perm, long nr,
                              int n_block, int* __restrict__ flag) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr) return;
    const bool head = i == 0 || ob[i] != ob[i - 1] || ot[i] != ot[i - 1];
    flag[i] = head && perm[i] >= n_block;
}

__global__ void accept_scatter_kernel(const int* __restrict__ flag, const int* __restrict__ pos, const int* __restrict__ perm,
                                      long nr, int n_block, int* __restrict__ acc) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr || !flag[i]) return;
    acc[pos[i]] = perm[i] - n_block;
}

// create_batch.py:883-884 (partial_2hop): only the first path of every source
__global__ void first_of_source_kernel(const int* __restrict__ acc, int na, const int* __restrict__ c_l1,
                                       const int* __restrict__ l1_b, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na) return;
    flag[i] = i == 0 || l1_b[c_l1[acc[i]]] != l1_b[c_l1[acc[i - 1]]];
}

__global__ void nhop_rows_kernel(const int* __restrict__ acc, int na, const int* __restrict__ c_l1, const int* __restrict__ c_c,
                                 const int* __restrict__ l1_b, const int* __restrict__ l1_a, const int* __restrict__ s32,
                                 const int* __restrict__ ur0, const int* __restrict__ ut, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na) return;
    const int c = acc[i];
    const int l = c_l1[c], cc = c_c[c];
    out[4 * i + 0] = s32[l1_b[l]];
    out[4 * i + 1] = ur0[l1_a[l]];
    out[4 * i + 2] = ur0[cc];
    out[4 * i + 3] = ut[cc];
}

// Scratch is carved out of a few large blocks obtained from the caller's allocator (one callback per ~64 MB instead of one
// per array); results are separate allocations so that the caller can hand them out as tensors of their own.
struct Arena {
    spk_alloc_fn fn; void* ctx; bool failed = false;
    char* cur = nullptr; size_t left = 0;
    static constexpr size_t BLOCK = 64u << 20;
    template <class T> T* get(long n) {
        size_t bytes = ((size_t)(n > 0 ? n : 1) * sizeof(T) + 255) & ~(size_t)255;
        if (bytes > left) {
            const size_t want = bytes > BLOCK ? bytes : BLOCK;
            cur = reinterpret_cast<char*>(fn(ctx, (int64_t)want));
            if (!cur) { failed = true; left = 0; return nullptr; }
            left = want;
        }
        T* p = reinterpret_cast<T*>(cur);
        cur += bytes; left -= bytes;
        return p;
    }
    template <class T> T* result(long n) {
        void* p = fn(ctx, (int64_t)(n > 0 ? n : 1) * (int64_t)sizeof(T));
        if (!p) failed = true;
        return reinterpret_cast<T*>(p);
    }
};

inline unsigned blocks(long n, int t = 256) { return (unsigned)((n + t - 1) / t); }

int exscan(const int* in, long n, int* out, Arena& ar, cudaStream_t s) {        // out has n + 1 entries
    const int ntiles = (int)((n + SC_TILE - 1) / SC_TILE);
    int* sums = ar.get<int>(ntiles + 1);
    if (ar.failed) return 4;
    if (ntiles > 0) {
        scan_tile_sums_kernel<<<ntiles, SC_THREADS, 0, s>>>(in, n, sums);
        if (int rc = check_launch("scan_tile_sums")) return rc;
    }
    scan_sums_kernel<<<1, SC_THREADS, 0, s>>>(sums, ntiles);
    if (int rc = check_launch("scan_sums")) return rc;
    if (ntiles > 0) {
        scan_apply_kernel<<<ntiles, SC_THREADS, 0, s>>>(in, n, sums, ntiles, out);
        return check_launch("scan_apply");
    }
    cudaMemcpyAsync(out, sums, sizeof(int), cudaMemcpyDeviceToDevice, s);          // n == 0: out[0] = 0
    return 0;
}

int read_int(const int* dev, cudaStream_t s, long* value) {
    int v = 0;
    if (cudaMemcpyAsync(&v, dev, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) {
        set_error("nhop_build: size read-back failed (%s)", cudaGetErrorString(cudaGetLastError()));
        return 3;
    }
    *value = v;
    return 0;
}

int key_bits_of(long n) { int b = 1; while ((1L << b) < n && b < 31) ++b; return b; }

// stable sort of (key, val) pairs; returns the buffers holding the result
int sort_stable(int*& keys, int*& vals, long n, int bits, Arena& ar, cudaStream_t s) {
    if (n <= 1) return 0;
    int* kt = ar.get<int>(n); int* vt = ar.get<int>(n);
    void* ws = ar.get<char>(radix_sort_workspace_bytes(n));
    if (ar.failed) return 4;
    int in_tmp = 0;
    if (int rc = radix_sort_pairs(keys, vals, kt, vt, n, bits, ws, &in_tmp, s)) return rc;
    if (in_tmp) { keys = kt; vals = vt; }
    return 0;
}

}  // namespace
}  // namespace spk

extern "C" int spk_nhop_build(const spk_triple_graph* g, const int64_t* sources, int64_t n_sources, int32_t flags,
                              spk_alloc_fn alloc, void* alloc_ctx, spk_nhop_result* res, spk_stream_t stream) {
    using namespace spk;
    if (!g || !res || !alloc || n_sources < 0 || (n_sources > 0 && !sources)) { set_error("nhop_build: bad arguments"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    Arena ar{alloc, alloc_ctx};
    const bool partial = flags & 1, want_nhop = !(flags & 2);
    res->adj_idx = nullptr; res->adj_val = nullptr; res->nhop = nullptr; res->e1 = 0; res->e2 = 0;
    const long nb = (long)n_sources;
    if (nb >= (1L << 31)) { set_error("nhop_build: too many sources"); return 1; }
    int* err = ar.get<int>(1);
    int* s32 = ar.get<int>(nb); int* cnt1 = ar.get<int>(nb); int* off1 = ar.get<int>(nb + 1);
    if (ar.failed) { set_error("nhop_build: allocation failed"); return 4; }
    cudaMemsetAsync(err, 0, sizeof(int), s);
    long L1 = 0, K1 = 0, E1 = 0, Cn = 0, NA = 0;
    if (nb > 0) {
        sources_kernel<<<blocks(nb), 256, 0, s>>>(reinterpret_cast<const long long*>(sources), nb, g->n_nodes, g->uptr, s32, cnt1, err);
        if (int rc = check_launch("nhop_sources")) return rc;
    }
    if (int rc = exscan(cnt1, nb, off1, ar, s)) return rc;
    if (int rc = read_int(off1 + nb, s, &L1)) return rc;
    long bad = 0;
    if (int rc = read_int(err, s, &bad)) return rc;
    if (bad) { set_error("nhop_build: source id out of range"); return 5; }
    // ---- level 1
    int* t_b = ar.get<int>(L1); int* t_a = ar.get<int>(L1); int* t_m = ar.get<int>(L1); int* keep = ar.get<int>(L1);
    int* kpos = ar.get<int>(L1 + 1);
    if (ar.failed) { set_error("nhop_build: allocation failed"); return 4; }
    if (L1 > 0) {
        level1_kernel<<<blocks(L1), 256, 0, s>>>(off1, (int)nb, (int)L1, s32, g->uptr, g->ut, t_b, t_a, t_m, keep);
        if (int rc = check_launch("nhop_level1")) return rc;
    }
    if (int rc = exscan(keep, L1, kpos, ar, s)) return rc;
    if (int rc = read_int(kpos + L1, s, &K1)) return rc;
    int* l1_b = ar.get<int>(K1); int* l1_a = ar.get<int>(K1); int* l1_m = ar.get<int>(K1);
    int* ce = ar.get<int>(K1); int* offe = ar.get<int>(K1 + 1);
    int* cnt2 = want_nhop ? ar.get<int>(K1) : nullptr; int* off2 = want_nhop ? ar.get<int>(K1 + 1) : nullptr;
    if (ar.failed) { set_error("nhop_build: allocation failed"); return 4; }
    if (L1 > 0) {
        compact3_kernel<<<blocks(L1), 256, 0, s>>>(keep, kpos, (int)L1, t_b, t_a, t_m, l1_b, l1_a, l1_m);
        if (int rc = check_launch("nhop_compact")) return rc;
    }
    if (K1 > 0) {
        pair_counts_kernel<<<blocks(K1), 256, 0, s>>>(l1_a, l1_m, (int)K1, g->ugs, g->uge, g->uptr, ce, cnt2);
        if (int rc = check_launch("nhop_pair_counts")) return rc;
    }
    // ---- batch adjacency
    if (int rc = exscan(ce, K1, offe, ar, s)) return rc;
    if (int rc = read_int(offe + K1, s, &E1)) return rc;
    res->adj_idx = ar.result<int64_t>(2 * E1); res->adj_val = ar.result<int64_t>(E1); res->e1 = E1;
    if (ar.failed) { set_error("nhop_build: allocation failed"); return 4; }
    if (E1 > 0) {
        adjacency_kernel<<<blocks(E1), 256, 0, s>>>(offe, (int)K1, (int)E1, l1_b, l1_a, l1_m, s32, g->ugs, g->rs,
                                                     reinterpret_cast<long long*>(res->adj_idx), reinterpret_cast<long long*>(res->adj_val));
        if (int rc = check_launch("nhop_adjacency")) return rc;
    }
    if (!want_nhop) { res->nhop = ar.result<int32_t>(0); return ar.failed ? 4 : 0; }
    // ---- level-2 candidates + blockers, stable sort by (slot, node)
    if (int rc = exscan(cnt2, K1, off2, ar, s)) return rc;
    if (int rc = read_int(off2 + K1, s, &Cn)) return rc;
    const long n_block = nb + K1, NR = n_block + Cn;
    if (NR >= (1L << 31)) { set_error("nhop_build: %ld records exceed int32; use smaller batches", NR); return 1; }
    int* rb = ar.get<int>(NR); int* rt = ar.get<int>(NR); int* c_l1 = ar.get<int>(Cn); int* c_c = ar.get<int>(Cn);
    int* perm = ar.get<int>(NR); int* key = ar.get<int>(NR);
    if (ar.failed) { set_error("nhop_build: allocation failed"); return 4; }
    if (NR > 0) {
        records_kernel<<<blocks(NR), 256, 0, s>>>(off2, (int)nb, (int)K1, (int)Cn, s32, l1_b, l1_m, g->uptr, g->ut, rb, rt, c_l1, c_c);
        if (int rc = check_launch("nhop_records")) return rc;
        if (int rc = iota_i32(perm, NR, s)) return rc;
        cudaMemcpyAsync(key, rt, NR * sizeof(int), cudaMemcpyDeviceToDevice, s);
        if (int rc = sort_stable(key, perm, NR, key_bits_of(g->n_nodes), ar, s)) return rc;         // minor key: node
        int* key2 = ar.get<int>(NR);
        if (ar.failed) { set_error("nhop_build: allocation failed"); return 4; }
        if (int rc = gather_i32(rb, perm, NR, key2, s)) return rc;
        if (int rc = sort_stable(key2, perm, NR, key_bits_of(nb), ar, s)) return rc;                // major key: batch slot
        int* ot = ar.get<int>(NR); int* flag = ar.get<int>(NR); int* fpos = ar.get<int>(NR + 1);
        if (ar.failed) { set_error("nhop_build: allocation failed"); return 4; }
        if (int rc = gather_i32(rt, perm, NR, ot, s)) return rc;
        accept_kernel<<<blocks(NR), 256, 0, s>>>(key2, ot, perm, NR, (int)n_block, flag);
        if (int rc = check_launch("nhop_accept")) return rc;
        if (int rc = exscan(flag, NR, fpos, ar, s)) return rc;
        if (int rc = read_int(fpos + NR, s, &NA)) return rc;
        int* acc = ar.get<int>(NA); int* dummy = ar.get<int>(NA);
        if (ar.failed) { set_error("nhop_build: allocation failed"); return 4; }
        if (NA > 0) {
            accept_scatter_kernel<<<blocks(NR), 256, 0, s>>>(flag, fpos, perm, NR, (int)n_block, acc);
            if (int rc = check_launch("nhop_accept_scatter")) return rc;
            cudaMemsetAsync(dummy, 0, NA * sizeof(int), s);
            if (int rc = sort_stable(acc, dummy, NA, key_bits_of(Cn + 1), ar, s)) return rc;          // back to discovery order
            if (partial) {
                int* f2 = ar.get<int>(NA); int* p2 = ar.get<int>(NA + 1);
                if (ar.failed) { set_error("nhop_build: allocation failed"); return 4; }
                first_of_source_kernel<<<blocks(NA), 256, 0, s>>>(acc, (int)NA, c_l1, l1_b, f2);
                if (int rc = check_launch("nhop_first_of_source")) return rc;
                if (int rc = exscan(f2, NA, p2, ar, s)) return rc;
                long NA2 = 0;
                if (int rc = read_int(p2 + NA, s, &NA2)) return rc;
                int* acc2 = ar.get<int>(NA2);
                if (ar.failed) { set_error("nhop_build: allocation failed"); return 4; }
                compact3_kernel<<<blocks(NA), 256, 0, s>>>(f2, p2, (int)NA, acc, nullptr, nullptr, acc2, nullptr, nullptr);
                if (int rc = check_launch("nhop_compact")) return rc;
                acc = acc2; NA = NA2;
            }
        }
        res->nhop = ar.result<int32_t>(4 * NA); res->e2 = NA;
        if (ar.failed) { set_error("nhop_build: allocation failed"); return 4; }
        if (NA > 0) {
            nhop_rows_kernel<<<blocks(NA), 256, 0, s>>>(acc, (int)NA, c_l1, c_c, l1_b, l1_a, s32, g->ur0, g->ut, res->nhop);
            if (int rc = check_launch("nhop_rows")) return rc;
        }
    } else {
        res->nhop = ar.result<int32_t>(0);
    }
    return ar.failed ? 4 : 0;
}
