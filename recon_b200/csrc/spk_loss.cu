// N1 (the step right after the hot path in every training iteration):
//   * batch_gat_loss (GAT/main.py:344-376): TransE L1 distance of every positive / corrupted triple on the hot path's
//     outputs, nn.MarginRankingLoss(margin) with y = -1 (main.py:367-372, 451), mean over the 2*ratio*P pairs,
//   * its backward, which produces the gradient of out_entity / out_relation that SpKBGATModified's backward consumes
//     (replaces the index_put_(accumulate=True) atomics of autograd's IndexBackward: here every entity / relation row
//     sums its incidences in a fixed order over a radix-sorted incidence list, so the result is run-to-run identical),
//   * the SGD step on the parameter tensors (main.py:445-446, 524).
//
// Forward, one warp per triple (HBM-bound gathers of three `width`-float rows):
//     x = (ent[h] + rel[r]) - ent[t];  norm = sum |x|;  sgn = sign(x) packed as int8 (the only thing backward needs
//     from x: d|x|/dx, torch.norm(p=1) uses sign with sign(0) = 0)
// Pairs, one thread per positive p (pos_triples.repeat(2*ratio, 1) pairs negative k with positive k mod P):
//     v = (pos[p] - neg[k]) + margin;  l = clamp_min(v, 0);  active = v >= 0 (clamp_min's backward passes the tie)
//     coef[neg k] = -active / M,  coef[pos p] = sum_j active / M          (dl/dx of each triple = coef * sgn)
// Backward, one warp per entity (relation) segment of the sorted incidence list, hub segments as tasks + finalize:
//     d ent[i] = g * sum_{inc} (+-) coef[t] * sgn[t]   (+ for heads and relations, - for tails)
#include "../../include/spkbgat.h"
#include "spk_common.cuh"

namespace spk {
namespace {

constexpr int LOSS_MAX_CH = 4;            // a lane owns up to 4 chunks of 4 columns: width <= 512
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int sgn_i(float v) { return (v > 0.f) - (v < 0.f); }
__device__ __forceinline__ unsigned pack_sgn(float4 x) {
    return (unsigned)(sgn_i(x.x) & 0xff) | ((unsigned)(sgn_i(x.y) & 0xff) << 8) |
           ((unsigned)(sgn_i(x.z) & 0xff) << 16) | ((unsigned)(sgn_i(x.w) & 0xff) << 24);
}
__device__ __forceinline__ float4 unpack_sgn(unsigned u) {
    return make_float4((float)(signed char)(u & 0xffu), (float)(signed char)((u >> 8) & 0xffu),
                       (float)(signed char)((u >> 16) & 0xffu), (float)(signed char)((u >> 24) & 0xffu));
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
// chunk c (columns 4c .. 4c+3) of a row of `width` floats; columns past the end read as 0
__device__ __forceinline__ float4 load_chunk(const float* __restrict__ row, int c, int width, int vec) {
    if (vec) return ldg4(row + 4 * c);
    const int j = 4 * c;
    return make_float4(j < width ? __ldg(row + j) : 0.f, j + 1 < width ? __ldg(row + j + 1) : 0.f,
                       j + 2 < width ? __ldg(row + j + 2) : 0.f, j + 3 < width ? __ldg(row + j + 3) : 0.f);
}
__device__ __forceinline__ void store_chunk(float* __restrict__ row, int c, int width, int vec, float4 v) {
    if (vec) { *reinterpret_cast<float4*>(row + 4 * c) = v; return; }
    const int j = 4 * c;
    if (j < width) row[j] = v.x;
    if (j + 1 < width) row[j + 1] = v.y;
    if (j + 2 < width) row[j + 2] = v.z;
    if (j + 3 < width) row[j + 3] = v.w;
}

// ---- incidence lists: entity keys (head, tail of every triple) and relation keys --------------------------------
__global__ void __launch_bounds__(256)
triple_incidence_kernel(const long long* __restrict__ tri, long T, long n_ent, long n_rel,
                        int* __restrict__ ent_keys, int* __restrict__ ent_vals,
                        int* __restrict__ rel_keys, int* __restrict__ rel_vals, int* __restrict__ err) {
    const long t = (long)blockIdx.x * 256 + threadIdx.x;
    if (t >= T) return;
    long long h = tri[3 * t], r = tri[3 * t + 1], tl = tri[3 * t + 2];
    if (h < 0 || h >= n_ent || tl < 0 || tl >= n_ent || r < 0 || r >= n_rel) { *err = 1; h = 0; r = 0; tl = 0; }
    ent_keys[2 * t] = (int)h;      ent_vals[2 * t] = (int)(2 * t);
    ent_keys[2 * t + 1] = (int)tl; ent_vals[2 * t + 1] = (int)(2 * t + 1);
    rel_keys[t] = (int)r;          rel_vals[t] = (int)t;
}

// ---- forward: L1 norm and sign pattern of every triple ---------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(256)
loss_norm_kernel(const long long* __restrict__ tri, long T, const float* __restrict__ ent, long lde, long n_ent,
                 const float* __restrict__ rel, long ldr, long n_rel, int width, int W4, int vec,
                 float* __restrict__ norm, unsigned* __restrict__ sgn, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const long t = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (t >= T) return;
    const long long h = tri[3 * t], r = tri[3 * t + 1], tl = tri[3 * t + 2];
    if (h < 0 || h >= n_ent || tl < 0 || tl >= n_ent || r < 0 || r >= n_rel) {
        if (lane == 0) { *err = 1; norm[t] = 0.f; }
#pragma unroll
        for (int k = 0; k < NCH; ++k) { const int c = lane + 32 * k; if (c < W4) sgn[t * W4 + c] = 0u; }
        return;
    }
    const float* ph = ent + h * lde;
    const float* pr = rel + r * ldr;
    const float* pt = ent + tl * lde;
    float4 a[NCH], b[NCH], d[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c < W4) { a[k] = load_chunk(ph, c, width, vec); b[k] = load_chunk(pr, c, width, vec); d[k] = load_chunk(pt, c, width, vec); }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c < W4) {
            // (source + relation) - tail, main.py:357,364
            const float4 x = make_float4((a[k].x + b[k].x) - d[k].x, (a[k].y + b[k].y) - d[k].y,
                                         (a[k].z + b[k].z) - d[k].z, (a[k].w + b[k].w) - d[k].w);
            s += (fabsf(x.x) + fabsf(x.y)) + (fabsf(x.z) + fabsf(x.w));
            sgn[t * W4 + c] = pack_sgn(x);
        }
    }
    s = warp_sum(s);
    if (lane == 0) norm[t] = s;
}

// ---- forward: hinge over the pairs, per-triple coefficients, per-CTA partial sums (double, fixed order) ---------------
__global__ void __launch_bounds__(256)
loss_pairs_kernel(const float* __restrict__ norm, long P, int reps, float margin, float inv_m,
                  float* __restrict__ coef, double* __restrict__ partial) {
    __shared__ double sm[8];
    const long p = (long)blockIdx.x * 256 + threadIdx.x;
    double ls = 0.0;
    if (p < P) {
        const float pn = norm[p];
        float cp = 0.f;
        for (int j = 0; j < reps; ++j) {
            const long k = p + (long)j * P;
            const float v = (pn - norm[P + k]) + margin;      // -y (x1 - x2) + margin with y = -1
            const float l = (v < 0.f) ? 0.f : v;              // clamp_min(0); NaN stays NaN
            const float c = (v >= 0.f) ? inv_m : 0.f;
            ls += (double)l;
            coef[P + k] = -c;
            cp += c;
        }
        coef[p] = cp;
    }
    ls = warp_sum_d(ls);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = ls;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sm[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(1024)
loss_final_kernel(const double* __restrict__ partial, long n, double scale, float* __restrict__ loss) {
    __shared__ double sm[32];
    double s = 0.0;
    for (long i = threadIdx.x; i < n; i += 1024) s += partial[i];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = sm[threadIdx.x];
        v = warp_sum_d(v);
        if (threadIdx.x == 0) *loss = (float)(v * scale);
    }
}

// ---- backward ------------------------------------------------------------------------------------------------------------
struct LossBwdP {
    const int* segptr; const int* inc; const float* coef; const unsigned* sgn; const float* gscale;
    float* out; long ldo; int n_seg; int width; int W4; int mode; int vec_out;
};

// acc += sum over the incidences [beg, end) (in list order) of (+-) coef[t] * sgn[t]
template <int NCH>
__device__ __forceinline__ void loss_accum(const LossBwdP& p, int beg, int end, int lane, float4 (&acc)[NCH]) {
    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int tv = 0;
        float cv = 0.f;
        if (lane < n) {
            const int v = __ldg(p.inc + base + lane);
            if (p.mode == 0) { tv = v >> 1; const float c = __ldg(p.coef + tv); cv = (v & 1) ? -c : c; }
            else { tv = v; cv = __ldg(p.coef + tv); }
        }
        for (int u = 0; u < n; u += 4) {
            unsigned w[4][NCH];
            float c[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {              // lanes >= n carry (t = 0, coef = 0): a harmless + 0
                const int tq = __shfl_sync(FULL, tv, u + q);
                c[q] = __shfl_sync(FULL, cv, u + q);
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    const int ci = lane + 32 * k;
                    w[q][k] = ci < p.W4 ? __ldg(p.sgn + (long)tq * p.W4 + ci) : 0u;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int k = 0; k < NCH; ++k) f4fma(acc[k], c[q], unpack_sgn(w[q][k]));
        }
    }
}

template <int NCH>
__global__ void __launch_bounds__(256)
loss_bwd_seg_kernel(LossBwdP p, int hub_thresh) {
    const int lane = threadIdx.x & 31;
    const int seg = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (seg >= p.n_seg) return;
    const int beg = p.segptr[seg], end = p.segptr[seg + 1];
    if (end - beg > hub_thresh) return;                 // written by the task + finalize kernels
    float4 acc[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    loss_accum<NCH>(p, beg, end, lane, acc);
    const float g = __ldg(p.gscale);
    float* o = p.out + (long)seg * p.ldo;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c < p.W4) store_chunk(o, c, p.width, p.vec_out, make_float4(g * acc[k].x, g * acc[k].y, g * acc[k].z, g * acc[k].w));
    }
}

template <int NCH>
__global__ void __launch_bounds__(256)
loss_bwd_task_kernel(LossBwdP p, const int* __restrict__ task_beg, const int* __restrict__ task_end, int n_tasks,
                     float* __restrict__ partial, long ldpart) {
    const int lane = threadIdx.x & 31;
    const int task = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (task >= n_tasks) return;
    float4 acc[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    loss_accum<NCH>(p, task_beg[task], task_end[task], lane, acc);
    float* o = partial + (long)task * ldpart;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c = lane + 32 * k;
        if (c < p.W4) *reinterpret_cast<float4*>(o + 4 * c) = acc[k];
    }
}

// one CTA per hub segment: column-parallel sum of its task partials in task order
__global__ void __launch_bounds__(256)
loss_bwd_finalize_kernel(LossBwdP p, const int* __restrict__ hub_seg, const int* __restrict__ hub_task_ptr,
                         const float* __restrict__ partial, long ldpart) {
    const int seg = hub_seg[blockIdx.x];
    const int t0 = hub_task_ptr[blockIdx.x], t1 = hub_task_ptr[blockIdx.x + 1];
    const float g = __ldg(p.gscale);
    for (int c = threadIdx.x; c < p.width; c += 256) {
        float s = 0.f;
        int t = t0;
        for (; t + 8 <= t1; t += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = partial[(long)(t + u) * ldpart + c];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; t < t1; ++t) s += partial[(long)t * ldpart + c];
        p.out[(long)seg * p.ldo + c] = g * s;
    }
}

// ---- SGD: p <- p - lr * g on up to 16 tensors in one launch (torch.optim.SGD without momentum / weight decay) -------------
struct SgdP { float* p[16]; const float* g[16]; long n[16]; float lr; };
__global__ void __launch_bounds__(256) sgd_kernel(SgdP a) {
    float* __restrict__ p = a.p[blockIdx.y];
    const float* __restrict__ g = a.g[blockIdx.y];
    const long n = a.n[blockIdx.y];
    const float nlr = -a.lr;
    for (long j = (long)blockIdx.x * 256 + threadIdx.x; j < n; j += (long)gridDim.x * 256) p[j] = fmaf(nlr, g[j], p[j]);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace spk

using namespace spk;

extern "C" {

int spk_triple_incidence(const int64_t* triples, int64_t n_triples, int64_t n_ent, int64_t n_rel,
                         int32_t* ent_keys, int32_t* ent_vals, int32_t* rel_keys, int32_t* rel_vals,
                         int32_t* err_flag, spk_stream_t stream) {
    if (n_triples <= 0) return 0;
    if (n_triples >= (1LL << 30) || n_ent >= (1LL << 31) || n_rel >= (1LL << 31)) {
        set_error("triple_incidence: sizes exceed the int32 incidence encoding (T=%lld)", (long long)n_triples);
        return 2;
    }
    triple_incidence_kernel<<<(unsigned)((n_triples + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)triples, n_triples, n_ent, n_rel, ent_keys, ent_vals, rel_keys, rel_vals, err_flag);
    return check_launch("triple_incidence");
}

int64_t spk_margin_loss_partials(int64_t n_pos) { return n_pos > 0 ? (n_pos + 255) / 256 : 1; }

int spk_margin_loss_fwd(const int64_t* triples, int64_t n_triples, int64_t n_pos,
                        const float* ent, int64_t lde, int64_t n_ent, const float* rel, int64_t ldr, int64_t n_rel,
                        int32_t width, float margin, int32_t mean,
                        float* norm, uint32_t* sgn, float* coef, double* partial, float* loss,
                        int32_t* err_flag, spk_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (n_pos <= 0 || n_triples <= n_pos || (n_triples - n_pos) % n_pos != 0) {
        set_error("margin_loss_fwd: %lld triples are not P positives followed by a multiple of P negatives (P=%lld)",
                  (long long)n_triples, (long long)n_pos);
        return 2;
    }
    if (width < 1 || width > 128 * LOSS_MAX_CH || n_triples >= (1LL << 30)) {
        set_error("margin_loss_fwd: width %d (1..512) or triple count %lld out of range", width, (long long)n_triples);
        return 2;
    }
    const int W4 = (width + 3) / 4;
    const int vec = (width % 4 == 0) && (lde % 4 == 0) && (ldr % 4 == 0) && aligned16(ent) && aligned16(rel);
    const unsigned grid = (unsigned)((n_triples + 7) / 8);
#define SPK_LOSS_NORM(NCH)                                                                                              \
    loss_norm_kernel<NCH><<<grid, 256, 0, s>>>((const long long*)triples, n_triples, ent, lde, n_ent, rel, ldr, n_rel, \
                                               width, W4, vec, norm, sgn, err_flag)
    if (W4 <= 32) SPK_LOSS_NORM(1);
    else if (W4 <= 64) SPK_LOSS_NORM(2);
    else SPK_LOSS_NORM(4);
#undef SPK_LOSS_NORM
    int rc = check_launch("loss_norm");
    if (rc) return rc;
    const int64_t m = n_triples - n_pos;
    const int reps = (int)(m / n_pos);
    const float inv_m = mean ? (float)(1.0 / (double)m) : 1.0f;
    const int64_t nblk = spk_margin_loss_partials(n_pos);
    loss_pairs_kernel<<<(unsigned)nblk, 256, 0, s>>>(norm, n_pos, reps, margin, inv_m, coef, partial);
    rc = check_launch("loss_pairs");
    if (rc) return rc;
    loss_final_kernel<<<1, 1024, 0, s>>>(partial, nblk, mean ? 1.0 / (double)m : 1.0, loss);
    return check_launch("loss_final");
}

int spk_margin_loss_bwd(const spk_loss_bwd_args* a, spk_stream_t stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (a->n_seg <= 0) return 0;
    if (a->width < 1 || a->width > 128 * LOSS_MAX_CH || (a->mode != 0 && a->mode != 1)) {
        set_error("margin_loss_bwd: bad width %d / mode %d", a->width, a->mode);
        return 2;
    }
    LossBwdP p;
    p.segptr = a->segptr; p.inc = a->inc; p.coef = a->coef; p.sgn = a->sgn; p.gscale = a->gscale;
    p.out = a->out; p.ldo = a->ldo; p.n_seg = a->n_seg; p.width = a->width; p.W4 = (a->width + 3) / 4; p.mode = a->mode;
    p.vec_out = (a->width % 4 == 0) && (a->ldo % 4 == 0) && aligned16(a->out);
    const spk_hub_tasks& h = a->hub;
    const int hub_thresh = h.n_tasks > 0 ? h.hub_thresh : 0x7fffffff;
    if (h.n_tasks > 0 && (h.partial == nullptr || h.ldpart < 4 * p.W4 || (h.ldpart & 3) || !aligned16(h.partial))) {
        set_error("margin_loss_bwd: hub partial buffer missing or narrower than %d floats", 4 * p.W4);
        return 2;
    }
    const unsigned grid = (unsigned)((a->n_seg + 7) / 8);
    int rc;
#define SPK_LOSS_BWD(NCH)                                                                                               \
    do {                                                                                                                \
        loss_bwd_seg_kernel<NCH><<<grid, 256, 0, s>>>(p, hub_thresh);                                                   \
        rc = check_launch("loss_bwd_seg");                                                                              \
        if (rc == 0 && h.n_tasks > 0) {                                                                                 \
            loss_bwd_task_kernel<NCH><<<(unsigned)((h.n_tasks + 7) / 8), 256, 0, s>>>(p, h.task_beg, h.task_end,        \
                                                                                      h.n_tasks, h.partial, h.ldpart);  \
            rc = check_launch("loss_bwd_task");                                                                         \
        }                                                                                                               \
    } while (0)
    if (p.W4 <= 32) SPK_LOSS_BWD(1);
    else if (p.W4 <= 64) SPK_LOSS_BWD(2);
    else SPK_LOSS_BWD(4);
#undef SPK_LOSS_BWD
    if (rc) return rc;
    if (h.n_tasks > 0 && h.n_hubs > 0) {
        loss_bwd_finalize_kernel<<<(unsigned)h.n_hubs, 256, 0, s>>>(p, h.hub_seg, h.hub_task_ptr, h.partial, h.ldpart);
        rc = check_launch("loss_bwd_finalize");
    }
    return rc;
}

int spk_sgd_step(const spk_sgd_args* a, spk_stream_t stream) {
    if (a->count < 0 || a->count > 16) { set_error("sgd_step: %d tensors (max 16 per call)", a->count); return 2; }
    SgdP p;
    long nmax = 0;
    int cnt = 0;
    for (int i = 0; i < a->count; ++i) {
        if (a->numel[i] <= 0) continue;
        p.p[cnt] = a->param[i]; p.g[cnt] = a->grad[i]; p.n[cnt] = (long)a->numel[i];
        nmax = a->numel[i] > nmax ? (long)a->numel[i] : nmax;
        ++cnt;
    }
    if (cnt == 0) return 0;
    p.lr = a->lr;
    long gx = (nmax + 255) / 256;
    if (gx > 148 * 16) gx = 148 * 16;
    sgd_kernel<<<dim3((unsigned)gx, (unsigned)cnt), 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("sgd_step");
}

}  // extern "C"
