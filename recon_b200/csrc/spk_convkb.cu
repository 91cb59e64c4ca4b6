// N4: the ConvKB scoring stage that consumes the hot path's output embeddings.
//   reference: ConvKB.forward            GAT/layers.py:31-48   (live path: fc2(LeakyReLU(fc1([h | r | t]))))
//              SpKBGATConvOnly.forward   GAT/models.py:294-304 (gather + concatenate the three embedding rows)
//              relation ranking          GAT/create_batch.py:1367-1393 (every test pair scored under every relation)
//              tanh(e . W_ent2rel[r])    GAT_sep_space/models.py:316-320
// The dense products run on the library GEMMs (tcgen05, spk_gemm_tc.cu); this file holds the row gathers, the fused
// bias + LeakyReLU + fc2 head (forward and backward), tanh, and the streaming all-relations ranking pass that uses the
// re-association fc1([h|r|t]) = A[h] + B[r] + C[t]. All HBM / L2 streaming work: one warp per row, 128-bit accesses.
#include "../../include/spkbgat.h"
#include "spk_common.cuh"

namespace spk {
namespace {

// out[b, p*D : (p+1)*D] = src_p[row_p(b), :D],  row_p(b) = idx_p ? idx_p[b * stride_p] : b   (p = 0..np-1)
struct ConcatArgs {
    const float* src[3]; long ld[3]; const long long* idx[3]; long stride[3]; long rows[3];
    int np;
};

__global__ void __launch_bounds__(256)
gather_concat_kernel(ConcatArgs a, long B, int D, float* __restrict__ out, long ldo, int* __restrict__ err) {
    const int lane = threadIdx.x & 31;
    const long b = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    for (int p = 0; p < a.np; ++p) {
        long r = a.idx[p] ? (long)a.idx[p][b * a.stride[p]] : b;
        if (r < 0 || r >= a.rows[p]) { if (lane == 0) atomicOr(err, 1); r = 0; }
        const float* s = a.src[p] + r * a.ld[p];
        float* o = out + b * ldo + (long)p * D;
        for (int c = lane; c < D; c += 32) o[c] = __ldg(s + c);
    }
}

// out[b] = b2 + sum_d w2[d] * lrelu(H1[b, d] + b1[d])
__global__ void __launch_bounds__(256)
mlp_head_fwd_kernel(const float* __restrict__ H1, long ldh, const float* __restrict__ b1, const float* __restrict__ w2,
                    const float* __restrict__ b2, float slope, long B, int D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long b = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) {
        const float x = H1[b * ldh + c] + __ldg(b1 + c);
        s = fmaf(__ldg(w2 + c), x > 0.f ? x : slope * x, s);
    }
    s = warp_sum(s);
    if (lane == 0) out[b] = s + __ldg(b2);
}

// dH1[b, d] = dout[b] * w2[d] * lrelu'(pre),  act[b, d] = lrelu(pre),  pre = H1[b, d] + b1[d]
__global__ void __launch_bounds__(256)
mlp_head_bwd_kernel(const float* __restrict__ H1, long ldh, const float* __restrict__ b1, const float* __restrict__ w2,
                    float slope, const float* __restrict__ dout, long B, int D, float* __restrict__ dH1, long ldd,
                    float* __restrict__ act, long lda) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * D) return;
    const long b = i / D;
    const int c = (int)(i % D);
    const float x = H1[b * ldh + c] + __ldg(b1 + c);
    const float g = __ldg(dout + b) * __ldg(w2 + c);
    dH1[b * ldd + c] = x > 0.f ? g : slope * g;
    act[b * lda + c] = x > 0.f ? x : slope * x;
}

__global__ void __launch_bounds__(256)
tanh_fwd_kernel(float* __restrict__ x, long ld, long n, int w) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * w) return;
    float* p = x + (i / w) * ld + (i % w);
    *p = tanhf(*p);
}

// dpre = dout * (1 - y^2)
__global__ void __launch_bounds__(256)
tanh_bwd_kernel(const float* __restrict__ y, long ldy, const float* __restrict__ dout, long ldo, long n, int w,
                float* __restrict__ dpre, long ldp) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * w) return;
    const long r = i / w;
    const int c = (int)(i % w);
    const float v = y[r * ldy + c];
    dpre[r * ldp + c] = dout[r * ldo + c] * (1.f - v * v);
}

// scores[i, r] = b2 + sum_d w2[d] * lrelu(U[i, d] + Bt[r, d])      (U = A[h] + C[t] + b1 per test pair, Bt = Rel . W1b^T)
// A CTA of 8 warps takes 8 pairs; a warp keeps its pair's U row and w2 in registers (D <= 32 * RK_REG) and streams all
// relation rows (the same rows for every warp: L1 / L2 hits), one warp-wide reduction per score.
constexpr int RK_REG = 16;
__global__ void __launch_bounds__(256)
rank_scores_kernel(const float* __restrict__ U, long ldu, const float* __restrict__ Bt, long ldb, const float* __restrict__ w2,
                   const float* __restrict__ b2, float slope, long T, int R, int D, float* __restrict__ out, long ldo) {
    const int lane = threadIdx.x & 31;
    const long i = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= T) return;
    float u[RK_REG], w[RK_REG];
#pragma unroll
    for (int k = 0; k < RK_REG; ++k) {
        const int c = lane + 32 * k;
        u[k] = c < D ? U[i * ldu + c] : 0.f;
        w[k] = c < D ? __ldg(w2 + c) : 0.f;
    }
    const float bias = __ldg(b2);
    for (int r = 0; r < R; ++r) {
        const float* br = Bt + (long)r * ldb;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < RK_REG; ++k) {
            const int c = lane + 32 * k;
            if (c < D) {
                const float x = u[k] + __ldg(br + c);
                s = fmaf(w[k], x > 0.f ? x : slope * x, s);
            }
        }
        s = warp_sum(s);
        if (lane == 0) out[i * ldo + r] = s + bias;
    }
}

}  // namespace

int launch_gather_concat(const float* const* src, const long* ld, const long long* const* idx, const long* stride,
                         const long* rows, int np, long B, int D, float* out, long ldo, int* err, cudaStream_t s) {
    if (B <= 0 || D <= 0) return 0;
    ConcatArgs a;
    a.np = np;
    for (int p = 0; p < 3; ++p) {
        a.src[p] = p < np ? src[p] : nullptr; a.ld[p] = p < np ? ld[p] : 0; a.idx[p] = p < np ? idx[p] : nullptr;
        a.stride[p] = p < np ? stride[p] : 0; a.rows[p] = p < np ? rows[p] : 0;
    }
    gather_concat_kernel<<<(unsigned)((B + 7) / 8), 256, 0, s>>>(a, B, D, out, ldo, err);
    return check_launch("gather_concat");
}

int launch_mlp_head_fwd(const float* H1, long ldh, const float* b1, const float* w2, const float* b2, float slope, long B,
                        int D, float* out, cudaStream_t s) {
    if (B <= 0) return 0;
    mlp_head_fwd_kernel<<<(unsigned)((B + 7) / 8), 256, 0, s>>>(H1, ldh, b1, w2, b2, slope, B, D, out);
    return check_launch("mlp_head_fwd");
}

int launch_mlp_head_bwd(const float* H1, long ldh, const float* b1, const float* w2, float slope, const float* dout, long B,
                        int D, float* dH1, long ldd, float* act, long lda, cudaStream_t s) {
    if (B <= 0 || D <= 0) return 0;
    mlp_head_bwd_kernel<<<(unsigned)((B * D + 255) / 256), 256, 0, s>>>(H1, ldh, b1, w2, slope, dout, B, D, dH1, ldd, act, lda);
    return check_launch("mlp_head_bwd");
}

int launch_tanh_fwd(float* x, long ld, long n, int w, cudaStream_t s) {
    if (n <= 0 || w <= 0) return 0;
    tanh_fwd_kernel<<<(unsigned)((n * w + 255) / 256), 256, 0, s>>>(x, ld, n, w);
    return check_launch("tanh_fwd");
}

int launch_tanh_bwd(const float* y, long ldy, const float* dout, long ldo, long n, int w, float* dpre, long ldp, cudaStream_t s) {
    if (n <= 0 || w <= 0) return 0;
    tanh_bwd_kernel<<<(unsigned)((n * w + 255) / 256), 256, 0, s>>>(y, ldy, dout, ldo, n, w, dpre, ldp);
    return check_launch("tanh_bwd");
}

int launch_rank_scores(const float* U, long ldu, const float* Bt, long ldb, const float* w2, const float* b2, float slope,
                       long T, int R, int D, float* out, long ldo, cudaStream_t s) {
    if (T <= 0 || R <= 0) return 0;
    if (D > 32 * RK_REG) { set_error("rank_scores: width %d exceeds %d", D, 32 * RK_REG); return 2; }
    rank_scores_kernel<<<(unsigned)((T + 7) / 8), 256, 0, s>>>(U, ldu, Bt, ldb, w2, b2, slope, T, R, D, out, ldo);
    return check_launch("rank_scores");
}

}  // namespace spk

extern "C" {

int spk_gather_concat(const float* const* src, const int64_t* ld, const int64_t* const* idx, const int64_t* stride,
                      const int64_t* rows, int32_t n_pieces, int64_t n_out, int32_t D, float* out, int64_t ldo,
                      int32_t* err_flag, spk_stream_t stream) {
    if (n_pieces < 1 || n_pieces > 3 || D < 1 || ldo < (int64_t)n_pieces * D || !err_flag) {
        spk::set_error("gather_concat: bad arguments");
        return 1;
    }
    const float* s[3]; long l[3]; const long long* ix[3]; long st[3]; long rw[3];
    for (int p = 0; p < n_pieces; ++p) {
        if (!src[p] || ld[p] < D || rows[p] < 1) { spk::set_error("gather_concat: bad piece %d", p); return 1; }
        s[p] = src[p]; l[p] = (long)ld[p]; ix[p] = reinterpret_cast<const long long*>(idx[p]); st[p] = (long)stride[p]; rw[p] = (long)rows[p];
    }
    return spk::launch_gather_concat(s, l, ix, st, rw, n_pieces, (long)n_out, D, out, (long)ldo, err_flag, (cudaStream_t)stream);
}

int spk_mlp_head_fwd(const float* H1, int64_t ldh, const float* b1, const float* w2, const float* b2, float slope,
                     int64_t n_rows, int32_t D, float* out, spk_stream_t stream) {
    if (D < 1 || ldh < D) { spk::set_error("mlp_head_fwd: bad shape"); return 1; }
    return spk::launch_mlp_head_fwd(H1, (long)ldh, b1, w2, b2, slope, (long)n_rows, D, out, (cudaStream_t)stream);
}

int spk_mlp_head_bwd(const float* H1, int64_t ldh, const float* b1, const float* w2, float slope, const float* dout,
                     int64_t n_rows, int32_t D, float* dH1, int64_t ldd, float* act, int64_t lda, spk_stream_t stream) {
    if (D < 1 || ldh < D || ldd < D || lda < D) { spk::set_error("mlp_head_bwd: bad shape"); return 1; }
    return spk::launch_mlp_head_bwd(H1, (long)ldh, b1, w2, slope, dout, (long)n_rows, D, dH1, (long)ldd, act, (long)lda,
                                    (cudaStream_t)stream);
}

int spk_tanh_fwd(float* x, int64_t ld, int64_t n_rows, int32_t width, spk_stream_t stream) {
    if (ld < width) { spk::set_error("tanh_fwd: bad shape"); return 1; }
    return spk::launch_tanh_fwd(x, (long)ld, (long)n_rows, width, (cudaStream_t)stream);
}

int spk_tanh_bwd(const float* y, int64_t ldy, const float* dout, int64_t ldo, int64_t n_rows, int32_t width, float* dpre,
                 int64_t ldp, spk_stream_t stream) {
    if (ldy < width || ldo < width || ldp < width) { spk::set_error("tanh_bwd: bad shape"); return 1; }
    return spk::launch_tanh_bwd(y, (long)ldy, dout, (long)ldo, (long)n_rows, width, dpre, (long)ldp, (cudaStream_t)stream);
}

int spk_rank_scores(const float* U, int64_t ldu, const float* Bt, int64_t ldb, const float* w2, const float* b2, float slope,
                    int64_t n_pairs, int32_t n_rel, int32_t D, float* out, int64_t ldo, spk_stream_t stream) {
    if (D < 1 || ldu < D || ldb < D || ldo < n_rel) { spk::set_error("rank_scores: bad shape"); return 1; }
    return spk::launch_rank_scores(U, (long)ldu, Bt, (long)ldb, w2, b2, slope, (long)n_pairs, n_rel, D, out, (long)ldo,
                                   (cudaStream_t)stream);
}

}  // extern "C"
