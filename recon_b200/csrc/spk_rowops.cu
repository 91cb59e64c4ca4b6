// K6: row-wise wrappers around the attention layers (GAT/models.py:160-161,167-179):
//   rownorm        : y = x / max(||x||_2, 1e-12)                         (F.normalize, in place allowed)
//   residual_norm  : v = EW + mask*x2 ; out = v / max(||v||, 1e-12)      (models.py:175-179), saves 1/||v||
//   residual_norm_bwd : dv = (g - out*(out.g)) / ||v|| ; dEW = dv ; dx2 = mask*dv
//   mask_from_index: mask[idx[b]] = 1                                    (models.py:167-173)
// One warp per row, float4 when the row allows it; HBM-bound streaming.
#include "spk_common.cuh"
#include "spk_rowops.cuh"

namespace spk {
namespace {

__global__ void __launch_bounds__(256)
rownorm_kernel(const float* __restrict__ x, long ldx, float* __restrict__ y, long ldy, long n_rows, int width) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float* xr = x + row * ldx;
    float ss = 0.f;
    for (int c = lane; c < width; c += 32) { const float v = xr[c]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    float* yr = y + row * ldy;
    for (int c = lane; c < width; c += 32) yr[c] = xr[c] * inv;
}

// even width, 8-byte aligned rows, width <= 64 * NV: a lane keeps its NV float2 of a row in registers (one pass over
// memory); a warp handles R consecutive rows with all their loads issued up front (narrow rows need the extra loads in flight)
template <int NV, int R>
__global__ void __launch_bounds__(256)
rownorm_vec2_kernel(const float* __restrict__ x, long ldx, float* __restrict__ y, long ldy, long n_rows, int w2) {
    const int lane = threadIdx.x & 31;
    const long row0 = ((long)blockIdx.x * 8 + (threadIdx.x >> 5)) * R;
    float2 v[R][NV];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float2* xr = reinterpret_cast<const float2*>(x + (row0 + r) * ldx);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = lane + 32 * k;
            v[r][k] = (row0 + r < n_rows && c < w2) ? xr[c] : make_float2(0.f, 0.f);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (row0 + r >= n_rows) break;
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) ss = fmaf(v[r][k].x, v[r][k].x, fmaf(v[r][k].y, v[r][k].y, ss));
        ss = warp_sum(ss);
        const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
        float2* yr = reinterpret_cast<float2*>(y + (row0 + r) * ldy);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = lane + 32 * k;
            if (c < w2) yr[c] = make_float2(v[r][k].x * inv, v[r][k].y * inv);
        }
    }
}

__global__ void __launch_bounds__(256)
residual_norm_kernel(const float* __restrict__ ew, long lde, const float* __restrict__ x2, long ldx,
                     const float* __restrict__ mask, float* __restrict__ out, long ldo,
                     float* __restrict__ inv_norm, long n_rows, int width) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float m = mask[row];
    const float* e = ew + row * lde;
    const float* x = x2 + row * ldx;
    float ss = 0.f;
    for (int c = lane; c < width; c += 32) { const float v = e[c] + m * x[c]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    float* o = out + row * ldo;
    for (int c = lane; c < width; c += 32) o[c] = (e[c] + m * x[c]) * inv;
    if (lane == 0) inv_norm[row] = inv;
}

__global__ void __launch_bounds__(256)
residual_norm_bwd_kernel(const float* __restrict__ g, long ldg, const float* __restrict__ out, long ldo,
                         const float* __restrict__ mask, const float* __restrict__ inv_norm,
                         float* __restrict__ dew, long lde, float* __restrict__ dx2, long ldx,
                         long n_rows, int width) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float* gr = g + row * ldg;
    const float* o = out + row * ldo;
    float dot = 0.f;
    for (int c = lane; c < width; c += 32) dot = fmaf(gr[c], o[c], dot);
    dot = warp_sum(dot);
    const float inv = inv_norm[row];
    const float m = mask[row];
    // ||v|| < eps -> F.normalize divides by the constant eps: dv = g / eps
    const bool clamped = inv >= 1e12f;
    for (int c = lane; c < width; c += 32) {
        const float dv = clamped ? gr[c] * inv : (gr[c] - o[c] * dot) * inv;
        dew[row * lde + c] = dv;
        dx2[row * ldx + c] = m * dv;
    }
}

// negative indices count from the end (the reference indexes a tensor with them); an index outside [-n_rows, n_rows) sets
// bit 1 of *flag (the reference raises IndexError there) instead of costing the caller a min / max read-back per step
__global__ void mask_from_index_kernel(const long long* __restrict__ idx, long n_idx, float* __restrict__ mask, long n_rows,
                                       int* __restrict__ flag) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_idx) return;
    long long r = idx[i];
    if (r < 0) r += n_rows;
    if (r >= 0 && r < n_rows) mask[r] = 1.0f;        // idempotent store: duplicates are harmless
    else if (flag) atomicOr(flag, 2);
}

// <a, b> over n contiguous floats, deterministic: every block sums a fixed contiguous chunk (float4 loads, fp32 per thread,
// fp64 across the block), the last kernel adds the per-block partials in a fixed tree. Serves the linear probe loss
// <out, G> of bench.py / the tests without a library dot.
constexpr int DOT_BLOCKS = 148 * 8;
__global__ void __launch_bounds__(256)
inner_product_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long n, double* __restrict__ partial) {
    const long n4 = n >> 2;
    const long per = (n4 + gridDim.x - 1) / gridDim.x;
    const long lo = (long)blockIdx.x * per, hi = lo + per < n4 ? lo + per : n4;
    float acc0 = 0.f, acc1 = 0.f;
    long i = lo + threadIdx.x;
    for (; i + 256 < hi; i += 512) {
        const float4 x0 = ldg4_stream(a + 4 * i), y0 = ldg4_stream(b + 4 * i), x1 = ldg4_stream(a + 4 * (i + 256)), y1 = ldg4_stream(b + 4 * (i + 256));
        acc0 = fmaf(x0.x, y0.x, fmaf(x0.y, y0.y, fmaf(x0.z, y0.z, fmaf(x0.w, y0.w, acc0))));
        acc1 = fmaf(x1.x, y1.x, fmaf(x1.y, y1.y, fmaf(x1.z, y1.z, fmaf(x1.w, y1.w, acc1))));
    }
    for (; i < hi; i += 256) {
        const float4 x0 = ldg4_stream(a + 4 * i), y0 = ldg4_stream(b + 4 * i);
        acc0 = fmaf(x0.x, y0.x, fmaf(x0.y, y0.y, fmaf(x0.z, y0.z, fmaf(x0.w, y0.w, acc0))));
    }
    if (blockIdx.x == gridDim.x - 1)                                   // scalar tail (n % 4 elements)
        for (long t = (n4 << 2) + threadIdx.x; t < n; t += 256) acc1 = fmaf(a[t], b[t], acc1);
    __shared__ double sh[256];
    sh[threadIdx.x] = (double)acc0 + (double)acc1;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256)
inner_product_final_kernel(const double* __restrict__ partial, int n_part, float* __restrict__ out, int accumulate) {
    __shared__ double sh[256];
    double v = 0.0;
    for (int i = threadIdx.x; i < n_part; i += 256) v += partial[i];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + (float)sh[0];
}
}  // namespace

long inner_product_workspace_bytes() { return (long)DOT_BLOCKS * sizeof(double); }
int inner_product(const float* a, const float* b, long n, void* workspace, float* out, int accumulate, cudaStream_t s) {
    if (n < 0 || (n > 0 && (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) != 0))) {
        set_error("inner_product: operands must be 16-byte aligned");
        return 1;
    }
    int blocks = (int)((n / 4 + 2047) / 2048);
    if (blocks < 1) blocks = 1;
    if (blocks > DOT_BLOCKS) blocks = DOT_BLOCKS;
    inner_product_partial_kernel<<<blocks, 256, 0, s>>>(a, b, n, reinterpret_cast<double*>(workspace));
    if (int rc = check_launch("inner_product")) return rc;
    inner_product_final_kernel<<<1, 256, 0, s>>>(reinterpret_cast<const double*>(workspace), blocks, out, accumulate);
    return check_launch("inner_product_final");
}

int rownorm(const float* x, long ldx, float* y, long ldy, long n_rows, int width, cudaStream_t s) {
    if (n_rows <= 0) return 0;
    const bool vec2 = (width % 2 == 0) && width <= 256 && (ldx % 2 == 0) && (ldy % 2 == 0) &&
                      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 7) == 0;
    if (vec2 && width <= 64) rownorm_vec2_kernel<1, 4><<<(unsigned)((n_rows + 31) / 32), 256, 0, s>>>(x, ldx, y, ldy, n_rows, width / 2);
    else if (vec2) rownorm_vec2_kernel<4, 1><<<(unsigned)((n_rows + 7) / 8), 256, 0, s>>>(x, ldx, y, ldy, n_rows, width / 2);
    else rownorm_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, s>>>(x, ldx, y, ldy, n_rows, width);
    return check_launch("rownorm");
}
int residual_norm(const float* ew, long lde, const float* x2, long ldx, const float* mask, float* out, long ldo,
                  float* inv_norm, long n_rows, int width, cudaStream_t s) {
    if (n_rows <= 0) return 0;
    residual_norm_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, s>>>(ew, lde, x2, ldx, mask, out, ldo, inv_norm, n_rows, width);
    return check_launch("residual_norm");
}
int residual_norm_bwd(const float* g, long ldg, const float* out, long ldo, const float* mask, const float* inv_norm,
                      float* dew, long lde, float* dx2, long ldx, long n_rows, int width, cudaStream_t s) {
    if (n_rows <= 0) return 0;
    residual_norm_bwd_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, s>>>(g, ldg, out, ldo, mask, inv_norm, dew, lde, dx2, ldx, n_rows, width);
    return check_launch("residual_norm_bwd");
}
int mask_from_index(const long long* idx, long n_idx, float* mask, long n_rows, int* flag, cudaStream_t s) {
    if (n_idx <= 0) return 0;
    mask_from_index_kernel<<<(unsigned)((n_idx + 255) / 256), 256, 0, s>>>(idx, n_idx, mask, n_rows, flag);
    return check_launch("mask_from_index");
}
}  // namespace spk
