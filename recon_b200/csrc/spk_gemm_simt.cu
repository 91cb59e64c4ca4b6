// K1 / K5 (first implementation): exact-fp32 SIMT GEMMs for the projection and gradient products.
//   nn : C[M,N] (+)= A[M,K] * B[K,N]           (projections P~ = X * W~, dX = dP~ * W~^T with W~^T prebuilt)
//   tn : C[Ka,Nb] (+)= sum_m A[m,Ka] * B[m,Nb]  (weight gradients dW~ = X^T * dP~; split over m with
//                                               per-split partials added in split order -> deterministic)
// Row-major, arbitrary leading dimensions. These replace the reference's edge-sized `a.mm(edge_h)`
// (GAT/layers.py:137) and its autograd after the re-association of SURVEY.md 8 a-4.
// The tcgen05 3xTF32 path (spk_gemm_tc.cu) takes over for the large shapes when enabled.
#include "spk_common.cuh"
#include "spk_gemm.cuh"

namespace spk {

namespace {
constexpr int BM = 128, BN = 128, BK = 8;

__global__ void __launch_bounds__(256)
sgemm_nn_kernel(const float* __restrict__ A, long lda, const float* __restrict__ B, long ldb,
                float* __restrict__ C, long ldc, long M, int N, int K, int accumulate, int a_vec, int b_vec) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long m0 = (long)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;

    const int a_row = tid >> 1, a_k = (tid & 1) * 4;       // A tile: 128 rows x 8 k
    const int b_k = tid >> 5, b_n = (tid & 31) * 4;        // B tile: 8 k x 128 cols
    float ra[4], rb[4];

    auto load_tiles = [&](int k0) {
        const long gm = m0 + a_row;
        const int gk = k0 + a_k;
        if (gm < M && a_vec && gk + 3 < K) {
            const float4 t = ldg4(A + gm * lda + gk);
            ra[0] = t.x; ra[1] = t.y; ra[2] = t.z; ra[3] = t.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) ra[i] = (gm < M && gk + i < K) ? __ldg(A + gm * lda + gk + i) : 0.f;
        }
        const int bk = k0 + b_k, bn = n0 + b_n;
        if (bk < K && b_vec && bn + 3 < N) {
            const float4 t = ldg4(B + (long)bk * ldb + bn);
            rb[0] = t.x; rb[1] = t.y; rb[2] = t.z; rb[3] = t.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) rb[i] = (bk < K && bn + i < N) ? __ldg(B + (long)bk * ldb + bn + i) : 0.f;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) As[buf][a_k + i][a_row] = ra[i];
        *reinterpret_cast<float4*>(&Bs[buf][b_k][b_n]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int jh = 0; jh < 2; ++jh) {
            const int gn = n0 + jh * 64 + tx * 4;
            float* c = C + gm * ldc + gn;
            if (gn + 3 < N && (ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0)) {
                float4 o = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
                if (accumulate) { const float4 p = *reinterpret_cast<const float4*>(c); o = f4add(o, p); }
                *reinterpret_cast<float4*>(c) = o;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (gn + j < N) c[j] = accumulate ? c[j] + acc[i][jh * 4 + j] : acc[i][jh * 4 + j];
            }
        }
    }
}

constexpr int TK = 64, TN = 64, TM = 16;

// partial[z][ka][nb] = sum over m in split z
__global__ void __launch_bounds__(256)
sgemm_tn_partial_kernel(const float* __restrict__ A, long lda, const float* __restrict__ B, long ldb,
                        float* __restrict__ part, long M, int Ka, int Nb, long m_per_split, int a_vec, int b_vec) {
    __shared__ __align__(16) float As[TM][TK];
    __shared__ __align__(16) float Bs[TM][TN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int k0 = blockIdx.y * TK, n0 = blockIdx.x * TN;
    const long mbeg = (long)blockIdx.z * m_per_split;
    const long mend = min(M, mbeg + m_per_split);
    const int lr = tid >> 4, lc = (tid & 15) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long m = mbeg; m < mend; m += TM) {
        const long gm = m + lr;
        float4 ta = make_float4(0.f, 0.f, 0.f, 0.f), tb = ta;
        if (gm < mend) {
            const int gk = k0 + lc, gn = n0 + lc;
            if (a_vec && gk + 3 < Ka) ta = ldg4(A + gm * lda + gk);
            else {
                ta.x = gk + 0 < Ka ? __ldg(A + gm * lda + gk + 0) : 0.f;
                ta.y = gk + 1 < Ka ? __ldg(A + gm * lda + gk + 1) : 0.f;
                ta.z = gk + 2 < Ka ? __ldg(A + gm * lda + gk + 2) : 0.f;
                ta.w = gk + 3 < Ka ? __ldg(A + gm * lda + gk + 3) : 0.f;
            }
            if (b_vec && gn + 3 < Nb) tb = ldg4(B + gm * ldb + gn);
            else {
                tb.x = gn + 0 < Nb ? __ldg(B + gm * ldb + gn + 0) : 0.f;
                tb.y = gn + 1 < Nb ? __ldg(B + gm * ldb + gn + 1) : 0.f;
                tb.z = gn + 2 < Nb ? __ldg(B + gm * ldb + gn + 2) : 0.f;
                tb.w = gn + 3 < Nb ? __ldg(B + gm * ldb + gn + 3) : 0.f;
            }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&As[lr][lc]) = ta;
        *reinterpret_cast<float4*>(&Bs[lr][lc]) = tb;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < TM; ++r) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[r][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[r][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
    float* p = part + (long)blockIdx.z * Ka * Nb;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gk = k0 + ty * 4 + i;
        if (gk >= Ka) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn < Nb) p[(long)gk * Nb + gn] = acc[i][j];
        }
    }
}

__global__ void tn_reduce_kernel(const float* __restrict__ part, int splits, long elems, int Nb,
                                 float* __restrict__ C, long ldc, int accumulate) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= elems) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(long)z * elems + idx];     // fixed order
    float* c = C + (idx / Nb) * ldc + (idx % Nb);
    *c = accumulate ? *c + s : s;
}
}  // namespace

int gemm_nn_simt(const float* A, long lda, const float* B, long ldb, float* C, long ldc,
                 long M, int N, int K, int accumulate, cudaStream_t s) {
    if (M <= 0 || N <= 0) return 0;
    const int a_vec = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const int b_vec = (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    dim3 grid((N + BN - 1) / BN, (unsigned)((M + BM - 1) / BM));
    if (grid.y > 65535u) {
        // split very tall problems into bands of 65535 row tiles
        const long band = 65535L * BM;
        for (long m = 0; m < M; m += band) {
            const long mm = (M - m < band) ? (M - m) : band;
            dim3 g((N + BN - 1) / BN, (unsigned)((mm + BM - 1) / BM));
            sgemm_nn_kernel<<<g, 256, 0, s>>>(A + m * lda, lda, B, ldb, C + m * ldc, ldc, mm, N, K, accumulate, a_vec, b_vec);
        }
    } else {
        sgemm_nn_kernel<<<grid, 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, a_vec, b_vec);
    }
    return check_launch("sgemm_nn");
}

long gemm_tn_workspace_floats(long M, int Ka, int Nb) {
    return (long)gemm_tn_splits(M, Ka, Nb) * Ka * Nb;
}

int gemm_tn_splits(long M, int Ka, int Nb) {
    const long tiles = (long)((Ka + TK - 1) / TK) * ((Nb + TN - 1) / TN);
    long want = (148L * 8 + tiles - 1) / tiles;              // ~8 CTAs per SM in flight
    const long max_by_m = (M + 1023) / 1024;                 // at least 1024 rows per split
    if (want > max_by_m) want = max_by_m;
    if (want < 1) want = 1;
    if (want > 512) want = 512;
    return (int)want;
}

int gemm_tn_simt(const float* A, long lda, const float* B, long ldb, float* C, long ldc,
                 long M, int Ka, int Nb, int accumulate, float* workspace, cudaStream_t s) {
    if (Ka <= 0 || Nb <= 0) return 0;
    const int splits = gemm_tn_splits(M, Ka, Nb);
    long mps = (M + splits - 1) / splits;
    mps = ((mps + TM - 1) / TM) * TM;
    if (mps < TM) mps = TM;
    const int a_vec = (lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const int b_vec = (ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    dim3 grid((Nb + TN - 1) / TN, (Ka + TK - 1) / TK, splits);
    sgemm_tn_partial_kernel<<<grid, 256, 0, s>>>(A, lda, B, ldb, workspace, M, Ka, Nb, mps, a_vec, b_vec);
    if (int rc = check_launch("sgemm_tn_partial")) return rc;
    const long elems = (long)Ka * Nb;
    tn_reduce_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, s>>>(workspace, splits, elems, Nb, C, ldc, accumulate);
    return check_launch("tn_reduce");
}

}  // namespace spk
