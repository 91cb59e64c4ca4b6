#pragma once
#include <cuda_runtime.h>
namespace spk {
int rownorm(const float* x, long ldx, float* y, long ldy, long n_rows, int width, cudaStream_t s);
int residual_norm(const float* ew, long lde, const float* x2, long ldx, const float* mask, float* out, long ldo,
                  float* inv_norm, long n_rows, int width, cudaStream_t s);
int residual_norm_bwd(const float* g, long ldg, const float* out, long ldo, const float* mask, const float* inv_norm,
                      float* dew, long lde, float* dx2, long ldx, long n_rows, int width, cudaStream_t s);
int mask_from_index(const long long* idx, long n_idx, float* mask, long n_rows, int* flag, cudaStream_t s);
// extended weight assembly (spk_weights.cu)
struct AttnWeightsArgs {
    const float* a[4]; const float* a2[4];
    float* da[4]; float* da2[4];
    int H, F, Rd, D, mode;
    int Dp, Dt, Wd;            // mode 0
    int Fp, LZ;                // mode 1
    float* W0; long ld0; float* W1; long ld1; float* W2;
};
int launch_attn_weights_fwd(const AttnWeightsArgs& w, cudaStream_t s);
int launch_attn_weights_bwd(const AttnWeightsArgs& w, cudaStream_t s);
long inner_product_workspace_bytes();
int inner_product(const float* a, const float* b, long n, void* workspace, float* out, int accumulate, cudaStream_t s);
}  // namespace spk
