#pragma once
#include <cuda_runtime.h>
namespace spk {
int edges_concat(const long long* edge, long e1, const long long* etype, const long long* nhop, long e2,
                 int* row, int* col, int* t1, int* t2, long n_nodes, long n_rel, int* err, cudaStream_t s);
int iota_i32(int* v, long n, cudaStream_t s);
long radix_sort_workspace_bytes(long n);
int radix_sort_pairs(int* keys, int* vals, int* keys_tmp, int* vals_tmp, long n, int key_bits,
                     void* workspace, int* result_in_tmp, cudaStream_t s);
int segment_ptr(const int* sorted_keys, long n, int n_seg, int* ptr, cudaStream_t s);
int gather_i32(const int* src, const int* idx, long n, int* out, cudaStream_t s);
int rel_incidence(const int* t1, const int* t2, long e, int n_rel, int* keys, int* vals, cudaStream_t s);
}  // namespace spk
