// N3: the per-iteration triple sampler, Corpus.get_iteration_triples_batch
// (GAT/create_batch.py:262-351), on the device. The positives come from the batch adjacency (K0b, nhop.py); this file
// builds the membership structure for valid_triples_dict (create_batch.py:82-83) and corrupts the tiled copies.
//
//   key(h, r, t) = (h * R + r) * N + t  (int64; the caller passes the triples already in (h, r, t) lexicographic order,
//   so the keys come out sorted) and membership is a binary search: HBM/L2-bound integer work, ~log2(M) dependent loads.
//
// One thread per negative slot s in [0, 2*ratio*P); row P + s starts as a copy of positive s mod P (np.tile,
// create_batch.py:298-301) and, with half = ratio / 2:
//   s in [0, P*half)            head  <- candidate entity   while (cand, r, t) is valid: redraw        value -1   (304-313)
//   s in [P*half, 2*P*half)     tail  <- candidate entity   while (h, r, cand) is valid: redraw        value -1   (315-326)
//   s in [2*P*half, P*ratio)    untouched copy (odd ratio)                                              value +1
//   s in [P*ratio, 2*P*ratio)   rel   <- candidate relation while (h, cand, t) valid: redraw, at most R redraws;
//                               after R redraws the row stays a +1 copy of the positive                 value -1   (328-347)
// The first candidate of a slot is the caller's draw (init_ent / init_rel, the reference's random_entities /
// random_relations arrays, index s resp. s - P*ratio) when given, else generated; redraws come from a counter-based
// generator keyed on (seed, slot, attempt) -- the reference consumes numpy's global stream sequentially instead, so
// rows are bit-identical to the reference's exactly when it needed no redraw for them (tests/test_sampler.py).
#include "../../include/spkbgat.h"
#include "spk_common.cuh"

namespace spk {
namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {      // splitmix64 finalizer
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
// uniform integer in [0, n) from (seed, slot, attempt): multiply-shift of a 64-bit hash
__device__ __forceinline__ long long draw(unsigned long long seed, unsigned long long slot, unsigned attempt, long long n) {
    const unsigned long long x = mix64(mix64(seed ^ (slot * 0xd1342543de82ef95ULL)) + attempt);
    return (long long)__umul64hi(x, (unsigned long long)n);
}
__device__ __forceinline__ bool is_valid(const long long* __restrict__ keys, long long m, long long key) {
    long long lo = 0, hi = m;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        const long long v = __ldg(keys + mid);
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return lo < m && __ldg(keys + lo) == key;
}

__global__ void __launch_bounds__(256)
triple_keys_kernel(const long long* __restrict__ tri, long long m, long long n_ent, long long n_rel,
                   long long* __restrict__ keys, int* __restrict__ err) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    long long h = tri[3 * i], r = tri[3 * i + 1], t = tri[3 * i + 2];
    if (h < 0 || h >= n_ent || t < 0 || t >= n_ent || r < 0 || r >= n_rel) { *err = 1; h = r = t = 0; }
    keys[i] = (h * n_rel + r) * n_ent + t;
}

constexpr int MAX_ENTITY_REDRAWS = 4096;      // the reference loops forever when every entity is valid; we stop here

__global__ void __launch_bounds__(256)
corrupt_triples_kernel(const long long* __restrict__ pos, long long P, int ratio,
                       const long long* __restrict__ keys, long long m, long long n_ent, long long n_rel,
                       const long long* __restrict__ init_ent, const long long* __restrict__ init_rel,
                       unsigned long long seed, long long* __restrict__ out, float* __restrict__ val) {
    const long long s = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long n_neg = 2LL * ratio * P;
    if (s >= P + n_neg) return;
    if (s < P) {                                   // rows [0, P): the positives themselves
        out[3 * s] = pos[3 * s]; out[3 * s + 1] = pos[3 * s + 1]; out[3 * s + 2] = pos[3 * s + 2];
        val[s] = 1.0f;
        return;
    }
    const long long q = s - P;                     // negative slot
    const long long p = q % P;
    long long h = pos[3 * p], r = pos[3 * p + 1], t = pos[3 * p + 2];
    float v = 1.0f;
    const long long half = ratio / 2;
    if (q < 2 * P * half) {
        const bool head = q < P * half;
        long long cand = init_ent ? init_ent[q] : draw(seed, (unsigned long long)q, 0u, n_ent);
        if (cand < 0 || cand >= n_ent) cand = draw(seed, (unsigned long long)q, 0u, n_ent);
        for (int a = 1; a <= MAX_ENTITY_REDRAWS; ++a) {
            const long long key = head ? (cand * n_rel + r) * n_ent + t : (h * n_rel + r) * n_ent + cand;
            if (!is_valid(keys, m, key)) break;
            cand = draw(seed, (unsigned long long)q, (unsigned)a, n_ent);
        }
        if (head) h = cand; else t = cand;
        v = -1.0f;
    } else if (q >= P * (long long)ratio) {
        const long long cr = q - P * (long long)ratio;
        long long cand = init_rel ? init_rel[cr] : draw(seed, (unsigned long long)q, 0u, n_rel);
        if (cand < 0 || cand >= n_rel) cand = draw(seed, (unsigned long long)q, 0u, n_rel);
        long long rel_count = 0;
        while (is_valid(keys, m, (h * n_rel + cand) * n_ent + t)) {
            ++rel_count;
            cand = draw(seed, (unsigned long long)q, (unsigned)rel_count, n_rel);
            if (rel_count >= n_rel) break;
        }
        if (rel_count < n_rel) { r = cand; v = -1.0f; }
    }
    out[3 * s] = h; out[3 * s + 1] = r; out[3 * s + 2] = t;
    val[s] = v;
}

}  // namespace
}  // namespace spk

using namespace spk;

extern "C" {

int spk_triple_keys(const int64_t* triples, int64_t n_triples, int64_t n_ent, int64_t n_rel, int64_t* keys,
                    int32_t* err_flag, spk_stream_t stream) {
    if (n_triples <= 0) return 0;
    if (n_ent <= 0 || n_rel <= 0 || (double)n_ent * (double)n_rel * (double)n_ent >= 9.0e18) {
        set_error("triple_keys: N*R*N = %lld*%lld*%lld does not fit the int64 key", (long long)n_ent, (long long)n_rel, (long long)n_ent);
        return 2;
    }
    triple_keys_kernel<<<(unsigned)((n_triples + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)triples, n_triples, n_ent, n_rel, (long long*)keys, err_flag);
    return check_launch("triple_keys");
}

int spk_corrupt_triples(const int64_t* positives, int64_t n_pos, int32_t ratio, const int64_t* valid_keys, int64_t n_valid,
                        int64_t n_ent, int64_t n_rel, const int64_t* init_entities, const int64_t* init_relations,
                        uint64_t seed, int64_t* out_indices, float* out_values, spk_stream_t stream) {
    if (n_pos <= 0) return 0;
    if (ratio < 0 || n_ent <= 0 || n_rel <= 0 || (double)n_ent * (double)n_rel * (double)n_ent >= 9.0e18) {
        set_error("corrupt_triples: bad ratio %d or table sizes", ratio);
        return 2;
    }
    const long long total = n_pos * (2LL * ratio + 1);
    corrupt_triples_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const long long*)positives, n_pos, ratio, (const long long*)valid_keys, n_valid, n_ent, n_rel,
        (const long long*)init_entities, (const long long*)init_relations, (unsigned long long)seed,
        (long long*)out_indices, out_values);
    return check_launch("corrupt_triples");
}

}  // extern "C"
