// On-device filtered negative sampling for the completion scorer.
//
// Reference being replaced: TrainDataset.get_neg_sample (modules/load/data_loader.py:36-47) -- per training triple an
// O(num_ent) boolean mask and np.random.choice(all_ent[mask], neg_num, replace=False) on the CPU, run in 12 DataLoader
// workers (train.py:288-294).  Here one thread per (h, r) row draws candidates from a counter-based Philox-4x32-10 stream
// (seed, row) and rejects known tails (binary search in the sorted key list (h*n_rel + r)*n_ent + t of all known triples) and
// repeats of an earlier accepted candidate of the row.  Sequential rejection of an i.i.d. uniform stream yields the uniform
// distribution over ordered k-subsets of the valid entities, i.e. the reference's distribution; the random stream itself is
// Philox, not numpy's Mersenne twister, so parity is distributional (the scorer's parity tests feed precomputed negatives).
// Integer work, a few KB of traffic per row: latency-bound, one launch per batch.
#include <stdint.h>

#include "common.cuh"

namespace jmac {

// Philox-4x32-10 (Salmon et al., SC'11): counter (row, draw index), key (seed)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ bool is_known(const int64_t* __restrict__ keys, int64_t n_keys, int64_t key) {
    int64_t lo = 0, hi = n_keys;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo < n_keys && __ldg(keys + lo) == key;
}

__global__ void __launch_bounds__(128) sample_negatives_kernel(const int64_t* __restrict__ keys, int64_t n_keys, const int64_t* __restrict__ h,
                                                               const int64_t* __restrict__ r, int64_t B, int64_t n_ent, int64_t n_rel, int k,
                                                               uint64_t seed, int64_t* __restrict__ out, int* __restrict__ fail) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int64_t hh = h[b], rr = r[b];
    int64_t* row = out + b * k;
    if ((uint64_t)hh >= (uint64_t)n_ent || (uint64_t)rr >= (uint64_t)n_rel) {
        *reinterpret_cast<volatile int*>(fail) = 1;
        for (int j = 0; j < k; ++j) row[j] = 0;
        return;
    }
    const int64_t base = (hh * n_rel + rr) * n_ent;
    // unbiased integers in [0, n_ent): multiply-shift with rejection of the short first interval (Lemire)
    const uint32_t n32 = (uint32_t)n_ent;
    const uint32_t thresh = (uint32_t)(-n32) % n32;
    int got = 0;
    const uint32_t max_blocks = 1u << 20;                       // 4 M draws per row: unreachable when k <= number of valid entities
    for (uint32_t blk = 0; blk < max_blocks && got < k; ++blk) {
        uint32_t c[4] = {blk, 0u, (uint32_t)b, (uint32_t)((uint64_t)b >> 32)};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
        for (int q = 0; q < 4 && got < k; ++q) {
            const uint64_t prod = (uint64_t)c[q] * n32;
            if ((uint32_t)prod < thresh) continue;
            const int64_t cand = (int64_t)(prod >> 32);
            bool dup = false;
            for (int j = 0; j < got; ++j) dup |= row[j] == cand;
            if (dup || is_known(keys, n_keys, base + cand)) continue;
            row[got++] = cand;
        }
    }
    if (got < k) {
        *reinterpret_cast<volatile int*>(fail) = 1;
        for (int j = got; j < k; ++j) row[j] = 0;
    }
}

}  // namespace jmac

using namespace jmac;

extern "C" int jmac_sample_negatives(const int64_t* keys, int64_t n_keys, const int64_t* h, const int64_t* r, int64_t B, int64_t n_ent,
                                     int64_t n_rel, int k, uint64_t seed, int64_t* out, void* stream) {
    JMAC_ARG(h && r && out && B >= 0 && k >= 0 && n_keys >= 0 && (n_keys == 0 || keys != nullptr));
    JMAC_ARG(n_ent > 0 && n_ent < (int64_t(1) << 31) && n_rel > 0);
    if (B == 0 || k == 0) return 0;
    int* fail = status_words() ? status_words() + JMAC_ST_SAMPLER : nullptr;
    JMAC_ARG(fail != nullptr);
    sample_negatives_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(keys, n_keys, h, r, B, n_ent, n_rel, k, seed, out, fail);
    JMAC_LAUNCH_CHECK();
    return 0;
}
