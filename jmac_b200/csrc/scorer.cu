// Completion scorer (TransE-L1 margin loss) and pair-cosine kernels.
//
// Reference being replaced: the scorer part of JMAC.completion_loss (src/jmac_model.py:345-374) on
// the batch layout of train.py:347-352, and the F.normalize + row-dot of alignment_loss_simple /
// alignment_loss (src/jmac_model.py:243-248, 268-287).
//
// Forward: one warp per positive b walks its K+1 triples m = k*B + b.  The (h, r) rows are reloaded only
// when the index changes (they are tiled K+1 times in the reference layout), so the HBM traffic is
// the algorithmic (K+3) rows per positive per layer.  When a gradient will be asked for, the same pass
// records everything backward needs WITHOUT re-gathering a single embedding row:
//   coef[l][m]  d loss / d score_m up to the scale: the number of active hinge terms for a positive, -1 / 0 for an
//               active / inactive negative (the positive is the first triple of its warp, so activity is known on the fly);
//   mask[l][m]  the sign pattern of h + r - t (two bits per feature: > 0, < 0; sign(0) = 0 like torch.abs' gradient);
//   dhr[l][b]   sum_k coef_k sgn_k over the leading run of equal (h, r) of positive b (all K+1 triples in the reference
//               layout): the dense gradient row that h_b and r_b receive from that run.
// Backward is then an atomics-free, bit-reproducible segmented reduction: the occurrences of every table row -- tails
// (mask records), heads / relations outside a leading run (mask records), one dense record per positive for its head and
// its relation -- are radix-sorted by row id once per batch (all layers share the indices); warps reduce fixed chunks of
// the sorted list, rows whose occurrences straddle chunks are finished by a second, ordered pass over per-chunk carry rows.
// Every gradient row is written exactly once, by one warp, in a fixed order: no float REDs (round 1 used vector REDs: 8 GB of
// read-modify-write at B = 65536, K = 64, and run-to-run differences in the last bits).
#include <math.h>

#include <cub/cub.cuh>

#include "common.cuh"

namespace jmac {

constexpr int MAX_LAYERS = 4;

struct TransEParams {
    const float* ent[MAX_LAYERS];
    const float* rel[MAX_LAYERS];
    float* d_ent[MAX_LAYERS];
    float* d_rel[MAX_LAYERS];
    int n_layers, d4;
    const int64_t *h, *r, *t;
    int64_t M, B;
    int K1;  // M / B
    float margin;
    float* score;        // [n_layers, M]
    float* loss;         // [n_layers]
    const float* gloss;  // [n_layers]
    int64_t n_ent, n_rel;
    int* status;         // out-of-range index flag (common.cuh: status_words)
    // record for the atomics-free backward (all null when no gradient is needed)
    float* coef;         // [n_layers, B, K1]
    void* smask;         // [n_layers, B, K1, 32 lanes] SMaskT: low half = "h+r-t > 0" bits of the lane's features, high half = "< 0" bits
    float* dhr;          // [n_layers, B, d]
    float* loss_part;    // [n_layers, gridDim.x]: per-block hinge sums, added up in a fixed order by transe_loss_kernel
};

// sign pattern of a lane's 4*VPL features: VPL <= 2 -> 8 + 8 bits, else 16 + 16 bits
template <int VPL> struct SMaskT { typedef uint16_t type; static constexpr int HALF = 8; };
template <> struct SMaskT<3> { typedef uint32_t type; static constexpr int HALF = 16; };
template <> struct SMaskT<4> { typedef uint32_t type; static constexpr int HALF = 16; };
inline size_t transe_mask_bytes(int d) { return (((d / 4 + 31) / 32) <= 2) ? 64 : 128; }

// Sign pattern of a - b for the lane's features: bit f of the low half = (a - b > 0), of the high half = (a - b < 0).
// One funnel shift per feature collects the IEEE sign bits; exact zeros (sign(0) = 0, like torch.abs' gradient) are rare and are
// found with one running min of |a - b| per lane, then patched on a slow path.
template <int VPL>
__device__ __forceinline__ unsigned sign_bits(const Row<VPL>& a, const Row<VPL>& b) {
    constexpr int H = SMaskT<VPL>::HALF;
    unsigned sg = 0u;
    float mn = INFINITY;
#pragma unroll
    for (int i = VPL - 1; i >= 0; --i) {
        const float4 dlt = a.v[i] - b.v[i];
        sg = __funnelshift_l(__float_as_uint(dlt.w), sg, 1);
        sg = __funnelshift_l(__float_as_uint(dlt.z), sg, 1);
        sg = __funnelshift_l(__float_as_uint(dlt.y), sg, 1);
        sg = __funnelshift_l(__float_as_uint(dlt.x), sg, 1);
        mn = fminf(mn, fminf(fminf(fabsf(dlt.x), fabsf(dlt.y)), fminf(fabsf(dlt.z), fabsf(dlt.w))));
    }
    constexpr unsigned ALL = (1u << (4 * VPL)) - 1u;
    unsigned nz = ALL;
    if (mn == 0.f) {                                          // some feature of this lane is exactly zero
        nz = 0u;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 dlt = a.v[i] - b.v[i];
            nz |= (dlt.x != 0.f ? 1u : 0u) << (4 * i) | (dlt.y != 0.f ? 1u : 0u) << (4 * i + 1) | (dlt.z != 0.f ? 1u : 0u) << (4 * i + 2) |
                  (dlt.w != 0.f ? 1u : 0u) << (4 * i + 3);
        }
    }
    return ((~sg) & nz) | ((sg & nz) << H);
}
// acc += c * sgn, sgn decoded from the lane's bits
template <int VPL>
__device__ __forceinline__ void add_signed(Row<VPL>& acc, float c, unsigned m) {
    constexpr int H = SMaskT<VPL>::HALF;
    constexpr unsigned ALL = (1u << (4 * VPL)) - 1u;
    const unsigned pb = m & ALL, nb = (m >> H) & ALL;
    if ((pb | nb) == ALL) {                                   // no exact zero in this lane's features (the common case): +c or -c
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const unsigned q = nb >> (4 * i);
            acc.v[i].x += (q & 1u) ? -c : c;
            acc.v[i].y += (q & 2u) ? -c : c;
            acc.v[i].z += (q & 4u) ? -c : c;
            acc.v[i].w += (q & 8u) ? -c : c;
        }
    } else {
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const unsigned p4 = pb >> (4 * i), n4 = nb >> (4 * i);
            acc.v[i].x += ((p4 & 1u) ? c : 0.f) - ((n4 & 1u) ? c : 0.f);
            acc.v[i].y += ((p4 & 2u) ? c : 0.f) - ((n4 & 2u) ? c : 0.f);
            acc.v[i].z += ((p4 & 4u) ? c : 0.f) - ((n4 & 4u) ? c : 0.f);
            acc.v[i].w += ((p4 & 8u) ? c : 0.f) - ((n4 & 8u) ? c : 0.f);
        }
    }
}

template <int VPL>
__device__ __forceinline__ float l1_diff(const Row<VPL>& a, const Row<VPL>& b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
        s += fabsf(a.v[i].x - b.v[i].x) + fabsf(a.v[i].y - b.v[i].y) + fabsf(a.v[i].z - b.v[i].z) + fabsf(a.v[i].w - b.v[i].w);
    return s;
}

// Pull up to eight tail rows towards L2 ahead of their use: lane group g = lane / 4 takes row idx[first + g], its four lanes the
// row's 128-byte lines.  The warp's own loads then find them in L2 (~3x shorter latency), which is what the few rows a warp can
// hold in registers do not cover.
__device__ __forceinline__ void prefetch_tails_l2(const float* table, int tv, int first, int cnt, int row_f, int lane) {
    const int g = lane >> 2, ln = lane & 3;
    const int pos = first + g;
    const int row = __shfl_sync(FULL, tv, pos & 31);
    if (pos < cnt) {
        const char* base = reinterpret_cast<const char*>(table + (size_t)row * row_f);
        const int lines = (row_f * 4 + 127) >> 7;
        for (int q = ln; q < lines; q += 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + q * 128));
    }
}

template <int VPL, int U, bool SAVE, int MINB>
__global__ void __launch_bounds__(256, MINB) transe_fwd_kernel(TransEParams p) {
    typedef typename SMaskT<VPL>::type smask_t;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    __shared__ float block_loss[MAX_LAYERS][8];
    const bool live = b < p.B;
    const int d4 = p.d4;
    for (int l = 0; l < p.n_layers; ++l) {
        float lsum = 0.f;
        if (live) {
            const float* ent = p.ent[l];
            const float* rel = p.rel[l];
            // record layout: [layer][positive b][k]: a warp writes its K1 coefficients / sign records contiguously
            smask_t* maskl = SAVE ? reinterpret_cast<smask_t*>(p.smask) + ((size_t)l * p.M + (size_t)b * p.K1) * 32 : nullptr;
            float* coefl = SAVE ? p.coef + (size_t)l * p.M + (size_t)b * p.K1 : nullptr;
            int hprev = -1, rprev = -1, h0 = -1, r0 = -1;    // ids fit 31 bits (checked by the host): 32-bit keeps the register count down
            bool in_lead = true;                             // still inside the leading run of triples that share (h, r) with the positive
            Row<VPL> hr, dlead;
            hr.zero();
            dlead.zero();
            float pos = 0.f;
            unsigned m0 = 0u;
            int n_act = 0;
            for (int kb = 0; kb < p.K1; kb += 32) {
                const int cnt = min(32, p.K1 - kb);
                int hv = 0, rv = 0, tv = 0;
                if (lane < cnt) {
                    const int64_t m = (int64_t)(kb + lane) * p.B + b;
                    hv = (int)safe_index(__ldg(p.h + m), p.n_ent, p.status);
                    rv = (int)safe_index(__ldg(p.r + m), p.n_rel, p.status);
                    tv = (int)safe_index(__ldg(p.t + m), p.n_ent, p.status);
                }
                float sc_keep = 0.f, cf_keep = 0.f;
                prefetch_tails_l2(ent, tv, 0, cnt, d4 * 4, lane);
                for (int k = 0; k < cnt; k += U) {
                    if ((k & 7) == 0) prefetch_tails_l2(ent, tv, k + 8, cnt, d4 * 4, lane);       // eight rows ahead of the consumer
                    Row<VPL> tr[U];
                    int hh[U], rr[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int kk = min(k + u, cnt - 1);
                        hh[u] = __shfl_sync(FULL, hv, kk);
                        rr[u] = __shfl_sync(FULL, rv, kk);
                        const int tt = __shfl_sync(FULL, tv, kk);
                        load_row<VPL, true>(tr[u], ent, tt, d4, lane);
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (k + u < cnt) {
                            if (hh[u] != hprev || rr[u] != rprev) {  // warp-uniform
                                Row<VPL> hrow, rrow;
                                load_row<VPL, true>(hrow, ent, hh[u], d4, lane);
                                load_row<VPL, false>(rrow, rel, rr[u], d4, lane);
#pragma unroll
                                for (int i = 0; i < VPL; ++i) hr.v[i] = hrow.v[i] + rrow.v[i];
                                hprev = hh[u]; rprev = rr[u];
                            }
                            const bool first = kb + k + u == 0;
                            if (first) { h0 = hh[u]; r0 = rr[u]; }
                            if (hh[u] != h0 || rr[u] != r0) in_lead = false;
                            const float sc = warp_sum(l1_diff<VPL>(hr, tr[u]));
                            if (lane == k + u) sc_keep = sc;
                            unsigned bits = 0u;
                            if (SAVE) {
                                bits = sign_bits<VPL>(hr, tr[u]);
                                maskl[(size_t)(kb + k + u) * 32 + lane] = (smask_t)bits;
                            }
                            if (first) {
                                pos = sc;
                                m0 = bits;
                            } else {
                                lsum += fmaxf(pos - sc, -p.margin);
                                if (SAVE) {
                                    const bool active = pos - sc > -p.margin;      // warp-uniform
                                    if (lane == k + u) cf_keep = active ? -1.f : 0.f;
                                    if (active) {
                                        ++n_act;
                                        if (in_lead) add_signed<VPL>(dlead, -1.f, bits);
                                    }
                                }
                            }
                        }
                    }
                }
                if (lane < cnt) {
                    if (p.score) p.score[(int64_t)l * p.M + (int64_t)(kb + lane) * p.B + b] = sc_keep;
                    if (SAVE && kb + lane > 0) coefl[kb + lane] = cf_keep;
                }
            }
            if (SAVE) {
                if (lane == 0) coefl[0] = (float)n_act;           // the positive (k = 0): one unit per active hinge term
                add_signed<VPL>(dlead, (float)n_act, m0);
                store_row<VPL>(dlead, p.dhr + (size_t)l * p.B * d4 * 4, b, d4, lane);
            }
        }
        if (lane == 0) block_loss[l][wib] = lsum;
    }
    __syncthreads();
    if (threadIdx.x < p.n_layers) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += block_loss[threadIdx.x][w];
        p.loss_part[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
    }
}

// loss[l] = margin + (sum of the block partials, in a fixed order) / (B (K1 - 1)): one block per layer, bit-reproducible
__global__ void __launch_bounds__(256) transe_loss_kernel(const float* __restrict__ part, int n_blocks, float margin, float inv, float* __restrict__ loss) {
    __shared__ float sh[256];
    const float* pl = part + (size_t)blockIdx.x * n_blocks;
    float s = 0.f;
    for (int i = threadIdx.x; i < n_blocks; i += 256) s += pl[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[blockIdx.x] = margin + sh[0] * inv;
}

// ------------------------------------------------------------------------------------------------
// Backward: sorted occurrence lists + chunked segmented reduction (see the file header).
// Occurrence value word: bit 31 = dense record (payload = positive b: the row dhr[l][b]); bit 30 = head / relation role (sign +1),
// else tail role (sign -1); payload = triple position m.  Key = table row, or `sentinel` (one past the last row) for the
// head / relation occurrences that the dense record of their positive already covers.
// ------------------------------------------------------------------------------------------------
constexpr unsigned OCC_DENSE = 0x80000000u, OCC_HR = 0x40000000u, OCC_PAYLOAD = 0x3fffffffu;

// cover[b] = length of the leading run of triples k = 0.. that share (h, r) with the positive (K1 in the reference layout)
__global__ void transe_cover_kernel(const int64_t* __restrict__ h, const int64_t* __restrict__ r, int64_t B, int K1, int* __restrict__ cover) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const int64_t h0 = h[b], r0 = r[b];
    int first = K1;
    for (int k = lane; k < K1 && first == K1; k += 32) {
        if (h[(int64_t)k * B + b] != h0 || r[(int64_t)k * B + b] != r0) first = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(FULL, first, o));
    if (lane == 0) cover[b] = first;
}

__global__ void transe_keys_kernel(const int64_t* __restrict__ h, const int64_t* __restrict__ r, const int64_t* __restrict__ t, int64_t M,
                                   int64_t B, int64_t n_ent, int64_t n_rel, const int* __restrict__ cover, unsigned* __restrict__ ekeys,
                                   unsigned* __restrict__ evals, unsigned* __restrict__ rkeys, unsigned* __restrict__ rvals) {
    const int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int64_t b = m % B, k = m / B;
    const unsigned hh = (unsigned)safe_index(h[m], n_ent, nullptr), rr = (unsigned)safe_index(r[m], n_rel, nullptr),
                   tt = (unsigned)safe_index(t[m], n_ent, nullptr);          // flagged by the forward kernel already
    const bool covered = k < cover[b];
    ekeys[m] = tt;                                       evals[m] = (unsigned)m;
    ekeys[M + m] = covered ? (unsigned)n_ent : hh;       evals[M + m] = (unsigned)m | OCC_HR;
    rkeys[m] = covered ? (unsigned)n_rel : rr;           rvals[m] = (unsigned)m | OCC_HR;
    if (k == 0) {
        ekeys[2 * M + b] = hh;                           evals[2 * M + b] = (unsigned)b | OCC_DENSE;
        rkeys[M + b] = rr;                               rvals[M + b] = (unsigned)b | OCC_DENSE;
    }
}

struct ReduceParams {
    const unsigned* keys;     // sorted
    const unsigned* vals;
    int64_t n_occ;
    unsigned sentinel;
    int chunk;                // occurrences per warp (multiple of 32)
    int d4, K1;
    int64_t M, B;
    const float* coef;        // [L, B, K1]
    const void* smask;        // [L, B, K1, 32]
    const float* dhr;         // [L, B, d]
    const float* gloss;       // [L]
    float inv_pairs;          // 1 / (B (K1 - 1))
    float* out[MAX_LAYERS];   // gradient tables (zeroed by the caller)
    float* carry;             // [n_chunks, 2, L, d]: slot 0 = partial of the run continuing FROM the previous chunk, slot 1 = INTO the next
    int64_t n_chunks;
};

// Pass A: one warp per chunk of the sorted list, all NL layers in one walk (the index work is shared).  Runs of equal keys inside
// the chunk are written to the tables directly; the (at most two) runs that touch a chunk boundary and continue beyond it leave
// their partial sums in the carry rows.
template <int VPL, int NL>
__global__ void __launch_bounds__(256) transe_reduce_kernel(ReduceParams p) {
    typedef typename SMaskT<VPL>::type smask_t;
    const int lane = threadIdx.x & 31;
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (c >= p.n_chunks) return;
    const int64_t base = c * p.chunk;
    const int64_t end = min(p.n_occ, base + p.chunk);
    const unsigned first_key = p.keys[base];
    if (first_key == p.sentinel) return;
    const unsigned prev_last = c > 0 ? p.keys[base - 1] : 0xffffffffu;
    const unsigned next_first = end < p.n_occ ? p.keys[end] : 0xfffffffeu;
    const int d4 = p.d4;
    const size_t row_f = (size_t)d4 * 4;
    float gs[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) gs[l] = p.gloss[l] * p.inv_pairs;
    float* carry = p.carry + (size_t)c * 2 * NL * row_f;
    Row<VPL> acc[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) acc[l].zero();
    unsigned run_key = first_key;
    bool run_from_start = true;                            // the current run began at the first occurrence of the chunk

    auto finish_run = [&](bool to_head_slot, bool to_tail_slot) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            if (to_head_slot) store_row<VPL>(acc[l], carry, l, d4, lane);
            else if (to_tail_slot) store_row<VPL>(acc[l], carry, NL + l, d4, lane);
            else {
#pragma unroll
                for (int q = 0; q < VPL; ++q) acc[l].v[q] = gs[l] * acc[l].v[q];
                store_row<VPL>(acc[l], p.out[l], run_key, d4, lane);
            }
            acc[l].zero();
        }
    };

    for (int64_t b0 = base; b0 < end; b0 += 32) {
        const int cnt = (int)min((int64_t)32, end - b0);
        unsigned kv = p.sentinel, vv = 0u, rec = 0u;
        float cv[NL];
#pragma unroll
        for (int l = 0; l < NL; ++l) cv[l] = 0.f;
        if (lane < cnt) {
            kv = p.keys[b0 + lane];
            vv = p.vals[b0 + lane];
            if (kv != p.sentinel && !(vv & OCC_DENSE)) {
                const unsigned m = vv & OCC_PAYLOAD;
                rec = (unsigned)((m % (unsigned)p.B) * (unsigned)p.K1 + m / (unsigned)p.B);     // record index [b][k]
#pragma unroll
                for (int l = 0; l < NL; ++l) cv[l] = p.coef[(size_t)l * p.M + rec];
            } else {
                rec = vv & OCC_PAYLOAD;                                                            // dense record: positive b
            }
        }
        bool any = (vv & OCC_DENSE) != 0u;
#pragma unroll
        for (int l = 0; l < NL; ++l) any |= cv[l] != 0.f;
        const unsigned live = __ballot_sync(FULL, any && kv != p.sentinel);
        const unsigned valid = __ballot_sync(FULL, kv != p.sentinel);
        const int n_valid = __popc(valid);                 // the sentinels sort to the end: valid lanes are a prefix
        for (int i = 0; i < n_valid; ++i) {
            const unsigned ki = __shfl_sync(FULL, kv, i);
            if (ki != run_key) {
                // the finished run lies inside the chunk unless it began at the chunk start AND continues the previous chunk
                finish_run(run_from_start && run_key == prev_last, false);
                run_key = ki;
                run_from_start = false;
            }
            if (!((live >> i) & 1u)) continue;             // inactive negative in every layer: nothing to add (warp-uniform)
            const unsigned vi = __shfl_sync(FULL, vv, i);
            const unsigned ri = __shfl_sync(FULL, rec, i);
            if (vi & OCC_DENSE) {
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    Row<VPL> dr;
                    load_row<VPL, true>(dr, p.dhr + (size_t)l * p.B * row_f, ri, d4, lane);
#pragma unroll
                    for (int q = 0; q < VPL; ++q) acc[l].v[q] = acc[l].v[q] + dr.v[q];
                }
            } else {
                unsigned bits[NL];
#pragma unroll
                for (int l = 0; l < NL; ++l)
                    bits[l] = (unsigned)reinterpret_cast<const smask_t*>(p.smask)[((size_t)l * p.M + ri) * 32 + lane];
                const float sgn = (vi & OCC_HR) ? 1.f : -1.f;
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    const float ci = __shfl_sync(FULL, cv[l], i);
                    if (ci != 0.f) add_signed<VPL>(acc[l], sgn * ci, bits[l]);
                }
            }
        }
        if (n_valid < cnt) break;                          // the rest of the chunk is sentinels
    }
    // the last run of the chunk
    const bool head = run_from_start && run_key == prev_last;
    const bool tail = run_key == next_first;
    finish_run(head, !head && tail);                       // (head && tail: the whole chunk is one through-run, slot 0)
}

// Pass B: the chunk in which a multi-chunk run STARTS finishes it: its own tail partial plus the head partials of the following
// chunks, in list order, while they continue the same key.
template <int VPL>
__global__ void __launch_bounds__(256) transe_fixup_kernel(ReduceParams p, int NL) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (c >= p.n_chunks) return;
    const int64_t base = c * p.chunk;
    const int64_t end = min(p.n_occ, base + p.chunk);
    if (end >= p.n_occ) return;                               // no next chunk: nothing continues
    const unsigned key = p.keys[end - 1];                     // last key of this chunk
    if (key == p.sentinel || p.keys[end] != key) return;      // its last run does not continue
    // does the run start in this chunk?  Not if the whole chunk is that key and it already continued the previous chunk.
    if (p.keys[base] == key && c > 0 && p.keys[base - 1] == key) return;
    const int d4 = p.d4;
    const size_t row_f = (size_t)d4 * 4;
    for (int l = 0; l < NL; ++l) {
        const float gs = p.gloss[l] * p.inv_pairs;
        Row<VPL> acc;
        load_row<VPL, true>(acc, p.carry + ((size_t)c * 2 + 1) * NL * row_f, l, d4, lane);
        for (int64_t c2 = c + 1; c2 < p.n_chunks; ++c2) {
            const int64_t b2 = c2 * p.chunk, e2 = min(p.n_occ, b2 + p.chunk);
            if (p.keys[b2] != key) break;
            Row<VPL> part;
            load_row<VPL, true>(part, p.carry + (size_t)c2 * 2 * NL * row_f, l, d4, lane);
#pragma unroll
            for (int q = 0; q < VPL; ++q) acc.v[q] = acc.v[q] + part.v[q];
            const bool through = p.keys[e2 - 1] == key && e2 < p.n_occ && p.keys[e2] == key;
            if (!through) break;
        }
#pragma unroll
        for (int q = 0; q < VPL; ++q) acc.v[q] = gs * acc.v[q];
        store_row<VPL>(acc, p.out[l], key, d4, lane);
    }
}

// ------------------------------------------------------------------------------------------------
// Pair cosine (alignment losses).  Its backward scatters with vector REDs: a few thousand seed links, collisions are rare.
template <int VPL>
__device__ __forceinline__ void red_row(float* base, int64_t row, int d4, int lane, const Row<VPL>& v) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = i * 32 + lane;
        if (c < d4) red_add_v4(base + (row * d4 + c) * 4, v.v[i]);
    }
}

template <int VPL>
__global__ void __launch_bounds__(256) pair_cos_fwd_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int d4,
                                                           const int64_t* __restrict__ ia, const int64_t* __restrict__ ib, int64_t M,
                                                           int64_t nA, int64_t nB, int* status, float* __restrict__ cosv) {
    const int lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    Row<VPL> a, b;
    load_row<VPL, true>(a, A, safe_index(ia[m], nA, status), d4, lane);
    load_row<VPL, true>(b, Bm, safe_index(ib[m], nB, status), d4, lane);
    float dt = 0.f, na = 0.f, nb = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) { dt += dot4(a.v[i], b.v[i]); na += dot4(a.v[i], a.v[i]); nb += dot4(b.v[i], b.v[i]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dt += __shfl_xor_sync(FULL, dt, o); na += __shfl_xor_sync(FULL, na, o); nb += __shfl_xor_sync(FULL, nb, o);
    }
    if (lane == 0) cosv[m] = dt / (fmaxf(sqrtf(na), 1e-12f) * fmaxf(sqrtf(nb), 1e-12f));
}

template <int VPL>
__global__ void __launch_bounds__(256) pair_cos_bwd_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int d4,
                                                           const int64_t* __restrict__ ia, const int64_t* __restrict__ ib, int64_t M,
                                                           int64_t nA, int64_t nB, int* status, const float* __restrict__ gcos,
                                                           float* __restrict__ dA, float* __restrict__ dB) {
    const int lane = threadIdx.x & 31;
    const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= M) return;
    const float g = gcos[m];
    if (g == 0.f) return;  // hinge-inactive pair
    const int64_t ra = safe_index(ia[m], nA, status), rb = safe_index(ib[m], nB, status);
    Row<VPL> a, b;
    load_row<VPL, true>(a, A, ra, d4, lane);
    load_row<VPL, true>(b, Bm, rb, d4, lane);
    float dt = 0.f, na = 0.f, nb = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) { dt += dot4(a.v[i], b.v[i]); na += dot4(a.v[i], a.v[i]); nb += dot4(b.v[i], b.v[i]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dt += __shfl_xor_sync(FULL, dt, o); na += __shfl_xor_sync(FULL, na, o); nb += __shfl_xor_sync(FULL, nb, o);
    }
    const float la = sqrtf(na), lb = sqrtf(nb);
    const float ca = fmaxf(la, 1e-12f), cb = fmaxf(lb, 1e-12f);
    const float inv = 1.f / (ca * cb);
    const float cosv = dt * inv;
    // d cos / d a = b/(|a||b|) - cos * a/|a|^2   (the clamp branch has zero norm-gradient)
    const float ka = la > 1e-12f ? cosv / na : 0.f;
    const float kb = lb > 1e-12f ? cosv / nb : 0.f;
    Row<VPL> ga, gb;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        ga.v[i] = g * (inv * b.v[i] - ka * a.v[i]);
        gb.v[i] = g * (inv * a.v[i] - kb * b.v[i]);
    }
    red_row<VPL>(dA, ra, d4, lane, ga);
    red_row<VPL>(dB, rb, d4, lane, gb);
}

}  // namespace jmac

using namespace jmac;

static int fill_transe(TransEParams& p, const float* const* ent, const float* const* rel, int n_layers, int d, const int64_t* h,
                       const int64_t* r, const int64_t* t, int64_t M, int64_t B, int64_t n_ent, int64_t n_rel, float margin) {
    JMAC_ARG(ent && rel && h && r && t);
    JMAC_ARG(n_layers >= 1 && n_layers <= MAX_LAYERS && dim_ok(d));
    JMAC_ARG(B > 0 && M >= 2 * B && (M % B) == 0 && M / B < (1 << 20));
    for (int l = 0; l < MAX_LAYERS; ++l) { p.ent[l] = p.rel[l] = nullptr; p.d_ent[l] = p.d_rel[l] = nullptr; }
    for (int l = 0; l < n_layers; ++l) {
        JMAC_ARG(ent[l] && rel[l]);
        p.ent[l] = ent[l]; p.rel[l] = rel[l];
    }
    p.n_layers = n_layers; p.d4 = d / 4; p.h = h; p.r = r; p.t = t; p.M = M; p.B = B; p.K1 = (int)(M / B); p.margin = margin;
    p.score = nullptr; p.loss = nullptr; p.gloss = nullptr;
    JMAC_ARG(n_ent > 0 && n_ent < (int64_t(1) << 31) && n_rel > 0 && n_rel < (int64_t(1) << 31));      // the kernels keep ids in 32 bits
    p.n_ent = n_ent; p.n_rel = n_rel; p.status = status_words() ? status_words() + JMAC_ST_TRIPLE : nullptr;
    p.coef = nullptr; p.smask = nullptr; p.dhr = nullptr; p.loss_part = nullptr;
    return 0;
}

static int bits_for_u(uint64_t n) {  // bits needed to represent values in [0, n]
    int b = 1;
    while ((uint64_t(1) << b) <= n) ++b;
    return b;
}

// occurrences per warp of the reduction: 32 on small batches (more warps), 128 on large ones (fewer carry rows)
static int reduce_chunk(int64_t n_occ) { return n_occ >= (int64_t(1) << 20) ? 128 : 32; }

struct TransEBwdLayout {
    size_t off_cover, off_ekeys, off_evals, off_ekeys2, off_evals2, off_rkeys, off_rvals, off_rkeys2, off_rvals2, off_carry, off_cub, total;
    size_t cub_bytes;
    int64_t n_eocc, n_rocc, e_chunks, r_chunks;
    int e_chunk, r_chunk;
};

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

static TransEBwdLayout transe_bwd_layout(int64_t M, int64_t B, int n_layers, int d) {
    TransEBwdLayout L;
    L.n_eocc = 2 * M + B;
    L.n_rocc = M + B;
    L.e_chunk = reduce_chunk(L.n_eocc);
    L.r_chunk = reduce_chunk(L.n_rocc);
    L.e_chunks = (L.n_eocc + L.e_chunk - 1) / L.e_chunk;
    L.r_chunks = (L.n_rocc + L.r_chunk - 1) / L.r_chunk;
    size_t o = 0;
    L.off_cover = o;  o += align256(sizeof(int) * (size_t)B);
    L.off_ekeys = o;  o += align256(4 * (size_t)L.n_eocc);
    L.off_evals = o;  o += align256(4 * (size_t)L.n_eocc);
    L.off_ekeys2 = o; o += align256(4 * (size_t)L.n_eocc);
    L.off_evals2 = o; o += align256(4 * (size_t)L.n_eocc);
    L.off_rkeys = o;  o += align256(4 * (size_t)L.n_rocc);
    L.off_rvals = o;  o += align256(4 * (size_t)L.n_rocc);
    L.off_rkeys2 = o; o += align256(4 * (size_t)L.n_rocc);
    L.off_rvals2 = o; o += align256(4 * (size_t)L.n_rocc);
    const int64_t max_chunks = L.e_chunks > L.r_chunks ? L.e_chunks : L.r_chunks;      // the carry rows are reused by the two reductions
    L.off_carry = o;  o += align256(sizeof(float) * (size_t)n_layers * (size_t)max_chunks * 2 * (size_t)d);
    size_t c1 = 0, c2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, c1, (const unsigned*)nullptr, (unsigned*)nullptr, (const unsigned*)nullptr, (unsigned*)nullptr,
                                    (int)L.n_eocc, 0, 32);
    cub::DeviceRadixSort::SortPairs(nullptr, c2, (const unsigned*)nullptr, (unsigned*)nullptr, (const unsigned*)nullptr, (unsigned*)nullptr,
                                    (int)L.n_rocc, 0, 32);
    L.cub_bytes = c1 > c2 ? c1 : c2;
    L.off_cub = o;    o += align256(L.cub_bytes);
    L.total = o;
    return L;
}

extern "C" size_t jmac_transe_record_bytes(int64_t M, int64_t B, int n_layers, int d) {
    if (M <= 0 || B <= 0 || n_layers <= 0 || !dim_ok(d)) return 0;
    // coef [L,M] f32 | mask [L,M,64|128 B] | dhr [L,B,d] f32 | per-block loss partials [L, ceil(B/8)] f32, each 256-byte aligned
    return align256(4 * (size_t)n_layers * M) + align256(transe_mask_bytes(d) * (size_t)n_layers * M) + align256(4 * (size_t)n_layers * B * d) +
           align256(4 * (size_t)n_layers * ((B + 7) / 8));
}

extern "C" size_t jmac_transe_bwd_workspace_bytes(int64_t M, int64_t B, int n_layers, int d) {
    if (M <= 0 || B <= 0 || n_layers <= 0 || !dim_ok(d) || 2 * M + B >= (int64_t(1) << 30)) return 0;
    return transe_bwd_layout(M, B, n_layers, d).total;
}

static void transe_record_ptrs(TransEParams& p, void* record, int d) {
    char* w = reinterpret_cast<char*>(record);
    p.coef = reinterpret_cast<float*>(w);  w += align256(4 * (size_t)p.n_layers * p.M);
    p.smask = w;                           w += align256(transe_mask_bytes(d) * (size_t)p.n_layers * p.M);
    p.dhr = reinterpret_cast<float*>(w);   w += align256(4 * (size_t)p.n_layers * p.B * (size_t)d);
    p.loss_part = reinterpret_cast<float*>(w);
}

extern "C" int jmac_transe_margin_fwd(const float* const* ent, const float* const* rel, int n_layers, int d, const int64_t* h,
                                      const int64_t* r, const int64_t* t, int64_t M, int64_t B, int64_t n_ent, int64_t n_rel,
                                      float margin, float* score, float* loss, void* record, void* stream) {
    TransEParams p;
    int rc = fill_transe(p, ent, rel, n_layers, d, h, r, t, M, B, n_ent, n_rel, margin);
    if (rc) return rc;
    JMAC_ARG(loss != nullptr);
    JMAC_ARG((reinterpret_cast<uintptr_t>(record) & 255) == 0);
    p.score = score; p.loss = loss;
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = 256;
    const int64_t blocks = (B + 7) / 8;
    float* part_tmp = nullptr;
    if (record) transe_record_ptrs(p, record, d);
    else {                                                   // no record buffer to host the loss partials: stream-ordered scratch
        JMAC_CUDA(cudaMallocAsync(&part_tmp, sizeof(float) * (size_t)n_layers * (size_t)blocks, st));
        p.loss_part = part_tmp;
    }
    JMAC_DISPATCH_VPL(d, {
        constexpr int U = VPL <= 2 ? 4 : 2;
        // recording variant: 2 rows per warp in registers, 24 warps / SM (measured against 4 rows with 16 warps or with spills: 2.11 vs
        // 2.52 / 2.36 ms at B = 65536, K = 64; 1.77 ms with the L2 prefetch, which supplies the memory-level parallelism)
        if (!record) transe_fwd_kernel<VPL, U, false, 3><<<(unsigned)blocks, threads, 0, st>>>(p);
        else transe_fwd_kernel<VPL, 2, true, 3><<<(unsigned)blocks, threads, 0, st>>>(p);
    });
    JMAC_LAUNCH_CHECK();
    transe_loss_kernel<<<n_layers, 256, 0, st>>>(p.loss_part, (int)blocks, margin, 1.f / ((float)B * (float)(p.K1 - 1)), loss);
    JMAC_LAUNCH_CHECK();
    if (part_tmp) JMAC_CUDA(cudaFreeAsync(part_tmp, st));
    return 0;
}

extern "C" int jmac_transe_margin_bwd(int n_layers, int d, const int64_t* h, const int64_t* r, const int64_t* t, int64_t M, int64_t B,
                                      int64_t n_ent, int64_t n_rel, const void* record, const float* gloss, float* const* d_ent,
                                      float* const* d_rel, void* workspace, size_t workspace_bytes, void* stream) {
    JMAC_ARG(h && r && t && record && gloss && d_ent && d_rel && workspace);
    JMAC_ARG(n_layers >= 1 && n_layers <= MAX_LAYERS && dim_ok(d));
    JMAC_ARG(B > 0 && M >= 2 * B && (M % B) == 0 && M / B < (1 << 20) && 2 * M + B < (int64_t(1) << 30));
    JMAC_ARG(n_ent > 0 && n_ent < (int64_t(1) << 31) - 1 && n_rel > 0 && n_rel < (int64_t(1) << 31) - 1);
    JMAC_ARG((reinterpret_cast<uintptr_t>(record) & 255) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0);
    const TransEBwdLayout L = transe_bwd_layout(M, B, n_layers, d);
    JMAC_ARG(workspace_bytes >= L.total);
    for (int l = 0; l < n_layers; ++l) JMAC_ARG(d_ent[l] && d_rel[l]);
    cudaStream_t st = (cudaStream_t)stream;
    char* w = reinterpret_cast<char*>(workspace);
    int* cover = reinterpret_cast<int*>(w + L.off_cover);
    unsigned* ekeys = reinterpret_cast<unsigned*>(w + L.off_ekeys);   unsigned* evals = reinterpret_cast<unsigned*>(w + L.off_evals);
    unsigned* ekeys2 = reinterpret_cast<unsigned*>(w + L.off_ekeys2); unsigned* evals2 = reinterpret_cast<unsigned*>(w + L.off_evals2);
    unsigned* rkeys = reinterpret_cast<unsigned*>(w + L.off_rkeys);   unsigned* rvals = reinterpret_cast<unsigned*>(w + L.off_rvals);
    unsigned* rkeys2 = reinterpret_cast<unsigned*>(w + L.off_rkeys2); unsigned* rvals2 = reinterpret_cast<unsigned*>(w + L.off_rvals2);
    const int K1 = (int)(M / B);
    transe_cover_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, st>>>(h, r, B, K1, cover);
    JMAC_LAUNCH_CHECK();
    transe_keys_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(h, r, t, M, B, n_ent, n_rel, cover, ekeys, evals, rkeys, rvals);
    JMAC_LAUNCH_CHECK();
    size_t cub_bytes = L.cub_bytes;
    JMAC_CUDA(cub::DeviceRadixSort::SortPairs(w + L.off_cub, cub_bytes, ekeys, ekeys2, evals, evals2, (int)L.n_eocc, 0, bits_for_u((uint64_t)n_ent), st));
    cub_bytes = L.cub_bytes;
    JMAC_CUDA(cub::DeviceRadixSort::SortPairs(w + L.off_cub, cub_bytes, rkeys, rkeys2, rvals, rvals2, (int)L.n_rocc, 0, bits_for_u((uint64_t)n_rel), st));
    ReduceParams q;
    q.d4 = d / 4; q.K1 = K1; q.M = M; q.B = B;
    {
        const char* rw = reinterpret_cast<const char*>(record);
        q.coef = reinterpret_cast<const float*>(rw);  rw += align256(4 * (size_t)n_layers * M);
        q.smask = rw;                                 rw += align256(transe_mask_bytes(d) * (size_t)n_layers * M);
        q.dhr = reinterpret_cast<const float*>(rw);
    }
    q.gloss = gloss;
    q.inv_pairs = 1.f / ((float)B * (float)(K1 - 1));
    q.carry = reinterpret_cast<float*>(w + L.off_carry);
    for (int l = 0; l < MAX_LAYERS; ++l) q.out[l] = nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        // pass 0: entity tables (tails, uncovered heads, one dense record per positive); pass 1: relation tables
        q.keys = pass == 0 ? ekeys2 : rkeys2;
        q.vals = pass == 0 ? evals2 : rvals2;
        q.n_occ = pass == 0 ? L.n_eocc : L.n_rocc;
        q.sentinel = (unsigned)(pass == 0 ? n_ent : n_rel);
        q.chunk = pass == 0 ? L.e_chunk : L.r_chunk;
        q.n_chunks = pass == 0 ? L.e_chunks : L.r_chunks;
        for (int l = 0; l < n_layers; ++l) q.out[l] = pass == 0 ? d_ent[l] : d_rel[l];
        const unsigned blocks = (unsigned)((q.n_chunks * 32 + 255) / 256);
        JMAC_DISPATCH_VPL(d, {
            if (n_layers == 1) transe_reduce_kernel<VPL, 1><<<blocks, 256, 0, st>>>(q);
            else if (n_layers == 2) transe_reduce_kernel<VPL, 2><<<blocks, 256, 0, st>>>(q);
            else if (n_layers == 3) transe_reduce_kernel<VPL, 3><<<blocks, 256, 0, st>>>(q);
            else transe_reduce_kernel<VPL, 4><<<blocks, 256, 0, st>>>(q);
        });
        JMAC_LAUNCH_CHECK();
        JMAC_DISPATCH_VPL(d, { transe_fixup_kernel<VPL><<<blocks, 256, 0, st>>>(q, n_layers); });
        JMAC_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int jmac_pair_cos_fwd(const float* A, const float* B, int d, const int64_t* ia, const int64_t* ib, int64_t M, int64_t nA,
                                 int64_t nB, float* cosv, void* stream) {
    JMAC_ARG(A && B && dim_ok(d) && M >= 0 && nA > 0 && nB > 0);
    if (M == 0) return 0;
    JMAC_ARG(ia && ib && cosv);
    const int64_t blocks = (M + 7) / 8;
    int* status = status_words() ? status_words() + JMAC_ST_PAIR : nullptr;
    JMAC_DISPATCH_VPL(d, { pair_cos_fwd_kernel<VPL><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A, B, d / 4, ia, ib, M, nA, nB, status, cosv); });
    JMAC_LAUNCH_CHECK();
    return 0;
}

extern "C" int jmac_pair_cos_bwd(const float* A, const float* B, int d, const int64_t* ia, const int64_t* ib, int64_t M, int64_t nA,
                                 int64_t nB, const float* gcos, float* dA, float* dB, void* stream) {
    JMAC_ARG(A && B && dim_ok(d) && M >= 0 && nA > 0 && nB > 0);
    if (M == 0) return 0;
    JMAC_ARG(ia && ib && gcos && dA && dB);
    const int64_t blocks = (M + 7) / 8;
    int* status = status_words() ? status_words() + JMAC_ST_PAIR : nullptr;
    JMAC_DISPATCH_VPL(d, { pair_cos_bwd_kernel<VPL><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A, B, d / 4, ia, ib, M, nA, nB, status, gcos, dA, dB); });
    JMAC_LAUNCH_CHECK();
    return 0;
}
