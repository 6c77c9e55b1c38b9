// Link prediction: L1 "cdist" of JMAC.forward_linkpred (src/jmac_model.py:301-313) and the ranking
// loop of CompletionEvaluator.test (src/validate.py:50-70: filter with -1e6, full descending sort,
// position of the target) done as rank-by-counting, without materialising or sorting [B,N].
//
// This is B*N*d abs-diff-adds on CUDA cores with both operands cache resident (HBM bytes are only
// (B+N)*d*4 per layer): FP32-pipe bound, not HBM bound, and not expressible on tensor cores.
// Tile: 128 queries x 128 candidates per CTA, 8x8 register micro-tile per thread, the feature
// dimension staged through shared memory 16 features at a time (transposed, padded).
// Every distance -- tile kernel, target prologue, filter epilogue -- is ONE running fp32 sum over the
// features of all layers in the same order, so equal inputs give bit-equal distances.
#include <math.h>

#include "common.cuh"

namespace jmac {

constexpr int LP_MAX_LAYERS = 4;
constexpr int TQ = 128, TC = 128, DK = 16, PADW = TQ + 4;

struct L1Params {
    const float* ent[LP_MAX_LAYERS];
    const float* rel[LP_MAX_LAYERS];
    int n_layers, d;
    const int64_t *e_idx, *r_idx, *t_idx;
    int64_t B, N, n_rel;
    int* status;          // out-of-range index flag
    float sign;
    float* dist;          // [B,N] or null
    const float* tdist;   // [B] or null
    int* rank;            // [B] or null
};

// sequential reference-order distance of query b to candidate n (used by prologue / filter)
__device__ __forceinline__ float l1_seq(const L1Params& p, int64_t e, int64_t r, int64_t n) {
    float acc = 0.f;                     // one running sum over the features of all layers, in layer order
    for (int l = 0; l < p.n_layers; ++l) {
        const float* er = p.ent[l] + e * p.d;
        const float* rr = p.rel[l] + r * p.d;
        const float* cn = p.ent[l] + n * p.d;
        for (int k = 0; k < p.d; ++k) {
            const float q = __fadd_rn(er[k], p.sign * rr[k]);
            acc = __fadd_rn(acc, fabsf(__fadd_rn(q, -cn[k])));
        }
    }
    return acc;
}

__global__ void l1_target_kernel(L1Params p, float* __restrict__ tdist, int* __restrict__ rank) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    tdist[b] = l1_seq(p, safe_index(p.e_idx[b], p.N, p.status), safe_index(p.r_idx[b], p.n_rel, p.status), safe_index(p.t_idx[b], p.N, p.status));
    rank[b] = 1;
}

__global__ void l1_filter_kernel(L1Params p, const int64_t* __restrict__ filt_ptr, const int64_t* __restrict__ filt_idx,
                                 const float* __restrict__ tdist, int* __restrict__ rank) {
    // one warp per query, lanes over its filter entries
    const int lane = threadIdx.x & 31;
    const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (b >= p.B) return;
    const int64_t e = safe_index(p.e_idx[b], p.N, p.status), r = safe_index(p.r_idx[b], p.n_rel, p.status), t = safe_index(p.t_idx[b], p.N, p.status);
    const float td = tdist[b];
    int cnt = 0;
    for (int64_t f = filt_ptr[b] + lane; f < filt_ptr[b + 1]; f += 32) {
        const int64_t n = safe_index(filt_idx[f], p.N, p.status);
        if (n == t || n < 0 || n >= p.N) continue;
        if (l1_seq(p, e, r, n) < td) ++cnt;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
    if (lane == 0 && cnt) atomicSub(rank + b, cnt);
}

template <bool RANK>
__global__ void __launch_bounds__(256, 2) l1_tile_kernel(L1Params p) {
    __shared__ float Qs[DK][PADW];
    __shared__ float Cs[DK][PADW];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t q0 = (int64_t)blockIdx.y * TQ, c0 = (int64_t)blockIdx.x * TC;
    // One accumulator per (query, candidate) of the 8x8 micro-tile, running over the features of all layers (64 registers; a
    // per-layer sum next to a total would make it 128 and halve the occupancy), two CTAs per SM.
    float total[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) total[i][j] = 0.f;

    // loader mapping: 128 rows x 4 float4 per chunk = 512 float4, two per thread.  The rows a thread loads never change: their table
    // rows are looked up once (the index -> row -> data chain used to be walked at every stage), and the loads of stage s + 1 are issued
    // into registers before the arithmetic of stage s, so that their latency hides behind it.
    const int lrow = threadIdx.x >> 2, lk4 = threadIdx.x & 3;
    int64_t erow[2], rrow[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int64_t b = q0 + lrow + half * 64;
        erow[half] = b < p.B ? safe_index(p.e_idx[b], p.N, p.status) : -1;
        rrow[half] = b < p.B ? safe_index(p.r_idx[b], p.n_rel, p.status) : 0;
    }
    const int spl = (p.d + DK - 1) / DK;                 // stages per layer
    const int n_stages = p.n_layers * spl;
    float4 qn[2], cn[2];
    auto fetch = [&](int s) {
        const int l = s / spl, kk = (s - l * spl) * DK + lk4 * 4;
        const float* ent = p.ent[l];
        const float* rel = p.rel[l];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = lrow + half * 64;
            qn[half] = f4_zero();
            cn[half] = f4_zero();
            if (kk < p.d) {
                if (erow[half] >= 0) {
                    const float4 ev = __ldg(reinterpret_cast<const float4*>(ent + erow[half] * p.d + kk));
                    const float4 rv = __ldg(reinterpret_cast<const float4*>(rel + rrow[half] * p.d + kk));
                    qn[half] = make_float4(__fadd_rn(ev.x, p.sign * rv.x), __fadd_rn(ev.y, p.sign * rv.y), __fadd_rn(ev.z, p.sign * rv.z),
                                           __fadd_rn(ev.w, p.sign * rv.w));
                }
                if (c0 + row < p.N) cn[half] = __ldg(reinterpret_cast<const float4*>(ent + (c0 + row) * p.d + kk));
            }
        }
    };
    fetch(0);
    for (int s = 0; s < n_stages; ++s) {
        __syncthreads();                                 // the previous stage's arithmetic has read the tiles
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int row = lrow + half * 64;
            Qs[lk4 * 4 + 0][row] = qn[half].x; Qs[lk4 * 4 + 1][row] = qn[half].y; Qs[lk4 * 4 + 2][row] = qn[half].z; Qs[lk4 * 4 + 3][row] = qn[half].w;
            Cs[lk4 * 4 + 0][row] = cn[half].x; Cs[lk4 * 4 + 1][row] = cn[half].y; Cs[lk4 * 4 + 2][row] = cn[half].z; Cs[lk4 * 4 + 3][row] = cn[half].w;
        }
        __syncthreads();
        if (s + 1 < n_stages) fetch(s + 1);
        const int kmax = min(DK, p.d - (s % spl) * DK);
        for (int k = 0; k < kmax; ++k) {
            // micro-tile rows/cols {4t..4t+3} and {64+4t..64+4t+3}: conflict-free LDS.128
            const float4 qa = *reinterpret_cast<const float4*>(&Qs[k][ty * 4]);
            const float4 qb = *reinterpret_cast<const float4*>(&Qs[k][64 + ty * 4]);
            const float4 ca = *reinterpret_cast<const float4*>(&Cs[k][tx * 4]);
            const float4 cb = *reinterpret_cast<const float4*>(&Cs[k][64 + tx * 4]);
            const float q[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
            const float c[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) total[i][j] = __fadd_rn(total[i][j], fabsf(__fadd_rn(q[i], -c[j])));
        }
    }

    if (!RANK) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t b = q0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            if (b >= p.B) continue;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t n = c0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                if (n < p.N) p.dist[b * p.N + n] = total[i][j];
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t b = q0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
            int cnt = 0;
            if (b < p.B) {
                const float td = p.tdist[b];
                const int64_t t = p.t_idx[b];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int64_t n = c0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                    if (n < p.N && n != t && total[i][j] < td) ++cnt;
                }
            }
            // the 16 threads sharing this query row are lanes (ty&1)*16 .. +15 of one warp
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
            if (tx == 0 && cnt && b < p.B) atomicAdd(p.rank + b, cnt);
        }
    }
}

}  // namespace jmac

using namespace jmac;

static int fill_l1(L1Params& p, const float* const* ent, const float* const* rel, int n_layers, int d, const int64_t* e_idx,
                   const int64_t* r_idx, int64_t B, int64_t N, int64_t n_rel, int pred_head) {
    JMAC_ARG(ent && rel && e_idx && r_idx);
    JMAC_ARG(n_layers >= 1 && n_layers <= LP_MAX_LAYERS && d >= 4 && (d % 4) == 0 && B > 0 && N > 0 && n_rel > 0);
    p.n_rel = n_rel;
    p.status = status_words() ? status_words() + JMAC_ST_LINKPRED : nullptr;
    for (int l = 0; l < LP_MAX_LAYERS; ++l) p.ent[l] = p.rel[l] = nullptr;
    for (int l = 0; l < n_layers; ++l) {
        JMAC_ARG(ent[l] && rel[l]);
        p.ent[l] = ent[l]; p.rel[l] = rel[l];
    }
    p.n_layers = n_layers; p.d = d; p.e_idx = e_idx; p.r_idx = r_idx; p.t_idx = nullptr; p.B = B; p.N = N;
    p.sign = pred_head ? -1.f : 1.f;
    p.dist = nullptr; p.tdist = nullptr; p.rank = nullptr;
    return 0;
}

extern "C" int jmac_l1_dist(const float* const* ent, const float* const* rel, int n_layers, int d, const int64_t* e_idx,
                            const int64_t* r_idx, int64_t B, int64_t N, int64_t n_rel, int pred_head, float* dist, void* stream) {
    L1Params p;
    int rc = fill_l1(p, ent, rel, n_layers, d, e_idx, r_idx, B, N, n_rel, pred_head);
    if (rc) return rc;
    JMAC_ARG(dist != nullptr);
    p.dist = dist;
    dim3 grid((unsigned)((N + TC - 1) / TC), (unsigned)((B + TQ - 1) / TQ));
    l1_tile_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    JMAC_LAUNCH_CHECK();
    return 0;
}

extern "C" int jmac_l1_rank(const float* const* ent, const float* const* rel, int n_layers, int d, const int64_t* e_idx,
                            const int64_t* r_idx, const int64_t* t_idx, int64_t B, int64_t N, int64_t n_rel, int pred_head,
                            const int64_t* filt_ptr, const int64_t* filt_idx, int32_t* rank, void* stream) {
    L1Params p;
    int rc = fill_l1(p, ent, rel, n_layers, d, e_idx, r_idx, B, N, n_rel, pred_head);
    if (rc) return rc;
    JMAC_ARG(t_idx && rank);
    JMAC_ARG((filt_ptr == nullptr) == (filt_idx == nullptr) || filt_ptr != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    p.t_idx = t_idx;
    float* tdist = nullptr;
    JMAC_CUDA(cudaMallocAsync(&tdist, sizeof(float) * (size_t)B, st));
    l1_target_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(p, tdist, rank);
    JMAC_LAUNCH_CHECK();
    p.tdist = tdist; p.rank = rank;
    dim3 grid((unsigned)((N + TC - 1) / TC), (unsigned)((B + TQ - 1) / TQ));
    l1_tile_kernel<true><<<grid, 256, 0, st>>>(p);
    JMAC_LAUNCH_CHECK();
    if (filt_ptr && filt_idx) {
        l1_filter_kernel<<<(unsigned)((B * 32 + 255) / 256), 256, 0, st>>>(p, filt_ptr, filt_idx, tdist, rank);
        JMAC_LAUNCH_CHECK();
    }
    JMAC_CUDA(cudaFreeAsync(tdist, st));
    return 0;
}
