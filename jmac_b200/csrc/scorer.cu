// Completion scorer (TransE-L1 margin loss) and pair-cosine kernels.
//
// Reference being replaced: the scorer part of JMAC.completion_loss (src/jmac_model.py:345-374) on
// the batch layout of train.py:347-352, and the F.normalize + row-dot of alignment_loss_simple /
// alignment_loss (src/jmac_model.py:243-248, 268-287).
//
// One warp per positive b walks its K+1 triples m = k*B + b.  The (h, r) rows are reloaded only
// when the index changes (they are tiled K+1 times in the reference layout), so the HBM traffic is
// the algorithmic (K+3) rows per positive per layer.  Backward skips hinge-inactive negatives
// entirely and scatters gradient rows with 128-bit vector REDs (the destination tables are dense
// autograd gradients; collisions between random negatives are rare).
#include <math.h>

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
};

__global__ void fill_kernel(float* p, int n, float v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

template <int VPL>
__device__ __forceinline__ float l1_diff(const Row<VPL>& a, const Row<VPL>& b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
        s += fabsf(a.v[i].x - b.v[i].x) + fabsf(a.v[i].y - b.v[i].y) + fabsf(a.v[i].z - b.v[i].z) + fabsf(a.v[i].w - b.v[i].w);
    return s;
}

template <int VPL, int U>
__global__ void __launch_bounds__(256, 3) transe_fwd_kernel(TransEParams p) {
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
            int hprev = -1, rprev = -1;                      // ids fit 31 bits (checked by the host): 32-bit keeps the kernel at 3 CTAs / SM
            Row<VPL> hr;
            hr.zero();
            float pos = 0.f;
            for (int kb = 0; kb < p.K1; kb += 32) {
                const int cnt = min(32, p.K1 - kb);
                int hv = 0, rv = 0, tv = 0;
                if (lane < cnt) {
                    const int64_t m = (int64_t)(kb + lane) * p.B + b;
                    hv = (int)safe_index(__ldg(p.h + m), p.n_ent, p.status);
                    rv = (int)safe_index(__ldg(p.r + m), p.n_rel, p.status);
                    tv = (int)safe_index(__ldg(p.t + m), p.n_ent, p.status);
                }
                float sc_keep = 0.f;
                for (int k = 0; k < cnt; k += U) {
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
                            const float sc = warp_sum(l1_diff<VPL>(hr, tr[u]));
                            if (lane == k + u) sc_keep = sc;
                            if (kb + k + u == 0) pos = sc;
                            else lsum += fmaxf(pos - sc, -p.margin);
                        }
                    }
                }
                if (lane < cnt) p.score[(int64_t)l * p.M + (int64_t)(kb + lane) * p.B + b] = sc_keep;
            }
        }
        if (lane == 0) block_loss[l][wib] = lsum;
    }
    __syncthreads();
    if (threadIdx.x < p.n_layers) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += block_loss[threadIdx.x][w];
        const float inv = 1.f / ((float)p.B * (float)(p.K1 - 1));
        atomicAdd(p.loss + threadIdx.x, s * inv);
    }
}

template <int VPL>
__device__ __forceinline__ void red_row(float* base, int64_t row, int d4, int lane, const Row<VPL>& v) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = i * 32 + lane;
        if (c < d4) red_add_v4(base + (row * d4 + c) * 4, v.v[i]);
    }
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f ? 1.f : 0.f) - (v < 0.f ? 1.f : 0.f); }

template <int VPL>
__global__ void __launch_bounds__(256) transe_bwd_kernel(TransEParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= p.B) return;
    const int d4 = p.d4;
    for (int l = 0; l < p.n_layers; ++l) {
        const float* ent = p.ent[l];
        const float* rel = p.rel[l];
        const float* score = p.score + (int64_t)l * p.M;
        const float gs = p.gloss[l] / ((float)p.B * (float)(p.K1 - 1));
        const float pos = score[b];
        // pass 1: number of active hinge terms of this positive
        int act = 0;
        for (int k = 1 + lane; k < p.K1; k += 32) act += (pos - score[(int64_t)k * p.B + b] > -p.margin) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) act += __shfl_xor_sync(FULL, act, o);
        int64_t hprev = -1, rprev = -1;
        Row<VPL> hr, dhr;
        hr.zero(); dhr.zero();
        bool dirty = false;
        for (int k = 0; k < p.K1; ++k) {
            const int64_t m = (int64_t)k * p.B + b;
            float ds;
            if (k == 0) ds = gs * (float)act;
            else ds = (pos - score[m] > -p.margin) ? -gs : 0.f;
            if (ds == 0.f) continue;  // warp-uniform: inactive negatives cost no row traffic
            const int64_t hh = safe_index(__ldg(p.h + m), p.n_ent, p.status), rr = safe_index(__ldg(p.r + m), p.n_rel, p.status),
                          tt = safe_index(__ldg(p.t + m), p.n_ent, p.status);
            Row<VPL> tr;
            load_row<VPL, true>(tr, ent, tt, d4, lane);
            if (hh != hprev || rr != rprev) {
                if (dirty) {
                    red_row<VPL>(p.d_ent[l], hprev, d4, lane, dhr);
                    red_row<VPL>(p.d_rel[l], rprev, d4, lane, dhr);
                    dhr.zero();
                    dirty = false;
                }
                Row<VPL> hrow, rrow;
                load_row<VPL, true>(hrow, ent, hh, d4, lane);
                load_row<VPL, false>(rrow, rel, rr, d4, lane);
#pragma unroll
                for (int i = 0; i < VPL; ++i) hr.v[i] = hrow.v[i] + rrow.v[i];
                hprev = hh; rprev = rr;
            }
            Row<VPL> dt;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const float4 s4 = make_float4(sgn(hr.v[i].x - tr.v[i].x), sgn(hr.v[i].y - tr.v[i].y), sgn(hr.v[i].z - tr.v[i].z), sgn(hr.v[i].w - tr.v[i].w));
                fma4(dhr.v[i], ds, s4);
                dt.v[i] = (-ds) * s4;
            }
            dirty = true;
            red_row<VPL>(p.d_ent[l], tt, d4, lane, dt);
        }
        if (dirty) {
            red_row<VPL>(p.d_ent[l], hprev, d4, lane, dhr);
            red_row<VPL>(p.d_rel[l], rprev, d4, lane, dhr);
        }
    }
}

// ------------------------------------------------------------------------------------------------
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
    return 0;
}

extern "C" int jmac_transe_margin_fwd(const float* const* ent, const float* const* rel, int n_layers, int d, const int64_t* h,
                                      const int64_t* r, const int64_t* t, int64_t M, int64_t B, int64_t n_ent, int64_t n_rel,
                                      float margin, float* score, float* loss, void* stream) {
    TransEParams p;
    int rc = fill_transe(p, ent, rel, n_layers, d, h, r, t, M, B, n_ent, n_rel, margin);
    if (rc) return rc;
    JMAC_ARG(score && loss);
    p.score = score; p.loss = loss;
    cudaStream_t st = (cudaStream_t)stream;
    fill_kernel<<<1, 32, 0, st>>>(loss, n_layers, margin);
    JMAC_LAUNCH_CHECK();
    const int threads = 256;
    const int64_t blocks = (B + 7) / 8;
    JMAC_DISPATCH_VPL(d, {
        constexpr int U = VPL <= 2 ? 4 : 2;
        transe_fwd_kernel<VPL, U><<<(unsigned)blocks, threads, 0, st>>>(p);
    });
    JMAC_LAUNCH_CHECK();
    return 0;
}

extern "C" int jmac_transe_margin_bwd(const float* const* ent, const float* const* rel, int n_layers, int d, const int64_t* h,
                                      const int64_t* r, const int64_t* t, int64_t M, int64_t B, int64_t n_ent, int64_t n_rel, float margin,
                                      const float* score, const float* gloss, float* const* d_ent, float* const* d_rel, void* stream) {
    TransEParams p;
    int rc = fill_transe(p, ent, rel, n_layers, d, h, r, t, M, B, n_ent, n_rel, margin);
    if (rc) return rc;
    JMAC_ARG(score && gloss && d_ent && d_rel);
    for (int l = 0; l < n_layers; ++l) {
        JMAC_ARG(d_ent[l] && d_rel[l]);
        p.d_ent[l] = d_ent[l]; p.d_rel[l] = d_rel[l];
    }
    p.score = const_cast<float*>(score); p.gloss = gloss;
    const int threads = 256;
    const int64_t blocks = (B + 7) / 8;
    JMAC_DISPATCH_VPL(d, { transe_bwd_kernel<VPL><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(p); });
    JMAC_LAUNCH_CHECK();
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
