// Backward of the edge phase (autograd of src/jmac_model.py:56-89 + message_passing.py:24-28, which
// the reference gets from index_put_(accumulate) atomics and E-row GEMM transposes).
//
// Forward saves, per CSR position, the attention logit s_e (4 B) and the sign mask of
// z_e = P[i]+Q[j]-S[t] (1 bit per feature: 32 B at d=256).  With those, three atomics-free segmented
// passes over the three edge orderings of the plan need ONE gathered node row per edge each:
//   dst pass : per destination i -> alpha_e = exp(s_e - lse_i), u_e = G[i].(X[j]-R'[t]),
//              ds_e = alpha_e (sqrt(deg_i) u_e - G[i].agg[i]);  mu_i = sum_e ds_e lrelu'(z_e);
//              dP[i] = a (.) mu_i;  record {w_e = alpha_e sqrt(deg_i), ds_e}.       gathers X[j]
//   src pass : per source j  -> dXe[j] = sum w_e G[i],  nu_j = sum ds_e lrelu'(z_e), dQ[j] = a (.) nu_j.
//                                                                                    gathers G[i]
//   rel pass : per relation t -> dR'[t] = -sum w_e G[i], dS[t] = -a (.) omega_t.      gathers G[i]
// The sign mask is what lets every pass rebuild d z_e = ds_e a (.) lrelu'(z_e) without re-gathering
// P[i], Q[j], S[t].  The gradient of a_att needs lrelu(z_e) itself; LeakyReLU is positively
// homogeneous (lrelu(z) = z lrelu'(z)), so
//     d a = sum_e ds_e lrelu(z_e) = sum_i P[i] (.) mu_i + sum_j Q[j] (.) nu_j - sum_t S[t] (.) omega_t,
// three node-level terms that each pass accumulates at segment end (one extra row read per segment).
// Algorithmic HBM bytes per edge (d=256): dst 1 KiB + 56, src 1 KiB + 48, rel 1 KiB + 48.
#include <math.h>

#include "common.cuh"

namespace jmac {

struct EdgeBwdDstParams {
    const Item* items;
    int n_items;
    int E;
    const int* csr_src;
    const int* csr_type;
    const float* sqrt_deg;
    const float *P, *X, *Rp, *a;
    float slope;
    int d4;
    const float *agg, *lse, *G;
    float* dP;
    float2* ew;          // [E] (w_e, ds_e)                      out
    const void* mask;    // [E][32] lanes x mask_t               in (forward)
    const float* logit;  // [E]                                  in (forward)
    float* da;           // [d] accumulated
    float* gsum;         // [d] accumulated column sums of G over the destinations (null: not wanted): the gradient of the self-loop
                         // relation row that the fused forward subtracts from every node (T = agg + X - R'[loop])
    float* partial;      // [n_slots, d]
};

// block-level reduction of the per-warp d a_att contributions, then one vector RED per feature per block
template <int VPL, int NW>
__device__ __forceinline__ void flush_da_nw(const Row<VPL>& da_acc, float* da, int d4) {
    __shared__ float4 red[NW][VPL * 32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < VPL; ++i) red[wib][i * 32 + lane] = da_acc.v[i];
    __syncthreads();
    for (int c4 = threadIdx.x; c4 < d4; c4 += blockDim.x) {
        float4 s = red[0][c4];
        for (int w = 1; w < NW; ++w) s = s + red[w][c4];
        red_add_v4(da + 4 * c4, s);
    }
}
template <int VPL>
__device__ __forceinline__ void flush_da(const Row<VPL>& da_acc, float* da, int d4) { flush_da_nw<VPL, 8>(da_acc, da, d4); }
template <int VPL>
__device__ __forceinline__ void flush_da16(const Row<VPL>& da_acc, float* da, int d4) { flush_da_nw<VPL, RING_WARPS>(da_acc, da, d4); }

template <int VPL, int U>
__global__ void __launch_bounds__(256) edge_bwd_dst_kernel(EdgeBwdDstParams p) {
    typedef typename MaskT<VPL>::type mask_t;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(FULL, (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), 0);   // warp-uniform for the compiler
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int d4 = p.d4;
    Row<VPL> av, da_acc, gs_acc;
    load_row<VPL, false>(av, p.a, 0, d4, lane);
    da_acc.zero();
    gs_acc.zero();
    const mask_t* maskp = reinterpret_cast<const mask_t*>(p.mask);
    const float hi = 1.f - p.slope;

    for (int it = warp; it < p.n_items; it += nwarps) {
        const Item item = p.items[it];
        Row<VPL> gi, mu;          // mu: sum over edges with z > 0 of ds_e, per feature
        load_row<VPL, true>(gi, p.G, item.seg, d4, lane);
        if (p.gsum && (it == 0 || p.items[it - 1].seg != item.seg)) {       // every destination once (the first item of its segment)
#pragma unroll
            for (int i = 0; i < VPL; ++i) gs_acc.v[i] = gs_acc.v[i] + gi.v[i];
        }
        mu.zero();
        float gdot = 0.f, ds_sum = 0.f;
        {
            Row<VPL> ai;
            load_row<VPL, true>(ai, p.agg, item.seg, d4, lane);
#pragma unroll
            for (int i = 0; i < VPL; ++i) gdot += dot4(gi.v[i], ai.v[i]);
            gdot = warp_sum(gdot);
        }
        const float lse = p.lse[item.seg];
        const float c = p.sqrt_deg[item.seg];
        for (int base = item.beg; base < item.end; base += 32) {
            const int cnt = min(32, item.end - base);
            int jv = 0, tv = 0;
            float sv = 0.f;
            if (lane < cnt) {
                jv = __ldg(p.csr_src + base + lane);
                tv = __ldg(p.csr_type + base + lane);
                sv = __ldg(p.logit + base + lane);
            }
            float w_keep = 0.f, ds_keep = 0.f;      // lane k keeps the record of edge base+k
            for (int k = 0; k < cnt; k += U) {
                Row<VPL> x[U], r[U];
                unsigned bits[U];
                float sc[U], uu[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int kk = min(k + u, cnt - 1);     // clamp: tail edges are recomputed, never recorded twice
                    const int j = __shfl_sync(FULL, jv, kk);
                    const int t = __shfl_sync(FULL, tv, kk);
                    sc[u] = __shfl_sync(FULL, sv, kk);
                    load_row<VPL, true>(x[u], p.X, j, d4, lane);
                    load_row<VPL, false>(r[u], p.Rp, t, d4, lane);
                    bits[u] = (unsigned)maskp[(size_t)(base + kk) * 32 + lane];
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uu[u] = 0.f;
#pragma unroll
                    for (int i = 0; i < VPL; ++i) uu[u] += dot4(gi.v[i], x[u].v[i] - r[u].v[i]);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int u = 0; u < U; ++u) uu[u] += __shfl_xor_sync(FULL, uu[u], o);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (k + u < cnt) {                      // warp-uniform
                        const float alpha = __expf(sc[u] - lse);
                        const float w = c * alpha;
                        const float ds = alpha * (c * uu[u] - gdot);
                        ds_sum += ds;
#pragma unroll
                        for (int i = 0; i < VPL; ++i) {
                            const unsigned b = bits[u] >> (4 * i);
                            if (b & 1u) mu.v[i].x += ds;
                            if (b & 2u) mu.v[i].y += ds;
                            if (b & 4u) mu.v[i].z += ds;
                            if (b & 8u) mu.v[i].w += ds;
                        }
                        if (lane == k + u) { w_keep = w; ds_keep = ds; }
                    }
                }
            }
            if (lane < cnt) p.ew[base + lane] = make_float2(w_keep, ds_keep);
        }
        // mu' = slope * sum ds + (1 - slope) * sum_{z>0} ds;  dP = a (.) mu';  d a += P[i] (.) mu'
        const float lo = p.slope * ds_sum;
        Row<VPL> pi;
        load_row<VPL, true>(pi, p.P, item.seg, d4, lane);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            mu.v[i].x = fmaf(hi, mu.v[i].x, lo); mu.v[i].y = fmaf(hi, mu.v[i].y, lo);
            mu.v[i].z = fmaf(hi, mu.v[i].z, lo); mu.v[i].w = fmaf(hi, mu.v[i].w, lo);
            da_acc.v[i].x = fmaf(pi.v[i].x, mu.v[i].x, da_acc.v[i].x); da_acc.v[i].y = fmaf(pi.v[i].y, mu.v[i].y, da_acc.v[i].y);
            da_acc.v[i].z = fmaf(pi.v[i].z, mu.v[i].z, da_acc.v[i].z); da_acc.v[i].w = fmaf(pi.v[i].w, mu.v[i].w, da_acc.v[i].w);
            mu.v[i].x *= av.v[i].x; mu.v[i].y *= av.v[i].y; mu.v[i].z *= av.v[i].z; mu.v[i].w *= av.v[i].w;
        }
        if (item.slot < 0) store_row<VPL>(mu, p.dP, item.seg, d4, lane);
        else store_row<VPL>(mu, p.partial, item.slot, d4, lane);
    }
    flush_da<VPL>(da_acc, p.da, d4);
    if (p.gsum) {
        __syncthreads();
        flush_da<VPL>(gs_acc, p.gsum, d4);
    }
}

// Sum of the partial rows of split segments: out[seg] = sum_k partial[first + k]; `rows` rows per slot.
// One CTA of 8 warps per split: warp w adds the slots w, w + 8, ..., the eight warp sums are combined in a fixed order.
template <int VPL>
__global__ void __launch_bounds__(256) sum_partials_kernel(const Split* __restrict__ splits, int n_splits, const float* __restrict__ partial,
                                                           int d4, int rows, float* __restrict__ out0, float* __restrict__ out1) {
    __shared__ float4 s_acc[8][VPL * 32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if ((int)blockIdx.x >= n_splits) return;
    const Split sp = splits[blockIdx.x];
    for (int r = 0; r < rows; ++r) {
        Row<VPL> acc;
        acc.zero();
        for (int k = wib; k < sp.count; k += 8) {
            Row<VPL> part;
            load_row<VPL, true>(part, partial, (int64_t)(sp.first + k) * rows + r, d4, lane);
#pragma unroll
            for (int i = 0; i < VPL; ++i) acc.v[i] = acc.v[i] + part.v[i];
        }
        __syncthreads();                                   // previous round's readers are done with s_acc
#pragma unroll
        for (int i = 0; i < VPL; ++i) s_acc[wib][i * 32 + lane] = acc.v[i];
        __syncthreads();
        if (wib == 0) {
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                float4 t = s_acc[0][i * 32 + lane];
#pragma unroll
                for (int w = 1; w < 8; ++w) t = t + s_acc[w][i * 32 + lane];
                acc.v[i] = t;
            }
            store_row<VPL>(acc, r == 0 ? out0 : out1, sp.seg, d4, lane);
        }
    }
}

struct EdgeBwdGatherParams {
    const Item* items;
    int n_items;
    const int* pos;    // ordering position -> CSR position
    const int* dst;    // ordering position -> destination node
    const float* G;
    const float2* ew;
    const void* mask;
    const float* a;
    const float* Z;    // Q (src pass, rows = sources) or S (rel pass, rows = relations): the d a_att term
    float slope;
    int d4;
    float* out_w;      // sum w_e G[i]      (dXe for the src pass, dR' for the rel pass)
    float* out_z;      // sum d z_e         (dQ  for the src pass, dS  for the rel pass)
    float* da;         // [d] accumulated
    float* partial;    // [n_slots, 2, d]   (src pass only)
};

// MODE 0: source pass (rows written / partials); MODE 1: relation pass (negated; split segments RED
// into the zeroed tables).
template <int VPL, int MODE, int U>
__global__ void __launch_bounds__(256) edge_bwd_gather_kernel(EdgeBwdGatherParams p) {
    typedef typename MaskT<VPL>::type mask_t;
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(FULL, (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), 0);   // warp-uniform for the compiler
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int d4 = p.d4;
    Row<VPL> av, da_acc;
    load_row<VPL, false>(av, p.a, 0, d4, lane);
    da_acc.zero();
    const mask_t* maskp = reinterpret_cast<const mask_t*>(p.mask);
    const float sign = MODE == 0 ? 1.f : -1.f;

    for (int it = warp; it < p.n_items; it += nwarps) {
        const Item item = p.items[it];
        Row<VPL> accw, accz;  // accz holds sum over edges with a positive z of ds_e (per feature)
        accw.zero();
        accz.zero();
        float ds_sum = 0.f;
        for (int base = item.beg; base < item.end; base += 32) {
            const int cnt = min(32, item.end - base);
            int pv = 0, iv = 0;
            float2 e2 = make_float2(0.f, 0.f);
            if (lane < cnt) {
                pv = __ldg(p.pos + base + lane);
                iv = __ldg(p.dst + base + lane);
                e2 = __ldg(p.ew + pv);
            }
            for (int k = 0; k < cnt; k += U) {
                Row<VPL> g[U];
                unsigned bits[U];
                float w[U], ds[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int kk = min(k + u, cnt - 1);
                    const int i = __shfl_sync(FULL, iv, kk);
                    const int pp = __shfl_sync(FULL, pv, kk);
                    const bool live = (k + u) < cnt;
                    const float wv = __shfl_sync(FULL, e2.x, kk);
                    const float dv = __shfl_sync(FULL, e2.y, kk);
                    w[u] = live ? wv : 0.f;   // clamped tail edges contribute nothing
                    ds[u] = live ? dv : 0.f;
                    load_row<VPL, true>(g[u], p.G, i, d4, lane);
                    bits[u] = (unsigned)maskp[(size_t)pp * 32 + lane];
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    ds_sum += ds[u];
#pragma unroll
                    for (int i = 0; i < VPL; ++i) {
                        fma4(accw.v[i], w[u], g[u].v[i]);
                        const unsigned b = bits[u] >> (4 * i);
                        if (b & 1u) accz.v[i].x += ds[u];
                        if (b & 2u) accz.v[i].y += ds[u];
                        if (b & 4u) accz.v[i].z += ds[u];
                        if (b & 8u) accz.v[i].w += ds[u];
                    }
                }
            }
        }
        // nu' = slope * sum ds + (1 - slope) * sum_{z>0} ds;  d a += sign * Z[seg] (.) nu';  out_z = sign * a (.) nu'
        const float lo = p.slope * ds_sum, hi = 1.f - p.slope;
        Row<VPL> zr;
        load_row<VPL, MODE == 0>(zr, p.Z, item.seg, d4, lane);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            accz.v[i].x = sign * fmaf(hi, accz.v[i].x, lo); accz.v[i].y = sign * fmaf(hi, accz.v[i].y, lo);
            accz.v[i].z = sign * fmaf(hi, accz.v[i].z, lo); accz.v[i].w = sign * fmaf(hi, accz.v[i].w, lo);
            da_acc.v[i].x = fmaf(zr.v[i].x, accz.v[i].x, da_acc.v[i].x); da_acc.v[i].y = fmaf(zr.v[i].y, accz.v[i].y, da_acc.v[i].y);
            da_acc.v[i].z = fmaf(zr.v[i].z, accz.v[i].z, da_acc.v[i].z); da_acc.v[i].w = fmaf(zr.v[i].w, accz.v[i].w, da_acc.v[i].w);
            accz.v[i].x *= av.v[i].x; accz.v[i].y *= av.v[i].y; accz.v[i].z *= av.v[i].z; accz.v[i].w *= av.v[i].w;
            accw.v[i] = sign * accw.v[i];
        }
        if (item.slot < 0) {
            store_row<VPL>(accw, p.out_w, item.seg, d4, lane);
            store_row<VPL>(accz, p.out_z, item.seg, d4, lane);
        } else if (MODE == 0) {
            store_row<VPL>(accw, p.partial, (int64_t)item.slot * 2, d4, lane);
            store_row<VPL>(accz, p.partial, (int64_t)item.slot * 2 + 1, d4, lane);
        } else {
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c = i * 32 + lane;
                if (c < d4) {
                    red_add_v4(p.out_w + ((size_t)item.seg * d4 + c) * 4, accw.v[i]);
                    red_add_v4(p.out_z + ((size_t)item.seg * d4 + c) * 4, accz.v[i]);
                }
            }
        }
    }
    flush_da<VPL>(da_acc, p.da, d4);
}

// ------------------------------------------------------------------------------------------------
// Ring variants (default): the one gathered node row per edge (X[j] in the dst pass, G[i] in the src / rel
// passes) is fetched by lane 0 with cp.async.bulk into a per-warp ring of shared-memory slots, BWD_RING_S-1
// edges ahead of its use, two edges are consumed per iteration.  16 warps x 12 slots x 1 KB = 192 KB of
// gathers in flight per SM (d = 256) -- the register-staged kernels above top out at ~64 KB, which Little's
// law (2-2.7 us loaded latency) turns into ~4.5 TB/s.
// ------------------------------------------------------------------------------------------------
// ring depth by row width: 12 slots of 1 KB + mask at d <= 256, 7 / 6 for the wider rows of VPL = 3 / 4
template <int VPL> struct BwdRing { static constexpr int S = VPL <= 2 ? 12 : (VPL == 3 ? 7 : 6); };

template <int BWD_RING_S>
struct Ring {
    uint8_t* slots;
    uint32_t slots_s, bars, row_bytes, mask_b, slot_bytes;
    int row_f;
    uint32_t issue_slot, use_slot, use_phase;                  // maintained incrementally: no modulo by BWD_RING_S in the loops
    __device__ __forceinline__ void init(uint8_t* raw, int wib, int lane, int d4, int mask_bytes) {
        row_bytes = (uint32_t)d4 * 16;
        mask_b = (uint32_t)mask_bytes;
        slot_bytes = row_bytes + mask_b;                       // {gathered row, sign mask of the edge}
        row_f = d4 * 4;
        slots = raw + (size_t)wib * BWD_RING_S * slot_bytes;
        slots_s = smem_addr(slots);
        bars = smem_addr(raw) + (uint32_t)RING_WARPS * BWD_RING_S * slot_bytes + (uint32_t)wib * BWD_RING_S * 8;
        issue_slot = use_slot = use_phase = 0;
        if (lane == 0) {
            for (int s = 0; s < BWD_RING_S; ++s) mbar_init(bars + 8 * s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    // one edge: the node row `row` of `table` and the 32/64-byte mask record at CSR position `pos`
    __device__ __forceinline__ void issue(const float* table, int row, const void* mask, int pos, int lane) {
        const uint32_t s = issue_slot;
        issue_slot = s + 1 == BWD_RING_S ? 0 : s + 1;
        if (lane == 0) {
            const uint32_t dst = slots_s + s * slot_bytes;
            mbar_expect_tx(bars + 8 * s, slot_bytes);
            bulk_row(dst, table + (size_t)row * row_f, row_bytes, bars + 8 * s);
            bulk_row(dst + row_bytes, reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(mask) + (size_t)pos * mask_b), mask_b, bars + 8 * s);
        }
    }
    // two edges under one elected-lane region (steady state of the pair loops)
    __device__ __forceinline__ void issue2(const float* table, int row0, int row1, const void* mask, int pos0, int pos1, int lane) {
        const uint32_t s0 = issue_slot;
        const uint32_t s1 = s0 + 1 == BWD_RING_S ? 0 : s0 + 1;
        issue_slot = s1 + 1 == BWD_RING_S ? 0 : s1 + 1;
        if (lane == 0) {
            const uint32_t d0 = slots_s + s0 * slot_bytes, d1 = slots_s + s1 * slot_bytes;
            const uint8_t* mb = reinterpret_cast<const uint8_t*>(mask);
            mbar_expect_tx(bars + 8 * s0, slot_bytes);
            bulk_row(d0, table + (size_t)row0 * row_f, row_bytes, bars + 8 * s0);
            bulk_row(d0 + row_bytes, reinterpret_cast<const float*>(mb + (size_t)pos0 * mask_b), mask_b, bars + 8 * s0);
            mbar_expect_tx(bars + 8 * s1, slot_bytes);
            bulk_row(d1, table + (size_t)row1 * row_f, row_bytes, bars + 8 * s1);
            bulk_row(d1 + row_bytes, reinterpret_cast<const float*>(mb + (size_t)pos1 * mask_b), mask_b, bars + 8 * s1);
        }
    }
    // waits for the next slot in consumption order and returns its shared-memory address
    __device__ __forceinline__ const uint8_t* wait_next() {
        const uint32_t s = use_slot;
        mbar_wait(bars + 8 * s, use_phase);
        if (s + 1 == BWD_RING_S) { use_slot = 0; use_phase ^= 1; } else use_slot = s + 1;
        return slots + (size_t)s * slot_bytes;
    }
};

// Contiguous run of items per warp (see edge_fwd.cu): the index batches and the ring run ahead across segments.
__device__ __forceinline__ int first_item_at_or_after(const Item* items, int n_items, int pos) {
    int lo = 0, hi = n_items;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&items[mid].beg) < pos) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// pull the per-segment rows of the NEXT segment towards L2 while the current one is processed (no registers held)
__device__ __forceinline__ void prefetch_rows_l2(const float* a, const float* b, const float* c, int row, int row_f, int lane) {
    const int lines = (row_f * 4 + 127) >> 7;                    // 128-byte lines per row (<= 16)
    const float* t = lane < 8 ? a : (lane < 16 ? b : c);
    const int ln = lane & 7;
    if (t != nullptr && lane < 24) {
        for (int q = ln; q < lines; q += 8)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(t + (size_t)row * row_f + q * 32));
    }
}

template <int VPL, bool FULLROW>
__global__ void __launch_bounds__(RING_WARPS * 32, 1) edge_bwd_dst_ring_kernel(EdgeBwdDstParams p) {
    typedef typename MaskT<VPL>::type mask_t;
    extern __shared__ __align__(16) uint8_t ring_raw[];
    const int lane = threadIdx.x & 31, wib = __shfl_sync(FULL, threadIdx.x >> 5, 0);   // broadcast: lets the compiler see a warp-uniform value
    const int warp = blockIdx.x * RING_WARPS + wib;
    const int nwarps = gridDim.x * RING_WARPS;
    const int d4 = FULLROW ? VPL * 32 : p.d4;      // see edge_fwd_ring_kernel
    constexpr int BWD_RING_S = BwdRing<VPL>::S;
    Ring<BWD_RING_S> ring;
    ring.init(ring_raw, wib, lane, d4, (int)sizeof(mask_t) * 32);
    Row<VPL> av, da_acc, gs_acc;
    load_row<VPL, false>(av, p.a, 0, d4, lane);
    da_acc.zero();
    gs_acc.zero();
    const float hi = 1.f - p.slope;
    constexpr int ADV = (32 - BWD_RING_S) & ~1;

    const int per = (p.E + nwarps - 1) / nwarps;
    const int it0 = first_item_at_or_after(p.items, p.n_items, min(p.E, warp * per));
    const int it1 = warp == nwarps - 1 ? p.n_items : first_item_at_or_after(p.items, p.n_items, min(p.E, (warp + 1) * per));
    if (it0 < it1) {
        const int stream_end = p.items[it1 - 1].end;
        Item item = p.items[it0];
        int batch_base = item.beg, issued = item.beg;             // absolute CSR positions
        int jv = 0, tv = 0;
        float sv = 0.f;
        if (batch_base + lane < stream_end) {
            jv = __ldg(p.csr_src + batch_base + lane);
            tv = __ldg(p.csr_type + batch_base + lane);
            sv = __ldg(p.logit + batch_base + lane);
        }
        prefetch_idx_l2(p.csr_src, p.csr_type, p.logit, batch_base + 32, stream_end, lane);
        prefetch_idx_l2(p.csr_src, p.csr_type, p.logit, batch_base + 64, stream_end, lane);
        float w_keep = 0.f, ds_keep = 0.f;
        // The relation rows R'[t] of a pair come from L2 (~700 cycles under load) and are needed as soon as the pair's gathered rows are
        // there: they are loaded one pair ahead (by stream position -- the stream runs across segment boundaries) and carried in registers.
        // (rows of up to 8 registers; the wider ones would spill, they load at the point of use.)
        constexpr bool REL_AHEAD = VPL <= 2;
        Row<REL_AHEAD ? VPL : 1> rn0, rn1;
        if (REL_AHEAD) {
            load_row<REL_AHEAD ? VPL : 1, false>(rn0, p.Rp, __shfl_sync(FULL, tv, 0), d4, lane);
            load_row<REL_AHEAD ? VPL : 1, false>(rn1, p.Rp, __shfl_sync(FULL, tv, 1), d4, lane);
        }
        for (int it = it0; it < it1; ++it) {
            Item nxt = item;
            if (it + 1 < it1) {
                nxt = p.items[it + 1];
                prefetch_rows_l2(p.G, p.agg, p.P, nxt.seg, d4 * 4, lane);
            }
            Row<VPL> gi, mu;
            load_row<VPL, true>(gi, p.G, item.seg, d4, lane);
            if (p.gsum && (it == 0 || __ldg(&p.items[it - 1].seg) != item.seg)) {   // every destination once (first item of its segment)
#pragma unroll
                for (int i = 0; i < VPL; ++i) gs_acc.v[i] = gs_acc.v[i] + gi.v[i];
            }
            mu.zero();
            float gdot = 0.f, ds_sum = 0.f;
            {
                Row<VPL> ai;
                load_row<VPL, true>(ai, p.agg, item.seg, d4, lane);
#pragma unroll
                for (int i = 0; i < VPL; ++i) gdot += dot4(gi.v[i], ai.v[i]);
                gdot = warp_sum(gdot);
            }
            const float lse = p.lse[item.seg];
            const float c = p.sqrt_deg[item.seg];
            for (int k = item.beg; k < item.end; k += 2) {
                if (k - batch_base >= ADV) {                      // next index batch (warp-uniform)
                    if (batch_base + lane < k) p.ew[batch_base + lane] = make_float2(w_keep, ds_keep);
                    batch_base = k;
                    jv = tv = 0;
                    sv = 0.f;
                    if (batch_base + lane < stream_end) {
                        jv = __ldg(p.csr_src + batch_base + lane);
                        tv = __ldg(p.csr_type + batch_base + lane);
                        sv = __ldg(p.logit + batch_base + lane);
                    }
                    prefetch_idx_l2(p.csr_src, p.csr_type, p.logit, batch_base + 96, stream_end, lane);   // a batch advances < 32 entries: every line is covered
                }
                const int want = min(stream_end, k + BWD_RING_S);
                while (issued + 1 < want) {                      // warp-uniform; steady state: one pair per iteration
                    ring.issue2(p.X, __shfl_sync(FULL, jv, issued - batch_base), __shfl_sync(FULL, jv, issued + 1 - batch_base), p.mask, issued,
                                issued + 1, lane);
                    issued += 2;
                }
                if (issued < want) {
                    ring.issue(p.X, __shfl_sync(FULL, jv, issued - batch_base), p.mask, issued, lane);
                    ++issued;
                }
                const bool two = k + 1 < item.end;
                const int kb = k - batch_base, kb1 = two ? kb + 1 : kb;
                const float sc0 = __shfl_sync(FULL, sv, kb), sc1 = __shfl_sync(FULL, sv, kb1);
                Row<VPL> r0, r1;
                if constexpr (REL_AHEAD) {
                    r0 = rn0;
                    r1 = rn1;                                     // r1 of an unpaired last edge is some other (finite) row: its weight below is 0
                    const int nb = kb + (two ? 2 : 1);            // the next pair of the stream; kb < ADV, so nb + 1 < 32 (lanes past the stream hold type 0)
                    load_row<VPL, false>(rn0, p.Rp, __shfl_sync(FULL, tv, nb), d4, lane);
                    load_row<VPL, false>(rn1, p.Rp, __shfl_sync(FULL, tv, nb + 1), d4, lane);
                } else {
                    load_row<VPL, false>(r0, p.Rp, __shfl_sync(FULL, tv, kb), d4, lane);
                    load_row<VPL, false>(r1, p.Rp, __shfl_sync(FULL, tv, kb1), d4, lane);
                }
                Row<VPL> x0, x1;
                const uint8_t* sl0 = ring.wait_next();
                lds_row<VPL>(x0, reinterpret_cast<const float*>(sl0), d4, lane);
                const unsigned b0 = (unsigned)reinterpret_cast<const mask_t*>(sl0 + ring.row_bytes)[lane];
                unsigned b1 = b0;
                if (two) {
                    const uint8_t* sl1 = ring.wait_next();
                    lds_row<VPL>(x1, reinterpret_cast<const float*>(sl1), d4, lane);
                    b1 = (unsigned)reinterpret_cast<const mask_t*>(sl1 + ring.row_bytes)[lane];
                } else {
                    x1 = x0;
                }
                __syncwarp();                                    // slots may be refilled
                float u0 = 0.f, u1 = 0.f;
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    u0 += dot4(gi.v[i], x0.v[i] - r0.v[i]);
                    u1 += dot4(gi.v[i], x1.v[i] - r1.v[i]);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    u0 += __shfl_xor_sync(FULL, u0, o);
                    u1 += __shfl_xor_sync(FULL, u1, o);
                }
                const float al0 = __expf(sc0 - lse), al1 = two ? __expf(sc1 - lse) : 0.f;
                const float ds0 = al0 * (c * u0 - gdot), ds1 = al1 * (c * u1 - gdot);
                ds_sum += ds0 + ds1;
                // one bit test (predicate) + one predicated add per feature and edge
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    if (b0 & (1u << (4 * i))) mu.v[i].x += ds0;
                    if (b0 & (2u << (4 * i))) mu.v[i].y += ds0;
                    if (b0 & (4u << (4 * i))) mu.v[i].z += ds0;
                    if (b0 & (8u << (4 * i))) mu.v[i].w += ds0;
                }
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    if (b1 & (1u << (4 * i))) mu.v[i].x += ds1;
                    if (b1 & (2u << (4 * i))) mu.v[i].y += ds1;
                    if (b1 & (4u << (4 * i))) mu.v[i].z += ds1;
                    if (b1 & (8u << (4 * i))) mu.v[i].w += ds1;
                }
                if (lane == kb) { w_keep = c * al0; ds_keep = ds0; }
                if (two && lane == kb + 1) { w_keep = c * al1; ds_keep = ds1; }
            }
            const float lo = p.slope * ds_sum;
            Row<VPL> pi;
            load_row<VPL, true>(pi, p.P, item.seg, d4, lane);
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                mu.v[i].x = fmaf(hi, mu.v[i].x, lo); mu.v[i].y = fmaf(hi, mu.v[i].y, lo);
                mu.v[i].z = fmaf(hi, mu.v[i].z, lo); mu.v[i].w = fmaf(hi, mu.v[i].w, lo);
                da_acc.v[i].x = fmaf(pi.v[i].x, mu.v[i].x, da_acc.v[i].x); da_acc.v[i].y = fmaf(pi.v[i].y, mu.v[i].y, da_acc.v[i].y);
                da_acc.v[i].z = fmaf(pi.v[i].z, mu.v[i].z, da_acc.v[i].z); da_acc.v[i].w = fmaf(pi.v[i].w, mu.v[i].w, da_acc.v[i].w);
                mu.v[i].x *= av.v[i].x; mu.v[i].y *= av.v[i].y; mu.v[i].z *= av.v[i].z; mu.v[i].w *= av.v[i].w;
            }
            if (item.slot < 0) store_row<VPL>(mu, p.dP, item.seg, d4, lane);
            else store_row<VPL>(mu, p.partial, item.slot, d4, lane);
            item = nxt;
        }
        if (batch_base + lane < stream_end) p.ew[batch_base + lane] = make_float2(w_keep, ds_keep);
    }
    flush_da16<VPL>(da_acc, p.da, d4);
    if (p.gsum) {
        __syncthreads();
        flush_da16<VPL>(gs_acc, p.gsum, d4);
    }
}

template <int VPL, int MODE, bool FULLROW>
__global__ void __launch_bounds__(RING_WARPS * 32, 1) edge_bwd_gather_ring_kernel(EdgeBwdGatherParams p) {
    typedef typename MaskT<VPL>::type mask_t;
    extern __shared__ __align__(16) uint8_t ring_raw[];
    const int lane = threadIdx.x & 31, wib = __shfl_sync(FULL, threadIdx.x >> 5, 0);   // broadcast: lets the compiler see a warp-uniform value
    const int warp = blockIdx.x * RING_WARPS + wib;
    const int nwarps = gridDim.x * RING_WARPS;
    const int d4 = FULLROW ? VPL * 32 : p.d4;      // see edge_fwd_ring_kernel
    constexpr int BWD_RING_S = BwdRing<VPL>::S;
    Ring<BWD_RING_S> ring;
    ring.init(ring_raw, wib, lane, d4, (int)sizeof(mask_t) * 32);
    Row<VPL> av, da_acc;
    load_row<VPL, false>(av, p.a, 0, d4, lane);
    da_acc.zero();
    const float sign = MODE == 0 ? 1.f : -1.f;
    constexpr int ADV = (32 - BWD_RING_S) & ~1;

    for (int it = warp; it < p.n_items; it += nwarps) {
        const Item item = p.items[it];
        const int n = item.end - item.beg;
        Row<VPL> accw, accz;
        accw.zero();
        accz.zero();
        float ds_sum = 0.f;
        int issued = 0;
        for (int base = 0; base < n; base += ADV) {
            int pv = 0, iv = 0;
            float2 e2 = make_float2(0.f, 0.f);
            if (base + lane < n) {
                pv = __ldg(p.pos + item.beg + base + lane);
                iv = __ldg(p.dst + item.beg + base + lane);
                e2 = __ldg(p.ew + pv);
            }
            const int lim = min(n, base + ADV);
            for (int k = base; k < lim; k += 2) {
                const int want = min(n, k + BWD_RING_S);
                while (issued + 1 < want) {                          // warp-uniform; steady state: one pair per iteration
                    ring.issue2(p.G, __shfl_sync(FULL, iv, issued - base), __shfl_sync(FULL, iv, issued + 1 - base), p.mask,
                                __shfl_sync(FULL, pv, issued - base), __shfl_sync(FULL, pv, issued + 1 - base), lane);
                    issued += 2;
                }
                if (issued < want) {
                    ring.issue(p.G, __shfl_sync(FULL, iv, issued - base), p.mask, __shfl_sync(FULL, pv, issued - base), lane);
                    ++issued;
                }
                const bool two = k + 1 < lim;
                const int k1 = two ? k + 1 : k;
                const float w0 = __shfl_sync(FULL, e2.x, k - base), d0 = __shfl_sync(FULL, e2.y, k - base);
                float w1 = __shfl_sync(FULL, e2.x, k1 - base), d1 = __shfl_sync(FULL, e2.y, k1 - base);
                if (!two) { w1 = 0.f; d1 = 0.f; }
                Row<VPL> g0, g1;
                const uint8_t* sl0 = ring.wait_next();
                lds_row<VPL>(g0, reinterpret_cast<const float*>(sl0), d4, lane);
                const unsigned b0 = (unsigned)reinterpret_cast<const mask_t*>(sl0 + ring.row_bytes)[lane];
                unsigned b1 = b0;
                if (two) {
                    const uint8_t* sl1 = ring.wait_next();
                    lds_row<VPL>(g1, reinterpret_cast<const float*>(sl1), d4, lane);
                    b1 = (unsigned)reinterpret_cast<const mask_t*>(sl1 + ring.row_bytes)[lane];
                } else {
                    g1 = g0;
                }
                __syncwarp();
                ds_sum += d0 + d1;
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    fma4(accw.v[i], w0, g0.v[i]);
                    fma4(accw.v[i], w1, g1.v[i]);
                }
                // one bit test (predicate) + one predicated add per feature and edge
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    if (b0 & (1u << (4 * i))) accz.v[i].x += d0;
                    if (b0 & (2u << (4 * i))) accz.v[i].y += d0;
                    if (b0 & (4u << (4 * i))) accz.v[i].z += d0;
                    if (b0 & (8u << (4 * i))) accz.v[i].w += d0;
                }
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    if (b1 & (1u << (4 * i))) accz.v[i].x += d1;
                    if (b1 & (2u << (4 * i))) accz.v[i].y += d1;
                    if (b1 & (4u << (4 * i))) accz.v[i].z += d1;
                    if (b1 & (8u << (4 * i))) accz.v[i].w += d1;
                }
            }
        }
        const float lo = p.slope * ds_sum, hi = 1.f - p.slope;
        Row<VPL> zr;
        load_row<VPL, MODE == 0>(zr, p.Z, item.seg, d4, lane);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            accz.v[i].x = sign * fmaf(hi, accz.v[i].x, lo); accz.v[i].y = sign * fmaf(hi, accz.v[i].y, lo);
            accz.v[i].z = sign * fmaf(hi, accz.v[i].z, lo); accz.v[i].w = sign * fmaf(hi, accz.v[i].w, lo);
            da_acc.v[i].x = fmaf(zr.v[i].x, accz.v[i].x, da_acc.v[i].x); da_acc.v[i].y = fmaf(zr.v[i].y, accz.v[i].y, da_acc.v[i].y);
            da_acc.v[i].z = fmaf(zr.v[i].z, accz.v[i].z, da_acc.v[i].z); da_acc.v[i].w = fmaf(zr.v[i].w, accz.v[i].w, da_acc.v[i].w);
            accz.v[i].x *= av.v[i].x; accz.v[i].y *= av.v[i].y; accz.v[i].z *= av.v[i].z; accz.v[i].w *= av.v[i].w;
            accw.v[i] = sign * accw.v[i];
        }
        if (item.slot < 0) {
            store_row<VPL>(accw, p.out_w, item.seg, d4, lane);
            store_row<VPL>(accz, p.out_z, item.seg, d4, lane);
        } else if (MODE == 0) {
            store_row<VPL>(accw, p.partial, (int64_t)item.slot * 2, d4, lane);
            store_row<VPL>(accz, p.partial, (int64_t)item.slot * 2 + 1, d4, lane);
        } else {
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                const int c = i * 32 + lane;
                if (c < d4) {
                    red_add_v4(p.out_w + ((size_t)item.seg * d4 + c) * 4, accw.v[i]);
                    red_add_v4(p.out_z + ((size_t)item.seg * d4 + c) * 4, accz.v[i]);
                }
            }
        }
    }
    flush_da16<VPL>(da_acc, p.da, d4);
}

static inline int bwd_ring_depth(int d) {
    const int vpl = (d / 4 + 31) / 32;
    return vpl <= 2 ? BwdRing<2>::S : (vpl == 3 ? BwdRing<3>::S : BwdRing<4>::S);
}
static inline size_t bwd_ring_smem(int d) { return (size_t)RING_WARPS * bwd_ring_depth(d) * ((size_t)d * 4 + edge_mask_bytes(d) + 8); }
static inline bool use_ring(int d) {
    static const int on = getenv("JMAC_EDGE_RING") ? atoi(getenv("JMAC_EDGE_RING")) : 1;
    const int vpl = (d / 4 + 31) / 32;
    const size_t static_red = (size_t)RING_WARPS * vpl * 32 * sizeof(float4);      // flush_da16's block reduction buffer
    return on && bwd_ring_smem(d) + static_red + 1024 <= 227 * 1024;
}

}  // namespace jmac

using namespace jmac;

extern "C" size_t jmac_edge_record_bytes(int d) { return dim_ok(d) ? 8 + edge_mask_bytes(d) + 4 : 0; }

extern "C" int jmac_edge_bwd_dst(jmac_plan* plan, const float* P, const float* X, const float* Rp, const float* a_att, float leaky_slope,
                                 int d, const float* agg, const float* lse, const float* G, float* dP, void* erec, float* da_att,
                                 float* g_colsum, void* stream) {
    JMAC_ARG(plan && P && X && Rp && a_att && agg && lse && G && dP && da_att);
    JMAC_ARG(plan->E == 0 || erec != nullptr);
    JMAC_ARG((reinterpret_cast<uintptr_t>(erec) & 31) == 0);
    JMAC_ARG(dim_ok(d));
    cudaStream_t st = (cudaStream_t)stream;
    EdgeBwdDstParams p;
    p.items = plan->dst_items.items; p.n_items = plan->dst_items.n_items; p.E = (int)plan->E;
    p.csr_src = plan->csr_src; p.csr_type = plan->csr_type; p.sqrt_deg = plan->sqrt_deg;
    p.P = P; p.X = X; p.Rp = Rp; p.a = a_att;
    p.slope = leaky_slope; p.d4 = d / 4;
    p.agg = agg; p.lse = lse; p.G = G; p.dP = dP;
    p.ew = erec_ew(erec, plan->E, d);
    p.mask = erec_mask(erec, plan->E);
    p.logit = erec_logit(erec, plan->E, d);
    p.da = da_att;
    p.gsum = g_colsum;
    p.partial = nullptr;
    if (plan->dst_items.n_slots > 0) {
        int rc = plan_scratch(plan, sizeof(float) * (size_t)plan->dst_items.n_slots * (size_t)d, &p.partial, st);
        if (rc) return rc;
    }
    if (p.n_items == 0) return 0;
    const int threads = 256;
    JMAC_DISPATCH_VPL(d, {
        constexpr int U = VPL <= 2 ? 4 : 2;
        if (use_ring(d) && plan->E > 0) {
            auto kern = d == VPL * 128 ? edge_bwd_dst_ring_kernel<VPL, true> : edge_bwd_dst_ring_kernel<VPL, false>;
            JMAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_ring_smem(d)));
            long long want = ((long long)p.n_items + RING_WARPS - 1) / RING_WARPS;
            kern<<<(int)(want < sm_count() ? want : sm_count()), RING_WARPS * 32, bwd_ring_smem(d), st>>>(p);
        } else {
            auto kern = edge_bwd_dst_kernel<VPL, U>;
            int blocks = persistent_blocks(kern, threads, p.n_items);
            kern<<<blocks, threads, 0, st>>>(p);
        }
        JMAC_LAUNCH_CHECK();
        if (plan->dst_items.n_splits > 0) {
            sum_partials_kernel<VPL><<<plan->dst_items.n_splits, threads, 0, st>>>(plan->dst_items.splits, plan->dst_items.n_splits, p.partial, p.d4, 1, dP, nullptr);
            JMAC_LAUNCH_CHECK();
        }
    });
    plan_mark_use(plan, st);
    return 0;
}

static int gather_pass(jmac_plan* plan, int mode, const float* G, const void* erec, const float* a_att, float slope, int d, const float* Z,
                       float* out_w, float* out_z, float* da_att, cudaStream_t st) {
    JMAC_ARG(plan && G && a_att && Z && out_w && out_z && da_att);
    JMAC_ARG(plan->E == 0 || erec != nullptr);
    JMAC_ARG((reinterpret_cast<uintptr_t>(erec) & 31) == 0);
    JMAC_ARG(dim_ok(d));
    const ItemList& il = mode == 0 ? plan->src_items : plan->typ_items;
    EdgeBwdGatherParams p;
    p.items = il.items; p.n_items = il.n_items;
    p.pos = mode == 0 ? plan->csc_pos : plan->typ_pos;
    p.dst = mode == 0 ? plan->csc_dst : plan->typ_dst;
    p.G = G;
    p.ew = erec_ew(const_cast<void*>(erec), plan->E, d);
    p.mask = erec_mask(const_cast<void*>(erec), plan->E);
    p.a = a_att; p.Z = Z; p.slope = slope; p.d4 = d / 4;
    p.out_w = out_w; p.out_z = out_z; p.da = da_att;
    p.partial = nullptr;
    if (mode == 1) {
        JMAC_CUDA(cudaMemsetAsync(out_w, 0, sizeof(float) * (size_t)plan->n_rel * d, st));
        JMAC_CUDA(cudaMemsetAsync(out_z, 0, sizeof(float) * (size_t)plan->n_rel * d, st));
    } else if (il.n_slots > 0) {
        int rc = plan_scratch(plan, sizeof(float) * (size_t)il.n_slots * 2 * (size_t)d, &p.partial, st);
        if (rc) return rc;
    }
    if (p.n_items == 0) return 0;
    const int threads = 256;
    JMAC_DISPATCH_VPL(d, {
        constexpr int U = VPL == 1 ? 4 : (VPL == 2 ? 4 : 2);
        const bool ring = use_ring(d);
        const long long want = ((long long)p.n_items + RING_WARPS - 1) / RING_WARPS;
        const int rblocks = (int)(want < sm_count() ? want : sm_count());
        if (mode == 0) {
            if (ring) {
                auto kern = d == VPL * 128 ? edge_bwd_gather_ring_kernel<VPL, 0, true> : edge_bwd_gather_ring_kernel<VPL, 0, false>;
                JMAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_ring_smem(d)));
                kern<<<rblocks, RING_WARPS * 32, bwd_ring_smem(d), st>>>(p);
            } else {
                auto kern = edge_bwd_gather_kernel<VPL, 0, U>;
                int blocks = persistent_blocks(kern, threads, p.n_items);
                kern<<<blocks, threads, 0, st>>>(p);
            }
            JMAC_LAUNCH_CHECK();
            if (il.n_splits > 0) {
                sum_partials_kernel<VPL><<<il.n_splits, threads, 0, st>>>(il.splits, il.n_splits, p.partial, p.d4, 2, out_w, out_z);
                JMAC_LAUNCH_CHECK();
            }
        } else {
            if (ring) {
                auto kern = d == VPL * 128 ? edge_bwd_gather_ring_kernel<VPL, 1, true> : edge_bwd_gather_ring_kernel<VPL, 1, false>;
                JMAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd_ring_smem(d)));
                kern<<<rblocks, RING_WARPS * 32, bwd_ring_smem(d), st>>>(p);
            } else {
                auto kern = edge_bwd_gather_kernel<VPL, 1, U>;
                int blocks = persistent_blocks(kern, threads, p.n_items);
                kern<<<blocks, threads, 0, st>>>(p);
            }
            JMAC_LAUNCH_CHECK();
        }
    });
    plan_mark_use(plan, st);
    return 0;
}

extern "C" int jmac_edge_bwd_src(jmac_plan* plan, const float* G, const void* erec, const float* a_att, float leaky_slope, int d,
                                 const float* Q, float* dXe, float* dQ, float* da_att, void* stream) {
    return gather_pass(plan, 0, G, erec, a_att, leaky_slope, d, Q, dXe, dQ, da_att, (cudaStream_t)stream);
}

extern "C" int jmac_edge_bwd_rel(jmac_plan* plan, const float* G, const void* erec, const float* a_att, float leaky_slope, int d,
                                 const float* S, float* dRp, float* dS, float* da_att, void* stream) {
    return gather_pass(plan, 1, G, erec, a_att, leaky_slope, d, S, dRp, dS, da_att, (cudaStream_t)stream);
}
