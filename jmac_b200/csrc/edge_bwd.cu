// Backward of the edge phase (autograd of src/jmac_model.py:56-89 + message_passing.py:24-28, which
// the reference gets from index_put_(accumulate) atomics and E-row GEMM transposes).
//
// Three atomics-free segmented passes over the three edge orderings of the plan:
//   dst pass : per destination i  -> dP[i]; per-edge record {w_e = alpha_e*sqrt(deg_i), ds_e,
//              sign mask of z_e = P[i]+Q[j]-S[t]}; da_att.  Gathers Q[j], X[j] (2 rows / edge).
//   src pass : per source j       -> dXe[j] = sum w_e G[i],  dQ[j] = sum ds_e a (.) lrelu'(z_e).
//              Gathers G[i] (1 row / edge) + the 8+MB byte record.
//   rel pass : per relation t     -> dR'[t] = -sum w_e G[i], dS[t] = -sum ds_e a (.) lrelu'(z_e).
//              Gathers G[i] (1 row / edge) + the record.
// The sign mask (1 bit per feature, 32 B per edge at d=256) is what lets the src and rel passes
// rebuild d z_e without re-gathering P[i], Q[j], S[t] (3 KB per edge).
// Algorithmic HBM bytes per edge: dst 2*d*4 + 8 + rec, src d*4 + 8 + rec, rel d*4 + 8 + rec.
#include <math.h>

#include "common.cuh"

namespace jmac {

template <int VPL> struct MaskT { typedef uint8_t type; };
template <> struct MaskT<3> { typedef uint16_t type; };
template <> struct MaskT<4> { typedef uint16_t type; };

static inline size_t mask_bytes(int d) { return (((d / 4 + 31) / 32) <= 2) ? 32 : 64; }

struct EdgeBwdDstParams {
    const Item* items;
    int n_items;
    const int* csr_src;
    const int* csr_type;
    const float* sqrt_deg;
    const float *P, *Q, *X, *S, *Rp, *a;
    float slope;
    int d4;
    const float *agg, *lse, *G;
    float* dP;
    float2* ew;     // [E] (w_e, ds_e)
    void* mask;     // [E][32] lanes x MaskT
    float* da;      // [d] accumulated
    float* partial; // [n_slots, d]
};

template <int VPL, int U>
__global__ void __launch_bounds__(256) edge_bwd_dst_kernel(EdgeBwdDstParams p) {
    typedef typename MaskT<VPL>::type mask_t;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int d4 = p.d4;
    Row<VPL> av, da_acc;
    load_row<VPL, false>(av, p.a, 0, d4, lane);
    da_acc.zero();
    mask_t* maskp = reinterpret_cast<mask_t*>(p.mask);

    for (int it = warp; it < p.n_items; it += nwarps) {
        const Item item = p.items[it];
        Row<VPL> pi, gi, dp_acc;
        load_row<VPL, true>(pi, p.P, item.seg, d4, lane);
        load_row<VPL, true>(gi, p.G, item.seg, d4, lane);
        dp_acc.zero();
        float gdot = 0.f;
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
            if (lane < cnt) {
                jv = __ldg(p.csr_src + base + lane);
                tv = __ldg(p.csr_type + base + lane);
            }
            float w_keep = 0.f, ds_keep = 0.f;  // lane k keeps the record of edge base+k
            int k = 0;
            for (; k < cnt; k += U) {
                Row<VPL> q[U], x[U], s[U], r[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int kk = min(k + u, cnt - 1);  // clamp: tail edges are recomputed, never recorded twice
                    const int j = __shfl_sync(FULL, jv, kk);
                    const int t = __shfl_sync(FULL, tv, kk);
                    load_row<VPL, true>(q[u], p.Q, j, d4, lane);
                    load_row<VPL, true>(x[u], p.X, j, d4, lane);
                    load_row<VPL, false>(s[u], p.S, t, d4, lane);
                    load_row<VPL, false>(r[u], p.Rp, t, d4, lane);
                }
                float sc[U], uu[U];
                unsigned bits[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    sc[u] = 0.f; uu[u] = 0.f; bits[u] = 0u;
#pragma unroll
                    for (int i = 0; i < VPL; ++i) {
                        float4 z = pi.v[i] + q[u].v[i] - s[u].v[i];
                        bits[u] |= ((z.x > 0.f ? 1u : 0u) | (z.y > 0.f ? 2u : 0u) | (z.z > 0.f ? 4u : 0u) | (z.w > 0.f ? 8u : 0u)) << (4 * i);
                        z.x = z.x > 0.f ? z.x : p.slope * z.x;
                        z.y = z.y > 0.f ? z.y : p.slope * z.y;
                        z.z = z.z > 0.f ? z.z : p.slope * z.z;
                        z.w = z.w > 0.f ? z.w : p.slope * z.w;
                        sc[u] += dot4(av.v[i], z);
                        uu[u] += dot4(gi.v[i], x[u].v[i] - r[u].v[i]);
                        q[u].v[i] = z;  // keep LeakyReLU(z) for da_att
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        sc[u] += __shfl_xor_sync(FULL, sc[u], o);
                        uu[u] += __shfl_xor_sync(FULL, uu[u], o);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (k + u < cnt) {  // warp-uniform
                        const float alpha = __expf(sc[u] - lse);
                        const float w = c * alpha;
                        const float ds = alpha * (c * uu[u] - gdot);
                        const float dsl = ds * p.slope;
#pragma unroll
                        for (int i = 0; i < VPL; ++i) {
                            const unsigned b = bits[u] >> (4 * i);
                            dp_acc.v[i].x = fmaf((b & 1u) ? ds : dsl, av.v[i].x, dp_acc.v[i].x);
                            dp_acc.v[i].y = fmaf((b & 2u) ? ds : dsl, av.v[i].y, dp_acc.v[i].y);
                            dp_acc.v[i].z = fmaf((b & 4u) ? ds : dsl, av.v[i].z, dp_acc.v[i].z);
                            dp_acc.v[i].w = fmaf((b & 8u) ? ds : dsl, av.v[i].w, dp_acc.v[i].w);
                            fma4(da_acc.v[i], ds, q[u].v[i]);
                        }
                        maskp[(size_t)(base + k + u) * 32 + lane] = (mask_t)bits[u];
                        if (lane == k + u) { w_keep = w; ds_keep = ds; }
                    }
                }
            }
            if (lane < cnt) p.ew[base + lane] = make_float2(w_keep, ds_keep);
        }
        if (item.slot < 0) store_row<VPL>(dp_acc, p.dP, item.seg, d4, lane);
        else store_row<VPL>(dp_acc, p.partial, item.slot, d4, lane);
    }
    // da_att: block-level reduction in shared memory, then one RED per feature per block
    __shared__ float4 red[8][VPL * 32];
    const int wib = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < VPL; ++i) red[wib][i * 32 + lane] = da_acc.v[i];
    __syncthreads();
    for (int c4 = threadIdx.x; c4 < d4; c4 += blockDim.x) {
        float4 s = red[0][c4];
        for (int w = 1; w < 8; ++w) s = s + red[w][c4];
        red_add_v4(p.da + 4 * c4, s);
    }
}

// Sum of the partial rows of split segments: out[seg] = sum_k partial[first + k]; `rows` rows per slot.
template <int VPL>
__global__ void __launch_bounds__(256) sum_partials_kernel(const Split* __restrict__ splits, int n_splits, const float* __restrict__ partial,
                                                           int d4, int rows, float* __restrict__ out0, float* __restrict__ out1) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_splits) return;
    const Split sp = splits[w];
    for (int r = 0; r < rows; ++r) {
        Row<VPL> acc;
        acc.zero();
        for (int k = 0; k < sp.count; ++k) {
            Row<VPL> part;
            load_row<VPL, true>(part, partial, (int64_t)(sp.first + k) * rows + r, d4, lane);
#pragma unroll
            for (int i = 0; i < VPL; ++i) acc.v[i] = acc.v[i] + part.v[i];
        }
        store_row<VPL>(acc, r == 0 ? out0 : out1, sp.seg, d4, lane);
    }
}

struct EdgeBwdGatherParams {
    const Item* items;
    int n_items;
    const int* pos;   // ordering position -> CSR position
    const int* dst;   // ordering position -> destination node
    const float* G;
    const float2* ew;
    const void* mask;
    const float* a;
    float slope;
    int d4;
    float* out_w;     // sum w_e G[i]      (dXe for the src pass, dR' for the rel pass)
    float* out_z;     // sum d z_e         (dQ  for the src pass, dS  for the rel pass)
    float* partial;   // [n_slots, 2, d]   (src pass only)
};

// MODE 0: source pass (rows written / partials); MODE 1: relation pass (negated; split segments RED
// into the zeroed tables).
template <int VPL, int MODE, int U>
__global__ void __launch_bounds__(256) edge_bwd_gather_kernel(EdgeBwdGatherParams p) {
    typedef typename MaskT<VPL>::type mask_t;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int d4 = p.d4;
    Row<VPL> av;
    load_row<VPL, false>(av, p.a, 0, d4, lane);
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
                        accz.v[i].x += (b & 1u) ? ds[u] : 0.f;
                        accz.v[i].y += (b & 2u) ? ds[u] : 0.f;
                        accz.v[i].z += (b & 4u) ? ds[u] : 0.f;
                        accz.v[i].w += (b & 8u) ? ds[u] : 0.f;
                    }
                }
            }
        }
        // d z summed = a (.) (slope * sum ds + (1 - slope) * sum_{z>0} ds)
        const float lo = p.slope * ds_sum, hi = 1.f - p.slope;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            accz.v[i].x = sign * av.v[i].x * fmaf(hi, accz.v[i].x, lo);
            accz.v[i].y = sign * av.v[i].y * fmaf(hi, accz.v[i].y, lo);
            accz.v[i].z = sign * av.v[i].z * fmaf(hi, accz.v[i].z, lo);
            accz.v[i].w = sign * av.v[i].w * fmaf(hi, accz.v[i].w, lo);
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
}

}  // namespace jmac

using namespace jmac;

extern "C" size_t jmac_edge_record_bytes(int d) { return dim_ok(d) ? 8 + mask_bytes(d) : 0; }

extern "C" int jmac_edge_bwd_dst(jmac_plan* plan, const float* P, const float* Q, const float* X, const float* S, const float* Rp,
                                 const float* a_att, float leaky_slope, int d, const float* agg, const float* lse, const float* G,
                                 float* dP, void* erec, float* da_att, void* stream) {
    JMAC_ARG(plan && P && Q && X && S && Rp && a_att && agg && lse && G && dP && da_att);
    JMAC_ARG(plan->E == 0 || erec != nullptr);
    JMAC_ARG(dim_ok(d));
    cudaStream_t st = (cudaStream_t)stream;
    EdgeBwdDstParams p;
    p.items = plan->dst_items.items; p.n_items = plan->dst_items.n_items;
    p.csr_src = plan->csr_src; p.csr_type = plan->csr_type; p.sqrt_deg = plan->sqrt_deg;
    p.P = P; p.Q = Q; p.X = X; p.S = S; p.Rp = Rp; p.a = a_att;
    p.slope = leaky_slope; p.d4 = d / 4;
    p.agg = agg; p.lse = lse; p.G = G; p.dP = dP;
    p.ew = reinterpret_cast<float2*>(erec);
    p.mask = reinterpret_cast<char*>(erec) + sizeof(float2) * (size_t)plan->E;
    p.da = da_att;
    p.partial = nullptr;
    if (plan->dst_items.n_slots > 0) {
        int rc = plan_scratch(plan, sizeof(float) * (size_t)plan->dst_items.n_slots * (size_t)d, &p.partial);
        if (rc) return rc;
    }
    if (p.n_items == 0) return 0;
    const int threads = 256;
    JMAC_DISPATCH_VPL(d, {
        constexpr int U = VPL == 1 ? 4 : 2;
        auto kern = edge_bwd_dst_kernel<VPL, U>;
        int blocks = persistent_blocks(kern, threads, p.n_items);
        kern<<<blocks, threads, 0, st>>>(p);
        JMAC_LAUNCH_CHECK();
        if (plan->dst_items.n_splits > 0) {
            int mb = (plan->dst_items.n_splits * 32 + threads - 1) / threads;
            sum_partials_kernel<VPL><<<mb, threads, 0, st>>>(plan->dst_items.splits, plan->dst_items.n_splits, p.partial, p.d4, 1, dP, nullptr);
            JMAC_LAUNCH_CHECK();
        }
    });
    return 0;
}

static int gather_pass(jmac_plan* plan, int mode, const float* G, const void* erec, const float* a_att, float slope, int d,
                       float* out_w, float* out_z, cudaStream_t st) {
    JMAC_ARG(plan && G && a_att && out_w && out_z);
    JMAC_ARG(plan->E == 0 || erec != nullptr);
    JMAC_ARG(dim_ok(d));
    const ItemList& il = mode == 0 ? plan->src_items : plan->typ_items;
    EdgeBwdGatherParams p;
    p.items = il.items; p.n_items = il.n_items;
    p.pos = mode == 0 ? plan->csc_pos : plan->typ_pos;
    p.dst = mode == 0 ? plan->csc_dst : plan->typ_dst;
    p.G = G;
    p.ew = reinterpret_cast<const float2*>(erec);
    p.mask = reinterpret_cast<const char*>(erec) + sizeof(float2) * (size_t)plan->E;
    p.a = a_att; p.slope = slope; p.d4 = d / 4;
    p.out_w = out_w; p.out_z = out_z;
    p.partial = nullptr;
    if (mode == 1) {
        JMAC_CUDA(cudaMemsetAsync(out_w, 0, sizeof(float) * (size_t)plan->n_rel * d, st));
        JMAC_CUDA(cudaMemsetAsync(out_z, 0, sizeof(float) * (size_t)plan->n_rel * d, st));
    } else if (il.n_slots > 0) {
        int rc = plan_scratch(plan, sizeof(float) * (size_t)il.n_slots * 2 * (size_t)d, &p.partial);
        if (rc) return rc;
    }
    if (p.n_items == 0) return 0;
    const int threads = 256;
    JMAC_DISPATCH_VPL(d, {
        constexpr int U = VPL == 1 ? 4 : (VPL == 2 ? 4 : 2);
        if (mode == 0) {
            auto kern = edge_bwd_gather_kernel<VPL, 0, U>;
            int blocks = persistent_blocks(kern, threads, p.n_items);
            kern<<<blocks, threads, 0, st>>>(p);
            JMAC_LAUNCH_CHECK();
            if (il.n_splits > 0) {
                int mb = (il.n_splits * 32 + threads - 1) / threads;
                sum_partials_kernel<VPL><<<mb, threads, 0, st>>>(il.splits, il.n_splits, p.partial, p.d4, 2, out_w, out_z);
                JMAC_LAUNCH_CHECK();
            }
        } else {
            auto kern = edge_bwd_gather_kernel<VPL, 1, U>;
            int blocks = persistent_blocks(kern, threads, p.n_items);
            kern<<<blocks, threads, 0, st>>>(p);
            JMAC_LAUNCH_CHECK();
        }
    });
    return 0;
}

extern "C" int jmac_edge_bwd_src(jmac_plan* plan, const float* G, const void* erec, const float* a_att, float leaky_slope, int d,
                                 float* dXe, float* dQ, void* stream) {
    return gather_pass(plan, 0, G, erec, a_att, leaky_slope, d, dXe, dQ, (cudaStream_t)stream);
}

extern "C" int jmac_edge_bwd_rel(jmac_plan* plan, const float* G, const void* erec, const float* a_att, float leaky_slope, int d,
                                 float* dRp, float* dS, void* stream) {
    return gather_pass(plan, 1, G, erec, a_att, leaky_slope, d, dRp, dS, (cudaStream_t)stream);
}
