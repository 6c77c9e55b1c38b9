// Edge phase of RelationAwareLayer.forward (comp_op='sub'), CSR-by-destination, one warp per work
// item (a destination segment or a <=chunk slice of a hub segment).
//
// Reference being replaced: src/jmac_model.py:49,56-89 (message: gather x_i/x_j/rel, compose,
// attention logit, message transform) + modules/helper/message_passing.py:4-29 (scatter_softmax,
// division by edge_norm, weighted scatter-sum).  The E-row GEMMs of the reference are hoisted to
// node level by the caller (P = X W_top, Q = X W_bot, S = R' W_bot); what is left per edge is two
// gathered node rows (Q[j], X[j]), two cache-resident relation rows (S[t], R'[t]), a d-length dot,
// an online segmented softmax and a weighted segmented sum: HBM-bound gather/reduce.
//
// Data layout: every table is fp32 row-major [rows, d]; a warp reads a row as VPL x float4 per lane
// (512 B per load instruction, fully coalesced).  Algorithmic HBM bytes: 2*d*4 + 8 per edge,
// 2*d*4 + 28 per destination.
#include <math.h>

#include "common.cuh"

namespace jmac {

struct EdgeFwdParams {
    const Item* items;
    int n_items;
    int E;
    const int* csr_src;
    const int* csr_type;
    const float* sqrt_deg;
    const float *P, *Q, *X, *S, *Rp, *a;
    float slope;
    int d4;
    float* agg;
    float* lse;
    float* partial;  // [n_slots, d + 4]: (m, l, -, -) then the un-normalised weighted sum
    // optional fused self-loop term: T[i] = agg[i] + Xself[i] - loop_row  (all three null or all set)
    const float* Xself;
    const float* loop_row;
    float* T;
    // optional per-edge record for backward (both null when no gradient is needed)
    void* mask;      // [E][32 lanes] sign mask of z_e
    float* logit;    // [E] attention logit s_e
};

// T = agg + X[i] - R'[loop]: the operand of the message GEMM ((neigh + self) W_g, src/jmac_model.py:49-52)
template <int VPL>
__device__ __forceinline__ void store_self_fused(const EdgeFwdParams& p, const Row<VPL>& agg_row, int seg, int d4, int lane) {
    Row<VPL> xs, lr;
    load_row<VPL, true>(xs, p.Xself, seg, d4, lane);
    load_row<VPL, false>(lr, p.loop_row, 0, d4, lane);
#pragma unroll
    for (int i = 0; i < VPL; ++i) xs.v[i] = agg_row.v[i] + xs.v[i] - lr.v[i];
    store_row<VPL>(xs, p.T, seg, d4, lane);
}

// a . lrelu(z) for z = P[i] + Q[j] - S[t], and the sign mask of z (bit 4*i + c <=> component c of float4 #i is > 0).
// Works on nz = S - (P + Q) = -z (exact): the IEEE sign bit of nz is set exactly when z > 0 (z == 0 gives +0, "not
// positive", the reference's `z > 0` test), so the mask is built by shifting one sign bit per feature into a
// register (one funnel shift each), and -lrelu(z) = min(nz, slope * nz) for slope <= 1 (max otherwise) is one
// multiply + one min/max: 6 instructions per feature instead of 8, same rounding as the select form.
template <int VPL>
__device__ __forceinline__ float leaky_dot(const Row<VPL>& pi, const Row<VPL>& q, const Row<VPL>& s, const Row<VPL>& a, float slope,
                                           unsigned& bits) {
    float part = 0.f;
    unsigned b = 0u;
    const bool use_min = slope <= 1.f;
#pragma unroll
    for (int i = VPL - 1; i >= 0; --i) {
        const float4 nz = s.v[i] - (pi.v[i] + q.v[i]);
        b = __funnelshift_l(__float_as_uint(nz.w), b, 1);
        b = __funnelshift_l(__float_as_uint(nz.z), b, 1);
        b = __funnelshift_l(__float_as_uint(nz.y), b, 1);
        b = __funnelshift_l(__float_as_uint(nz.x), b, 1);
        float4 m;
        m.x = use_min ? fminf(nz.x, slope * nz.x) : fmaxf(nz.x, slope * nz.x);
        m.y = use_min ? fminf(nz.y, slope * nz.y) : fmaxf(nz.y, slope * nz.y);
        m.z = use_min ? fminf(nz.z, slope * nz.z) : fmaxf(nz.z, slope * nz.z);
        m.w = use_min ? fminf(nz.w, slope * nz.w) : fmaxf(nz.w, slope * nz.w);
        float pp = -a.v[i].x * m.x;                     // one short chain per float4, summed at the end
        pp = fmaf(-a.v[i].y, m.y, pp);
        pp = fmaf(-a.v[i].z, m.z, pp);
        pp = fmaf(-a.v[i].w, m.w, pp);
        part += pp;
    }
    bits = b;
    return part;
}

// Sign mask of one edge: lane l owns the (VPL <= 2 ? 1 : 2) bytes at offset l of the edge's 32/64-byte record.
// Packed through two (one) shuffles so that 8 (16) lanes issue 32-bit stores instead of 32 byte stores.
template <int VPL>
__device__ __forceinline__ void store_mask(void* mask, int edge, unsigned bits, int lane) {
    if (VPL <= 2) {
        unsigned v = bits | (__shfl_down_sync(FULL, bits, 1) << 8);
        v |= __shfl_down_sync(FULL, v, 2) << 16;
        if ((lane & 3) == 0) reinterpret_cast<unsigned*>(mask)[(size_t)edge * 8 + (lane >> 2)] = v;
    } else {
        const unsigned v = bits | (__shfl_down_sync(FULL, bits, 1) << 16);
        if ((lane & 1) == 0) reinterpret_cast<unsigned*>(mask)[(size_t)edge * 16 + (lane >> 1)] = v;
    }
}

template <int VPL, int U>
__global__ void __launch_bounds__(256) edge_fwd_kernel(EdgeFwdParams p) {
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(FULL, (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), 0);   // warp-uniform for the compiler
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int d4 = p.d4;
    const bool save = p.logit != nullptr;
    Row<VPL> av;
    load_row<VPL, false>(av, p.a, 0, d4, lane);

    for (int it = warp; it < p.n_items; it += nwarps) {
        const Item item = p.items[it];
        Row<VPL> pi;
        load_row<VPL, true>(pi, p.P, item.seg, d4, lane);
        float m = -INFINITY, l = 0.f;
        Row<VPL> acc;
        acc.zero();
        for (int base = item.beg; base < item.end; base += 32) {
            const int cnt = min(32, item.end - base);
            int jv = 0, tv = 0;
            if (lane < cnt) {
                jv = __ldg(p.csr_src + base + lane);
                tv = __ldg(p.csr_type + base + lane);
            }
            float sc_keep = 0.f;                      // lane k keeps the logit of edge base+k
            int k = 0;
            for (; k + U <= cnt; k += U) {
                Row<VPL> q[U], x[U], s[U], r[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = __shfl_sync(FULL, jv, k + u);
                    const int t = __shfl_sync(FULL, tv, k + u);
                    load_row<VPL, true>(q[u], p.Q, j, d4, lane);
                    load_row<VPL, true>(x[u], p.X, j, d4, lane);
                    load_row<VPL, false>(s[u], p.S, t, d4, lane);
                    load_row<VPL, false>(r[u], p.Rp, t, d4, lane);
                }
                float sc[U];
                unsigned bits[U];
#pragma unroll
                for (int u = 0; u < U; ++u) sc[u] = leaky_dot<VPL>(pi, q[u], s[u], av, p.slope, bits[u]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                    for (int u = 0; u < U; ++u) sc[u] += __shfl_xor_sync(FULL, sc[u], o);
                }
                if (save) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        store_mask<VPL>(p.mask, base + k + u, bits[u], lane);
                        if (lane == k + u) sc_keep = sc[u];
                    }
                }
                float mn = m;
#pragma unroll
                for (int u = 0; u < U; ++u) mn = fmaxf(mn, sc[u]);
                const float corr = __expf(m - mn);
                l *= corr;
#pragma unroll
                for (int i = 0; i < VPL; ++i) acc.v[i] = corr * acc.v[i];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float pe = __expf(sc[u] - mn);
                    l += pe;
#pragma unroll
                    for (int i = 0; i < VPL; ++i) fma4(acc.v[i], pe, x[u].v[i] - r[u].v[i]);
                }
                m = mn;
            }
            for (; k < cnt; ++k) {
                const int j = __shfl_sync(FULL, jv, k);
                const int t = __shfl_sync(FULL, tv, k);
                Row<VPL> q, x, s, r;
                load_row<VPL, true>(q, p.Q, j, d4, lane);
                load_row<VPL, true>(x, p.X, j, d4, lane);
                load_row<VPL, false>(s, p.S, t, d4, lane);
                load_row<VPL, false>(r, p.Rp, t, d4, lane);
                unsigned bits1;
                const float sc = warp_sum(leaky_dot<VPL>(pi, q, s, av, p.slope, bits1));
                if (save) {
                    store_mask<VPL>(p.mask, base + k, bits1, lane);
                    if (lane == k) sc_keep = sc;
                }
                const float mn = fmaxf(m, sc);
                const float corr = __expf(m - mn);
                const float pe = __expf(sc - mn);
                l = l * corr + pe;
#pragma unroll
                for (int i = 0; i < VPL; ++i) {
                    acc.v[i] = corr * acc.v[i];
                    fma4(acc.v[i], pe, x.v[i] - r.v[i]);
                }
                m = mn;
            }
            if (save && lane < cnt) p.logit[base + lane] = sc_keep;
        }
        if (item.slot < 0) {
            const float c = p.sqrt_deg[item.seg];
            const float inv = l > 0.f ? c / l : 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) acc.v[i] = inv * acc.v[i];
            store_row<VPL>(acc, p.agg, item.seg, d4, lane);
            if (p.T) store_self_fused<VPL>(p, acc, item.seg, d4, lane);
            if (lane == 0) p.lse[item.seg] = l > 0.f ? m + logf(l) : 0.f;
        } else {
            float* ps = p.partial + (size_t)item.slot * (size_t)(d4 * 4 + 4);
            if (lane == 0) *reinterpret_cast<float4*>(ps) = make_float4(m, l, 0.f, 0.f);
            store_row<VPL>(acc, ps + 4, 0, d4, lane);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Same computation with the row gathers moved off the register file (the default): every warp owns a ring
// of RING_S shared-memory slots; lane 0 issues one cp.async.bulk (TMA bulk copy, 1 instruction per row) for
// each of Q[j], X[j] of an edge into a slot and the whole warp waits on the slot's mbarrier RING_S-1 edges
// later.  Memory latency is hidden by bytes in flight in shared memory (16 warps x 6 slots x 2 KB = 192 KB
// per SM) instead of by occupancy; the LSU only sees conflict-free LDS.128.  The relation rows S[t], R'[t]
// are L2-resident: plain loads issued before the wait on the ring.
// ------------------------------------------------------------------------------------------------
constexpr int RING_ROWS = 2;               // slot = {Q[j], X[j]}
// ring depth by row width: 16 warps x 6 x 2 KB = 192 KB at d = 256; wider rows (VPL >= 3, e.g. JMAC_DBPv1's d = 300) get 4 slots
template <int VPL> struct FwdRing { static constexpr int S = VPL <= 2 ? 6 : 4; };

// one edge of the ring: logit from the shared-memory copies of Q[j] (row 0 of the slot) and the relation row S[t]
template <int VPL>
__device__ __forceinline__ float ring_logit(const Row<VPL>& pi, const float* slot, const Row<VPL>& sr, const Row<VPL>& av, float slope,
                                            int d4, int lane, unsigned& bits) {
    Row<VPL> q;
    lds_row<VPL>(q, slot, d4, lane);
    return leaky_dot<VPL>(pi, q, sr, av, slope, bits);
}

// Work decomposition of the ring kernels: every warp owns a CONTIGUOUS run of work items whose edges form one
// contiguous stream of CSR positions (~E / #warps edges per warp), so the index batches and the gather ring run
// ahead across segment boundaries -- the per-segment start-up latency of short segments (mean in-degree 3-20
// on KGs) is paid once per warp instead of once per destination.
__device__ __forceinline__ int first_item_at_or_after(const Item* items, int n_items, int pos) {
    int lo = 0, hi = n_items;                      // items are sorted by beg
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(&items[mid].beg) < pos) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// FULLROW: d == 128 * VPL, every lane owns VPL complete float4 -- the row loads lose their tail predicates (and the
// zero-fill moves that go with them), ~10 % fewer instructions per edge in these issue-bound loops.
template <int VPL, bool FULLROW>
__global__ void __launch_bounds__(RING_WARPS * 32, 1) edge_fwd_ring_kernel(EdgeFwdParams p) {
    extern __shared__ __align__(16) uint8_t ring_raw[];
    constexpr int RING_S = FwdRing<VPL>::S;
    const int lane = threadIdx.x & 31, wib = __shfl_sync(FULL, threadIdx.x >> 5, 0);   // broadcast: lets the compiler see a warp-uniform value
    const int warp = blockIdx.x * RING_WARPS + wib;
    const int nwarps = gridDim.x * RING_WARPS;
    const int d4 = FULLROW ? VPL * 32 : p.d4;
    const uint32_t row_bytes = (uint32_t)d4 * 16;
    const int row_f = d4 * 4;                                   // floats per row
    float* slots = reinterpret_cast<float*>(ring_raw) + (size_t)wib * RING_S * RING_ROWS * row_f;
    const uint32_t slots_s = smem_addr(slots);
    const uint32_t bars = smem_addr(ring_raw) + (uint32_t)RING_WARPS * RING_S * RING_ROWS * row_bytes + (uint32_t)wib * RING_S * 8;
    if (lane == 0) {
        for (int s = 0; s < RING_S; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const bool save = p.logit != nullptr;
    Row<VPL> av;
    load_row<VPL, false>(av, p.a, 0, d4, lane);

    // this warp's run of items: those starting in [warp * per, (warp + 1) * per); the last warp also takes the
    // trailing empty segments
    const int per = (p.E + nwarps - 1) / nwarps;
    const int it0 = first_item_at_or_after(p.items, p.n_items, min(p.E, warp * per));
    const int it1 = warp == nwarps - 1 ? p.n_items : first_item_at_or_after(p.items, p.n_items, min(p.E, (warp + 1) * per));
    if (it0 >= it1) return;
    const int stream_end = p.items[it1 - 1].end;
    constexpr int ADV = (32 - RING_S) & ~1;
    uint32_t issue_slot = 0, use_slot = 0, use_phase = 0;         // ring positions, maintained incrementally (no modulo in the loops)
    Item item = p.items[it0];
    Row<VPL> pi;
    load_row<VPL, true>(pi, p.P, item.seg, d4, lane);
    int batch_base = item.beg, issued = item.beg;                 // absolute CSR positions
    int jv = 0, tv = 0;
    if (batch_base + lane < stream_end) {
        jv = __ldg(p.csr_src + batch_base + lane);
        tv = __ldg(p.csr_type + batch_base + lane);
    }
    prefetch_idx_l2(p.csr_src, p.csr_type, nullptr, batch_base + 32, stream_end, lane);
    prefetch_idx_l2(p.csr_src, p.csr_type, nullptr, batch_base + 64, stream_end, lane);
    float sc_keep = 0.f;
    // The relation rows S[t] of a pair come from L2 (~700 cycles under load) and feed the logits as soon as the pair's gathered rows have
    // landed: they are loaded one pair ahead (by stream position -- the stream runs across segment boundaries) and carried in registers;
    // the R'[t] rows, first used after the softmax update, are loaded at the top of their own iteration.  (Rows of up to 8 registers;
    // wider ones would spill and load at the point of use.)
    constexpr bool REL_AHEAD = VPL <= 2;
    Row<REL_AHEAD ? VPL : 1> sn0, sn1;
    if (REL_AHEAD) {
        load_row<REL_AHEAD ? VPL : 1, false>(sn0, p.S, __shfl_sync(FULL, tv, 0), d4, lane);
        load_row<REL_AHEAD ? VPL : 1, false>(sn1, p.S, __shfl_sync(FULL, tv, 1), d4, lane);
    }

    for (int it = it0; it < it1; ++it) {
        // next segment's descriptor and P row are fetched while this one is processed
        Item nxt = item;
        Row<VPL> pn;
        pn.zero();
        if (it + 1 < it1) {
            nxt = p.items[it + 1];
            load_row<VPL, true>(pn, p.P, nxt.seg, d4, lane);
            // the epilogue rows of the NEXT segment (its own X row for the fused self-loop term, its sqrt(deg)) are pulled towards L2 now:
            // they are read at the very end of a segment and used at once, i.e. at full DRAM latency otherwise
            if (p.T) {
                const int lines = (row_f * 4 + 127) >> 7;
                if (lane < lines) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.Xself + (size_t)nxt.seg * row_f + lane * 32));
            }
            if (lane == 31) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.sqrt_deg + nxt.seg));
        }
        float m = -INFINITY, l = 0.f;
        Row<VPL> acc;
        acc.zero();
        for (int k = item.beg; k < item.end; k += 2) {
            if (k - batch_base >= ADV) {                          // next index batch (warp-uniform)
                if (save && batch_base + lane < k) p.logit[batch_base + lane] = sc_keep;
                batch_base = k;
                jv = tv = 0;
                if (batch_base + lane < stream_end) {
                    jv = __ldg(p.csr_src + batch_base + lane);
                    tv = __ldg(p.csr_type + batch_base + lane);
                }
                prefetch_idx_l2(p.csr_src, p.csr_type, nullptr, batch_base + 96, stream_end, lane);       // a batch advances < 32 entries: every line is covered
            }
            // keep the ring full: issue the Q[j], X[j] gathers up to position k + RING_S - 1 (may belong to later segments)
            const int want = min(stream_end, k + RING_S);
            while (issued + 1 < want) {                           // warp-uniform; steady state: one pair per iteration
                const int j0 = __shfl_sync(FULL, jv, issued - batch_base), j1 = __shfl_sync(FULL, jv, issued + 1 - batch_base);
                const uint32_t s0 = issue_slot, s1 = s0 + 1 == RING_S ? 0 : s0 + 1;
                issue_slot = s1 + 1 == RING_S ? 0 : s1 + 1;
                if (lane == 0) {
                    const uint32_t bar0 = bars + 8 * s0, dst0 = slots_s + s0 * RING_ROWS * row_bytes;
                    const uint32_t bar1 = bars + 8 * s1, dst1 = slots_s + s1 * RING_ROWS * row_bytes;
                    mbar_expect_tx(bar0, RING_ROWS * row_bytes);
                    bulk_row(dst0, p.Q + (size_t)j0 * row_f, row_bytes, bar0);
                    bulk_row(dst0 + row_bytes, p.X + (size_t)j0 * row_f, row_bytes, bar0);
                    mbar_expect_tx(bar1, RING_ROWS * row_bytes);
                    bulk_row(dst1, p.Q + (size_t)j1 * row_f, row_bytes, bar1);
                    bulk_row(dst1 + row_bytes, p.X + (size_t)j1 * row_f, row_bytes, bar1);
                }
                issued += 2;
            }
            if (issued < want) {
                const int j = __shfl_sync(FULL, jv, issued - batch_base);
                const uint32_t s = issue_slot;
                issue_slot = s + 1 == RING_S ? 0 : s + 1;
                if (lane == 0) {
                    const uint32_t bar = bars + 8 * s, dst = slots_s + s * RING_ROWS * row_bytes;
                    mbar_expect_tx(bar, RING_ROWS * row_bytes);
                    bulk_row(dst, p.Q + (size_t)j * row_f, row_bytes, bar);
                    bulk_row(dst + row_bytes, p.X + (size_t)j * row_f, row_bytes, bar);
                }
                ++issued;
            }
            const bool two = k + 1 < item.end;                    // warp-uniform
            const int kb = k - batch_base;
            // relation rows (L2-resident) of the two edges: plain loads, in flight while we wait on the ring
            const int t0 = __shfl_sync(FULL, tv, kb);
            const int t1 = __shfl_sync(FULL, tv, two ? kb + 1 : kb);
            Row<VPL> s0, r0, s1, r1;
            if constexpr (REL_AHEAD) {
                s0 = sn0;
                s1 = sn1;                                         // s1 of an unpaired last edge is some other (finite) row: its logit is replaced by -inf below
                const int nb = kb + (two ? 2 : 1);                // the next pair of the stream; kb < ADV, so nb + 1 < 32 (lanes past the stream hold type 0)
                load_row<VPL, false>(sn0, p.S, __shfl_sync(FULL, tv, nb), d4, lane);
                load_row<VPL, false>(sn1, p.S, __shfl_sync(FULL, tv, nb + 1), d4, lane);
            } else {
                load_row<VPL, false>(s0, p.S, t0, d4, lane);
                load_row<VPL, false>(s1, p.S, t1, d4, lane);
            }
            load_row<VPL, false>(r0, p.Rp, t0, d4, lane);
            load_row<VPL, false>(r1, p.Rp, t1, d4, lane);
            const uint32_t sa = use_slot, sb = sa + 1 == RING_S ? 0 : sa + 1;
            const uint32_t pb = sa + 1 == RING_S ? use_phase ^ 1 : use_phase;
            mbar_wait(bars + 8 * sa, use_phase);
            if (two) mbar_wait(bars + 8 * sb, pb);
            const float* sl0 = slots + (size_t)sa * RING_ROWS * row_f;
            const float* sl1 = slots + (size_t)(two ? sb : sa) * RING_ROWS * row_f;
            unsigned b0, b1;
            float sc0 = ring_logit<VPL>(pi, sl0, s0, av, p.slope, d4, lane, b0);
            float sc1 = ring_logit<VPL>(pi, sl1, s1, av, p.slope, d4, lane, b1);
            Row<VPL> x0, x1;
            lds_row<VPL>(x0, sl0 + row_f, d4, lane);
            lds_row<VPL>(x1, sl1 + row_f, d4, lane);
            __syncwarp();                                         // every lane has its copies: the slots may be refilled
            if (two) { use_slot = sb + 1 == RING_S ? 0 : sb + 1; use_phase = sb + 1 == RING_S ? pb ^ 1 : pb; }
            else { use_slot = sb; use_phase = pb; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sc0 += __shfl_xor_sync(FULL, sc0, o);
                sc1 += __shfl_xor_sync(FULL, sc1, o);
            }
            if (save) {
                store_mask<VPL>(p.mask, k, b0, lane);
                if (lane == kb) sc_keep = sc0;
                if (two) {
                    store_mask<VPL>(p.mask, k + 1, b1, lane);
                    if (lane == kb + 1) sc_keep = sc1;
                }
            }
            if (!two) sc1 = -INFINITY;                             // contributes exp(-inf) = 0
            const float mn = fmaxf(m, fmaxf(sc0, sc1));
            const float corr = __expf(m - mn);
            const float pe0 = __expf(sc0 - mn), pe1 = __expf(sc1 - mn);
            l = l * corr + pe0 + pe1;
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                acc.v[i] = corr * acc.v[i];
                fma4(acc.v[i], pe0, x0.v[i] - r0.v[i]);
                fma4(acc.v[i], pe1, x1.v[i] - r1.v[i]);
            }
            m = mn;
        }
        if (item.slot < 0) {
            const float c = p.sqrt_deg[item.seg];
            const float inv = l > 0.f ? c / l : 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) acc.v[i] = inv * acc.v[i];
            store_row<VPL>(acc, p.agg, item.seg, d4, lane);
            if (p.T) store_self_fused<VPL>(p, acc, item.seg, d4, lane);
            if (lane == 0) p.lse[item.seg] = l > 0.f ? m + logf(l) : 0.f;
        } else {
            float* ps = p.partial + (size_t)item.slot * (size_t)(d4 * 4 + 4);
            if (lane == 0) *reinterpret_cast<float4*>(ps) = make_float4(m, l, 0.f, 0.f);
            store_row<VPL>(acc, ps + 4, 0, d4, lane);
        }
        item = nxt;
        pi = pn;
    }
    if (save && batch_base + lane < stream_end) p.logit[batch_base + lane] = sc_keep;
}

// Two-level softmax merge for split (hub) destinations: one CTA of 8 warps per split segment.  Warp w combines the slots
// w, w + 8, ... (a hub of the reference's graphs has up to ~90 slots of 32 edges; walking them with one warp serialised ~90
// dependent row loads), the eight warp results are then added in a fixed order (deterministic).
template <int VPL>
__global__ void __launch_bounds__(256) edge_fwd_merge_kernel(EdgeFwdParams fp, const Split* __restrict__ splits, int n_splits,
                                                             const float* __restrict__ partial, const float* __restrict__ sqrt_deg,
                                                             int d4, float* __restrict__ agg, float* __restrict__ lse) {
    __shared__ float s_max[8], s_l[8];
    __shared__ float4 s_acc[8][VPL * 32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if ((int)blockIdx.x >= n_splits) return;
    const Split sp = splits[blockIdx.x];
    const size_t stride = (size_t)(d4 * 4 + 4);
    const float* base = partial + (size_t)sp.first * stride;
    float M = -INFINITY;
    for (int k = threadIdx.x; k < sp.count; k += 256) M = fmaxf(M, base[k * stride]);
    M = warp_max(M);
    if (lane == 0) s_max[wib] = M;
    __syncthreads();
    M = s_max[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) M = fmaxf(M, s_max[w]);
    float L = 0.f;
    Row<VPL> acc;
    acc.zero();
    for (int k = wib; k < sp.count; k += 8) {
        const float4 ml = *reinterpret_cast<const float4*>(base + k * stride);
        const float sc = ml.y > 0.f ? __expf(ml.x - M) : 0.f;
        L += sc * ml.y;
        Row<VPL> part;
        load_row<VPL, true>(part, base + k * stride + 4, 0, d4, lane);
#pragma unroll
        for (int i = 0; i < VPL; ++i) fma4(acc.v[i], sc, part.v[i]);
    }
    if (lane == 0) s_l[wib] = L;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s_acc[wib][i * 32 + lane] = acc.v[i];
    __syncthreads();
    if (wib == 0) {
        L = s_l[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) L += s_l[w];
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            float4 t = s_acc[0][i * 32 + lane];
#pragma unroll
            for (int w = 1; w < 8; ++w) t = t + s_acc[w][i * 32 + lane];
            acc.v[i] = t;
        }
        const float inv = L > 0.f ? sqrt_deg[sp.seg] / L : 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) acc.v[i] = inv * acc.v[i];
        store_row<VPL>(acc, agg, sp.seg, d4, lane);
        if (fp.T) store_self_fused<VPL>(fp, acc, sp.seg, d4, lane);
        if (lane == 0) lse[sp.seg] = L > 0.f ? M + logf(L) : 0.f;
    }
}

}  // namespace jmac

using namespace jmac;

extern "C" int jmac_edge_fwd(jmac_plan* plan, const float* P, const float* Q, const float* X, const float* S, const float* Rp,
                             const float* a_att, float leaky_slope, int d, float* agg, float* lse, void* stream) {
    return jmac_edge_fwd_fused(plan, P, Q, X, S, Rp, a_att, leaky_slope, d, agg, lse, nullptr, nullptr, nullptr, stream);
}

extern "C" int jmac_edge_fwd_fused(jmac_plan* plan, const float* P, const float* Q, const float* X, const float* S, const float* Rp,
                                   const float* a_att, float leaky_slope, int d, float* agg, float* lse, const float* Xself, float* T,
                                   void* erec, void* stream) {
    JMAC_ARG(plan && P && Q && X && S && Rp && a_att && agg && lse);
    JMAC_ARG((Xself == nullptr) == (T == nullptr));
    JMAC_ARG(dim_ok(d));
    cudaStream_t st = (cudaStream_t)stream;
    EdgeFwdParams p;
    JMAC_ARG((reinterpret_cast<uintptr_t>(erec) & 31) == 0);
    p.items = plan->dst_items.items;
    p.n_items = plan->dst_items.n_items;
    p.E = (int)plan->E;
    p.csr_src = plan->csr_src;
    p.csr_type = plan->csr_type;
    p.sqrt_deg = plan->sqrt_deg;
    p.P = P; p.Q = Q; p.X = X; p.S = S; p.Rp = Rp; p.a = a_att;
    p.slope = leaky_slope;
    p.d4 = d / 4;
    p.agg = agg; p.lse = lse;
    p.Xself = Xself; p.T = T;
    p.mask = erec ? erec_mask(erec, plan->E) : nullptr;
    p.logit = erec ? erec_logit(erec, plan->E, d) : nullptr;
    p.loop_row = Rp + (size_t)(plan->n_rel - 1) * d;     // the self-loop relation is the last row of R'
    p.partial = nullptr;
    if (plan->dst_items.n_slots > 0) {
        int rc = plan_scratch(plan, sizeof(float) * (size_t)plan->dst_items.n_slots * (size_t)(d + 4), &p.partial, st);
        if (rc) return rc;
    }
    if (p.n_items == 0) return 0;
    const int threads = 256;
    // The bulk-copy ring kernel is the default (C4 forward: 15.9-16.6 ms against ~20 ms for the register-staged one,
    // which can keep only ~64 KB of gathers in flight per SM -- DESIGN.md section 4); JMAC_EDGE_RING=0 selects the
    // register-staged kernel, which is also the fallback when a ring of d-wide rows does not fit in shared memory.
    static const int use_ring = getenv("JMAC_EDGE_RING") ? atoi(getenv("JMAC_EDGE_RING")) : 1;
    JMAC_DISPATCH_VPL(d, {
        const size_t ring_smem = (size_t)RING_WARPS * FwdRing<VPL>::S * (RING_ROWS * (size_t)d * 4 + 8);
        constexpr int U = VPL == 1 ? 4 : 2;
        if (use_ring && plan->E > 0 && ring_smem <= 220 * 1024) {
            // one CTA of RING_WARPS warps per SM
            auto kern = d == VPL * 128 ? edge_fwd_ring_kernel<VPL, true> : edge_fwd_ring_kernel<VPL, false>;
            JMAC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_smem));
            long long want = ((long long)p.n_items + RING_WARPS - 1) / RING_WARPS;
            int blocks = (int)(want < sm_count() ? want : sm_count());
            kern<<<blocks, RING_WARPS * 32, ring_smem, st>>>(p);
        } else {
            auto kern = edge_fwd_kernel<VPL, U>;
            int blocks = persistent_blocks(kern, threads, p.n_items);
            kern<<<blocks, threads, 0, st>>>(p);
        }
        JMAC_LAUNCH_CHECK();
        if (plan->dst_items.n_splits > 0) {
            edge_fwd_merge_kernel<VPL><<<plan->dst_items.n_splits, threads, 0, st>>>(p, plan->dst_items.splits, plan->dst_items.n_splits, p.partial,
                                                                plan->sqrt_deg, p.d4, agg, lse);
            JMAC_LAUNCH_CHECK();
        }
    });
    plan_mark_use(plan, st);
    return 0;
}
