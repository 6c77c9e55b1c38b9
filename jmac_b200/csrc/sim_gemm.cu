// Alignment similarity: S = E1 . E2^T as a bf16 tcgen05 GEMM (fp32 accumulation in TMEM) with a fused
// STREAMING epilogue -- the n1 x n2 matrix is never written.  One kernel serves
//   * sim + calculate_nearest_k (modules/finding/similarity.py:13-84): row top-k values (-> r1, r2),
//   * csls_sim + calculate_rank   (similarity.py:58-78, alignment.py:87-112): v = 2S - r1[i] - r2[j],
//     rank of the gold column by counting v > v_gold, optional top-k of v,
//   * get_neg (modules/utils/util.py:31-54): row top-k indices,
//   * the diagonal / gold-column raw similarity (needed to form v_gold bit-consistently).
//
// Shape of the machine (B200, sm_100a):
//   CTA = 320 threads: warp 0 TMA producer, warp 1 tcgen05.mma issuer (+TMEM allocator), warps 2..9
//   epilogue (one thread per output row, 256 rows per CTA).
//   A block: 256 rows of E1 x K(<=256) bf16, resident in shared memory for a whole row sweep (128 KB).
//   B tiles: 128 rows of E2, streamed as 64-wide K slices through a 6-slot TMA ring (16 KB each).
//   MMA: cta_group::1, M=128 (two row halves), N=128, K=16, accumulators double-buffered in TMEM
//   (2 stages x 2 halves x 128 columns = all 512 columns).
//   Per 256x128 output tile: 2*16 MMAs = 2048 tensor cycles; B traffic 32 B/cycle/SM (L2-resident).
//   The epilogue owns one row per thread, keeps its running top-k in registers across the whole sweep
//   (no cross-CTA merge) and reads accumulators with tcgen05.ld 32x32b.x32.
// Shared-memory operands use the canonical K-major SWIZZLE_128B layout written by TMA.
//
// Work decomposition.  The 256x128 output tiles are numbered row-block-major; a work unit is a contiguous run of
// tiles, a "piece" the part of a unit inside one row block (the granularity of A residency and of the per-row
// epilogue state).  Large problems (>= 4 row blocks per SM): unit = one row block, results written directly.
// Small problems -- the sizes the reference actually evaluates, 2.6k-30k rows, where 256-row blocks would occupy
// 11-120 of the 148 SMs -- are cut stream-K style into one equal run of tiles per SM; every piece writes its partial
// top-k / count / softmax statistics to a slot of a caller-provided workspace and sim_merge_kernel combines the (at most
// P) slots of each row in ascending column order (top-k, counts and the gold entry come out bit-identical to an unsplit
// sweep; the softmax statistics are re-associated sums).
//
// Precision.  PREC = 0: bf16 operands, one kind::f16 MMA per K step (C5: 1M x 1M).  PREC = 1: fp32-accurate 3xTF32 --
// operands arrive pre-split (hi = RN-to-TF32(x), lo = RN-to-TF32(x - hi), both exactly representable so the tensor core
// drops nothing), three kind::tf32 MMAs per K step (hi.hi + lo.hi + hi.lo); both operands are streamed through a
// 2-stage ring of 32-wide K slices (no resident A block, any K).  The reference evaluates in f64/f32
// (similarity.py:44-46, util.py:49-53, train.py:239-257): this is the default of jmac_b200.finding.
#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>

#include "common.cuh"

namespace jmac {

constexpr int SIM_BM = 256, SIM_BN = 128, SIM_BK = 64, SIM_NS = 6;
constexpr int SIM_A_SLICE = SIM_BM * SIM_BK * 2;   // 32 KB
constexpr int SIM_B_SLICE = SIM_BN * SIM_BK * 2;   // 16 KB
constexpr int SIM_MAX_KS = 5;                       // K <= 320 (JMAC_DBPv1's d = 300); the A block stays resident
constexpr int SIM_BASE_KS = 4;                      // up to K = 256 the B ring has SIM_NS slots, beyond it two fewer (same total)
constexpr int SIM_THREADS = 320;
constexpr int SIM_SMEM = SIM_BASE_KS * SIM_A_SLICE + SIM_NS * SIM_B_SLICE + 1024 /*align*/ + 256 /*barriers*/ + 2 * SIM_BN * 4 /*r2 ring*/;
static_assert(SIM_SMEM <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
// 3xTF32: a stage holds one 32-wide fp32 K slice of A_hi, A_lo (256 rows x 128 B each) and B_hi, B_lo (128 rows x 128 B each)
constexpr int SIM_T_BK = 32, SIM_T_NS = 2;
constexpr int SIM_T_A = SIM_BM * SIM_T_BK * 4;      // 32 KB
constexpr int SIM_T_B = SIM_BN * SIM_T_BK * 4;      // 16 KB
constexpr int SIM_T_STAGE = 2 * SIM_T_A + 2 * SIM_T_B;
constexpr int SIM_SMEM_T = SIM_T_NS * SIM_T_STAGE + 1024 + 256 + 2 * SIM_BN * 4;
static_assert(SIM_SMEM_T <= 232448, "exceeds the 227 KB per-CTA shared memory limit");

struct SimParams {
    int n1, n2, nks, k;
    float alpha;
    const float* r1;
    const float* r2;
    const int* gold_col;
    const float* gold_val;
    int* gold_cnt;
    float* diag;
    float* topv;
    int* topi;
    // softmax-statistics mode (STATS): r1 = row mask, r2 = column mask (non-zero = masked)
    float masked;
    float *st_m, *st_Z, *st_W;
    int* st_arg;
    // work decomposition: n_units runs of tpc tiles; split != 0: every piece writes to slot (unit - first unit of its row block)
    int n_ct, tpc, n_units, split;
    size_t slot_rows;     // rows per slot of the partial arrays (= n1)
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format: version 1, layout type 2),
// 8-row groups 1024 B apart.  Advancing K by 16 bf16 = +32 B on the start address.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;              // leading byte offset (unused for swizzled K-major) = 1
    d |= (uint64_t)(1024 >> 4) << 32;    // stride byte offset
    d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;              // SWIZZLE_128B
    return d;
}

template <int KMAX>
struct TopK {
    float v[KMAX];
    int i[KMAX];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < KMAX; ++j) { v[j] = -INFINITY; i[j] = -1; }
    }
    // sorted descending; strict '>' keeps the earlier (lower) column on exact ties
    __device__ __forceinline__ void insert(float x, int col) {
#pragma unroll
        for (int j = KMAX - 1; j >= 1; --j) {
            const bool up = x > v[j - 1];
            const bool here = !up && x > v[j];
            i[j] = up ? i[j - 1] : (here ? col : i[j]);
            v[j] = up ? v[j - 1] : (here ? x : v[j]);
        }
        if (x > v[0]) { v[0] = x; i[0] = col; }
    }
};

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// The pieces of this CTA, in the order all three roles walk them.
struct PieceIter {
    int u, t, t_end, T, n_ct, tpc, n_units, stride;
    __device__ __forceinline__ PieceIter(const SimParams& p, int n_rb) {
        n_ct = p.n_ct; tpc = p.tpc; n_units = p.n_units; stride = gridDim.x; T = n_rb * n_ct;
        u = blockIdx.x;
        t = u * tpc;
        t_end = min(t + tpc, T);
    }
    __device__ __forceinline__ bool valid() const { return u < n_units; }
    // current piece: row block rb, column tiles [ct0, ct1)
    __device__ __forceinline__ void piece(int& rb, int& ct0, int& ct1) const {
        rb = t / n_ct;
        ct0 = t - rb * n_ct;
        ct1 = min(n_ct, ct0 + (t_end - t));
    }
    __device__ __forceinline__ int slot(int rb) const { return u - (rb * n_ct) / tpc; }
    __device__ __forceinline__ void next(int ct0, int ct1) {
        t += ct1 - ct0;
        if (t >= t_end) {
            u += stride;
            t = u * tpc;
            t_end = min(t + tpc, T);
        }
    }
};

template <int KMAX, bool CSLS, bool STATS, int PREC>
__global__ void __launch_bounds__(SIM_THREADS, 1)
sim_topk_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_b_lo, SimParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // SWIZZLE_128B needs 1024 B alignment
    const uint32_t smem_a = base;
    // bf16, K <= 256: 4 resident A slices + a 6-slot B ring; K = 320: 5 A slices + 4 B slots (one A slice = two B slices)
    const uint32_t a_slices = p.nks > SIM_BASE_KS ? (uint32_t)p.nks : (uint32_t)SIM_BASE_KS;
    const uint32_t ns = PREC == 0 ? SIM_NS - 2 * (a_slices - SIM_BASE_KS) : (uint32_t)SIM_T_NS;
    const uint32_t smem_b = base + a_slices * SIM_A_SLICE;
    const uint32_t bars = PREC == 0 ? smem_b + ns * SIM_B_SLICE : base + SIM_T_NS * SIM_T_STAGE;
    // barrier map (8 B each)
    const uint32_t full0 = bars, empty0 = bars + 8 * SIM_NS, a_full = bars + 16 * SIM_NS, a_empty = a_full + 8;
    const uint32_t t_full0 = a_empty + 8, t_empty0 = t_full0 + 16, tmem_slot = t_empty0 + 16;
    const uint32_t r2_full0 = tmem_slot + 8, r2_empty0 = r2_full0 + 16, r2_slots = bars + 256;   // 2 x 512 B of r2, staged by TMA

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_rb = (p.n1 + SIM_BM - 1) / SIM_BM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < SIM_NS; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(t_full0 + 8 * s, 1); mbar_init(t_empty0 + 8 * s, 8); }
        for (int s = 0; s < 2; ++s) { mbar_init(r2_full0 + 8 * s, 1); mbar_init(r2_empty0 + 8 * s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0, itp = 0, pcount = 0, tcount = 0;      // it: ring slot, itp: its phase
            for (PieceIter pi(p, n_rb); pi.valid(); ++pcount) {
                int rb, ct0, ct1;
                pi.piece(rb, ct0, ct1);
                if (PREC == 0) {
                    mbar_wait(a_empty, (pcount & 1) ^ 1);
                    mbar_expect_tx(a_full, (uint32_t)p.nks * SIM_A_SLICE);
                    for (int ks = 0; ks < p.nks; ++ks) tma_load_2d(smem_a + ks * SIM_A_SLICE, &map_a, a_full, ks * SIM_BK, rb * SIM_BM);
                }
                for (int ct = ct0; ct < ct1; ++ct, ++tcount) {
                    for (int ks = 0; ks < p.nks; ++ks, it = it + 1 == ns ? 0 : it + 1, itp ^= (it == 0)) {
                        const uint32_t slot = it, ph = itp;
                        mbar_wait(empty0 + 8 * slot, ph ^ 1);
                        if (PREC == 0) {
                            mbar_expect_tx(full0 + 8 * slot, SIM_B_SLICE);
                            tma_load_2d(smem_b + slot * SIM_B_SLICE, &map_b, full0 + 8 * slot, ks * SIM_BK, ct * SIM_BN);
                        } else {
                            const uint32_t st = base + slot * SIM_T_STAGE;
                            mbar_expect_tx(full0 + 8 * slot, SIM_T_STAGE);
                            tma_load_2d(st, &map_a, full0 + 8 * slot, ks * SIM_T_BK, rb * SIM_BM);
                            tma_load_2d(st + SIM_T_A, &map_a_lo, full0 + 8 * slot, ks * SIM_T_BK, rb * SIM_BM);
                            tma_load_2d(st + 2 * SIM_T_A, &map_b, full0 + 8 * slot, ks * SIM_T_BK, ct * SIM_BN);
                            tma_load_2d(st + 2 * SIM_T_A + SIM_T_B, &map_b_lo, full0 + 8 * slot, ks * SIM_T_BK, ct * SIM_BN);
                        }
                    }
                    if (CSLS) {
                        // this tile's 128 column offsets r2[j] -> shared memory (r2 is padded to a multiple of 128).
                        // Issued AFTER the tile's operand slices so that waiting for the epilogue of tile t-2 never delays them.
                        const uint32_t rs = tcount & 1, rph = (tcount >> 1) & 1;
                        mbar_wait(r2_empty0 + 8 * rs, rph ^ 1);
                        mbar_expect_tx(r2_full0 + 8 * rs, SIM_BN * 4);
                        bulk_load_1d(r2_slots + rs * SIM_BN * 4, p.r2 + (size_t)ct * SIM_BN, SIM_BN * 4, r2_full0 + 8 * rs);
                    }
                }
                pi.next(ct0, ct1);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // kind::f16: A=B=BF16 (1); kind::tf32: A=B=TF32 (2); D=F32, both K-major, M=128, N=128
            const uint32_t fmt = PREC == 0 ? 1u : 2u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(SIM_BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint32_t it = 0, itp = 0, tcount = 0, pcount = 0;
            for (PieceIter pi(p, n_rb); pi.valid(); ++pcount) {
                int rb, ct0, ct1;
                pi.piece(rb, ct0, ct1);
                if (PREC == 0) {
                    mbar_wait(a_full, pcount & 1);
                    tc_fence_after();
                }
                for (int ct = ct0; ct < ct1; ++ct, ++tcount) {
                    const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
                    mbar_wait(t_empty0 + 8 * acc, aph ^ 1);
                    tc_fence_after();
                    for (int ks = 0; ks < p.nks; ++ks, it = it + 1 == ns ? 0 : it + 1, itp ^= (it == 0)) {
                        const uint32_t slot = it, ph = itp;
                        mbar_wait(full0 + 8 * slot, ph);
                        tc_fence_after();
                        if (PREC == 0) {
                            const uint64_t bdesc = umma_desc(smem_b + slot * SIM_B_SLICE);
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                const uint64_t adesc = umma_desc(smem_a + ks * SIM_A_SLICE + half * (128 * 128));
                                const uint32_t d_tmem = tmem_base + acc * 256 + half * 128;
#pragma unroll
                                for (int kk = 0; kk < SIM_BK / 16; ++kk)
                                    tc_mma_bf16(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (ks | kk) != 0 ? 1u : 0u);
                            }
                        } else {
                            const uint32_t st = base + slot * SIM_T_STAGE;
                            const uint64_t bhi = umma_desc(st + 2 * SIM_T_A), blo = umma_desc(st + 2 * SIM_T_A + SIM_T_B);
#pragma unroll
                            for (int half = 0; half < 2; ++half) {
                                const uint64_t ahi = umma_desc(st + half * (128 * 128)), alo = umma_desc(st + SIM_T_A + half * (128 * 128));
                                const uint32_t d_tmem = tmem_base + acc * 256 + half * 128;
#pragma unroll
                                for (int kk = 0; kk < SIM_T_BK / 8; ++kk) {          // K = 8 tf32 = 32 B per MMA
                                    const uint64_t o = (uint64_t)(kk * 2);
                                    tc_mma_tf32(d_tmem, alo + o, bhi + o, idesc, (ks | kk) != 0 ? 1u : 0u);   // small terms first
                                    tc_mma_tf32(d_tmem, ahi + o, blo + o, idesc, 1u);
                                    tc_mma_tf32(d_tmem, ahi + o, bhi + o, idesc, 1u);
                                }
                            }
                        }
                        tc_commit(empty0 + 8 * slot);     // slot reusable once these MMAs retire
                    }
                    tc_commit(t_full0 + 8 * acc);         // accumulators of this tile complete
                }
                if (PREC == 0) tc_commit(a_empty);         // A block may be overwritten
                pi.next(ct0, ct1);
            }
        }
    } else {
        // ===================== epilogue: one thread per output row =====================
        const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;
        const int row_in_blk = half * 128 + quarter * 32 + lane;
        const bool want_topk = p.k > 0;
        const bool want_cnt = p.gold_cnt != nullptr;
        const float* r2_smem = reinterpret_cast<const float*>(smem_raw + (r2_slots - smem_u32(smem_raw)));
        uint32_t tcount = 0;
        for (PieceIter pi(p, n_rb); pi.valid();) {
            int rb, ct0, ct1;
            pi.piece(rb, ct0, ct1);
            const int row = rb * SIM_BM + row_in_blk;
            const bool rvalid = row < p.n1;
            const size_t orow = (p.split ? (size_t)pi.slot(rb) * p.slot_rows : (size_t)0) + (size_t)row;
            TopK<KMAX> tk;
            tk.init();
            const float r1 = (p.r1 && rvalid) ? p.r1[row] : 0.f;
            const float gold = (p.gold_val && rvalid) ? p.gold_val[row] : INFINITY;
            const int gcol = (p.diag && rvalid) ? p.gold_col[row] : -1;
            int cnt = 0;
            float dg = 0.f;
            float st_m = -INFINITY, st_Z = 0.f, st_W = 0.f;
            int st_arg = 0;
            for (int ct = ct0; ct < ct1; ++ct, ++tcount) {
                const uint32_t acc = tcount & 1, aph = (tcount >> 1) & 1;
                const uint32_t rs = tcount & 1, rph = (tcount >> 1) & 1;
                mbar_wait(t_full0 + 8 * acc, aph);
                if (CSLS) mbar_wait(r2_full0 + 8 * rs, rph);
                tc_fence_after();
                const bool ragged = (ct + 1) * SIM_BN > p.n2;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256 + half * 128;
                const float* r2t = r2_smem + rs * SIM_BN;
                uint32_t raw[2][32];
                tc_ld32_issue(taddr, raw[0]);
#pragma unroll
                for (int c = 0; c < SIM_BN / 32; ++c) {
                    tc_ld_wait();
                    if (c + 1 < SIM_BN / 32) tc_ld32_issue(taddr + (c + 1) * 32, raw[(c + 1) & 1]);   // overlaps the math below
                    const uint32_t(&rw)[32] = raw[c & 1];
                    const int col0 = ct * SIM_BN + c * 32;
                    if (gcol >= col0 && gcol < col0 + 32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j == gcol) dg = __uint_as_float(rw[j]);
                    }
                    float v[32];
                    if (STATS) {
                        // softmax statistics of x = alpha*S with masked rows / columns forced to p.masked
                        // (compute_alignment_quality, train.py:231-259): running max m, argmax, Z = sum e^(x-m),
                        // W = sum e^(x-m) (x-m).  Entropy = log Z - W/Z, max probability = 1/Z.
                        bool ok[32];
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            float4 cm = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (CSLS) cm = *reinterpret_cast<const float4*>(r2t + c * 32 + 4 * j4);
                            const float cmv[4] = {cm.x, cm.y, cm.z, cm.w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int j = 4 * j4 + q;
                                const bool masked = (r1 != 0.f) || (cmv[q] != 0.f);
                                v[j] = masked ? p.masked : p.alpha * __uint_as_float(rw[j]);
                                ok[j] = !ragged || (col0 + j < p.n2);
                            }
                        }
                        float cmax = -INFINITY;
                        int carg = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (ok[j] && v[j] > cmax) { cmax = v[j]; carg = j; }
                        if (cmax > st_m) {
                            const float sc = __expf(st_m - cmax);                   // 0 on the first chunk (st_m = -inf)
                            st_W = st_Z > 0.f ? (st_W + (st_m - cmax) * st_Z) * sc : 0.f;
                            st_Z *= sc;
                            st_m = cmax;
                            st_arg = col0 + carg;
                        }
                        // the 32 terms of the chunk are summed on their own first: adding thousands of tiny terms one by one to a
                        // Z ~ 1 dominated by the gold column would round each of them to ulp(1)
                        float cz = 0.f, cw = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float t = ok[j] ? v[j] - st_m : 0.f;
                            const float e = ok[j] ? __expf(t) : 0.f;
                            cz += e;
                            cw = fmaf(e, t, cw);
                        }
                        st_Z += cz;
                        st_W += cw;
                        continue;
                    }
                    if (CSLS) {
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 rr = *reinterpret_cast<const float4*>(r2t + c * 32 + 4 * j4);   // LDS.128 broadcast
                            v[4 * j4 + 0] = __fsub_rn(__fmaf_rn(p.alpha, __uint_as_float(rw[4 * j4 + 0]), -r1), rr.x);
                            v[4 * j4 + 1] = __fsub_rn(__fmaf_rn(p.alpha, __uint_as_float(rw[4 * j4 + 1]), -r1), rr.y);
                            v[4 * j4 + 2] = __fsub_rn(__fmaf_rn(p.alpha, __uint_as_float(rw[4 * j4 + 2]), -r1), rr.z);
                            v[4 * j4 + 3] = __fsub_rn(__fmaf_rn(p.alpha, __uint_as_float(rw[4 * j4 + 3]), -r1), rr.w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __fmaf_rn(p.alpha, __uint_as_float(rw[j]), -r1);
                    }
                    if (ragged) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j >= p.n2) v[j] = -INFINITY;
                    }
                    if (want_cnt) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) cnt += (v[j] > gold) ? 1 : 0;
                    }
                    if (want_topk) {
                        float m = v[0];
#pragma unroll
                        for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
                        if (m > tk.v[KMAX - 1]) {               // rare after the first few tiles
                            // compact slow path (one copy of the insertion code, not 32: keeps the I-cache footprint small)
                            unsigned cand = 0u;
#pragma unroll
                            for (int j = 0; j < 32; ++j) cand |= (v[j] > tk.v[KMAX - 1] ? 1u : 0u) << j;
                            while (cand) {
                                const int j = __ffs(cand) - 1;
                                cand &= cand - 1;
                                float x = v[0];
#pragma unroll
                                for (int jj = 1; jj < 32; ++jj) x = (j == jj) ? v[jj] : x;
                                if (x > tk.v[KMAX - 1]) tk.insert(x, col0 + j);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(t_empty0 + 8 * acc);
                    if (CSLS) mbar_arrive(r2_empty0 + 8 * rs);
                }
            }
            if (rvalid) {
                if (want_topk) {
#pragma unroll
                    for (int j = 0; j < KMAX; ++j)
                        if (j < p.k) { p.topv[orow * p.k + j] = tk.v[j]; p.topi[orow * p.k + j] = tk.i[j]; }
                }
                if (want_cnt) p.gold_cnt[orow] = cnt;
                if (p.diag) p.diag[orow] = dg;
                if (STATS) { p.st_m[orow] = st_m; p.st_Z[orow] = st_Z; p.st_W[orow] = st_W; p.st_arg[orow] = st_arg; }
            }
            pi.next(ct0, ct1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// Combines the per-piece partial results of a split launch: one thread per row, slots visited in ascending column
// order (strict '>' comparisons keep the lower column on exact ties, as the unsplit sweep does).
struct MergeParams {
    int n1, k, n_ct, tpc;
    const float* p_topv; const int* p_topi; const int* p_cnt; const float* p_diag;
    const float *p_m, *p_Z, *p_W; const int* p_arg;
    float* topv; int* topi; int* cnt; float* diag;
    float *st_m, *st_Z, *st_W; int* st_arg;
};

template <int KMAX>
__global__ void __launch_bounds__(256) sim_merge_kernel(MergeParams q) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= q.n1) return;
    const int rb = row / SIM_BM;
    const int first = (rb * q.n_ct) / q.tpc, last = ((rb + 1) * q.n_ct - 1) / q.tpc;
    const int np = last - first + 1;
    const size_t n1 = (size_t)q.n1;
    if (q.topv) {
        TopK<KMAX> tk;
        tk.init();
        for (int s = 0; s < np; ++s) {
            const float* pv = q.p_topv + ((size_t)s * n1 + row) * q.k;
            const int* pi = q.p_topi + ((size_t)s * n1 + row) * q.k;
            for (int j = 0; j < q.k; ++j) {
                const float x = pv[j];
                if (!(x > tk.v[KMAX - 1])) break;          // the slot's list is descending
                tk.insert(x, pi[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j < q.k) { q.topv[(size_t)row * q.k + j] = tk.v[j]; q.topi[(size_t)row * q.k + j] = tk.i[j]; }
    }
    if (q.cnt) {
        int c = 0;
        for (int s = 0; s < np; ++s) c += q.p_cnt[(size_t)s * n1 + row];
        q.cnt[row] = c;
    }
    if (q.diag) {
        float d = 0.f;
        for (int s = 0; s < np; ++s) d += q.p_diag[(size_t)s * n1 + row];   // exactly one slot holds the gold column, the others wrote 0
        q.diag[row] = d;
    }
    if (q.st_m) {
        float M = -INFINITY;
        int arg = 0;
        for (int s = 0; s < np; ++s) {
            const float m = q.p_m[(size_t)s * n1 + row];
            if (m > M) { M = m; arg = q.p_arg[(size_t)s * n1 + row]; }
        }
        float Z = 0.f, W = 0.f;
        for (int s = 0; s < np; ++s) {
            const size_t o = (size_t)s * n1 + row;
            const float m = q.p_m[o], z = q.p_Z[o];
            if (z > 0.f) {
                const float sc = __expf(m - M);
                Z = fmaf(z, sc, Z);
                W += (q.p_W[o] + (m - M) * z) * sc;
            }
        }
        q.st_m[row] = M; q.st_Z[row] = Z; q.st_W[row] = W; q.st_arg[row] = arg;
    }
}

// fp32 rows (optionally gathered by idx, optionally L2-normalised) -> bf16 rows padded to dpad columns
__global__ void to_bf16_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t n_rows, int d, int dpad,
                                    int normalize, uint16_t* __restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rows) return;
    const float* s = src + (idx ? idx[r] : r) * (int64_t)d;
    float scale = 1.f;
    if (normalize) {
        float ss = 0.f;
        for (int c = lane; c < d; c += 32) ss = fmaf(s[c], s[c], ss);
        ss = warp_sum(ss);
        scale = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    }
    for (int c = lane; c < dpad; c += 32) {
        const float v = c < d ? s[c] * scale : 0.f;
        dst[r * dpad + c] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    }
}

__global__ void csls_gold_kernel(float alpha, const float* __restrict__ diag, const float* __restrict__ r1, const float* __restrict__ r2,
                                 const int* __restrict__ gold_col, int64_t n, float* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = __fmaf_rn(alpha, diag[i], r1 ? -r1[i] : -0.f);
    if (r2) v = __fsub_rn(v, r2[gold_col[i]]);
    out[i] = v;
}

// fp32 rows (optionally gathered / L2-normalised, norm accumulated in fp64) -> the two TF32 terms of the 3xTF32 product:
// hi = RN-to-TF32(v), lo = RN-to-TF32(v - hi); both have zero low mantissa bits, so the tensor core truncates nothing and
// |v - hi - lo| <= 2^-22 |v|.  Columns d..dpad-1 are zero.
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r & 0xFFFFE000u);
}
__global__ void to_tf32_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t n_rows, int d, int dpad,
                                    int normalize, float* __restrict__ hi, float* __restrict__ lo) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rows) return;
    const float* s = src + (idx ? idx[r] : r) * (int64_t)d;
    float scale = 1.f;
    if (normalize) {
        double ss = 0.0;
        for (int c = lane; c < d; c += 32) ss += (double)s[c] * (double)s[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(FULL, ss, o);
        scale = (float)(1.0 / fmax(sqrt(ss), 1e-12));
    }
    for (int c = lane; c < dpad; c += 32) {
        const float v = c < d ? s[c] * scale : 0.f;
        const float h = tf32_rna(v);
        hi[r * dpad + c] = h;
        lo[r * dpad + c] = tf32_rna(v - h);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows x dpad table (bf16: 64-wide K slices, fp32: 32-wide), box = one K slice x box_rows, 128-byte swizzle
static int make_map(CUtensorMap* map, const void* ptr, int64_t rows, int dpad, int box_rows, bool f32) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return JMAC_EINVAL; }
    cuuint64_t dims[2] = {(cuuint64_t)dpad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)dpad * (f32 ? 4 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)(f32 ? SIM_T_BK : SIM_BK), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return JMAC_EINVAL; }
    return 0;
}

struct SimDecomp { int n_rb, n_ct, tpc, n_units, split, P, grid; };

// b_bytes: size of the column-side operand.  Unsplit, the CTAs sweep the column tiles in lockstep and share every B tile through
// L2; stream-K runs start at different columns, so B is re-read from HBM unless it is L2-resident (measured at 131072 x 1M bf16:
// 70-72 ms split against 64.5 ms unsplit, although the split removes a 13 % wave-quantisation tail).  Split therefore only when
// there are fewer row blocks than SMs, or when B fits in L2 and the last wave would be badly filled.
static SimDecomp sim_decomp(int64_t n1, int64_t n2, bool allow_split, size_t b_bytes = 0) {
    SimDecomp dc;
    const int sms = sm_count();
    dc.n_rb = (int)((n1 + SIM_BM - 1) / SIM_BM);
    dc.n_ct = (int)((n2 + SIM_BN - 1) / SIM_BN);
    const long long T = (long long)dc.n_rb * dc.n_ct;
    const bool few_blocks = dc.n_rb < sms;
    const bool cheap_b = b_bytes <= (size_t(64) << 20) && dc.n_rb < 4 * sms;
    if (allow_split && (few_blocks || cheap_b) && dc.n_ct > 1) {
        const int G = (int)(T < sms ? T : sms);
        dc.tpc = (int)((T + G - 1) / G);
        dc.n_units = (int)((T + dc.tpc - 1) / dc.tpc);
        dc.split = 1;
        dc.P = (dc.n_ct + dc.tpc - 1) / dc.tpc + 1;
        dc.grid = dc.n_units;
    } else {
        dc.tpc = dc.n_ct; dc.n_units = dc.n_rb; dc.split = 0; dc.P = 1;
        dc.grid = dc.n_rb < sms ? dc.n_rb : sms;
    }
    return dc;
}

static size_t sim_ws_bytes(const SimDecomp& dc, int64_t n1, int k) {
    return dc.split ? (size_t)dc.P * (size_t)n1 * ((size_t)(k > 0 ? k : 1) * 8 + 24) : 0;
}

// One launch of the similarity kernel (+ the merge of a split launch).  PREC 0: A/B are bf16 tables, A_lo/B_lo unused.
static int sim_launch(int prec, const void* A, const void* A_lo, const void* B, const void* B_lo, int64_t n1, int64_t n2, int dpad, SimParams p,
                      bool stats, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    const bool f32 = prec == 1;
    CUtensorMap map_a, map_b, map_a_lo, map_b_lo;
    int rc = make_map(&map_a, A, n1, dpad, SIM_BM, f32);
    if (rc) return rc;
    rc = make_map(&map_b, B, n2, dpad, SIM_BN, f32);
    if (rc) return rc;
    map_a_lo = map_a; map_b_lo = map_b;
    if (f32) {
        rc = make_map(&map_a_lo, A_lo, n1, dpad, SIM_BM, true);
        if (rc) return rc;
        rc = make_map(&map_b_lo, B_lo, n2, dpad, SIM_BN, true);
        if (rc) return rc;
    }
    const size_t b_bytes = (size_t)n2 * (size_t)dpad * (f32 ? 8 : 2);
    SimDecomp dc = sim_decomp(n1, n2, workspace != nullptr, b_bytes);
    if (dc.split && sim_ws_bytes(dc, n1, p.k) > workspace_bytes) dc = sim_decomp(n1, n2, false);     // workspace too small: unsplit launch
    p.n_ct = dc.n_ct; p.tpc = dc.tpc; p.n_units = dc.n_units; p.split = dc.split; p.slot_rows = (size_t)n1;
    p.nks = dpad / (f32 ? SIM_T_BK : SIM_BK);
    MergeParams q;
    if (dc.split) {
        // carve the partial arrays out of the workspace; the kernel's output pointers are redirected to them
        const size_t rows = (size_t)dc.P * (size_t)n1, kk = (size_t)(p.k > 0 ? p.k : 1);
        char* w = reinterpret_cast<char*>(workspace);
        float* w_topv = reinterpret_cast<float*>(w); w += rows * kk * 4;
        int* w_topi = reinterpret_cast<int*>(w); w += rows * kk * 4;
        int* w_cnt = reinterpret_cast<int*>(w); w += rows * 4;
        float* w_diag = reinterpret_cast<float*>(w); w += rows * 4;
        float* w_m = reinterpret_cast<float*>(w); w += rows * 4;
        float* w_Z = reinterpret_cast<float*>(w); w += rows * 4;
        float* w_W = reinterpret_cast<float*>(w); w += rows * 4;
        int* w_arg = reinterpret_cast<int*>(w);
        q.n1 = (int)n1; q.k = p.k; q.n_ct = dc.n_ct; q.tpc = dc.tpc;
        q.p_topv = w_topv; q.p_topi = w_topi; q.p_cnt = w_cnt; q.p_diag = w_diag; q.p_m = w_m; q.p_Z = w_Z; q.p_W = w_W; q.p_arg = w_arg;
        q.topv = p.k > 0 ? p.topv : nullptr; q.topi = p.topi; q.cnt = p.gold_cnt; q.diag = p.diag;
        q.st_m = stats ? p.st_m : nullptr; q.st_Z = p.st_Z; q.st_W = p.st_W; q.st_arg = p.st_arg;
        if (p.k > 0) { p.topv = w_topv; p.topi = w_topi; }
        if (p.gold_cnt) p.gold_cnt = w_cnt;
        if (p.diag) p.diag = w_diag;
        if (stats) { p.st_m = w_m; p.st_Z = w_Z; p.st_W = w_W; p.st_arg = w_arg; }
    }
    const int smem = f32 ? SIM_SMEM_T : SIM_SMEM;
#define JMAC_SIM_LAUNCH(KM, CS, STT, PR)                                                                                         \
    do {                                                                                                                         \
        JMAC_CUDA(cudaFuncSetAttribute(sim_topk_kernel<KM, CS, STT, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));     \
        sim_topk_kernel<KM, CS, STT, PR><<<dc.grid, SIM_THREADS, smem, st>>>(map_a, map_b, map_a_lo, map_b_lo, p);               \
    } while (0)
#define JMAC_SIM_PREC(KM, CS, STT) do { if (f32) JMAC_SIM_LAUNCH(KM, CS, STT, 1); else JMAC_SIM_LAUNCH(KM, CS, STT, 0); } while (0)
    const bool cs = p.r2 != nullptr;
    if (stats)          { if (cs) JMAC_SIM_PREC(1, true, true);   else JMAC_SIM_PREC(1, false, true); }
    else if (p.k <= 1)  { if (cs) JMAC_SIM_PREC(1, true, false);  else JMAC_SIM_PREC(1, false, false); }
    else if (p.k <= 16) { if (cs) JMAC_SIM_PREC(16, true, false); else JMAC_SIM_PREC(16, false, false); }
    else                { if (cs) JMAC_SIM_PREC(32, true, false); else JMAC_SIM_PREC(32, false, false); }
#undef JMAC_SIM_PREC
#undef JMAC_SIM_LAUNCH
    JMAC_LAUNCH_CHECK();
    if (dc.split) {
        const unsigned blocks = (unsigned)((n1 + 255) / 256);
        if (p.k <= 1) sim_merge_kernel<1><<<blocks, 256, 0, st>>>(q);
        else if (p.k <= 16) sim_merge_kernel<16><<<blocks, 256, 0, st>>>(q);
        else sim_merge_kernel<32><<<blocks, 256, 0, st>>>(q);
        JMAC_LAUNCH_CHECK();
    }
    return 0;
}

static int sim_topk_common(int prec, const void* A, const void* A_lo, const void* B, const void* B_lo, int64_t n1, int64_t n2, int dpad, int k,
                           float alpha, const float* r1, const float* r2, const int32_t* gold_col, const float* gold_val, int32_t* gold_cnt,
                           float* diag, float* topv, int32_t* topi, void* workspace, size_t workspace_bytes, void* stream) {
    JMAC_ARG(A && B && n1 > 0 && n2 > 0 && n1 < (int64_t(1) << 31) && n2 < (int64_t(1) << 31));
    if (prec == 0) JMAC_ARG(dpad >= 64 && dpad <= SIM_MAX_KS * SIM_BK && (dpad % 64) == 0);
    else JMAC_ARG(A_lo && B_lo && dpad >= 32 && dpad <= 4096 && (dpad % 32) == 0);
    JMAC_ARG(k >= 0 && k <= 32 && (k == 0 || (topv && topi)));
    JMAC_ARG((gold_cnt == nullptr) || gold_val != nullptr);
    JMAC_ARG((diag == nullptr) || gold_col != nullptr);
    JMAC_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)A_lo % 16) == 0 && ((uintptr_t)B_lo % 16) == 0);
    JMAC_ARG(r2 == nullptr || ((uintptr_t)r2 % 16) == 0);
    JMAC_ARG(((uintptr_t)workspace % 16) == 0);
    SimParams p;
    p.n1 = (int)n1; p.n2 = (int)n2; p.nks = 0; p.k = k; p.alpha = alpha; p.r1 = r1; p.r2 = r2;
    p.gold_col = gold_col; p.gold_val = gold_val; p.gold_cnt = gold_cnt; p.diag = diag; p.topv = topv; p.topi = topi;
    p.masked = 0.f; p.st_m = p.st_Z = p.st_W = nullptr; p.st_arg = nullptr;
    return sim_launch(prec, A, A_lo, B, B_lo, n1, n2, dpad, p, false, workspace, workspace_bytes, (cudaStream_t)stream);
}

static int sim_stats_common(int prec, const void* A, const void* A_lo, const void* B, const void* B_lo, int64_t n1, int64_t n2, int dpad,
                            float alpha, const float* row_mask, const float* col_mask, float masked_value, float* out_max, int32_t* out_arg,
                            float* out_Z, float* out_W, void* workspace, size_t workspace_bytes, void* stream) {
    JMAC_ARG(A && B && out_max && out_arg && out_Z && out_W);
    JMAC_ARG(n1 > 0 && n2 > 0 && n1 < (int64_t(1) << 31) && n2 < (int64_t(1) << 31));
    if (prec == 0) JMAC_ARG(dpad >= 64 && dpad <= SIM_MAX_KS * SIM_BK && (dpad % 64) == 0);
    else JMAC_ARG(A_lo && B_lo && dpad >= 32 && dpad <= 4096 && (dpad % 32) == 0);
    JMAC_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)A_lo % 16) == 0 && ((uintptr_t)B_lo % 16) == 0);
    JMAC_ARG(col_mask == nullptr || ((uintptr_t)col_mask % 16) == 0);
    JMAC_ARG(((uintptr_t)workspace % 16) == 0);
    SimParams p;
    p.n1 = (int)n1; p.n2 = (int)n2; p.nks = 0; p.k = 0; p.alpha = alpha; p.r1 = row_mask; p.r2 = col_mask;
    p.gold_col = nullptr; p.gold_val = nullptr; p.gold_cnt = nullptr; p.diag = nullptr; p.topv = nullptr; p.topi = nullptr;
    p.masked = masked_value; p.st_m = out_max; p.st_Z = out_Z; p.st_W = out_W; p.st_arg = out_arg;
    return sim_launch(prec, A, A_lo, B, B_lo, n1, n2, dpad, p, true, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // namespace jmac

using namespace jmac;

extern "C" int jmac_sm_count(void) { return sm_count(); }

extern "C" size_t jmac_sim_workspace_bytes(int64_t n1, int64_t n2, int k) {
    if (n1 <= 0 || n2 <= 0) return 0;
    return sim_ws_bytes(sim_decomp(n1, n2, true), n1, k);
}

extern "C" int jmac_sim_grid(int64_t n1, int64_t n2, int split) {
    if (n1 <= 0 || n2 <= 0) return 0;
    return sim_decomp(n1, n2, split != 0).grid;
}

extern "C" int jmac_to_bf16_rows(const float* src, int d, const int64_t* idx, int64_t n_rows, int normalize, uint16_t* dst, int dpad,
                                 void* stream) {
    JMAC_ARG(src && dst && d > 0 && dpad >= d && (dpad % 64) == 0 && n_rows >= 0);
    if (n_rows == 0) return 0;
    const int64_t blocks = (n_rows * 32 + 255) / 256;
    to_bf16_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, idx, n_rows, d, dpad, normalize, dst);
    JMAC_LAUNCH_CHECK();
    return 0;
}

extern "C" int jmac_to_tf32_rows(const float* src, int d, const int64_t* idx, int64_t n_rows, int normalize, float* hi, float* lo, int dpad,
                                 void* stream) {
    JMAC_ARG(src && hi && lo && d > 0 && dpad >= d && (dpad % 32) == 0 && n_rows >= 0);
    if (n_rows == 0) return 0;
    const int64_t blocks = (n_rows * 32 + 255) / 256;
    to_tf32_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, idx, n_rows, d, dpad, normalize, hi, lo);
    JMAC_LAUNCH_CHECK();
    return 0;
}

extern "C" int jmac_csls_gold(float alpha, const float* diag, const float* r1, const float* r2, const int32_t* gold_col, int64_t n,
                              float* out, void* stream) {
    JMAC_ARG(diag && out && n >= 0 && (r2 == nullptr || gold_col != nullptr));
    if (n == 0) return 0;
    csls_gold_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(alpha, diag, r1, r2, gold_col, n, out);
    JMAC_LAUNCH_CHECK();
    return 0;
}

extern "C" int jmac_sim_topk(const uint16_t* A, const uint16_t* B, int64_t n1, int64_t n2, int dpad, int k, float alpha, const float* r1,
                             const float* r2, const int32_t* gold_col, const float* gold_val, int32_t* gold_cnt, float* diag, float* topv,
                             int32_t* topi, void* workspace, size_t workspace_bytes, void* stream) {
    return sim_topk_common(0, A, nullptr, B, nullptr, n1, n2, dpad, k, alpha, r1, r2, gold_col, gold_val, gold_cnt, diag, topv, topi, workspace,
                           workspace_bytes, stream);
}

extern "C" int jmac_sim_topk_f32(const float* A_hi, const float* A_lo, const float* B_hi, const float* B_lo, int64_t n1, int64_t n2, int dpad,
                                 int k, float alpha, const float* r1, const float* r2, const int32_t* gold_col, const float* gold_val,
                                 int32_t* gold_cnt, float* diag, float* topv, int32_t* topi, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    return sim_topk_common(1, A_hi, A_lo, B_hi, B_lo, n1, n2, dpad, k, alpha, r1, r2, gold_col, gold_val, gold_cnt, diag, topv, topi, workspace,
                           workspace_bytes, stream);
}

extern "C" int jmac_sim_softmax_stats(const uint16_t* A, const uint16_t* B, int64_t n1, int64_t n2, int dpad, float alpha,
                                      const float* row_mask, const float* col_mask, float masked_value, float* out_max, int32_t* out_arg,
                                      float* out_Z, float* out_W, void* workspace, size_t workspace_bytes, void* stream) {
    return sim_stats_common(0, A, nullptr, B, nullptr, n1, n2, dpad, alpha, row_mask, col_mask, masked_value, out_max, out_arg, out_Z, out_W,
                            workspace, workspace_bytes, stream);
}

extern "C" int jmac_sim_softmax_stats_f32(const float* A_hi, const float* A_lo, const float* B_hi, const float* B_lo, int64_t n1, int64_t n2,
                                          int dpad, float alpha, const float* row_mask, const float* col_mask, float masked_value,
                                          float* out_max, int32_t* out_arg, float* out_Z, float* out_W, void* workspace,
                                          size_t workspace_bytes, void* stream) {
    return sim_stats_common(1, A_hi, A_lo, B_hi, B_lo, n1, n2, dpad, alpha, row_mask, col_mask, masked_value, out_max, out_arg, out_Z, out_W,
                            workspace, workspace_bytes, stream);
}
