// Shared device/host helpers for libjmac_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/jmac_b200.h"

namespace jmac {

void set_error(const char* fmt, ...);
int sm_count();
void count_launch();
// Device-visible status words in mapped, pinned host memory: kernels that consume caller-provided index lists store 1 into
// word `which` when an index is out of range (and use row 0 instead, so nothing is read or written out of bounds);
// jmac_status_poll() reads and clears them without synchronising anything.  nullptr if the allocation failed.
int* status_words();
enum { JMAC_ST_TRIPLE = 0, JMAC_ST_PAIR = 1, JMAC_ST_LINKPRED = 2, JMAC_ST_SAMPLER = 3, JMAC_ST_WORDS = 8 };

#define JMAC_CUDA(expr)                                                                    \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            jmac::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (int)_e;                                                                \
        }                                                                                  \
    } while (0)

#define JMAC_ARG(cond)                                                                     \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            jmac::set_error("%s:%d bad argument: %s", __FILE__, __LINE__, #cond);          \
            return JMAC_EINVAL;                                                            \
        }                                                                                  \
    } while (0)

// every kernel launch of the library goes through this: error check + launch counter
#define JMAC_LAUNCH_CHECK()              \
    do {                                 \
        jmac::count_launch();            \
        JMAC_CUDA(cudaGetLastError());   \
    } while (0)

// A work item of a segmented pass: segment id, [beg,end) in that ordering, and the partial slot
// (-1 when the segment is handled by exactly one item and can be finalised in place).
struct __align__(16) Item { int seg, beg, end, slot; };
// A split segment: segment id, first partial slot, number of slots.
struct __align__(16) Split { int seg, first, count, pad; };

struct ItemList {
    Item* items = nullptr;
    Split* splits = nullptr;
    int n_items = 0, n_splits = 0, n_slots = 0;
};

}  // namespace jmac

struct jmac_plan {
    int64_t E = 0, N = 0, N_src = 0, n_rel = 0;   // N = destination (aggregation) nodes, N_src = source table rows
    int chunk = 256;
    int max_deg = 0;
    int device = 0;                               // the device the plan's arrays live on (restored around destroy)
    // by destination (CSR)
    int *rowptr = nullptr, *csr_src = nullptr, *csr_type = nullptr, *csr_eid = nullptr, *csr_dst = nullptr;
    float* sqrt_deg = nullptr;
    // by source
    int *colptr = nullptr, *csc_pos = nullptr, *csc_dst = nullptr;
    // by relation
    int *typptr = nullptr, *typ_pos = nullptr, *typ_dst = nullptr;
    jmac::ItemList dst_items, src_items, typ_items;
    // scratch for split segments (partials), grown lazily (stream-ordered allocation on the stream of the call that grows it)
    float* scratch = nullptr;
    size_t scratch_bytes = 0;
    // last use of the plan's arrays per stream (up to 4 streams): jmac_plan_destroy orders the frees behind these events instead
    // of synchronising the device
    cudaEvent_t use_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t use_st[4] = {nullptr, nullptr, nullptr, nullptr};
    int n_use = 0;
    bool many_streams = false;
};

namespace jmac {

int plan_scratch(jmac_plan* p, size_t bytes, float** out, cudaStream_t st);
// to be called after the last launch of an entry point that read the plan's arrays on `st`
void plan_mark_use(jmac_plan* p, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// 128-bit streaming load of a gathered node row (read once per SM: keep it out of L1).
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// 128-bit load of a hot (relation / per-node) row through the read-only path, L1-allocating.
__device__ __forceinline__ float4 ld_hot(const float4* p) { return __ldg(p); }

__device__ __forceinline__ void st_stream(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// vector reduction (no return) -- sm_90+: one 16-byte RED instead of four.
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// an index taken from the caller: flagged and replaced by 0 when it is outside [0, n)
__device__ __forceinline__ int64_t safe_index(int64_t v, int64_t n, int* status_word) {
    if ((uint64_t)v >= (uint64_t)n) {
        if (status_word) *reinterpret_cast<volatile int*>(status_word) = 1;
        return 0;
    }
    return v;
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float s, float4 a) { return make_float4(s * a.x, s * a.y, s * a.z, s * a.w); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ void fma4(float4& acc, float s, float4 a) {
    acc.x = fmaf(s, a.x, acc.x); acc.y = fmaf(s, a.y, acc.y); acc.z = fmaf(s, a.z, acc.z); acc.w = fmaf(s, a.w, acc.w);
}

// A row of d floats held by one warp: lane l owns float4 #(v*32 + l), v < VPL (predicated on d4).
template <int VPL>
struct Row {
    float4 v[VPL];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[i] = f4_zero();
    }
};

template <int VPL, bool STREAM>
__device__ __forceinline__ void load_row(Row<VPL>& r, const float* base, int64_t row, int d4, int lane) {
    const float4* p = reinterpret_cast<const float4*>(base) + row * d4;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        int c = i * 32 + lane;
        if (c < d4) r.v[i] = STREAM ? ld_stream(p + c) : ld_hot(p + c);
        else r.v[i] = f4_zero();
    }
}
template <int VPL>
__device__ __forceinline__ void store_row(const Row<VPL>& r, float* base, int64_t row, int d4, int lane) {
    float4* p = reinterpret_cast<float4*>(base) + row * d4;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        int c = i * 32 + lane;
        if (c < d4) p[c] = r.v[i];
    }
}
// ---- per-warp shared-memory rings fed by cp.async.bulk (TMA bulk copies): latency is hidden by bytes in
// flight in shared memory instead of registers/occupancy.  One CTA of RING_WARPS warps per SM.
constexpr int RING_WARPS = 16;
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_row(uint32_t dst, const float* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

// Pull the cache lines that hold entry `pos` of up to three 4-byte-per-entry streams (index lists, logits) from HBM towards L2:
// nothing prefetches sequential streams in hardware, and a warp consumes ~20 entries of each per batch, so without this every
// batch load pays the full DRAM latency.  Lanes 0..2 issue one prefetch each; `pos` >= n is ignored.
__device__ __forceinline__ void prefetch_idx_l2(const void* a, const void* b, const void* c, long long pos, long long n, int lane) {
    const void* t = lane == 0 ? a : (lane == 1 ? b : c);
    if (lane < 3 && t != nullptr && pos < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(t) + pos * 4));
}

template <int VPL>
__device__ __forceinline__ void lds_row(Row<VPL>& r, const float* base, int d4, int lane) {
    const float4* p = reinterpret_cast<const float4*>(base);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int c = i * 32 + lane;
        r.v[i] = c < d4 ? p[c] : f4_zero();
    }
}


// grid for a persistent warp-per-item kernel: enough blocks to fill every SM at the kernel's
// occupancy, never more than the work needs.
template <typename K>
inline int persistent_blocks(K kernel, int threads, int n_warp_items) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int wpb = threads / 32;
    long long want = ((long long)n_warp_items + wpb - 1) / wpb;
    long long cap = (long long)sm_count() * per_sm;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

#endif  // __CUDACC__

// host-side dispatch on the number of float4 per lane
#define JMAC_DISPATCH_VPL(d, ...)                                   \
    do {                                                            \
        int _vpl = ((d) / 4 + 31) / 32;                             \
        if (_vpl == 1) { constexpr int VPL = 1; __VA_ARGS__; }      \
        else if (_vpl == 2) { constexpr int VPL = 2; __VA_ARGS__; } \
        else if (_vpl == 3) { constexpr int VPL = 3; __VA_ARGS__; } \
        else { constexpr int VPL = 4; __VA_ARGS__; }                \
    } while (0)

inline bool dim_ok(int d) { return d >= 4 && d <= 512 && (d % 4) == 0; }

// Per-edge record shared by the forward (writes logit + mask) and backward passes, in CSR order:
//   [0, MB*E)              sign mask of z_e: lane l of the warp owns MB/32 bytes   (forward)
//   [MB*E, MB*E + 8E)      float2 {w_e = alpha_e sqrt(deg_i), ds_e}                 (backward dst pass)
//   then                   float logit s_e                                          (forward)
// The mask region comes first: its rows are fetched by cp.async.bulk, which needs 16-byte aligned
// sources for every E (MB is 32 or 64, the buffer itself is at least 32-byte aligned).
template <int VPL> struct MaskT { typedef uint8_t type; };
template <> struct MaskT<3> { typedef uint16_t type; };
template <> struct MaskT<4> { typedef uint16_t type; };
inline size_t edge_mask_bytes(int d) { return (((d / 4 + 31) / 32) <= 2) ? 32 : 64; }
inline void* erec_mask(void* erec, int64_t) { return erec; }
inline float2* erec_ew(void* erec, int64_t E, int d) {
    return reinterpret_cast<float2*>(reinterpret_cast<char*>(erec) + edge_mask_bytes(d) * (size_t)E);
}
inline float* erec_logit(void* erec, int64_t E, int d) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(erec) + (sizeof(float2) + edge_mask_bytes(d)) * (size_t)E);
}

}  // namespace jmac
