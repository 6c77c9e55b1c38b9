// Graph plan: COO (int64, as the reference passes it) -> three device-resident edge orderings.
//   by destination (CSR, secondary key = relation type)  -> forward + backward dst pass
//   by source                                            -> backward src pass (atomics-free dX, dQ)
//   by relation type                                     -> backward relation pass
// plus the work-item lists that split long segments (hubs) into warp-sized chunks.
// Replaces RelationAwareLayer.compute_norm (src/jmac_model.py:99-109) and the index work of
// MessagePassing.propagate (modules/helper/message_passing.py:69-81).
#include <stdarg.h>

#include <cub/cub.cuh>

#include <stdlib.h>

#include <stdlib.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace jmac {

static thread_local char g_err[768] = "";
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int* status_words() {
    static int* host = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* q = nullptr;
        if (cudaHostAlloc(&q, sizeof(int) * JMAC_ST_WORDS, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
            host = static_cast<int*>(q);
            for (int i = 0; i < JMAC_ST_WORDS; ++i) host[i] = 0;
        } else {
            cudaGetLastError();
        }
    }
    return host;      // unified addressing: the same pointer is valid on every device
}

// SMs the persistent kernels may occupy.  jmac_set_sm_reserve(r) withholds r SMs: a communication kernel running next to a
// persistent one-CTA-per-SM compute kernel (which fills every SM's shared memory) can then always be scheduled (jmac_b200.dist).
static int g_sm_reserve = 0;
int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    const int r = __atomic_load_n(&g_sm_reserve, __ATOMIC_RELAXED);
    return n - r > 8 ? n - r : n;
}

// Stream-ordered allocation for the plan and its build temporaries, from a block cache owned by the library.
// cudaMalloc / cudaFree serialise the device and cost milliseconds per 100 MB block, which made rebuilding a 40M-edge plan (a
// triple-transfer round, train.py:471-472, or the host-fed bench step) slower than the sorts themselves.  The driver's own pool
// (cudaMallocAsync) removes that in the common case but trims and re-grows itself at synchronisation points: intermittent 0.2-0.8 s
// host stalls in the two-deep host-fed pipeline (bench.py e2e: 127 / 144 / 191 / 289 ms per step, run to run).  So freed blocks
// are kept here: a free records an event on the freeing stream, a later allocation of a similar size on ANY stream takes the
// block and makes its stream wait for that event on the device -- no host wait, no driver allocation after the first build of a
// given graph size.  Bounded: at most 32 GiB (JMAC_POOL_KEEP_GIB) of free blocks stay cached per device, beyond that the largest
// ones go back to the driver.  While a stream is being captured into a CUDA graph the driver's stream-ordered allocator is used
// (its nodes can be captured; events recorded outside the capture cannot be waited for inside it).
namespace {
struct Block { void* ptr; size_t bytes; cudaEvent_t ev; bool pool; };
struct DevCache {
    std::mutex mu;
    std::vector<Block> free_blocks;
    std::unordered_map<void*, Block> live;
    size_t cached = 0;
};
DevCache* dev_cache() {
    static DevCache caches[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    return &caches[dev];
}
size_t cache_cap() {
    static const size_t cap = [] {
        const char* env = getenv("JMAC_POOL_KEEP_GIB");
        const long gib = env ? atol(env) : 32;
        return (size_t)(gib > 0 ? gib : 1) << 30;
    }();
    return cap;
}
bool capturing(cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
    return cs != cudaStreamCaptureStatusNone;
}
void drop_block(const Block& b) {          // back to the driver (host-synchronous: only when the cache is over its bound or out of memory)
    cudaEventSynchronize(b.ev);
    cudaFree(b.ptr);
    cudaEventDestroy(b.ev);
}
}  // namespace

static cudaError_t dmalloc_bytes(void** p, size_t bytes, cudaStream_t st) {
    const size_t need = ((bytes > 0 ? bytes : 1) + 511) & ~(size_t)511;
    DevCache* dc = dev_cache();
    if (capturing(st)) {
        cudaError_t e = cudaMallocAsync(p, need, st);
        if (e == cudaSuccess) {
            std::lock_guard<std::mutex> g(dc->mu);
            dc->live[*p] = Block{*p, need, nullptr, true};
        }
        return e;
    }
    std::lock_guard<std::mutex> g(dc->mu);
    int best = -1;
    for (int i = 0; i < (int)dc->free_blocks.size(); ++i) {
        const size_t b = dc->free_blocks[i].bytes;
        if (b >= need && b <= need + need / 4 + (1u << 20) && (best < 0 || b < dc->free_blocks[best].bytes)) best = i;
    }
    if (best >= 0) {
        Block blk = dc->free_blocks[best];
        dc->free_blocks.erase(dc->free_blocks.begin() + best);
        dc->cached -= blk.bytes;
        cudaError_t e = cudaStreamWaitEvent(st, blk.ev, 0);      // the block's previous user (any stream) finishes first -- on the device
        if (e != cudaSuccess) { drop_block(blk); return e; }
        dc->live[blk.ptr] = blk;
        *p = blk.ptr;
        return cudaSuccess;
    }
    Block blk{nullptr, need, nullptr, false};
    cudaError_t e = cudaMalloc(&blk.ptr, need);
    if (e != cudaSuccess) {                                       // out of memory: give the cached blocks back and retry once
        cudaGetLastError();
        for (const Block& b : dc->free_blocks) drop_block(b);
        dc->free_blocks.clear();
        dc->cached = 0;
        e = cudaMalloc(&blk.ptr, need);
        if (e != cudaSuccess) return e;
    }
    e = cudaEventCreateWithFlags(&blk.ev, cudaEventDisableTiming);
    if (e != cudaSuccess) { cudaFree(blk.ptr); return e; }
    dc->live[blk.ptr] = blk;
    *p = blk.ptr;
    return cudaSuccess;
}
template <typename T>
static cudaError_t dmalloc(T** p, size_t bytes, cudaStream_t st) {
    return dmalloc_bytes(reinterpret_cast<void**>(p), bytes, st);
}
static void dfree(void* p, cudaStream_t st) {
    if (!p) return;
    DevCache* dc = dev_cache();
    std::lock_guard<std::mutex> g(dc->mu);
    auto it = dc->live.find(p);
    if (it == dc->live.end()) return;                             // not ours (cannot happen): leave it alone
    Block blk = it->second;
    dc->live.erase(it);
    if (blk.pool) { cudaFreeAsync(p, st); return; }
    if (capturing(st) || cudaEventRecord(blk.ev, st) != cudaSuccess) {   // cannot be ordered by an event: hand it back synchronously
        cudaGetLastError();
        cudaStreamSynchronize(st);
        cudaFree(blk.ptr);
        cudaEventDestroy(blk.ev);
        return;
    }
    dc->free_blocks.push_back(blk);
    dc->cached += blk.bytes;
    while (dc->cached > cache_cap() && !dc->free_blocks.empty()) {
        int big = 0;
        for (int i = 1; i < (int)dc->free_blocks.size(); ++i)
            if (dc->free_blocks[i].bytes > dc->free_blocks[big].bytes) big = i;
        Block d = dc->free_blocks[big];
        dc->free_blocks.erase(dc->free_blocks.begin() + big);
        dc->cached -= d.bytes;
        drop_block(d);
    }
}
// Build temporaries of one scope: returned to the pool (stream-ordered) when the scope ends, including on the
// early-return error paths of JMAC_CUDA / JMAC_LAUNCH_CHECK.
struct Temps {
    cudaStream_t st;
    void* ptrs[12];
    int n = 0;
    explicit Temps(cudaStream_t s) : st(s) {}
    Temps(const Temps&) = delete;
    Temps& operator=(const Temps&) = delete;
    ~Temps() { for (int i = 0; i < n; ++i) dfree(ptrs[i], st); }
    template <typename T>
    cudaError_t alloc(T** p, size_t bytes) {
        if (n >= 12) return cudaErrorMemoryAllocation;
        cudaError_t e = dmalloc(p, bytes, st);
        if (e == cudaSuccess) ptrs[n++] = *p;
        return e;
    }
};

int plan_scratch(jmac_plan* p, size_t bytes, float** out, cudaStream_t st) {
    if (bytes > p->scratch_bytes) {
        // stream-ordered: earlier kernels of this stream that still read the old block finish before it is recycled
        if (p->scratch) dfree(p->scratch, st);
        p->scratch = nullptr;
        p->scratch_bytes = 0;
        JMAC_CUDA(dmalloc(&p->scratch, bytes, st));
        p->scratch_bytes = bytes;
    }
    *out = p->scratch;
    return 0;
}

void plan_mark_use(jmac_plan* p, cudaStream_t st) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return; }
    if (cs != cudaStreamCaptureStatusNone) return;      // captured launches: the owner of the graph keeps the plan alive (jmac_b200.graphs)
    int i = 0;
    while (i < p->n_use && p->use_st[i] != st) ++i;
    if (i == p->n_use) {
        if (p->n_use == 4) { p->many_streams = true; return; }
        if (cudaEventCreateWithFlags(&p->use_ev[i], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); p->many_streams = true; return; }
        p->use_st[i] = st;
        ++p->n_use;
    }
    cudaEventRecord(p->use_ev[i], st);
}

// one non-blocking stream per device on which destroyed plans are freed, behind their last-use events
static cudaStream_t free_stream(int dev) {
    static cudaStream_t fs[64] = {};
    if (dev < 0 || dev >= 64) return nullptr;
    if (!fs[dev] && cudaStreamCreateWithFlags(&fs[dev], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); fs[dev] = nullptr; }
    return fs[dev];
}

static int bits_for(int64_t n) {  // bits needed to represent values in [0, n)
    int b = 1;
    while ((int64_t(1) << b) < n) ++b;
    return b;
}

__global__ void make_keys_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                 const int64_t* __restrict__ typ, int64_t E, int64_t N, int64_t NS, int64_t n_rel, int tb,
                                 uint64_t* __restrict__ keys, int* __restrict__ vals, int* __restrict__ bad) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t r = row[e], c = col[e], t = typ[e];
    if (r < 0 || r >= N || c < 0 || c >= NS || t < 0 || t >= n_rel) {
        atomicOr(bad, 1);
        r = 0; t = 0;
    }
    keys[e] = ((uint64_t)r << tb) | (uint64_t)t;
    vals[e] = (int)e;
}

__global__ void fill_csr_kernel(const uint64_t* __restrict__ keys, const int* __restrict__ eid,
                                const int64_t* __restrict__ col, int64_t E, int tb, int* __restrict__ csr_src,
                                int* __restrict__ csr_type, int* __restrict__ csr_dst) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= E) return;
    uint64_t k = keys[p];
    csr_dst[p] = (int)(k >> tb);
    csr_type[p] = (int)(k & ((uint64_t(1) << tb) - 1));
    csr_src[p] = (int)col[eid[p]];
}

// ptr[s] = first position whose (sorted) key is >= s, for s in [0, nseg]; ptr[nseg] = E.
__global__ void build_ptr_kernel(const int* __restrict__ keys, int64_t E, int nseg, int* __restrict__ ptr) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p > E) return;
    int prev = (p == 0) ? -1 : keys[p - 1];
    int cur = (p == E) ? nseg : keys[p];
    for (int s = prev + 1; s <= cur; ++s) ptr[s] = (int)p;
}

__global__ void iota_kernel(int* v, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int)i;
}

__global__ void gather_int_kernel(const int* __restrict__ src, const int* __restrict__ idx, int64_t n, int* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

__global__ void sqrt_deg_kernel(const int* __restrict__ rowptr, int64_t N, float* __restrict__ out, int* __restrict__ max_deg) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    int deg = rowptr[i + 1] - rowptr[i];
    out[i] = sqrtf((float)deg);
    int m = deg;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_deg, m);
}

__global__ void item_counts_kernel(const int* __restrict__ ptr, int nseg, int chunk, int keep_empty,
                                   int* __restrict__ n_it, int* __restrict__ n_sl, int* __restrict__ n_sp) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    int len = ptr[s + 1] - ptr[s];
    int k = len == 0 ? keep_empty : (len + chunk - 1) / chunk;
    n_it[s] = k;
    n_sl[s] = k > 1 ? k : 0;
    n_sp[s] = k > 1 ? 1 : 0;
}

__global__ void item_fill_kernel(const int* __restrict__ ptr, int nseg, int chunk, const int* __restrict__ n_it,
                                 const int* __restrict__ off_it, const int* __restrict__ off_sl,
                                 const int* __restrict__ off_sp, Item* __restrict__ items, Split* __restrict__ splits) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    int k = n_it[s];
    if (k == 0) return;
    int beg = ptr[s], end = ptr[s + 1];
    int o = off_it[s];
    if (k == 1) {
        items[o] = Item{s, beg, end, -1};
        return;
    }
    int so = off_sl[s];
    for (int i = 0; i < k; ++i) {
        int b = beg + i * chunk;
        items[o + i] = Item{s, b, min(end, b + chunk), so + i};
    }
    splits[off_sp[s]] = Split{s, so, k, 0};
}

static int build_items(const int* ptr, int nseg, int chunk, int keep_empty, cudaStream_t st, ItemList* out) {
    int *n_it = nullptr, *n_sl = nullptr, *n_sp = nullptr, *o_it = nullptr, *o_sl = nullptr, *o_sp = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    size_t nb = sizeof(int) * (size_t)(nseg + 1);
    Temps tmps(st);
    JMAC_CUDA(tmps.alloc(&n_it, nb)); JMAC_CUDA(tmps.alloc(&n_sl, nb)); JMAC_CUDA(tmps.alloc(&n_sp, nb));
    JMAC_CUDA(tmps.alloc(&o_it, nb)); JMAC_CUDA(tmps.alloc(&o_sl, nb)); JMAC_CUDA(tmps.alloc(&o_sp, nb));
    JMAC_CUDA(cudaMemsetAsync(n_it, 0, nb, st)); JMAC_CUDA(cudaMemsetAsync(n_sl, 0, nb, st)); JMAC_CUDA(cudaMemsetAsync(n_sp, 0, nb, st));
    int threads = 256, blocks = (nseg + threads - 1) / threads;
    if (nseg > 0) {
        item_counts_kernel<<<blocks, threads, 0, st>>>(ptr, nseg, chunk, keep_empty, n_it, n_sl, n_sp);
        JMAC_LAUNCH_CHECK();
    }
    // exclusive scans over nseg+1 entries (last entry of the counts is 0) -> totals land at [nseg]
    JMAC_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, n_it, o_it, nseg + 1, st));
    JMAC_CUDA(tmps.alloc(&tmp, tmp_bytes));
    JMAC_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, n_it, o_it, nseg + 1, st));
    JMAC_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, n_sl, o_sl, nseg + 1, st));
    JMAC_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, n_sp, o_sp, nseg + 1, st));
    int tot[3] = {0, 0, 0};
    JMAC_CUDA(cudaMemcpyAsync(&tot[0], o_it + nseg, sizeof(int), cudaMemcpyDeviceToHost, st));
    JMAC_CUDA(cudaMemcpyAsync(&tot[1], o_sl + nseg, sizeof(int), cudaMemcpyDeviceToHost, st));
    JMAC_CUDA(cudaMemcpyAsync(&tot[2], o_sp + nseg, sizeof(int), cudaMemcpyDeviceToHost, st));
    JMAC_CUDA(cudaStreamSynchronize(st));
    out->n_items = tot[0]; out->n_slots = tot[1]; out->n_splits = tot[2];
    JMAC_CUDA(dmalloc(&out->items, sizeof(Item) * (size_t)(tot[0] > 0 ? tot[0] : 1), st));
    JMAC_CUDA(dmalloc(&out->splits, sizeof(Split) * (size_t)(tot[2] > 0 ? tot[2] : 1), st));
    if (nseg > 0) {
        item_fill_kernel<<<blocks, threads, 0, st>>>(ptr, nseg, chunk, n_it, o_it, o_sl, o_sp, out->items, out->splits);
        JMAC_LAUNCH_CHECK();
    }
    return 0;
}

static int sort_by_int_key(const int* keys_in, int64_t E, int end_bit, cudaStream_t st, int* keys_out, int* pos_out) {
    int* iota = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    Temps tmps(st);
    JMAC_CUDA(tmps.alloc(&iota, sizeof(int) * (size_t)(E > 0 ? E : 1)));
    int threads = 256;
    int64_t blocks = (E + threads - 1) / threads;
    if (E > 0) { iota_kernel<<<(unsigned)blocks, threads, 0, st>>>(iota, E); JMAC_LAUNCH_CHECK(); }
    JMAC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const unsigned*)keys_in, (unsigned*)keys_out, iota, pos_out, (int)E, 0, end_bit, st));
    JMAC_CUDA(tmps.alloc(&tmp, tmp_bytes > 0 ? tmp_bytes : 1));
    JMAC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const unsigned*)keys_in, (unsigned*)keys_out, iota, pos_out, (int)E, 0, end_bit, st));
    return 0;
}

}  // namespace jmac

using namespace jmac;

extern "C" const char* jmac_version(void) { return "jmac_b200 0.1 (sm_100a)"; }
extern "C" const char* jmac_last_error(void) { return jmac::g_err; }
extern "C" uint64_t jmac_launch_count(void) { return __atomic_load_n(&jmac::g_launches, __ATOMIC_RELAXED); }

extern "C" int jmac_set_sm_reserve(int n_sms) {
    const int old = __atomic_exchange_n(&jmac::g_sm_reserve, n_sms > 0 ? n_sms : 0, __ATOMIC_RELAXED);
    return old;
}

extern "C" int jmac_status_poll(void) {
    int* w = jmac::status_words();
    int bits = 0;
    if (w) {
        for (int i = 0; i < jmac::JMAC_ST_WORDS; ++i)
            if (__atomic_exchange_n(&w[i], 0, __ATOMIC_RELAXED)) bits |= 1 << i;
    }
    return bits;
}

extern "C" int jmac_plan_destroy(jmac_plan* p) {
    if (!p) return 0;
    // No host-side wait: the frees are stream-ordered on a per-device "free" stream that first waits for the last kernel that read
    // the plan on every stream it was used on (plan_mark_use).  The memory goes back to the library's block cache (not to the driver),
    // so a rebuilt plan reuses it.  Only a plan used from more than four streams falls back to one device-wide wait.
    int cur = p->device;
    cudaGetDevice(&cur);
    if (cur != p->device) cudaSetDevice(p->device);      // a process that drives several devices may destroy the plan from another one
    cudaStream_t fs = free_stream(p->device);
    if (p->many_streams || fs == nullptr) cudaDeviceSynchronize();
    for (int i = 0; i < p->n_use; ++i) {
        if (fs) cudaStreamWaitEvent(fs, p->use_ev[i], 0);
    }
    void* ptrs[] = {p->rowptr, p->csr_src, p->csr_type, p->csr_eid, p->csr_dst, p->sqrt_deg, p->colptr, p->csc_pos, p->csc_dst, p->typptr,
                    p->typ_pos, p->typ_dst, p->dst_items.items, p->dst_items.splits, p->src_items.items, p->src_items.splits,
                    p->typ_items.items, p->typ_items.splits, p->scratch};
    for (void* q : ptrs) dfree(q, fs);
    // No wait at all: each block carries the event recorded on the free stream; whoever takes it from the cache waits for it on the device.
    for (int i = 0; i < p->n_use; ++i) cudaEventDestroy(p->use_ev[i]);
    if (cur != p->device) cudaSetDevice(cur);
    delete p;
    return 0;
}

static int plan_build(jmac_plan* p, const int64_t* d_row, const int64_t* d_col, const int64_t* d_type, cudaStream_t st) {
    const int64_t E = p->E, N = p->N, NS = p->N_src, R = p->n_rel;
    const size_t Eb = (size_t)(E > 0 ? E : 1);
    const int tb = bits_for(R), nbits = bits_for(N), sbits = bits_for(NS);
    JMAC_CUDA(dmalloc(&p->rowptr, sizeof(int) * (size_t)(N + 1), st));
    JMAC_CUDA(dmalloc(&p->colptr, sizeof(int) * (size_t)(NS + 1), st));
    JMAC_CUDA(dmalloc(&p->typptr, sizeof(int) * (size_t)(R + 1), st));
    JMAC_CUDA(dmalloc(&p->sqrt_deg, sizeof(float) * (size_t)(N > 0 ? N : 1), st));
    JMAC_CUDA(dmalloc(&p->csr_src, sizeof(int) * Eb, st)); JMAC_CUDA(dmalloc(&p->csr_type, sizeof(int) * Eb, st));
    JMAC_CUDA(dmalloc(&p->csr_eid, sizeof(int) * Eb, st)); JMAC_CUDA(dmalloc(&p->csr_dst, sizeof(int) * Eb, st));
    JMAC_CUDA(dmalloc(&p->csc_pos, sizeof(int) * Eb, st)); JMAC_CUDA(dmalloc(&p->csc_dst, sizeof(int) * Eb, st));
    JMAC_CUDA(dmalloc(&p->typ_pos, sizeof(int) * Eb, st)); JMAC_CUDA(dmalloc(&p->typ_dst, sizeof(int) * Eb, st));

    const int threads = 256;
    const unsigned eblocks = (unsigned)((E + threads - 1) / threads);
    const unsigned e1blocks = (unsigned)((E + 1 + threads - 1) / threads);
    int* flags = nullptr;  // [0] = bad index flag, [1] = max degree
    Temps keep(st);
    JMAC_CUDA(keep.alloc(&flags, 2 * sizeof(int)));
    JMAC_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int), st));

    // ---- by destination, secondary key relation type
    {
        uint64_t *k_in = nullptr, *k_out = nullptr;
        int* v_in = nullptr;
        void* tmp = nullptr;
        size_t tmp_bytes = 0;
        Temps tmps(st);
        JMAC_CUDA(tmps.alloc(&k_in, sizeof(uint64_t) * Eb)); JMAC_CUDA(tmps.alloc(&k_out, sizeof(uint64_t) * Eb));
        JMAC_CUDA(tmps.alloc(&v_in, sizeof(int) * Eb));
        if (E > 0) {
            make_keys_kernel<<<eblocks, threads, 0, st>>>(d_row, d_col, d_type, E, N, NS, R, tb, k_in, v_in, flags);
            JMAC_LAUNCH_CHECK();
        }
        JMAC_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in, k_out, v_in, p->csr_eid, (int)E, 0, tb + nbits, st));
        JMAC_CUDA(tmps.alloc(&tmp, tmp_bytes > 0 ? tmp_bytes : 1));
        JMAC_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k_in, k_out, v_in, p->csr_eid, (int)E, 0, tb + nbits, st));
        if (E > 0) {
            fill_csr_kernel<<<eblocks, threads, 0, st>>>(k_out, p->csr_eid, d_col, E, tb, p->csr_src, p->csr_type, p->csr_dst);
            JMAC_LAUNCH_CHECK();
        }
        int bad = 0;
        JMAC_CUDA(cudaMemcpyAsync(&bad, flags, sizeof(int), cudaMemcpyDeviceToHost, st));
        JMAC_CUDA(cudaStreamSynchronize(st));
        if (bad) {
            set_error("jmac_plan_create: edge_index / edge_type entry out of range (N=%lld, n_rel=%lld)", (long long)N, (long long)R);
            return JMAC_ERANGE;
        }
    }
    build_ptr_kernel<<<e1blocks, threads, 0, st>>>(p->csr_dst, E, (int)N, p->rowptr);
    JMAC_LAUNCH_CHECK();
    if (N > 0) {
        sqrt_deg_kernel<<<(unsigned)((N + threads - 1) / threads), threads, 0, st>>>(p->rowptr, N, p->sqrt_deg, flags + 1);
        JMAC_LAUNCH_CHECK();
    }
    JMAC_CUDA(cudaMemcpyAsync(&p->max_deg, flags + 1, sizeof(int), cudaMemcpyDeviceToHost, st));

    // ---- by source (stable: positions stay in destination order inside a source segment)
    {
        int* sorted = nullptr;
        Temps tmps(st);
        JMAC_CUDA(tmps.alloc(&sorted, sizeof(int) * Eb));
        int rc = sort_by_int_key(p->csr_src, E, sbits, st, sorted, p->csc_pos);
        if (rc) return rc;
        build_ptr_kernel<<<e1blocks, threads, 0, st>>>(sorted, E, (int)NS, p->colptr);
        JMAC_LAUNCH_CHECK();
        if (E > 0) { gather_int_kernel<<<eblocks, threads, 0, st>>>(p->csr_dst, p->csc_pos, E, p->csc_dst); JMAC_LAUNCH_CHECK(); }
        // ---- by relation type
        rc = sort_by_int_key(p->csr_type, E, tb, st, sorted, p->typ_pos);
        if (rc) return rc;
        build_ptr_kernel<<<e1blocks, threads, 0, st>>>(sorted, E, (int)R, p->typptr);
        JMAC_LAUNCH_CHECK();
        if (E > 0) { gather_int_kernel<<<eblocks, threads, 0, st>>>(p->csr_dst, p->typ_pos, E, p->typ_dst); JMAC_LAUNCH_CHECK(); }
    }
    int rc = build_items(p->rowptr, (int)N, p->chunk, 1, st, &p->dst_items);
    if (rc) return rc;
    rc = build_items(p->colptr, (int)NS, p->chunk, 1, st, &p->src_items);
    if (rc) return rc;
    // relation pass: larger chunks (one CTA-wide warp loop per item), empty relations skipped
    rc = build_items(p->typptr, (int)R, p->chunk, 0, st, &p->typ_items);
    return rc;
}

extern "C" int jmac_plan_create_bipartite(const int64_t* d_row, const int64_t* d_col, const int64_t* d_type, int64_t E, int64_t N_dst,
                                          int64_t N_src, int64_t n_rel, int chunk, void* stream, jmac_plan** out) {
    JMAC_ARG(out != nullptr);
    *out = nullptr;
    JMAC_ARG(E >= 0 && E < (int64_t(1) << 31) - 1);
    JMAC_ARG(N_dst > 0 && N_dst < (int64_t(1) << 31) - 1);
    JMAC_ARG(N_src > 0 && N_src < (int64_t(1) << 31) - 1);
    JMAC_ARG(n_rel > 0 && n_rel < (int64_t(1) << 24));
    JMAC_ARG(E == 0 || (d_row && d_col && d_type));
    JMAC_ARG(chunk >= 0 && chunk <= (1 << 20));
    jmac_plan* p = new (std::nothrow) jmac_plan();
    if (!p) return JMAC_ENOMEM;
    p->E = E; p->N = N_dst; p->N_src = N_src; p->n_rel = n_rel;
    // Work-item size: 256 edges on large graphs; on the reference's graph sizes (1e4-1e5 edges) a warp's fair share of the
    // edges is a few dozen, and a 256-edge slice of a hub would be 8x that -- the whole launch would wait for the hub warps.
    if (chunk > 0) p->chunk = chunk;
    else {
        const int64_t per = E / ((int64_t)sm_count() * RING_WARPS);
        int c = 32;
        while (c < 256 && c < per) c <<= 1;
        p->chunk = c;
    }
    if (cudaGetDevice(&p->device) != cudaSuccess) p->device = 0;
    int rc = plan_build(p, d_row, d_col, d_type, (cudaStream_t)stream);
    if (rc == 0 && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) rc = JMAC_EINVAL;   // complete on return: usable from any stream
    if (rc) { jmac_plan_destroy(p); return rc; }
    *out = p;
    return 0;
}

extern "C" int jmac_plan_create(const int64_t* d_row, const int64_t* d_col, const int64_t* d_type, int64_t E, int64_t N,
                                int64_t n_rel, int chunk, void* stream, jmac_plan** out) {
    return jmac_plan_create_bipartite(d_row, d_col, d_type, E, N, N, n_rel, chunk, stream, out);
}

extern "C" int jmac_plan_info(const jmac_plan* p, int64_t info[8]) {
    JMAC_ARG(p && info);
    info[0] = p->E; info[1] = p->N; info[2] = p->n_rel; info[3] = p->chunk;
    info[4] = p->dst_items.n_items; info[5] = p->src_items.n_items; info[6] = p->typ_items.n_items; info[7] = p->max_deg;
    return 0;
}

extern "C" int jmac_plan_export(const jmac_plan* p, int32_t* rowptr, int32_t* csr_src, int32_t* csr_type, int32_t* csr_eid,
                                float* sqrt_deg, void* stream) {
    JMAC_ARG(p != nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    if (rowptr) JMAC_CUDA(cudaMemcpyAsync(rowptr, p->rowptr, sizeof(int) * (size_t)(p->N + 1), cudaMemcpyDeviceToDevice, st));
    if (csr_src && p->E) JMAC_CUDA(cudaMemcpyAsync(csr_src, p->csr_src, sizeof(int) * (size_t)p->E, cudaMemcpyDeviceToDevice, st));
    if (csr_type && p->E) JMAC_CUDA(cudaMemcpyAsync(csr_type, p->csr_type, sizeof(int) * (size_t)p->E, cudaMemcpyDeviceToDevice, st));
    if (csr_eid && p->E) JMAC_CUDA(cudaMemcpyAsync(csr_eid, p->csr_eid, sizeof(int) * (size_t)p->E, cudaMemcpyDeviceToDevice, st));
    if (sqrt_deg) JMAC_CUDA(cudaMemcpyAsync(sqrt_deg, p->sqrt_deg, sizeof(float) * (size_t)p->N, cudaMemcpyDeviceToDevice, st));
    plan_mark_use(const_cast<jmac_plan*>(p), st);
    return 0;
}
