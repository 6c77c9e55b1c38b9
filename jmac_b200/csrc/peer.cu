// Peer-memory plumbing for the multi-GPU exchange (jmac_b200/peer.py): buffers that other ranks of the node read directly over
// NVLink with copy-engine transfers (cudaMemcpyAsync on peer-mapped pointers), so that the halo all-gather and the gradient
// "reduce-scatter" (peer pulls + a local sum) cost no SM time next to the persistent one-CTA-per-SM compute kernels.
// NCCL's kernels occupy SMs for the whole transfer; measured at 8 GPUs (profiles/multi_gpu_overlap_r02.md) that cost could only be
// moved between "exposed communication" and "slower compute", not removed.
// Plain cudaMalloc allocations (one allocation = one IPC handle, offset 0), CUDA IPC handles exchanged by the host side.
#include <string.h>

#include "common.cuh"

using namespace jmac;

extern "C" int jmac_peer_alloc(size_t bytes, void** out) {
    JMAC_ARG(out != nullptr && bytes > 0);
    *out = nullptr;
    JMAC_CUDA(cudaMalloc(out, bytes));
    return 0;
}

extern "C" int jmac_peer_free(void* ptr) {
    if (ptr) JMAC_CUDA(cudaFree(ptr));
    return 0;
}

// handle: 64 bytes (cudaIpcMemHandle_t)
extern "C" int jmac_ipc_get_handle(void* ptr, unsigned char* handle) {
    JMAC_ARG(ptr && handle);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    JMAC_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, 64);
    return 0;
}

// Maps another process's allocation into this process (peer access between the two devices is enabled lazily by the driver).
extern "C" int jmac_ipc_open(const unsigned char* handle, void** out) {
    JMAC_ARG(handle && out);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    *out = nullptr;
    JMAC_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int jmac_ipc_close(void* ptr) {
    if (ptr) JMAC_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

// Device-to-device copy by the copy engines (unified addressing resolves local / peer pointers).
extern "C" int jmac_memcpy_async(void* dst, const void* src, size_t bytes, void* stream) {
    JMAC_ARG(dst && src);
    if (bytes == 0) return 0;
    JMAC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

// ---- cross-process ordering without an SM: stream memory operations on flag words ------------------------------------------
// A rank announces "my producer kernels before this point are done" by a 4-byte copy-engine write of an epoch number into a
// flag word of every peer (jmac_memcpy_async from a local word set by jmac_stream_write_u32), and waits for its peers with
// jmac_stream_wait_geq_u32 on its OWN flag words.  Neither needs an SM, so the handshake cannot be held up by the persistent
// one-CTA-per-SM compute kernels the way a (tiny) NCCL kernel is.
#include <cuda.h>

namespace {
typedef CUresult (*wait32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
void* driver_fn(const char* name) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return p;
}
}  // namespace

extern "C" int jmac_stream_write_u32(void* addr, unsigned int value, void* stream) {
    JMAC_ARG(addr && ((uintptr_t)addr & 3) == 0);
    static wait32_fn fn = (wait32_fn)driver_fn("cuStreamWriteValue32");
    if (!fn) { set_error("cuStreamWriteValue32 not available from the driver"); return JMAC_EINVAL; }
    CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WRITE_VALUE_DEFAULT);
    if (r != CUDA_SUCCESS) { set_error("cuStreamWriteValue32 failed with CUresult %d", (int)r); return JMAC_EINVAL; }
    return 0;
}

// The stream waits until (int32)(*addr - value) >= 0 (cyclic comparison: epochs may wrap).
extern "C" int jmac_stream_wait_geq_u32(void* addr, unsigned int value, void* stream) {
    JMAC_ARG(addr && ((uintptr_t)addr & 3) == 0);
    static wait32_fn fn = (wait32_fn)driver_fn("cuStreamWaitValue32");
    if (!fn) { set_error("cuStreamWaitValue32 not available from the driver"); return JMAC_EINVAL; }
    CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) { set_error("cuStreamWaitValue32 failed with CUresult %d", (int)r); return JMAC_EINVAL; }
    return 0;
}

// ---- the local half of the peer "reduce-scatter": dst[i] (+)= sum_g src[g * stride + i], g ascending (deterministic) --------
namespace {
__global__ void __launch_bounds__(256) peer_sum_kernel(float4* __restrict__ dst, const float4* __restrict__ src, int n_slices, size_t stride4,
                                                       size_t n4, int accumulate) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = accumulate ? dst[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4* p = src + i;
#pragma unroll 4
        for (int g = 0; g < n_slices; ++g) {
            const float4 v = __ldcs(p + (size_t)g * stride4);      // streamed once: do not keep in L2
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        dst[i] = acc;
    }
}
}  // namespace

extern "C" int jmac_peer_sum(float* dst, const float* src, int n_slices, int64_t slice_stride, int64_t n, int accumulate, void* stream) {
    JMAC_ARG(dst && src && n_slices >= 1 && n >= 0 && slice_stride >= n && n % 4 == 0 && slice_stride % 4 == 0);
    JMAC_ARG(((uintptr_t)dst & 15) == 0 && ((uintptr_t)src & 15) == 0);
    if (n == 0) return 0;
    const size_t n4 = (size_t)n / 4;
    long long blocks = (long long)((n4 + 255) / 256);
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    peer_sum_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float4*)dst, (const float4*)src, n_slices, (size_t)slice_stride / 4, n4,
                                                                       accumulate);
    JMAC_LAUNCH_CHECK();
    return 0;
}
