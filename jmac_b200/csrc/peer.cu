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
