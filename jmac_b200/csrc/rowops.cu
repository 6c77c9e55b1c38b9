// Row-wise L2 normalisation with its backward: F.normalize(x, 2, -1) of JMAC.forward_name / get_emb / the DBPv1 scorer
// (src/jmac_model.py:183,192,230-231; JMAC_DBPv1/models/jmac_model.py:245-247), one warp per row, one pass each way
// (the eager op is a norm reduction, a clamp and a division forward, five elementwise / reduction kernels backward).
// HBM-bound: 8 B/element forward, 12 B/element backward.
#include <math.h>

#include "common.cuh"

namespace jmac {

// y = x / max(|x|, eps); inv[row] = 1 / max(|x|, eps), stored NEGATIVE when the clamp was active (the norm then has no gradient)
template <int VPL>
__global__ void __launch_bounds__(256) row_normalize_fwd_kernel(const float* __restrict__ x, int64_t N, int d4, float eps, float* __restrict__ y,
                                                                float* __restrict__ inv) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    Row<VPL> r;
    load_row<VPL, true>(r, x, row, d4, lane);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) ss += dot4(r.v[i], r.v[i]);
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float s = 1.f / fmaxf(nrm, eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) r.v[i] = s * r.v[i];
    store_row<VPL>(r, y, row, d4, lane);
    if (lane == 0 && inv) inv[row] = nrm > eps ? s : -s;
}

// dx = inv (g - y (y.g))   (clamped rows: dx = |inv| g)
template <int VPL>
__global__ void __launch_bounds__(256) row_normalize_bwd_kernel(const float* __restrict__ y, const float* __restrict__ inv,
                                                                const float* __restrict__ g, int64_t N, int d4, float* __restrict__ dx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= N) return;
    Row<VPL> yr, gr;
    load_row<VPL, true>(yr, y, row, d4, lane);
    load_row<VPL, true>(gr, g, row, d4, lane);
    const float s = inv[row];
    float dt = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) dt += dot4(yr.v[i], gr.v[i]);
    dt = s > 0.f ? warp_sum(dt) : 0.f;
    const float a = fabsf(s);
#pragma unroll
    for (int i = 0; i < VPL; ++i) gr.v[i] = a * (gr.v[i] - dt * yr.v[i]);
    store_row<VPL>(gr, dx, row, d4, lane);
}

}  // namespace jmac

using namespace jmac;

extern "C" int jmac_row_normalize_fwd(const float* x, int64_t N, int d, float eps, float* y, float* inv, void* stream) {
    JMAC_ARG(x && y && N >= 0 && dim_ok(d) && eps > 0.f);
    if (N == 0) return 0;
    const unsigned blocks = (unsigned)((N * 32 + 255) / 256);
    JMAC_DISPATCH_VPL(d, { row_normalize_fwd_kernel<VPL><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, N, d / 4, eps, y, inv); });
    JMAC_LAUNCH_CHECK();
    return 0;
}

extern "C" int jmac_row_normalize_bwd(const float* y, const float* inv, const float* g, int64_t N, int d, float* dx, void* stream) {
    JMAC_ARG(y && inv && g && dx && N >= 0 && dim_ok(d));
    if (N == 0) return 0;
    const unsigned blocks = (unsigned)((N * 32 + 255) / 256);
    JMAC_DISPATCH_VPL(d, { row_normalize_bwd_kernel<VPL><<<blocks, 256, 0, (cudaStream_t)stream>>>(y, inv, g, N, d / 4, dx); });
    JMAC_LAUNCH_CHECK();
    return 0;
}
