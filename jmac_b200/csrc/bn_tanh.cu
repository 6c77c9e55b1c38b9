// BatchNorm1d over the N nodes + tanh (src/jmac_model.py:52: act(bn((neigh + self)/2))), forward and
// backward, split into a per-feature statistics kernel and an elementwise apply kernel so that a
// multi-GPU caller can all-reduce the 2*d sums in between (BN couples all destination shards).
// Streaming, HBM-bound: forward reads pre twice + writes out (12 B/elem), backward reads
// (gout, out, pre) twice + writes gpre (28 B/elem).
#include <math.h>

#include "common.cuh"

namespace jmac {

// Column sums in double: each thread owns one float4 column group and strides over rows.
// blockDim = (CX, CY) with CX >= d4 rounded to a multiple of 32 (<=128), CY = 256 / CX.
template <int NV>
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd, int64_t N, int d4,
                                                        double* __restrict__ sums) {
    // NV == 1: sums of (a, a^2).  NV == 3: a=gout, b=out, c=pre -> sums of (dy, dy*xhat).
    const int cx = threadIdx.x, cy = threadIdx.y;
    const int CY = blockDim.y;
    double s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
    float4 mu = f4_zero(), rs = f4_zero();
    if (NV == 3 && cx < d4) {
        mu = reinterpret_cast<const float4*>(mean)[cx];
        rs = reinterpret_cast<const float4*>(rstd)[cx];
    }
    if (cx < d4) {
        for (int64_t r = (int64_t)blockIdx.x * CY + cy; r < N; r += (int64_t)gridDim.x * CY) {
            const float4 va = ld_stream(reinterpret_cast<const float4*>(a) + r * d4 + cx);
            if (NV == 1) {
                s0[0] += va.x; s0[1] += va.y; s0[2] += va.z; s0[3] += va.w;
                s1[0] += (double)va.x * va.x; s1[1] += (double)va.y * va.y; s1[2] += (double)va.z * va.z; s1[3] += (double)va.w * va.w;
            } else {
                const float4 vo = ld_stream(reinterpret_cast<const float4*>(b) + r * d4 + cx);
                const float4 vp = ld_stream(reinterpret_cast<const float4*>(c) + r * d4 + cx);
                const float dy0 = va.x * (1.f - vo.x * vo.x), dy1 = va.y * (1.f - vo.y * vo.y);
                const float dy2 = va.z * (1.f - vo.z * vo.z), dy3 = va.w * (1.f - vo.w * vo.w);
                s0[0] += dy0; s0[1] += dy1; s0[2] += dy2; s0[3] += dy3;
                s1[0] += (double)dy0 * ((vp.x - mu.x) * rs.x); s1[1] += (double)dy1 * ((vp.y - mu.y) * rs.y);
                s1[2] += (double)dy2 * ((vp.z - mu.z) * rs.z); s1[3] += (double)dy3 * ((vp.w - mu.w) * rs.w);
            }
        }
    }
    __shared__ double sh[2][1024];  // [which][(cy * CX + cx) * 4 + k], CX * CY == 256
    const int CX = blockDim.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sh[0][(cy * CX + cx) * 4 + k] = s0[k];
        sh[1][(cy * CX + cx) * 4 + k] = s1[k];
    }
    __syncthreads();
    if (cy == 0 && cx < d4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double t0 = 0, t1 = 0;
            for (int y = 0; y < CY; ++y) { t0 += sh[0][(y * CX + cx) * 4 + k]; t1 += sh[1][(y * CX + cx) * 4 + k]; }
            atomicAdd(sums + cx * 4 + k, t0);
            atomicAdd(sums + d4 * 4 + cx * 4 + k, t1);
        }
    }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, int64_t n_total, int d, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ save_mean, float* __restrict__ save_rstd, int training) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    if (training) {
        const double n = (double)n_total;
        const double mu = sums[c] / n;
        double var = sums[d + c] / n - mu * mu;
        if (var < 0) var = 0;
        save_mean[c] = (float)mu;
        save_rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
        if (running_var) {
            const double unb = n_total > 1 ? var * n / (n - 1.0) : var;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
        }
    } else {
        save_mean[c] = running_mean[c];
        save_rstd[c] = 1.f / sqrtf(running_var[c] + eps);
    }
}

__global__ void __launch_bounds__(256) bn_tanh_apply_kernel(const float* __restrict__ pre, int64_t n4, int d4, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % d4);
        const float4 x = ld_stream(reinterpret_cast<const float4*>(pre) + i);
        const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c), rs = __ldg(reinterpret_cast<const float4*>(rstd) + c);
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c), b = __ldg(reinterpret_cast<const float4*>(beta) + c);
        float4 y;
        y.x = tanhf(fmaf((x.x - mu.x) * rs.x, g.x, b.x));
        y.y = tanhf(fmaf((x.y - mu.y) * rs.y, g.y, b.y));
        y.z = tanhf(fmaf((x.z - mu.z) * rs.z, g.z, b.z));
        y.w = tanhf(fmaf((x.w - mu.w) * rs.w, g.w, b.w));
        reinterpret_cast<float4*>(out)[i] = y;
    }
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ bsums, int d, float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d) return;
    if (dbeta) dbeta[c] = (float)bsums[c];
    if (dgamma) dgamma[c] = (float)bsums[d + c];
}

__global__ void __launch_bounds__(256) bn_tanh_bwd_apply_kernel(const float* __restrict__ gout, const float* __restrict__ out,
                                                                const float* __restrict__ pre, int64_t n4, int d4,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                const float* __restrict__ gamma, const double* __restrict__ bsums,
                                                                double inv_n, int training, float* __restrict__ gpre) {
    const int d = d4 * 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % d4);
        const float4 go = ld_stream(reinterpret_cast<const float4*>(gout) + i);
        const float4 o = ld_stream(reinterpret_cast<const float4*>(out) + i);
        const float4 rs = __ldg(reinterpret_cast<const float4*>(rstd) + c), g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        float4 dy = make_float4(go.x * (1.f - o.x * o.x), go.y * (1.f - o.y * o.y), go.z * (1.f - o.z * o.z), go.w * (1.f - o.w * o.w));
        float4 r;
        if (training) {
            const float4 x = ld_stream(reinterpret_cast<const float4*>(pre) + i);
            const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c);
            const float sb[4] = {(float)(bsums[4 * c] * inv_n), (float)(bsums[4 * c + 1] * inv_n), (float)(bsums[4 * c + 2] * inv_n), (float)(bsums[4 * c + 3] * inv_n)};
            const float sg[4] = {(float)(bsums[d + 4 * c] * inv_n), (float)(bsums[d + 4 * c + 1] * inv_n), (float)(bsums[d + 4 * c + 2] * inv_n), (float)(bsums[d + 4 * c + 3] * inv_n)};
            r.x = g.x * rs.x * (dy.x - sb[0] - (x.x - mu.x) * rs.x * sg[0]);
            r.y = g.y * rs.y * (dy.y - sb[1] - (x.y - mu.y) * rs.y * sg[1]);
            r.z = g.z * rs.z * (dy.z - sb[2] - (x.z - mu.z) * rs.z * sg[2]);
            r.w = g.w * rs.w * (dy.w - sb[3] - (x.w - mu.w) * rs.w * sg[3]);
        } else {
            r = make_float4(g.x * rs.x * dy.x, g.y * rs.y * dy.y, g.z * rs.z * dy.z, g.w * rs.w * dy.w);
        }
        reinterpret_cast<float4*>(gpre)[i] = r;
    }
}

static void stats_dims(int d4, dim3* block) {
    int cx = ((d4 + 31) / 32) * 32;
    if (cx > 128) cx = 128;
    *block = dim3(cx, 256 / cx, 1);
}

static int stream_blocks(int64_t n_threads_wanted) {
    int64_t b = (n_threads_wanted + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace jmac

using namespace jmac;

extern "C" int jmac_bn_stats(const float* pre, int64_t N, int d, double* sums, void* stream) {
    JMAC_ARG(pre && sums && N >= 0 && dim_ok(d));
    if (N == 0) return 0;
    dim3 block;
    stats_dims(d / 4, &block);
    int64_t want = (N + block.y * 32 - 1) / (block.y * 32);          // >= 32 rows per thread: the double atomics at the end stay few
    int blocks = (int)(want < (int64_t)sm_count() * 4 ? want : (int64_t)sm_count() * 4);
    col_stats_kernel<1><<<blocks, block, 0, (cudaStream_t)stream>>>(pre, nullptr, nullptr, nullptr, nullptr, N, d / 4, sums);
    JMAC_LAUNCH_CHECK();
    return 0;
}

extern "C" int jmac_bn_tanh_apply(const float* pre, int64_t N, int d, const double* sums, int64_t n_total, const float* gamma,
                                  const float* beta, float* running_mean, float* running_var, float momentum, float eps, int training,
                                  float* out, float* save_mean, float* save_rstd, void* stream) {
    JMAC_ARG(pre && gamma && beta && out && save_mean && save_rstd && N >= 0 && dim_ok(d));
    JMAC_ARG(training ? (sums != nullptr && n_total > 0) : (running_mean != nullptr && running_var != nullptr));
    cudaStream_t st = (cudaStream_t)stream;
    bn_finalize_kernel<<<(d + 127) / 128, 128, 0, st>>>(sums, n_total, d, eps, momentum, running_mean, running_var, save_mean, save_rstd, training);
    JMAC_LAUNCH_CHECK();
    if (N == 0) return 0;
    const int64_t n4 = N * (d / 4);
    bn_tanh_apply_kernel<<<stream_blocks(n4), 256, 0, st>>>(pre, n4, d / 4, save_mean, save_rstd, gamma, beta, out);
    JMAC_LAUNCH_CHECK();
    return 0;
}

extern "C" int jmac_bn_tanh_bwd_stats(const float* gout, const float* out, const float* pre, int64_t N, int d, const float* mean,
                                      const float* rstd, double* bsums, void* stream) {
    JMAC_ARG(gout && out && pre && mean && rstd && bsums && N >= 0 && dim_ok(d));
    if (N == 0) return 0;
    dim3 block;
    stats_dims(d / 4, &block);
    int64_t want = (N + block.y * 32 - 1) / (block.y * 32);
    int blocks = (int)(want < (int64_t)sm_count() * 4 ? want : (int64_t)sm_count() * 4);
    col_stats_kernel<3><<<blocks, block, 0, (cudaStream_t)stream>>>(gout, out, pre, mean, rstd, N, d / 4, bsums);
    JMAC_LAUNCH_CHECK();
    return 0;
}

extern "C" int jmac_bn_tanh_bwd_apply(const float* gout, const float* out, const float* pre, int64_t N, int d, const float* mean,
                                      const float* rstd, const float* gamma, const double* bsums, int64_t n_total, int training,
                                      float* gpre, float* dgamma, float* dbeta, void* stream) {
    JMAC_ARG(gout && out && pre && mean && rstd && gamma && bsums && gpre && N >= 0 && dim_ok(d) && n_total > 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (dgamma || dbeta) {
        bn_bwd_finalize_kernel<<<(d + 127) / 128, 128, 0, st>>>(bsums, d, dgamma, dbeta);
        JMAC_LAUNCH_CHECK();
    }
    if (N == 0) return 0;
    const int64_t n4 = N * (d / 4);
    bn_tanh_bwd_apply_kernel<<<stream_blocks(n4), 256, 0, st>>>(gout, out, pre, n4, d / 4, mean, rstd, gamma, bsums, 1.0 / (double)n_total,
                                                                 training, gpre);
    JMAC_LAUNCH_CHECK();
    return 0;
}
