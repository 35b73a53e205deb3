// Point gather forward / backward and ball query, sm_100a.
//
// Replaces in sampling/sampling_cuda.cu of the reference:
//   gather_points_forward_kernel   :26-62   out[b,c,j] = points[b,c,idx[b,j]]
//   gather_points_backward_kernel  :64-100  grad_points[b,c,idx[b,j]] += grad_out[b,c,j]
//   query_ball_point_kernel        :267-314 (dead code in the reference, kept for the surface)
// The reference launches grid (b,c) with <=512 threads on the legacy stream; here one flat
// grid-stride launch covers all (b,c,j) with j fastest, so index reads and output writes are
// coalesced and the launch fills the chip whatever the (b,c) split is.  HBM/latency-bound.
#include <cuda_fp16.h>

#include "pu3_common.cuh"

namespace pu3 {

template <typename T>
__global__ void __launch_bounds__(256) gather_fwd_kernel(int c, int n, int m, long long total,
                                                        const T *__restrict__ points,
                                                        const int32_t *__restrict__ idx, T *__restrict__ out) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % m);
        const long long bc = t / m;          // = b*c + l
        const long long b = bc / c;
        const int a = __ldg(idx + b * m + j);
        out[t] = __ldg(points + bc * n + a);
    }
}

template <typename T>
__device__ __forceinline__ void atomic_add_any(T *p, T v) { atomicAdd(p, v); }

template <typename T>
__global__ void __launch_bounds__(256) gather_bwd_kernel(int c, int n, int m, long long total,
                                                        const T *__restrict__ grad_out,
                                                        const int32_t *__restrict__ idx,
                                                        T *__restrict__ grad_points) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % m);
        const long long bc = t / m;
        const long long b = bc / c;
        const int a = __ldg(idx + b * m + j);
        atomic_add_any(grad_points + bc * n + a, grad_out[t]);
    }
}

__global__ void __launch_bounds__(128) ball_query_kernel(int n, int m, float radius2, int nsample,
                                                        const float *__restrict__ new_xyz,
                                                        const float *__restrict__ xyz, int32_t *__restrict__ idx) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const float *q = new_xyz + ((size_t)b * m + j) * 3;
    const float qx = q[0], qy = q[1], qz = q[2];
    const float *p = xyz + (size_t)b * n * 3;
    int32_t *o = idx + ((size_t)b * m + j) * nsample;
    int cnt = 0;
    for (int k = 0; k < n && cnt < nsample; ++k) {
        const float d2 = sqdist3(qx - p[k * 3 + 0], qy - p[k * 3 + 1], qz - p[k * 3 + 2]);
        if (d2 < radius2) {
            if (cnt == 0)
                for (int l = 0; l < nsample; ++l) o[l] = k;
            o[cnt] = k;
            ++cnt;
        }
    }
}

static int flat_grid(long long total) {
    const long long blocks = (total + 255) / 256;
    const long long cap = (long long)device_info().sm_count * 8;  // 8 CTAs of 256 threads per SM
    return (int)(blocks < cap ? blocks : cap);
}

}  // namespace pu3

using namespace pu3;

extern "C" int pu3_gather_fwd(int b, int c, int n, int m, int elem_bytes, const void *points,
                              const int32_t *idx, void *out, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && c >= 0 && n >= 0 && m >= 0, "gather_fwd: negative size");
    const long long total = (long long)b * c * m;
    if (total == 0) return PU3_OK;
    PU3_ARG_CHECK(n > 0, "gather_fwd: gathering %d indices from an empty point set", m);
    PU3_ARG_CHECK(points && idx && out, "gather_fwd: null pointer");
    cudaStream_t s = as_stream(stream);
    const int g = flat_grid(total);
    switch (elem_bytes) {
        case 2: gather_fwd_kernel<uint16_t><<<g, 256, 0, s>>>(c, n, m, total, (const uint16_t *)points, idx, (uint16_t *)out); break;
        case 4: gather_fwd_kernel<uint32_t><<<g, 256, 0, s>>>(c, n, m, total, (const uint32_t *)points, idx, (uint32_t *)out); break;
        case 8: gather_fwd_kernel<uint64_t><<<g, 256, 0, s>>>(c, n, m, total, (const uint64_t *)points, idx, (uint64_t *)out); break;
        default: set_error("gather_fwd: elem_bytes=%d (want 2, 4 or 8)", elem_bytes); return PU3_E_ARG;
    }
    PU3_LAUNCH_CHECK("gather_fwd_kernel");
    return PU3_OK;
}

extern "C" int pu3_gather_bwd(int b, int c, int n, int m, int dtype, const void *grad_out, const int32_t *idx,
                              void *grad_points, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && c >= 0 && n >= 0 && m >= 0, "gather_bwd: negative size");
    const long long total = (long long)b * c * m;
    if (total == 0) return PU3_OK;
    PU3_ARG_CHECK(n > 0 && grad_out && idx && grad_points, "gather_bwd: null pointer or empty target");
    cudaStream_t s = as_stream(stream);
    const int g = flat_grid(total);
    switch (dtype) {
        case 0: gather_bwd_kernel<float><<<g, 256, 0, s>>>(c, n, m, total, (const float *)grad_out, idx, (float *)grad_points); break;
        case 1: gather_bwd_kernel<double><<<g, 256, 0, s>>>(c, n, m, total, (const double *)grad_out, idx, (double *)grad_points); break;
        case 2: gather_bwd_kernel<__half><<<g, 256, 0, s>>>(c, n, m, total, (const __half *)grad_out, idx, (__half *)grad_points); break;
        default: set_error("gather_bwd: dtype=%d (0 f32, 1 f64, 2 f16)", dtype); return PU3_E_ARG;
    }
    PU3_LAUNCH_CHECK("gather_bwd_kernel");
    return PU3_OK;
}

extern "C" int pu3_ball_query_f32(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                                  const float *xyz, int32_t *idx, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && n >= 0 && m >= 0 && nsample >= 0, "ball_query: negative size");
    if (b == 0 || m == 0 || nsample == 0) return PU3_OK;
    PU3_ARG_CHECK(b <= 65535 && new_xyz && xyz && idx, "ball_query: null pointer or b > 65535");
    ball_query_kernel<<<dim3((m + 127) / 128, b), 128, 0, as_stream(stream)>>>(n, m, radius * radius, nsample,
                                                                                new_xyz, xyz, idx);
    PU3_LAUNCH_CHECK("ball_query_kernel");
    return PU3_OK;
}
