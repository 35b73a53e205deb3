// The small data-movement / reduction steps between the big kernels of the eval path, as kernels of their own so that a
// whole Net.forward is a fixed sequence of launches (CUDA-graph capturable, no eager-framework elementwise ops):
//   normalize_point_batch                      network/operations.py:12-30
//   outlier mask + masked_select compaction    network/upsampler.py:63-76
//   seeds of the tiles (gather after FPS)      network/upsampler.py:78 (+ operations.py:320)
//   tiles -> flat list of patches + normalise  network/upsampler.py:83-85, 138
//   de-normalise + merge the tiles of a shape  network/upsampler.py:144, 149-155
// Arithmetic is spelled with round-to-nearest intrinsics in the operator order of the reference's torch expressions
// (x*r + c is a multiply and an add, not an FMA), so results are those of the eager composition.
#include "pu3_common.cuh"

namespace pu3 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum of three values + (second phase) max of one; blockDim.x a multiple of 32, <= 1024
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float *red /* NV*32 floats */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * 32 + warp] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float t = lane < nw ? red[i * 32 + lane] : 0.f;
        v[i] = warp_sum(t);
    }
}
__device__ __forceinline__ float block_max(float v, float *red /* 32 floats */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = lane < nw ? red[lane] : -INFINITY;
    return warp_max(t);
}

// ---- normalize_point_batch (operations.py:12-30) ---------------------------------------------------------------------------
// cloud b: element (c, i) at in[b*bs + c*cs + i*ps]; same layout out.  centroid (B,3), radius (B).
__global__ void __launch_bounds__(256) normalize_kernel(int n, long long bs, long long cs, long long ps, const float *__restrict__ in,
                                                        float *__restrict__ out, float *__restrict__ centroid,
                                                        float *__restrict__ radius) {
    __shared__ float red[96];
    const float *src = in + (size_t)blockIdx.x * bs;
    float *dst = out + (size_t)blockIdx.x * bs;
    float s[3] = {0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        s[0] += src[i * ps]; s[1] += src[cs + i * ps]; s[2] += src[2 * cs + i * ps];
    }
    block_sum<3>(s, red);
    const float inv_n = (float)n;
    const float cx = __fdiv_rn(s[0], inv_n), cy = __fdiv_rn(s[1], inv_n), cz = __fdiv_rn(s[2], inv_n);   // torch.mean = sum / n
    float far = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float x = __fsub_rn(src[i * ps], cx), y = __fsub_rn(src[cs + i * ps], cy), z = __fsub_rn(src[2 * cs + i * ps], cz);
        far = fmaxf(far, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))));
    }
    far = block_max(far, red);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        dst[i * ps] = __fdiv_rn(__fsub_rn(src[i * ps], cx), far);
        dst[cs + i * ps] = __fdiv_rn(__fsub_rn(src[cs + i * ps], cy), far);
        dst[2 * cs + i * ps] = __fdiv_rn(__fsub_rn(src[2 * cs + i * ps], cz), far);
    }
    if (threadIdx.x == 0) {
        centroid[blockIdx.x * 3 + 0] = cx; centroid[blockIdx.x * 3 + 1] = cy; centroid[blockIdx.x * 3 + 2] = cz;
        radius[blockIdx.x] = far;
    }
}

// ---- outlier filter (upsampler.py:63-76) ---------------------------------------------------------------------------------
// d (B,N,dk): distance to the dk nearest neighbours of every point, column 1 = nearest OTHER point.  A point is kept when
// d < 5 * mean_N(d).  Kept points first, order preserved (masked_select), removed points behind them.  Outputs: compacted cloud
// channel-major (B,3,N) and point-major (B,N,3); n_arr[b] = max(count, k) (a request with count < k is flagged in *bad and its
// result discarded by the caller, but it must still read valid memory); p_arr[b] = int(count / k * 5) computed in double like the
// host expression of :76.
constexpr int OC_THREADS = 1024;
__global__ void __launch_bounds__(OC_THREADS) outlier_compact_kernel(int n, int dk, int k, int r, const float *__restrict__ d,
                                                                     const float *__restrict__ xyz, float *__restrict__ out_cm,
                                                                     float *__restrict__ out_pm, int32_t *__restrict__ n_arr,
                                                                     int32_t *__restrict__ p_arr, int32_t *__restrict__ pk_arr,
                                                                     int32_t *__restrict__ pkr_arr, int32_t *__restrict__ bad) {
    __shared__ float red[32];
    __shared__ int wcount[32];
    __shared__ int s_total;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *dist = d + (size_t)b * n * dk + 1;
    const float *src = xyz + (size_t)b * 3 * n;
    float s[1] = {0.f};
    for (int i = tid; i < n; i += OC_THREADS) s[0] += dist[(size_t)i * dk];
    block_sum<1>(s, red);
    const float thr = __fmul_rn(5.f, __fdiv_rn(s[0], (float)n));
    // contiguous segment per thread -> order-preserving compaction with one block scan
    const int seg = (n + OC_THREADS - 1) / OC_THREADS;
    const int i0 = min(tid * seg, n), i1 = min(i0 + seg, n);
    int mine = 0;
    for (int i = i0; i < i1; ++i) mine += dist[(size_t)i * dk] < thr;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wcount[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = wcount[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        wcount[lane] = wi - w;           // exclusive prefix of the warps
        if (lane == 31) s_total = wi;
    }
    __syncthreads();
    const int count = s_total;
    int kept_pos = wcount[warp] + incl - mine;             // kept points before my segment
    int drop_pos = count + (i0 - kept_pos);                // removed points before my segment, behind all kept ones
    float *ocm = out_cm + (size_t)b * 3 * n;
    float *opm = out_pm + (size_t)b * 3 * n;
    for (int i = i0; i < i1; ++i) {
        const bool keep = dist[(size_t)i * dk] < thr;
        const int pos = keep ? kept_pos++ : drop_pos++;
        const float x = src[i], y = src[n + i], z = src[2 * n + i];
        ocm[pos] = x; ocm[n + pos] = y; ocm[2 * n + pos] = z;
        opm[3 * pos] = x; opm[3 * pos + 1] = y; opm[3 * pos + 2] = z;
    }
    if (tid == 0) {
        n_arr[b] = max(count, min(k, n));
        const int p = (int)((double)count / (double)k * 5.0);
        p_arr[b] = p;
        if (pk_arr) pk_arr[b] = p * k;            // points of the request's tiles side by side (next level's skip search)
        if (pkr_arr) pkr_arr[b] = p * k * r;      // ... and of its upsampled tiles (merge FPS, :158)
        if (count < k) atomicOr(bad, 1);
    }
}

// ---- seeds of the tiles: seeds[b,c,j] = xyz[b,c,idx[b, j < p_arr[b] ? j : 0]] (spare slots repeat the first tile) -----------
__global__ void tile_seeds_kernel(int n, int p, const float *__restrict__ xyz, const int32_t *__restrict__ idx,
                                  const int32_t *__restrict__ p_arr, float *__restrict__ seeds) {
    const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    const int src = idx[(size_t)b * p + (j < p_arr[b] ? j : 0)];
#pragma unroll
    for (int c = 0; c < 3; ++c) seeds[((size_t)b * 3 + c) * p + j] = xyz[((size_t)b * 3 + c) * n + src];
}

// ---- tiles (B,3,P,k) -> patches (B*P,3,k) raw + normalised, centroid (B*P,3), radius (B*P), and the tiles of a request side by
// side (B,3,P*k): the cloud the next level's skip connection searches (upsampler.py:85,138,148-152) ---------------------------
// (256 threads like normalize_kernel: the same per-thread strides and reduction tree, hence bit-identical centroids / radii)
__global__ void __launch_bounds__(256) tiles_normalize_kernel(int p, int k, const float *__restrict__ tiles, float *__restrict__ patch,
                                                              float *__restrict__ patch_norm, float *__restrict__ centroid,
                                                              float *__restrict__ radius, float *__restrict__ side_by_side) {
    __shared__ float red[96];
    const int t = blockIdx.x, b = t / p, j = t - b * p;
    const float *src = tiles + ((size_t)b * 3 * p + j) * k;          // channel c at + c*p*k
    const size_t cs = (size_t)p * k;
    float s[3] = {0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < k; i += blockDim.x) { s[0] += src[i]; s[1] += src[cs + i]; s[2] += src[2 * cs + i]; }
    block_sum<3>(s, red);
    const float fk = (float)k;
    const float cx = __fdiv_rn(s[0], fk), cy = __fdiv_rn(s[1], fk), cz = __fdiv_rn(s[2], fk);
    float far = 0.f;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const float x = __fsub_rn(src[i], cx), y = __fsub_rn(src[cs + i], cy), z = __fsub_rn(src[2 * cs + i], cz);
        far = fmaxf(far, __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))));
    }
    far = block_max(far, red);
    float *raw = patch + (size_t)t * 3 * k, *nrm = patch_norm + (size_t)t * 3 * k;
    float *sbs = side_by_side ? side_by_side + (size_t)b * 3 * cs + (size_t)j * k : nullptr;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = src[c * cs + i];
            const float ctr = c == 0 ? cx : (c == 1 ? cy : cz);
            raw[c * k + i] = v;
            nrm[c * k + i] = __fdiv_rn(__fsub_rn(v, ctr), far);
            if (sbs) sbs[c * cs + i] = v;
        }
    }
    if (threadIdx.x == 0) {
        centroid[t * 3 + 0] = cx; centroid[t * 3 + 1] = cy; centroid[t * 3 + 2] = cz;
        radius[t] = far;
    }
}

// ---- de-normalise (xyz * radius + centroid, :144) and put the tiles of a request side by side (:149-155), POINT-major
// (B, P*kr, 3): the layout the FPS kernels read ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) denorm_merge_kernel(int p, int kr, const float *__restrict__ xyz_norm,
                                                           const float *__restrict__ centroid, const float *__restrict__ radius,
                                                           float *__restrict__ merged_pm) {
    const int t = blockIdx.y, b = t / p, j = t - b * p;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kr) return;
    const float r = radius[t];
    const float *src = xyz_norm + (size_t)t * 3 * kr;
    float *dst = merged_pm + (((size_t)b * p + j) * kr + i) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) dst[c] = __fadd_rn(__fmul_rn(src[c * kr + i], r), centroid[t * 3 + c]);
}

// ---- gather from a point-major cloud into channel-major samples: out[b,c,j] = pts[b, idx[b,j], c] ----------------------------
__global__ void gather_pm_kernel(int n, int m, const float *__restrict__ pts, const int32_t *__restrict__ idx, float *__restrict__ out) {
    const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const float *src = pts + ((size_t)b * n + idx[(size_t)b * m + j]) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[((size_t)b * 3 + c) * m + j] = src[c];
}

// ---- backward helpers of the expansion head (upsampler.py:349-372) -------------------------------------------------------------
// out[t,c,i] (+)= sum_{j<r} in[t,c,i*r+j]: the gradient of "replicate every point r times" (:356-358, :371)
__global__ void __launch_bounds__(256) replica_sum_kernel(long long rows_n, int r, const float *__restrict__ in, float *__restrict__ out,
                                                          int accumulate) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows_n) return;
    float s = 0.f;
    for (int j = 0; j < r; ++j) s += in[e * r + j];
    out[e] = accumulate ? out[e] + s : s;
}
// g[e] = act[e] > 0 ? g[e] : 0: the ReLU derivative applied to a gradient in place
__global__ void __launch_bounds__(256) relu_mask_kernel(long long total, float *__restrict__ g, const float *__restrict__ act) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < total && !(act[e] > 0.f)) g[e] = 0.f;
}

}  // namespace pu3

using namespace pu3;

extern "C" int pu3_replica_sum_f32(long long rows, int n, int r, const float *in, float *out, int accumulate, pu3_stream_t stream) {
    PU3_ARG_CHECK(rows >= 0 && n >= 0 && r > 0, "replica_sum: bad size");
    const long long total = rows * n;
    if (total == 0) return PU3_OK;
    PU3_ARG_CHECK(in && out, "replica_sum: null pointer");
    replica_sum_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(total, r, in, out, accumulate);
    PU3_LAUNCH_CHECK("replica_sum_kernel");
    return PU3_OK;
}

extern "C" int pu3_relu_mask_f32(long long total, float *g, const float *act, pu3_stream_t stream) {
    PU3_ARG_CHECK(total >= 0, "relu_mask: bad size");
    if (total == 0) return PU3_OK;
    PU3_ARG_CHECK(g && act, "relu_mask: null pointer");
    relu_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(total, g, act);
    PU3_LAUNCH_CHECK("relu_mask_kernel");
    return PU3_OK;
}

extern "C" int pu3_normalize_f32(int b, int n, int nchw, const float *pc, float *out, float *centroid, float *radius,
                                 pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && n > 0, "normalize: bad size b=%d n=%d", b, n);
    if (b == 0) return PU3_OK;
    PU3_ARG_CHECK(pc && out && centroid && radius, "normalize: null pointer");
    const long long bs = 3LL * n, cs = nchw ? n : 1, ps = nchw ? 1 : 3;
    normalize_kernel<<<b, 256, 0, as_stream(stream)>>>(n, bs, cs, ps, pc, out, centroid, radius);
    PU3_LAUNCH_CHECK("normalize_kernel");
    return PU3_OK;
}

extern "C" int pu3_outlier_compact_f32(int b, int n, int dk, int k, int r, const float *dist, const float *xyz, float *out_cm,
                                       float *out_pm, int32_t *n_arr, int32_t *p_arr, int32_t *pk_arr, int32_t *pkr_arr,
                                       int32_t *bad, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && n > 0 && dk >= 2 && k > 0, "outlier_compact: bad size b=%d n=%d dk=%d k=%d", b, n, dk, k);
    if (b == 0) return PU3_OK;
    PU3_ARG_CHECK(dist && xyz && out_cm && out_pm && n_arr && p_arr && bad, "outlier_compact: null pointer");
    outlier_compact_kernel<<<b, OC_THREADS, 0, as_stream(stream)>>>(n, dk, k, r, dist, xyz, out_cm, out_pm, n_arr, p_arr, pk_arr, pkr_arr, bad);
    PU3_LAUNCH_CHECK("outlier_compact_kernel");
    return PU3_OK;
}

extern "C" int pu3_tile_seeds_f32(int b, int n, int p, const float *xyz, const int32_t *idx, const int32_t *p_arr, float *seeds,
                                  pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && n > 0 && p >= 0, "tile_seeds: bad size");
    if (b == 0 || p == 0) return PU3_OK;
    PU3_ARG_CHECK(xyz && idx && p_arr && seeds && b <= 65535, "tile_seeds: null pointer or b > 65535");
    tile_seeds_kernel<<<dim3((p + 127) / 128, b), 128, 0, as_stream(stream)>>>(n, p, xyz, idx, p_arr, seeds);
    PU3_LAUNCH_CHECK("tile_seeds_kernel");
    return PU3_OK;
}

extern "C" int pu3_tiles_normalize_f32(int b, int p, int k, const float *tiles, float *patch, float *patch_norm, float *centroid,
                                       float *radius, float *side_by_side, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && p >= 0 && k > 0, "tiles_normalize: bad size");
    if (b == 0 || p == 0) return PU3_OK;
    PU3_ARG_CHECK(tiles && patch && patch_norm && centroid && radius, "tiles_normalize: null pointer");
    tiles_normalize_kernel<<<b * p, 256, 0, as_stream(stream)>>>(p, k, tiles, patch, patch_norm, centroid, radius, side_by_side);
    PU3_LAUNCH_CHECK("tiles_normalize_kernel");
    return PU3_OK;
}

extern "C" int pu3_denorm_merge_f32(int b, int p, int kr, const float *xyz_norm, const float *centroid, const float *radius,
                                    float *merged_pm, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && p >= 0 && kr > 0, "denorm_merge: bad size");
    if (b == 0 || p == 0) return PU3_OK;
    PU3_ARG_CHECK(xyz_norm && centroid && radius && merged_pm && (long long)b * p <= 65535, "denorm_merge: null pointer or too many tiles");
    denorm_merge_kernel<<<dim3((kr + 255) / 256, b * p), 256, 0, as_stream(stream)>>>(p, kr, xyz_norm, centroid, radius, merged_pm);
    PU3_LAUNCH_CHECK("denorm_merge_kernel");
    return PU3_OK;
}

extern "C" int pu3_gather_pm_f32(int b, int n, int m, const float *pts, const int32_t *idx, float *out, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && n > 0 && m >= 0, "gather_pm: bad size");
    if (b == 0 || m == 0) return PU3_OK;
    PU3_ARG_CHECK(pts && idx && out && b <= 65535, "gather_pm: null pointer or b > 65535");
    gather_pm_kernel<<<dim3((m + 255) / 256, b), 256, 0, as_stream(stream)>>>(n, m, pts, idx, out);
    PU3_LAUNCH_CHECK("gather_pm_kernel");
    return PU3_OK;
}
