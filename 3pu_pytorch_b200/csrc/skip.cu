// Inter-level skip connection (bilateral feature interpolation), fused, fp32, sm_100a.
//
// Replaces network/upsampler.py:317-347 + exponential_distance (:232-250) of the reference:
//     gather the K=5 nearest previous-level points' coordinates and 264-channel features,
//     spatial weights  ws = exp(-ds / (hs/2)),  ds = |xyz - nb_xyz|^2,  hs = mean_N(min_K ds)   per patch
//     feature weights  wf = exp(-df / (hf/2)),  df = |x - nb_feat|^2,   hf = mean_N(min_K df)   per patch
//     w = ws*wf / sum_K(ws*wf + 1e-5);   x += 0.2 * sum_K w * nb_feat
// The reference materialises the gathered (B,264,N,K) tensor (and several same-sized temporaries) in HBM.
// Here one CTA owns one patch: pass 1 computes ds/df for its N points into shared memory and reduces the two
// per-patch means, pass 2 re-gathers and writes x in place.  The previous level's features are kept
// POINT-major ((B,No,C): one neighbour = one contiguous 1 KB row) so every gather is a coalesced row read;
// the patch's own features (channel-major (T,C,N), as the convolutions want them) go through a transposing
// shared-memory tile.  Bound: L2/HBM gather bandwidth (2 x N*K*C*4 bytes per patch).
#include "pu3_common.cuh"

namespace pu3 {

constexpr int SK_THREADS = 256;
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int SK_PT = 32;     // points per tile
constexpr int SK_KMAX = 8;

struct SkipArgs {
    int t, n, c, k, p_div, no;
    float *x;                  // (t,c,n) in/out, channel-major
    const float *xyz;          // (t,3,n)
    const int64_t *idx;        // (t,n,k) neighbour indices into the owner's previous cloud
    const float *prev_xyz;     // (clouds,3,no) channel-major
    const float *prev_feat;    // (clouds,no,c) POINT-major
    const int32_t *owner;      // (t) or null -> t / p_div
    float *w_out;              // (t,n,k) or null: the normalised interpolation weights (saved for the backward pass)
    float *pm_out;             // (t,n,c) or null: the updated features once more, POINT-major (what the next level gathers from)
};

__global__ void __launch_bounds__(SK_THREADS) skip_fuse_kernel(SkipArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int C = a.c, N = a.n, K = a.k;
    float *xt = sm;                                  // [C][SK_PT+1] transposing tile of the patch's features
    float *sds = xt + (size_t)C * (SK_PT + 1);       // [N][K]
    float *sdf = sds + (size_t)N * K;                // [N][K]
    float *red = sdf + (size_t)N * K;                // [2][SK_WARPS] partial sums of the minima
    __shared__ float s_h[2];

    const int ti = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cloud = a.owner ? __ldg(a.owner + ti) : ti / a.p_div;
    float *xb = a.x + (size_t)ti * C * N;
    const float *pxyz = a.prev_xyz + (size_t)cloud * 3 * a.no;
    const float *pf = a.prev_feat + (size_t)cloud * a.no * C;
    const float *q = a.xyz + (size_t)ti * 3 * N;
    const int64_t *ib = a.idx + (size_t)ti * N * K;

    float sum_s = 0.f, sum_f = 0.f;   // lane 0 of each warp accumulates the minima of the warp's points
    // ---------------- pass 1: squared distances -------------------------------------------------------
    for (int p0 = 0; p0 < N; p0 += SK_PT) {
        const int pc = min(SK_PT, N - p0);
        __syncthreads();
        for (int t = threadIdx.x; t < C * SK_PT; t += SK_THREADS) {
            const int ch = t / SK_PT, pl = t % SK_PT;
            xt[ch * (SK_PT + 1) + pl] = pl < pc ? xb[(size_t)ch * N + p0 + pl] : 0.f;
        }
        __syncthreads();
        for (int pl = warp; pl < pc; pl += SK_WARPS) {
            const int i = p0 + pl;
            float mins = INFINITY, minf = INFINITY;
            for (int kk = 0; kk < K; ++kk) {
                const int j = (int)ib[(size_t)i * K + kk];
                const float *row = pf + (size_t)j * C;
                float acc = 0.f;
                for (int ch = lane; ch < C; ch += 32) {
                    const float d = xt[ch * (SK_PT + 1) + pl] - __ldg(row + ch);
                    acc = __fmaf_rn(d, d, acc);
                }
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
                const float dx = __ldg(q + i) - __ldg(pxyz + j);
                const float dy = __ldg(q + N + i) - __ldg(pxyz + a.no + j);
                const float dz = __ldg(q + 2 * N + i) - __ldg(pxyz + 2 * (size_t)a.no + j);
                const float ds = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                if (lane == 0) { sds[i * K + kk] = ds; sdf[i * K + kk] = acc; }
                mins = fminf(mins, ds); minf = fminf(minf, acc);
            }
            sum_s += mins; sum_f += minf;
        }
    }
    if (lane == 0) { red[warp] = sum_s; red[SK_WARPS + warp] = sum_f; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ts = 0.f, tf = 0.f;
        for (int w = 0; w < SK_WARPS; ++w) { ts += red[w]; tf += red[SK_WARPS + w]; }
        s_h[0] = ts / (float)N;    // h = mean_N(min_K d)   (:247)
        s_h[1] = tf / (float)N;
    }
    __syncthreads();
    const float hs2 = s_h[0] / 2.0f, hf2 = s_h[1] / 2.0f;   // h / 2 (:249)

    // ---------------- pass 2: weights, interpolation, in-place update ------------------------------------
    for (int p0 = 0; p0 < N; p0 += SK_PT) {
        const int pc = min(SK_PT, N - p0);
        __syncthreads();
        for (int t = threadIdx.x; t < C * SK_PT; t += SK_THREADS) {
            const int ch = t / SK_PT, pl = t % SK_PT;
            xt[ch * (SK_PT + 1) + pl] = pl < pc ? xb[(size_t)ch * N + p0 + pl] : 0.f;
        }
        __syncthreads();
        for (int pl = warp; pl < pc; pl += SK_WARPS) {
            const int i = p0 + pl;
            float w[SK_KMAX];
            float wsum = 0.f;
#pragma unroll
            for (int kk = 0; kk < SK_KMAX; ++kk) {
                w[kk] = 0.f;
                if (kk < K) {
                    const float ws = expf(-sds[i * K + kk] / hs2);
                    const float wf = expf(-sdf[i * K + kk] / hf2);
                    w[kk] = ws * wf;
                    wsum += w[kk] + 1e-5f;             // sum(average_weight + 1e-5) (:341-342)
                }
            }
#pragma unroll
            for (int kk = 0; kk < SK_KMAX; ++kk) w[kk] = w[kk] / wsum;
            if (a.w_out && lane == 0) {
#pragma unroll
                for (int kk = 0; kk < SK_KMAX; ++kk) if (kk < K) a.w_out[((size_t)ti * N + i) * K + kk] = w[kk];
            }
            for (int ch = lane; ch < C; ch += 32) {
                float acc = 0.f;
#pragma unroll
                for (int kk = 0; kk < SK_KMAX; ++kk) {
                    if (kk < K) {
                        const int j = (int)ib[(size_t)i * K + kk];
                        acc = __fmaf_rn(w[kk], __ldg(pf + (size_t)j * C + ch), acc);
                    }
                }
                float *cell = &xt[ch * (SK_PT + 1) + pl];
                const float nv = __fmaf_rn(0.2f, acc, *cell);   // x = 0.2 * knnIdx_feats + x (:347)
                *cell = nv;
                if (a.pm_out) a.pm_out[((size_t)ti * N + i) * C + ch] = nv;
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < C * SK_PT; t += SK_THREADS) {
            const int ch = t / SK_PT, pl = t % SK_PT;
            if (pl < pc) xb[(size_t)ch * N + p0 + pl] = xt[ch * (SK_PT + 1) + pl];
        }
    }
}

// Specialisation for the reference configuration (K = fm_knn = 5 neighbours, C = 264 channels), same arithmetic in
// the same order as the generic kernel above (results are bit-identical), but with every trip count known at
// compile time: the 5 neighbour indices are read first and the 5 x 9 row loads of a point are all in flight before
// the first use.  The generic kernel exposes one row (8-9 loads) at a time and is bound by L2 gather latency
// (ncu: long-scoreboard 6.3 warps per issue, profiles/r1f_ncu_summary.md).
// 3 CTAs per SM (ncu, profiles/r2: at 115 registers and the default carve-out only 2 CTAs = 16 warps were resident and the kernel
// waited on its gathers: long-scoreboard 4.8 warps per issue at 41 % issue-active).
template <int K, int C>
__global__ void __launch_bounds__(SK_THREADS, 3) skip_fuse_fixed_kernel(SkipArgs a) {
    extern __shared__ __align__(16) float sm[];
    constexpr int U = (C + 31) / 32;                 // channels per lane
    const int N = a.n;
    float *xt = sm;                                  // [C][SK_PT+1]
    float *sds = xt + (size_t)C * (SK_PT + 1);       // [N][K]
    float *sdf = sds + (size_t)N * K;                // [N][K]
    float *red = sdf + (size_t)N * K;                // [2][SK_WARPS]
    __shared__ float s_h[2];

    const int ti = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cloud = a.owner ? __ldg(a.owner + ti) : ti / a.p_div;
    float *xb = a.x + (size_t)ti * C * N;
    const float *pxyz = a.prev_xyz + (size_t)cloud * 3 * a.no;
    const float *pf = a.prev_feat + (size_t)cloud * a.no * C;
    const float *q = a.xyz + (size_t)ti * 3 * N;
    const int64_t *ib = a.idx + (size_t)ti * N * K;

    float sum_s = 0.f, sum_f = 0.f;
    for (int p0 = 0; p0 < N; p0 += SK_PT) {
        const int pc = min(SK_PT, N - p0);
        __syncthreads();
        for (int t = threadIdx.x; t < C * SK_PT; t += SK_THREADS) {
            const int ch = t / SK_PT, pl = t % SK_PT;
            xt[ch * (SK_PT + 1) + pl] = pl < pc ? xb[(size_t)ch * N + p0 + pl] : 0.f;
        }
        __syncthreads();
        for (int pl = warp; pl < pc; pl += SK_WARPS) {
            const int i = p0 + pl;
            int j[K];
#pragma unroll
            for (int kk = 0; kk < K; ++kk) j[kk] = (int)ib[(size_t)i * K + kk];
            float nb[K][U];
#pragma unroll
            for (int kk = 0; kk < K; ++kk)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int ch = lane + 32 * u;
                    nb[kk][u] = ch < C ? __ldg(pf + (size_t)j[kk] * C + ch) : 0.f;
                }
            float xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int ch = lane + 32 * u;
                xv[u] = ch < C ? xt[ch * (SK_PT + 1) + pl] : 0.f;
            }
            const float qx = __ldg(q + i), qy = __ldg(q + N + i), qz = __ldg(q + 2 * N + i);
            float mins = INFINITY, minf = INFINITY;
#pragma unroll
            for (int kk = 0; kk < K; ++kk) {
                float acc = 0.f;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (lane + 32 * u < C) {
                        const float d = xv[u] - nb[kk][u];
                        acc = __fmaf_rn(d, d, acc);
                    }
                }
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
                const float dx = qx - __ldg(pxyz + j[kk]);
                const float dy = qy - __ldg(pxyz + a.no + j[kk]);
                const float dz = qz - __ldg(pxyz + 2 * (size_t)a.no + j[kk]);
                const float ds = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                if (lane == 0) { sds[i * K + kk] = ds; sdf[i * K + kk] = acc; }
                mins = fminf(mins, ds); minf = fminf(minf, acc);
            }
            sum_s += mins; sum_f += minf;
        }
    }
    if (lane == 0) { red[warp] = sum_s; red[SK_WARPS + warp] = sum_f; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ts = 0.f, tf = 0.f;
        for (int w = 0; w < SK_WARPS; ++w) { ts += red[w]; tf += red[SK_WARPS + w]; }
        s_h[0] = ts / (float)N;
        s_h[1] = tf / (float)N;
    }
    __syncthreads();
    const float hs2 = s_h[0] / 2.0f, hf2 = s_h[1] / 2.0f;

    for (int p0 = 0; p0 < N; p0 += SK_PT) {
        const int pc = min(SK_PT, N - p0);
        __syncthreads();
        for (int t = threadIdx.x; t < C * SK_PT; t += SK_THREADS) {
            const int ch = t / SK_PT, pl = t % SK_PT;
            xt[ch * (SK_PT + 1) + pl] = pl < pc ? xb[(size_t)ch * N + p0 + pl] : 0.f;
        }
        __syncthreads();
        for (int pl = warp; pl < pc; pl += SK_WARPS) {
            const int i = p0 + pl;
            int j[K];
#pragma unroll
            for (int kk = 0; kk < K; ++kk) j[kk] = (int)ib[(size_t)i * K + kk];
            float nb[K][U];
#pragma unroll
            for (int kk = 0; kk < K; ++kk)
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int ch = lane + 32 * u;
                    nb[kk][u] = ch < C ? __ldg(pf + (size_t)j[kk] * C + ch) : 0.f;
                }
            float w[K];
            float wsum = 0.f;
#pragma unroll
            for (int kk = 0; kk < K; ++kk) {
                const float ws = expf(-sds[i * K + kk] / hs2);
                const float wf = expf(-sdf[i * K + kk] / hf2);
                w[kk] = ws * wf;
                wsum += w[kk] + 1e-5f;
            }
#pragma unroll
            for (int kk = 0; kk < K; ++kk) w[kk] = w[kk] / wsum;
            if (a.w_out && lane == 0) {
#pragma unroll
                for (int kk = 0; kk < K; ++kk) a.w_out[((size_t)ti * N + i) * K + kk] = w[kk];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int ch = lane + 32 * u;
                if (ch < C) {
                    float acc = 0.f;
#pragma unroll
                    for (int kk = 0; kk < K; ++kk) acc = __fmaf_rn(w[kk], nb[kk][u], acc);
                    float *cell = &xt[ch * (SK_PT + 1) + pl];
                    const float nv = __fmaf_rn(0.2f, acc, *cell);
                    *cell = nv;
                    // the warp holds the point's whole updated row, channel = lane + 32 u: the point-major copy the next level's skip
                    // connection gathers from is written from here (coalesced), instead of by a transposing pass over x afterwards
                    if (a.pm_out) a.pm_out[((size_t)ti * N + i) * C + ch] = nv;
                }
            }
        }
        __syncthreads();
        for (int t = threadIdx.x; t < C * SK_PT; t += SK_THREADS) {
            const int ch = t / SK_PT, pl = t % SK_PT;
            if (pl < pc) xb[(size_t)ch * N + p0 + pl] = xt[ch * (SK_PT + 1) + pl];
        }
    }
}

// Backward of the skip connection with respect to the previous level's features (the weights are detached in the
// reference, upsampler.py:243,249, so x' = x + 0.2 sum_k w_k f_{j_k} is linear in f):
//     dprev[cloud, j_k(i), :] += 0.2 * w_k(i) * dx'[:, i]        (dx = dx' passes through unchanged)
// One warp per point, the tile's gradient columns through a transposing shared-memory tile, whole-row atomics.
__global__ void __launch_bounds__(SK_THREADS) skip_bwd_kernel(int n, int c, int k, int p_div, int no, const float *__restrict__ dx,
                                                              const int64_t *__restrict__ idx, const float *__restrict__ w,
                                                              const int32_t *__restrict__ owner, float *__restrict__ dprev) {
    extern __shared__ __align__(16) float sm[];
    float *xt = sm;                                  // [c][SK_PT+1]
    const int ti = blockIdx.y, p0 = blockIdx.x * SK_PT;
    const int pc = min(SK_PT, n - p0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cloud = owner ? __ldg(owner + ti) : ti / p_div;
    const float *db = dx + (size_t)ti * c * n;
    for (int t = threadIdx.x; t < c * SK_PT; t += SK_THREADS) {
        const int ch = t / SK_PT, pl = t % SK_PT;
        xt[ch * (SK_PT + 1) + pl] = pl < pc ? db[(size_t)ch * n + p0 + pl] : 0.f;
    }
    __syncthreads();
    float *pf = dprev + (size_t)cloud * no * c;
    for (int pl = warp; pl < pc; pl += SK_WARPS) {
        const int i = p0 + pl;
        for (int kk = 0; kk < k; ++kk) {
            const int j = (int)idx[((size_t)ti * n + i) * k + kk];
            const float wk = 0.2f * __ldg(w + ((size_t)ti * n + i) * k + kk);
            if (wk == 0.f) continue;
            for (int ch = lane; ch < c; ch += 32) atomicAdd(pf + (size_t)j * c + ch, wk * xt[ch * (SK_PT + 1) + pl]);
        }
    }
}

// (T,C,N) channel-major -> rows of a point-major (rows, C) buffer: out[slot[t]*N + i][c] = in[t][c][i]
// (the features a level hands to the next one, laid out for skip_fuse_kernel's row gathers)
__global__ void __launch_bounds__(256) to_point_major_kernel(int c, int n, const float *__restrict__ in,
                                                            const int64_t *__restrict__ slot, float *__restrict__ out) {
    __shared__ float tile[32][33];
    const int t = blockIdx.z;
    const int c0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const float *src = in + (size_t)t * c * n;
    for (int r = ty; r < 32; r += 8) {
        const int ch = c0 + r, i = i0 + tx;
        tile[r][tx] = (ch < c && i < n) ? src[(size_t)ch * n + i] : 0.f;
    }
    __syncthreads();
    const long long row0 = (slot ? slot[t] : (long long)t) * n;
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, ch = c0 + tx;
        if (i < n && ch < c) out[(size_t)(row0 + i) * c + ch] = tile[tx][r];
    }
}

}  // namespace pu3

using namespace pu3;

// Test hook: force the generic (runtime K, C) kernel.
static int g_skip_force_generic = 0;
extern "C" void pu3_skip_force_generic(int on) { g_skip_force_generic = on; }

extern "C" int pu3_skip_fuse_f32(int t, int n, int c, int k, int p_div, int no, float *x, const float *xyz,
                                 const int64_t *idx, const float *prev_xyz, const float *prev_feat_pm,
                                 const int32_t *owner, pu3_stream_t stream) {
    return pu3_skip_fuse_ex_f32(t, n, c, k, p_div, no, x, xyz, idx, prev_xyz, prev_feat_pm, owner, nullptr, stream);
}

extern "C" int pu3_skip_bwd_f32(int t, int n, int c, int k, int p_div, int no, const float *dx, const int64_t *idx,
                                const float *w, const int32_t *owner, float *dprev_feat_pm, pu3_stream_t stream) {
    PU3_ARG_CHECK(t >= 0 && n > 0 && c > 0 && k > 0 && no > 0, "skip_bwd: bad size t=%d n=%d c=%d k=%d no=%d", t, n, c, k, no);
    if (t == 0) return PU3_OK;
    PU3_ARG_CHECK(owner || p_div >= 1, "skip_bwd: need owner or p_div");
    PU3_ARG_CHECK(dx && idx && w && dprev_feat_pm && t <= 65535, "skip_bwd: null pointer or t > 65535");
    const size_t smem = (size_t)c * (SK_PT + 1) * sizeof(float);
    PU3_ARG_CHECK(smem <= (size_t)device_info().smem_optin, "skip_bwd: c=%d needs %zu bytes of shared memory", c, smem);
    int st = cuda_status(cudaFuncSetAttribute(skip_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "skip_bwd: smem attr");
    if (st) return st;
    skip_bwd_kernel<<<dim3((n + SK_PT - 1) / SK_PT, t), SK_THREADS, smem, as_stream(stream)>>>(n, c, k, p_div, no, dx, idx, w, owner, dprev_feat_pm);
    PU3_LAUNCH_CHECK("skip_bwd_kernel");
    return PU3_OK;
}

extern "C" int pu3_skip_fuse_ex_f32(int t, int n, int c, int k, int p_div, int no, float *x, const float *xyz,
                                    const int64_t *idx, const float *prev_xyz, const float *prev_feat_pm,
                                    const int32_t *owner, float *w_out, pu3_stream_t stream) {
    return pu3_skip_fuse_pm_f32(t, n, c, k, p_div, no, x, xyz, idx, prev_xyz, prev_feat_pm, owner, w_out, nullptr, stream);
}

extern "C" int pu3_skip_fuse_pm_f32(int t, int n, int c, int k, int p_div, int no, float *x, const float *xyz,
                                    const int64_t *idx, const float *prev_xyz, const float *prev_feat_pm,
                                    const int32_t *owner, float *w_out, float *x_pm_out, pu3_stream_t stream) {
    PU3_ARG_CHECK(t >= 0 && n > 0 && c > 0 && k > 0 && no > 0, "skip_fuse: bad size t=%d n=%d c=%d k=%d no=%d", t, n, c, k, no);
    if (t == 0) return PU3_OK;
    PU3_ARG_CHECK(k <= SK_KMAX, "skip_fuse: k=%d exceeds %d", k, SK_KMAX);
    PU3_ARG_CHECK(owner || p_div >= 1, "skip_fuse: need owner or p_div");
    PU3_ARG_CHECK(x && xyz && idx && prev_xyz && prev_feat_pm, "skip_fuse: null pointer");
    SkipArgs a{t, n, c, k, p_div, no, x, xyz, idx, prev_xyz, prev_feat_pm, owner, w_out, x_pm_out};
    const size_t smem = ((size_t)c * (SK_PT + 1) + 2 * (size_t)n * k + 2 * SK_WARPS) * sizeof(float);
    PU3_ARG_CHECK(smem <= (size_t)device_info().smem_optin, "skip_fuse: c=%d n=%d k=%d needs %zu bytes of shared memory", c, n, k, smem);
    int st = cuda_status(cudaFuncSetAttribute(skip_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "skip_fuse: smem attr");
    if (st) return st;
    if (k == 5 && c == 264 && !g_skip_force_generic) {
        st = cuda_status(cudaFuncSetAttribute(skip_fuse_fixed_kernel<5, 264>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "skip_fuse: smem attr");
        if (st) return st;
        st = cuda_status(cudaFuncSetAttribute(skip_fuse_fixed_kernel<5, 264>, cudaFuncAttributePreferredSharedMemoryCarveout, 100), "skip_fuse: carve-out");
        if (st) return st;
        skip_fuse_fixed_kernel<5, 264><<<t, SK_THREADS, smem, as_stream(stream)>>>(a);
        PU3_LAUNCH_CHECK("skip_fuse_fixed_kernel");
        return PU3_OK;
    }
    skip_fuse_kernel<<<t, SK_THREADS, smem, as_stream(stream)>>>(a);
    PU3_LAUNCH_CHECK("skip_fuse_kernel");
    return PU3_OK;
}

extern "C" int pu3_to_point_major_f32(int t, int c, int n, const float *in, const int64_t *slot, float *out,
                                      pu3_stream_t stream) {
    PU3_ARG_CHECK(t >= 0 && c > 0 && n > 0, "to_point_major: bad size");
    if (t == 0) return PU3_OK;
    PU3_ARG_CHECK(t <= 65535 && in && out, "to_point_major: null pointer or t > 65535");
    dim3 grid((n + 31) / 32, (c + 31) / 32, t);
    to_point_major_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, n, in, slot, out);
    PU3_LAUNCH_CHECK("to_point_major_kernel");
    return PU3_OK;
}
