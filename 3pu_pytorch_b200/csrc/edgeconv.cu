// Fused DenseEdgeConv forward, fp32, sm_100a.
//
// Replaces DenseEdgeConv.forward / get_local_graph of the reference (network/layers.py:22-64), which
// materialises in HBM, per block and per batch of B clouds of N points with K neighbours:
//     the gathered neighbours (B,24,N,K+1), the edge tensor (B,48,N,K), a repeat() of x (B,24,N,K),
//     and the growing concatenations (B,36,N,K), (B,48,N,K), (B,60,N,K)           (:40-61)
// i.e. ~250 MB of traffic for B=32, N=312, K=32 to produce a (B,60,N) result of 2.4 MB, through
// 3 cuDNN launches + cat/relu/max kernels.  Here one kernel reads x (B,24,N) and the neighbour indices
// and writes y (B,60,N); nothing per-edge ever leaves the SM.
//
// Algebra (exact re-association of the reference's 1x1 convolutions, no approximation):
//   mlp0([c, n-c]) = W0[:, :24] c + W0[:, 24:] (n-c) + b0           -> per point: A0 = W0[:, :24] c + b0
//   mlp1([h0, c])  = W1[:, :12] h0 + W1[:, 12:] c + b1              -> per point: A1 = W1[:, 12:] c + b1
//   mlp2([h1, h0, c]) = W2[:, :12] h1 + W2[:, 12:24] h0 + W2[:, 24:] c + b2 -> A2 = W2[:, 24:] c + b2
//   y = max_k [h2, h1, h0, c] = [max_k h2, max_k h1, max_k h0, c]    (h0, h1 after ReLU; h2 without: :58-59)
// The per-point terms are computed once per point instead of once per edge: 720 instead of 1584 MACs per edge.
//
// Mapping: one warp per point, one lane per edge (K = 32 in the reference configuration); the cloud's
// features sit in shared memory (point-major, padded) so the neighbour gather is an LDS; weights are read
// as broadcast LDS.128.  Bound: fp32 FFMA issue, not HBM.
#include "pu3_common.cuh"

namespace pu3 {

constexpr int EC_C = 24;       // input channels
constexpr int EC_G = 12;       // growth rate
constexpr int EC_OUT = 60;     // 3*G + C
constexpr int EC_XS = 25;      // padded row length of a point in shared memory (odd: conflict-free scalar LDS)
constexpr int EC_WARPS = 8;
constexpr int EC_PT = 32;      // points per output staging tile

struct EcWeights {             // all (out,in) row-major as in the state_dict
    const float *w0, *b0;      // (12,48)
    const float *w1, *b1;      // (12,36)
    const float *w2, *b2;      // (12,48)
};

struct EcSmemW {
    float w0b[EC_C][EC_G];     // [in][out] : W0[:, 24+in]
    float w1a[EC_G][EC_G];     // W1[:, in]
    float w2a[EC_G][EC_G];     // W2[:, in]
    float w2b[EC_G][EC_G];     // W2[:, 12+in]
    float wp[EC_C][36];        // per-point terms: [in][0..11]=W0[:, in], [12..23]=W1[:, 12+in], [24..35]=W2[:, 24+in]
    float bp[36];              // b0, b1, b2
};

template <bool X_IN_SMEM>
__global__ void __launch_bounds__(EC_WARPS * 32) edgeconv_kernel(int n, int k, int pts_per_cta, const float *__restrict__ x,
                                                                 long long x_bstride, const int32_t *__restrict__ idx,
                                                                 int idx_stride, int idx_off, EcWeights W,
                                                                 float *__restrict__ y, long long y_bstride) {
    extern __shared__ __align__(16) unsigned char raw[];
    EcSmemW &sw = *reinterpret_cast<EcSmemW *>(raw);
    float *s_out = reinterpret_cast<float *>(raw + sizeof(EcSmemW));        // [EC_OUT][EC_PT+1]
    float *s_a = s_out + EC_OUT * (EC_PT + 1);                                // [EC_WARPS][36]
    float *xs = s_a + EC_WARPS * 36;                                          // [n][EC_XS] when X_IN_SMEM

    const int bi = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *xb = x + bi * x_bstride;
    float *yb = y + bi * y_bstride;
    const int32_t *ib = idx + (size_t)bi * n * idx_stride;

    // ---- stage weights (transposed) and the cloud ---------------------------------------------------
    for (int t = threadIdx.x; t < EC_C * EC_G; t += blockDim.x) {
        const int in = t / EC_G, o = t % EC_G;
        sw.w0b[in][o] = __ldg(W.w0 + o * 48 + 24 + in);
    }
    for (int t = threadIdx.x; t < EC_G * EC_G; t += blockDim.x) {
        const int in = t / EC_G, o = t % EC_G;
        sw.w1a[in][o] = __ldg(W.w1 + o * 36 + in);
        sw.w2a[in][o] = __ldg(W.w2 + o * 48 + in);
        sw.w2b[in][o] = __ldg(W.w2 + o * 48 + 12 + in);
    }
    for (int t = threadIdx.x; t < EC_C * 36; t += blockDim.x) {
        const int in = t / 36, o = t % 36;
        float v;
        if (o < 12) v = __ldg(W.w0 + o * 48 + in);
        else if (o < 24) v = __ldg(W.w1 + (o - 12) * 36 + 12 + in);
        else v = __ldg(W.w2 + (o - 24) * 48 + 24 + in);
        sw.wp[in][o] = v;
    }
    if (threadIdx.x < 36) {
        const int o = threadIdx.x;
        sw.bp[o] = o < 12 ? __ldg(W.b0 + o) : (o < 24 ? __ldg(W.b1 + o - 12) : __ldg(W.b2 + o - 24));
    }
    if (X_IN_SMEM) {
        for (int t = threadIdx.x; t < EC_C * n; t += blockDim.x) {
            const int c = t / n, p = t - c * n;  // coalesced global read, transposing store
            xs[p * EC_XS + c] = __ldg(xb + (size_t)c * n + p);
        }
    }
    __syncthreads();

    const int p_begin = blockIdx.x * pts_per_cta;
    const int p_end = min(n, p_begin + pts_per_cta);
    float *wa = s_a + warp * 36;

    for (int t0 = p_begin; t0 < p_end; t0 += EC_PT) {
        const int tcnt = min(EC_PT, p_end - t0);
        for (int lp = warp; lp < tcnt; lp += EC_WARPS) {
            const int i = t0 + lp;
            // ---- centre features + per-point terms A0|A1|A2 -------------------------------------------
            float c[EC_C];
#pragma unroll
            for (int ch = 0; ch < EC_C; ++ch) c[ch] = X_IN_SMEM ? xs[i * EC_XS + ch] : __ldg(xb + (size_t)ch * n + i);
            {
                float a0 = sw.bp[lane], a1 = lane < 4 ? sw.bp[32 + lane] : 0.f;
#pragma unroll
                for (int ch = 0; ch < EC_C; ++ch) {
                    a0 = __fmaf_rn(sw.wp[ch][lane], c[ch], a0);
                    if (lane < 4) a1 = __fmaf_rn(sw.wp[ch][32 + lane], c[ch], a1);
                }
                __syncwarp();
                wa[lane] = a0;
                if (lane < 4) wa[32 + lane] = a1;
                __syncwarp();
            }
            float m0[EC_G], m1[EC_G], m2[EC_G];
#pragma unroll
            for (int o = 0; o < EC_G; ++o) m0[o] = m1[o] = m2[o] = -INFINITY;

            for (int e0 = 0; e0 < k; e0 += 32) {
                const int e = e0 + lane;
                const bool live = e < k;
                const int j = live ? __ldg(ib + (size_t)i * idx_stride + idx_off + e) : i;
                float h0[EC_G], h1[EC_G], h2[EC_G];
#pragma unroll
                for (int o = 0; o < EC_G; ++o) h0[o] = wa[o];
#pragma unroll
                for (int ch = 0; ch < EC_C; ++ch) {
                    const float nb = X_IN_SMEM ? xs[j * EC_XS + ch] : __ldg(xb + (size_t)ch * n + j);
                    const float d = nb - c[ch];                       // edge feature n - c (layers.py:41)
                    const float4 wA = *reinterpret_cast<const float4 *>(&sw.w0b[ch][0]);
                    const float4 wB = *reinterpret_cast<const float4 *>(&sw.w0b[ch][4]);
                    const float4 wC = *reinterpret_cast<const float4 *>(&sw.w0b[ch][8]);
                    h0[0] = __fmaf_rn(wA.x, d, h0[0]); h0[1] = __fmaf_rn(wA.y, d, h0[1]);
                    h0[2] = __fmaf_rn(wA.z, d, h0[2]); h0[3] = __fmaf_rn(wA.w, d, h0[3]);
                    h0[4] = __fmaf_rn(wB.x, d, h0[4]); h0[5] = __fmaf_rn(wB.y, d, h0[5]);
                    h0[6] = __fmaf_rn(wB.z, d, h0[6]); h0[7] = __fmaf_rn(wB.w, d, h0[7]);
                    h0[8] = __fmaf_rn(wC.x, d, h0[8]); h0[9] = __fmaf_rn(wC.y, d, h0[9]);
                    h0[10] = __fmaf_rn(wC.z, d, h0[10]); h0[11] = __fmaf_rn(wC.w, d, h0[11]);
                }
#pragma unroll
                for (int o = 0; o < EC_G; ++o) { h0[o] = fmaxf(h0[o], 0.f); h1[o] = wa[12 + o]; h2[o] = wa[24 + o]; }
#pragma unroll
                for (int in = 0; in < EC_G; ++in) {
                    const float v = h0[in];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const float4 w1 = *reinterpret_cast<const float4 *>(&sw.w1a[in][q * 4]);
                        const float4 w2 = *reinterpret_cast<const float4 *>(&sw.w2b[in][q * 4]);
                        h1[q * 4 + 0] = __fmaf_rn(w1.x, v, h1[q * 4 + 0]); h1[q * 4 + 1] = __fmaf_rn(w1.y, v, h1[q * 4 + 1]);
                        h1[q * 4 + 2] = __fmaf_rn(w1.z, v, h1[q * 4 + 2]); h1[q * 4 + 3] = __fmaf_rn(w1.w, v, h1[q * 4 + 3]);
                        h2[q * 4 + 0] = __fmaf_rn(w2.x, v, h2[q * 4 + 0]); h2[q * 4 + 1] = __fmaf_rn(w2.y, v, h2[q * 4 + 1]);
                        h2[q * 4 + 2] = __fmaf_rn(w2.z, v, h2[q * 4 + 2]); h2[q * 4 + 3] = __fmaf_rn(w2.w, v, h2[q * 4 + 3]);
                    }
                }
#pragma unroll
                for (int o = 0; o < EC_G; ++o) h1[o] = fmaxf(h1[o], 0.f);
#pragma unroll
                for (int in = 0; in < EC_G; ++in) {
                    const float v = h1[in];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const float4 w2 = *reinterpret_cast<const float4 *>(&sw.w2a[in][q * 4]);
                        h2[q * 4 + 0] = __fmaf_rn(w2.x, v, h2[q * 4 + 0]); h2[q * 4 + 1] = __fmaf_rn(w2.y, v, h2[q * 4 + 1]);
                        h2[q * 4 + 2] = __fmaf_rn(w2.z, v, h2[q * 4 + 2]); h2[q * 4 + 3] = __fmaf_rn(w2.w, v, h2[q * 4 + 3]);
                    }
                }
                if (live) {
#pragma unroll
                    for (int o = 0; o < EC_G; ++o) {
                        m0[o] = fmaxf(m0[o], h0[o]); m1[o] = fmaxf(m1[o], h1[o]); m2[o] = fmaxf(m2[o], h2[o]);
                    }
                }
            }
            // ---- max over the edges (lanes) ----------------------------------------------------------
#pragma unroll
            for (int o = 0; o < EC_G; ++o) {
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    m0[o] = fmaxf(m0[o], __shfl_xor_sync(0xffffffffu, m0[o], s));
                    m1[o] = fmaxf(m1[o], __shfl_xor_sync(0xffffffffu, m1[o], s));
                    m2[o] = fmaxf(m2[o], __shfl_xor_sync(0xffffffffu, m2[o], s));
                }
            }
            // output channel order [h2, h1, h0, x] (layers.py:57-61 put the new features first)
            if (lane == 0) {
#pragma unroll
                for (int o = 0; o < EC_G; ++o) {
                    s_out[(o) * (EC_PT + 1) + lp] = m2[o];
                    s_out[(12 + o) * (EC_PT + 1) + lp] = m1[o];
                    s_out[(24 + o) * (EC_PT + 1) + lp] = m0[o];
                }
            }
            if (lane < EC_C) s_out[(36 + lane) * (EC_PT + 1) + lp] = X_IN_SMEM ? xs[i * EC_XS + lane] : __ldg(xb + (size_t)lane * n + i);
        }
        __syncthreads();
        for (int t = threadIdx.x; t < EC_OUT * EC_PT; t += blockDim.x) {
            const int ch = t / EC_PT, lp = t % EC_PT;
            if (lp < tcnt) yb[(size_t)ch * n + t0 + lp] = s_out[ch * (EC_PT + 1) + lp];
        }
        __syncthreads();
    }
}

}  // namespace pu3

using namespace pu3;

extern "C" int pu3_edgeconv_f32(int b, int n, int k, const float *x, long long x_bstride, const int32_t *idx,
                                int idx_stride, int idx_off, const float *w0, const float *b0, const float *w1,
                                const float *b1, const float *w2, const float *b2, float *y, long long y_bstride,
                                pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && n >= 0 && k > 0, "edgeconv: bad size b=%d n=%d k=%d", b, n, k);
    if (b == 0 || n == 0) return PU3_OK;
    PU3_ARG_CHECK(b <= 65535, "edgeconv: b=%d exceeds 65535", b);
    PU3_ARG_CHECK(x && idx && w0 && b0 && w1 && b1 && w2 && b2 && y, "edgeconv: null pointer");
    PU3_ARG_CHECK(idx_stride >= idx_off + k && idx_off >= 0, "edgeconv: idx_stride=%d too small for idx_off=%d + k=%d", idx_stride, idx_off, k);
    EcWeights W{w0, b0, w1, b1, w2, b2};
    const size_t fixed = sizeof(EcSmemW) + (size_t)(EC_OUT * (EC_PT + 1) + EC_WARPS * 36) * sizeof(float);
    const size_t with_x = fixed + (size_t)n * EC_XS * sizeof(float);
    const bool in_smem = with_x <= (size_t)device_info().smem_optin && with_x <= 100 * 1024;  // keep >= 2 CTAs per SM
    const size_t smem = in_smem ? with_x : fixed;
    // points per CTA: cover the chip at least twice, in multiples of the staging tile
    const int sms = device_info().sm_count;
    int pts = n;
    if ((long long)b < 2LL * sms) {
        const int split = (int)((2LL * sms + b - 1) / b);
        pts = (n + split - 1) / split;
        pts = ((pts + EC_PT - 1) / EC_PT) * EC_PT;
    }
    dim3 grid((n + pts - 1) / pts, b);
    cudaStream_t s = as_stream(stream);
    int st;
    if (in_smem) {
        st = cuda_status(cudaFuncSetAttribute(edgeconv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "edgeconv: smem attr");
        if (st) return st;
        edgeconv_kernel<true><<<grid, EC_WARPS * 32, smem, s>>>(n, k, pts, x, x_bstride, idx, idx_stride, idx_off, W, y, y_bstride);
    } else {
        st = cuda_status(cudaFuncSetAttribute(edgeconv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "edgeconv: smem attr");
        if (st) return st;
        edgeconv_kernel<false><<<grid, EC_WARPS * 32, smem, s>>>(n, k, pts, x, x_bstride, idx, idx_stride, idx_off, W, y, y_bstride);
    }
    PU3_LAUNCH_CHECK("edgeconv_kernel");
    return PU3_OK;
}
