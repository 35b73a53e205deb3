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


// ------------------------------------------------------------------------------------------------------
// Fast kernel (k <= 32, the cloud fits shared memory).  Differences to edgeconv_kernel above, all driven by the
// round-1 profile (fp32-issue bound, FMA pipe 20 % busy, 1 CTA/SM, 21 M shared-memory bank conflicts):
//   * 2 edges per lane (a half-warp of 16 lanes owns one point): every broadcast LDS.128 of 4 weights feeds
//     8 FMAs, issued as 4 FFMA2 -> the FMA pipe, not instruction issue, is the limiter
//   * neighbour rows are fetched as 6 LDS.128 from a 28-float row stride (conflict-free quarter-warp phases)
//     instead of 24 scalar LDS
//   * the max over the 32 edges goes through shared memory (36 STS + ~60 LDS/FMNMX per warp) instead of a
//     5-step shuffle butterfly on 36 values
//   * <= 128 registers: 2 CTAs per SM
// ------------------------------------------------------------------------------------------------------
constexpr int EF_XS = 24;            // row stride of a point's features in shared memory (only centre rows are read: broadcasts)
constexpr int EB_XS = 28;             // backward kernel: row stride of a point (7 quads, odd -> neighbour rows on distinct banks)
constexpr int EF_PS0 = 12;           // row stride of the per-point layer-0 products W0[:,24:] . x_j gathered per edge
constexpr int EF_RS = 20;            // row stride of the reduction scratch (16 lanes + pad)
constexpr int EF_PS = 36 * EF_RS;    // per-point scratch; 720 % 32 == 16 keeps the two half-warps on disjoint banks

struct EfSmemW {
    float w0b[EC_C][EC_G];
    float w1a[EC_G][EC_G];
    float w2a[EC_G][EC_G];
    float w2b[EC_G][EC_G];
    float wp[EC_C][36];
    float bp[36];
};

__device__ __forceinline__ void ef_layer(f32x2 (&h)[6], const float *wrow, float x) {
    const float4 a = *reinterpret_cast<const float4 *>(wrow);
    const float4 b = *reinterpret_cast<const float4 *>(wrow + 4);
    const float4 c = *reinterpret_cast<const float4 *>(wrow + 8);
    const f32x2 xx = pack2(x, x);
    h[0] = fma2(pack2(a.x, a.y), xx, h[0]); h[1] = fma2(pack2(a.z, a.w), xx, h[1]);
    h[2] = fma2(pack2(b.x, b.y), xx, h[2]); h[3] = fma2(pack2(b.z, b.w), xx, h[3]);
    h[4] = fma2(pack2(c.x, c.y), xx, h[4]); h[5] = fma2(pack2(c.z, c.w), xx, h[5]);
}
// two edges share the weight loads
__device__ __forceinline__ void ef_layer2(f32x2 (&ha)[6], f32x2 (&hb)[6], const float *wrow, float xa, float xb) {
    const float4 a = *reinterpret_cast<const float4 *>(wrow);
    const float4 b = *reinterpret_cast<const float4 *>(wrow + 4);
    const float4 c = *reinterpret_cast<const float4 *>(wrow + 8);
    const f32x2 w0 = pack2(a.x, a.y), w1 = pack2(a.z, a.w), w2 = pack2(b.x, b.y), w3 = pack2(b.z, b.w),
                w4 = pack2(c.x, c.y), w5 = pack2(c.z, c.w);
    const f32x2 xxa = pack2(xa, xa), xxb = pack2(xb, xb);
    ha[0] = fma2(w0, xxa, ha[0]); ha[1] = fma2(w1, xxa, ha[1]); ha[2] = fma2(w2, xxa, ha[2]);
    ha[3] = fma2(w3, xxa, ha[3]); ha[4] = fma2(w4, xxa, ha[4]); ha[5] = fma2(w5, xxa, ha[5]);
    hb[0] = fma2(w0, xxb, hb[0]); hb[1] = fma2(w1, xxb, hb[1]); hb[2] = fma2(w2, xxb, hb[2]);
    hb[3] = fma2(w3, xxb, hb[3]); hb[4] = fma2(w4, xxb, hb[4]); hb[5] = fma2(w5, xxb, hb[5]);
}

// FULLK: k == 32 (the reference configuration): every lane's two edges exist, the validity selects disappear
template <bool FULLK>
__global__ void __launch_bounds__(EC_WARPS * 32, 2) edgeconv_fast_kernel(int n, int k, int pts_per_cta,
                                                                         const float *__restrict__ x, long long x_bstride,
                                                                         const int32_t *__restrict__ idx, int idx_stride,
                                                                         int idx_off, EcWeights W, float *__restrict__ y,
                                                                         long long y_bstride) {
    extern __shared__ __align__(16) unsigned char raw[];
    EfSmemW &sw = *reinterpret_cast<EfSmemW *>(raw);
    float *s_out = reinterpret_cast<float *>(raw + sizeof(EfSmemW));   // [EC_OUT][EC_PT+1]
    float *s_a = s_out + EC_OUT * (EC_PT + 1);                           // [EC_WARPS][2][36]
    float *s_red = s_a + EC_WARPS * 2 * 36;                              // [EC_WARPS][2][EF_PS]
    float *xs = s_red + EC_WARPS * 2 * EF_PS;                            // [n][EF_XS]
    float *ps = xs + (size_t)n * EF_XS;                                  // [n][EF_PS0]: P_j = W0[:, 24:] . x_j

    const int bi = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hl = lane & 15, hp = lane >> 4;
    const float *xb = x + bi * x_bstride;
    float *yb = y + bi * y_bstride;
    const int32_t *ib = idx + (size_t)bi * n * idx_stride;

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
    for (int t = threadIdx.x; t < EC_C * n; t += blockDim.x) {
        const int c = t / n, p = t - c * n;
        xs[p * EF_XS + c] = __ldg(xb + (size_t)c * n + p);
    }
    __syncthreads();
    // The neighbour enters the block only through layer 0, and there linearly: W0 [c, n - c] + b0 =
    // (W0a - W0b) c + b0 + W0b n.  W0b x_j is a per-POINT quantity, so it is computed once per point of the cloud
    // (288 MAC) instead of once per edge (288 MAC x 32 edges); an edge then costs one 48-byte gather and 12 adds
    // for layer 0, and 432 MAC for layers 1 and 2.
    for (int t = threadIdx.x; t < n * 3; t += blockDim.x) {
        const int p = t / 3, q = t - p * 3;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *row = xs + p * EF_XS;
#pragma unroll
        for (int ch = 0; ch < EC_C; ++ch) {
            const float4 w = *reinterpret_cast<const float4 *>(&sw.w0b[ch][q * 4]);
            const float v = row[ch];
            acc.x = __fmaf_rn(w.x, v, acc.x); acc.y = __fmaf_rn(w.y, v, acc.y);
            acc.z = __fmaf_rn(w.z, v, acc.z); acc.w = __fmaf_rn(w.w, v, acc.w);
        }
        *reinterpret_cast<float4 *>(ps + p * EF_PS0 + q * 4) = acc;
    }
    __syncthreads();

    const int p_begin = blockIdx.x * pts_per_cta;
    const int p_end = min(n, p_begin + pts_per_cta);
    float *wa = s_a + (warp * 2 + hp) * 36;
    float *wred = s_red + (size_t)(warp * 2 + hp) * EF_PS;

    // the lane's two edges are fixed; their neighbour indices are the only global loads of the main loop and sit at the
    // head of its dependency chain (ncu: long-scoreboard was the top stall), so they are fetched one point ahead
    const int ea = hl, eb = 16 + hl;
    const bool va = FULLK || ea < k, vb = FULLK || eb < k;
    auto fetch_idx = [&](int t0x, int lp0x, int &ja_out, int &jb_out) {
        const int tc = min(EC_PT, p_end - t0x);
        const int ix = (lp0x + hp < tc) ? t0x + lp0x + hp : t0x + lp0x;
        ja_out = va ? __ldg(ib + (size_t)ix * idx_stride + idx_off + ea) : ix;
        jb_out = vb ? __ldg(ib + (size_t)ix * idx_stride + idx_off + eb) : ix;
    };
    int ja_next = 0, jb_next = 0;
    if (p_begin < p_end && warp * 2 < min(EC_PT, p_end - p_begin)) fetch_idx(p_begin, warp * 2, ja_next, jb_next);

    for (int t0 = p_begin; t0 < p_end; t0 += EC_PT) {
        const int tcnt = min(EC_PT, p_end - t0);
        for (int lp0 = warp * 2; lp0 < tcnt; lp0 += EC_WARPS * 2) {
            const int lp = lp0 + hp;
            const bool pvalid = lp < tcnt;                 // the second point of the pair may not exist
            const int i = pvalid ? t0 + lp : t0 + lp0;
            const float *ci = xs + i * EF_XS;
            const int ja = ja_next, jb = jb_next;
            {   // next work item of this warp: same tile, or the first pair of the next tile
                int n_t0 = t0, n_lp0 = lp0 + EC_WARPS * 2;
                if (n_lp0 >= tcnt) { n_t0 = t0 + EC_PT; n_lp0 = warp * 2; }
                if (n_t0 < p_end && n_lp0 < min(EC_PT, p_end - n_t0)) fetch_idx(n_t0, n_lp0, ja_next, jb_next);
            }
            // ---- per-point terms A0|A1|A2 (36 values) by the 16 lanes of the half-warp -------------------
            {
                float a0 = sw.bp[hl], a1 = sw.bp[16 + hl], a2 = hl < 4 ? sw.bp[32 + hl] : 0.f;
#pragma unroll
                for (int ch = 0; ch < EC_C; ++ch) {
                    const float cv = ci[ch];
                    a0 = __fmaf_rn(sw.wp[ch][hl], cv, a0);
                    a1 = __fmaf_rn(sw.wp[ch][16 + hl], cv, a1);
                    if (hl < 4) a2 = __fmaf_rn(sw.wp[ch][32 + hl], cv, a2);
                }
                if (hl < EC_G) a0 -= ps[i * EF_PS0 + hl];   // (W0a - W0b) c + b0: the centre's own W0b c goes out here
                __syncwarp();
                wa[hl] = a0; wa[16 + hl] = a1;
                if (hl < 4) wa[32 + hl] = a2;
                __syncwarp();
            }
            // ---- the lane's two edges -------------------------------------------------------------------------
            const float *pa = ps + ja * EF_PS0, *pb = ps + jb * EF_PS0;
            float r0a[EC_G], r0b[EC_G];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float4 fa = *reinterpret_cast<const float4 *>(pa + 4 * q);
                const float4 fb = *reinterpret_cast<const float4 *>(pb + 4 * q);
                r0a[4 * q + 0] = fmaxf(wa[4 * q + 0] + fa.x, 0.f); r0b[4 * q + 0] = fmaxf(wa[4 * q + 0] + fb.x, 0.f);
                r0a[4 * q + 1] = fmaxf(wa[4 * q + 1] + fa.y, 0.f); r0b[4 * q + 1] = fmaxf(wa[4 * q + 1] + fb.y, 0.f);
                r0a[4 * q + 2] = fmaxf(wa[4 * q + 2] + fa.z, 0.f); r0b[4 * q + 2] = fmaxf(wa[4 * q + 2] + fb.z, 0.f);
                r0a[4 * q + 3] = fmaxf(wa[4 * q + 3] + fa.w, 0.f); r0b[4 * q + 3] = fmaxf(wa[4 * q + 3] + fb.w, 0.f);
            }
            f32x2 h1a[6], h1b[6], h2a[6], h2b[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                h1a[q] = h1b[q] = pack2(wa[12 + 2 * q], wa[12 + 2 * q + 1]);
                h2a[q] = h2b[q] = pack2(wa[24 + 2 * q], wa[24 + 2 * q + 1]);
            }
#pragma unroll
            for (int in = 0; in < EC_G; ++in) {
                ef_layer2(h1a, h1b, &sw.w1a[in][0], r0a[in], r0b[in]);
                ef_layer2(h2a, h2b, &sw.w2b[in][0], r0a[in], r0b[in]);
            }
            float m[36];   // max over the lane's two edges: [h2 | h1 | h0]
#pragma unroll
            for (int o = 0; o < EC_G; ++o) {
                const float xa = va ? r0a[o] : -INFINITY, xb2 = vb ? r0b[o] : -INFINITY;
                m[24 + o] = fmaxf(xa, xb2);
            }
            {
                float r1a[EC_G], r1b[EC_G];
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    unpack2(h1a[q], r1a[2 * q], r1a[2 * q + 1]);
                    unpack2(h1b[q], r1b[2 * q], r1b[2 * q + 1]);
                }
#pragma unroll
                for (int o = 0; o < EC_G; ++o) { r1a[o] = fmaxf(r1a[o], 0.f); r1b[o] = fmaxf(r1b[o], 0.f); }
#pragma unroll
                for (int in = 0; in < EC_G; ++in) ef_layer2(h2a, h2b, &sw.w2a[in][0], r1a[in], r1b[in]);
#pragma unroll
                for (int o = 0; o < EC_G; ++o) m[12 + o] = fmaxf(va ? r1a[o] : -INFINITY, vb ? r1b[o] : -INFINITY);
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                float a0, a1, b0, b1;
                unpack2(h2a[q], a0, a1);
                unpack2(h2b[q], b0, b1);
                m[2 * q] = fmaxf(va ? a0 : -INFINITY, vb ? b0 : -INFINITY);
                m[2 * q + 1] = fmaxf(va ? a1 : -INFINITY, vb ? b1 : -INFINITY);
            }
            // ---- max over the 16 lanes of the point through shared memory -------------------------------------------
            __syncwarp();
#pragma unroll
            for (int o = 0; o < 36; ++o) wred[o * EF_RS + hl] = m[o];
            __syncwarp();
            const float *wbase = s_red + (size_t)(warp * 2) * EF_PS;
            for (int o = lane; o < 72; o += 32) {
                const int pp = o / 36, ch = o - pp * 36;
                const float *rrow = wbase + (size_t)pp * EF_PS + ch * EF_RS;
                const float4 v0 = *reinterpret_cast<const float4 *>(rrow), v1 = *reinterpret_cast<const float4 *>(rrow + 4);
                const float4 v2 = *reinterpret_cast<const float4 *>(rrow + 8), v3 = *reinterpret_cast<const float4 *>(rrow + 12);
                const float mx = fmaxf(fmaxf(fmaxf(fmaxf(v0.x, v0.y), fmaxf(v0.z, v0.w)), fmaxf(fmaxf(v1.x, v1.y), fmaxf(v1.z, v1.w))),
                                       fmaxf(fmaxf(fmaxf(v2.x, v2.y), fmaxf(v2.z, v2.w)), fmaxf(fmaxf(v3.x, v3.y), fmaxf(v3.z, v3.w))));
                if (lp0 + pp < tcnt) s_out[ch * (EC_PT + 1) + lp0 + pp] = mx;   // rows 0..35 = [h2, h1, h0] already in output order
            }
            if (pvalid) {
                for (int ch = hl; ch < EC_C; ch += 16) s_out[(36 + ch) * (EC_PT + 1) + lp] = ci[ch];
            }
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

// Test hook: force the generic kernel.
static int g_ec_force_generic = 0;
extern "C" void pu3_edgeconv_force_generic(int on) { g_ec_force_generic = on; }
// A/B hook: k == 32 runs the per-edge layers on the tensor cores -- 2 = every operand in tensor memory (edgeconv_ts.cu), 1 = layer-1
// operand images in shared memory (edgeconv_tc.cu), 0 = FFMA kernels only
static int g_ec_tc = 2;
extern "C" void pu3_edgeconv_set_tc(int on) { g_ec_tc = on; }
namespace pu3 {
bool edgeconv_tc_launch(int b, int n, const float *x, long long x_bstride, const int32_t *idx, int idx_stride, int idx_off,
                        const float *w0, const float *b0, const float *w1, const float *b1, const float *w2, const float *b2,
                        float *y, long long y_bstride, cudaStream_t s, int *status);
bool edgeconv_ts_launch(int b, int n, const float *x, long long x_bstride, const int32_t *idx, int idx_stride, int idx_off,
                        const float *w0, const float *b0, const float *w1, const float *b1, const float *w2, const float *b2,
                        float *y, long long y_bstride, cudaStream_t s, int *status);
}

static int edgeconv_forward(bool allow_tc, int b, int n, int k, const float *x, long long x_bstride, const int32_t *idx,
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
    if (k == 32 && allow_tc && g_ec_tc && !g_ec_force_generic) {
        if ((g_ec_tc >= 2 && edgeconv_ts_launch(b, n, x, x_bstride, idx, idx_stride, idx_off, w0, b0, w1, b1, w2, b2, y, y_bstride, s, &st)) ||
            edgeconv_tc_launch(b, n, x, x_bstride, idx, idx_stride, idx_off, w0, b0, w1, b1, w2, b2, y, y_bstride, s, &st)) {
            if (st) return st;
            PU3_LAUNCH_CHECK("edgeconv_tc_kernel");
            return PU3_OK;
        }
    }
    const size_t fast_smem = sizeof(EfSmemW) + (size_t)(EC_OUT * (EC_PT + 1) + EC_WARPS * 2 * 36 + EC_WARPS * 2 * EF_PS + (size_t)n * (EF_XS + EF_PS0)) * sizeof(float);
    if (k <= 32 && fast_smem <= 110 * 1024 && g_ec_force_generic == 0) {
        auto kern = k == 32 ? edgeconv_fast_kernel<true> : edgeconv_fast_kernel<false>;
        st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem), "edgeconv: smem attr");
        if (st) return st;
        kern<<<grid, EC_WARPS * 32, fast_smem, s>>>(n, k, pts, x, x_bstride, idx, idx_stride, idx_off, W, y, y_bstride);
        PU3_LAUNCH_CHECK("edgeconv_fast_kernel");
        return PU3_OK;
    }
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

extern "C" int pu3_edgeconv_f32(int b, int n, int k, const float *x, long long x_bstride, const int32_t *idx,
                                int idx_stride, int idx_off, const float *w0, const float *b0, const float *w1,
                                const float *b1, const float *w2, const float *b2, float *y, long long y_bstride,
                                pu3_stream_t stream) {
    return edgeconv_forward(true, b, n, k, x, x_bstride, idx, idx_stride, idx_off, w0, b0, w1, b1, w2, b2, y, y_bstride, stream);
}
// the FFMA kernels only: the arithmetic pu3_edgeconv_bwd_f32 recomputes, so a train-mode forward returns exactly the function its
// backward differentiates
extern "C" int pu3_edgeconv_ffma_f32(int b, int n, int k, const float *x, long long x_bstride, const int32_t *idx,
                                     int idx_stride, int idx_off, const float *w0, const float *b0, const float *w1,
                                     const float *b1, const float *w2, const float *b2, float *y, long long y_bstride,
                                     pu3_stream_t stream) {
    return edgeconv_forward(false, b, n, k, x, x_bstride, idx, idx_stride, idx_off, w0, b0, w1, b1, w2, b2, y, y_bstride, stream);
}

// ------------------------------------------------------------------------------------------------------
// DenseEdgeConv backward (k <= 32, cloud in shared memory).  What autograd derives for layers.py:44-64 in the
// reference's train step, as one kernel: per point the forward activations of its k edges are recomputed
// (cheaper than storing (B,36,N,K)), the max() routes each output channel's gradient to the FIRST edge that
// attains the maximum (torch.max semantics), the chain rule runs back through the three 1x1 layers, and
//   dx  (B,24,N)  += neighbour / centre contributions (shared-memory accumulation per cloud, one atomic flush)
//   dW*, db*       += outer products, kept in registers per lane across all points of the warp
// One warp per point, one lane per edge.
// ------------------------------------------------------------------------------------------------------
namespace pu3 {

constexpr int EB_ST = 85;    // staged floats per edge (84 used, odd stride: conflict-free rows): d[24] | h0[12] | h1[12] | g0[12] | g1[12] | g2[12]
constexpr int EB_ACC = 900;  // centre-part weight gradients (36 x 24) + biases (36)
constexpr int EB_DXS = 25;   // row stride of the dx accumulator

__device__ __forceinline__ float dot12(const float *row, const float (&g)[EC_G]) {
    const float4 a = *reinterpret_cast<const float4 *>(row), b = *reinterpret_cast<const float4 *>(row + 4),
                 c = *reinterpret_cast<const float4 *>(row + 8);
    float s = a.x * g[0];
    s = __fmaf_rn(a.y, g[1], s); s = __fmaf_rn(a.z, g[2], s); s = __fmaf_rn(a.w, g[3], s);
    s = __fmaf_rn(b.x, g[4], s); s = __fmaf_rn(b.y, g[5], s); s = __fmaf_rn(b.z, g[6], s); s = __fmaf_rn(b.w, g[7], s);
    s = __fmaf_rn(c.x, g[8], s); s = __fmaf_rn(c.y, g[9], s); s = __fmaf_rn(c.z, g[10], s); s = __fmaf_rn(c.w, g[11], s);
    return s;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    return v;
}

struct EcGrads { float *dw0, *db0, *dw1, *db1, *dw2, *db2; };

__global__ void __launch_bounds__(EC_WARPS * 32, 1)
edgeconv_bwd_kernel(int n, int k, int pts_per_cta, const float *__restrict__ x, long long x_bstride,
                    const int32_t *__restrict__ idx, int idx_stride, int idx_off, EcWeights W,
                    const float *__restrict__ dy, long long dy_bstride, float *__restrict__ dx, long long dx_bstride,
                    EcGrads G) {
    extern __shared__ __align__(16) unsigned char raw[];
    EfSmemW &sw = *reinterpret_cast<EfSmemW *>(raw);
    float *s_a = reinterpret_cast<float *>(raw + sizeof(EfSmemW));   // [EC_WARPS][36]  A0|A1|A2 of the warp's point
    float *s_dy = s_a + EC_WARPS * 36;                                 // [EC_WARPS][60]
    float *s_da = s_dy + EC_WARPS * 60;                                // [EC_WARPS][36]  dA0|dA1|dA2
    float *s_st = s_da + EC_WARPS * 36;                                // [EC_WARPS][32][EB_ST]
    float *s_acc = s_st + EC_WARPS * 32 * EB_ST;                       // [EC_WARPS][EB_ACC]
    float *dxs = s_acc + EC_WARPS * EB_ACC;                            // [n][EB_DXS]
    float *xs = dxs + (size_t)n * EB_DXS;                              // [n][EB_XS]

    const int bi = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *xb = x + bi * x_bstride;
    const float *dyb = dy + bi * dy_bstride;
    float *dxb = dx + bi * dx_bstride;
    const int32_t *ib = idx + (size_t)bi * n * idx_stride;

    for (int t = threadIdx.x; t < EC_C * EC_G; t += blockDim.x) { const int in = t / EC_G, o = t % EC_G; sw.w0b[in][o] = __ldg(W.w0 + o * 48 + 24 + in); }
    for (int t = threadIdx.x; t < EC_G * EC_G; t += blockDim.x) {
        const int in = t / EC_G, o = t % EC_G;
        sw.w1a[in][o] = __ldg(W.w1 + o * 36 + in); sw.w2a[in][o] = __ldg(W.w2 + o * 48 + in); sw.w2b[in][o] = __ldg(W.w2 + o * 48 + 12 + in);
    }
    for (int t = threadIdx.x; t < EC_C * 36; t += blockDim.x) {
        const int in = t / 36, o = t % 36;
        sw.wp[in][o] = o < 12 ? __ldg(W.w0 + o * 48 + in) : (o < 24 ? __ldg(W.w1 + (o - 12) * 36 + 12 + in) : __ldg(W.w2 + (o - 24) * 48 + 24 + in));
    }
    if (threadIdx.x < 36) { const int o = threadIdx.x; sw.bp[o] = o < 12 ? __ldg(W.b0 + o) : (o < 24 ? __ldg(W.b1 + o - 12) : __ldg(W.b2 + o - 24)); }
    for (int t = threadIdx.x; t < EC_C * n; t += blockDim.x) { const int c = t / n, p = t - c * n; xs[p * EB_XS + c] = __ldg(xb + (size_t)c * n + p); }
    for (int t = threadIdx.x; t < n * EB_DXS; t += blockDim.x) dxs[t] = 0.f;
    __syncthreads();

    // weight-gradient accumulators live in registers for all points of the warp (profiles/r2: the former per-warp shared-memory
    // rows, updated by single-accumulator loops over the staged edges, left the kernel latency-bound at 8 warps per SM)
    const int og = lane >> 3, cg = lane & 7;      // dW0[:, 24:] tile: o = 3 og + u, c = 3 cg + t
    const int gg = lane >> 2, hg = lane & 3;      // [g1|g2] x [h0|h1] tile: row = 3 gg + u, col = 6 hg + t
    float acc0[3][3], acc1[3][6];
    // centre parts dA (x) x_i (36 x 24) + biases (36): one read-modify-write per entry and point, per-warp rows in shared memory
    float *accC = s_acc + (size_t)warp * EB_ACC;
    for (int e = lane; e < EB_ACC; e += 32) accC[e] = 0.f;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
#pragma unroll
        for (int t = 0; t < 3; ++t) acc0[u][t] = 0.f;
#pragma unroll
        for (int t = 0; t < 6; ++t) acc1[u][t] = 0.f;
    }

    float *wa = s_a + warp * 36, *wdy = s_dy + warp * 60, *wda = s_da + warp * 36, *st = s_st + (size_t)warp * 32 * EB_ST;
    float *me = st + lane * EB_ST;
    const int p_begin = blockIdx.x * pts_per_cta, p_end = min(n, p_begin + pts_per_cta);
    for (int i = p_begin + warp; i < p_end; i += EC_WARPS) {
        const float *ci = xs + i * EB_XS;
        __syncwarp();
        {   // per-point terms and the incoming gradient of the point's 60 output channels
            float a0 = sw.bp[lane], a1 = lane < 4 ? sw.bp[32 + lane] : 0.f;
#pragma unroll
            for (int ch = 0; ch < EC_C; ++ch) {
                const float cv = ci[ch];
                a0 = __fmaf_rn(sw.wp[ch][lane], cv, a0);
                if (lane < 4) a1 = __fmaf_rn(sw.wp[ch][32 + lane], cv, a1);
            }
            wa[lane] = a0;
            if (lane < 4) wa[32 + lane] = a1;
            wdy[lane] = __ldg(dyb + (size_t)lane * n + i);
            if (lane < 28) wdy[32 + lane] = __ldg(dyb + (size_t)(32 + lane) * n + i);
        }
        __syncwarp();
        // ---- forward recompute of this lane's edge; d, h0, h1 go straight to the staging row ------------------------
        const bool live = lane < k;
        const int j = live ? __ldg(ib + (size_t)i * idx_stride + idx_off + lane) : i;
        const float *nj = xs + j * EB_XS;
        float h0[EC_G], h1[EC_G], h2[EC_G];
#pragma unroll
        for (int o = 0; o < EC_G; ++o) h0[o] = wa[o];
#pragma unroll
        for (int c = 0; c < EC_C; ++c) {
            const float dc = nj[c] - ci[c];
            me[c] = live ? dc : 0.f;
#pragma unroll
            for (int o = 0; o < EC_G; ++o) h0[o] = __fmaf_rn(sw.w0b[c][o], dc, h0[o]);
        }
#pragma unroll
        for (int o = 0; o < EC_G; ++o) { h0[o] = fmaxf(h0[o], 0.f); h1[o] = wa[12 + o]; h2[o] = wa[24 + o]; }
#pragma unroll
        for (int in = 0; in < EC_G; ++in)
#pragma unroll
            for (int o = 0; o < EC_G; ++o) { h1[o] = __fmaf_rn(sw.w1a[in][o], h0[in], h1[o]); h2[o] = __fmaf_rn(sw.w2b[in][o], h0[in], h2[o]); }
#pragma unroll
        for (int o = 0; o < EC_G; ++o) h1[o] = fmaxf(h1[o], 0.f);
#pragma unroll
        for (int in = 0; in < EC_G; ++in)
#pragma unroll
            for (int o = 0; o < EC_G; ++o) h2[o] = __fmaf_rn(sw.w2a[in][o], h1[in], h2[o]);
        // (compiler fences between the phases: without them ptxas hoists the weight loads of later phases above the
        // earlier ones, holds them in registers and spills ~3 KB per thread)
        asm volatile("" ::: "memory");
        // ---- max() routing: the first edge attaining the maximum takes the channel's gradient -------------------
        float g2[EC_G], g1[EC_G], g0[EC_G];
        unsigned pos0 = 0, pos1 = 0;      // ReLU masks of h0 / h1
#pragma unroll
        for (int o = 0; o < EC_G; ++o) {
            me[24 + o] = live ? h0[o] : 0.f;
            me[36 + o] = live ? h1[o] : 0.f;
            pos0 |= (h0[o] > 0.f ? 1u : 0u) << o;
            pos1 |= (h1[o] > 0.f ? 1u : 0u) << o;
            float v2 = live ? h2[o] : -INFINITY, v1 = live ? h1[o] : -INFINITY, v0 = live ? h0[o] : -INFINITY;
            float m2 = v2, m1 = v1, m0 = v0;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) {
                m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, s));
                m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, s));
                m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, s));
            }
            const int w2 = __ffs(__ballot_sync(0xffffffffu, v2 == m2)) - 1;
            const int w1 = __ffs(__ballot_sync(0xffffffffu, v1 == m1)) - 1;
            const int w0 = __ffs(__ballot_sync(0xffffffffu, v0 == m0)) - 1;
            g2[o] = lane == w2 ? wdy[o] : 0.f;          // y = [h2 | h1 | h0 | x]
            g1[o] = lane == w1 ? wdy[12 + o] : 0.f;
            g0[o] = lane == w0 ? wdy[24 + o] : 0.f;
        }
        // ---- chain rule --------------------------------------------------------------------------------------------
        asm volatile("" ::: "memory");
#pragma unroll
        for (int in = 0; in < EC_G; ++in) g1[in] = (pos1 >> in) & 1u ? g1[in] + dot12(&sw.w2a[in][0], g2) : 0.f;
        asm volatile("" ::: "memory");
#pragma unroll
        for (int in = 0; in < EC_G; ++in)
            g0[in] = (pos0 >> in) & 1u ? g0[in] + dot12(&sw.w1a[in][0], g1) + dot12(&sw.w2b[in][0], g2) : 0.f;
        asm volatile("" ::: "memory");
#pragma unroll
        for (int o = 0; o < EC_G; ++o) { me[48 + o] = g0[o]; me[60 + o] = g1[o]; me[72 + o] = g2[o]; }
        // neighbour side of d = x_j - x_i: dx_j += W0b^T g0
#pragma unroll 4
        for (int c = 0; c < EC_C; ++c) {
            if (live) atomicAdd(&dxs[j * EB_DXS + c], dot12(&sw.w0b[c][0], g0));
        }
        __syncwarp();
        // dA0 = sum_k g0, dA1 = sum_k g1: column sums of the staged rows (odd row stride: conflict-free);
        // dA2 = the incoming gradient of h2 (every channel's maximum is attained by exactly one edge)
        if (lane < 24) {
            float sacc = 0.f;
#pragma unroll 8
            for (int kk = 0; kk < 32; ++kk) sacc += st[kk * EB_ST + 48 + lane];
            wda[lane] = sacc;
        } else {
            wda[lane] = wdy[lane - 24];                      // lanes 24..31 -> dA2[0..8)
        }
        if (lane < 4) wda[32 + lane] = wdy[8 + lane];       // dA2[8..12)
        __syncwarp();
        // centre: dx_i += dy_x - W0b dA0 (= - sum_k gd, by linearity) + Wp^T dA
        if (lane < EC_C) {
            float v = wdy[36 + lane];
#pragma unroll
            for (int o = 0; o < EC_G; ++o) v = __fmaf_rn(-sw.w0b[lane][o], wda[o], v);
#pragma unroll
            for (int o = 0; o < 36; ++o) v = __fmaf_rn(sw.wp[lane][o], wda[o], v);
            atomicAdd(&dxs[i * EB_DXS + lane], v);
        }
        // weight gradients, register tiles over the 32 staged edges:
        //   dW0[:, 24:]  = g0 (x) d          lane tile 3 (o) x 3 (c)
        //   [g1|g2] (x) [h0|h1]              lane tile 3 x 6 (the g1 (x) h1 quadrant is not a gradient and is dropped at the flush)
#pragma unroll 4
        for (int kk = 0; kk < 32; ++kk) {
            const float *r = st + kk * EB_ST;
            float ga[3], da[3], gb[3], hb[6];
#pragma unroll
            for (int t = 0; t < 3; ++t) { ga[t] = r[48 + 3 * og + t]; da[t] = r[3 * cg + t]; gb[t] = r[60 + 3 * gg + t]; }
#pragma unroll
            for (int t = 0; t < 6; ++t) hb[t] = r[24 + 6 * hg + t];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
#pragma unroll
                for (int t = 0; t < 3; ++t) acc0[u][t] = __fmaf_rn(ga[u], da[t], acc0[u][t]);
#pragma unroll
                for (int t = 0; t < 6; ++t) acc1[u][t] = __fmaf_rn(gb[u], hb[t], acc1[u][t]);
            }
        }
        // centre parts dA (x) x_i (36 x 24, entry e = lane + 32 q) and the biases
#pragma unroll
        for (int q = 0; q < 27; ++q) {
            const int e = lane + 32 * q, o = e / 24, c = e % 24;
            accC[e] = __fmaf_rn(wda[o], ci[c], accC[e]);
        }
        accC[864 + lane] += wda[lane];
        if (lane < 4) accC[896 + lane] += wda[32 + lane];
    }
    __syncwarp();
    // ---- flush the warp's accumulators -----------------------------------------------------------------------------
#pragma unroll
    for (int u = 0; u < 3; ++u) {
#pragma unroll
        for (int t = 0; t < 3; ++t) atomicAdd(G.dw0 + (3 * og + u) * 48 + 24 + 3 * cg + t, acc0[u][t]);
        const int row = 3 * gg + u;
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            const int col = 6 * hg + t;
            if (row < 12) {
                if (col < 12) atomicAdd(G.dw1 + row * 36 + col, acc1[u][t]);                 // dW1[:, :12]   = g1 (x) h0
            } else if (col < 12) {
                atomicAdd(G.dw2 + (row - 12) * 48 + 12 + col, acc1[u][t]);                   // dW2[:, 12:24] = g2 (x) h0
            } else {
                atomicAdd(G.dw2 + (row - 12) * 48 + (col - 12), acc1[u][t]);                 // dW2[:, :12]   = g2 (x) h1
            }
        }
    }
    for (int e = lane; e < 864; e += 32) {
        const int o = e / 24, c = e % 24;
        float *dst = o < 12 ? G.dw0 + o * 48 + c : (o < 24 ? G.dw1 + (o - 12) * 36 + 12 + c : G.dw2 + (o - 24) * 48 + 24 + c);
        atomicAdd(dst, accC[e]);
    }
    for (int o = lane; o < 36; o += 32) atomicAdd(o < 12 ? G.db0 + o : (o < 24 ? G.db1 + o - 12 : G.db2 + o - 24), accC[864 + o]);
    __syncthreads();
    for (int t = threadIdx.x; t < EC_C * n; t += blockDim.x) {
        const int c = t / n, p = t - c * n;
        const float v = dxs[p * EB_DXS + c];
        if (v != 0.f) atomicAdd(dxb + (size_t)c * n + p, v);
    }
}

}  // namespace pu3

extern "C" int pu3_edgeconv_bwd_f32(int b, int n, int k, const float *x, long long x_bstride, const int32_t *idx,
                                    int idx_stride, int idx_off, const float *w0, const float *b0, const float *w1,
                                    const float *b1, const float *w2, const float *b2, const float *dy,
                                    long long dy_bstride, float *dx, long long dx_bstride, float *dw0, float *db0,
                                    float *dw1, float *db1, float *dw2, float *db2, pu3_stream_t stream) {
    using namespace pu3;
    PU3_ARG_CHECK(b >= 0 && n >= 0 && k > 0, "edgeconv_bwd: bad size b=%d n=%d k=%d", b, n, k);
    if (b == 0 || n == 0) return PU3_OK;
    PU3_ARG_CHECK(b <= 65535 && k <= 32, "edgeconv_bwd: b=%d (max 65535), k=%d (max 32)", b, k);
    PU3_ARG_CHECK(x && idx && w0 && b0 && w1 && b1 && w2 && b2 && dy && dx && dw0 && db0 && dw1 && db1 && dw2 && db2,
                  "edgeconv_bwd: null pointer");
    PU3_ARG_CHECK(idx_stride >= idx_off + k && idx_off >= 0, "edgeconv_bwd: idx_stride too small");
    const size_t smem = sizeof(EfSmemW) + (size_t)(EC_WARPS * (36 + 60 + 36) + EC_WARPS * 32 * EB_ST + EC_WARPS * EB_ACC + (size_t)n * (EB_DXS + EB_XS)) * sizeof(float);
    if (smem > (size_t)device_info().smem_optin) {
        set_error("edgeconv_bwd: n=%d does not fit shared memory (%zu bytes)", n, smem);
        return PU3_E_UNSUPPORTED;
    }
    const int sms = device_info().sm_count;
    int pts = n;
    if (b < sms) {
        // one resident CTA per SM (registers + shared memory): the grid must fit ONE wave.  ncu (profiles/r2): ceil(148/32) = 5
        // splits gave 160 CTAs, i.e. a second wave of 12 that doubled the kernel's duration (SMs active 63 % of the time).
        const int split = sms / b;
        pts = (n + split - 1) / split;
        pts = ((pts + EC_WARPS - 1) / EC_WARPS) * EC_WARPS;
    }
    EcWeights W{w0, b0, w1, b1, w2, b2};
    EcGrads G{dw0, db0, dw1, db1, dw2, db2};
    int st = cuda_status(cudaFuncSetAttribute(edgeconv_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "edgeconv_bwd: smem attr");
    if (st) return st;
    edgeconv_bwd_kernel<<<dim3((n + pts - 1) / pts, b), EC_WARPS * 32, smem, as_stream(stream)>>>(
        n, k, pts, x, x_bstride, idx, idx_stride, idx_off, W, dy, dy_bstride, dx, dx_bstride, G);
    PU3_LAUNCH_CHECK("edgeconv_bwd_kernel");
    return PU3_OK;
}
