// 1x1 convolution over points (per-point MLP layer), fp32, sm_100a.
//
// Replaces the nn.Conv1d / nn.Conv2d 1x1 layers of the reference (network/layers.py:115-204 used at
// network/upsampler.py:209-230,288-369), which run as cuDNN/cuBLAS library calls plus separate bias /
// ReLU / cat kernels.  Here one kernel computes
//     Y[b, co, p] = act( sum_ci W[co,ci] * X[b, ci, p] + bias[co] )  (+ residual R[b, co, p / rdiv])
// reading X from and writing Y into CHANNEL SLICES of larger (B, Ctot, N) buffers (pointer + batch
// stride), so the dense concatenations of Level.forward (torch.cat at upsampler.py:293,299,305,311)
// cost nothing: every layer writes its output where the next layer expects it.
//
// fp32 FFMA on purpose: the tolerance of the path is 1e-5 relative, which rules out TF32/BF16 tensor
// core inputs (hard parts).  It is a register-tiled SGEMM: the CTA tile is
// TM output channels x TN points, each thread owns 8x8 (or RM x 8) accumulators, K (= Cin) is streamed
// through shared memory in chunks of KC.  Columns are the flattened (batch, point) index, so tiles
// are full even when N (312) is not a multiple of the tile width.
#include "pu3_common.cuh"

namespace pu3 {

constexpr int PW_KC = 16;

struct PwArgs {
    int b, n, cin, cout;
    const float *x; long long x_bstride;   // X[b] = x + b*x_bstride, channel stride n
    const float *w;                        // (cout, cin) row-major
    const float *bias;                     // (cout) or null
    float *y; long long y_bstride;
    const float *res; long long res_bstride; int res_n, res_div;  // optional residual (b, cout, res_n), column p / res_div
    int relu;
    // backward use: Y = [Y +] (mask > 0 ? v : 0).  mask (b, cout, n) laid out like y (its own batch stride): the ReLU
    // derivative of the activation this gradient flows into; accumulate: add to what y already holds (a gradient slice
    // that several consumers contribute to).
    const float *mask; long long mask_bstride;
    int accumulate;
};

// TM = RM * TY output channels, TN = 8 * TX columns, threads = TX * TY
template <int RM, int TY, int TX>
__global__ void __launch_bounds__(TX * TY) pointwise_conv_kernel(PwArgs a) {
    constexpr int TM = RM * TY, TN = 8 * TX, NT = TX * TY;
    __shared__ __align__(16) float Ws[PW_KC][TM + 4];
    __shared__ __align__(16) float Xs[PW_KC][TN];
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const long long cols = (long long)a.b * a.n;
    const long long col0 = (long long)blockIdx.x * TN;
    const int co0 = blockIdx.y * TM;

    // column bookkeeping for the loads: thread loads column (tid % TN) for rows tid / TN, ...
    float acc[RM][8];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.cin; k0 += PW_KC) {
        // ---- stage W chunk (transposed: Ws[k][co]) and X chunk ---------------------------------
        for (int t = threadIdx.x; t < PW_KC * TM; t += NT) {
            const int k = t % PW_KC, co = t / PW_KC;  // consecutive threads walk k: contiguous in W rows
            const int gk = k0 + k, gco = co0 + co;
            Ws[k][co] = (gk < a.cin && gco < a.cout) ? __ldg(a.w + (size_t)gco * a.cin + gk) : 0.f;
        }
        for (int t = threadIdx.x; t < PW_KC * TN; t += NT) {
            const int c = t % TN, k = t / TN;
            const long long col = col0 + c;
            const int gk = k0 + k;
            float v = 0.f;
            if (col < cols && gk < a.cin) {
                const long long bi = col / a.n;
                const int p = (int)(col - bi * a.n);
                v = __ldg(a.x + bi * a.x_bstride + (size_t)gk * a.n + p);
            }
            Xs[k][c] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PW_KC; ++k) {
            float wv[RM], xv[8];
#pragma unroll
            for (int i = 0; i < RM; i += 4) {
                if (RM - i >= 4) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(&Ws[k][ty * RM + i]);
                    wv[i] = w4.x; wv[i + 1] = w4.y; wv[i + 2] = w4.z; wv[i + 3] = w4.w;
                } else {
#pragma unroll
                    for (int r = i; r < RM; ++r) wv[r] = Ws[k][ty * RM + r];
                }
            }
            const float4 x0 = *reinterpret_cast<const float4 *>(&Xs[k][tx * 8]);
            const float4 x1 = *reinterpret_cast<const float4 *>(&Xs[k][tx * 8 + 4]);
            xv[0] = x0.x; xv[1] = x0.y; xv[2] = x0.z; xv[3] = x0.w;
            xv[4] = x1.x; xv[5] = x1.y; xv[6] = x1.z; xv[7] = x1.w;
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(wv[i], xv[j], acc[i][j]);
        }
        __syncthreads();
    }
    // ---- epilogue: bias, residual, ReLU, store ---------------------------------------------------
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int co = co0 + ty * RM + i;
        if (co >= a.cout) continue;
        const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long long col = col0 + tx * 8 + j;
            if (col >= cols) continue;
            const long long bi = col / a.n;
            const int p = (int)(col - bi * a.n);
            float v = acc[i][j] + bv;
            if (a.relu) v = fmaxf(v, 0.f);
            if (a.mask && !(__ldg(a.mask + bi * a.mask_bstride + (size_t)co * a.n + p) > 0.f)) v = 0.f;
            if (a.res) v += __ldg(a.res + bi * a.res_bstride + (size_t)co * a.res_n + p / a.res_div);
            float *dst = a.y + bi * a.y_bstride + (size_t)co * a.n + p;
            if (a.accumulate) v += *dst;
            *dst = v;
        }
    }
}


// ------------------------------------------------------------------------------------------------------
// Fast path: n % 4 == 0, 16-byte aligned slices.  Double-buffered shared memory (one barrier per K chunk),
// global->register prefetch of the next chunk under the FMAs of the current one, conflict-free 128-bit
// fragment loads (a thread's 8 columns are two groups of 4, TN/2 apart), FFMA2 inner product.
// ------------------------------------------------------------------------------------------------------
template <int RM, int TY, int TX>
__global__ void __launch_bounds__(TX * TY) pointwise_conv_fast_kernel(PwArgs a) {
    constexpr int TM = RM * TY, TN = 8 * TX, NT = TX * TY, KC = PW_KC;
    constexpr int WV = (TM * KC / 4 + NT - 1) / NT;      // float4 of W per thread per chunk
    constexpr int XV = (KC * TN / 4 + NT - 1) / NT;      // float4 of X per thread per chunk
    __shared__ __align__(16) float Ws[2][KC][TM + 4];
    __shared__ __align__(16) float Xs[2][KC][TN];
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const long long cols = (long long)a.b * a.n;
    const long long col0 = (long long)blockIdx.x * TN;
    const int co0 = blockIdx.y * TM;
    const bool w_vec = (a.cin % 4) == 0 && (reinterpret_cast<uintptr_t>(a.w) & 15) == 0;   // parameter views may be unaligned

    f32x2 acc[RM][4];                                    // [row][column pair]: pairs (0,1)(2,3) | (4,5)(6,7)
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;

    // the (batch, point) of this thread's X loads: float4 index v -> row k = v / (TN/4), column group c4 = v % (TN/4)
    float4 xr[XV], wr[WV];
    auto load_chunk = [&](int k0) {
#pragma unroll
        for (int q = 0; q < XV; ++q) {
            const int v = threadIdx.x + q * NT;
            const int k = v / (TN / 4), c4 = v % (TN / 4);
            const long long col = col0 + c4 * 4;
            const int gk = k0 + k;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < KC * TN / 4 && col < cols && gk < a.cin) {
                const long long bi = col / a.n;
                const int p = (int)(col - bi * a.n);     // n % 4 == 0: the 4 columns share the batch element
                val = __ldg(reinterpret_cast<const float4 *>(a.x + bi * a.x_bstride + (size_t)gk * a.n + p));
            }
            xr[q] = val;
        }
#pragma unroll
        for (int q = 0; q < WV; ++q) {
            const int v = threadIdx.x + q * NT;
            const int co = v / (KC / 4), k4 = v % (KC / 4);
            const int gco = co0 + co, gk = k0 + k4 * 4;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < TM * KC / 4 && gco < a.cout) {
                const float *src = a.w + (size_t)gco * a.cin + gk;
                if (w_vec && gk + 3 < a.cin) val = __ldg(reinterpret_cast<const float4 *>(src));
                else {
                    if (gk < a.cin) val.x = __ldg(src);
                    if (gk + 1 < a.cin) val.y = __ldg(src + 1);
                    if (gk + 2 < a.cin) val.z = __ldg(src + 2);
                    if (gk + 3 < a.cin) val.w = __ldg(src + 3);
                }
            }
            wr[q] = val;
        }
    };
    auto store_chunk = [&](int buf) {
#pragma unroll
        for (int q = 0; q < XV; ++q) {
            const int v = threadIdx.x + q * NT;
            if (v < KC * TN / 4) *reinterpret_cast<float4 *>(&Xs[buf][v / (TN / 4)][(v % (TN / 4)) * 4]) = xr[q];
        }
#pragma unroll
        for (int q = 0; q < WV; ++q) {
            const int v = threadIdx.x + q * NT;
            if (v < TM * KC / 4) {
                const int co = v / (KC / 4), k = (v % (KC / 4)) * 4;
                Ws[buf][k][co] = wr[q].x; Ws[buf][k + 1][co] = wr[q].y; Ws[buf][k + 2][co] = wr[q].z; Ws[buf][k + 3][co] = wr[q].w;
            }
        }
    };

    const int nchunks = (a.cin + KC - 1) / KC;
    load_chunk(0);
    store_chunk(0);
    __syncthreads();
    for (int ch = 0; ch < nchunks; ++ch) {
        const int buf = ch & 1;
        if (ch + 1 < nchunks) load_chunk((ch + 1) * KC);   // in flight during the FMAs below
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const float4 xa = *reinterpret_cast<const float4 *>(&Xs[buf][k][tx * 4]);
            const float4 xb = *reinterpret_cast<const float4 *>(&Xs[buf][k][TN / 2 + tx * 4]);
            const f32x2 x01 = pack2(xa.x, xa.y), x23 = pack2(xa.z, xa.w), x45 = pack2(xb.x, xb.y), x67 = pack2(xb.z, xb.w);
            float wv[RM];
            if constexpr (RM % 4 == 0) {
#pragma unroll
                for (int i = 0; i < RM; i += 4) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(&Ws[buf][k][ty * RM + i]);
                    wv[i] = w4.x; wv[i + 1] = w4.y; wv[i + 2] = w4.z; wv[i + 3] = w4.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < RM; ++i) wv[i] = Ws[buf][k][ty * RM + i];
            }
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                const f32x2 ww = pack2(wv[i], wv[i]);
                acc[i][0] = fma2(ww, x01, acc[i][0]);
                acc[i][1] = fma2(ww, x23, acc[i][1]);
                acc[i][2] = fma2(ww, x45, acc[i][2]);
                acc[i][3] = fma2(ww, x67, acc[i][3]);
            }
        }
        if (ch + 1 < nchunks) {
            store_chunk(buf ^ 1);     // the other buffer: nobody reads it during this chunk
            __syncthreads();
        }
    }
    // ---- epilogue: bias, ReLU, residual, 128-bit stores --------------------------------------------------
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int co = co0 + ty * RM + i;
        if (co >= a.cout) continue;
        const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long col = col0 + h * (TN / 2) + tx * 4;
            if (col >= cols) continue;
            const long long bi = col / a.n;
            const int p = (int)(col - bi * a.n);
            float v[4];
            unpack2(acc[i][h * 2], v[0], v[1]);
            unpack2(acc[i][h * 2 + 1], v[2], v[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] += bv;
                if (a.relu) v[j] = fmaxf(v[j], 0.f);
                if (a.res) v[j] += __ldg(a.res + bi * a.res_bstride + (size_t)co * a.res_n + (p + j) / a.res_div);
            }
            float4 *dst = reinterpret_cast<float4 *>(a.y + bi * a.y_bstride + (size_t)co * a.n + p);
            if (a.mask) {
                const float4 mk = __ldg(reinterpret_cast<const float4 *>(a.mask + bi * a.mask_bstride + (size_t)co * a.n + p));
                if (!(mk.x > 0.f)) v[0] = 0.f;
                if (!(mk.y > 0.f)) v[1] = 0.f;
                if (!(mk.z > 0.f)) v[2] = 0.f;
                if (!(mk.w > 0.f)) v[3] = 0.f;
            }
            if (a.accumulate) { const float4 o = *dst; v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w; }
            *dst = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// Feature expansion of the up-sampling head (upsampler.py:349-366): the reference replicates every point's
// 264 features r times, appends one code channel and runs the 265->128 convolution on the (B,265,N*r) tensor.
// The replicas share W[:, :264] . x, so that product is computed once per point (pre, (B,cout,N)) and this
// kernel only adds the code column:  Y[b,co,p*r+j] = relu(pre[b,co,p] + wcode[co] * code[j]).
__global__ void __launch_bounds__(256) expand_code_kernel(int b, int cout, int n, int r, const float *__restrict__ pre,
                                                         const float *__restrict__ w, int w_stride, int code_col,
                                                         const float *__restrict__ code, float *__restrict__ y) {
    const long long total = (long long)b * cout * n * r;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % r);
        const long long q = t / r;           // (b*cout + co)*n + p
        const int co = (int)((q / n) % cout);
        const float v = __fmaf_rn(__ldg(w + (size_t)co * w_stride + code_col), __ldg(code + j), __ldg(pre + q));
        y[t] = fmaxf(v, 0.f);
    }
}

}  // namespace pu3

using namespace pu3;

// Test hook: force the generic (unaligned-safe) kernel.
static int g_pw_force_generic = 0;
extern "C" void pu3_pointwise_force_generic(int on) { g_pw_force_generic = on; }
#define RM4_OK(cout) true

extern "C" int pu3_pointwise_conv_f32(int b, int n, int cin, int cout, const float *x, long long x_bstride,
                                      const float *w, const float *bias, float *y, long long y_bstride,
                                      const float *res, long long res_bstride, int res_n, int res_div, int relu,
                                      pu3_stream_t stream) {
    return pu3_pointwise_conv_ex_f32(b, n, cin, cout, x, x_bstride, w, bias, y, y_bstride, res, res_bstride, res_n, res_div, relu,
                                     nullptr, 0, 0, stream);
}

extern "C" int pu3_pointwise_conv_ex_f32(int b, int n, int cin, int cout, const float *x, long long x_bstride,
                                         const float *w, const float *bias, float *y, long long y_bstride,
                                         const float *res, long long res_bstride, int res_n, int res_div, int relu,
                                         const float *mask, long long mask_bstride, int accumulate, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && n >= 0 && cin > 0 && cout > 0, "pointwise_conv: bad size b=%d n=%d cin=%d cout=%d", b, n, cin, cout);
    if (b == 0 || n == 0) return PU3_OK;
    PU3_ARG_CHECK(x && w && y, "pointwise_conv: null pointer");
    PU3_ARG_CHECK(!res || (res_div >= 1 && res_n >= 1), "pointwise_conv: bad residual description");
    PwArgs a{b, n, cin, cout, x, x_bstride, w, bias, y, y_bstride, res, res_bstride, res_n, res_div > 0 ? res_div : 1, relu,
             mask, mask_bstride, accumulate};
    const long long cols = (long long)b * n;
    cudaStream_t s = as_stream(stream);
    const bool aligned = (n % 4 == 0) && (x_bstride % 4 == 0) && (y_bstride % 4 == 0) && (!mask || (mask_bstride % 4 == 0 && ((uintptr_t)mask & 15) == 0)) &&
                         (((uintptr_t)x | (uintptr_t)y) & 15) == 0 && g_pw_force_generic == 0;
    if (aligned && RM4_OK(cout)) {
        if (cout <= 4) {
            dim3 grid((unsigned)((cols + 255) / 256), 1);
            pointwise_conv_fast_kernel<1, 4, 32><<<grid, 128, 0, s>>>(a);
        } else if (cout <= 24) {
            dim3 grid((unsigned)((cols + 255) / 256), (cout + 23) / 24);
            pointwise_conv_fast_kernel<8, 3, 32><<<grid, 96, 0, s>>>(a);
        } else if (cout <= 64) {
            dim3 grid((unsigned)((cols + 127) / 128), (cout + 63) / 64);
            pointwise_conv_fast_kernel<8, 8, 16><<<grid, 128, 0, s>>>(a);
        } else {
            dim3 grid((unsigned)((cols + 127) / 128), (cout + 127) / 128);
            pointwise_conv_fast_kernel<8, 16, 16><<<grid, 256, 0, s>>>(a);
        }
        PU3_LAUNCH_CHECK("pointwise_conv_fast_kernel");
        return PU3_OK;
    }
    if (cout <= 4) {          // 64 -> 3 regressor: 4 x 256-column tiles, 32 threads
        dim3 grid((unsigned)((cols + 255) / 256), (cout + 3) / 4);
        pointwise_conv_kernel<4, 1, 32><<<grid, 32, 0, s>>>(a);
    } else if (cout <= 24) {  // 3->24, 84/144/204 -> 24: 24 x 256 tile, 96 threads
        dim3 grid((unsigned)((cols + 255) / 256), (cout + 23) / 24);
        pointwise_conv_kernel<8, 3, 32><<<grid, 96, 0, s>>>(a);
    } else if (cout <= 64) {  // 128 -> 64: 64 x 256 tile
        dim3 grid((unsigned)((cols + 255) / 256), (cout + 63) / 64);
        pointwise_conv_kernel<8, 8, 32><<<grid, 256, 0, s>>>(a);
    } else {                  // 264/128 -> 128: 128 x 128 tile
        dim3 grid((unsigned)((cols + 127) / 128), (cout + 127) / 128);
        pointwise_conv_kernel<8, 16, 16><<<grid, 256, 0, s>>>(a);
    }
    PU3_LAUNCH_CHECK("pointwise_conv_kernel");
    return PU3_OK;
}

extern "C" int pu3_expand_code_f32(int b, int cout, int n, int r, const float *pre, const float *w, int w_stride,
                                   int code_col, const float *code, float *y, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && cout > 0 && n >= 0 && r > 0, "expand_code: bad size");
    const long long total = (long long)b * cout * n * r;
    if (total == 0) return PU3_OK;
    PU3_ARG_CHECK(pre && w && code && y, "expand_code: null pointer");
    const long long blocks = (total + 255) / 256;
    const long long cap = (long long)device_info().sm_count * 16;
    expand_code_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(b, cout, n, r, pre, w, w_stride,
                                                                                         code_col, code, y);
    PU3_LAUNCH_CHECK("expand_code_kernel");
    return PU3_OK;
}

// ------------------------------------------------------------------------------------------------------
// Backward of the 1x1 convolution with respect to the weights and the bias:
//     dW[co,ci] += sum_{b,p} dY[b,co,p] * X[b,ci,p]        db[co] += sum_{b,p} dY[b,co,p]
// (what autograd computes for nn.Conv1d/Conv2d in the reference's train step, model.py:62).  A GEMM whose
// reduction dimension is the flattened (batch, point) axis: split-K over column chunks, 64x64 output tile per
// CTA, 4x4 per thread, partial tiles combined with fp32 atomics (summation order unspecified, like cuDNN's).
// dX needs no kernel of its own: it is the forward kernel applied to dY with W transposed.
// ------------------------------------------------------------------------------------------------------
namespace pu3 {

constexpr int BW_T = 64;       // output tile (co x ci)
constexpr int BW_KC = 16;      // columns per shared-memory chunk
constexpr int BW_COLS = 128;   // columns per CTA (split-K granularity).  ncu (profiles/r2): with 512 every launch took ~120 us whatever its size --
                               // 39 CTAs each walking 32 serial load -> barrier -> FMA chunks; 128 gives 4x the CTAs and a 4x shorter chain

__global__ void __launch_bounds__(256) pointwise_conv_bwd_w_kernel(int b, int n, int cin, int cout,
                                                                  const float *__restrict__ x, long long x_bstride,
                                                                  const float *__restrict__ dy, long long dy_bstride,
                                                                  float *__restrict__ dw, int dw_stride, float *__restrict__ db) {
    __shared__ float As[BW_KC][BW_T + 1];   // dY chunk, [col][co]
    __shared__ float Bs[BW_KC][BW_T + 1];   // X chunk,  [col][ci]
    const long long cols = (long long)b * n;
    const long long c_begin = (long long)blockIdx.x * BW_COLS;
    const long long c_end = min(cols, c_begin + BW_COLS);
    const int co0 = blockIdx.y * BW_T, ci0 = blockIdx.z * BW_T;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 x 16 threads, 4x4 outputs each
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;                        // threads 0..63 of the ci-tile-0 CTAs also sum dY rows (bias gradient)
    const bool do_bias = db != nullptr && blockIdx.z == 0;
    for (long long c0 = c_begin; c0 < c_end; c0 += BW_KC) {
        for (int t = threadIdx.x; t < BW_KC * BW_T; t += 256) {
            const int k = t % BW_KC, r = t / BW_KC;           // consecutive threads walk the columns: coalesced
            const long long col = c0 + k;
            float av = 0.f, bv = 0.f;
            if (col < c_end) {
                const long long bi = col / n;
                const int p = (int)(col - bi * n);
                if (co0 + r < cout) av = __ldg(dy + bi * dy_bstride + (size_t)(co0 + r) * n + p);
                if (ci0 + r < cin) bv = __ldg(x + bi * x_bstride + (size_t)(ci0 + r) * n + p);
            }
            As[k][r] = av; Bs[k][r] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BW_KC; ++k) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; bb[i] = Bs[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(a[i], bb[j], acc[i][j]);
        }
        if (do_bias && threadIdx.x < BW_T) {
#pragma unroll
            for (int k = 0; k < BW_KC; ++k) bsum += As[k][threadIdx.x];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = ci0 + tx * 4 + j;
            if (ci < cin) atomicAdd(dw + (size_t)co * dw_stride + ci, acc[i][j]);
        }
    }
    if (do_bias && threadIdx.x < BW_T && co0 + (int)threadIdx.x < cout) atomicAdd(db + co0 + threadIdx.x, bsum);
}

}  // namespace pu3

extern "C" int pu3_pointwise_conv_bwd_w_f32(int b, int n, int cin, int cout, const float *x, long long x_bstride,
                                            const float *dy, long long dy_bstride, float *dw, float *db,
                                            pu3_stream_t stream) {
    return pu3_pointwise_conv_bwd_w_ex_f32(b, n, cin, cout, x, x_bstride, dy, dy_bstride, dw, cin, db, stream);
}

extern "C" int pu3_pointwise_conv_bwd_w_ex_f32(int b, int n, int cin, int cout, const float *x, long long x_bstride,
                                               const float *dy, long long dy_bstride, float *dw, int dw_stride, float *db,
                                               pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && n >= 0 && cin > 0 && cout > 0 && dw_stride >= cin, "pointwise_conv_bwd_w: bad size");
    if (b == 0 || n == 0) return PU3_OK;
    PU3_ARG_CHECK(x && dy && dw, "pointwise_conv_bwd_w: null pointer");
    const long long cols = (long long)b * n;
    dim3 grid((unsigned)((cols + pu3::BW_COLS - 1) / pu3::BW_COLS), (cout + pu3::BW_T - 1) / pu3::BW_T, (cin + pu3::BW_T - 1) / pu3::BW_T);
    pu3::pointwise_conv_bwd_w_kernel<<<grid, 256, 0, pu3::as_stream(stream)>>>(b, n, cin, cout, x, x_bstride, dy, dy_bstride, dw, dw_stride, db);
    PU3_LAUNCH_CHECK("pointwise_conv_bwd_w_kernel");
    return PU3_OK;
}
