// DenseEdgeConv forward with the per-edge layers on the tensor cores (tcgen05, 3xTF32), sm_100a.
//
// Same function and the same exact re-association as edgeconv.cu (network/layers.py:22-64 of the reference): per point
// the centre terms A0|A1|A2 and P_j = W0[:, 24:] x_j are computed once, an edge (i, j) then is
//     h0 = relu(A0_i + P_j)                     gather + 12 adds (SIMT)
//     h1 = relu(W1[:, :12] h0 + A1_i)           12 x 12
//     h2 = W2[:, :12] h1 + W2[:, 12:24] h0 + A2_i   12 x 24
// and y = [max_k h2, max_k h1, max_k h0, c].  The FFMA kernel spends 432 MAC per edge in FFMA2 at 48 % issue efficiency;
// here the two small layers are GEMMs over the EDGES: a tile is 128 edges = 4 points x 32 neighbours (one warp per point,
// one lane per edge = one row of the MMA = one lane of tensor memory), K = 12 padded to 16, and
//     stage 1:  [h1 | h2 part] (128 x 64) = h0_hi . [W1a_hi | W1a_lo | W2b_hi | W2b_lo]^T ,  (128 x 32) = h0_lo . [W1a_hi | W2b_hi]^T
//     stage 2:  h2 (128 x 32, accumulating) += h1_hi . [W2a_hi | W2a_lo]^T ,  (128 x 16) += h1_lo . W2a_hi^T
// i.e. 8 small tcgen05.mma per tile (3xTF32: hi.hi in a main accumulator, hi.lo + lo.hi in correction columns, summed in
// the epilogue -- profiles/r1g: the accuracy of the FFMA kernel).  The threads only add the centre terms, apply ReLU, split
// hi/lo, write the next operand image (K-major, no swizzle) and take the maximum over the 32 edges of a point with
// redux.sync.max.f32 (the 32 edges of a point are the 32 lanes of a warp).  A CTA is two warpgroups working on their own
// tiles; two CTAs per SM (2 x 256 tensor-memory columns), so four tiles per SM are in flight and hide each other's MMA waits.
#include "tc_common.cuh"

namespace pu3 {
using namespace tc;

constexpr int ET_C = 24, ET_G = 12;
constexpr int ET_WGS = 2;                 // warpgroups per CTA
constexpr int ET_THREADS = ET_WGS * 128;
constexpr int ET_BLK = 32;                // points per output staging block
constexpr int ET_XS = 25;                 // row stride of the staged cloud (prolog only)
constexpr int ET_IMG = 8192;              // one operand image: 2 k-steps x (128 rows x 8 tf32)
constexpr int ET_KSTEP = 4096;
constexpr int ET_SO = 36 * (ET_BLK + 1);  // floats of one warpgroup's output block
constexpr int ET_TMEM_WG = 128;           // tensor-memory columns per warpgroup (96 used)
constexpr int ET_WB = 8192;               // weight images: B1 (64 rows) 4096 | Bc1 (32 rows) 2048 | B2 (32 rows) 2048

struct EtWeights { const float *w0, *b0, *w1, *b1, *w2, *b2; };   // (12,48) (12,36) (12,48), row-major as in the state_dict

__host__ __device__ inline size_t et_oper_bytes(int n) {
    const size_t xs = ((size_t)n * ET_XS * sizeof(float) + 127) & ~(size_t)127;
    const size_t op = (size_t)ET_WGS * 2 * ET_IMG;
    return xs > op ? xs : op;
}
__host__ inline size_t et_smem_bytes(int n, int pts) {
    return 128 + et_oper_bytes(n) + ET_WB + (size_t)n * ET_G * 4 + (size_t)pts * 36 * 4 + (size_t)ET_WGS * ET_SO * 4 + 64;
}

__device__ __forceinline__ constexpr uint32_t et_idesc(uint32_t ncols) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((ncols >> 3) << 17) | ((128u >> 4) << 24);   // f32 accumulate, tf32 x tf32, K-major both
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float warp_max_f32(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void wg_sync(int wg) {
    if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
}

// bring-up timeline (build with `make -B EXTRA=-DPU3_ET_TIMELINE`, then profiles/edgeconv_tc_timeline.py): SM cycle counter at the
// phases of tiles [8, 16) of CTA (0,0), warpgroup 0, thread 0.  Compiled out by default.
#ifdef PU3_ET_TIMELINE
__device__ unsigned int *g_et_timeline = nullptr;
__device__ __forceinline__ void et_mark(int tile, int phase) {
    if (g_et_timeline && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && tile >= 8 && tile < 16)
        g_et_timeline[(tile - 8) * 16 + phase] = (unsigned int)clock64();
}
__device__ __forceinline__ void et_mark_cta(int slot) {
    if (g_et_timeline && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_et_timeline[128 + slot] = (unsigned int)clock64();
}
#else
__device__ __forceinline__ void et_mark(int, int) {}
__device__ __forceinline__ void et_mark_cta(int) {}
#endif

// the row's 12 values as hi / lo tf32 chunks of the K-major operand images (chunk 3 = K 12..15 stays zero)
__device__ __forceinline__ void et_store_split(unsigned char *img_hi, unsigned char *img_lo, uint32_t row_off, const float (&r)[ET_G]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float4 hi, lo;
        hi.x = to_tf32(r[4 * c + 0]); hi.y = to_tf32(r[4 * c + 1]); hi.z = to_tf32(r[4 * c + 2]); hi.w = to_tf32(r[4 * c + 3]);
        lo.x = r[4 * c + 0] - hi.x; lo.y = r[4 * c + 1] - hi.y; lo.z = r[4 * c + 2] - hi.z; lo.w = r[4 * c + 3] - hi.w;
        const uint32_t off = (uint32_t)(c >> 1) * ET_KSTEP + (uint32_t)(c & 1) * 128u + row_off;
        *reinterpret_cast<float4 *>(img_hi + off) = hi;
        *reinterpret_cast<float4 *>(img_lo + off) = lo;
    }
}

__global__ void __launch_bounds__(ET_THREADS, 2) edgeconv_tc_kernel(int n, int pts_per_cta, const float *__restrict__ x, long long x_bstride,
                                                                    const int32_t *__restrict__ idx, int idx_stride, int idx_off,
                                                                    EtWeights W, float *__restrict__ y, long long y_bstride) {
    extern __shared__ unsigned char raw_[];
    unsigned char *sm = raw_ + ((128u - (smem_u32(raw_) & 127u)) & 127u);
    unsigned char *oper = sm;                                                    // [wg][hi|lo][ET_IMG]; the prolog keeps the cloud here
    unsigned char *wb = oper + et_oper_bytes(n);
    float *sP = reinterpret_cast<float *>(wb + ET_WB);                           // [n][12]
    float *sA = sP + (size_t)n * ET_G;                                           // [pts][36]: A0 | A1 | A2
    float *s_out = sA + (size_t)pts_per_cta * 36;                                // [wg][36][33]; the prolog keeps wcat | bias here
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_out + ET_WGS * ET_SO);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + ET_WGS);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wg = warp >> 2, wq = warp & 3;
    const int bi = blockIdx.y;
    const float *xb = x + bi * x_bstride;
    float *yb = y + bi * y_bstride;
    const int32_t *ib = idx + (size_t)bi * n * idx_stride + idx_off;
    const int p_begin = blockIdx.x * pts_per_cta, p_end = min(n, p_begin + pts_per_cta);

    // ---------------- prolog: cloud, weights, per-point terms ----------------------------------------------------------
    et_mark_cta(0);
    if (tid == 0) {
        for (int g = 0; g < ET_WGS; ++g) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[g])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(ET_WGS * ET_TMEM_WG)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    et_mark_cta(1);
    float *xs = reinterpret_cast<float *>(oper);
    float *wcat = s_out;                 // [24][48]: W0b | W0a - W0b | W1[:, 12:] | W2[:, 24:]   (in-major)
    float *bias = s_out + ET_C * 48;     // [48]: 0 | b0 | b1 | b2
    for (int t = tid; t < ET_C * n; t += ET_THREADS) {
        const int c = t / n, p = t - c * n;
        xs[p * ET_XS + c] = __ldg(xb + (size_t)c * n + p);
    }
    for (int t = tid; t < ET_C * 48; t += ET_THREADS) {
        const int ch = t / 48, o = t - ch * 48;
        float v;
        if (o < 12) v = __ldg(W.w0 + o * 48 + 24 + ch);
        else if (o < 24) v = __ldg(W.w0 + (o - 12) * 48 + ch) - __ldg(W.w0 + (o - 12) * 48 + 24 + ch);
        else if (o < 36) v = __ldg(W.w1 + (o - 24) * 36 + 12 + ch);
        else v = __ldg(W.w2 + (o - 36) * 48 + 24 + ch);
        wcat[t] = v;
    }
    if (tid < 48) bias[tid] = tid < 12 ? 0.f : (tid < 24 ? __ldg(W.b0 + tid - 12) : (tid < 36 ? __ldg(W.b1 + tid - 24) : __ldg(W.b2 + tid - 36)));
    et_mark_cta(2);
    // weight images, K-major without swizzle: element (row r, k) of a k-step at (r / 8) * 256 + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4
    for (int t = tid; t < 128 * 16; t += ET_THREADS) {
        const int rr = t >> 4, kk = t & 15;                  // rr: 0..63 B1, 64..95 Bc1, 96..127 B2
        int img_off, nr, r;
        if (rr < 64) { img_off = 0; nr = 64; r = rr; }
        else if (rr < 96) { img_off = 4096; nr = 32; r = rr - 64; }
        else { img_off = 6144; nr = 32; r = rr - 96; }
        const int blk = r >> 4, o = r & 15;
        float w = 0.f;
        bool lo = false;
        if (o < ET_G && kk < ET_G) {
            if (rr < 64) { w = blk < 2 ? __ldg(W.w1 + o * 36 + kk) : __ldg(W.w2 + o * 48 + 12 + kk); lo = blk & 1; }
            else if (rr < 96) { w = blk == 0 ? __ldg(W.w1 + o * 36 + kk) : __ldg(W.w2 + o * 48 + 12 + kk); }
            else { w = __ldg(W.w2 + o * 48 + kk); lo = blk & 1; }
        }
        const float hi = to_tf32(w);
        const float v = lo ? to_tf32(w - hi) : hi;
        const int off = img_off + (kk >> 3) * (nr * 32) + (r >> 3) * 256 + ((kk & 7) >> 2) * 128 + (r & 7) * 16 + (kk & 3) * 4;
        *reinterpret_cast<float *>(wb + off) = v;
    }
    et_mark_cta(3);
    __syncthreads();
    et_mark_cta(4);
    for (int t = tid; t < n * 3; t += ET_THREADS) {          // P_j for the whole cloud (any point can be a neighbour)
        const int p = t / 3, q = t - p * 3;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float *row = xs + p * ET_XS;
#pragma unroll
        for (int ch = 0; ch < ET_C; ++ch) {
            const float4 w = *reinterpret_cast<const float4 *>(wcat + ch * 48 + q * 4);
            const float v = row[ch];
            acc.x = __fmaf_rn(w.x, v, acc.x); acc.y = __fmaf_rn(w.y, v, acc.y); acc.z = __fmaf_rn(w.z, v, acc.z); acc.w = __fmaf_rn(w.w, v, acc.w);
        }
        *reinterpret_cast<float4 *>(sP + p * ET_G + q * 4) = acc;
    }
    const int pcnt = p_end - p_begin;
    et_mark_cta(5);
    for (int t = tid; t < pcnt * 9; t += ET_THREADS) {       // centre terms of the CTA's own points
        const int p = t / 9, q = t - p * 9;
        float4 acc = *reinterpret_cast<const float4 *>(bias + 12 + q * 4);
        const float *row = xs + (p_begin + p) * ET_XS;
#pragma unroll
        for (int ch = 0; ch < ET_C; ++ch) {
            const float4 w = *reinterpret_cast<const float4 *>(wcat + ch * 48 + 12 + q * 4);
            const float v = row[ch];
            acc.x = __fmaf_rn(w.x, v, acc.x); acc.y = __fmaf_rn(w.y, v, acc.y); acc.z = __fmaf_rn(w.z, v, acc.z); acc.w = __fmaf_rn(w.w, v, acc.w);
        }
        *reinterpret_cast<float4 *>(sA + p * 36 + q * 4) = acc;
    }
    et_mark_cta(6);
    for (int t = tid; t < ET_C * pcnt; t += ET_THREADS) {    // y[36..59] = the centre itself
        const int c = t / pcnt, p = t - c * pcnt;
        yb[(size_t)(36 + c) * n + p_begin + p] = xs[(p_begin + p) * ET_XS + c];
    }
    et_mark_cta(7);
    __syncthreads();
    et_mark_cta(8);
    for (int t = tid; t < ET_WGS * 2 * ET_IMG / 16; t += ET_THREADS) reinterpret_cast<float4 *>(oper)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    et_mark_cta(9);

    // ---------------- main loop: a warpgroup walks its blocks of 32 points, 4 points (128 edges) per tile -----------------
    const uint32_t tmem_wg = *tmem_slot + (uint32_t)wg * ET_TMEM_WG;
    const uint32_t tmem_rd = tmem_wg + ((uint32_t)(wq * 32) << 16);
    unsigned char *img_hi = oper + (size_t)wg * 2 * ET_IMG, *img_lo = img_hi + ET_IMG;
    const uint32_t row = (uint32_t)(wq * 32 + lane);
    const uint32_t row_off = (row >> 3) * 256u + (row & 7u) * 16u;
    const uint64_t d_ah = smem_desc(smem_u32(img_hi), 128, 256, 0), d_al = smem_desc(smem_u32(img_lo), 128, 256, 0);
    const uint64_t d_b1 = smem_desc(smem_u32(wb), 128, 256, 0), d_bc1 = smem_desc(smem_u32(wb + 4096), 128, 256, 0),
                   d_b2 = smem_desc(smem_u32(wb + 6144), 128, 256, 0);
    uint64_t *bar = &bars[wg];
    uint32_t phase = 0;
    float *so = s_out + wg * ET_SO;
    const bool issuer = (wq == 0 && lane == 0);
    // the prolog's wcat | bias live where the output blocks go: every thread is past its reads (barriers above)

    const int nblk = (pcnt + ET_BLK - 1) / ET_BLK;
    int jn = 0, tile_no = 0;
    if (wg < nblk) jn = __ldg(ib + (size_t)min(p_begin + wg * ET_BLK + wq, p_end - 1) * idx_stride + lane);
    for (int blk = wg; blk < nblk; blk += ET_WGS) {
        const int b0p = p_begin + blk * ET_BLK;
        const int bcnt = min(ET_BLK, p_end - b0p);
        const int ntile = (bcnt + 3) >> 2;
        for (int t = 0; t < ntile; ++t) {
            const int lp = t * 4 + wq;
            const bool pv = lp < bcnt;
            const int i = pv ? b0p + lp : p_end - 1;
            const int j = jn;
            {   // the neighbour index of this thread's next edge, one tile ahead
                int nb = blk, nt = t + 1;
                if (nt >= ntile) { nb = blk + ET_WGS; nt = 0; }
                if (nb < nblk) jn = __ldg(ib + (size_t)min(p_begin + nb * ET_BLK + nt * 4 + wq, p_end - 1) * idx_stride + lane);
            }
            const float4 *ai = reinterpret_cast<const float4 *>(sA + (size_t)(i - p_begin) * 36);
            float k0 = 0.f, k1 = 0.f, k2 = 0.f;
            et_mark(tile_no, 0);
            float r[ET_G];
            // ---- layer 0: gather + add + ReLU -------------------------------------------------------------------------------
            {
                const float4 *pj = reinterpret_cast<const float4 *>(sP + (size_t)j * ET_G);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float4 p = pj[q], a = ai[q];
                    r[4 * q + 0] = fmaxf(p.x + a.x, 0.f); r[4 * q + 1] = fmaxf(p.y + a.y, 0.f);
                    r[4 * q + 2] = fmaxf(p.z + a.z, 0.f); r[4 * q + 3] = fmaxf(p.w + a.w, 0.f);
                }
#pragma unroll
                for (int o = 0; o < ET_G; ++o) { const float m = warp_max_f32(r[o]); if (lane == o) k0 = m; }
                et_store_split(img_hi, img_lo, row_off, r);
            }
            et_mark(tile_no, 1);
            fence_proxy_async();
            tc_fence_before();
            wg_sync(wg);
            et_mark(tile_no, 2);
            if (issuer) {
                tc_fence_after();
#pragma unroll
                for (uint32_t ks = 0; ks < 2; ++ks) {
                    umma_tf32(tmem_wg + 0, d_ah + ks * (ET_KSTEP >> 4), d_b1 + ks * (2048 >> 4), et_idesc(64), ks);
                    umma_tf32(tmem_wg + 64, d_al + ks * (ET_KSTEP >> 4), d_bc1 + ks * (1024 >> 4), et_idesc(32), ks);
                }
                tc_commit(bar);
            }
            et_mark(tile_no, 3);
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after();
            et_mark(tile_no, 4);
            // ---- layer 1 epilogue: + A1, ReLU, max, next operand ----------------------------------------------------------------
            {
                float a[16], b[16], c[16];
                tmem_ld16(tmem_rd + 0, a); tmem_ld16(tmem_rd + 16, b); tmem_ld16(tmem_rd + 64, c);
                tmem_ld_wait();
                et_mark(tile_no, 5);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float4 a1 = ai[3 + q];
                    r[4 * q + 0] = fmaxf((a[4 * q + 0] + a1.x) + (b[4 * q + 0] + c[4 * q + 0]), 0.f);
                    r[4 * q + 1] = fmaxf((a[4 * q + 1] + a1.y) + (b[4 * q + 1] + c[4 * q + 1]), 0.f);
                    r[4 * q + 2] = fmaxf((a[4 * q + 2] + a1.z) + (b[4 * q + 2] + c[4 * q + 2]), 0.f);
                    r[4 * q + 3] = fmaxf((a[4 * q + 3] + a1.w) + (b[4 * q + 3] + c[4 * q + 3]), 0.f);
                }
#pragma unroll
                for (int o = 0; o < ET_G; ++o) { const float m = warp_max_f32(r[o]); if (lane == o) k1 = m; }
                et_store_split(img_hi, img_lo, row_off, r);     // stage 1 has completed: its operand images are free
            }
            et_mark(tile_no, 6);
            fence_proxy_async();
            tc_fence_before();
            wg_sync(wg);
            et_mark(tile_no, 7);
            if (issuer) {
                tc_fence_after();
#pragma unroll
                for (uint32_t ks = 0; ks < 2; ++ks) {
                    umma_tf32(tmem_wg + 32, d_ah + ks * (ET_KSTEP >> 4), d_b2 + ks * (1024 >> 4), et_idesc(32), 1u);
                    umma_tf32(tmem_wg + 80, d_al + ks * (ET_KSTEP >> 4), d_b2 + ks * (1024 >> 4), et_idesc(16), 1u);
                }
                tc_commit(bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
            tc_fence_after();
            et_mark(tile_no, 8);
            // ---- layer 2 epilogue: + A2, max (no ReLU: layers.py:58-59) -------------------------------------------------------
            {
                float a[16], b[16], c[16];
                tmem_ld16(tmem_rd + 32, a); tmem_ld16(tmem_rd + 48, b); tmem_ld16(tmem_rd + 80, c);
                tmem_ld_wait();
                et_mark(tile_no, 9);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float4 a2 = ai[6 + q];
                    r[4 * q + 0] = (a[4 * q + 0] + a2.x) + (b[4 * q + 0] + c[4 * q + 0]);
                    r[4 * q + 1] = (a[4 * q + 1] + a2.y) + (b[4 * q + 1] + c[4 * q + 1]);
                    r[4 * q + 2] = (a[4 * q + 2] + a2.z) + (b[4 * q + 2] + c[4 * q + 2]);
                    r[4 * q + 3] = (a[4 * q + 3] + a2.w) + (b[4 * q + 3] + c[4 * q + 3]);
                }
#pragma unroll
                for (int o = 0; o < ET_G; ++o) { const float m = warp_max_f32(r[o]); if (lane == o) k2 = m; }
            }
            if (pv && lane < ET_G) {
                so[lane * (ET_BLK + 1) + lp] = k2;
                so[(12 + lane) * (ET_BLK + 1) + lp] = k1;
                so[(24 + lane) * (ET_BLK + 1) + lp] = k0;
            }
            et_mark(tile_no, 10);
            ++tile_no;
        }
        wg_sync(wg);
        for (int e = (tid & 127); e < 36 * ET_BLK; e += 128) {
            const int ch = e >> 5, c = e & 31;
            if (c < bcnt) yb[(size_t)ch * n + b0p + c] = so[ch * (ET_BLK + 1) + c];
        }
        // the next block's first write to `so` comes after two more warpgroup barriers: no barrier needed here
    }

    et_mark_cta(10);
    tc_fence_before();
    __syncthreads();
    et_mark_cta(11);
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"((uint32_t)(ET_WGS * ET_TMEM_WG)) : "memory");
    }
}

// launched by pu3_edgeconv_f32 (edgeconv.cu) for k == 32; returns false when the shape does not fit (caller falls back)
bool edgeconv_tc_launch(int b, int n, const float *x, long long x_bstride, const int32_t *idx, int idx_stride, int idx_off,
                        const float *w0, const float *b0, const float *w1, const float *b1, const float *w2, const float *b2,
                        float *y, long long y_bstride, cudaStream_t s, int *status) {
    const int sms = device_info().sm_count;
    int pts = n;
    if ((long long)b < 2LL * sms) {                     // few clouds: several CTAs per cloud, each at least two blocks per warpgroup pair
        const int split = (int)((2LL * sms + b - 1) / b);
        pts = (n + split - 1) / split;
        pts = ((pts + 2 * ET_BLK - 1) / (2 * ET_BLK)) * (2 * ET_BLK);
        if (pts > n) pts = n;
    }
    const size_t smem = et_smem_bytes(n, pts);
    if (smem > 112 * 1024) return false;
    static bool attr_done = false;
    if (!attr_done) {
        *status = cuda_status(cudaFuncSetAttribute(edgeconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024), "edgeconv_tc: smem attr");
        if (*status) return true;
        attr_done = true;
    }
    EtWeights W{w0, b0, w1, b1, w2, b2};
    dim3 grid((n + pts - 1) / pts, b);
    edgeconv_tc_kernel<<<grid, ET_THREADS, smem, s>>>(n, pts, x, x_bstride, idx, idx_stride, idx_off, W, y, y_bstride);
    *status = PU3_OK;
    return true;
}

}  // namespace pu3

#ifdef PU3_ET_TIMELINE
extern "C" void pu3_edgeconv_tc_set_timeline(unsigned int *buf) { cudaMemcpyToSymbol(pu3::g_et_timeline, &buf, sizeof(buf)); }   // bring-up hook
#endif
