// DenseEdgeConv forward, every GEMM operand of the edge threads in TENSOR MEMORY (tcgen05.mma TS form), sm_100a.
//
// Same decomposition as edgeconv_tc.cu (network/layers.py:22-64 of the reference; a tile = 128 edges = 4 points x 32 neighbours, one warp
// per point, one lane per edge = one MMA row = one lane of tensor memory, max over a point's edges by redux.sync.max.f32, warp-specialised
// MMA issue, two tiles in flight per warpgroup, persistent CTAs).  What changes: the [hi | lo] operand rows of BOTH per-edge layers and of
// the prolog GEMM are written with tcgen05.st into tensor memory and read by the MMA from there -- no operand images in shared memory at
// all (edgeconv_tc.cu: shared-memory pipe 71 % busy, half of it operand images written by STS.128 and read back by the tensor core).
// To make room in the 128 tensor-memory columns of a warpgroup the three 3xTF32 products of a layer go into ONE accumulator
// (K <= 36: the accumulate-rounding of the tensor core stays ~1e-7, profiles/r1g/ncu_summary.md) as two passes:
//     [hi | lo] (K = 24) . [W_hi ; W_hi]^T      3 MMAs        +        [hi | lo(0..3)] (K = 16) . [W_lo ; 0]^T      2 MMAs
// Per tile slot: h1 accumulator 16 columns | h2 accumulator 16 | operand 24 = 56; layer 2's operand overwrites layer 1's.
// Shared memory then only holds the weights, P (n x 12), the centre terms (n x 36) and an output block: a WHOLE cloud per work item
// fits (edgeconv_tc.cu needed two items per cloud, each repeating the cloud's prolog).
#include <cstdlib>

#include "tc_common.cuh"

namespace pu3 {
using namespace tc;

constexpr int EU_C = 24, EU_G = 12;
constexpr int EU_WGS = 2;                      // warpgroups of edge threads per CTA
constexpr int EU_SIMT = EU_WGS * 128;
constexpr int EU_THREADS = EU_SIMT + 32 * EU_WGS;   // + one MMA-issuing warp per warpgroup
constexpr int EU_BLK = 16;                     // points per output staging block (4 tiles)
// weight images, K-major without swizzle, one [rows][8] block per k-step (bytes)
constexpr int EU_B1HH = 0, EU_B1HL = EU_B1HH + 32 * 24 * 4, EU_B2HH = EU_B1HL + 32 * 16 * 4, EU_B2HL = EU_B2HH + 16 * 24 * 4,
              EU_BPHH = EU_B2HL + 16 * 16 * 4, EU_BPHL = EU_BPHH + 48 * 48 * 4, EU_WB = EU_BPHL + 48 * 24 * 4;
constexpr int EU_SO = 36 * (EU_BLK + 1);       // floats of one output block
constexpr int EU_SOUT = EU_WGS * EU_SO * 4;    // one block per warpgroup (bytes)
constexpr int EU_TMEM_WG = 128;                // tensor-memory columns per warpgroup
constexpr int EU_SLOT = 64;                    // a tile slot: +0 h1 accumulator, later h1 . W2a (16), +16 h0 . W2b (16), +32 operand [hi | lo] (24)
constexpr int EU_XCOL = 64;                    // prolog: accumulator [P | A] at +0 (48), operand [x_hi | x_lo] at +64 (48)
constexpr int EU_RAW = 576 + 432 + 576 + 36;   // floats of the raw weights staged by the prolog

struct EuWeights { const float *w0, *b0, *w1, *b1, *w2, *b2; };   // (12,48) (12,36) (12,48), row-major as in the state_dict

__host__ inline size_t eu_smem_bytes(int n, int pts) {
    return 128 + EU_WB + (size_t)n * EU_G * 4 + (size_t)pts * 36 * 4 + EU_SOUT + 48 * 4 + 8 * 8 + 16;
}

__device__ __forceinline__ constexpr uint32_t eu_idesc(uint32_t ncols) {
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
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// mbarrier wait that parks the thread in hardware (suspend-time hint, CUTLASS' ClusterBarrier::wait form) instead of polling: the
// polling loop of tc::mbar_wait costs issue slots -- ncu counted 644 warp instructions per point with it, the MMA lanes spinning all
// the time on two of the four schedulers.  Still bounded: traps after ~4 s.
__device__ __forceinline__ void eu_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity), "r"(10000000u) : "memory");
        if (!done && spins > 400u) __trap();
    }
}
__device__ __forceinline__ float warp_max_f32(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
// lane o (< 12) keeps m[o]: a select tree on the lane's bits (11 selects, 4 predicates) instead of 12 compare-and-select pairs
__device__ __forceinline__ float eu_pick12(const float (&m)[EU_G], int lane) {
    const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4, b3 = lane & 8;
    const float a0 = b0 ? m[1] : m[0], a1 = b0 ? m[3] : m[2], a2 = b0 ? m[5] : m[4], a3 = b0 ? m[7] : m[6], a4 = b0 ? m[9] : m[8], a5 = b0 ? m[11] : m[10];
    const float c0 = b1 ? a1 : a0, c1 = b1 ? a3 : a2, c2 = b1 ? a5 : a4;
    const float d0 = b2 ? c1 : c0;
    return b3 ? c2 : d0;
}
__device__ __forceinline__ float eu_max12(const float (&r)[EU_G], int lane) {
    float m[EU_G];
#pragma unroll
    for (int o = 0; o < EU_G; ++o) m[o] = warp_max_f32(r[o]);
    return eu_pick12(m, lane);
}
__device__ __forceinline__ void wg_sync(int wg) {
    if (wg == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
}

// element (row r, k) of a weight image with 8 k per k-step: k-step stride rows * 32 B, 8-row groups 256 B apart, the two 4-k halves 128 B apart
__device__ __forceinline__ int eu_w_off(int rows, int r, int kk) {
    return (kk >> 3) * (rows * 32) + (r >> 3) * 256 + ((kk & 7) >> 2) * 128 + (r & 7) * 16 + (kk & 3) * 4;
}
__device__ __forceinline__ float eu_hi_or_lo(float w, bool lo) {
    const float hi = to_tf32(w);
    return lo ? to_tf32(w - hi) : hi;
}


// the row's 12 values as [hi(12) | lo(12)] tf32 into 24 columns of tensor memory
__device__ __forceinline__ void eu_store_split(uint32_t taddr, const float (&r)[EU_G]) {
    uint32_t w0[16], w1[8];
#pragma unroll
    for (int o = 0; o < EU_G; ++o) {
        const float h = to_tf32(r[o]);
        w0[o] = __float_as_uint(h);
        const uint32_t l = __float_as_uint(r[o] - h);
        if (o < 4) w0[12 + o] = l; else w1[o - 4] = l;
    }
    tmem_st16(taddr, w0);
    tmem_st8(taddr + 16, w1);
    tmem_st_wait();
}

struct EuSlot {            // what an edge thread carries for a tile in flight
    int j;                 // neighbour index of the thread's edge (fetched one round ahead)
    float k0, k1;          // lanes 0..11: max over the point's edges of h0[lane], h1[lane]
    uint32_t phase;        // parity of the slot's `done` barrier
    const float4 *ai;      // the point's centre terms A0 | A1 | A2 (9 float4)
    int lp;                // the point's column in the output block, -1 beyond the item's last point
};

__global__ void __launch_bounds__(EU_THREADS, 2) edgeconv_ts_kernel(int n, int pts_per_cta, int splits, int items, const float *__restrict__ x,
                                                                    long long x_bstride, const int32_t *__restrict__ idx, int idx_stride, int idx_off,
                                                                    EuWeights W, float *__restrict__ y, long long y_bstride) {
    extern __shared__ unsigned char raw_[];
    unsigned char *sm = raw_ + ((128u - (smem_u32(raw_) & 127u)) & 127u);
    unsigned char *wb = sm;                                                      // weight images
    float *sP = reinterpret_cast<float *>(wb + EU_WB);                           // [n][12]
    float *sA = sP + (size_t)n * EU_G;                                           // [pts][36]: A0 | A1 | A2; the raw weights first
    float *s_out = sA + (size_t)pts_per_cta * 36;                                // [wg][36][17]
    float *bias = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(s_out) + EU_SOUT);   // [48]: 0 | b0 | b1 | b2
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias + 48);                    // ready[wg][slot] (128 arrivals), done[wg][slot] (commit)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool mma_warp = warp >= EU_SIMT / 32;
    const int wg = mma_warp ? warp - EU_SIMT / 32 : warp >> 2, wq = warp & 3;
    uint64_t *ready = bars + wg * 2, *done = bars + 4 + wg * 2;

    if (tid == 0) {
        for (int g = 0; g < 4; ++g) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(smem_u32(&bars[g])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[4 + g])) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(EU_WGS * EU_TMEM_WG)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---------------- once per CTA: raw weights -> shared memory -> the tf32 weight images ------------------------------------
    float *rw = sA;                                      // w0 (576) | w1 (432) | w2 (576) | b0 b1 b2 (36); needs pts >= 46
    for (int t = tid; t < EU_RAW; t += EU_THREADS) {
        float v;
        if (t < 576) v = __ldg(W.w0 + t);
        else if (t < 1008) v = __ldg(W.w1 + t - 576);
        else if (t < 1584) v = __ldg(W.w2 + t - 1008);
        else if (t < 1596) v = __ldg(W.b0 + t - 1584);
        else if (t < 1608) v = __ldg(W.b1 + t - 1596);
        else v = __ldg(W.b2 + t - 1608);
        rw[t] = v;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const float *rw0 = rw, *rw1 = rw + 576, *rw2 = rw + 1008;
    // prolog weights Wp[o][ch] = W0b | W0a - W0b | W1[:, 12:] | W2[:, 24:] (48 x 24): BPHH K 0..23 hi, K 24..47 hi again (against x_lo);
    // BPHL K 0..23 lo
    for (int t = tid; t < 48 * 72; t += EU_THREADS) {
        const int o = t / 72, kk = t - o * 72, ch = kk % 24;
        float v;
        if (o < 12) v = rw0[o * 48 + 24 + ch];
        else if (o < 24) v = rw0[(o - 12) * 48 + ch] - rw0[(o - 12) * 48 + 24 + ch];
        else if (o < 36) v = rw1[(o - 24) * 36 + 12 + ch];
        else v = rw2[(o - 36) * 48 + 24 + ch];
        if (kk < 48) *reinterpret_cast<float *>(wb + EU_BPHH + eu_w_off(48, o, kk)) = to_tf32(v);
        else *reinterpret_cast<float *>(wb + EU_BPHL + eu_w_off(48, o, kk - 48)) = eu_hi_or_lo(v, true);
    }
    // layer 1 | h0 part of layer 2: rows 0..15 W1a = W1[:, :12], rows 16..31 W2b = W2[:, 12:24]; layer 2: rows 0..15 W2a = W2[:, :12].
    // HH images: K 0..11 hi, K 12..23 hi again (against the lo half of the operand); HL images: K 0..11 lo, K 12..15 zero
    for (int t = tid; t < 48 * 40; t += EU_THREADS) {
        const int rr = t / 40, kk = t - rr * 40;          // rr 0..31: layer 1 images, 32..47: layer 2 images; kk 0..23 HH, 24..39 HL
        const int r = rr < 32 ? rr : rr - 32, o = r & 15, in = kk < 24 ? kk % 12 : kk - 24;
        float v = 0.f;
        if (o < EU_G && in < EU_G) {
            const float w = rr < 16 ? rw1[o * 36 + in] : (rr < 32 ? rw2[o * 48 + 12 + in] : rw2[o * 48 + in]);
            v = eu_hi_or_lo(w, kk >= 24);
        }
        const int rows = rr < 32 ? 32 : 16;
        const int base = rr < 32 ? (kk < 24 ? EU_B1HH : EU_B1HL) : (kk < 24 ? EU_B2HH : EU_B2HL);
        *reinterpret_cast<float *>(wb + base + eu_w_off(rows, r, kk < 24 ? kk : kk - 24)) = v;
    }
    if (tid < 48) bias[tid] = tid < 12 ? 0.f : rw[1584 + tid - 12];
    fence_proxy_async();
    __syncthreads();          // the raw weights are dead: sA is free

    const uint32_t tmem_wg = *tmem_slot + (uint32_t)wg * EU_TMEM_WG;
    const int m_tiles = (n + 127) >> 7;
    const int row = wq * 32 + lane;                       // row of the warpgroup's tile = lane of tensor memory

    if (mma_warp) {
        // =============================== MMA issuer of warpgroup `wg` (one lane) ===============================================
        uint32_t rph[2] = {0u, 0u};
        const uint64_t d_b1hh = smem_desc(smem_u32(wb + EU_B1HH), 128, 256, 0), d_b1hl = smem_desc(smem_u32(wb + EU_B1HL), 128, 256, 0),
                       d_b2hh = smem_desc(smem_u32(wb + EU_B2HH), 128, 256, 0), d_b2hl = smem_desc(smem_u32(wb + EU_B2HL), 128, 256, 0),
                       d_bphh = smem_desc(smem_u32(wb + EU_BPHH), 128, 256, 0), d_bphl = smem_desc(smem_u32(wb + EU_BPHL), 128, 256, 0);
        auto tile_of = [&](int q) { return (q >> 2) * 8 + wg * 4 + (q & 3); };
        auto issue = [&](int sl, bool second) {
            eu_wait(&ready[sl], rph[sl]); rph[sl] ^= 1u;
            tc_fence_after();
            const uint32_t base = tmem_wg + sl * EU_SLOT, a = base + 32;
            // One accumulator per product, but the accumulate-rounding of the tensor core (a truncation at the accumulator's magnitude per
            // MMA) is kept to two steps: the small terms first -- hi . W_lo (2 MMAs), then lo . W_hi (k-step 2) -- and the two k-steps
            // that carry hi . W_hi last.  Measured against float64: the same error as separate main / correction accumulators.
            if (!second) {      // [h1 | h2a] (32 columns) = [h0_hi | .] . [W_lo ; 0]^T + [h0_hi | h0_lo] . [W_hi ; W_hi]^T
#pragma unroll
                for (uint32_t s = 0; s < 2; ++s) umma_tf32_ts(base, a + s * 8, d_b1hl + s * 64, eu_idesc(32), s);
#pragma unroll
                for (int s = 2; s >= 0; --s) umma_tf32_ts(base, a + s * 8, d_b1hh + s * 64, eu_idesc(32), 1u);
            } else {            // h2b (16 columns, over the dead h1 accumulator) = the same passes with [h1_hi | h1_lo] and W2a
#pragma unroll
                for (uint32_t s = 0; s < 2; ++s) umma_tf32_ts(base, a + s * 8, d_b2hl + s * 32, eu_idesc(16), s);
#pragma unroll
                for (int s = 2; s >= 0; --s) umma_tf32_ts(base, a + s * 8, d_b2hh + s * 32, eu_idesc(16), 1u);
            }
            tc_commit(&done[sl]);
        };
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int p_begin = (item % splits) * pts_per_cta, pcnt = min(n, p_begin + pts_per_cta) - p_begin;
            if (lane == 0) {
                for (int mt = wg; mt < m_tiles; mt += EU_WGS) {      // prolog: [P | A] (48 columns) = x . Wp^T per 128 points, two passes
                    eu_wait(&ready[0], rph[0]); rph[0] ^= 1u;
                    tc_fence_after();
#pragma unroll
                    for (uint32_t s = 0; s < 3; ++s) umma_tf32_ts(tmem_wg, tmem_wg + EU_XCOL + s * 8, d_bphl + s * 96, eu_idesc(48), s);      // x_hi . W_lo
#pragma unroll
                    for (int s = 5; s >= 0; --s) umma_tf32_ts(tmem_wg, tmem_wg + EU_XCOL + s * 8, d_bphh + s * 96, eu_idesc(48), 1u);        // x_lo . W_hi, then x_hi . W_hi
                    tc_commit(&done[0]);
                }
            }
            __syncwarp();
            asm volatile("bar.sync 3, %0;" ::"r"(EU_THREADS) : "memory");      // with every thread of the CTA: P of the whole cloud is in place
            if (lane == 0) {
                const int tiles_total = (pcnt + 3) >> 2;
                for (int q = 0; tile_of(q) < tiles_total; q += 2) {
                    const bool two = tile_of(q + 1) < tiles_total;
                    issue(0, false); if (two) issue(1, false);
                    issue(0, true); if (two) issue(1, true);
                }
            }
            __syncwarp();
            asm volatile("bar.sync 3, %0;" ::"r"(EU_THREADS) : "memory");      // the item is finished: P | A may be overwritten
        }
    } else {
        // =============================== edge threads ==========================================================================
        EuSlot S[2];
        S[0].phase = S[1].phase = 0u;
        const uint32_t tmem_rd = tmem_wg + ((uint32_t)(wq * 32) << 16);
        float *so = s_out + (size_t)wg * EU_SO;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int bi = item / splits;
            const float *xb = x + bi * x_bstride;
            float *yb = y + bi * y_bstride;
            const int32_t *ib = idx + (size_t)bi * n * idx_stride + idx_off;
            const int p_begin = (item - bi * splits) * pts_per_cta, p_end = min(n, p_begin + pts_per_cta);
            const int pcnt = p_end - p_begin;
            // ---- prolog: the thread's point of every 128-point tile of this warpgroup: centre copy, [hi | lo] operand row, P | A row
            for (int mt = wg; mt < m_tiles; mt += EU_WGS) {
                const int p = mt * 128 + row;
                float xv[EU_C];
#pragma unroll
                for (int c = 0; c < EU_C; ++c) xv[c] = p < n ? __ldg(xb + (size_t)c * n + p) : 0.f;
                const bool own = p >= p_begin && p < p_end;
                if (own) {
#pragma unroll
                    for (int c = 0; c < EU_C; ++c) yb[(size_t)(36 + c) * n + p] = xv[c];
                }
                {
                    uint32_t w[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) w[c] = __float_as_uint(to_tf32(xv[c]));
                    tmem_st16(tmem_rd + EU_XCOL, w);
#pragma unroll
                    for (int c = 0; c < 8; ++c) w[c] = __float_as_uint(to_tf32(xv[16 + c]));
#pragma unroll
                    for (int c = 0; c < 8; ++c) w[8 + c] = __float_as_uint(xv[c] - to_tf32(xv[c]));
                    tmem_st16(tmem_rd + EU_XCOL + 16, w);
#pragma unroll
                    for (int c = 0; c < 16; ++c) w[c] = __float_as_uint(xv[8 + c] - to_tf32(xv[8 + c]));
                    tmem_st16(tmem_rd + EU_XCOL + 32, w);
                    tmem_st_wait();
                }
                tc_fence_before();
                mbar_arrive(&ready[0]);
                eu_wait(&done[0], S[0].phase); S[0].phase ^= 1u;
                tc_fence_after();
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    float a[16];
                    tmem_ld16(tmem_rd + ch * 16, a);
                    tmem_ld_wait();
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const int o = ch * 16 + q4 * 4;          // outputs o..o+3: P for o < 12, A[o - 12] above
                        const float4 bb = *reinterpret_cast<const float4 *>(bias + o);
                        float4 v;
                        v.x = a[q4 * 4 + 0] + bb.x; v.y = a[q4 * 4 + 1] + bb.y; v.z = a[q4 * 4 + 2] + bb.z; v.w = a[q4 * 4 + 3] + bb.w;
                        if (o < 12) { if (p < n) *reinterpret_cast<float4 *>(sP + (size_t)p * EU_G + o) = v; }
                        else if (own) *reinterpret_cast<float4 *>(sA + (size_t)(p - p_begin) * 36 + (o - 12)) = v;
                    }
                }
                tc_fence_before();
            }
            asm volatile("bar.sync 3, %0;" ::"r"(EU_THREADS) : "memory");

            // ---- main loop: two tiles (4 points x 32 edges each) in flight per warpgroup ---------------------------------------------
            const int tiles_total = (pcnt + 3) >> 2;
            auto tile_of = [&](int q) { return (q >> 2) * 8 + wg * 4 + (q & 3); };
            auto point_of = [&](int q) { return min(p_begin + tile_of(q) * 4 + wq, p_end - 1); };
            auto fetch = [&](int q) { return __ldg(ib + (size_t)point_of(q) * idx_stride + lane); };
            // layer 0 of tile q: gather + add + ReLU, max, operand row into tensor memory, hand-over to the MMA warp
            auto layer0 = [&](int q, int sl) {
                const float4 *ai = reinterpret_cast<const float4 *>(sA + (size_t)(point_of(q) - p_begin) * 36);
                S[sl].ai = ai;
                S[sl].lp = p_begin + tile_of(q) * 4 + wq < p_end ? (q & 3) * 4 + wq : -1;
                const float4 *pj = reinterpret_cast<const float4 *>(sP + (size_t)S[sl].j * EU_G);
                if (tile_of(q + 2) < tiles_total) S[sl].j = fetch(q + 2);
                float r[EU_G];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 p = pj[c], a = ai[c];
                    r[4 * c + 0] = fmaxf(p.x + a.x, 0.f); r[4 * c + 1] = fmaxf(p.y + a.y, 0.f);
                    r[4 * c + 2] = fmaxf(p.z + a.z, 0.f); r[4 * c + 3] = fmaxf(p.w + a.w, 0.f);
                }
                S[sl].k0 = eu_max12(r, lane);
                eu_store_split(tmem_rd + sl * EU_SLOT + 32, r);
                tc_fence_before();
                mbar_arrive(&ready[sl]);
            };
            // layer 1 of tile q: accumulator + A1, ReLU, max; [hi | lo] over layer 1's operand columns as the operand of layer 2
            auto layer1 = [&](int q, int sl) {
                const float4 *ai = S[sl].ai + 3;
                eu_wait(&done[sl], S[sl].phase); S[sl].phase ^= 1u;
                tc_fence_after();
                float a[16], r[EU_G];
                tmem_ld16(tmem_rd + sl * EU_SLOT, a);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 a1 = ai[c];
                    r[4 * c + 0] = fmaxf(a[4 * c + 0] + a1.x, 0.f); r[4 * c + 1] = fmaxf(a[4 * c + 1] + a1.y, 0.f);
                    r[4 * c + 2] = fmaxf(a[4 * c + 2] + a1.z, 0.f); r[4 * c + 3] = fmaxf(a[4 * c + 3] + a1.w, 0.f);
                }
                S[sl].k1 = eu_max12(r, lane);
                eu_store_split(tmem_rd + sl * EU_SLOT + 32, r);
                tc_fence_before();
                mbar_arrive(&ready[sl]);
            };
            // layer 2 of tile q: accumulator + A2 (no ReLU: layers.py:58-59), max, the point's 36 values into the output block
            auto layer2 = [&](int q, int sl) {
                const int g = tile_of(q);
                const float4 *ai = S[sl].ai + 6;
                eu_wait(&done[sl], S[sl].phase); S[sl].phase ^= 1u;
                tc_fence_after();
                float a[16], b[16];
                tmem_ld16(tmem_rd + sl * EU_SLOT + 16, a); tmem_ld16(tmem_rd + sl * EU_SLOT, b);      // h0 . W2b (stage 1), h1 . W2a (stage 2)
                tmem_ld_wait();
                tc_fence_before();
                float r[EU_G];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float4 a2 = ai[c];
                    r[4 * c + 0] = (a[4 * c + 0] + a2.x) + b[4 * c + 0]; r[4 * c + 1] = (a[4 * c + 1] + a2.y) + b[4 * c + 1];
                    r[4 * c + 2] = (a[4 * c + 2] + a2.z) + b[4 * c + 2]; r[4 * c + 3] = (a[4 * c + 3] + a2.w) + b[4 * c + 3];
                }
                const float k = eu_max12(r, lane);
                const int lp = S[sl].lp;
                if (lane < EU_G && lp >= 0) {
                    so[lane * (EU_BLK + 1) + lp] = k;
                    so[(12 + lane) * (EU_BLK + 1) + lp] = S[sl].k1;
                    so[(24 + lane) * (EU_BLK + 1) + lp] = S[sl].k0;
                }
                if ((q & 3) == 3 || tile_of(q + 1) >= tiles_total) {       // the block is complete: 36 rows of up to 16 points
                    wg_sync(wg);
                    const int b0p = p_begin + (g >> 2) * EU_BLK, bcnt = min(EU_BLK, p_end - b0p);
                    for (int e = (tid & 127); e < 36 * EU_BLK; e += 128) {
                        const int ch = e >> 4, c = e & 15;
                        if (c < bcnt) yb[(size_t)ch * n + b0p + c] = so[ch * (EU_BLK + 1) + c];
                    }
                    wg_sync(wg);
                }
            };
            if (tile_of(0) < tiles_total) S[0].j = fetch(0);
            if (tile_of(1) < tiles_total) S[1].j = fetch(1);
            if (tile_of(0) < tiles_total) layer0(0, 0);
            if (tile_of(1) < tiles_total) layer0(1, 1);
            for (int q = 0; tile_of(q) < tiles_total; q += 2) {
                const bool two = tile_of(q + 1) < tiles_total;
                layer1(q, 0);
                if (two) layer1(q + 1, 1);
                layer2(q, 0);
                if (tile_of(q + 2) < tiles_total) layer0(q + 2, 0);
                if (two) {
                    layer2(q + 1, 1);
                    if (tile_of(q + 3) < tiles_total) layer0(q + 3, 1);
                }
            }
            asm volatile("bar.sync 3, %0;" ::"r"(EU_THREADS) : "memory");      // every warp is done with P | A of this item
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"((uint32_t)(EU_WGS * EU_TMEM_WG)) : "memory");
    }
}

// launched by pu3_edgeconv_f32 (edgeconv.cu) for k == 32; returns false when the shape does not fit (caller falls back)
bool edgeconv_ts_launch(int b, int n, const float *x, long long x_bstride, const int32_t *idx, int idx_stride, int idx_off,
                        const float *w0, const float *b0, const float *w1, const float *b1, const float *w2, const float *b2,
                        float *y, long long y_bstride, cudaStream_t s, int *status) {
    const int sms = device_info().sm_count;
    // points per work item: the whole cloud when there are clouds enough to fill two CTAs per SM, else a multiple of 16 chosen by a
    // small cost model -- rounds of the persistent grid x (prolog of the whole cloud, about two 16-point blocks' worth (fitted to the
    // measured 1275 / 640-cloud launches) + the busier warpgroup's 16-point blocks)
    const int n16 = ((n + 15) / 16) * 16;
    int pts = n16;
    {
        double best = 1e30;
        const long long slots = 2LL * sms;
        for (int cand = 48; cand <= n16; cand += 16) {
            const long long its = (long long)((n + cand - 1) / cand) * b;
            const long long rounds = (its + slots - 1) / slots;
            const double cost = (double)rounds * (2.0 + (double)(((cand + 15) / 16 + 1) / 2));
            if (cost < best - 1e-9) { best = cost; pts = cand; }      // ties: the smaller items (more CTAs busy in the last round)
        }
    }
    static const int forced_pts = [] { const char *e = getenv("PU3_EC_PTS"); return e ? atoi(e) : 0; }();     // tuning hook
    if (forced_pts >= 48 && forced_pts % 16 == 0) pts = forced_pts < n16 ? forced_pts : n16;
    const int splits = (n + pts - 1) / pts;
    const size_t smem = eu_smem_bytes(n, pts);
    if (smem > 113 * 1024 || n < 48) return false;
    // per launch, not once per process: the attribute belongs to the device the caller has made current
    *status = cuda_status(cudaFuncSetAttribute(edgeconv_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024), "edgeconv_ts: smem attr");
    if (*status) return true;
    EuWeights W{w0, b0, w1, b1, w2, b2};
    const long long items = (long long)splits * b;
    const int grid = (int)(items < 2LL * sms ? items : 2LL * sms);
    edgeconv_ts_kernel<<<grid, EU_THREADS, smem, s>>>(n, pts, splits, (int)items, x, x_bstride, idx, idx_stride, idx_off, W, y, y_bstride);
    *status = PU3_OK;
    return true;
}

}  // namespace pu3
