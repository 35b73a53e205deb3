// The whole expansion head of a Level (network/upsampler.py:349-372) as ONE persistent tcgen05 kernel:
//
//     features (t, cin=264, n)  --up_layer1 (+ code column, x2 replication, ReLU)-->  (128, 2n)
//                               --up_layer2 + ReLU-->  (128, 2n)  --fc_layer1 + ReLU-->  (64, 2n)
//                               --fc_layer2 + residual-->  out (t, 3, 2n)
//
// Three separate launches (csrc/conv_tc.cu) move ~4.1 KB per point through HBM (write 1 KB, read 1 KB, write 1 KB, read 1 KB)
// and cannot exceed ~55 % tensor-pipe activity even at full HBM speed; chained on chip the head reads the 264 features once
// (1.06 KB per input point) and writes 24 bytes, so the tensor cores are the bound.
//
// Design (sm_100a):
//   * tile = 128 input points = four TMA boxes of 32 points x 32 channels; the boxes of a tile may belong to different clouds
//     (312 points = 9.75 boxes: 2.5 % padding instead of the 18 % of 128-point sub-tiles per cloud), and TMEM lane quarter q of an
//     accumulator = box q, so every epilogue warp owns the 32 points of one box.
//   * 3xTF32 as in conv_tc.cu.  Three chained GEMMs per tile, all M = 128:
//       P1  up1   A = features, MN-major straight from channel-major HBM (TMA, 128B swizzle / 32B atoms), converted to hi/lo
//                 tf32 by 4 converter warps;  D1 = main | correction accumulators (2 x 128 TMEM columns)
//       P2  up2   A = relu(D1 + b1 + w_code * code[j]) for replica j = 0, 1: epilogue group j (4 warps, one per lane quarter)
//                 reads D1 from TMEM and WRITES the hi/lo operand tiles (K-major, 128B swizzle) into shared memory, 32 channels
//                 at a time -- the epilogue of one layer is the operand producer of the next;  D2 = 2 replicas x 128 columns
//                 (one accumulator per replica: TMEM holds 512 columns and D1 must stay readable while D2 fills)
//       P3  fc1   A = relu(D2 + b2), produced the same way;  D3 = 2 replicas x (main 64 | correction 64) in D1's columns
//       E3        relu(D3 + b3) stays in registers, fc_layer2 (64 -> 3) + bias + residual per point, 12 bytes stored.
//     The two 256-column TMEM regions swap roles every tile (D1/D3 of tile i in region i&1, D2 in the other) so that up1 of
//     tile i+1 runs while E3 of tile i drains.
//   * rings (shared memory, 208 KB): raw activations 3 x 16 KB (TMA landing zone), operand tiles 2 x (hi 16 KB | lo 16 KB)
//     shared by the three producers (converters, epilogue group 0, epilogue group 1) in one global order, weights 3 x 32 KB
//     (9 + 4 + 4 k-blocks per tile streamed from L2 by bulk copies; the next tile's activations are prefetched into L2).
//   * 16 warps: 0 activation TMA, 1 MMA issue, 2 weight copies, 4-7 converters, 8-11 / 12-15 epilogue groups 0 / 1.
//   * mbarrier parity waits are only safe when the waiter is at most one phase behind.  Ring slots alternate between producers,
//     so a producer that starts a new phase of the tile first waits for an event that implies "everything up to four uses before
//     mine has been consumed": accumulator-complete commits for the epilogue groups, a dedicated commit (p3_gate) for the
//     converters' first block of the next tile.
#include "tc_common.cuh"

namespace pu3 {
namespace tc {
namespace head {

constexpr int KB = 32;                         // channels per k-block
constexpr int NA = 3, NO = 2, NW = 3;          // ring depths: raw activations, operand tiles, weights
constexpr int RAW_BYTES = 128 * KB * 4;        // 16 KB: 128 points x 32 channels
constexpr int O_BYTES = 2 * RAW_BYTES;         // hi | lo
constexpr int W_BYTES = 2 * 128 * 128;         // [W_hi | W_lo], 128 rows of 128 B each
constexpr int W3_BYTES = 2 * 64 * 128;
constexpr int SMEM_BYTES = NA * RAW_BYTES + NO * O_BYTES + NW * W_BYTES + 1024;
constexpr int NUM_THREADS = 512;
constexpr int CONV_WARP0 = 4, EPI_WARP0 = 8;
constexpr int C1 = 128, C2 = 128, C3 = 64;     // channels of up1, up2, fc1

// instruction descriptors (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, M = 128, B K-major; bit 15 = A MN-major
constexpr uint32_t IDESC_K = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t IDESC_MN = IDESC_K | (1u << 15);
constexpr uint32_t idesc_n(uint32_t base, int n) { return base | ((uint32_t)(n >> 3) << 17); }

struct Args {
    int b, n, cin, bpc;                 // bpc = boxes of 32 points per cloud
    long long nboxes, ntiles;
    const unsigned char *w1, *w2, *w3;  // split weight images (pu3_conv_tc_prepare_f32)
    const float *bias1, *wfull; int w_stride, code_col; const float *code;
    const float *bias2, *bias3, *w4, *b4;
    const float *res; long long res_bstride;
    float *y; long long y_bstride;
    unsigned int *dbg;
};

// bring-up / tuning: SM cycle counter of events of CTA 0 (kind-major, 512 slots per kind)
__device__ __forceinline__ void tl_mark(unsigned int *dbg, int kind, int slot) {
    if (dbg && blockIdx.x == 0 && slot < 512) dbg[kind * 512 + slot] = (unsigned int)clock64();   // 16 kinds
}

__global__ void __launch_bounds__(NUM_THREADS, 1) head_tc_kernel(const __grid_constant__ CUtensorMap xmap, const Args a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_A[NA], empty_A[NA], full_O[NO], empty_O[NO], full_W[NW], empty_W[NW];
    __shared__ uint64_t acc_full[3], e2_done, e3_done, p3_gate;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias1_s[C1], wcode_s[C1], bias2_s[C2], bias3_s[C3], w4_s[3 * C3];
    __shared__ float b4_s[4], code_s[2];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
    const uint32_t sA = smem0, sO = smem0 + NA * RAW_BYTES, sW = sO + NO * O_BYTES;
    unsigned char *gA = smem_gen, *gO = smem_gen + NA * RAW_BYTES;

    const int nkb1 = (a.cin + KB - 1) / KB;
    const int last_ksteps = ((a.cin - (nkb1 - 1) * KB) + 7) / 8;
    const uint32_t U = (uint32_t)nkb1 + 16u;                 // operand-tile uses per tile: P1 | P2 (4 chunks x 2 replicas) | P3

    for (int i = threadIdx.x; i < C1; i += NUM_THREADS) {
        bias1_s[i] = a.bias1 ? a.bias1[i] : 0.f;
        wcode_s[i] = a.wfull[(size_t)i * a.w_stride + a.code_col];
        bias2_s[i] = a.bias2 ? a.bias2[i] : 0.f;
    }
    for (int i = threadIdx.x; i < C3; i += NUM_THREADS) bias3_s[i] = a.bias3 ? a.bias3[i] : 0.f;
    for (int i = threadIdx.x; i < 3 * C3; i += NUM_THREADS) w4_s[i] = a.w4[i];
    if (threadIdx.x < 4) b4_s[threadIdx.x] = (threadIdx.x < 3 && a.b4) ? a.b4[threadIdx.x] : 0.f;
    if (threadIdx.x < 2) code_s[threadIdx.x] = a.code[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(&full_A[s], 1); mbar_init(&empty_A[s], 4); }
        for (int s = 0; s < NO; ++s) { mbar_init(&full_O[s], 4); mbar_init(&empty_O[s], 1); }
        for (int s = 0; s < NW; ++s) { mbar_init(&full_W[s], 1); mbar_init(&empty_W[s], 1); }
        for (int s = 0; s < 3; ++s) mbar_init(&acc_full[s], 1);
        mbar_init(&e2_done, 8);
        mbar_init(&e3_done, 8);
        mbar_init(&p3_gate, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== activation TMA: 4 boxes of 32 points x 32 channels per k-block into the raw ring =====
        uint32_t slot = 0, phase = 0;
        for (long long t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
            const long long nt = t + gridDim.x;                       // HBM -> L2 for the next tile of this CTA
            if (nt < a.ntiles)
                for (int i = lane; i < nkb1 * 4; i += 32) {
                    const long long box = nt * 4 + (i & 3);
                    if (box < a.nboxes) tma_prefetch_3d(&xmap, (int)(box % a.bpc) * 32, (i >> 2) * KB, (int)(box / a.bpc));
                }
            for (int kb = 0; kb < nkb1; ++kb) {
                mbar_wait(&empty_A[slot], phase ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_A[slot], (uint32_t)RAW_BYTES);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const long long box = t * 4 + q;              // box >= nboxes: cloud index >= b, out of bounds, zeros
                        tma_load_3d(sA + slot * RAW_BYTES + q * 4096, &xmap, (int)(box % a.bpc) * 32, kb * KB, (int)(box / a.bpc), &full_A[slot]);
                    }
                }
                __syncwarp();
                if (++slot == NA) { slot = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 2) {
        // ===== weight k-blocks: W1[0..nkb1), W2[0..4), W3[0..4) per tile, bulk copies from L2 =====
        uint32_t slot = 0, phase = 0;
        for (long long t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
            for (int j = 0; j < nkb1 + 8; ++j) {
                mbar_wait(&empty_W[slot], phase ^ 1u);
                if (elect_one()) {
                    const unsigned char *src;
                    uint32_t bytes = W_BYTES;
                    if (j < nkb1) src = a.w1 + (size_t)j * W_BYTES;
                    else if (j < nkb1 + 4) src = a.w2 + (size_t)(j - nkb1) * W_BYTES;
                    else { src = a.w3 + (size_t)(j - nkb1 - 4) * W3_BYTES; bytes = W3_BYTES; }
                    mbar_arrive_expect_tx(&full_W[slot], bytes);
                    bulk_load(sW + slot * W_BYTES, src, bytes, &full_W[slot]);
                }
                __syncwarp();
                if (++slot == NW) { slot = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp waits, one elected lane issues) =====
        uint32_t u = 0, wslot = 0, wphase = 0, it = 0;
        for (long long t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
            const uint32_t RX = tmem_base + (it & 1u) * 256u, RY = tmem_base + ((it & 1u) ^ 1u) * 256u;
            if (it > 0) mbar_wait(&e2_done, (it - 1u) & 1u);          // D2 of the previous tile (this tile's D1 region) has been read
            tc_fence_after();
            if (lane == 0) tl_mark(a.dbg, 0, (int)it);
            // ---- P1: up1, two accumulators (main = hi.hi, correction = hi.lo + lo.hi)
            for (int kb = 0; kb < nkb1; ++kb, ++u) {
                const uint32_t os = u & 1u;
                mbar_wait(&full_W[wslot], wphase);
                mbar_wait(&full_O[os], (u >> 1) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t oa = sO + os * O_BYTES, wb = sW + wslot * W_BYTES;
                    const int nks = kb == nkb1 - 1 ? last_ksteps : KB / 8;
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint64_t ahi = smem_desc(oa + ks * 1024, 4096, 512, LAYOUT_SW128_32B);
                        const uint64_t alo = smem_desc(oa + RAW_BYTES + ks * 1024, 4096, 512, LAYOUT_SW128_32B);
                        const uint64_t bhl = smem_desc(wb + ks * 32, 16, 1024, LAYOUT_SW128);     // rows [W_hi; W_lo]
                        umma_tf32(RX, ahi, bhl, idesc_n(IDESC_MN, 2 * C1), (kb | ks) != 0);
                        umma_tf32(RX + C1, alo, bhl, idesc_n(IDESC_MN, C1), 1u);
                    }
                    tc_commit(&empty_O[os]);
                    tc_commit(&empty_W[wslot]);
                    if (kb == nkb1 - 1) tc_commit(&acc_full[0]);
                }
                __syncwarp();
                if (++wslot == NW) { wslot = 0; wphase ^= 1u; }
            }
            if (lane == 0) tl_mark(a.dbg, 1, (int)it);
            if (it > 0) mbar_wait(&e3_done, (it - 1u) & 1u);          // D3 of the previous tile (this tile's D2 region) has been read
            tc_fence_after();
            if (lane == 0) tl_mark(a.dbg, 2, (int)it);
            // ---- P2: up2, one accumulator per replica
            for (int ch = 0; ch < C1 / KB; ++ch) {
                mbar_wait(&full_W[wslot], wphase);
                for (int g = 0; g < 2; ++g, ++u) {
                    const uint32_t os = u & 1u;
                    mbar_wait(&full_O[os], (u >> 1) & 1u);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t oa = sO + os * O_BYTES, wb = sW + wslot * W_BYTES, d = RY + g * C2;
#pragma unroll
                        for (int ks = 0; ks < KB / 8; ++ks) {
                            const uint64_t ahi = smem_desc(oa + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint64_t alo = smem_desc(oa + RAW_BYTES + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint64_t bhi = smem_desc(wb + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint64_t blo = smem_desc(wb + C2 * 128 + ks * 32, 16, 1024, LAYOUT_SW128);
                            umma_tf32(d, ahi, blo, idesc_n(IDESC_K, C2), (ch | ks) != 0);   // small terms first
                            umma_tf32(d, alo, bhi, idesc_n(IDESC_K, C2), 1u);
                            umma_tf32(d, ahi, bhi, idesc_n(IDESC_K, C2), 1u);
                        }
                        tc_commit(&empty_O[os]);
                        if (g == 1) {
                            tc_commit(&empty_W[wslot]);
                            if (ch == C1 / KB - 1) tc_commit(&acc_full[1]);
                        }
                    }
                    __syncwarp();
                }
                if (++wslot == NW) { wslot = 0; wphase ^= 1u; }
            }
            if (lane == 0) tl_mark(a.dbg, 3, (int)it);
            // ---- P3: fc1, main | correction per replica, in D1's columns (all of D1 was read before D2 completed)
            for (int ch = 0; ch < C2 / KB; ++ch) {
                mbar_wait(&full_W[wslot], wphase);
                for (int g = 0; g < 2; ++g, ++u) {
                    const uint32_t os = u & 1u;
                    mbar_wait(&full_O[os], (u >> 1) & 1u);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t oa = sO + os * O_BYTES, wb = sW + wslot * W_BYTES, d = RX + g * (2 * C3);
#pragma unroll
                        for (int ks = 0; ks < KB / 8; ++ks) {
                            const uint64_t ahi = smem_desc(oa + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint64_t alo = smem_desc(oa + RAW_BYTES + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint64_t bhl = smem_desc(wb + ks * 32, 16, 1024, LAYOUT_SW128);   // rows [W_hi (64); W_lo (64)]
                            umma_tf32(d, ahi, bhl, idesc_n(IDESC_K, 2 * C3), (ch | ks) != 0);
                            umma_tf32(d + C3, alo, bhl, idesc_n(IDESC_K, C3), 1u);
                        }
                        tc_commit(&empty_O[os]);
                        if (g == 1) {
                            tc_commit(&empty_W[wslot]);
                            if (ch == C2 / KB - 2) tc_commit(&p3_gate);       // all but the last two operand tiles of this tile consumed
                            if (ch == C2 / KB - 1) tc_commit(&acc_full[2]);
                        }
                    }
                    __syncwarp();
                }
                if (++wslot == NW) { wslot = 0; wphase ^= 1u; }
            }
            if (lane == 0) tl_mark(a.dbg, 4, (int)it);
        }
    } else if (warp >= CONV_WARP0 && warp < EPI_WARP0) {
        // ===== converters (P1): raw fp32 -> hi = tf32(x), lo = x - hi, same (TMA-swizzled MN-major) layout =====
        const int ct = threadIdx.x - CONV_WARP0 * 32;        // 0..127
        uint32_t aslot = 0, aphase = 0, it = 0;
        for (long long t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
            if (it > 0) mbar_wait(&p3_gate, (it - 1u) & 1u);
            for (int kb = 0; kb < nkb1; ++kb) {
                const uint32_t u = it * U + (uint32_t)kb, os = u & 1u;
                mbar_wait(&full_A[aslot], aphase);
                const float4 *src = reinterpret_cast<const float4 *>(gA + aslot * RAW_BYTES);
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = src[ct + i * 128];
                mbar_wait(&empty_O[os], ((u >> 1) & 1u) ^ 1u);
                float4 *hi = reinterpret_cast<float4 *>(gO + os * O_BYTES);
                float4 *lo = reinterpret_cast<float4 *>(gO + os * O_BYTES + RAW_BYTES);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 h, l;
                    h.x = to_tf32(v[i].x); h.y = to_tf32(v[i].y); h.z = to_tf32(v[i].z); h.w = to_tf32(v[i].w);
                    l.x = v[i].x - h.x; l.y = v[i].y - h.y; l.z = v[i].z - h.z; l.w = v[i].w - h.w;
                    hi[ct + i * 128] = h;
                    lo[ct + i * 128] = l;
                }
                fence_proxy_async();                         // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) { mbar_arrive(&full_O[os]); mbar_arrive(&empty_A[aslot]); }
                if (++aslot == NA) { aslot = 0; aphase ^= 1u; }
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===== epilogue groups: group g = replica g; warp quarter q = TMEM lanes 32q.. = box q of the tile =====
        const int g = (warp - EPI_WARP0) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lanebits = (uint32_t)(q * 32) << 16;
        const uint32_t swz = (uint32_t)(row & 7);
        unsigned char *orow_base = gO + row * 128;
        const float code_g = code_s[g];
        uint32_t it = 0;
        for (long long t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
            const uint32_t par = it & 1u;
            const uint32_t RX = tmem_base + par * 256u + lanebits, RY = tmem_base + (par ^ 1u) * 256u + lanebits;
            const long long box = t * 4 + q;
            const long long bi = box / a.bpc;
            const int p = (int)(box % a.bpc) * 32 + lane;
            const bool valid = box < a.nboxes && p < a.n;
            const uint32_t ubase = it * U + (uint32_t)nkb1 + (uint32_t)g;
            // ---- E1: D1 -> relu(up1 + bias + w_code * code[g]) -> operand tiles of up2
            mbar_wait(&acc_full[0], par);
            tc_fence_after();
            if (warp == EPI_WARP0 && lane == 0) tl_mark(a.dbg, 5, (int)it);
#pragma unroll 1
            for (int ch = 0; ch < C1 / KB; ++ch) {
                uint32_t v[32], vc[32];
                tmem_ld32(RX + ch * 32, v);
                tmem_ld32(RX + C1 + ch * 32, vc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int co = ch * 32 + j;
                    const float pre = (__uint_as_float(v[j]) + __uint_as_float(vc[j])) + bias1_s[co];
                    v[j] = __float_as_uint(fmaxf(__fmaf_rn(wcode_s[co], code_g, pre), 0.f));
                }
                const uint32_t u = ubase + 2u * ch, os = u & 1u;
                mbar_wait(&empty_O[os], ((u >> 1) & 1u) ^ 1u);
                unsigned char *orow = orow_base + os * O_BYTES;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 h, l;
                    const float x0 = __uint_as_float(v[4 * c]), x1 = __uint_as_float(v[4 * c + 1]), x2 = __uint_as_float(v[4 * c + 2]), x3 = __uint_as_float(v[4 * c + 3]);
                    h.x = to_tf32(x0); h.y = to_tf32(x1); h.z = to_tf32(x2); h.w = to_tf32(x3);
                    l.x = x0 - h.x; l.y = x1 - h.y; l.z = x2 - h.z; l.w = x3 - h.w;
                    const uint32_t off = ((uint32_t)c ^ swz) << 4;
                    *reinterpret_cast<float4 *>(orow + off) = h;
                    *reinterpret_cast<float4 *>(orow + RAW_BYTES + off) = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_O[os]);
            }
            // ---- E2: D2 (replica g) -> relu(up2 + bias) -> operand tiles of fc1
            mbar_wait(&acc_full[1], par);
            tc_fence_after();
            if (warp == EPI_WARP0 && lane == 0) tl_mark(a.dbg, 6, (int)it);
#pragma unroll 1
            for (int ch = 0; ch < C2 / KB; ++ch) {
                uint32_t v[32];
                tmem_ld32(RY + g * C2 + ch * 32, v);
                tmem_ld_wait();
                if (ch == C2 / KB - 1) {                     // D2 fully read: the MMA warp may start up1 of the next tile in this region
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&e2_done);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]) + bias2_s[ch * 32 + j], 0.f));
                const uint32_t u = ubase + 8u + 2u * ch, os = u & 1u;
                mbar_wait(&empty_O[os], ((u >> 1) & 1u) ^ 1u);
                unsigned char *orow = orow_base + os * O_BYTES;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 h, l;
                    const float x0 = __uint_as_float(v[4 * c]), x1 = __uint_as_float(v[4 * c + 1]), x2 = __uint_as_float(v[4 * c + 2]), x3 = __uint_as_float(v[4 * c + 3]);
                    h.x = to_tf32(x0); h.y = to_tf32(x1); h.z = to_tf32(x2); h.w = to_tf32(x3);
                    l.x = x0 - h.x; l.y = x1 - h.y; l.z = x2 - h.z; l.w = x3 - h.w;
                    const uint32_t off = ((uint32_t)c ^ swz) << 4;
                    *reinterpret_cast<float4 *>(orow + off) = h;
                    *reinterpret_cast<float4 *>(orow + RAW_BYTES + off) = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_O[os]);
            }
            // ---- E3: D3 (replica g) -> relu(fc1 + bias) in registers -> fc_layer2 + bias + residual
            mbar_wait(&acc_full[2], par);
            tc_fence_after();
            if (warp == EPI_WARP0 && lane == 0) tl_mark(a.dbg, 7, (int)it);
            float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
            for (int ch = 0; ch < C3 / 32; ++ch) {
                uint32_t v[32], vc[32];
                tmem_ld32(RX + g * (2 * C3) + ch * 32, v);
                tmem_ld32(RX + g * (2 * C3) + C3 + ch * 32, vc);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int co = ch * 32 + j;
                    const float h = fmaxf((__uint_as_float(v[j]) + __uint_as_float(vc[j])) + bias3_s[co], 0.f);
                    o0 = __fmaf_rn(w4_s[co], h, o0);
                    o1 = __fmaf_rn(w4_s[C3 + co], h, o1);
                    o2 = __fmaf_rn(w4_s[2 * C3 + co], h, o2);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&e3_done);
            if (valid) {
                const float o[3] = {o0 + b4_s[0], o1 + b4_s[1], o2 + b4_s[2]};
#pragma unroll
                for (int c3 = 0; c3 < 3; ++c3) {
                    float r = o[c3];
                    if (a.res) r += __ldg(a.res + bi * a.res_bstride + (size_t)c3 * a.n + p);
                    a.y[bi * a.y_bstride + (size_t)c3 * (2 * a.n) + 2 * p + g] = r;
                }
            }
            if (warp == EPI_WARP0 && lane == 0) tl_mark(a.dbg, 8, (int)it);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// =====================================================================================================================
// Version 2 (default): activation operands in TENSOR MEMORY (tcgen05.mma "TS" form: A from TMEM, B = weights from smem).
//
// ncu of version 1 (profiles/r2/ncu_summary.md): 4.2 MB of shared-memory traffic per 128-point tile (tensor-core operand
// reads 1.9 MB, converter / epilogue loads and stores 1.7 MB, TMA writes 0.6 MB) at 128 B/clk = the measured 32.5 K cycles per
// tile: shared-memory-bandwidth bound at 48 % tensor-pipe activity.  Here the hi/lo activation tiles are written with
// tcgen05.st into tensor memory (lane = point, one column per channel) and read by the MMA from there, so shared memory only
// carries the raw TMA landing zone (read once by the converters) and the weight k-blocks: ~2.0 MB per tile.
//
// TMEM (512 columns) cannot hold both replicas' accumulators plus operands, so the replicas run one after the other (the
// weights of up2 / fc1 are streamed twice per tile) and every operand tile is written over columns that have just died:
//     tile parity p:  R = columns [256p, 256p+256),  O = the other half,  S = O[128:256)
//       up1      A: converters -> ring of 2 slots x (hi 32 | lo 32) in S            D1 = R[0:128) main | R[128:256) correction
//       E1(0)    group 0 reads main + correction chunk by chunk: pre = main + corr + bias -> over main (replica 1 needs it),
//                h = relu(pre + w_code code[0]): hi -> over the correction chunk just read, lo -> S[32 ch ..]
//       up2(0)   A = R[128:256) | S                                                  D2 = O[0:128)   (one accumulator)
//       E2(0)    relu(D2 + bias): hi -> over D2 in place, lo -> S
//       fc1(0)   A = O[0:128) | S                                                    D3 = R[128:256) (main 64 | correction 64)
//       E1(1)    group 1 reads pre: hi -> over pre in place, lo -> S chunk by chunk as fc1(0) releases it (lo_free)
//       up2(1)   A = R[0:128) | S                                                    D2 = O[0:128)
//       E2(1), fc1(1) as for replica 0                                               D3 = R[0:128)
//       E3(g)    relu(D3 + bias) in registers, fc_layer2 + bias + residual, 12 bytes per point stored
//     The tensor pipe executes MMAs in issue order, so "operand consumed before its columns are overwritten by a later MMA's
//     accumulator" needs no barrier; epilogue writes are ordered by the accumulator-complete commits.  up1 of the next tile
//     (D1 in O) overlaps E3 of replica 1; no operand production ever waits for a ring slot except the converters'.
// =====================================================================================================================
// L2 -> SM delivery (~30-38 B/clk per SM measured) bounds the kernel once the operands are out of shared memory, so fc_layer1's
// split image (64 KB, used twice per tile) stays resident; up_layer1 / up_layer2 k-blocks stream through a 3-slot ring.
constexpr int TS_NA = 3, TS_NW = 3;
constexpr int TS_W3_BYTES = 4 * W3_BYTES;
constexpr int TS_SMEM_BYTES = TS_NA * RAW_BYTES + TS_NW * W_BYTES + TS_W3_BYTES + 1024;

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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 activations of this thread's point -> tf32 hi part to 32 columns at hi_addr, remainder to 32 columns at lo_addr
__device__ __forceinline__ void stage_hi_lo(uint32_t hi_addr, uint32_t lo_addr, const uint32_t (&v)[32]) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float x = __uint_as_float(v[half * 16 + j]);
            const float h = to_tf32(x);
            hi[j] = __float_as_uint(h);
            lo[j] = __float_as_uint(x - h);
        }
        tmem_st16(hi_addr + half * 16, hi);
        tmem_st16(lo_addr + half * 16, lo);
    }
}

__global__ void __launch_bounds__(NUM_THREADS, 1) head_ts_kernel(const __grid_constant__ CUtensorMap xmap, const Args a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_A[TS_NA], empty_A[TS_NA], full_W[TS_NW], empty_W[TS_NW], full_S[2], empty_S[2];
    __shared__ uint64_t full_C[4][4];                       // [production phase E1(0), E2(0), E1(1), E2(1)][32-channel chunk]
    __shared__ uint64_t lo_free[4];                         // fc1(0) has consumed chunk ch: its lo columns may take replica 1's
    __shared__ uint64_t acc1_full, acc2_full[2], acc3_full[2], e3_done[2], w3_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias1_s[C1], wcode_s[C1], bias2_s[C2], bias3_s[C3], w4_s[3 * C3];
    __shared__ float b4_s[4], code_s[2];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
    const uint32_t sA = smem0, sW = smem0 + TS_NA * RAW_BYTES, sW3 = sW + TS_NW * W_BYTES;

    const int nkb1 = (a.cin + KB - 1) / KB;
    const int last_ksteps = ((a.cin - (nkb1 - 1) * KB) + 7) / 8;

    for (int i = threadIdx.x; i < C1; i += NUM_THREADS) {
        bias1_s[i] = a.bias1 ? a.bias1[i] : 0.f;
        wcode_s[i] = a.wfull[(size_t)i * a.w_stride + a.code_col];
        bias2_s[i] = a.bias2 ? a.bias2[i] : 0.f;
    }
    for (int i = threadIdx.x; i < C3; i += NUM_THREADS) bias3_s[i] = a.bias3 ? a.bias3[i] : 0.f;
    for (int i = threadIdx.x; i < 3 * C3; i += NUM_THREADS) w4_s[i] = a.w4[i];
    if (threadIdx.x < 4) b4_s[threadIdx.x] = (threadIdx.x < 3 && a.b4) ? a.b4[threadIdx.x] : 0.f;
    if (threadIdx.x < 2) code_s[threadIdx.x] = a.code[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int s = 0; s < TS_NA; ++s) { mbar_init(&full_A[s], 1); mbar_init(&empty_A[s], 4); }
        for (int s = 0; s < TS_NW; ++s) { mbar_init(&full_W[s], 1); mbar_init(&empty_W[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full_S[s], 4); mbar_init(&empty_S[s], 1);
            mbar_init(&acc2_full[s], 1); mbar_init(&acc3_full[s], 1);
            mbar_init(&e3_done[s], 4);
        }
        for (int p = 0; p < 4; ++p) {
            for (int c = 0; c < 4; ++c) mbar_init(&full_C[p][c], 4);
            mbar_init(&lo_free[p], 1);
        }
        mbar_init(&acc1_full, 1);
        mbar_init(&w3_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== activation TMA (no swizzle: the converters, not the tensor core, read this): [box q][channel][32 points] =====
        uint32_t slot = 0, phase = 0;
        for (long long t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
            const long long nt = t + gridDim.x;
            if (nt < a.ntiles)
                for (int i = lane; i < nkb1 * 4; i += 32) {
                    const long long box = nt * 4 + (i & 3);
                    if (box < a.nboxes) tma_prefetch_3d(&xmap, (int)(box % a.bpc) * 32, (i >> 2) * KB, (int)(box / a.bpc));
                }
            for (int kb = 0; kb < nkb1; ++kb) {
                mbar_wait(&empty_A[slot], phase ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_A[slot], (uint32_t)RAW_BYTES);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const long long box = t * 4 + q;
                        tma_load_3d(sA + slot * RAW_BYTES + q * 4096, &xmap, (int)(box % a.bpc) * 32, kb * KB, (int)(box / a.bpc), &full_A[slot]);
                    }
                }
                __syncwarp();
                if (++slot == TS_NA) { slot = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 2) {
        // ===== weights: fc_layer1's image once (resident); per tile the k-blocks W1[0..nkb1) | W2[0..4) | W2[0..4) =====
        uint32_t slot = 0, phase = 0;
        if (elect_one()) {
            mbar_arrive_expect_tx(&w3_full, (uint32_t)TS_W3_BYTES);
            for (int j = 0; j < 4; ++j) bulk_load(sW3 + j * W3_BYTES, a.w3 + (size_t)j * W3_BYTES, W3_BYTES, &w3_full);
        }
        __syncwarp();
        for (long long t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
            for (int j = 0; j < nkb1 + 8; ++j) {
                mbar_wait(&empty_W[slot], phase ^ 1u);
                if (elect_one()) {
                    const unsigned char *src = j < nkb1 ? a.w1 + (size_t)j * W_BYTES : a.w2 + (size_t)((j - nkb1) & 3) * W_BYTES;
                    mbar_arrive_expect_tx(&full_W[slot], (uint32_t)W_BYTES);
                    bulk_load(sW + slot * W_BYTES, src, W_BYTES, &full_W[slot]);
                }
                __syncwarp();
                if (++slot == TS_NW) { slot = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        uint32_t u = 0, wslot = 0, wphase = 0, it = 0;
        const bool tl = a.dbg != nullptr && blockIdx.x == 0;            // tuning: cycles this warp waits for weights / operands / epilogues
#define TL_WAIT(acc, bar, parity) do { if (tl) { const long long c0_ = clock64(); mbar_wait(bar, parity); acc += (unsigned int)(clock64() - c0_); } else mbar_wait(bar, parity); } while (0)
        for (long long t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
            const uint32_t par = it & 1u, prev = (it - 1u) & 1u;
            const uint32_t R = tmem_base + par * 256u, O = tmem_base + (par ^ 1u) * 256u, S = O + 128u;
            unsigned int wait_w = 0, wait_s = 0, wait_e = 0, wait_s1 = 0;
            if (lane == 0) tl_mark(a.dbg, 0, (int)it);
            // D1 columns R = last tile's O: its operands (h2 hi | lo) were consumed by MMAs issued before these
            for (int kb = 0; kb < nkb1; ++kb, ++u) {                    // ---- up1: main = hi.hi, correction = hi.lo + lo.hi
                const uint32_t ss = u & 1u;
                TL_WAIT(wait_w, &full_W[wslot], wphase);
                TL_WAIT(wait_s1, &full_S[ss], (u >> 1) & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = S + ss * 64u, wb = sW + wslot * W_BYTES;
                    const int nks = kb == nkb1 - 1 ? last_ksteps : KB / 8;
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint64_t bhl = smem_desc(wb + ks * 32, 16, 1024, LAYOUT_SW128);     // rows [W_hi; W_lo]
                        umma_tf32_ts(R, sa + ks * 8, bhl, idesc_n(IDESC_K, 2 * C1), (kb | ks) != 0);
                        umma_tf32_ts(R + C1, sa + 32 + ks * 8, bhl, idesc_n(IDESC_K, C1), 1u);
                    }
                    tc_commit(&empty_S[ss]);
                    tc_commit(&empty_W[wslot]);
                    if (kb == nkb1 - 1) tc_commit(&acc1_full);
                }
                __syncwarp();
                if (++wslot == TS_NW) { wslot = 0; wphase ^= 1u; }
            }
            if (lane == 0) tl_mark(a.dbg, 1, (int)it);
            if (it > 0) TL_WAIT(wait_e, &e3_done[1], prev);           // D2 columns O[0:128) = D3(1) of the previous tile: has been read
            else mbar_wait(&w3_full, 0);                              // fc_layer1's weights have landed (once)
            for (int g = 0; g < 2; ++g) {
                const uint32_t Ahi1 = g == 0 ? R + 128u : R;            // up2 operand: hi over D1's correction (g = 0) / over pre (g = 1)
                if (lane == 0) tl_mark(a.dbg, 2 + 3 * g, (int)it);
                for (int ch = 0; ch < C1 / KB; ++ch) {                  // ---- up2(g): one accumulator in O[0:128)
                    TL_WAIT(wait_w, &full_W[wslot], wphase);
                    TL_WAIT(wait_s, &full_C[2 * g][ch], par);
                    if (tl && lane == 0 && it == 3) a.dbg[23 * 512 + g * 8 + ch] = (unsigned int)clock64();
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t ahi = Ahi1 + ch * 32, alo = S + ch * 32, wb = sW + wslot * W_BYTES;
#pragma unroll
                        for (int ks = 0; ks < KB / 8; ++ks) {
                            const uint64_t bhi = smem_desc(wb + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint64_t blo = smem_desc(wb + C2 * 128 + ks * 32, 16, 1024, LAYOUT_SW128);
                            umma_tf32_ts(O, ahi + ks * 8, blo, idesc_n(IDESC_K, C2), (ch | ks) != 0);
                            umma_tf32_ts(O, alo + ks * 8, bhi, idesc_n(IDESC_K, C2), 1u);
                            umma_tf32_ts(O, ahi + ks * 8, bhi, idesc_n(IDESC_K, C2), 1u);
                        }
                        tc_commit(&empty_W[wslot]);
                        if (ch == C1 / KB - 1) tc_commit(&acc2_full[g]);
                    }
                    __syncwarp();
                    if (++wslot == TS_NW) { wslot = 0; wphase ^= 1u; }
                }
                if (lane == 0) tl_mark(a.dbg, 3 + 3 * g, (int)it);
                const uint32_t D3 = g == 0 ? R + 128u : R;              // over the up2 operand of this replica (consumed: issue order)
                for (int ch = 0; ch < C2 / KB; ++ch) {                  // ---- fc1(g): main | correction; weights resident
                    TL_WAIT(wait_s, &full_C[2 * g + 1][ch], par);
                    if (tl && lane == 0 && it == 3) a.dbg[23 * 512 + g * 8 + 4 + ch] = (unsigned int)clock64();
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t ahi = O + ch * 32, alo = S + ch * 32, wb = sW3 + ch * W3_BYTES;
#pragma unroll
                        for (int ks = 0; ks < KB / 8; ++ks) {
                            const uint64_t bhl = smem_desc(wb + ks * 32, 16, 1024, LAYOUT_SW128);   // rows [W_hi (64); W_lo (64)]
                            umma_tf32_ts(D3, ahi + ks * 8, bhl, idesc_n(IDESC_K, 2 * C3), (ch | ks) != 0);
                            umma_tf32_ts(D3 + C3, alo + ks * 8, bhl, idesc_n(IDESC_K, C3), 1u);
                        }
                        if (g == 0) tc_commit(&lo_free[ch]);
                        if (ch == C2 / KB - 1) tc_commit(&acc3_full[g]);
                    }
                    __syncwarp();
                }
                if (lane == 0) tl_mark(a.dbg, 4 + 3 * g, (int)it);
            }
            if (tl && lane == 0 && it < 512) {
                a.dbg[16 * 512 + it] = wait_w; a.dbg[17 * 512 + it] = wait_s1; a.dbg[18 * 512 + it] = wait_s; a.dbg[19 * 512 + it] = wait_e;
            }
        }
#undef TL_WAIT
    } else if (warp >= CONV_WARP0 && warp < EPI_WARP0) {
        // ===== converters (up1): this thread's point, 32 channels: raw fp32 -> hi / lo -> staging slot in TMEM =====
        const int q = warp & 3;
        const uint32_t lanebits = (uint32_t)(q * 32) << 16;
        uint32_t aslot = 0, aphase = 0, it = 0, u = 0;
        const bool tl = a.dbg != nullptr && blockIdx.x == 0 && warp == CONV_WARP0;
        for (long long t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
            const uint32_t S = tmem_base + ((it & 1u) ^ 1u) * 256u + 128u + lanebits;
            unsigned int wait_a = 0, wait_slot = 0;
            if (it > 0) {
                mbar_wait(&e3_done[0], (it - 1u) & 1u);                // this tile's staging columns = D3(0) of the previous tile
                tc_fence_after();
            }
            for (int kb = 0; kb < nkb1; ++kb, ++u) {
                const uint32_t ss = u & 1u;
                { const long long c0 = tl ? clock64() : 0; mbar_wait(&full_A[aslot], aphase); if (tl) wait_a += (unsigned int)(clock64() - c0); }
                const float *src = reinterpret_cast<const float *>(smem_gen + aslot * RAW_BYTES + q * 4096) + lane;
                uint32_t v[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(src[c * 32]);
                { const long long c0 = tl ? clock64() : 0; mbar_wait(&empty_S[ss], ((u >> 1) & 1u) ^ 1u); if (tl) wait_slot += (unsigned int)(clock64() - c0); }
                tc_fence_after();
                stage_hi_lo(S + ss * 64u, S + ss * 64u + 32u, v);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&full_S[ss]); mbar_arrive(&empty_A[aslot]); }
                if (++aslot == TS_NA) { aslot = 0; aphase ^= 1u; }
            }
            if (tl && lane == 0 && it < 512) { a.dbg[20 * 512 + it] = wait_a; a.dbg[21 * 512 + it] = wait_slot; }
        }
    } else if (warp >= EPI_WARP0) {
        // ===== epilogue groups: group g = replica g (E1, E2, E3 of that replica); warp quarter q = box q of the tile =====
        const int g = (warp - EPI_WARP0) >> 2, q = warp & 3;
        const uint32_t lanebits = (uint32_t)(q * 32) << 16;
        const float code_g = code_s[g];
        const float4 *bias1_4 = reinterpret_cast<const float4 *>(bias1_s), *wcode_4 = reinterpret_cast<const float4 *>(wcode_s);
        const float4 *bias2_4 = reinterpret_cast<const float4 *>(bias2_s), *bias3_4 = reinterpret_cast<const float4 *>(bias3_s);
        const float4 *w4_4 = reinterpret_cast<const float4 *>(w4_s);
        uint32_t it = 0;
        for (long long t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it) {
            const uint32_t par = it & 1u;
            const uint32_t R = tmem_base + par * 256u + lanebits, O = tmem_base + (par ^ 1u) * 256u + lanebits, S = O + 128u;
            const long long box = t * 4 + q;
            const long long bi = box / a.bpc;
            const int p = (int)(box % a.bpc) * 32 + lane;
            const bool valid = box < a.nboxes && p < a.n;
            const bool fine = a.dbg != nullptr && blockIdx.x == 0 && g == 0 && q == 0 && lane == 0 && it == 3;
#define FINE_MARK(idx) do { if (fine) a.dbg[22 * 512 + (idx)] = (unsigned int)clock64(); } while (0)
            // ---- E1(g): relu(pre + w_code * code[g]) -> operand of up2(g)
            if (g == 0) {
                mbar_wait(&acc1_full, par);
                tc_fence_after();
            }
            if (q == 0 && lane == 0) tl_mark(a.dbg, 8 + 3 * g, (int)it);
#pragma unroll 1
            for (int ch = 0; ch < C1 / KB; ++ch) {
                uint32_t v[32];
                FINE_MARK(ch * 5 + 0);
                if (g == 0) {
                    uint32_t vc[32];
                    tmem_ld32(R + ch * 32, v);
                    tmem_ld32(R + C1 + ch * 32, vc);
                    tmem_ld_wait();
                    FINE_MARK(ch * 5 + 1);
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 b = bias1_4[ch * 8 + j4];
                        v[4 * j4 + 0] = __float_as_uint((__uint_as_float(v[4 * j4 + 0]) + __uint_as_float(vc[4 * j4 + 0])) + b.x);
                        v[4 * j4 + 1] = __float_as_uint((__uint_as_float(v[4 * j4 + 1]) + __uint_as_float(vc[4 * j4 + 1])) + b.y);
                        v[4 * j4 + 2] = __float_as_uint((__uint_as_float(v[4 * j4 + 2]) + __uint_as_float(vc[4 * j4 + 2])) + b.z);
                        v[4 * j4 + 3] = __float_as_uint((__uint_as_float(v[4 * j4 + 3]) + __uint_as_float(vc[4 * j4 + 3])) + b.w);
                    }
                    {   // pre replaces D1's main columns: replica 1 reads it from there
                        uint32_t lo16[16], hi16[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) { lo16[j] = v[j]; hi16[j] = v[16 + j]; }
                        tmem_st16(R + ch * 32, lo16);
                        tmem_st16(R + ch * 32 + 16, hi16);
                    }
                } else {
                    mbar_wait(&full_C[0][ch], par);           // group 0 has written pre chunk ch
                    tc_fence_after();
                    tmem_ld32(R + ch * 32, v);
                    tmem_ld_wait();
                }
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 w = wcode_4[ch * 8 + j4];
                    v[4 * j4 + 0] = __float_as_uint(fmaxf(__fmaf_rn(w.x, code_g, __uint_as_float(v[4 * j4 + 0])), 0.f));
                    v[4 * j4 + 1] = __float_as_uint(fmaxf(__fmaf_rn(w.y, code_g, __uint_as_float(v[4 * j4 + 1])), 0.f));
                    v[4 * j4 + 2] = __float_as_uint(fmaxf(__fmaf_rn(w.z, code_g, __uint_as_float(v[4 * j4 + 2])), 0.f));
                    v[4 * j4 + 3] = __float_as_uint(fmaxf(__fmaf_rn(w.w, code_g, __uint_as_float(v[4 * j4 + 3])), 0.f));
                }
                if (g == 1) {
                    mbar_wait(&lo_free[ch], par);             // fc1(0) has consumed the lo columns of chunk ch
                    tc_fence_after();
                }
                // g = 0: hi over the correction chunk just read; g = 1: hi over the pre chunk just read; lo -> S
                FINE_MARK(ch * 5 + 2);
                stage_hi_lo((g == 0 ? R + C1 : R) + ch * 32, S + ch * 32, v);
                FINE_MARK(ch * 5 + 3);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_C[2 * g][ch]);
                FINE_MARK(ch * 5 + 4);
            }
            // ---- E2(g): relu(D2 + bias): hi over D2 in place, lo -> S (up2(g) is complete: its operands are dead)
            mbar_wait(&acc2_full[g], par);
            tc_fence_after();
            if (q == 0 && lane == 0) tl_mark(a.dbg, 9 + 3 * g, (int)it);
#pragma unroll 1
            for (int ch = 0; ch < C2 / KB; ++ch) {
                uint32_t v[32];
                FINE_MARK(20 + ch * 5 + 0);
                tmem_ld32(O + ch * 32, v);
                tmem_ld_wait();
                FINE_MARK(20 + ch * 5 + 1);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 b = bias2_4[ch * 8 + j4];
                    v[4 * j4 + 0] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j4 + 0]) + b.x, 0.f));
                    v[4 * j4 + 1] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j4 + 1]) + b.y, 0.f));
                    v[4 * j4 + 2] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j4 + 2]) + b.z, 0.f));
                    v[4 * j4 + 3] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j4 + 3]) + b.w, 0.f));
                }
                FINE_MARK(20 + ch * 5 + 2);
                stage_hi_lo(O + ch * 32, S + ch * 32, v);
                FINE_MARK(20 + ch * 5 + 3);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full_C[2 * g + 1][ch]);
                FINE_MARK(20 + ch * 5 + 4);
            }
            // ---- E3(g): relu(D3 + bias) in registers -> fc_layer2 + bias + residual
            mbar_wait(&acc3_full[g], par);
            tc_fence_after();
            if (q == 0 && lane == 0) tl_mark(a.dbg, 10 + 3 * g, (int)it);
            const uint32_t D3 = g == 0 ? R + 128u : R;
            float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
            for (int ch = 0; ch < C3 / 32; ++ch) {
                uint32_t v[32], vc[32];
                tmem_ld32(D3 + ch * 32, v);
                tmem_ld32(D3 + C3 + ch * 32, vc);
                tmem_ld_wait();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 b = bias3_4[ch * 8 + j4];
                    const float4 wa = w4_4[ch * 8 + j4], wb = w4_4[C3 / 4 + ch * 8 + j4], wc = w4_4[2 * (C3 / 4) + ch * 8 + j4];
                    const float bb[4] = {b.x, b.y, b.z, b.w}, w0[4] = {wa.x, wa.y, wa.z, wa.w}, w1[4] = {wb.x, wb.y, wb.z, wb.w}, w2[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float h = fmaxf((__uint_as_float(v[4 * j4 + e]) + __uint_as_float(vc[4 * j4 + e])) + bb[e], 0.f);
                        o0 = __fmaf_rn(w0[e], h, o0);
                        o1 = __fmaf_rn(w1[e], h, o1);
                        o2 = __fmaf_rn(w2[e], h, o2);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&e3_done[g]);
            if (valid) {
                const float o[3] = {o0 + b4_s[0], o1 + b4_s[1], o2 + b4_s[2]};
#pragma unroll
                for (int c3 = 0; c3 < 3; ++c3) {
                    float r = o[c3];
                    if (a.res) r += __ldg(a.res + bi * a.res_bstride + (size_t)c3 * a.n + p);
                    a.y[bi * a.y_bstride + (size_t)c3 * (2 * a.n) + 2 * p + g] = r;
                }
            }
            if (q == 0 && lane == 0) tl_mark(a.dbg, 14 + g, (int)it);
#undef FINE_MARK
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}



// ---------------------------------------------------------------------------------------------------------------------
// Version 3: the same dataflow on CTA PAIRS (tcgen05.mma.cta_group::2, M = 256 = two 128-point tiles, one per SM).
// What bounds version 2 is the delivery of weight k-blocks (L2 -> SM at ~25-38 B/clk per SM, profiles/r2/ncu_summary.md): every SM
// streams 676 KB per tile.  In a pair each CTA holds the weight rows of HALF of the output channels and the tensor cores of
// both SMs read both halves, so weight bytes per SM (L2 traffic, TMA writes and tensor-core operand reads) are halved.  The
// leader CTA (rank 0) issues every MMA and commits with a multicast arrive to both CTAs' barriers; everything a producer of
// either CTA has to tell the MMA warp is an arrive on the LEADER's barrier (remote arrive from the peer).
// ---------------------------------------------------------------------------------------------------------------------
#ifndef PU3_TS2_NA
#define PU3_TS2_NA 5
#endif
#ifndef PU3_TS2_NW
#define PU3_TS2_NW 5
#endif
constexpr int TS2_NA = PU3_TS2_NA;                      // raw activation ring: 5 x 16 KB (what is in flight bounds the L2 -> SM rate)
constexpr int TS2_NW = PU3_TS2_NW;                      // weight ring: 5 slots x [hi 64 rows | lo 64 rows] = 16 KB
constexpr int W2H_BYTES = W_BYTES / 2;         // this CTA's half of a 128-channel k-block
constexpr int W3H_BYTES = W3_BYTES / 2;        // ... of a fc_layer1 k-block: [hi 32 rows | lo 32 rows] = 8 KB
constexpr int W3R_BYTES = W3_BYTES / 2 + W3_BYTES / 4;   // resident fc_layer1 k-block of one CTA: 64 + 32 rows = 12 KB
constexpr int TS2_SMEM_BYTES = TS2_NA * RAW_BYTES + TS2_NW * W2H_BYTES + 4 * W3R_BYTES + 1024;
constexpr uint32_t IDESC2_K = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(256 >> 4) << 24);   // M = 256 over the pair

__device__ __forceinline__ void umma2_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
// completion of all MMAs issued so far -> arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit2(uint64_t *bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// arrive on `bar` of CTA `cta` of the cluster (local fast path)
__device__ __forceinline__ void arrive_at(uint64_t *bar, uint32_t cta, bool local) {
    if (local) { mbar_arrive(bar); return; }
    asm volatile(
        "{\n\t"
        ".reg .b32 rem;\n\t"
        "mapa.shared::cluster.u32 rem, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [rem];\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
// wait on a barrier that also receives arrivals from the peer CTA (cluster-scope acquire)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    unsigned long long t0 = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (!done && (spins & 255u) == 255u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) __trap();
        }
    }
}

// One 32-channel chunk of E1 for replica 0, straight from the up1 accumulators: pre = main + correction + bias -> over the main
// columns (replica 1 reads it there), relu(pre + w_code code[0]): hi -> over the correction columns just read, lo -> S.
// Ends with the stores complete and fenced; the caller arrives on the barriers.
__device__ __forceinline__ void e1_chunk_from_d1(uint32_t R0, uint32_t R1, uint32_t S, int ch, float code0, const float4 *bias1_4, const float4 *wcode_4) {
    uint32_t v[32];
    {
        uint32_t vc[32];
        tmem_ld32(R0 + ch * 32, v);
        tmem_ld32(R1 + ch * 32, vc);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b = bias1_4[ch * 8 + j4];
            v[4 * j4 + 0] = __float_as_uint((__uint_as_float(v[4 * j4 + 0]) + __uint_as_float(vc[4 * j4 + 0])) + b.x);
            v[4 * j4 + 1] = __float_as_uint((__uint_as_float(v[4 * j4 + 1]) + __uint_as_float(vc[4 * j4 + 1])) + b.y);
            v[4 * j4 + 2] = __float_as_uint((__uint_as_float(v[4 * j4 + 2]) + __uint_as_float(vc[4 * j4 + 2])) + b.z);
            v[4 * j4 + 3] = __float_as_uint((__uint_as_float(v[4 * j4 + 3]) + __uint_as_float(vc[4 * j4 + 3])) + b.w);
        }
    }
    {
        uint32_t lo16[16], hi16[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) { lo16[j] = v[j]; hi16[j] = v[16 + j]; }
        tmem_st16(R0 + ch * 32, lo16);
        tmem_st16(R0 + ch * 32 + 16, hi16);
    }
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 w = wcode_4[ch * 8 + j4];
        v[4 * j4 + 0] = __float_as_uint(fmaxf(__fmaf_rn(w.x, code0, __uint_as_float(v[4 * j4 + 0])), 0.f));
        v[4 * j4 + 1] = __float_as_uint(fmaxf(__fmaf_rn(w.y, code0, __uint_as_float(v[4 * j4 + 1])), 0.f));
        v[4 * j4 + 2] = __float_as_uint(fmaxf(__fmaf_rn(w.z, code0, __uint_as_float(v[4 * j4 + 2])), 0.f));
        v[4 * j4 + 3] = __float_as_uint(fmaxf(__fmaf_rn(w.w, code0, __uint_as_float(v[4 * j4 + 3])), 0.f));
    }
    stage_hi_lo(R1 + ch * 32, S + ch * 32, v);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
}
// E1 values of replica 1 for one chunk: relu(pre + w_code code[1]) into registers (staged later, when the columns are free)
__device__ __forceinline__ void e1_values_from_pre(uint32_t R, int ch, float code1, const float4 *wcode_4, uint32_t (&v)[32]) {
    tmem_ld32(R + ch * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 w = wcode_4[ch * 8 + j4];
        v[4 * j4 + 0] = __float_as_uint(fmaxf(__fmaf_rn(w.x, code1, __uint_as_float(v[4 * j4 + 0])), 0.f));
        v[4 * j4 + 1] = __float_as_uint(fmaxf(__fmaf_rn(w.y, code1, __uint_as_float(v[4 * j4 + 1])), 0.f));
        v[4 * j4 + 2] = __float_as_uint(fmaxf(__fmaf_rn(w.z, code1, __uint_as_float(v[4 * j4 + 2])), 0.f));
        v[4 * j4 + 3] = __float_as_uint(fmaxf(__fmaf_rn(w.w, code1, __uint_as_float(v[4 * j4 + 3])), 0.f));
    }
}
// One chunk of E2: relu(D2 + bias): hi over the D2 chunk in place, lo -> S once `free_bar` says the MMAs that read the previous
// content of those columns have completed
__device__ __forceinline__ void e2_chunk(uint32_t D2, uint32_t S, int ch, const float4 *bias2_4, uint64_t *free_bar, uint32_t parity) {
    uint32_t v[32];
    tmem_ld32(D2 + ch * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 b = bias2_4[ch * 8 + j4];
        v[4 * j4 + 0] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j4 + 0]) + b.x, 0.f));
        v[4 * j4 + 1] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j4 + 1]) + b.y, 0.f));
        v[4 * j4 + 2] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j4 + 2]) + b.z, 0.f));
        v[4 * j4 + 3] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j4 + 3]) + b.w, 0.f));
    }
    mbar_wait(free_bar, parity);
    tc_fence_after();
    stage_hi_lo(D2 + ch * 32, S + ch * 32, v);
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
}

__global__ void __launch_bounds__(NUM_THREADS, 1) head_ts2_kernel(const __grid_constant__ CUtensorMap xmap, const Args a) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t full_A[TS2_NA], empty_A[TS2_NA], full_W[TS2_NW], empty_W[TS2_NW], full_S[2], empty_S[2];
    __shared__ uint64_t full_C[4][4];                       // [production phase E1(0), E2(0), E1(1), E2(1)][32-channel chunk]
    __shared__ uint64_t s_free[4][4];                       // [up2(0), up2(1), fc1(0), fc1(1)][chunk]: consumed, its lo columns may be rewritten
    __shared__ uint64_t acc1_full, acc2_full[2], acc3_full[2], e3_done[2], w3_full;
    __shared__ uint64_t full_Wl[TS2_NW], w3l_full, pre_ok[4];   // local: this CTA's weight copies have landed; pre chunk written
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias1_s[C1], wcode_s[C1], bias2_s[C2], bias3_s[C3], w4_s[3 * C3];
    __shared__ float b4_s[4], code_s[2];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
    const uint32_t sA = smem0, sW = smem0 + TS2_NA * RAW_BYTES, sW3 = sW + TS2_NW * W2H_BYTES;
    const uint32_t rank = cluster_ctarank();                 // 0 = leader (issues every MMA of the pair), 1 = peer
    const bool leader = rank == 0;

    const int nkb1 = (a.cin + KB - 1) / KB;
    const int last_ksteps = ((a.cin - (nkb1 - 1) * KB) + 7) / 8;
    // both CTAs of a pair run the same number of tiles (the leader issues the MMAs of both): tiles >= ntiles are all padding
    const long long tiles_end = (a.ntiles + gridDim.x - 1) / gridDim.x * gridDim.x;

    for (int i = threadIdx.x; i < C1; i += NUM_THREADS) {
        bias1_s[i] = a.bias1 ? a.bias1[i] : 0.f;
        wcode_s[i] = a.wfull[(size_t)i * a.w_stride + a.code_col];
        bias2_s[i] = a.bias2 ? a.bias2[i] : 0.f;
    }
    for (int i = threadIdx.x; i < C3; i += NUM_THREADS) bias3_s[i] = a.bias3 ? a.bias3[i] : 0.f;
    for (int i = threadIdx.x; i < 3 * C3; i += NUM_THREADS) w4_s[i] = a.w4[i];
    if (threadIdx.x < 4) b4_s[threadIdx.x] = (threadIdx.x < 3 && a.b4) ? a.b4[threadIdx.x] : 0.f;
    if (threadIdx.x < 2) code_s[threadIdx.x] = a.code[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int s = 0; s < TS2_NA; ++s) { mbar_init(&full_A[s], 1); mbar_init(&empty_A[s], 4); }
        for (int s = 0; s < TS2_NW; ++s) { mbar_init(&full_W[s], 2); mbar_init(&empty_W[s], 1); mbar_init(&full_Wl[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full_S[s], 8); mbar_init(&empty_S[s], 1);
            mbar_init(&acc2_full[s], 1); mbar_init(&acc3_full[s], 1);
            mbar_init(&e3_done[s], 8);                       // the leader's MMA warp waits for both CTAs' E3 of replica s
        }
        for (int p = 0; p < 4; ++p) {
            for (int c = 0; c < 4; ++c) mbar_init(&full_C[p][c], 8);
            for (int c = 0; c < 4; ++c) mbar_init(&s_free[p][c], 1);
            mbar_init(&pre_ok[p], 4);
        }
        mbar_init(&acc1_full, 1);
        mbar_init(&w3_full, 2);
        mbar_init(&w3l_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_arrive_release();                                // both CTAs' barriers are initialised before anything arrives remotely
    cluster_wait_acquire();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== activation TMA (no swizzle: the converters, not the tensor core, read this): [box q][channel][32 points] =====
        uint32_t slot = 0, phase = 0;
        for (long long t = blockIdx.x; t < tiles_end; t += gridDim.x) {
            const long long nt = t + gridDim.x;
            if (nt < a.ntiles)
                for (int i = lane; i < nkb1 * 4; i += 32) {
                    const long long box = nt * 4 + (i & 3);
                    if (box < a.nboxes) tma_prefetch_3d(&xmap, (int)(box % a.bpc) * 32, (i >> 2) * KB, (int)(box / a.bpc));
                }
            for (int kb = 0; kb < nkb1; ++kb) {
                mbar_wait(&empty_A[slot], phase ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_A[slot], (uint32_t)RAW_BYTES);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const long long box = t * 4 + q;
                        tma_load_3d(sA + slot * RAW_BYTES + q * 4096, &xmap, (int)(box % a.bpc) * 32, kb * KB, (int)(box / a.bpc), &full_A[slot]);
                    }
                }
                __syncwarp();
                if (++slot == TS2_NA) { slot = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 2) {
        // ===== weights, this CTA's half of the output channels (rows 64 rank .. of the hi and of the lo image): fc_layer1 once
        // (resident, [k-block][hi 32 rows | lo 32 rows]); per tile the k-blocks W1[0..nkb1) | W2[0..4) | W2[0..4) as [hi 64 | lo 64] =====
        uint32_t slot = 0, phase = 0;
        if (elect_one()) {
            // fc_layer1 (64 output channels), per k-block: X0 = 64 rows: W_hi in the leader, W_lo in the peer -> ONE N = 128 instruction
            // A_hi . [W_hi; W_lo] fills main | correction; X1 = 32 rows: W_hi(0:32) in the leader, W_hi(32:64) in the peer -> N = 64 for A_lo . W_hi
            mbar_arrive_expect_tx(&w3l_full, (uint32_t)(4 * W3R_BYTES));
            for (int j = 0; j < 4; ++j) {
                bulk_load(sW3 + j * W3R_BYTES, a.w3 + (size_t)j * W3_BYTES + rank * (W3_BYTES / 2), W3_BYTES / 2, &w3l_full);
                bulk_load(sW3 + j * W3R_BYTES + W3_BYTES / 2, a.w3 + (size_t)j * W3_BYTES + rank * (W3_BYTES / 4), W3_BYTES / 4, &w3l_full);
            }
        }
        __syncwarp();
        for (long long t = blockIdx.x; t < tiles_end; t += gridDim.x) {
            for (int j = 0; j < nkb1 + 8; ++j) {
                mbar_wait(&empty_W[slot], phase ^ 1u);
                if (elect_one()) {
                    const unsigned char *src = j < nkb1 ? a.w1 + (size_t)j * W_BYTES : a.w2 + (size_t)((j - nkb1) & 3) * W_BYTES;
                    mbar_arrive_expect_tx(&full_Wl[slot], (uint32_t)W2H_BYTES);
                    bulk_load(sW + slot * W2H_BYTES, src + rank * (W2H_BYTES / 2), W2H_BYTES / 2, &full_Wl[slot]);
                    bulk_load(sW + slot * W2H_BYTES + W2H_BYTES / 2, src + W_BYTES / 2 + rank * (W2H_BYTES / 2), W2H_BYTES / 2, &full_Wl[slot]);
                }
                __syncwarp();
                if (++slot == TS2_NW) { slot = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 3) {
        // ===== forwarder: "this CTA's copy of weight block j has landed" -> the leader's full_W (the MMA needs both halves) =====
        uint32_t slot = 0, phase = 0;
        mbar_wait(&w3l_full, 0);
        if (lane == 0) arrive_at(&w3_full, 0u, leader);
        for (long long t = blockIdx.x; t < tiles_end; t += gridDim.x) {
            for (int j = 0; j < nkb1 + 8; ++j) {
                mbar_wait(&full_Wl[slot], phase);
                if (lane == 0) arrive_at(&full_W[slot], 0u, leader);
                __syncwarp();
                if (++slot == TS2_NW) { slot = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1 && leader) {
        // ===== MMA issuer (leader CTA only).  Order per tile: up1 | up2(0) | up2(1) | fc1(0) | fc1(1): the epilogue that turns an
        // accumulator into the next operand (E2 of replica 0 / 1) runs while the MMAs of the OTHER replica execute =====
        uint32_t wslot = 0, wphase = 0, it = 0, cnt[2] = {0, 0};
        const uint32_t R0 = tmem_base, R1 = tmem_base + 128u, O0 = tmem_base + 256u, S = tmem_base + 384u;
        const bool tl = a.dbg != nullptr && blockIdx.x == 0;            // tuning: cycles this warp waits for weights / operands / epilogues
#define TL_WAIT(acc, bar, parity) do { if (tl) { const long long c0_ = clock64(); mbar_wait_cluster(bar, parity); acc += (unsigned int)(clock64() - c0_); } else mbar_wait_cluster(bar, parity); } while (0)
        for (long long t = blockIdx.x; t < tiles_end; t += gridDim.x, ++it) {
            const uint32_t par = it & 1u, prev = (it - 1u) & 1u;
            unsigned int wait_w = 0, wait_s = 0, wait_e = 0, wait_s1 = 0;
            if (it > 0) TL_WAIT(wait_e, &e3_done[0], prev);           // R0 = D3(0) of the previous tile: has been read
            if (lane == 0) tl_mark(a.dbg, 0, (int)it);
            for (int kb = 0; kb < nkb1; ++kb) {                         // ---- up1: main (R0) = hi.hi, correction (R1) = hi.lo + lo.hi
                const uint32_t ss = kb & 1u;
                TL_WAIT(wait_w, &full_W[wslot], wphase);
                TL_WAIT(wait_s1, &full_S[ss], cnt[ss] & 1u);
                ++cnt[ss];
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = S + ss * 64u, wb = sW + wslot * W2H_BYTES;
                    const int nks = kb == nkb1 - 1 ? last_ksteps : KB / 8;
                    for (int ks = 0; ks < nks; ++ks) {    // each CTA supplies its 64 rows of W_hi / W_lo: N = 128 per instruction
                        const uint64_t bhi = smem_desc(wb + ks * 32, 16, 1024, LAYOUT_SW128);
                        const uint64_t blo = smem_desc(wb + W2H_BYTES / 2 + ks * 32, 16, 1024, LAYOUT_SW128);
                        umma2_tf32_ts(R0, sa + ks * 8, bhi, idesc_n(IDESC2_K, C1), (kb | ks) != 0);
                        umma2_tf32_ts(R1, sa + ks * 8, blo, idesc_n(IDESC2_K, C1), (kb | ks) != 0);
                        umma2_tf32_ts(R1, sa + 32 + ks * 8, bhi, idesc_n(IDESC2_K, C1), 1u);
                    }
                    tc_commit2(&empty_S[ss]);
                    tc_commit2(&empty_W[wslot]);
                    if (kb == nkb1 - 1) tc_commit2(&acc1_full);
                }
                __syncwarp();
                if (++wslot == TS2_NW) { wslot = 0; wphase ^= 1u; }
            }
            if (lane == 0) tl_mark(a.dbg, 1, (int)it);
            if (it > 0) TL_WAIT(wait_e, &e3_done[1], prev);           // O0 = D3(1) of the previous tile: has been read
            else mbar_wait_cluster(&w3_full, 0);                      // fc_layer1's weights have landed in both CTAs (once)
            for (int g = 0; g < 2; ++g) {                               // ---- up2(g): one accumulator; g = 0 -> O0, g = 1 -> R1 (hi1(0) consumed)
                const uint32_t Ahi = g == 0 ? R1 : R0, D2 = g == 0 ? O0 : R1;
                if (lane == 0) tl_mark(a.dbg, 2 + 2 * g, (int)it);
                for (int ch = 0; ch < C1 / KB; ++ch) {
                    TL_WAIT(wait_w, &full_W[wslot], wphase);
                    TL_WAIT(wait_s, &full_C[2 * g][ch], par);
                    if (tl && lane == 0 && it == 3) a.dbg[23 * 512 + g * 8 + ch] = (unsigned int)clock64();
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t ahi = Ahi + ch * 32, alo = S + ch * 32, wb = sW + wslot * W2H_BYTES;
#pragma unroll
                        for (int ks = 0; ks < KB / 8; ++ks) {
                            const uint64_t bhi = smem_desc(wb + ks * 32, 16, 1024, LAYOUT_SW128);
                            const uint64_t blo = smem_desc(wb + W2H_BYTES / 2 + ks * 32, 16, 1024, LAYOUT_SW128);
                            umma2_tf32_ts(D2, ahi + ks * 8, blo, idesc_n(IDESC2_K, C2), (ch | ks) != 0);
                            umma2_tf32_ts(D2, alo + ks * 8, bhi, idesc_n(IDESC2_K, C2), 1u);
                            umma2_tf32_ts(D2, ahi + ks * 8, bhi, idesc_n(IDESC2_K, C2), 1u);
                        }
                        tc_commit2(&empty_W[wslot]);
                        tc_commit2(&s_free[g][ch]);                  // the lo columns of chunk ch may take the next operand
                        if (ch == C1 / KB - 1) tc_commit2(&acc2_full[g]);
                    }
                    __syncwarp();
                    if (++wslot == TS2_NW) { wslot = 0; wphase ^= 1u; }
                }
                if (lane == 0) tl_mark(a.dbg, 3 + 2 * g, (int)it);
            }
            for (int g = 0; g < 2; ++g) {                               // ---- fc1(g): main | correction; weights resident
                // g = 0: operand hi over D2(0) in O0, D3 -> R0 (hi1(1) consumed); g = 1: operand hi over D2(1) in R1, D3 -> O0 (hi2(0) consumed)
                const uint32_t Ahi = g == 0 ? O0 : R1, D3 = g == 0 ? R0 : O0;
                for (int ch = 0; ch < C2 / KB; ++ch) {
                    TL_WAIT(wait_s, &full_C[2 * g + 1][ch], par);
                    if (tl && lane == 0 && it == 3) a.dbg[23 * 512 + g * 8 + 4 + ch] = (unsigned int)clock64();
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t ahi = Ahi + ch * 32, alo = S + ch * 32, wb = sW3 + ch * W3R_BYTES;
#pragma unroll
                        for (int ks = 0; ks < KB / 8; ++ks) {   // an MMA costs >= ~70 cycles whatever its N (A read from TMEM): 2 instead of 3 per k-step
                            const uint64_t bhl = smem_desc(wb + ks * 32, 16, 1024, LAYOUT_SW128);                    // pair: [W_hi (64); W_lo (64)]
                            const uint64_t bh2 = smem_desc(wb + W3_BYTES / 2 + ks * 32, 16, 1024, LAYOUT_SW128);     // pair: W_hi (32 + 32)
                            umma2_tf32_ts(D3, ahi + ks * 8, bhl, idesc_n(IDESC2_K, 2 * C3), (ch | ks) != 0);
                            umma2_tf32_ts(D3 + C3, alo + ks * 8, bh2, idesc_n(IDESC2_K, C3), 1u);
                        }
                        tc_commit2(&s_free[2 + g][ch]);
                        if (ch == C2 / KB - 1) tc_commit2(&acc3_full[g]);
                    }
                    __syncwarp();
                }
                if (lane == 0) tl_mark(a.dbg, 6 + g, (int)it);
            }
            if (tl && lane == 0 && it < 512) {
                a.dbg[16 * 512 + it] = wait_w; a.dbg[17 * 512 + it] = wait_s1; a.dbg[18 * 512 + it] = wait_s; a.dbg[19 * 512 + it] = wait_e;
            }
        }
#undef TL_WAIT
    } else if (warp >= CONV_WARP0 && warp < EPI_WARP0) {
        // ===== converters (up1): this thread's point, 32 channels: raw fp32 -> hi / lo -> staging slot in TMEM (slot = k-block & 1:
        // columns S[0:64) / S[64:128), i.e. the lo columns of chunks 0-1 / 2-3).  Afterwards they help the epilogue groups where one
        // warp per lane quarter is too slow for the MMAs it feeds: chunks 1 and 3 of E1(0) and of E2(1) =====
        const int q = warp & 3;
        const uint32_t lanebits = (uint32_t)(q * 32) << 16;
        const uint32_t R0 = tmem_base + lanebits, R1 = R0 + 128u, S = R0 + 384u;
        const float4 *bias1_4 = reinterpret_cast<const float4 *>(bias1_s), *wcode_4 = reinterpret_cast<const float4 *>(wcode_s);
        const float4 *bias2_4 = reinterpret_cast<const float4 *>(bias2_s);
        const float code0 = code_s[0];
        uint32_t aslot = 0, aphase = 0, it = 0, cnt[2] = {0, 0};
        const bool tl = a.dbg != nullptr && blockIdx.x == 0 && warp == CONV_WARP0;
        for (long long t = blockIdx.x; t < tiles_end; t += gridDim.x, ++it) {
            const uint32_t par = it & 1u;
            unsigned int wait_a = 0, wait_slot = 0;
            for (int kb = 0; kb < nkb1; ++kb) {
                const uint32_t ss = kb & 1u;
                { const long long c0 = tl ? clock64() : 0; mbar_wait(&full_A[aslot], aphase); if (tl) wait_a += (unsigned int)(clock64() - c0); }
                const float *src = reinterpret_cast<const float *>(smem_gen + aslot * RAW_BYTES + q * 4096) + lane;
                uint32_t v[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(src[c * 32]);
                { const long long c0 = tl ? clock64() : 0;
                  if (it > 0 && kb < 2) mbar_wait(&s_free[3][2 * kb + 1], (it - 1u) & 1u);   // fc1(1) of the previous tile has consumed these lo columns
                  mbar_wait(&empty_S[ss], (cnt[ss] & 1u) ^ 1u);
                  if (tl) wait_slot += (unsigned int)(clock64() - c0); }
                ++cnt[ss];
                tc_fence_after();
                stage_hi_lo(S + ss * 64u, S + ss * 64u + 32u, v);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { arrive_at(&full_S[ss], 0u, leader); mbar_arrive(&empty_A[aslot]); }
                if (++aslot == TS2_NA) { aslot = 0; aphase ^= 1u; }
            }
            if (tl && lane == 0 && it < 512) { a.dbg[20 * 512 + it] = wait_a; a.dbg[21 * 512 + it] = wait_slot; }
            mbar_wait(&acc1_full, par);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 1; ch < C1 / KB; ch += 2) {
                e1_chunk_from_d1(R0, R1, S, ch, code0, bias1_4, wcode_4);
                if (lane == 0) { mbar_arrive(&pre_ok[ch]); arrive_at(&full_C[0][ch], 0u, leader); }
            }
            mbar_wait(&acc2_full[1], par);
            tc_fence_after();
#pragma unroll 1
            for (int ch = 1; ch < C2 / KB; ch += 2) {
                e2_chunk(R1, S, ch, bias2_4, &s_free[2][ch], par);
                if (lane == 0) arrive_at(&full_C[3][ch], 0u, leader);
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===== epilogue groups: group g = replica g (E1, E2, E3 of that replica); warp quarter q = box q of the tile =====
        const int g = (warp - EPI_WARP0) >> 2, q = warp & 3;
        const uint32_t lanebits = (uint32_t)(q * 32) << 16;
        const uint32_t R0 = tmem_base + lanebits, R1 = R0 + 128u, O0 = R0 + 256u, S = R0 + 384u;
        const float code_g = code_s[g];
        const float4 *bias1_4 = reinterpret_cast<const float4 *>(bias1_s), *wcode_4 = reinterpret_cast<const float4 *>(wcode_s);
        const float4 *bias2_4 = reinterpret_cast<const float4 *>(bias2_s), *bias3_4 = reinterpret_cast<const float4 *>(bias3_s);
        const float4 *w4_4 = reinterpret_cast<const float4 *>(w4_s);
        uint32_t it = 0;
        for (long long t = blockIdx.x; t < tiles_end; t += gridDim.x, ++it) {
            const uint32_t par = it & 1u;
            const long long box = t * 4 + q;
            const long long bi = box / a.bpc;
            const int p = (int)(box % a.bpc) * 32 + lane;
            const bool valid = box < a.nboxes && p < a.n;
            if (g == 0) {
                // ---- E1(0), chunks 0 and 2 (1 and 3: converter warps): pre over D1's main columns, hi over the correction columns, lo -> S
                mbar_wait(&acc1_full, par);
                tc_fence_after();
                if (q == 0 && lane == 0) tl_mark(a.dbg, 8, (int)it);
#pragma unroll 1
                for (int ch = 0; ch < C1 / KB; ch += 2) {
                    e1_chunk_from_d1(R0, R1, S, ch, code_g, bias1_4, wcode_4);
                    if (lane == 0) { mbar_arrive(&pre_ok[ch]); arrive_at(&full_C[0][ch], 0u, leader); }
                }
                // ---- E2(0) (while up2(1) runs): relu(D2(0) + bias): hi over D2(0) in place, lo -> S as up2(1) releases the columns
                mbar_wait(&acc2_full[0], par);
                tc_fence_after();
                if (q == 0 && lane == 0) tl_mark(a.dbg, 9, (int)it);
#pragma unroll 1
                for (int ch = 0; ch < C2 / KB; ++ch) {
                    e2_chunk(O0, S, ch, bias2_4, &s_free[1][ch], par);
                    if (lane == 0) arrive_at(&full_C[1][ch], 0u, leader);
                }
            } else {
                // ---- E1(1) (while up2(0) runs): values from pre one chunk ahead; hi over pre in place, lo -> S as up2(0) releases the columns
                if (q == 0 && lane == 0) tl_mark(a.dbg, 11, (int)it);
                uint32_t v[32];
                mbar_wait(&pre_ok[0], par);
                tc_fence_after();
                e1_values_from_pre(R0, 0, code_g, wcode_4, v);
#pragma unroll
                for (int ch = 0; ch < C1 / KB; ++ch) {
                    uint32_t vn[32];
                    if (ch + 1 < C1 / KB) {
                        mbar_wait(&pre_ok[ch + 1], par);
                        tc_fence_after();
                        e1_values_from_pre(R0, ch + 1, code_g, wcode_4, vn);
                    }
                    mbar_wait(&s_free[0][ch], par);           // up2(0) has consumed the lo columns of chunk ch
                    tc_fence_after();
                    stage_hi_lo(R0 + ch * 32, S + ch * 32, v);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_at(&full_C[2][ch], 0u, leader);
                    if (ch + 1 < C1 / KB) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = vn[j];
                    }
                }
                // ---- E2(1), chunks 0 and 2 (1 and 3: converter warps; while fc1(0) runs): hi over D2(1) in place (R1), lo -> S
                mbar_wait(&acc2_full[1], par);
                tc_fence_after();
                if (q == 0 && lane == 0) tl_mark(a.dbg, 12, (int)it);
#pragma unroll 1
                for (int ch = 0; ch < C2 / KB; ch += 2) {
                    e2_chunk(R1, S, ch, bias2_4, &s_free[2][ch], par);
                    if (lane == 0) arrive_at(&full_C[3][ch], 0u, leader);
                }
            }
            // ---- E3(g): relu(D3 + bias) in registers -> fc_layer2 + bias + residual
            mbar_wait(&acc3_full[g], par);
            tc_fence_after();
            if (q == 0 && lane == 0) tl_mark(a.dbg, 10 + 3 * g, (int)it);
            const uint32_t D3 = g == 0 ? R0 : O0;
            float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
            for (int ch = 0; ch < C3 / 32; ++ch) {
                uint32_t v[32], vc[32];
                tmem_ld32(D3 + ch * 32, v);
                tmem_ld32(D3 + C3 + ch * 32, vc);
                tmem_ld_wait();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 b = bias3_4[ch * 8 + j4];
                    const float4 wa = w4_4[ch * 8 + j4], wb = w4_4[C3 / 4 + ch * 8 + j4], wc = w4_4[2 * (C3 / 4) + ch * 8 + j4];
                    const float bb[4] = {b.x, b.y, b.z, b.w}, w0[4] = {wa.x, wa.y, wa.z, wa.w}, w1[4] = {wb.x, wb.y, wb.z, wb.w}, w2[4] = {wc.x, wc.y, wc.z, wc.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float h = fmaxf((__uint_as_float(v[4 * j4 + e]) + __uint_as_float(vc[4 * j4 + e])) + bb[e], 0.f);
                        o0 = __fmaf_rn(w0[e], h, o0);
                        o1 = __fmaf_rn(w1[e], h, o1);
                        o2 = __fmaf_rn(w2[e], h, o2);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_at(&e3_done[g], 0u, leader);
            if (valid) {
                const float o[3] = {o0 + b4_s[0], o1 + b4_s[1], o2 + b4_s[2]};
#pragma unroll
                for (int c3 = 0; c3 < 3; ++c3) {
                    float r = o[c3];
                    if (a.res) r += __ldg(a.res + bi * a.res_bstride + (size_t)c3 * a.n + p);
                    a.y[bi * a.y_bstride + (size_t)c3 * (2 * a.n) + 2 * p + g] = r;
                }
            }
            if (q == 0 && lane == 0) tl_mark(a.dbg, 14 + g, (int)it);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_arrive_release();                                // no CTA leaves while its partner may still signal it or read its shared memory
    cluster_wait_acquire();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


static int g_mode = 2;    // 2 = CTA pairs (head_ts2_kernel), 1 = activation operands in TMEM (head_ts_kernel), 0 = in shared memory (head_tc_kernel)
static unsigned int *g_dbg = nullptr;

}  // namespace head
}  // namespace tc
}  // namespace pu3

using namespace pu3;

extern "C" void pu3_head_tc_set_debug(void *buf) { tc::head::g_dbg = static_cast<unsigned int *>(buf); }
extern "C" void pu3_head_tc_set_mode(int mode) { tc::head::g_mode = mode; }

extern "C" int pu3_head_tc_f32(int b, int n, int cin, const float *x, long long x_bstride, const void *ws1, const void *ws2,
                               const void *ws3, const float *w1, int w1_stride, int code_col, const float *b1, const float *code,
                               const float *b2, const float *b3, const float *w4, const float *b4, const float *res,
                               long long res_bstride, float *y, long long y_bstride, pu3_stream_t stream) {
    namespace H = tc::head;
    PU3_ARG_CHECK(b >= 0 && n >= 0 && cin > 0, "head_tc: bad size b=%d n=%d cin=%d", b, n, cin);
    if (b == 0 || n == 0) return PU3_OK;
    PU3_ARG_CHECK(x && ws1 && ws2 && ws3 && w1 && code && w4 && y, "head_tc: null pointer");
    PU3_ARG_CHECK(n % 4 == 0 && x_bstride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                  "head_tc: the TMA path needs n %% 4 == 0 and 16-byte aligned slices (n=%d)", n);
    PU3_ARG_CHECK(((reinterpret_cast<uintptr_t>(ws1) | reinterpret_cast<uintptr_t>(ws2) | reinterpret_cast<uintptr_t>(ws3)) & 15) == 0,
                  "head_tc: split weights must be 16-byte aligned");
    tc::EncodeTiledFn enc = tc::encode_fn();
    if (!enc) { set_error("head_tc: cuTensorMapEncodeTiled not available"); return PU3_E_ARG; }
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)n, (cuuint64_t)cin, (cuuint64_t)b};
    const cuuint64_t gstride[2] = {(cuuint64_t)n * 4, (cuuint64_t)x_bstride * 4};
    const cuuint32_t box[3] = {32, H::KB, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const bool ts = H::g_mode != 0;     // TS form: the converters read the landing zone, so it needs no swizzle
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(x), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, ts ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("head_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return PU3_E_ARG; }
    H::Args a{};
    a.b = b; a.n = n; a.cin = cin; a.bpc = (n + 31) / 32;
    a.nboxes = (long long)b * a.bpc;
    a.ntiles = (a.nboxes + 3) / 4;
    a.w1 = static_cast<const unsigned char *>(ws1); a.w2 = static_cast<const unsigned char *>(ws2); a.w3 = static_cast<const unsigned char *>(ws3);
    a.bias1 = b1; a.wfull = w1; a.w_stride = w1_stride; a.code_col = code_col; a.code = code;
    a.bias2 = b2; a.bias3 = b3; a.w4 = w4; a.b4 = b4;
    a.res = res; a.res_bstride = res_bstride; a.y = y; a.y_bstride = y_bstride;
    a.dbg = H::g_dbg;
    const int grid = (int)(a.ntiles < device_info().sm_count ? a.ntiles : device_info().sm_count);
    if (H::g_mode == 2) {
        int st = cuda_status(cudaFuncSetAttribute(H::head_ts2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H::TS2_SMEM_BYTES),
                             "head_tc: shared memory opt-in");
        if (st) return st;
        const int sms = device_info().sm_count & ~1;
        const long long want = (a.ntiles + 1) & ~1LL;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(want < sms ? want : sms));
        cfg.blockDim = dim3((unsigned)H::NUM_THREADS);
        cfg.dynamicSmemBytes = H::TS2_SMEM_BYTES;
        cfg.stream = as_stream(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cuda_status(cudaLaunchKernelEx(&cfg, H::head_ts2_kernel, map, a), "head_ts2_kernel launch");
    }
    if (ts) {
        int st = cuda_status(cudaFuncSetAttribute(H::head_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H::TS_SMEM_BYTES),
                             "head_tc: shared memory opt-in");
        if (st) return st;
        H::head_ts_kernel<<<grid, H::NUM_THREADS, H::TS_SMEM_BYTES, as_stream(stream)>>>(map, a);
        PU3_LAUNCH_CHECK("head_ts_kernel");
        return PU3_OK;
    }
    {
        int st = cuda_status(cudaFuncSetAttribute(H::head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H::SMEM_BYTES),
                             "head_tc: shared memory opt-in");
        if (st) return st;
    }
    H::head_tc_kernel<<<grid, H::NUM_THREADS, H::SMEM_BYTES, as_stream(stream)>>>(map, a);
    PU3_LAUNCH_CHECK("head_tc_kernel");
    return PU3_OK;
}
