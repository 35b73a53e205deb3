// Tensor-core path of the 1x1 convolutions of the expansion head (network/upsampler.py:349-372, the one true
// dense contraction of the hot path: 265 -> 128 -> 128 -> 64 -> 3 over every up-sampled point).
//
//     Y[b, :, p] = act(W . X[b, :, p] + bias)            X (b, cin, n) channel-major fp32, W (cout, cin) fp32
//
// Blackwell-native design (sm_100a only):
//   * 3xTF32 on tcgen05: every fp32 operand is split into hi = tf32(v) and lo = v - hi, and the product is
//     accumulated as hi.hi + lo.hi + hi.lo in the fp32 TMEM accumulator.  That keeps the result inside the
//     path's 1e-5 relative tolerance (a plain TF32 product is ~5e-4) at 3 tensor-core
//     passes, an order of magnitude above the 72 TFLOP/s FFMA roof of the SGEMM it replaces.
//   * D (points x cout) = A (points x cin) . B (cout x cin)^T with UMMA M = 128 points, N = cout, K = 8 per
//     instruction.  A is MN-major: a TMA box of 32 points x 32 channels with the "128-byte swizzle, 32-byte
//     atoms" mode (the one MN-major layout tf32 accepts) lands exactly as eight canonical 4-channel swizzle
//     atoms (512 B each), so channel-major activations need no transpose.
//     B is K-major with the 128-byte swizzle, pre-split and pre-swizzled once per call by a tiny kernel into
//     [k-block][hi|lo][cout x 128 B] and brought in by one bulk copy per stage.
//   * warp-specialised persistent CTA, one per SM: warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
//     warps 2-5 = converters (raw -> hi in place, lo beside it, then fence.proxy.async), warps 6-9 = epilogue
//     (tcgen05.ld, bias / ReLU, fused variants, coalesced stores: lane = point).  Two 128-point sub-tiles share
//     one weight stage; accumulators are double buffered in TMEM (2 x 2 x cout columns) so the epilogue of tile
//     i overlaps the MMAs of tile i+1.
//   * fused epilogues: (1) the feature-expansion layer writes both replicas relu(acc + b + w_code*code[j]) as
//     one 8-byte store per point, so the (B,265,N*r) tensor and the separate "pre" tensor never exist;
//     (2) the last two layers: relu(W3 h + b3) stays in registers and the 64 -> 3 projection + residual is
//     finished per point, so the 64-channel tensor is never written.
#include "tc_common.cuh"

namespace pu3 {
namespace tc {

constexpr int KB = 32;                        // channels per pipeline stage (one 128-byte weight row)
constexpr int SUBM = 128;                     // points per MMA (UMMA M)
constexpr int SUB = 2;                        // sub-tiles per CTA tile
constexpr int STAGES = 2;
constexpr int PREFETCH = 4;                   // stages the L2 prefetch runs ahead of the TMA loads
constexpr int A_SUB_BYTES = SUBM * KB * 4;    // 16 KB: 4 TMA boxes of 32 points x 32 channels
constexpr int NUM_THREADS = 448;             // TMA warp, MMA warp, 4 converter warps, 8 epilogue warps
constexpr int CONV_WARP0 = 2, EPI_WARP0 = 6;

enum Mode { MODE_PLAIN = 0, MODE_EXPAND = 1, MODE_PROJECT = 2 };

struct Args {
    int b, n, cin, cout;
    const unsigned char *wsplit;
    const float *bias;
    float *y; long long y_bstride;
    int relu;
    // MODE_EXPAND
    int r; const float *wfull; int w_stride, code_col; const float *code;
    // MODE_PROJECT
    const float *w4, *b4; int cout4; const float *res; long long res_bstride; int res_n, res_div;
    int variant;
    float *dbg;
};


template <int NC>
struct Cfg {
    static constexpr int B_HALF = NC * 128;                            // one weight tile (hi or lo): NC rows of 32 floats
    static constexpr int A_BYTES = SUB * A_SUB_BYTES;                  // raw / hi activations of a stage
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_HALF;       // [A hi][A lo][B hi][B lo]
    static constexpr int TX_BYTES = A_BYTES + 2 * B_HALF;              // what TMA delivers per stage
    // TMEM: per sub-tile a main accumulator (hi.hi) and a correction accumulator (hi.lo + lo.hi), NC columns each.
    // The tensor core truncates when it adds into the fp32 accumulator (measured: the error of a single-accumulator
    // 3xTF32 grows linearly with the number of MMAs, profiles/r1g); keeping the ~2^-11-times-smaller correction
    // terms out of the main accumulator cuts the number of truncating adds it sees by 3.
    static constexpr int ACC_COLS = SUB * 2 * NC;                      // one accumulator stage: [s0 main|s0 corr|s1 main|s1 corr]
    static constexpr int ACC_STAGES = 512 / ACC_COLS;                  // NC = 64: double buffered, NC = 128: single
    static constexpr int TMEM_COLS = 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;     // + slack for the 1 KB alignment
    // instruction descriptors (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, A MN-major, B K-major, M = 128;
    // IDESC2 multiplies by [W_hi; W_lo] (N = 2 NC, main and correction columns in one instruction), IDESC1 by W_hi
    static constexpr uint32_t IDESC_BASE = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(SUBM >> 4) << 24);
    static constexpr uint32_t IDESC1 = IDESC_BASE | ((uint32_t)(NC >> 3) << 17);
    static constexpr uint32_t IDESC2 = IDESC_BASE | ((uint32_t)((2 * NC) >> 3) << 17);
};

// ---- weights: fp32 (cout, cin) -> [k-block][hi | lo][NC rows x 128 B], 16-byte chunks XOR-swizzled by row ----
__global__ void __launch_bounds__(256) split_weights_kernel(int cin, int cout, int nc, const float *__restrict__ w, int w_stride,
                                                           unsigned char *__restrict__ out) {
    const int nkb = (cin + KB - 1) / KB;
    const int total = nkb * nc * 8;           // one thread per 16-byte chunk
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const int c = t & 7, row = (t >> 3) % nc, kb = (t >> 3) / nc;
        float hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ch = kb * KB + c * 4 + i;
            const float v = (row < cout && ch < cin) ? __ldg(w + (size_t)row * w_stride + ch) : 0.f;
            hi[i] = to_tf32(v);
            lo[i] = v - hi[i];
        }
        unsigned char *tile = out + (size_t)kb * (2 * nc * 128);
        const int off = row * 128 + ((c ^ (row & 7)) << 4);
        *reinterpret_cast<float4 *>(tile + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4 *>(tile + nc * 128 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// bring-up timeline: SM cycle counter of pipeline events of CTA 0 (test hook buffer: 8 event kinds x 256 slots after
// the 4096-float stage dump)
__device__ __forceinline__ void tl_mark(float *dbg, int kind, int slot) {
    if (dbg && blockIdx.x == 0 && slot < 256) {
        const unsigned long long c = clock64();
        reinterpret_cast<unsigned int *>(dbg)[4096 + kind * 256 + slot] = (unsigned int)c;
    }
}

template <int NC, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap xmap, const Args a) {
    using C = Cfg<NC>;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint64_t bar_full[STAGES], bar_conv[STAGES], bar_empty[STAGES], bar_tfull[2], bar_tempty[2];   // accumulator stages: C::ACC_STAGES of them in use
    __shared__ uint32_t tmem_base_s;
    __shared__ float bias_s[NC], wcode_s[NC], code_s[8], w4_s[MODE == MODE_PROJECT ? 3 * NC : 1], b4_s[4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));

    // A 128-point sub-tile is four TMA boxes of 32 points, and the boxes of a sub-tile may belong to different clouds: the
    // (cloud, point) list is cut into boxes, not into 128-point pieces per cloud (312 points = 9.75 boxes: 2.5 % padding
    // instead of 18 %).  TMEM lane quarter q of a sub-tile = its box q = one epilogue warp.
    const int bpc = (a.n + 31) / 32;                         // boxes per cloud
    const long long nboxes = (long long)a.b * bpc;
    const long long nsub = (nboxes + 3) / 4;
    const long long ntiles = (nsub + SUB - 1) / SUB;
    const int nkb = (a.cin + KB - 1) / KB;
    const int last_ksteps = ((a.cin - (nkb - 1) * KB) + 7) / 8;

    for (int i = threadIdx.x; i < NC; i += NUM_THREADS) {
        bias_s[i] = (a.bias && i < a.cout) ? a.bias[i] : 0.f;
        if (MODE == MODE_EXPAND) wcode_s[i] = i < a.cout ? a.wfull[(size_t)i * a.w_stride + a.code_col] : 0.f;
    }
    if (MODE == MODE_EXPAND && threadIdx.x < 8) code_s[threadIdx.x] = (int)threadIdx.x < a.r ? a.code[threadIdx.x] : 0.f;
    if (MODE == MODE_PROJECT) {
        for (int i = threadIdx.x; i < 3 * NC; i += NUM_THREADS) {
            const int c3 = i / NC, co = i % NC;
            w4_s[i] = (c3 < a.cout4 && co < a.cout) ? a.w4[(size_t)c3 * a.cout + co] : 0.f;
        }
        if (threadIdx.x < 4) b4_s[threadIdx.x] = ((int)threadIdx.x < a.cout4 && a.b4) ? a.b4[threadIdx.x] : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_conv[s], 4);       // one arrival per converter warp
            mbar_init(&bar_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_tfull[s], 1);
            mbar_init(&bar_tempty[s], 8);     // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer (whole warp waits, one elected lane issues) =====
        // Only two 96 KB stages fit shared memory, i.e. ~32 KB of activations in flight per SM -- too little to cover
        // HBM latency (profiles/r1g: the load pipeline alone topped out near 3 TB/s).  So lanes 0-7 also prefetch the
        // eight boxes of the stage PREFETCH steps ahead into L2 (cp.async.bulk.prefetch.tensor): the TMA loads that
        // fill shared memory then hit L2, and the depth of the HBM queue no longer depends on shared-memory capacity.
        uint32_t stage = 0, phase = 0;
        long long li = 0;                                    // this CTA's tile counter
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++li) {
            int cb[SUB][4], cp[SUB][4];
#pragma unroll
            for (int s = 0; s < SUB; ++s)
#pragma unroll
                for (int mb = 0; mb < 4; ++mb) {
                    const long long box = (t * SUB + s) * 4 + mb;   // box >= nboxes: cloud index >= b -> fully out of bounds -> zeros
                    cb[s][mb] = (int)(box / bpc);
                    cp[s][mb] = (int)(box % bpc) * 32;
                }
            for (int kb = 0; kb < nkb; ++kb) {
                if (lane < SUB * 4 && !(a.variant & 64)) {
                    const long long f = li * nkb + kb + PREFETCH;
                    const long long lt = f / nkb, tt = blockIdx.x + lt * gridDim.x;
                    if (tt < ntiles) {
                        const long long box = tt * SUB * 4 + lane;      // lanes 0..7: the eight boxes of that tile
                        if (box < nboxes)
                            tma_prefetch_3d(&xmap, (int)(box % bpc) * 32, (int)(f - lt * nkb) * KB, (int)(box / bpc));
                    }
                }
                mbar_wait(&bar_empty[stage], phase ^ 1u);
                if (lane == 0) tl_mark(a.dbg, 0, (int)(li * nkb + kb));
                if (elect_one()) {
                    const bool no_w = a.variant & 256, no_a = a.variant & 512;    // timing experiments only
                    mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)((no_a ? 0 : C::A_BYTES) + (no_w ? 0 : 2 * C::B_HALF)));
                    const uint32_t sa = smem0 + stage * C::STAGE_BYTES;
                    if (!no_a) {
#pragma unroll
                        for (int s = 0; s < SUB; ++s)
#pragma unroll
                            for (int mb = 0; mb < 4; ++mb)
                                tma_load_3d(sa + s * A_SUB_BYTES + mb * 4096, &xmap, cp[s][mb], kb * KB, cb[s][mb], &bar_full[stage]);
                    }
                    if (!no_w) bulk_load(sa + 2 * C::A_BYTES, a.wsplit + (size_t)kb * (2 * C::B_HALF), 2 * C::B_HALF, &bar_full[stage]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp waits, one elected lane issues) =====
        uint32_t stage = 0, phase = 0, it = 0;
        // A stage: 4 TMA boxes (32 points each, LBO apart) of 32 channel rows x 128 B; an MMA (K = 8) reads two
        // 4-row swizzle atoms SBO apart
        const uint32_t a_lbo = (a.variant & 1) ? 512u : 4096u, a_sbo = (a.variant & 1) ? 4096u : 512u;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const uint32_t acc = it % C::ACC_STAGES, acc_phase = (it / C::ACC_STAGES) & 1u;
            mbar_wait(&bar_tempty[acc], acc_phase ^ 1u);
            if (lane == 0) tl_mark(a.dbg, 5, (int)it);
            tc_fence_after();
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&bar_full[stage], phase);
                if (lane == 0) tl_mark(a.dbg, 1, (int)(it * nkb + kb));
                mbar_wait(&bar_conv[stage], phase);
                if (lane == 0) tl_mark(a.dbg, 3, (int)(it * nkb + kb));
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem0 + stage * C::STAGE_BYTES;
                    const uint32_t sb = sa + 2 * C::A_BYTES;
                    const int nks = kb == nkb - 1 ? last_ksteps : KB / 8;
#pragma unroll
                    for (int s = 0; s < SUB; ++s) {
                        const uint32_t d = tmem_base + acc * C::ACC_COLS + s * 2 * NC;   // [main | correction]
                        for (int ks = 0; ks < nks; ++ks) {
                            if (a.variant & (8 | 32)) continue;   // bring-up / timing: no MMA
                            const uint64_t ahi = smem_desc(sa + s * A_SUB_BYTES + ks * 1024, a_lbo, a_sbo, LAYOUT_SW128_32B);
                            const uint64_t alo = smem_desc(sa + C::A_BYTES + s * A_SUB_BYTES + ks * 1024, a_lbo, a_sbo, LAYOUT_SW128_32B);
                            const uint64_t bhl = smem_desc(sb + ks * 32, 16, 1024, LAYOUT_SW128);    // rows [W_hi; W_lo]
                            umma_tf32(d, ahi, bhl, C::IDESC2, (kb | ks) != 0);        // main += hi.hi, correction += hi.lo
                            if (!(a.variant & 2)) umma_tf32(d + NC, alo, bhl, C::IDESC1, 1u);   // correction += lo.hi
                        }
                    }
                    tc_commit(&bar_empty[stage]);            // frees the stage when these MMAs have read it
                    if (kb == nkb - 1) tc_commit(&bar_tfull[acc]);   // accumulators of this tile are complete
                }
                __syncwarp();
                if (lane == 0) tl_mark(a.dbg, 4, (int)(it * nkb + kb));
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp < EPI_WARP0) {
        // ===== converters: raw fp32 -> hi (in place) + lo =====
        const int ct = threadIdx.x - CONV_WARP0 * 32;        // 0..127
        int cstep = 0;
        uint32_t stage = 0, phase = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&bar_full[stage], phase);
                float4 *hi = reinterpret_cast<float4 *>(smem_gen + stage * C::STAGE_BYTES);
                float4 *lo = reinterpret_cast<float4 *>(smem_gen + stage * C::STAGE_BYTES + C::A_BYTES);
                if (a.dbg && blockIdx.x == 0 && t == blockIdx.x && kb == 0) {   // bring-up: what TMA delivered
                    const float *ar = reinterpret_cast<const float *>(hi);
                    const float *br = reinterpret_cast<const float *>(smem_gen + stage * C::STAGE_BYTES + 2 * C::A_BYTES);
                    for (int i = ct; i < 1024; i += 128) { a.dbg[i] = ar[i]; a.dbg[1024 + i] = br[i]; }
                    if (ct == 0) { a.dbg[2048] = __uint_as_float(tmem_base); a.dbg[2049] = __uint_as_float(smem0); }
                    __syncwarp();
                }
#pragma unroll 4
                for (int i = ct; i < ((a.variant & 128) ? 0 : C::A_BYTES / 16); i += 128) {
                    const float4 v = hi[i];
                    float4 h, l;
                    h.x = to_tf32(v.x); h.y = to_tf32(v.y); h.z = to_tf32(v.z); h.w = to_tf32(v.w);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    if (a.variant & 4) { h = make_float4(1.f, 1.f, 1.f, 1.f); l = make_float4(0.f, 0.f, 0.f, 0.f); }
                    hi[i] = h;
                    lo[i] = l;
                }
                fence_proxy_async();                         // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_conv[stage]);
                if (ct == 0) tl_mark(a.dbg, 2, cstep);
                ++cstep;
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global (lane = point) =====
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int s = (warp - EPI_WARP0) >> 2;               // the sub-tile this warp drains
        uint32_t it = 0;
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
            const uint32_t acc = it % C::ACC_STAGES, acc_phase = (it / C::ACC_STAGES) & 1u;
            mbar_wait(&bar_tfull[acc], acc_phase);
            if (warp == EPI_WARP0 && lane == 0) tl_mark(a.dbg, 6, (int)it);
            tc_fence_after();
            if (a.variant & 8) {   // bring-up: TMEM store/load self test, value = lane * 1000 + column
                for (int c = 0; c < 2 * NC; ++c) {
                    const uint32_t val = __float_as_uint(c < NC ? (float)((q * 32 + lane) * 1000 + c) : 0.f);
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::ACC_COLS + s * 2 * NC + c), "r"(val) : "memory");
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            {
                const long long box = (t * SUB + s) * 4 + q;
                const long long bi = box / bpc;
                const int p = (int)(box % bpc) * 32 + lane;
                const bool valid = box < nboxes && p < a.n && !(a.variant & 16);   // 16: timing without stores
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::ACC_COLS + s * 2 * NC;
                float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll(MODE == MODE_PROJECT ? NC / 32 : 1)
                for (int ch = 0; ch < NC / 32; ++ch) {       // store modes stay rolled: large body (several store paths), keep it in the i-cache
                    uint32_t v[32], vc[32];
                    tmem_ld32(taddr + ch * 32, v);               // main and correction columns: two loads in flight, one wait
                    tmem_ld32(taddr + NC + ch * 32, vc);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(vc[j]));
                    // stores: ONE divergent branch per chunk (valid = this lane's point exists); inside it only the
                    // warp-uniform co < cout test remains, so every store is a predicated STG instead of a branch region
                    if (MODE == MODE_PLAIN) {
                        float *yp = a.y + bi * a.y_bstride + p + (size_t)(ch * 32) * a.n;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float r = __uint_as_float(v[j]) + bias_s[ch * 32 + j];
                            if (a.relu) r = fmaxf(r, 0.f);
                            v[j] = __float_as_uint(r);
                        }
                        if (valid) {
                            const int lim = a.cout - ch * 32;        // channels of this chunk that exist (warp-uniform)
                            if (lim >= 32) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) yp[(size_t)j * a.n] = __uint_as_float(v[j]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (j < lim) yp[(size_t)j * a.n] = __uint_as_float(v[j]);
                            }
                        }
                    } else if (MODE == MODE_EXPAND) {
                        const int nr = a.n * a.r;
                        float *yp = a.y + bi * a.y_bstride + (size_t)p * a.r;
                        if (a.r == 2) {      // the reference's step ratio: both replicas in one 8-byte store per point
                            const float c0 = code_s[0], c1 = code_s[1];
                            if (valid) {
                                const int lim = a.cout - ch * 32;
                                float *yq = yp + (size_t)(ch * 32) * nr;
                                if (lim >= 32) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) {
                                        const int co = ch * 32 + j;
                                        const float pre = __uint_as_float(v[j]) + bias_s[co];
                                        float2 o;
                                        o.x = fmaxf(__fmaf_rn(wcode_s[co], c0, pre), 0.f);
                                        o.y = fmaxf(__fmaf_rn(wcode_s[co], c1, pre), 0.f);
                                        *reinterpret_cast<float2 *>(yq + (size_t)j * nr) = o;
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) {
                                        const int co = ch * 32 + j;
                                        const float pre = __uint_as_float(v[j]) + bias_s[co];
                                        float2 o;
                                        o.x = fmaxf(__fmaf_rn(wcode_s[co], c0, pre), 0.f);
                                        o.y = fmaxf(__fmaf_rn(wcode_s[co], c1, pre), 0.f);
                                        if (j < lim) *reinterpret_cast<float2 *>(yq + (size_t)j * nr) = o;
                                    }
                                }
                            }
                        } else {
#pragma unroll 1
                            for (int jj = 0; jj < a.r; ++jj) {
                                const float cj = code_s[jj];
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const int co = ch * 32 + j;
                                    const float pre = __uint_as_float(v[j]) + bias_s[co];
                                    if (valid && co < a.cout) yp[(size_t)co * nr + jj] = fmaxf(__fmaf_rn(wcode_s[co], cj, pre), 0.f);
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int co = ch * 32 + j;
                            float h = __uint_as_float(v[j]) + bias_s[co];
                            if (a.relu) h = fmaxf(h, 0.f);
                            o0 = __fmaf_rn(w4_s[co], h, o0);
                            o1 = __fmaf_rn(w4_s[NC + co], h, o1);
                            o2 = __fmaf_rn(w4_s[2 * NC + co], h, o2);
                        }
                    }
                }
                if (MODE == MODE_PROJECT && valid) {
                    float o[3] = {o0 + b4_s[0], o1 + b4_s[1], o2 + b4_s[2]};
                    for (int c3 = 0; c3 < a.cout4; ++c3) {
                        float r = o[c3];
                        if (a.res) r += __ldg(a.res + bi * a.res_bstride + (size_t)c3 * a.res_n + p / a.res_div);
                        a.y[bi * a.y_bstride + (size_t)c3 * a.n + p] = r;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[acc]);
            if (warp == EPI_WARP0 && lane == 0) tl_mark(a.dbg, 7, (int)it);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// ---- host ------------------------------------------------------------------------------------------

static int nc_for(int cout) { return cout <= 64 ? 64 : 128; }

static int g_variant = -1;
static float *g_dbg = nullptr;
static int variant() {
    if (g_variant < 0) {
        const char *e = getenv("PU3_TC_VARIANT");
        g_variant = e ? atoi(e) : 0;
    }
    return g_variant;
}

template <int NC, int MODE>
static int launch(const CUtensorMap &map, const Args &a, cudaStream_t s) {
    using C = Cfg<NC>;
    // the attribute is per device/context (tensors on other devices are supported): set it on every launch, like the other kernels
    {
        int st = cuda_status(cudaFuncSetAttribute(conv_tc_kernel<NC, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES),
                             "conv_tc: shared memory opt-in");
        if (st) return st;
    }
    const long long nsubs = ((long long)a.b * ((a.n + 31) / 32) + 3) / 4;
    const long long ntiles = (nsubs + SUB - 1) / SUB;
    const int grid = (int)(ntiles < device_info().sm_count ? ntiles : device_info().sm_count);
    conv_tc_kernel<NC, MODE><<<grid, NUM_THREADS, C::SMEM_BYTES, s>>>(map, a);
    PU3_LAUNCH_CHECK("conv_tc_kernel");
    return PU3_OK;
}

static int run(int mode, Args a, const float *x, long long x_bstride, cudaStream_t s) {
    PU3_ARG_CHECK(a.b >= 0 && a.n >= 0 && a.cin > 0 && a.cout > 0 && a.cout <= 128, "conv_tc: bad size b=%d n=%d cin=%d cout=%d", a.b, a.n, a.cin, a.cout);
    if (a.b == 0 || a.n == 0) return PU3_OK;
    PU3_ARG_CHECK(x && a.wsplit && a.y, "conv_tc: null pointer");
    PU3_ARG_CHECK(a.n % 4 == 0 && x_bstride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0,
                  "conv_tc: the TMA path needs n %% 4 == 0 and 16-byte aligned slices (n=%d)", a.n);
    PU3_ARG_CHECK((reinterpret_cast<uintptr_t>(a.wsplit) & 15) == 0, "conv_tc: split weights must be 16-byte aligned");
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled not available"); return PU3_E_ARG; }
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)a.n, (cuuint64_t)a.cin, (cuuint64_t)a.b};
    const cuuint64_t gstride[2] = {(cuuint64_t)a.n * 4, (cuuint64_t)x_bstride * 4};
    const cuuint32_t box[3] = {32, KB, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(x), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return PU3_E_ARG; }
    a.variant = variant();
    a.dbg = g_dbg;
    const int nc = nc_for(a.cout);
    if (mode == MODE_PLAIN) return nc == 64 ? launch<64, MODE_PLAIN>(map, a, s) : launch<128, MODE_PLAIN>(map, a, s);
    if (mode == MODE_EXPAND) return nc == 64 ? launch<64, MODE_EXPAND>(map, a, s) : launch<128, MODE_EXPAND>(map, a, s);
    PU3_ARG_CHECK(nc == 64 && a.cout4 >= 1 && a.cout4 <= 3, "conv_tc_project: needs cout <= 64 and 1..3 projected channels");
    return launch<64, MODE_PROJECT>(map, a, s);
}

}  // namespace tc
}  // namespace pu3

using namespace pu3;

// Test hook: descriptor variant (bit 0 swaps the leading/stride byte offsets of the MN-major operand).
extern "C" void pu3_conv_tc_set_variant(int v) { tc::g_variant = v; }
extern "C" void pu3_conv_tc_set_debug(float *buf) { tc::g_dbg = buf; }

extern "C" size_t pu3_conv_tc_wsplit_bytes(int cin, int cout) {
    if (cin <= 0 || cout <= 0 || cout > 128) return 0;
    return (size_t)((cin + tc::KB - 1) / tc::KB) * 2 * tc::nc_for(cout) * 128;
}

extern "C" int pu3_conv_tc_prepare_f32(int cin, int cout, const float *w, int w_stride, void *wsplit, pu3_stream_t stream) {
    PU3_ARG_CHECK(cin > 0 && cout > 0 && cout <= 128 && w_stride >= cin, "conv_tc_prepare: bad size cin=%d cout=%d stride=%d", cin, cout, w_stride);
    PU3_ARG_CHECK(w && wsplit && (reinterpret_cast<uintptr_t>(wsplit) & 15) == 0, "conv_tc_prepare: null or unaligned pointer");
    const int nc = tc::nc_for(cout);
    const int total = (cin + tc::KB - 1) / tc::KB * nc * 8;
    tc::split_weights_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(cin, cout, nc, w, w_stride, static_cast<unsigned char *>(wsplit));
    PU3_LAUNCH_CHECK("split_weights_kernel");
    return PU3_OK;
}

extern "C" int pu3_conv_tc_f32(int b, int n, int cin, int cout, const float *x, long long x_bstride, const void *wsplit,
                               const float *bias, float *y, long long y_bstride, int relu, pu3_stream_t stream) {
    tc::Args a{};
    a.b = b; a.n = n; a.cin = cin; a.cout = cout; a.wsplit = static_cast<const unsigned char *>(wsplit); a.bias = bias;
    a.y = y; a.y_bstride = y_bstride; a.relu = relu;
    return tc::run(tc::MODE_PLAIN, a, x, x_bstride, as_stream(stream));
}

extern "C" int pu3_conv_tc_expand_f32(int b, int n, int cin, int cout, int r, const float *x, long long x_bstride,
                                      const void *wsplit, const float *w, int w_stride, int code_col, const float *bias,
                                      const float *code, float *y, long long y_bstride, pu3_stream_t stream) {
    PU3_ARG_CHECK(r >= 1 && r <= 8 && w && code, "conv_tc_expand: bad expansion arguments (r=%d)", r);
    PU3_ARG_CHECK(r != 2 || ((reinterpret_cast<uintptr_t>(y) & 7) == 0 && y_bstride % 2 == 0), "conv_tc_expand: y must be 8-byte aligned");
    tc::Args a{};
    a.b = b; a.n = n; a.cin = cin; a.cout = cout; a.wsplit = static_cast<const unsigned char *>(wsplit); a.bias = bias;
    a.y = y; a.y_bstride = y_bstride; a.relu = 1; a.r = r; a.wfull = w; a.w_stride = w_stride; a.code_col = code_col; a.code = code;
    return tc::run(tc::MODE_EXPAND, a, x, x_bstride, as_stream(stream));
}

extern "C" int pu3_conv_tc_project_f32(int b, int n, int cin, int cmid, int cout, const float *x, long long x_bstride,
                                       const void *wsplit, const float *bias_mid, const float *w_out, const float *b_out,
                                       float *y, long long y_bstride, const float *res, long long res_bstride, int res_n,
                                       int res_div, pu3_stream_t stream) {
    PU3_ARG_CHECK(w_out && (!res || (res_div >= 1 && res_n >= 1)), "conv_tc_project: bad arguments");
    tc::Args a{};
    a.b = b; a.n = n; a.cin = cin; a.cout = cmid; a.wsplit = static_cast<const unsigned char *>(wsplit); a.bias = bias_mid;
    a.y = y; a.y_bstride = y_bstride; a.relu = 1; a.w4 = w_out; a.b4 = b_out; a.cout4 = cout;
    a.res = res; a.res_bstride = res_bstride; a.res_n = res_n; a.res_div = res_div > 0 ? res_div : 1;
    return tc::run(tc::MODE_PROJECT, a, x, x_bstride, as_stream(stream));
}
