// Farthest point sampling, sm_100a.
//
// Replaces furthest_point_sampling_forward_kernel / furthest_sampling_cuda_forward
// (sampling/sampling_cuda.cu:103-265 of the reference).
//
// The reference runs ONE CTA of <=512 threads per cloud; every one of the m-1 dependent rounds
// re-reads the running distances (and every point past the first 512) from global memory and
// ends in a log2(T)-step shared-memory tree with a __syncthreads per step.  FPS is a chain of
// m-1 dependent arg-max rounds, so it is latency-bound: the design goal is the shortest round.
//
//   * the cloud lives ON CHIP for the whole kernel: each thread keeps its points' running
//     distance (and, when they fit, the coordinates) in registers; a shared-memory copy of the
//     coordinates serves the one winner lookup per round.  HBM is touched once.
//   * a cloud too large for one SM is spread over a thread-block CLUSTER (2..16 CTAs); per round
//     each CTA publishes its local winner into every peer's shared memory (DSMEM) and one
//     barrier.cluster makes them visible.
//   * the arg-max inside a CTA is two redux.sync per warp + ONE __syncthreads per round.
//
// Bit-exactness with the reference (indices are compared for equality):
//   - distance: FMUL(dy,dy), FFMA(dx,dx,.), FFMA(dz,dz,.) like the reference SASS (sqdist3)
//   - running distance t = fminf(d, t)                               (:144)
//   - winner: largest t; among equal t the smallest (k mod T), then the smallest k, where
//     T = 2^floor(log2 n) <= 512 is the reference's block size (cuda_utils.h:9-14).  That is what
//     its per-thread strided scan (:133-150, first strictly greater) followed by its tree
//     (:155-167, lower slot keeps ties) computes.  Here it is a lexicographic max on
//     (t, -rank), rank = (k mod T) * ceil(n/T) + k / T.
//   - deliberately NOT reproduced: the reference indexes temp rows by blockIdx.x (:131,146), which
//     corrupts results for b > 32; every cloud owns its row here.
#include "pu3_common.cuh"

namespace pu3 {

constexpr int FPS_MAX_CLUSTER = 16;

// what a CTA publishes per round: 32 bytes
struct __align__(16) FpsCand {
    uint32_t dkey;  // float bits of the running distance (>= 0, so uint order == float order)
    uint32_t rank;  // tie rank, smaller wins
    int32_t k;      // point index
    float x, y, z;
    uint32_t pad0, pad1;
};

// Speculative rounds (cluster flavour with <= 64 warps).  FPS is a chain of m-1 dependent arg-max rounds and a round costs one
// cluster exchange (~1 us) whatever the work, so the only way to go faster is FEWER EXCHANGES.  Every warp publishes its best
// candidate AND the key (t, rank) of its second best (in the two spare words of the 32-byte slot).  After the exchange all
// warps know the candidates sorted by key, c1 >= c2 >= ..., and B = the best UNPUBLISHED key.  c1 is this round's sample.
// c2 is provably the NEXT sample, without another exchange, when (a) key(c2) > B -- it is the true global runner-up -- and
// (b) its running distance does not change when c1 is added: d(c2, c1) >= t(c2), evaluated with the very expression the update
// uses, and (c) t(c2) > 0.  Proof: every other point i had key(i) < key(c2); adding c1 can only lower t(i), and a lowered t(i)
// is strictly below its old value <= t(c2), an unchanged one keeps key(i) < key(c2); c1 itself drops to t = 0 < t(c2): c2 is
// the lexicographic maximum after the update, which is what the reference's next round selects.  The same argument accepts c3 (unaffected by c1 and c2), c4, ...; the chain stops at
// the first candidate that fails (a) or (b).  The accepted samples' updates are then applied in ONE pass over the points.
// Measured on merged tile clouds: ~1.9 samples per exchange at depth 2, ~3.5 at depth 4 (profiles/r2).  Indices and temp stay
// bit-identical to the reference: the same samples are chosen in the same order.
constexpr int FPS_SPEC = 4;
// tuning / test hooks: run-time cap of the speculation depth (1 = one sample per exchange, the round-1 behaviour) and the number
// of exchanges cloud 0 of the last launch needed (samples per exchange = (m - 1) / exchanges)
__device__ int g_fps_spec_cap = FPS_SPEC;
__device__ unsigned int g_fps_exchanges = 0;
__device__ int g_fps_select_all = -1;       // tuning: 1 = every warp runs the selection itself, 0 = warp 0 + hand-over through shared memory, -1 = by cloud size

constexpr int FPS_MAX_CAND = 256;          // candidates a CTA receives per round: cluster size x warps per CTA
struct FpsSmem {
    FpsCand cluster_slot[2][FPS_MAX_CAND];     // [round parity][source CTA rank * warps + source warp]
    FpsCand warp_slot[2][32];                  // [round parity][warp]: the warp's winner, coordinates included
    FpsCand cta_slot[2];                       // [round parity] the CTA's winner (wide clusters: leader-warp exchange)
    uint64_t mbar[2];                          // [round parity] "all S candidates of the round have landed"
    float res_w[2][FPS_SPEC][4];               // [round parity] samples chosen by warp 0 (x, y, z, -) ...
    int res_a[2];                              // ... and how many
};

// lexicographic arg-max over the full warp; returns true in exactly one lane (the winner)
__device__ __forceinline__ bool warp_argmax(uint32_t dkey, uint32_t rank, uint32_t &dmax, uint32_t &rmin) {
    dmax = __reduce_max_sync(0xffffffffu, dkey);
    const uint32_t cand = (dkey == dmax) ? rank : 0xffffffffu;
    rmin = __reduce_min_sync(0xffffffffu, cand);
    return cand == rmin && rmin != 0xffffffffu;
}

// bring-up timeline (build with `make -B EXTRA=-DPU3_FPS_TIMELINE`, then profiles/fps_timeline.py): SM cycle counter at the
// phases of rounds [1000, 1008) of cluster 0, CTA 0, thread 0.  Compiled out by default.
#ifdef PU3_FPS_TIMELINE
__device__ unsigned int *g_fps_timeline = nullptr;
__device__ __forceinline__ void fps_mark(int j, int phase) {
    if (g_fps_timeline && blockIdx.x == 0 && threadIdx.x == 0 && j >= 1000 && j < 1008)
        g_fps_timeline[(j - 1000) * 8 + phase] = (unsigned int)clock64();
}
#else
__device__ __forceinline__ void fps_mark(int, int) {}
#endif

// After an exchange: `total` (<= 64) candidates in `slots`, each with the key of the best point its publisher did NOT publish in
// the two spare words.  Chooses this exchange's samples (see FPS_SPEC): always the best candidate, then up to spec_cap - 1 more
// while each is (a) above every unpublished key, (b) unchanged by the samples accepted before it, (c) at t > 0.  Every warp runs
// this on the same data and reaches the same result.  Returns the number of samples; their coordinates go to wx/wy/wz, their
// indices to out[j ...] (written by the thread for which `writer` is set).
template <int SPEC>
__device__ __forceinline__ int fps_select(const FpsCand *slots, int total, int lane, int spec_cap, int j, int m, bool writer,
                                          int32_t *out, float (&wx)[SPEC], float (&wy)[SPEC], float (&wz)[SPEC]) {
    // Round 2, second version.  The first one popped the candidates one after the other with the slot read, the distance tests
    // and the break conditions of a sample on the chain before the next reduce-max: ~650 cycles per accepted sample, 2 600 of an
    // exchange's 5 600 cycles at depth 4 (in-kernel timeline, profiles/r2) -- latency, not issue slots: running it in one warp
    // instead of all changed nothing, and ranking every candidate against all others (no collectives, but ~900 instructions in
    // one warp) was slower still.  Now the SPEC best candidates are found first by SPEC rounds of {reduce-max, ballot, shuffle}
    // and nothing else; their slots are then read and all acceptance tests evaluated side by side.  Same candidates, same order,
    // same tests as before: results are bit-identical.
    uint32_t ct[2], cr[2];
    uint32_t lbt = 0u, lbr = 0xffffffffu;                      // lane-local best of the unpublished keys
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int c = lane + 32 * u;
        ct[u] = 0u; cr[u] = 0xffffffffu;
        if (c < total) {
            const uint2 kr = *reinterpret_cast<const uint2 *>(&slots[c]);           // (dkey, rank)
            const uint2 s2 = *(reinterpret_cast<const uint2 *>(&slots[c]) + 3);     // best unpublished key of that publisher
            ct[u] = kr.x; cr[u] = kr.y;
            if (s2.x > lbt || (s2.x == lbt && s2.y < lbr)) { lbt = s2.x; lbr = s2.y; }
        }
    }
    // B = the best key that was NOT published
    const uint32_t bt = __reduce_max_sync(0xffffffffu, lbt);
    const uint32_t br = __reduce_min_sync(0xffffffffu, lbt == bt ? lbr : 0xffffffffu);
    // the SPEC best candidates in order: SPEC rounds of {reduce-max, ballot, shuffle}, NOTHING else on the chain (the slot reads
    // and the acceptance tests of all of them follow side by side)
    int cwin[SPEC];
    bool cvalid[SPEC];
    int taken = 0;
#pragma unroll
    for (int mth = 0; mth < SPEC; ++mth) {
        uint32_t lt = 0u, lr = 0xffffffffu;
        int lc = lane;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const bool free_u = !((taken >> u) & 1);
            const bool take = free_u & ((ct[u] > lt) | ((ct[u] == lt) & (cr[u] < lr)));
            lt = take ? ct[u] : lt; lr = take ? cr[u] : lr; lc = take ? lane + 32 * u : lc;
        }
        const uint32_t tmax = __reduce_max_sync(0xffffffffu, lt);
        unsigned cands = __ballot_sync(0xffffffffu, lt == tmax && lr != 0xffffffffu);
        if (__popc(cands) > 1) {                                // tie on the distance (rare): smallest rank wins
            const uint32_t rmin = __reduce_min_sync(0xffffffffu, lt == tmax ? lr : 0xffffffffu);
            cands = __ballot_sync(0xffffffffu, lt == tmax && lr == rmin);
        }
        cvalid[mth] = cands != 0u;
        const int csrc = cands ? __ffs(cands) - 1 : 0;
        cwin[mth] = __shfl_sync(0xffffffffu, lc, csrc);
        if (lane == csrc) taken |= 1 << (cwin[mth] >> 5);
    }
    uint32_t kt[SPEC], krk[SPEC];
    int kidx[SPEC];
    float cx[SPEC], cy[SPEC], cz[SPEC];
#pragma unroll
    for (int mth = 0; mth < SPEC; ++mth) {
        const uint4 wlo = *reinterpret_cast<const uint4 *>(&slots[cwin[mth]]);      // broadcast reads, independent of each other
        const uint2 whi = *(reinterpret_cast<const uint2 *>(&slots[cwin[mth]]) + 2);
        kt[mth] = wlo.x; krk[mth] = cvalid[mth] ? wlo.y : 0xffffffffu; kidx[mth] = (int)wlo.z;
        cx[mth] = __uint_as_float(wlo.w); cy[mth] = __uint_as_float(whi.x); cz[mth] = __uint_as_float(whi.y);
    }
    int A = 0;
    bool go = true;
#pragma unroll
    for (int mth = 0; mth < SPEC; ++mth) {
        if (mth > 0) {
            go = go && krk[mth] != 0xffffffffu;                                             // a candidate is left
            go = go && (kt[mth] > bt || (kt[mth] == bt && krk[mth] < br));                  // (a) above every unpublished key
            go = go && kt[mth] != 0u;                                                       // (c) t > 0
            const float tm = __uint_as_float(kt[mth]);
#pragma unroll
            for (int a2 = 0; a2 < SPEC; ++a2)                                               // (b) unchanged by the samples before it
                if (a2 < mth) go = go && (fminf(sqdist3(cx[mth] - cx[a2], cy[mth] - cy[a2], cz[mth] - cz[a2]), tm) == tm);
        }
        if (go) {
            wx[mth] = cx[mth]; wy[mth] = cy[mth]; wz[mth] = cz[mth];
            if (writer) out[j + mth] = kidx[mth];
            A = mth + 1;
            if (j + A >= m || A >= spec_cap) go = false;
        }
    }
    return A;
}

// PPT points per thread; XYZ_REGS: coordinates in registers (else read from shared memory each round)
// DIRECT: flavour of the cluster exchange (host: S x warps <= 64), a template parameter so that each flavour is compiled alone
template <int PPT, bool XYZ_REGS, int MAX_THREADS, bool DIRECT>
__global__ void __launch_bounds__(MAX_THREADS, 1)
fps_kernel(int n_stride, int m_stride, const int32_t *__restrict__ n_arr, const int32_t *__restrict__ m_arr,
           const float *__restrict__ xyz, float *__restrict__ temp, int32_t *__restrict__ idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FpsSmem &sm = *reinterpret_cast<FpsSmem *>(smem_raw);
    float *sx = reinterpret_cast<float *>(smem_raw + sizeof(FpsSmem));

    const uint32_t S = cluster_nctarank();
    const uint32_t crank = cluster_ctarank();
    const int cloud = blockIdx.x / S;
    // ragged batches: every cloud may have its own point / sample count (rows are n_stride / m_stride apart)
    const int n = n_arr ? min(__ldg(n_arr + cloud), n_stride) : n_stride;
    const int m = m_arr ? min(__ldg(m_arr + cloud), m_stride) : m_stride;
    const int t_ref = n > 0 ? min(512, 1 << (31 - __clz(n))) : 1;  // the reference's block size for THIS cloud
    const int t_shift = 31 - __clz(t_ref);                         // t_ref is a power of two: k / t_ref = k >> t_shift
    const int nthreads = blockDim.x;
    const int GT = nthreads * S;                     // threads per cloud; a multiple of t_ref
    const int g = crank * nthreads + threadIdx.x;    // this thread's id within the cloud
    const int cap = nthreads * PPT;                  // points this CTA can hold
    float *sy = sx + cap, *sz = sy + cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = nthreads >> 5;

    const float *p = xyz + (size_t)cloud * n_stride * 3;
    float *trow = temp ? temp + (size_t)cloud * n_stride : nullptr;
    int32_t *out = idx + (size_t)cloud * m_stride;
    const uint32_t rank_rows = (n + t_ref - 1) / t_ref;

    // ---- load: thread owns points k = g + i*GT; local slot i*nthreads + tid -------------------
    float px[XYZ_REGS ? PPT : 1], py[XYZ_REGS ? PPT : 1], pz[XYZ_REGS ? PPT : 1], t[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = g + i * GT;
        float x = 0.f, y = 0.f, z = 0.f, tv = -1.f;  // tv < 0 marks "no point": never wins (real t >= 0)
        if (k < n) {
            x = __ldg(p + (size_t)k * 3 + 0);
            y = __ldg(p + (size_t)k * 3 + 1);
            z = __ldg(p + (size_t)k * 3 + 2);
            tv = trow ? trow[k] : 1e10f;
        }
        const int slot = i * nthreads + threadIdx.x;
        sx[slot] = x; sy[slot] = y; sz[slot] = z;
        if (XYZ_REGS) { px[i] = x; py[i] = y; pz[i] = z; }
        t[i] = tv;
    }
    if (g == 0 && m > 0) out[0] = 0;                         // :115 first sample is point 0
    constexpr int SPEC = FPS_SPEC;                           // samples one exchange may yield (all three exchange flavours)
    float wx[SPEC], wy[SPEC], wz[SPEC];                      // samples chosen by the last exchange, their update still pending
#pragma unroll
    for (int a = 0; a < SPEC; ++a) { wx[a] = 0.f; wy[a] = 0.f; wz[a] = 0.f; }
    int A = 1;                                               // how many of them
    if (n > 0) { wx[0] = __ldg(p + 0); wy[0] = __ldg(p + 1); wz[0] = __ldg(p + 2); }
    if (S > 1 && threadIdx.x == 0) {
        mbar_init(&sm.mbar[0], 1);
        mbar_init(&sm.mbar[1], 1);
        mbar_fence_init_cluster();
    }
    __syncthreads();
    if (S > 1) { cluster_arrive_release(); cluster_wait_acquire(); }  // peers' smem + barriers exist before remote stores

    const int spec_cap = min(SPEC, max(1, g_fps_spec_cap));
    // measured (profiles/debug/fps_spec_ab.py): the hand-over wins on the 24 960-point clouds (3.57 against 3.74 ms), the redundant
    // selection on the smaller ones (0.70 / 1.38 against 0.76 / 1.46 ms)
    const bool select_all = g_fps_select_all < 0 ? n < 20000 : g_fps_select_all != 0;
    int j = 1;                                               // next sample to choose
    uint32_t e = 1;
    for (; j < m; ++e) {                                     // e: exchange counter (buffer parity / barrier phase)
        const int par = (int)(e & 1u);
        fps_mark(j, 0);
        // ---- 1. apply the pending samples to the running distances (one pass), thread-local best and second best
        //         (first strictly greater: within a thread a lower i is a lower tie rank) -------
        float best = -1.f, sec = -1.f;
        int besti = 0, seci = 0;
        // (all SPEC distance evaluations are predicated on a < A rather than dispatched on A: measured, profiles/debug/fps_spec_ab.py --
        // a per-count copy of this loop is faster at depth 1 but slower at depth 4, where ~3.5 of the 4 are live anyway)
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            float x, y, z;
            if (XYZ_REGS) { x = px[i]; y = py[i]; z = pz[i]; }
            else { const int slot = i * nthreads + threadIdx.x; x = sx[slot]; y = sy[slot]; z = sz[slot]; }
            // padding slots keep t = -1: fminf(d,-1) = -1
            float d2 = fminf(sqdist3(x - wx[0], y - wy[0], z - wz[0]), t[i]);
#pragma unroll
            for (int a = 1; a < SPEC; ++a)
                if (a < A) d2 = fminf(sqdist3(x - wx[a], y - wy[a], z - wz[a]), d2);
            t[i] = d2;
            if (d2 > best) {
                if (SPEC > 1) { sec = best; seci = besti; }
                best = d2; besti = i;
            } else if (SPEC > 1 && d2 > sec) {
                sec = d2; seci = i;
            }
        }
        fps_mark(j, 1);
        // ---- 2. warp arg-max: the winning lane publishes (distance, tie rank, index, coordinates) ----------
        const uint32_t dkey = best >= 0.f ? __float_as_uint(best) : 0u;
        const uint32_t dmax_w = __reduce_max_sync(0xffffffffu, dkey);
        uint32_t rank = 0xffffffffu;
        int kbest = 0;
        if (dkey == dmax_w && best >= 0.f) {     // rank only where it can matter
            kbest = g + besti * GT;
            rank = (uint32_t)(kbest & (t_ref - 1)) * rank_rows + (uint32_t)(kbest >> t_shift);
        }
        // ties on the distance are rare: when exactly one lane holds the maximum the rank reduction is skipped
        const unsigned holders = __ballot_sync(0xffffffffu, rank != 0xffffffffu);
        const uint32_t rmin_w = __popc(holders) == 1 ? __shfl_sync(0xffffffffu, rank, __ffs(holders) - 1)
                                                     : __reduce_min_sync(0xffffffffu, rank);
        const bool publisher = rank == rmin_w && (rmin_w != 0xffffffffu || lane == 0);   // the winner, or lane 0 of an empty warp
        // the warp's SECOND best key: the winner lane offers its second point, every other lane its best
        uint32_t o_t = 0u, o_r = 0xffffffffu;
        if (SPEC > 1) {
            const bool iswin = rank == rmin_w && rmin_w != 0xffffffffu;
            const float off = iswin ? sec : best;
            const int offi = iswin ? seci : besti;
            const uint32_t okey = off >= 0.f ? __float_as_uint(off) : 0u;
            o_t = __reduce_max_sync(0xffffffffu, okey);
            uint32_t orank = 0xffffffffu;
            if (okey == o_t && off >= 0.f) {
                const int kk = g + offi * GT;
                orank = (uint32_t)(kk & (t_ref - 1)) * rank_rows + (uint32_t)(kk >> t_shift);
            }
            o_r = __reduce_min_sync(0xffffffffu, orank);
        }
        uint4 plo = make_uint4(0u, 0xffffffffu, 0u, 0u), phi = make_uint4(0u, 0u, o_t, o_r);
        if (publisher) {
            const int slot = besti * nthreads + threadIdx.x;            // coordinates from the smem copy: no dynamic register indexing
            plo.x = dmax_w; plo.y = rmin_w; plo.z = (uint32_t)kbest; plo.w = __float_as_uint(sx[slot]);
            phi.x = __float_as_uint(sy[slot]); phi.y = __float_as_uint(sz[slot]);
        }
        fps_mark(j, 2);
        if (S == 1) {
            // ---- 3a. one CTA per cloud: warp slots + ONE __syncthreads, every warp reduces the slots itself ----------
            if (publisher) {
                uint4 *dst = reinterpret_cast<uint4 *>(&sm.warp_slot[par][warp]);
                dst[0] = plo; dst[1] = phi;
            }
            __syncthreads();
            A = fps_select<SPEC>(&sm.warp_slot[par][0], nwarps, lane, spec_cap, j, m, threadIdx.x == 0, out, wx, wy, wz);
            j += A;
        } else if (DIRECT) {
            // ---- 3b. cluster of <= 64 warps: EVERY warp sends its candidate straight into every CTA's slot array by
            //          async DSMEM stores that complete on the receiver's mbarrier.  No CTA-level reduction, no
            //          __syncthreads, no leader warp: the in-kernel timeline (profiles/r1g) showed those cost ~1000 of the
            //          ~2400 cycles of a round.  The mbarrier wait is the only synchronisation: an exchange's slots are complete
            //          when all S x nwarps candidates have landed, and a warp cannot overwrite a slot of exchange e+2 before
            //          every warp of the cluster has read exchange e (it must first pass the barrier of exchange e+1, which
            //          needs their e+1 sends).  The candidate goes through the warp's own slot so that lane r sends it to
            //          CTA r: the S x 2 remote stores are issued by S lanes at once instead of one lane after the other.
            if (publisher) {
                uint4 *dst = reinterpret_cast<uint4 *>(&sm.warp_slot[par][warp]);
                dst[0] = plo; dst[1] = phi;
            }
            __syncwarp();
            if (lane < (int)S) {
                const uint4 *src = reinterpret_cast<const uint4 *>(&sm.warp_slot[par][warp]);
                const uint4 lo = src[0], hi = src[1];
                const uint32_t dst = map_to_cta(&sm.cluster_slot[par][0], (uint32_t)lane) + (crank * (uint32_t)nwarps + (uint32_t)warp) * 32u;
                const uint32_t rbar = map_to_cta(&sm.mbar[par], (uint32_t)lane);
                st_async_v4(dst, rbar, lo.x, lo.y, lo.z, lo.w);
                st_async_v4(dst + 16, rbar, hi.x, hi.y, hi.z, hi.w);
            }
            if (threadIdx.x == 0) mbar_arrive_expect_tx(&sm.mbar[par], S * (uint32_t)nwarps * 32u);
            fps_mark(j, 4);
            // ---- 4. every warp: wait for the S x nwarps candidates, sort out up to SPEC samples (see FPS_SPEC above) ----------
            mbar_wait_parity(&sm.mbar[par], ((e - 1u) >> 1) & 1u);   // phase = earlier uses of this buffer
            fps_mark(j, 5);
            // The selection is a chain of warp-wide reductions (~650 cycles per accepted sample when all 16 warps of an SM run it
            // at the same time, in-kernel timeline profiles/r2: 2 600 of an exchange's 5 600 cycles).  Every warp would reach the same
            // result, so warp 0 alone computes it and hands it over through shared memory: one __syncthreads instead of 7
            // redundant copies competing for the issue slots.
            if (select_all) {
                A = fps_select<SPEC>(&sm.cluster_slot[par][0], (int)S * nwarps, lane, spec_cap, j, m, g == 0, out, wx, wy, wz);
                fps_mark(j, 6);
                j += A;
                continue;
            }
            if (warp == 0) {
                A = fps_select<SPEC>(&sm.cluster_slot[par][0], (int)S * nwarps, lane, spec_cap, j, m, g == 0, out, wx, wy, wz);
                if (lane == 0) {
                    sm.res_a[par] = A;
#pragma unroll
                    for (int a = 0; a < SPEC; ++a) *reinterpret_cast<float4 *>(&sm.res_w[par][a][0]) = make_float4(wx[a], wy[a], wz[a], 0.f);
                }
            }
            __syncthreads();
            if (warp != 0) {
                A = sm.res_a[par];
#pragma unroll
                for (int a = 0; a < SPEC; ++a) {
                    const float4 w4 = *reinterpret_cast<const float4 *>(&sm.res_w[par][a][0]);
                    wx[a] = w4.x; wy[a] = w4.y; wz[a] = w4.z;
                }
            }
            fps_mark(j, 6);
            j += A;
        } else {
            // ---- 3c. wide cluster (> 64 warps: the 16-CTA whole-shape call): scanning S x nwarps candidates in every warp
            //          would cost more than it saves, so the CTA reduces first (warp slots, __syncthreads, leader warp) and
            //          only S candidates travel; lane r of the leader warp sends to CTA r.
            if (publisher) {
                uint4 *dst = reinterpret_cast<uint4 *>(&sm.warp_slot[par][warp]);
                dst[0] = plo; dst[1] = phi;
            }
            __syncthreads();
            if (warp == 0) {
                uint4 lo = make_uint4(0u, 0xffffffffu, 0u, 0u), hi = make_uint4(0u, 0u, 0u, 0u);
                if (lane < nwarps) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(&sm.warp_slot[par][lane]);
                    lo = src[0]; hi = src[1];
                }
                uint32_t dmax, rmin;
                const bool win = warp_argmax(lo.x, lo.y, dmax, rmin);
                const unsigned wb = __ballot_sync(0xffffffffu, win);
                // the CTA's best UNPUBLISHED key: the winner warp's second best or another warp's best, whichever is larger
                const uint32_t ot = win ? hi.z : lo.x, orr = win ? hi.w : lo.y;
                const uint32_t c_t = __reduce_max_sync(0xffffffffu, ot);
                const uint32_t c_r = __reduce_min_sync(0xffffffffu, ot == c_t ? orr : 0xffffffffu);
                if (wb == 0u ? lane == 0 : win) {
                    uint4 *dst = reinterpret_cast<uint4 *>(&sm.cta_slot[par]);
                    hi.z = c_t; hi.w = c_r;
                    dst[0] = lo; dst[1] = hi;
                    mbar_arrive_expect_tx(&sm.mbar[par], S * 32u);
                }
                __syncwarp();
                if (lane < (int)S) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(&sm.cta_slot[par]);
                    const uint4 clo = src[0], chi = src[1];
                    const uint32_t dst = map_to_cta(&sm.cluster_slot[par][crank], (uint32_t)lane);
                    const uint32_t rbar = map_to_cta(&sm.mbar[par], (uint32_t)lane);
                    st_async_v4(dst, rbar, clo.x, clo.y, clo.z, clo.w);
                    st_async_v4(dst + 16, rbar, chi.x, chi.y, chi.z, chi.w);
                }
            }
            mbar_wait_parity(&sm.mbar[par], ((e - 1u) >> 1) & 1u);
            A = fps_select<SPEC>(&sm.cluster_slot[par][0], (int)S, lane, spec_cap, j, m, g == 0, out, wx, wy, wz);
            j += A;
            fps_mark(j, 6);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) g_fps_exchanges = e - 1u;
    // temp is an in/out buffer of the reference ABI: it ends as the running distance to samples 0 .. m-2 (the reference's round
    // updates BEFORE it selects, so its last sample is never applied, :133-170).  The last exchange may have chosen several
    // samples (j + A == m, the last of them is sample m-1): apply all but that one.
    if (trow) {
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            if (SPEC > 1 && A > 1) {
                float x, y, z;
                if (XYZ_REGS) { x = px[i]; y = py[i]; z = pz[i]; }
                else { const int slot = i * nthreads + threadIdx.x; x = sx[slot]; y = sy[slot]; z = sz[slot]; }
                float d2 = t[i];
#pragma unroll
                for (int a = 0; a < SPEC - 1; ++a)
                    if (a + 1 < A) d2 = fminf(sqdist3(x - wx[a], y - wy[a], z - wz[a]), d2);
                t[i] = d2;
            }
            const int k = g + i * GT;
            if (k < n) trow[k] = t[i];
        }
    }
    // no CTA may exit while a peer can still write into its shared memory
    if (S > 1) { cluster_arrive_release(); cluster_wait_acquire(); }
}

// Any-size fallback: one CTA per cloud, running distances in global memory (workspace = temp or
// a caller-invisible requirement that temp != NULL).  Only used when the cloud does not fit a
// 16-CTA cluster (n > 262144).
__global__ void __launch_bounds__(1024, 1)
fps_fallback_kernel(int n, int m, int t_ref, const float *__restrict__ xyz, float *__restrict__ temp,
                    int32_t *__restrict__ idx) {
    __shared__ uint32_t s_d[2][32], s_r[2][32];
    __shared__ int32_t s_k[2][32];
    const int cloud = blockIdx.x;
    const float *p = xyz + (size_t)cloud * n * 3;
    float *trow = temp + (size_t)cloud * n;
    int32_t *out = idx + (size_t)cloud * m;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t rank_rows = (n + t_ref - 1) / t_ref;
    int old = 0;
    if (threadIdx.x == 0) out[0] = 0;
    for (int j = 1; j < m; ++j) {
        const int par = j & 1;
        const float x1 = __ldg(p + (size_t)old * 3), y1 = __ldg(p + (size_t)old * 3 + 1), z1 = __ldg(p + (size_t)old * 3 + 2);
        float best = -1.f;
        int kbest = 0;
        // blockDim.x is a multiple of t_ref, so a thread's points share k mod t_ref
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const float d = sqdist3(__ldg(p + (size_t)k * 3) - x1, __ldg(p + (size_t)k * 3 + 1) - y1,
                                    __ldg(p + (size_t)k * 3 + 2) - z1);
            const float td = trow[k];
            const float d2 = fminf(d, td);
            if (d2 != td) trow[k] = d2;
            if (d2 > best) { best = d2; kbest = k; }
        }
        const bool has = best >= 0.f;
        const uint32_t dkey = has ? __float_as_uint(best) : 0u;
        const uint32_t rank = has ? (uint32_t)(kbest & (t_ref - 1)) * rank_rows + (uint32_t)(kbest / t_ref) : 0xffffffffu;
        uint32_t dmax, rmin;
        if (warp_argmax(dkey, rank, dmax, rmin)) { s_d[par][warp] = dmax; s_r[par][warp] = rmin; s_k[par][warp] = kbest; }
        if (!__any_sync(0xffffffffu, has) && lane == 0) { s_d[par][warp] = 0u; s_r[par][warp] = 0xffffffffu; s_k[par][warp] = 0; }
        __syncthreads();
        const uint32_t wd = lane < nwarps ? s_d[par][lane] : 0u;
        const uint32_t wr = lane < nwarps ? s_r[par][lane] : 0xffffffffu;
        const bool wwin = warp_argmax(wd, wr, dmax, rmin);
        const uint32_t wsrc = __ffs(__ballot_sync(0xffffffffu, wwin)) - 1;
        old = wsrc != 0xffffffffu ? s_k[par][wsrc] : 0;
        if (threadIdx.x == 0) out[j] = old;
    }
}

static int floor_pow2(int v) {
    int p = 1;
    while (p * 2 <= v) p *= 2;
    return p;
}
static int ceil_pow2(int v) {
    int p = 1;
    while (p < v) p *= 2;
    return p;
}

// the reference's opt_n_threads (cuda_utils.h:9-14); n >= 1
static int ref_block_size(int n) {
    int t = floor_pow2(n);
    return t > 512 ? 512 : t;
}

template <int PPT, bool XYZ_REGS, int MAX_THREADS, bool DIRECT>
static int launch_fps_impl(int b, int n, int m, const int32_t *n_arr, const int32_t *m_arr, int S, int threads,
                           const float *xyz, float *temp, int32_t *idx, cudaStream_t stream);

template <int PPT, bool XYZ_REGS, int MAX_THREADS>
static int launch_fps(int b, int n, int m, const int32_t *n_arr, const int32_t *m_arr, int S, int threads,
                      const float *xyz, float *temp, int32_t *idx, cudaStream_t stream) {
    // every warp of the cluster exchanges directly while the candidates stay few (<= 64 warps); wider clusters reduce per CTA first
    if (S * (threads / 32) <= 64)
        return launch_fps_impl<PPT, XYZ_REGS, MAX_THREADS, true>(b, n, m, n_arr, m_arr, S, threads, xyz, temp, idx, stream);
    return launch_fps_impl<PPT, XYZ_REGS, MAX_THREADS, false>(b, n, m, n_arr, m_arr, S, threads, xyz, temp, idx, stream);
}

template <int PPT, bool XYZ_REGS, int MAX_THREADS, bool DIRECT>
static int launch_fps_impl(int b, int n, int m, const int32_t *n_arr, const int32_t *m_arr, int S, int threads,
                           const float *xyz, float *temp, int32_t *idx, cudaStream_t stream) {
    auto kern = fps_kernel<PPT, XYZ_REGS, MAX_THREADS, DIRECT>;
    if (S > 1 && S * (threads / 32) > FPS_MAX_CAND) {
        set_error("fps: internal: cluster of %d CTAs x %d warps exceeds %d candidate slots", S, threads / 32, FPS_MAX_CAND);
        return PU3_E_UNSUPPORTED;
    }
    const size_t smem = sizeof(FpsSmem) + (size_t)threads * PPT * 3 * sizeof(float);
    int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                         "fps: set max dynamic smem");
    if (st) return st;
    if (S > 8) {
        st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1),
                         "fps: allow cluster size 16");
        if (st) return st;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(b * S));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cuda_status(cudaLaunchKernelEx(&cfg, kern, n, m, n_arr, m_arr, xyz, temp, idx), "fps_kernel launch");
}

}  // namespace pu3

using namespace pu3;

// Test/tuning hook: force the cluster size (0 = heuristic).  Not part of the public header.
static int g_fps_force_cluster = 0;
static int g_fps_force_threads = 0;
static int g_fps_sm_budget = 0;   // SMs FPS may spread over (0 = all): leaves room for kernels of a concurrent stream
extern "C" void pu3_fps_set_sm_budget(int n) { g_fps_sm_budget = n; }
#ifdef PU3_FPS_TIMELINE
extern "C" void pu3_fps_set_timeline(unsigned int *buf) { cudaMemcpyToSymbol(pu3::g_fps_timeline, &buf, sizeof(buf)); }   // bring-up hook
#endif
extern "C" void pu3_fps_set_cluster(int s) { g_fps_force_cluster = s; }
extern "C" void pu3_fps_set_threads(int t) { g_fps_force_threads = t; }   // tuning hook (profiles/tune_fps.py)

static int fps_dispatch(int b, int n, int m, const int32_t *n_arr, const int32_t *m_arr, const float *xyz, float *temp,
                        int32_t *idx, pu3_stream_t stream);

extern "C" void pu3_fps_set_select_all(int on) { cudaMemcpyToSymbol(pu3::g_fps_select_all, &on, sizeof(int)); }
extern "C" void pu3_fps_set_spec(int depth) { cudaMemcpyToSymbol(pu3::g_fps_spec_cap, &depth, sizeof(int)); }
extern "C" unsigned int pu3_fps_last_exchanges(void) {
    unsigned int v = 0;
    cudaMemcpyFromSymbol(&v, pu3::g_fps_exchanges, sizeof(v));
    return v;
}

extern "C" int pu3_fps_f32(int b, int n, int m, const float *xyz, float *temp, int32_t *idx,
                           pu3_stream_t stream) {
    return fps_dispatch(b, n, m, nullptr, nullptr, xyz, temp, idx, stream);
}

extern "C" int pu3_fps_ragged_f32(int b, int n_stride, int m_stride, const int32_t *n_arr, const int32_t *m_arr,
                                  const float *xyz, float *temp, int32_t *idx, pu3_stream_t stream) {
    return fps_dispatch(b, n_stride, m_stride, n_arr, m_arr, xyz, temp, idx, stream);
}

static int fps_dispatch(int b, int n, int m, const int32_t *n_arr, const int32_t *m_arr, const float *xyz, float *temp,
                        int32_t *idx, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && n >= 0 && m >= 0, "fps: negative size b=%d n=%d m=%d", b, n, m);
    if (b == 0 || m == 0) return PU3_OK;  // the reference kernel returns at once for m <= 0 (:106)
    PU3_ARG_CHECK(n > 0, "fps: sampling %d points from an empty cloud", m);
    PU3_ARG_CHECK(xyz && idx, "fps: null pointer");
    cudaStream_t s = as_stream(stream);
    const int t_ref = ref_block_size(n);
    const int sms = device_info().sm_count;

    // Launch shape.  A round costs every warp (10 instructions per point + ~45 of reduction / hand-shake) issue
    // slots, plus ~350 cycles of DSMEM exchange when the cloud is spread over S > 1 CTAs.  Few fat warps beat many
    // thin ones (the per-warp overhead is paid by every warp): pick the (threads, S, points per thread) with the
    // smallest modelled round among the shapes that (a) keep S*threads a multiple of the reference block size
    // (tie rule), (b) hold the cloud in registers (<= 24 points per thread, fewer at high thread counts),
    // (c) keep all clouds co-resident (b*S <= SMs).
    // Two CTAs of DIFFERENT clouds per SM let one cloud's serial phase (reduction + exchange, issue slots idle)
    // overlap the other's point loop, so the cluster may be twice as wide as "one CTA per SM" allows, provided
    // two CTAs fit an SM (threads <= 512 and 2 * threads * registers <= 64 K).
    const int sm_budget = g_fps_sm_budget > 0 ? (g_fps_sm_budget < sms ? g_fps_sm_budget : sms) : sms;
    int s_cap = floor_pow2((2 * sm_budget) / b > 0 ? (2 * sm_budget) / b : 1);
    if (s_cap > 8) s_cap = 8;
    int threads = 0, S = 1, ppt = 0;
    long long best_cost = -1;
    static const int kThreads[] = {1024, 896, 768, 640, 512, 384, 256, 128, 64, 32};
    static const int kPpt[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 14, 16, 20, 24};
    auto ppt_slot = [&](long long need, int th) -> int {   // smallest instantiated points-per-thread >= need that fits the registers
        for (int v : kPpt) {
            if (v < need) continue;
            if (v > 8 && th > 512) return 0;
            if (v > 16 && th > 384) return 0;
            return v;
        }
        return 0;
    };
    for (int cs = 1; cs <= s_cap; cs *= 2) {
        if (g_fps_force_cluster > 0 && cs != g_fps_force_cluster) continue;
        const bool two_per_sm = (long long)b * cs > sm_budget;
        for (int th : kThreads) {
            if (g_fps_force_threads > 0 && th != g_fps_force_threads) continue;
            const long long gt = (long long)cs * th;
            if (gt % t_ref != 0) continue;
            const int p = ppt_slot((n + gt - 1) / gt, th);
            if (p == 0) continue;
            if (two_per_sm && (th > 512 || (long long)2 * th * (4 * p + 40) > 65536)) continue;
            const long long issue = (long long)(th / 32) * (p * 10 + 45) / 4;
            const long long lat = cs > 1 ? 1100 : 600;
            const long long cost = two_per_sm ? (2 * issue > issue + lat ? 2 * issue : issue + lat) + 100 : issue + lat;
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; threads = th; S = cs; ppt = p; }
        }
    }
    if (g_fps_force_cluster > 0 && threads == 0) {   // forced cluster size (tests): any shape that fits
        S = g_fps_force_cluster;
        for (int th : kThreads) {
            const long long gt = (long long)S * th;
            if (gt % t_ref != 0) continue;
            const int p = ppt_slot((n + gt - 1) / gt, th);
            if (p) { threads = th; ppt = p; break; }
        }
    }
    int st;
    if (threads > 0) {
#define PU3_FPS_CASE(P, MT) case P: st = launch_fps<P, true, MT>(b, n, m, n_arr, m_arr, S, threads, xyz, temp, idx, s); break;
        switch (ppt) {
            PU3_FPS_CASE(1, 1024) PU3_FPS_CASE(2, 1024) PU3_FPS_CASE(3, 1024) PU3_FPS_CASE(4, 1024)
            PU3_FPS_CASE(5, 1024) PU3_FPS_CASE(6, 1024) PU3_FPS_CASE(7, 1024) PU3_FPS_CASE(8, 1024)
            PU3_FPS_CASE(10, 512) PU3_FPS_CASE(12, 512) PU3_FPS_CASE(14, 512) PU3_FPS_CASE(16, 512)
            PU3_FPS_CASE(20, 384) PU3_FPS_CASE(24, 384)
            default: set_error("fps: internal: no kernel for %d points per thread", ppt); return PU3_E_UNSUPPORTED;
        }
#undef PU3_FPS_CASE
        return st;
    }
    // does not fit 8 points per thread in the allowed cluster: shared-memory coordinate variants, up to 32 per thread
    S = 1;
    while (S < 16 && (long long)S * 512 * 32 < n) S *= 2;
    if ((long long)S * 512 * 32 >= n) {
        ppt = (int)((n + (long long)S * 512 - 1) / ((long long)S * 512));
        if (ppt <= 16) st = launch_fps<16, false, 512>(b, n, m, n_arr, m_arr, S, 512, xyz, temp, idx, s);
        else st = launch_fps<32, false, 512>(b, n, m, n_arr, m_arr, S, 512, xyz, temp, idx, s);
    } else {
        PU3_ARG_CHECK(n_arr == nullptr && m_arr == nullptr, "fps: ragged batches beyond 262144 points per cloud are not supported");
        PU3_ARG_CHECK(temp != nullptr, "fps: n=%d needs the temp buffer (clouds beyond 262144 points run from global memory)", n);
        fps_fallback_kernel<<<b, 1024, 0, s>>>(n, m, t_ref, xyz, temp, idx);
        st = cuda_status(cudaGetLastError(), "fps_fallback_kernel");
    }
    return st;
}
