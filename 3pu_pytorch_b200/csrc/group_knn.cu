// k nearest neighbours + neighbour grouping, sm_100a.
//
// Replaces network.operations.group_knn / __batch_distance_matrix_general of the reference
// (network/operations.py:151-216), which is a chain of PyTorch library calls with a blocking
// device->host->device round trip in the middle:
//     matmul + 2 broadcast adds  -> (B,M,N) distance matrix in HBM          (:158-161, :191)
//     points.cpu().numpy(), np.unique per cloud, back to the device         (:192-204)
//     topk(-D, k, sorted)         -> radix select + sort over the matrix    (:207)
//     gather on an expanded view  -> (B,M,k,C), returned as a permuted view (:209-214)
// Here the distance matrix never exists: distances are produced tile by tile from shared memory
// and consumed by the selection in registers; duplicate detection runs on the device; the
// neighbour features are written once, already in (B,C,M,k) order.
//
// Semantics kept from the reference:
//   D = |q|^2 - 2 q.p + |p|^2 in fp32 ("expanded form", can be slightly negative)
//   unique: every point equal (all channels) to an EARLIER point of its cloud gets max(D) added
//           (np.unique keeps first occurrences; max over the reference's whole batch, here over
//           `max_group` consecutive batch elements so that independent requests can share a launch)
//   output sorted by ascending D; torch.topk leaves the order of equal distances unspecified,
//   here equal distances are ordered by ascending point index.
//
// Kernels
//   knn_dup_kernel       duplicate flags per point + "group has duplicates" flags
//   knn_maxd_kernel      max(D) per group; exits at once when its group has no duplicates
//   knn_small_kernel     k <= 64: one warp per query, candidates streamed from a shared-memory
//                        tile, top-k kept sorted across the lanes of the warp
//   knn_large_kernel     k  > 64: one CTA per query, keys in shared memory, radix select of the
//                        k-th key, ordered compaction, bitonic sort of the k survivors
#include "pu3_common.cuh"

namespace pu3 {

struct KnnArgs {
    int b, c, m, n, k, p_div, max_group;
    // ragged batches (all optional): cloud read by a batch element, duplicate-penalty group of a batch element,
    // valid points per cloud (<= n), valid queries per batch element (<= m); n and m stay the row strides
    const int32_t *owner, *group_of, *n_arr, *m_arr;
    const float *query;   // (b,c,m)
    const float *points;  // (b/p_div,c,n)
    const uint8_t *dup;   // (b/p_div,n) or null
    const int *cloud_dups; // (clouds) number of duplicate points per cloud, or null
    const int *group_any; // (groups) or null: group needs the exact max(D) penalty
    const uint32_t *maxd; // (groups) ordered keys
    float *knn;           // (b,c,m,k) or null
    int64_t *idx64;       // (b,m,k) or null
    int32_t *idx32;       // (b,m,k) or null
    float *dist;          // (b,m,k) or null
    int exact_pops;       // knn_feat_kernel, indices-only mode: 1 = ranks 1..k-1 in exact order too (test hook)
    int prefilter;        // knn_thread_kernel: exact bounding-sphere candidate pre-filter allowed (queries are not the cloud itself)
    int fused_dup;        // knn_feat_kernel: `unique` requested and NO duplicate pre-pass ran -- the kernel finds the duplicates itself
    int grid_target;      // knn_grid_kernel: candidates per cell the grid is sized for
};

__device__ __forceinline__ int knn_cloud(const KnnArgs &a, int bi) { return a.owner ? __ldg(a.owner + bi) : bi / a.p_div; }
__device__ __forceinline__ int knn_group(const KnnArgs &a, int bi) { return a.group_of ? __ldg(a.group_of + bi) : bi / a.max_group; }
__device__ __forceinline__ int knn_n(const KnnArgs &a, int cloud) { return a.n_arr ? min(a.n, __ldg(a.n_arr + cloud)) : a.n; }
__device__ __forceinline__ int knn_m(const KnnArgs &a, int bi) { return a.m_arr ? min(a.m, __ldg(a.m_arr + bi)) : a.m; }

// Duplicate handling of one batch element (operations.py:192-204).  The reference adds max(D) to every
// duplicate candidate, which sorts ALL duplicates after ALL first occurrences (max(D) >= any D).  So:
//   mode 0  no duplicates in the cloud: nothing to do
//   mode 1  the cloud has at least k first occurrences: duplicates can never be selected -> skip them
//           (no max(D) pass at all; the merged previous-level clouds of the eval path are full of duplicates)
//   mode 2  fewer than k first occurrences: exact penalty, max(D) computed by knn_maxd_kernel for the group
__device__ __forceinline__ int knn_dup_mode(const KnnArgs &a, int cloud, int grp) {
    if (a.dup == nullptr) return 0;
    if (a.group_any[grp] != 0) return 2;
    return a.cloud_dups[cloud] > 0 ? 1 : 0;
}

// the reference's D = r_A - 2*m + r_B, evaluated left to right (operations.py:161)
// (2 * dot is exact in binary floating point, so rq - 2 * dot rounds once either way: the fused form is bit-identical and one
// instruction shorter)
__device__ __forceinline__ float expanded_dist(float rq, float dot, float rp) {
    return __fadd_rn(__fmaf_rn(-2.0f, dot, rq), rp);
}

// --------------------------------------------------------------------------------------------
// duplicates
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) knn_dup_kernel(int c, int n, const int32_t *__restrict__ n_arr,
                                                     const float *__restrict__ points,
                                                     uint8_t *__restrict__ dup, int *__restrict__ cloud_any) {
    const int cloud = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int nv = n_arr ? min(n, __ldg(n_arr + cloud)) : n;
    if (j >= nv) return;
    const float *p = points + (size_t)cloud * c * n;
    const float v0 = p[j];
    bool found = false;
    for (int e = 0; e < j && !found; ++e) {
        if (__ldg(p + e) != v0) continue;  // warp-uniform address: one broadcast load
        bool same = true;
        for (int ch = 1; ch < c && same; ++ch) same = p[(size_t)ch * n + e] == p[(size_t)ch * n + j];
        found = same;
    }
    dup[(size_t)cloud * n + j] = found ? 1 : 0;
    if (found) atomicAdd(cloud_any + cloud, 1);   // number of duplicates of the cloud
}


// Clouds of <= 1024 points (every feature-space search): one CTA per cloud, a 32-bit hash of each point's
// channels in shared memory, candidates compared hash first.  Comparing channel 0 first (kernel above) is
// useless on post-ReLU features, where a third of the values are exactly 0.
constexpr int KD_MAXN = 1024;
__global__ void __launch_bounds__(256) knn_dup_small_kernel(int c, int n, const int32_t *__restrict__ n_arr,
                                                           const float *__restrict__ points,
                                                           uint8_t *__restrict__ dup, int *__restrict__ cloud_any) {
    __shared__ uint32_t hsh[KD_MAXN];
    const int cloud = blockIdx.x;
    const int nv = n_arr ? min(n, __ldg(n_arr + cloud)) : n;
    const float *p = points + (size_t)cloud * c * n;
    for (int j = threadIdx.x; j < nv; j += blockDim.x) {
        uint32_t h = 2166136261u;
        for (int ch = 0; ch < c; ++ch) {
            uint32_t u = __float_as_uint(__ldg(p + (size_t)ch * n + j));
            if ((u << 1) == 0u) u = 0u;                  // -0.0 == +0.0 (np.unique compares values)
            h = (h ^ u) * 16777619u;
            h ^= h >> 15;
        }
        hsh[j] = h;
    }
    __syncthreads();
    int found_any = 0;
    for (int j = threadIdx.x; j < nv; j += blockDim.x) {
        const uint32_t hj = hsh[j];
        bool found = false;
        for (int e = 0; e < j && !found; ++e) {
            if (hsh[e] != hj) continue;
            bool same = true;
            for (int ch = 0; ch < c && same; ++ch) same = p[(size_t)ch * n + e] == p[(size_t)ch * n + j];
            found = same;
        }
        dup[(size_t)cloud * n + j] = found ? 1 : 0;
        found_any += found ? 1 : 0;
    }
    if (found_any) atomicAdd(cloud_any + cloud, found_any);
}

// Clouds of up to 16384 points (every feature-space search: 312 points; the previous-level clouds of the skip connection:
// 3120 and 6240 points at levels 3 and 4): one CTA per cloud builds an open-addressing hash table of point indices in shared memory.  Equal points always
// meet in the same slot (slots never change their class of equal points once claimed), atomicMin keeps the smallest index
// of the class there, so "duplicate" = "not the representative of my slot" -- np.unique's first occurrence
// (operations.py:199).  O(n) instead of the O(n^2) scan of knn_dup_kernel (0.65 -> ~0.05 ms per eval step).
constexpr int KH_MAXN = 16384;
__global__ void __launch_bounds__(1024) knn_dup_hash_kernel(int c, int n, int table_size, const int32_t *__restrict__ n_arr,
                                                           const float *__restrict__ points, uint8_t *__restrict__ dup,
                                                           int *__restrict__ cloud_any) {
    extern __shared__ uint32_t table[];
    const int cloud = blockIdx.x;
    const int nv = n_arr ? min(n, __ldg(n_arr + cloud)) : n;
    const float *p = points + (size_t)cloud * c * n;
    const uint32_t mask = (uint32_t)table_size - 1u, EMPTY = 0xffffffffu;
    for (int i = threadIdx.x; i < table_size; i += blockDim.x) table[i] = EMPTY;
    __syncthreads();
    auto hash_of = [&](int j) {
        uint32_t h = 2166136261u;
        for (int ch = 0; ch < c; ++ch) {
            uint32_t u = __float_as_uint(__ldg(p + (size_t)ch * n + j));
            if ((u << 1) == 0u) u = 0u;                  // -0.0 == +0.0 (np.unique compares values)
            h = (h ^ u) * 16777619u;
            h ^= h >> 15;
        }
        return h * 2654435761u;
    };
    auto same = [&](int a, int b) {
        for (int ch = 0; ch < c; ++ch)
            if (__ldg(p + (size_t)ch * n + a) != __ldg(p + (size_t)ch * n + b)) return false;
        return true;
    };
    for (int j = threadIdx.x; j < nv; j += blockDim.x) {
        uint32_t s = hash_of(j) & mask;
        while (true) {
            uint32_t cur = table[s];
            if (cur == EMPTY) {
                cur = atomicCAS(&table[s], EMPTY, (uint32_t)j);
                if (cur == EMPTY) break;                 // claimed
            }
            if (same((int)cur, j)) { atomicMin(&table[s], (uint32_t)j); break; }
            s = (s + 1u) & mask;
        }
    }
    __syncthreads();
    int found_any = 0;
    for (int j = threadIdx.x; j < nv; j += blockDim.x) {
        uint32_t s = hash_of(j) & mask;
        uint32_t rep;
        while (true) {
            rep = table[s];
            if (rep == (uint32_t)j || same((int)rep, j)) break;
            s = (s + 1u) & mask;
        }
        const bool found = rep != (uint32_t)j;
        dup[(size_t)cloud * n + j] = found ? 1 : 0;
        found_any += found ? 1 : 0;
    }
    if (found_any) atomicAdd(cloud_any + cloud, found_any);
}

// a group needs max(D) as soon as one of its batch elements reads a cloud with duplicates
__global__ void __launch_bounds__(256) knn_groupflag_kernel(KnnArgs a, const int *__restrict__ cloud_any,
                                                           int *__restrict__ group_any) {
    const int bi = blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= a.b) return;
    const int cloud = knn_cloud(a, bi);
    const int dups = cloud_any[cloud];
    // the exact max(D) penalty is needed only when the k nearest cannot be filled with first occurrences
    if (dups > 0 && knn_n(a, cloud) - dups < a.k) group_any[knn_group(a, bi)] = 1;
}

__global__ void __launch_bounds__(128) knn_maxd_kernel(KnnArgs a, uint32_t *__restrict__ maxd) {
    const int bi = blockIdx.y;
    const int g = knn_group(a, bi);
    if (a.group_any[g] == 0) return;  // the common case: nothing to do
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t best = 0;
    const int cloud = knn_cloud(a, bi);
    const int nv = knn_n(a, cloud);
    if (qi < knn_m(a, bi)) {
        const float *q = a.query + (size_t)bi * a.c * a.m;
        const float *p = a.points + (size_t)cloud * a.c * a.n;
        float rq = 0.f;
        for (int ch = 0; ch < a.c; ++ch) { const float v = q[(size_t)ch * a.m + qi]; rq = __fmaf_rn(v, v, rq); }
        for (int j = 0; j < nv; ++j) {
            float dot = 0.f, rp = 0.f;
            for (int ch = 0; ch < a.c; ++ch) {
                const float pv = __ldg(p + (size_t)ch * a.n + j);
                dot = __fmaf_rn(q[(size_t)ch * a.m + qi], pv, dot);
                rp = __fmaf_rn(pv, pv, rp);
            }
            best = max(best, float_to_ordered(expanded_dist(rq, dot, rp)));
        }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0 && best) atomicMax(maxd + g, best);
}

// --------------------------------------------------------------------------------------------
// k <= 64: warp per query
// --------------------------------------------------------------------------------------------
constexpr int KS_WARPS = 8;
constexpr int KS_THREADS = KS_WARPS * 32;

// Sorted top-k of a warp: position p = lane*E + e holds the p-th smallest (key, index).
template <int E>
struct WarpTopK {
    uint32_t key[E];
    int32_t id[E];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int e = 0; e < E; ++e) { key[e] = 0xffffffffu; id[e] = 0; }
    }
    // value at sorted position p, broadcast to the warp
    __device__ __forceinline__ uint32_t key_at(int p) const {
        uint32_t v = key[0];
#pragma unroll
        for (int e = 1; e < E; ++e) if ((p % E) == e) v = key[e];
        return __shfl_sync(0xffffffffu, v, p / E);
    }
    // insert (kd, ki) keeping order; equal keys stay in arrival (= index) order; the last of
    // the 32*E positions falls off
    __device__ __forceinline__ void insert(uint32_t kd, int32_t ki) {
        const int lane = threadIdx.x & 31;
        int ins = 0;
#pragma unroll
        for (int e = 0; e < E; ++e) ins += __popc(__ballot_sync(0xffffffffu, key[e] <= kd));
        const uint32_t up_k = __shfl_up_sync(0xffffffffu, key[E - 1], 1);
        const int32_t up_i = __shfl_up_sync(0xffffffffu, id[E - 1], 1);
#pragma unroll
        for (int e = E - 1; e >= 0; --e) {
            const int p = lane * E + e;
            const uint32_t prev_k = e == 0 ? up_k : key[e - 1];
            const int32_t prev_i = e == 0 ? up_i : id[e - 1];
            if (p > ins) { key[e] = prev_k; id[e] = prev_i; }
            else if (p == ins) { key[e] = kd; id[e] = ki; }
        }
    }
};

// CT > 0: channel count known at compile time, query kept in registers.  CT == 0: generic.
template <int CT, int E>
__global__ void __launch_bounds__(KS_THREADS) knn_small_kernel(KnnArgs a, int tile_n, int q_per_cta) {
    extern __shared__ __align__(16) float smem[];
    const int C = CT > 0 ? CT : a.c;
    float *sp = smem;                      // [C][tile_n] candidate tile, channel-major
    float *srp = sp + (size_t)C * tile_n;  // [tile_n] squared norms
    float *sq = srp + tile_n;              // [KS_WARPS][C] queries (generic path)

    const int bi = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cloud = knn_cloud(a, bi);
    const int grp = knn_group(a, bi);
    const int nv = knn_n(a, cloud);          // valid candidates of this cloud
    const float *qb = a.query + (size_t)bi * C * a.m;
    const float *pb = a.points + (size_t)cloud * C * a.n;
    const int dmode = knn_dup_mode(a, cloud, grp);
    const bool penal = dmode != 0;
    const float maxd = dmode == 2 ? ordered_to_float(a.maxd[grp]) : 0.f;
    const uint8_t *dupb = penal ? a.dup + (size_t)cloud * a.n : nullptr;

    const int q_begin = blockIdx.x * q_per_cta;
    const int q_end = min(knn_m(a, bi), q_begin + q_per_cta);
    if (q_begin >= q_end) return;            // block-uniform
    const int passes = (q_end - q_begin + KS_WARPS - 1) / KS_WARPS;

    for (int pass = 0; pass < passes; ++pass) {
        const int qi = q_begin + pass * KS_WARPS + warp;
        const bool active = qi < q_end;  // warp-uniform
        float qreg[CT > 0 ? CT : 1];
        float rq = 0.f;
        if (active) {
            if (CT > 0) {
#pragma unroll
                for (int ch = 0; ch < (CT > 0 ? CT : 1); ++ch) {
                    qreg[ch] = __ldg(qb + (size_t)ch * a.m + qi);
                    rq = __fmaf_rn(qreg[ch], qreg[ch], rq);
                }
            } else {
                for (int ch = lane; ch < C; ch += 32) sq[warp * C + ch] = __ldg(qb + (size_t)ch * a.m + qi);
                __syncwarp();
                for (int ch = 0; ch < C; ++ch) rq = __fmaf_rn(sq[warp * C + ch], sq[warp * C + ch], rq);
            }
        }
        WarpTopK<E> top;
        top.init();
        uint32_t thr = 0xffffffffu;  // key at position k-1

        for (int n0 = 0; n0 < nv; n0 += tile_n) {
            const int cnt = min(tile_n, nv - n0);
            // a cloud that fits one tile is staged once per CTA, not once per pass
            if (n0 > 0 || pass == 0 || nv > tile_n) {
                __syncthreads();
                for (int t = threadIdx.x; t < cnt; t += KS_THREADS) {
                    float r = 0.f;
                    for (int ch = 0; ch < C; ++ch) {
                        const float v = __ldg(pb + (size_t)ch * a.n + n0 + t);
                        sp[(size_t)ch * tile_n + t] = v;
                        r = __fmaf_rn(v, v, r);
                    }
                    srp[t] = r;
                }
                __syncthreads();
            }
            if (!active) continue;
            for (int j0 = 0; j0 < cnt; j0 += 32) {
                const int j = j0 + lane;
                uint32_t key = 0xffffffffu;
                if (j < cnt) {
                    float dot = 0.f;
                    if (CT > 0) {
#pragma unroll
                        for (int ch = 0; ch < (CT > 0 ? CT : 1); ++ch) dot = __fmaf_rn(qreg[ch], sp[ch * tile_n + j], dot);
                    } else {
                        for (int ch = 0; ch < C; ++ch) dot = __fmaf_rn(sq[warp * C + ch], sp[(size_t)ch * tile_n + j], dot);
                    }
                    float d = expanded_dist(rq, dot, srp[j]);
                    const bool isdup = penal && dupb[n0 + j];
                    if (isdup) d = __fadd_rn(d, maxd);  // D += max(D) * duplicated (:204)
                    key = (isdup && dmode == 1) ? 0xffffffffu : float_to_ordered(d);
                }
                unsigned pend = __ballot_sync(0xffffffffu, key < thr);
                while (pend) {
                    const int src = __ffs(pend) - 1;
                    pend &= pend - 1;
                    const uint32_t kd = __shfl_sync(0xffffffffu, key, src);
                    if (kd < thr) {  // thr may have dropped since the ballot
                        top.insert(kd, n0 + j0 + src);
                        thr = top.key_at(a.k - 1);
                    }
                }
            }
        }
        if (!active) continue;
        // ---- outputs: position p = lane*E + e ----------------------------------------------
        const size_t row = ((size_t)bi * a.m + qi) * a.k;
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int p = lane * E + e;
            if (p < a.k) {
                if (a.idx64) a.idx64[row + p] = top.id[e];
                if (a.idx32) a.idx32[row + p] = top.id[e];
                if (a.dist) a.dist[row + p] = ordered_to_float(top.key[e]);
            }
        }
        if (a.knn) {
            for (int ch = 0; ch < C; ++ch) {
                float *o = a.knn + (((size_t)bi * C + ch) * a.m + qi) * a.k;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int p = lane * E + e;
                    if (p < a.k) o[p] = __ldg(pb + (size_t)ch * a.n + top.id[e]);
                }
            }
        }
    }
}


// --------------------------------------------------------------------------------------------
// k <= 64, n <= 320: the feature-space kNN of DenseEdgeConv (layers.py:33: 24 channels, 312 points, k+1 = 33)
// --------------------------------------------------------------------------------------------
// One CTA per cloud.  Queries are processed in blocks of KF_QB:
//   phase 1  the KF_QB x n block of distances as a register-tiled (8 queries x 5 candidates per thread, FFMA2)
//            product from the shared-memory copy of the cloud, stored as ordered keys in shared memory
//   phase 2  one warp per query: every lane takes the 10 keys of its column stripe, sorts them in registers
//            (29-comparator network on packed (key,stripe) words), writes the sorted keys back over its own
//            slots of the row and the warp pops the global minimum k times: redux.min on the key, redux.min on
//            the index among the lanes that hold that key, the winning lane advances its list head.
// Cost per query is independent of the data (the streaming-insertion kernel above degrades to one insertion per
// candidate on sorted input) and about 4x lower at k = 33.
constexpr int KF_QB = 32;       // queries per block
constexpr int KF_NMAX = 320;    // candidates per cloud (10 per lane)
constexpr int KF_S = KF_NMAX / 32;
constexpr int KF_THREADS = 256;
#ifndef KF_MINB
#define KF_MINB 3
#endif
constexpr int KF_JG = 64;       // candidate groups: thread tile = 8 queries x 5 candidates (jg, jg+64, ..., jg+256)

__device__ __forceinline__ void cex(unsigned long long &a, unsigned long long &b) {
    const unsigned long long lo = a < b ? a : b, hi = a < b ? b : a;
    a = lo; b = hi;
}

template <int CT, bool HOT>
__global__ void __launch_bounds__(KF_THREADS, KF_MINB) knn_feat_kernel(KnnArgs a) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int C = CT > 0 ? CT : a.c;
    float *sx = reinterpret_cast<float *>(raw);                        // [C][KF_NMAX] cloud, channel-major, zero padded
    float *snorm = sx + (size_t)C * KF_NMAX;                           // [KF_NMAX]
    uint32_t *skeys = reinterpret_cast<uint32_t *>(snorm + KF_NMAX);   // [KF_QB][KF_NMAX]

    const int bi = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = knn_cloud(a, bi), grp = knn_group(a, bi);
    const int nv = knn_n(a, cloud), mv = knn_m(a, bi);
    const float *qb = a.query + (size_t)bi * C * a.m;
    const float *pb = a.points + (size_t)cloud * C * a.n;
    const bool self = (qb == pb) && (a.m == a.n);                      // DenseEdgeConv: queries are the cloud itself
    int dmode = knn_dup_mode(a, cloud, grp);
    bool penal = dmode != 0;
    float maxd = dmode == 2 ? ordered_to_float(a.maxd[grp]) : 0.f;
    const uint8_t *dupb = penal ? a.dup + (size_t)cloud * a.n : nullptr;

    // ---- stage the cloud: TMA bulk copies (one per channel row, completing on an mbarrier) when the rows are 16-byte aligned
    // -- n = 312: 1248-byte rows -- with the row tails and the zero padding up to KF_NMAX by ordinary stores
    __shared__ uint64_t tma_bar;
    const bool bulk_ok = ((reinterpret_cast<uintptr_t>(pb) & 15u) == 0) && ((a.n & 3) == 0);
    const int nb = bulk_ok ? (nv & ~3) : 0;                  // floats per row moved by the copy engine
    if (nb > 0) {
        if (tid == 0) { mbar_init(&tma_bar, 1); mbar_fence_init_cluster(); }
        __syncthreads();
        if (tid == 0) mbar_arrive_expect_tx(&tma_bar, (uint32_t)(C * nb * 4));
        if (tid < C) bulk_copy_g2s(sx + tid * KF_NMAX, pb + (size_t)tid * a.n, (uint32_t)(nb * 4), &tma_bar);
        for (int ch = KF_THREADS + tid; ch < C; ch += KF_THREADS)     // (generic C > 256 only)
            bulk_copy_g2s(sx + ch * KF_NMAX, pb + (size_t)ch * a.n, (uint32_t)(nb * 4), &tma_bar);
    }
    {
        const int rest = KF_NMAX - nb;
        for (int t = tid; t < C * rest; t += KF_THREADS) {
            const int ch = t / rest, j = nb + (t - ch * rest);
            sx[ch * KF_NMAX + j] = j < nv ? __ldg(pb + (size_t)ch * a.n + j) : 0.f;
        }
    }
    if (nb > 0) mbar_wait_parity(&tma_bar, 0u);
    __syncthreads();
    for (int j = tid; j < KF_NMAX; j += KF_THREADS) {
        float r = 0.f;
        for (int ch = 0; ch < C; ++ch) r = __fmaf_rn(sx[ch * KF_NMAX + j], sx[ch * KF_NMAX + j], r);
        snorm[j] = r;
    }
    __syncthreads();

    // ---- duplicates (operations.py:192-204), found HERE when no pre-pass ran: the cloud is in shared memory, a 32-bit hash per
    // point and ~n^2/2 hash compares per CTA replace three side launches per call (hash table + group flags + max(D): 57 launches
    // and 0.55 ms per eval step, profiles/r2).  A cloud with >= k first occurrences never selects a duplicate: they are dropped
    // (mode 1).  Only a degenerate cloud (< k distinct points) needs the reference's exact penalty max(D over its group): this CTA
    // then computes it itself with the arithmetic of knn_maxd_kernel -- slow, correct, and never on the hot path.
    __shared__ __align__(16) uint32_t s_hash[KF_NMAX];
    __shared__ uint8_t s_dup[KF_NMAX];
    __shared__ int s_ndup;
    __shared__ uint32_t s_maxd;
    if (a.fused_dup) {
        if (tid == 0) { s_ndup = 0; s_maxd = 0u; }
        for (int j = tid; j < nv; j += KF_THREADS) {
            uint32_t h = 2166136261u;
            for (int ch = 0; ch < C; ++ch) {
                uint32_t u = __float_as_uint(sx[ch * KF_NMAX + j]);
                if ((u << 1) == 0u) u = 0u;                  // -0.0 == +0.0 (np.unique compares values)
                h = (h ^ u) * 16777619u;
                h ^= h >> 15;
            }
            s_hash[j] = h;
        }
        __syncthreads();
        int mine = 0;
        for (int j = tid; j < nv; j += KF_THREADS) {
            const uint32_t hj = s_hash[j];
            bool found = false;
            // four hashes per 128-bit load, independent compares: the scan is throughput-, not latency-bound
            for (int e0 = 0; e0 < j && !found; e0 += 4) {
                const uint4 h4 = *reinterpret_cast<const uint4 *>(&s_hash[e0]);
                unsigned hit = (h4.x == hj ? 1u : 0u) | (h4.y == hj ? 2u : 0u) | (h4.z == hj ? 4u : 0u) | (h4.w == hj ? 8u : 0u);
                while (hit && !found) {                  // rare: equal hashes -> compare the points, earliest first
                    const int o = __ffs(hit) - 1;
                    hit &= hit - 1u;
                    const int e = e0 + o;
                    if (e >= j) break;
                    bool same = true;
                    for (int ch = 0; ch < C && same; ++ch) same = sx[ch * KF_NMAX + e] == sx[ch * KF_NMAX + j];
                    found = same;
                }
            }
            s_dup[j] = found ? 1 : 0;
            mine += found ? 1 : 0;
        }
        if (mine) atomicAdd(&s_ndup, mine);
        __syncthreads();
        const int ndup = s_ndup;
        dmode = ndup == 0 ? 0 : (nv - ndup >= a.k ? 1 : 2);
        penal = dmode != 0;
        dupb = s_dup;
        if (dmode == 2) {
            uint32_t best = 0u;
            for (int b2 = 0; b2 < a.b; ++b2) {
                if (knn_group(a, b2) != grp) continue;
                const int c2 = knn_cloud(a, b2);
                const int nv2 = knn_n(a, c2), mv2 = knn_m(a, b2);
                const float *q2 = a.query + (size_t)b2 * C * a.m;
                const float *p2 = a.points + (size_t)c2 * C * a.n;
                for (int qi = tid; qi < mv2; qi += KF_THREADS) {
                    float rq2 = 0.f;
                    for (int ch = 0; ch < C; ++ch) { const float v = q2[(size_t)ch * a.m + qi]; rq2 = __fmaf_rn(v, v, rq2); }
                    for (int j = 0; j < nv2; ++j) {
                        float dot = 0.f, rp = 0.f;
                        for (int ch = 0; ch < C; ++ch) {
                            const float pv = __ldg(p2 + (size_t)ch * a.n + j);
                            dot = __fmaf_rn(q2[(size_t)ch * a.m + qi], pv, dot);
                            rp = __fmaf_rn(pv, pv, rp);
                        }
                        best = max(best, float_to_ordered(expanded_dist(rq2, dot, rp)));
                    }
                }
            }
            best = __reduce_max_sync(0xffffffffu, best);
            if (lane == 0 && best) atomicMax(&s_maxd, best);
            __syncthreads();
            maxd = ordered_to_float(s_maxd);
        }
    }

    const int qg = tid / KF_JG, jg = tid % KF_JG;    // 4 query groups of 8 x 64 candidate groups of 5
    // few clouds (train step, level 1 of eval, single requests): gridDim.y CTAs share a cloud, each takes every gridDim.y-th
    // block of 32 queries (every CTA stages the whole cloud: 30 KB)
    for (int q0 = blockIdx.y * KF_QB; q0 < mv; q0 += KF_QB * gridDim.y) {
        // ---- phase 1: keys of queries [q0, q0+32) x candidates [0, 320): one 8x5 tile per thread, FFMA2 -------
        {
            const int ql = qg * 8;
            f32x2 acc[4][5];                          // [query pair][candidate]
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 5; ++v) acc[u][v] = 0ull;
            float rq[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) rq[u] = 0.f;
            if (self) {
                const int qq = q0 + ql;               // multiple of 8, qq + 7 < KF_NMAX (q0 < n <= 320, q0 % 32 == 0)
#pragma unroll 4
                for (int ch = 0; ch < C; ++ch) {
                    const float *row = sx + ch * KF_NMAX;
                    const float4 qa = *reinterpret_cast<const float4 *>(row + qq);
                    const float4 qc = *reinterpret_cast<const float4 *>(row + qq + 4);
                    const f32x2 q01 = pack2(qa.x, qa.y), q23 = pack2(qa.z, qa.w), q45 = pack2(qc.x, qc.y), q67 = pack2(qc.z, qc.w);
#pragma unroll
                    for (int v = 0; v < 5; ++v) {
                        const float pv = row[jg + v * KF_JG];
                        const f32x2 pp = pack2(pv, pv);
                        acc[0][v] = fma2(q01, pp, acc[0][v]); acc[1][v] = fma2(q23, pp, acc[1][v]);
                        acc[2][v] = fma2(q45, pp, acc[2][v]); acc[3][v] = fma2(q67, pp, acc[3][v]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) rq[u] = snorm[qq + u];
            } else {
                for (int ch = 0; ch < C; ++ch) {
                    const float *row = sx + ch * KF_NMAX;
                    float qv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int qi = q0 + ql + u;
                        qv[u] = qi < mv ? __ldg(qb + (size_t)ch * a.m + qi) : 0.f;
                        rq[u] = __fmaf_rn(qv[u], qv[u], rq[u]);
                    }
                    const f32x2 q01 = pack2(qv[0], qv[1]), q23 = pack2(qv[2], qv[3]), q45 = pack2(qv[4], qv[5]), q67 = pack2(qv[6], qv[7]);
#pragma unroll
                    for (int v = 0; v < 5; ++v) {
                        const float pv = row[jg + v * KF_JG];
                        const f32x2 pp = pack2(pv, pv);
                        acc[0][v] = fma2(q01, pp, acc[0][v]); acc[1][v] = fma2(q23, pp, acc[1][v]);
                        acc[2][v] = fma2(q45, pp, acc[2][v]); acc[3][v] = fma2(q67, pp, acc[3][v]);
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < 5; ++v) {
                const int j = jg + v * KF_JG;
                const float rp = snorm[j];
                const bool live = j < nv;
                const bool isdup = penal && live && dupb[j];
                const bool drop = !live || (isdup && dmode == 1);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float d0, d1;
                    unpack2(acc[u][v], d0, d1);
                    float e0 = expanded_dist(rq[2 * u], d0, rp), e1 = expanded_dist(rq[2 * u + 1], d1, rp);
                    if (isdup) { e0 = __fadd_rn(e0, maxd); e1 = __fadd_rn(e1, maxd); }
                    skeys[(ql + 2 * u) * KF_NMAX + j] = drop ? 0xffffffffu : float_to_ordered(e0);
                    skeys[(ql + 2 * u + 1) * KF_NMAX + j] = drop ? 0xffffffffu : float_to_ordered(e1);
                }
            }
        }
        __syncthreads();
        // ---- phase 2: selection, one warp per query ----------------------------------------------------
        if (HOT && a.exact_pops == 0) {
            // Indices-only calls of the fused DenseEdgeConv (rank 0 exact, ranks 1..k-1 as a set, see below): a warp works on TWO
            // queries at once (ql and ql + 8).  The k-1 rounds are a dependent chain of {reduce-min, compare, shared-memory load}
            // per query -- ~60 cycles of latency per round against ~7 instructions -- so independent chains interleaved in the
            // same warp fill each other's bubbles (two at once: 5.08 -> 4.85 ms per step) (the per-lane sorting networks of the two queries interleave the same way).
            constexpr int NZ = 2;                                     // queries a warp works on at once (4: 4.89 ms, 2: 4.85 ms, 1: 5.08 ms per step)
            for (int base = 0; base < KF_QB && q0 + base + warp < mv; base += NZ * (KF_THREADS / 32)) {
                int qlz[NZ];
                uint32_t *krowz[NZ];
                int nz = 0;
#pragma unroll
                for (int z = 0; z < NZ; ++z) {
                    qlz[z] = base + warp + z * (KF_THREADS / 32);
                    krowz[z] = skeys + qlz[z] * KF_NMAX;
                    if (q0 + qlz[z] < mv) nz = z + 1;
                }
                uint32_t hk[NZ], perm_lo[NZ], perm_hi[NZ];
#pragma unroll
                for (int z = 0; z < NZ; ++z) {
                    hk[z] = 0xffffffffu; perm_lo[z] = 0; perm_hi[z] = 0;
                    if (z < nz) {
                        uint32_t *krow = krowz[z];
                        // The lane's 10 keys are first sorted on 32-bit words (key with its low 4 bits replaced by the stripe: a
                        // compare-exchange is one IMNMX pair instead of a 64-bit compare and two 64-bit selects), the exact keys are
                        // fetched in that order and checked; a lane with two keys closer than 16 ulp may come out in the wrong
                        // order, and then (any lane of the warp) the query is sorted again on the packed 64-bit words.
                        uint32_t w[KF_S];
#pragma unroll
                        for (int s2 = 0; s2 < KF_S; ++s2) w[s2] = (krow[s2 * 32 + lane] & ~15u) | (uint32_t)s2;
#define PU3_CEX32(a_, b_) do { const uint32_t lo_ = min(a_, b_), hi_ = max(a_, b_); a_ = lo_; b_ = hi_; } while (0)
                        PU3_CEX32(w[4], w[9]); PU3_CEX32(w[3], w[8]); PU3_CEX32(w[2], w[7]); PU3_CEX32(w[1], w[6]); PU3_CEX32(w[0], w[5]);
                        PU3_CEX32(w[1], w[4]); PU3_CEX32(w[6], w[9]); PU3_CEX32(w[0], w[3]); PU3_CEX32(w[5], w[8]);
                        PU3_CEX32(w[0], w[2]); PU3_CEX32(w[3], w[6]); PU3_CEX32(w[7], w[9]);
                        PU3_CEX32(w[0], w[1]); PU3_CEX32(w[2], w[4]); PU3_CEX32(w[5], w[7]); PU3_CEX32(w[8], w[9]);
                        PU3_CEX32(w[1], w[2]); PU3_CEX32(w[4], w[6]); PU3_CEX32(w[7], w[8]); PU3_CEX32(w[3], w[5]);
                        PU3_CEX32(w[2], w[5]); PU3_CEX32(w[6], w[8]); PU3_CEX32(w[1], w[3]); PU3_CEX32(w[4], w[7]);
                        PU3_CEX32(w[2], w[3]); PU3_CEX32(w[6], w[7]);
                        PU3_CEX32(w[3], w[4]); PU3_CEX32(w[5], w[6]);
                        PU3_CEX32(w[4], w[5]);
#undef PU3_CEX32
                        uint32_t ex[KF_S];
#pragma unroll
                        for (int s2 = 0; s2 < KF_S; ++s2) ex[s2] = krow[(w[s2] & 15u) * 32 + lane];
                        bool ok = true;
#pragma unroll
                        for (int s2 = 0; s2 + 1 < KF_S; ++s2)
                            ok = ok & ((ex[s2] < ex[s2 + 1]) | ((ex[s2] == ex[s2 + 1]) & ((w[s2] & 15u) < (w[s2 + 1] & 15u))));
                        if (__all_sync(0xffffffffu, ok)) {
#pragma unroll
                            for (int s2 = 0; s2 < KF_S; ++s2) {
                                const uint32_t tag = w[s2] & 15u;
                                if (s2 < 8) perm_lo[z] |= tag << (4 * s2); else perm_hi[z] |= tag << (4 * (s2 - 8));
                                if (s2 > 0) krow[s2 * 32 + lane] = ex[s2];
                            }
                            hk[z] = ex[0];
                        } else {
                            unsigned long long v[KF_S];
#pragma unroll
                            for (int s2 = 0; s2 < KF_S; ++s2) v[s2] = ((unsigned long long)krow[s2 * 32 + lane] << 32) | (uint32_t)s2;
                            cex(v[4], v[9]); cex(v[3], v[8]); cex(v[2], v[7]); cex(v[1], v[6]); cex(v[0], v[5]);
                            cex(v[1], v[4]); cex(v[6], v[9]); cex(v[0], v[3]); cex(v[5], v[8]);
                            cex(v[0], v[2]); cex(v[3], v[6]); cex(v[7], v[9]);
                            cex(v[0], v[1]); cex(v[2], v[4]); cex(v[5], v[7]); cex(v[8], v[9]);
                            cex(v[1], v[2]); cex(v[4], v[6]); cex(v[7], v[8]); cex(v[3], v[5]);
                            cex(v[2], v[5]); cex(v[6], v[8]); cex(v[1], v[3]); cex(v[4], v[7]);
                            cex(v[2], v[3]); cex(v[6], v[7]);
                            cex(v[3], v[4]); cex(v[5], v[6]);
                            cex(v[4], v[5]);
#pragma unroll
                            for (int s2 = 0; s2 < KF_S; ++s2) {
                                const uint32_t tag = (uint32_t)v[s2] & 15u;
                                if (s2 < 8) perm_lo[z] |= tag << (4 * s2); else perm_hi[z] |= tag << (4 * (s2 - 8));
                                if (s2 > 0) krow[s2 * 32 + lane] = (uint32_t)(v[s2] >> 32);
                            }
                            hk[z] = (uint32_t)(v[0] >> 32);
                        }
                    }
                }
                uint32_t hk_first[NZ];
#pragma unroll
                for (int z = 0; z < NZ; ++z) hk_first[z] = hk[z];
                __syncwarp();
                // rank 0 exactly (key, then lowest index); then k-1 cheap rounds: every lane whose head equals the warp minimum advances
                bool first[NZ];
                int cnt[NZ];
#pragma unroll
                for (int z = 0; z < NZ; ++z) {
                    const uint32_t hj = (perm_lo[z] & 15u) * 32u + lane;
                    const uint32_t kmin0 = __reduce_min_sync(0xffffffffu, hk[z]);
                    const uint32_t jmin0 = __reduce_min_sync(0xffffffffu, hk[z] == kmin0 ? hj : 0xffffffffu);
                    first[z] = (hj == jmin0 && hk[z] == kmin0);
                    cnt[z] = 0;
                    if (first[z]) { hk[z] = krowz[z][32 + lane]; cnt[z] = 1; }
                }
                for (int r = 1; r < a.k; ++r) {
#pragma unroll
                    for (int z = 0; z < NZ; ++z) {
                        const uint32_t kmin = __reduce_min_sync(0xffffffffu, hk[z]);
                        if (hk[z] == kmin) {
                            ++cnt[z];                             // index clamped: a lane that runs out of keys (or massive ties)
                            hk[z] = krowz[z][min(cnt[z], KF_S - 1) * 32 + lane];   // re-reads its last key; that case is caught below
                        }
                    }
                }
#pragma unroll
                for (int z = 0; z < NZ; ++z) {
                    if (z >= nz) continue;                            // warp-uniform
                    uint32_t *krow = krowz[z];
                    uint16_t *stage = reinterpret_cast<uint16_t *>(krow);   // stripe 0 of the row is free: heads live in registers
                    const int total = __reduce_add_sync(0xffffffffu, cnt[z]);
                    const int most = __reduce_max_sync(0xffffffffu, cnt[z]);
                    if (total == a.k && most < KF_S) {        // exactly k pops, and no lane exhausted its 10 keys
                        const int mine = cnt[z] - (first[z] ? 1 : 0);
                        int incl = mine;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const int up = __shfl_up_sync(0xffffffffu, incl, d);
                            if (lane >= d) incl += up;
                        }
                        int pos = 1 + incl - mine;
                        unsigned long long perm = ((unsigned long long)perm_hi[z] << 32) | perm_lo[z];
                        if (first[z]) { stage[0] = (uint16_t)(((uint32_t)perm & 15u) * 32u + lane); perm >>= 4; }
                        for (int e = 0; e < mine; ++e) {
                            stage[pos + e] = (uint16_t)(((uint32_t)perm & 15u) * 32u + lane);
                            perm >>= 4;
                        }
                    } else {
                        // ties straddling rank k-1 (duplicates, clamped distances) or an exhausted lane: exact pops
                        uint32_t hkx = hk_first[z], plo = perm_lo[z], phi = perm_hi[z];
                        uint32_t hjx = (plo & 15u) * 32u + lane;
                        int h = 1;
                        for (int r = 0; r < a.k; ++r) {
                            const uint32_t kmin = __reduce_min_sync(0xffffffffu, hkx);
                            const uint32_t jmin = __reduce_min_sync(0xffffffffu, hkx == kmin ? hjx : 0xffffffffu);
                            if (hjx == jmin && hkx == kmin) {       // exactly one lane (indices are unique)
                                stage[r] = (uint16_t)hjx;
                                hkx = h < KF_S ? krow[h * 32 + lane] : 0xffffffffu;
                                plo = __funnelshift_r(plo, phi, 4);
                                phi >>= 4;
                                hjx = (plo & 15u) * 32u + lane;
                                ++h;
                            }
                        }
                    }
                    __syncwarp();
                    int32_t *o = a.idx32 + ((size_t)bi * a.m + q0 + qlz[z]) * a.k;
                    for (int r = lane; r < a.k; r += 32) o[r] = stage[r];
                    __syncwarp();
                }
            }
        } else
        for (int ql = warp; ql < KF_QB; ql += KF_THREADS / 32) {
            const int qi = q0 + ql;
            if (qi >= mv) break;  // warp-uniform
            uint32_t *krow = skeys + ql * KF_NMAX;
            // packed (key, stripe): a candidate's index is stripe*32 + lane, so 4 bits identify it inside the lane
            unsigned long long v[KF_S];
#pragma unroll
            for (int s = 0; s < KF_S; ++s) v[s] = ((unsigned long long)krow[s * 32 + lane] << 32) | (uint32_t)s;
            // optimal 29-comparator sorting network for 10 inputs
            cex(v[4], v[9]); cex(v[3], v[8]); cex(v[2], v[7]); cex(v[1], v[6]); cex(v[0], v[5]);
            cex(v[1], v[4]); cex(v[6], v[9]); cex(v[0], v[3]); cex(v[5], v[8]);
            cex(v[0], v[2]); cex(v[3], v[6]); cex(v[7], v[9]);
            cex(v[0], v[1]); cex(v[2], v[4]); cex(v[5], v[7]); cex(v[8], v[9]);
            cex(v[1], v[2]); cex(v[4], v[6]); cex(v[7], v[8]); cex(v[3], v[5]);
            cex(v[2], v[5]); cex(v[6], v[8]); cex(v[1], v[3]); cex(v[4], v[7]);
            cex(v[2], v[3]); cex(v[6], v[7]);
            cex(v[3], v[4]); cex(v[5], v[6]);
            cex(v[4], v[5]);
            // the sorted keys go back into the lane's own 10 slots of the row, the stripe order into two registers
            uint32_t perm_lo = 0, perm_hi = 0;
#pragma unroll
            for (int s = 0; s < KF_S; ++s) {
                const uint32_t tag = (uint32_t)v[s] & 15u;
                if (s < 8) perm_lo |= tag << (4 * s); else perm_hi |= tag << (4 * (s - 8));
                if (s > 0) krow[s * 32 + lane] = (uint32_t)(v[s] >> 32);
            }
            uint32_t hk = (uint32_t)(v[0] >> 32);
            uint32_t hj = (perm_lo & 15u) * 32u + lane;
            int h = 1;
            const size_t row = ((size_t)bi * a.m + qi) * a.k;
            if (HOT) {
                // Only int32 indices wanted, and their consumer (the fused DenseEdgeConv) takes a max over the
                // neighbours after dropping rank 0: rank 0 must be exact, the other k-1 are needed as a SET.
                // So: one exact pop (key, then lowest index) for rank 0, then k-1 cheap rounds in which every lane
                // whose head equals the warp minimum advances (1 redux + 4 instructions instead of 2 redux + 20;
                // ncu: the exact pop loop was 47 % of the kernel's instructions).  A lane always contributes a
                // prefix of its sorted list, so the set is written afterwards from per-lane counts.  If equal keys
                // made the rounds pop more than k candidates (ties straddling rank k-1: duplicates, clamped
                // distances), the exact loop below redoes the query.
                uint16_t *stage = reinterpret_cast<uint16_t *>(krow);   // stripe 0 of the row is free: heads live in registers
                const uint32_t hk_first = hk, perm_lo_first = perm_lo, perm_hi_first = perm_hi;
                __syncwarp();
                bool exact = a.exact_pops != 0;
                if (!exact) {
                    const uint32_t kmin0 = __reduce_min_sync(0xffffffffu, hk);
                    const uint32_t jmin0 = __reduce_min_sync(0xffffffffu, hk == kmin0 ? hj : 0xffffffffu);
                    const bool first = (hj == jmin0 && hk == kmin0);
                    int cnt = 0;                                  // candidates this lane contributes
                    if (first) { hk = krow[32 + lane]; cnt = 1; }
                    for (int r = 1; r < a.k; ++r) {
                        const uint32_t kmin = __reduce_min_sync(0xffffffffu, hk);
                        if (hk == kmin) {
                            ++cnt;                                // index clamped: a lane that runs out of keys (or massive ties)
                            hk = krow[min(cnt, KF_S - 1) * 32 + lane];   // re-reads its last key; that case is caught below
                        }
                    }
                    const int total = __reduce_add_sync(0xffffffffu, cnt);
                    const int most = __reduce_max_sync(0xffffffffu, cnt);
                    if (total == a.k && most < KF_S) {        // exactly k pops, and no lane exhausted its 10 keys
                        // exclusive prefix of the per-lane counts (rank 0 goes to slot 0, outside the prefix)
                        const int mine = cnt - (first ? 1 : 0);
                        int incl = mine;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const int up = __shfl_up_sync(0xffffffffu, incl, d);
                            if (lane >= d) incl += up;
                        }
                        int pos = 1 + incl - mine;
                        unsigned long long perm = ((unsigned long long)perm_hi << 32) | perm_lo;
                        if (first) { stage[0] = (uint16_t)(((uint32_t)perm & 15u) * 32u + lane); perm >>= 4; }
                        for (int e = 0; e < mine; ++e) {
                            stage[pos + e] = (uint16_t)(((uint32_t)perm & 15u) * 32u + lane);
                            perm >>= 4;
                        }
                    } else {
                        exact = true;                             // warp-uniform
                        hk = hk_first; perm_lo = perm_lo_first; perm_hi = perm_hi_first;
                        hj = (perm_lo & 15u) * 32u + lane;
                    }
                }
                if (exact) {
                    for (int r = 0; r < a.k; ++r) {
                        const uint32_t kmin = __reduce_min_sync(0xffffffffu, hk);
                        const uint32_t jmin = __reduce_min_sync(0xffffffffu, hk == kmin ? hj : 0xffffffffu);
                        if (hj == jmin && hk == kmin) {       // exactly one lane (indices are unique)
                            stage[r] = (uint16_t)hj;
                            hk = h < KF_S ? krow[h * 32 + lane] : 0xffffffffu;
                            perm_lo = __funnelshift_r(perm_lo, perm_hi, 4);
                            perm_hi >>= 4;
                            hj = (perm_lo & 15u) * 32u + lane;
                            ++h;
                        }
                    }
                }
                __syncwarp();
                int32_t *o = a.idx32 + row;
                for (int r = lane; r < a.k; r += 32) o[r] = stage[r];
                __syncwarp();
            } else {
                for (int r = 0; r < a.k; ++r) {
                    const uint32_t kmin = __reduce_min_sync(0xffffffffu, hk);
                    const uint32_t jmin = __reduce_min_sync(0xffffffffu, hk == kmin ? hj : 0xffffffffu);
                    if (hj == jmin && hk == kmin) {
                        if (a.idx32) a.idx32[row + r] = (int32_t)hj;
                        if (a.idx64) a.idx64[row + r] = (int64_t)hj;
                        if (a.dist) a.dist[row + r] = ordered_to_float(hk);
                        if (a.knn)
                            for (int ch = 0; ch < C; ++ch)
                                a.knn[(((size_t)bi * C + ch) * a.m + qi) * a.k + r] = sx[ch * KF_NMAX + hj];
                        hk = h < KF_S ? krow[h * 32 + lane] : 0xffffffffu;
                        perm_lo = __funnelshift_r(perm_lo, perm_hi, 4);
                        perm_hi >>= 4;
                        hj = (perm_lo & 15u) * 32u + lane;
                        ++h;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// --------------------------------------------------------------------------------------------
// k <= 8, 3 channels: the xyz-space searches (inter-level skip k = 5, upsampler.py:325; outlier filter k = 2, :63)
// --------------------------------------------------------------------------------------------
// Many candidates (up to 6240 per query at level 4), tiny k: one THREAD per query keeps its k best in
// registers as a sorted list; candidates stream through shared memory as (x, y, z, |p|^2) so that the inner
// loop is one broadcast LDS.128 + 5 FP32 ops + one compare per pair, shared by KT_QPT queries per thread.
// An insertion (rare after the first few hundred candidates: ~k ln(n/k) per query) is a short bubble pass.
constexpr int KT_KMAX = 8;
constexpr int KT_THREADS = 128;
constexpr int KT_QPT = 2;
constexpr int KT_TILE = 1024;   // candidates per shared-memory tile (16 KB)

template <int KK>
__global__ void __launch_bounds__(KT_THREADS) knn_thread_kernel(KnnArgs a) {
    __shared__ float4 tile[KT_TILE];
    __shared__ int tidx[KT_TILE];
    __shared__ uint8_t tdup[KT_TILE];
    __shared__ int wtot[KT_THREADS / 32];
    const int bi = blockIdx.y;
    const int cloud = knn_cloud(a, bi), grp = knn_group(a, bi);
    const int nv = knn_n(a, cloud), mv = knn_m(a, bi);
    const int q0 = blockIdx.x * (KT_THREADS * KT_QPT);
    if (q0 >= mv) return;  // block-uniform
    const float *qb = a.query + (size_t)bi * 3 * a.m;
    const float *pb = a.points + (size_t)cloud * 3 * a.n;
    const int dmode = knn_dup_mode(a, cloud, grp);
    const bool penal = dmode != 0;
    const float maxd = dmode == 2 ? ordered_to_float(a.maxd[grp]) : 0.f;
    const uint8_t *dupb = penal ? a.dup + (size_t)cloud * a.n : nullptr;

    float qx[KT_QPT], qy[KT_QPT], qz[KT_QPT], rq[KT_QPT];
    float bd[KT_QPT][KK];
    int bj[KT_QPT][KK];
#pragma unroll
    for (int r = 0; r < KT_QPT; ++r) {
        const int qi = min(q0 + r * KT_THREADS + (int)threadIdx.x, mv - 1);
        qx[r] = __ldg(qb + qi); qy[r] = __ldg(qb + a.m + qi); qz[r] = __ldg(qb + 2 * (size_t)a.m + qi);
        rq[r] = __fmaf_rn(qz[r], qz[r], __fmaf_rn(qy[r], qy[r], __fmul_rn(qx[r], qx[r])));  // same chain as the other kernels: fma over ch 0,1,2 from 0
#pragma unroll
        for (int e = 0; e < KK; ++e) { bd[r][e] = __int_as_float(0x7f800000); bj[r][e] = 0; }
    }
    // ---- exact candidate pre-filter (skip connection: the 256 queries of a CTA are one tile's points, a small region of the
    // previous level's cloud).  Sphere around the CTA's queries: centre c = centre of their bounding box, radius
    // R = 1.5 max|q - c|.  A candidate p outside it has |p - q| > R - |q - c| for every query q of the CTA, so if q's k-th best
    // distance (found among the candidates inside) is below that bound -- checked with slack after the pass -- no excluded
    // candidate can belong to its k nearest and the result is exactly the unfiltered one.  Otherwise the CTA redoes the search
    // without the filter.  profiles/r2: 1280 x 312 queries x 6240 candidates were 2.5 G pair evaluations, issue bound.
    __shared__ float sred[6][KT_THREADS / 32];
    const bool can_filter = a.prefilter != 0 && dmode != 2 && nv >= 1024;
    float cx = 0.f, cy = 0.f, cz = 0.f, R1 = 0.f, R1sq = 0.f, dq[KT_QPT];
#pragma unroll
    for (int r = 0; r < KT_QPT; ++r) dq[r] = 0.f;
    if (can_filter) {
        float lo[3] = {qx[0], qy[0], qz[0]}, hi[3] = {qx[0], qy[0], qz[0]};
#pragma unroll
        for (int r = 1; r < KT_QPT; ++r) {
            lo[0] = fminf(lo[0], qx[r]); lo[1] = fminf(lo[1], qy[r]); lo[2] = fminf(lo[2], qz[r]);
            hi[0] = fmaxf(hi[0], qx[r]); hi[1] = fmaxf(hi[1], qy[r]); hi[2] = fmaxf(hi[2], qz[r]);
        }
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[c3] = fminf(lo[c3], __shfl_xor_sync(0xffffffffu, lo[c3], o));
                hi[c3] = fmaxf(hi[c3], __shfl_xor_sync(0xffffffffu, hi[c3], o));
            }
            if ((threadIdx.x & 31) == 0) { sred[c3][threadIdx.x >> 5] = lo[c3]; sred[3 + c3][threadIdx.x >> 5] = hi[c3]; }
        }
        __syncthreads();
#pragma unroll
        for (int w = 0; w < KT_THREADS / 32; ++w) {
            lo[0] = fminf(lo[0], sred[0][w]); lo[1] = fminf(lo[1], sred[1][w]); lo[2] = fminf(lo[2], sred[2][w]);
            hi[0] = fmaxf(hi[0], sred[3][w]); hi[1] = fmaxf(hi[1], sred[4][w]); hi[2] = fmaxf(hi[2], sred[5][w]);
        }
        cx = 0.5f * (lo[0] + hi[0]); cy = 0.5f * (lo[1] + hi[1]); cz = 0.5f * (lo[2] + hi[2]);
        // every query lies in the box, so max|q - c| <= half the box diagonal: no second reduction needed
        const float hx = 0.5f * (hi[0] - lo[0]), hy = 0.5f * (hi[1] - lo[1]), hz = 0.5f * (hi[2] - lo[2]);
        R1 = 1.5f * sqrtf(hx * hx + hy * hy + hz * hz) + 1e-6f;
        R1sq = R1 * R1;
#pragma unroll
        for (int r = 0; r < KT_QPT; ++r) {
            const float ex = qx[r] - cx, ey = qy[r] - cy, ez = qz[r] - cz;
            dq[r] = sqrtf(ex * ex + ey * ey + ez * ez);
        }
    }
    for (int pass = 0; pass < 2; ++pass) {
    const bool filt = can_filter && pass == 0;
    for (int n0 = 0; n0 < nv; n0 += KT_TILE) {
        const int src = min(KT_TILE, nv - n0);
        __syncthreads();
        int cnt;
        if (dmode == 1 || filt) {
            // stage only first occurrences (and, with the pre-filter, only candidates inside the sphere), in index order
            // (ordered compaction: ballot prefix inside the warp, warp totals through shared memory).  In the merged
            // previous-level clouds ~4 of 5 points are duplicates.
            int base = 0;
            for (int c0 = 0; c0 < src; c0 += KT_THREADS) {
                const int t = c0 + threadIdx.x;
                bool keep = t < src && (dmode != 1 || dupb[n0 + t] == 0);
                float x = 0.f, y = 0.f, z = 0.f;
                if (keep) {
                    x = __ldg(pb + n0 + t); y = __ldg(pb + a.n + n0 + t); z = __ldg(pb + 2 * (size_t)a.n + n0 + t);
                    if (filt) {
                        const float ex = x - cx, ey = y - cy, ez = z - cz;
                        keep = ex * ex + ey * ey + ez * ez <= R1sq;
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if ((threadIdx.x & 31) == 0) wtot[threadIdx.x >> 5] = __popc(bal);
                __syncthreads();
                int off = base;
                for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += wtot[w];
                int total = 0;
                for (int w = 0; w < KT_THREADS / 32; ++w) total += wtot[w];
                if (keep) {
                    const int slot = off + __popc(bal & ((1u << (threadIdx.x & 31)) - 1u));
                    tile[slot] = make_float4(x, y, z, __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
                    tidx[slot] = n0 + t;
                }
                base += total;
                __syncthreads();
            }
            cnt = base;
        } else {
            for (int t = threadIdx.x; t < src; t += KT_THREADS) {
                const float x = __ldg(pb + n0 + t), y = __ldg(pb + a.n + n0 + t), z = __ldg(pb + 2 * (size_t)a.n + n0 + t);
                tile[t] = make_float4(x, y, z, __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
                tidx[t] = n0 + t;
                if (dmode == 2) tdup[t] = dupb[n0 + t];
            }
            cnt = src;
            __syncthreads();
        }
#pragma unroll 2
        for (int j = 0; j < cnt; ++j) {
            const float4 p = tile[j];
            const bool dp = dmode == 2 && tdup[j];
#pragma unroll
            for (int r = 0; r < KT_QPT; ++r) {
                const float dot = __fmaf_rn(qz[r], p.z, __fmaf_rn(qy[r], p.y, __fmul_rn(qx[r], p.x)));
                float d = expanded_dist(rq[r], dot, p.w);
                if (dp) d = __fadd_rn(d, maxd);
                if (d < bd[r][KK - 1]) {           // strict: an equal distance with a larger index stays behind
                    bd[r][KK - 1] = d; bj[r][KK - 1] = tidx[j];
#pragma unroll
                    for (int e = KK - 1; e > 0; --e) {
                        if (bd[r][e] < bd[r][e - 1]) {
                            const float td = bd[r][e]; bd[r][e] = bd[r][e - 1]; bd[r][e - 1] = td;
                            const int tj = bj[r][e]; bj[r][e] = bj[r][e - 1]; bj[r][e - 1] = tj;
                        }
                    }
                }
            }
        }
    }
    if (!filt) break;
    // ---- verify: every query's k-th best must lie strictly inside what the sphere guarantees (1e-3 relative slack covers the
    // fp32 rounding of the expanded-form distances); a query that found fewer than k candidates (distance still +inf) fails
    bool fail = false;
#pragma unroll
    for (int r = 0; r < KT_QPT; ++r) {
        const float lim = (R1 - dq[r]) * (1.f - 1e-3f) - 1e-6f;
        fail |= !(lim > 0.f && bd[r][KK - 1] <= lim * lim);
    }
    if (!__syncthreads_or(fail)) break;
#pragma unroll
    for (int r = 0; r < KT_QPT; ++r)
#pragma unroll
        for (int e = 0; e < KK; ++e) { bd[r][e] = __int_as_float(0x7f800000); bj[r][e] = 0; }
    }
#pragma unroll
    for (int r = 0; r < KT_QPT; ++r) {
        const int qi = q0 + r * KT_THREADS + threadIdx.x;
        if (qi >= mv) continue;
        const size_t row = ((size_t)bi * a.m + qi) * a.k;
#pragma unroll
        for (int e = 0; e < KK; ++e) {
            if (e < a.k) {
                if (a.idx64) a.idx64[row + e] = bj[r][e];
                if (a.idx32) a.idx32[row + e] = bj[r][e];
                if (a.dist) a.dist[row + e] = bd[r][e];
                if (a.knn) {
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        a.knn[(((size_t)bi * 3 + ch) * a.m + qi) * a.k + e] = __ldg(pb + (size_t)ch * a.n + bj[r][e]);
                }
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// k <= 8, 3 channels, >= 1024 candidates: uniform-grid search (inter-level skip at levels 3 and 4, outlier filter of the
// merged clouds).  knn_thread_kernel evaluates every (query, candidate) pair: 0.5 G pairs for the level-4 skip search even
// after dropping the duplicates and the bounding-sphere pre-filter, fp32-issue bound.  Here a CTA
//   1. stages the cloud's candidates (first occurrences only when the cloud has duplicates) into shared memory SORTED BY CELL
//      of a G^3 grid over their bounding box (~6 per cell: count, histogram, scan, scatter -- ~1 k instructions per thread,
//      paid once per CTA and amortised over several tiles of queries: gridDim.x CTAs share a cloud's query blocks),
//   2. answers a query from the 27 cells around it (a few hundred candidates instead of thousands), with the SAME distance
//      expression and the same (distance, index) order as the exhaustive kernel, and
//   3. verifies the result: every candidate outside the scanned block is farther than the distance to the block's faces, so
//      if the k-th best distance (plus slack for the fp32 rounding of the expanded-form distances and of the cell
//      assignment) is below that bound the answer is exactly the exhaustive one; otherwise the query is redone over all cells.
// Batch elements in duplicate mode 2 (fewer than k distinct points: exact max(D) penalty) take the exhaustive path from
// global memory.  Results are bit-identical to knn_thread_kernel (tests/test_gpu_group_knn.py, A/B hook pu3_knn_set_grid).
// --------------------------------------------------------------------------------------------
constexpr int KG_TEAM = 320;      // threads of a team: one tile of 312 queries = one pass
constexpr int KG_TEAMS = 2;       // teams per CTA (the staged cloud fills shared memory: one CTA per SM, so two tiles at a time)
constexpr int KG_THREADS = KG_TEAM * KG_TEAMS;
constexpr int KG_GMAX = 12;       // cells per axis (<= 1728 cells)
constexpr int KG_MIN_N = 1024;

template <int KK>
__device__ __forceinline__ void kg_insert(float (&bd)[KK], int (&bj)[KK], float d, int j) {
    // (distance, index) lexicographic: the exhaustive kernel visits candidates in index order with a strict compare
    if (d < bd[KK - 1] || (d == bd[KK - 1] && j < bj[KK - 1])) {
        bd[KK - 1] = d; bj[KK - 1] = j;
#pragma unroll
        for (int e = KK - 1; e > 0; --e) {
            if (bd[e] < bd[e - 1] || (bd[e] == bd[e - 1] && bj[e] < bj[e - 1])) {
                const float td = bd[e]; bd[e] = bd[e - 1]; bd[e - 1] = td;
                const int tj = bj[e]; bj[e] = bj[e - 1]; bj[e - 1] = tj;
            }
        }
    }
}

__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(KG_TEAM) : "memory"); }

template <int KK>
__global__ void __launch_bounds__(KG_THREADS) knn_grid_kernel(KnnArgs a) {
    extern __shared__ __align__(16) unsigned char kg_raw[];
    float4 *pts = reinterpret_cast<float4 *>(kg_raw);                      // [n] (x, y, z, |p|^2), sorted by cell
    int *pidx = reinterpret_cast<int *>(pts + a.n);                          // [n] original index
    int *cstart = pidx + a.n;                                                // [G^3 + 1]
    int *qh_all = cstart + (KG_GMAX * KG_GMAX * KG_GMAX + 1);                // per team: [G^3 + 1] cell histogram of its queries
    int *qord_all = qh_all + KG_TEAMS * (KG_GMAX * KG_GMAX * KG_GMAX + 1);   // per team: [KG_TEAM] queries sorted by cell
    __shared__ float s_red[8][KG_THREADS / 32];
    __shared__ int s_cnt;
    __shared__ float s_box[8];                                               // lo xyz, cell xyz, max |p|^2, -
    __shared__ int s_g;
    const int cloud = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nv = knn_n(a, cloud);
    const float *pb = a.points + (size_t)cloud * 3 * a.n;
    const bool drop_dups = a.dup != nullptr && a.cloud_dups[cloud] > 0;     // duplicate mode 1 (mode 2 elements: exhaustive path below)
    const uint8_t *dupb = a.dup ? a.dup + (size_t)cloud * a.n : nullptr;

    // ---- 1a. bounding box, count and max |p|^2 of the candidates that take part
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, rmax = 0.f;
    int mine = 0;
    for (int j = tid; j < nv; j += KG_THREADS) {
        if (drop_dups && dupb[j]) continue;
        const float x = __ldg(pb + j), y = __ldg(pb + a.n + j), z = __ldg(pb + 2 * (size_t)a.n + j);
        lo[0] = fminf(lo[0], x); lo[1] = fminf(lo[1], y); lo[2] = fminf(lo[2], z);
        hi[0] = fmaxf(hi[0], x); hi[1] = fmaxf(hi[1], y); hi[2] = fmaxf(hi[2], z);
        rmax = fmaxf(rmax, __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
        ++mine;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
            lo[c3] = fminf(lo[c3], __shfl_xor_sync(0xffffffffu, lo[c3], o));
            hi[c3] = fmaxf(hi[c3], __shfl_xor_sync(0xffffffffu, hi[c3], o));
        }
        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        mine += __shfl_xor_sync(0xffffffffu, mine, o);
    }
    if (tid == 0) s_cnt = 0;
    if (lane == 0) {
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) { s_red[c3][warp] = lo[c3]; s_red[3 + c3][warp] = hi[c3]; }
        s_red[6][warp] = rmax;
    }
    __syncthreads();
    if (lane == 0 && mine) atomicAdd(&s_cnt, mine);
    if (tid == 0) {
        for (int w = 1; w < KG_THREADS / 32; ++w) {
#pragma unroll
            for (int c3 = 0; c3 < 3; ++c3) { s_red[c3][0] = fminf(s_red[c3][0], s_red[c3][w]); s_red[3 + c3][0] = fmaxf(s_red[3 + c3][0], s_red[3 + c3][w]); }
            s_red[6][0] = fmaxf(s_red[6][0], s_red[6][w]);
        }
    }
    __syncthreads();
    const int cnt = s_cnt;
    if (tid == 0) {
        int g = (int)cbrtf((float)cnt / (float)a.grid_target);
        g = g < 1 ? 1 : (g > KG_GMAX ? KG_GMAX : g);
        s_g = g;
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
            const float ext = fmaxf(s_red[3 + c3][0] - s_red[c3][0], 1e-12f);
            s_box[c3] = s_red[c3][0];
            s_box[3 + c3] = ext / (float)g;
        }
        s_box[6] = s_red[6][0];
    }
    __syncthreads();
    const int G = s_g, GC = G * G * G;
    const float blo[3] = {s_box[0], s_box[1], s_box[2]}, cell[3] = {s_box[3], s_box[4], s_box[5]};
    const float inv[3] = {1.0f / cell[0], 1.0f / cell[1], 1.0f / cell[2]};
    const float rp_max = s_box[6];
    const float ext_max = fmaxf(fmaxf(cell[0], cell[1]), cell[2]) * (float)G;
    auto cell_of = [&](float v, int c3) { int c = (int)floorf((v - blo[c3]) * inv[c3]); return c < 0 ? 0 : (c >= G ? G - 1 : c); };

    // ---- 1b. histogram, exclusive scan, scatter
    for (int c = tid; c <= GC; c += KG_THREADS) cstart[c] = 0;
    __syncthreads();
    for (int j = tid; j < nv; j += KG_THREADS) {
        if (drop_dups && dupb[j]) continue;
        const float x = __ldg(pb + j), y = __ldg(pb + a.n + j), z = __ldg(pb + 2 * (size_t)a.n + j);
        atomicAdd(&cstart[(cell_of(z, 2) * G + cell_of(y, 1)) * G + cell_of(x, 0) + 1], 1);
    }
    __syncthreads();
    if (warp == 0) {                                   // inclusive scan of cstart[1..GC] by one warp, 32 cells at a time
        int carry = 0;
        for (int c0 = 1; c0 <= GC; c0 += 32) {
            const int c = c0 + lane;
            int v = c <= GC ? cstart[c] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += up;
            }
            if (c <= GC) cstart[c] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    // scatter: cstart[c] = first slot of cell c; a per-cell fill counter lives in pidx's tail?  No spare room in general: use a
    // second pass with atomics on a copy of the starts kept in the (not yet written) pidx array's upper... -> simplest: cursor array
    // in registers is impossible, so cstart[] doubles as the cursor and is restored afterwards by a shift.
    for (int j = tid; j < nv; j += KG_THREADS) {
        if (drop_dups && dupb[j]) continue;
        const float x = __ldg(pb + j), y = __ldg(pb + a.n + j), z = __ldg(pb + 2 * (size_t)a.n + j);
        const int c = (cell_of(z, 2) * G + cell_of(y, 1)) * G + cell_of(x, 0);
        const int slot = atomicAdd(&cstart[c], 1);     // cursor of cell c runs from its start to the start of cell c+1
        pts[slot] = make_float4(x, y, z, __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
        pidx[slot] = j;
    }
    __syncthreads();
    // after the scatter cstart[c] holds the END of cell c = the start of cell c+1: shift by one cell to restore the starts
    {
        int keep[(KG_GMAX * KG_GMAX * KG_GMAX + KG_THREADS) / KG_THREADS];
        int e = 0;
        for (int c = tid; c < GC; c += KG_THREADS) keep[e++] = cstart[c];
        __syncthreads();
        e = 0;
        for (int c = tid; c < GC; c += KG_THREADS) cstart[c + 1] = keep[e++];
        if (tid == 0) cstart[0] = 0;
    }
    __syncthreads();

    // ---- 2. queries: batch elements that read this cloud, their query blocks dealt round-robin to the gridDim.x CTAs of the cloud
    const int team = tid / KG_TEAM, tt = tid - team * KG_TEAM;
    int *qh = qh_all + team * (KG_GMAX * KG_GMAX * KG_GMAX + 1), *qord = qord_all + team * KG_TEAM;
    int item = 0;
    const int bi_lo = a.owner ? 0 : cloud * a.p_div, bi_hi = a.owner ? a.b : (cloud + 1) * a.p_div;
    for (int bi0 = bi_lo; bi0 < bi_hi; bi0 += 32) {
      // which of the next 32 batch elements read this cloud (one owner load per lane instead of a dependent load per element)
      unsigned todo = __ballot_sync(0xffffffffu, bi0 + lane < bi_hi && (a.owner == nullptr || __ldg(a.owner + bi0 + lane) == cloud));
      while (todo) {                                                         // block-uniform: every warp sees the same mask
        const int bi = bi0 + __ffs(todo) - 1;
        todo &= todo - 1u;
        const int mv = knn_m(a, bi);
        const int grp = knn_group(a, bi);
        const bool mode2 = a.dup != nullptr && a.group_any[grp] != 0;
        const float maxd = mode2 ? ordered_to_float(a.maxd[grp]) : 0.f;
        const float *qb = a.query + (size_t)bi * 3 * a.m;
        for (int q0 = 0; q0 < mv; q0 += KG_TEAM, ++item) {
            if (item % ((int)gridDim.x * KG_TEAMS) != (int)blockIdx.x * KG_TEAMS + team) continue;   // team-uniform
            // threads of a warp should look at the same cells (candidate loads become broadcasts instead of 32 different
            // addresses): the team's queries are sorted by cell first (counting sort, ~100 instructions per thread)
            int qs = tt;                                                   // the query (offset in this block) this thread answers
            const int nq = min(KG_TEAM, mv - q0);
            if (!mode2) {
                for (int c = tt; c <= GC; c += KG_TEAM) qh[c] = 0;
                team_sync(team);
                int mycell = 0;
                if (tt < nq) {
                    const float x = __ldg(qb + q0 + tt), y = __ldg(qb + a.m + q0 + tt), z = __ldg(qb + 2 * (size_t)a.m + q0 + tt);
                    mycell = (cell_of(z, 2) * G + cell_of(y, 1)) * G + cell_of(x, 0);
                    atomicAdd(&qh[mycell + 1], 1);
                }
                team_sync(team);
                if (tt < 32) {
                    int carry = 0;
                    for (int c0 = 1; c0 <= GC; c0 += 32) {
                        const int c = c0 + tt;
                        int v = c <= GC ? qh[c] : 0;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const int up = __shfl_up_sync(0xffffffffu, v, d);
                            if (tt >= d) v += up;
                        }
                        if (c <= GC) qh[c] = v + carry;
                        carry += __shfl_sync(0xffffffffu, v, 31);
                    }
                }
                team_sync(team);
                if (tt < nq) qord[atomicAdd(&qh[mycell], 1)] = tt;
                team_sync(team);
                if (tt < nq) qs = qord[tt];
            }
            const int qi = q0 + qs;
            if (tt >= nq) continue;
            const float qx = __ldg(qb + qi), qy = __ldg(qb + a.m + qi), qz = __ldg(qb + 2 * (size_t)a.m + qi);
            const float rq = __fmaf_rn(qz, qz, __fmaf_rn(qy, qy, __fmul_rn(qx, qx)));
            float bd[KK];
            int bj[KK];
#pragma unroll
            for (int e = 0; e < KK; ++e) { bd[e] = __int_as_float(0x7f800000); bj[e] = 0; }
            if (mode2) {
                // degenerate cloud: all candidates in index order from global memory, duplicates penalised by max(D)
                for (int j = 0; j < nv; ++j) {
                    const float x = __ldg(pb + j), y = __ldg(pb + a.n + j), z = __ldg(pb + 2 * (size_t)a.n + j);
                    const float rp = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
                    const float dot = __fmaf_rn(qz, z, __fmaf_rn(qy, y, __fmul_rn(qx, x)));
                    float d = expanded_dist(rq, dot, rp);
                    if (dupb[j]) d = __fadd_rn(d, maxd);
                    kg_insert<KK>(bd, bj, d, j);
                }
            } else {
                const int cx = cell_of(qx, 0), cy = cell_of(qy, 1), cz = cell_of(qz, 2);
                const float slack = 8e-6f * (rq + rp_max) + 1e-9f;        // fp32 rounding of two expanded-form distances
                bool done = false;
                // the 3x3x3 block around the query's cell, then -- if the verification fails -- the shell up to 5x5x5
                int px0 = 0, px1 = -1, py0 = 0, py1 = -1, pz0 = 0, pz1 = -1;   // block scanned so far (empty)
#pragma unroll 1
                for (int ring = 1; ring <= 2 && !done; ++ring) {
                    const int x0 = max(cx - ring, 0), x1 = min(cx + ring, G - 1), y0 = max(cy - ring, 0), y1 = min(cy + ring, G - 1);
                    const int z0 = max(cz - ring, 0), z1 = min(cz + ring, G - 1);
                    for (int zz = z0; zz <= z1; ++zz)
                        for (int yy = y0; yy <= y1; ++yy) {
                            // the cells x0..x1 of a row are contiguous in the sorted array; rows inside the previous block only
                            // contribute their two ends
                            const bool inner = zz >= pz0 && zz <= pz1 && yy >= py0 && yy <= py1;
                            const int rowc = (zz * G + yy) * G;
                            int s0 = cstart[rowc + x0], s1 = cstart[rowc + x1 + 1];
                            int h0 = s1, h1 = s1;                           // hole [h0, h1) = what the previous block already covered
                            if (inner) { h0 = cstart[rowc + px0]; h1 = cstart[rowc + px1 + 1]; }
                            for (int sl = s0; sl < s1; ++sl) {
                                if (sl == h0) { sl = h1; if (sl >= s1) break; }
                                const float4 p = pts[sl];
                                const float dot = __fmaf_rn(qz, p.z, __fmaf_rn(qy, p.y, __fmul_rn(qx, p.x)));
                                kg_insert<KK>(bd, bj, expanded_dist(rq, dot, p.w), pidx[sl]);
                            }
                        }
                    px0 = x0; px1 = x1; py0 = y0; py1 = y1; pz0 = z0; pz1 = z1;
                    // verify: distance from the query to the nearest face of the scanned block that has cells behind it
                    float lim = INFINITY;
                    if (x0 > 0) lim = fminf(lim, qx - (blo[0] + (float)x0 * cell[0]));
                    if (x1 < G - 1) lim = fminf(lim, (blo[0] + (float)(x1 + 1) * cell[0]) - qx);
                    if (y0 > 0) lim = fminf(lim, qy - (blo[1] + (float)y0 * cell[1]));
                    if (y1 < G - 1) lim = fminf(lim, (blo[1] + (float)(y1 + 1) * cell[1]) - qy);
                    if (z0 > 0) lim = fminf(lim, qz - (blo[2] + (float)z0 * cell[2]));
                    if (z1 < G - 1) lim = fminf(lim, (blo[2] + (float)(z1 + 1) * cell[2]) - qz);
                    // slack: a candidate may sit in the neighbouring cell of where exact arithmetic would put it (1e-4 of the box)
                    const float lim_eff = lim * (1.f - 1e-3f) - 1e-4f * ext_max - 1e-6f;
                    float dk = bd[0];                             // the k-th best decides (k <= KK)
#pragma unroll
                    for (int e = 1; e < KK; ++e) if (e < a.k) dk = bd[e];
                    done = dk < __int_as_float(0x7f800000) && (lim == INFINITY || (lim_eff > 0.f && dk + slack <= lim_eff * lim_eff));
                }
                if (!done) {                                      // exhaustive over the staged candidates (any order: (d, j) insert)
#pragma unroll
                    for (int e = 0; e < KK; ++e) { bd[e] = __int_as_float(0x7f800000); bj[e] = 0; }
                    for (int sl = 0; sl < cnt; ++sl) {
                        const float4 p = pts[sl];
                        const float dot = __fmaf_rn(qz, p.z, __fmaf_rn(qy, p.y, __fmul_rn(qx, p.x)));
                        kg_insert<KK>(bd, bj, expanded_dist(rq, dot, p.w), pidx[sl]);
                    }
                }
            }
            const size_t row = ((size_t)bi * a.m + qi) * a.k;
#pragma unroll
            for (int e = 0; e < KK; ++e) {
                if (e < a.k) {
                    if (a.idx64) a.idx64[row + e] = bj[e];
                    if (a.idx32) a.idx32[row + e] = bj[e];
                    if (a.dist) a.dist[row + e] = bd[e];
                    if (a.knn) {
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch)
                            a.knn[(((size_t)bi * 3 + ch) * a.m + qi) * a.k + e] = __ldg(pb + (size_t)ch * a.n + bj[e]);
                    }
                }
            }
        }
      }
    }
}

// --------------------------------------------------------------------------------------------
// k > 64: CTA per query
// --------------------------------------------------------------------------------------------
constexpr int KL_THREADS = 256;

__global__ void __launch_bounds__(KL_THREADS) knn_large_kernel(KnnArgs a, int k2, uint32_t *__restrict__ gkeys) {
    extern __shared__ __align__(16) unsigned char raw[];
    // layout: sel[k2] u64 | hist[256] | misc[8] | keys[n] (unless gkeys)
    unsigned long long *sel = reinterpret_cast<unsigned long long *>(raw);
    uint32_t *hist = reinterpret_cast<uint32_t *>(sel + k2);
    uint32_t *misc = hist + 256;
    const int qi = blockIdx.x, bi = blockIdx.y;
    if (qi >= knn_m(a, bi)) return;          // ragged: this batch element has fewer queries (block-uniform)
    uint32_t *keys = gkeys ? gkeys + ((size_t)bi * a.m + qi) * a.n : misc + 8;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = a.c;
    const int cloud = knn_cloud(a, bi);
    const int grp = knn_group(a, bi);
    const int nv = knn_n(a, cloud);
    const int kv = min(a.k, nv);             // a ragged cloud may hold fewer than k points: the tail stays unwritten
    const float *qb = a.query + (size_t)bi * C * a.m;
    const float *pb = a.points + (size_t)cloud * C * a.n;
    const int dmode = knn_dup_mode(a, cloud, grp);
    const bool penal = dmode != 0;
    const float maxd = dmode == 2 ? ordered_to_float(a.maxd[grp]) : 0.f;
    const uint8_t *dupb = penal ? a.dup + (size_t)cloud * a.n : nullptr;
    if (kv <= 0) return;

    // ---- 1. keys --------------------------------------------------------------------------
    float rq = 0.f;
    for (int ch = 0; ch < C; ++ch) { const float v = __ldg(qb + (size_t)ch * a.m + qi); rq = __fmaf_rn(v, v, rq); }
    for (int j = tid; j < nv; j += KL_THREADS) {
        float dot = 0.f, rp = 0.f;
        for (int ch = 0; ch < C; ++ch) {
            const float pv = __ldg(pb + (size_t)ch * a.n + j);
            dot = __fmaf_rn(__ldg(qb + (size_t)ch * a.m + qi), pv, dot);
            rp = __fmaf_rn(pv, pv, rp);
        }
        float d = expanded_dist(rq, dot, rp);
        const bool isdup = penal && dupb[j];
        if (isdup) d = __fadd_rn(d, maxd);
        keys[j] = (isdup && dmode == 1) ? 0xffffffffu : float_to_ordered(d);
    }
    // ---- 2. radix select: key of rank k-1, MSB first, 8 bits per pass ------------------------
    uint32_t prefix = 0, pmask = 0;
    uint32_t want = kv - 1;  // rank wanted among keys matching the prefix
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[tid] = 0;
        __syncthreads();
        for (int j = tid; j < nv; j += KL_THREADS) {
            const uint32_t key = keys[j];
            if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (warp == 0) {  // 256 buckets: 8 per lane, exclusive scan
            uint32_t loc[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { loc[i] = hist[lane * 8 + i]; sum += loc[i]; }
            uint32_t inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
            uint32_t run = inc - sum;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (want >= run && want < run + loc[i]) { misc[0] = lane * 8 + i; misc[1] = want - run; }
                run += loc[i];
            }
        }
        __syncthreads();
        prefix |= misc[0] << shift;
        pmask |= 255u << shift;
        want = misc[1];
        __syncthreads();
    }
    const uint32_t kth = prefix;         // key value at rank k-1
    const uint32_t ties_wanted = want + 1;  // how many keys == kth belong to the top k (lowest indices)

    // ---- 3. ordered compaction ------------------------------------------------------------
    if (tid == 0) { misc[2] = 0; misc[3] = 0; }  // [2] slots used, [3] ties taken so far
    for (int i = tid; i < k2; i += KL_THREADS) sel[i] = ~0ull;
    __syncthreads();
    for (int base = 0; base < nv; base += KL_THREADS) {
        const int j = base + tid;
        const uint32_t key = j < nv ? keys[j] : 0xffffffffu;
        const bool less = j < nv && key < kth;
        const bool tie = j < nv && key == kth;
        // ties must be taken in index order: block-wide exclusive count of earlier ties
        const unsigned tb = __ballot_sync(0xffffffffu, tie);
        if (lane == 0) hist[warp] = __popc(tb);
        __syncthreads();
        uint32_t before = misc[3];
        for (int w = 0; w < warp; ++w) before += hist[w];
        before += __popc(tb & ((1u << lane) - 1u));
        uint32_t total = 0;
        for (int w = 0; w < KL_THREADS / 32; ++w) total += hist[w];
        if (less || (tie && before < ties_wanted)) {
            const uint32_t slot = atomicAdd(&misc[2], 1u);
            sel[slot] = ((unsigned long long)key << 32) | (uint32_t)j;
        }
        __syncthreads();
        if (tid == 0) misc[3] += total;
        __syncthreads();
    }
    // ---- 4. bitonic sort of the packed (key,index) pairs, ascending ----------------------------
    for (int size = 2; size <= k2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (k2 >> 1); t += KL_THREADS) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const unsigned long long x = sel[lo], y = sel[hi];
                if ((x > y) == up) { sel[lo] = y; sel[hi] = x; }
            }
            __syncthreads();
        }
    }
    // ---- 5. outputs ---------------------------------------------------------------------------
    const size_t row = ((size_t)bi * a.m + qi) * a.k;
    for (int p = tid; p < kv; p += KL_THREADS) {
        const unsigned long long v = sel[p];
        const int32_t j = (int32_t)(uint32_t)v;
        if (a.idx64) a.idx64[row + p] = j;
        if (a.idx32) a.idx32[row + p] = j;
        if (a.dist) a.dist[row + p] = ordered_to_float((uint32_t)(v >> 32));
    }
    if (a.knn) {
        for (int ch = 0; ch < C; ++ch) {
            float *o = a.knn + (((size_t)bi * C + ch) * a.m + qi) * a.k;
            for (int p = tid; p < kv; p += KL_THREADS) o[p] = __ldg(pb + (size_t)ch * a.n + (uint32_t)sel[p]);
        }
    }
}

// grad_points[b/p_div, c, idx[b,m,kk]] += grad_knn[b,c,m,kk]
__global__ void __launch_bounds__(256) group_gather_bwd_kernel(int c, int m, int n, int k, int p_div, long long total,
                                                              const float *__restrict__ grad_knn,
                                                              const int64_t *__restrict__ idx,
                                                              float *__restrict__ grad_points) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const long long mk = (long long)m * k;
        const long long bc = t / mk;
        const long long r = t - bc * mk;  // m*k offset
        const long long bi = bc / c;
        const int ch = (int)(bc - bi * c);
        const int j = (int)idx[bi * mk + r];
        atomicAdd(grad_points + ((bi / p_div) * c + ch) * (long long)n + j, grad_knn[t]);
    }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static int ceil_pow2_i(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct KnnPlan {
    bool large;
    int k2;
    bool keys_global;
    size_t smem;
    int tile_n;
    size_t off_dup, off_any, off_cany, off_maxd, off_keys, total;
    int groups;
};

static bool make_plan(int b, int c, int m, int n, int k, int clouds, int groups, int unique, KnnPlan &pl) {
    const int smem_cap = device_info().smem_optin;
    pl.large = k > 64;
    pl.k2 = 0; pl.keys_global = false; pl.tile_n = 0; pl.smem = 0;
    if (pl.large) {
        pl.k2 = ceil_pow2_i(k);
        const size_t fixed = (size_t)pl.k2 * 8 + (256 + 8) * 4;
        if (fixed > (size_t)smem_cap) return false;
        pl.keys_global = fixed + (size_t)n * 4 > (size_t)smem_cap;
        pl.smem = pl.keys_global ? fixed : fixed + (size_t)n * 4;
    } else {
        // ~48 KB candidate tile: several CTAs per SM; at least one warp-width of candidates
        int tn = (int)((48 * 1024) / ((size_t)(c + 1) * 4));
        tn = (tn / 32) * 32;
        if (tn < 32) tn = 32;
        const int n_up = ((n + 31) / 32) * 32;
        if (tn > n_up) tn = n_up;
        pl.tile_n = tn;
        pl.smem = ((size_t)(c + 1) * tn + (size_t)KS_WARPS * c) * 4;
        if (pl.smem > (size_t)smem_cap) return false;
    }
    pl.groups = groups;
    size_t off = 0;
    pl.off_any = off;  off += align256(unique ? (size_t)pl.groups * 4 : 0);
    pl.off_cany = off; off += align256(unique ? (size_t)clouds * 4 : 0);
    pl.off_maxd = off; off += align256(unique ? (size_t)pl.groups * 4 : 0);
    pl.off_dup = off;  off += align256(unique ? (size_t)clouds * n : 0);
    pl.off_keys = off; off += align256(pl.keys_global ? (size_t)b * m * n * 4 : 0);
    pl.total = off;
    return true;
}

}  // namespace pu3

using namespace pu3;

// Test hook: force the streaming-insertion kernel for k <= 64 even when the cloud fits the tiled kernel.
static int g_knn_force_stream = 0;
extern "C" void pu3_knn_force_stream(int on) { g_knn_force_stream = on; }
// Test hook: 1 = the indices-only feature kNN orders ranks 1..k-1 exactly (default: rank 0 exact, the rest as a set).
static int g_knn_exact_pops = 0;
extern "C" void pu3_knn_exact_pops(int on) { g_knn_exact_pops = on; }
// Test hook: 1 = duplicate detection by the O(n^2) scans (hash-first in shared memory up to 1024 points) instead of the hash table.
// Test hook: 1 = knn_thread_kernel never pre-filters its candidates (A/B of the exact bounding-sphere filter).
static int g_knn_no_prefilter = 0;
extern "C" void pu3_knn_no_prefilter(int on) { g_knn_no_prefilter = on; }
// Test / A-B hook: 0 = the xyz searches (c = 3, k <= 8) always run the exhaustive knn_thread_kernel, 1 (default) = clouds of
// >= 1024 points go through the uniform-grid kernel (bit-identical results).
static int g_knn_grid = 1;
extern "C" void pu3_knn_set_grid(int on) { g_knn_grid = on; }
static int knn_grid_target() {
    static int t = 0;
    if (t == 0) { const char *e = getenv("PU3_KNN_GRID_TARGET"); t = e ? atoi(e) : 3; if (t < 1) t = 1; }
    return t;
}
// Test hook: 1 = the feature-space kernel does not look for duplicates itself (the three pre-pass kernels run, as in round 1).
static int g_knn_no_fused_dup = 0;
extern "C" void pu3_knn_no_fused_dup(int on) { g_knn_no_fused_dup = on; }
static int g_knn_dup_scan = 0;
extern "C" void pu3_knn_dup_scan(int on) { g_knn_dup_scan = on; }

extern "C" size_t pu3_group_knn_workspace(int b, int c, int m, int n, int k, int p_div, int unique) {
    if (b <= 0 || c <= 0 || m <= 0 || n <= 0 || k <= 0 || p_div <= 0) return 0;
    KnnPlan pl;
    // upper bound for every grouping: one cloud and one group per batch element
    (void)p_div;
    if (!make_plan(b, c, m, n, k, b, b, unique, pl)) return 0;
    return pl.total;
}

static int group_knn_impl(int b, int c, int m, int n, int k, int p_div, int clouds, int groups, const int32_t *owner,
                          const int32_t *group_of, const int32_t *n_arr, const int32_t *m_arr, const float *query,
                          const float *points, int unique, int max_group, float *knn, int64_t *idx64,
                          int32_t *idx32, float *dist, void *workspace, size_t workspace_bytes, pu3_stream_t stream);

extern "C" int pu3_group_knn_f32(int b, int c, int m, int n, int k, int p_div, const float *query,
                                 const float *points, int unique, int max_group, float *knn, int64_t *idx64,
                                 int32_t *idx32, float *dist, void *workspace, size_t workspace_bytes,
                                 pu3_stream_t stream) {
    if (b > 0 && p_div >= 1 && b % p_div == 0) {
        if (max_group <= 0 || max_group > b) max_group = b;
        return group_knn_impl(b, c, m, n, k, p_div, b / p_div, (b + max_group - 1) / max_group, nullptr, nullptr, nullptr,
                              nullptr, query, points, unique, max_group, knn, idx64, idx32, dist, workspace,
                              workspace_bytes, stream);
    }
    return group_knn_impl(b, c, m, n, k, p_div, 0, 0, nullptr, nullptr, nullptr, nullptr, query, points, unique, max_group,
                          knn, idx64, idx32, dist, workspace, workspace_bytes, stream);
}

extern "C" int pu3_group_knn_ragged_f32(int b, int c, int m, int n, int k, int clouds, int groups, const int32_t *owner,
                                        const int32_t *group_of, const int32_t *n_arr, const int32_t *m_arr,
                                        const float *query, const float *points, int unique, float *knn,
                                        int64_t *idx64, int32_t *idx32, float *dist, void *workspace,
                                        size_t workspace_bytes, pu3_stream_t stream) {
    PU3_ARG_CHECK(owner && group_of, "group_knn_ragged: owner and group_of are required");
    PU3_ARG_CHECK(clouds > 0 && groups > 0 && clouds <= 65535, "group_knn_ragged: clouds=%d groups=%d", clouds, groups);
    return group_knn_impl(b, c, m, n, k, 1, clouds, groups, owner, group_of, n_arr, m_arr, query, points, unique, 1, knn,
                          idx64, idx32, dist, workspace, workspace_bytes, stream);
}

static int group_knn_impl(int b, int c, int m, int n, int k, int p_div, int clouds, int groups, const int32_t *owner,
                          const int32_t *group_of, const int32_t *n_arr, const int32_t *m_arr, const float *query,
                          const float *points, int unique, int max_group, float *knn, int64_t *idx64,
                          int32_t *idx32, float *dist, void *workspace, size_t workspace_bytes, pu3_stream_t stream) {
    const bool unordered = (unique & PU3_KNN_SET_ORDER) != 0;   // ranks 1..k-1 as a set (indices-only calls of the tiled kernel)
    unique &= 1;
    PU3_ARG_CHECK(b >= 0 && c > 0 && m >= 0 && n >= 0 && k >= 0, "group_knn: bad size b=%d c=%d m=%d n=%d k=%d", b, c, m, n, k);
    PU3_ARG_CHECK(k <= n, "group_knn: points size must be greater or equal to k (n=%d, k=%d)", n, k);  // operations.py:186
    if (b == 0 || m == 0 || k == 0) return PU3_OK;
    PU3_ARG_CHECK(p_div >= 1 && b % p_div == 0, "group_knn: p_div=%d must divide b=%d", p_div, b);
    PU3_ARG_CHECK(b <= 65535, "group_knn: b=%d exceeds 65535", b);
    PU3_ARG_CHECK(query && points, "group_knn: null input pointer");
    KnnPlan pl;
    if (!make_plan(b, c, m, n, k, clouds, groups, unique, pl)) {
        set_error("group_knn: c=%d k=%d does not fit shared memory", c, k);
        return PU3_E_UNSUPPORTED;
    }
    if (pl.total > 0) {
        if (!workspace || workspace_bytes < pl.total) {
            set_error("group_knn: workspace %zu bytes, need %zu", workspace_bytes, pl.total);
            return PU3_E_WORKSPACE;
        }
        PU3_ARG_CHECK(((uintptr_t)workspace & 255) == 0, "group_knn: workspace must be 256-byte aligned");
    }
    cudaStream_t s = as_stream(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    KnnArgs a{b, c, m, n, k, p_div, max_group, owner, group_of, n_arr, m_arr,
              query, points, nullptr, nullptr, nullptr, nullptr, knn, idx64, idx32, dist,
              (unordered && g_knn_exact_pops == 0) ? 0 : 1,
              (query != points && g_knn_no_prefilter == 0) ? 1 : 0, 0, knn_grid_target()};
    // the tiled feature-space kernel (n <= 320, k <= 64) finds duplicates itself: no pre-pass
    const bool feat_path = !pl.large && !(c == 3 && k <= KT_KMAX && g_knn_force_stream == 0) && n <= KF_NMAX && g_knn_force_stream == 0 &&
                           ((size_t)(c + 1) * KF_NMAX) * 4 + (size_t)KF_QB * KF_NMAX * 4 <= (size_t)device_info().smem_optin;
    if (unique && feat_path && g_knn_no_fused_dup == 0) {
        a.fused_dup = 1;
    } else if (unique) {
        int *group_any = reinterpret_cast<int *>(ws + pl.off_any);
        int *cloud_any = reinterpret_cast<int *>(ws + pl.off_cany);
        uint32_t *maxd = reinterpret_cast<uint32_t *>(ws + pl.off_maxd);
        uint8_t *dup = ws + pl.off_dup;
        int st = cuda_status(cudaMemsetAsync(ws + pl.off_any, 0, pl.off_dup - pl.off_any, s), "group_knn: memset");
        if (st) return st;
        if (n <= KH_MAXN && g_knn_dup_scan == 0) {
            int table = 512;
            while (table < 2 * n) table *= 2;            // load factor <= 0.5
            st = cuda_status(cudaFuncSetAttribute(knn_dup_hash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, table * 4), "group_knn: smem attr");
            if (st) return st;
            knn_dup_hash_kernel<<<clouds, n <= 512 ? 256 : 1024, table * 4, s>>>(c, n, table, n_arr, points, dup, cloud_any);
        } else if (n <= KD_MAXN) knn_dup_small_kernel<<<clouds, 256, 0, s>>>(c, n, n_arr, points, dup, cloud_any);
        else knn_dup_kernel<<<dim3((n + 255) / 256, clouds), 256, 0, s>>>(c, n, n_arr, points, dup, cloud_any);
        PU3_LAUNCH_CHECK("knn_dup_kernel");
        knn_groupflag_kernel<<<(b + 255) / 256, 256, 0, s>>>(a, cloud_any, group_any);
        PU3_LAUNCH_CHECK("knn_groupflag_kernel");
        a.dup = dup; a.cloud_dups = cloud_any; a.group_any = group_any; a.maxd = maxd;
        knn_maxd_kernel<<<dim3((m + 127) / 128, b), 128, 0, s>>>(a, maxd);
        PU3_LAUNCH_CHECK("knn_maxd_kernel");
    }
    if (pl.large) {
        PU3_ARG_CHECK(m <= 2147483647 / 1, "group_knn: m too large");
        uint32_t *gkeys = pl.keys_global ? reinterpret_cast<uint32_t *>(ws + pl.off_keys) : nullptr;
        int st = cuda_status(cudaFuncSetAttribute(knn_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem),
                             "group_knn: smem attr");
        if (st) return st;
        knn_large_kernel<<<dim3(m, b), KL_THREADS, pl.smem, s>>>(a, pl.k2, gkeys);
        PU3_LAUNCH_CHECK("knn_large_kernel");
        return PU3_OK;
    }
    if (c == 3 && k <= KT_KMAX && g_knn_force_stream == 0 && g_knn_grid != 0 && n >= KG_MIN_N) {
        // uniform-grid search: the cloud's candidates sorted by cell in shared memory (20 bytes per point + the cell starts)
        const size_t smem = (size_t)n * 20 + (size_t)(KG_GMAX * KG_GMAX * KG_GMAX + 1) * 4 * (1 + KG_TEAMS) + (size_t)KG_TEAMS * KG_TEAM * 4 + 16;
        if (smem <= (size_t)device_info().smem_optin) {
            // CTAs per cloud: enough to cover the chip, never more than there are query blocks
            const long long blocks_per_cloud = ((long long)((m + KG_TEAM - 1) / KG_TEAM) * (b / clouds > 0 ? b / clouds : 1) + KG_TEAMS - 1) / KG_TEAMS;
            long long parts = (device_info().sm_count + clouds - 1) / clouds;
            if (parts > blocks_per_cloud) parts = blocks_per_cloud;
            if (parts < 1) parts = 1;
#define PU3_KG_LAUNCH(KKV)                                                                                              \
    do {                                                                                                                \
        auto kern = knn_grid_kernel<KKV>;                                                                                \
        int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),        \
                             "group_knn: smem attr");                                                                   \
        if (st) return st;                                                                                              \
        kern<<<dim3((unsigned)parts, clouds), KG_THREADS, smem, s>>>(a);                                                \
    } while (0)
            if (k <= 2) PU3_KG_LAUNCH(2);
            else if (k <= 5) PU3_KG_LAUNCH(5);
            else PU3_KG_LAUNCH(8);
#undef PU3_KG_LAUNCH
            PU3_LAUNCH_CHECK("knn_grid_kernel");
            return PU3_OK;
        }
    }
    if (c == 3 && k <= KT_KMAX && g_knn_force_stream == 0) {
        dim3 grid((m + KT_THREADS * KT_QPT - 1) / (KT_THREADS * KT_QPT), b);
        if (k <= 2) knn_thread_kernel<2><<<grid, KT_THREADS, 0, s>>>(a);
        else if (k <= 5) knn_thread_kernel<5><<<grid, KT_THREADS, 0, s>>>(a);
        else knn_thread_kernel<8><<<grid, KT_THREADS, 0, s>>>(a);
        PU3_LAUNCH_CHECK("knn_thread_kernel");
        return PU3_OK;
    }
    if (n <= KF_NMAX && g_knn_force_stream == 0) {
        const size_t smem = ((size_t)(c + 1) * KF_NMAX) * 4 + (size_t)KF_QB * KF_NMAX * 4;
        if (smem <= (size_t)device_info().smem_optin) {
            const bool hot = idx32 && !idx64 && !dist && !knn;
            // 3 resident CTAs per SM: split the query blocks of a cloud over several CTAs until the chip is covered
            const int qblocks = (m + KF_QB - 1) / KF_QB;
            int qsplit = (int)((3LL * device_info().sm_count + b - 1) / b);
            if (qsplit > qblocks) qsplit = qblocks;
            if (qsplit < 1) qsplit = 1;
#define PU3_KF_LAUNCH(CTV, HOTV)                                                                                        \
    do {                                                                                                                \
        auto kern = knn_feat_kernel<CTV, HOTV>;                                                                          \
        int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),        \
                             "group_knn: smem attr");                                                                   \
        if (st) return st;                                                                                              \
        kern<<<dim3(b, qsplit), KF_THREADS, smem, s>>>(a);                                                              \
    } while (0)
            if (c == 24 && hot) PU3_KF_LAUNCH(24, true);
            else if (c == 24) PU3_KF_LAUNCH(24, false);
            else if (hot) PU3_KF_LAUNCH(0, true);
            else PU3_KF_LAUNCH(0, false);
#undef PU3_KF_LAUNCH
            PU3_LAUNCH_CHECK("knn_feat_kernel");
            return PU3_OK;
        }
    }
    // queries per CTA: enough CTAs to cover the chip twice, but never fewer than one pass of 8 warps
    const int sms = device_info().sm_count;
    int q_per_cta = m;
    const long long want_ctas = 2LL * sms;
    if ((long long)b < want_ctas) {
        const int split = (int)((want_ctas + b - 1) / b);
        q_per_cta = (m + split - 1) / split;
        q_per_cta = ((q_per_cta + KS_WARPS - 1) / KS_WARPS) * KS_WARPS;
    }
    if (q_per_cta < KS_WARPS) q_per_cta = KS_WARPS;
    dim3 grid((m + q_per_cta - 1) / q_per_cta, b);
    const int E = k <= 32 ? 1 : 2;
#define PU3_KNN_LAUNCH(CT, EE)                                                                               \
    do {                                                                                                      \
        auto kern = knn_small_kernel<CT, EE>;                                                                 \
        int st = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem), \
                             "group_knn: smem attr");                                                       \
        if (st) return st;                                                                                    \
        kern<<<grid, KS_THREADS, pl.smem, s>>>(a, pl.tile_n, q_per_cta);                                      \
    } while (0)
    if (c == 3 && E == 1) PU3_KNN_LAUNCH(3, 1);
    else if (c == 3) PU3_KNN_LAUNCH(3, 2);
    else if (c == 24 && E == 1) PU3_KNN_LAUNCH(24, 1);
    else if (c == 24) PU3_KNN_LAUNCH(24, 2);
    else if (E == 1) PU3_KNN_LAUNCH(0, 1);
    else PU3_KNN_LAUNCH(0, 2);
#undef PU3_KNN_LAUNCH
    PU3_LAUNCH_CHECK("knn_small_kernel");
    return PU3_OK;
}

extern "C" int pu3_group_gather_bwd_f32(int b, int c, int m, int n, int k, int p_div, const float *grad_knn,
                                        const int64_t *idx64, float *grad_points, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && c >= 0 && m >= 0 && n >= 0 && k >= 0 && p_div >= 1, "group_gather_bwd: bad size");
    const long long total = (long long)b * c * m * k;
    if (total == 0) return PU3_OK;
    PU3_ARG_CHECK(grad_knn && idx64 && grad_points && n > 0, "group_gather_bwd: null pointer");
    const long long blocks = (total + 255) / 256;
    const long long cap = (long long)device_info().sm_count * 8;
    group_gather_bwd_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(c, m, n, k, p_div, total,
                                                                                              grad_knn, idx64, grad_points);
    PU3_LAUNCH_CHECK("group_gather_bwd_kernel");
    return PU3_OK;
}
