// Chamfer / nearest-neighbour distance (forward + backward), sm_100a.
//
// Replaces losses/nmdistance_cuda.cu of the reference:
//   NmDistanceKernel      :11-133  -> nmdist_fwd_kernel (both directions in ONE launch)
//   NmDistanceGradKernel  :154-173 -> nmdist_bwd_kernel (both directions in ONE launch)
// The reference launches a fixed (32,16)x512 grid per direction (at the training shape
// b=32, n=m=624 only 64 of its 512 CTAs have work) and runs the backward serially over
// the batch (grid (1,16)).  Here the grid covers (query tile, batch, direction).
//
// Bound: all-pairs FP32 ALU work, not HBM (arithmetic intensity ~ 8*m/20 flop/B); the
// candidates are staged once per CTA in shared memory as float4 so that the inner loop is
// one broadcast LDS.128 per candidate shared by QPT queries.
#include "pu3_common.cuh"

namespace pu3 {

constexpr int NMD_THREADS = 128;
constexpr int NMD_QPT = 2;      // queries per thread: each LDS.128 feeds 2 distance evaluations
constexpr int NMD_TILE = 2048;  // candidates per shared-memory tile (32 KB as float4)

struct NmdDir {
    const float *q;  // (b,nq,3) queries
    const float *p;  // (b,np,3) candidates
    float *dist;     // (b,nq)
    int32_t *idx;    // (b,nq)
    int nq, np;
};

__global__ void __launch_bounds__(NMD_THREADS) nmdist_fwd_kernel(NmdDir d0, NmdDir d1) {
    const NmdDir D = blockIdx.z == 0 ? d0 : d1;
    const int q0 = blockIdx.x * (NMD_THREADS * NMD_QPT);
    if (q0 >= D.nq) return;  // block-uniform
    const int b = blockIdx.y;
    __shared__ float4 tile[NMD_TILE];

    const float *qb = D.q + (size_t)b * D.nq * 3;
    const float *pb = D.p + (size_t)b * D.np * 3;

    float qx[NMD_QPT], qy[NMD_QPT], qz[NMD_QPT], best[NMD_QPT];
    int besti[NMD_QPT];
#pragma unroll
    for (int r = 0; r < NMD_QPT; ++r) {
        const int j = q0 + r * NMD_THREADS + threadIdx.x;  // coalesced across the CTA
        const int jj = j < D.nq ? j : D.nq - 1;
        qx[r] = __ldg(qb + jj * 3 + 0);
        qy[r] = __ldg(qb + jj * 3 + 1);
        qz[r] = __ldg(qb + jj * 3 + 2);
        best[r] = __int_as_float(0x7f800000);  // +inf; the reference seeds with candidate 0 (:33)
        besti[r] = 0;
    }

    for (int k2 = 0; k2 < D.np; k2 += NMD_TILE) {
        const int cnt = min(NMD_TILE, D.np - k2);
        __syncthreads();
        const float *src = pb + (size_t)k2 * 3;          // the tile's 3*cnt packed floats
        if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
            // 128-bit global loads over the packed (x,y,z) stream; every float lands in its point's float4 slot
            float *tf = reinterpret_cast<float *>(tile);
            const int nflt = 3 * cnt, nvec = nflt >> 2;
            for (int t = threadIdx.x; t < nvec; t += NMD_THREADS) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + t);
                const int f = 4 * t;
                tf[((f) / 3) * 4 + (f) % 3] = v.x;
                tf[((f + 1) / 3) * 4 + (f + 1) % 3] = v.y;
                tf[((f + 2) / 3) * 4 + (f + 2) % 3] = v.z;
                tf[((f + 3) / 3) * 4 + (f + 3) % 3] = v.w;
            }
            for (int f = 4 * nvec + threadIdx.x; f < nflt; f += NMD_THREADS) tf[(f / 3) * 4 + f % 3] = __ldg(src + f);
        } else {
            for (int t = threadIdx.x; t < cnt; t += NMD_THREADS) {
                const float *s = src + (size_t)t * 3;
                tile[t] = make_float4(__ldg(s), __ldg(s + 1), __ldg(s + 2), 0.f);
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            const float4 c = tile[k];
#pragma unroll
            for (int r = 0; r < NMD_QPT; ++r) {
                // x2 = buf - x1 ... d = x2*x2+y2*y2+z2*z2 (:27-30), in the reference's SASS order
                const float d = sqdist3(c.x - qx[r], c.y - qy[r], c.z - qz[r]);
                if (d < best[r]) {  // strict: the lowest index keeps a tie (:31,:45,:125)
                    best[r] = d;
                    besti[r] = k2 + k;
                }
            }
        }
    }
    if (D.np <= 0) return;
#pragma unroll
    for (int r = 0; r < NMD_QPT; ++r) {
        const int j = q0 + r * NMD_THREADS + threadIdx.x;
        if (j < D.nq) {
            D.dist[(size_t)b * D.nq + j] = best[r];
            D.idx[(size_t)b * D.nq + j] = besti[r];
        }
    }
}

struct NmdGradDir {
    const float *p1;  // (b,n1,3) the set the distances belong to
    const float *p2;  // (b,n2,3) the set idx points into
    const float *gdist;
    const int32_t *idx;
    float *g1, *g2;
    int n1, n2;
};

__global__ void __launch_bounds__(256) nmdist_bwd_kernel(NmdGradDir d0, NmdGradDir d1) {
    const NmdGradDir D = blockIdx.z == 0 ? d0 : d1;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= D.n1) return;
    const int b = blockIdx.y;
    const size_t o1 = ((size_t)b * D.n1 + j) * 3;
    const float x1 = __ldg(D.p1 + o1), y1 = __ldg(D.p1 + o1 + 1), z1 = __ldg(D.p1 + o1 + 2);
    const int j2 = __ldg(D.idx + (size_t)b * D.n1 + j);
    const size_t o2 = ((size_t)b * D.n2 + j2) * 3;
    const float x2 = __ldg(D.p2 + o2), y2 = __ldg(D.p2 + o2 + 1), z2 = __ldg(D.p2 + o2 + 2);
    const float g = __fmul_rn(__ldg(D.gdist + (size_t)b * D.n1 + j), 2.0f);  // :164
    const float gx = __fmul_rn(g, x1 - x2), gy = __fmul_rn(g, y1 - y2), gz = __fmul_rn(g, z1 - z2);
    // accumulate: the other direction adds into the same buffers (:165-170)
    atomicAdd(D.g1 + o1 + 0, gx);
    atomicAdd(D.g1 + o1 + 1, gy);
    atomicAdd(D.g1 + o1 + 2, gz);
    atomicAdd(D.g2 + o2 + 0, -gx);
    atomicAdd(D.g2 + o2 + 1, -gy);
    atomicAdd(D.g2 + o2 + 2, -gz);
}

}  // namespace pu3

using namespace pu3;

extern "C" int pu3_nmdist_fwd_f32(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1,
                                  float *dist2, int32_t *idx1, int32_t *idx2, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && n >= 0 && m >= 0, "nmdist_fwd: negative size b=%d n=%d m=%d", b, n, m);
    if (b == 0 || (n == 0 && m == 0)) return PU3_OK;
    PU3_ARG_CHECK(b <= 65535, "nmdist_fwd: b=%d exceeds 65535", b);
    PU3_ARG_CHECK(xyz1 && xyz2 && dist1 && dist2 && idx1 && idx2, "nmdist_fwd: null pointer");
    NmdDir d0{xyz1, xyz2, dist1, idx1, n, m}, d1{xyz2, xyz1, dist2, idx2, m, n};
    const int per = NMD_THREADS * NMD_QPT;
    dim3 grid((max(n, m) + per - 1) / per, b, 2);
    nmdist_fwd_kernel<<<grid, NMD_THREADS, 0, as_stream(stream)>>>(d0, d1);
    PU3_LAUNCH_CHECK("nmdist_fwd_kernel");
    return PU3_OK;
}

extern "C" int pu3_nmdist_bwd_f32(int b, int n, int m, const float *xyz1, const float *xyz2, float *gradxyz1,
                                  float *gradxyz2, const float *graddist1, const float *graddist2,
                                  const int32_t *idx1, const int32_t *idx2, pu3_stream_t stream) {
    PU3_ARG_CHECK(b >= 0 && n >= 0 && m >= 0, "nmdist_bwd: negative size b=%d n=%d m=%d", b, n, m);
    if (b == 0 || n == 0 || m == 0) return PU3_OK;
    PU3_ARG_CHECK(b <= 65535, "nmdist_bwd: b=%d exceeds 65535", b);
    PU3_ARG_CHECK(xyz1 && xyz2 && gradxyz1 && gradxyz2 && graddist1 && graddist2 && idx1 && idx2,
                  "nmdist_bwd: null pointer");
    NmdGradDir d0{xyz1, xyz2, graddist1, idx1, gradxyz1, gradxyz2, n, m};
    NmdGradDir d1{xyz2, xyz1, graddist2, idx2, gradxyz2, gradxyz1, m, n};
    dim3 grid((max(n, m) + 255) / 256, b, 2);
    nmdist_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(d0, d1);
    PU3_LAUNCH_CHECK("nmdist_bwd_kernel");
    return PU3_OK;
}
