/*
 * oracle_c.c -- CPU restatement of the reference's two CUDA extensions.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may link, import
 * or call this file; it is the checker the CUDA kernels are compared with
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference).
 *
 * Each function follows one reference kernel statement by statement and
 * emulates the reference's launch geometry where the geometry is observable
 * in the results (FPS tie rule).  Floating point: the reference is compiled by
 * nvcc with default -fmad=true; its SASS for sm_100a (checked with cuobjdump)
 * evaluates   dx*dx + dy*dy + dz*dz   as   FMUL(dy,dy) -> FFMA(dx,dx,.) ->
 * FFMA(dz,dz,.)   in BOTH nmdistance_cuda.cu and sampling_cuda.cu, so that is
 * what sqdist3() spells out with fmaf().  Build with -ffp-contract=off so gcc
 * adds no contraction of its own.
 *
 * Parity pin: the reference has no tests and no golden vectors for this path
 * (SURVEY.md section 4).  This file is pinned against tests/golden/ref_cuda_*.npz,
 * which are outputs of the reference's own .cu files compiled for sm_100a and
 * run on a B200 (recipe: oracle/build_ref.py + tests/golden/make_golden_gpu.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float sqdist3(float dx, float dy, float dz) {
    /* reference SASS order, see header */
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* cuda_utils.h:9-14  opt_n_threads: 2^floor(log2(work)) clamped to [1,512] */
int oracle_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 512) t = 512;
    if (t < 1) t = 1;
    return t;
}

/*
 * sampling_cuda.cu:103-174 furthest_point_sampling_forward_kernel and its
 * launcher :176-265.  xyz (b,n,3), temp (b,n) in/out (caller fills 1e10),
 * idx (b,m) int32.
 *
 * The kernel is run by T = opt_n_threads(n) threads; thread t scans
 * k = t, t+T, ... keeping the first strictly greater value (:147), then a
 * shared-memory tree keeps the LOWER slot on ties (:162 strict '<').  The
 * emulation below keeps per-thread (best,besti) slots and runs the same tree.
 *
 * legacy_temp_rows != 0 reproduces the reference's indexing of temp by
 * blockIdx.x (:131,146) with gridDim.x = min(32,(n*b+T/2)/T) -- identical to
 * per-batch rows while b <= gridDim.x, wrong beyond (documented reference bug).
 * With 0, every batch element owns its temp row (what the product does).
 */
void oracle_fps(int b, int n, int m, const float *xyz, float *temp, int32_t *idx,
                int legacy_temp_rows) {
    if (m <= 0 || n <= 0) return;
    const int T = oracle_opt_n_threads(n);
    int grid = (int)(((long long)n * b + T / 2) / T);
    if (grid > 32) grid = 32;
    if (grid < 1) grid = 1;
    float *dists = (float *)malloc(sizeof(float) * T);
    int *dists_i = (int *)malloc(sizeof(int) * T);
    /* blocks run in an unspecified order on the GPU; rows only collide in the
     * buggy b > grid case, where we pick ascending batch order. */
    for (int i = 0; i < b; ++i) {
        const float *p = xyz + (size_t)i * n * 3;
        float *trow = temp + (size_t)(legacy_temp_rows ? (i % grid) : i) * n;
        int old = 0;
        idx[(size_t)i * m + 0] = old;
        for (int j = 1; j < m; ++j) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int t = 0; t < T; ++t) {
                int besti = 0;
                float best = -1.0f;
                for (int k = t; k < n; k += T) {
                    const float td = trow[k];
                    const float d = sqdist3(p[k * 3 + 0] - x1, p[k * 3 + 1] - y1, p[k * 3 + 2] - z1);
                    const float d2 = fminf(d, td);      /* CUDA min(float,float) == fminf */
                    if (d2 != td) trow[k] = d2;
                    if (d2 > best) { best = d2; besti = k; }
                }
                dists[t] = best;
                dists_i[t] = besti;
            }
            for (int u = 0; (1 << u) < T; ++u) {
                const int active = T >> (u + 1);
                for (int t = 0; t < active; ++t) {
                    const int i1 = (t * 2) << u, i2 = (t * 2 + 1) << u;
                    if (dists[i1] < dists[i2]) { dists[i1] = dists[i2]; dists_i[i1] = dists_i[i2]; }
                }
            }
            old = dists_i[0];
            idx[(size_t)i * m + j] = old;
        }
    }
    free(dists);
    free(dists_i);
}

/* sampling_cuda.cu:26-41 gather_points_forward_kernel: out[b,c,j]=points[b,c,idx[b,j]] */
#define DEF_GATHER_FWD(NAME, T)                                                              \
    void NAME(int b, int c, int n, int m, const T *points, const int32_t *idx, T *out) {     \
        for (int i = 0; i < b; ++i)                                                          \
            for (int l = 0; l < c; ++l)                                                      \
                for (int j = 0; j < m; ++j) {                                                \
                    const int a = idx[(size_t)i * m + j];                                    \
                    out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + a];  \
                }                                                                            \
    }
DEF_GATHER_FWD(oracle_gather_fwd_f32, float)
DEF_GATHER_FWD(oracle_gather_fwd_f64, double)
DEF_GATHER_FWD(oracle_gather_fwd_u16, uint16_t) /* f16 is a pure copy: move the bits */

/* sampling_cuda.cu:64-80 gather_points_backward_kernel: atomicAdd scatter into a
 * caller-zeroed grad_points.  Sequential j order here; the GPU order is
 * unspecified, so float results agree only up to summation order. */
#define DEF_GATHER_BWD(NAME, T)                                                              \
    void NAME(int b, int c, int n, int m, const T *grad_out, const int32_t *idx,             \
              T *grad_points) {                                                              \
        for (int i = 0; i < b; ++i)                                                          \
            for (int l = 0; l < c; ++l)                                                      \
                for (int j = 0; j < m; ++j) {                                                \
                    const int a = idx[(size_t)i * m + j];                                    \
                    grad_points[((size_t)i * c + l) * n + a] +=                              \
                        grad_out[((size_t)i * c + l) * m + j];                               \
                }                                                                            \
    }
DEF_GATHER_BWD(oracle_gather_bwd_f32, float)
DEF_GATHER_BWD(oracle_gather_bwd_f64, double)

/*
 * nmdistance_cuda.cu:11-133 NmDistanceKernel, one direction: for every point of
 * xyz (b,n,3) the smallest squared distance to xyz2 (b,m,3) and its index.
 * Tiles of 512 candidates; inside a tile a strict '<' scan seeded by the tile's
 * first candidate (:33 'k==0 ||'), across tiles strict '>' (:125) -- together:
 * the lowest index that attains the minimum.  The 4x unrolling of the
 * reference does not change the visiting order, so it is not reproduced.
 */
void oracle_nmdist_dir(int b, int n, const float *xyz, int m, const float *xyz2,
                       float *result, int32_t *result_i) {
    const int batch = 512;
    for (int i = 0; i < b; ++i) {
        for (int k2 = 0; k2 < m; k2 += batch) {
            const int end_k = (m < k2 + batch ? m : k2 + batch) - k2;
            const float *buf = xyz2 + ((size_t)i * m + k2) * 3;
            for (int j = 0; j < n; ++j) {
                const float x1 = xyz[((size_t)i * n + j) * 3 + 0];
                const float y1 = xyz[((size_t)i * n + j) * 3 + 1];
                const float z1 = xyz[((size_t)i * n + j) * 3 + 2];
                int best_i = 0;
                float best = 0;
                for (int k = 0; k < end_k; ++k) {
                    const float d = sqdist3(buf[k * 3 + 0] - x1, buf[k * 3 + 1] - y1, buf[k * 3 + 2] - z1);
                    if (k == 0 || d < best) { best = d; best_i = k + k2; }
                }
                if (k2 == 0 || result[(size_t)i * n + j] > best) {
                    result[(size_t)i * n + j] = best;
                    result_i[(size_t)i * n + j] = best_i;
                }
            }
        }
    }
}

/* nmdistance_cuda.cu:135-153 chamfer_cuda_forward: both directions */
void oracle_nmdist_fwd(int b, int n, int m, const float *xyz1, const float *xyz2,
                       float *dist1, int32_t *idx1, float *dist2, int32_t *idx2) {
    oracle_nmdist_dir(b, n, xyz1, m, xyz2, dist1, idx1);
    oracle_nmdist_dir(b, m, xyz2, n, xyz1, dist2, idx2);
}

/* nmdistance_cuda.cu:154-173 NmDistanceGradKernel, one direction, accumulating
 * into caller-zeroed grads (model_loss.py:25-26).  g = 2*grad_dist (SASS: FADD g,g);
 * +g*(p1-p2) to grad_xyz1[j], the exact negation to grad_xyz2[idx[j]]. */
static void oracle_nmdist_grad_dir(int b, int n, const float *xyz1, int m, const float *xyz2,
                                   const float *grad_dist1, const int32_t *idx1,
                                   float *grad_xyz1, float *grad_xyz2) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const float x1 = xyz1[((size_t)i * n + j) * 3 + 0];
            const float y1 = xyz1[((size_t)i * n + j) * 3 + 1];
            const float z1 = xyz1[((size_t)i * n + j) * 3 + 2];
            const int j2 = idx1[(size_t)i * n + j];
            const float x2 = xyz2[((size_t)i * m + j2) * 3 + 0];
            const float y2 = xyz2[((size_t)i * m + j2) * 3 + 1];
            const float z2 = xyz2[((size_t)i * m + j2) * 3 + 2];
            const float g = grad_dist1[(size_t)i * n + j] * 2;
            grad_xyz1[((size_t)i * n + j) * 3 + 0] += g * (x1 - x2);
            grad_xyz1[((size_t)i * n + j) * 3 + 1] += g * (y1 - y2);
            grad_xyz1[((size_t)i * n + j) * 3 + 2] += g * (z1 - z2);
            grad_xyz2[((size_t)i * m + j2) * 3 + 0] += -(g * (x1 - x2));
            grad_xyz2[((size_t)i * m + j2) * 3 + 1] += -(g * (y1 - y2));
            grad_xyz2[((size_t)i * m + j2) * 3 + 2] += -(g * (z1 - z2));
        }
}

/* nmdistance_cuda.cu:175-194 chamfer_cuda_backward */
void oracle_nmdist_bwd(int b, int n, int m, const float *xyz1, const float *xyz2,
                       const float *graddist1, const float *graddist2,
                       const int32_t *idx1, const int32_t *idx2,
                       float *gradxyz1, float *gradxyz2) {
    oracle_nmdist_grad_dir(b, n, xyz1, m, xyz2, graddist1, idx1, gradxyz1, gradxyz2);
    oracle_nmdist_grad_dir(b, m, xyz2, n, xyz1, graddist2, idx2, gradxyz2, gradxyz1);
}

/* sampling_cuda.cu:267-305 query_ball_point_kernel (dead in the reference: exported at
 * sampling.cpp:88, never called from Python).  Restated for surface completeness. */
void oracle_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                       const float *xyz, int32_t *idx) {
    const float radius2 = radius * radius;
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            const float *q = new_xyz + ((size_t)i * m + j) * 3;
            int32_t *o = idx + ((size_t)i * m + j) * nsample;
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                const float *p = xyz + ((size_t)i * n + k) * 3;
                /* the reference writes (q-p)*(q-p)+... ; SASS contraction as sqdist3 */
                const float d2 = sqdist3(q[0] - p[0], q[1] - p[1], q[2] - p[2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                }
            }
        }
}
