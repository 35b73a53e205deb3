// Level engine: ONE C-ABI call enqueues every kernel of a Level.forward (network/upsampler.py:272-374) --
// layer0, 4 x {prep conv, feature-space kNN, fused DenseEdgeConv}, inter-level skip connection, expansion head.
//
// Why it exists: with the kernels at a few hundred microseconds each, the eval step became bound by the host
// (one Python -> ctypes crossing, several torch.empty() and stream look-ups per kernel; a B=1 forward spent as
// long issuing ~290 launches as the GPU spent running them, profiles/r1_bench_history.md).  The engine carves all
// intermediates out of one caller-provided workspace and launches back to back from C++; nothing in it
// synchronises, so a level is also CUDA-graph capturable.
#include "pu3_common.cuh"

extern "C" {
int pu3_iota_i32(int n, int32_t *out, pu3_stream_t stream);
}

namespace pu3 {
__global__ void iota_kernel(int n, int32_t *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}
static size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

struct LevelPlan {
    size_t off_h, off_idx, off_me, off_pre, off_h1, off_h2, off_h3, off_skipidx, off_knnws, off_wsplit[3], off_wprep[3], total;
    size_t knn_ws;
};

static LevelPlan plan_level(int t, int n, int r, int knn, int fm_knn, int clouds, int no, bool has_prev) {
    LevelPlan p;
    size_t off = 0;
    p.off_h = off;   off += al256((size_t)t * 24 * n * 4);
    p.off_idx = off; off += al256((size_t)t * n * (knn + 1) * 4);
    p.off_me = off;  off += al256((size_t)t * 4);
    p.off_pre = off; off += al256((size_t)t * 128 * n * 4);
    p.off_h1 = off;  off += al256((size_t)t * 128 * n * r * 4);
    p.off_h2 = off;  off += al256((size_t)t * 128 * n * r * 4);
    p.off_h3 = off;  off += al256((size_t)t * 64 * n * r * 4);
    p.off_skipidx = off; off += al256(has_prev ? (size_t)t * n * fm_knn * 8 : 0);
    size_t w1 = pu3_group_knn_workspace(t, 24, n, n, knn + 1, 1, 1);
    size_t w2 = has_prev ? pu3_group_knn_workspace(t, 3, n, no, fm_knn, 1, 1) : 0;
    p.knn_ws = w1 > w2 ? w1 : w2;
    p.off_knnws = off; off += al256(p.knn_ws);
    // tf32 hi/lo images of the head weights for the tensor-core path (csrc/conv_tc.cu): up1 (264 feature columns), up2, fc1
    p.off_wsplit[0] = off; off += al256(pu3_conv_tc_wsplit_bytes(264, 128));
    p.off_wsplit[1] = off; off += al256(pu3_conv_tc_wsplit_bytes(128, 128));
    p.off_wsplit[2] = off; off += al256(pu3_conv_tc_wsplit_bytes(128, 64));
    for (int i = 0; i < 3; ++i) { p.off_wprep[i] = off; off += al256(pu3_conv_tc_wsplit_bytes(84 + 60 * i, 24)); }   // layerK_prep
    p.total = off;
    (void)clouds;
    return p;
}
}  // namespace pu3

using namespace pu3;

// Test / A-B hook: 3 (default) = the expansion head as one fused tcgen05 kernel (eval) + the three 24-channel prep convolutions
// on the tcgen05 conv kernel, 2 = head as three tcgen05 kernels + prep convolutions, 1 = three-kernel head only,
// 0 = everything on the fp32 FFMA SGEMM.
static int g_level_tc = 3;
extern "C" void pu3_level_set_tc(int on) { g_level_tc = on; }

// Test hook (teacher forcing): neighbour lists to use INSTEAD of the engine's own searches -- idx[blk] (t,n,knn+1) i32 for the four
// dense blocks, skip (t,n,fm_knn) i64 for the skip connection; NULL entries keep the search.  With the oracle's lists injected,
// every continuous stage of a Level can be compared at the full 1e-5 tolerance on every element (no near-tie excuses).
static const int32_t *g_knn_override[4] = {nullptr, nullptr, nullptr, nullptr};
static const int64_t *g_skip_override = nullptr;
extern "C" void pu3_level_set_knn_override(const int32_t *b0, const int32_t *b1, const int32_t *b2, const int32_t *b3, const int64_t *skip) {
    g_knn_override[0] = b0; g_knn_override[1] = b1; g_knn_override[2] = b2; g_knn_override[3] = b3; g_skip_override = skip;
}

extern "C" int pu3_iota_i32(int n, int32_t *out, pu3_stream_t stream) {
    if (n <= 0) return PU3_OK;
    iota_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(n, out);
    PU3_LAUNCH_CHECK("iota_kernel");
    return PU3_OK;
}

extern "C" size_t pu3_level_workspace(int t, int n, int r, int knn, int fm_knn, int clouds, int no, int has_prev) {
    if (t <= 0 || n <= 0 || r <= 0 || knn <= 0) return 0;
    return plan_level(t, n, r, knn, fm_knn, clouds, no, has_prev != 0).total;
}

static int level_forward_impl(const pu3_level_weights *w, int t, int n, const float *xyz, const float *xyz_norm,
                              const int32_t *owner, int groups, int max_group, const float *prev_xyz,
                              const float *prev_feat_pm, int clouds, int no, const int32_t *prev_n,
                              float *feat, float *out_xyz, void *workspace, size_t workspace_bytes,
                              const pu3_level_saved *saved, float *feat_pm_out, pu3_stream_t stream);

extern "C" int pu3_level_forward_f32(const pu3_level_weights *w, int t, int n, const float *xyz, const float *xyz_norm,
                                     const int32_t *owner, int groups, int max_group, const float *prev_xyz,
                                     const float *prev_feat_pm, int clouds, int no, const int32_t *prev_n,
                                     float *feat, float *out_xyz, void *workspace, size_t workspace_bytes,
                                     pu3_stream_t stream) {
    return level_forward_impl(w, t, n, xyz, xyz_norm, owner, groups, max_group, prev_xyz, prev_feat_pm, clouds, no, prev_n,
                              feat, out_xyz, workspace, workspace_bytes, nullptr, nullptr, stream);
}

// ... and the features once more point-major (t,n,264) for the next level's skip connection: written by the skip kernel of THIS
// level while the rows are in registers (levels without a previous level: a transposing pass)
extern "C" int pu3_level_forward_pm_f32(const pu3_level_weights *w, int t, int n, const float *xyz, const float *xyz_norm,
                                        const int32_t *owner, int groups, int max_group, const float *prev_xyz,
                                        const float *prev_feat_pm, int clouds, int no, const int32_t *prev_n,
                                        float *feat, float *out_xyz, float *feat_pm_out, void *workspace, size_t workspace_bytes,
                                        pu3_stream_t stream) {
    return level_forward_impl(w, t, n, xyz, xyz_norm, owner, groups, max_group, prev_xyz, prev_feat_pm, clouds, no, prev_n,
                              feat, out_xyz, workspace, workspace_bytes, nullptr, feat_pm_out, stream);
}

// The same forward keeping what the backward pass needs (train step, model.py:53-66): the 24-channel input of every dense
// block, the neighbour indices of every block, the skip connection's indices and weights, and the two 128-channel activations of
// the head.  `saved` NULL = plain forward.
extern "C" int pu3_level_forward_train_f32(const pu3_level_weights *w, int t, int n, const float *xyz, const float *xyz_norm,
                                           const int32_t *owner, int groups, int max_group, const float *prev_xyz,
                                           const float *prev_feat_pm, int clouds, int no, const int32_t *prev_n,
                                           float *feat, float *out_xyz, void *workspace, size_t workspace_bytes,
                                           const pu3_level_saved *saved, pu3_stream_t stream) {
    return level_forward_impl(w, t, n, xyz, xyz_norm, owner, groups, max_group, prev_xyz, prev_feat_pm, clouds, no, prev_n,
                              feat, out_xyz, workspace, workspace_bytes, saved, nullptr, stream);
}

static int level_forward_impl(const pu3_level_weights *w, int t, int n, const float *xyz, const float *xyz_norm,
                              const int32_t *owner, int groups, int max_group, const float *prev_xyz,
                              const float *prev_feat_pm, int clouds, int no, const int32_t *prev_n,
                              float *feat, float *out_xyz, void *workspace, size_t workspace_bytes,
                              const pu3_level_saved *saved, float *feat_pm_out, pu3_stream_t stream) {
    PU3_ARG_CHECK(w && t >= 0 && n > 0, "level_forward: bad arguments");
    if (t == 0) return PU3_OK;
    PU3_ARG_CHECK(xyz_norm && feat && out_xyz, "level_forward: null pointer");
    const int r = w->r, K = w->knn, C = 264;
    const bool has_prev = prev_xyz != nullptr && w->fm_knn > 0;
    PU3_ARG_CHECK(!has_prev || (xyz && prev_feat_pm && clouds > 0 && no > 0), "level_forward: incomplete previous level");
    PU3_ARG_CHECK(n >= K + 1, "level_forward: points size must be greater or equal to k (n=%d, k=%d)", n, K + 1);
    const LevelPlan p = plan_level(t, n, r, K, w->fm_knn, clouds, no, has_prev);
    if (!workspace || workspace_bytes < p.total) {
        set_error("level_forward: workspace %zu bytes, need %zu", workspace_bytes, p.total);
        return PU3_E_WORKSPACE;
    }
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    float *h = saved ? saved->h[0] : reinterpret_cast<float *>(ws + p.off_h);
    int32_t *idx = saved ? saved->idx[0] : reinterpret_cast<int32_t *>(ws + p.off_idx);
    PU3_ARG_CHECK(!saved || (saved->h[0] && saved->h[1] && saved->h[2] && saved->h[3] && saved->idx[0] && saved->idx[1] && saved->idx[2] &&
                             saved->idx[3] && saved->h1 && saved->h2), "level_forward: incomplete saved buffers");
    int32_t *me = reinterpret_cast<int32_t *>(ws + p.off_me);
    float *pre = reinterpret_cast<float *>(ws + p.off_pre);
    float *h1 = saved ? saved->h1 : reinterpret_cast<float *>(ws + p.off_h1);
    float *h2 = saved ? saved->h2 : reinterpret_cast<float *>(ws + p.off_h2);
    float *h3 = reinterpret_cast<float *>(ws + p.off_h3);
    int64_t *skipidx = (saved && saved->skip_idx) ? saved->skip_idx : reinterpret_cast<int64_t *>(ws + p.off_skipidx);
    void *knnws = ws + p.off_knnws;
    const long long fs = (long long)C * n;   // batch stride of the feature buffer
    int st;
#define PU3_TRY(call) do { st = (call); if (st) return st; } while (0)
#define PU3_TRYT(tag, call) do { const int _pid = prof_begin(tag, as_stream(stream)); st = (call); prof_end(_pid, as_stream(stream)); if (st) return st; } while (0)
    if (owner) PU3_TRY(pu3_iota_i32(t, me, stream));
    // layer0 (upsampler.py:288): 3 -> 24, no activation; x0 also lives in the last 24 channels of feat
    PU3_TRYT(PROF_CONV, pu3_pointwise_conv_f32(t, n, 3, 24, xyz_norm, 3LL * n, w->layer0_w, w->layer0_b, h, 24LL * n, nullptr, 0, 1, 1, 0, stream));
    PU3_TRY(cuda_status(cudaMemcpy2DAsync(feat + (size_t)(C - 24) * n, fs * 4, h, 24LL * n * 4, 24LL * n * 4, t,
                                          cudaMemcpyDeviceToDevice, as_stream(stream)), "level_forward: copy x0"));
    int lo = C - 24;
    for (int blk = 0; blk < 4; ++blk) {
        if (saved && blk > 0) { h = saved->h[blk]; idx = saved->idx[blk]; }
        if (blk > 0) {   // layerK_prep (:213-221): Conv1d + ReLU over everything produced so far
            if (g_level_tc >= 2 && n % 4 == 0) {   // tensor cores (3xTF32), cout 24 padded to the 64-column MMA
                void *wsp = ws + p.off_wprep[blk - 1];
                PU3_TRYT(PROF_CONV_TC_PREP, pu3_conv_tc_prepare_f32(C - lo, 24, w->prep_w[blk - 1], C - lo, wsp, stream));
                PU3_TRYT(PROF_CONV_TC_PREP, pu3_conv_tc_f32(t, n, C - lo, 24, feat + (size_t)lo * n, fs, wsp, w->prep_b[blk - 1], h, 24LL * n, 1, stream));
            } else {
                PU3_TRYT(PROF_CONV, pu3_pointwise_conv_f32(t, n, C - lo, 24, feat + (size_t)lo * n, fs, w->prep_w[blk - 1], w->prep_b[blk - 1],
                                               h, 24LL * n, nullptr, 0, 1, 1, 1, stream));
            }
        }
        // dynamic graph in feature space (layers.py:33): k+1 nearest, duplicates pushed back, rank 0 dropped by idx_off=1;
        // the edge-conv takes a max over the other k, so they are requested as a set (PU3_KNN_SET_ORDER)
        if (g_knn_override[blk])
            PU3_TRY(cuda_status(cudaMemcpyAsync(idx, g_knn_override[blk], (size_t)t * n * (K + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                                as_stream(stream)), "level_forward: neighbour-list override"));
        else if (owner)
            PU3_TRYT(PROF_KNN_FEAT, pu3_group_knn_ragged_f32(t, 24, n, n, K + 1, t, groups, me, owner, nullptr, nullptr, h, h, 1 | PU3_KNN_SET_ORDER, nullptr,
                                             nullptr, idx, nullptr, knnws, p.knn_ws, stream));
        else
            PU3_TRYT(PROF_KNN_FEAT, pu3_group_knn_f32(t, 24, n, n, K + 1, 1, h, h, 1 | PU3_KNN_SET_ORDER, max_group, nullptr, nullptr, idx, nullptr, knnws,
                                      p.knn_ws, stream));
        PU3_TRYT(PROF_EDGECONV, (saved ? pu3_edgeconv_ffma_f32 : pu3_edgeconv_f32)(t, n, K, h, 24LL * n, idx, K + 1, 1, w->ec_w[blk][0], w->ec_b[blk][0], w->ec_w[blk][1],
                                 w->ec_b[blk][1], w->ec_w[blk][2], w->ec_b[blk][2], feat + (size_t)(lo - 60) * n, fs, stream));
        lo -= 60;
    }
    if (has_prev) {   // inter-level skip connection (:317-347)
        if (g_skip_override)
            PU3_TRY(cuda_status(cudaMemcpyAsync(skipidx, g_skip_override, (size_t)t * n * w->fm_knn * sizeof(int64_t), cudaMemcpyDeviceToDevice,
                                                as_stream(stream)), "level_forward: skip neighbour-list override"));
        else if (owner)
            PU3_TRYT(PROF_KNN_SKIP, pu3_group_knn_ragged_f32(t, 3, n, no, w->fm_knn, clouds, groups, owner, owner, prev_n, nullptr, xyz, prev_xyz,
                                             1, nullptr, skipidx, nullptr, nullptr, knnws, p.knn_ws, stream));
        else
            PU3_TRYT(PROF_KNN_SKIP, pu3_group_knn_f32(t, 3, n, no, w->fm_knn, t / clouds, xyz, prev_xyz, 1, max_group, nullptr, skipidx, nullptr,
                                      nullptr, knnws, p.knn_ws, stream));
        if (saved && saved->feat_pre)
            PU3_TRY(cuda_status(cudaMemcpyAsync(saved->feat_pre, feat, (size_t)t * C * n * sizeof(float), cudaMemcpyDeviceToDevice,
                                                as_stream(stream)), "level_forward: copy of the pre-skip features"));
        PU3_TRYT(PROF_SKIP_FUSE, pu3_skip_fuse_pm_f32(t, n, C, w->fm_knn, owner ? 1 : t / clouds, no, feat, xyz, skipidx, prev_xyz, prev_feat_pm,
                                  owner, saved ? saved->skip_w : nullptr, feat_pm_out, stream));
    } else if (feat_pm_out) {
        PU3_TRYT(PROF_MISC, pu3_to_point_major_f32(t, C, n, feat, nullptr, feat_pm_out, stream));
    }
    // expansion head (:349-372)
    if (g_level_tc >= 3 && n % 4 == 0 && r == 2 && !saved) {
        // the whole head as ONE persistent tcgen05 kernel (csrc/head_tc.cu): activations chained on chip, HBM sees the
        // features once and the 3-channel result (the train-mode forward keeps the 3-kernel path: its backward needs h1, h2)
        void *ws1 = ws + p.off_wsplit[0], *ws2 = ws + p.off_wsplit[1], *ws3 = ws + p.off_wsplit[2];
        PU3_TRYT(PROF_CONV_TC, pu3_conv_tc_prepare_f32(C, 128, w->up1_w, C + 1, ws1, stream));
        PU3_TRYT(PROF_CONV_TC, pu3_conv_tc_prepare_f32(128, 128, w->up2_w, 128, ws2, stream));
        PU3_TRYT(PROF_CONV_TC, pu3_conv_tc_prepare_f32(128, 64, w->fc1_w, 128, ws3, stream));
        PU3_TRYT(PROF_CONV_TC, pu3_head_tc_f32(t, n, C, feat, fs, ws1, ws2, ws3, w->up1_w, C + 1, C, w->up1_b, w->code, w->up2_b, w->fc1_b,
                                               w->fc2_w, w->fc2_b, xyz_norm, 3LL * n, out_xyz, 3LL * n * r, stream));
    } else if (g_level_tc >= 1 && n % 4 == 0 && r <= 8) {
        // tensor cores (tcgen05, 3xTF32): up1 + code column + replication | up2 | fc1 + fc2 + residual -- 3 kernels, the
        // (t,265,n*r) input, the 128-channel "pre" tensor and the 64-channel activation never exist
        void *ws1 = ws + p.off_wsplit[0], *ws2 = ws + p.off_wsplit[1], *ws3 = ws + p.off_wsplit[2];
        PU3_TRYT(PROF_CONV_TC, pu3_conv_tc_prepare_f32(C, 128, w->up1_w, C + 1, ws1, stream));
        PU3_TRYT(PROF_CONV_TC, pu3_conv_tc_prepare_f32(128, 128, w->up2_w, 128, ws2, stream));
        PU3_TRYT(PROF_CONV_TC, pu3_conv_tc_prepare_f32(128, 64, w->fc1_w, 128, ws3, stream));
        PU3_TRYT(PROF_CONV_TC, pu3_conv_tc_expand_f32(t, n, C, 128, r, feat, fs, ws1, w->up1_w, C + 1, C, w->up1_b, w->code, h1, 128LL * n * r, stream));
        PU3_TRYT(PROF_CONV_TC, pu3_conv_tc_f32(t, n * r, 128, 128, h1, 128LL * n * r, ws2, w->up2_b, h2, 128LL * n * r, 1, stream));
        PU3_TRYT(PROF_CONV_TC, pu3_conv_tc_project_f32(t, n * r, 128, 64, 3, h2, 128LL * n * r, ws3, w->fc1_b, w->fc2_w, w->fc2_b, out_xyz,
                                                       3LL * n * r, xyz_norm, 3LL * n, n, r, stream));
    } else {
        PU3_TRYT(PROF_CONV, pu3_pointwise_conv_f32(t, n, C, 128, feat, fs, w->up1_w_feat, w->up1_b, pre, 128LL * n, nullptr, 0, 1, 1, 0, stream));
        PU3_TRYT(PROF_EXPAND, pu3_expand_code_f32(t, 128, n, r, pre, w->up1_w, C + 1, C, w->code, h1, stream));
        PU3_TRYT(PROF_CONV, pu3_pointwise_conv_f32(t, n * r, 128, 128, h1, 128LL * n * r, w->up2_w, w->up2_b, h2, 128LL * n * r, nullptr, 0, 1, 1, 1, stream));
        PU3_TRYT(PROF_CONV, pu3_pointwise_conv_f32(t, n * r, 128, 64, h2, 128LL * n * r, w->fc1_w, w->fc1_b, h3, 64LL * n * r, nullptr, 0, 1, 1, 1, stream));
        PU3_TRYT(PROF_CONV, pu3_pointwise_conv_f32(t, n * r, 64, 3, h3, 64LL * n * r, w->fc2_w, w->fc2_b, out_xyz, 3LL * n * r, xyz_norm, 3LL * n, n, r, 0, stream));
    }
#undef PU3_TRY
#undef PU3_TRYT
    return PU3_OK;
}
