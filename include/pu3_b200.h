/*
 * pu3_b200.h -- C ABI of the B200-native 3PU patch-upsampling hot path.
 *
 * One shared library (3pu_pytorch_b200/lib/libpu3_b200.so), plain pointers and
 * sizes, no torch types.  Every pointer is a DEVICE pointer on the device that
 * is current for the calling thread unless stated otherwise; every call only
 * ENQUEUES work on `stream` (a cudaStream_t passed as void*, NULL = legacy
 * default stream) and never synchronises, so all entry points can be captured
 * in CUDA graphs.  Return value: 0 on success, otherwise a negative PU3_E_*
 * code (argument error, nothing launched) or a positive cudaError_t; the text
 * is available from pu3_last_error().  Nothing here ever calls exit(): the
 * reference kills the process on a launch failure (sampling_cuda.cu:56-60,
 * 259-263) or prints and carries on (nmdistance_cuda.cu:144-149).
 *
 * Section A replaces the reference's two pybind extension modules one entry
 * point for one; section B is the part of the path the reference runs as
 * PyTorch library calls (network/operations.py, layers.py, upsampler.py) and
 * that this library runs as its own kernels.
 */
#ifndef PU3_B200_H
#define PU3_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *pu3_stream_t; /* cudaStream_t */

#define PU3_OK 0
#define PU3_E_ARG (-1)         /* bad size / null pointer / unsupported combination */
#define PU3_E_WORKSPACE (-2)   /* workspace too small: call the *_workspace() query */
#define PU3_E_UNSUPPORTED (-3) /* valid request this build cannot serve */

const char *pu3_last_error(void); /* thread-local text of the last non-zero status */
int pu3_version(void);            /* ABI version, bumped on any signature change */
/* SM count, max opt-in shared memory per block and compute capability (major*10+minor) of the current device */
int pu3_device_info(int *sm_count, int *smem_optin_bytes, int *cc);

/* ------------------------------------------------------------------------------------------
 * A. replacements of the reference's native entry points
 * ---------------------------------------------------------------------------------------- */

/*
 * Farthest point sampling.  Replaces sampling.furthest_sampling
 * (sampling/sampling.cpp:26-35 -> furthest_sampling_cuda_forward,
 * sampling/sampling_cuda.cu:103-265).
 *   xyz  (b,n,3) f32 contiguous      idx (b,m) i32 out, idx[:,0] = 0
 *   temp (b,n)   f32 in/out running min distance; the reference's caller fills it with 1e10
 *        (network/operations.py:291).  NULL = "filled with 1e10, do not write back".
 * Bit-exact with the reference kernel including its tie rule (lowest k mod T, then lowest k,
 * T = 2^floor(log2 n) capped at 512) and its FMUL/FFMA distance; differs on purpose only where
 * the reference is wrong: every batch element owns its temp row, so b > 32 is correct
 * (the reference indexes temp by blockIdx.x, sampling_cuda.cu:131,146).
 */
int pu3_fps_f32(int b, int n, int m, const float *xyz, float *temp, int32_t *idx, pu3_stream_t stream);

/*
 * Ragged batch of clouds for the batched eval path: cloud i has n_arr[i] <= n_stride points and wants
 * m_arr[i] <= m_stride samples (either array may be NULL = the stride); rows of xyz / temp / idx are
 * n_stride / m_stride apart and entries past a cloud's own counts are neither read nor written.  Each cloud
 * gets exactly the result pu3_fps_f32 would give for it alone (the tie rule uses the cloud's own n).
 * n_arr / m_arr are DEVICE arrays of b int32.
 */
int pu3_fps_ragged_f32(int b, int n_stride, int m_stride, const int32_t *n_arr, const int32_t *m_arr,
                       const float *xyz, float *temp, int32_t *idx, pu3_stream_t stream);

/*
 * Point gather, out[b,c,j] = points[b,c,idx[b,j]].  Replaces sampling.gather_forward
 * (sampling/sampling.cpp:37-45, sampling_cuda.cu:26-62).  points (b,c,n), idx (b,m) i32, out (b,c,m).
 * elem_bytes = 2, 4 or 8 (the reference dispatches half/float/double; a gather only moves bits).
 */
int pu3_gather_fwd(int b, int c, int n, int m, int elem_bytes, const void *points, const int32_t *idx,
                   void *out, pu3_stream_t stream);

/*
 * Gather backward, grad_points[b,c,idx[b,j]] += grad_out[b,c,j] into a caller-zeroed buffer.
 * Replaces sampling.gather_backward (sampling/sampling.cpp:47-53, sampling_cuda.cu:64-100).
 * dtype: 0 = f32, 1 = f64, 2 = f16.  Atomic adds, summation order unspecified like the reference.
 */
int pu3_gather_bwd(int b, int c, int n, int m, int dtype, const void *grad_out, const int32_t *idx,
                   void *grad_points, pu3_stream_t stream);

/*
 * Ball query.  Replaces sampling.ball_query (sampling/sampling.cpp:59-81, sampling_cuda.cu:267-314),
 * which no reference Python calls; kept for surface completeness.  idx (b,m,nsample) i32 must arrive
 * zero-filled (the reference allocates it with torch::zeros).
 */
int pu3_ball_query_f32(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                       const float *xyz, int32_t *idx, pu3_stream_t stream);

/*
 * Chamfer / nearest-neighbour distance, both directions.  Replaces losses.nmdistance_forward
 * (losses/nmdistance.cpp:12-14 -> chamfer_cuda_forward, nmdistance_cuda.cu:11-153).
 *   xyz1 (b,n,3), xyz2 (b,m,3) f32;  dist1,idx1 (b,n);  dist2,idx2 (b,m); idx = lowest index at the minimum.
 * Bit-exact with the reference kernel (same FMUL/FFMA order).
 */
int pu3_nmdist_fwd_f32(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1,
                       float *dist2, int32_t *idx1, int32_t *idx2, pu3_stream_t stream);

/*
 * Chamfer backward into caller-zeroed gradxyz1 (b,n,3), gradxyz2 (b,m,3).  Replaces
 * losses.nmdistance_backward (losses/nmdistance.cpp:17-21, nmdistance_cuda.cu:154-194).
 */
int pu3_nmdist_bwd_f32(int b, int n, int m, const float *xyz1, const float *xyz2, float *gradxyz1,
                       float *gradxyz2, const float *graddist1, const float *graddist2,
                       const int32_t *idx1, const int32_t *idx2, pu3_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * B. the PyTorch-level part of the path, as kernels
 * ---------------------------------------------------------------------------------------- */

/*
 * k nearest neighbours + neighbour grouping.  Replaces network.operations.group_knn
 * (network/operations.py:151-216: matmul distance matrix, CPU np.unique round trip, topk, gather).
 *   query  (b,c,m) f32, points (b/p_div,c,n) f32 -- channel-major ("NCHW"); batch element i reads
 *          points[i / p_div] (p_div > 1 = the reference's expand() of the previous level, upsampler.py:319-323)
 *   k <= n; distances are the reference's expanded form |q|^2 - 2 q.p + |p|^2 in f32
 *   unique != 0: every point that equals an earlier point of its cloud in all c channels gets
 *          max(D) added, max taken over groups of `max_group` consecutive batch elements
 *          (the reference takes it over its whole batch: max_group = b; operations.py:204)
 *   unique | PU3_KNN_SET_ORDER (bit 1, opt-in): when only idx32 is requested (the fused DenseEdgeConv: rank 0 is
 *          dropped and a max over the other neighbours follows, layers.py:33-35,63), rank 0 is exact and ranks
 *          1..k-1 are returned as a SET in unspecified order -- the k-1 serial arg-min rounds become cheap ones
 *   outputs, each may be NULL: knn (b,c,m,k) f32 contiguous; idx64 (b,m,k) i64; idx32 (b,m,k) i32;
 *          dist (b,m,k) f32, ascending; ties are ordered by ascending point index.
 *   workspace: pu3_group_knn_workspace() bytes, 256-byte aligned.
 */
#define PU3_KNN_SET_ORDER 2
size_t pu3_group_knn_workspace(int b, int c, int m, int n, int k, int p_div, int unique);
void pu3_knn_set_grid(int on);     /* test / A-B hook: 1 (default) = xyz searches (c = 3, k <= 8) over >= 1024 points use the uniform-grid kernel, 0 = always the exhaustive kernel; results are bit-identical */
void pu3_knn_no_prefilter(int on); /* test / A-B hook: 1 = the exhaustive xyz kernel without its bounding-sphere pre-filter */
int pu3_group_knn_f32(int b, int c, int m, int n, int k, int p_div, const float *query, const float *points,
                      int unique, int max_group, float *knn, int64_t *idx64, int32_t *idx32, float *dist,
                      void *workspace, size_t workspace_bytes, pu3_stream_t stream);

/*
 * group_knn over a ragged batch (batched eval, where the outlier filter of network/upsampler.py:63-73 leaves
 * every cloud its own size): batch element i reads cloud owner[i] (clouds clouds of n_arr[c] <= n valid points,
 * row stride n), has m_arr[i] <= m valid queries, and belongs to duplicate-penalty group group_of[i]
 * (0 <= group_of[i] < groups).  n_arr / m_arr may be NULL.  Rows of batch elements / queries past the valid
 * counts are not written; when a cloud holds fewer than k points only its first n_arr[c] result columns are.
 * All index arrays are DEVICE int32.  Workspace as for pu3_group_knn_f32.
 */
int pu3_group_knn_ragged_f32(int b, int c, int m, int n, int k, int clouds, int groups, const int32_t *owner,
                             const int32_t *group_of, const int32_t *n_arr, const int32_t *m_arr,
                             const float *query, const float *points, int unique, float *knn, int64_t *idx64,
                             int32_t *idx32, float *dist, void *workspace, size_t workspace_bytes,
                             pu3_stream_t stream);

/*
 * Backward of the neighbour gather of group_knn: grad_points[b/p_div, c, idx[b,m,kk]] += grad_knn[b,c,m,kk]
 * into a caller-zeroed (b/p_div,c,n) buffer (autograd of torch.gather, operations.py:209-211).
 */
int pu3_group_gather_bwd_f32(int b, int c, int m, int n, int k, int p_div, const float *grad_knn,
                             const int64_t *idx64, float *grad_points, pu3_stream_t stream);

/*
 * 1x1 convolution over points: Y[b,co,p] = act(sum_ci W[co,ci] X[b,ci,p] + bias[co]) (+ R[b,co,p/res_div]).
 * Replaces the nn.Conv1d/nn.Conv2d 1x1 layers of the reference (network/layers.py:115-204; used at
 * network/upsampler.py:209-230) including their bias, ReLU and the torch.cat that follows them: X and Y are
 * channel slices of larger (B,Ctot,N) buffers, addressed as base pointer + batch stride (elements), channel
 * stride n.  w (cout,cin) row-major as in the state_dict, bias may be NULL, relu != 0 applies max(.,0) before
 * the optional residual res (b,cout,res_n), read at column p / res_div (upsampler.py:371-372).  fp32 FFMA.
 */
int pu3_pointwise_conv_f32(int b, int n, int cin, int cout, const float *x, long long x_bstride, const float *w,
                           const float *bias, float *y, long long y_bstride, const float *res,
                           long long res_bstride, int res_n, int res_div, int relu, pu3_stream_t stream);
/*
 * The same kernel with the backward-pass epilogue: y = [y +] (mask > 0 ? act(W x + b) [+ res] : 0).  mask (b,cout,n) (its own
 * batch stride, channel stride n) is the forward activation whose ReLU derivative gates this gradient; accumulate != 0 adds to what
 * y already holds (a gradient slice with several contributors).  With W = the transposed forward weight and x = dY this is the
 * input gradient of the 1x1 convolution (model.py:62 autograd of network/layers.py:115-204).
 */
int pu3_pointwise_conv_ex_f32(int b, int n, int cin, int cout, const float *x, long long x_bstride, const float *w,
                              const float *bias, float *y, long long y_bstride, const float *res, long long res_bstride,
                              int res_n, int res_div, int relu, const float *mask, long long mask_bstride, int accumulate,
                              pu3_stream_t stream);

/*
 * Code column of the feature-expansion layer (network/upsampler.py:349-366):
 * Y[b,co,p*r+j] = relu(pre[b,co,p] + w[co*w_stride + code_col] * code[j]) with pre = W[:, :code_col] x + bias
 * computed once per point by pu3_pointwise_conv_f32; y is (b,cout,n*r) contiguous.
 */
int pu3_expand_code_f32(int b, int cout, int n, int r, const float *pre, const float *w, int w_stride, int code_col,
                        const float *code, float *y, pu3_stream_t stream);

/*
 * Tensor-core (tcgen05, 3xTF32 split: fp32-faithful) variants of the 1x1 convolution for the expansion head
 * (network/upsampler.py:349-372; nn.Conv2d 1x1 layers up_layer1/up_layer2/fc_layer1/fc_layer2, layers.py:161-204).
 * Same tensors as pu3_pointwise_conv_f32 (x: channel slice, pointer + batch stride, channel stride n), cout <= 128,
 * n % 4 == 0 and 16-byte aligned slices (TMA).  The weights are first split into tf32 hi/lo parts in the
 * shared-memory layout of the MMA by pu3_conv_tc_prepare_f32 (w (cout, >= cin) row-major with row stride
 * w_stride) into a caller buffer of pu3_conv_tc_wsplit_bytes() bytes, 16-byte aligned.
 *   pu3_conv_tc_f32          y[b,co,p] = act(W x + bias)
 *   pu3_conv_tc_expand_f32   the feature-expansion layer (upsampler.py:349-366): y[b,co,p*r+j] =
 *                            relu(W[:, :cin] x[b,:,p] + bias + w[co*w_stride+code_col] * code[j]); y (b,cout,n*r)
 *   pu3_conv_tc_project_f32  two layers (upsampler.py:369-372): h = relu(W x + bias_mid) (cmid <= 64) stays on
 *                            chip, y[b,c,p] = w_out[c,:] h + b_out[c] (+ res[b,c,p/res_div]), cout <= 3
 */
size_t pu3_conv_tc_wsplit_bytes(int cin, int cout);
void pu3_conv_tc_set_variant(int v); /* test hook (shared-memory descriptor variant); 0 = default */
void pu3_conv_tc_set_debug(float *buf); /* test hook: device buffer (>= 2050 floats) receiving the first stage; NULL = off */
int pu3_conv_tc_prepare_f32(int cin, int cout, const float *w, int w_stride, void *wsplit, pu3_stream_t stream);
int pu3_conv_tc_f32(int b, int n, int cin, int cout, const float *x, long long x_bstride, const void *wsplit,
                    const float *bias, float *y, long long y_bstride, int relu, pu3_stream_t stream);
int pu3_conv_tc_expand_f32(int b, int n, int cin, int cout, int r, const float *x, long long x_bstride,
                           const void *wsplit, const float *w, int w_stride, int code_col, const float *bias,
                           const float *code, float *y, long long y_bstride, pu3_stream_t stream);
int pu3_conv_tc_project_f32(int b, int n, int cin, int cmid, int cout, const float *x, long long x_bstride,
                            const void *wsplit, const float *bias_mid, const float *w_out, const float *b_out,
                            float *y, long long y_bstride, const float *res, long long res_bstride, int res_n,
                            int res_div, pu3_stream_t stream);

/*
 * The whole expansion head (network/upsampler.py:349-372: replicate x2 + 1-D code, up_layer1, up_layer2, fc_layer1,
 * fc_layer2, residual) as ONE persistent tcgen05 kernel: the 128/128/64-channel activations never leave the SM (the epilogue
 * of a layer writes the tf32 hi/lo operand tiles of the next one into shared memory), HBM sees the (b,cin,n) features once
 * and the (b,3,2n) result.  Step ratio 2 and the reference's head widths (128, 128, 64, 3) only; other configurations use
 * the three pu3_conv_tc_* calls.  x: channel slice (pointer + batch stride, channel stride n); ws1/ws2/ws3: split images of
 * up_layer1 (cin feature columns of its (128, >= cin+1) weight w1, code column code_col), up_layer2 (128,128), fc_layer1
 * (64,128) from pu3_conv_tc_prepare_f32; code (2); w4 (3,64) / b4 (3) = fc_layer2; res (b,3,n) or NULL; y (b,3,2n) with
 * y[b,c,2p+j] = fc2(relu(fc1(relu(up2(relu(up1([x[b,:,p]; code[j]])))))))[c] + res[b,c,p].
 */
int pu3_head_tc_f32(int b, int n, int cin, const float *x, long long x_bstride, const void *ws1, const void *ws2,
                    const void *ws3, const float *w1, int w1_stride, int code_col, const float *b1, const float *code,
                    const float *b2, const float *b3, const float *w4, const float *b4, const float *res,
                    long long res_bstride, float *y, long long y_bstride, pu3_stream_t stream);
void pu3_head_tc_set_debug(void *buf); /* test hook: device buffer (24 x 512 u32) receiving CTA 0's event timeline; NULL = off */
void pu3_head_tc_set_mode(int mode);   /* A/B hook: 2 (default) = CTA pairs (tcgen05.mma.cta_group::2), operands in tensor memory; 1 = single CTAs, operands in tensor memory (TS form); 0 = operands in shared memory */

/*
 * Fused DenseEdgeConv forward for the reference configuration (24 input channels, growth 12, 3 layers):
 * replaces network/layers.py:22-64 (neighbour gather, edge feature [c, n-c], three 1x1 convolutions with dense
 * concatenation, max over the k edges).  x (b,24,n) slice (batch stride x_bstride), idx (b,n,idx_stride) i32 of which
 * entries [idx_off, idx_off+k) are the neighbours (idx_off = 1 drops rank 0, layers.py:34-35), weights as in
 * the state_dict (w0 (12,48), w1 (12,36), w2 (12,48)), y (b,60,n) slice = [max h2, max h1, max h0, x].
 */
int pu3_edgeconv_f32(int b, int n, int k, const float *x, long long x_bstride, const int32_t *idx, int idx_stride,
                     int idx_off, const float *w0, const float *b0, const float *w1, const float *b1,
                     const float *w2, const float *b2, float *y, long long y_bstride, pu3_stream_t stream);
/* The same function on the FFMA kernels only -- the arithmetic pu3_edgeconv_bwd_f32 recomputes.  The train-mode forward uses it so that
 * it returns exactly the function its backward differentiates (ReLU masks and arg-max edges included). */
int pu3_edgeconv_ffma_f32(int b, int n, int k, const float *x, long long x_bstride, const int32_t *idx, int idx_stride,
                          int idx_off, const float *w0, const float *b0, const float *w1, const float *b1,
                          const float *w2, const float *b2, float *y, long long y_bstride, pu3_stream_t stream);
void pu3_edgeconv_set_tc(int on);   /* test / A-B hook: 2 (default) = for k == 32 the two per-edge layers run on the tensor cores with every operand in tensor memory (tcgen05 TS form, 3xTF32; edgeconv_ts.cu), 1 = first version with layer-1 operand images in shared memory (edgeconv_tc.cu), 0 = FFMA kernels only; results agree to 1e-5 */

/*
 * Inter-level skip connection, fused (network/upsampler.py:317-347 and exponential_distance :232-250):
 * for each of t patches of n points, given the k nearest previous-level points idx (t,n,k) i64 (from
 * pu3_group_knn*), x (t,c,n) is updated in place:  x += 0.2 * sum_k w_k * prev_feat[idx_k],
 * w = ws*wf / sum_k(ws*wf + 1e-5), ws/wf the exponential spatial/feature weights with the per-patch bandwidths
 * h = mean_n(min_k d).  prev_xyz (clouds,3,no) channel-major; prev_feat_pm (clouds,no,c) POINT-major (see
 * pu3_to_point_major_f32); patch i reads cloud owner[i] (or i / p_div when owner is NULL).
 */
int pu3_skip_fuse_f32(int t, int n, int c, int k, int p_div, int no, float *x, const float *xyz, const int64_t *idx,
                      const float *prev_xyz, const float *prev_feat_pm, const int32_t *owner, pu3_stream_t stream);
/* ... the same, also writing the normalised weights w (t,n,k) (NULL = not wanted): what the backward pass needs. */
int pu3_skip_fuse_ex_f32(int t, int n, int c, int k, int p_div, int no, float *x, const float *xyz, const int64_t *idx,
                         const float *prev_xyz, const float *prev_feat_pm, const int32_t *owner, float *w_out, pu3_stream_t stream);
/* the same, also writing the updated features point-major, x_pm_out (t,n,c) (NULL = not wanted): what the NEXT level's skip
 * connection gathers from, produced while the rows are in registers instead of by pu3_to_point_major_f32 afterwards */
int pu3_skip_fuse_pm_f32(int t, int n, int c, int k, int p_div, int no, float *x, const float *xyz, const int64_t *idx,
                         const float *prev_xyz, const float *prev_feat_pm, const int32_t *owner, float *w_out, float *x_pm_out,
                         pu3_stream_t stream);
/*
 * Backward of the skip connection (autograd of network/upsampler.py:334-347 in the train step; the weights are detached there,
 * :243,249): dprev_feat_pm[cloud, idx[i,kk], :] += 0.2 * w[i,kk] * dx[:, i] with fp32 atomics into a caller-zeroed POINT-major
 * buffer (clouds,no,c); the gradient of x itself passes through unchanged.  dx (t,c,n) channel-major.
 */
int pu3_skip_bwd_f32(int t, int n, int c, int k, int p_div, int no, const float *dx, const int64_t *idx, const float *w,
                     const int32_t *owner, float *dprev_feat_pm, pu3_stream_t stream);
void pu3_skip_force_generic(int on); /* test hook: 1 = runtime-(k,c) kernel even for the k=5, c=264 configuration */

/*
 * Layout change for the features handed to the next level: in (t,c,n) channel-major ->
 * out[(slot[t]*n + i)*c + ch] = in[t][ch][i]  (slot NULL = identity), i.e. tiles placed side by side along the
 * point axis of a point-major (rows,c) buffer (the merge of network/upsampler.py:149-155 for the features).
 */
int pu3_to_point_major_f32(int t, int c, int n, const float *in, const int64_t *slot, float *out, pu3_stream_t stream);

/*
 * Train-step tail on one flat buffer (model.py:63-65 of the reference: clip_grad_value_(params, clip) followed by
 * Adam.step()): g = clamp(grad * grad_scale, -clip, clip) (clip <= 0: no clipping), then the Adam update with
 * torch.optim.Adam's formulas (no weight decay, no amsgrad); step = 1 for the first update.  grad_scale = 1/world
 * turns an all-reduce(sum) into the DDP mean.
 */
int pu3_clip_adam_f32(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                      float grad_scale, float clip, float lr, float beta1, float beta2, float eps, int step,
                      pu3_stream_t stream);

/*
 * Level engine: every kernel of one Level.forward (network/upsampler.py:272-374) enqueued by one call.
 * Weights in the reference's state_dict layouts; up1_w_feat is a contiguous (128,264)
 * copy of up_layer1's weight without its code column, code the (r) expansion code (upsampler.py:264-270).
 */
typedef struct {
    const float *layer0_w, *layer0_b;           /* (24,3), (24) */
    const float *ec_w[4][3], *ec_b[4][3];       /* layer{1..4}.mlps.{0,1,2}: (12,48) (12,36) (12,48) */
    const float *prep_w[3], *prep_b[3];         /* layer{2,3,4}_prep: (24,84) (24,144) (24,204) */
    const float *up1_w, *up1_w_feat, *up1_b;    /* (128,265), (128,264), (128) */
    const float *up2_w, *up2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
    const float *code;                          /* (r) */
    int r, knn, fm_knn, reserved;
} pu3_level_weights;

/*
 * t patches of n points.  xyz_norm (t,3,n) normalised input; xyz (t,3,n) un-normalised (only read by the skip
 * connection).  Previous level (optional, prev_xyz NULL = none): prev_xyz (clouds,3,no) channel-major,
 * prev_feat_pm (clouds,no,264) point-major, prev_n (clouds) valid sizes or NULL.  Patch i belongs to request
 * owner[i] (ragged batches, groups = number of requests) or, with owner NULL, to request i / (t/clouds) with
 * duplicate-penalty groups of max_group consecutive patches.  Outputs: feat (t,264,n) = [y4|y3|y2|y1|x0],
 * out_xyz (t,3,n*r) in the normalised frame.  workspace: pu3_level_workspace() bytes, 256-byte aligned.
 */
size_t pu3_level_workspace(int t, int n, int r, int knn, int fm_knn, int clouds, int no, int has_prev);
int pu3_level_forward_f32(const pu3_level_weights *w, int t, int n, const float *xyz, const float *xyz_norm,
                          const int32_t *owner, int groups, int max_group, const float *prev_xyz,
                          const float *prev_feat_pm, int clouds, int no, const int32_t *prev_n, float *feat,
                          float *out_xyz, void *workspace, size_t workspace_bytes, pu3_stream_t stream);
/* the same, also handing the level's features over point-major for the next level: feat_pm_out (t,n,264) or NULL */
int pu3_level_forward_pm_f32(const pu3_level_weights *w, int t, int n, const float *xyz, const float *xyz_norm,
                             const int32_t *owner, int groups, int max_group, const float *prev_xyz,
                             const float *prev_feat_pm, int clouds, int no, const int32_t *prev_n, float *feat,
                             float *out_xyz, float *feat_pm_out, void *workspace, size_t workspace_bytes, pu3_stream_t stream);
/*
 * Buffers the train-mode forward fills for the backward pass (all caller-allocated): h[blk] (t,24,n) the input of dense block blk
 * (layer0 output, then the three prep outputs after ReLU), idx[blk] (t,n,knn+1) i32 its neighbour lists (column 0 = the dropped
 * rank 0), skip_idx (t,n,fm_knn) i64 and skip_w (t,n,fm_knn) of the skip connection (NULL without a previous level), h1 / h2
 * (t,128,n*r) the activations after up_layer1 / up_layer2, feat_pre see below.
 */
typedef struct pu3_level_saved {
    float *h[4];
    int32_t *idx[4];
    int64_t *skip_idx;
    float *skip_w;
    float *h1, *h2;
    float *feat_pre; /* (t,264,n) copy of the features BEFORE the skip connection's in-place update (what the prep convolutions
                        read in the forward); NULL without a previous level (feat itself is then unchanged) */
} pu3_level_saved;
int pu3_level_forward_train_f32(const pu3_level_weights *w, int t, int n, const float *xyz, const float *xyz_norm,
                                const int32_t *owner, int groups, int max_group, const float *prev_xyz,
                                const float *prev_feat_pm, int clouds, int no, const int32_t *prev_n, float *feat,
                                float *out_xyz, void *workspace, size_t workspace_bytes, const pu3_level_saved *saved,
                                pu3_stream_t stream);
int pu3_iota_i32(int n, int32_t *out, pu3_stream_t stream); /* out[i] = i */
/* test hook (teacher forcing): neighbour lists the level engine uses instead of its own searches -- b0..b3 (t,n,knn+1) i32 for the
 * four dense blocks, skip (t,n,fm_knn) i64; NULL = search as usual.  Global, not thread-safe: tests only. */
void pu3_level_set_knn_override(const int32_t *b0, const int32_t *b1, const int32_t *b2, const int32_t *b3, const int64_t *skip);
void pu3_level_set_tc(int mode); /* test / A-B hook: 3 (default) = fused tcgen05 head (pu3_head_tc_f32) + prep convolutions on tcgen05, 2 = three-kernel tcgen05 head + prep convolutions, 1 = three-kernel head only, 0 = fp32 FFMA kernels */

/*
 * Weight / bias gradient of the 1x1 convolution: dw[co,ci] += sum_{b,p} dy[b,co,p] x[b,ci,p], db[co] += sum dy
 * (db may be NULL), accumulated into caller-zeroed buffers with fp32 atomics.  x, dy: channel slices like the
 * forward (pointer + batch stride, channel stride n).  The input gradient is pu3_pointwise_conv_f32 applied to dy
 * with the transposed weight.
 */
int pu3_pointwise_conv_bwd_w_f32(int b, int n, int cin, int cout, const float *x, long long x_bstride, const float *dy,
                                 long long dy_bstride, float *dw, float *db, pu3_stream_t stream);
/* ... with an explicit row stride of dw (>= cin): gradients of a column block of a wider weight matrix (the 264 feature columns
 * and the code column of up_layer1's (128,265) weight, upsampler.py:225,354-361). */
int pu3_pointwise_conv_bwd_w_ex_f32(int b, int n, int cin, int cout, const float *x, long long x_bstride, const float *dy,
                                    long long dy_bstride, float *dw, int dw_stride, float *db, pu3_stream_t stream);
/* out[e] (+)= sum_{j<r} in[e*r + j] for e < rows*n: gradient of the r-fold point replication of the expansion head (:356-358,:371). */
int pu3_replica_sum_f32(long long rows, int n, int r, const float *in, float *out, int accumulate, pu3_stream_t stream);
/* g[e] = act[e] > 0 ? g[e] : 0 (ReLU derivative applied to a gradient in place). */
int pu3_relu_mask_f32(long long total, float *g, const float *act, pu3_stream_t stream);

/*
 * DenseEdgeConv backward (autograd of network/layers.py:44-64 in the reference's train step), k <= 32:
 * given x, the neighbour indices and the weights of the forward call and dy (b,60,n), accumulates into the
 * caller-zeroed dx (b,24,n) and dw0 (12,48), db0, dw1 (12,36), db1, dw2 (12,48), db2.  The forward activations
 * are recomputed per edge; max() routes a channel's gradient to the first edge attaining the maximum.
 */
int pu3_edgeconv_bwd_f32(int b, int n, int k, const float *x, long long x_bstride, const int32_t *idx, int idx_stride,
                         int idx_off, const float *w0, const float *b0, const float *w1, const float *b1,
                         const float *w2, const float *b2, const float *dy, long long dy_bstride, float *dx,
                         long long dx_bstride, float *dw0, float *db0, float *dw1, float *db1, float *dw2, float *db2,
                         pu3_stream_t stream);

/*
 * The small steps between the big kernels of the eval path, as kernels (csrc/glue.cu), so that a whole Net.forward is a
 * fixed sequence of launches.  Arithmetic follows the operator order of the reference's torch expressions.
 *
 * normalize_point_batch (network/operations.py:12-30): centroid = mean over points, out = (pc - centroid) / max ||pc - centroid||.
 * pc / out (b,3,n) when nchw != 0, else (b,n,3); centroid (b,3), radius (b).
 */
int pu3_normalize_f32(int b, int n, int nchw, const float *pc, float *out, float *centroid, float *radius, pu3_stream_t stream);
/*
 * Outlier filter of eval-mode patch extraction (network/upsampler.py:63-76).  dist (b,n,dk) = distances to the dk >= 2 nearest
 * neighbours of every point of xyz (b,3,n) (pu3_group_knn_f32 with k = dk); a point is kept when dist[..,1] < 5 * mean_n(dist[..,1]).
 * Kept points first, order preserved (torch.masked_select), removed points behind them: out_cm (b,3,n) channel-major and out_pm
 * (b,n,3) point-major.  n_arr[i] = max(kept_i, min(k, n)); p_arr[i] = int(kept_i / k * 5) (:76, evaluated in double);
 * pk_arr[i] = p_arr[i]*k and pkr_arr[i] = p_arr[i]*k*r (either may be NULL): the valid sizes of the request's tiles side by side
 * before / after the r-fold upsampling; *bad |= 1 when some kept_i < k (the tile size itself would change: the caller redoes
 * that forward request by request).
 */
int pu3_outlier_compact_f32(int b, int n, int dk, int k, int r, const float *dist, const float *xyz, float *out_cm, float *out_pm,
                            int32_t *n_arr, int32_t *p_arr, int32_t *pk_arr, int32_t *pkr_arr, int32_t *bad, pu3_stream_t stream);
/* seeds[i,c,j] = xyz[i,c,idx[i, j < p_arr[i] ? j : 0]]: the gather after the seed FPS (:78, operations.py:320); tile slots past a
 * request's own count repeat its first tile.  xyz (b,3,n), idx (b,p) i32, seeds (b,3,p). */
int pu3_tile_seeds_f32(int b, int n, int p, const float *xyz, const int32_t *idx, const int32_t *p_arr, float *seeds,
                       pu3_stream_t stream);
/* tiles (b,3,p,k) (pu3_group_knn* neighbour output) -> patch (b*p,3,k) (torch.cat(torch.unbind(.,2),0), :85), patch_norm + centroid
 * (b*p,3) + radius (b*p) (normalize_point_batch, :138) and, unless NULL, side_by_side (b,3,p*k): the tiles of a request along the point
 * axis (:148-152), the cloud the next level's skip connection searches. */
int pu3_tiles_normalize_f32(int b, int p, int k, const float *tiles, float *patch, float *patch_norm, float *centroid,
                            float *radius, float *side_by_side, pu3_stream_t stream);
/* merged_pm[i, j*kr + q, c] = xyz_norm[i*p + j, c, q] * radius[i*p + j] + centroid[i*p + j, c]  (:144, :149-155), POINT-major (b,p*kr,3). */
int pu3_denorm_merge_f32(int b, int p, int kr, const float *xyz_norm, const float *centroid, const float *radius,
                         float *merged_pm, pu3_stream_t stream);
/* out[i,c,j] = pts[i, idx[i,j], c]: gather_points (operations.py:320) reading a point-major cloud pts (b,n,3); out (b,3,m). */
int pu3_gather_pm_f32(int b, int n, int m, const float *pts, const int32_t *idx, float *out, pu3_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PU3_B200_H */
