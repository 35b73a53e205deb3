// Train-step tail on one flat fp32 buffer: clip_grad_value_ + Adam in a single pass, sm_100a.
//
// Replaces, for the 304 108 parameters of the 4-level network (one 1.2 MB buffer), the sequence of
// model.py:63-65 of the reference: torch.nn.utils.clip_grad_value_(params, 1) then optimizer.step()
// (torch.optim.Adam, lr 5e-4, betas (0.9, 0.999), eps 1e-8, no weight decay; model.py:21-23), which runs as
// 2 x 160 tiny per-tensor kernels (or a multi-tensor foreach).  With DDP the gradient arrives already averaged
// in the same flat buffer (one NCCL all-reduce), optionally still to be scaled by 1/world.
// HBM-bound elementwise: 16 B read + 12 B written per parameter.
#include "pu3_common.cuh"

namespace pu3 {

__global__ void __launch_bounds__(256) clip_adam_kernel(long long n, float *__restrict__ p, const float *__restrict__ g,
                                                       float *__restrict__ m, float *__restrict__ v, float grad_scale,
                                                       float clip, float lr_over_bc1, float beta1, float beta2,
                                                       float inv_sqrt_bc2, float eps) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * grad_scale;
        if (clip > 0.f) gi = fminf(fmaxf(gi, -clip), clip);              // clip_grad_value_
        const float mi = __fmaf_rn(gi - m[i], 1.0f - beta1, m[i]);       // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = __fmaf_rn(gi * gi, 1.0f - beta2, v[i] * beta2); // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;              // (sqrt(v) / sqrt(bias_correction2)) + eps
        m[i] = mi;
        v[i] = vi;
        p[i] = p[i] - lr_over_bc1 * (mi / denom);                        // param.addcdiv_(exp_avg, denom, value=-lr/bias_correction1)
    }
}

}  // namespace pu3

using namespace pu3;

extern "C" int pu3_clip_adam_f32(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                                 float grad_scale, float clip, float lr, float beta1, float beta2, float eps, int step,
                                 pu3_stream_t stream) {
    PU3_ARG_CHECK(n >= 0 && step >= 1, "clip_adam: n=%lld step=%d", n, step);
    if (n == 0) return PU3_OK;
    PU3_ARG_CHECK(param && grad && exp_avg && exp_avg_sq, "clip_adam: null pointer");
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    const long long blocks = (n + 255) / 256;
    const long long cap = (long long)device_info().sm_count * 8;
    clip_adam_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, as_stream(stream)>>>(
        n, param, grad, exp_avg, exp_avg_sq, grad_scale, clip, (float)(lr / bc1), beta1, beta2, (float)(1.0 / sqrt(bc2)), eps);
    PU3_LAUNCH_CHECK("clip_adam_kernel");
    return PU3_OK;
}
