// Fused Adam step over one flat FP32 parameter buffer (train.py:183-187,263-265).
//
// torch.optim.Adam(params, lr, weight_decay) as the reference configures it (betas 0.9 / 0.999, eps 1e-8, no amsgrad),
// with the learning rate of torch.optim.lr_scheduler.ExponentialLR passed in by the host (lr0 * gamma^epoch):
//   g  = grad * grad_scale + weight_decay * p          (coupled L2, as torch.optim.Adam)
//   m  = m + (1 - beta1) (g - m)                       (lerp_)
//   v  = beta2 v + (1 - beta2) g^2
//   p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// One launch for the whole model (1.6 M parameters at L=128): vectorised, grid-stride, HBM bound (28 B per parameter).
// `grad_scale` folds the 1/world of an averaged all-reduce (or a loss scale) into the same pass.
#include "common.cuh"

namespace cgnn {
namespace {

// `live` (nullable): 0 where the parameter got no gradient this step -- torch.optim.Adam skips such parameters
// entirely (no weight decay, no moment update), and so does this kernel.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, const float* __restrict__ live, int64_t n, float step_size,
                                                   float beta1, float beta2, float eps, float weight_decay, float inv_bc2_sqrt,
                                                   float grad_scale) {
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 lv = make_float4(1.f, 1.f, 1.f, 1.f);
        if (live != nullptr) {
            lv = reinterpret_cast<const float4*>(live)[i];
            if (lv.x == 0.f && lv.y == 0.f && lv.z == 0.f && lv.w == 0.f) continue;
        }
        const float* le = &lv.x;
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        float* pe = &pp.x; const float* ge = &gg.x; float* me = &mm.x; float* ve = &vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (le[j] == 0.f) continue;
            const float gr = fmaf(weight_decay, pe[j], ge[j] * grad_scale);
            me[j] = fmaf(1.0f - beta1, gr - me[j], me[j]);
            ve[j] = fmaf(1.0f - beta2, gr * gr, beta2 * ve[j]);
            pe[j] -= step_size * (me[j] / (sqrtf(ve[j]) * inv_bc2_sqrt + eps));
        }
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    // tail (n not a multiple of 4)
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (live != nullptr && live[i] == 0.f) continue;
        const float gr = fmaf(weight_decay, p[i], g[i] * grad_scale);
        const float mi = fmaf(1.0f - beta1, gr - m[i], m[i]);
        const float vi = fmaf(1.0f - beta2, gr * gr, beta2 * v[i]);
        m[i] = mi; v[i] = vi;
        p[i] -= step_size * (mi / (sqrtf(vi) * inv_bc2_sqrt + eps));
    }
}

}  // namespace
}  // namespace cgnn

using namespace cgnn;

extern "C" int cgnn_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int32_t step, float grad_scale,
                              cgnn_stream stream_) {
    return cgnn_adam_step_masked(params, grads, exp_avg, exp_avg_sq, nullptr, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, stream_);
}

extern "C" int cgnn_adam_step_masked(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, const float* live,
                                     int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                                     float grad_scale, cgnn_stream stream_) {
    CGNN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && n >= 1, "cgnn_adam_step: bad arguments");
    CGNN_CHECK_ARG((reinterpret_cast<uintptr_t>(live) & 15) == 0, "cgnn_adam_step: the mask must be 16-byte aligned");
    CGNN_CHECK_ARG(step >= 1 && beta1 >= 0.0f && beta1 < 1.0f && beta2 >= 0.0f && beta2 < 1.0f, "cgnn_adam_step: step >= 1 and betas in [0, 1)");
    CGNN_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                     reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "cgnn_adam_step: buffers must be 16-byte aligned");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    adam_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(params, grads, exp_avg, exp_avg_sq, live, n, step_size, beta1, beta2, eps,
                                                                      weight_decay, inv_bc2_sqrt, grad_scale);
    CGNN_LAUNCH_CHECK();
    return CGNN_OK;
}
