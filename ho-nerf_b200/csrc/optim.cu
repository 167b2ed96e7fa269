// Adam over one flat fp32 parameter buffer (torch.optim.Adam semantics, exp_runner.py:83 builds one Adam over every
// network's parameters): one launch per step instead of a multi-tensor sweep over ~90 small tensors.
// The step count AND the learning rate live in device memory so the launch is identical every step (CUDA-graph replay)
// while the caller's schedule (exp_runner.py:update_learning_rate writes param_groups[i]['lr'] every iteration) still
// takes effect: the host updates the device scalar outside the graph.  A parameter that received no gradient in some
// earlier steps has a smaller step count than the others (torch keeps `step` per parameter): `skipped` holds, per
// 32-element block of the flat buffer, how many steps its parameter sat out.
#include <algorithm>

#include "common.cuh"

namespace hn {

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, const float* __restrict__ step,
                                                        const float* __restrict__ lr_dev, const float* __restrict__ skipped,
                                                        float lr, float beta1, float beta2, float omb1, float omb2, float eps,
                                                        float weight_decay, float grad_scale) {
    const float t0 = *step;
    if (lr_dev) lr = *lr_dev;
    float t_cur = t0;
    float bc1 = 1.0f - powf(beta1, t0), bc2 = 1.0f - powf(beta2, t0);
    float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (skipped) {
            const float t = t0 - skipped[i >> 5];
            if (t != t_cur) {
                t_cur = t;
                bc1 = 1.0f - powf(beta1, t);
                bc2 = 1.0f - powf(beta2, t);
                step_size = lr / bc1;
                inv_sqrt_bc2 = rsqrtf(bc2);
            }
        }
        float gi = g[i] * grad_scale;
        const float pi = p[i];
        if (weight_decay != 0.0f) gi += weight_decay * pi;
        const float mi = beta1 * m[i] + omb1 * gi;
        const float vi = beta2 * v[i] + omb2 * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}

}  // namespace hn

using namespace hn;

extern "C" int hn_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* step, const float* lr_dev,
                            const float* skipped, double lr, double beta1, double beta2, double eps, double weight_decay,
                            double grad_scale, hn_stream_t stream) {
    HN_REQUIRE(p && g && m && v && step && n >= 0, "hn_adam_flat: null argument");
    if (n == 0) return HN_OK;
    const int grid = (int)std::min<int64_t>(ceil_div(n, 256), 4 * 148);
    // 1 - beta in double, then rounded once (torch.optim.Adam does the same: 1 - 0.999 evaluated in fp32 is off by 1.3e-5)
    adam_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, step, lr_dev, skipped, (float)lr, (float)beta1, (float)beta2,
                                                             (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
                                                             (float)weight_decay, (float)grad_scale);
    count_launch();
    HN_CHECK_LAUNCH();
    return HN_OK;
}
