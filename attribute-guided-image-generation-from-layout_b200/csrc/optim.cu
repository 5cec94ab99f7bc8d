// optim.cu — multi-tensor Adam: every parameter of an optimizer (torch.optim.Adam(params, lr, betas=(0.5, 0.999)),
// train64.py:111-114) updated by ONE launch.  The step counter lives on the device, so the update is CUDA-graph capturable.
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// (torch's non-amsgrad, weight_decay = 0 formulation, same operation order).  HBM bound: 4 reads + 3 writes of 4 bytes per
// parameter.
#include "common.cuh"

namespace b200 {

// one work item = one chunk of one tensor; items are laid out by the host (b200_adam_entry) and walked by blockIdx.x
__global__ void __launch_bounds__(256) adam_multi_kernel(const b200_adam_entry* __restrict__ entries, int n_entries,
                                                        const float* __restrict__ step, double lr_d, double beta1_d,
                                                        double beta2_d, double eps_d) {
    // scalar preparation in double, as torch.optim.Adam does on the host (1 - beta2 is not representable from an fp32 beta2)
    const double t = (double)*step;
    const double bc1 = 1.0 - pow(beta1_d, t), bc2 = 1.0 - pow(beta2_d, t);
    const float step_size = (float)(lr_d / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    const float beta2 = (float)beta2_d, omb1 = (float)(1.0 - beta1_d), omb2 = (float)(1.0 - beta2_d), eps = (float)eps_d;
    for (int e = blockIdx.x; e < n_entries; e += gridDim.x) {
        const b200_adam_entry en = entries[e];
        float* __restrict__ p = en.param;
        const float* __restrict__ g = en.grad;
        float* __restrict__ m = en.exp_avg;
        float* __restrict__ v = en.exp_avg_sq;
        const int n = en.n;
        if ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0) {
            const int n4 = n >> 2;
            for (int i = threadIdx.x; i < n4; i += 256) {
                float4 pv = reinterpret_cast<float4*>(p)[i];
                const float4 gv = reinterpret_cast<const float4*>(g)[i];
                float4 mv = reinterpret_cast<float4*>(m)[i];
                float4 vv = reinterpret_cast<float4*>(v)[i];
                float* pp = &pv.x; const float* gg = &gv.x; float* mm = &mv.x; float* vq = &vv.x;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    mm[k] = mm[k] + (gg[k] - mm[k]) * omb1;                         // lerp, as torch's exp_avg.lerp_(grad, 1 - beta1)
                    vq[k] = vq[k] * beta2 + omb2 * (gg[k] * gg[k]);
                    const float denom = sqrtf(vq[k]) / bc2_sqrt + eps;
                    pp[k] = pp[k] - step_size * (mm[k] / denom);
                }
                reinterpret_cast<float4*>(p)[i] = pv;
                reinterpret_cast<float4*>(m)[i] = mv;
                reinterpret_cast<float4*>(v)[i] = vv;
            }
            for (int i = (n4 << 2) + threadIdx.x; i < n; i += 256) {
                const float gk = g[i];
                const float mk = m[i] + (gk - m[i]) * omb1;
                const float vk = v[i] * beta2 + omb2 * (gk * gk);
                m[i] = mk; v[i] = vk;
                p[i] = p[i] - step_size * (mk / (sqrtf(vk) / bc2_sqrt + eps));
            }
        } else {
            for (int i = threadIdx.x; i < n; i += 256) {
                const float gk = g[i];
                const float mk = m[i] + (gk - m[i]) * omb1;
                const float vk = v[i] * beta2 + omb2 * (gk * gk);
                m[i] = mk; v[i] = vk;
                p[i] = p[i] - step_size * (mk / (sqrtf(vk) / bc2_sqrt + eps));
            }
        }
    }
}

__global__ void adam_step_inc_kernel(float* step) { *step += 1.f; }

}  // namespace b200

using namespace b200;

extern "C" int b200_adam_multi(const b200_adam_entry* entries_dev, int n_entries, float* step_dev, double lr, double beta1,
                               double beta2, double eps, b200_stream_t stream) {
    if (n_entries <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    adam_step_inc_kernel<<<1, 1, 0, st>>>(step_dev);
    B200_CHECK_LAUNCH();
    const int grid = n_entries < kNumSMs * 8 ? n_entries : kNumSMs * 8;
    adam_multi_kernel<<<grid, 256, 0, st>>>(entries_dev, n_entries, step_dev, lr, beta1, beta2, eps);
    B200_CHECK_LAUNCH();
    return 0;
}
