// data.cu — the device side of the reference's data contract either side of the training step:
//  * one-hot attribute vectors from the loader's -1 terminated index lists (data/vg_custom_mask.py:160-171),
//  * imagenet_deprocess_batch (data/utils.py:32-66): un-normalise, per-image min/max rescale, *255, clamp, uint8.
// Both are bit-exact restatements of fp32 / integer arithmetic (no FMA contraction, IEEE division).
#include "common.cuh"

namespace b200 {

// one thread per object: ones at att[o, 0 .. k) where k = index of the first -1 (vg_custom_mask.py:162-168)
__global__ void one_hot_attributes_kernel(const int64_t* __restrict__ att, int O, int A, int n_att, float* __restrict__ out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    float* row = out + (int64_t)o * n_att;
    for (int k = 0; k < n_att; ++k) row[k] = 0.f;
    for (int k = 0; k < A; ++k) {
        const int64_t a = att[(int64_t)o * A + k];
        if (a == -1) break;
        if (a >= 0 && a < n_att) row[a] = 1.f;
    }
}

// T.Normalize(mean=0, std=1/s) then T.Normalize(mean=-m, std=1):  (x / inv_std) - (-m) in fp32, as torchvision evaluates it
__device__ __forceinline__ float denorm(float x, float inv_std, float neg_mean) {
    return __fdiv_rn(__fsub_rn(__fdiv_rn(__fsub_rn(x, 0.f), inv_std), neg_mean), 1.f);
}

// pass 1: per-image min / max of the un-normalised values (min / max are order independent: bit-exact by construction)
__global__ void __launch_bounds__(256) deprocess_minmax_kernel(const float* __restrict__ imgs, int C, int HW,
                                                              const float* __restrict__ inv_std,
                                                              const float* __restrict__ neg_mean, float* __restrict__ mm) {
    __shared__ float slo[8], shi[8];
    const int n = blockIdx.x;
    const float* p = imgs + (int64_t)n * C * HW;
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < C * HW; i += 256) {
        const int c = i / HW;
        const float v = denorm(p[i], inv_std[c], neg_mean[c]);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
        mm[2 * n] = lo;
        mm[2 * n + 1] = hi;
    }
}

// pass 2: ((v - lo) / (hi - lo)) * 255, clamp to [0, 255], truncate to uint8 (Tensor.byte())
__global__ void deprocess_apply_kernel(const float* __restrict__ imgs, int N, int C, int HW,
                                       const float* __restrict__ inv_std, const float* __restrict__ neg_mean,
                                       const float* __restrict__ mm, int rescale, uint8_t* __restrict__ out) {
    const int64_t total = (int64_t)N * C * HW;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(t / ((int64_t)C * HW));
        const int c = (int)((t / HW) % C);
        float v = denorm(imgs[t], inv_std[c], neg_mean[c]);
        if (rescale) {
            const float lo = mm[2 * n], hi = mm[2 * n + 1];
            v = __fdiv_rn(__fsub_rn(v, lo), __fsub_rn(hi, lo));
        }
        v = __fmul_rn(v, 255.f);
        v = fminf(fmaxf(v, 0.f), 255.f);          // NaN (hi == lo) clamps like torch.clamp: stays NaN -> byte() gives 0
        out[t] = (v == v) ? (uint8_t)v : (uint8_t)0;
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_one_hot_attributes(const int64_t* att_idx, int O, int A, int n_att, float* out, b200_stream_t stream) {
    if (O == 0) return 0;
    one_hot_attributes_kernel<<<(O + 127) / 128, 128, 0, as_stream(stream)>>>(att_idx, O, A, n_att, out);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_imagenet_deprocess(const float* imgs, int N, int C, int HW, const float* inv_std, const float* neg_mean,
                                       int rescale, uint8_t* out, float* ws, b200_stream_t stream) {
    if (N == 0 || C == 0 || HW == 0) return 0;
    B200_REQUIRE((int64_t)C * HW < (1ll << 31), "imagenet_deprocess: image too large");
    if (rescale) {
        deprocess_minmax_kernel<<<N, 256, 0, as_stream(stream)>>>(imgs, C, HW, inv_std, neg_mean, ws);
        B200_CHECK_LAUNCH();
    }
    const int64_t total = (int64_t)N * C * HW;
    deprocess_apply_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(imgs, N, C, HW, inv_std, neg_mean, ws, rescale,
                                                                                out);
    B200_CHECK_LAUNCH();
    return 0;
}
