// layout.cu — the integer box-to-layout work that feeds the generator: box -> mask rasterisation and the shifted boxes.
// Replaces data/vg_custom_mask.py:120,136,157 (masks[i, :, round(y0*H):round(y1*H), round(x0*W):round(x1*W)] = 1 with
// Python's round — half-to-even on double — and Python slice semantics) and :139-158 (narrow boxes move 0.8x of the way to
// the farther horizontal border, arithmetic in double, stored as fp32).  Bit-exact contract.
#include "common.cuh"

namespace b200 {

// Python: round(v * size) on double (round half to even), then slice-bound normalisation of a[start:stop]:
// negative indices count from the end, everything is clamped to [0, size]
__device__ __forceinline__ int py_slice_bound(float v, int size) {
    const double r = rint((double)v * (double)size);       // rint: round half to even in the default rounding mode
    long long i = (long long)r;
    if (i < 0) i += size;
    if (i < 0) i = 0;
    if (i > size) i = size;
    return (int)i;
}

// masks (O, 1, H, W) fp32; one thread per 4 consecutive x (W % 4 == 0) or per pixel
template <int V>
__global__ void rasterize_boxes_kernel(const float* __restrict__ boxes, int O, int H, int W, float* __restrict__ masks) {
    const int Wv = W / V;
    const int64_t total = (int64_t)O * H * Wv;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int xv = (int)(t % Wv);
        const int y = (int)((t / Wv) % H);
        const int o = (int)(t / ((int64_t)Wv * H));
        const float* b = boxes + (int64_t)o * 4;
        const int xa = py_slice_bound(b[0], W), ya = py_slice_bound(b[1], H);
        const int xb = py_slice_bound(b[2], W), yb = py_slice_bound(b[3], H);
        const bool row = y >= ya && y < yb;
        float v[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const int x = xv * V + e;
            v[e] = (row && x >= xa && x < xb) ? 1.f : 0.f;
        }
        float* p = masks + ((int64_t)o * H + y) * W + xv * V;
        if (V == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        else p[0] = v[0];
    }
}

__global__ void shift_boxes_kernel(const float* __restrict__ boxes, int O, float* __restrict__ out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= O) return;
    double x0 = (double)boxes[o * 4 + 0], x1 = (double)boxes[o * 4 + 2];
    if (x1 - x0 < 0.5) {
        const double left = x0, right = 1.0 - x1;
        if (left > right) {
            const double s = left * 0.8;
            x0 = x0 - s;
            x1 = x1 - s;
        } else if (right > left) {
            const double s = right * 0.8;
            x0 = x0 + s;
            x1 = x1 + s;
        }
    }
    out[o * 4 + 0] = (float)x0;
    out[o * 4 + 1] = boxes[o * 4 + 1];
    out[o * 4 + 2] = (float)x1;
    out[o * 4 + 3] = boxes[o * 4 + 3];
}

}  // namespace b200

using namespace b200;

extern "C" int b200_rasterize_boxes(const float* boxes, int O, int H, int W, float* masks, b200_stream_t stream) {
    if (O == 0 || H == 0 || W == 0) return 0;
    const bool vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(masks) & 15) == 0;
    const int64_t total = (int64_t)O * H * (vec ? W / 4 : W);
    if (vec) rasterize_boxes_kernel<4><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(boxes, O, H, W, masks);
    else rasterize_boxes_kernel<1><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(boxes, O, H, W, masks);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_shift_boxes(const float* boxes, int O, float* out, b200_stream_t stream) {
    if (O == 0) return 0;
    shift_boxes_kernel<<<(O + 127) / 128, 128, 0, as_stream(stream)>>>(boxes, O, out);
    B200_CHECK_LAUNCH();
    return 0;
}
