// crop.cu — box crops (K8) and their deterministic backward (K9).
// Replaces models/bilinear.py:26-41,67-104,107-136 (crop_bbox_batch -> crop_bbox -> F.grid_sample, bilinear,
// zeros padding, align_corners=False).  Index arithmetic reproduces the reference's fp32 operation order without
// FMA contraction so floor(ix), floor(iy) are bit-exact.
#include "common.cuh"

namespace b200 {

// 2*bbox-1 (bilinear.py:127); sw*start + ew*end (bilinear.py:280); ATen grid_sampler_unnormalize, align_corners=False:
// ((coord + 1) * size - 1) / 2.
__device__ __forceinline__ float crop_coord(float lo, float hi, float sw, float ew, int size) {
    float a = __fadd_rn(__fmul_rn(2.f, lo), -1.f);
    float b = __fadd_rn(__fmul_rn(2.f, hi), -1.f);
    float g = __fadd_rn(__fmul_rn(sw, a), __fmul_rn(ew, b));
    return __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), -1.f), 0.5f);
}

__global__ void crop_taps_kernel(const float* __restrict__ boxes, const float* __restrict__ wx,
                                 const float* __restrict__ wy, int32_t* ix0, int32_t* iy0, float* fx, float* fy, int H,
                                 int W, int B, int HH, int WW) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < B * WW) {
        int b = t / WW, j = t % WW;
        float ix = crop_coord(boxes[b * 4 + 0], boxes[b * 4 + 2], wx[j], wx[WW + j], W);
        float f = floorf(ix);
        ix0[t] = (int)f;
        fx[t] = ix - f;
    }
    if (t < B * HH) {
        int b = t / HH, i = t % HH;
        float iy = crop_coord(boxes[b * 4 + 1], boxes[b * 4 + 3], wy[i], wy[HH + i], H);
        float f = floorf(iy);
        iy0[t] = (int)f;
        fy[t] = iy - f;
    }
}

// one thread per (b, i, j); channels looped (C = 3 on the path). Lanes run along j: coalesced NCHW stores.
__global__ void crop_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ boxes,
                                const int32_t* __restrict__ box_to_img, const float* __restrict__ wx,
                                const float* __restrict__ wy, float* __restrict__ crops, int C, int H, int W, int B,
                                int HH, int WW) {
    int64_t total = (int64_t)B * HH * WW;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(t % WW);
        int i = (int)((t / WW) % HH);
        int b = (int)(t / ((int64_t)WW * HH));
        const float* bx = boxes + b * 4;
        float ix = crop_coord(bx[0], bx[2], wx[j], wx[WW + j], W);
        float iy = crop_coord(bx[1], bx[3], wy[i], wy[HH + i], H);
        float x0f = floorf(ix), y0f = floorf(iy);
        int x0 = (int)x0f, y0 = (int)y0f;
        float wx1 = ix - x0f, wy1 = iy - y0f;
        float wx0 = (x0f + 1.f) - ix, wy0 = (y0f + 1.f) - iy;
        bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
        bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
        const float* img = feats + (int64_t)box_to_img[b] * C * H * W;
        for (int c = 0; c < C; ++c) {
            const float* p = img + (int64_t)c * H * W;
            float v = 0.f;
            if (vy0 && vx0) v += p[y0 * W + x0] * (wx0 * wy0);
            if (vy0 && vx1) v += p[y0 * W + x0 + 1] * (wx1 * wy0);
            if (vy1 && vx0) v += p[(y0 + 1) * W + x0] * (wx0 * wy1);
            if (vy1 && vx1) v += p[(y0 + 1) * W + x0 + 1] * (wx1 * wy1);
            crops[(((int64_t)b * C + c) * HH + i) * WW + j] = v;
        }
    }
}

// The sampling coordinate of crop row i (column j) is monotone in i: only the few rows whose floor lands on y - 1 or y
// contribute to image row y.  [lo, hi) is a conservative superset of them from the linear end points (one extra index
// either side, the whole range for degenerate / non-increasing boxes); the exact floor predicate below still decides,
// so the contributing set and its ascending summation order are those of the full scan.
__device__ __forceinline__ void tap_range(float c_first, float c_last, int S, int target, int& lo, int& hi) {
    lo = 0;
    hi = S;
    const float step = S > 1 ? (c_last - c_first) / (float)(S - 1) : 0.f;
    if (step > 1e-3f) {
        const float a = ((float)target - 1.f - c_first) / step, b = ((float)target + 1.f - c_first) / step;
        const int l = (int)floorf(a) - 1, h = (int)ceilf(b) + 2;
        lo = l < 0 ? 0 : (l > S ? S : l);
        hi = h < 0 ? 0 : (h > S ? S : h);
    }
}

// pass 1: T[b,c,y,j] = sum_i wy(i -> y) * dcrops[b,c,i,j]   (rows of the crop that touch image row y, ascending i)
__global__ void crop_bwd_rows_kernel(const float* __restrict__ dcrops, const float* __restrict__ boxes,
                                     const float* __restrict__ wy, float* __restrict__ T, int C, int H, int B, int HH,
                                     int WW) {
    int64_t total = (int64_t)B * C * H * WW;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(t % WW);
        int y = (int)((t / WW) % H);
        int c = (int)((t / ((int64_t)WW * H)) % C);
        int b = (int)(t / ((int64_t)WW * H * C));
        const float* bx = boxes + b * 4;
        const float* d = dcrops + ((int64_t)b * C + c) * HH * WW + j;
        int lo, hi;
        tap_range(crop_coord(bx[1], bx[3], wy[0], wy[HH], H), crop_coord(bx[1], bx[3], wy[HH - 1], wy[2 * HH - 1], H), HH, y,
                  lo, hi);
        float acc = 0.f;
        for (int i = lo; i < hi; ++i) {
            float iy = crop_coord(bx[1], bx[3], wy[i], wy[HH + i], H);
            float y0f = floorf(iy);
            int y0 = (int)y0f;
            if (y0 == y) acc += ((y0f + 1.f) - iy) * d[(int64_t)i * WW];
            else if (y0 + 1 == y) acc += (iy - y0f) * d[(int64_t)i * WW];
        }
        T[t] = acc;
    }
}

// pass 2: dfeats[n,c,y,x] = sum_{b in image n, ascending} sum_j wx(j -> x) * T[b,c,y,j]
__global__ void crop_bwd_cols_kernel(const float* __restrict__ T, const float* __restrict__ boxes,
                                     const int32_t* __restrict__ img_box_start, const int32_t* __restrict__ box_order,
                                     const float* __restrict__ wx, float* __restrict__ dfeats, int N, int C, int H,
                                     int W, int WW) {
    int64_t total = (int64_t)N * C * H * W;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int x = (int)(t % W);
        int y = (int)((t / W) % H);
        int c = (int)((t / ((int64_t)W * H)) % C);
        int n = (int)(t / ((int64_t)W * H * C));
        float acc = 0.f;
        for (int k = img_box_start[n]; k < img_box_start[n + 1]; ++k) {
            int b = box_order[k];
            const float* bx = boxes + b * 4;
            const float* row = T + (((int64_t)b * C + c) * H + y) * WW;
            int lo, hi;
            tap_range(crop_coord(bx[0], bx[2], wx[0], wx[WW], W), crop_coord(bx[0], bx[2], wx[WW - 1], wx[2 * WW - 1], W), WW,
                      x, lo, hi);
            for (int j = lo; j < hi; ++j) {
                float ix = crop_coord(bx[0], bx[2], wx[j], wx[WW + j], W);
                float x0f = floorf(ix);
                int x0 = (int)x0f;
                if (x0 == x) acc += ((x0f + 1.f) - ix) * row[j];
                else if (x0 + 1 == x) acc += (ix - x0f) * row[j];
            }
        }
        dfeats[t] = acc;
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_crop_fwd(const float* feats, const float* boxes, const int32_t* box_to_img, const float* wx,
                             const float* wy, float* crops, int N, int C, int H, int W, int B, int HH, int WW,
                             b200_stream_t stream) {
    (void)N;
    if (B == 0) return 0;
    int64_t total = (int64_t)B * HH * WW;
    crop_fwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(feats, boxes, box_to_img, wx, wy, crops, C, H,
                                                                         W, B, HH, WW);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_crop_taps(const float* boxes, const float* wx, const float* wy, int32_t* ix0, int32_t* iy0,
                              float* fx, float* fy, int H, int W, int B, int HH, int WW, b200_stream_t stream) {
    if (B == 0) return 0;
    int n = B * (HH > WW ? HH : WW);
    crop_taps_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(boxes, wx, wy, ix0, iy0, fx, fy, H, W, B, HH, WW);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_crop_bwd(const float* dcrops, const float* boxes, const int32_t* img_box_start,
                             const int32_t* box_order, const float* wx, const float* wy, float* dfeats, float* ws,
                             int N, int C, int H, int W, int B, int HH, int WW, b200_stream_t stream) {
    if (B > 0) {
        int64_t t1 = (int64_t)B * C * H * WW;
        crop_bwd_rows_kernel<<<grid_for(t1, 256), 256, 0, as_stream(stream)>>>(dcrops, boxes, wy, ws, C, H, B, HH, WW);
        B200_CHECK_LAUNCH();
    }
    int64_t t2 = (int64_t)N * C * H * W;
    crop_bwd_cols_kernel<<<grid_for(t2, 256), 256, 0, as_stream(stream)>>>(ws, boxes, img_box_start, box_order, wx,
                                                                           dfeats, N, C, H, W, WW);
    B200_CHECK_LAUNCH();
    return 0;
}
