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
    const uint32_t total = (uint32_t)B * HH * WW;               // < 2^31, checked by the launcher
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const uint32_t q = t / (uint32_t)WW;
        const int j = (int)(t - q * (uint32_t)WW);
        const int i = (int)(q % (uint32_t)HH);
        const int b = (int)(q / (uint32_t)HH);
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

// pixel footprint of box b: rows / columns of the image any bilinear tap of the crop can touch (one pixel of slack),
// ext[b] = {ylo, yhi, xlo, xhi} inclusive, clamped to the image.  Pixels outside it receive nothing from the box.
__global__ void crop_extents_kernel(const float* __restrict__ boxes, const float* __restrict__ wx,
                                    const float* __restrict__ wy, int4* __restrict__ ext, int B, int H, int W, int HH,
                                    int WW) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* bx = boxes + b * 4;
    const float ya = crop_coord(bx[1], bx[3], wy[0], wy[HH], H), yb = crop_coord(bx[1], bx[3], wy[HH - 1], wy[2 * HH - 1], H);
    const float xa = crop_coord(bx[0], bx[2], wx[0], wx[WW], W), xb = crop_coord(bx[0], bx[2], wx[WW - 1], wx[2 * WW - 1], W);
    int ylo = (int)floorf(fminf(ya, yb)) - 1, yhi = (int)floorf(fmaxf(ya, yb)) + 2;
    int xlo = (int)floorf(fminf(xa, xb)) - 1, xhi = (int)floorf(fmaxf(xa, xb)) + 2;
    ylo = ylo < 0 ? 0 : ylo; xlo = xlo < 0 ? 0 : xlo;
    yhi = yhi > H - 1 ? H - 1 : yhi; xhi = xhi > W - 1 ? W - 1 : xhi;
    ext[b] = make_int4(ylo, yhi, xlo, xhi);
}

// pass 1: T[b,c,y,j] = sum_i wy(i -> y) * dcrops[b,c,i,j]   (rows of the crop that touch image row y, ascending i);
// one block per (box, channel) walks only the rows of the box's footprint (pass 2 reads no others)
__global__ void crop_bwd_rows_kernel(const float* __restrict__ dcrops, const float* __restrict__ boxes,
                                     const float* __restrict__ wy, const int4* __restrict__ ext, float* __restrict__ T,
                                     int C, int H, int B, int HH, int WW) {
    const int b = blockIdx.x / C, c = blockIdx.x - b * C;
    (void)B;
    const int4 e = ext[b];
    const float* bx = boxes + b * 4;
    const float c_first = crop_coord(bx[1], bx[3], wy[0], wy[HH], H);
    const float c_last = crop_coord(bx[1], bx[3], wy[HH - 1], wy[2 * HH - 1], H);
    const float* dbase = dcrops + ((int64_t)b * C + c) * HH * WW;
    float* Tb = T + ((int64_t)b * C + c) * H * WW;
    const int n = (e.y - e.x + 1) * WW;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int r = idx / WW, j = idx - r * WW;
        const int y = e.x + r;
        const float* d = dbase + j;
        int lo, hi;
        tap_range(c_first, c_last, HH, y, lo, hi);
        float acc = 0.f;
        for (int i = lo; i < hi; ++i) {
            float iy = crop_coord(bx[1], bx[3], wy[i], wy[HH + i], H);
            float y0f = floorf(iy);
            int y0 = (int)y0f;
            if (y0 == y) acc += ((y0f + 1.f) - iy) * d[(int64_t)i * WW];
            else if (y0 + 1 == y) acc += (iy - y0f) * d[(int64_t)i * WW];
        }
        Tb[(int64_t)y * WW + j] = acc;
    }
}

// pass 2: dfeats[n,c,y,x] = sum_{b in image n, ascending} sum_j wx(j -> x) * T[b,c,y,j]
__global__ void crop_bwd_cols_kernel(const float* __restrict__ T, const float* __restrict__ boxes,
                                     const int32_t* __restrict__ img_box_start, const int32_t* __restrict__ box_order,
                                     const float* __restrict__ wx, const int4* __restrict__ ext,
                                     float* __restrict__ dfeats, int N, int C, int H, int W, int WW) {
    const uint32_t total = (uint32_t)N * C * H * W;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        uint32_t q = t / (uint32_t)W;
        const int x = (int)(t - q * (uint32_t)W);
        const int y = (int)(q % (uint32_t)H);
        q /= (uint32_t)H;
        const int c = (int)(q % (uint32_t)C);
        const int n = (int)(q / (uint32_t)C);
        float acc = 0.f;
        for (int k = img_box_start[n]; k < img_box_start[n + 1]; ++k) {
            int b = box_order[k];
            const int4 e = ext[b];
            if (y < e.x || y > e.y || x < e.z || x > e.w) continue;      // outside the box's footprint
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

// ---------------------------------------------------------------------------------------------------------------------
// Shared-memory staged variants (the default): one block per (box, channel).
// Forward: the box's source coordinates are computed ONCE per block (HH + WW evaluations instead of two per output pixel),
// the pixel footprint of the box in the image plane is staged in shared memory with coalesced row reads, and every output
// pixel takes its four taps from there; stores are coalesced along j.  Arithmetic (coordinates, floors, weights, tap order)
// is exactly that of crop_fwd_kernel, so the results are bit-identical to it.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kCropMaxS = 128;                 // crop side the staged kernels take
constexpr int kCropFootFloats = 8192;          // 32 KB footprint tile (e.g. 90 x 90 pixels); larger boxes read global memory

__global__ void __launch_bounds__(256) crop_fwd_staged_kernel(const float* __restrict__ feats, const float* __restrict__ boxes,
                                                             const int32_t* __restrict__ box_to_img,
                                                             const float* __restrict__ wx, const float* __restrict__ wy,
                                                             float* __restrict__ crops, int C, int H, int W, int HH, int WW) {
    __shared__ float sx[kCropMaxS], sy[kCropMaxS];           // source coordinates per crop column / row
    __shared__ float foot[kCropFootFloats];
    __shared__ int ext[4];
    const int b = blockIdx.x / C, c = blockIdx.x - b * C;
    const float* bx = boxes + b * 4;
    for (int t = threadIdx.x; t < WW + HH; t += 256) {
        if (t < WW) sx[t] = crop_coord(bx[0], bx[2], wx[t], wx[WW + t], W);
        else sy[t - WW] = crop_coord(bx[1], bx[3], wy[t - WW], wy[HH + t - WW], H);
    }
    __syncthreads();
    if (threadIdx.x < 32) {                                    // footprint = [min floor, max floor + 1] clamped to the plane
        float ylo = INFINITY, yhi = -INFINITY, xlo = INFINITY, xhi = -INFINITY;
        for (int i = threadIdx.x; i < HH; i += 32) { const float f = floorf(sy[i]); ylo = fminf(ylo, f); yhi = fmaxf(yhi, f); }
        for (int j = threadIdx.x; j < WW; j += 32) { const float f = floorf(sx[j]); xlo = fminf(xlo, f); xhi = fmaxf(xhi, f); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ylo = fminf(ylo, __shfl_xor_sync(0xffffffffu, ylo, o)); yhi = fmaxf(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
            xlo = fminf(xlo, __shfl_xor_sync(0xffffffffu, xlo, o)); xhi = fmaxf(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
        }
        if (threadIdx.x == 0) {
            // NaN / huge coordinates (degenerate boxes) collapse to an empty footprint: every tap is out of bounds anyway
            const bool ok = ylo == ylo && xlo == xlo && fabsf(ylo) < 1e8f && fabsf(yhi) < 1e8f && fabsf(xlo) < 1e8f && fabsf(xhi) < 1e8f;
            ext[0] = ok ? max((int)ylo, 0) : 0;
            ext[1] = ok ? min((int)yhi + 1, H - 1) : -1;
            ext[2] = ok ? max((int)xlo, 0) : 0;
            ext[3] = ok ? min((int)xhi + 1, W - 1) : -1;
        }
    }
    __syncthreads();
    const int y0f = ext[0], y1f = ext[1], x0f = ext[2], x1f = ext[3];
    const int fh = y1f - y0f + 1, fw = x1f - x0f + 1;
    const float* plane = feats + ((int64_t)box_to_img[b] * C + c) * H * W;
    const bool staged = fh > 0 && fw > 0 && fh * fw <= kCropFootFloats;
    if (staged) {
        for (int t = threadIdx.x; t < fh * fw; t += 256) {
            const int r = t / fw, q = t - r * fw;
            foot[t] = plane[(y0f + r) * W + x0f + q];
        }
    }
    __syncthreads();
    float* out = crops + ((int64_t)b * C + c) * HH * WW;
    for (int t = threadIdx.x; t < HH * WW; t += 256) {
        const int i = t / WW, j = t - i * WW;
        const float ix = sx[j], iy = sy[i];
        const float xf = floorf(ix), yf = floorf(iy);
        float v = 0.f;
        if (xf >= -1.f && xf < (float)W && yf >= -1.f && yf < (float)H) {       // (also false for NaN coordinates)
            const int x0 = (int)xf, y0 = (int)yf;
            const float wx1 = ix - xf, wy1 = iy - yf;
            const float wx0 = (xf + 1.f) - ix, wy0 = (yf + 1.f) - iy;
            const bool vx0 = x0 >= 0 && x0 < W, vx1 = x0 + 1 >= 0 && x0 + 1 < W;
            const bool vy0 = y0 >= 0 && y0 < H, vy1 = y0 + 1 >= 0 && y0 + 1 < H;
            if (staged) {
                const float* p = foot + (y0 - y0f) * fw + (x0 - x0f);
                if (vy0 && vx0) v += p[0] * (wx0 * wy0);
                if (vy0 && vx1) v += p[1] * (wx1 * wy0);
                if (vy1 && vx0) v += p[fw] * (wx0 * wy1);
                if (vy1 && vx1) v += p[fw + 1] * (wx1 * wy1);
            } else {
                const float* p = plane + y0 * W + x0;
                if (vy0 && vx0) v += p[0] * (wx0 * wy0);
                if (vy0 && vx1) v += p[1] * (wx1 * wy0);
                if (vy1 && vx0) v += p[W] * (wx0 * wy1);
                if (vy1 && vx1) v += p[W + 1] * (wx1 * wy1);
            }
        }
        out[t] = v;
    }
}

// Backward, staged: one block per (image, channel, row tile).  The gradient tile of the image plane is accumulated in shared
// memory while the block walks the image's boxes in ascending order: per box the crop gradient (HH x WW) is staged with
// coalesced reads, pass 1 folds its rows into the tile rows (T[y][j] = sum_i wy(i -> y) d[i][j], ascending i), pass 2 folds
// the columns (acc[y][x] += sum_j wx(j -> x) T[y][j], ascending j, continuing the running sum) — the summation order of
// crop_bwd_rows/cols_kernel, so the result is bit-identical to them; nothing but the final tile touches global memory.
constexpr int kCropTileRows = 8;        // short tiles: more blocks, and a block walks only the boxes that overlap its rows
__global__ void __launch_bounds__(256) crop_bwd_staged_kernel(const float* __restrict__ dcrops, const float* __restrict__ boxes,
                                                             const int32_t* __restrict__ img_box_start,
                                                             const int32_t* __restrict__ box_order,
                                                             const float* __restrict__ wx, const float* __restrict__ wy,
                                                             const int4* __restrict__ ext, float* __restrict__ dfeats, int C,
                                                             int H, int W, int HH, int WW) {
    extern __shared__ float cb_smem[];
    float* acc = cb_smem;                                  // [kCropTileRows][W]
    float* dt = acc + kCropTileRows * W;                    // [HH][WW]   crop gradient of the current box
    float* T = dt + HH * WW;                                // [kCropTileRows][WW]
    float* sx = T + kCropTileRows * WW;                     // [WW] source x coordinate per crop column
    float* sy = sx + WW;                                    // [HH]
    const int n = blockIdx.x / C, c = blockIdx.x - n * C;
    const int ty0 = blockIdx.y * kCropTileRows;
    const int trows = min(kCropTileRows, H - ty0);
    for (int t = threadIdx.x; t < trows * W; t += 256) acc[t] = 0.f;
    for (int k = img_box_start[n]; k < img_box_start[n + 1]; ++k) {
        const int b = box_order[k];
        const int4 e = ext[b];
        const int ya = max(e.x, ty0), yb = min(e.y, ty0 + trows - 1);
        if (ya > yb || e.z > e.w) continue;                   // (uniform per block)
        const float* bx = boxes + b * 4;
        __syncthreads();
        for (int t = threadIdx.x; t < WW + HH; t += 256) {
            if (t < WW) sx[t] = crop_coord(bx[0], bx[2], wx[t], wx[WW + t], W);
            else sy[t - WW] = crop_coord(bx[1], bx[3], wy[t - WW], wy[HH + t - WW], H);
        }
        const float* d = dcrops + ((int64_t)b * C + c) * HH * WW;
        for (int t = threadIdx.x; t < HH * WW; t += 256) dt[t] = d[t];
        __syncthreads();
        const float cy0 = sy[0], cy1 = sy[HH - 1], cx0 = sx[0], cx1 = sx[WW - 1];
        const int nr = yb - ya + 1;
        for (int t = threadIdx.x; t < nr * WW; t += 256) {        // pass 1
            const int r = t / WW, j = t - r * WW;
            const int y = ya + r;
            int lo, hi;
            tap_range(cy0, cy1, HH, y, lo, hi);
            float a = 0.f;
            for (int i = lo; i < hi; ++i) {
                const float iy = sy[i], yf = floorf(iy);
                const int y0 = (int)yf;
                if (y0 == y) a += ((yf + 1.f) - iy) * dt[i * WW + j];
                else if (y0 + 1 == y) a += (iy - yf) * dt[i * WW + j];
            }
            T[r * WW + j] = a;
        }
        __syncthreads();
        const int xa = e.z, nx = e.w - e.z + 1;
        for (int t = threadIdx.x; t < nr * nx; t += 256) {        // pass 2
            const int r = t / nx, x = xa + (t - r * nx);
            int lo, hi;
            tap_range(cx0, cx1, WW, x, lo, hi);
            float a = acc[(ya - ty0 + r) * W + x];
            const float* row = T + r * WW;
            for (int j = lo; j < hi; ++j) {
                const float ix = sx[j], xf = floorf(ix);
                const int x0 = (int)xf;
                if (x0 == x) a += ((xf + 1.f) - ix) * row[j];
                else if (x0 + 1 == x) a += (ix - xf) * row[j];
            }
            acc[(ya - ty0 + r) * W + x] = a;
        }
    }
    __syncthreads();
    float* out = dfeats + (((int64_t)n * C + c) * H + ty0) * W;
    for (int t = threadIdx.x; t < trows * W; t += 256) out[t] = acc[t];
}

}  // namespace b200

using namespace b200;

static int g_crop_staged = 1;
extern "C" int b200_crop_set_staged(int enable) {
    const int prev = g_crop_staged;
    g_crop_staged = enable ? 1 : 0;
    return prev;
}

extern "C" int b200_crop_fwd(const float* feats, const float* boxes, const int32_t* box_to_img, const float* wx,
                             const float* wy, float* crops, int N, int C, int H, int W, int B, int HH, int WW,
                             b200_stream_t stream) {
    (void)N;
    if (B == 0) return 0;
    int64_t total = (int64_t)B * HH * WW;
    B200_REQUIRE(total < (1ll << 31), "crop_fwd: too many crop pixels");
    if (g_crop_staged && HH <= kCropMaxS && WW <= kCropMaxS && (int64_t)B * C < (1ll << 31))
        crop_fwd_staged_kernel<<<(unsigned)(B * C), 256, 0, as_stream(stream)>>>(feats, boxes, box_to_img, wx, wy, crops, C, H,
                                                                                W, HH, WW);
    else
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
    // ws: B*C*H*WW floats (pass-1 rows) followed by B int4 footprints (16-byte aligned: the float count is rounded up)
    const int64_t t1 = (int64_t)B * C * H * WW;
    int4* ext = reinterpret_cast<int4*>(ws + (t1 + 3) / 4 * 4);
    B200_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "crop_bwd: workspace must be 16-byte aligned");
    B200_REQUIRE(t1 < (1ll << 31) && (int64_t)N * C * H * W < (1ll << 31), "crop_bwd: tensors too large");
    const size_t smem = ((size_t)kCropTileRows * W + (size_t)HH * WW + (size_t)kCropTileRows * WW + WW + HH) * sizeof(float);
    if (g_crop_staged && smem <= 200 * 1024 && (int64_t)N * C < (1ll << 31)) {
        if (B > 0) {
            crop_extents_kernel<<<(B + 127) / 128, 128, 0, as_stream(stream)>>>(boxes, wx, wy, ext, B, H, W, HH, WW);
            B200_CHECK_LAUNCH();
        }
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(crop_bwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            B200_REQUIRE(e == cudaSuccess, "crop_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        }
        crop_bwd_staged_kernel<<<dim3((unsigned)(N * C), (unsigned)((H + kCropTileRows - 1) / kCropTileRows)), 256, smem,
                                 as_stream(stream)>>>(dcrops, boxes, img_box_start, box_order, wx, wy, ext, dfeats, C, H, W, HH, WW);
        B200_CHECK_LAUNCH();
        return 0;
    }
    if (B > 0) {
        crop_extents_kernel<<<(B + 127) / 128, 128, 0, as_stream(stream)>>>(boxes, wx, wy, ext, B, H, W, HH, WW);
        B200_CHECK_LAUNCH();
        crop_bwd_rows_kernel<<<B * C, 256, 0, as_stream(stream)>>>(dcrops, boxes, wy, ext, ws, C, H, B, HH, WW);
        B200_CHECK_LAUNCH();
    }
    int64_t t2 = (int64_t)N * C * H * W;
    crop_bwd_cols_kernel<<<grid_for(t2, 256), 256, 0, as_stream(stream)>>>(ws, boxes, img_box_start, box_order, wx, ext,
                                                                           dfeats, N, C, H, W, WW);
    B200_CHECK_LAUNCH();
    return 0;
}
