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


// ---------------------------------------------------------------------------------------------------------------------
// masks_to_layout (sg2im layout.py semantics; called by the reference at utils/draw_box.py:482-483, BASELINE config 5):
//   out[n, d, y, x] = sum_{o in image n, ascending} vecs[o, d] * bilinear(masks[o], grid_o(y, x))
// grid_o: X = (linspace(0,1,W)[x] - x0) / (x1 - x0), g = 2X - 1, source index ((g + 1) * M - 1) / 2 (align_corners=False),
// zeros padding.  The sample is rank-1 in (d, pixel): the mask is sampled ONCE per (object, pixel) and the object's
// embedding is scaled by it — the (O, D, M, M) product tensor and the (O, D, H, W) sampled tensor of the formulation above
// are never materialised.  The scatter-sum is a gather: each output pixel walks its image's objects in ascending order
// (deterministic, no atomics).  fp32 operation order of the index arithmetic is torch's (no FMA contraction): floors and
// in-bounds predicates are bit-exact.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float m2l_coord(float lin, float lo, float hi, int M) {
    const float X = __fdiv_rn(__fsub_rn(lin, lo), __fsub_rn(hi, lo));
    const float g = __fsub_rn(__fmul_rn(X, 2.f), 1.f);
    return __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)M), -1.f), 0.5f);
}

// bilinear sample of one M x M mask at source index (ix, iy): ATen grid_sampler_2d, zeros padding (nw, ne, sw, se order)
__device__ __forceinline__ float m2l_sample(const float* __restrict__ mask, int M, float ix, float iy) {
    const float x0f = floorf(ix), y0f = floorf(iy);
    // NaN / far-out coordinates (degenerate boxes): no tap is in bounds
    if (!(x0f >= -1.f && x0f < (float)M && y0f >= -1.f && y0f < (float)M)) return 0.f;
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float wx1 = ix - x0f, wy1 = iy - y0f;
    const float wx0 = (x0f + 1.f) - ix, wy0 = (y0f + 1.f) - iy;
    const bool vx0 = x0 >= 0, vx1 = x0 + 1 < M, vy0 = y0 >= 0, vy1 = y0 + 1 < M;
    float v = 0.f;
    if (vy0 && vx0) v += mask[y0 * M + x0] * (wx0 * wy0);
    if (vy0 && vx1) v += mask[y0 * M + x0 + 1] * (wx1 * wy0);
    if (vy1 && vx0) v += mask[(y0 + 1) * M + x0] * (wx0 * wy1);
    if (vy1 && vx1) v += mask[(y0 + 1) * M + x0 + 1] * (wx1 * wy1);
    return v;
}

__global__ void m2l_taps_kernel(const float* __restrict__ boxes, const float* __restrict__ linx,
                                const float* __restrict__ liny, int32_t* ix0, int32_t* iy0, float* fx, float* fy, int M,
                                int H, int W, int O) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < O * W) {
        const int o = t / W, x = t % W;
        const float ix = m2l_coord(linx[x], boxes[o * 4 + 0], boxes[o * 4 + 2], M);
        const float f = floorf(ix);
        ix0[t] = (int)f;
        fx[t] = ix - f;
    }
    if (t < O * H) {
        const int o = t / H, y = t % H;
        const float iy = m2l_coord(liny[y], boxes[o * 4 + 1], boxes[o * 4 + 3], M);
        const float f = floorf(iy);
        iy0[t] = (int)f;
        fy[t] = iy - f;
    }
}

// conservative pixel footprint of object o along one axis: every output index whose source coordinate can have an
// in-bounds tap (source index in (-1, M)  <=>  X in (-1/(2M), 1 + 1/(2M))), one pixel of slack
__device__ __forceinline__ void m2l_extent(float lo, float hi, int M, int S, int& a, int& b) {
    const float ext = hi - lo;
    if (!(ext > 0.f)) { a = 0; b = S - 1; return; }
    const float m = ext / (float)(2 * M);
    const float fa = floorf((lo - m) * (float)(S - 1)) - 1.f, fb = ceilf((hi + m) * (float)(S - 1)) + 1.f;
    a = fa < 0.f ? 0 : (fa > (float)(S - 1) ? S - 1 : (int)fa);
    b = fb < 0.f ? 0 : (fb > (float)(S - 1) ? S - 1 : (int)fb);
}

constexpr int kM2LObj = 32;      // objects per pass: their per-pixel mask samples live in registers

// block = 256 consecutive pixels of image blockIdx.y; the image's embeddings (<= 32 per pass) are staged in shared
// memory and read as 16-byte broadcasts; stores are coalesced along x in every channel plane
__global__ void __launch_bounds__(256) m2l_fwd_kernel(const float* __restrict__ vecs, const float* __restrict__ boxes,
                                                      const float* __restrict__ masks,
                                                      const int32_t* __restrict__ img_obj_start,
                                                      const int32_t* __restrict__ obj_order,
                                                      const float* __restrict__ linx, const float* __restrict__ liny,
                                                      float* __restrict__ out, int D, int M, int H, int W) {
    extern __shared__ float4 m2l_smem[];
    float4* vec_s = m2l_smem;                                     // [kM2LObj][D / 4]
    const int n = blockIdx.y;
    const int HW = H * W, D4 = D >> 2;
    const int p = blockIdx.x * 256 + threadIdx.x;
    const bool live = p < HW;
    const int y = live ? p / W : 0, x = live ? p - y * W : 0;
    const float lx = linx[x], ly = liny[y];
    const int k0 = img_obj_start[n], k1 = img_obj_start[n + 1];
    float* obase = out + (int64_t)n * D * HW + p;
    if (k0 == k1) {                                               // image without objects: zeros (scatter_add target)
        if (live)
            for (int d = 0; d < D; ++d) obase[(int64_t)d * HW] = 0.f;
        return;
    }
    __shared__ int sel[kM2LObj];
    __shared__ int nsel_s;
    const int p_last = min(blockIdx.x * 256 + 255, HW - 1);
    const int ty0 = (blockIdx.x * 256) / W, ty1 = p_last / W;       // image rows this block's pixels lie in
    bool first = true;
    for (int kb = k0; kb < k1; kb += kM2LObj) {
        const int cnt0 = min(kM2LObj, k1 - kb);
        __syncthreads();
        // cull: only the objects whose (conservative) row footprint meets the block's rows can contribute; the survivors
        // keep their ascending order (ballot compaction), so the summation order of the full walk is preserved
        if (threadIdx.x < 32) {
            bool hit = false;
            int o = 0;
            if ((int)threadIdx.x < cnt0) {
                o = obj_order[kb + threadIdx.x];
                const float4 bx = *reinterpret_cast<const float4*>(boxes + (int64_t)o * 4);
                int ya, yb;
                m2l_extent(bx.y, bx.w, M, H, ya, yb);
                hit = ya <= ty1 && yb >= ty0;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (hit) sel[__popc(bal & ((1u << threadIdx.x) - 1u))] = o;
            if (threadIdx.x == 0) nsel_s = __popc(bal);
        }
        __syncthreads();
        const int cnt = nsel_s;
        if (cnt == 0) continue;                                      // (uniform)
        for (int i = threadIdx.x; i < cnt * D4; i += 256) {
            const int j = i / D4, d4 = i - j * D4;
            vec_s[j * D4 + d4] = reinterpret_cast<const float4*>(vecs + (int64_t)sel[j] * D)[d4];
        }
        float s[kM2LObj];
#pragma unroll
        for (int j = 0; j < kM2LObj; ++j) {
            s[j] = 0.f;
            if (j < cnt) {
                const int o = sel[j];
                const float4 bx = *reinterpret_cast<const float4*>(boxes + (int64_t)o * 4);
                s[j] = m2l_sample(masks + (int64_t)o * M * M, M, m2l_coord(lx, bx.x, bx.z, M), m2l_coord(ly, bx.y, bx.w, M));
            }
        }
        __syncthreads();
        const bool had = !first;
        first = false;
        if (!live) continue;
        for (int d4 = 0; d4 < D4; ++d4) {
            float* o4 = obase + (int64_t)(4 * d4) * HW;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (had) acc = make_float4(o4[0], o4[HW], o4[2 * (int64_t)HW], o4[3 * (int64_t)HW]);
#pragma unroll
            for (int j = 0; j < kM2LObj; ++j) {
                if (j < cnt) {
                    const float4 v = vec_s[j * D4 + d4];
                    acc.x += s[j] * v.x; acc.y += s[j] * v.y; acc.z += s[j] * v.z; acc.w += s[j] * v.w;
                }
            }
            o4[0] = acc.x; o4[HW] = acc.y; o4[2 * (int64_t)HW] = acc.z; o4[3 * (int64_t)HW] = acc.w;
        }
    }
    if (first && live)                                            // no object reaches these rows: zeros
        for (int d = 0; d < D; ++d) obase[(int64_t)d * HW] = 0.f;
}

// backward, pass A — one block per object, restricted to the object's pixel footprint:
//   dvecs[o, d] = sum_pix dout[n, d, pix] * s_o(pix)          (fixed-order block reduction)
//   g[o, pix]   = sum_d   dout[n, d, pix] * vecs[o, d]        (the gradient reaching the sampled mask; pass B spreads it)
__global__ void __launch_bounds__(256) m2l_bwd_obj_kernel(const float* __restrict__ dout, const float* __restrict__ vecs,
                                                          const float* __restrict__ boxes, const float* __restrict__ masks,
                                                          const int32_t* __restrict__ obj_to_img,
                                                          const float* __restrict__ linx, const float* __restrict__ liny,
                                                          float* __restrict__ dvecs, float* __restrict__ g, int D, int M,
                                                          int H, int W) {
    extern __shared__ float4 m2l_smem[];
    float* vec_s = reinterpret_cast<float*>(m2l_smem);            // [D]
    __shared__ float red[8][32];
    const int o = blockIdx.x, n = obj_to_img[o];
    const int HW = H * W;
    const float4 bx = *reinterpret_cast<const float4*>(boxes + (int64_t)o * 4);
    const float* mask = masks + (int64_t)o * M * M;
    for (int d = threadIdx.x; d < D; d += 256) vec_s[d] = vecs[(int64_t)o * D + d];
    int xa, xb, ya, yb;
    m2l_extent(bx.x, bx.z, M, W, xa, xb);
    m2l_extent(bx.y, bx.w, M, H, ya, yb);
    const int fw = xb - xa + 1, npix = fw * (yb - ya + 1);
    const float* dbase = dout + (int64_t)n * D * HW;
    float* gb = g ? g + (int64_t)o * HW : nullptr;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d0 = 0; d0 < D; d0 += 32) {
        const int dn = min(32, D - d0);
        float acc[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = 0.f;
        for (int i = threadIdx.x; i < npix; i += 256) {
            const int r = i / fw, y = ya + r, x = xa + (i - r * fw);
            const float s = m2l_sample(mask, M, m2l_coord(linx[x], bx.x, bx.z, M), m2l_coord(liny[y], bx.y, bx.w, M));
            const float* dp = dbase + (int64_t)d0 * HW + y * W + x;
            float gs = 0.f;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                if (k < dn) {
                    const float v = dp[(int64_t)k * HW];
                    acc[k] += s * v;
                    gs += v * vec_s[d0 + k];
                }
            }
            if (gb) gb[y * W + x] = (d0 == 0 ? 0.f : gb[y * W + x]) + gs;
        }
        if (dvecs) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const float w = warp_sum(acc[k]);
                if (lane == 0) red[warp][k] = w;
            }
            __syncthreads();
            if (threadIdx.x < dn) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
                dvecs[(int64_t)o * D + d0 + threadIdx.x] = t;
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ void m2l_tap_range(float c_first, float c_last, int S, int target, int& lo, int& hi) {
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

// backward, pass B — dmasks[o, my, mx] = sum_y sum_x wy(y -> my) * wx(x -> mx) * g[o, y, x]: a gather over the output rows /
// columns whose bilinear taps land on (my, mx), ascending (the source coordinate is monotone in y and in x)
__global__ void m2l_bwd_mask_kernel(const float* __restrict__ g, const float* __restrict__ boxes,
                                    const float* __restrict__ linx, const float* __restrict__ liny,
                                    float* __restrict__ dmasks, int O, int M, int H, int W) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= O * M * M) return;
    const int mx = t % M, my = (t / M) % M, o = t / (M * M);
    const float4 bx = *reinterpret_cast<const float4*>(boxes + (int64_t)o * 4);
    const float* go = g + (int64_t)o * H * W;
    int ylo, yhi, xlo, xhi;
    m2l_tap_range(m2l_coord(liny[0], bx.y, bx.w, M), m2l_coord(liny[H - 1], bx.y, bx.w, M), H, my, ylo, yhi);
    m2l_tap_range(m2l_coord(linx[0], bx.x, bx.z, M), m2l_coord(linx[W - 1], bx.x, bx.z, M), W, mx, xlo, xhi);
    float acc = 0.f;
    for (int y = ylo; y < yhi; ++y) {
        const float iy = m2l_coord(liny[y], bx.y, bx.w, M);
        const float y0f = floorf(iy);
        float wy;
        if (y0f == (float)my) wy = (y0f + 1.f) - iy;
        else if (y0f + 1.f == (float)my) wy = iy - y0f;
        else continue;
        for (int x = xlo; x < xhi; ++x) {
            const float ix = m2l_coord(linx[x], bx.x, bx.z, M);
            const float x0f = floorf(ix);
            float wx;
            if (x0f == (float)mx) wx = (x0f + 1.f) - ix;
            else if (x0f + 1.f == (float)mx) wx = ix - x0f;
            else continue;
            acc += (wx * wy) * go[y * W + x];
        }
    }
    dmasks[t] = acc;
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

extern "C" int b200_masks_to_layout_taps(const float* boxes, const float* linx, const float* liny, int32_t* ix0,
                                         int32_t* iy0, float* fx, float* fy, int M, int H, int W, int O,
                                         b200_stream_t stream) {
    if (O == 0) return 0;
    const int n = O * (H > W ? H : W);
    m2l_taps_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(boxes, linx, liny, ix0, iy0, fx, fy, M, H, W, O);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_masks_to_layout_fwd(const float* vecs, const float* boxes, const float* masks,
                                        const int32_t* img_obj_start, const int32_t* obj_order, const float* linx,
                                        const float* liny, float* out, int N, int O, int D, int M, int H, int W,
                                        b200_stream_t stream) {
    (void)O;
    if (N == 0 || D == 0 || H == 0 || W == 0) return 0;
    B200_REQUIRE((D & 3) == 0 && D <= 1024, "masks_to_layout: D must be a multiple of 4, <= 1024");
    B200_REQUIRE(((reinterpret_cast<uintptr_t>(vecs) | reinterpret_cast<uintptr_t>(boxes)) & 15) == 0,
                 "masks_to_layout: vecs / boxes must be 16-byte aligned");
    B200_REQUIRE((int64_t)H * W < (1ll << 31) && N <= 65535, "masks_to_layout: sizes out of range");
    const size_t smem = (size_t)kM2LObj * D * sizeof(float);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(m2l_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((unsigned)((H * W + 255) / 256), (unsigned)N);
    m2l_fwd_kernel<<<grid, 256, smem, as_stream(stream)>>>(vecs, boxes, masks, img_obj_start, obj_order, linx, liny, out, D,
                                                           M, H, W);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_masks_to_layout_bwd(const float* dout, const float* vecs, const float* boxes, const float* masks,
                                        const int32_t* obj_to_img, const float* linx, const float* liny, float* dvecs,
                                        float* dmasks, float* ws, int N, int O, int D, int M, int H, int W,
                                        b200_stream_t stream) {
    (void)N;
    if (O == 0 || (dvecs == nullptr && dmasks == nullptr)) return 0;
    B200_REQUIRE(D <= 12288, "masks_to_layout_bwd: D too large");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "masks_to_layout_bwd: boxes must be 16-byte aligned");
    B200_REQUIRE(dmasks == nullptr || ws != nullptr, "masks_to_layout_bwd: workspace (O*H*W floats) required for dmasks");
    m2l_bwd_obj_kernel<<<O, 256, (size_t)D * sizeof(float), as_stream(stream)>>>(dout, vecs, boxes, masks, obj_to_img, linx,
                                                                                   liny, dvecs, dmasks ? ws : nullptr, D, M,
                                                                                   H, W);
    B200_CHECK_LAUNCH();
    if (dmasks) {
        const int n = O * M * M;
        m2l_bwd_mask_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(ws, boxes, linx, liny, dmasks, O, M, H, W);
        B200_CHECK_LAUNCH();
    }
    return 0;
}
