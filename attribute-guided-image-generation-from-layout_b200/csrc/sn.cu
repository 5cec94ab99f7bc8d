// sn.cu — spectral normalisation (K5): one power iteration per forward call and the gradient through W/sigma.
// Replaces torch.nn.utils.spectral_norm's pre-forward hook as installed by add_sn (discriminator.py:15-22):
//   v <- normalize(W^T u, eps), u <- normalize(W v, eps)  (in place), sigma = u . (W v), weight = weight_orig / sigma.
// W is the (h, w) row-major view of weight_orig.  All reductions are fixed order.
#include "common.cuh"

namespace b200 {

constexpr int kSnRowChunks = 8;      // row chunks of the W^T u pass (partial sums per chunk, summed in order)

// partial[j] = sum_{i in [i0, i1)} W[i][j] * u[i] for the 128 columns of block column `bx`; blockDim.x = 128.
// Vector layout (w % 4 == 0, 16-byte aligned): lane = column quad (a warp reads 512 contiguous bytes of a row), the 4
// warps take rows i0 + warp, step 4, and their sums are added in warp order through shared memory.  The single-layer and
// the whole-network kernels share this function (their results are compared bit for bit).
__device__ __forceinline__ void wtu_chunk(const float* __restrict__ W, const float* __restrict__ u, int w, int i0, int i1,
                                          int bx, float* __restrict__ partial) {
    __shared__ float4 red[3][32];
    if ((w & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(partial)) & 15) == 0) {
        const int q = threadIdx.x & 31, rl = threadIdx.x >> 5;
        const int j = (bx * 32 + q) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < w) {
            const float* Wp = W + j;
#pragma unroll 4
            for (int i = i0 + rl; i < i1; i += 4) {
                const float4 v = *reinterpret_cast<const float4*>(Wp + (int64_t)i * w);
                const float ui = u[i];
                acc.x += v.x * ui; acc.y += v.y * ui; acc.z += v.z * ui; acc.w += v.w * ui;
            }
        }
        if (rl > 0) red[rl - 1][q] = acc;
        __syncthreads();
        if (rl == 0 && j < w) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 o = red[k][q];
                acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            }
            *reinterpret_cast<float4*>(partial + j) = acc;
        }
        return;
    }
    const int j = bx * 128 + threadIdx.x;
    if (j >= w) return;
    float acc = 0.f;
    for (int i = i0; i < i1; ++i) acc += W[(int64_t)i * w + j] * u[i];
    partial[j] = acc;
}

// partial[chunk][j] = sum_{i in chunk} W[i][j] * u[i]
__global__ void sn_wtu_kernel(const float* __restrict__ W, const float* __restrict__ u, int h, int w, int rows_per_chunk,
                              float* __restrict__ partial) {
    int i0 = blockIdx.y * rows_per_chunk;
    int i1 = i0 + rows_per_chunk < h ? i0 + rows_per_chunk : h;
    wtu_chunk(W, u, w, i0, i1, blockIdx.x, partial + (int64_t)blockIdx.y * w);
}

__device__ float block_sum_1024(float v, float* red) {
    // fixed-order block reduction (blockDim.x == 1024)
    v = warp_sum(v);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = 0.f;
    if (wid == 0) {
        t = red[lane];
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// t = sum_chunks partial; v = t / max(||t||, eps)
__global__ void sn_norm_v_kernel(const float* __restrict__ partial, int nchunks, int w, float eps, float* __restrict__ v) {
    __shared__ float red[33];
    float ss = 0.f;
    for (int j = threadIdx.x; j < w; j += 1024) {
        float t = 0.f;
        for (int k = 0; k < nchunks; ++k) t += partial[(int64_t)k * w + j];
        v[j] = t;  // unnormalised for now
        ss += t * t;
    }
    float tot = block_sum_1024(ss, red);
    float inv = 1.f / fmaxf(sqrtf(tot), eps);
    for (int j = threadIdx.x; j < w; j += 1024) v[j] *= inv;
}

// one warp: sum_j Wr[j] * v[j]; 4 consecutive columns per lane (16-byte loads) when the row layout allows it.
// The single-layer and the whole-network kernels share this order (their results are compared bit for bit).
__device__ __forceinline__ float warp_row_dot(const float* __restrict__ Wr, const float* __restrict__ v, int w, int lane,
                                              bool vec) {
    float acc = 0.f;
    if (vec) {
#pragma unroll 4
        for (int j = lane * 4; j < w; j += 128) {
            const float4 q = *reinterpret_cast<const float4*>(Wr + j);
            const float4 p = *reinterpret_cast<const float4*>(v + j);
            acc += q.x * p.x + q.y * p.y + q.z * p.z + q.w * p.w;
        }
    } else {
        for (int j = lane; j < w; j += 32) acc += Wr[j] * v[j];
    }
    return warp_sum(acc);
}
__device__ __forceinline__ bool row_vec_ok(const float* W, const float* v, int w) {
    return (w & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
}

// wv[i] = sum_j W[i][j] v[j]; one warp per row
__global__ void sn_wv_kernel(const float* __restrict__ W, const float* __restrict__ v, int h, int w,
                             float* __restrict__ wv) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= h) return;
    const float acc = warp_row_dot(W + (int64_t)row * w, v, w, lane, row_vec_ok(W, v, w));
    if (lane == 0) wv[row] = acc;
}

// do_iter: u = wv / max(||wv||, eps); sigma = u . wv; out2 = [sigma, 1/sigma]
__global__ void sn_norm_u_kernel(const float* __restrict__ wv, int h, float eps, int do_iter, float* __restrict__ u,
                                 float* __restrict__ sigma_out, float* __restrict__ inv_out) {
    __shared__ float red[33];
    float ss = 0.f;
    for (int i = threadIdx.x; i < h; i += 1024) ss += wv[i] * wv[i];
    float tot = block_sum_1024(ss, red);
    float inv = 1.f / fmaxf(sqrtf(tot), eps);
    float dot = 0.f;
    for (int i = threadIdx.x; i < h; i += 1024) {
        float un = do_iter ? wv[i] * inv : u[i];
        if (do_iter) u[i] = un;
        dot += un * wv[i];
    }
    float sigma = block_sum_1024(dot, red);
    if (threadIdx.x == 0) {
        if (sigma_out) *sigma_out = sigma;
        *inv_out = 1.f / sigma;
    }
}

__global__ void sn_dot_partial_kernel(const float* __restrict__ g, const float* __restrict__ W, int64_t n,
                                      double* __restrict__ part) {
    __shared__ double red[8];
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += (double)g[i] * (double)W[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += red[k];
        part[1 + blockIdx.x] = s;
    }
}
__global__ void sn_dot_final_kernel(double* part, int nblocks) {
    // one warp, fixed order: lanes stride over the block partials, shuffle tree
    double s = 0.0;
    for (int k = threadIdx.x; k < nblocks; k += 32) s += part[1 + k];
    s = warp_sum(s);
    if (threadIdx.x == 0) part[0] = s;
}
// dW = g / sigma - <g, W> / sigma^2 * u v^T.  Every block first sums the partial dot products itself (one warp, lanes
// stride over the block partials, shuffle tree — the fixed order sn_dot_final_kernel uses), so no separate launch.
__global__ void sn_grad_apply_kernel(const float* __restrict__ g, const float* __restrict__ u,
                                     const float* __restrict__ v, const float* __restrict__ inv_sigma,
                                     const double* __restrict__ part, int nblocks, float* __restrict__ dW, int h, int w,
                                     int accumulate) {
    __shared__ float coef_s;
    const float inv = *inv_sigma;
    if (threadIdx.x < 32) {
        double s = 0.0;
        for (int k = threadIdx.x; k < nblocks; k += 32) s += part[1 + k];
        s = warp_sum(s);
        if (threadIdx.x == 0) coef_s = (float)s * inv * inv;
    }
    __syncthreads();
    const float coef = coef_s;
    const uint32_t n = (uint32_t)h * (uint32_t)w;
    if ((w & 3) == 0 && ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dW)) & 15) == 0) {
        const uint32_t n4 = n >> 2, w4 = (uint32_t)w >> 2;
        for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += gridDim.x * blockDim.x) {
            const uint32_t i = t / w4, j = (t - i * w4) << 2;
            const float4 gv = *reinterpret_cast<const float4*>(g + (size_t)t * 4);
            const float4 vv = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);      // history rows: 4-byte aligned only
            const float cu = coef * u[i];
            float4 o = make_float4(gv.x * inv - cu * vv.x, gv.y * inv - cu * vv.y, gv.z * inv - cu * vv.z, gv.w * inv - cu * vv.w);
            float4* dst = reinterpret_cast<float4*>(dW + (size_t)t * 4);
            if (accumulate) {
                const float4 p = *dst;
                o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
            }
            *dst = o;
        }
        return;
    }
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const uint32_t i = t / (uint32_t)w, j = t - i * (uint32_t)w;
        const float val = g[t] * inv - coef * u[i] * v[j];
        dW[t] = accumulate ? dW[t] + val : val;
    }
}


// ---------------------------------------------------------------------------------------------------------
// Grouped weight gradient through W / sigma_g: `groups` calls of a spectral-normalised layer batched along dim 0 share
// ONE weight-gradient GEMM whose pixel splits are aligned with the call boundaries (splits [g*spg, (g+1)*spg) belong to
// call g).  Phase A sums each call's splits into G_g (parameter layout (M, C, T); the partials are (M, tap, C)) and
// accumulates the partial dot products <G_g, W>; phase B forms
//     dW = sum_g ( G_g / sigma_g - <G_g, W> / sigma_g^2 * u_g v_g^T ).
// Replaces, per call, wgrad + wgrad_reduce + sn_dot_partial + sn_grad_apply.  All sums in a fixed order.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSnMaxGroups = 8;

// Pass A.  grid (ceil(C/32), row blocks), 256 threads; a block owns 32 channels x R rows (m) x all taps per pass.  The W tile
// of the pass is staged in shared memory (coalesced read, parameter order).  Each thread owns one (row, tap, channel quad)
// item: it issues the 16-byte loads of ALL calls' splits back to back (groups * spg independent loads in flight), forms
// G_g per call in registers, accumulates <G_g, W> against the staged tile and sum_g G_g / sigma_g, and writes the latter into
// the transposing tile; one coalesced store of the (m, c, tap)-ordered result per pass.  Two block syncs per pass.
// Output: dW_main = sum_g G_g / sigma_g and the per-block partial dots <G_g, W>; the per-call G_g is never stored.
//
// FOLD (the pooled convolution, see b200_fold_pool_weight): the partials are those of the (Th+1) x (Tw+1) stride-2 convolution
// with the folded weight; the gradient of the Th x Tw parameter tap (k, l) is 0.25 * the sum of the partials' taps
// (k+i, l+j), i, j in {0, 1} — the transpose of the fold, applied while the splits are summed.  T = Th * Tw, Tw = fold_tw.
template <bool FOLD>
__global__ void __launch_bounds__(256) sn_wgrad_reduce_kernel(const float* __restrict__ ws, int groups, int spg,
                                                             int64_t split_stride, int M, int T, int C, int R,
                                                             const float* __restrict__ W, const float* __restrict__ inv,
                                                             float* __restrict__ dW, double* __restrict__ dot_part,
                                                             int fold_tw) {
    extern __shared__ float sn_smem[];
    const int TP = T | 1;                         // odd tap pitch: conflict-free transposed access
    float* wtile = sn_smem;                       // [R][32][TP]  W, parameter order
    float* otile = sn_smem + R * 32 * TP;         // [R][32][TP]  sum_g G_g / sigma_g
    __shared__ double red[8];
    const int c0 = blockIdx.x * 32;
    const int nvalid = C - c0 < 32 ? C - c0 : 32;                 // a multiple of 4
    const int per_row = T * 8;                                     // (tap, channel quad) items per row
    const int items = R * per_row;
    const int cells = R * 32 * T;
    const int nparts = gridDim.x * gridDim.y;
    float invs[kSnMaxGroups];
    double dot[kSnMaxGroups];
#pragma unroll
    for (int g = 0; g < kSnMaxGroups; ++g) { dot[g] = 0.0; invs[g] = g < groups ? inv[g] : 0.f; }
    for (int m0 = blockIdx.y * R; m0 < M; m0 += gridDim.y * R) {
        for (int i = threadIdx.x; i < cells; i += 256) {
            const int r = i / (32 * T), j = i - r * 32 * T;
            const int c = j / T, tap = j - c * T;
            wtile[(r * 32 + c) * TP + tap] = (m0 + r < M && c < nvalid) ? W[((int64_t)(m0 + r) * C + c0) * T + j] : 0.f;
        }
        __syncthreads();
        for (int it = threadIdx.x; it < items; it += 256) {
            const int r = it / per_row, rem = it - r * per_row;
            const int tap = rem >> 3, q = rem & 7;
            const int m = m0 + r;
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M && q * 4 < nvalid) {
                int64_t poff;                      // this item's offset inside one split's partial matrix
                int T4 = T, tw4 = 0;
                if (FOLD) {
                    const int th = T / fold_tw;
                    tw4 = fold_tw + 1;
                    T4 = (th + 1) * tw4;
                    const int k = tap / fold_tw, l = tap - k * fold_tw;
                    poff = ((int64_t)m * T4 + k * tw4 + l) * C + c0 + q * 4;
                } else {
                    poff = ((int64_t)m * T + tap) * C + c0 + q * 4;
                }
                const float* p = ws + poff;
                const float* wt = wtile + (r * 32 + q * 4) * TP + tap;
                const float w0 = wt[0], w1 = wt[TP], w2 = wt[2 * TP], w3 = wt[3 * TP];
#pragma unroll
                for (int g = 0; g < kSnMaxGroups; ++g) {
                    if (g < groups) {
                        const float* pg = p + (int64_t)g * spg * split_stride;
                        float4 a;
                        if (FOLD) {
                            a = make_float4(0.f, 0.f, 0.f, 0.f);
                            for (int sidx = 0; sidx < spg; ++sidx) {
                                const float* ps = pg + (int64_t)sidx * split_stride;
                                const float4 v0 = *reinterpret_cast<const float4*>(ps);
                                const float4 v1 = *reinterpret_cast<const float4*>(ps + C);
                                const float4 v2 = *reinterpret_cast<const float4*>(ps + (int64_t)tw4 * C);
                                const float4 v3 = *reinterpret_cast<const float4*>(ps + (int64_t)(tw4 + 1) * C);
                                a.x += (v0.x + v1.x) + (v2.x + v3.x); a.y += (v0.y + v1.y) + (v2.y + v3.y);
                                a.z += (v0.z + v1.z) + (v2.z + v3.z); a.w += (v0.w + v1.w) + (v2.w + v3.w);
                            }
                            a.x *= 0.25f; a.y *= 0.25f; a.z *= 0.25f; a.w *= 0.25f;
                        } else {
                            a = *reinterpret_cast<const float4*>(pg);
                            for (int sidx = 1; sidx < spg; ++sidx) {
                                const float4 v = *reinterpret_cast<const float4*>(pg + (int64_t)sidx * split_stride);
                                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
                            }
                        }
                        dot[g] += (double)a.x * w0 + (double)a.y * w1 + (double)a.z * w2 + (double)a.w * w3;
                        o.x += a.x * invs[g]; o.y += a.y * invs[g]; o.z += a.z * invs[g]; o.w += a.w * invs[g];
                    }
                }
            }
            float* t = otile + (r * 32 + q * 4) * TP + tap;
            t[0] = o.x; t[TP] = o.y; t[2 * TP] = o.z; t[3 * TP] = o.w;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cells; i += 256) {
            const int r = i / (32 * T), j = i - r * 32 * T;
            const int c = j / T, tap = j - c * T;
            if (m0 + r < M && c < nvalid) dW[((int64_t)(m0 + r) * C + c0) * T + j] = otile[(r * 32 + c) * TP + tap];
        }
    }
    for (int g = 0; g < groups; ++g) {
        __syncthreads();
        const double d = warp_sum(dot[g]);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int k = 0; k < 8; ++k) t += red[k];
            dot_part[(int64_t)g * nparts + blockIdx.y * gridDim.x + blockIdx.x] = t;
        }
    }
}

// Pass B.  dW -= sum_g (<G_g, W> / sigma_g^2) u_g v_g^T   (rank-1 corrections; one read + one write of |W|)
__global__ void __launch_bounds__(256) sn_grad_groups_kernel(int groups, const float* __restrict__ u_hist,
                                                            const float* __restrict__ v_hist,
                                                            const float* __restrict__ inv, const double* __restrict__ dot_part,
                                                            int nparts, float* __restrict__ dW, int h, int w) {
    __shared__ float coef[kSnMaxGroups];
    for (int g = threadIdx.x >> 5; g < groups; g += 8) {          // warp g: fixed-order sum of call g's partial dots
        double sdot = 0.0;
        for (int k = threadIdx.x & 31; k < nparts; k += 32) sdot += dot_part[(int64_t)g * nparts + k];
        sdot = warp_sum(sdot);
        if ((threadIdx.x & 31) == 0) {
            const float iv = inv[g];
            coef[g] = (float)sdot * iv * iv;
        }
    }
    __syncthreads();
    const uint32_t n = (uint32_t)h * (uint32_t)w;
    const uint32_t n4 = n >> 2, w4 = (uint32_t)w >> 2;             // w % 4 == 0
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += gridDim.x * blockDim.x) {
        const uint32_t i = t / w4, j = (t - i * w4) << 2;
        float4 o = *reinterpret_cast<const float4*>(dW + (size_t)t * 4);
        for (int g = 0; g < groups; ++g) {
            const float* v = v_hist + (size_t)g * w + j;          // (the staged history is only 4-byte aligned)
            const float cu = coef[g] * u_hist[(size_t)g * h + i];
            o.x -= cu * v[0];
            o.y -= cu * v[1];
            o.z -= cu * v[2];
            o.w -= cu * v[3];
        }
        *reinterpret_cast<float4*>(dW + (size_t)t * 4) = o;
    }
}

// ---------------------------------------------------------------------------------------------------------
// whole-network power iteration: every spectral-normalised layer of a discriminator in one launch sequence
// (blockIdx.z / blockIdx.y = layer; blocks beyond a layer's extent exit).  Same arithmetic and reduction order
// per layer as the single-layer kernels above.
// ---------------------------------------------------------------------------------------------------------
__global__ void snm_wtu_kernel(const b200_sn_layer* __restrict__ layers) {
    const b200_sn_layer l = layers[blockIdx.z];
    const int nch = l.h < 64 ? 1 : kSnRowChunks;
    if ((int)blockIdx.y >= nch || (int)blockIdx.x * 128 >= l.w) return;
    const int rpc = (l.h + nch - 1) / nch;
    const int i0 = blockIdx.y * rpc;
    const int i1 = i0 + rpc < l.h ? i0 + rpc : l.h;
    wtu_chunk(l.W, l.u, l.w, i0, i1, blockIdx.x, l.ws + (int64_t)blockIdx.y * l.w);
}

__global__ void snm_norm_v_kernel(const b200_sn_layer* __restrict__ layers, float eps, int it) {
    __shared__ float red[33];
    const b200_sn_layer l = layers[blockIdx.x];
    const int nch = l.h < 64 ? 1 : kSnRowChunks;
    float ss = 0.f;
    for (int j = threadIdx.x; j < l.w; j += 1024) {
        float t = 0.f;
        for (int k = 0; k < nch; ++k) t += l.ws[(int64_t)k * l.w + j];
        l.v[j] = t;
        ss += t * t;
    }
    float tot = block_sum_1024(ss, red);
    float inv = 1.f / fmaxf(sqrtf(tot), eps);
    for (int j = threadIdx.x; j < l.w; j += 1024) l.v[j] *= inv;
    (void)it;
}

__global__ void snm_wv_kernel(const b200_sn_layer* __restrict__ layers) {
    const b200_sn_layer l = layers[blockIdx.y];
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= l.h) return;
    const float acc = warp_row_dot(l.W + (int64_t)row * l.w, l.v, l.w, lane, row_vec_ok(l.W, l.v, l.w));
    if (lane == 0) l.ws[(int64_t)kSnRowChunks * l.w + row] = acc;
}

// u update, sigma, and the per-iteration record: inv[it], u_hist[it], v_hist[it]
__global__ void snm_norm_u_kernel(const b200_sn_layer* __restrict__ layers, float eps, int do_iter, int it) {
    __shared__ float red[33];
    const b200_sn_layer l = layers[blockIdx.x];
    const float* wv = l.ws + (int64_t)kSnRowChunks * l.w;
    float ss = 0.f;
    for (int i = threadIdx.x; i < l.h; i += 1024) ss += wv[i] * wv[i];
    float tot = block_sum_1024(ss, red);
    float inv = 1.f / fmaxf(sqrtf(tot), eps);
    float dot = 0.f;
    for (int i = threadIdx.x; i < l.h; i += 1024) {
        float un = do_iter ? wv[i] * inv : l.u[i];
        if (do_iter) l.u[i] = un;
        if (l.u_hist) l.u_hist[(int64_t)it * l.h + i] = un;
        dot += un * wv[i];
    }
    if (l.v_hist)
        for (int j = threadIdx.x; j < l.w; j += 1024) l.v_hist[(int64_t)it * l.w + j] = l.v[j];
    float sigma = block_sum_1024(dot, red);
    if (threadIdx.x == 0) l.inv[it] = 1.f / sigma;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_sn_power_iter(const float* W, int h, int w, float* u, float* v, int do_iter, float eps,
                                  float* sigma_out, float* inv_sigma_out, float* ws, b200_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    float* partial = ws;                       // kSnRowChunks * w
    float* wv = ws + (int64_t)kSnRowChunks * w;  // h
    if (do_iter) {
        int nch = h < 64 ? 1 : kSnRowChunks;
        int rpc = (h + nch - 1) / nch;
        dim3 grid((w + 127) / 128, nch);
        sn_wtu_kernel<<<grid, 128, 0, st>>>(W, u, h, w, rpc, partial);
        B200_CHECK_LAUNCH();
        sn_norm_v_kernel<<<1, 1024, 0, st>>>(partial, nch, w, eps, v);
        B200_CHECK_LAUNCH();
    }
    sn_wv_kernel<<<(h + 7) / 8, 256, 0, st>>>(W, v, h, w, wv);
    B200_CHECK_LAUNCH();
    sn_norm_u_kernel<<<1, 1024, 0, st>>>(wv, h, eps, do_iter, u, sigma_out, inv_sigma_out);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_sn_grad(const float* g, const float* W, const float* u, const float* v, const float* inv_sigma,
                            float* dW, int h, int w, int accumulate, double* ws, b200_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    int64_t n = (int64_t)h * w;
    int nblocks = grid_for(n, 256, 4);
    if (nblocks > 1023) nblocks = 1023;
    B200_REQUIRE(n < (1ll << 32), "sn_grad: weight too large");
    sn_dot_partial_kernel<<<nblocks, 256, 0, st>>>(g, W, n, ws);
    B200_CHECK_LAUNCH();
    sn_grad_apply_kernel<<<grid_for((w & 3) == 0 ? n / 4 : n, 256), 256, 0, st>>>(g, u, v, inv_sigma, ws, nblocks, dW, h, w, accumulate);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_sn_wgrad_parts(int M, int C) {
    const int cb = (C + 31) / 32;
    int rows = (kNumSMs * 8 + cb - 1) / cb;
    if (rows > M) rows = M;
    if (rows < 1) rows = 1;
    return cb * rows;
}

static int sn_wgrad_finish_impl(const float* ws, int groups, int splits_per_group, int64_t split_stride, int M, int T,
                                int C, const float* W, const float* u_hist, const float* v_hist, const float* inv,
                                float* Gbuf, double* dot_part, float* dW, int fold_tw, b200_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    B200_REQUIRE(fold_tw == 0 || (fold_tw >= 1 && T % fold_tw == 0), "sn_wgrad_finish: fold_tw must divide T");
    B200_REQUIRE(groups >= 1 && groups <= kSnMaxGroups && splits_per_group >= 1, "sn_wgrad_finish: bad groups / splits");
    B200_REQUIRE((C & 3) == 0 && T >= 1 && T <= 64 && M >= 1 && M < 65536 && (split_stride & 3) == 0 &&
                     (int64_t)M * C * T < (1ll << 31),
                 "sn_wgrad_finish: needs C %% 4 == 0, T <= 64, M < 65536");
    B200_REQUIRE(((reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(Gbuf) | reinterpret_cast<uintptr_t>(dW)) & 15) == 0,
                 "sn_wgrad_finish: buffers must be 16-byte aligned");
    (void)Gbuf;                                   // kept in the signature; the per-call gradients are no longer materialised
    const int cb = (C + 31) / 32;
    const int nparts = b200_sn_wgrad_parts(M, C);
    const int rows = nparts / cb;
    int R = 256 / (T * 8);                        // rows per pass: about one (tap, channel quad) item per thread
    if (R < 1) R = 1;
    if (R > 32) R = 32;
    if (R > M) R = M;
    const size_t smem = (size_t)2 * R * 32 * (T | 1) * sizeof(float);
    B200_REQUIRE(smem <= 48 * 1024, "sn_wgrad_finish: tile too large");
    int rblocks = (rows + R - 1) / R;
    if (rblocks < 1) rblocks = 1;
    if (fold_tw)
        sn_wgrad_reduce_kernel<true><<<dim3(cb, rblocks), 256, smem, st>>>(ws, groups, splits_per_group, split_stride, M, T, C,
                                                                          R, W, inv, dW, dot_part, fold_tw);
    else
        sn_wgrad_reduce_kernel<false><<<dim3(cb, rblocks), 256, smem, st>>>(ws, groups, splits_per_group, split_stride, M, T, C,
                                                                           R, W, inv, dW, dot_part, 0);
    B200_CHECK_LAUNCH();
    const int64_t n = (int64_t)M * C * T;
    sn_grad_groups_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(groups, u_hist, v_hist, inv, dot_part, cb * rblocks, dW, M, C * T);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_sn_wgrad_finish(const float* ws, int groups, int splits_per_group, int64_t split_stride, int M, int T,
                                    int C, const float* W, const float* u_hist, const float* v_hist, const float* inv,
                                    float* Gbuf, double* dot_part, float* dW, b200_stream_t stream) {
    return sn_wgrad_finish_impl(ws, groups, splits_per_group, split_stride, M, T, C, W, u_hist, v_hist, inv, Gbuf, dot_part, dW,
                                0, stream);
}

extern "C" int b200_sn_wgrad_finish_pooled(const float* ws, int groups, int splits_per_group, int64_t split_stride, int M,
                                           int Th, int Tw, int C, const float* W, const float* u_hist, const float* v_hist,
                                           const float* inv, double* dot_part, float* dW, b200_stream_t stream) {
    B200_REQUIRE(Th >= 1 && Tw >= 1, "sn_wgrad_finish_pooled: bad taps");
    return sn_wgrad_finish_impl(ws, groups, splits_per_group, split_stride, M, Th * Tw, C, W, u_hist, v_hist, inv, nullptr,
                                dot_part, dW, Tw, stream);
}

extern "C" int b200_sn_power_iter_multi(const b200_sn_layer* layers, int n_layers, int max_h, int max_w, int iters,
                                        int do_iter, float eps, b200_stream_t stream) {
    if (n_layers <= 0 || iters <= 0) return 0;
    B200_REQUIRE(n_layers < 65536 && max_h > 0 && max_w > 0, "sn_power_iter_multi: bad sizes");
    cudaStream_t st = as_stream(stream);
    for (int it = 0; it < iters; ++it) {
        if (do_iter) {
            snm_wtu_kernel<<<dim3((max_w + 127) / 128, kSnRowChunks, n_layers), 128, 0, st>>>(layers);
            B200_CHECK_LAUNCH();
            snm_norm_v_kernel<<<n_layers, 1024, 0, st>>>(layers, eps, it);
            B200_CHECK_LAUNCH();
        }
        snm_wv_kernel<<<dim3((max_h + 7) / 8, n_layers), 256, 0, st>>>(layers);
        B200_CHECK_LAUNCH();
        snm_norm_u_kernel<<<n_layers, 1024, 0, st>>>(layers, eps, do_iter, it);
        B200_CHECK_LAUNCH();
    }
    return 0;
}
