// elementwise.cu — HBM-bound pointwise / pooling / layout kernels of the G+D step (channel-last activations stored as fp32 or bf16, fp32 arithmetic).
// Reference call sites are cited per entry point in include/b200gan.h.
#include "common.cuh"

namespace b200 {

#define GRID_STRIDE(i, n) \
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

// T = storage type of the channel-last activations (float or bf16); arithmetic is fp32 throughout.
template <typename T>
__global__ void relu_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t nv, int tail) {
    constexpr int V = VecIO<T>::V;
    GRID_STRIDE(i, nv) {
        float v[V];
        VecIO<T>::load(x + i * V, v);
#pragma unroll
        for (int e = 0; e < V; ++e) v[e] = fmaxf(v[e], 0.f);
        VecIO<T>::store(y + i * V, v);
    }
    if (blockIdx.x == 0 && threadIdx.x < tail) stf(y + nv * V + threadIdx.x, fmaxf(ldf(x + nv * V + threadIdx.x), 0.f));
}

template <typename T>
__global__ void relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, int64_t nv,
                                int tail) {
    constexpr int V = VecIO<T>::V;
    GRID_STRIDE(i, nv) {
        float g[V], v[V];
        VecIO<T>::load(dy + i * V, g);
        VecIO<T>::load(y + i * V, v);
#pragma unroll
        for (int e = 0; e < V; ++e) g[e] = v[e] > 0.f ? g[e] : 0.f;
        VecIO<T>::store(dx + i * V, g);
    }
    if (blockIdx.x == 0 && threadIdx.x < tail) {
        const int64_t t = nv * V + threadIdx.x;
        stf(dx + t, ldf(y + t) > 0.f ? ldf(dy + t) : 0.f);
    }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, int64_t nv, int tail,
                           int relu) {
    constexpr int V = VecIO<T>::V;
    GRID_STRIDE(i, nv) {
        float u[V], v[V];
        VecIO<T>::load(a + i * V, u);
        VecIO<T>::load(b + i * V, v);
#pragma unroll
        for (int e = 0; e < V; ++e) {
            u[e] += v[e];
            if (relu) u[e] = fmaxf(u[e], 0.f);
        }
        VecIO<T>::store(o + i * V, u);
    }
    if (blockIdx.x == 0 && threadIdx.x < tail) {
        const int64_t t = nv * V + threadIdx.x;
        const float s = ldf(a + t) + ldf(b + t);
        stf(o + t, relu ? fmaxf(s, 0.f) : s);
    }
}

// y[n,qy,qx,c] = scale * sum_{dy,dx<f} x[n,qy*f+dy,qx*f+dx,c];  V = channels per thread (16 bytes when C % V == 0, else
// 1); I = index type (32-bit when the element count fits: no 64-bit divisions)
template <typename T, int V, typename I>
__global__ void pool_fwd_kernel(const T* __restrict__ x, const T* __restrict__ x2, T* __restrict__ y, int N, int H, int W,
                                int C, int f, float scale, int relu) {
    const int Ho = H / f, Wo = W / f, Cv = C / V;
    const I total = (I)N * Ho * Wo * Cv;
    for (I t = blockIdx.x * (I)blockDim.x + threadIdx.x; t < total; t += (I)gridDim.x * blockDim.x) {
        const int c = (int)(t % Cv) * V;
        I q = t / Cv;
        const int qx = (int)(q % Wo);
        q /= Wo;
        const int qy = (int)(q % Ho);
        const int n = (int)(q / Ho);
        const T* p = x + (((int64_t)n * H + (int64_t)qy * f) * W + (int64_t)qx * f) * C + c;
        if constexpr (V > 1) {
            static_assert(V == VecIO<T>::V, "vector width");
            float acc[V];
#pragma unroll
            for (int e = 0; e < V; ++e) acc[e] = 0.f;
            for (int dy = 0; dy < f; ++dy)
                for (int dx = 0; dx < f; ++dx) {
                    float v[V];
                    VecIO<T>::load(p + ((int64_t)dy * W + dx) * C, v);
#pragma unroll
                    for (int e = 0; e < V; ++e) acc[e] += v[e];
                    if (x2) {           // pooled SUM of two tensors (residual branch + shortcut of a downsampling block)
                        VecIO<T>::load(x2 + (p - x) + ((int64_t)dy * W + dx) * C, v);
#pragma unroll
                        for (int e = 0; e < V; ++e) acc[e] += v[e];
                    }
                }
#pragma unroll
            for (int e = 0; e < V; ++e) {
                acc[e] *= scale;
                if (relu) acc[e] = fmaxf(acc[e], 0.f);
            }
            VecIO<T>::store(y + (int64_t)t * V, acc);
        } else {
            float acc = 0.f;
            for (int dy = 0; dy < f; ++dy)
                for (int dx = 0; dx < f; ++dx) {
                    acc += ldf(p + ((int64_t)dy * W + dx) * C);
                    if (x2) acc += ldf(x2 + (p - x) + ((int64_t)dy * W + dx) * C);
                }
            stf(y + (int64_t)t, relu ? fmaxf(acc * scale, 0.f) : acc * scale);
        }
    }
}

// y[n,oy,ox,c] = scale * x[n,oy/f,ox/f,c]  (nearest upsampling; the adjoint of pool_fwd)
template <typename T, int V, typename I>
__global__ void unpool_fwd_kernel(const T* __restrict__ x, const T* __restrict__ mask, T* __restrict__ y, int N, int H,
                                  int W, int C, int f, float scale) {
    const int Ho = H * f, Wo = W * f, Cv = C / V;
    const I total = (I)N * Ho * Wo * Cv;
    for (I t = blockIdx.x * (I)blockDim.x + threadIdx.x; t < total; t += (I)gridDim.x * blockDim.x) {
        const int c = (int)(t % Cv) * V;
        I q = t / Cv;
        const int ox = (int)(q % Wo);
        q /= Wo;
        const int oy = (int)(q % Ho);
        const int n = (int)(q / Ho);
        const T* p = x + (((int64_t)n * H + oy / f) * W + ox / f) * C + c;
        if constexpr (V > 1) {
            static_assert(V == VecIO<T>::V, "vector width");
            float v[V];
            VecIO<T>::load(p, v);
            if (mask) {          // gradient of relu(pool(.)): zero where the pooled output was not positive
                float mk[V];
                VecIO<T>::load(mask + (p - x), mk);
#pragma unroll
                for (int e = 0; e < V; ++e)
                    if (!(mk[e] > 0.f)) v[e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < V; ++e) v[e] *= scale;
            VecIO<T>::store(y + (int64_t)t * V, v);
        } else {
            float v = ldf(p);
            if (mask && !(ldf(mask + (p - x)) > 0.f)) v = 0.f;
            stf(y + (int64_t)t, scale * v);
        }
    }
}

template <typename T>
__global__ void concat_fwd_kernel(const T* __restrict__ a, int Ca, int a_div, const T* __restrict__ b, int Cb,
                                  int b_div, T* __restrict__ out, int64_t rows) {
    int Ct = Ca + Cb;
    int64_t total = rows * Ct;
    GRID_STRIDE(t, total) {
        int c = (int)(t % Ct);
        int64_t r = t / Ct;
        out[t] = c < Ca ? a[(r / a_div) * Ca + c] : b[(r / b_div) * Cb + (c - Ca)];
    }
}

// adjoint of concat: da[ra][c] = sum_{r in [ra*a_div, (ra+1)*a_div)} dout[r][c] (ascending r), same for b
template <typename T>
__global__ void concat_bwd_kernel(const T* __restrict__ dout, int Ca, int a_div, T* __restrict__ da, int Cb,
                                  int b_div, T* __restrict__ db, int64_t rows) {
    int Ct = Ca + Cb;
    int64_t na = da ? (rows / a_div) * Ca : 0;
    int64_t nb = db ? (rows / b_div) * Cb : 0;
    GRID_STRIDE(t, na + nb) {
        if (t < na) {
            int c = (int)(t % Ca);
            int64_t r0 = (t / Ca) * a_div;
            float acc = 0.f;
            for (int k = 0; k < a_div; ++k) acc += ldf(dout + (r0 + k) * Ct + c);
            stf(da + t, acc);
        } else {
            int64_t u = t - na;
            int c = (int)(u % Cb);
            int64_t r0 = (u / Cb) * b_div;
            float acc = 0.f;
            for (int k = 0; k < b_div; ++k) acc += ldf(dout + (r0 + k) * Ct + Ca + c);
            stf(db + u, acc);
        }
    }
}

__global__ void gather_rows_kernel(const float* __restrict__ table, const int32_t* __restrict__ idx,
                                   float* __restrict__ out, int rows, int D) {
    int64_t total = (int64_t)rows * D;
    GRID_STRIDE(t, total) {
        int c = (int)(t % D);
        int r = (int)(t / D);
        out[t] = table[(int64_t)idx[r] * D + c];
    }
}

__global__ void scatter_rows_kernel(const float* __restrict__ dout, const int32_t* __restrict__ idx,
                                    float* __restrict__ dtable, int rows, int D, int num_classes) {
    int64_t total = (int64_t)num_classes * D;
    GRID_STRIDE(t, total) {
        int c = (int)(t % D);
        int k = (int)(t / D);
        float acc = 0.f;
        for (int r = 0; r < rows; ++r)
            if (idx[r] == k) acc += dout[(int64_t)r * D + c];
        dtable[t] = acc;
    }
}

// rows of 16-byte units (any element type)
__global__ void permute_rows_kernel(const uint4* __restrict__ x, const int32_t* __restrict__ src_row,
                                    uint4* __restrict__ out, int rows, int rowlen16) {
    int64_t total = (int64_t)rows * rowlen16;
    GRID_STRIDE(t, total) {
        int c = (int)(t % rowlen16);
        int r = (int)(t / rowlen16);
        int s = src_row[r];
        out[t] = s >= 0 ? x[(int64_t)s * rowlen16 + c] : make_uint4(0u, 0u, 0u, 0u);
    }
}

template <typename T>
__global__ void mask_outer_fwd_kernel(const float* __restrict__ v, const float* __restrict__ mask,
                                      T* __restrict__ out, int O, int H, int W, int C) {
    constexpr int V = VecIO<T>::V;
    const int Hp = H + 2, Wp = W + 2, Cv = C / V;
    const int64_t total = (int64_t)O * Hp * Wp * Cv;
    GRID_STRIDE(t, total) {
        // total < 2^31 is checked by the launcher: 32-bit divisions
        const uint32_t u = (uint32_t)t;
        const int c = (int)(u % Cv) * V;
        uint32_t q = u / Cv;
        const int x = (int)(q % Wp);
        q /= Wp;
        const int y = (int)(q % Hp);
        const int o = (int)(q / Hp);
        float r[V];
#pragma unroll
        for (int e = 0; e < V; ++e) r[e] = 0.f;
        if (y >= 1 && y <= H && x >= 1 && x <= W) {
            const float m = mask[((int64_t)o * H + (y - 1)) * W + (x - 1)];
            if (m != 0.f) {
                ldp<V>(v + (int64_t)o * C + c, r);
#pragma unroll
                for (int e = 0; e < V; ++e) r[e] *= m;
            }
        }
        VecIO<T>::store(out + t * V, r);
    }
}

// dv[o,c] = sum_{y,x} mask[o,y,x] * dout[o,y+1,x+1,c]; one block per (o, 32-channel group), fixed-order tree
template <typename T>
__global__ void mask_outer_bwd_kernel(const T* __restrict__ dout, const float* __restrict__ mask,
                                      float* __restrict__ dv, int O, int H, int W, int C) {
    __shared__ float red[8][32];
    int o = blockIdx.x;
    int c = blockIdx.y * 32 + threadIdx.x;
    int Wp = W + 2;
    float acc = 0.f;
    if (c < C) {
        for (int p = threadIdx.y; p < H * W; p += 8) {
            int y = p / W, x = p % W;
            float m = mask[((int64_t)o * H + y) * W + x];
            if (m != 0.f) acc += m * ldf(dout + (((int64_t)o * (H + 2) + (y + 1)) * Wp + (x + 1)) * C + c);
        }
    }
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
        dv[(int64_t)o * C + c] = s;
    }
}

// vector variant: one block per object; a thread owns V channels (16 bytes) of a fixed slot and walks the pixels with
// stride 256 / (C / V); fixed-order tree over the pixel lanes (warp shuffles, then the 8 warps).  C / V a power of two <= 32.
template <typename T>
__global__ void __launch_bounds__(256) mask_outer_bwd_vec_kernel(const T* __restrict__ dout, const float* __restrict__ mask,
                                                                float* __restrict__ dv, int H, int W, int C) {
    constexpr int V = VecIO<T>::V;
    __shared__ float sm[8][32 * V];
    const int o = blockIdx.x;
    const int ct = C / V;
    const int slot = threadIdx.x % ct;
    const int pstep = 256 / ct;
    const int Wp = W + 2;
    const T* base = dout + (int64_t)o * (H + 2) * Wp * C + slot * V;
    const float* mk = mask + (int64_t)o * H * W;
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll 4
    for (int p = threadIdx.x / ct; p < H * W; p += pstep) {
        const float m = mk[p];
        if (m != 0.f) {
            const int y = p / W, x = p - y * W;
            float v[V];
            VecIO<T>::load(base + ((int64_t)(y + 1) * Wp + (x + 1)) * C, v);
#pragma unroll
            for (int e = 0; e < V; ++e) acc[e] += m * v[e];
        }
    }
    for (int off = 16; off >= ct; off >>= 1) {
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < ct) {
#pragma unroll
        for (int e = 0; e < V; ++e) sm[warp][lane * V + e] = acc[e];
    }
    __syncthreads();
    if (threadIdx.x < C) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += sm[w][threadIdx.x];
        dv[(int64_t)o * C + threadIdx.x] = s;
    }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// pre-activations and the hidden state are activations (T); the cell state and the saved gates stay fp32
template <typename T>
__global__ void lstm_gates_fwd_kernel(const T* __restrict__ pre_x, const T* __restrict__ pre_h,
                                      const float* __restrict__ c_prev, float* __restrict__ gates,
                                      float* __restrict__ c_out, T* __restrict__ h_out, int64_t rows, int hid) {
    int64_t total = rows * hid;
    GRID_STRIDE(t, total) {
        int k = (int)(t % hid);
        int64_t r = t / hid;
        int64_t base = r * 4 * hid + k;
        float pi = ldf(pre_x + base), pf = ldf(pre_x + base + hid), po = ldf(pre_x + base + 2 * hid),
              pg = ldf(pre_x + base + 3 * hid);
        if (pre_h) {
            pi += ldf(pre_h + base); pf += ldf(pre_h + base + hid); po += ldf(pre_h + base + 2 * hid);
            pg += ldf(pre_h + base + 3 * hid);
        }
        float i = sigmoidf_(pi), f = sigmoidf_(pf), o = sigmoidf_(po), g = tanhf(pg);
        float cp = c_prev ? c_prev[t] : 0.f;
        float cn = f * cp + i * g;
        gates[base] = i; gates[base + hid] = f; gates[base + 2 * hid] = o; gates[base + 3 * hid] = g;
        c_out[t] = cn;
        stf(h_out + t, o * tanhf(cn));
    }
}

template <typename T>
__global__ void lstm_gates_bwd_kernel(const T* __restrict__ dh, const float* __restrict__ dc_next,
                                      const float* __restrict__ gates, const float* __restrict__ c_prev,
                                      const float* __restrict__ c_out, T* __restrict__ dpre,
                                      float* __restrict__ dc_prev, int64_t rows, int hid) {
    int64_t total = rows * hid;
    GRID_STRIDE(t, total) {
        int k = (int)(t % hid);
        int64_t r = t / hid;
        int64_t base = r * 4 * hid + k;
        float i = gates[base], f = gates[base + hid], o = gates[base + 2 * hid], g = gates[base + 3 * hid];
        float tc = tanhf(c_out[t]);
        float dhv = ldf(dh + t);
        float dc = (dc_next ? dc_next[t] : 0.f) + dhv * o * (1.f - tc * tc);
        float cp = c_prev ? c_prev[t] : 0.f;
        stf(dpre + base, dc * g * i * (1.f - i));
        stf(dpre + base + hid, dc * cp * f * (1.f - f));
        stf(dpre + base + 2 * hid, dhv * tc * o * (1.f - o));
        stf(dpre + base + 3 * hid, dc * i * (1.f - g * g));
        dc_prev[t] = dc * f;
    }
}

__global__ void reparam_fwd_kernel(const float* mu, const float* logvar, const float* eps, float* z, int64_t n) {
    GRID_STRIDE(t, n) { z[t] = eps[t] * expf(logvar[t] * 0.5f) + mu[t]; }
}
__global__ void reparam_bwd_kernel(const float* dz, const float* logvar, const float* eps, float* dmu, float* dlogvar,
                                   int64_t n) {
    GRID_STRIDE(t, n) {
        float g = dz[t];
        dmu[t] = g;
        dlogvar[t] = g * eps[t] * 0.5f * expf(logvar[t] * 0.5f);
    }
}

// column sums: stage 1 per-chunk partial (double), stage 2 fixed-order combine
template <typename T>
__global__ void colsum_partial_kernel(const T* __restrict__ x, int64_t rows, int C, int64_t rows_per_chunk,
                                      double* __restrict__ ws) {
    __shared__ double red[8][33];
    int c = blockIdx.y * 32 + threadIdx.x;
    int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
    int64_t r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
    double acc = 0.0;
    if (c < C)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) acc += (double)ldf(x + r * C + c);
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
        ws[(int64_t)blockIdx.x * C + c] = s;
    }
}
// vector variant of the column sums: a block covers CT*V channels (CT channel threads, a power of two <= 16) x 256/CT
// rows per pass; per-thread fp64 accumulation, fixed-order tree (warp shuffles across the lanes sharing a channel slot,
// then the 8 warps through shared memory)
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_vec_kernel(const T* __restrict__ x, int64_t rows, int C,
                                                                int64_t rows_per_chunk, int ct,
                                                                double* __restrict__ ws) {
    constexpr int V = VecIO<T>::V;
    __shared__ double sm[8 * 16 * V];
    const int c0 = blockIdx.y * ct * V;
    const int c = c0 + (threadIdx.x % ct) * V;
    const int rstep = 256 / ct;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_chunk;
    const int64_t r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
    double s[V];
#pragma unroll
    for (int i = 0; i < V; ++i) s[i] = 0.0;
#pragma unroll 4
    for (int64_t r = r0 + threadIdx.x / ct; r < r1; r += rstep) {
        float v[V];
        VecIO<T>::load(x + r * C + c, v);
#pragma unroll
        for (int i = 0; i < V; ++i) s[i] += (double)v[i];
    }
    for (int o = 16; o >= ct; o >>= 1) {
#pragma unroll
        for (int i = 0; i < V; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < ct) {
#pragma unroll
        for (int i = 0; i < V; ++i) sm[(warp * 16 + lane) * V + i] = s[i];
    }
    __syncthreads();
    if (threadIdx.x < ct * V) {
        const int slot = threadIdx.x / V, i = threadIdx.x % V;
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[(w * 16 + slot) * V + i];
        ws[(int64_t)blockIdx.x * C + c0 + threadIdx.x] = t;
    }
}

// out[r] = sum_l x[r*L + l]: one block per row (bias gradient of an NCHW output: rows = (n, c), L = H*W), fixed order
template <typename T>
__global__ void __launch_bounds__(256) rowsum_kernel(const T* __restrict__ x, int64_t L, float* __restrict__ out) {
    __shared__ double red[8];
    const T* p = x + (int64_t)blockIdx.x * L;
    double acc = 0.0;
    for (int64_t l = threadIdx.x; l < L; l += 256) acc += (double)ldf(p + l);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += red[k];
        out[blockIdx.x] = (float)s;
    }
}

// one warp per column: lanes stride over the chunks, fixed-order shuffle tree
__global__ void colsum_final_kernel(const double* __restrict__ ws, int nchunks, int C, float* __restrict__ out) {
    int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    int lane = threadIdx.x & 31;
    double s = 0.0;
#pragma unroll 4
    for (int k = lane; k < nchunks; k += 32) s += ws[(int64_t)k * C + c];
    s = warp_sum(s);
    if (lane == 0) out[c] = (float)s;
}

__device__ __forceinline__ uint16_t f32_to_bf16_rn(float v) {
    uint32_t u = __float_as_uint(v);
    uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
    if ((u & 0x7F800000u) == 0x7F800000u) r = u;  // inf / nan pass through
    return (uint16_t)(r >> 16);
}

// full entry (C_dst == C, c_off == 0): every element of the Mpad x ldw matrix is written (zeros in the padding);
// partial entry (a channel slice [c_off, c_off + C) of a wider matrix): only its valid elements are written
__global__ void pack_weight_kernel(const float* __restrict__ src, void* __restrict__ dst, int dst_bf16, int M, int Mpad,
                                   int Th, int Tw, int C, int64_t ldw, int64_t s_m, int64_t s_ky, int64_t s_kx,
                                   int64_t s_c, int ky0, int kx0, int kstep, int C_dst, int c_off) {
    int64_t total = (int64_t)Mpad * ldw;
    const int K = Th * Tw * C_dst;
    const bool full = (C_dst == C && c_off == 0);
    GRID_STRIDE(t, total) {
        int64_t m = t / ldw;
        int k = (int)(t % ldw);
        float v = 0.f;
        bool valid = false;
        if (m < M && k < K) {
            int c = k % C_dst - c_off;
            int tap = k / C_dst;
            if (c >= 0 && c < C) {
                int i = tap % Tw, j = tap / Tw;
                v = src[m * s_m + (int64_t)(ky0 + kstep * j) * s_ky + (int64_t)(kx0 + kstep * i) * s_kx + (int64_t)c * s_c];
                valid = true;
            }
        }
        if (!valid && !full) continue;
        if (dst_bf16) reinterpret_cast<uint16_t*>(dst)[t] = f32_to_bf16_rn(v);
        else reinterpret_cast<float*>(dst)[t] = v;
    }
}

// every packed operand of a network in one launch.  Work unit ("chunk") = a tile of 8 rows (m) x 64 channels (c) x all taps
// of one entry, staged through shared memory so that BOTH sides are coalesced: the parameter is read in contiguous runs
// (the (c, tap) plane of a row when s_c is the tap count — forward operands — or the (m, tap) plane of a channel when s_m is
// — data-gradient operands, whose rows are the convolution's INPUT channels), the operand matrix is written in 64-element
// row segments.  Entries with another stride pattern (none on the path) take the element-wise fallback.
// Block b -> (entry, tile) by binary search over the chunk prefix sums; same element mapping as pack_weight_kernel.
constexpr int kPackMT = B200_PACK_MT, kPackCT = B200_PACK_CT;
constexpr int kPackSmemFloats = 11264;                       // 44 KB staging tile (+ 3.2 KB of index tables): wider kernels go in row sub-passes

__global__ void __launch_bounds__(256) pack_weight_multi_kernel(const b200_pack_entry* __restrict__ entries, int n_entries,
                                                               const int32_t* __restrict__ chunk_entry) {
    __shared__ float tile[kPackSmemFloats];
    const int b = blockIdx.x;
    int lo;
    if (chunk_entry) {
        lo = chunk_entry[b];                            // host-built block -> entry map: one load instead of a search
    } else {
        lo = 0;
        int hi = n_entries - 1;
        while (lo < hi) {                               // last entry with chunk_begin <= b
            int mid = (lo + hi + 1) >> 1;
            if (entries[mid].chunk_begin <= b) lo = mid; else hi = mid - 1;
        }
    }
    const b200_pack_entry en = entries[lo];
    const int T = en.Th * en.Tw;
    const int ctiles = (en.C + kPackCT - 1) / kPackCT;
    const int tb = b - en.chunk_begin;
    const int m0 = (tb / ctiles) * kPackMT, c0 = (tb % ctiles) * kPackCT;
    const int mt = min(kPackMT, en.M - m0), ct = min(kPackCT, en.C - c0);
    // full tap count of the parameter (the contiguous innermost run): s_c for forward operands, s_m for data-gradient ones
    const int64_t kk = en.s_c < en.s_m ? en.s_c : en.s_m;
    const bool inner_c = en.s_c == kk;                   // (c, tap) contiguous per row m; else (m, tap) contiguous per channel c
    const bool staged = en.s_kx == 1 && kk <= 25 && T <= 32 && (kk | 1) * kPackCT <= kPackSmemFloats &&
                        kk >= (int64_t)(en.ky0 + en.kstep * (en.Th - 1)) * en.s_ky + en.kx0 + en.kstep * (en.Tw - 1) + 1;
    if (staged) {
        const int K = (int)kk;
        const int Kp = K | 1;                                    // odd tap pitch: the transposed shared reads spread over the banks
        const int mp = min(mt, (kPackSmemFloats - kPackCT) / (ct * Kp));     // rows per sub-pass (>= 1)
        // index tables (integer divisions once per block, not per element): position of run element r inside a staged slab,
        // and per selected tap its source offset inside the K taps of the parameter
        __shared__ uint16_t rpos[kPackCT * 25];                  // run <= 64 channels (or <= 32 rows) x 25 taps
        __shared__ uint8_t tapk[32];
        const int inner_max = inner_c ? ct : min(mp, mt);
        for (int r = threadIdx.x; r < inner_max * K; r += 256) {
            const int in = r / K, k = r - in * K;
            rpos[r] = (uint16_t)(in * Kp + k);
        }
        for (int tap = threadIdx.x; tap < T; tap += 256) {
            const int j = tap / en.Tw, ii = tap - j * en.Tw;
            tapk[tap] = (uint8_t)((en.ky0 + en.kstep * j) * (int)en.s_ky + (en.kx0 + en.kstep * ii));
        }
        __syncthreads();
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int ms = 0; ms < mt; ms += mp) {
            const int mc = min(mp, mt - ms);
            // load: `outer` contiguous runs of inner * K floats -> tile[outer][inner][Kp]; one warp per run, lanes along it
            const int outer = inner_c ? mc : ct, inner = inner_c ? ct : mc, run = inner * K;
            const int P = (inner * Kp) | 1;                      // odd pitch between outer slabs as well
            const int64_t s_outer = inner_c ? en.s_m : en.s_c;
            const float* base = en.src + (int64_t)(m0 + ms) * en.s_m + (int64_t)c0 * en.s_c;
            for (int o = warp; o < outer; o += 8) {
                const float* src = base + (int64_t)o * s_outer;
                float* dstt = tile + o * P;
                for (int r = lane; r < run; r += 32) dstt[rpos[r]] = src[r];
            }
            __syncthreads();
            // store: rows (m, tap), ct contiguous channels per row: 4 rows per pass, 64 threads along the channels
            const int c = threadIdx.x & (kPackCT - 1);
            if (c < ct) {
                for (int m = 0; m < mc; ++m) {
                    const int64_t drow = (int64_t)(m0 + ms + m) * en.ldw + en.c_off + c0 + c;
                    const float* tsrc = inner_c ? tile + m * P + c * Kp : tile + c * P + m * Kp;
                    for (int tap = threadIdx.x >> 6; tap < T; tap += 4) {
                        const float v = tsrc[tapk[tap]];
                        const int64_t d = drow + (int64_t)tap * en.C_dst;
                        if (en.dst_bf16) reinterpret_cast<uint16_t*>(en.dst)[d] = f32_to_bf16_rn(v);
                        else reinterpret_cast<float*>(en.dst)[d] = v;
                    }
                }
            }
            __syncthreads();
        }
        return;
    }
    for (int i = threadIdx.x; i < mt * T * ct; i += 256) {
        const int c = i % ct, rw = i / ct;
        const int m = rw / T, tap = rw - m * T;
        const int j = tap / en.Tw, ii = tap - j * en.Tw;
        const float v = en.src[(int64_t)(m0 + m) * en.s_m + (int64_t)(en.ky0 + en.kstep * j) * en.s_ky +
                               (int64_t)(en.kx0 + en.kstep * ii) * en.s_kx + (int64_t)(c0 + c) * en.s_c];
        const int64_t d = (int64_t)(m0 + m) * en.ldw + (int64_t)tap * en.C_dst + en.c_off + c0 + c;
        if (en.dst_bf16) reinterpret_cast<uint16_t*>(en.dst)[d] = f32_to_bf16_rn(v);
        else reinterpret_cast<float*>(en.dst)[d] = v;
    }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int64_t split_stride, int M, int Th, int Tw, int C,
                                    float* __restrict__ dst, int64_t s_m, int64_t s_ty, int64_t s_tx, int64_t s_c,
                                    const float* __restrict__ scale, int accumulate) {
    int64_t K = (int64_t)Th * Tw * C;
    int64_t total = (int64_t)M * K;
    float alpha = scale ? *scale : 1.f;
    GRID_STRIDE(t, total) {
        int64_t m = t / K;
        int k = (int)(t % K);
        int c = k % C;
        int tap = k / C;
        int tx = tap % Tw, ty = tap / Tw;
        float acc = 0.f;
        for (int s = 0; s < splits; ++s) acc += ws[(int64_t)s * split_stride + t];
        float* p = dst + m * s_m + (int64_t)ty * s_ty + (int64_t)tx * s_tx + (int64_t)c * s_c;
        *p = accumulate ? (*p + alpha * acc) : alpha * acc;
    }
}

// 4 consecutive channels per thread (16-byte loads of the partials) and 4 split lanes per element: lane = (element e =
// lane & 7, split lane sl = lane >> 3); split lane sl sums splits sl, sl + 4, ... in order, the four partial sums are
// combined as (p0 + p2) + (p1 + p3) by two shuffles — a fixed order.  C % 4 == 0, total < 2^31.
__global__ void wgrad_reduce_vec_kernel(const float* __restrict__ ws, int splits, int64_t split_stride, int M, int Th, int Tw,
                                        int C, float* __restrict__ dst, int64_t s_m, int64_t s_ty, int64_t s_tx,
                                        int64_t s_c, const float* __restrict__ scale, int accumulate) {
    const uint32_t K = (uint32_t)(Th * Tw * C);
    const uint32_t total4 = (uint32_t)M * K / 4;
    const float alpha = scale ? *scale : 1.f;
    const int lane = threadIdx.x & 31, e = lane & 7, sl = lane >> 3;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5);
    const uint32_t groups = (total4 + 7) / 8;                 // 8 elements per warp pass
    for (uint32_t gidx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); gidx < groups; gidx += warps) {
        const uint32_t t4 = gidx * 8 + e;
        const bool live = t4 < total4;
        const uint32_t t = (live ? t4 : 0) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* p = ws + t;
        int s = sl;
        for (; s + 12 < splits; s += 16) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(p + (int64_t)(s + 4 * u) * split_stride);
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; s < splits; s += 4) {
            const float4 v = *reinterpret_cast<const float4*>(p + (int64_t)s * split_stride);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
#pragma unroll
        for (int o = 16; o >= 8; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        if (live && sl == 0) {
            const uint32_t m = t / K, k = t - m * K;
            const uint32_t tap = k / (uint32_t)C, c = k - tap * (uint32_t)C;
            const uint32_t ty = tap / (uint32_t)Tw, tx = tap - ty * (uint32_t)Tw;
            float* q = dst + (int64_t)m * s_m + (int64_t)ty * s_ty + (int64_t)tx * s_tx + (int64_t)c * s_c;
            const float o4[4] = {alpha * acc.x, alpha * acc.y, alpha * acc.z, alpha * acc.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i * s_c] = accumulate ? q[i * s_c] + o4[i] : o4[i];
        }
    }
}

// Transposing variant for the parameter layout (M, C, Th, Tw) (s_c = Th*Tw, s_ty = Tw, s_tx = 1): the partial results are
// (M, tap, C).  A block owns 32 channels x R rows (m) x all taps per pass: one (row, tap, channel quad) item per thread sums
// the splits with 16-byte loads (4 independent accumulators in flight, combined in a fixed order), the tile is transposed
// through shared memory and written — contiguous in the destination — with coalesced stores.
__global__ void __launch_bounds__(256) wgrad_reduce_tr_kernel(const float* __restrict__ ws, int splits, int64_t split_stride,
                                                             int M, int T, int C, int R, float* __restrict__ dst, int64_t s_m,
                                                             const float* __restrict__ scale, int accumulate) {
    extern __shared__ float wr_tile[];                             // [R][32][T]
    const int c0 = blockIdx.x * 32;
    const float alpha = scale ? *scale : 1.f;
    const int nvalid = C - c0 < 32 ? C - c0 : 32;                 // a multiple of 4
    const int per_row = T * 8;                                     // (tap, channel quad)
    const int items = R * per_row, cells = R * 32 * T;
    for (int m0 = blockIdx.y * R; m0 < M; m0 += gridDim.y * R) {
        for (int it = threadIdx.x; it < items; it += 256) {
            const int r = it / per_row, rem = it - r * per_row;
            const int tap = rem >> 3, q = rem & 7;
            const int m = m0 + r;
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
            if (m < M && q * 4 < nvalid) {
                const float* p = ws + ((int64_t)m * T + tap) * C + c0 + q * 4;
                int sidx = 0;
                for (; sidx + 4 <= splits; sidx += 4) {
                    const float4 v0 = *reinterpret_cast<const float4*>(p + (int64_t)sidx * split_stride);
                    const float4 v1 = *reinterpret_cast<const float4*>(p + (int64_t)(sidx + 1) * split_stride);
                    const float4 v2 = *reinterpret_cast<const float4*>(p + (int64_t)(sidx + 2) * split_stride);
                    const float4 v3 = *reinterpret_cast<const float4*>(p + (int64_t)(sidx + 3) * split_stride);
                    a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
                    a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
                    a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
                    a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
                }
                for (; sidx < splits; ++sidx) {
                    const float4 v0 = *reinterpret_cast<const float4*>(p + (int64_t)sidx * split_stride);
                    a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
                }
                a0.x += a1.x + (a2.x + a3.x); a0.y += a1.y + (a2.y + a3.y);
                a0.z += a1.z + (a2.z + a3.z); a0.w += a1.w + (a2.w + a3.w);
            }
            float* t = wr_tile + (r * 32 + q * 4) * T + tap;
            t[0] = a0.x; t[T] = a0.y; t[2 * T] = a0.z; t[3 * T] = a0.w;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < cells; i += 256) {
            const int r = i / (32 * T), j = i - r * 32 * T;
            if (m0 + r < M && j < nvalid * T) {
                float* out = dst + (int64_t)(m0 + r) * s_m + (int64_t)c0 * T + j;
                const float v = alpha * wr_tile[i];
                *out = accumulate ? *out + v : v;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// im2col packing of a FEW-channel convolution input (the 3-channel image / crop convolutions): row m = output pixel
// (n, qy, qx), column k = (ky*kw + kx)*Cx + c, bf16, zero padded to Kp (a multiple of 64).  The packed matrix is the
// channel-last activation of an equivalent 1x1 convolution with Kp input channels, which runs on the tcgen05
// gather-GEMMs (forward and weight gradient) — K = Cx*kh*kw is far too small per tap to feed them directly.
// blockDim = (Kp/8, rows per block): a thread owns 8 consecutive columns (fixed taps), 16-byte stores.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void im2col_pack_kernel(const T* __restrict__ x, int64_t M, int Hx, int Wx, int Cx, int64_t sn, int64_t sh,
                                   int64_t sw, int64_t sc, int kh, int kw, int stride, int pad, int Hy, int Wy, int Kp,
                                   int flip, bf16* __restrict__ out) {
    const int kg = threadIdx.x;
    int64_t off[8];
    int dy[8], dx[8];
    const int K = Cx * kh * kw;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = kg * 8 + j;
        if (k < K) {
            const int tap = k / Cx, c = k - tap * Cx;
            const int ky = tap / kw, kx = tap - ky * kw;
            dy[j] = flip ? pad - ky : ky - pad;      // flip: the transposed (data-gradient) window walk
            dx[j] = flip ? pad - kx : kx - pad;
            off[j] = (int64_t)c * sc;
        } else {
            dy[j] = -(1 << 28);          // always out of bounds -> zero
            dx[j] = 0;
            off[j] = 0;
        }
    }
    for (int64_t m = (int64_t)blockIdx.x * blockDim.y + threadIdx.y; m < M; m += (int64_t)gridDim.x * blockDim.y) {
        const int mi = (int)m;
        const int qx = mi % Wy;
        const int t = mi / Wy;
        const int qy = t % Hy;
        const int n = t / Hy;
        const T* xb = x + (int64_t)n * sn;
        const int iy0 = qy * stride, ix0 = qx * stride;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int iy = iy0 + dy[j], ix = ix0 + dx[j];
            const bool ok = iy >= 0 && iy < Hx && ix >= 0 && ix < Wx;
            v[j] = ok ? ldf(xb + off[j] + (int64_t)iy * sh + (int64_t)ix * sw) : 0.f;
        }
        uint4 u;
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]),
                       c2 = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
        u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
        u.z = *reinterpret_cast<uint32_t*>(&c2); u.w = *reinterpret_cast<uint32_t*>(&d);
        *reinterpret_cast<uint4*>(out + m * Kp + kg * 8) = u;
    }
}

// Staged variant (unit input x-stride, i.e. NCHW inputs): one warp packs a strip of 32 consecutive output pixels.
// Phase 1: lane = pixel, loop over the K columns (warp-uniform tap / channel): consecutive lanes read consecutive input
// columns (coalesced), results go to shared memory as bf16 [32][Kp + 8].  Phase 2: the strip's 32 x Kp block is
// contiguous in the output: 16-byte shared loads, fully coalesced 16-byte stores.
template <typename T>
__global__ void __launch_bounds__(128) im2col_pack_staged_kernel(const T* __restrict__ x, int64_t M, int Hx, int Wx, int Cx,
                                                                int64_t sn, int64_t sh, int64_t sw, int64_t sc, int kh,
                                                                int kw, int stride, int pad, int Hy, int Wy, int Kp,
                                                                int flip, bf16* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t pack_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ld = Kp + 8;                                   // row pitch in bf16: 16-byte aligned rows, spread banks
    bf16* tile = reinterpret_cast<bf16*>(pack_smem) + (size_t)warp * 32 * ld;
    const int64_t strips = (M + 31) / 32;
    for (int64_t sidx = (int64_t)blockIdx.x * 4 + warp; sidx < strips; sidx += (int64_t)gridDim.x * 4) {
        const int64_t m = sidx * 32 + lane;
        const bool live = m < M;
        int qx = 0, qy = 0, n = 0;
        if (live) {
            const int mi = (int)m;
            qx = mi % Wy;
            const int t = mi / Wy;
            qy = t % Hy;
            n = t / Hy;
        }
        const T* xb = x + (int64_t)n * sn;
        const int iy0 = qy * stride + (flip ? pad : -pad), ix0 = qx * stride + (flip ? pad : -pad);
        // two consecutive K columns per step -> one 32-bit shared store (half the stores / bank conflicts of bf16 stores)
        uint32_t* row2 = reinterpret_cast<uint32_t*>(tile + lane * ld);
        const int K = kh * kw * Cx;
        int ky = 0, kx = 0, c = 0;
        for (int k = 0; k < Kp; k += 2) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                v[e] = 0.f;
                if (k + e < K) {
                    const int iy = flip ? iy0 - ky : iy0 + ky, ix = flip ? ix0 - kx : ix0 + kx;
                    if (live && iy >= 0 && iy < Hx && ix >= 0 && ix < Wx)
                        v[e] = ldf(xb + (int64_t)iy * sh + (int64_t)ix * sw + (int64_t)c * sc);
                    if (++c == Cx) { c = 0; if (++kx == kw) { kx = 0; ++ky; } }
                }
            }
            __nv_bfloat162 pr = __floats2bfloat162_rn(v[0], v[1]);
            row2[k >> 1] = *reinterpret_cast<uint32_t*>(&pr);
        }
        __syncwarp();
        const int cpr = Kp / 8;                               // 16-byte chunks per row
        const int64_t m0 = sidx * 32;
        const int rows = (int)(M - m0 < 32 ? M - m0 : 32);
        for (int j = lane; j < rows * cpr; j += 32) {
            const int r = j / cpr, q = j - r * cpr;
            *reinterpret_cast<uint4*>(out + (m0 + r) * Kp + q * 8) = *reinterpret_cast<const uint4*>(tile + r * ld + q * 8);
        }
        __syncwarp();
    }
}

// Patch variant (unit input x-stride, stride-1 window walk, output rows made of whole 32-pixel strips — every few-channel
// layer of the training step): the staged kernel above is instruction bound (one global load, bounds test and index update
// per matrix element: ncu issue-active 75-78 %, profiles/r02m_mem_kernels.md).  Here a warp first copies the strip's INPUT
// patch — kh rows x (32 + kw - 1) columns x Cx channels, channel-interleaved bf16 — into shared memory (kh*Cx coalesced row
// reads, zero outside the image), then every 16-byte chunk of the strip's contiguous 32 x Kp output block is gathered from
// the patch through a per-launch column -> patch-offset table (pixel r adds r*Cx).  ~3x fewer instructions per strip.
template <typename T>
__global__ void __launch_bounds__(128) im2col_pack_patch_kernel(const T* __restrict__ x, int64_t M, int Hx, int Wx, int Cx,
                                                               int64_t sn, int64_t sh, int64_t sc, int kh, int kw, int pad,
                                                               int Hy, int Wy, int Kp, int flip, bf16* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t pack_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int RW = 32 + kw - 1;                                   // patch columns
    const int K = kh * kw * Cx;
    const int row_elems = RW * Cx;
    const int patch_pitch = (kh * row_elems + 7) & ~7;
    uint16_t* koff = reinterpret_cast<uint16_t*>(pack_smem);      // [Kp]: patch offset of column k for pixel 0 (0xFFFF = padding)
    uint16_t* patch = koff + Kp + (size_t)warp * patch_pitch;
    for (int k = threadIdx.x; k < Kp; k += 128) {
        uint16_t v = 0xFFFFu;
        if (k < K) {
            const int tap = k / Cx, c = k - tap * Cx;
            const int ky = tap / kw, kx = tap - ky * kw;
            v = (uint16_t)((ky * RW + (flip ? kw - 1 - kx : kx)) * Cx + c);
        }
        koff[k] = v;
    }
    __syncthreads();
    const int cpr = Kp >> 3;                                      // 16-byte chunks per output row
    const int64_t strips = M >> 5;
    for (int64_t sidx = (int64_t)blockIdx.x * 4 + warp; sidx < strips; sidx += (int64_t)gridDim.x * 4) {
        const int mi = (int)(sidx << 5);
        const int qx0 = mi % Wy;
        const int t0 = mi / Wy;
        const int qy = t0 % Hy;
        const int n = t0 / Hy;
        const T* xb = x + (int64_t)n * sn;
        // patch column j holds input column ixb + j; pixel r, tap kx reads column r + kx (flip: r + kw - 1 - kx)
        const int ixb = flip ? qx0 + pad - (kw - 1) : qx0 - pad;
        for (int ky = 0; ky < kh; ++ky) {
            const int iy = flip ? qy + pad - ky : qy - pad + ky;
            const bool rowok = iy >= 0 && iy < Hx;
            const T* src = xb + (int64_t)iy * sh;
            uint16_t* prow = patch + ky * row_elems;
            int c = 0, j = lane;                                   // flattened (c, j) walk, lanes along the input row
            while (j >= RW) { j -= RW; ++c; }
            while (c < Cx) {
                const int ix = ixb + j;
                const float v = (rowok && ix >= 0 && ix < Wx) ? ldf(src + (int64_t)c * sc + ix) : 0.f;
                prow[j * Cx + c] = f32_to_bf16_rn(v);
                j += 32;
                while (j >= RW) { j -= RW; ++c; }
            }
        }
        __syncwarp();
        bf16* o = out + ((int64_t)sidx << 5) * Kp;
        int r = 0, q = lane;
        while (q >= cpr) { q -= cpr; ++r; }
        for (int t = lane; t < 32 * cpr; t += 32) {
            const uint4 kt = *reinterpret_cast<const uint4*>(koff + q * 8);
            const uint16_t* pr = patch + r * Cx;
            const uint32_t ks[4] = {kt.x, kt.y, kt.z, kt.w};
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t k0 = ks[e] & 0xFFFFu, k1 = ks[e] >> 16;
                const uint32_t lo = k0 == 0xFFFFu ? 0u : pr[k0];
                const uint32_t hi = k1 == 0xFFFFu ? 0u : pr[k1];
                w[e] = lo | (hi << 16);
            }
            *reinterpret_cast<uint4*>(o + (int64_t)t * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            q += 32;
            while (q >= cpr) { q -= cpr; ++r; }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// avg_pool2(conv_{kh x kw, stride 1, pad p}(x; W)) == conv_{(kh+1) x (kw+1), stride 2, pad p}(x; W4) with
//     W4[f][a][b] = 0.25 * sum_{i,j in {0,1}} W[f][a-i][b-j]        (terms outside the kh x kw window dropped)
// (both are linear in x and the pooled output pixel (Y, X) averages the stride-1 outputs (2Y+i, 2X+j)).  The discriminator
// blocks' second convolution is followed by that pooling (discriminator.py:52-57, 90-96): the folded form does 16/36 of the
// multiply-adds of a 3x3 layer and never materialises the full-resolution output.  f = (cout, cin) filter index.
// ---------------------------------------------------------------------------------------------------------
__global__ void fold_pool_weight_kernel(const float* __restrict__ w, float* __restrict__ w4, int64_t filters, int kh, int kw) {
    const int kh4 = kh + 1, kw4 = kw + 1;
    const int64_t n = filters * kh4 * kw4;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = t / (kh4 * kw4);
        const int r = (int)(t - f * (kh4 * kw4));
        const int a = r / kw4, b = r - a * kw4;
        const float* wf = w + f * kh * kw;
        float acc = 0.f;                      // fixed order: (i, j) = (0,0), (0,1), (1,0), (1,1)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int k = a - i, l = b - j;
                if (k >= 0 && k < kh && l >= 0 && l < kw) acc += wf[k * kw + l];
            }
        w4[t] = 0.25f * acc;
    }
}

// y[b][c][r] = x[b][r][c]  (batched R x C -> C x R transpose; NCHW <-> channel-last)
__global__ void transpose_kernel(const float* __restrict__ x, float* __restrict__ y, int R, int C) {
    __shared__ float tile[32][33];
    const float* xb = x + (int64_t)blockIdx.z * R * C;
    float* yb = y + (int64_t)blockIdx.z * R * C;
    int c = blockIdx.x * 32 + threadIdx.x;
    for (int k = threadIdx.y; k < 32; k += 8) {
        int r = blockIdx.y * 32 + k;
        if (r < R && c < C) tile[k][threadIdx.x] = xb[(int64_t)r * C + c];
    }
    __syncthreads();
    int r2 = blockIdx.y * 32 + threadIdx.x;
    for (int k = threadIdx.y; k < 32; k += 8) {
        int c2 = blockIdx.x * 32 + k;
        if (r2 < R && c2 < C) yb[(int64_t)c2 * R + r2] = tile[threadIdx.x][k];
    }
}

}  // namespace b200

using namespace b200;

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <typename T>
static inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & (4 * sizeof(T) - 1)) == 0; }

extern "C" int b200_relu_fwd(const void* x, void* y, int64_t n, int dt, b200_stream_t stream) {
    if (n == 0) return 0;
    B200_DISPATCH_DT(dt, T, {
        int64_t n4 = (aligned16(x) && aligned16(y)) ? n / VecIO<T>::V : 0;
        int tail = (int)(n - n4 * VecIO<T>::V);
        B200_REQUIRE(tail < 256, "relu_fwd: unaligned large tensor");
        relu_fwd_kernel<T><<<grid_for(n4 > 0 ? n4 : 1, 256), 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, n4, tail);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, int dt, b200_stream_t stream) {
    if (n == 0) return 0;
    B200_DISPATCH_DT(dt, T, {
        int64_t n4 = (aligned16(dy) && aligned16(y) && aligned16(dx)) ? n / VecIO<T>::V : 0;
        int tail = (int)(n - n4 * VecIO<T>::V);
        B200_REQUIRE(tail < 256, "relu_bwd: unaligned large tensor");
        relu_bwd_kernel<T><<<grid_for(n4 > 0 ? n4 : 1, 256), 256, 0, as_stream(stream)>>>((const T*)dy, (const T*)y, (T*)dx,
                                                                                      n4, tail);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

static int add_launch(const void* a, const void* b, void* out, int64_t n, int dt, int relu, b200_stream_t stream) {
    if (n == 0) return 0;
    B200_DISPATCH_DT(dt, T, {
        int64_t n4 = (aligned16(a) && aligned16(b) && aligned16(out)) ? n / VecIO<T>::V : 0;
        int tail = (int)(n - n4 * VecIO<T>::V);
        B200_REQUIRE(tail < 256, "add: unaligned large tensor");
        add_kernel<T><<<grid_for(n4 > 0 ? n4 : 1, 256), 256, 0, as_stream(stream)>>>((const T*)a, (const T*)b, (T*)out, n4,
                                                                                 tail, relu);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_add(const void* a, const void* b, void* out, int64_t n, int dt, b200_stream_t stream) {
    return add_launch(a, b, out, n, dt, 0, stream);
}

extern "C" int b200_add_relu(const void* a, const void* b, void* out, int64_t n, int dt, b200_stream_t stream) {
    return add_launch(a, b, out, n, dt, 1, stream);
}

static int pool_launch(const void* x, const void* x2, void* y, int N, int H, int W, int C, int f, float scale, int dt,
                       int relu, b200_stream_t stream) {
    B200_REQUIRE(f >= 1 && H % f == 0 && W % f == 0, "pool_fwd: H=%d W=%d not divisible by f=%d", H, W, f);
    int64_t total = (int64_t)N * (H / f) * (W / f) * C;
    if (total == 0) return 0;
    B200_DISPATCH_DT(dt, T, {
        constexpr int V = VecIO<T>::V;
        const bool small = (int64_t)N * H * W * C < (1ll << 31);
        const T* a = (const T*)x;
        const T* b = (const T*)x2;
        if (C % V == 0 && aligned16(x) && aligned16(y) && (x2 == nullptr || aligned16(x2))) {
            if (small) pool_fwd_kernel<T, V, uint32_t><<<grid_for(total / V, 256), 256, 0, as_stream(stream)>>>(a, b, (T*)y, N, H, W, C, f, scale, relu);
            else pool_fwd_kernel<T, V, int64_t><<<grid_for(total / V, 256), 256, 0, as_stream(stream)>>>(a, b, (T*)y, N, H, W, C, f, scale, relu);
        } else {
            if (small) pool_fwd_kernel<T, 1, uint32_t><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(a, b, (T*)y, N, H, W, C, f, scale, relu);
            else pool_fwd_kernel<T, 1, int64_t><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(a, b, (T*)y, N, H, W, C, f, scale, relu);
        }
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_pool_fwd(const void* x, void* y, int N, int H, int W, int C, int f, float scale, int dt,
                             b200_stream_t stream) {
    return pool_launch(x, nullptr, y, N, H, W, C, f, scale, dt, 0, stream);
}

extern "C" int b200_pool_add_fwd(const void* a, const void* b, void* y, int N, int H, int W, int C, int f, float scale,
                                 int relu, int dt, b200_stream_t stream) {
    B200_REQUIRE(a != nullptr && b != nullptr, "pool_add_fwd: two inputs required");
    return pool_launch(a, b, y, N, H, W, C, f, scale, dt, relu, stream);
}

static int unpool_launch(const void* x, const void* mask, void* y, int N, int H, int W, int C, int f, float scale, int dt,
                         b200_stream_t stream) {
    int64_t total = (int64_t)N * H * f * W * f * C;
    if (total == 0) return 0;
    B200_DISPATCH_DT(dt, T, {
        constexpr int V = VecIO<T>::V;
        const bool small = total < (1ll << 31);
        const T* xp = (const T*)x;
        const T* mp = (const T*)mask;
        if (C % V == 0 && aligned16(x) && aligned16(y) && (mask == nullptr || aligned16(mask))) {
            if (small) unpool_fwd_kernel<T, V, uint32_t><<<grid_for(total / V, 256), 256, 0, as_stream(stream)>>>(xp, mp, (T*)y, N, H, W, C, f, scale);
            else unpool_fwd_kernel<T, V, int64_t><<<grid_for(total / V, 256), 256, 0, as_stream(stream)>>>(xp, mp, (T*)y, N, H, W, C, f, scale);
        } else {
            if (small) unpool_fwd_kernel<T, 1, uint32_t><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(xp, mp, (T*)y, N, H, W, C, f, scale);
            else unpool_fwd_kernel<T, 1, int64_t><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(xp, mp, (T*)y, N, H, W, C, f, scale);
        }
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_unpool_fwd(const void* x, void* y, int N, int H, int W, int C, int f, float scale, int dt,
                               b200_stream_t stream) {
    return unpool_launch(x, nullptr, y, N, H, W, C, f, scale, dt, stream);
}

extern "C" int b200_unpool_masked_fwd(const void* x, const void* mask, void* y, int N, int H, int W, int C, int f, float scale,
                                      int dt, b200_stream_t stream) {
    B200_REQUIRE(mask != nullptr, "unpool_masked_fwd: mask required");
    return unpool_launch(x, mask, y, N, H, W, C, f, scale, dt, stream);
}

extern "C" int b200_concat_fwd(const void* a, int Ca, int a_div, const void* b, int Cb, int b_div, void* out,
                               int64_t rows, int dt, b200_stream_t stream) {
    if (rows == 0) return 0;
    B200_DISPATCH_DT(dt, T, {
        concat_fwd_kernel<T><<<grid_for(rows * (Ca + Cb), 256), 256, 0, as_stream(stream)>>>((const T*)a, Ca, a_div, (const T*)b,
                                                                                         Cb, b_div, (T*)out, rows);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_concat_bwd(const void* dout, int Ca, int a_div, void* da, int Cb, int b_div, void* db,
                               int64_t rows, int dt, b200_stream_t stream) {
    if (rows == 0) return 0;
    B200_REQUIRE(rows % a_div == 0 && rows % b_div == 0, "concat_bwd: rows not divisible by broadcast factors");
    B200_DISPATCH_DT(dt, T, {
        concat_bwd_kernel<T><<<grid_for(rows * (Ca + Cb), 256), 256, 0, as_stream(stream)>>>((const T*)dout, Ca, a_div, (T*)da,
                                                                                         Cb, b_div, (T*)db, rows);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_gather_rows(const float* table, const int32_t* idx, float* out, int rows, int D,
                                b200_stream_t stream) {
    if (rows == 0) return 0;
    gather_rows_kernel<<<grid_for((int64_t)rows * D, 256), 256, 0, as_stream(stream)>>>(table, idx, out, rows, D);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_scatter_rows(const float* dout, const int32_t* idx, float* dtable, int rows, int D,
                                 int num_classes, b200_stream_t stream) {
    scatter_rows_kernel<<<grid_for((int64_t)num_classes * D, 128), 128, 0, as_stream(stream)>>>(dout, idx, dtable, rows,
                                                                                                D, num_classes);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_permute_rows(const void* x, const int32_t* src_row, void* out, int rows, int64_t row_bytes,
                                 b200_stream_t stream) {
    if (rows == 0) return 0;
    B200_REQUIRE(row_bytes % 16 == 0 && aligned16(x) && aligned16(out), "permute_rows: row_bytes %% 16 != 0 or unaligned");
    permute_rows_kernel<<<grid_for((int64_t)rows * (row_bytes / 16), 256), 256, 0, as_stream(stream)>>>(
        (const uint4*)x, src_row, (uint4*)out, rows, (int)(row_bytes / 16));
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_mask_outer_fwd(const float* v, const float* mask, void* out, int O, int H, int W, int C, int dt,
                                   b200_stream_t stream) {
    if (O == 0) return 0;
    B200_REQUIRE(C % 8 == 0, "mask_outer_fwd: C=%d must be a multiple of 8", C);
    B200_REQUIRE((int64_t)O * (H + 2) * (W + 2) * (C / 4) < (1ll << 31), "mask_outer_fwd: tensor too large");
    B200_REQUIRE(aligned16(v) && aligned16(out), "mask_outer_fwd: operands must be 16-byte aligned");
    B200_DISPATCH_DT(dt, T, {
        int64_t total = (int64_t)O * (H + 2) * (W + 2) * (C / VecIO<T>::V);
        mask_outer_fwd_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(v, mask, (T*)out, O, H, W, C);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_mask_outer_bwd(const void* dout, const float* mask, float* dv, int O, int H, int W, int C, int dt,
                                   b200_stream_t stream) {
    if (O == 0) return 0;
    dim3 grid(O, (C + 31) / 32), block(32, 8);
    B200_DISPATCH_DT(dt, T, {
        constexpr int V = VecIO<T>::V;
        const int ct = C / V;
        if (C % V == 0 && ct >= 1 && ct <= 32 && (ct & (ct - 1)) == 0 && aligned16(dout))
            mask_outer_bwd_vec_kernel<T><<<O, 256, 0, as_stream(stream)>>>((const T*)dout, mask, dv, H, W, C);
        else
            mask_outer_bwd_kernel<T><<<grid, block, 0, as_stream(stream)>>>((const T*)dout, mask, dv, O, H, W, C);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_lstm_gates_fwd(const void* pre_x, const void* pre_h, const float* c_prev, float* gates,
                                   float* c_out, void* h_out, int64_t rows, int hid, int dt, b200_stream_t stream) {
    if (rows == 0) return 0;
    B200_DISPATCH_DT(dt, T, {
        lstm_gates_fwd_kernel<T><<<grid_for(rows * hid, 256), 256, 0, as_stream(stream)>>>(
            (const T*)pre_x, (const T*)pre_h, c_prev, gates, c_out, (T*)h_out, rows, hid);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_lstm_gates_bwd(const void* dh, const float* dc_next, const float* gates, const float* c_prev,
                                   const float* c_out, void* dpre, float* dc_prev, int64_t rows, int hid, int dt,
                                   b200_stream_t stream) {
    if (rows == 0) return 0;
    B200_DISPATCH_DT(dt, T, {
        lstm_gates_bwd_kernel<T><<<grid_for(rows * hid, 256), 256, 0, as_stream(stream)>>>(
            (const T*)dh, dc_next, gates, c_prev, c_out, (T*)dpre, dc_prev, rows, hid);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_reparam_fwd(const float* mu, const float* logvar, const float* eps, float* z, int64_t n,
                                b200_stream_t stream) {
    if (n == 0) return 0;
    reparam_fwd_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(mu, logvar, eps, z, n);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_reparam_bwd(const float* dz, const float* logvar, const float* eps, float* dmu_add,
                                float* dlogvar_add, int64_t n, b200_stream_t stream) {
    if (n == 0) return 0;
    reparam_bwd_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(dz, logvar, eps, dmu_add, dlogvar_add, n);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_bn_chunks(int64_t rows, int C) {
    int cblocks = (C + 31) / 32;
    int64_t target = (int64_t)kNumSMs * 8 / cblocks;
    if (target < 1) target = 1;
    int64_t by_rows = (rows + 31) / 32;
    int64_t n = by_rows < target ? by_rows : target;
    if (n < 1) n = 1;
    if (n > 1024) n = 1024;
    return (int)n;
}

extern "C" int b200_colsum(const void* x, int64_t rows, int C, int dt, float* out, double* ws, b200_stream_t stream) {
    int nchunks = b200_bn_chunks(rows, C);
    int64_t rpc = (rows + nchunks - 1) / nchunks;
    if (rpc < 1) rpc = 1;
    dim3 grid(nchunks, (C + 31) / 32), block(32, 8);
    B200_DISPATCH_DT(dt, T, {
        constexpr int V = VecIO<T>::V;
        const int tpr = C / V;
        if (C % V == 0 && tpr <= 256 && (tpr & (tpr - 1)) == 0 && aligned16(x)) {
            const int ct = tpr < 16 ? tpr : 16;
            colsum_partial_vec_kernel<T><<<dim3(nchunks, tpr / ct), 256, 0, as_stream(stream)>>>((const T*)x, rows, C, rpc, ct, ws);
        } else {
            colsum_partial_kernel<T><<<grid, block, 0, as_stream(stream)>>>((const T*)x, rows, C, rpc, ws);
        }
    });
    B200_CHECK_LAUNCH();
    colsum_final_kernel<<<(C + 7) / 8, 256, 0, as_stream(stream)>>>(ws, nchunks, C, out);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_rowsum(const void* x, int dt, int64_t rows, int64_t L, float* out, b200_stream_t stream) {
    if (rows == 0) return 0;
    B200_REQUIRE(rows < (1ll << 31) && L > 0, "rowsum: bad sizes");
    B200_DISPATCH_DT(dt, T, { rowsum_kernel<T><<<(unsigned)rows, 256, 0, as_stream(stream)>>>((const T*)x, L, out); });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_pack_weight(const float* src, void* dst, int dst_bf16, int M, int Mpad, int Th, int Tw, int C,
                                int64_t ldw, int64_t s_m, int64_t s_ky, int64_t s_kx, int64_t s_c, int ky0, int kx0,
                                int kstep, int C_dst, int c_off, b200_stream_t stream) {
    if (C_dst <= 0) { C_dst = C; c_off = 0; }
    B200_REQUIRE(ldw >= (int64_t)Th * Tw * C_dst && Mpad >= M && c_off >= 0 && c_off + C <= C_dst,
                 "pack_weight: ldw/Mpad too small or bad channel slice");
    int64_t total = (int64_t)Mpad * ldw;
    if (total == 0) return 0;
    pack_weight_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(src, dst, dst_bf16, M, Mpad, Th, Tw, C, ldw,
                                                                            s_m, s_ky, s_kx, s_c, ky0, kx0, kstep, C_dst,
                                                                            c_off);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_pack_weight_multi(const b200_pack_entry* entries_dev, int n_entries, int total_chunks,
                                      const int32_t* chunk_entry_dev, b200_stream_t stream) {
    if (n_entries <= 0 || total_chunks <= 0) return 0;
    B200_REQUIRE(entries_dev != nullptr, "pack_weight_multi: null table");
    pack_weight_multi_kernel<<<total_chunks, 256, 0, as_stream(stream)>>>(entries_dev, n_entries, chunk_entry_dev);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_wgrad_reduce(const float* ws, int splits, int64_t split_stride, int M, int Th, int Tw, int C,
                                 float* dst, int64_t s_m, int64_t s_ty, int64_t s_tx, int64_t s_c, const float* scale,
                                 int accumulate, b200_stream_t stream) {
    int64_t total = (int64_t)M * Th * Tw * C;
    if (total == 0) return 0;
    const int T = Th * Tw;
    if ((C & 3) == 0 && (split_stride & 3) == 0 && aligned16(ws) && T > 1 && T <= 64 && s_tx == 1 && s_ty == Tw && s_c == T &&
        M < 65536) {
        const int cb = (C + 31) / 32;
        int R = 256 / (T * 8);                               // rows per pass: about one (tap, channel quad) item per thread
        if (R < 1) R = 1;
        if (R > M) R = M;
        int rows = (kNumSMs * 8 + cb - 1) / cb;              // ~8 blocks per SM in total, each walking its share of the rows
        if (rows > (M + R - 1) / R) rows = (M + R - 1) / R;
        wgrad_reduce_tr_kernel<<<dim3(cb, rows), 256, (size_t)R * 32 * T * sizeof(float), as_stream(stream)>>>(
            ws, splits, split_stride, M, T, C, R, dst, s_m, scale, accumulate);
    } else if ((C & 3) == 0 && (split_stride & 3) == 0 && total < (1ll << 31) && aligned16(ws))
        wgrad_reduce_vec_kernel<<<grid_for(total / 4 * 4, 128, 16), 128, 0, as_stream(stream)>>>(ws, splits, split_stride, M, Th, Tw, C,
                                                                                        dst, s_m, s_ty, s_tx, s_c, scale, accumulate);
    else
        wgrad_reduce_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(ws, splits, split_stride, M, Th, Tw, C, dst,
                                                                                 s_m, s_ty, s_tx, s_c, scale, accumulate);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_im2col_pack(const void* x, int x_dt, int64_t N, int Hx, int Wx, int Cx, int64_t sn, int64_t sh,
                                int64_t sw, int64_t sc, int kh, int kw, int stride, int pad, int Hy, int Wy, int Kp,
                                int flip, void* out_bf16, b200_stream_t stream) {
    const int64_t M = N * Hy * Wy;
    if (M == 0) return 0;
    B200_REQUIRE(Kp % 64 == 0 && Kp >= Cx * kh * kw && Kp <= 512, "im2col_pack: Kp=%d must be a multiple of 64 covering K=%d", Kp, Cx * kh * kw);
    B200_REQUIRE(M < (1ll << 31), "im2col_pack: too many output pixels");
    const int patch_smem = Kp * 2 + 4 * (((kh * (32 + kw - 1) * Cx) + 7) & ~7) * 2;
    if (sw == 1 && stride == 1 && Wy % 32 == 0 && patch_smem <= 48 * 1024 && kh * (32 + kw - 1) * Cx < 0xFFFF) {
        // whole 32-pixel strips per output row: input patch staged in shared memory (see im2col_pack_patch_kernel)
        const int64_t strips = M / 32;
        const int grid = grid_for((strips + 3) / 4, 1, 12);
        B200_DISPATCH_DT(x_dt, T, {
            im2col_pack_patch_kernel<T><<<grid, 128, patch_smem, as_stream(stream)>>>((const T*)x, M, Hx, Wx, Cx, sn, sh, sc, kh, kw,
                                                                                   pad, Hy, Wy, Kp, flip, (bf16*)out_bf16);
        });
        B200_CHECK_LAUNCH();
        return 0;
    }
    if (sw == 1) {
        // NCHW-style input (unit stride along x): staged kernel, coalesced on both sides
        const int smem = 4 * 32 * (Kp + 8) * 2;
        const int64_t strips = (M + 31) / 32;
        const int grid = grid_for((strips + 3) / 4, 1, 12);      // ~48 warps per SM: the strip loop is latency bound
        B200_DISPATCH_DT(x_dt, T, {
            static bool configured = false;
            if (!configured) {
                cudaError_t e = cudaFuncSetAttribute(im2col_pack_staged_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32 * (512 + 8) * 2);
                if (e != cudaSuccess) return set_error("im2col_pack: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                configured = true;
            }
            im2col_pack_staged_kernel<T><<<grid, 128, smem, as_stream(stream)>>>((const T*)x, M, Hx, Wx, Cx, sn, sh, sw, sc, kh, kw,
                                                                              stride, pad, Hy, Wy, Kp, flip, (bf16*)out_bf16);
        });
        B200_CHECK_LAUNCH();
        return 0;
    }
    const int G = Kp / 8;
    const int rows = 256 / G > 0 ? 256 / G : 1;
    dim3 block(G, rows);
    const int grid = grid_for((M + rows - 1) / rows, 1, 16);
    B200_DISPATCH_DT(x_dt, T, {
        im2col_pack_kernel<T><<<grid, block, 0, as_stream(stream)>>>((const T*)x, M, Hx, Wx, Cx, sn, sh, sw, sc, kh, kw, stride,
                                                                  pad, Hy, Wy, Kp, flip, (bf16*)out_bf16);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_fold_pool_weight(const float* w, float* w4, int64_t filters, int kh, int kw, b200_stream_t stream) {
    if (filters == 0) return 0;
    B200_REQUIRE(kh >= 1 && kw >= 1 && kh <= 15 && kw <= 15, "fold_pool_weight: bad kernel size");
    const int64_t n = filters * (kh + 1) * (kw + 1);
    fold_pool_weight_kernel<<<grid_for(n, 256, 4), 256, 0, as_stream(stream)>>>(w, w4, filters, kh, kw);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_transpose(const float* x, float* y, int B, int R, int C, b200_stream_t stream) {
    if (B == 0 || R == 0 || C == 0) return 0;
    B200_REQUIRE(B < 65536, "transpose: batch too large");
    dim3 grid((C + 31) / 32, (R + 31) / 32, B), block(32, 8);
    B200_REQUIRE(grid.y < 65536, "transpose: R too large");
    transpose_kernel<<<grid, block, 0, as_stream(stream)>>>(x, y, R, C);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_copy(void* dst, const void* src, size_t bytes, b200_stream_t stream) {
    if (bytes == 0) return 0;
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream));
    if (e != cudaSuccess) return set_error("copy: %s", cudaGetErrorString(e));
    count_launch();
    return 0;
}
