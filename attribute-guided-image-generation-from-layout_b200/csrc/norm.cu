// norm.cu — batch-norm statistics (K6) and the fused normalise/modulate/ReLU/residual apply (K7), fwd + bwd.
// Replaces nn.BatchNorm1d/2d (F.batch_norm training semantics: biased var for normalisation, unbiased for
// running_var, momentum 0.1), ConditionalBatchNorm2d (generator_obj_att.py:31-44) and SPADE's modulation
// (models/spade/networks/normalization.py:94-108).  x is (rows, C) channel-last; all reductions are two-stage,
// fixed order, accumulated in fp64 (deterministic).
#include "common.cuh"

namespace b200 {

// the vector kernels take layouts where a 256-thread block covers whole rows: C/V threads per row, a power of two <= 256
template <typename T>
static inline bool vec_ok(int64_t rows, int C) {
    constexpr int V = VecIO<T>::V;
    if (C % V != 0 || rows >= (1ll << 31)) return false;
    const int tpr = C / V;
    return tpr >= 1 && tpr <= 256 && (tpr & (tpr - 1)) == 0;
}

// ---------------------------------------------------------------------------------------------------------
// statistics.  `groups` independent calls batched along the row dimension (rows = groups * rows_per_group) keep
// separate statistics: grid.z = group, chunks never straddle a group.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void bn_stats_partial_kernel(const T* __restrict__ x, int64_t rows_per_group, int C,
                                        int64_t rows_per_chunk, double* __restrict__ ws) {
    __shared__ double r1[8][33], r2[8][33];
    int c = blockIdx.y * 32 + threadIdx.x;
    int64_t g0 = (int64_t)blockIdx.z * rows_per_group;
    int64_t a = g0 + (int64_t)blockIdx.x * rows_per_chunk;
    int64_t e = g0 + rows_per_group;
    int64_t b = a + rows_per_chunk < e ? a + rows_per_chunk : e;
    double s1 = 0.0, s2 = 0.0;
    if (c < C) {
        for (int64_t r = a + threadIdx.y; r < b; r += 8) {
            double v = (double)ldf(x + r * C + c);
            s1 += v;
            s2 += v * v;
        }
    }
    r1[threadIdx.y][threadIdx.x] = s1;
    r2[threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double t1 = 0.0, t2 = 0.0;
        for (int k = 0; k < 8; ++k) { t1 += r1[k][threadIdx.x]; t2 += r2[k][threadIdx.x]; }
        int64_t chunk = (int64_t)blockIdx.z * gridDim.x + blockIdx.x;
        ws[(chunk * C + c) * 2 + 0] = t1;
        ws[(chunk * C + c) * 2 + 1] = t2;
    }
}

// one block (128 threads) per channel: threads stride over the chunks of every group at once (all loads independent), then
// a fixed-order tree (warp shuffles, 4 warps through shared memory); the groups' running-statistics updates are applied
// one after the other, as the separate calls would have.  groups <= 8 per pass.
__global__ void __launch_bounds__(128) bn_stats_final_kernel(const double* __restrict__ ws, int nchunks, int C, int groups,
                                                            int64_t rows_per_group, float* mean, float* var,
                                                            float* running_mean, float* running_var, float momentum) {
    constexpr int GB = 8;
    __shared__ double red[4][GB][2];
    const int c = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float rm = 0.f, rv = 0.f;
    if (running_mean && threadIdx.x == 0) { rm = running_mean[c]; rv = running_var[c]; }
    for (int g0 = 0; g0 < groups; g0 += GB) {
        double s1[GB], s2[GB];
#pragma unroll
        for (int u = 0; u < GB; ++u) s1[u] = s2[u] = 0.0;
        for (int k = threadIdx.x; k < nchunks; k += 128) {
#pragma unroll
            for (int u = 0; u < GB; ++u) {
                if (g0 + u < groups) {
                    const double* p = ws + (((int64_t)(g0 + u) * nchunks + k) * C + c) * 2;
                    s1[u] += p[0];
                    s2[u] += p[1];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < GB; ++u) {
            const double t1 = warp_sum(s1[u]), t2 = warp_sum(s2[u]);
            if (lane == 0) { red[warp][u][0] = t1; red[warp][u][1] = t2; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int u = 0; u < GB && g0 + u < groups; ++u) {
                const int g = g0 + u;
                const double t1 = (red[0][u][0] + red[1][u][0]) + (red[2][u][0] + red[3][u][0]);
                const double t2 = (red[0][u][1] + red[1][u][1]) + (red[2][u][1] + red[3][u][1]);
                double m = t1 / (double)rows_per_group;
                double v = t2 / (double)rows_per_group - m * m;
                if (v < 0.0) v = 0.0;
                mean[(int64_t)g * C + c] = (float)m;
                var[(int64_t)g * C + c] = (float)v;
                double unb = rows_per_group > 1 ? v * (double)rows_per_group / (double)(rows_per_group - 1) : v;
                rm = (1.f - momentum) * rm + momentum * (float)m;
                rv = (1.f - momentum) * rv + momentum * (float)unb;
            }
        }
        __syncthreads();
    }
    if (running_mean && threadIdx.x == 0) { running_mean[c] = rm; running_var[c] = rv; }
}

// statistics from the per-slab (sum, sum of squares) pairs a convolution epilogue produced (32 rows per slab): one block per
// channel, threads stride over the slabs of a group (fp64 accumulation, fixed-order tree), running statistics updated per
// group in call order — the same outputs as bn_stats_partial + bn_stats_final without reading the activation again
__global__ void __launch_bounds__(128) bn_stats_slabs_kernel(const float* __restrict__ cs, int64_t slabs_per_group, int ld,
                                                            int C, int groups, int64_t rows_per_group, float* mean, float* var,
                                                            float* running_mean, float* running_var, float momentum) {
    __shared__ double red[4][2];
    const int c = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float rm = 0.f, rv = 0.f;
    if (running_mean && threadIdx.x == 0) { rm = running_mean[c]; rv = running_var[c]; }
    for (int g = 0; g < groups; ++g) {
        double s1 = 0.0, s2 = 0.0;
        const float* base = cs + ((int64_t)g * slabs_per_group * ld + c) * 2;
#pragma unroll 4
        for (int64_t k = threadIdx.x; k < slabs_per_group; k += 128) {
            const float2 v = *reinterpret_cast<const float2*>(base + k * ld * 2);
            s1 += (double)v.x;
            s2 += (double)v.y;
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) { red[warp][0] = s1; red[warp][1] = s2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double t1 = (red[0][0] + red[1][0]) + (red[2][0] + red[3][0]);
            const double t2 = (red[0][1] + red[1][1]) + (red[2][1] + red[3][1]);
            double m = t1 / (double)rows_per_group;
            double v = t2 / (double)rows_per_group - m * m;
            if (v < 0.0) v = 0.0;
            mean[(int64_t)g * C + c] = (float)m;
            var[(int64_t)g * C + c] = (float)v;
            const double unb = rows_per_group > 1 ? v * (double)rows_per_group / (double)(rows_per_group - 1) : v;
            rm = (1.f - momentum) * rm + momentum * (float)m;
            rv = (1.f - momentum) * rv + momentum * (float)unb;
        }
        __syncthreads();
    }
    if (running_mean && threadIdx.x == 0) { running_mean[c] = rm; running_var[c] = rv; }
}

// ---------------------------------------------------------------------------------------------------------
// forward apply
// ---------------------------------------------------------------------------------------------------------
// T: storage type of x, y, residual and (SPADE) the gamma|beta activation; parameter tables are fp32
template <typename T, int MODE>
__global__ void norm_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int C,
                                const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                const void* __restrict__ gamma_v, const float* __restrict__ beta,
                                const int32_t* __restrict__ idx, int rows_per_seg, const T* __restrict__ residual,
                                int relu, int64_t rows_per_group) {
    const float* gamma = reinterpret_cast<const float*>(gamma_v);
    const T* gb = reinterpret_cast<const T*>(gamma_v);
    int C4 = C >> 2;
    int64_t total = rows * C4;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)(t % C4) << 2;
        int64_t r = t / C4;
        float4 v = ld4(x + t * 4);
        const int64_t go = (r / rows_per_group) * C;
        float4 m = *reinterpret_cast<const float4*>(mean + go + c);
        float4 vv = *reinterpret_cast<const float4*>(var + go + c);
        float4 o;
        o.x = (v.x - m.x) * (1.f / sqrtf(vv.x + eps));
        o.y = (v.y - m.y) * (1.f / sqrtf(vv.y + eps));
        o.z = (v.z - m.z) * (1.f / sqrtf(vv.z + eps));
        o.w = (v.w - m.w) * (1.f / sqrtf(vv.w + eps));
        if (MODE == B200_NORM_AFFINE) {
            float4 g = *reinterpret_cast<const float4*>(gamma + c);
            float4 b = *reinterpret_cast<const float4*>(beta + c);
            o.x = o.x * g.x + b.x; o.y = o.y * g.y + b.y; o.z = o.z * g.z + b.z; o.w = o.w * g.w + b.w;
        } else if (MODE == B200_NORM_CBN) {
            const float* row = gamma + (int64_t)idx[r / rows_per_seg] * 2 * C;
            float4 g = *reinterpret_cast<const float4*>(row + c);
            float4 b = *reinterpret_cast<const float4*>(row + C + c);
            o.x = g.x * o.x + b.x; o.y = g.y * o.y + b.y; o.z = g.z * o.z + b.z; o.w = g.w * o.w + b.w;
        } else if (MODE == B200_NORM_SPADE) {
            const T* row = gb + r * 2 * C;
            float4 g = ld4(row + c);
            float4 b = ld4(row + C + c);
            o.x = o.x * (1.f + g.x) + b.x; o.y = o.y * (1.f + g.y) + b.y;
            o.z = o.z * (1.f + g.z) + b.z; o.w = o.w * (1.f + g.w) + b.w;
        }
        if (residual) {
            float4 q = ld4(residual + t * 4);
            o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
        }
        if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
        st4(y + t * 4, o);
    }
}

// vector variant: a thread owns V consecutive channels (16 bytes) of a fixed channel slot and walks rows; no 64-bit
// divisions, per-row parameters (group statistics, CBN table row) resolved once per row
template <typename T, int MODE>
__global__ void __launch_bounds__(256) norm_fwd_vec_kernel(const T* __restrict__ x, T* __restrict__ y, int rows, int C,
                                                          const float* __restrict__ mean,
                                                          const float* __restrict__ var, float eps,
                                                          const void* __restrict__ gamma_v,
                                                          const float* __restrict__ beta,
                                                          const int32_t* __restrict__ idx, int rows_per_seg,
                                                          const T* __restrict__ residual, int relu,
                                                          int rows_per_group) {
    constexpr int V = VecIO<T>::V;
    const float* gamma = reinterpret_cast<const float*>(gamma_v);
    const T* gb = reinterpret_cast<const T*>(gamma_v);
    const int tpr = C / V;
    const int c = (threadIdx.x % tpr) * V;
    const int rpb = blockDim.x / tpr;
    float ga[V], be[V];
    if (MODE == B200_NORM_AFFINE) { ldp<V>(gamma + c, ga); ldp<V>(beta + c, be); }
    float m[V], rs[V];
    // a block owns a CONTIGUOUS range of rows: the group statistics and (conditional batch norm) the object's gamma / beta
    // table row change only every rows_per_group / rows_per_seg rows: the group and segment of a row are tracked
    // incrementally (no integer division per row — the kernel was issue bound: ncu issue active 65 %, DRAM 21 %,
    // profiles/r02m_mem_kernels.md) and their parameters re-loaded on change
    const int rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
    const int r_begin = blockIdx.x * rows_per_block;
    const int r_end = min(rows, r_begin + rows_per_block);
    int r = r_begin + threadIdx.x / tpr;
    if (r >= r_end) return;
    int g = r / rows_per_group;
    int64_t g_next = (int64_t)(g + 1) * rows_per_group;
    int seg = MODE == B200_NORM_CBN ? r / rows_per_seg : 0;
    int64_t seg_next = MODE == B200_NORM_CBN ? (int64_t)(seg + 1) * rows_per_seg : (int64_t)1 << 62;
    bool new_g = true, new_seg = true;
#pragma unroll 2
    for (; r < r_end; r += rpb) {
        const int64_t o = (int64_t)r * C + c;
        float v[V];
        VecIO<T>::load(x + o, v);
        while (r >= g_next) { ++g; g_next += rows_per_group; new_g = true; }
        if (new_g) {
            new_g = false;
            float vv[V];
            ldp<V>(mean + (int64_t)g * C + c, m);
            ldp<V>(var + (int64_t)g * C + c, vv);
#pragma unroll
            for (int e = 0; e < V; ++e) rs[e] = 1.f / sqrtf(vv[e] + eps);
        }
        float out[V];
#pragma unroll
        for (int e = 0; e < V; ++e) out[e] = (v[e] - m[e]) * rs[e];
        if (MODE == B200_NORM_AFFINE) {
#pragma unroll
            for (int e = 0; e < V; ++e) out[e] = out[e] * ga[e] + be[e];
        } else if (MODE == B200_NORM_CBN) {
            while (r >= seg_next) { ++seg; seg_next += rows_per_seg; new_seg = true; }
            if (new_seg) {
                new_seg = false;
                const float* row = gamma + (int64_t)idx[seg] * 2 * C;
                ldp<V>(row + c, ga);
                ldp<V>(row + C + c, be);
            }
#pragma unroll
            for (int e = 0; e < V; ++e) out[e] = ga[e] * out[e] + be[e];
        } else if (MODE == B200_NORM_SPADE) {
            const T* row = gb + (int64_t)r * 2 * C;
            VecIO<T>::load(row + c, ga);
            VecIO<T>::load(row + C + c, be);
#pragma unroll
            for (int e = 0; e < V; ++e) out[e] = out[e] * (1.f + ga[e]) + be[e];
        }
        if (residual) {
            float q[V];
            VecIO<T>::load(residual + o, q);
#pragma unroll
            for (int e = 0; e < V; ++e) out[e] += q[e];
        }
        if (relu) {
#pragma unroll
            for (int e = 0; e < V; ++e) out[e] = fmaxf(out[e], 0.f);
        }
        VecIO<T>::store(y + o, out);
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward stage 1: per-segment sums
// ---------------------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void norm_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                       const T* __restrict__ y, int64_t rows, int C, const float* __restrict__ mean,
                                       const float* __restrict__ var, float eps, const void* __restrict__ gamma,
                                       int rows_per_seg, int relu, double* __restrict__ seg_sums,
                                       int64_t rows_per_group, int segs_per_group) {
    __shared__ double r1[8][33], r2[8][33];
    int c = blockIdx.y * 32 + threadIdx.x;
    const int g = blockIdx.x / segs_per_group;
    const int64_t gend = (int64_t)(g + 1) * rows_per_group;
    int64_t a = (int64_t)g * rows_per_group + (int64_t)(blockIdx.x - g * segs_per_group) * rows_per_seg;
    int64_t b = a + rows_per_seg < gend ? a + rows_per_seg : gend;
    (void)rows;
    double s1 = 0.0, s2 = 0.0;
    if (c < C) {
        float m = mean[(int64_t)g * C + c], rstd = 1.f / sqrtf(var[(int64_t)g * C + c] + eps);
        for (int64_t r = a + threadIdx.y; r < b; r += 8) {
            float g = ldf(dy + r * C + c);
            if (relu && !(ldf(y + r * C + c) > 0.f)) g = 0.f;
            float xh = (ldf(x + r * C + c) - m) * rstd;
            if (MODE == B200_NORM_SPADE) g *= 1.f + ldf(reinterpret_cast<const T*>(gamma) + r * 2 * C + c);
            s1 += (double)g;
            s2 += (double)g * (double)xh;
        }
    }
    r1[threadIdx.y][threadIdx.x] = s1;
    r2[threadIdx.y][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        double t1 = 0.0, t2 = 0.0;
        for (int k = 0; k < 8; ++k) { t1 += r1[k][threadIdx.x]; t2 += r2[k][threadIdx.x]; }
        seg_sums[((int64_t)blockIdx.x * C + c) * 2 + 0] = t1;
        seg_sums[((int64_t)blockIdx.x * C + c) * 2 + 1] = t2;
    }
}

// ---------------------------------------------------------------------------------------------------------
// vector reductions over rows (statistics, backward segment sums): a block covers CT*V channels (CT = channel
// threads, a power of two <= 16) x 256/CT rows per pass; per-thread fp64 accumulation, then a fixed-order tree:
// lanes of a warp that share a channel slot (shuffle), then the 8 warps (shared memory).  Deterministic.
// ---------------------------------------------------------------------------------------------------------
template <int V>
__device__ __forceinline__ void block_reduce_rows(double (&s1)[V], double (&s2)[V], int ct, double* sm /* [8][16*V*2] */,
                                                  double* out1, double* out2, int64_t stride, bool valid) {
    // lanes l and l ^ o with o >= ct hold the same channel slot, different rows
    for (int o = 16; o >= ct; o >>= 1) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o);
            s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < ct) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            sm[(warp * 16 + lane) * V * 2 + e * 2] = s1[e];
            sm[(warp * 16 + lane) * V * 2 + e * 2 + 1] = s2[e];
        }
    }
    __syncthreads();
    if (threadIdx.x < ct * V && valid) {
        const int slot = threadIdx.x / V, e = threadIdx.x % V;
        double t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < 8; ++w) {
            t1 += sm[(w * 16 + slot) * V * 2 + e * 2];
            t2 += sm[(w * 16 + slot) * V * 2 + e * 2 + 1];
        }
        out1[(int64_t)threadIdx.x * stride] = t1;
        out2[(int64_t)threadIdx.x * stride] = t2;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_stats_partial_vec_kernel(const T* __restrict__ x, int64_t rows_per_group, int C,
                                                                  int64_t rows_per_chunk, int ct,
                                                                  double* __restrict__ ws) {
    constexpr int V = VecIO<T>::V;
    __shared__ double sm[8 * 16 * V * 2];
    const int c0 = blockIdx.y * ct * V;
    const int c = c0 + (threadIdx.x % ct) * V;
    const int rstep = 256 / ct;
    const int64_t g0 = (int64_t)blockIdx.z * rows_per_group;
    const int64_t a = g0 + (int64_t)blockIdx.x * rows_per_chunk;
    const int64_t e = g0 + rows_per_group;
    const int64_t b = a + rows_per_chunk < e ? a + rows_per_chunk : e;
    double s1[V], s2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.0;
    if (sizeof(T) == 2) {
        // bf16 storage: runs of kRun rows are summed in fp32 (x*x is exact: 16 significant bits) and each run is folded into
        // the fp64 accumulators — a quarter of the fp64 conversions / additions of the per-element form, and the run's loads
        // are issued back to back (the kernel was latency bound: ncu issue active 43 %, DRAM 32 %, profiles/r02m_mem_kernels.md)
        constexpr int kRun = 4;
        int64_t r = a + threadIdx.x / ct;
        for (; r + (int64_t)(kRun - 1) * rstep < b; r += (int64_t)kRun * rstep) {
            float v[kRun][V];
#pragma unroll
            for (int k = 0; k < kRun; ++k) VecIO<T>::load(x + (r + (int64_t)k * rstep) * C + c, v[k]);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float f1 = v[0][i], f2 = v[0][i] * v[0][i];
#pragma unroll
                for (int k = 1; k < kRun; ++k) { f1 += v[k][i]; f2 = __fmaf_rn(v[k][i], v[k][i], f2); }
                s1[i] += (double)f1;
                s2[i] += (double)f2;
            }
        }
        for (; r < b; r += rstep) {
            float v[V];
            VecIO<T>::load(x + r * C + c, v);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                s1[i] += (double)v[i];
                s2[i] += (double)(v[i] * v[i]);
            }
        }
    } else {
#pragma unroll 2
        for (int64_t r = a + threadIdx.x / ct; r < b; r += rstep) {
            float v[V];
            VecIO<T>::load(x + r * C + c, v);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const double d = (double)v[i];
                s1[i] += d;
                s2[i] += d * d;
            }
        }
    }
    const int64_t chunk = (int64_t)blockIdx.z * gridDim.x + blockIdx.x;
    double* o = ws + (chunk * C + c0) * 2;
    block_reduce_rows<V>(s1, s2, ct, sm, o, o + 1, 2, true);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256) norm_bwd_reduce_vec_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                 const T* __restrict__ y, int C,
                                                                 const float* __restrict__ mean,
                                                                 const float* __restrict__ var, float eps,
                                                                 const void* __restrict__ gamma,
                                                                 const int32_t* __restrict__ idx, int rows_per_seg,
                                                                 int relu, double* __restrict__ seg_sums,
                                                                 int64_t rows_per_group, int segs_per_group, int ct) {
    constexpr int V = VecIO<T>::V;
    __shared__ double sm[8 * 16 * V * 2];
    const int c0 = blockIdx.y * ct * V;
    const int c = c0 + (threadIdx.x % ct) * V;
    const int rstep = 256 / ct;
    const int g = blockIdx.x / segs_per_group;
    const int64_t gend = (int64_t)(g + 1) * rows_per_group;
    const int64_t a = (int64_t)g * rows_per_group + (int64_t)(blockIdx.x - g * segs_per_group) * rows_per_seg;
    const int64_t b = a + rows_per_seg < gend ? a + rows_per_seg : gend;
    float m[V], rs[V];
    {
        float vv[V];
        ldp<V>(mean + (int64_t)g * C + c, m);
        ldp<V>(var + (int64_t)g * C + c, vv);
#pragma unroll
        for (int i = 0; i < V; ++i) rs[i] = 1.f / sqrtf(vv[i] + eps);
    }
    double s1[V], s2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.0;
    // relu == 2 (CBN): the ReLU mask is recomputed from x (one segment = one object = one table row) instead of
    // reading the saved output: gamma * xhat + beta > 0, the exact fp32 expression of the forward kernel
    float cg[V], cb[V];
    if (MODE == B200_NORM_CBN && relu == 2) {
        const float* row = reinterpret_cast<const float*>(gamma) + (int64_t)idx[a / rows_per_seg] * 2 * C;
        ldp<V>(row + c, cg);
        ldp<V>(row + C + c, cb);
    }
#pragma unroll 2
    for (int64_t r = a + threadIdx.x / ct; r < b; r += rstep) {
        float gv[V], xv[V];
        VecIO<T>::load(dy + r * C + c, gv);
        VecIO<T>::load(x + r * C + c, xv);
        if (MODE == B200_NORM_CBN && relu == 2) {
#pragma unroll
            for (int i = 0; i < V; ++i)
                if (!(cg[i] * ((xv[i] - m[i]) * rs[i]) + cb[i] > 0.f)) gv[i] = 0.f;
        } else if (relu) {
            float yv[V];
            VecIO<T>::load(y + r * C + c, yv);
#pragma unroll
            for (int i = 0; i < V; ++i)
                if (!(yv[i] > 0.f)) gv[i] = 0.f;
        }
        if (MODE == B200_NORM_SPADE) {
            float q[V];
            VecIO<T>::load(reinterpret_cast<const T*>(gamma) + r * 2 * C + c, q);
#pragma unroll
            for (int i = 0; i < V; ++i) gv[i] *= 1.f + q[i];
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float xh = (xv[i] - m[i]) * rs[i];
            s1[i] += (double)gv[i];
            s2[i] += (double)gv[i] * (double)xh;
        }
    }
    double* o = seg_sums + ((int64_t)blockIdx.x * C + c0) * 2;
    block_reduce_rows<V>(s1, s2, ct, sm, o, o + 1, 2, true);
}

// stage 2: s[g][c] = (sum dxhat, sum dxhat*xhat) per group; parameter gradients summed over the groups.
// One warp per channel, fixed-order shuffle tree.
__global__ void __launch_bounds__(128) norm_bwd_finalize_kernel(const double* __restrict__ seg, int nseg, int C, int mode,
                                                               const float* __restrict__ gamma,
                                                               const int32_t* __restrict__ idx, float* s, float* dgamma,
                                                               float* dbeta, int groups) {
    // one block (128 threads) per channel: the segments of a group are strided over the threads (independent loads), then a
    // fixed-order tree (warp shuffles, 4 warps through shared memory)
    __shared__ double red[4][2];
    const int c = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int spg = nseg / groups;
    double t1 = 0.0, t2 = 0.0;
    for (int g = 0; g < groups; ++g) {
        double s1 = 0.0, s2 = 0.0;
        if (mode == B200_NORM_CBN) {
#pragma unroll 2
            for (int k = g * spg + threadIdx.x; k < (g + 1) * spg; k += 128) {
                double w = (double)gamma[(int64_t)idx[k] * 2 * C + c];
                s1 += w * seg[((int64_t)k * C + c) * 2];
                s2 += w * seg[((int64_t)k * C + c) * 2 + 1];
            }
        } else {
#pragma unroll 2
            for (int k = g * spg + threadIdx.x; k < (g + 1) * spg; k += 128) {
                s1 += seg[((int64_t)k * C + c) * 2];
                s2 += seg[((int64_t)k * C + c) * 2 + 1];
            }
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) { red[warp][0] = s1; red[warp][1] = s2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            s1 = (red[0][0] + red[1][0]) + (red[2][0] + red[3][0]);
            s2 = (red[0][1] + red[1][1]) + (red[2][1] + red[3][1]);
            t1 += s1;
            t2 += s2;
            if (mode == B200_NORM_AFFINE) {
                double w = (double)gamma[c];
                s1 *= w;
                s2 *= w;
            }
            s[((int64_t)g * C + c) * 2] = (float)s1;
            s[((int64_t)g * C + c) * 2 + 1] = (float)s2;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && mode == B200_NORM_AFFINE) {
        if (dbeta) dbeta[c] = (float)t1;
        if (dgamma) dgamma[c] = (float)t2;
    }
}

// CBN embedding gradient: dtable[k][c] = sum_{o: idx[o]==k} sum(dyr*xhat); dtable[k][C+c] = sum_{o} sum(dyr).
// One block per class k: warp 0 compacts the segments of that class into shared memory in ascending order (ballot
// prefix), then the threads walk that short list per channel — the summation order of the plain scan, without every
// (class, channel) thread scanning all segments.
constexpr int kCbnListMax = 2048;
__global__ void __launch_bounds__(256) cbn_dtable_kernel(const double* __restrict__ seg, int nseg, int C,
                                                        const int32_t* __restrict__ idx, int num_classes,
                                                        float* __restrict__ dtable) {
    __shared__ int list[kCbnListMax];
    __shared__ int cnt;
    const int k = blockIdx.x;
    (void)num_classes;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int n = 0;
        for (int base = 0; base < nseg; base += 32) {
            const int o = base + lane;
            const bool m = o < nseg && idx[o] == k;
            const unsigned bal = __ballot_sync(0xffffffffu, m);
            if (m) {
                const int pos = n + __popc(bal & ((1u << lane) - 1u));
                if (pos < kCbnListMax) list[pos] = o;
            }
            n += __popc(bal);
        }
        if (lane == 0) cnt = n;
    }
    __syncthreads();
    const int n = cnt;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double a = 0.0, b = 0.0;
        if (n <= kCbnListMax) {
            for (int i = 0; i < n; ++i) {
                const int o = list[i];
                a += seg[((int64_t)o * C + c) * 2 + 1];
                b += seg[((int64_t)o * C + c) * 2];
            }
        } else {                                  // more matches than the list holds: plain ordered scan
            for (int o = 0; o < nseg; ++o)
                if (idx[o] == k) { a += seg[((int64_t)o * C + c) * 2 + 1]; b += seg[((int64_t)o * C + c) * 2]; }
        }
        dtable[(int64_t)k * 2 * C + c] = (float)a;
        dtable[(int64_t)k * 2 * C + C + c] = (float)b;
    }
}

// stage 3 (4 channels per thread)
template <typename T, int MODE>
__global__ void norm_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                      const T* __restrict__ y, T* __restrict__ dx, int64_t rows, int C,
                                      const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                      const void* __restrict__ gamma_v, const int32_t* __restrict__ idx,
                                      int rows_per_seg, int relu, const float* __restrict__ s, T* __restrict__ dgb,
                                      int64_t rows_per_group) {
    const float* gamma = reinterpret_cast<const float*>(gamma_v);
    const int C4 = C >> 2;
    const int64_t total = rows * C4;
    const float inv_rows = 1.f / (float)rows_per_group;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C4) << 2;
        const int64_t r = t / C4;
        const float4 g4 = ld4(dy + t * 4);
        const float4 x4 = ld4(x + t * 4);
        float g[4] = {g4.x, g4.y, g4.z, g4.w};
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
        if (relu) {
            const float4 y4 = ld4(y + t * 4);
            if (!(y4.x > 0.f)) g[0] = 0.f;
            if (!(y4.y > 0.f)) g[1] = 0.f;
            if (!(y4.z > 0.f)) g[2] = 0.f;
            if (!(y4.w > 0.f)) g[3] = 0.f;
        }
        const int64_t go = (r / rows_per_group) * C;
        const float4 m4 = *reinterpret_cast<const float4*>(mean + go + c);
        const float4 v4 = *reinterpret_cast<const float4*>(var + go + c);
        const float mv[4] = {m4.x, m4.y, m4.z, m4.w};
        const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
        float ga[4] = {1.f, 1.f, 1.f, 1.f};
        if (MODE == B200_NORM_AFFINE) {
            const float4 q = *reinterpret_cast<const float4*>(gamma + c);
            ga[0] = q.x; ga[1] = q.y; ga[2] = q.z; ga[3] = q.w;
        } else if (MODE == B200_NORM_CBN) {
            const float4 q = *reinterpret_cast<const float4*>(gamma + (int64_t)idx[r / rows_per_seg] * 2 * C + c);
            ga[0] = q.x; ga[1] = q.y; ga[2] = q.z; ga[3] = q.w;
        } else if (MODE == B200_NORM_SPADE) {
            const float4 q = ld4(reinterpret_cast<const T*>(gamma_v) + r * 2 * C + c);
            ga[0] = 1.f + q.x; ga[1] = 1.f + q.y; ga[2] = 1.f + q.z; ga[3] = 1.f + q.w;
        }
        float o[4], gx[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float rstd = 1.f / sqrtf(vv[e] + eps);
            const float xh = (xv[e] - mv[e]) * rstd;
            const float dxh = g[e] * ga[e];
            gx[e] = g[e] * xh;
            o[e] = rstd * (dxh - s[(go + c + e) * 2] * inv_rows - xh * s[(go + c + e) * 2 + 1] * inv_rows);
        }
        if (MODE == B200_NORM_SPADE) {
            st4(dgb + r * 2 * C + c, make_float4(gx[0], gx[1], gx[2], gx[3]));
            st4(dgb + r * 2 * C + C + c, make_float4(g[0], g[1], g[2], g[3]));
        }
        st4(dx + t * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
}


// stage 3, vector variant (V channels of a fixed slot per thread, rows walked; see norm_fwd_vec_kernel)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) norm_bwd_apply_vec_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                const T* __restrict__ y, T* __restrict__ dx, int rows,
                                                                int C, const float* __restrict__ mean,
                                                                const float* __restrict__ var, float eps,
                                                                const void* __restrict__ gamma_v,
                                                                const int32_t* __restrict__ idx, int rows_per_seg,
                                                                int relu, const float* __restrict__ s,
                                                                T* __restrict__ dgb, int rows_per_group) {
    constexpr int V = VecIO<T>::V;
    const float* gamma = reinterpret_cast<const float*>(gamma_v);
    const int tpr = C / V;
    const int c = (threadIdx.x % tpr) * V;
    const int rpb = blockDim.x / tpr;
    const float inv_rows = 1.f / (float)rows_per_group;
    float ga[V];
#pragma unroll
    for (int e = 0; e < V; ++e) ga[e] = 1.f;
    if (MODE == B200_NORM_AFFINE) ldp<V>(gamma + c, ga);
    float m[V], rs[V], sa[V], sb[V], be[V];
    const int rows_per_block = (rows + gridDim.x - 1) / gridDim.x;      // contiguous rows per block (see norm_fwd_vec_kernel)
    const int r_begin = blockIdx.x * rows_per_block;
    const int r_end = min(rows, r_begin + rows_per_block);
    int r = r_begin + threadIdx.x / tpr;
    if (r >= r_end) return;
    int grp = r / rows_per_group;                                       // tracked incrementally, as in norm_fwd_vec_kernel
    int64_t g_next = (int64_t)(grp + 1) * rows_per_group;
    int seg = MODE == B200_NORM_CBN ? r / rows_per_seg : 0;
    int64_t seg_next = MODE == B200_NORM_CBN ? (int64_t)(seg + 1) * rows_per_seg : (int64_t)1 << 62;
    bool new_g = true, new_seg = true;
#pragma unroll 2
    for (; r < r_end; r += rpb) {
        const int64_t o = (int64_t)r * C + c;
        float g[V], xv[V];
        VecIO<T>::load(dy + o, g);
        VecIO<T>::load(x + o, xv);
        if (relu == 1) {
            float yv[V];
            VecIO<T>::load(y + o, yv);
#pragma unroll
            for (int e = 0; e < V; ++e)
                if (!(yv[e] > 0.f)) g[e] = 0.f;
        }
        while (r >= g_next) { ++grp; g_next += rows_per_group; new_g = true; }
        if (new_g) {
            new_g = false;
            float vv[V], s2[2 * V];
            ldp<V>(mean + (int64_t)grp * C + c, m);
            ldp<V>(var + (int64_t)grp * C + c, vv);
            ldp<2 * V>(s + ((int64_t)grp * C + c) * 2, s2);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                rs[e] = 1.f / sqrtf(vv[e] + eps);
                sa[e] = s2[2 * e] * inv_rows;
                sb[e] = s2[2 * e + 1] * inv_rows;
            }
        }
        if (MODE == B200_NORM_CBN) {
            while (r >= seg_next) { ++seg; seg_next += rows_per_seg; new_seg = true; }
            if (new_seg) {
                new_seg = false;
                const float* row = gamma + (int64_t)idx[seg] * 2 * C;
                ldp<V>(row + c, ga);
                if (relu == 2) ldp<V>(row + C + c, be);
            }
            if (relu == 2) {      // recomputed ReLU mask: gamma * xhat + beta > 0 (the forward kernel's fp32 expression)
#pragma unroll
                for (int e = 0; e < V; ++e)
                    if (!(ga[e] * ((xv[e] - m[e]) * rs[e]) + be[e] > 0.f)) g[e] = 0.f;
            }
        } else if (MODE == B200_NORM_SPADE) {
            float q[V];
            VecIO<T>::load(reinterpret_cast<const T*>(gamma_v) + (int64_t)r * 2 * C + c, q);
#pragma unroll
            for (int e = 0; e < V; ++e) ga[e] = 1.f + q[e];
        }
        float out[V], gx[V];
#pragma unroll
        for (int e = 0; e < V; ++e) {
            const float xh = (xv[e] - m[e]) * rs[e];
            const float dxh = g[e] * ga[e];
            gx[e] = g[e] * xh;
            out[e] = rs[e] * (dxh - sa[e] - xh * sb[e]);
        }
        if (MODE == B200_NORM_SPADE) {
            VecIO<T>::store(dgb + (int64_t)r * 2 * C + c, gx);
            VecIO<T>::store(dgb + (int64_t)r * 2 * C + C + c, g);
        }
        VecIO<T>::store(dx + o, out);
    }
}

static inline bool vec_ok_dt(int dt, int64_t rows, int C) {
    return dt == B200_BF16 ? vec_ok<bf16>(rows, C) : vec_ok<float>(rows, C);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_bn_stats(const void* x, int dt, int64_t rows, int C, int groups, float* mean, float* var,
                             float* running_mean, float* running_var, float momentum, double* ws,
                             b200_stream_t stream) {
    B200_REQUIRE(rows > 0 && C > 0 && groups >= 1 && rows % groups == 0, "bn_stats: bad rows=%lld groups=%d", (long long)rows, groups);
    B200_REQUIRE(groups < 65536, "bn_stats: too many groups");
    int64_t rpg = rows / groups;
    int nchunks = b200_bn_chunks(rpg, C);
    int64_t rpc = (rpg + nchunks - 1) / nchunks;
    dim3 grid(nchunks, (C + 31) / 32, groups), block(32, 8);
    B200_DISPATCH_DT(dt, T, {
        if (vec_ok<T>(rows, C)) {
            const int tpr = C / VecIO<T>::V, ct = tpr < 16 ? tpr : 16;
            dim3 vgrid(nchunks, tpr / ct, groups);
            bn_stats_partial_vec_kernel<T><<<vgrid, 256, 0, as_stream(stream)>>>((const T*)x, rpg, C, rpc, ct, ws);
        } else {
            bn_stats_partial_kernel<T><<<grid, block, 0, as_stream(stream)>>>((const T*)x, rpg, C, rpc, ws);
        }
    });
    B200_CHECK_LAUNCH();
    bn_stats_final_kernel<<<C, 128, 0, as_stream(stream)>>>(ws, nchunks, C, groups, rpg, mean, var, running_mean, running_var,
                                                            momentum);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_bn_stats_slabs(const float* col_stats, int64_t n_slabs, int ld, int64_t rows, int C, int groups, float* mean,
                                   float* var, float* running_mean, float* running_var, float momentum,
                                   b200_stream_t stream) {
    B200_REQUIRE(rows > 0 && C > 0 && groups >= 1 && rows % groups == 0 && ld >= C, "bn_stats_slabs: bad sizes");
    const int64_t rpg = rows / groups;
    B200_REQUIRE(rpg % 32 == 0, "bn_stats_slabs: rows per group must be a multiple of 32 (slabs must not straddle groups)");
    B200_REQUIRE(n_slabs >= rows / 32 && (reinterpret_cast<uintptr_t>(col_stats) & 7) == 0, "bn_stats_slabs: slab buffer too small");
    bn_stats_slabs_kernel<<<C, 128, 0, as_stream(stream)>>>(col_stats, rpg / 32, ld, C, groups, rpg, mean, var, running_mean,
                                                            running_var, momentum);
    B200_CHECK_LAUNCH();
    return 0;
}

template <typename T>
static int norm_fwd_launch(const void* x, void* y, int64_t rows, int C, int64_t rpg, const float* mean, const float* var,
                           float eps, int mode, const void* gamma, const float* beta, const int32_t* idx, int rows_per_seg,
                           const void* residual, int relu, cudaStream_t st) {
    const T* xp = (const T*)x;
    T* yp = (T*)y;
    const T* rp = (const T*)residual;
    if (vec_ok<T>(rows, C) && rpg < (1ll << 31)) {
        const int rpb = 256 / (C / VecIO<T>::V);
        const int vg = grid_for((rows + rpb - 1) / rpb, 1, 8);
        const int ri = (int)rows, gi = (int)rpg;
        switch (mode) {
            case B200_NORM_PLAIN:
                norm_fwd_vec_kernel<T, B200_NORM_PLAIN><<<vg, 256, 0, st>>>(xp, yp, ri, C, mean, var, eps, gamma, beta, idx, rows_per_seg, rp, relu, gi);
                break;
            case B200_NORM_AFFINE:
                norm_fwd_vec_kernel<T, B200_NORM_AFFINE><<<vg, 256, 0, st>>>(xp, yp, ri, C, mean, var, eps, gamma, beta, idx, rows_per_seg, rp, relu, gi);
                break;
            case B200_NORM_CBN:
                norm_fwd_vec_kernel<T, B200_NORM_CBN><<<vg, 256, 0, st>>>(xp, yp, ri, C, mean, var, eps, gamma, beta, idx, rows_per_seg, rp, relu, gi);
                break;
            case B200_NORM_SPADE:
                norm_fwd_vec_kernel<T, B200_NORM_SPADE><<<vg, 256, 0, st>>>(xp, yp, ri, C, mean, var, eps, gamma, beta, idx, rows_per_seg, rp, relu, gi);
                break;
            default:
                return set_error("norm_fwd: bad mode %d", mode);
        }
        return 0;
    }
    const int g = grid_for(rows * (C / 4), 256);
    switch (mode) {
        case B200_NORM_PLAIN:
            norm_fwd_kernel<T, B200_NORM_PLAIN><<<g, 256, 0, st>>>(xp, yp, rows, C, mean, var, eps, gamma, beta, idx, rows_per_seg, rp, relu, rpg);
            break;
        case B200_NORM_AFFINE:
            norm_fwd_kernel<T, B200_NORM_AFFINE><<<g, 256, 0, st>>>(xp, yp, rows, C, mean, var, eps, gamma, beta, idx, rows_per_seg, rp, relu, rpg);
            break;
        case B200_NORM_CBN:
            norm_fwd_kernel<T, B200_NORM_CBN><<<g, 256, 0, st>>>(xp, yp, rows, C, mean, var, eps, gamma, beta, idx, rows_per_seg, rp, relu, rpg);
            break;
        case B200_NORM_SPADE:
            norm_fwd_kernel<T, B200_NORM_SPADE><<<g, 256, 0, st>>>(xp, yp, rows, C, mean, var, eps, gamma, beta, idx, rows_per_seg, rp, relu, rpg);
            break;
        default:
            return set_error("norm_fwd: bad mode %d", mode);
    }
    return 0;
}

extern "C" int b200_norm_fwd(const void* x, void* y, int dt, int64_t rows, int C, int groups, const float* mean,
                             const float* var, float eps, int mode, const void* gamma, const float* beta,
                             const int32_t* idx, int rows_per_seg, const void* residual, int relu,
                             b200_stream_t stream) {
    B200_REQUIRE(C % 4 == 0, "norm_fwd: C=%d must be a multiple of 4", C);
    if (rows == 0) return 0;
    B200_REQUIRE(groups >= 1 && rows % groups == 0, "norm_fwd: rows %% groups != 0");
    const int64_t rpg = rows / groups;
    if (rows_per_seg < 1) rows_per_seg = 1;
    int rc;
    B200_DISPATCH_DT(dt, T, { rc = norm_fwd_launch<T>(x, y, rows, C, rpg, mean, var, eps, mode, gamma, beta, idx, rows_per_seg, residual, relu, as_stream(stream)); });
    if (rc) return rc;
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_norm_bwd_reduce(const void* dy, const void* x, const void* y, int dt, int64_t rows, int C, int groups,
                                    const float* mean, const float* var, float eps, int mode, const void* gamma,
                                    const int32_t* idx, int rows_per_seg, int relu, double* seg_sums,
                                    b200_stream_t stream) {
    B200_REQUIRE(rows > 0 && rows_per_seg > 0 && groups >= 1 && rows % groups == 0, "norm_bwd_reduce: bad sizes");
    B200_REQUIRE(relu != 2 || (mode == B200_NORM_CBN && idx != nullptr && vec_ok_dt(dt, rows, C) && (rows / groups) % rows_per_seg == 0),
                 "norm_bwd_reduce: relu = 2 (recomputed mask) needs conditional batch norm on a vector-kernel layout");
    const int64_t rpg = rows / groups;
    const int spg = (int)((rpg + rows_per_seg - 1) / rows_per_seg);
    dim3 grid(spg * groups, (C + 31) / 32), block(32, 8);
    cudaStream_t st = as_stream(stream);
    B200_DISPATCH_DT(dt, T, {
        if (vec_ok<T>(rows, C)) {
            const int tpr = C / VecIO<T>::V, ct = tpr < 16 ? tpr : 16;
            dim3 vgrid(spg * groups, tpr / ct);
            if (mode == B200_NORM_SPADE)
                norm_bwd_reduce_vec_kernel<T, B200_NORM_SPADE><<<vgrid, 256, 0, st>>>((const T*)dy, (const T*)x, (const T*)y, C, mean, var, eps, gamma, idx, rows_per_seg, relu, seg_sums, rpg, spg, ct);
            else if (mode == B200_NORM_CBN && relu == 2)
                norm_bwd_reduce_vec_kernel<T, B200_NORM_CBN><<<vgrid, 256, 0, st>>>((const T*)dy, (const T*)x, (const T*)y, C, mean, var, eps, gamma, idx, rows_per_seg, relu, seg_sums, rpg, spg, ct);
            else
                norm_bwd_reduce_vec_kernel<T, B200_NORM_PLAIN><<<vgrid, 256, 0, st>>>((const T*)dy, (const T*)x, (const T*)y, C, mean, var, eps, gamma, idx, rows_per_seg, relu, seg_sums, rpg, spg, ct);
        } else
        if (mode == B200_NORM_SPADE)
            norm_bwd_reduce_kernel<T, B200_NORM_SPADE><<<grid, block, 0, st>>>((const T*)dy, (const T*)x, (const T*)y, rows, C, mean, var, eps, gamma, rows_per_seg, relu, seg_sums, rpg, spg);
        else
            norm_bwd_reduce_kernel<T, B200_NORM_PLAIN><<<grid, block, 0, st>>>((const T*)dy, (const T*)x, (const T*)y, rows, C, mean, var, eps, gamma, rows_per_seg, relu, seg_sums, rpg, spg);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_norm_bwd_finalize(const double* seg_sums, int nseg, int C, int groups, int mode, const float* gamma,
                                      const int32_t* idx, int num_classes, float* s, float* dgamma, float* dbeta,
                                      float* dtable, b200_stream_t stream) {
    cudaStream_t st = as_stream(stream);
    B200_REQUIRE(groups >= 1 && nseg % groups == 0, "norm_bwd_finalize: nseg %% groups != 0");
    norm_bwd_finalize_kernel<<<C, 128, 0, st>>>(seg_sums, nseg, C, mode, gamma, idx, s, dgamma, dbeta, groups);
    B200_CHECK_LAUNCH();
    if (mode == B200_NORM_CBN && dtable) {
        cbn_dtable_kernel<<<num_classes, 256, 0, st>>>(seg_sums, nseg, C, idx, num_classes, dtable);
        B200_CHECK_LAUNCH();
    }
    return 0;
}

template <typename T>
static int norm_bwd_apply_launch(const void* dy, const void* x, const void* y, void* dx, int64_t rows, int C, int64_t rpg,
                                 const float* mean, const float* var, float eps, int mode, const void* gamma,
                                 const int32_t* idx, int rows_per_seg, int relu, const float* s, void* dgb,
                                 cudaStream_t st) {
    const T *dyp = (const T*)dy, *xp = (const T*)x, *yp = (const T*)y;
    T *dxp = (T*)dx, *dgbp = (T*)dgb;
    if (vec_ok<T>(rows, C) && rpg < (1ll << 31)) {
        const int rpb = 256 / (C / VecIO<T>::V);
        const int vg = grid_for((rows + rpb - 1) / rpb, 1, 8);
        const int ri = (int)rows, gi = (int)rpg;
        switch (mode) {
            case B200_NORM_PLAIN:
                norm_bwd_apply_vec_kernel<T, B200_NORM_PLAIN><<<vg, 256, 0, st>>>(dyp, xp, yp, dxp, ri, C, mean, var, eps, gamma, idx, rows_per_seg, relu, s, dgbp, gi);
                break;
            case B200_NORM_AFFINE:
                norm_bwd_apply_vec_kernel<T, B200_NORM_AFFINE><<<vg, 256, 0, st>>>(dyp, xp, yp, dxp, ri, C, mean, var, eps, gamma, idx, rows_per_seg, relu, s, dgbp, gi);
                break;
            case B200_NORM_CBN:
                norm_bwd_apply_vec_kernel<T, B200_NORM_CBN><<<vg, 256, 0, st>>>(dyp, xp, yp, dxp, ri, C, mean, var, eps, gamma, idx, rows_per_seg, relu, s, dgbp, gi);
                break;
            case B200_NORM_SPADE:
                if (dgb == nullptr) return set_error("norm_bwd_apply: SPADE needs dgb");
                norm_bwd_apply_vec_kernel<T, B200_NORM_SPADE><<<vg, 256, 0, st>>>(dyp, xp, yp, dxp, ri, C, mean, var, eps, gamma, idx, rows_per_seg, relu, s, dgbp, gi);
                break;
            default:
                return set_error("norm_bwd_apply: bad mode %d", mode);
        }
        return 0;
    }
    const int g = grid_for(rows * (C / 4), 256);
    switch (mode) {
        case B200_NORM_PLAIN:
            norm_bwd_apply_kernel<T, B200_NORM_PLAIN><<<g, 256, 0, st>>>(dyp, xp, yp, dxp, rows, C, mean, var, eps, gamma, idx, rows_per_seg, relu, s, dgbp, rpg);
            break;
        case B200_NORM_AFFINE:
            norm_bwd_apply_kernel<T, B200_NORM_AFFINE><<<g, 256, 0, st>>>(dyp, xp, yp, dxp, rows, C, mean, var, eps, gamma, idx, rows_per_seg, relu, s, dgbp, rpg);
            break;
        case B200_NORM_CBN:
            norm_bwd_apply_kernel<T, B200_NORM_CBN><<<g, 256, 0, st>>>(dyp, xp, yp, dxp, rows, C, mean, var, eps, gamma, idx, rows_per_seg, relu, s, dgbp, rpg);
            break;
        case B200_NORM_SPADE:
            if (dgb == nullptr) return set_error("norm_bwd_apply: SPADE needs dgb");
            norm_bwd_apply_kernel<T, B200_NORM_SPADE><<<g, 256, 0, st>>>(dyp, xp, yp, dxp, rows, C, mean, var, eps, gamma, idx, rows_per_seg, relu, s, dgbp, rpg);
            break;
        default:
            return set_error("norm_bwd_apply: bad mode %d", mode);
    }
    return 0;
}

extern "C" int b200_norm_bwd_apply(const void* dy, const void* x, const void* y, void* dx, int dt, int64_t rows, int C,
                                   int groups, const float* mean, const float* var, float eps, int mode,
                                   const void* gamma, const int32_t* idx, int rows_per_seg, int relu, const float* s,
                                   void* dgb, b200_stream_t stream) {
    if (rows == 0) return 0;
    B200_REQUIRE(C % 4 == 0, "norm_bwd_apply: C=%d must be a multiple of 4", C);
    B200_REQUIRE(groups >= 1 && rows % groups == 0, "norm_bwd_apply: rows %% groups != 0");
    B200_REQUIRE(relu != 2 || (mode == B200_NORM_CBN && vec_ok_dt(dt, rows, C) && rows / groups < (1ll << 31)),
                 "norm_bwd_apply: relu = 2 (recomputed mask) needs conditional batch norm on a vector-kernel layout");
    const int64_t rpg = rows / groups;
    if (rows_per_seg < 1) rows_per_seg = 1;
    int rc;
    B200_DISPATCH_DT(dt, T, { rc = norm_bwd_apply_launch<T>(dy, x, y, dx, rows, C, rpg, mean, var, eps, mode, gamma, idx, rows_per_seg, relu, s, dgb, as_stream(stream)); });
    if (rc) return rc;
    B200_CHECK_LAUNCH();
    return 0;
}
