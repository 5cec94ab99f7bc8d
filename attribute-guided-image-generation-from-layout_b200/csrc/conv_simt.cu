// conv_simt.cu — fp32 CUDA-core gather-GEMM convolutions (forward-type and weight-gradient).
// This is the bit-tight parity path (plain fp32 FMA accumulation) and the handler of the skinny cases the
// tcgen05 path does not take (Cin = 3 first layers, Linear heads).  Descriptor semantics: include/b200gan.h.
#include "common.cuh"

namespace b200 {

constexpr int BM = 64, BN = 64, BK = 16;

struct RowGeom {
    int64_t in_base;   // n * in_sn
    int iy0, ix0;      // qy*in_sy + tap_oy, qx*in_sx + tap_ox
    int valid;
};

__device__ __forceinline__ void decode_row(const b200_conv_desc& d, int64_t m, int64_t M, RowGeom& g) {
    g.valid = m < M;
    if (!g.valid) { g.in_base = 0; g.iy0 = g.ix0 = 0; return; }
    int qx = (int)(m % d.Qw);
    int qy = (int)((m / d.Qw) % d.Qh);
    int64_t n = m / ((int64_t)d.Qw * d.Qh);
    g.in_base = n * d.in_sn;
    g.iy0 = qy * d.in_sy + d.tap_oy;
    g.ix0 = qx * d.in_sx + d.tap_ox;
}

__device__ __forceinline__ int64_t out_offset(const b200_conv_desc& d, int64_t m, int64_t M) {
    if (m >= M) return -1;
    int qx = (int)(m % d.Qw);
    int qy = (int)((m / d.Qw) % d.Qh);
    int64_t n = m / ((int64_t)d.Qw * d.Qh);
    int oy = qy * d.out_sy + d.out_oy, ox = qx * d.out_sx + d.out_ox;
    if (oy < 0 || oy >= d.Ho || ox < 0 || ox >= d.Wo) return -1;
    return n * d.out_sn + (int64_t)oy * d.out_sh + (int64_t)ox * d.out_sw;
}

// ----------------------------------------------------------------------------------------------------------
// forward-type gather GEMM: out[m, co] = sum_k A[m, k] * wmat[co, k]
// VEC: Cin % 16 == 0, in_sc == 1, 16-byte aligned rows -> one tap per K step and float4 gathers
// ----------------------------------------------------------------------------------------------------------
// TI / TO: storage types of the gathered input and of the output (float or bf16); fp32 arithmetic
template <bool VEC, typename TI, typename TO>
__global__ void __launch_bounds__(256) conv_gemm_f32_kernel(b200_conv_desc d, const TI* __restrict__ in,
                                                            const float* __restrict__ wmat,
                                                            const float* __restrict__ bias,
                                                            const float* __restrict__ scale, TO* __restrict__ out) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    __shared__ int64_t row_out[BM];

    const int tid = threadIdx.x;
    const int64_t M = (int64_t)d.B * d.Qh * d.Qw;
    const int K = d.Th * d.Tw * d.Cin;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // loader mapping: 64 rows x 4 k-quads
    const int lr = tid >> 2;
    const int lk = (tid & 3) * 4;
    RowGeom rg;
    decode_row(d, m0 + lr, M, rg);
    if (tid < BM) row_out[tid] = out_offset(d, m0 + tid, M);

    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int wrow = n0 + lr;
    const bool wvalid = wrow < d.Cout;
    const bool wvec = (d.ldw % 4 == 0) && ((reinterpret_cast<uintptr_t>(wmat) & 15) == 0);

    for (int k0 = 0; k0 < K; k0 += BK) {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        if (VEC) {
            int tap = k0 / d.Cin, c0 = k0 - tap * d.Cin;
            int tyy = tap / d.Tw, txx = tap - tyy * d.Tw;
            int iy = rg.iy0 + tyy * d.tap_sy, ix = rg.ix0 + txx * d.tap_sx;
            if (rg.valid && iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi) {
                const TI* p = in + rg.in_base + (int64_t)(iy >> d.up_shift) * d.in_sh +
                              (int64_t)(ix >> d.up_shift) * d.in_sw + c0 + lk;
                float4 v = ld4(p);
                a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int k = k0 + lk + e;
                if (rg.valid && k < K) {
                    int tap = k / d.Cin, c = k - tap * d.Cin;
                    int tyy = tap / d.Tw, txx = tap - tyy * d.Tw;
                    int iy = rg.iy0 + tyy * d.tap_sy, ix = rg.ix0 + txx * d.tap_sx;
                    if (iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi)
                        a[e] = ldf(in + rg.in_base + (int64_t)(iy >> d.up_shift) * d.in_sh +
                                   (int64_t)(ix >> d.up_shift) * d.in_sw + (int64_t)c * d.in_sc);
                }
            }
        }
        float b[4] = {0.f, 0.f, 0.f, 0.f};
        if (wvalid) {
            const float* p = wmat + (int64_t)wrow * d.ldw + k0 + lk;
            if (wvec && k0 + lk + 3 < K) {
                float4 v = *reinterpret_cast<const float4*>(p);
                b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (k0 + lk + e < K) b[e] = p[e];
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            As[lk + e][lr] = a[e];
            Bs[lk + e][lr] = b[e];
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float ar[4] = {av.x, av.y, av.z, av.w};
            float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t ro = row_out[ty * 4 + i];
        if (ro < 0) continue;
        const int64_t mrow = m0 + ty * 4 + i;
        const float alpha = scale ? scale[d.scale_rows > 0 ? mrow / d.scale_rows : 0] : 1.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = n0 + tx * 4 + j;
            if (co >= d.Cout) continue;
            float v = acc[i][j] * alpha + (bias ? bias[co] : 0.f);
            if (d.relu) v = fmaxf(v, 0.f);
            stf(out + ro + (int64_t)co * d.out_sc, v);
        }
    }
}

// ----------------------------------------------------------------------------------------------------------
// weight-gradient gather GEMM: R[m, tap*Cin + c] = sum_q P[q, m] * G_tap[q, c]
// grid: (ceil(Cout/64), taps * ceil(Cin/64), splits)
// ----------------------------------------------------------------------------------------------------------
template <typename TP, typename TG>
__global__ void __launch_bounds__(256) wgrad_gemm_f32_kernel(b200_conv_desc d, const TP* __restrict__ P,
                                                             const TG* __restrict__ G, float* __restrict__ ws,
                                                             int64_t rows_per_split) {
    __shared__ __align__(16) float Ps[BK][BM + 4];
    __shared__ __align__(16) float Gs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t Q = (int64_t)d.B * d.Qh * d.Qw;
    const int ctiles = (d.Cin + BN - 1) / BN;
    const int tap = blockIdx.y / ctiles;
    const int c0 = (blockIdx.y % ctiles) * BN;
    const int m0 = blockIdx.x * BM;
    const int tyy = tap / d.Tw, txx = tap % d.Tw;
    const int64_t q_begin = (int64_t)blockIdx.z * rows_per_split;
    const int64_t q_end = q_begin + rows_per_split < Q ? q_begin + rows_per_split : Q;

    // loader mapping: 16 pixels x 16 channel-quads
    const int lq = tid >> 4;
    const int lc = (tid & 15) * 4;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t q0 = q_begin; q0 < q_end; q0 += BK) {
        int64_t q = q0 + lq;
        float p[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
        if (q < q_end) {
            int qx = (int)(q % d.Qw);
            int qy = (int)((q / d.Qw) % d.Qh);
            int64_t n = q / ((int64_t)d.Qw * d.Qh);
            int oy = qy * d.out_sy + d.out_oy, ox = qx * d.out_sx + d.out_ox;
            if (oy >= 0 && oy < d.Ho && ox >= 0 && ox < d.Wo) {
                const TP* pp = P + n * d.out_sn + (int64_t)oy * d.out_sh + (int64_t)ox * d.out_sw;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (m0 + lc + e < d.Cout) p[e] = ldf(pp + (int64_t)(m0 + lc + e) * d.out_sc);
            }
            int iy = qy * d.in_sy + d.tap_oy + tyy * d.tap_sy, ix = qx * d.in_sx + d.tap_ox + txx * d.tap_sx;
            if (iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi) {
                const TG* gp = G + n * d.in_sn + (int64_t)(iy >> d.up_shift) * d.in_sh +
                               (int64_t)(ix >> d.up_shift) * d.in_sw;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (c0 + lc + e < d.Cin) g[e] = ldf(gp + (int64_t)(c0 + lc + e) * d.in_sc);
            }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&Ps[lq][lc]) = make_float4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<float4*>(&Gs[lq][lc]) = make_float4(g[0], g[1], g[2], g[3]);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 av = *reinterpret_cast<const float4*>(&Ps[kk][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Gs[kk][tx * 4]);
            float ar[4] = {av.x, av.y, av.z, av.w};
            float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }
    const int64_t Kt = (int64_t)d.Th * d.Tw * d.Cin;
    float* dst = ws + (int64_t)blockIdx.z * d.Cout * Kt;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= d.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = c0 + tx * 4 + j;
            if (c >= d.Cin) continue;
            dst[(int64_t)m * Kt + (int64_t)tap * d.Cin + c] = acc[i][j];
        }
    }
}

// ----------------------------------------------------------------------------------------------------------
// weight-gradient for a skinny gathered operand (Cin <= 8: the 3-channel image / crop side).  The 64-wide column tile
// runs over the flattened (tap, channel) index instead of one tap's channels, so a 7x7x3 layer needs 3 column tiles
// instead of 49 nearly empty ones.  grid: (ceil(Cout/64), ceil(Th*Tw*Cin/64), splits)
// ----------------------------------------------------------------------------------------------------------
template <typename TP, typename TG>
__global__ void __launch_bounds__(256) wgrad_gemm_f32_flatk_kernel(b200_conv_desc d, const TP* __restrict__ P,
                                                                   const TG* __restrict__ G, float* __restrict__ ws,
                                                                   int64_t rows_per_split) {
    __shared__ __align__(16) float Ps[BK][BM + 4];
    __shared__ __align__(16) float Gs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t Q = (int64_t)d.B * d.Qh * d.Qw;
    const int Kt = d.Th * d.Tw * d.Cin;
    const int k0 = blockIdx.y * BN;
    const int m0 = blockIdx.x * BM;
    const int64_t q_begin = (int64_t)blockIdx.z * rows_per_split;
    const int64_t q_end = q_begin + rows_per_split < Q ? q_begin + rows_per_split : Q;

    const int lq = tid >> 4;
    const int lc = (tid & 15) * 4;
    const int ty = tid >> 4, tx = tid & 15;
    // this thread's 4 gathered columns: (tap offset, channel offset), fixed over the pixel loop
    int goy[4], gox[4];
    int64_t gco[4];
    bool gok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int k = k0 + lc + e;
        gok[e] = k < Kt;
        int tap = gok[e] ? k / d.Cin : 0;
        int c = gok[e] ? k - tap * d.Cin : 0;
        int tyy = tap / d.Tw, txx = tap - tyy * d.Tw;
        goy[e] = d.tap_oy + tyy * d.tap_sy;
        gox[e] = d.tap_ox + txx * d.tap_sx;
        gco[e] = (int64_t)c * d.in_sc;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t q0 = q_begin; q0 < q_end; q0 += BK) {
        int64_t q = q0 + lq;
        float p[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
        if (q < q_end) {
            int qx = (int)(q % d.Qw);
            int qy = (int)((q / d.Qw) % d.Qh);
            int64_t n = q / ((int64_t)d.Qw * d.Qh);
            int oy = qy * d.out_sy + d.out_oy, ox = qx * d.out_sx + d.out_ox;
            if (oy >= 0 && oy < d.Ho && ox >= 0 && ox < d.Wo) {
                const TP* pp = P + n * d.out_sn + (int64_t)oy * d.out_sh + (int64_t)ox * d.out_sw;
                if (d.out_sc == 1 && m0 + lc + 3 < d.Cout && ((d.out_sn | d.out_sh | d.out_sw) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(P) & 15) == 0) {
                    float4 v = ld4(pp + m0 + lc);
                    p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (m0 + lc + e < d.Cout) p[e] = ldf(pp + (int64_t)(m0 + lc + e) * d.out_sc);
                }
            }
            const TG* gb = G + n * d.in_sn;
            const int by = qy * d.in_sy, bx = qx * d.in_sx;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int iy = by + goy[e], ix = bx + gox[e];
                if (gok[e] && iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi)
                    g[e] = ldf(gb + (int64_t)(iy >> d.up_shift) * d.in_sh + (int64_t)(ix >> d.up_shift) * d.in_sw + gco[e]);
            }
        }
        __syncthreads();
        *reinterpret_cast<float4*>(&Ps[lq][lc]) = make_float4(p[0], p[1], p[2], p[3]);
        *reinterpret_cast<float4*>(&Gs[lq][lc]) = make_float4(g[0], g[1], g[2], g[3]);
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 av = *reinterpret_cast<const float4*>(&Ps[kk][ty * 4]);
            float4 bv = *reinterpret_cast<const float4*>(&Gs[kk][tx * 4]);
            float ar[4] = {av.x, av.y, av.z, av.w};
            float br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
    }
    float* dst = ws + (int64_t)blockIdx.z * d.Cout * Kt;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= d.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + tx * 4 + j;
            if (k >= Kt) continue;
            dst[(int64_t)m * Kt + k] = acc[i][j];
        }
    }
}

}  // namespace b200

using namespace b200;

template <typename TI, typename TO>
static void launch_conv_f32(const b200_conv_desc* d, const void* in, const float* wmat, const float* bias,
                            const float* scale, void* out, dim3 grid, cudaStream_t st) {
    // 4-element vector gathers need 4-element-aligned strides and base (16 B for fp32, 8 B for bf16)
    bool vec = d->Cin % 16 == 0 && d->in_sc == 1 && d->in_sn % 4 == 0 && d->in_sh % 4 == 0 && d->in_sw % 4 == 0 &&
               (reinterpret_cast<uintptr_t>(in) & (4 * sizeof(TI) - 1)) == 0;
    if (vec)
        conv_gemm_f32_kernel<true, TI, TO><<<grid, 256, 0, st>>>(*d, (const TI*)in, wmat, bias, scale, (TO*)out);
    else
        conv_gemm_f32_kernel<false, TI, TO><<<grid, 256, 0, st>>>(*d, (const TI*)in, wmat, bias, scale, (TO*)out);
}

extern "C" int b200_conv_gemm_f32(const b200_conv_desc* d, const void* in, int in_dt, const float* wmat,
                                  const float* bias, const float* scale, void* out, int out_dt, b200_stream_t stream) {
    B200_REQUIRE(d->relu_mask == nullptr, "conv_gemm_f32: relu_mask is a tcgen05-path epilogue");
    int64_t M = (int64_t)d->B * d->Qh * d->Qw;
    if (M == 0 || d->Cout == 0) return 0;
    B200_REQUIRE(d->Cin > 0 && d->Th > 0 && d->Tw > 0, "conv_gemm_f32: bad descriptor");
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((d->Cout + BN - 1) / BN));
    B200_REQUIRE(grid.y < 65536, "conv_gemm_f32: Cout too large");
    cudaStream_t st = as_stream(stream);
    B200_DISPATCH_DT(in_dt, TI, {
        if (out_dt == B200_BF16) launch_conv_f32<TI, bf16>(d, in, wmat, bias, scale, out, grid, st);
        else launch_conv_f32<TI, float>(d, in, wmat, bias, scale, out, grid, st);
    });
    B200_CHECK_LAUNCH();
    return 0;
}

template <typename TP, typename TG>
static int launch_wgrad_f32(const b200_conv_desc* d, const void* P, const void* G, float* ws, int splits, int64_t rps,
                            cudaStream_t st) {
    if (d->Cin <= 8) {
        dim3 grid((unsigned)((d->Cout + BM - 1) / BM), (unsigned)((d->Th * d->Tw * d->Cin + BN - 1) / BN), (unsigned)splits);
        wgrad_gemm_f32_flatk_kernel<TP, TG><<<grid, 256, 0, st>>>(*d, (const TP*)P, (const TG*)G, ws, rps);
        return 0;
    }
    int ctiles = (d->Cin + BN - 1) / BN;
    dim3 grid((unsigned)((d->Cout + BM - 1) / BM), (unsigned)(d->Th * d->Tw * ctiles), (unsigned)splits);
    if (grid.y >= 65536) return set_error("wgrad_gemm_f32: too many tap tiles");
    wgrad_gemm_f32_kernel<TP, TG><<<grid, 256, 0, st>>>(*d, (const TP*)P, (const TG*)G, ws, rps);
    return 0;
}

extern "C" int b200_wgrad_gemm_f32(const b200_conv_desc* d, const void* P, int p_dt, const void* G, int g_dt, float* ws,
                                   int splits, b200_stream_t stream) {
    int64_t Q = (int64_t)d->B * d->Qh * d->Qw;
    B200_REQUIRE(splits >= 1 && splits < 65536, "wgrad_gemm_f32: bad splits");
    int64_t rps = (Q + splits - 1) / splits;
    rps = (rps + BK - 1) / BK * BK;
    if (rps < BK) rps = BK;
    cudaStream_t st = as_stream(stream);
    int rc;
    B200_DISPATCH_DT(p_dt, TP, {
        if (g_dt == B200_BF16) rc = launch_wgrad_f32<TP, bf16>(d, P, G, ws, splits, rps, st);
        else rc = launch_wgrad_f32<TP, float>(d, P, G, ws, splits, rps, st);
    });
    if (rc) return rc;
    B200_CHECK_LAUNCH();
    return 0;
}
