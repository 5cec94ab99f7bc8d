// conv_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM convolutions for sm_100a (K1/K2 forward-type, K3 weight gradient).
//
// Forward-type GEMMs (conv fwd, dgrad, ConvTranspose phases, 1x1 GEMMs over im2col-packed matrices): a 128 x BN output
// tile per accumulator, operands in the canonical K-major SWIZZLE_128B layout, bf16 x bf16 -> fp32 in TMEM.
//   conv_gemm_tc_persist_kernel (default): persistent CTAs (2 per SM at BN = 128, 3 at BN = 64) walk the tiles; warp 4
//       issues one TMA im2col load (A: 128 pixels x 64 channels of one tap) and one tiled TMA load (B: BN x 64 weights) per
//       k-block into a 3-stage ring that runs ahead across tiles; warp 5 issues the tcgen05.mma instructions into one of
//       two TMEM accumulators; warps 0-3 run the epilogue (tcgen05.ld, 1/sigma scale, bias, ReLU, optional ReLU mask of
//       the layer input, strided stores) of tile j while tile j+1 is being accumulated.
//   conv_gemm_tc_kernel: one tile per CTA; split-K launches (fp32 partials + conv_splitk_reduce_kernel) and gathers the
//       im2col tensor map cannot express (fused nearest upsampling: warps 0-3 gather with 16-byte cp.async instead).
//   conv_halo_tc_kernel: shifted-window experiment (one zero-padded slab per tile, taps as row-shifted descriptors);
//       correct but slower, disabled by default.
// The TMA / MMA issuing roles run warp-uniform code with one elected lane issuing (descriptors stay in uniform registers);
// every mbarrier wait is bounded.
// Weight-gradient kernel: R[m, c] = sum_pixels P[pix, m] * G[pix, c]; both operands are pixel-major tiles
//   (64 pixels x 128 B of channels) = the canonical MN-major SWIZZLE_128B layout, loaded by two TMA im2col maps (or
//   cp.async gathers); split over the pixel range with fp32 partials reduced by b200_wgrad_reduce / b200_sn_wgrad_finish
//   (deterministic).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace b200 {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;                 // bf16 elements = 128 bytes = one swizzle row
constexpr int A_BYTES = BM * BK * 2;   // 16 KB
constexpr int kSpinLimit = 1 << 26;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    int spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > kSpinLimit) __trap();   // a broken pipeline must not hang the GPU
    }
}
// one lane of a converged warp (the tcgen05 / TMA issue instructions take their operands from uniform registers: the
// issuing roles run warp-uniform code and predicate only the issue itself, so no per-instruction broadcast is needed)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t uniform(uint32_t x) { return __shfl_sync(0xffffffffu, x, 0); }
__device__ __forceinline__ int uniform(int x) { return __shfl_sync(0xffffffffu, x, 0); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(bar)
        : "memory");
}

// TMA im2col load: `pixelsPerColumn` consecutive output pixels (w fastest, wrapping over h and n inside the tensor map's
// bounding box) x `channelsPerPixel` channels starting at channel c; {w, h, n} = input coordinates of tap (0,0) of the
// first pixel, {off_w, off_h} = tap offsets; out-of-image taps are zero filled by the hardware
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t smem_dst, const CUtensorMap* tmap, int c, int w, int h, int n,
                                                   uint16_t off_w, uint16_t off_h, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor), M = 128, fp32 accumulate (c_format = 1 at bit 4); operand formats at
// bits [7,10) / [10,13): kind::f16 -> 1 = bf16, kind::tf32 -> 2 = tf32.  EB = operand element size in bytes (2 or 4).
template <int EB = 2>
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((EB == 2 ? 1u : 2u) << 7) | ((EB == 2 ? 1u : 2u) << 10) | ((uint32_t)a_mn_major << 15) |
           ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one tcgen05.mma of 32 bytes of K per operand row: K = 16 bf16 (kind::f16) or K = 8 tf32 (kind::tf32)
template <int EB>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (EB == 2) umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
    else umma_tf32(tmem_d, adesc, bdesc, idesc, accumulate);
}

// 16-byte LDGSTS; src_bytes = 0 zero-fills the destination (padding taps, rows beyond M)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival when every cp.async issued so far by this thread has landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

struct SmemTail {            // lives after the operand ring
    uint64_t full[8];
    uint64_t empty[8];
    uint64_t tmem_full;
    uint32_t tmem_base;
    uint32_t pad;
};

// ------------------------------------------------------------------------------------------------------------
// forward-type kernel
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_bf16x8(void* p, const float* o) {
    uint4 u;
    u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// fused ReLU backward of the layer input: zero the 16 outputs whose mask value (same address as the output) is not > 0
__device__ __forceinline__ void relu_mask16(float (&o)[16], const void* mask, int64_t off, int out_bf16) {
    if (out_bf16) {
        const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(mask) + off);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint4 q = p[h];
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (!(__uint_as_float(w[i] << 16) > 0.f)) o[h * 8 + 2 * i] = 0.f;
                if (!(__uint_as_float(w[i] & 0xFFFF0000u) > 0.f)) o[h * 8 + 2 * i + 1] = 0.f;
            }
        }
    } else {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(mask) + off);
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float4 q = p[h];
            if (!(q.x > 0.f)) o[h * 4] = 0.f;
            if (!(q.y > 0.f)) o[h * 4 + 1] = 0.f;
            if (!(q.z > 0.f)) o[h * 4 + 2] = 0.f;
            if (!(q.w > 0.f)) o[h * 4 + 3] = 0.f;
        }
    }
}
__device__ __forceinline__ bool relu_mask1(const void* mask, int64_t off, int out_bf16) {
    return out_bf16 ? (__bfloat162float(reinterpret_cast<const __nv_bfloat16*>(mask)[off]) > 0.f)
                    : (reinterpret_cast<const float*>(mask)[off] > 0.f);
}

template <int BN, int STAGES, bool ATMA, int EB = 2>
__global__ void __launch_bounds__(192) conv_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap,
                                                           const __grid_constant__ CUtensorMap tmap_a, b200_conv_desc d,
                                                           const __nv_bfloat16* __restrict__ in,
                                                           const float* __restrict__ bias,
                                                           const float* __restrict__ scale, void* __restrict__ out_v,
                                                           int out_bf16, float* __restrict__ split_ws,
                                                           int kb_per_split) {
    constexpr int BKE = 128 / EB;              // operand elements per 128-byte swizzle row (k-block = one row per pixel)
    constexpr int B_BYTES = BN * 128;
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
    static_assert(EB == 2 || ATMA, "the cp.async gather exists for bf16 operands only");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    SmemTail* tail = reinterpret_cast<SmemTail*>(smem_b + STAGES * B_BYTES);
    int64_t* row_base = reinterpret_cast<int64_t*>(tail + 1);
    int64_t* row_out = row_base + BM;
    int* row_iy = reinterpret_cast<int*>(row_out + BM);
    int* row_ix = row_iy + BM;

    const int tid = threadIdx.x;
    const int warp = uniform(tid >> 5), lane = tid & 31;
    const int64_t M = (int64_t)d.B * d.Qh * d.Qw;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int cpb = d.Cin / BKE;                    // channel blocks per tap
    const int num_kb_total = d.Th * d.Tw * cpb;
    const int kb_begin = blockIdx.z * kb_per_split;
    const int num_kb = (kb_begin + kb_per_split < num_kb_total ? kb_begin + kb_per_split : num_kb_total) - kb_begin;

    if (tid < BM) {
        int64_t m = m0 + tid;
        if (m < M) {
            int qx = (int)(m % d.Qw);
            int qy = (int)((m / d.Qw) % d.Qh);
            int64_t n = m / ((int64_t)d.Qw * d.Qh);
            row_base[tid] = n * d.in_sn;
            row_iy[tid] = qy * d.in_sy + d.tap_oy;
            row_ix[tid] = qx * d.in_sx + d.tap_ox;
            int oy = qy * d.out_sy + d.out_oy, ox = qx * d.out_sx + d.out_ox;
            row_out[tid] = (oy >= 0 && oy < d.Ho && ox >= 0 && ox < d.Wo)
                               ? n * d.out_sn + (int64_t)oy * d.out_sh + (int64_t)ox * d.out_sw
                               : -1;
        } else {
            row_base[tid] = -1;
            row_iy[tid] = row_ix[tid] = 0;
            row_out[tid] = -1;
        }
    }
    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
        if (ATMA) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    }
    if (warp == 5) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                // cp.async path: 128 cp.async completions + the TMA expect_tx arrival; im2col path: the expect_tx arrival only
                mbar_init(smem_u32(&tail->full[s]), ATMA ? 1 : 128 + 1);
                mbar_init(smem_u32(&tail->empty[s]), 1);
            }
            mbar_init(smem_u32(&tail->tmem_full), 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(&tail->tmem_base), TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp < 4) {
        // ===================== A producers =====================
        // 8 lanes cover one 128-byte row (64 channels); 16 rows per pass, 8 passes per k-block
        const int j = tid & 7;
        const int rg = tid >> 3;
        int64_t rb[8];
        int riy[8], rix[8];
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int r = p * 16 + rg;
            rb[p] = row_base[r];
            riy[p] = row_iy[r];
            rix[p] = row_ix[r];
        }
        const uint32_t dst_off = (uint32_t)rg * 128u + ((((uint32_t)j) ^ ((uint32_t)rg & 7u)) << 4);
        int tap = kb_begin / cpb, cc = kb_begin - tap * cpb;
        for (int i = 0; i < (ATMA ? 0 : num_kb); ++i) {
            const int s = i % STAGES;
            const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
            mbar_wait(smem_u32(&tail->empty[s]), ph ^ 1u);
            const int tyy = tap / d.Tw, txx = tap - tyy * d.Tw;
            const int dy = tyy * d.tap_sy, dx = txx * d.tap_sx;
            const int c0 = cc * BK + j * 8;
            const uint32_t a_dst = smem_u32(smem_a + s * A_BYTES) + dst_off;
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int iy = riy[p] + dy, ix = rix[p] + dx;
                const bool ok = rb[p] >= 0 && iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi;
                const __nv_bfloat16* src = ok ? in + rb[p] + (int64_t)(iy >> d.up_shift) * d.in_sh +
                                                    (int64_t)(ix >> d.up_shift) * d.in_sw + c0
                                              : in;
                cp_async16(a_dst + p * 2048, src, ok ? 16u : 0u);
            }
            cp_async_arrive_noinc(smem_u32(&tail->full[s]));
            if (++cc == cpb) { cc = 0; ++tap; }
        }
        // ===================== epilogue =====================
        mbar_wait(smem_u32(&tail->tmem_full), 0);
        tc_fence_after();
        const int row = warp * 32 + lane;
        if (gridDim.z > 1) {
            // split-K: raw fp32 partials, [split][M][gridDim.y * BN]; conv_splitk_reduce_kernel applies the epilogue
            const int64_t m = m0 + row;
            const int ldo = gridDim.y * BN;
            float* dst = split_ws + ((int64_t)blockIdx.z * M + m) * ldo + n0;
#pragma unroll 1
            for (int cb = 0; cb < BN; cb += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, r);
                if (m < M) {
                    float4* p = reinterpret_cast<float4*>(dst + cb);
                    p[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                    p[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
                    p[2] = make_float4(__uint_as_float(r[8]), __uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]));
                    p[3] = make_float4(__uint_as_float(r[12]), __uint_as_float(r[13]), __uint_as_float(r[14]), __uint_as_float(r[15]));
                }
            }
        } else {
            const int64_t ro = row_out[row];
            const int64_t mrow = m0 + row;
            const float alpha = scale ? scale[(d.scale_rows > 0 && mrow < M) ? mrow / d.scale_rows : 0] : 1.f;
            float* out = reinterpret_cast<float*>(out_v);
            __nv_bfloat16* outh = reinterpret_cast<__nv_bfloat16*>(out_v);
            const bool vec = d.out_sc == 1 && (d.Cout & 7) == 0 &&
                             (out_bf16 ? (((d.out_sn | d.out_sh | d.out_sw) & 7) == 0 && (reinterpret_cast<uintptr_t>(out_v) & 15) == 0)
                                       : (((d.out_sn | d.out_sh | d.out_sw) & 3) == 0 && (reinterpret_cast<uintptr_t>(out_v) & 15) == 0));
#pragma unroll 1
            for (int cb = 0; cb < BN; cb += 16) {
                uint32_t r[16];
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, r);
                if (ro >= 0) {
                    float o[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        int co = n0 + cb + e;
                        float val = __uint_as_float(r[e]) * alpha + ((bias && co < d.Cout) ? bias[co] : 0.f);
                        o[e] = d.relu ? fmaxf(val, 0.f) : val;
                    }
                    if (vec && n0 + cb + 16 <= d.Cout) {
                        if (d.relu_mask) relu_mask16(o, d.relu_mask, ro + n0 + cb, out_bf16);
                        if (out_bf16) {
                            st_bf16x8(outh + ro + n0 + cb, o);
                            st_bf16x8(outh + ro + n0 + cb + 8, o + 8);
                        } else {
                            float4* p = reinterpret_cast<float4*>(out + ro + n0 + cb);
                            p[0] = make_float4(o[0], o[1], o[2], o[3]);
                            p[1] = make_float4(o[4], o[5], o[6], o[7]);
                            p[2] = make_float4(o[8], o[9], o[10], o[11]);
                            p[3] = make_float4(o[12], o[13], o[14], o[15]);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            int co = n0 + cb + e;
                            if (co < d.Cout) {
                                const int64_t oo = ro + (int64_t)co * d.out_sc;
                                const float val = (d.relu_mask && !relu_mask1(d.relu_mask, oo, out_bf16)) ? 0.f : o[e];
                                if (out_bf16) outh[oo] = __float2bfloat16_rn(val);
                                else out[oo] = val;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 4) {
        // ===================== B producer (TMA), and the A producer on the im2col path =====================
        // warp-uniform; one elected lane issues the TMA requests
        const uint32_t bar0 = smem_u32(tail);
        const uint32_t a0 = smem_u32(smem_a), b0 = smem_u32(smem_b);
        // im2col base coordinates of the tile's first output pixel (tap (0,0) in the tensor map's offset convention)
        int aw = 0, ah = 0, an = 0, tap = 0, cc = 0, tyy = 0, txx = 0;
        if (ATMA) {
            const int qx = (int)(m0 % d.Qw);
            const int qy = (int)((m0 / d.Qw) % d.Qh);
            an = (int)(m0 / ((int64_t)d.Qw * d.Qh));
            aw = qx * d.in_sx + (d.tap_sx > 0 ? d.tap_ox : d.tap_ox - (d.Tw - 1));
            ah = qy * d.in_sy + (d.tap_sy > 0 ? d.tap_oy : d.tap_oy - (d.Th - 1));
            tap = kb_begin / cpb;
            cc = kb_begin - tap * cpb;
            tyy = tap / d.Tw;
            txx = tap - tyy * d.Tw;
        }
        uint32_t s = 0, ph = 0;
        for (int i = 0; i < num_kb; ++i) {
            mbar_wait(bar0 + (uint32_t)offsetof(SmemTail, empty) + 8u * s, ph ^ 1u);
            const uint16_t ow = (uint16_t)(d.tap_sx > 0 ? txx : d.Tw - 1 - txx);
            const uint16_t oh = (uint16_t)(d.tap_sy > 0 ? tyy : d.Th - 1 - tyy);
            if (elect_one()) {
                const uint32_t bar = bar0 + (uint32_t)offsetof(SmemTail, full) + 8u * s;
                mbar_arrive_expect_tx(bar, ATMA ? A_BYTES + B_BYTES : B_BYTES);
                if (ATMA) tma_load_im2col_4d(a0 + s * A_BYTES, &tmap_a, cc * BKE, aw, ah, an, ow, oh, bar);
                tma_load_2d(b0 + s * B_BYTES, &tmap, (kb_begin + i) * BKE, n0, bar);
            }
            if (ATMA && ++cc == cpb) {
                cc = 0;
                if (++txx == d.Tw) { txx = 0; ++tyy; }
            }
            if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
    } else {
        // ===================== MMA issuer (warp-uniform; one elected lane issues) =====================
        constexpr uint32_t idesc = make_idesc<EB>(BN, 0, 0);
        const uint64_t desc_hi = make_desc(0, 16, 1024);
        const uint32_t bar0 = smem_u32(tail);
        const uint32_t a0 = smem_u32(smem_a) >> 4, b0 = smem_u32(smem_b) >> 4;
        uint32_t s = 0, ph = 0;
        for (int i = 0; i < num_kb; ++i) {
            mbar_wait(bar0 + (uint32_t)offsetof(SmemTail, full) + 8u * s, ph);
            tc_fence_after();
            const uint64_t adesc = desc_hi | (uint64_t)((a0 + s * (A_BYTES >> 4)) & 0x3FFFu);
            const uint64_t bdesc = desc_hi | (uint64_t)((b0 + s * (B_BYTES >> 4)) & 0x3FFFu);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)         // 4 x 32 bytes of K per 128-byte row
                    umma<EB>(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)(i | k));
                umma_commit(bar0 + (uint32_t)offsetof(SmemTail, empty) + 8u * s);
                if (i == num_kb - 1) umma_commit(bar0 + (uint32_t)offsetof(SmemTail, tmem_full));
            }
            if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// persistent forward-type kernel (TMA im2col operands, no split-K): one CTA per SM walks the output tiles
// t = blockIdx.x, blockIdx.x + gridDim.x, ... (N tile fastest, so concurrently running CTAs share the activation tile
// in L2).  The operand ring runs ahead across tile boundaries and two TMEM accumulators alternate, so the epilogue of
// tile j (warps 0-3: tcgen05.ld, scale / bias / ReLU, store) overlaps the loads and MMAs of tile j+1.
//   warp 4 lane 0 : producer — per k-block one im2col TMA (A, 128 pixels x 64 channels) + one tiled TMA (B, BN x 64)
//   warp 5 lane 0 : tcgen05.mma issuer; tcgen05.commit frees the stage / publishes the accumulator
//   warps 0-3     : epilogue
// ------------------------------------------------------------------------------------------------------------
// column totals of a 32-row x 16-column register tile: v[e] = this lane's (row's) value of column e.  Four exchange rounds
// halve the columns a lane still carries (lane bit 4 picks the upper / lower 8, bit 3 the next 4, ...), a fifth adds the two
// lanes of a pair: afterwards v[0] of lane l is the total of column l >> 1 over the 32 rows.  16 shuffles (not 16 x 5).
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
    for (int off = 16, n = 8; n >= 1; off >>= 1, n >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < n) {
                const float send = up ? v[i] : v[i + n];
                const float keep = up ? v[i + n] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

struct PersistTail {
    uint64_t full[8];
    uint64_t empty[8];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN, int STAGES, int EB = 2>
__global__ void __launch_bounds__(192, 3) conv_gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                      const __grid_constant__ CUtensorMap tmap_a,
                                                                      b200_conv_desc d, const float* __restrict__ bias,
                                                                      const float* __restrict__ scale,
                                                                      void* __restrict__ out_v, int out_bf16,
                                                                      int num_n_tiles, int num_tiles) {
    constexpr int BKE = 128 / EB;
    constexpr int B_BYTES = BN * 128;
    constexpr int ACC_COLS = BN < 16 ? 16 : BN;
    constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
    constexpr int CW = BN >= 32 ? 32 : 16;      // accumulator columns per tcgen05.ld
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    PersistTail* tail = reinterpret_cast<PersistTail*>(smem_b + STAGES * B_BYTES);

    const int tid = threadIdx.x;
    const int warp = uniform(tid >> 5), lane = tid & 31;
    const int64_t M = (int64_t)d.B * d.Qh * d.Qw;
    const int cpb = d.Cin / BKE;
    const int num_kb = d.Th * d.Tw * cpb;

    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    }
    if (warp == 5) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(smem_u32(&tail->full[s]), 1);
                mbar_init(smem_u32(&tail->empty[s]), 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(smem_u32(&tail->tmem_full[b]), 1);
                mbar_init(smem_u32(&tail->tmem_empty[b]), 4);      // one arrival per epilogue warp
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(&tail->tmem_base), TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp < 4) {
        // ===================== epilogue =====================
        const int row = warp * 32 + lane;
        float* out = reinterpret_cast<float*>(out_v);
        __nv_bfloat16* outh = reinterpret_cast<__nv_bfloat16*>(out_v);
        const bool vec = d.out_sc == 1 && (d.Cout & 7) == 0 &&
                         (out_bf16 ? (((d.out_sn | d.out_sh | d.out_sw) & 7) == 0 && (reinterpret_cast<uintptr_t>(out_v) & 15) == 0)
                                   : (((d.out_sn | d.out_sh | d.out_sw) & 3) == 0 && (reinterpret_cast<uintptr_t>(out_v) & 15) == 0));
        int j = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++j) {
            const int buf = j & 1;
            const int n0 = (t % num_n_tiles) * BN;
            const uint32_t m = (uint32_t)(t / num_n_tiles) * BM + (uint32_t)row;      // M < 2^31 (checked at launch)
            int64_t ro = -1;
            float alpha = 1.f;
            if (m < (uint32_t)M) {
                const uint32_t q = m / (uint32_t)d.Qw;
                const int qx = (int)(m - q * (uint32_t)d.Qw);
                const uint32_t n = q / (uint32_t)d.Qh;
                const int qy = (int)(q - n * (uint32_t)d.Qh);
                const int oy = qy * d.out_sy + d.out_oy, ox = qx * d.out_sx + d.out_ox;
                if (oy >= 0 && oy < d.Ho && ox >= 0 && ox < d.Wo)
                    ro = (int64_t)n * d.out_sn + (int64_t)oy * d.out_sh + (int64_t)ox * d.out_sw;
                if (scale) alpha = scale[d.scale_rows > 0 ? m / (uint32_t)d.scale_rows : 0];
            }
            mbar_wait(smem_u32(&tail->tmem_full[buf]), (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * ACC_COLS);
#pragma unroll 1
            for (int cb = 0; cb < BN; cb += CW) {
                uint32_t r[CW];
                if constexpr (CW == 32) tmem_ld32(tacc + (uint32_t)cb, r);
                else tmem_ld16(tacc + (uint32_t)cb, r);
                const bool stats = d.col_stats != nullptr;          // (uniform)
                if (ro >= 0 || stats) {
#pragma unroll
                    for (int h = 0; h < CW; h += 16) {
                        float o[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const int co = n0 + cb + h + e;
                            const float val = __uint_as_float(r[h + e]) * alpha + ((bias && co < d.Cout) ? bias[co] : 0.f);
                            o[e] = d.relu ? fmaxf(val, 0.f) : val;
                        }
                        const int c0 = n0 + cb + h;
                        if (stats) {
                            // batch-norm statistics of the STORED values, per 32-row slab (= this warp's rows): every lane takes
                            // part in the shuffles; rows outside the output contribute zeros
                            float sv[16], sq[16];
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                float q = out_bf16 ? __bfloat162float(__float2bfloat16_rn(o[e])) : o[e];
                                if (ro < 0) q = 0.f;
                                sv[e] = q;
                                sq[e] = q * q;
                            }
                            const float ts = warp_colsum16(sv, lane), tq = warp_colsum16(sq, lane);
                            const int col = c0 + (lane >> 1);
                            if ((lane & 1) == 0 && col < d.col_stats_ld) {
                                const int64_t slab = (int64_t)(t / num_n_tiles) * 4 + warp;
                                *reinterpret_cast<float2*>(d.col_stats + (slab * d.col_stats_ld + col) * 2) = make_float2(ts, tq);
                            }
                        }
                        if (ro < 0) continue;
                        if (vec && c0 + 16 <= d.Cout) {
                            if (d.relu_mask) relu_mask16(o, d.relu_mask, ro + c0, out_bf16);
                            if (out_bf16) {
                                st_bf16x8(outh + ro + c0, o);
                                st_bf16x8(outh + ro + c0 + 8, o + 8);
                            } else {
                                float4* p = reinterpret_cast<float4*>(out + ro + c0);
                                p[0] = make_float4(o[0], o[1], o[2], o[3]);
                                p[1] = make_float4(o[4], o[5], o[6], o[7]);
                                p[2] = make_float4(o[8], o[9], o[10], o[11]);
                                p[3] = make_float4(o[12], o[13], o[14], o[15]);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const int co = c0 + e;
                                if (co < d.Cout) {
                                    const int64_t oo = ro + (int64_t)co * d.out_sc;
                                    const float val = (d.relu_mask && !relu_mask1(d.relu_mask, oo, out_bf16)) ? 0.f : o[e];
                                    if (out_bf16) outh[oo] = __float2bfloat16_rn(val);
                                    else out[oo] = val;
                                }
                            }
                        }
                    }
                }
            }
            // the accumulator has been read: hand the TMEM buffer back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tail->tmem_empty[buf]));
        }
    } else if (warp == 4) {
        // ===================== producer (warp-uniform; one elected lane issues the TMA requests) =====================
        const uint32_t bar0 = smem_u32(tail);
        const uint32_t a0 = smem_u32(smem_a), b0 = smem_u32(smem_b);
        const int lw = d.tap_sx > 0 ? d.tap_ox : d.tap_ox - (d.Tw - 1);
        const int lh = d.tap_sy > 0 ? d.tap_oy : d.tap_oy - (d.Th - 1);
        uint32_t s = 0, ph = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int n0 = (t % num_n_tiles) * BN;
            const int64_t m0 = (int64_t)(t / num_n_tiles) * BM;
            const int qx = (int)(m0 % d.Qw);
            const int qy = (int)((m0 / d.Qw) % d.Qh);
            const int an = (int)(m0 / ((int64_t)d.Qw * d.Qh));
            const int aw = qx * d.in_sx + lw;
            const int ah = qy * d.in_sy + lh;
            int kb = 0;
            for (int ty = 0; ty < d.Th; ++ty) {
                const uint16_t oh = (uint16_t)(d.tap_sy > 0 ? ty : d.Th - 1 - ty);
                for (int tx = 0; tx < d.Tw; ++tx) {
                    const uint16_t ow = (uint16_t)(d.tap_sx > 0 ? tx : d.Tw - 1 - tx);
                    for (int cc = 0; cc < cpb; ++cc, ++kb) {
                        mbar_wait(bar0 + (uint32_t)offsetof(PersistTail, empty) + 8u * s, ph ^ 1u);
                        if (elect_one()) {
                            const uint32_t bar = bar0 + (uint32_t)offsetof(PersistTail, full) + 8u * s;
                            mbar_arrive_expect_tx(bar, A_BYTES + B_BYTES);
                            tma_load_im2col_4d(a0 + s * A_BYTES, &tmap_a, cc * BKE, aw, ah, an, ow, oh, bar);
                            tma_load_2d(b0 + s * B_BYTES, &tmap, kb * BKE, n0, bar);
                        }
                        if (++s == STAGES) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else {
        // ===================== MMA issuer (warp-uniform; one elected lane issues) =====================
        constexpr uint32_t idesc = make_idesc<EB>(BN, 0, 0);
        const uint64_t desc_hi = make_desc(0, 16, 1024);
        const uint32_t bar0 = smem_u32(tail);
        const uint32_t a0 = smem_u32(smem_a) >> 4, b0 = smem_u32(smem_b) >> 4;
        uint32_t s = 0, ph = 0;
        int j = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++j) {
            const int buf = j & 1;
            mbar_wait(bar0 + (uint32_t)offsetof(PersistTail, tmem_empty) + 8u * buf, ((uint32_t)(j >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(buf * ACC_COLS);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(bar0 + (uint32_t)offsetof(PersistTail, full) + 8u * s, ph);
                tc_fence_after();
                const uint64_t adesc = desc_hi | (uint64_t)((a0 + s * (A_BYTES >> 4)) & 0x3FFFu);
                const uint64_t bdesc = desc_hi | (uint64_t)((b0 + s * (B_BYTES >> 4)) & 0x3FFFu);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma<EB>(tacc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (uint32_t)(kb | k));
                    umma_commit(bar0 + (uint32_t)offsetof(PersistTail, empty) + 8u * s);
                    if (kb == num_kb - 1) umma_commit(bar0 + (uint32_t)offsetof(PersistTail, tmem_full) + 8u * buf);
                }
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// shifted-window kernel for stride-1 k x k convolutions (forward and their dgrad): the activation is read ONCE.
//
// Output pixels of one image are numbered along PADDED rows of width Wp = Qw + Tw - 1: v = qy * Wp + qxv (columns
// qxv >= Qw are dummies, dropped by the epilogue).  For tap (ty, tx) the input pixel of output v is then
// u = v + dy * Wp + dx — a constant row offset — in the zero-padded input plane P (rows of Wp pixels).  So a tile of 128
// consecutive v needs, for ALL taps, one contiguous slab of P rows [r0, r0 + RB): a single tiled TMA box
// {64 channels, Wp pixels, RB rows} with hardware zero fill outside the image, written to shared memory in the K-major
// SWIZZLE_128B layout (one 128-byte row per pixel).  The A operand of tap (ty, tx) is that slab read from row
// voff + dy*Wp + dx on: tcgen05.mma applies the swizzle to absolute shared-memory addresses, so a descriptor whose
// start address is shifted by whole 128-byte rows addresses the shifted window directly (tools/probes/
// umma_rowshift_probe.cu verifies this on the hardware).  Per tile the activation traffic drops from
// taps x 16 KB to one slab (e.g. 144 KB -> 35 KB for a 3x3 layer at 32x32) — these layers are L2->SM bandwidth
// bound, not tensor bound.  Persistent CTAs, weights streamed through a TMA ring, double-buffered slab and TMEM.
// ------------------------------------------------------------------------------------------------------------
struct HaloTail {
    uint64_t a_full[4];
    uint64_t a_empty[4];
    uint64_t b_full[8];
    uint64_t b_empty[8];
    uint64_t b_res_full;
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
};

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* tmap, int c, int w, int h, int n,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
        : "memory");
}

struct HaloGeom {
    int Wp, RB;             // padded row width, slab rows
    int lw, lh;             // input coordinates of P[0][0]
    int slab_bytes;         // RB * Wp * 128 rounded up to 1024
    int slab_tx;            // RB * Wp * 128
    int na;                 // slab buffers (2..4)
    int b_stages;           // weight ring stages (streaming mode)
    int b_resident;         // 1: the CTA keeps its N tile's whole weight matrix in shared memory
    int tiles_per_img;
    int num_n_tiles, num_m_tiles;
};

// tile i of CTA b: streaming mode walks t = b + i*G with the N tile fastest; resident mode pins the CTA to N tile
// b % NT and walks the M tiles b / NT + i * (G / NT)   (G % NT == 0)
__device__ __forceinline__ bool halo_tile(const HaloGeom& hg, int i, int& mt, int& nt) {
    if (hg.b_resident) {
        nt = (int)blockIdx.x % hg.num_n_tiles;
        mt = (int)blockIdx.x / hg.num_n_tiles + i * ((int)gridDim.x / hg.num_n_tiles);
    } else {
        const int64_t t = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
        nt = (int)(t % hg.num_n_tiles);
        const int64_t m = t / hg.num_n_tiles;
        mt = m > 0x7fffffff ? 0x7fffffff : (int)m;
    }
    return mt < hg.num_m_tiles;
}

template <int BN>
__global__ void __launch_bounds__(192, 1) conv_halo_tc_kernel(const __grid_constant__ CUtensorMap tmap,
                                                              const __grid_constant__ CUtensorMap tmap_in,
                                                              b200_conv_desc d, HaloGeom hg,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ scale,
                                                              void* __restrict__ out_v, int out_bf16) {
    constexpr int B_BYTES = BN * BK * 2;
    constexpr int ACC_COLS = BN < 16 ? 16 : BN;
    constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
    constexpr int CW = BN >= 32 ? 32 : 16;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int cpb = d.Cin / BK;
    const int taps = d.Th * d.Tw;
    const int num_kb = taps * cpb;
    uint8_t* smem_a = smem;                                   // na slabs
    uint8_t* smem_b = smem + hg.na * hg.slab_bytes;           // ring (b_stages) or resident (num_kb) x B_BYTES
    HaloTail* tail = reinterpret_cast<HaloTail*>(smem_b + (hg.b_resident ? num_kb : hg.b_stages) * B_BYTES);

    const int tid = threadIdx.x;
    const int warp = uniform(tid >> 5), lane = tid & 31;
    const uint32_t SB = (uint32_t)hg.b_stages, NA = (uint32_t)hg.na;

    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_in)) : "memory");
    }
    if (warp == 5) {
        if (lane == 0) {
            for (int b = 0; b < 4; ++b) {
                mbar_init(smem_u32(&tail->a_full[b]), 1);
                mbar_init(smem_u32(&tail->a_empty[b]), 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(smem_u32(&tail->tmem_full[b]), 1);
                mbar_init(smem_u32(&tail->tmem_empty[b]), 4);
            }
            for (int s = 0; s < 8; ++s) {
                mbar_init(smem_u32(&tail->b_full[s]), 1);
                mbar_init(smem_u32(&tail->b_empty[s]), 1);
            }
            mbar_init(smem_u32(&tail->b_res_full), 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(&tail->tmem_base), TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp < 4) {
        // ===================== epilogue =====================
        const int row = warp * 32 + lane;
        float* out = reinterpret_cast<float*>(out_v);
        __nv_bfloat16* outh = reinterpret_cast<__nv_bfloat16*>(out_v);
        const bool vec = d.out_sc == 1 && (d.Cout & 7) == 0 &&
                         (out_bf16 ? (((d.out_sn | d.out_sh | d.out_sw) & 7) == 0 && (reinterpret_cast<uintptr_t>(out_v) & 15) == 0)
                                   : (((d.out_sn | d.out_sh | d.out_sw) & 3) == 0 && (reinterpret_cast<uintptr_t>(out_v) & 15) == 0));
        int mt, nt;
        for (int j = 0; halo_tile(hg, j, mt, nt); ++j) {
            const int buf = j & 1;
            const int n0 = nt * BN;
            const int img = mt / hg.tiles_per_img;
            const int v = (mt - img * hg.tiles_per_img) * BM + row;
            const int qy = v / hg.Wp, qx = v - qy * hg.Wp;
            int64_t ro = -1;
            float alpha = 1.f;
            if (qy < d.Qh && qx < d.Qw) {
                ro = (int64_t)img * d.out_sn + (int64_t)qy * d.out_sh + (int64_t)qx * d.out_sw;
                if (scale) {
                    const int64_t m = ((int64_t)img * d.Qh + qy) * d.Qw + qx;
                    alpha = scale[d.scale_rows > 0 ? m / d.scale_rows : 0];
                }
            }
            mbar_wait(smem_u32(&tail->tmem_full[buf]), (uint32_t)(j >> 1) & 1u);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * ACC_COLS);
#pragma unroll 1
            for (int cb = 0; cb < BN; cb += CW) {
                uint32_t r[CW];
                if constexpr (CW == 32) tmem_ld32(tacc + (uint32_t)cb, r);
                else tmem_ld16(tacc + (uint32_t)cb, r);
                if (ro >= 0) {
#pragma unroll
                    for (int h = 0; h < CW; h += 16) {
                        float o[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const int co = n0 + cb + h + e;
                            const float val = __uint_as_float(r[h + e]) * alpha + ((bias && co < d.Cout) ? bias[co] : 0.f);
                            o[e] = d.relu ? fmaxf(val, 0.f) : val;
                        }
                        const int c0 = n0 + cb + h;
                        if (vec && c0 + 16 <= d.Cout) {
                            if (d.relu_mask) relu_mask16(o, d.relu_mask, ro + c0, out_bf16);
                            if (out_bf16) {
                                st_bf16x8(outh + ro + c0, o);
                                st_bf16x8(outh + ro + c0 + 8, o + 8);
                            } else {
                                float4* p = reinterpret_cast<float4*>(out + ro + c0);
                                p[0] = make_float4(o[0], o[1], o[2], o[3]);
                                p[1] = make_float4(o[4], o[5], o[6], o[7]);
                                p[2] = make_float4(o[8], o[9], o[10], o[11]);
                                p[3] = make_float4(o[12], o[13], o[14], o[15]);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const int co = c0 + e;
                                if (co < d.Cout) {
                                    const int64_t oo = ro + (int64_t)co * d.out_sc;
                                    const float val = (d.relu_mask && !relu_mask1(d.relu_mask, oo, out_bf16)) ? 0.f : o[e];
                                    if (out_bf16) outh[oo] = __float2bfloat16_rn(val);
                                    else out[oo] = val;
                                }
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tail->tmem_empty[buf]));
        }
    } else if (warp == 4) {
        // ===================== producer (warp-uniform; one elected lane issues the TMA requests) =====================
        const uint32_t bar0 = smem_u32(tail);
        const uint32_t sa0 = smem_u32(smem_a), sb0 = smem_u32(smem_b);
        uint32_t ab = 0, aph = 0, bs = 0, bph = 0;
        int mt, nt;
        if (hg.b_resident && halo_tile(hg, 0, mt, nt)) {
            if (elect_one()) {
                const uint32_t bbar = bar0 + (uint32_t)offsetof(HaloTail, b_res_full);
                mbar_arrive_expect_tx(bbar, (uint32_t)(num_kb * B_BYTES));
                for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(sb0 + kb * B_BYTES, &tmap, kb * BK, nt * BN, bbar);
            }
            __syncwarp();
        }
        for (int i = 0; halo_tile(hg, i, mt, nt); ++i) {
            const int n0 = nt * BN;
            const int img = mt / hg.tiles_per_img;
            const int v0 = (mt - img * hg.tiles_per_img) * BM;
            const int r0 = v0 / hg.Wp;
            for (int cc = 0; cc < cpb; ++cc) {
                mbar_wait(bar0 + (uint32_t)offsetof(HaloTail, a_empty) + 8u * ab, aph ^ 1u);
                if (elect_one()) {
                    const uint32_t abar = bar0 + (uint32_t)offsetof(HaloTail, a_full) + 8u * ab;
                    mbar_arrive_expect_tx(abar, (uint32_t)hg.slab_tx);
                    tma_load_4d(sa0 + ab * (uint32_t)hg.slab_bytes, &tmap_in, cc * BK, hg.lw, hg.lh + r0, img, abar);
                }
                if (++ab == NA) { ab = 0; aph ^= 1u; }
                if (!hg.b_resident) {
                    for (int tap = 0; tap < taps; ++tap) {
                        mbar_wait(bar0 + (uint32_t)offsetof(HaloTail, b_empty) + 8u * bs, bph ^ 1u);
                        if (elect_one()) {
                            const uint32_t bbar = bar0 + (uint32_t)offsetof(HaloTail, b_full) + 8u * bs;
                            mbar_arrive_expect_tx(bbar, B_BYTES);
                            tma_load_2d(sb0 + bs * B_BYTES, &tmap, (tap * cpb + cc) * BK, n0, bbar);
                        }
                        if (++bs == SB) { bs = 0; bph ^= 1u; }
                    }
                }
            }
        }
    } else {
        // ===================== MMA issuer (warp-uniform; one elected lane issues) =====================
        constexpr uint32_t idesc = make_idesc(BN, 0, 0);
        const uint64_t desc_hi = make_desc(0, 16, 1024);
        const uint32_t slab0 = smem_u32(smem_a) >> 4, slab_step = (uint32_t)hg.slab_bytes >> 4;
        const uint32_t b0 = smem_u32(smem_b) >> 4;
        const uint32_t bar0 = smem_u32(tail);
        // 16-byte units: one pixel row of the slab = 8
        const int row_step = d.tap_sy * hg.Wp * 8, col_step = d.tap_sx * 8;
        const int off00 = ((d.tap_oy - hg.lh) * hg.Wp + (d.tap_ox - hg.lw)) * 8;
        uint32_t ab = 0, aph = 0, bs = 0, bph = 0;
        int mt, nt;
        if (hg.b_resident && halo_tile(hg, 0, mt, nt)) {
            mbar_wait(bar0 + (uint32_t)offsetof(HaloTail, b_res_full), 0);
            tc_fence_after();
        }
        for (int j = 0; halo_tile(hg, j, mt, nt); ++j) {
            const int buf = j & 1;
            const int img = mt / hg.tiles_per_img;
            const int v0 = (mt - img * hg.tiles_per_img) * BM;
            const int voff = v0 % hg.Wp;
            mbar_wait(bar0 + (uint32_t)offsetof(HaloTail, tmem_empty) + 8u * buf, ((uint32_t)(j >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(buf * ACC_COLS);
            uint32_t acc = 0;
            for (int cc = 0; cc < cpb; ++cc) {
                mbar_wait(bar0 + (uint32_t)offsetof(HaloTail, a_full) + 8u * ab, aph);
                tc_fence_after();
                const uint32_t a_base = slab0 + ab * slab_step + (uint32_t)(voff * 8 + off00);
                uint32_t b_res = b0 + (uint32_t)cc * (B_BYTES >> 4);
                int row_off = 0;
                for (int ty = 0; ty < d.Th; ++ty, row_off += row_step) {
                    int off = row_off;
                    for (int tx = 0; tx < d.Tw; ++tx, off += col_step) {
                        uint32_t b_lo;
                        if (hg.b_resident) {
                            b_lo = b_res;
                            b_res += (uint32_t)cpb * (B_BYTES >> 4);
                        } else {
                            mbar_wait(bar0 + (uint32_t)offsetof(HaloTail, b_full) + 8u * bs, bph);
                            tc_fence_after();
                            b_lo = b0 + bs * (B_BYTES >> 4);
                        }
                        const uint64_t adesc = desc_hi | (uint64_t)((a_base + (uint32_t)off) & 0x3FFFu);
                        const uint64_t bdesc = desc_hi | (uint64_t)(b_lo & 0x3FFFu);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_bf16(tacc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, acc | (uint32_t)k);
                            if (!hg.b_resident) umma_commit(bar0 + (uint32_t)offsetof(HaloTail, b_empty) + 8u * bs);
                        }
                        acc = 1;
                        if (!hg.b_resident && ++bs == SB) { bs = 0; bph ^= 1u; }
                    }
                }
                if (elect_one()) {
                    umma_commit(bar0 + (uint32_t)offsetof(HaloTail, a_empty) + 8u * ab);
                    if (cc == cpb - 1) umma_commit(bar0 + (uint32_t)offsetof(HaloTail, tmem_full) + 8u * buf);
                }
                if (++ab == NA) { ab = 0; aph ^= 1u; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, TMEM_COLS);
}

// split-K second pass: out = epilogue( sum_z ws[z][m][co] ), fixed summation order
__global__ void conv_splitk_reduce_kernel(b200_conv_desc d, const float* __restrict__ ws, int splits, int ldo,
                                          const float* __restrict__ bias, const float* __restrict__ scale,
                                          void* __restrict__ out_v, int out_bf16) {
    const int64_t M = (int64_t)d.B * d.Qh * d.Qw;
    const int C4 = (d.Cout + 3) >> 2;
    const int64_t total = M * C4;
    float* out = reinterpret_cast<float*>(out_v);
    __nv_bfloat16* outh = reinterpret_cast<__nv_bfloat16*>(out_v);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int co = (int)(t % C4) << 2;
        const int64_t m = t / C4;
        const float alpha = scale ? scale[d.scale_rows > 0 ? m / d.scale_rows : 0] : 1.f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int z = 0; z < splits; ++z) {
            float4 v = *reinterpret_cast<const float4*>(ws + ((int64_t)z * M + m) * ldo + co);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        int qx = (int)(m % d.Qw);
        int qy = (int)((m / d.Qw) % d.Qh);
        int64_t n = m / ((int64_t)d.Qw * d.Qh);
        int oy = qy * d.out_sy + d.out_oy, ox = qx * d.out_sx + d.out_ox;
        if (oy < 0 || oy >= d.Ho || ox < 0 || ox >= d.Wo) continue;
        const int64_t ro = n * d.out_sn + (int64_t)oy * d.out_sh + (int64_t)ox * d.out_sw;
        float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (co + e >= d.Cout) break;
            float v = o[e] * alpha + (bias ? bias[co + e] : 0.f);
            if (d.relu) v = fmaxf(v, 0.f);
            const int64_t oo = ro + (int64_t)(co + e) * d.out_sc;
            if (d.relu_mask && !relu_mask1(d.relu_mask, oo, out_bf16)) v = 0.f;
            if (out_bf16) outh[oo] = __float2bfloat16_rn(v);
            else out[oo] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// weight-gradient kernel: R[m0:m0+128, tap, c0:c0+BNW] over a pixel split
// ------------------------------------------------------------------------------------------------------------
struct PixInfo {
    int64_t p_off;   // offset of P row (or -1)
    int64_t g_off;   // offset of G pixel for this tap (or -1)
};

template <int BNW, int STAGES, bool WTMA, int EB = 2>
__global__ void __launch_bounds__(192) wgrad_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_p,
                                                            const __grid_constant__ CUtensorMap tmap_g,
                                                            b200_conv_desc d, const __nv_bfloat16* __restrict__ P,
                                                            const __nv_bfloat16* __restrict__ G,
                                                            float* __restrict__ ws, int64_t rows_per_split) {
    constexpr int BKE = 128 / EB;                      // channels per MN atom row (128 bytes): 64 bf16 / 32 tf32
    constexpr int PA_ATOMS = 128 / BKE, GB_ATOMS = BNW / BKE;
    constexpr int PA_BYTES = PA_ATOMS * 64 * 128;      // 128 P channels x 64 pixels (one 8 KB MN atom column per BKE channels)
    constexpr int GB_BYTES = GB_ATOMS * 64 * 128;      // BNW G channels x 64 pixels
    constexpr int KMMA = 32 / EB;                      // pixels (K) per tcgen05.mma: 16 bf16 / 8 tf32
    static_assert(EB == 2 || WTMA, "the cp.async gather exists for bf16 operands only");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * PA_BYTES;
    SmemTail* tail = reinterpret_cast<SmemTail*>(smem_b + STAGES * GB_BYTES);
    PixInfo* pix = reinterpret_cast<PixInfo*>(tail + 1);   // [STAGES][64]

    const int tid = threadIdx.x;
    const int warp = uniform(tid >> 5), lane = tid & 31;
    const int64_t Q = (int64_t)d.B * d.Qh * d.Qw;
    const int ctiles = (d.Cin + BNW - 1) / BNW;
    const int tap = blockIdx.y / ctiles;
    const int c0 = (blockIdx.y % ctiles) * BNW;
    const int m0 = blockIdx.x * BM;
    const int tyy = tap / d.Tw, txx = tap % d.Tw;
    const int64_t q_begin = (int64_t)blockIdx.z * rows_per_split;
    const int64_t q_end = q_begin + rows_per_split < Q ? q_begin + rows_per_split : Q;
    const int num_kb = q_end > q_begin ? (int)((q_end - q_begin + 63) / 64) : 0;

    if (WTMA && warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_p)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_g)) : "memory");
    }
    if (warp == 5) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(smem_u32(&tail->full[s]), WTMA ? 1 : 128);
                mbar_init(smem_u32(&tail->empty[s]), 1);
            }
            mbar_init(smem_u32(&tail->tmem_full), 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(&tail->tmem_base), BNW < 32 ? 32 : BNW);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp < 4) {
        // 16 lanes cover the 256 bytes (128 channels) of one pixel; 8 pixels per pass, 8 passes per 64-pixel k-block
        const int j16 = tid & 15;
        const int pg = tid >> 4;
        const uint32_t p_dst_off = (uint32_t)(j16 >> 3) * 8192u + (uint32_t)pg * 128u + ((((uint32_t)j16 & 7u) ^ (uint32_t)pg) << 4);
        const int pch = m0 + j16 * 8;             // first of this lane's 8 P channels
        const bool p_ch_ok = pch < d.Cout;
        // G tile: BNW == 128 -> same mapping; BNW == 64 -> 8 lanes per pixel, 16 pixels per pass, 4 passes
        const int gj = BNW == 128 ? j16 : (tid & 7);
        const int gpg = BNW == 128 ? pg : (tid >> 3);
        const uint32_t g_dst_off = BNW == 128 ? p_dst_off
                                              : ((uint32_t)gpg * 128u + ((((uint32_t)gj) ^ ((uint32_t)gpg & 7u)) << 4));
        const int gch = c0 + gj * 8;
        const bool g_ch_ok = gch < d.Cin;
        for (int kb = 0; kb < (WTMA ? 0 : num_kb); ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
            mbar_wait(smem_u32(&tail->empty[s]), ph ^ 1u);
            PixInfo* pi = pix + s * 64;
            if (tid < 64) {
                int64_t q = q_begin + (int64_t)kb * 64 + tid;
                PixInfo info;
                info.p_off = -1;
                info.g_off = -1;
                if (q < q_end) {
                    int qx = (int)(q % d.Qw);
                    int qy = (int)((q / d.Qw) % d.Qh);
                    int64_t n = q / ((int64_t)d.Qw * d.Qh);
                    int oy = qy * d.out_sy + d.out_oy, ox = qx * d.out_sx + d.out_ox;
                    if (oy >= 0 && oy < d.Ho && ox >= 0 && ox < d.Wo)
                        info.p_off = n * d.out_sn + (int64_t)oy * d.out_sh + (int64_t)ox * d.out_sw;
                    int iy = qy * d.in_sy + d.tap_oy + tyy * d.tap_sy, ix = qx * d.in_sx + d.tap_ox + txx * d.tap_sx;
                    if (iy >= 0 && iy < d.Hi && ix >= 0 && ix < d.Wi)
                        info.g_off = n * d.in_sn + (int64_t)(iy >> d.up_shift) * d.in_sh +
                                     (int64_t)(ix >> d.up_shift) * d.in_sw;
                }
                pi[tid] = info;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            {
                const uint32_t a_dst = smem_u32(smem_a + s * PA_BYTES) + p_dst_off;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int64_t off = pi[p * 8 + pg].p_off;
                    const bool ok = off >= 0 && p_ch_ok;
                    cp_async16(a_dst + p * 1024, ok ? P + off + pch : P, ok ? 16u : 0u);
                }
            }
            {
                const uint32_t b_dst = smem_u32(smem_b + s * GB_BYTES) + g_dst_off;
                if (BNW == 128) {
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const int64_t off = pi[p * 8 + gpg].g_off;
                        const bool ok = off >= 0 && g_ch_ok;
                        cp_async16(b_dst + p * 1024, ok ? G + off + gch : G, ok ? 16u : 0u);
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int64_t off = pi[p * 16 + gpg].g_off;
                        const bool ok = off >= 0 && g_ch_ok;
                        cp_async16(b_dst + p * 2048, ok ? G + off + gch : G, ok ? 16u : 0u);
                    }
                }
            }
            cp_async_arrive_noinc(smem_u32(&tail->full[s]));
        }
        // ---- epilogue: TMEM lane = P channel m, column = G channel c
        const int64_t Kt = (int64_t)d.Th * d.Tw * d.Cin;
        float* dst = ws + (int64_t)blockIdx.z * d.Cout * Kt;
        const int m = m0 + warp * 32 + lane;
        if (num_kb > 0) {
            mbar_wait(smem_u32(&tail->tmem_full), 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int cb = 0; cb < BNW; cb += 16) {
            uint32_t r[16];
            if (num_kb > 0) {
                tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, r);
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) r[e] = 0u;
            }
            if (m < d.Cout) {
                float* p = dst + (int64_t)m * Kt + (int64_t)tap * d.Cin + c0 + cb;
                if (c0 + cb + 16 <= d.Cin && ((Kt | d.Cin) & 3) == 0) {
                    float4* p4 = reinterpret_cast<float4*>(p);
                    p4[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                    p4[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
                    p4[2] = make_float4(__uint_as_float(r[8]), __uint_as_float(r[9]), __uint_as_float(r[10]), __uint_as_float(r[11]));
                    p4[3] = make_float4(__uint_as_float(r[12]), __uint_as_float(r[13]), __uint_as_float(r[14]), __uint_as_float(r[15]));
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (c0 + cb + e < d.Cin) p[e] = __uint_as_float(r[e]);
                }
            }
        }
    } else if (warp == 4) {
        // ===================== TMA producer (im2col maps of dY and of the gathered activation) =====================
        // warp-uniform; one elected lane issues the TMA requests; pixel coordinates advance incrementally
        if (WTMA) {
            const uint16_t ow = (uint16_t)(d.tap_sx > 0 ? txx : d.Tw - 1 - txx);
            const uint16_t oh = (uint16_t)(d.tap_sy > 0 ? tyy : d.Th - 1 - tyy);
            const int lw = d.tap_sx > 0 ? d.tap_ox : d.tap_ox - (d.Tw - 1);
            const int lh = d.tap_sy > 0 ? d.tap_oy : d.tap_oy - (d.Th - 1);
            const uint32_t bar0 = smem_u32(tail);
            const uint32_t a0 = smem_u32(smem_a), b0 = smem_u32(smem_b);
            int qx = (int)(q_begin % d.Qw);
            int qy = (int)((q_begin / d.Qw) % d.Qh);
            int n = (int)(q_begin / ((int64_t)d.Qw * d.Qh));
            uint32_t s = 0, ph = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(bar0 + (uint32_t)offsetof(SmemTail, empty) + 8u * s, ph ^ 1u);
                if (elect_one()) {
                    const uint32_t bar = bar0 + (uint32_t)offsetof(SmemTail, full) + 8u * s;
                    mbar_arrive_expect_tx(bar, PA_BYTES + GB_BYTES);
                    const uint32_t a_dst = a0 + s * PA_BYTES;
                    const int pw = qx * d.out_sx + d.out_ox, phh = qy * d.out_sy + d.out_oy;
#pragma unroll
                    for (int a = 0; a < PA_ATOMS; ++a)
                        tma_load_im2col_4d(a_dst + a * 8192, &tmap_p, m0 + a * BKE, pw, phh, n, 0, 0, bar);
                    const uint32_t b_dst = b0 + s * GB_BYTES;
                    const int gw = qx * d.in_sx + lw, gh = qy * d.in_sy + lh;
#pragma unroll
                    for (int a = 0; a < GB_ATOMS; ++a)
                        tma_load_im2col_4d(b_dst + a * 8192, &tmap_g, c0 + a * BKE, gw, gh, n, ow, oh, bar);
                }
                // advance 64 pixels
                qx += 64;
                if (qx >= d.Qw) {
                    const int cy = qx / d.Qw;
                    qx -= cy * d.Qw;
                    qy += cy;
                    if (qy >= d.Qh) {
                        const int cn = qy / d.Qh;
                        qy -= cn * d.Qh;
                        n += cn;
                    }
                }
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer (warp-uniform; one elected lane issues) =====================
        constexpr uint32_t idesc = make_idesc<EB>(BNW, 1, 1);
        // MN-major SWIZZLE_128B: LBO = stride between 64-element MN atoms (8192 B), SBO = stride between
        // 8-row K groups (1024 B); each UMMA (K = 16 pixels) advances 16 rows = 2048 B
        const uint64_t desc_hi = make_desc(0, 8192, 1024);
        const uint32_t bar0 = smem_u32(tail);
        const uint32_t a0 = smem_u32(smem_a) >> 4, b0 = smem_u32(smem_b) >> 4;
        uint32_t s = 0, ph = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(bar0 + (uint32_t)offsetof(SmemTail, full) + 8u * s, ph);
            tc_fence_after();
            const uint64_t adesc = desc_hi | (uint64_t)((a0 + s * (PA_BYTES >> 4)) & 0x3FFFu);
            const uint64_t bdesc = desc_hi | (uint64_t)((b0 + s * (GB_BYTES >> 4)) & 0x3FFFu);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 64 / KMMA; ++k)      // each MMA advances KMMA pixel rows of 128 bytes
                    umma<EB>(tmem_base, adesc + (uint64_t)(k * KMMA * 8), bdesc + (uint64_t)(k * KMMA * 8), idesc,
                             (uint32_t)(kb | k));
                umma_commit(bar0 + (uint32_t)offsetof(SmemTail, empty) + 8u * s);
                if (kb == num_kb - 1) umma_commit(bar0 + (uint32_t)offsetof(SmemTail, tmem_full));
            }
            if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, BNW < 32 ? 32 : BNW);
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static int g_use_persist = 2;      // 0 off, 1 one CTA per SM, 2 two CTAs per SM, 3 two CTAs per SM for short K loops
static int g_two_cta_max_kb = 4;
static int g_use_halo = 0;     // measured slower than the persistent im2col kernel on every step shape (profiles/r01e_bench_conv.md)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

static EncodeIm2colFn get_encode_im2col_fn() {
    static EncodeIm2colFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeIm2colFn>(p);
    }
    return fn;
}

// The activation operand can be fetched by TMA im2col when the gather is a plain strided window walk:
// unit tap steps, no fused nearest-upsampling, corners / offsets inside the 4-D im2col encoding limits.
static bool im2col_eligible(const b200_conv_desc* d) {
    if (d->up_shift != 0) return false;
    if ((d->tap_sy != 1 && d->tap_sy != -1) || (d->tap_sx != 1 && d->tap_sx != -1)) return false;
    if (d->in_sy < 1 || d->in_sy > 8 || d->in_sx < 1 || d->in_sx > 8) return false;
    if (d->Th > 255 || d->Tw > 255) return false;
    const int lw = d->tap_sx > 0 ? d->tap_ox : d->tap_ox - (d->Tw - 1);
    const int lh = d->tap_sy > 0 ? d->tap_oy : d->tap_oy - (d->Th - 1);
    const int uw = lw + (d->Qw - 1) * d->in_sx + 1 - d->Wi;
    const int uh = lh + (d->Qh - 1) * d->in_sy + 1 - d->Hi;
    if (lw < -128 || lw > 127 || lh < -128 || lh > 127 || uw < -128 || uw > 127 || uh < -128 || uh > 127) return false;
    // the bounding box [lower, dim + upper) must be non-empty
    if (d->Wi + uw - lw < 1 || d->Hi + uh - lh < 1) return false;
    return get_encode_im2col_fn() != nullptr;
}

static int encode_im2col(const b200_conv_desc* d, const void* in, CUtensorMap* tmap, int eb = 2) {
    EncodeIm2colFn enc = get_encode_im2col_fn();
    const int lw = d->tap_sx > 0 ? d->tap_ox : d->tap_ox - (d->Tw - 1);
    const int lh = d->tap_sy > 0 ? d->tap_oy : d->tap_oy - (d->Th - 1);
    const int uw = lw + (d->Qw - 1) * d->in_sx + 1 - d->Wi;
    const int uh = lh + (d->Qh - 1) * d->in_sy + 1 - d->Hi;
    cuuint64_t gdim[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->B};
    cuuint64_t gstr[3] = {(cuuint64_t)d->in_sw * eb, (cuuint64_t)d->in_sh * eb, (cuuint64_t)d->in_sn * eb};
    int lower[2] = {lw, lh}, upper[2] = {uw, uh};
    cuuint32_t estr[4] = {1, (cuuint32_t)d->in_sx, (cuuint32_t)d->in_sy, 1};
    // fp32 tensors feeding kind::tf32 MMAs are fetched as TFLOAT32: the TMA unit rounds to the 10-bit mantissa on the way in
    CUresult r = enc(tmap, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4,
                     const_cast<void*>(in), gdim, gstr, lower, upper, (cuuint32_t)(128 / eb), (cuuint32_t)BM, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "conv_gemm_tc: cuTensorMapEncodeIm2col failed (%d)", (int)r);
    // drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KiB (descriptor word 1, bit 21 must be
    // clear) — the same fix-up the CUTLASS im2col traits apply
    static int drv = -1;
    if (drv < 0 && cudaDriverGetVersion(&drv) != cudaSuccess) drv = 0;
    const uint64_t bytes = (uint64_t)d->B * (uint64_t)d->in_sn * (uint64_t)eb;
    if (drv <= 13010 && bytes < 131072) reinterpret_cast<uint64_t*>(tmap)[1] &= ~(1ull << 21);
    return 0;
}

// shifted-window plan: stride-1 window walk with unit tap steps onto a dense output at a resolution where the
// dummy-column / last-tile waste is small, slabs small enough to keep >= 2 in flight.  If the N tile's whole weight
// matrix fits next to the slabs it stays RESIDENT (BN = 64 is tried first for that), else it streams through a ring.
static bool halo_plan(const b200_conv_desc* d, int BN, bool allow_stream, HaloGeom* hg) {
    if (!g_use_halo) return false;
    if (d->up_shift != 0 || d->in_sy != 1 || d->in_sx != 1) return false;
    if ((d->tap_sy != 1 && d->tap_sy != -1) || (d->tap_sx != 1 && d->tap_sx != -1)) return false;
    if (d->out_sy != 1 || d->out_sx != 1 || d->out_oy != 0 || d->out_ox != 0 || d->Ho != d->Qh || d->Wo != d->Qw) return false;
    const int taps = d->Th * d->Tw;
    if (taps < 2) return false;
    const int Wp = d->Qw + d->Tw - 1;
    const int RB = (Wp - 1 + 127 + d->Tw - 1) / Wp + 1 + (d->Th - 1);
    if (Wp > 256 || RB > 256) return false;
    const int slab_tx = RB * Wp * 128;
    const int slab_bytes = (slab_tx + 1023) & ~1023;
    const int b_bytes = BN * BK * 2;
    const int cpb = d->Cin / BK;
    const int num_kb = taps * cpb;
    const int total = 226 * 1024 - 1024 - (int)sizeof(HaloTail);
    const int tiles_per_img = (d->Qh * Wp + BM - 1) / BM;
    // waste of the padded-row tiling against the dense M = Qh*Qw tiling of the im2col kernels
    const double waste = (double)tiles_per_img * BM / ((double)d->Qh * d->Qw);
    if (waste > 1.25) return false;
    int na, stages = 0, resident = 0;
    if (2 * slab_bytes + num_kb * b_bytes <= total) {
        resident = 1;
        na = (total - num_kb * b_bytes) / slab_bytes;
    } else {
        if (!allow_stream) return false;
        stages = 4;
        na = (total - stages * b_bytes) / slab_bytes;
        if (na < 2) return false;
        if (na > 4) na = 4;
        stages = (total - na * slab_bytes) / b_bytes;
        if (stages > 8) stages = 8;
        // streaming only pays when the activation re-read dominates the weight traffic
        const double K = (double)taps * d->Cin;
        const double bytes_halo = waste * (cpb * (double)slab_tx + BN * K * 2.0);
        const double bytes_im2col = K / BK * A_BYTES + BN * K * 2.0;
        if (bytes_halo > 0.75 * bytes_im2col) return false;
    }
    if (na > 4) na = 4;
    hg->Wp = Wp; hg->RB = RB;
    hg->lw = d->tap_sx > 0 ? d->tap_ox : d->tap_ox - (d->Tw - 1);
    hg->lh = d->tap_sy > 0 ? d->tap_oy : d->tap_oy - (d->Th - 1);
    hg->slab_bytes = slab_bytes; hg->slab_tx = slab_tx; hg->na = na; hg->b_stages = stages; hg->b_resident = resident;
    hg->tiles_per_img = tiles_per_img;
    return true;
}

template <int BN>
static int launch_halo(const b200_conv_desc* d, const HaloGeom& hg_in, const void* wmat, const void* in,
                       const float* bias, const float* scale, void* out, int out_bf16, cudaStream_t st) {
    EncodeTiledFn enc = get_encode_fn();
    HaloGeom hg = hg_in;
    const int pad_tile = b200_conv_tc_ntile(d->Cout);                       // rows of the packed weight matrix
    const int rows = (d->Cout + pad_tile - 1) / pad_tile * pad_tile;
    const int ntiles = (d->Cout + BN - 1) / BN;
    CUtensorMap tmap_w, tmap_in;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)d->ldw, (cuuint64_t)rows};
        cuuint64_t gstr[1] = {(cuuint64_t)d->ldw * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&tmap_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wmat), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        B200_REQUIRE(r == CUDA_SUCCESS, "conv_gemm_tc: cuTensorMapEncodeTiled(weights) failed (%d)", (int)r);
    }
    {
        cuuint64_t gdim[4] = {(cuuint64_t)d->Cin, (cuuint64_t)d->Wi, (cuuint64_t)d->Hi, (cuuint64_t)d->B};
        cuuint64_t gstr[3] = {(cuuint64_t)d->in_sw * 2, (cuuint64_t)d->in_sh * 2, (cuuint64_t)d->in_sn * 2};
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)hg.Wp, (cuuint32_t)hg.RB, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmap_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        B200_REQUIRE(r == CUDA_SUCCESS, "conv_gemm_tc: cuTensorMapEncodeTiled(4D slab) failed (%d)", (int)r);
    }
    const int num_kb = d->Th * d->Tw * (d->Cin / BK);
    const int smem_bytes = 1024 + hg.na * hg.slab_bytes + (hg.b_resident ? num_kb : hg.b_stages) * BN * BK * 2 +
                           (int)sizeof(HaloTail);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        B200_REQUIRE(e == cudaSuccess, "conv_gemm_tc: cudaFuncSetAttribute(halo): %s", cudaGetErrorString(e));
        configured = true;
    }
    const int64_t mtiles = (int64_t)d->B * hg.tiles_per_img;
    B200_REQUIRE(mtiles * ntiles < (1ll << 31), "conv_gemm_tc: too many tiles");
    hg.num_n_tiles = ntiles;
    hg.num_m_tiles = (int)mtiles;
    int grid;
    if (hg.b_resident) {
        int per_n = kNumSMs / ntiles;
        if (per_n < 1) return set_error("conv_gemm_tc: too many N tiles for the resident shifted-window kernel");
        if (per_n > mtiles) per_n = (int)mtiles;
        grid = per_n * ntiles;
    } else {
        const int64_t tiles = mtiles * ntiles;
        grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
    }
    conv_halo_tc_kernel<BN><<<grid, 192, smem_bytes, st>>>(tmap_w, tmap_in, *d, hg, bias, scale, out, out_bf16);
    B200_CHECK_LAUNCH();
    return 0;
}

template <int BN, int STAGES, int EB = 2>
static int launch_fwd(const b200_conv_desc* d, const __nv_bfloat16* in, const void* wmat, const float* bias,
                      const float* scale, void* out, int out_bf16, float* split_ws, int splits, int use_im2col,
                      cudaStream_t st) {
    EncodeTiledFn enc = get_encode_fn();
    B200_REQUIRE(enc != nullptr, "conv_gemm_tc: cuTensorMapEncodeTiled unavailable");
    int64_t M = (int64_t)d->B * d->Qh * d->Qw;
    int ntiles = (d->Cout + BN - 1) / BN;
    CUtensorMap tmap, tmap_a;
    cuuint64_t gdim[2] = {(cuuint64_t)d->ldw, (cuuint64_t)ntiles * BN};
    constexpr int BKE = 128 / EB;
    cuuint64_t gstr[1] = {(cuuint64_t)d->ldw * EB};
    cuuint32_t box[2] = {(cuuint32_t)BKE, (cuuint32_t)BN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmap, EB == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2,
                     const_cast<void*>(wmat), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "conv_gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    HaloGeom hg;
    if (EB == 2 && use_im2col && splits == 1 && d->col_stats == nullptr) {
        // resident weights with a 64-wide N tile first (fits for Cin = 64 3x3 layers of any Cout), then the native tile
        if (BN == 128 && d->Cout / 64 <= kNumSMs && halo_plan(d, 64, false, &hg))
            return launch_halo<64>(d, hg, wmat, in, bias, scale, out, out_bf16, st);
        if (halo_plan(d, BN, true, &hg)) return launch_halo<BN>(d, hg, wmat, in, bias, scale, out, out_bf16, st);
    }
    const bool atma = (use_im2col || EB != 2) && im2col_eligible(d);
    B200_REQUIRE(EB == 2 || atma, "conv_gemm_tf32: the gather is not expressible as a TMA im2col walk");
    if (atma) {
        if (encode_im2col(d, in, &tmap_a, EB) != 0) return -1;
    } else {
        tmap_a = tmap;
    }
    constexpr int smem_bytes = 1024 + STAGES * (A_BYTES + BN * 128) + (int)sizeof(SmemTail) + BM * 24;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaSuccess;
        if constexpr (EB == 2)
            e = cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, STAGES, false, 2>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv_gemm_tc_kernel<BN, STAGES, true, EB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     smem_bytes);
        B200_REQUIRE(e == cudaSuccess, "conv_gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    const int num_kb = d->Th * d->Tw * (d->Cin / BKE);
    int kbps = (num_kb + splits - 1) / splits;
    splits = (num_kb + kbps - 1) / kbps;          // no empty split
    static const int persist_max_kb = []() {
        const char* e = getenv("B200_TC_PERSIST_MAXKB");      // experiments only
        return e ? atoi(e) : (1 << 30);       // measured: the persistent kernel wins for every K-loop length
    }();
    if (atma && splits == 1 && g_use_persist && BN >= 64 && num_kb <= persist_max_kb) {
        // persistent path: operand ring running ahead across tiles, double-buffered TMEM accumulator.  Two co-resident
        // CTAs per SM with a shallower ring each (g_use_persist == 2, short K loops: two epilogue warp sets interleave)
        // or one CTA per SM with a deep ring.
        const int64_t mtiles = (M + BM - 1) / BM;
        const int64_t tiles = mtiles * ntiles;
        B200_REQUIRE(tiles < (1ll << 31), "conv_gemm_tc: too many tiles");
        const bool two = g_use_persist == 2 || (g_use_persist == 3 && num_kb <= g_two_cta_max_kb);
        if (two) {
            // BN = 128: two CTAs per SM (3 x 32 KB stages each); BN <= 64: three CTAs per SM (3 x 24 KB stages each) — the
            // short-K, epilogue-heavy launches gain from a third epilogue warp set
            constexpr int PSTAGES = 3;
            constexpr int PER_SM = BN == 128 ? 2 : 3;
            constexpr int psmem = 1024 + PSTAGES * (A_BYTES + BN * 128) + (int)sizeof(PersistTail);
            static bool pconfigured = false;
            if (!pconfigured) {
                cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_persist_kernel<BN, PSTAGES, EB>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, psmem);
                B200_REQUIRE(e == cudaSuccess, "conv_gemm_tc: cudaFuncSetAttribute(persist): %s", cudaGetErrorString(e));
                pconfigured = true;
            }
            const int grid_p = (int)(tiles < PER_SM * kNumSMs ? tiles : PER_SM * kNumSMs);
            conv_gemm_tc_persist_kernel<BN, PSTAGES, EB><<<grid_p, 192, psmem, st>>>(tmap, tmap_a, *d, bias, scale, out,
                                                                               out_bf16, ntiles, (int)tiles);
        } else {
            constexpr int PSTAGES = BN == 128 ? 6 : 8;
            constexpr int psmem = 1024 + PSTAGES * (A_BYTES + BN * 128) + (int)sizeof(PersistTail);
            static bool pconfigured = false;
            if (!pconfigured) {
                cudaError_t e = cudaFuncSetAttribute(conv_gemm_tc_persist_kernel<BN, PSTAGES, EB>,
                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, psmem);
                B200_REQUIRE(e == cudaSuccess, "conv_gemm_tc: cudaFuncSetAttribute(persist): %s", cudaGetErrorString(e));
                pconfigured = true;
            }
            const int grid_p = (int)(tiles < kNumSMs ? tiles : kNumSMs);
            conv_gemm_tc_persist_kernel<BN, PSTAGES, EB><<<grid_p, 192, psmem, st>>>(tmap, tmap_a, *d, bias, scale, out,
                                                                               out_bf16, ntiles, (int)tiles);
        }
        B200_CHECK_LAUNCH();
        return 0;
    }
    B200_REQUIRE(d->col_stats == nullptr, "conv_gemm_tc: col_stats needs the persistent kernel (check b200_conv_tc_stats_ok)");
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)ntiles, (unsigned)splits);
    if (atma) {
        conv_gemm_tc_kernel<BN, STAGES, true, EB><<<grid, 192, smem_bytes, st>>>(tmap, tmap_a, *d, in, bias, scale, out,
                                                                                out_bf16, split_ws, kbps);
    } else {
        if constexpr (EB == 2)
            conv_gemm_tc_kernel<BN, STAGES, false, 2><<<grid, 192, smem_bytes, st>>>(tmap, tmap_a, *d, in, bias, scale, out,
                                                                                    out_bf16, split_ws, kbps);
    }
    B200_CHECK_LAUNCH();
    if (splits > 1) {
        int64_t total = M * ((d->Cout + 3) / 4);
        conv_splitk_reduce_kernel<<<grid_for(total, 256), 256, 0, st>>>(*d, split_ws, splits, ntiles * BN, bias, scale,
                                                                        out, out_bf16);
        B200_CHECK_LAUNCH();
    }
    return 0;
}

// generic 4-D im2col map of a channel-last tensor walked with unit tap steps: `pixels` x 64 channels per request
static int encode_im2col_raw(const void* base, int C, int W, int H, int N, int64_t sw, int64_t sh, int64_t sn, int lw,
                             int lh, int uw, int uh, int step_x, int step_y, int pixels, CUtensorMap* tmap, int eb = 2) {
    EncodeIm2colFn enc = get_encode_im2col_fn();
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)sw * eb, (cuuint64_t)sh * eb, (cuuint64_t)sn * eb};
    int lower[2] = {lw, lh}, upper[2] = {uw, uh};
    cuuint32_t estr[4] = {1, (cuuint32_t)step_x, (cuuint32_t)step_y, 1};
    CUresult r = enc(tmap, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 4,
                     const_cast<void*>(base), gdim, gstr, lower, upper, (cuuint32_t)(128 / eb), (cuuint32_t)pixels, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "tcgen05 gather-GEMM: cuTensorMapEncodeIm2col failed (%d)", (int)r);
    static int drv = -1;
    if (drv < 0 && cudaDriverGetVersion(&drv) != cudaSuccess) drv = 0;
    if (drv <= 13010 && (uint64_t)N * (uint64_t)sn * (uint64_t)eb < 131072) reinterpret_cast<uint64_t*>(tmap)[1] &= ~(1ull << 21);
    return 0;
}

static bool wgrad_im2col_eligible(const b200_conv_desc* d) {
    if (!im2col_eligible(d)) return false;          // the gathered (G) side
    // the P side: pixel (qy*out_sy + out_oy, qx*out_sx + out_ox), no taps
    if (d->out_sy < 1 || d->out_sy > 8 || d->out_sx < 1 || d->out_sx > 8) return false;
    const int uw = d->out_ox + (d->Qw - 1) * d->out_sx + 1 - d->Wo;
    const int uh = d->out_oy + (d->Qh - 1) * d->out_sy + 1 - d->Ho;
    if (d->out_ox < -128 || d->out_ox > 127 || d->out_oy < -128 || d->out_oy > 127 || uw < -128 || uw > 127 ||
        uh < -128 || uh > 127)
        return false;
    if (d->Wo + uw - d->out_ox < 1 || d->Ho + uh - d->out_oy < 1) return false;
    return true;
}

template <int BNW, int STAGES, int EB = 2>
static int launch_wgrad(const b200_conv_desc* d, const __nv_bfloat16* P, const __nv_bfloat16* G, float* ws, int splits,
                        int use_im2col, cudaStream_t st) {
    int64_t Q = (int64_t)d->B * d->Qh * d->Qw;
    int64_t rps = (Q + splits - 1) / splits;
    rps = (rps + 63) / 64 * 64;
    if (rps < 64) rps = 64;
    constexpr int BKE = 128 / EB;
    constexpr int smem_bytes = 1024 + STAGES * ((128 / BKE) * 8192 + (BNW / BKE) * 8192) + (int)sizeof(SmemTail) + STAGES * 64 * 16;
    static_assert(smem_bytes <= 232448, "wgrad_gemm_tc: operand ring does not fit in shared memory");
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaSuccess;
        if constexpr (EB == 2)
            e = cudaFuncSetAttribute(wgrad_gemm_tc_kernel<BNW, STAGES, false, 2>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(wgrad_gemm_tc_kernel<BNW, STAGES, true, EB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     smem_bytes);
        B200_REQUIRE(e == cudaSuccess, "wgrad_gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    int ctiles = (d->Cin + BNW - 1) / BNW;
    dim3 grid((unsigned)((d->Cout + BM - 1) / BM), (unsigned)(d->Th * d->Tw * ctiles), (unsigned)splits);
    B200_REQUIRE(grid.y < 65536 && grid.z < 65536, "wgrad_gemm_tc: grid too large");
    CUtensorMap tmap_p, tmap_g;
    memset(&tmap_p, 0, sizeof(tmap_p));
    memset(&tmap_g, 0, sizeof(tmap_g));
    // the TMA path reads whole 64-channel atoms: both channel counts must cover the tiles (zero padding comes from
    // the tensor map's channel bound, so only a partial LAST atom is fine)
    const bool wtma = (use_im2col || EB != 2) && wgrad_im2col_eligible(d);
    B200_REQUIRE(EB == 2 || wtma, "wgrad_gemm_tf32: the gather is not expressible as a TMA im2col walk");
    if (wtma) {
        const int lw = d->tap_sx > 0 ? d->tap_ox : d->tap_ox - (d->Tw - 1);
        const int lh = d->tap_sy > 0 ? d->tap_oy : d->tap_oy - (d->Th - 1);
        const int uw = lw + (d->Qw - 1) * d->in_sx + 1 - d->Wi;
        const int uh = lh + (d->Qh - 1) * d->in_sy + 1 - d->Hi;
        if (encode_im2col_raw(G, d->Cin, d->Wi, d->Hi, d->B, d->in_sw, d->in_sh, d->in_sn, lw, lh, uw, uh, d->in_sx,
                              d->in_sy, 64, &tmap_g, EB) != 0)
            return -1;
        const int puw = d->out_ox + (d->Qw - 1) * d->out_sx + 1 - d->Wo;
        const int puh = d->out_oy + (d->Qh - 1) * d->out_sy + 1 - d->Ho;
        if (encode_im2col_raw(P, d->Cout, d->Wo, d->Ho, d->B, d->out_sw, d->out_sh, d->out_sn, d->out_ox, d->out_oy, puw,
                              puh, d->out_sx, d->out_sy, 64, &tmap_p, EB) != 0)
            return -1;
        wgrad_gemm_tc_kernel<BNW, STAGES, true, EB><<<grid, 192, smem_bytes, st>>>(tmap_p, tmap_g, *d, P, G, ws, rps);
    } else {
        if constexpr (EB == 2)
            wgrad_gemm_tc_kernel<BNW, STAGES, false, 2><<<grid, 192, smem_bytes, st>>>(tmap_p, tmap_g, *d, P, G, ws, rps);
    }
    B200_CHECK_LAUNCH();
    return 0;
}

__global__ void cast_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ y, int64_t n4) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
        float4 v = x[t];
        uint2 u;
        u.x = pack_bf16(v.x, v.y);
        u.y = pack_bf16(v.z, v.w);
        y[t] = u;
    }
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits of x survive; hi*hi' + hi*lo' + lo*hi' reproduces an
// fp32 product to ~2^-16 (the weight-gradient GEMM of the tf32 mode runs as three bf16 tensor-core GEMMs on these halves)
__global__ void split_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo, int64_t n4) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n4; t += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = x[t];
        const float f[4] = {v.x, v.y, v.z, v.w};
        float r[4];
        uint32_t h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat16 b = __float2bfloat16_rn(f[i]);
            h[i] = (uint32_t)__bfloat16_as_ushort(b);
            r[i] = f[i] - __bfloat162float(b);
        }
        hi[t] = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
        uint2 u;
        u.x = pack_bf16(r[0], r[1]);
        u.y = pack_bf16(r[2], r[3]);
        lo[t] = u;
    }
}

}  // namespace tc
}  // namespace b200

using namespace b200;

static int g_use_im2col = []() {
    const char* e = getenv("B200_NO_IM2COL");      // diagnostic switch: force the cp.async gather everywhere
    return (e && e[0] == '1') ? 0 : 1;
}();
extern "C" int b200_conv_tc_set_im2col(int enable) {
    int old = g_use_im2col;
    g_use_im2col = enable ? 1 : 0;
    return old;
}

extern "C" int b200_conv_tc_set_persistent(int enable) {
    int prev = tc::g_use_persist;
    tc::g_use_persist = enable < 0 ? 0 : (enable > 3 ? 3 : enable);
    return prev;
}

extern "C" int b200_conv_tc_set_halo(int enable) {
    int prev = tc::g_use_halo;
    tc::g_use_halo = enable ? 1 : 0;
    return prev;
}

extern "C" int b200_conv_tc_ntile(int Cout) { return Cout >= 128 ? 128 : (Cout >= 64 ? 64 : 16); }

static int conv_splits_impl(const b200_conv_desc* d, int bke);
extern "C" int b200_conv_tc_stats_ok(const b200_conv_desc* d, int elem_bytes) {
    const int bke = elem_bytes == 4 ? 32 : 64;
    if (!tc::g_use_persist || !(g_use_im2col || elem_bytes == 4) || b200_conv_tc_ntile(d->Cout) < 64) return 0;
    if (d->Cin % bke != 0 || !tc::im2col_eligible(d)) return 0;
    if (d->relu_mask != nullptr) return 0;
    return conv_splits_impl(d, bke) == 1 ? 1 : 0;
}

/* split-K plan: enough CTAs to fill the 148 SMs twice when the output tile grid alone cannot */
extern "C" int b200_conv_tc_splits(const b200_conv_desc* d) { return conv_splits_impl(d, 64); }

static int conv_splits_impl(const b200_conv_desc* d, int bke) {
    static int forced = []() {
        const char* e = getenv("B200_TC_SPLITS");      // experiments only: force the split-K factor
        return e ? atoi(e) : 0;
    }();
    if (forced > 0) {
        const int nkb = d->Th * d->Tw * (d->Cin / bke);
        return forced < nkb ? forced : (nkb > 0 ? nkb : 1);
    }
    int64_t M = (int64_t)d->B * d->Qh * d->Qw;
    int bn = b200_conv_tc_ntile(d->Cout);
    int64_t tiles = ((M + 127) / 128) * ((d->Cout + bn - 1) / bn);
    int num_kb = d->Th * d->Tw * (d->Cin / bke) * bke / 64;      // in 128-byte bf16-equivalent blocks: same work threshold
    // measured (tools/bench_conv.py, B200_TC_SPLITS): the fp32 partials' round trip plus the reduce launch cost more
    // than idle SMs unless fewer than half of them have a tile and the K loop is long
    if (tiles * 2 > kNumSMs || num_kb < 16) return 1;
    int64_t s = kNumSMs / tiles;
    if (s > num_kb / 8) s = num_kb / 8;
    if (s > 8) s = 8;
    return s < 1 ? 1 : (int)s;
}

extern "C" int b200_conv_tf32_splits(const b200_conv_desc* d) { return conv_splits_impl(d, 32); }

/* can the tf32 kernels (TMA im2col operands only) run this descriptor?  conv: kind 0; weight gradient: kind 1 */
extern "C" int b200_conv_tf32_ok(const b200_conv_desc* d, int wgrad) {
    if (d->Cin % 32 != 0 || d->in_sc != 1 || ((d->in_sn | d->in_sh | d->in_sw) & 3) != 0) return 0;
    if (wgrad) {
        if (d->Cout % 32 != 0 || d->out_sc != 1 || ((d->out_sn | d->out_sh | d->out_sw) & 3) != 0) return 0;
        return tc::wgrad_im2col_eligible(d) ? 1 : 0;
    }
    return tc::im2col_eligible(d) ? 1 : 0;
}

extern "C" int b200_conv_gemm_tf32(const b200_conv_desc* d, const float* in, const float* wmat, const float* bias,
                                   const float* scale, float* out, float* split_ws, int splits, b200_stream_t stream) {
    int64_t M = (int64_t)d->B * d->Qh * d->Qw;
    if (M == 0 || d->Cout == 0) return 0;
    B200_REQUIRE(d->Cin % 32 == 0 && d->in_sc == 1, "conv_gemm_tf32: needs Cin %% 32 == 0 and channel-last input");
    B200_REQUIRE(((d->in_sn | d->in_sh | d->in_sw) & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0,
                 "conv_gemm_tf32: input rows must be 16-byte aligned");
    B200_REQUIRE(d->ldw % 32 == 0 && d->ldw >= (int64_t)d->Th * d->Tw * d->Cin, "conv_gemm_tf32: bad ldw");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(wmat) & 15) == 0, "conv_gemm_tf32: wmat must be 16-byte aligned");
    B200_REQUIRE(splits >= 1 && (splits == 1 || split_ws != nullptr), "conv_gemm_tf32: split-K needs a workspace");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(d->relu_mask) & 15) == 0, "conv_gemm_tf32: relu_mask must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const __nv_bfloat16* inp = reinterpret_cast<const __nv_bfloat16*>(in);      // opaque to the TMA path
    int bn = b200_conv_tc_ntile(d->Cout);
    if (bn == 128) return tc::launch_fwd<128, 3, 4>(d, inp, wmat, bias, scale, out, 0, split_ws, splits, 1, st);
    if (bn == 64) return tc::launch_fwd<64, 4, 4>(d, inp, wmat, bias, scale, out, 0, split_ws, splits, 1, st);
    return tc::launch_fwd<16, 4, 4>(d, inp, wmat, bias, scale, out, 0, split_ws, splits, 1, st);
}

extern "C" int b200_wgrad_gemm_tf32(const b200_conv_desc* d, const float* P, const float* G, float* ws, int splits,
                                    b200_stream_t stream) {
    B200_REQUIRE(splits >= 1, "wgrad_gemm_tf32: bad splits");
    B200_REQUIRE(d->Cin % 32 == 0 && d->in_sc == 1 && d->Cout % 32 == 0 && d->out_sc == 1,
                 "wgrad_gemm_tf32: needs channel-last operands with channels %% 32 == 0");
    B200_REQUIRE(((d->in_sn | d->in_sh | d->in_sw | d->out_sn | d->out_sh | d->out_sw) & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(P) & 15) == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0,
                 "wgrad_gemm_tf32: operand rows must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const __nv_bfloat16* Pp = reinterpret_cast<const __nv_bfloat16*>(P);
    const __nv_bfloat16* Gp = reinterpret_cast<const __nv_bfloat16*>(G);
    if (d->Cin >= 128) return tc::launch_wgrad<128, 3, 4>(d, Pp, Gp, ws, splits, 1, st);
    return tc::launch_wgrad<64, 4, 4>(d, Pp, Gp, ws, splits, 1, st);
}

extern "C" int b200_cast_bf16(const float* x, void* y, int64_t n, b200_stream_t stream) {
    if (n == 0) return 0;
    B200_REQUIRE(n % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0,
                 "cast_bf16: needs n %% 4 == 0 and aligned pointers");
    tc::cast_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(y), n / 4);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_split_bf16(const float* x, void* hi, void* lo, int64_t n, b200_stream_t stream) {
    if (n == 0) return 0;
    B200_REQUIRE(n % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 &&
                     (reinterpret_cast<uintptr_t>(lo) & 7) == 0, "split_bf16: needs n %% 4 == 0 and aligned pointers");
    tc::split_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo), n / 4);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_conv_gemm_tc(const b200_conv_desc* d, const void* in_bf16, const void* wmat_bf16, const float* bias,
                                 const float* scale, void* out, int out_bf16, float* split_ws, int splits,
                                 b200_stream_t stream) {
    int64_t M = (int64_t)d->B * d->Qh * d->Qw;
    if (M == 0 || d->Cout == 0) return 0;
    B200_REQUIRE(d->Cin % 64 == 0 && d->in_sc == 1, "conv_gemm_tc: needs Cin %% 64 == 0 and channel-last input");
    B200_REQUIRE(((d->in_sn | d->in_sh | d->in_sw) & 7) == 0 && (reinterpret_cast<uintptr_t>(in_bf16) & 15) == 0,
                 "conv_gemm_tc: input rows must be 16-byte aligned");
    B200_REQUIRE(d->ldw % 64 == 0 && d->ldw >= (int64_t)d->Th * d->Tw * d->Cin, "conv_gemm_tc: bad ldw");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(wmat_bf16) & 15) == 0, "conv_gemm_tc: wmat must be 16-byte aligned");
    B200_REQUIRE(splits >= 1 && (splits == 1 || split_ws != nullptr), "conv_gemm_tc: split-K needs a workspace");
    B200_REQUIRE((reinterpret_cast<uintptr_t>(d->relu_mask) & 15) == 0, "conv_gemm_tc: relu_mask must be 16-byte aligned");
    B200_REQUIRE(splits == 1 || (reinterpret_cast<uintptr_t>(split_ws) & 15) == 0, "conv_gemm_tc: unaligned workspace");
    cudaStream_t st = as_stream(stream);
    const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(in_bf16);
    int bn = b200_conv_tc_ntile(d->Cout);
    const int use_im2col = g_use_im2col;
    if (bn == 128) return tc::launch_fwd<128, 3>(d, in, wmat_bf16, bias, scale, out, out_bf16, split_ws, splits, use_im2col, st);
    if (bn == 64) return tc::launch_fwd<64, 4>(d, in, wmat_bf16, bias, scale, out, out_bf16, split_ws, splits, use_im2col, st);
    return tc::launch_fwd<16, 4>(d, in, wmat_bf16, bias, scale, out, out_bf16, split_ws, splits, use_im2col, st);
}

extern "C" int b200_wgrad_gemm_tc(const b200_conv_desc* d, const void* P_bf16, const void* G_bf16, float* ws, int splits,
                                  b200_stream_t stream) {
    B200_REQUIRE(splits >= 1, "wgrad_gemm_tc: bad splits");
    B200_REQUIRE(d->Cin % 64 == 0 && d->in_sc == 1 && d->Cout % 64 == 0 && d->out_sc == 1,
                 "wgrad_gemm_tc: needs channel-last operands with channels %% 64 == 0");
    B200_REQUIRE(((d->in_sn | d->in_sh | d->in_sw | d->out_sn | d->out_sh | d->out_sw) & 7) == 0 &&
                     (reinterpret_cast<uintptr_t>(P_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(G_bf16) & 15) == 0,
                 "wgrad_gemm_tc: operand rows must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    const __nv_bfloat16* P = reinterpret_cast<const __nv_bfloat16*>(P_bf16);
    const __nv_bfloat16* G = reinterpret_cast<const __nv_bfloat16*>(G_bf16);
    if (d->Cin >= 128) return tc::launch_wgrad<128, 3>(d, P, G, ws, splits, g_use_im2col, st);
    return tc::launch_wgrad<64, 4>(d, P, G, ws, splits, g_use_im2col, st);
}
