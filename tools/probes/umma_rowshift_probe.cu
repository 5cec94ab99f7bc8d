// Probe: does tcgen05.mma accept an A descriptor whose start address is shifted by r x 128 bytes (r rows) inside a
// SWIZZLE_128B K-major tile that TMA wrote at a 1024-byte aligned base?  (needed by the shifted-window 3x3 convolution)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rowshift_probe umma_rowshift_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0; int spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (++spins > (1 << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"((uint64_t)tmap), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7u) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// A source: 160 rows x 64 bf16 (K-major), B: 64 rows (N) x 64 bf16.  out[variant][r][128][64] fp32
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tma, const __grid_constant__ CUtensorMap tmb, float* out, int variant, int shift) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                 // 160 rows * 128 B = 20480
    uint8_t* sb = smem + 20480;         // 64 * 128 = 8192
    uint64_t* bars = (uint64_t*)(smem + 20480 + 8192);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        mbar_expect(smem_u32(&bars[0]), 20480 + 8192);
        tma_load_2d(smem_u32(sa), &tma, 0, 0, smem_u32(&bars[0]));       // box 64 x 160? (two loads of 80 rows keep box <= 256)
        tma_load_2d(smem_u32(sb), &tmb, 0, 0, smem_u32(&bars[0]));
        mbar_wait(smem_u32(&bars[0]), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = smem_u32(sa) + (uint32_t)shift * 128u;
        const uint32_t boff = variant == 0 ? 0u : ((a0 >> 7) & 7u);
        for (int k = 0; k < 4; ++k) {
            uint64_t ad = make_desc(a0 + k * 32, 16, 1024, boff);
            uint64_t bd = make_desc(smem_u32(sb) + k * 32, 16, 1024, 0);
            uint32_t acc = k != 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(make_idesc(64)), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[1])) : "memory");
    }
    mbar_wait(smem_u32(&bars[1]), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int cb = 0; cb < 64; cb += 16) {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + cb) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 16; ++e) out[(size_t)tid * 64 + cb + e] = __uint_as_float(r[e]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int RA = 160, K = 64, N = 64;
    std::vector<__nv_bfloat16> ha(RA * K), hb(N * K);
    std::vector<float> fa(RA * K), fb(N * K);
    srand(1);
    for (int i = 0; i < RA * K; ++i) { float v = (float)((rand() % 17) - 8) / 8.f; ha[i] = __float2bfloat16(v); fa[i] = __bfloat162float(ha[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (float)((rand() % 13) - 6) / 4.f; hb[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hb[i]); }
    __nv_bfloat16 *da, *db; float* dout;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    CUtensorMap ta, tb;
    cuuint64_t gda[2] = {K, RA}, gsa[1] = {K * 2}; cuuint32_t boxa[2] = {64, RA}, es[2] = {1, 1};
    CUresult r1 = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, gda, gsa, boxa, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t gdb[2] = {K, N}; cuuint32_t boxb[2] = {64, N};
    CUresult r2 = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, gdb, gsa, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d %d\n", (int)r1, (int)r2);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    std::vector<float> ho(128 * 64);
    for (int variant = 0; variant < 2; ++variant)
        for (int shift = 0; shift <= 17; ++shift) {
            cudaMemset(dout, 0, 128 * 64 * 4);
            probe<<<1, 128, 40000>>>(ta, tb, dout, variant, shift);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant %d shift %d: CUDA error %s\n", variant, shift, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < K; ++k) ref += (double)fa[(m + shift) * K + k] * fb[n * K + k];
                    maxerr = fmax(maxerr, fabs(ref - ho[m * 64 + n]));
                }
            printf("variant %d (base_offset %s) shift %2d rows: max abs err %.4g %s\n", variant, variant ? "=(addr>>7)&7" : "=0", shift, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
        }
    return 0;
}
