// loss.cu — the step arithmetic of the training loop (train64.py:195-252, 284-364) as a handful of fused kernels: every
// loss term is ONE launch that produces the term's partial sums AND the gradient with respect to its logits / tensors
// (the terms are closed-form: BCE-with-logits against constant targets, cross entropy, BCE with pos_weight on annotated
// rows, masked L1, KL), and one final launch combines the partial sums in fixed order into the term values and the total.
// Replaces ~350 elementwise / reduction launches of the PyTorch formulation per iteration.  Deterministic: per-block partial
// sums in double, summed in index order by b200_loss_total.
//
// Partial-sum buffer layout: `partials` is [n_terms][B200_LOSS_MAX_BLOCKS] doubles; a term kernel writes one value per block
// into row `slot` and the number of blocks it used into counts[slot].
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < 8; ++w) t += sh[w];
    __syncthreads();
    return t;           // valid in thread 0
}

// F.binary_cross_entropy_with_logits(x, t) = (1 - t) * x + softplus(-x), softplus(-x) = max(-x, 0) + log1p(exp(-|x|))
__device__ __forceinline__ float softplus_neg(float x) { return fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x))); }
__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + expf(-x)); }

// term = scale * sum_g weight[g] * mean_{i in group g} BCE(x_i, target[g]);  x: groups * n logits
// groups >= split_group are summed into the NEXT slot (the fake / real halves of a discriminator's adversarial loss are one
// launch but two reported terms, train64.py:195-212)
__global__ void __launch_bounds__(256) loss_bce_groups_kernel(const float* __restrict__ x, int n, int groups, int split_group,
                                                              const float* __restrict__ target,
                                                              const float* __restrict__ weight, float scale,
                                                              float* __restrict__ grad, double* __restrict__ partial,
                                                              int* __restrict__ count) {
    __shared__ double sh[8];
    const int total = n * groups;
    double acc = 0.0, acc2 = 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
        const int g = i / n;
        const float t = target[g], w = weight[g] * scale / (float)n, v = x[i];
        const double l = (double)(w * ((1.f - t) * v + softplus_neg(v)));
        if (g < split_group) acc += l; else acc2 += l;
        grad[i] = w * (sigmoidf(v) - t);
    }
    const double s = block_sum_256(acc, sh);
    const double s2 = block_sum_256(acc2, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = s;
        if (blockIdx.x == 0) *count = gridDim.x;
        if (split_group < groups) {
            partial[B200_LOSS_MAX_BLOCKS + blockIdx.x] = s2;
            if (blockIdx.x == 0) count[1] = gridDim.x;
        }
    }
}

// term = scale * sum_g weight[g] * mean_{r in group g} CE(x[g*n + r, :], label[r]); one warp per row; rows of groups with
// weight 0 get a zero gradient (the D-step classifies only the real crops, train64.py:236-238)
__global__ void __launch_bounds__(256) loss_ce_groups_kernel(const float* __restrict__ x, const int64_t* __restrict__ label,
                                                             int n, int groups, int C, const float* __restrict__ weight,
                                                             float scale, float* __restrict__ grad,
                                                             double* __restrict__ partial, int* __restrict__ count) {
    __shared__ double sh[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows = n * groups;
    double acc = 0.0;
    for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
        const int g = r / n;
        const float w = weight[g] * scale / (float)n;
        const float* xr = x + (int64_t)r * C;
        float* gr = grad + (int64_t)r * C;
        if (w == 0.f) {
            for (int c = lane; c < C; c += 32) gr[c] = 0.f;
            continue;
        }
        float m = -INFINITY;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, xr[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += expf(xr[c] - m);
        se = warp_sum(se);
        const float lse = m + logf(se);
        const int lab = (int)label[r - g * n];
        for (int c = lane; c < C; c += 32) gr[c] = w * (expf(xr[c] - lse) - (c == lab ? 1.f : 0.f));
        if (lane == 0) acc += (double)(w * (lse - xr[lab]));
    }
    const double s = block_sum_256(acc, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = s;
        if (blockIdx.x == 0) *count = gridDim.x;
    }
}

// term = scale * sum_g weight[g] * mean over {selected rows r, all A columns} of BCE(x[g*n + r, a], t[r, a]; pos_weight[a]);
// F.binary_cross_entropy_with_logits with pos_weight: lw = 1 + (pw - 1) * t; (1 - t) * x + lw * softplus(-x).
// sel[r] in {0, 1} marks the annotated objects (train64.py:241-246, 323-349: index_select by att_idx); n_sel = their number.
__global__ void __launch_bounds__(256) loss_bce_pw_rows_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                               const float* __restrict__ sel, int n, int groups, int A,
                                                               int n_sel, const float* __restrict__ pos_weight,
                                                               const float* __restrict__ weight, float scale,
                                                               float* __restrict__ grad, double* __restrict__ partial,
                                                               int* __restrict__ count) {
    __shared__ double sh[8];
    const int64_t total = (int64_t)n * groups * A;
    double acc = 0.0;
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int a = (int)(i % A);
        const int row = (int)(i / A);
        const int g = row / n, r = row - g * n;
        float gv = 0.f;
        if (sel[r] != 0.f && n_sel > 0) {
            const float w = weight[g] * scale / ((float)n_sel * (float)A);
            const float tv = t[(int64_t)r * A + a], v = x[i];
            const float lw = 1.f + (pos_weight[a] - 1.f) * tv;
            acc += (double)(w * ((1.f - tv) * v + lw * softplus_neg(v)));
            gv = w * ((1.f - tv) - lw * (1.f - sigmoidf(v)));
        }
        grad[i] = gv;
    }
    const double s = block_sum_256(acc, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = s;
        if (blockIdx.x == 0) *count = gridDim.x;
    }
}

// term = scale * sum_n mask[n] * mean_L |a[n, :] - b[n, :]| / denom   (b_stride_n = 0 broadcasts one b row; mask may be NULL)
// grid = (chunks, N)
__global__ void __launch_bounds__(256) loss_l1_rows_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           int64_t L, int64_t b_stride_n, const float* __restrict__ mask,
                                                           float denom, float scale, float* __restrict__ grad,
                                                           double* __restrict__ partial, int* __restrict__ count) {
    __shared__ double sh[8];
    const int n = blockIdx.y;
    const float w = (mask ? mask[n] : 1.f) * scale / ((float)L * denom);
    const float* an = a + (int64_t)n * L;
    const float* bn = b + (int64_t)n * b_stride_n;
    float* gn = grad + (int64_t)n * L;
    double acc = 0.0;
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < L; i += (int64_t)gridDim.x * 256) {
        const float dlt = an[i] - bn[i];
        acc += (double)(w * fabsf(dlt));
        gn[i] = dlt > 0.f ? w : (dlt < 0.f ? -w : 0.f);
    }
    const double s = block_sum_256(acc, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.y * gridDim.x + blockIdx.x] = s;
        if (blockIdx.x == 0 && blockIdx.y == 0) *count = gridDim.x * gridDim.y;
    }
}

// term = scale * (-0.5) * sum (1 + logvar - mu^2 - exp(logvar))   (train64.py:293)
__global__ void __launch_bounds__(256) loss_kl_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, int64_t n,
                                                      float scale, float* __restrict__ dmu, float* __restrict__ dlogvar,
                                                      double* __restrict__ partial, int* __restrict__ count) {
    __shared__ double sh[8];
    double acc = 0.0;
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const float m = mu[i], lv = logvar[i], e = expf(lv);
        acc += (double)(-0.5f * scale * (1.f + lv - m * m - e));
        dmu[i] = scale * m;
        dlogvar[i] = -0.5f * scale * (1.f - e);
    }
    const double s = block_sum_256(acc, sh);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = s;
        if (blockIdx.x == 0) *count = gridDim.x;
    }
}

// terms[s] = sum of slot s's partials (index order); terms[n_terms] = sum of all terms
__global__ void loss_total_kernel(const double* __restrict__ partials, const int* __restrict__ counts, int n_terms,
                                  float* __restrict__ terms) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double total = 0.0;
    for (int s = 0; s < n_terms; ++s) {
        double t = 0.0;
        const int c = counts[s];
        for (int k = 0; k < c; ++k) t += partials[(int64_t)s * B200_LOSS_MAX_BLOCKS + k];
        terms[s] = (float)t;
        total += t;
    }
    terms[n_terms] = (float)total;
}

static inline int loss_blocks(int64_t work, int per_block) {
    int64_t g = (work + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > B200_LOSS_MAX_BLOCKS) g = B200_LOSS_MAX_BLOCKS;
    return (int)g;
}

}  // namespace b200

using namespace b200;

#define SLOT_ARGS double* partials, int* counts, int slot
#define SLOT_PTRS partials + (int64_t)slot * B200_LOSS_MAX_BLOCKS, counts + slot

extern "C" int b200_loss_bce_groups(const float* x, int n, int groups, int split_group, const float* target,
                                    const float* weight, float scale, float* grad, SLOT_ARGS, b200_stream_t stream) {
    B200_REQUIRE(n > 0 && groups > 0, "loss_bce_groups: empty input");
    loss_bce_groups_kernel<<<loss_blocks((int64_t)n * groups, 1024), 256, 0, as_stream(stream)>>>(
        x, n, groups, split_group, target, weight, scale, grad, SLOT_PTRS);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_loss_ce_groups(const float* x, const int64_t* label, int n, int groups, int C, const float* weight,
                                   float scale, float* grad, SLOT_ARGS, b200_stream_t stream) {
    B200_REQUIRE(n > 0 && groups > 0 && C > 0, "loss_ce_groups: empty input");
    loss_ce_groups_kernel<<<loss_blocks((int64_t)n * groups, 8), 256, 0, as_stream(stream)>>>(x, label, n, groups, C, weight,
                                                                                               scale, grad, SLOT_PTRS);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_loss_bce_pw_rows(const float* x, const float* t, const float* sel, int n, int groups, int A, int n_sel,
                                     const float* pos_weight, const float* weight, float scale, float* grad, SLOT_ARGS,
                                     b200_stream_t stream) {
    B200_REQUIRE(n > 0 && groups > 0 && A > 0, "loss_bce_pw_rows: empty input");
    loss_bce_pw_rows_kernel<<<loss_blocks((int64_t)n * groups * A, 1024), 256, 0, as_stream(stream)>>>(
        x, t, sel, n, groups, A, n_sel, pos_weight, weight, scale, grad, SLOT_PTRS);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_loss_l1_rows(const float* a, const float* b, int N, int64_t L, int64_t b_stride_n, const float* mask,
                                 float denom, float scale, float* grad, SLOT_ARGS, b200_stream_t stream) {
    B200_REQUIRE(N > 0 && L > 0 && N <= B200_LOSS_MAX_BLOCKS, "loss_l1_rows: bad sizes");
    int chunks = loss_blocks(L, 2048);
    if (chunks > B200_LOSS_MAX_BLOCKS / N) chunks = B200_LOSS_MAX_BLOCKS / N;
    if (chunks < 1) chunks = 1;
    loss_l1_rows_kernel<<<dim3((unsigned)chunks, (unsigned)N), 256, 0, as_stream(stream)>>>(a, b, L, b_stride_n, mask, denom,
                                                                                           scale, grad, SLOT_PTRS);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_loss_kl(const float* mu, const float* logvar, int64_t n, float scale, float* dmu, float* dlogvar, SLOT_ARGS,
                            b200_stream_t stream) {
    B200_REQUIRE(n > 0, "loss_kl: empty input");
    loss_kl_kernel<<<loss_blocks(n, 1024), 256, 0, as_stream(stream)>>>(mu, logvar, n, scale, dmu, dlogvar, SLOT_PTRS);
    B200_CHECK_LAUNCH();
    return 0;
}

extern "C" int b200_loss_total(const double* partials, const int* counts, int n_terms, float* terms, b200_stream_t stream) {
    B200_REQUIRE(n_terms > 0, "loss_total: no terms");
    loss_total_kernel<<<1, 32, 0, as_stream(stream)>>>(partials, counts, n_terms, terms);
    B200_CHECK_LAUNCH();
    return 0;
}
