/* b200gan.h — C ABI of libb200gan.so: the sm_100a kernels behind the G+D training step of the
 * attribute-guided layout-to-image GAN (reference: ubc-vision/attribute-guided-image-generation-from-layout).
 *
 * Conventions
 *  - every pointer is DEVICE memory owned by the caller (PyTorch), including workspaces;
 *    the library never allocates, frees, synchronises or throws;
 *  - every entry point is asynchronous on `stream` and returns 0, or a negative code with a
 *    thread-local message retrievable through b200_last_error();
 *  - activations are fp32; "rows x C" tensors are channel-last (N*H*W rows, C contiguous); conv
 *    tensors carry explicit element strides so NCHW (3-channel images / crops, the layout at the
 *    reference boundary) and NHWC (internal) are both addressed without a transpose;
 *  - reductions are deterministic (fixed order, no floating point atomics).
 *
 * Each entry cites the reference interface it replaces (file:line under the reference root).
 */
#ifndef B200GAN_H
#define B200GAN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200_stream_t; /* cudaStream_t */

/* Storage type of a channel-last ACTIVATION tensor (and of its gradient): fp32, or bf16 in the bf16 training mode.
 * Wherever an entry point takes `void*` activations it also takes their type code `dt`; arithmetic, statistics,
 * reductions, parameters and parameter gradients are always fp32 (or wider).  NCHW boundary tensors (images, crops)
 * are always fp32. */
enum { B200_F32 = 0, B200_BF16 = 1 };

const char* b200_last_error(void);
int b200_version(void);
/* number of kernel launches issued through this library by the calling process (for bench.py's gpu_launches) */
int64_t b200_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Box -> layout integer work (bit-exact contract): the rasterised box masks the LayoutEncoder broadcasts the object
 * embedding into (generator_obj_att.py:489-490) and the shifted boxes of the "shift" pass.
 * b200_rasterize_boxes: masks (O,1,H,W) fp32 = 1 inside [round(y0*H), round(y1*H)) x [round(x0*W), round(x1*W)), else 0,
 *   with Python's round (half to even, on double) and Python slice-bound semantics (data/vg_custom_mask.py:120,136,157);
 *   boxes (O,4) fp32 [x0,y0,x1,y1].
 * b200_shift_boxes: data/vg_custom_mask.py:139-158 — boxes narrower than 0.5 move 0.8x of the distance to the farther
 *   horizontal border (double arithmetic, fp32 result). */
int b200_rasterize_boxes(const float* boxes, int O, int H, int W, float* masks, b200_stream_t stream);
int b200_shift_boxes(const float* boxes, int O, float* out, b200_stream_t stream);

/* masks_to_layout — the scatter-sum object layout the reference calls at utils/draw_box.py:482-483
 * (`masks_to_layout(vecs, boxes, masks, obj_to_img, H=..., N=...)`; the definition is sg2im's layout.py, not shipped in the
 * reference tree — see oracle/layout_oracle.py): out (N,D,H,W) fp32,
 *   out[n,d,y,x] = sum over the objects o of image n (ascending) of vecs[o,d] * bilinear(masks[o] (M,M), grid_o(y,x)),
 * grid_o from the box [x0,y0,x1,y1] as X = (linspace(0,1,W)[x] - x0)/(x1 - x0) (zeros padding, align_corners=False).
 * linx (W) / liny (H): torch.linspace(0,1,.) built on the CPU by the caller (the reference's tables, bit for bit).
 * img_obj_start (N+1) / obj_order (O): objects grouped by image, ascending inside an image; obj_to_img (O) int32.
 * vecs (O,D) with D % 4 == 0.  Deterministic gather (no atomics); floors / in-bounds predicates bit-exact
 * (b200_masks_to_layout_taps exposes them: ix0 (O,W), iy0 (O,H) int32 and the fractional weights).
 * bwd: dvecs (O,D) and/or dmasks (O,M,M) (either may be NULL); ws: O*H*W floats (needed for dmasks). */
int b200_masks_to_layout_taps(const float* boxes, const float* linx, const float* liny, int32_t* ix0, int32_t* iy0,
                              float* fx, float* fy, int M, int H, int W, int O, b200_stream_t stream);
int b200_masks_to_layout_fwd(const float* vecs, const float* boxes, const float* masks, const int32_t* img_obj_start,
                             const int32_t* obj_order, const float* linx, const float* liny, float* out, int N, int O,
                             int D, int M, int H, int W, b200_stream_t stream);
int b200_masks_to_layout_bwd(const float* dout, const float* vecs, const float* boxes, const float* masks,
                             const int32_t* obj_to_img, const float* linx, const float* liny, float* dvecs, float* dmasks,
                             float* ws, int N, int O, int D, int M, int H, int W, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Box crops — replaces models/bilinear.py:26-41,67-104,107-136 (crop_bbox_batch -> F.grid_sample,
 * bilinear, zeros padding, align_corners=False) and its autograd backward.
 * feats (N,C,H,W) NCHW; boxes (B,4) [x0,y0,x1,y1] in [0,1]; box_to_img (B) int32;
 * wx (2*WW) = [linspace(1,0,WW) | linspace(0,1,WW)] and wy (2*HH) built on the HOST in fp32 exactly
 * as bilinear.py:272-275 does; crops (B,C,HH,WW).
 * b200_crop_taps exports the integer part (floor indices) for the bit-exact index contract.
 * b200_crop_bwd is a deterministic two-pass gather restricted to each box's pixel footprint: ws needs B*C*H*WW floats
 * (rounded up to a multiple of 4) + 4*B more (the footprints), 16-byte aligned; img_box_start (N+1) int32
 * are offsets into box_order (B) int32 = box ids grouped by image (ascending box id inside an image).
 */
int b200_crop_fwd(const float* feats, const float* boxes, const int32_t* box_to_img, const float* wx, const float* wy,
                  float* crops, int N, int C, int H, int W, int B, int HH, int WW, b200_stream_t stream);
int b200_crop_taps(const float* boxes, const float* wx, const float* wy, int32_t* ix0, int32_t* iy0, float* fx,
                   float* fy, int H, int W, int B, int HH, int WW, b200_stream_t stream);
int b200_crop_bwd(const float* dcrops, const float* boxes, const int32_t* img_box_start, const int32_t* box_order,
                  const float* wx, const float* wy, float* dfeats, float* ws, int N, int C, int H, int W, int B,
                  int HH, int WW, b200_stream_t stream);
/* The crop kernels stage each box's source footprint (forward) / the image-plane gradient tile and the crop gradient
 * (backward) in shared memory — coalesced global traffic, coordinates computed once per box — and are bit-identical to the
 * element-wise kernels they replace; b200_crop_set_staged(0) selects the element-wise kernels (parity tests compare the two).
 * Returns the previous setting. */
int b200_crop_set_staged(int enable);

/* ------------------------------------------------------------------------------------------------
 * Convolutions as gather-GEMMs — replace every nn.Conv2d / nn.ConvTranspose2d / nn.Linear forward,
 * dgrad and wgrad on the path (generator_obj_att.py:93,374-393,432-435,474-483,528-544,582-586;
 * generator_obj_att128.py:549-557; discriminator.py:36-44,70-79,128,168,218,252-253;
 * models/spade/networks/normalization.py:88-92).
 *
 * One descriptor covers forward convs (any stride), stride-1 dgrad, the four output phases of a
 * stride-2 dgrad / ConvTranspose forward, nearest-upsampled inputs and NCHW/NHWC addressing:
 *   out[n, qy*out_sy+out_oy, qx*out_sx+out_ox, co] =
 *       epilogue( sum_{ty<Th, tx<Tw, c<Cin} in[n, (qy*in_sy + ty*tap_sy + tap_oy) >> up, (qx*in_sx + ...) >> up, c]
 *                                          * wmat[co, (ty*Tw+tx)*Cin + c] )
 * taps outside [0,Hi)x[0,Wi) (logical, i.e. after upsampling) read zero; outputs outside [0,Ho)x[0,Wo) are dropped.
 * epilogue(v) = relu?( v * (scale ? scale[scale_rows ? m / scale_rows : 0] : 1) + bias[co] ), m = output row index.
 */
typedef struct {
    int B, Qh, Qw;                 /* output grid: rows M = B*Qh*Qw */
    int Cin, Cout;
    int Th, Tw;                    /* tap grid */
    int in_sy, in_sx;              /* input step per output grid step */
    int tap_sy, tap_sx;            /* input step per tap (may be negative) */
    int tap_oy, tap_ox;            /* input offset of tap 0 */
    int Hi, Wi;                    /* logical input extent (after nearest upsampling) */
    int up_shift;                  /* physical input coordinate = logical >> up_shift */
    int64_t in_sn, in_sh, in_sw, in_sc;     /* input element strides (physical) */
    int out_sy, out_sx, out_oy, out_ox;    /* output coordinate = q*out_s + out_o */
    int Ho, Wo;
    int64_t out_sn, out_sh, out_sw, out_sc; /* output element strides */
    int64_t ldw;                   /* row stride (elements) of wmat */
    int relu;
    int scale_rows;                /* 0: one scalar *scale; > 0: scale[m / scale_rows] (per-group 1/sigma of batched calls) */
    const void* relu_mask;         /* tcgen05 path only, may be NULL: tensor with out's type and strides; outputs are zeroed
                                      where it is not > 0 — the ReLU backward of a layer whose INPUT is a ReLU output, fused
                                      into that layer's data-gradient GEMM (out = dX, relu_mask = X) */
    float* col_stats;              /* persistent tcgen05 kernel only (b200_conv_tc_stats_ok), may be NULL: per 32-row slab of the
                                      output (row r = m / 32, the warp that owns those rows) and per output channel c the pair
                                      (sum, sum of squares) of the STORED values: col_stats[(r * col_stats_ld + c) * 2 + {0,1}],
                                      ceil(M/128)*4 slabs — batch-norm statistics accumulated in the convolution's epilogue
                                      (b200_bn_stats_slabs finishes them), so the normalisation reads its input once */
    int col_stats_ld;              /* channels per slab row (>= the N-tile-padded Cout) */
} b200_conv_desc;

/* fp32 CUDA-core path (bit-tight parity mode; also the 3-channel layers and Linear heads). wmat fp32 [Cout][ldw];
 * in / out stored as in_dt / out_dt, strides in elements of that type. */
int b200_conv_gemm_f32(const b200_conv_desc* d, const void* in, int in_dt, const float* wmat, const float* bias,
                       const float* scale, void* out, int out_dt, b200_stream_t stream);
/* tcgen05 path: bf16 operands, fp32 accumulate in TMEM.  `in_bf16` is the channel-last activation stored as bf16
 * (b200_cast_bf16 produces it from an fp32 tensor); the descriptor's input strides are in bf16 elements and must be
 * multiples of 8 (16-byte rows).  wmat bf16 [Cout_pad][ldw] with Cout_pad a multiple of the N tile
 * (b200_conv_tc_ntile), ldw a multiple of 64, zero padded; requires Cin % 64 == 0 and in_sc == 1.
 * out is fp32 (out_bf16 = 0) or bf16 (out_bf16 = 1) with the descriptor's output strides in elements of that type.
 * splits > 1 divides the K range (taps x channel blocks) over gridDim.z: partial sums go to split_ws
 * (splits * M * Cout_pad floats, caller-owned) and a second kernel reduces them in fixed order and applies the
 * epilogue; b200_conv_tc_splits returns the library's choice for a descriptor (1 = no split). */
int b200_conv_gemm_tc(const b200_conv_desc* d, const void* in_bf16, const void* wmat_bf16, const float* bias,
                      const float* scale, void* out, int out_bf16, float* split_ws, int splits, b200_stream_t stream);
int b200_conv_tc_ntile(int Cout);
/* The activation operand is fetched by TMA im2col (cuTensorMapEncodeIm2col: 128 output pixels x 64 channels per
 * request, zero fill outside the image) whenever the descriptor is a plain strided window walk, else by 16-byte
 * cp.async gathers.  b200_conv_tc_set_im2col(0) forces the cp.async path (parity tests compare the two); returns the
 * previous setting. */
int b200_conv_tc_set_im2col(int enable);
/* im2col-eligible, unsplit launches run the PERSISTENT kernel (one CTA per SM walking the output tiles, operand ring
 * running ahead across tiles, double-buffered TMEM accumulator so the epilogue overlaps the next tile's MMAs).
 * b200_conv_tc_set_persistent(mode): 0 forces the one-tile-per-CTA kernel (parity tests compare the two), 1 = one
 * persistent CTA per SM with a deep operand ring, 2 = two co-resident CTAs per SM with shallower rings, 3 = two CTAs per
 * SM for short K loops only; returns the previous mode. */
int b200_conv_tc_set_persistent(int enable);
/* Stride-1 k x k window walks (conv forward and its dgrad) whose activation slab fits shared memory run the
 * SHIFTED-WINDOW kernel: one tiled TMA box per 64-channel slab brings the zero-padded input rows of a 128-pixel output
 * tile into shared memory once, and every tap's A operand is that slab addressed from a row-shifted start (UMMA
 * descriptors swizzle on absolute shared-memory addresses), instead of one im2col load per tap.
 * b200_conv_tc_set_halo(0) disables it (parity tests compare with the im2col kernels); returns the previous setting. */
int b200_conv_tc_set_halo(int enable);
int b200_conv_tc_splits(const b200_conv_desc* d);
/* 1 if b200_conv_gemm_tc / _tf32 (elem_bytes 2 / 4) would run this descriptor on the persistent kernel — the one whose epilogue
 * can produce col_stats — else 0 */
int b200_conv_tc_stats_ok(const b200_conv_desc* d, int elem_bytes);
/* y[i] = bf16(x[i]) (round to nearest even), n % 4 == 0 */
int b200_cast_bf16(const float* x, void* y_bf16, int64_t n, b200_stream_t stream);
/* tcgen05 kind::tf32 variants — the "fp32 on tensor cores" mode (BASELINE config 2, fp32 half): operands stay fp32 tensors in
 * memory (strides in fp32 elements, multiples of 4), the TMA unit rounds them to tf32 (10-bit mantissa) on the way into shared
 * memory, accumulation is fp32 in TMEM.  Same descriptor / epilogue semantics as the bf16 entry points; wmat fp32, ldw a
 * multiple of 32, rows padded to b200_conv_tc_ntile; requires Cin % 32 == 0 (and Cout % 32 == 0 for the weight gradient) and
 * a gather the TMA im2col mode can express — b200_conv_tf32_ok(d, wgrad) tells (callers use the fp32 CUDA-core kernels
 * otherwise).  Stated operation-level bound: 2e-3 relative. */
/* hi = bf16(x), lo = bf16(x - hi): the two-term bf16 split of an fp32 tensor (n % 4 == 0).  The tf32 mode's WEIGHT gradient
 * runs as three bf16 tensor-core GEMMs on these halves (hi*hi + hi*lo + lo*hi, ~2^-16 relative — tighter than tf32): the
 * pixel-major (MN-major) operands that GEMM needs returned all-zero accumulators with kind::tf32 on this hardware / toolchain
 * (tools/probes/tf32_wgrad_probe.py), so b200_wgrad_gemm_tf32 is kept for probing only. */
int b200_split_bf16(const float* x, void* hi_bf16, void* lo_bf16, int64_t n, b200_stream_t stream);
int b200_conv_tf32_ok(const b200_conv_desc* d, int wgrad);
int b200_conv_tf32_splits(const b200_conv_desc* d);
int b200_conv_gemm_tf32(const b200_conv_desc* d, const float* in, const float* wmat, const float* bias, const float* scale,
                        float* out, float* split_ws, int splits, b200_stream_t stream);
int b200_wgrad_gemm_tf32(const b200_conv_desc* d, const float* P, const float* G, float* ws, int splits,
                         b200_stream_t stream);

/* Weight-gradient gather-GEMM:  R[m, (ty*Tw+tx)*Cg + c] = sum_{n,qy,qx} P[n,qy,qx,m] * G[n, gather(qy,qx,ty,tx), c]
 * where `d` describes the gather of G exactly as above (d->Cin = Cg) and P is addressed with d's out_* fields
 * (d->Cout = number of P channels m).  Split over the row range into `splits` partial results written to
 * ws ([splits][Cout][Th*Tw*Cin] fp32); b200_wgrad_reduce sums them in fixed order into the parameter layout. */
int b200_wgrad_gemm_f32(const b200_conv_desc* d, const void* P, int p_dt, const void* G, int g_dt, float* ws, int splits,
                        b200_stream_t stream);
/* tcgen05 variant: P and G are bf16 channel-last tensors (strides in bf16 elements, multiples of 8) */
int b200_wgrad_gemm_tc(const b200_conv_desc* d, const void* P_bf16, const void* G_bf16, float* ws, int splits,
                       b200_stream_t stream);
/* dst[m*s_m + ty*s_ty + tx*s_tx + c*s_c] (=|+=) alpha * sum_s ws[s*split_stride + m*Th*Tw*C + (ty*Tw+tx)*C + c];
 * alpha = *scale or 1; split_stride = elements between consecutive partial results (rows_total*Th*Tw*C), so a row
 * window of a taller partial buffer can be reduced by offsetting ws. */
int b200_wgrad_reduce(const float* ws, int splits, int64_t split_stride, int M, int Th, int Tw, int C, float* dst,
                      int64_t s_m, int64_t s_ty, int64_t s_tx, int64_t s_c, const float* scale, int accumulate,
                      b200_stream_t stream);
/* Weight packing from the parameter layout into GEMM matrices (fp32 or bf16), taps (ky0+kstep*j, kx0+kstep*i):
 * dst[m*ldw + (j*Tw+i)*C + c] = src[m*s_m + (ky0+kstep*j)*s_ky + (kx0+kstep*i)*s_kx + c*s_c]; rows m in [M,Mpad) and
 * columns beyond Th*Tw*C are zero filled.
 * C_dst > C: the entry is the channel slice [c_off, c_off + C) of a C_dst-channel matrix (column (j*Tw+i)*C_dst + c_off + c)
 * and writes only its valid elements; C_dst <= 0 means C_dst = C, c_off = 0. */
int b200_pack_weight(const float* src, void* dst, int dst_bf16, int M, int Mpad, int Th, int Tw, int C, int64_t ldw,
                     int64_t s_m, int64_t s_ky, int64_t s_kx, int64_t s_c, int ky0, int kx0, int kstep, int C_dst,
                     int c_off, b200_stream_t stream);
/* The same packing for MANY matrices in one launch: the refresh of every packed GEMM operand of a network after its
 * optimizer step (train64.py:258-262 / 366-370 update the weights every iteration; the operand matrices above must follow).
 * Entry e writes ONLY the valid elements dst[m*ldw + (j*Tw+i)*C_dst + c_off + c], m < M, c < C — padding (rows >= M,
 * columns >= Th*Tw*C_dst) is left untouched (the caller zero-fills it once at allocation), so several entries can fill
 * one matrix (SPADE's fused gamma|beta operand is packed from two parameters).  The kernel maps one block to one chunk = a
 * tile of B200_PACK_MT rows x B200_PACK_CT channels x all taps (staged through shared memory: coalesced on both sides); entry
 * e owns ceil(M / B200_PACK_MT) * ceil(C / B200_PACK_CT) chunks and chunk_begin = the chunks of the entries before it. */
#define B200_PACK_MT 32
#define B200_PACK_CT 64
typedef struct {
    const float* src;
    void* dst;
    int64_t ldw, s_m, s_ky, s_kx, s_c;
    int32_t dst_bf16, M, Th, Tw, C, C_dst, c_off, ky0, kx0, kstep;
    int32_t chunk_begin, pad;
} b200_pack_entry;
/* chunk_entry_dev (optional, total_chunks int32): entry index of every chunk — saves each block the binary search */
int b200_pack_weight_multi(const b200_pack_entry* entries_dev, int n_entries, int total_chunks,
                           const int32_t* chunk_entry_dev, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Normalisation family — replaces nn.BatchNorm1d/2d, ConditionalBatchNorm2d (generator_obj_att.py:31-44),
 * SPADE's param-free BN + modulation (normalization.py:94-108), fused with ReLU / residual add.
 * x, y, residual, dy, dx (rows, C) and SPADE's gamma|beta activation gb / its gradient dgb (rows, 2C) are channel-last
 * activations of storage type dt; statistics, parameter tables and parameter gradients are fp32.
 */
/* `groups`: several independent calls of the same layer batched along the row dimension (rows = groups *
 * rows_per_group, group g = rows [g*rows_per_group, (g+1)*rows_per_group)) keep SEPARATE batch statistics — this is how
 * the three generator passes / the fake+real discriminator passes of one step share a launch without merging
 * their statistics.  mean, var are (groups, C); running statistics receive the groups' updates in order.
 *
 * batch statistics (biased var) + running-stat update (momentum, unbiased var) — F.batch_norm training semantics.
 * ws: 2*C*groups*b200_bn_chunks(rows/groups, C) doubles. running_* may be NULL. */
int b200_bn_chunks(int64_t rows, int C);
/* batch statistics from the per-slab (sum, sum of squares) pairs a convolution epilogue produced (b200_conv_desc.col_stats):
 * same outputs and running-statistics update as b200_bn_stats; rows_per_group must be a multiple of 32 (slabs never straddle a
 * group); n_slabs = ceil(rows / 128) * 4 (slabs beyond the last row hold zeros). */
int b200_bn_stats_slabs(const float* col_stats, int64_t n_slabs, int ld, int64_t rows, int C, int groups, float* mean,
                        float* var, float* running_mean, float* running_var, float momentum, b200_stream_t stream);
int b200_bn_stats(const void* x, int dt, int64_t rows, int C, int groups, float* mean, float* var, float* running_mean,
                  float* running_var, float momentum, double* ws, b200_stream_t stream);
enum { B200_NORM_PLAIN = 0, B200_NORM_AFFINE = 1, B200_NORM_CBN = 2, B200_NORM_SPADE = 3 };
/* y = relu?( residual? + (x-mean)*rsqrt(var+eps) * g + b ):
 *   PLAIN g=1,b=0 | AFFINE g=gamma[c], b=beta[c] | CBN g=table[idx[row/rows_per_seg]][c], b=table[..][C+c]
 *   | SPADE g=1+gb[row][c], b=gb[row][C+c]  (gb = the fused gamma|beta conv output, (rows,2C)) */
int b200_norm_fwd(const void* x, void* y, int dt, int64_t rows, int C, int groups, const float* mean, const float* var,
                  float eps, int mode, const void* gamma, const float* beta, const int32_t* idx, int rows_per_seg,
                  const void* residual, int relu, b200_stream_t stream);
/* backward, stage 1: per-segment sums of (dyr*g, dyr*g*xhat) [and for CBN the raw (dyr, dyr*xhat)] where dyr = dy masked by
 * y>0 when relu; segments never straddle a group: seg_sums is (groups * ceil(rows_per_group / rows_per_seg), C, 2)
 * doubles.  For CBN pass rows_per_seg = H*W.  relu = 2 (CBN on a 16-byte vector layout only): the mask is recomputed as
 * g * xhat + b > 0 — the forward kernel's fp32 expression — and y may be NULL (one activation read less per pass). */
int b200_norm_bwd_reduce(const void* dy, const void* x, const void* y, int dt, int64_t rows, int C, int groups,
                         const float* mean, const float* var, float eps, int mode, const void* gamma,
                         const int32_t* idx, int rows_per_seg, int relu, double* seg_sums, b200_stream_t stream);
/* stage 2: combine segments in fixed order -> s (groups, C, 2) floats = (sum dxhat, sum dxhat*xhat); parameter gradients
 * (summed over groups): AFFINE: dgamma[c], dbeta[c]; CBN: dtable (num_classes, 2C) deterministic segmented sum by class
 * (dtable is overwritten). */
int b200_norm_bwd_finalize(const double* seg_sums, int nseg, int C, int groups, int mode, const float* gamma,
                           const int32_t* idx, int num_classes, float* s, float* dgamma, float* dbeta, float* dtable,
                           b200_stream_t stream);
/* stage 3: dx = rstd*(dxhat - s1/rows_per_group - xhat*s2/rows_per_group); SPADE additionally writes dgb (rows,2C) =
 * (dyr*xhat | dyr). */
int b200_norm_bwd_apply(const void* dy, const void* x, const void* y, void* dx, int dt, int64_t rows, int C, int groups,
                        const float* mean, const float* var, float eps, int mode, const void* gamma,
                        const int32_t* idx, int rows_per_seg, int relu, const float* s, void* dgb,
                        b200_stream_t stream);
/* eval-mode normalisation uses b200_norm_fwd with mean/var = running stats. */

/* ------------------------------------------------------------------------------------------------
 * Elementwise / pooling / layout — replace nn.ReLU, residual adds, F.avg_pool2d (discriminator.py:25-26),
 * F.interpolate nearest (normalization.py:100, generator_obj_att128.py:588), AdaptiveAvgPool2d, sum over (H,W)
 * (discriminator.py:226,270; generator_obj_att.py:445), torch.cat, nn.Embedding, ConvLSTM gate math
 * (generator_obj_att.py:102-112), the VAE reparameterisation (generator_obj_att.py:417-420) and the
 * embedding (x) mask broadcast of LayoutEncoder (generator_obj_att.py:489-490).
 */
int b200_relu_fwd(const void* x, void* y, int64_t n, int dt, b200_stream_t stream);
int b200_relu_bwd(const void* dy, const void* y, void* dx, int64_t n, int dt, b200_stream_t stream);
int b200_add(const void* a, const void* b, void* out, int64_t n, int dt, b200_stream_t stream);
/* y[n, qy, qx, c] = scale * sum_{f x f block} x[n, qy*f+dy, qx*f+dx, c]  (x: (N,H,W,C), H%f==0) */
int b200_pool_fwd(const void* x, void* y, int N, int H, int W, int C, int f, float scale, int dt, b200_stream_t stream);
/* y = scale * pooled sum of (a + b): the residual branch and the shortcut of a downsampling discriminator block
 * (discriminator.py:98-99: pool(resi(x)) + pool(sc(x)), pooling being linear) in one pass; relu != 0 applies the ReLU every
 * consumer of the block output starts with */
int b200_pool_add_fwd(const void* a, const void* b, void* y, int N, int H, int W, int C, int f, float scale, int relu,
                      int dt, b200_stream_t stream);
/* out = relu(a + b) (a block's output when every consumer applies ReLU first, discriminator.py:70-74,224) */
int b200_add_relu(const void* a, const void* b, void* out, int64_t n, int dt, b200_stream_t stream);
/* y = scale * x[.., iy/f, ix/f, ..] where mask[.., iy/f, ix/f, ..] > 0, else 0 (mask shaped like x): the gradient of
 * relu(pool(.)) in one pass */
int b200_unpool_masked_fwd(const void* x, const void* mask, void* y, int N, int H, int W, int C, int f, float scale, int dt,
                           b200_stream_t stream);
/* y[n, iy, ix, c] = scale * x[n, iy/f, ix/f, c]  (x: (N,H,W,C) -> y: (N,H*f,W*f,C)) */
int b200_unpool_fwd(const void* x, void* y, int N, int H, int W, int C, int f, float scale, int dt, b200_stream_t stream);
/* out[row] = [ a[row / a_div][0:Ca] | b[row / b_div][0:Cb] ] and its adjoint (sums over the broadcast rows, fixed order) */
int b200_concat_fwd(const void* a, int Ca, int a_div, const void* b, int Cb, int b_div, void* out, int64_t rows,
                    int dt, b200_stream_t stream);
int b200_concat_bwd(const void* dout, int Ca, int a_div, void* da, int Cb, int b_div, void* db, int64_t rows,
                    int dt, b200_stream_t stream);
int b200_gather_rows(const float* table, const int32_t* idx, float* out, int rows, int D, b200_stream_t stream);
/* dtable[k] = sum over rows with idx==k, ascending row order (deterministic); dtable (num_classes, D) overwritten */
int b200_scatter_rows(const float* dout, const int32_t* idx, float* dtable, int rows, int D, int num_classes,
                      b200_stream_t stream);
/* out[o, y+1, x+1, c] = mask[o,y,x] * v[o,c], zero ring (1x1 conv with padding 1 of a rank-1 tensor); v, mask, dv fp32,
 * out / dout activations of type dt */
int b200_mask_outer_fwd(const float* v, const float* mask, void* out, int O, int H, int W, int C, int dt,
                        b200_stream_t stream);
int b200_mask_outer_bwd(const void* dout, const float* mask, float* dv, int O, int H, int W, int C, int dt,
                        b200_stream_t stream);
/* ConvLSTM cell pointwise part; pre = pre_x + pre_h, channels ordered [i|f|o|g] each `hid` wide; rows = n*H*W.
 * gates (rows,4*hid) receives the activated gates for the backward. pre_h / c_prev may be NULL (t = 0). */
/* pre_x, pre_h, h_out, dh, dpre are activations of type dt; the cell state c and the saved gates stay fp32 */
int b200_lstm_gates_fwd(const void* pre_x, const void* pre_h, const float* c_prev, float* gates, float* c_out,
                        void* h_out, int64_t rows, int hid, int dt, b200_stream_t stream);
/* dpre (rows,4*hid), dc_prev (rows,hid) from dh, dc_next (NULL = 0), saved gates, c_prev (NULL = 0), c_out */
int b200_lstm_gates_bwd(const void* dh, const float* dc_next, const float* gates, const float* c_prev,
                        const float* c_out, void* dpre, float* dc_prev, int64_t rows, int hid, int dt,
                        b200_stream_t stream);
int b200_reparam_fwd(const float* mu, const float* logvar, const float* eps, float* z, int64_t n, b200_stream_t stream);
int b200_reparam_bwd(const float* dz, const float* logvar, const float* eps, float* dmu_add, float* dlogvar_add,
                     int64_t n, b200_stream_t stream);
/* out[c] = sum_rows x[row][c] (bias gradients), deterministic; ws: C*b200_bn_chunks(rows,C) doubles */
int b200_colsum(const void* x, int64_t rows, int C, int dt, float* out, double* ws, b200_stream_t stream);
/* im2col packing of a few-channel convolution input (the 3-channel image / crop convolutions; replaces the first
 * conv of CropEncoder generator_obj_att.py:371, OptimizedBlock discriminator.py:29-60 in the tcgen05 precision mode):
 * out[m][k] (bf16, row length Kp, a multiple of 64) = x[n, c, qy*stride + ky - pad, qx*stride + kx - pad] (0 outside the
 * image) with m = (n*Hy + qy)*Wy + qx and k = (ky*kw + kx)*Cx + c; columns k >= Cx*kh*kw are zero.  x is addressed with
 * element strides (sn, sh, sw, sc), storage type x_dt.  The result is the channel-last activation of an equivalent 1x1
 * convolution with Kp input channels.  flip != 0 walks the transposed window, x[n, c, qy + pad - ky, qx + pad - kx]
 * (stride 1): the im2col matrix of an output gradient, shared by the data gradient and the weight gradient of a
 * convolution with few OUTPUT channels (Decoder c4, generator_obj_att.py:572). */
int b200_im2col_pack(const void* x, int x_dt, int64_t N, int Hx, int Wx, int Cx, int64_t sn, int64_t sh, int64_t sw,
                     int64_t sc, int kh, int kw, int stride, int pad, int Hy, int Wy, int Kp, int flip, void* out_bf16,
                     b200_stream_t stream);
/* out[r] = sum_l x[r*L + l] (fixed order; the per-(n, c) sums of an NCHW output gradient: bias gradients) */
int b200_rowsum(const void* x, int dt, int64_t rows, int64_t L, float* out, b200_stream_t stream);
/* y[b][c][r] = x[b][r][c]: batched (R x C) transpose, i.e. NCHW <-> channel-last at module boundaries */
int b200_transpose(const float* x, float* y, int B, int R, int C, b200_stream_t stream);
/* row gather / scatter-overwrite for the time-major packing of ConvLSTM sequences: out[r] = x[src_row[r]] (rows of
 * row_bytes bytes, a multiple of 16; src_row < 0 writes zeros) */
int b200_permute_rows(const void* x, const int32_t* src_row, void* out, int rows, int64_t row_bytes, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Spectral normalisation — replaces torch.nn.utils.spectral_norm's pre-forward hook installed by add_sn
 * (discriminator.py:15-22): one power iteration in place on u (h) and v (w), sigma = u.(W v); W (h,w) row-major
 * view of weight_orig.  *sigma_out = sigma, *inv_sigma_out = 1/sigma.  ws: 8*w + h floats.
 * b200_sn_grad: dW = g*inv_sigma - (<g, W> * inv_sigma^2) * u v^T  (gradient through W/sigma with u, v constant),
 * g = gradient w.r.t. the normalised weight, same layout as W.  ws: 1024 doubles.
 */
int b200_sn_power_iter(const float* W, int h, int w, float* u, float* v, int do_iter, float eps, float* sigma_out,
                       float* inv_sigma_out, float* ws, b200_stream_t stream);
int b200_sn_grad(const float* g, const float* W, const float* u, const float* v, const float* inv_sigma, float* dW, int h,
                 int w, int accumulate, double* ws, b200_stream_t stream);

/* Weight gradient of `groups` (<= 8) batched calls of one spectral-normalised layer, each with its own sigma_g, u_g, v_g:
 * ws holds the fp32 partial results ([groups*splits_per_group][M][T*C], split_stride elements apart) of ONE
 * b200_wgrad_gemm_* launch whose pixel splits are aligned with the call boundaries (splits [g*spg, (g+1)*spg) = call g).
 *   G_g = sum of call g's splits, transposed to the parameter layout (M, C, T);
 *   dW  = sum_g ( G_g * inv[g] - <G_g, W> * inv[g]^2 * u_hist[g] v_hist[g]^T ),  W viewed (M, C*T).
 * Gbuf: unused (may be NULL; the per-call gradients are no longer materialised), dot_part: groups*b200_sn_wgrad_parts(M, C)
 * doubles (scratch).  C % 4 == 0, T <= 64. */
int b200_sn_wgrad_parts(int M, int C);
int b200_sn_wgrad_finish(const float* ws, int groups, int splits_per_group, int64_t split_stride, int M, int T, int C,
                         const float* W, const float* u_hist, const float* v_hist, const float* inv, float* Gbuf,
                         double* dot_part, float* dW, b200_stream_t stream);

/* The pooled convolution (discriminator.py:52-57, 90-96: the blocks' second convolution is followed by avg_pool2d(2)):
 * avg_pool2(conv_{kh x kw, stride 1, pad p}(x; W)) = conv_{(kh+1) x (kw+1), stride 2, pad p}(x; W4),
 * W4[f][a][b] = 0.25 * sum_{i,j in {0,1}} W[f][a-i][b-j] (f = filter (cout, cin); terms outside the window dropped).
 * b200_fold_pool_weight writes W4 (filters x (kh+1) x (kw+1) fp32) from W (filters x kh x kw fp32).
 * b200_sn_wgrad_finish_pooled is b200_sn_wgrad_finish for such a layer: `ws` holds the partials of the FOLDED convolution's
 * weight gradient ((M, (Th+1)*(Tw+1), C) per split, split_stride = M*(Th+1)*(Tw+1)*C) and the transpose of the fold
 * (G[m][c][k][l] = 0.25 * sum_{i,j} G4[m][k+i][l+j][c]) is applied while the splits are summed; W, dW: (M, C, Th, Tw). */
int b200_fold_pool_weight(const float* w, float* w4, int64_t filters, int kh, int kw, b200_stream_t stream);
int b200_sn_wgrad_finish_pooled(const float* ws, int groups, int splits_per_group, int64_t split_stride, int M, int Th,
                                int Tw, int C, const float* W, const float* u_hist, const float* v_hist, const float* inv,
                                double* dot_part, float* dW, b200_stream_t stream);

/* The same power iteration for EVERY spectral-normalised layer of a network at once (4 launches per iteration instead
 * of 4 per layer), `iters` times in sequence — one per batched call of the network.  `layers` is a DEVICE array of
 * n_layers descriptors; iteration `it` of layer l records inv[it] = 1/sigma, u_hist[it*h ..], v_hist[it*w ..] (the
 * vectors that sigma was computed with; either history pointer may be NULL).  ws: 8*w + h floats per layer. */
typedef struct {
    const float* W;     /* (h, w) row-major view of weight_orig */
    float* u;           /* (h,)  updated in place */
    float* v;           /* (w,)  updated in place */
    float* ws;          /* 8*w + h floats of scratch */
    float* inv;         /* (iters,) */
    float* u_hist;      /* (iters, h) or NULL */
    float* v_hist;      /* (iters, w) or NULL */
    int h, w;
} b200_sn_layer;
int b200_sn_power_iter_multi(const b200_sn_layer* layers, int n_layers, int max_h, int max_w, int iters, int do_iter,
                             float eps, b200_stream_t stream);

/* Multi-tensor Adam (torch.optim.Adam(params, lr, betas) as train64.py:111-114 builds it: no weight decay, no amsgrad):
 * ONE launch updates every chunk listed in the DEVICE array `entries` (a tensor is split into chunks of <= 65536 elements by
 * the host so that blocks are balanced).  *step_dev (float, on the device) is incremented first and supplies the bias
 * corrections, so the update can be captured in a CUDA graph.
 *   m = m + (g - m)*(1-b1);  v = v*b2 + (1-b2)*g*g;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps) */
typedef struct {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int32_t n;
    int32_t pad;
} b200_adam_entry;
int b200_adam_multi(const b200_adam_entry* entries_dev, int n_entries, float* step_dev, double lr, double beta1,
                    double beta2, double eps, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Data contract either side of the step, bit-exact integer / fp32 restatements:
 * b200_one_hot_attributes: data/vg_custom_mask.py:160-171 — att_idx (O,A) int64, -1 terminated per row ->
 *   out (O,n_att) fp32 multi-hot (entries after the first -1 are ignored, as the loader's while-loop does).
 * b200_imagenet_deprocess: data/utils.py:32-66 imagenet_deprocess_batch — imgs (N,C,H*W) fp32 normalised images ->
 *   out uint8: v = (x / inv_std[c]) - neg_mean[c]; rescale != 0: per image (v - min) / (max - min); then *255, clamp to
 *   [0,255], truncation.  inv_std / neg_mean (C floats, device): the caller passes fp32(1/std), fp32(-mean) as the
 *   reference builds them on the host.  ws: 2*N floats (per-image min, max). */
int b200_one_hot_attributes(const int64_t* att_idx, int O, int A, int n_att, float* out, b200_stream_t stream);
int b200_imagenet_deprocess(const float* imgs, int N, int C, int HW, const float* inv_std, const float* neg_mean,
                            int rescale, uint8_t* out, float* ws, b200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Step arithmetic (train64.py:195-252, 284-364): each loss term is one launch producing
 * its partial sums (doubles, one per block, row `slot` of partials[n_terms][B200_LOSS_MAX_BLOCKS], block count in
 * counts[slot]) AND the gradient w.r.t. its inputs; b200_loss_total sums every slot in index order into terms[slot] and the
 * grand total into terms[n_terms].  All tensors fp32; `scale` carries the lambda of the term; `weight` (per group) the
 * 0.4 / 0.4 / 0.2 mix of the rec / rand / shift passes (batched calls are groups of n rows).
 *   bce_groups  : scale * sum_g weight[g] * mean_i BCE_with_logits(x[g*n+i], target[g])            grad (groups*n)
 *                 (groups >= split_group are summed into slot + 1: the fake and the real half of an adversarial loss)
 *   ce_groups   : scale * sum_g weight[g] * mean_r CE(x[g*n+r, :C], label[r])                      grad (groups*n, C)
 *   bce_pw_rows : scale * sum_g weight[g] * mean_{r: sel[r] != 0, a} BCE(x[g*n+r, a], t[r, a]; pos_weight[a])   grad (groups*n, A)
 *   l1_rows     : scale * sum_n mask[n] * mean_L |a[n,:] - b[n,:]| / denom  (b_stride_n = 0: one shared b row; mask NULL = 1)
 *   kl          : scale * -0.5 * sum(1 + logvar - mu^2 - exp(logvar))                              dmu, dlogvar */
#define B200_LOSS_MAX_BLOCKS 1024
int b200_loss_bce_groups(const float* x, int n, int groups, int split_group, const float* target, const float* weight,
                         float scale, float* grad, double* partials, int* counts, int slot, b200_stream_t stream);
int b200_loss_ce_groups(const float* x, const int64_t* label, int n, int groups, int C, const float* weight, float scale,
                        float* grad, double* partials, int* counts, int slot, b200_stream_t stream);
int b200_loss_bce_pw_rows(const float* x, const float* t, const float* sel, int n, int groups, int A, int n_sel,
                          const float* pos_weight, const float* weight, float scale, float* grad, double* partials,
                          int* counts, int slot, b200_stream_t stream);
int b200_loss_l1_rows(const float* a, const float* b, int N, int64_t L, int64_t b_stride_n, const float* mask, float denom,
                      float scale, float* grad, double* partials, int* counts, int slot, b200_stream_t stream);
int b200_loss_kl(const float* mu, const float* logvar, int64_t n, float scale, float* dmu, float* dlogvar, double* partials,
                 int* counts, int slot, b200_stream_t stream);
int b200_loss_total(const double* partials, const int* counts, int n_terms, float* terms, b200_stream_t stream);

/* plain device-to-device copy on the stream (row concatenation of batched calls) */
int b200_copy(void* dst, const void* src, size_t bytes, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200GAN_H */
