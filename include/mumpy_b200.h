/*
 * mumpy_b200.h -- C ABI of libmumpy_b200.so: the sm_100a kernels behind the Mumpy inference forward.
 *
 * The reference (yuxiaoxiangyong/Multilateral-Temporal-view-Pyramid-Transformer-for-Video-Inpainting-
 * Detection) has no native layer: every function below replaces a group of ATen calls made by one
 * Python method of the reference, cited as <file>:<lines> relative to the reference root.  The Python
 * mirror classes in the package (same names / ctor / forward / state_dict keys as the reference) are
 * the only intended callers; INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - plain pointers + sizes; every pointer is DEVICE memory owned by the caller (incl. workspaces);
 *     the library never allocates, frees, synchronises or keeps references (CUDA-graph capturable);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, a negative MUMPY_ERR_* otherwise; mumpy_last_error() gives the text;
 *   - token tensors are "canvases": row-major (B, T*H, W, C), frame t in rows [t*H,(t+1)*H)
 *     (multiTemporalViewEncoder.py:614,707; swinTransformer.py:267); decoder maps are NHWC;
 *   - dtype codes: MUMPY_F32 / MUMPY_BF16 / MUMPY_F16.  MUMPY_F32 GEMMs run exact fp32 FMA kernels (the <=1e-4
 *     parity mode); MUMPY_BF16 and MUMPY_F16 GEMMs run the TMA + tcgen05 + TMEM kernel on 16-bit operands with fp32
 *     accumulation.  Wherever a prototype below says "bf16" for an operand buffer, MUMPY_F16 buffers are accepted too
 *     (same kernels instantiated for IEEE half); a call never mixes the two 16-bit types.
 *   - no CPU fallback exists anywhere in this library.
 */
#ifndef MUMPY_B200_H
#define MUMPY_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MUMPY_F32 0
#define MUMPY_BF16 1
#define MUMPY_F16 2      /* IEEE half operands: same tensor-core rate as bf16, 3 more mantissa bits (the "fp16" precision mode) */

#define MUMPY_ACT_NONE 0
#define MUMPY_ACT_GELU 1      /* exact erf GELU (nn.GELU default) */
#define MUMPY_ACT_RELU 2
#define MUMPY_ACT_SIGMOID 3

#define MUMPY_OK 0
#define MUMPY_ERR_ARG (-1)
#define MUMPY_ERR_CUDA (-2)
#define MUMPY_ERR_UNSUPPORTED (-3)

/* resample modes of mumpy_resample_nhwc */
#define MUMPY_RS_IDENTITY 0
#define MUMPY_RS_UP_ALIGNED 1     /* nn.Upsample(bilinear, align_corners=True)  decoder.py:72,79,86,93 */
#define MUMPY_RS_UP_HALFPIX 2     /* nn.Upsample(bilinear) default align_corners=False decoder.py:10,136-137 */
#define MUMPY_RS_AVGPOOL2 3       /* nn.AvgPool2d(2) decoder.py:149..176 */
#define MUMPY_RS_PIXEL_SHUFFLE2 4 /* nn.PixelShuffle(2) decoder.py:128 */

int mumpy_abi_version(void);
/* Binds the calling thread to `device` and resolves the driver entry point used to encode TMA maps. */
int mumpy_init(int device);
const char *mumpy_last_error(void);
/* Programmatic dependent launch of the library's kernels (default on; environment MUMPY_PDL=0 disables): each kernel
 * may be scheduled while its predecessor in the stream drains and synchronises on it in-kernel (griddepcontrol.wait). */
int mumpy_set_pdl(int enabled);
/* fp16 range guard.  In the "fp16" precision mode every fp32 -> IEEE-half conversion saturates to +-65504 instead of producing
 * inf, and the kernels that create new magnitudes in operand precision (GEMM / convolution epilogues, casts, gathers, LayerNorm,
 * the DCT repack) OR 1 into the caller-owned device word registered here whenever they had to saturate.  The caller reads the
 * word back with the step's results (ops.f16_overflowed(), evaluate.py raises) -- the reference computes in fp32
 * (test.py:94-95), so leaving the half range must be loud, never silent.  NULL disables reporting (saturation stays).
 * Synchronous (cudaMemcpyToSymbol); call once per device, outside stream capture. */
int mumpy_set_f16_overflow_flag(unsigned int *flag_dev);
/* CTA-pair (tcgen05 cta_group::2, 256 x BN tiles on a (2,1,1) cluster) policy of the bf16 GEMM / implicit-GEMM convolution:
 * 0 never, 1 the tile cost model decides, 2 whenever the shape allows, 3 cost model for K >= 1024 only, 4 (default) cost model
 * except for 16-bit outputs without residual (they keep the lean 1-CTA kernel).  Environment: MUMPY_TC_PAIR. */
int mumpy_set_gemm_pair_mode(int mode);
/* Tuning aid: force the GEMM tile width (a divisor of N; 0 = the cost model decides).  Environment: MUMPY_TC_BN. */
int mumpy_set_gemm_tile(int bn);
/* Kernel used by mumpy_window_attention in the 16-bit modes when the bias comes from the relative-position table:
 * 1 (default) tcgen05 / TMEM (two windows per 128-row accumulator), 0 the per-warp mma.sync kernel.  Environment: MUMPY_ATT_TC. */
int mumpy_set_attention_tc(int enabled);

/* nn.Linear / 1x1 conv:  out = act(A . W^T + bias) (+ residual).   swinTransformer.py:45-51,142,164,365;
 * blocks.py:28-34,56,71; deformableAttention.py:333,361-362,402; multiTemporalViewEncoder.py:283,740;
 * decoder.py:98-120 (Conv3d k=(3,1,1) as a per-pixel GEMM).
 * A (M,K) row-major with row stride lda, W (N,K) row-major, both of `ab_dtype`; bias (N) fp32 or NULL;
 * residual (M,N) fp32 with row stride ldo or NULL (may alias out); out (M,N) of `out_dtype`, row stride ldo. */
int mumpy_linear(const void *A, long lda, const void *W, const float *bias, const float *residual, void *out,
                 long ldo, long M, int N, int K, int ab_dtype, int out_dtype, int act, void *stream);

/* LayerNorm fused into the nn.Linear that consumes it (16-bit operand modes only):
 *   out = act( LayerNorm(x; gamma, beta, eps) . W^T + bias )
 * replaces `self.norm1(x)` + `self.qkv(x)` (swinTransformer.py:266,142; multiTemporalViewEncoder.py:246,
 * WindowAttention.forward) and `self.norm2(x)` + `self.fc1(x)` + `self.act(x)` (swinTransformer.py:305,47-48) of a Swin /
 * CrossSwin block: the normalised rows are produced in shared memory as the tcgen05 A operand and never written to global memory.
 * x (M,K) fp32 row-major (the residual stream), gamma/beta (K) fp32, W (N,K) row-major of `w_dtype` (MUMPY_BF16 / MUMPY_F16),
 * bias (N) fp32 or NULL, out (M,N) of `w_dtype` with row stride ldo; act MUMPY_ACT_NONE or MUMPY_ACT_GELU.
 * Shapes: K in {96,128,192,256,384,512} and N a multiple of 64 or 96 (mumpy_ln_linear_supported returns 1); anything else is an error --
 * the caller then runs mumpy_layernorm + mumpy_linear.  Statistics are exact two-pass fp32, bit-identical to mumpy_layernorm. */
int mumpy_ln_linear_supported(int N, int K);
/* Fused Swin MLP for the narrow stages (16-bit operand modes):  out = x + fc2( GELU( fc1( LayerNorm(x) ) ) )
 * replaces `x = x + self.drop_path(self.mlp(self.norm2(x)))` (swinTransformer.py:305 with Mlp.forward :45-51;
 * multiTemporalViewEncoder.py:289) in ONE kernel: the normalised rows and the 4C-wide hidden activations only exist in shared /
 * tensor memory.  x, out (M,C) fp32 row-major (the residual stream; out may not alias x), gamma/beta (C) fp32, W1 (4C,C) and
 * W2 (C,4C) row-major of `w_dtype`, b1 (4C), b2 (C) fp32.  C in {96,128,192,256} (mumpy_mlp_fused_supported returns 1); wider
 * blocks run mumpy_layernorm / mumpy_ln_linear + mumpy_linear.  Bit-identical to the unfused kernels. */
int mumpy_mlp_fused_supported(int C);
/* Kernel shape of mumpy_mlp_fused: 0 (default) the pipelined one (GELU / output / LayerNorm warp groups on neighbouring tiles),
 * 1 the serial one (one group of 16 warps runs the three phases back to back), 2 the serial one with two CTAs per SM where it fits
 * (C <= 128).  All three give bit-identical results.  Environment: MUMPY_MLP_SHAPE. */
int mumpy_set_mlp_fused_shape(int shape);
int mumpy_mlp_fused(const float *x, const float *gamma, const float *beta, float eps, const void *W1, const float *b1, const void *W2,
                    const float *b2, float *out, long M, int C, int w_dtype, void *stream);
/* CTA-pair policy of mumpy_ln_linear (two adjacent 128-row tiles on a (2,1,1) cluster, tcgen05 cta_group::2, each CTA streaming half of
 * every weight tile): 0 never, 1 the cost model decides (default), 2 whenever the shape allows.  Environment: MUMPY_LG_PAIR. */
int mumpy_set_ln_linear_pair_mode(int mode);
int mumpy_ln_linear(const float *x, const float *gamma, const float *beta, float eps, const void *W, const float *bias, void *out,
                    long ldo, long M, int N, int K, int w_dtype, int act, void *stream);
/* Same GEMM on bf16 operands with two results: out = act(A . W^T + bias) + residual (fp32) and
 * aux16 = (operand type)(act(A . W^T + bias)) -- the un-summed W-MSA branch that CrossSwinBlock returns as the next view's
 * key/value source (multiTemporalViewEncoder.py:275-276) together with the shortcut sum, in one pass. */
int mumpy_linear_dual(const void *A, long lda, const void *W, const float *bias, const float *residual, float *out,
                      void *aux16, long ldo, long M, int N, int K, int ab_dtype, int act, void *stream);

/* nn.LayerNorm over the last dim (eps inside the sqrt).  swinTransformer.py:266,305; blocks.py:88,92. */
int mumpy_layernorm(const float *x, const float *gamma, const float *beta, void *out, int out_dtype, long rows,
                    int C, float eps, void *stream);

/* PatchMerging gather + LayerNorm(4C): out row (b, r', c') = LN(cat[x(2r',2c'), x(2r'+1,2c'), x(2r',2c'+1),
 * x(2r'+1,2c'+1)]).  swinTransformer.py:355-364.  x canvas (B,TH,W,C) fp32 -> out (B*TH/2*W/2, 4C). */
int mumpy_patch_merge_norm(const float *x, const float *gamma, const float *beta, void *out, int out_dtype, int B,
                           int TH, int W, int C, float eps, void *stream);

/* Window attention core of WindowAttention.forward (swinTransformer.py:142-163) with the cyclic shift and
 * window partition/reverse of SwinTransformerBlock.forward (:270-301) folded into the addressing.
 * qkv (B*TH*W, 3C) canvas order, channel o -> (which=o/C, head=(o%C)/d, o%d); bias (heads,N,N) fp32 =
 * relative_position_bias_table gathered by relative_position_index; mask (nW,N,N) fp32 or NULL;
 * out (B*TH*W, C) canvas order; N = ws*ws (<=64), d = C/heads in {32,64}.
 * Optional fast path (bf16, d = 32): rel_table = the raw relative_position_bias_table ((2ws-1)^2, heads) fp32 and
 * standard_mask = 1 when `mask` is exactly the Swin shift mask of swinTransformer.py:233-252 for (TH, W, ws, shift): the
 * kernel then looks the bias up in a shared-memory copy of the table and recomputes the mask from region ids. */
int mumpy_window_attention(const void *qkv, const float *bias, const float *mask, const float *rel_table, int standard_mask,
                           void *out, int dtype, int B, int TH, int W, int C, int heads, int ws, int shift, void *stream);

/* Attention.forward core for short sequences (blocks.py:64-70): qkv (Bn*N, 3C) -> out (Bn*N, C), N <= 8. */
int mumpy_mha_short(const void *qkv, void *out, int dtype, long Bn, int N, int C, int heads, void *stream);
/* The attention map of the same call, softmax(q k^T d^-1/2): probs (Bn, heads, N, N) fp32 -- what Attention.forward returns next to x
 * (blocks.py:66-68,74) and Block.forward(return_attention=True) returns alone (:88-89).  Diagnostic, not on the hot path. */
int mumpy_mha_short_probs(const void *qkv, float *probs, int dtype, long Bn, int N, int C, int heads, void *stream);

/* CrossThreeViewTokenize, one view (multiTemporalViewEncoder.py:605-618): Conv3d(3->C, k=s=(kt,4,4)) + LayerNorm.
 * x (B,T,3,S,S) fp32; w_kc (3*kt*16, C) fp32 = weight.reshape(C,-1).T; out (B, To*(S/4)^2, C) fp32, To = T/kt. */
int mumpy_tokenize(const float *x, const float *w_kc, const float *bias, const float *gamma, const float *beta,
                   float *out, int B, int T, int S, int kt, int C, float eps, void *stream);

/* The same layer on the tensor cores (16-bit modes), step 1: the patches as a GEMM operand.  x (B,T,3,S,S) fp32 ->
 * out (B*To*(S/4)^2, 3K) of `out_dtype`, K = 3*kt*16 in Conv3d's weight.reshape(C,-1) order, each row = [hi | hi | lo] with
 * value = hi + lo; mumpy_linear against the weight rows [w_hi | w_lo | w_hi] (fp32 out, + bias) and mumpy_layernorm (fp32 out)
 * complete multiTemporalViewEncoder.py:605-618 at fp32 accuracy. */
int mumpy_patchify16(const float *x, void *out, int out_dtype, int B, int T, int S, int kt, void *stream);

/* FAF on the middle frame (dct.py:71-79; multiTemporalViewEncoder.py:734).  x (B,T,3,S,S) fp32, frame index
 * `frame`; dct (S,S) fp32 DCT-II matrix; ws fp32 workspace of 5*B*3*S*S floats; out (B,9,S,S) fp32,
 * channel = band*3 + rgb; band k keeps lo_k <= i+j <= hi_k, band_lo_hi6 = HOST array {lo0,hi0,lo1,hi1,lo2,hi2}
 * (read during the call; the only non-device pointer in this ABI). */
/* Workspace sizes: floats of mumpy_faf's `ws`, bytes of mumpy_faf16's `ws16`. */
long mumpy_faf_workspace_floats(int B, int S);
long mumpy_faf16_workspace_bytes(int B, int S);
int mumpy_faf(const float *x, const float *dct, float *ws, float *out, int B, int T, int frame, int S,
              const int *band_lo_hi6, void *stream);

/* The same transform on the tensor cores (16-bit modes): the four DCT passes run as tcgen05 GEMMs over split operands
 * (a = a_hi + a_lo in `dtype`, three partial products accumulated in fp32: error 2.9e-5 (bf16) / 2.5e-6 (f16) on outputs ~3),
 * whose epilogues write each result already transposed per image, band masked and split -- the operand of the next pass (one
 * split kernel for the input frame, no repack kernels in between).  dcat / dtcat (S, 3S) of `dtype` = [D_hi | D_lo | D_hi] for D
 * and for D^T; ws16: 2 x (9*B*S x 3S) 16-bit values (the passes ping-pong between the halves); ws32: unused, may be NULL. */
int mumpy_faf16(const float *x, const void *dcat, const void *dtcat, void *ws16, float *ws32, float *out, int B, int T,
                int frame, int S, const int *band_lo_hi6, int dtype, void *stream);

/* SwinDAttention pieces (deformableAttention.py:324-405); windows are addressed on canvases.
 * offsets: q (B*L1, C) fp32 canvas of the query view -> pix (N1, groups, P, 2) fp32 sampling positions in pixel
 *          units (y,x) of an aligned-corners ws x ws grid (:334-356).  dw_w (Cg,25), dw_b, ln_g, ln_b (Cg), pw (2,Cg). */
int mumpy_cva_offsets(const float *q, const float *dw_w, const float *dw_b, const float *ln_g, const float *ln_b,
                      const float *pw, float *pix, int B, int TH1, int W, int C, int groups, int ws, void *stream);
/* sample: x2 (B*L2, C) canvas of `x2_dtype` (after `pre`) -> sampled (N2*P, C) window-major rows of `out_dtype`
 *         (bf16 input needs bf16 output); kv window j uses the offsets of query window qidx(j) (see mumpy_cva_attention). */
int mumpy_cva_sample(const void *x2, int x2_dtype, const float *pix, void *sampled, int out_dtype, int B, int TH1, int TH2,
                     int W, int C, int groups, int ws, int per_clip_pairing, void *stream);
/* attention: q (B*L1,C) fp32 canvas, kv (N2*P, 2C) window-major of `kv_dtype` -> o (N1*P, C) of `out_dtype`
 *         o[i] = sum_t softmax(q[qidx(r*i+t)] k[r*i+t]^T * d^-1/2) v[r*i+t], r = N2/N1 (:329-330,364,390-395);
 *         qidx(j) = j mod N1 (reference, batch-global) or the same map applied inside each clip. */
int mumpy_cva_attention(const float *q, const void *kv, int kv_dtype, void *o, int out_dtype, int B, int TH1,
                        int TH2, int W, int C, int heads, int ws, int per_clip_pairing, void *stream);
/* The attention map of mumpy_cva_attention: probs (N2, heads, P, P) fp32 = the reference's (N1, r * heads, P, P) view
 * (deformableAttention.py:364,389,399; returned by SwinDAttention.forward and by CVAModule.forward(return_attention=True),
 * multiTemporalViewEncoder.py:134-137).  Diagnostic, not on the hot path. */
int mumpy_cva_attention_probs(const float *q, const void *kv, int kv_dtype, float *probs, int B, int TH1, int TH2, int W, int C,
                              int heads, int ws, int per_clip_pairing, void *stream);
/* residual: x_new[b,l,:] = h[b,l,:] + h[b, canvas(l/P, l%P), :] + reinterpret(y)[b,l,:]
 *         (deformableAttention.py:403 raw reshape; multiTemporalViewEncoder.py:138,284-286). y (N1*P, C) fp32. */
int mumpy_cva_residual(const float *h, const float *y, float *x_new, int B, int TH1, int W, int C, int ws,
                       void *stream);

/* Row gather into a column slice of a wider matrix (merge_views_along_channel_axis,
 * multiTemporalViewEncoder.py:710-718; decoder.py:43-53): for out row r = b*rows_out + q,
 *   src row = b*rows_src + (q / div) * mul_hi + (q % div) * mul_lo + add. */
int mumpy_gather_rows(const float *src, int C, void *dst, int dst_dtype, long dst_ld, int dst_col, int B,
                      int rows_out, int rows_src, int div, int mul_hi, int mul_lo, int add, void *stream);

/* Decoder primitives on NHWC maps (decoder.py:183-225). */
/* conv (stride 1): in (B,H,W,Cin) row stride ld_in, w (Cout, kh, kw, Cin) fp32, out (B,H,W,Cout) fp32. */
int mumpy_conv2d_nhwc(const float *in, long ld_in, const float *w, const float *bias, float *out, long ld_out, int B,
                      int H, int W, int Cin, int Cout, int kh, int kw, int ph, int pw, void *stream);
/* tensor-core implicit-GEMM convolution ('same', stride 1): in (B,H,W,Cin) bf16 NHWC with pixel stride ld_in (a channel
 * slice of a wider map is allowed); w_packed (Cout, kh*kw*ceil(Cin/64)*64) bf16, K order (ky,kx,c), each tap's channels
 * zero padded to a multiple of 64; out (B*H*W, Cout) fp32|bf16 = act(conv + bias) (+ residual).  The A tiles are
 * fetched by an im2col-mode TMA tensor map (padding = zero fill), no im2col buffer is materialised.
 * splitk_ws (optional, caller-owned, 16-byte aligned, splitk_ws_bytes long): when the map is small and the reduction long
 * (gcm1: 32 x 7 x 7 pixels, K = 7 * 2560) the k-blocks are split over otherwise idle SMs -- as many ranges as fit the
 * workspace at B*H*W*Cout*4 bytes each -- and summed by a second kernel that applies bias / act / residual. */
int mumpy_conv2d_nhwc_bf16(const void *in, long ld_in, const void *w_packed, const float *bias, const float *residual,
                           void *out, long ld_out, int B, int H, int W, int Cin, int Cout, int kh, int kw, int ph, int pw,
                           int in_dtype, int out_dtype, int act, float *splitk_ws, long splitk_ws_bytes, void *stream);
/* Cout == 1 convolution (final_out 3x3, decoder.py:95): in (B,H,W,Cin) fp32 contiguous, w (kh,kw,Cin), out (B,H,W). */
int mumpy_conv2d_nhwc_cout1(const float *in, const float *w, const float *bias, float *out, int B, int H, int W, int Cin,
                            int kh, int kw, int ph, int pw, void *stream);
/* im2col for the tensor-core path: out (B*H*W, Kpad) bf16, K order (ky,kx,c), zero padded to Kpad. */
int mumpy_im2col_nhwc(const float *in, long ld_in, void *out, int out_dtype, int B, int H, int W, int Cin, int kh, int kw,
                      int ph, int pw, int Kpad, void *stream);
/* GroupNorm + activation; stats over (H*W, C/groups) per (b, group), exact two-pass per pixel chunk + Chan combine.
 * stats_ws: 2*B*groups*(1 + nchunks) floats, nchunks = ceil(HW / min(64, 12288 / C, HW)).
 * quad_mean != 0: the output holds the mean of every 4 consecutive activated channels (C/4 per pixel; decoder.py:140-143 DAP =
 * PixelShuffle(2) -> AvgPool2d(2), which commutes with the bilinear upsample that follows) -- ld_out / out_col then count those. */
/* Floats of mumpy_groupnorm_nhwc's caller-owned `stats_ws` for a (B, HW, C) map with `groups` groups. */
long mumpy_groupnorm_workspace_floats(int B, int HW, int C, int groups);
int mumpy_groupnorm_nhwc(const float *x, const float *gamma, const float *beta, float *stats_ws, float *out,
                         long ld_out, int out_col, int B, int HW, int C, int groups, float eps, int act, int quad_mean, void *stream);
/* out[..., col:col+Cout] = resample(in) (* mul) (+ add); mul/add are (B,Ho,Wo,Cout) fp32 contiguous or NULL.  out is fp32
 * (out_dtype MUMPY_F32) or, when the only consumer is a tensor-core convolution, directly its 16-bit operand (MUMPY_BF16 / MUMPY_F16:
 * same rounding as mumpy_cast16, one pass over the map less; needs C, ld_out, out_col multiples of 4 and a mode other than pixel
 * shuffle).  ld_out / out_col count elements of out. */
int mumpy_resample_nhwc(const float *in, const float *mul, const float *add, void *out, int out_dtype, long ld_out, int out_col,
                        int B, int H, int W, int C, int mode, int scale, void *stream);
/* out = a * b (+ c), elementwise over n fp32 values (decoder.py:204-221 gates and skips); out fp32 or a 16-bit operand type. */
int mumpy_mul_add(const float *a, const float *b, const float *c, void *out, int out_dtype, long n, void *stream);
/* out = a + b elementwise over n floats (the `shortcut + attn` of CrossSwinBlock, whose un-summed attention
 * branch is also an output: multiTemporalViewEncoder.py:275-276). */
int mumpy_add(const float *a, const float *b, float *out, long n, void *stream);
/* layout changes: NCHW (B,C,H,W) <-> NHWC; `pool2` averages 2x2 blocks first (AvgPool2d on ffinfo). */
int mumpy_nchw_to_nhwc(const float *in, float *out, long ld_out, int out_col, int B, int C, int H, int W, int pool2,
                       void *stream);
int mumpy_nhwc_to_nchw(const float *in, long ld_in, float *out, int B, int C, int H, int W, void *stream);
/* DAP == mean over groups of `k` consecutive channels (decoder.py:140-143). */
int mumpy_channel_group_mean(const float *in, float *out, long pixels, int C, int k, void *stream);

/* Clip assembly on the device (test.py:22-25 ToTensor + Normalize; universaldataloader.py:45-48 sliding clip window):
 * frames (n_frames,H,W,3) uint8 HWC device images; clip_frames (B,T) int32 device frame indices (edge-clamped by the
 * caller, mumpy_b200.frontend.clip_frame_indices); out (B,T,3,H,W) fp32 = (frame/255 - mean[c]) / std[c], bit-identical to
 * torchvision's transform.  mean3/std3 are HOST arrays of 3 floats (read during the call). */
int mumpy_assemble_clips(const unsigned char *frames, const int *clip_frames, float *out, int B, int T, int H, int W,
                         const float *mean3, const float *std3, void *stream);

/* Frame resize of the loader on the device (dataloaders/universaldataset.py:68-79: PIL `img.resize(self.inputRes)`), bit-identical
 * to Pillow's 8-bit resampler.  filter: MUMPY_RESIZE_BICUBIC (Pillow >= 7 default; two separable passes, 22-bit fixed-point
 * coefficients, uint8 intermediate) or MUMPY_RESIZE_NEAREST (default of the pillow==4.0.0 pinned by requirements.txt:9).
 *
 * mumpy_resize_taps (HOST only, no GPU needed): per-axis tables exactly as Pillow builds them in double precision.
 *   bicubic: bounds = out_size x [first source index, tap count], coefs = out_size x ksize int32 (22 fractional bits);
 *            call with coefs = NULL to get *ksize only.   nearest: bounds = out_size source indices, coefs unused, *ksize = 1.
 * mumpy_resize_u8: in (n,in_h,in_w,C) uint8 -> out (n,out_h,out_w,C) uint8, C <= 4; the tables are DEVICE copies of the above
 *   (_h for the width axis, _v for the height axis); tmp = n*in_h*out_w*C bytes, needed when both sizes change. */
#define MUMPY_RESIZE_NEAREST 0
#define MUMPY_RESIZE_BICUBIC 3
int mumpy_resize_taps(int in_size, int out_size, int filter, int *bounds, int *coefs, int coef_capacity, int *ksize);
int mumpy_resize_u8(const unsigned char *in, unsigned char *out, unsigned char *tmp, int n, int in_h, int in_w, int out_h, int out_w,
                    int channels, int filter, const int *bounds_h, const int *coefs_h, int ksize_h, const int *bounds_v,
                    const int *coefs_v, int ksize_v, void *stream);

/* a20 + measure.py:77-91: thresholded mask (logit > 0 -> 255) and per-clip integer counts
 * [TP, n_pred, n_gt, n_union] (int64, accumulated with atomics; counts must be zeroed by the caller).
 * logits (B, HW) fp32; gt (B, HW) uint8 (non-zero = positive) or NULL; mask (B, HW) uint8 or NULL. */
int mumpy_mask_counts(const float *logits, const unsigned char *gt, unsigned char *mask, long long *counts, int B,
                      int HW, void *stream);

/* fp32 -> bf16 / f16 cast (weight packing, operand hand-over). */
int mumpy_cast16(const float *in, void *out, int out_dtype, long n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MUMPY_B200_H */
