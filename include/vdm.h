/*
 * libvdm.so -- C ABI of the B200-native video-diffusion hot path.
 *
 * The reference (cliangyu/video-diffusion) is pure Python/PyTorch and has no FFI layer;
 * its "interface" for this path is the set of ATen ops issued by
 *   improved_diffusion/unet.py            (CondMargVideoModel / UNetVideoModel forward)
 *   improved_diffusion/gaussian_diffusion.py (p_mean_variance, p_sample, ddim_sample, _vb_terms_bpd)
 * Each entry point below replaces the group of reference ops cited beside it.  The Python
 * host mirror (the video_diffusion_b200 package) binds them with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless named host_*;
 *   - the caller owns all memory (the library never allocates device memory, never frees,
 *     never synchronises); every call is asynchronous on `stream` and CUDA-graph capturable;
 *   - returns 0 on success, <0 for invalid arguments / unsupported shapes, >0 = cudaError_t;
 *     vdm_last_error_string() gives a thread-local message;
 *   - activations are channels-last: row m = (image n, y, x), n = b*F + f, C contiguous.
 */
#ifndef VDM_H_
#define VDM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vdm_stream_t; /* cudaStream_t */

enum { VDM_F32 = 0, VDM_BF16 = 1, VDM_F64 = 2, VDM_I64 = 3, VDM_F16 = 4 };

int vdm_version(void);
const char* vdm_last_error_string(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t vdm_launch_count(void);

/* Diagnostics: when buf (device, >= 8 * grid u64) is non-NULL every later bf16 vdm_gemm launch writes per-CTA cycle
 * counters [producer wait-empty, MMA wait-accumulator, MMA wait-operands, MMA total, epilogue wait, epilogue total].
 * NULL switches it off (default). */
void vdm_gemm_set_trace(void* buf);

/* ---- implicit-GEMM convolution / linear ------------------------------------------------
 * Replaces nn.Conv2d 3x3 / 1x1 (unet.py:61, 90-95, 141, 156-160, 172-173, 618, 748) and
 * nn.Linear (unet.py:143-150, 418-419, 606-610, 274-277) together with the ops fused around
 * them: bias, residual/skip add (unet.py:198, 538), emb add (unet.py:196), concat-free
 * second K range for the 1x1 skip projection, SiLU on a second output.
 *
 *   out[m][n] = sum_{tap,c} A1[pix(m)+tap][c] * W[n][tap*C1+c]  + sum_c A2[m][c] * W[n][taps*C1+c]
 *               + bias[n] + rowbias[img(m)][n] + residual[m][n]
 *
 * dtype VDM_BF16: A1/A2/W are bf16, the kernel is the tcgen05/TMEM/TMA one (fp32 accumulate).
 * dtype VDM_F32 : A1/A2/W are fp32, SIMT fp32 kernel (reference-accuracy mode).
 * dtype VDM_F16 : plain linears only (taps 1, no A2): A1 and W are IEEE half, same tcgen05 kernels with the f16 operand
 *                 format -- the normalised fp16 stream itself is the operand of the attention qkv projection.
 */
typedef struct {
  int32_t dtype;        /* VDM_F32 | VDM_BF16 | VDM_F16 (plain linears): type of a1, a2, w */
  int32_t taps;         /* 1 (linear / 1x1), 9 (3x3, pad 1) or 4 (a1_mode 3) */
  int32_t a1_mode;      /* 0: A1 is [n][H][W][C1] at output resolution (stride 1)
                           1: stride-2 conv.  bf16: A1 is parity planes [n][py][px][H][W][C1] of the
                              2H x 2W input; fp32: A1 is the raw [n][2H][2W][C1] input
                           2: nearest-x2 upsample folded in (fp32 only): A1 is [n][H/2][W/2][C1]
                           3: nearest-x2 upsample folded into four 2x2 parity convolutions (bf16 only):
                              A1 is [n][H/2][W/2][C1], taps = 4, w is [4][N][4*C1] (see fold in unet.py) */
  int32_t n_img, H, W;  /* output geometry; M = n_img*H*W.  For linear: n_img=M, H=W=1 */
  int32_t C1, C2;       /* channels of A1 / A2 (C2 = 0: no second range) */
  int32_t N;            /* output channels */
  const void* a1;
  const void* a2;       /* [M][C2] or NULL */
  const void* w;        /* [N][taps*C1 + C2], K contiguous */
  const float* bias;    /* [N] or NULL */
  const float* rowbias; /* [n_img][ld_rowbias] or NULL (per-image channel bias) */
  int32_t ld_rowbias;
  const float* residual; /* [M][ld_res] or NULL */
  int32_t ld_res;
  float* out_f32;       /* [M][ld_out] or NULL; with out_nchw: [n_img][N][H][W] */
  void* out_bf16;       /* [M][ld_out_bf16] or NULL */
  int32_t ld_out, ld_out_bf16;
  int32_t out_nchw;
  float* out_silu_f32;  /* optional second output silu(out) [M][ld_out] (fp32 kernel only) */
  int32_t lda1;         /* taps == 1 only: row stride of A1 in elements (0 = C1), so a column block of a wider
                           matrix (e.g. the q or k third of qkv) can be the operand */
  int32_t w_group_tiles; /* bf16 kernel, taps == 1: 0 = one weight matrix; g > 0 = grouped weights, w is
                           [groups][N][K] and the 128-row tile m uses group m / g (RPE tables differ per
                           (batch, frame) block of rows) */
  int64_t* stats_out;   /* optional (bf16 kernel only, H*W % 32 == 0): per-image, per-output-channel sums of the
                           stored values, [n_img][2][N] (plane 0: sum, plane 1: sum of squares) as 64-bit
                           fixed point (units of 2^-24), accumulated with integer atomics -- deterministic --
                           into a buffer the caller zeroed: the GroupNorm statistics of the NEXT layer come
                           out of this epilogue for free (nn.py:15-17) */
  int32_t n_prob;       /* bf16 kernel, taps == 1: > 1 runs that many same-shape plain linears in ONE launch: problem i
                           reads the A1 column block starting at i * prob_a_cols (lda1 spans them all), the weight
                           rows starting at i * prob_w_rows and writes at out + i * prob_out_stride elements.  Used for
                           the q -> Sk and k -> Sq RPE GEMMs of a temporal attention block (unet.py:357-366) */
  int32_t prob_a_cols;
  int64_t prob_w_rows;
  int64_t prob_out_stride;
  const float* a1_coef; /* bf16 kernel, taps == 9 stride 1: GroupNorm-apply (+ scale/shift) + SiLU fused into the A1
                           operand path (ResBlock `in_layers` / `out_layers`, unet.py:185-198, nn.py:10-17).  [n_img][C1]
                           pairs (a, b) from vdm_gn_coef: a1 then holds the RAW (un-normalised) bf16 activation and the
                           kernel's transform warps rewrite every activation tile in shared memory, between the TMA
                           load and the MMA, to act(a * x + b); conv padding stays zero.  NULL = A1 used as is.
                           Only the halo kernels take it: vdm_gemm_fused_norm_supported() tells; other shapes are an
                           error, never a silent slow path */
  int32_t a1_act;       /* with a1_coef: 1 = SiLU after the affine, 0 = affine only */
  const void* a2b;      /* bf16 kernel: optional SECOND tensor of the A2 range, [M][C2b], concatenated after a2 along
                           channels (the U-Net skip concat feeding the 1x1 skip projection, unet.py:826-828, 172-173):
                           w then is [N][taps*C1 + C2 + C2b] */
  int32_t C2b;
  int32_t a2_dtype;     /* bf16 kernel: VDM_BF16 (0 / default) or VDM_F16 = a2 / a2b AND the weight columns of that range
                           hold IEEE half.  The MMAs of the range then run with the f16 operand format, so the fp16
                           residual stream itself is the operand of the 1x1 skip projection -- no bf16 copy of it is ever
                           written -- and, with identity weight columns, of the residual add (`x + h`, unet.py:198) */
  int32_t io_dtype;     /* bf16 kernel: VDM_F32 (default) or VDM_F16 = `out_f32` and `residual` point to IEEE half tensors
                           (same leading dimensions, in elements).  The residual stream of the bf16 model is kept in fp16
                           (11-bit mantissa, conversions saturate): half the HBM traffic of every stream read / write;
                           accumulation, bias / residual adds and the GroupNorm statistics stay fp32 */
  uint32_t* img_done;   /* bf16 kernel, optional: per-image completion counters [n_img] (caller zeroes them).  Every epilogue
                           warp adds (rows x channels it stored) once those stores AND its statistics are visible device-wide,
                           so image n of the output -- and its GroupNorm statistics -- is complete when img_done[n] ==
                           H*W*N.  The kernel also releases its programmatic dependents at its very top (all its CTAs are
                           resident then), so a vdm_gn_apply launched behind it with `wait_done` runs BESIDE it on the SMs'
                           spare registers and normalises each image as soon as the convolution has finished it, while
                           the data is still in L2.  Only the transposed-role conv kernels take it
                           (vdm_gemm_img_done_supported); elsewhere it is an error */
  int32_t a1_raw_dtype; /* output head only (out_nchw, N <= 8, with a1_coef): VDM_F16 = a1 is the RAW fp16 residual stream
                           and GroupNorm-apply + SiLU (unet.py:745-748) happen while the conv stages its tile; 0 otherwise */
} vdm_gemm_args;

int vdm_gemm(const vdm_gemm_args* args, vdm_stream_t stream);
/* 1 if vdm_gemm would run `args` (with a1_coef set) on a kernel that has the fused-normalisation transform stage */
int vdm_gemm_fused_norm_supported(const vdm_gemm_args* args);
/* 1 if vdm_gemm would run `args` on a kernel that maintains `img_done` */
int vdm_gemm_img_done_supported(const vdm_gemm_args* args);

/* ---- GroupNorm32 (+SiLU, + scale/shift), producer of GEMM A operands --------------------
 * Replaces GroupNorm32 (nn.py:15-17) + SiLU (nn.py:10-12) + `h*(1+scale)+shift`
 * (unet.py:190-194) + th.cat of the U-Net skip (unet.py:826-828) + F.interpolate x2
 * (unet.py:63-69) + the dtype cast feeding the next conv.
 * Statistics are kept per (image, channel): [n_img][2][C] (plane 0: sum over H*W, plane 1: sum of
 * squares); the apply kernel folds channels into the 32 groups, so a skip tensor can be re-grouped by
 * whichever concat consumes it.  vdm_gn_stats fills a double-precision table (caller zeroes it); the
 * bf16 GEMM epilogue fills a 64-bit fixed-point one (vdm_gemm_args.stats_out). */
int vdm_gn_stats(const float* src, int32_t C, int32_t n_img, int32_t HW, double* stats, vdm_stream_t stream);
/* the same for an fp16 (VDM_F16) or fp32 source */
int vdm_gn_stats_t(const void* src, int32_t src_dtype, int32_t C, int32_t n_img, int32_t HW, double* stats,
                   vdm_stream_t stream);

typedef struct {
  const void* src1; int32_t C1;       /* [n_img*HW][C1], fp32 (or bf16 when src1_dtype == VDM_BF16 and C2 == 0) */
  const float* src2; int32_t C2;      /* optional second source, concatenated along C */
  int32_t n_img, H, W;
  const void* stats1;                 /* [n_img][2][C1] of src1; NULL: no normalisation (plain cast / concat) */
  const void* stats2;                 /* [n_img][2][C2] of src2 */
  int32_t stats_dtype;                /* of stats1: VDM_I64 (GEMM epilogue, fixed point) | VDM_F64 (vdm_gn_stats) */
  int32_t stats2_dtype;               /* of stats2 */
  const float* gamma; const float* beta;   /* [C1+C2] */
  const float* scale_shift;           /* optional [n_img][ld_ss]: scale = [0,C), shift = [C,2C) */
  int32_t ld_ss;
  int32_t silu;
  int32_t out_mode;                   /* 0 plain, 1 nearest-x2 upsampled, 2 stride-2 parity planes */
  int32_t out_dtype;                  /* VDM_F32 | VDM_BF16 */
  void* out;                          /* GEMM A operand; may be NULL with out_f32_copy (plain layout): only the copy is
                                         written -- in fp16 it is itself the operand of the qkv projection (VDM_F16) */
  int32_t src1_dtype;                 /* VDM_F32 | VDM_F16 (src2 alike) | VDM_BF16 (single source) */
  void* out_raw;                      /* optional second output: the un-normalised input cast to out_dtype, plain
                                         layout (A operand of the 1x1 skip projection, unet.py:172-173) */
  float* out_f32_copy;                /* optional fp32 copy of the plain output (attention residual) */
  int32_t copy_dtype;                 /* of out_f32_copy: VDM_F32 (default) | VDM_F16 */
  const uint32_t* wait_done;          /* optional: vdm_gemm_args.img_done of the kernel launched just before on the same
                                         stream, which produces src1 and stats1.  The launch then is a programmatic
                                         dependent of that kernel and every block waits for wait_done[image] >= wait_count
                                         (= H*W*C1) instead of for the whole producer grid */
  uint32_t wait_count;
} vdm_gn_apply_args;

int vdm_gn_apply(const vdm_gn_apply_args* args, vdm_stream_t stream);

/* Per-(image, channel) affine form of GroupNorm32 (+ scale/shift): coef[n][c] = (a, b) with
 *   a = rstd_g * gamma_c * (1 + scale_nc),   b = (beta_c - mean_g * rstd_g * gamma_c) * (1 + scale_nc) + shift_nc
 * from the same statistics tables vdm_gn_apply takes (two sources = channel concat).  A tiny launch (n_img blocks); its
 * output feeds vdm_gemm_args.a1_coef, so the normalised activation is never written to memory. */
int vdm_gn_coef(const void* stats1, int32_t stats_dtype, int32_t C1, const void* stats2, int32_t stats2_dtype,
                int32_t C2, int32_t n_img, int32_t HW, const float* gamma, const float* beta,
                const float* scale_shift, int32_t ld_ss, float* coef, vdm_stream_t stream);

/* GroupNorm over (C/32 channels x T frames) per (b, pixel): the temporal-attention norm
 * (unet.py:473-474 on x.reshape(B*D, C, T)).  x: [B][T][HW][C] fp32. */
int vdm_gn_temporal(const float* x, int32_t B, int32_t T, int32_t HW, int32_t C,
                    const float* gamma, const float* beta, float* out_f32,
                    void* out_a, int32_t out_dtype, vdm_stream_t stream);
/* the same with x and the normalised residual copy `out_res` in io_dtype (VDM_F32 | VDM_F16: the fp16 stream);
 * out_a may be NULL when out_res is given (the fp16 copy then feeds the qkv projection directly, vdm_gemm dtype VDM_F16) */
int vdm_gn_temporal_t(const void* x, int32_t io_dtype, int32_t B, int32_t T, int32_t HW, int32_t C,
                      const float* gamma, const float* beta, void* out_res, void* out_a, int32_t out_dtype,
                      vdm_stream_t stream);

/* out[m][c] = (h[m][c] + enc[m % HW][c]) + frame_emb[m / HW][c]: the learned spatial_encoding add
 * (unet.py:841-844) and, with use_frame_encoding, the per-frame sinusoid (unet.py:914-926; frame_emb is
 * vdm_timestep_embedding of the frame indices with max_period = 10 T).  Either addend may be NULL, not both;
 * out may alias h. */
int vdm_add_spatial_encoding(const float* h, const float* enc, const float* frame_emb, float* out,
                             int32_t n_img, int32_t HW, int32_t C, vdm_stream_t stream);
/* the same with h / out in io_dtype (VDM_F32 | VDM_F16) */
int vdm_add_spatial_encoding_t(const void* h, int32_t io_dtype, const float* enc, const float* frame_emb, void* out,
                               int32_t n_img, int32_t HW, int32_t C, vdm_stream_t stream);

/* ---- conditioning mix + input-conv im2col -----------------------------------------------
 * Replaces CondMargVideoModel.forward's masking / indicator channels / per-frame timesteps (unet.py:951-1019).
 * x, x0: [B][F][3][H][W] fp32 (reference layout; x0 = whichever tensor observed_frames selects); masks: [B][F] fp32.
 * mode follows cond_emb_type: 0 'channel' (5 input channels), 1 'duplicate' / 'all' (6), 2 't=0' (3, x unchanged).
 * a_out: im2col rows [B*F*H*W][64] (K index = tap*Cin + c, zero padded), dtype out_dtype.
 * t_frame[B*F] = t[b] * (1 - obs[b][f]) in mode 0 (observed_frames='x_0'), t[b] otherwise;
 * attn_mask[B*F] = min(obs+lat+kinda, 1). */
int vdm_cond_mix(const float* x, const float* x0, const float* obs_mask, const float* latent_mask,
                 const float* kinda_marg_mask, const float* t, int32_t B, int32_t F, int32_t H,
                 int32_t W, int32_t mode, void* a_out, int32_t out_dtype, float* t_frame, float* attn_mask,
                 vdm_stream_t stream);

/* `_WrappedModel.__call__` (respace.py:113-119) as one launch: out[i] = (float)timestep_map[clamp(t[i], 0, n_map-1)] * scale
 * (scale = 1000 / original_num_steps when rescale_timesteps, else 1).  An index outside the map is clamped here and
 * reported by the sampler kernels that receive the same t (vdm_sampler_error). */
int vdm_map_timesteps(const int64_t* t, const int64_t* timestep_map, int32_t n_map, float scale, float* out, int32_t B,
                      vdm_stream_t stream);

/* The inputs of one forward (the caller's x, x0, per-frame masks, per-video timesteps, frame indices) into the
 * address-stable workspace tensors the captured CUDA graph reads -- ONE launch instead of seven device-to-device copies
 * around every graph replay (the host side of UNetModel.forward's argument handling, unet.py:949-1013).  All tensors
 * contiguous; x / x0: `elems` fp32 each (a multiple of 4); masks [B*F] fp32; t [B] fp32; frame_indices [B*F] int64 or NULL. */
int vdm_stage_inputs(const float* x, const float* x0, const float* obs_mask, const float* latent_mask,
                     const float* kinda_marg_mask, const float* t, const int64_t* frame_indices, int32_t B, int32_t F,
                     int64_t elems, float* ws_x, float* ws_x0, float* ws_obs, float* ws_lat, float* ws_kinda, float* ws_t,
                     int64_t* ws_fi, vdm_stream_t stream);

/* sinusoidal embedding (nn.py:89-107, 110-122): out[n] = [cos(t*f_i) | sin(t*f_i)], f_i = exp(-ln(max_period) i / half);
 * max_period = 10000 for diffusion time, 10 T for frame indices */
int vdm_timestep_embedding(const float* t_frame, int32_t n, int32_t dim, double max_period, float* out,
                           vdm_stream_t stream);

/* ---- RPE net hidden layer (unet.py:283-296) ----------------------------------------------
 * hidden[net][(b*T+i)*T+j][c] = silu(e_t[b*T+i][net*C + c] + Wd[net][c][:]·feat(d_ij) + bd[net][c])
 * with d_ij = fi[b][i]-fi[b][j], feat = (log1p(max(d,0)), log1p(max(-d,0)), d==0).
 * e_t: [B*T][ld_et] holds embed_diffusion_time(temb) (+ its bias) of the 3 nets (q,k,v). */
int vdm_rpe_hidden(const float* e_t, int32_t ld_et,
                   const int32_t* et_offsets /* optional [n_blocks]: first e_t column of each block's (q,k,v) triple */,
                   int32_t n_blocks /* attention blocks batched in this call; wd, bd, out hold 3*n_blocks nets */,
                   const int64_t* frame_indices, const float* wd, const float* bd, int32_t B, int32_t T,
                   int32_t C, void* out, int32_t out_dtype, vdm_stream_t stream);

/* ---- head-averaged attention maps (logging only; return_attn_weights=True, unet.py:464-468) ----
 * out[g][i][j] = | mean over heads of softmax_j(logits[i][j]) |, g = outer * n_inner + inner, [n_outer*n_inner][L][L].
 * Sequence element l of g is the qkv row at qkv + outer*outer_stride + inner*inner_stride + l*seq_stride (strides in
 * elements; row columns (3, heads, hd)).  Temporal attention: outer = b, inner = pixel, r_q / r_k: [B*L*L][C] RPE
 * tables and mask [B][L] as for vdm_attn_temporal.  Spatial attention: outer = image, n_inner = 1, NULL tables / mask. */
int vdm_attn_weights_mean(const void* qkv, int32_t qkv_dtype, int64_t n_outer, int64_t n_inner,
                          int64_t outer_stride, int64_t inner_stride, int64_t seq_stride, int32_t L,
                          int32_t heads, int32_t hd, const float* r_q, const float* r_k, const float* mask,
                          int32_t allow_pad_interactions, float* out, vdm_stream_t stream);

/* ---- temporal attention with in-kernel RPE bias (unet.py:477-536) -------------------------
 * qkv: [B*T*HW][3C] fp32 (row (b,t,pix); columns (3, heads, hd)); R_q/R_k/R_v: [B*T*T][C] fp32
 * (row (b,i,j), columns (heads, hd)); mask: [B][T] fp32 (1 = real frame).
 * out_a: [B*T*HW][C] attention output as GEMM A operand for proj_out. */
int vdm_attn_temporal(const float* qkv, const float* r_q, const float* r_k, const float* r_v,
                      const float* mask, int32_t allow_pad_interactions, int32_t B, int32_t T,
                      int32_t HW, int32_t heads, int32_t hd, void* out_a, int32_t out_dtype,
                      vdm_stream_t stream);

/* Lookup-table RPE, use_rpe_net=False (replaces RPE.get_bucket_ids + lookup_table_weight[bucket_ids],
 * unet.py:326-347).  tables [3][n_buckets][C] (q, k, v nets; n_buckets = 2*beta+1), frame_indices [B][T] int64;
 * out [3][B*T*T][C] fp32 = the three R tables the attention kernels / vdm_rpe_expand consume. */
int vdm_rpe_lookup(const float* tables, const int64_t* frame_indices, int32_t B, int32_t T, int32_t C,
                   int32_t n_buckets, double alpha, double beta, double gamma, float* out, vdm_stream_t stream);

/* ---- temporal attention, tensor-core path (bf16 mode) ------------------------------------------
 * The RPE einsums contract against tables that depend on (batch, frame) but not on the pixel, so over
 * the pixels of one (b, t) they are GEMMs: vdm_rpe_expand builds block-diagonal-over-heads bf16 weight
 * blocks from R_q / R_k / R_v ([B*T*T][C] fp32), one per group g = b*T + t, `groups_per_tile` groups
 * per 128-row tile (1 if HW >= 128, 128/HW otherwise):
 *   bk, bq: [ceil(G/gpt)][128*gpt][C]   row sub*128 + h*T + j, column h*hd + f   (bq pre-scaled by hd^-0.5)
 *   bv    : [ceil(G/gpt)][C][128*gpt]
 * The host then runs vdm_gemm with w_group_tiles on q -> Sk and k -> Sq ([M][128*gpt] fp32), calls
 * vdm_attn_temporal_tc (scale*q.k^T + scale*Sk + Sq^T, mask, fp32 softmax, P.V on mma.sync; writes
 * P as bf16 [M][128*gpt] into a buffer whose padding columns the caller zeroed once, and PV fp32
 * [M][C]) and a last grouped vdm_gemm (P x bv, residual PV) for the attn.R_v term (unet.py:357-378). */
int vdm_rpe_expand(const float* r_q, const float* r_k, const float* r_v,
                   const float* bias /* optional [n_blocks][3][C] (q, k, v) added to the tables */,
                   int32_t n_blocks /* attention blocks batched in this call */,
                   int64_t r_block_stride /* elements between consecutive blocks' tables */,
                   int64_t qk_block_stride /* elements between consecutive blocks' bq (and bk) operands; 0 = packed */,
                   int32_t B, int32_t T,
                   int32_t heads, int32_t hd, int32_t groups_per_tile,
                   int32_t zero_fill /* 1: write every element; 0: the caller zeroed bq/bk/bv once and only this
                                        function (same shapes) writes them, so the structural zeros are skipped */,
                   void* bq, void* bk, void* bv, vdm_stream_t stream);
int vdm_attn_temporal_tc(const void* qkv, const float* sk, const float* sq, const float* mask,
                         int32_t allow_pad_interactions, int32_t B, int32_t T, int32_t HW,
                         int32_t heads, int32_t hd, int32_t groups_per_tile, void* pm, float* pv,
                         vdm_stream_t stream);

/* ---- temporal attention, fused block (bf16 mode, default) ------------------------------------------
 * Replaces RPEAttention._forward between the qkv projection and proj_out (unet.py:471-536) together with
 * RPE.forward_qk / forward_v (unet.py:357-378) in ONE kernel: per CTA (video, head, `pixels_per_cta` pixels) the
 * RPE score terms q.Rk and k.Rq (GEMMs over the pixels of the tile), q.k^T, mask, fp32 softmax, P.V and the
 * attn.R_v term, all on mma.sync with the intermediates in shared memory.  out: [B*T*HW][C] bf16.
 * vdm_rpe_pack converts the fp32 R tables ([B*T*T][C] each, + optional bias [n_blocks][3][C] (q, k, v)) into the
 * kernel's bf16 operands, for `n_blocks` attention blocks per call.  In the RPE products the table is the A operand of
 * mma.m16n8k16 and the pixels are the n = 8 dimension, so the tables are stored FRAGMENT-MAJOR: one 16-byte vector =
 * the four A registers of one lane for one k-step (lane = 4 g + t4; register r: row 16 mt + g + 8 (r & 1), contraction
 * elements 8 t4 + 4 kk2 + 2 (r >> 1) + {0, 1} -- the same permutation the kernel applies to its shared-memory operand):
 *   rq, rk: [n_blocks][B*T][heads][2][hd/32][2][32] vectors   rows = second frame index j (zero for j >= T),
 *                                                              contraction over the channels 32 p + c of the head
 *   rv    : [n_blocks][B*T][heads][hd/16][2][32] vectors      rows = channels, contraction over the second frame index
 * i.e. B*T*heads*hd*32 elements per table and block.  T <= 32, hd a multiple of 32.  The kernel: t_pad = 24 (T <= 24)
 * or 32 selects the key-tile count of the q.k^T phase; HW a multiple of pixels_per_cta (8, or 16 for hd = 96; 0 = 8). */
int vdm_rpe_pack(const float* r_q, const float* r_k, const float* r_v, const float* bias, int32_t n_blocks,
                 int64_t r_block_stride /* elements between consecutive blocks' fp32 tables */, int32_t B, int32_t T,
                 int32_t heads, int32_t hd, void* rq, void* rk, void* rv,
                 int64_t qk_block_stride /* elements between blocks' rq (and rk); 0 = packed */,
                 int64_t v_block_stride /* 0 = packed */, vdm_stream_t stream);
/* Dynamic shared memory (bytes) the fused kernel needs for this shape, or -1 if the shape is not instantiated; the
 * launch fails above 227 KB (the host mirror falls back to the three-launch path there). */
int64_t vdm_attn_temporal_fused_smem(int32_t T, int32_t hd, int32_t t_pad, int32_t pixels_per_cta);
/* Diagnostics: buf = 8 device uint64 counters (NULL = off).  Thread 0 of every CTA adds the cycles it spent in the six
 * phases of the kernel (staging wait, P1a, P1b, P2a, P2b, P3) to buf[0..5] and 1 to buf[7]. */
void vdm_attn_temporal_fused_set_trace(void* buf);
int vdm_attn_temporal_fused(const void* qkv, const void* rq, const void* rk, const void* rv, const float* mask,
                            int32_t allow_pad_interactions, int32_t B, int32_t T, int32_t HW, int32_t heads,
                            int32_t hd, int32_t t_pad, int32_t pixels_per_cta, void* out, vdm_stream_t stream);

/* ---- spatial attention (unet.py:258-266 -> 477-536 without RPE / mask) --------------------
 * qkv: [n_img*L][3C] (dtype qkv_dtype); softmax(q k^T / sqrt(hd)) v per (image, head).
 * qkv_dtype VDM_BF16 -> tensor-core flash kernel; VDM_F32 -> fp32 SIMT kernel. */
int vdm_attn_spatial(const void* qkv, int32_t qkv_dtype, int32_t n_img, int32_t L, int32_t heads,
                     int32_t hd, void* out_a, int32_t out_dtype, vdm_stream_t stream);

/* ---- sampler step (gaussian_diffusion.py:326-343, 374-382, 208-227, 438-443, 597-634) ----
 * tables: [VDM_TAB_COUNT][n_steps] fp32 (see enum), t: [B] int64 indices into the tables.
 * mode 0: ancestral p_sample; mode 1: DDIM with eta; mode 2: DDIM reverse ODE step x_t -> x_{t+1}
 * (:636-668, eta must be 0, noise unused).  Elementwise over B*per_batch.
 * A timestep outside [0, n_steps) (the reference's numpy table lookup raises IndexError there) is clamped,
 * so nothing is read out of bounds, and recorded: vdm_sampler_error() returns 1 once and clears the record. */
enum {
  VDM_TAB_SQRT_RECIP_ACP = 0, VDM_TAB_SQRT_RECIPM1_ACP, VDM_TAB_POST_C1, VDM_TAB_POST_C2,
  VDM_TAB_MODEL_LOGVAR, VDM_TAB_MODEL_VAR, VDM_TAB_ACP, VDM_TAB_ACP_PREV, VDM_TAB_POST_LOGVAR,
  VDM_TAB_SQRT_ACP, VDM_TAB_SQRT_1M_ACP, VDM_TAB_LOG_1M_ACP,
  VDM_TAB_POST_VAR,        /* posterior_variance (gaussian_diffusion.py:161-162) */
  VDM_TAB_RECIP_POST_C1,   /* 1 / posterior_mean_coef1 (:386) */
  VDM_TAB_POST_C2_DIV_C1,  /* posterior_mean_coef2 / posterior_mean_coef1 (:387-389) */
  VDM_TAB_ACP_NEXT,        /* alphas_cumprod_next (:149, DDIM reverse ODE :658) */
  VDM_TAB_COUNT
};

int vdm_sampler_error(void);
int vdm_sampler_step(int32_t mode, const float* x, const float* eps, const float* noise,
                     const int64_t* t, const float* tables, int32_t n_steps, int32_t B,
                     int64_t per_batch, int32_t clip_denoised, float eta, float* sample,
                     float* pred_xstart, float* mean, vdm_stream_t stream);

/* q_sample (gaussian_diffusion.py:190-206): out = sqrt_acp[t]*x0 + sqrt_1m_acp[t]*noise */
int vdm_q_sample(const float* x0, const float* noise, const int64_t* t, const float* tables,
                 int32_t n_steps, int32_t B, int64_t per_batch, float* out, vdm_stream_t stream);

/* Per-row linear combinations of the posterior / prediction helpers (gaussian_diffusion.py:148-157,
 * 208-227, 374-396), coefficients gathered from table rows by t[b]:
 *   op 0: out = tab[row_a]*a + tab[row_b]*b      q_posterior_mean_variance (POST_C1, POST_C2)
 *   op 1: out = tab[row_a]*a - tab[row_b]*b      _predict_xstart_from_eps / _from_xprev
 *   op 2: out = (tab[row_a]*a - b) / tab[row_b]  _predict_eps_from_xstart
 *   op 3: out = tab[row_a]*a                     q_mean_variance (b unused) */
int vdm_lincomb(int32_t op, const float* a, const float* b, const int64_t* t, const float* tables,
                int32_t n_steps, int32_t row_a, int32_t row_b, int32_t B, int64_t per_batch, float* out,
                vdm_stream_t stream);

/* ---- ELBO terms (gaussian_diffusion.py:750-788, 970-988; losses.py:12-70; nn.py:73-77) ----
 * acc[b][0..2] += (vb term in bits, xstart_mse, eps_mse), each already divided by the FULL
 * per-batch element count (masked sum / full count).  acc zeroed by the caller.
 * latent_mask: [B][F]; per_frame = per_batch / F. */
int vdm_vb_terms(const float* x0, const float* x_t, const float* eps, const float* noise,
                 const int64_t* t, const float* tables, int32_t n_steps, const float* latent_mask,
                 int32_t B, int32_t F, int64_t per_frame, int32_t clip_denoised, double* acc,
                 vdm_stream_t stream);

/* prior_bpd (gaussian_diffusion.py:909-926): acc[b] += KL(q(x_T|x0) || N(0,1)) in bits */
int vdm_prior_bpd(const float* x0, const float* tables, int32_t n_steps, const float* latent_mask,
                  int32_t B, int32_t F, int64_t per_frame, double* acc, vdm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VDM_H_ */
