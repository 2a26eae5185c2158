/*
 * stedm_b200 — C ABI of the sm_100a kernel library behind STEDM's sampling path.
 *
 * The reference (OettlM/STEDM) is 100 % Python and has no FFI of its own; this header is the
 * boundary a maintainer binds with ctypes.  Each entry point names the reference call
 * site(s) it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch allocates; the library never frees or
 *     retains memory, never allocates, never synchronises) unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it and is
 *     CUDA-graph capturable;
 *   - return value 0 = enqueued; negative = error (STEDM_ERR_*), message via stedm_last_error() (thread local);
 *   - activations inside the U-Net / VAE are NHWC ("channels last"), dtype tag STEDM_F32 or STEDM_BF16;
 *     tensors crossing the reference-facing Python API are NCHW fp32, exactly as in the reference.
 */
#ifndef STEDM_B200_H
#define STEDM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STEDM_ABI_VERSION 3

#define STEDM_F32 0
#define STEDM_BF16 1

#define STEDM_ACT_NONE 0
#define STEDM_ACT_GELU 1

#define STEDM_ERR_ARG (-1)
#define STEDM_ERR_CUDA (-2)
#define STEDM_ERR_UNSUPPORTED (-3)

int stedm_abi_version(void);
const char* stedm_last_error(void);
/* 1 if the current device is compute capability 10.x (tcgen05/TMEM available), else 0. */
int stedm_device_supported(void);

/* ----------------------------------------------------------------------------------------------------
 * K11  Classifier-free-guidance combine + std-rescale + DDIM update, one kernel.
 * Replaces DDIMSampler.p_sample_ddim's elementwise tail, ldm/models/diffusion/ddim.py:177-209:
 *   e_w = e_u + w (e_c - e_u);  e = phi * e_w * std_{C,H}(e_c)/std_{C,H}(e_w) + (1-phi) e_c   (guided != 0)
 *   pred_x0 = (x - sqrt(1-a_t) e)/sqrt(a_t);  x_prev = sqrt(a_prev) pred_x0 + sqrt(1-a_prev-sigma^2) e + sigma*noise
 * The std is unbiased and taken over dims (1,2) = channels and height only (shape (B,1,1,W)) as the reference does.
 * All tensors NCHW fp32 (B,C,H,W); e_u ignored when guided == 0; noise may be NULL (sigma*noise skipped).
 */
int stedm_cfg_ddim_step(const float* e_c, const float* e_u, const float* x, const float* noise, float* x_prev,
                        float* pred_x0, int batch, int channels, int height, int width, int guided, float cfg_scale,
                        float phi, float a_t, float a_prev, float sigma_t, float sqrt_one_minus_at, void* stream);

/* ----------------------------------------------------------------------------------------------------
 * K7  GroupNorm(32 groups) statistics and fused normalise + affine (+ SiLU) (+ channel concat).
 * Replaces GroupNorm32 + nn.SiLU (ldm/modules/diffusionmodules/util.py:199-216, openaimodel.py:214-216, 239-241, 729-731),
 * VAE Normalize + swish (ldm/modules/diffusionmodules/model.py:33-39) and the th.cat of skip tensors that
 * precedes them (openaimodel.py:800).
 * Input = channel concat [x0 (c0 ch) | x1 (c1 ch)] of NHWC tensors (x1 NULL when c1 == 0); x1 may have a smaller
 * batch x1_batch that is broadcast as b % x1_batch (shared encoder skips under batched guidance).
 * partials: double [batch][n_chunks][32][2] = per-chunk (sum, sum of squares), n_chunks = stedm_gn_num_chunks(hw,
 * c0 + c1) <= 128 (a function of the per-sample shape only).  No atomics and no zero-initialisation: the result
 * for a sample is bit-identical whatever batch it is launched in (rank-shard invariance).
 */
int stedm_gn_num_chunks(int hw, int channels);
int stedm_gn_stats(const void* x0, const void* x1, int in_dtype, int batch, int x1_batch, int hw, int c0, int c1,
                   double* partials, void* stream);
/* n_chunks: number of chunks in `partials` (0 => stedm_gn_num_chunks(hw, c0+c1), the layout stedm_gn_stats writes;
 * 1 => the folded layout stedm_gn_fold_tiles writes). */
int stedm_gn_apply(const void* x0, const void* x1, int in_dtype, int batch, int x1_batch, int hw, int c0, int c1,
                   const double* partials, int n_chunks, const float* gamma, const float* beta, float eps,
                   int apply_silu, void* out, int out_dtype, void* stream);
/* stedm_gn_apply for a concat whose second source is shared by the cond / uncond halves of a guided batch (x1 has
 * x1_batch < batch samples, broadcast as b % x1_batch; bf16 in, bf16 out).  Channels [split_c, c0 + c1) must lie in groups
 * made of x1 channels only: their normalised values repeat with period x1_batch and are written once per distinct sample
 * to out_hi [x1_batch][hw][c0 + c1 - split_c]; channels [0, split_c) go to out_lo [batch][hw][split_c].  The consumer
 * convolution takes (out_lo, out_hi) as its two sources, the second one broadcast (stedm_conv_desc.x1_batch).  Element
 * for element the values stedm_gn_apply writes (openaimodel.py:800 th.cat + util.py:199-216). */
int stedm_gn_apply_split(const void* x0, const void* x1, int batch, int x1_batch, int hw, int c0, int c1,
                         const double* partials, int n_chunks, const float* gamma, const float* beta, float eps,
                         int apply_silu, int split_c, void* out_lo, void* out_hi, void* stream);
/* Statistics pass folded into the producing convolutions: reduce the per-(128-pixel tile, channel) sums written by
 * stedm_conv_tc (stedm_conv_desc.stats_out) for the one or two producers of a GroupNorm input into
 * out = double [batch][1][32][2].  Source s: fp32 [reps_s][rep_stride_s tiles][c_s][2]; sample b owns tile rows
 * (b % batch_s)*tps_s + j, j < tps_s, in each of the reps_s repetitions (4 for the sub-pixel upsample phases).
 * coef (optional, with gamma / beta / eps / hw = pixels per sample of the normalised tensor): fp32 [batch][c0+c1][2] =
 * per-(sample, channel) (scale, shift) with y = x*scale + shift, computed exactly as stedm_gn_apply does — for consumers
 * that normalise in their own operand path (stedm_conv_desc.gn_coef).  `out` may be NULL when only coef is wanted. */
int stedm_gn_fold_tiles(const float* tiles0, int c0, int reps0, long long rep_stride0, int tps0, int batch0,
                        const float* tiles1, int c1, int reps1, long long rep_stride1, int tps1, int batch1, int batch,
                        double* out, const float* gamma, const float* beta, float eps, int hw, float* coef,
                        void* stream);

/* Row broadcast + per-sample embedding add (+ GroupNorm tile statistics): out[m][n] = src[m % src_rows][n] + emb[(m / hw) *
 * emb_stride + n], src fp32 [src_rows][c], out bf16 / fp32 [rows_out][c], stats_out optional fp32 [rows_out / 128][c][2].
 * Replaces the second evaluation of ResBlockStyle.in_layers (openaimodel.py:291-297, 268-287) for the unconditional half of
 * a guided batch: the convolution output is the same for both halves, only the style embedding added to it differs. */
int stedm_rows_add_emb(const float* src, long long src_rows, const float* emb, int emb_stride, void* out, int out_dtype,
                       long long rows_out, int hw, int c, float* stats_out, void* stream);

/* ----------------------------------------------------------------------------------------------------
 * K1/K2/K3/K4/K9/K10  Convolution as implicit GEMM, M = B*Ho*Wo pixels, N = Cout, K = k*k*(c0+c1).
 * Replaces nn.Conv2d 3x3 / 1x1 and nn.Conv1d k=1 call sites: openaimodel.py:120, 164-166, 217, 243, 254, 326, 334,
 * 542, 732; model.py:47-51, 92-119, 156-175, 487, 529; autoencoder.py:43 — with the epilogue fusing bias, the
 * per-sample embedding add (openaimodel.py:278-287), and the residual / skip add (openaimodel.py:288, model.py:141).
 */
typedef struct stedm_conv_desc {
  const void* x0;      /* NHWC source 0 [batch, in_h, in_w, c0] */
  const void* x1;      /* NHWC source 1 [x1_batch, in_h, in_w, c1] or NULL: input is the concat [x0 | x1] */
  const void* weight;  /* tensor-core path: bf16 [cout][k*k*(c0+c1)] (tap-major, channel-minor);
                          SIMT path: fp32 [k*k*(c0+c1)][cout] */
  const float* bias;   /* [cout] or NULL */
  const float* emb;    /* per-sample additive vector emb[b*emb_stride + n], or NULL */
  const void* residual;/* NHWC [batch, out_h, out_w, cout] added in the epilogue, or NULL */
  void* out;           /* NHWC [batch, out_h, out_w, cout], or NCHW when out_nchw != 0 */
  float* stats_out;    /* tensor-core path, optional: fp32 [tiles][cout][2] = per-(128-pixel tile, channel) sum and sum
                          of squares of the values written (tiles = batch*in_h*in_w/128, x4 phases in tap_mode 1):
                          the GroupNorm statistics pass folded into the producer; consumed by stedm_gn_fold_tiles */
  void* workspace;     /* tensor-core path, optional: scratch for split-K (launches with few output tiles and a deep
                          K loop, i.e. small batches); size from stedm_conv_tc_workspace_bytes(); NULL => single pass */
  int64_t workspace_bytes;
  int32_t c0, c1;
  int32_t in_dtype;    /* dtype of x0/x1 */
  int32_t batch, in_h, in_w;
  int32_t x1_batch;    /* 0 => batch */
  int32_t ksize;       /* 1 or 3 (padding = ksize/2) */
  int32_t stride;      /* 1 or 2 */
  int32_t upsample;    /* 1 => nearest x2 upsample fused in front of the conv (F.interpolate + conv) */
  int32_t emb_stride;
  int32_t res_dtype;
  int32_t out_dtype;
  int32_t out_nchw;
  int32_t cout;        /* GEMM N = rows of `weight` (tensor-core path: padded to a multiple of 16) */
  int32_t cout_store;  /* NCHW output only: number of leading output channels actually stored (0 => cout) */
  int32_t tap_mode;    /* tensor-core path: 0 = dense k x k taps; 1 = one 2x2 sub-pixel phase of "nearest x2 upsample
                          then 3x3 conv" (K9 folded into the conv: 4 taps on the LOW-resolution input instead of 9 on
                          the upsampled one): weight = bf16 [cout][4*c0] with taps (a,b) reading (y+a-1+py, x+b-1+px),
                          `out` = NHWC [batch, 2*in_h, 2*in_w, cout] written at (2y+py, 2x+px) */
  int32_t phase;       /* tap_mode 1: py*2 + px */
  int32_t act;         /* activation applied to (acc + bias + emb) before the residual add: STEDM_ACT_NONE, or
                          STEDM_ACT_GELU = exact erf GELU (the Swin-V2 MLP, torchvision swin_transformer.py MLP/nn.GELU) */
  int32_t skip_c0, skip_c1;  /* tensor-core path: channels of the fused skip input (see skip_x0) */
  int32_t skip_x1_batch;     /* 0 => batch; otherwise skip_x1 is broadcast as b % skip_x1_batch */
  const void* skip_x0; /* tensor-core path, optional: ResBlock.skip_connection / ResnetBlock.nin_shortcut (the 1x1 conv
                          on the block INPUT, openaimodel.py:246-256, model.py:104-119) fused into the block's last 3x3
                          conv: out = conv_kxk([x0 | x1]) + conv_1x1([skip_x0 | skip_x1]) accumulated in the same TMEM
                          tile.  skip inputs are NHWC [batch, in_h, in_w, skip_c*]; `weight` = bf16
                          [cout][k*k*(c0+c1) + skip_c0 + skip_c1] (the 1x1 weights appended along K), `bias` = the sum
                          of both biases.  NULL => no fused skip. */
  const void* skip_x1; /* second skip source (channel concat) or NULL */
  int32_t x0_pix_stride; /* tensor-core path: channels between consecutive pixels of x0 when x0 is a channel SLICE of a
                            wider NHWC tensor (0 => c0, dense) */
  int32_t res_batch;     /* tensor-core path: samples in `residual` when it has fewer than `batch` (broadcast as
                            b % res_batch; 0 => batch).  Together these let a convolution over a concat [h | skip] whose
                            skip half is shared by the two halves of a guided batch run as conv(h) + conv(skip), the
                            second term computed once and added here as a broadcast fp32 residual */
  int32_t x1_pix_stride; /* the same for x1 (0 => c1, dense) */
  int32_t gn_cstride;    /* channels per sample row of gn_coef (the GroupNorm's full channel count) */
  int32_t gn_c_off;      /* concat channel 0 of THIS convolution's input is channel gn_c_off of the GroupNorm */
  int32_t gn_silu;       /* 1 => SiLU after the normalisation (ResBlock in_layers / out_layers), 0 => none */
  const float* gn_coef;  /* tensor-core path, optional: GroupNorm (+ SiLU) applied to the RAW input inside the kernel's
                            operand path (util.py:199-216, openaimodel.py:268-288) — fp32 [samples][gn_cstride][2] =
                            (scale, shift) per (sample, channel) from stedm_gn_fold_tiles; the input tensors are then the
                            producer's unnormalised outputs and no normalised tensor is ever written.  Needs ksize 3,
                            cout % 256 == 0 and whole-row pixel tiles (stedm_conv_tc_plan reports whether a shape
                            qualifies: it fails with a message when gn_coef is set and the shape does not).  The fused
                            skip input is NOT normalised.  NULL => the input is used as given. */
} stedm_conv_desc;

/* tcgen05 + TMEM + TMA implicit GEMM (bf16 operands, fp32 accumulate).  Requires in_dtype == STEDM_BF16, stride 1,
 * upsample 0, c0 % 64 == 0, c1 % 64 == 0, cout % 16 == 0; output NHWC (bf16/fp32) or NCHW fp32 (the eps / image
 * heads with 3 real channels: cout = 16 zero-padded weight rows, cout_store = 3).
 * Plain GEMMs (ksize 1, one source: the token-major linear layers of the style encoder) relax this to c0 % 8 == 0
 * (the last K slab is zero-filled by TMA) and cout % 32 == 0 (the last output-channel tile is partially stored). */
int stedm_conv_tc(const stedm_conv_desc* d, void* stream);
/* Bytes of `workspace` stedm_conv_tc would use for this descriptor (0 when it runs in a single pass; always 0 when
 * stats_out is set: the fused statistics need the single-pass epilogue). */
long long stedm_conv_tc_workspace_bytes(const stedm_conv_desc* d);
/* The launch plan stedm_conv_tc would choose for this descriptor, computed on the host without touching the device
 * (no pointer of the descriptor is dereferenced; non-NULL-ness of x1 / skip_x0 / stats_out / workspace is what selects
 * the variants), so the tile-geometry checks and the mode selection are testable on a CPU-only machine:
 * plan8[0] = output-channel tile BN, [1] = cluster size, [2] = 1 for the cta_group::2 CTA pair, [3] = 1 for halo mode
 * (one activation box per (channel block, horizontal tap)), [4] = activation ring slots, [5] = bytes per slot,
 * [6] = split-K factor, [7] = K slabs per output tile.  Returns the same error codes as stedm_conv_tc. */
int stedm_conv_tc_plan(const stedm_conv_desc* d, int32_t* plan8);
/* General fp32-accumulate SIMT implicit GEMM: the fp32 parity mode and every shape the tensor-core path rejects. */
int stedm_conv_simt(const stedm_conv_desc* d, void* stream);

/* Batched GEMM C[z] = alpha * A[z] * op(B[z]) on CUDA cores (fp32 mode attention: QK^T and PV,
 * openaimodel.py:388-393, model.py:185-197).  z = (zb, zh): pointer offset = zb*stride_b + zh*stride_h (elements).
 * A is [m][k] row-major with leading dimension lda; B is [n][k] (b_is_nk != 0) or [k][n]; C is [m][n] with ldc. */
int stedm_gemm_simt(const void* a, const void* b, void* c, int dtype_a, int dtype_b, int dtype_c, int m, int n, int k,
                    int lda, int ldb, int ldc, int b_is_nk, int nb, int nh, long long a_sb, long long a_sh,
                    long long b_sb, long long b_sh, long long c_sb, long long c_sh, float alpha, void* stream);
/* Row softmax of scale * x over the last dimension of a [rows][cols] fp32 matrix (openaimodel.py:392, model.py:189-190).
 * x is used as scratch; the probabilities go to `out` (fp32 or bf16 [rows][cols]); out == NULL => in place. */
/* mask_diag_period = T > 0: the rows are the queries of [.., T, T] score matrices and column row % T (the token itself)
 * is excluded — sViT's LSA diagonal mask (networks/vit_set.py:52-54); 0 = plain softmax. */
int stedm_softmax_rows(float* x, void* out, int out_dtype, long long rows, int cols, float scale,
                       int mask_diag_period, void* stream);

/* GEGLU of the SpatialTransformer feed-forward (ldm/modules/attention.py:37-44): in = [rows][2f] = [x | gate] (the
 * proj Linear's output), out = [rows][f] = x * gelu(gate) (exact erf), fp32 or bf16. */
int stedm_geglu(const void* in, void* out, int dtype, long long rows, int f, void* stream);

/* K5  Fused flash-style self-attention on tcgen05 (S and O accumulators in TMEM, online fp32 softmax, P staged as
 * bf16 in shared memory): the U-Net AttentionBlock, head_dim 64 or 128, any token count.
 * q, k, v: bf16, token-major: element (b, h, t, c) at base + b*stride_b + h*stride_h + t*stride_t + c (strides in
 * elements, multiples of 8).  out: bf16 [batch][tokens][heads*head_dim].  scale multiplies q.k (= ch^-1/2).
 * Replaces QKVAttentionLegacy.forward (openaimodel.py:378-394): (q*ch^-1/4).(k*ch^-1/4), fp32 softmax, .v */
/* out_stride_b: elements between samples of `out` (0 => tokens*heads*head_dim, dense).  mask_diag != 0: a token does
 * not attend to itself — sViT's LSA (networks/vit_set.py:44-60: softmax(q.k^T * exp(temperature) with the diagonal
 * masked) . v), head_dim 64, scale = exp(temperature). */
/* Cross-attention (CrossAttention.forward with a context, ldm/modules/attention.py:169-193): tokens_kv > 0 keys /
 * values per sample with their own strides kv_stride_*; tokens_kv == 0 => self-attention (k, v use q's strides). */
int stedm_attention_tc(const void* q, const void* k, const void* v, void* out, int batch, int heads, int tokens,
                       int head_dim, long long stride_b, long long stride_h, long long stride_t, float scale,
                       long long out_stride_b, int mask_diag, int tokens_kv, long long kv_stride_b,
                       long long kv_stride_h, long long kv_stride_t, void* stream);


/* ----------------------------------------------------------------------------------------------------
 * Data movement helpers.
 */
/* K9 (unfused form): nearest x2 upsample, NHWC [b,h,w,c] -> [b,2h,2w,c] (openaimodel.py:129, model.py:54). */
int stedm_upsample_nearest2x(const void* x, void* out, int dtype, int batch, int h, int w, int c, void* stream);
/* K2 helper: gather the 3x3 / stride 2 / pad 1 receptive fields: NHWC [b,h,w,c] -> [b,h/2,w/2,9*c] (tap-major), so the
 * Downsample conv (openaimodel.py:164-166) runs as a plain GEMM on the tensor-core path. */
int stedm_im2col_3x3_s2(const void* x, void* out, int dtype, int batch, int h, int w, int c, void* stream);
/* DiffusionWrapper 'hybrid' input concat (ldm/models/diffusion/ddpm.py:1414-1415): NCHW fp32 sources [x0 | x1] ->
 * NHWC `out_dtype` with channels zero-padded to c_pad.  x1 may be NULL. */
int stedm_pack_nchw_to_nhwc(const float* x0, int c0, const float* x1, int c1, void* out, int out_dtype, int batch,
                            int hw, int c_pad, void* stream);
/* NHWC (any dtype) -> NCHW fp32 (API boundary). */
int stedm_nhwc_to_nchw_f32(const void* x, int dtype, float* out, int batch, int hw, int c, void* stream);

/* ----------------------------------------------------------------------------------------------------
 * K8  Embedding path.
 * timestep_embedding (ldm/modules/diffusionmodules/util.py:151-171): out[b] = [cos(t f) | sin(t f)], dim even. */
int stedm_timestep_embedding(const long long* t, float* out, int batch, int dim, void* stream);
/* The same for fractional timesteps: DPM-Solver feeds the U-Net t = (t_continuous - 1/N) * 1000 as a float
 * (ldm/models/diffusion/dpm_solver/dpm_solver.py:246-255; util.py:165 multiplies timesteps[:, None].float()). */
int stedm_timestep_embedding_f32(const float* t, float* out, int batch, int dim, void* stream);
/* out[b][n] = g(bias[n] + sum_k f(in[b][k]) * w[n][k]) (nn.Linear (out,in) layout); act is a bit set:
 * 1 = f is SiLU, 2 = f is ReLU, 4 = g is ReLU.  Covers time_embed (openaimodel.py:530-534), all emb_layers
 * (openaimodel.py:231-237) in one launch when their weights are stacked along n, the style encoder's head
 * (s_zss_dm.py:20) and Agg_Linear's ReLU-Linear-ReLU-Linear-ReLU block (agg_blocks.py:15-19). */
int stedm_linear(const float* in, const float* w, const float* bias, float* out, int batch, int k, int n, int act,
                 void* stream);

/* ----------------------------------------------------------------------------------------------------
 * Style encoder: torchvision swin_v2_t (networks/s_zss_dm.py:19-20, embedder.head = Linear(768, 512)) called by the
 * aggregation blocks (networks/agg_blocks.py:24-33, 47-54, 66-75).  The token-major linear layers (qkv, proj, MLP,
 * PatchMergingV2.reduction) run on stedm_conv_tc / stedm_conv_simt with ksize 1; these are the kernels between them.
 * Token maps are NHWC [batch, h, w, c]; torchvision file = torchvision/models/swin_transformer.py.
 */
/* features.0 = Conv2d(3, embed, 4, stride 4) + Permute + LayerNorm(embed).  img: NHWC fp32 [batch, p, p, 3] (the
 * layout style_imgs arrive in, modules/ldm_diffusion.py:51-60); w: fp32 [48][embed] with row (dy*4+dx)*3 + c;
 * outputs [batch, p/4, p/4, embed] as fp32 and/or bf16 (either may be NULL). */
int stedm_patch_embed_ln(const float* img, const float* w, const float* bias, const float* gamma, const float* beta,
                         float eps, float* out_f32, void* out_bf16, int batch, int p, int patch, int embed,
                         void* stream);
/* y = [residual +] LayerNorm(x) * gamma + beta over rows of c channels (c % 8 == 0, c <= 1024): Swin-V2's
 * res-post-norm x = x + norm(f(x)) (SwinTransformerBlockV2.forward), PatchMergingV2.norm, and sViT's pre-norms.
 * x: [rows][c] fp32 or bf16; residual: fp32 or NULL; the result is written as fp32 and/or bf16. */
int stedm_layernorm(const void* x, int x_dtype, const float* residual, const float* gamma, const float* beta, float eps,
                    float* out_f32, void* out_bf16, long long rows, int c, void* stream);
/* ShiftedWindowAttentionV2 core (shifted_window_attention with logit_scale): per 8x8 window and head,
 * softmax(normalize(q).normalize(k)^T * logit_scale[head] + rel_bias[head] + shift mask) . v, with the cyclic shift,
 * window partition and their inverses folded into the indexing.  qkv: [batch, h, w, 3*heads*32] = [q | k | v] (the
 * qkv Linear's output, k bias already zeroed by the caller as torchvision does); logit_scale: fp32 [heads] =
 * exp(min(param, log 100)); rel_bias: fp32 [heads][64][64] = 16*sigmoid(cpb_mlp(table))[index] (input independent,
 * computed at weight-pack time); out: [batch, h, w, heads*32].  A map that is not a whole number of windows is
 * zero-padded as torchvision does (F.pad before the qkv Linear): padded tokens act as keys with q, k, v = qkv_bias
 * (fp32 [3*heads*32], k part zero; NULL = no bias) and are dropped from the output; no shift along an axis one window wide. */
int stedm_window_attention(const void* qkv, int dtype, const float* logit_scale, const float* rel_bias,
                           const float* qkv_bias, void* out, int batch, int h, int w, int heads, int head_dim,
                           int window, int shift, void* stream);
/* PatchMergingV2's gather (_patch_merging_pad): [batch, h, w, c] -> [batch, h/2, w/2, 4c] in x0|x1|x2|x3 order. */
int stedm_patch_merge_gather(const void* x, void* out, int dtype, int batch, int h, int w, int c, void* stream);
/* SwinTransformer.forward tail: norm -> permute -> avgpool -> flatten: out[b][c] = mean_t LN(x[b][t])[c].
 * x: fp32 [batch][tokens][c], c % 32 == 0, c <= 1024. */
int stedm_ln_meanpool(const float* x, const float* gamma, const float* beta, float eps, float* out, int batch,
                      int tokens, int c, void* stream);
/* Agg_Mean (mode 0) / Agg_Max (mode 1) over the n style images of a sample: fp32 [b][n][f] -> [b][f]. */
int stedm_set_reduce(const float* x, float* out, int b, int n, int f, int mode, void* stream);

/* style_agg=svit (networks/vit_set.py sViT, built by networks/s_zss_dm.py:31-38): its Linear layers run on
 * stedm_conv_tc / stedm_conv_simt, the LayerNorms on stedm_layernorm, LSA on stedm_attention_tc (mask_diag) or
 * stedm_gemm_simt + stedm_softmax_rows (mask_diag_period); these are the remaining pieces.
 * SPT patch tokens (vit_set.py:82-107): img = style images NHWC fp32 [batch][ns][p_img][p_img][3]; out = fp32
 * [batch][(p_img/patch)^2][patch*patch*3*ns], patch element (p1, p2, c*ns + s) = img[b][s][h*patch+p1][w*patch+p2][c]. */
int stedm_spt_patchify(const float* img, float* out, int batch, int ns, int p_img, int patch, void* stream);
/* Token sequence of sViT.forward (vit_set.py:176-190): out fp32 [batch][t_pad][dim], row 0 = cls + pos[0], row 1 =
 * pos[1] (t_emb is None on the sampling path -> zeros), row 2+i = patches[b][i] + pos[2+i], rows >= n_patches+2 zero
 * (t_pad rounds the sequence up to whole 128-row GEMM tiles).  patches: [batch][n_patches][dim] fp32 or bf16. */
int stedm_svit_assemble(const void* patches, int dtype, const float* cls, const float* pos, float* out, int batch,
                        int n_patches, int t_pad, int dim, void* stream);
/* pool = 'mean' (vit_set.py:195-196): mean over the first `tokens` rows of each sample of fp32 [batch][t_pad][c]. */
int stedm_token_mean(const float* x, float* out, int batch, int tokens, int t_pad, int c, void* stream);

/* ----------------------------------------------------------------------------------------------------
 * K12  VQ nearest-code lookup (taming VectorQuantizer2.forward via ldm/models/autoencoder.py:277):
 * argmin_n |z|^2 + |e_n|^2 - 2 z.e_n over the codebook [n_codes][c] (fp32), first minimum wins (torch.argmin);
 * z, zq NCHW fp32 [batch, c, hw]; idx int32 [batch*hw] (may be NULL).  c <= 8. */
int stedm_vq_nearest(const float* z, const float* codebook, float* zq, int* idx, int batch, int c, int hw,
                     int n_codes, void* stream);

/* K13  SpatialRescaler (ldm/modules/encoders/modules.py:123-130): n_stages x bilinear 0.5 (= 2x2 mean each) then a
 * bias-free 1x1 conv.  seg: NCHW fp32 [b, cin, p, p]; w: [cout][cin]; out NCHW fp32 [b, cout, p>>n_stages, ...]. */
int stedm_spatial_rescale(const float* seg, const float* w, float* out, int batch, int cin, int cout, int p,
                          int n_stages, void* stream);

/* predict_step tail (modules/ldm_diffusion.py:94-96): clip to [-1,1], (x+1)*127.5, truncating uint8 cast,
 * NCHW fp32 -> NHWC uint8. */
int stedm_image_to_uint8(const float* img, uint8_t* out, int batch, int c, int hw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STEDM_B200_H */
