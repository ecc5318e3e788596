/*
 * ldmseg_b200.h -- C ABI of the B200-native LDMSeg sampling hot path.
 *
 * The reference (segments-ai/latent-diffusion-segmentation) is pure Python: its hot path has no
 * FFI of its own, every FLOP is a PyTorch / diffusers library call.  The entry points below are
 * therefore the operator set that the reference's Python call sites reduce to; each one cites the
 * reference call site (file:line under /root/reference) whose arithmetic it replaces.  The Python
 * mirror of `ldmseg.models` / `ldmseg.schedulers` in latent-diffusion-segmentation_b200/ldmseg
 * binds them with ctypes.
 *
 * Conventions
 *   - plain C types only: device pointers as void* / typed pointers, sizes as int / int64_t,
 *     the CUDA stream as void* (cudaStream_t).  No torch types.
 *   - every call is asynchronous on the given stream and returns 0 on success, non-zero on
 *     failure; ldmseg_last_error_string() describes the failure (thread-local).
 *   - no call allocates memory that outlives it; the caller owns all buffers.
 *   - activations are channel-last ("NHWC"): a row-major [rows = n*H*W pixels, C channels] matrix,
 *     bf16 unless stated; weights are bf16 [N out-channels, K] K-contiguous, K ordered
 *     [segment][tap ky,kx][channel], each (segment, tap) slice padded to a multiple of 64.
 */
#ifndef LDMSEG_B200_H_
#define LDMSEG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDMSEG_ABI_VERSION 4

/* ---- library ---------------------------------------------------------------------------- */
int ldmseg_version(void);
const char* ldmseg_last_error_string(void);
/* Number of kernel launches issued through this library since load (for bench accounting). */
int64_t ldmseg_launch_count(void);

/* ---- implicit GEMM on tcgen05 (conv3x3 / conv1x1 / linear) ------------------------------ *
 * Replaces F.conv2d / F.linear inside diffusers ResnetBlock2D, Transformer2DModel, Attention,
 * FeedForward, Downsample2D, Upsample2D (call sites ldmseg/models/unet.py:357,361-373,388-395,
 * 401-425,431) and the nn.Conv2d / nn.ConvTranspose2d of GeneralVAESeg.decode
 * (ldmseg/models/vae.py:133,155,164,267-271).
 *
 *   out[m, n] = epilogue( sum_seg sum_tap sum_c  A_seg[pixel(m) + tap, c] * Wt[n, k(seg,tap,c)] )
 *
 * Up to LDMSEG_MAX_SRC channel-last sources share the output's spatial geometry (nb, h, w);
 * a segment reads one source with 1 tap (1x1 / linear) or 9 taps (3x3, zero padding 1).
 * Several segments accumulate into the same output tile (concat-free skip connections, fused
 * 1x1 shortcut).  A plain [M, K] matrix is the geometry nb=1, h=1, w=M.
 */
#define LDMSEG_MAX_SRC 3
#define LDMSEG_MAX_SEG 4

enum { LDMSEG_OUT_BF16 = 0, LDMSEG_OUT_F32 = 1 };
enum { LDMSEG_ACT_NONE = 0, LDMSEG_ACT_SILU = 1, LDMSEG_ACT_GEGLU = 2 };

typedef struct ldmseg_igemm_params {
  /* A sources */
  const void* src[LDMSEG_MAX_SRC];   /* bf16 [nb*h*w, src_c[i]] */
  int src_c[LDMSEG_MAX_SRC];         /* channels (row stride, elements); multiple of 8 */
  int nsrc;
  int nb, h, w;                      /* output (= input) geometry; M = nb*h*w */
  /* K segments, in weight order */
  int nseg;
  int seg_src[LDMSEG_MAX_SEG];       /* index into src[] */
  int seg_taps[LDMSEG_MAX_SEG];      /* 1 or 9 (4: see upsample2) */
  /* B operand */
  const void* weight;                /* bf16 [n, ktot] */
  int n;                             /* output channels */
  int ktot;                          /* sum over segments of taps * roundup(src_c, 64) */
  /* epilogue */
  const float* bias;                 /* [n] or NULL */
  const float* rowbias;              /* [nb, rowbias_ld] added per image, or NULL */
  int rowbias_ld;
  const void* residual;              /* bf16 [M, res_ld] or NULL */
  int res_ld;
  void* out;                         /* bf16 or f32 [M, out_ld]; GEGLU writes n/2 columns */
  int out_ld;
  int out_dtype;                     /* LDMSEG_OUT_* */
  int act;                           /* LDMSEG_ACT_* */
  /* scheduling */
  int block_n;                       /* 0 = choose; else 64 / 128 / 160 / 256; 320 with `pair` only (two N = 160
                                        tcgen05.mma per k-step sharing the staged A rows; no split_k, no GEGLU) */
  int split_k;                       /* 0/1 = none; >1 needs workspace */
  float* workspace;                  /* split-K partial tiles, f32, tiles*split_k*128*block_n elements */
  int* tile_counters;                /* split-K: 8192 int32, zero-initialised once (self-resetting) */
  long long workspace_elems;         /* capacity of workspace in f32 elements */
  float* stats;                      /* optional f32 [nb, n, 2]: += per-(image, channel) sum and sum of
                                        squares of the stored (bf16-rounded) output -- GroupNorm statistics
                                        fused into the producer; caller zeroes it; needs h*w % 32 == 0 */
  int stats_hw;                      /* rows per image for `stats` (0 = h*w); lets a plain [M, K] GEMM
                                        (nb=1, h=1, w=M) produce per-image statistics */
  int weight_tiled;                  /* 1: weight is stored block-tiled [ceil(n/16)][ktot/64][16][64] (each 16x64
                                        block 2 KB contiguous, zero-padded rows) instead of row-major [n, ktot] */
  int pdl;                           /* 1: launch with programmatic dependent launch (overlap the prologue
                                        with the previous kernel's tail) */
  int pair;                          /* 1: CTA pairs (clusters of 2, tcgen05 cta_group::2): 256 x block_n tiles, each
                                        CTA stages its 128 rows of A and half of the B tile.  Needs weight_tiled,
                                        block_n in {128, 160, 256, 320} and M > 128 */
  /* ---- ABI version 2 ---- */
  int weight_static;                 /* 1: nothing on the stream writes `weight` (real parameters), so its first
                                        tiles may be fetched BEFORE the grid-dependency wait under `pdl`.  Must be 0
                                        when the B operand was produced by an earlier launch (VAE attention K / V^T) */
  const void* next_weight;           /* optional: weights of the NEXT igemm launch on the stream; each CTA pulls a
                                        slice of it into L2 once its own operand loads are in flight (hint only) */
  long long next_weight_bytes;
  int residual_f32;                  /* 1: `residual` is f32 [M, res_ld] (fp32 residual stream) */
  void* out2;                        /* optional bf16 [M, out2_ld] shadow of an f32 `out`: the copy that later launches
                                        read through TMA (shortcut / down-sampling operands) */
  int out2_ld;
  int conv_stride;                   /* 0/1: stride 1 with zero padding 1.  2: 3x3 stride-2 convolution; (nb, h, w) is
                                        the OUTPUT geometry, every source is [nb, 2h, 2w, c] and is read through the
                                        TMA traversal stride (diffusers Downsample2D, unet.py:361-373; the
                                        AutoencoderKL encoder's F.pad(0,1,0,1) + stride-2 conv) */
  int conv_pad;                      /* stride 2 only: zero padding before (top / left): 1 (UNet) or 0 (VAE) */
  /* LayerNorm folded into the GEMMs on either side of it (BasicTransformerBlock.norm1 / norm3 feed to_q/k/v and
   * ff.net.0.proj): the PRODUCER of the row x accumulates its per-row moments, the CONSUMER multiplies the raw x by
   * W' = W diag(gamma) and finishes  out = rstd_m * (acc - mean_m * colsum[n]) + bias[n],  bias = W beta + b,
   * colsum[n] = sum_k W'[n, k] (of the bf16-rounded W').  One launch per LayerNorm disappears. */
  float* rowstats_out;               /* producer: f32 [M, 2] += per-row {sum, sum of squares} of the stored bf16
                                        output (of the bf16 shadow `out2` for an f32 out); caller zeroes it */
  const float* ln_rowstats;          /* consumer: the producer's [M, 2] */
  const float* ln_colsum;            /* consumer: f32 [n] */
  int ln_channels;                   /* consumer: row width C of the LayerNorm (mean = sum / C) */
  float ln_eps;
  /* ---- ABI version 3 ---- */
  int stream_k;                      /* 1: stream-K tail.  The tiles past the last whole wave of the persistent grid are
                                        cut along K into one contiguous piece per CTA; pieces that end inside a tile
                                        go through `workspace` / `tile_counters` (same buffers as split_k, which must
                                        be 0/1) and the CTA holding a tile's last k-block finishes it.  Ignored (whole
                                        tiles) when the last wave is full, for GEGLU and for block_n 64 */
  int split_cluster;                 /* 1 (with split_k > 1, not with pair): launch the split_k CTAs of every tile as one
                                        thread-block cluster and exchange the partial tiles through distributed shared
                                        memory instead of `workspace` (which is then unused).  Falls back to the
                                        workspace exchange when the device cannot hold all the tiles' clusters at once
                                        (see ldmseg_igemm_max_split_clusters) */
  /* ---- ABI version 4 ---- */
  int upsample2;                     /* 1: nearest x2 up-sampling folded into the 3x3 convolution that follows it
                                        (diffusers Upsample2D = F.interpolate(scale_factor=2, mode="nearest") + conv,
                                        reached from ldmseg/models/unet.py:401-425).  (nb, h, w) is the INPUT geometry,
                                        `out` is [nb*2h*2w, out_ld].  The launch runs four phase GEMMs over the input
                                        pixels: output pixel (2y+py, 2x+px) is a 2x2 convolution of the input whose
                                        tap (a, b) reads (y+py+a-1, x+px+b-1) with the sum of the 3x3 taps that land
                                        there -- 4/9 of the multiply-adds and no up-sampled tensor.  One segment with
                                        seg_taps = 4; `weight` = the four phase matrices (phase = 2*py+px) stacked along
                                        n, block-tiled: [4*ceil(n/16)][4*roundup(c,64)/64][16][64]; ktot =
                                        4*roundup(c,64).  `stats_hw` counts INPUT rows per image.  Bias, SiLU, f32 /
                                        shadow outputs, split_k, stream_k, pair (even number of 128-row input tiles)
                                        and fused statistics are supported; residual / rowbias / LayerNorm folds /
                                        GEGLU / conv_stride 2 are not */
} ldmseg_igemm_params;

int ldmseg_igemm(const ldmseg_igemm_params* p, void* stream);

/* Number of clusters of `cluster_size` (2..16) split-K CTAs of the given tile width the device holds at once; a split
 * launch with split_cluster needs tiles <= this.  0 when such clusters cannot be scheduled.  Needs a CUDA device. */
int ldmseg_igemm_max_split_clusters(int block_n, int geglu, int cluster_size);

/* Reference-grade CUDA-core version of the same contract (fp32 accumulate, no tensor cores).
 * Test infrastructure for the tcgen05 kernel -- never on the product path. */
int ldmseg_igemm_simple(const ldmseg_igemm_params* p, void* stream);

/* ---- normalisation ---------------------------------------------------------------------- *
 * GroupNorm(32 groups)+optional SiLU over a (virtually concatenated) pair of channel-last
 * sources; replaces nn.GroupNorm + nn.SiLU in diffusers ResnetBlock2D.norm1/norm2,
 * Transformer2DModel.norm, UNet conv_norm_out (ldmseg/models/unet.py:428-430) and the GroupNorm
 * of the seg decoder (ldmseg/models/vae.py:162).  stats is a scratch f32 [nb, groups, 2]. */
int ldmseg_groupnorm(const void* src0, int c0, const void* src1, int c1, int nb, int hw, int groups,
                     const float* gamma, const float* beta, float eps, int silu, void* out,
                     float* stats, void* stream);
/* Apply pass only: statistics come from the producers' fused per-(image, channel) {sum, sumsq}
 * (ldmseg_igemm_params.stats), f32 [nb, c, 2] per source.  One launch per GroupNorm. */
int ldmseg_groupnorm_apply_cs(const void* src0, int c0, const float* chan_stats0, const void* src1,
                              int c1, const float* chan_stats1, int nb, int hw, int groups,
                              const float* gamma, const float* beta, float eps, int silu, void* out,
                              void* stream);
/* Same, both sources f32 (the fp32 residual stream written by ldmseg_igemm with out_dtype F32). */
int ldmseg_groupnorm_apply_cs_f32(const void* src0, int c0, const float* chan_stats0, const void* src1,
                                  int c1, const float* chan_stats1, int nb, int hw, int groups,
                                  const float* gamma, const float* beta, float eps, int silu, void* out,
                                  void* stream);
/* Launch every kernel of the library with programmatic dependent launch (prologue of kernel i+1
 * overlaps the tail of kernel i); returns the previous setting. */
int ldmseg_set_pdl(int enable);
/* Development switch for kernel experiments (0 in production); returns the previous value. */
int ldmseg_set_debug(int flags);
/* LayerNorm over the channel dim of each row (tokens or pixels); replaces nn.LayerNorm in
 * BasicTransformerBlock.norm1/norm3 and LayerNorm2d (ldmseg/models/vae.py:309-322). */
int ldmseg_layernorm(const void* src, int rows, int c, const float* gamma, const float* beta,
                     float eps, int silu, void* out, void* stream);
/* Same with an f32 source row (fp32 residual stream); output stays bf16 (a GEMM operand). */
int ldmseg_layernorm_f32(const void* src, int rows, int c, const float* gamma, const float* beta,
                         float eps, int silu, void* out, void* stream);

/* Row softmax with scale (f32 scores -> bf16 probabilities): the fp32 softmax of diffusers 0.16.1
 * AttentionBlock (single head, d = 512) in the AutoencoderKL mid block. */
int ldmseg_softmax_rows(const float* s, int rows, int cols, float scale, void* out, void* stream);

/* ---- attention -------------------------------------------------------------------------- *
 * softmax(Q K^T / sqrt(d)) V per (image, head); replaces F.scaled_dot_product_attention in
 * diffusers AttnProcessor2_0 (self-attention only; cross-attention is removed by
 * ldmseg/models/unet.py:83-105).  qkv is bf16 [nb*ntok, 3*heads*d] (q | k | v column blocks),
 * out is bf16 [nb*ntok, heads*d]. */
int ldmseg_attention(const void* qkv, int nb, int ntok, int heads, int d, void* out, void* stream);
int ldmseg_attention_simple(const void* qkv, int nb, int ntok, int heads, int d, void* out,
                            void* stream);
/* Cross-attention (diffusers BasicTransformerBlock.attn2; kept by the conditioned variants of
 * ldmseg/models/descriptors.py:67-105 and driven with a doubled batch by the guidance path of
 * ldmseg/trainers/trainers_ldm_cond.py:1098-1123,1143-1146): q is bf16 [nb*ntok_q, heads*d] (to_q of the
 * image tokens), kv is bf16 [nb*ntok_kv, 2*heads*d] (to_k | to_v of encoder_hidden_states; 77 text or 257 image
 * tokens: key columns beyond ntok_kv are masked), out is bf16 [nb*ntok_q, heads*d].  Same kernel as above. */
int ldmseg_cross_attention(const void* q, const void* kv, int nb, int ntok_q, int ntok_kv, int heads,
                           int d, void* out, void* stream);

/* ---- element-wise / layout -------------------------------------------------------------- */
/* h * gelu_erf(g) for x = [rows, 2*c] = (h | g): diffusers GEGLU. */
int ldmseg_geglu(const void* x, int rows, int c, void* out, void* stream);
/* nearest x2 upsample of a channel-last bf16 tensor (diffusers Upsample2D). */
int ldmseg_upsample2x(const void* src, int nb, int h, int w, int c, void* out, void* stream);
/* im2col for 3x3 stride-2 convs: out [nb*ho*wo, 9*c], taps ordered (ky,kx); pad_lo = zero padding
 * before (top/left): 1 for the UNet Downsample2D (padding=1), 0 for the VAE encoder
 * (F.pad(0,1,0,1) then padding=0). */
int ldmseg_im2col_s2(const void* src, int nb, int h, int w, int c, int pad_lo, void* out,
                     void* stream);
/* NCHW f32 -> channel-last bf16 with channel padding (cpad >= c, zeros), optional affine
 * (y = x*scale + shift) : UNet input / VAE input (2x-1, ldmseg/trainers/trainers_ldm_cond.py:369). */
int ldmseg_nchw_to_nhwc_bf16(const float* src, int nb, int c, int hw, int cpad, int coff,
                             float scale, float shift, void* out, void* stream);
/* channel-last f32 [nb*hw, ld] (first c columns) -> NCHW f32, optional scale. */
int ldmseg_nhwc_f32_to_nchw(const float* src, int nb, int c, int hw, int ld, float scale,
                            float* out, void* stream);
/* NCHW f32 -> channel-last f32 [nb*hw, ld] (first c columns), optional scale. */
int ldmseg_nchw_f32_to_nhwc(const float* src, int nb, int c, int hw, int ld, float scale, float* out,
                            void* stream);
/* channel-last bf16 [nb*hw, ld] -> NCHW f32 */
int ldmseg_nhwc_bf16_to_nchw(const void* src, int nb, int c, int hw, int ld, float scale,
                             float* out, void* stream);

/* ---- scheduler -------------------------------------------------------------------------- *
 * DDIM (eta = 0) update; replaces DDIMNoiseScheduler.step
 * (ldmseg/schedulers/ddim_scheduler.py:218-269).  All tensors f32, any layout (element-wise).
 * prediction_type: 0 epsilon, 1 sample, 2 v_prediction.  prev_sample / pred_x0 may be NULL.
 * Optional ancestral noise (DDPM extension, eta=1): sigma > 0 with noise != NULL. */
int ldmseg_ddim_step(const float* model_out, const float* sample, int64_t n, float alpha_t,
                     float alpha_prev, int prediction_type, int clip, float clip_range,
                     int use_clipped, float sigma, const float* noise, float* prev_sample,
                     float* pred_x0, void* stream);

/* Fused sampler step used by the CUDA-graph loop (ldmseg/trainers/trainers_ldm_cond.py:1127-1159):
 * reads eps f32 [M,4] (channel-last conv_out output) and the f32 latent state [M,4]; writes the
 * new state, x0, and the next UNet input row bf16 [M,16] = (x_{t-1} | rgb | x0 | 0).
 * The step index is read from device memory (*step_ptr) so one captured graph serves all steps;
 * coef is f32 [nsteps, 4] = (sqrt_a_t, sqrt_1ma_t, sqrt_a_prev, sqrt_1ma_prev); on the last step
 * the state becomes x0 (Q1 in SURVEY.md).  Optional inpainting blend (extension): mask f32 [M],
 * known f32 [nsteps, M, 4] (already noised to t_prev by the caller per step: known[step]); optional ancestral
 * noise f32 [nsteps, M, 4] with sigma f32 [nsteps] (DDPM extension).  prediction_type / clip / clip_range are the
 * scheduler's (ddim_scheduler.py:238-257).  cfg = 1: classifier-free guidance (trainers_ldm_cond.py:1143-1146):
 * eps is [2M,4] (uncond | cond rows), eps = e_u + guidance * (e_c - e_u), and the next UNet input is written for
 * both halves (rows i and M + i); not defined together with self_cond (the reference's shapes do not allow it). */
int ldmseg_sampler_step(const float* eps, float* latents, float* x0, const float* rgb_latents,
                        void* unet_in, int64_t m, const float* coef, const int* step_ptr,
                        int nsteps, int self_cond, const float* mask, const float* known,
                        const float* noise, const float* sigma, int prediction_type, int clip,
                        float clip_range, int cfg, float guidance, void* stream);
int ldmseg_advance_step(int* step_ptr, void* stream);
/* Same update as ldmseg_ddim_step, but the timestep is a 0-dim int64 tensor ON THE DEVICE and the
 * alphas_cumprod table (f32 [num_train_timesteps]) lives on the device too, so `step` needs no
 * device->host synchronisation (the reference pays three per call: ddim_scheduler.py:234-235). */
int ldmseg_ddim_step_indexed(const float* model_out, const float* sample, int64_t n,
                             const int64_t* timestep_dev, const float* alphas_cumprod_dev,
                             int step_ratio, float final_alpha, int prediction_type, int clip,
                             float clip_range, int use_clipped, float* prev_sample, float* pred_x0,
                             void* stream);

/* add_noise (mode 0) / remove_noise (mode 1) with PER-SAMPLE timesteps read on the device
 * (ddim_scheduler.py:155-216; the training-step no-grad forward, trainers_ldm_cond.py:813-831):
 *   mode 0: out = sqrt(a_t) * scale * x + sqrt(1 - a_t) * noise      mode 1: out = (x - sqrt(1 - a_t) * noise) / (sqrt(a_t) * scale)
 * x / noise / out f32 [nb, per_sample]; timesteps int64 [nb] and alphas_cumprod f32 on the device.  Optional
 * unet_in (mode 0, x = NCHW [nb, c, hw]): also writes channels [0, c) of the channel-last bf16 [nb*hw, cpad] UNet input. */
int ldmseg_noise_mix(const float* x, const float* noise, const int64_t* timesteps_dev,
                     const float* alphas_cumprod_dev, int nb, int64_t per_sample, float scale, int mode,
                     float* out, void* unet_in, int hw, int cpad, void* stream);

/* ---- time embedding --------------------------------------------------------------------- *
 * y[r, :] = act_out( W x_act(r) + b ), f32, small-batch (r = timesteps): Timesteps sinusoid,
 * TimestepEmbedding and the 22 time_emb_proj linears (ldmseg/models/unet.py:303-307). */
int ldmseg_timestep_sinusoid(const float* t, int rows, int dim, int flip_sin_to_cos,
                             float freq_shift, float* out, void* stream);
int ldmseg_small_linear(const float* x, int rows, int k, const float* w, const float* b, int n,
                        int silu_in, int silu_out, float* out, int out_ld, void* stream);
/* dst[b, :] = table[*step_ptr, :] for b < nb: selects the current step's precomputed time-embedding
 * biases inside a captured graph (the step index lives in device memory). */
int ldmseg_select_row(const float* table, int ncols, const int* step_ptr, int nb, float* dst,
                      void* stream);

/* ---- seg decoder tail ------------------------------------------------------------------- *
 * ConvTranspose2d(k2,s2) output comes out of ldmseg_igemm as [M, 4*c] (tap-major columns);
 * this scatters it to the 2x grid and applies LayerNorm2d + SiLU (vae.py:155-157, 309-322). */
int ldmseg_convt_shuffle_ln(const void* src, int nb, int h, int w, int c, const float* gamma,
                            const float* beta, float eps, int silu, void* out, void* stream);
/* bilinear x2 (align_corners=False) of channel-last bf16/f32 logits to NCHW f32 (vae.py:270). */
int ldmseg_bilinear2x_to_nchw(const void* src, int src_is_f32, int nb, int h, int w, int c, int ld,
                              float* out, void* stream);
/* fused fast path: bilinear x2 + argmax + softmax max-prob -> ids u8 [nb, 2h, 2w], prob f32
 * (ldmseg/trainers/trainers_ldm_cond.py:428-433). */
int ldmseg_bilinear2x_argmax(const void* src, int src_is_f32, int nb, int h, int w, int c, int ld,
                             uint8_t* ids, float* maxprob, void* stream);

/* ---- panoptic post-processing (the per-image tail of compute_pq) --------------------------------- *
 * Replaces ldmseg/trainers/trainers_ldm_cond.py:1261-1313, which resizes the 134 MB/image logits per image,
 * copies the full [128, h, w] sigmoid map to the host and filters segments in numpy.
 *   logits  f32 channel-last [nb, s, s, ld] (first c = 128 columns): the seg decoder's output BEFORE its bilinear
 *           x2 (vae.py:270); the x2 and the per-image resize (:1267-1272) are composed inside the kernel.
 *   geom    int32 [nb, 6] ON THE DEVICE = {h, w, crop_y0, crop_x0, crop_h, crop_w}: target size and the padding
 *           crop (crop_padding, :1172-1178) on the 2s x 2s grid.  max_hw >= max h*w; out_stride = pixels per image
 *           in pred / ids.
 *   pred    int16 [nb, out_stride]: argmax class, -1 where the max softmax probability < mask_th (:1275-1284)
 *   area / orig_area  int32 [nb, 128]: #pixels with pred == c / #pixels with sigmoid(logit_c) >= mask_th (:1301) */
int ldmseg_panoptic_resample(const float* logits, int nb, int s, int c, int ld, const int* geom_dev,
                             int max_hw, int out_stride, float mask_th, int threshold_output,
                             int16_t* pred, int* area, int* orig_area, void* stream);
/* keep[c] = area[c] >= count_th && c != ignore_label && area[c] / orig_area[c] >= overlap_th (:1293-1304);
 * ids (u8 [nb, out_stride]) = keep[pred] ? pred + 1 : 0 (:1296-1313); keep int32 [nb, 128] = segments_info. */
int ldmseg_panoptic_filter(const int16_t* pred, int nb, const int* geom_dev, int max_hw, int out_stride,
                           const int* area, const int* orig_area, int count_th, double overlap_th,
                           int ignore_label, uint8_t* ids, int* keep, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LDMSEG_B200_H_ */
