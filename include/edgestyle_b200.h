/* edgestyle_b200 -- C ABI of the B200 (sm_100a) kernels behind the EdgeStyle denoise hot path.
 *
 * The reference (andrei-ace/EdgeStyle) has no FFI of its own for this path: its boundary is the
 * Python `forward` of three classes, and everything below them is ATen library kernels
 * (SURVEY.md 2.1).  Each entry point here therefore cites the reference call site(s) whose
 * library dispatch it replaces.  All pointers are DEVICE pointers owned by the caller (PyTorch);
 * nothing allocates, nothing synchronises; every call enqueues on `stream` (a cudaStream_t passed
 * as void*).  Return value: 0 on success, negative on error; `es_last_error()` gives the text.
 * Thread-compatible: no global mutable state except the lazily resolved driver entry point.
 *
 * Activations are channels-last ("NHWC" == [tokens, channels]) fp16 or bf16 (`dtype`), all
 * accumulation and statistics are fp32.
 */
#ifndef EDGESTYLE_B200_H
#define EDGESTYLE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define ES_DTYPE_F16 0
#define ES_DTYPE_BF16 1

#define ES_ACT_NONE 0
#define ES_ACT_GEGLU 1 /* out[:, j] = v[:, j'] * gelu_erf(v[:, j' + tile/2]); weights pre-permuted per tile */
#define ES_ACT_SILU 2  /* out = silu(alpha * (acc + bias + rowvec)) + residual (ControlNetConditioningEmbedding convs) */

#define ES_MAX_SEG 4

const char* es_last_error(void);
int es_abi_version(void);
/* Programmatic dependent launch for every kernel of the library (off by default): each kernel's prologue then
 * overlaps the previous kernel's tail and it waits (griddepcontrol.wait) before its first global access.
 * Returns the previous setting. */
int es_set_pdl(int enabled);

/* ------------------------------------------------------------------------------------------
 * es_gemm: tcgen05/TMEM implicit-GEMM for conv3x3(stride 1, pad 1) / conv1x1 / Linear.
 *   replaces: every cuDNN conv and cuBLAS GEMM under UNet2DConditionModel.forward
 *   (/root/reference/model/edgestyle_pipeline.py:500-510) and CachedControlNetModel.forward
 *   (/root/reference/model/controllora.py:197-254): ResnetBlock2D conv1/conv2/shortcut,
 *   Transformer2DModel proj_in/out, attention to_q/k/v/out, GEGLU FF, zero-convs; and the
 *   LoRACompatibleLinear low-rank update (controllora.py:578-593) as a K-extension of the same
 *   accumulator (source 2 = x @ down^T, B2 = up).
 *
 *   out[m, n] = alpha * act( sum_{tap, c} A[pixel(m) + tap, c] * B[noff + n, tap, c]
 *                            + sum_{c2} A2[m, c2] * B2[noff2 + n, c2] + bias[noff + n] + rowvec[img(m), n] )
 *               + residual[m, n]
 *
 *   A is [n_img, h, w, c1] with pixel pitch `lda` elements (flat GEMM: w = M, h = n_img = 1).
 *   Segments: rows (flat GEMM) or IMAGES (taps == 9; boundaries on tile boundaries) [seg_row_start[g],
 *   seg_row_start[g+1]) use weight-row offset seg_b_noff[g] in B/bias and seg_b2_noff[g] in B2 (-1: no source 2
 *   for that segment).  This is how one batched pass serves the UNet rows and the ControlLoRA rows: a fused weight
 *   copy W + up_g down_g per LoRA group (Linear: controllora.py:578-593; conv: :561-575), or -- unfused -- the
 *   rank-r update as source 2 (A2 = x (*) down_g, B2 = up_g).
 * ------------------------------------------------------------------------------------------ */
typedef struct EsGemm {
  int dtype;
  const void* a;
  int c1;
  long long lda;
  int w, h, n_img;
  int taps; /* 1 or 9 */
  const void* b; /* [n_total_b][taps][c1] */
  int n_total_b;
  const void* a2; /* optional [same pixels as A][c2], pitch lda2 (1x1 / centre tap) */
  int c2;
  long long lda2;
  const void* b2; /* [n_total_b2][c2] */
  int n_total_b2;
  int n; /* GEMM N (columns of the accumulator) */
  int nseg;
  int seg_row_start[ES_MAX_SEG + 1];
  int seg_b_noff[ES_MAX_SEG];
  int seg_b2_noff[ES_MAX_SEG];
  const float* bias;   /* fp32, indexed noff + n; may be NULL */
  const float* rowvec; /* fp32 [imgs][rowvec_ld]; may be NULL */
  int rows_per_img;    /* flat mode: img(m) = m / rows_per_img */
  int rowvec_ld;
  const void* residual; /* same dtype as A; may be NULL */
  long long ldr;
  int act;
  float alpha;
  void* out;
  long long ldc;
  int out_fp32;
  int block_n; /* 0 = auto */
  int stages;  /* smem pipeline depth, 0 = auto */
  int split_k; /* 0 = auto, 1 = off, >1 = forced; < -1 = cooperative split-K with -split_k splits (every split CTA
                  reduces and finishes its own column chunks; at most 148 CTAs, long K); needs `workspace` */
  int b_blocked; /* 1: `b` is stored K-block-major, [taps * ceil(c1/64)][n_total_b][64] (K zero-padded to 64): every
                    weight tile is one contiguous chunk, which is what DRAM wants when weights are streamed once */
  float* gn_ws; /* optional: accumulate GroupNorm statistics of the OUTPUT here ([img][gn_groups][2] sum/sumsq, zeroed
                   by the caller) so the consuming GroupNorm skips its statistics pass; needs rows-per-image % 32 == 0
                   (flat mode: rows_per_img must be set) */
  int gn_groups;
  int gn_cpg;  /* 0: n / gn_groups.  > 0: this GEMM writes a column SLICE of the normalised tensor (one half of an
                  [x | skip] concat): channels per group of the whole tensor ... */
  int gn_col0; /* ... and the channel of that tensor its column 0 lands on; gn_groups = groups of the whole tensor */
  /* LayerNorm folded around the GEMM (BasicTransformerBlock norm1/2/3 never touch memory):
   *   producer: rowstat_out [M][2] fp32 (zeroed by the caller) receives per-row (sum, sumsq) of THIS GEMM's output
   *             (after bias / residual), accumulated over its column tiles;
   *   consumer: ln_rowstat [M][2] holds those sums for the rows of A; B must be W * gamma (per K column), `bias`
   *             bias + W beta, ln_colsum[noff + n] = sum_k B[n, k]; the epilogue then computes
   *             rstd_m * (acc - mean_m * ln_colsum[n]) + bias[n]  ==  LayerNorm(A)[m, :] . W[n, :] + bias. */
  float* rowstat_out;
  const float* ln_rowstat;
  const float* ln_colsum;
  int ln_features; /* number of features the statistics were taken over (C) */
  float ln_eps;
  const void* prefetch; /* optional: `prefetch_bytes` of constant global memory (the next layer's weights) that this
                           launch pulls into L2 while it runs (cp.async.bulk.prefetch.L2), spread over its CTAs */
  long long prefetch_bytes;
  void* workspace; /* optional split-K scratch: first 64 KiB = tile counters (zero-initialised ONCE by the caller,
                      self-resetting), rest = fp32 partial tiles.  Must not be shared by concurrent launches. */
  long long workspace_bytes;
} EsGemm;
int es_gemm(const EsGemm* g, void* stream);

/* ------------------------------------------------------------------------------------------
 * es_attention: flash-style softmax(Q K^T * scale) V on tcgen05/TMEM tiles fed by TMA.
 *   replaces: F.scaled_dot_product_attention under diffusers Attention (AttnProcessor2_0) in every
 *   BasicTransformerBlock attn1/attn2 (reached from controllora.py:205-238 and
 *   edgestyle_pipeline.py:500).
 *   q: [batch, nq, heads*d] pitch ldq; k, v: [batch, nkv, heads*d] pitch ldk / ldv; out like q.
 * ------------------------------------------------------------------------------------------ */
typedef struct EsAttention {
  int dtype;
  const void* q;
  const void* k;
  const void* v;
  void* out;
  long long ldq, ldk, ldv, ldo; /* elements between consecutive tokens */
  long long bsq, bsk, bsv, bso; /* elements between consecutive batch items */
  int batch, heads, d, nq, nkv;
  float scale; /* d^-0.5 */
} EsAttention;
int es_attention(const EsAttention* a, void* stream);

/* ------------------------------------------------------------------------------------------
 * GroupNorm(32 groups, affine) [+ SiLU] over NHWC, optionally over the channel concat of two sources.
 *   replaces: nn.GroupNorm + SiLU + torch.cat in ResnetBlock2D / Transformer2DModel.norm / conv_norm_out.
 *   stats: ws[img][group][2] (sum, sumsq) fp32, zeroed by the caller.
 * ------------------------------------------------------------------------------------------ */
typedef struct EsGroupNorm {
  int dtype;
  const void* x0; /* [n_img, hw, c0] pitch ld0 */
  int c0;
  long long ld0;
  const void* x1; /* optional second source (channel concat) */
  int c1;
  long long ld1;
  int n_img, hw, groups;
  float eps;
  const float* gamma; /* [c0+c1] */
  const float* beta;
  float* ws; /* [n_img][groups][2] */
  void* out; /* [n_img, hw, c0+c1] pitch ldo */
  long long ldo;
  int silu;
} EsGroupNorm;
int es_groupnorm_stats(const EsGroupNorm* g, void* stream);
int es_groupnorm_apply(const EsGroupNorm* g, void* stream);
/* Single-launch GroupNorm(+SiLU) for inputs whose statistics no producer accumulated: a cluster of 8 CTAs per
 * (image, group) holds the group's slab in registers and reduces through distributed shared memory (x1 must be NULL,
 * hw % 8 == 0, even channels per group, hw * channels-per-group <= 262144; `ws` is not used). */
int es_groupnorm_fused(const EsGroupNorm* g, void* stream);

/* LayerNorm over the channel dim of [rows, c] (eps 1e-5, affine): BasicTransformerBlock.norm1/2/3. */
int es_layernorm(int dtype, const void* x, long long ldx, void* out, long long ldo, const float* gamma,
                 const float* beta, int rows, int c, float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * EdgeStyle merge (ControlNetBlock, /root/reference/model/edgestyle_multicontrolnet.py:23-63 with
 * the interleave of :160-164,479-514 folded into indexing; closed form in SURVEY.md A.9), for up to
 * ES_MERGE_MAX_LEVELS residual levels per launch (the 13 blocks of :104-114 run as ONE grid per phase).
 *   res[k]  : residual of net k, [B, hw, C] (zero-conv outputs, NOT yet scaled); k = 0..5
 *   scale   : DEVICE vector [6] of conditioning_scale (controllora.py:267-270); a level multiplies it by
 *             `gain` (guess_mode: logspace(-1, 0, 13)[level], controllora.py:257-265).  Nets whose scale is 0
 *             are not read (control_guidance gating, edgestyle_pipeline.py:418-427).
 *   params (repacked channels-last by the host): w1 [3][2][C], b1 [3][C], g1/be1 [hw][3][C] (dtype),
 *           w2 [3][C], b2 [C], g2/be2 [hw][C] (dtype), w3 [C], b3 [C]
 *   phase 1: sums of u  -> stats[b][0..1];  phase 2: z ([B,hw,C], fp32 or dtype) + sums of z -> stats[b][2..3];
 *   phase 3: dst[b,p,c] = skip[b,p,c] + w3*SiLU(LN(z))+b3   (skip may be NULL; dst pitch ldd)
 * ------------------------------------------------------------------------------------------ */
#define ES_MERGE_MAX_LEVELS 16
typedef struct EsMergeLevel {
  const void* res[6];
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  const void *g1, *be1, *g2, *be2;
  double* stats; /* [B][4], zeroed by the caller */
  void* z;       /* [B][hw][C], fp32 when z_f32 else `dtype` */
  const void* skip;
  void* dst;
  /* optional (phase 3): accumulate GroupNorm statistics of what is written to dst, for a consumer that normalises the
   * tensor dst is a column slice of: ws [B][gn_groups][2] (sum, sumsq; zeroed by the caller), channels per group of the
   * whole tensor, channel of that tensor column 0 of dst lands on */
  float* gn_ws;
  long long lds, ldd;
  int hw, C;
  int gn_groups, gn_cpg, gn_col0;
  int z_f32;
  float gain;
} EsMergeLevel;
typedef struct EsMergeBatch {
  int dtype;
  int B;
  int n_levels;
  const float* scale; /* DEVICE pointer, 6 floats */
  EsMergeLevel levels[ES_MERGE_MAX_LEVELS];
} EsMergeBatch;
int es_merge_levels(const EsMergeBatch* m, int phase, void* stream);

/* ------------------------------------------------------------------------------------------ misc */
/* sinusoidal timestep embedding (flip_sin_to_cos, freq_shift 0): out fp32 [n][dim]; controllora.py:150 */
int es_timestep_embedding(const float* t, int n, int dim, float* out, void* stream);
/* y[r][n] = act_out( sum_k act_in(x[r][k]) * W[n][k] + bias[n] ) (+ y if accumulate); tiny-M linears:
 * time_embedding.linear_1/2, resnet time_emb_proj, and their LoRA down/up.  W in `dtype`, x/y fp32. */
int es_small_linear(int dtype, const float* x, int ldx, const void* w, const float* bias, float* y, int ldy, int rows,
                    int n, int k, int silu_in, int silu_out, int accumulate, void* stream);
/* NCHW fp32 [n,c,h,w] -> NHWC dtype [n,h,w,c_pad] (zero padded channels), and back. */
int es_nchw_to_nhwc(int dtype, const float* src, void* dst, int n, int c, int hw, long long ldd, void* stream);
int es_nhwc_to_nchw(int dtype, const void* src, long long lds, float* dst, int n, int c, int hw, void* stream);
/* im2col for 3x3 pad 1 stride s over NHWC: out [n*ho*wo][ldo] with column = tap*c + ch (zero padded to ldo). */
int es_im2col3x3(int dtype, const void* src, long long lds, void* dst, long long ldo, int n, int h, int w, int c,
                 int stride, void* stream);
/* same with explicit zero padding per side (0 or 1): pad_lo rows/columns before, pad_hi after the image.  The VAE
 * encoder's Downsample2D(padding=0) pads (0, 1, 0, 1) and then runs a stride-2 conv (diffusers AutoencoderKL, reached
 * from /root/reference/model/controllora.py:39). */
int es_im2col3x3_pad(int dtype, const void* src, long long lds, void* dst, long long ldo, int n, int h, int w, int c,
                     int stride, int pad_lo, int pad_hi, void* stream);
/* nearest-neighbour x2 upsample NHWC (Upsample2D before its conv). */
int es_upsample2x(int dtype, const void* src, long long lds, void* dst, long long ldd, int n, int h, int w, int c,
                  void* stream);
/* y = a + b elementwise over [rows, c] with pitches (skip + residual etc.). */
int es_add(int dtype, const void* a, long long lda, const void* b, long long ldb, void* out, long long ldo, int rows,
           int c, void* stream);
/* CFG combine + DDIM update (edgestyle_pipeline.py:513-522): eps NCHW fp32 [2*imgs, c, hw] (uncond rows first),
 * latents fp32 [imgs, c, hw] updated in place; guidance per image; coef = {sqrt(a_t), sqrt(1-a_t), sqrt(a_prev),
 * sqrt(1-a_prev)}.  eps_out (optional) receives the guided eps. */
int es_cfg_ddim(const float* eps, float* latents, const float* guidance, const float* coef, float* eps_out, int imgs,
                int chw, void* stream);

/* UniPC (reference default scheduler, /root/reference/app.py:118) on the device: x0-prediction from the guided
 * noise (convert_model_output), and the predictor / corrector updates as linear combinations whose scalar
 * coefficients the host derives from the sigma table (diffusers UniPCMultistepScheduler, bh2, order 2). */
int es_cfg_x0(const float* eps, const float* sample, const float* guidance, float alpha, float sigma, float* x0,
              int imgs, int chw, void* stream);
int es_lincomb4(float* out, float c0, const float* x0, float c1, const float* x1, float c2, const float* x2, float c3,
                const float* x3, long long n, void* stream);

/* ------------------------------------------------------------------------------------------ VAE (per-call stages)
 * The AutoencoderKL encoder (/root/reference/model/controllora.py:38-42, reached once per call from
 * /root/reference/model/edgestyle_pipeline.py:660-662) and decoder (edgestyle_pipeline.py:552-557) run on es_gemm /
 * es_groupnorm; their mid-block attention has ONE head as wide as the block (512), beyond es_attention's TMEM budget,
 * so its scores go through es_gemm (fp32 out), this row softmax, and es_gemm again.
 *   p[r][0..cols) = softmax(scale * s[r][0..cols)); s fp32 pitch lds, p `dtype` pitch ldp. */
int es_softmax_rows(int dtype, const float* s, long long lds, void* p, long long ldp, int rows, int cols, float scale,
                    void* stream);
/* DiagonalGaussianDistribution.sample() of the encoder moments (controllora.py:39) times `scale` (:40):
 * moments fp32 [n*hw][ldm] with columns (mean[0..L), logvar[0..L)); noise fp32 NCHW [n][L][hw] or NULL (= mode());
 * out fp32 NCHW [n][L][hw] = (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise) * scale. */
int es_gaussian_sample(const float* moments, long long ldm, const float* noise, float* out, int n, int latent_channels,
                       int hw, float scale, void* stream);

/* Profiling probe (tools/timeline.py): a 1-thread kernel that stores the GPU's %globaltimer (ns) into *slot once all
 * earlier work of `stream` has completed -- lets a captured multi-stream step be timed op by op. */
int es_stamp(unsigned long long* slot, void* stream);

#ifdef __cplusplus
}
#endif
#endif
