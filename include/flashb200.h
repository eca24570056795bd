/* flashb200.h — C ABI of libflashb200.so: the B200 (sm_100a) kernels under the Flash-Diffusion
 * distillation step.
 *
 * The reference (gojasper/flash-diffusion) has no native layer: every entry point below replaces
 * arithmetic that the reference reaches through diffusers / peft / torch library calls.  Each
 * prototype cites the reference call site (file:line under the reference tree) whose math it
 * implements; the UPSTREAM (diffusers/peft) math is restated in oracle/ and SURVEY.md §8a-L1.
 *
 * Conventions
 *  - plain pointers + sizes only; every pointer is DEVICE memory owned by the caller
 *    (PyTorch caching allocator on the Python host side);
 *  - asynchronous and stream-ordered on `stream` (a cudaStream_t passed as void*); no internal
 *    synchronisation; re-entrant across streams;
 *  - activations are bf16, channels-last: images are NHWC, token matrices are [rows, channels];
 *  - return 0 on success, negative on error; fd_last_error() gives a thread-local message;
 *  - the library never allocates persistent device memory.
 */
#ifndef FLASHB200_H
#define FLASHB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FD_MAX_TAPS 16

const char* fd_last_error(void);
int fd_version(void);
/* compute capability (major*10+minor) of the current device, or negative if no device */
int fd_sm_arch(void);
/* number of kernels this library has launched in this process (bench.py's `gpu_launches`) */
long long fd_launch_count(void);
/* Per-launch CUDA-event timing of the tensor-core kernels on their launch stream (bench.py roofline).
 * categories: 0 plain GEMM, 1 implicit-GEMM conv, 2 attention fwd, 3 attention bwd.
 * fd_profile_summary synchronises the device, fills ms / algorithmic FLOPs / launch counts per category
 * and clears the records. */
void fd_profile_enable(int on);
int fd_profile_summary(double* ms, double* flops, long long* counts, int ncat);
/* writes one CSV row per recorded launch (cat,M,N,K,ms,flops) WITHOUT clearing the records */
int fd_profile_dump(const char* path);

/* ------------------------------------------------------------------------------------------
 * fd_gemm — tcgen05/TMA GEMM with an optional second K segment and a fused epilogue.
 *
 *   acc[M,N] = A1[M,K1] * B1[N,K1]^T  (+ A2[M,K2] * B2[N,K2]^T)          (bf16 in, fp32 acc)
 *   LayerNorm fold (optional, ln_stats != NULL): the A operand is the RAW residual stream x and B is W*gamma;
 *     acc <- rstd[row] * (acc - mean[row] * ln_colsum[col])     with (mean, rstd) from ln_stats[row] =
 *     (sum x, sum x^2) over ln_inv_c^-1 channels — i.e. LayerNorm(x) W^T without a LayerNorm pass over memory
 *     (bias must then hold b + W beta).  UPSTREAM BasicTransformerBlock.norm1/2/3 -> to_q/k/v, ff.net.0.proj.
 *   acc += bias[N]                      (fp32, optional)
 *   acc += rowvec[row / rows_per_group, :N] (fp32, row stride ldrv; the ResnetBlock time-embedding add)
 *   geglu: out[:, 16j+i] = acc[:, 32j+i] * gelu(acc[:, 32j+16+i])   (weights packed interleaved)
 *   out (+)= residual[M, Nout]          (bf16, optional)
 *   out -> bf16 (or fp32 when out_fp32)
 *   rowstats_out (optional): [M,2] fp32 += (sum, sum of squares) of the stored (bf16-rounded) output row — the
 *     statistics the NEXT LayerNorm-folded GEMM consumes; zeroed by the call.
 *
 * Segment 2 is how LoRA (peft `y = base(x) + (x A^T) B^T * alpha/r`; reference call sites
 * examples/train_flash_sdxl.py:210-217) is folded in: A2 = x A^T [M,r], B2 = s*B [N,r]; it is also
 * how the ResnetBlock2D 1x1 conv_shortcut is accumulated into conv2 (UPSTREAM ResnetBlock2D).
 *
 * conv mode (conv_taps > 0): A1 is an NHWC activation [NB_in, H, W, C] and K1 = conv_taps * C;
 * output row r <-> pixel (n, h, w) of an [NB, H, W] grid, tap t reads input pixel
 * (n + tap_dn[t], h + tap_dh[t], w + tap_dw[t]) with zero fill outside [0,H)x[0,W) — i.e. an
 * implicit-GEMM convolution (UPSTREAM ResnetBlock2D / Downsample2D / Upsample2D convs,
 * reference call path src/flash/models/unets/unet.py:108-119).  B1 is [N, conv_taps*C] with K
 * ordered (tap, channel).  C must be a multiple of 8; W must divide or be a multiple of 128.
 *
 * Replaces: torch.nn.functional.linear / conv2d library calls under
 * src/flash/models/unets/unet.py:108-119 (denoiser forward).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t M, N;
    const void* a1; int64_t lda1; const void* b1; int64_t ldb1; int32_t K1;
    const void* a2; int64_t lda2; const void* b2; int64_t ldb2; int32_t K2;
    /* conv mode */
    int32_t conv_taps, NB_in, H, W, C;
    int32_t tap_dn[FD_MAX_TAPS], tap_dh[FD_MAX_TAPS], tap_dw[FD_MAX_TAPS];
    /* epilogue */
    const float* bias;
    const float* rowvec; int32_t rows_per_group; int64_t ldrv;   /* row stride of rowvec (elements) */
    int32_t geglu;
    const void* residual; int64_t ldr;
    void* out; int64_t ldo; int32_t out_fp32;
    int32_t force_bn;   /* 0 = heuristic, else 64/128/256 single-CTA, 512+128 / 512+160 / 512+256 CTA-pair; | 1024: cut
                         * tiles along K between CTA pairs (stream-K) whenever legal; | 2048: whole tiles only */
    /* LayerNorm fold */
    const float* ln_stats; const float* ln_colsum; float ln_inv_c; float ln_eps;
    float* rowstats_out;
    int32_t rowstats_prezeroed;   /* 1: the caller already zeroed rowstats_out (one memset for a whole arena of them) */
    /* DiT epilogue: act (0 none, 1 gelu-tanh, 2 ReLU) after the bias, then acc *= rowscale[row / rows_per_group_scale, :N]
     * (fp32, row stride ldrs) before the residual add — the AdaLN gate of PixArt / SD3 blocks
     * (UPSTREAM BasicTransformerBlock(ada_norm_single): hidden = gate * attn(...) + hidden). */
    int32_t act;
    const float* rowscale; int32_t rows_per_group_scale; int64_t ldrs;
    /* Caller-owned scratch for the stream-K schedule of the CTA-pair kernels (tiles cut along K between CTA pairs,
     * partial accumulators exchanged through it).  >= fd_gemm_workspace_bytes() bytes, 16-byte aligned, ZERO-FILLED
     * once by the caller (the kernels leave it zero-flagged); calls that may run CONCURRENTLY (different streams)
     * need different workspaces.  NULL: no K cuts (whole tiles per CTA pair, wave-quantised). */
    void* workspace; int64_t workspace_bytes;
    /* GroupNorm statistics from the producer: colstats_out [M / colstats_rows, N, 2] fp32 += per-image (sum, sum of
     * squares) of every stored (bf16-rounded) output column, accumulated by the epilogue (one 8-byte reduction per
     * column per 32 rows) — what fd_groupnorm_apply_cols of the FOLLOWING GroupNorm consumes instead of a reduction
     * pass over the activation.  The caller zero-fills it.  Needs bf16 output, N % 32 == 0, colstats_rows % 32 == 0,
     * no GEGLU; the call then always takes the CTA-pair kernel.  NULL: off. */
    float* colstats_out; int32_t colstats_rows;
} FdGemmArgs;
int fd_gemm(const FdGemmArgs* args, void* stream);
size_t fd_gemm_workspace_bytes(void);

/* ------------------------------------------------------------------------------------------
 * Normalisation / elementwise (HBM-bound) kernels — NHWC bf16 activations.
 * ------------------------------------------------------------------------------------------ */

/* GroupNorm apply from producer-side column statistics (FdGemmArgs.colstats_out of the GEMM / conv that wrote x):
 * colstats [NB, C, 2] fp32 (sum, sum of squares per image and channel) -> group mean / rstd formed per block,
 * y = act((x - mean) * rstd * gamma + beta).  One pass over x (read + write) and no reduction launch.
 * stats_out (optional) [NB, G, 2] (mean, rstd). */
int fd_groupnorm_apply_cols(const void* x, const float* colstats, const float* gamma, const float* beta, void* y,
                            float* stats_out, int32_t NB, int32_t HW, int32_t C, int32_t G, float eps,
                            int32_t silu_act, void* stream);

/* GroupNorm statistics: x [NB, HW, C] bf16 -> stats [NB, G, 2] fp32 (mean, rstd).
 * UPSTREAM torch.nn.GroupNorm inside ResnetBlock2D / Transformer2DModel. */
int fd_groupnorm_stats(const void* x, float* stats, int32_t NB, int32_t HW, int32_t C, int32_t G,
                       float eps, void* stream);
/* y = (x - mean) * rstd * gamma + beta, optionally followed by SiLU.  y bf16 [NB,HW,C]. */
int fd_groupnorm_apply(const void* x, const float* stats, const float* gamma, const float* beta,
                       void* y, int32_t NB, int32_t HW, int32_t C, int32_t G, int32_t silu,
                       void* stream);
/* Both of the above in two launches (reduce, then apply with the mean / rstd formed in the apply kernel): raw is a
 * caller scratch [NB, G, 2] fp32 (zeroed by the call); stats_out (optional, [NB, G, 2]) receives (mean, rstd) for
 * fd_groupnorm_bwd. */
int fd_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* raw, float* stats_out,
                     int32_t NB, int32_t HW, int32_t C, int32_t G, float eps, int32_t silu, void* stream);
/* GroupNorm(+SiLU) backward: dx from dy (both bf16 [NB,HW,C]); recomputes from x and stats. */
int fd_groupnorm_bwd(const void* x, const float* stats, const float* gamma, const float* beta,
                     const void* dy, void* dx, float* scratch /* [NB,G,2] */, int32_t NB,
                     int32_t HW, int32_t C, int32_t G, int32_t silu, void* stream);

/* LayerNorm over the last dim: x [rows, C] bf16 -> y bf16; optionally saves (mean, rstd) [rows,2].
 * UPSTREAM BasicTransformerBlock.norm1/2/3 (eps 1e-5, affine). */
int fd_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y,
                     float* stats /* may be NULL */, int32_t rows, int32_t C, float eps,
                     void* stream);
int fd_layernorm_bwd(const void* x, const float* stats, const float* gamma, const void* dy,
                     void* dx, int32_t rows, int32_t C, void* stream);
/* AdaLN modulation: y = LayerNorm(x) (no affine) * (1 + scale[b]) + shift[b], b = row / rows_per_batch;
 * scale / shift fp32 with row stride ld_mod.  UPSTREAM BasicTransformerBlock(norm_type="ada_norm_single") and
 * Transformer2DModel output norm (reference src/flash/models/transformers/tranformers.py:83-92). */
int fd_layernorm_modulate(const void* x, const float* scale, const float* shift, int64_t ld_mod, void* y,
                          int32_t rows, int32_t C, int32_t rows_per_batch, float eps, void* stream);
/* Backward of fd_layernorm_modulate (student LoRA backward through the MMDiT AdaLN blocks; reference autograd of
 * tranformers.py:103-150 under examples/train_flash_sd3.py:101-120, whose LoRA targets include every AdaLN linear):
 * dx [rows, C] bf16; dscale / dshift [rows/rows_per_batch, C] fp32 (dense, overwritten) = per-sample column sums of
 * dy * xhat and dy. */
int fd_layernorm_modulate_bwd(const void* x, const void* dy, const float* scale, int64_t ld_mod, void* dx,
                              float* dscale, float* dshift, int32_t rows, int32_t C, int32_t rows_per_batch,
                              float eps, void* stream);
/* AdaLN-Zero gate: out = res + gate[b] * h (training-time form of the fd_gemm rowscale + residual epilogue, which
 * keeps h for the gate gradient), and its backward dh = gate[b] * dout, dgate[b, c] = sum_rows dout * h
 * ([rows/rows_per_batch, C] fp32, overwritten).  gate fp32 with row stride ld_gate. */
int fd_gate_residual(const void* h, const float* gate, int64_t ld_gate, const void* res, void* out, int32_t rows,
                     int32_t C, int32_t rows_per_batch, void* stream);
int fd_gate_bwd(const void* dout, const void* h, const float* gate, int64_t ld_gate, void* dh, float* dgate,
                int32_t rows, int32_t C, int32_t rows_per_batch, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention (softmax(Q K^T / sqrt(d)) V), d = 64, bf16, fp32 softmax.
 * q [B, Nq, H*64] (row stride ldq), k/v [B, Nkv, H*64] (row strides ldk/ldv), o [B, Nq, H*64].
 * lse (optional) [B, H, Nq] fp32 for the backward.
 * UPSTREAM Attention + AttnProcessor2_0 (F.scaled_dot_product_attention), SURVEY §2.2.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const void* q; int64_t ldq; int64_t q_batch_stride;
    const void* k; int64_t ldk; int64_t k_batch_stride;
    const void* v; int64_t ldv; int64_t v_batch_stride;
    void* o; int64_t ldo; int64_t o_batch_stride;
    float* lse;
    int32_t B, H, Nq, Nkv;
    float scale;
} FdAttnArgs;
int fd_attn_fwd(const FdAttnArgs* args, void* stream);
/* Same contract for head dims other than 64 (multiple of 16, <= 192; q/k/v/o rows are H*head_dim wide) and an optional
 * key-padding mask kv_len[B] (keys >= kv_len[b] are ignored) — SD1.5 (d=40/80/160 zero-padded to 48/80/160 by the
 * weight packs) and PixArt-alpha (d=72 -> 80, masked T5 context).  Backward: fd_attn_bwd_generic. */
int fd_attn_fwd_generic(const FdAttnArgs* args, int32_t head_dim, const int32_t* kv_len, void* stream);

typedef struct {
    FdAttnArgs f;                 /* forward tensors (o = forward output, lse required) */
    const void* d_o; int64_t lddo; int64_t do_batch_stride;
    void* dq; int64_t lddq; int64_t dq_batch_stride;
    void* dk; int64_t lddk; int64_t dk_batch_stride;
    void* dv; int64_t lddv; int64_t dv_batch_stride;
    float* delta;                 /* scratch [B, H, Nq] fp32 */
    float* dq_accum;              /* scratch [B, Nq, H*64] fp32, zero-initialised by the call */
} FdAttnBwdArgs;
int fd_attn_bwd(const FdAttnBwdArgs* args, void* stream);
/* Backward for the shapes fd_attn_fwd_generic serves (same FdAttnBwdArgs; q/k/v/o/d_o/dq/dk/dv rows are H*head_dim wide,
 * dq_accum [B, Nq, H*head_dim] fp32): head_dim a multiple of 16 up to 80 runs on tcgen05 for any sequence length;
 * larger head dims (SD1.5's d = 160 at 16x16 / 8x8 tokens) run three CUDA-core passes for short sequences
 * (Nq*Nkv <= 2^20), for which dq_accum is a scratch of 2*B*H*Nq*Nkv floats.  kv_len as in the forward.  Student LoRA backward of the SD1.5 UNet and the PixArt-alpha DiT (reference unets/unet.py:108-119,
 * transformers/tranformers.py:58-92). */
int fd_attn_bwd_generic(const FdAttnBwdArgs* args, int32_t head_dim, const int32_t* kv_len, void* stream);

/* Row softmax y = softmax(scale * x) over the last dim of a bf16 [rows, L] matrix (row strides ldx / ldy, L % 8 == 0)
 * and its backward ds = p * (dp - sum(p * dp)) * scale.  Together with two fd_gemm calls (S = Q K^T, O = P V) this is
 * the single-head, 512-channel attention of the VAE mid block (UPSTREAM diffusers AutoencoderKL, reference
 * src/flash/models/vae/autoencoderKL.py:52-128), whose head dim is beyond what the fused attention kernels hold in TMEM. */
int fd_softmax_rows(const void* x, int64_t ldx, void* y, int64_t ldy, int32_t rows, int32_t L, float scale, void* stream);
int fd_softmax_rows_bwd(const void* p, const void* dp, void* ds, int64_t ld, int32_t rows, int32_t L, float scale,
                        void* stream);

/* LPIPS-VGG pieces of the `distill_loss_type="lpips"` objective (reference flash_diffusion_model.py:102-103,383-397;
 * lpips==0.1.4 restated in oracle/lpips.py): the VGG16 convolutions run on fd_gemm (conv mode, act = 2 ReLU), these are
 * the rest.  NHWC bf16 feature maps.
 *   fd_maxpool2x2 / _bwd   2x2 stride-2 max pooling and its gradient (first maximum wins, as torch)
 *   fd_relu_bwd            dx = dy where the ReLU output y > 0
 *   fd_lpips_layer         out[n] += mean_pixels sum_c w[c] (f0/(|f0|+1e-10) - f1/(|f1|+1e-10))^2   (out is NOT zeroed)
 *   fd_lpips_layer_bwd     df0 given gout[n] = dL/dout[n] */
int fd_maxpool2x2(const void* x, void* y, int32_t NB, int32_t H, int32_t W, int32_t C, void* stream);
int fd_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int32_t NB, int32_t H, int32_t W, int32_t C, void* stream);
int fd_relu_bwd(const void* y, const void* dy, void* dx, int64_t n, void* stream);
int fd_lpips_layer(const void* f0, const void* f1, const float* w, float* out, int32_t NB, int32_t HW, int32_t C,
                   void* stream);
int fd_lpips_layer_bwd(const void* f0, const void* f1, const float* w, const float* gout, void* df0, int32_t NB,
                       int32_t HW, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout / small elementwise helpers.
 * ------------------------------------------------------------------------------------------ */
/* NCHW fp32 -> NHWC bf16 with channel padding to Cpad (zeros). */
int fd_nchw_to_nhwc(const float* x, void* y, int32_t NB, int32_t C, int32_t H, int32_t W,
                    int32_t Cpad, void* stream);
/* NHWC (fp32 or bf16 rows of ld elements, first C valid) -> NCHW fp32. */
int fd_nhwc_to_nchw(const void* x, int32_t x_is_fp32, int64_t ld, float* y, int32_t NB, int32_t C,
                    int32_t H, int32_t W, void* stream);
/* DiT un-patchify: x [NB*h*w, p*p*Cout] fp32 (token rows, columns ordered (p, q, c)) -> y NCHW fp32
 * [NB, Ckeep, h*p, w*p] keeping the first Ckeep channels (reference tranformers.py:92 slices `[:, :in_channels]`). */
int fd_unpatchify(const float* x, float* y, int32_t NB, int32_t h, int32_t w, int32_t p, int32_t Cout, int32_t Ckeep,
                  void* stream);
/* gradient of fd_unpatchify: dy NCHW fp32 [NB, Ckeep, h*p, w*p] -> dx [NB*h*w, p*p*Cout] bf16 (zero for c >= Ckeep) */
int fd_patchify(const float* dy, void* dx, int32_t NB, int32_t h, int32_t w, int32_t p, int32_t Cout, int32_t Ckeep,
                void* stream);
/* nearest-neighbour 2x upsample, NHWC bf16 (UPSTREAM Upsample2D) */
int fd_upsample2x(const void* x, void* y, int32_t NB, int32_t H, int32_t W, int32_t C,
                  void* stream);
/* backward of the above: sums each 2x2 block */
int fd_upsample2x_bwd(const void* dy, void* dx, int32_t NB, int32_t H, int32_t W, int32_t C,
                      void* stream);
/* space-to-depth: x [NB,H,W,C] -> y [4*NB, H/2, W/2, C], phase p=(h&1)*2+(w&1) major.
 * Feeds the stride-2 convs (UPSTREAM Downsample2D, discriminator Conv 4x4 s2) in conv mode. */
int fd_space_to_depth(const void* x, void* y, int32_t NB, int32_t H, int32_t W, int32_t C,
                      void* stream);
/* inverse scatter used by the stride-2 conv backward */
int fd_depth_to_space(const void* x, void* y, int32_t NB, int32_t H, int32_t W, int32_t C,
                      void* stream);
/* copy rows [rows, C1] and [rows, C2] side by side into [rows, C1+C2] (skip-connection concat) */
int fd_concat_channels(const void* a, int32_t C1, const void* b, int32_t C2, void* y, int64_t rows,
                       void* stream);
/* y[rows, C] = x[rows, c0:c0+C] of a [rows, Ctot] matrix (skip-connection concat backward) */
int fd_slice_channels(const void* x, int32_t Ctot, int32_t c0, int32_t C, void* y, int64_t rows,
                      void* stream);
/* y = a + b (bf16, n elements) */
int fd_add(const void* a, const void* b, void* y, int64_t n, void* stream);
/* bf16 [rows, cols] -> bf16 [cols, rows] written with row stride ldy >= rows */
int fd_transpose(const void* x, void* y, int32_t rows, int32_t cols, int64_t ldy, void* stream);
/* fp32 -> bf16 with scale (weight packing, LoRA B * alpha/r) */
int fd_cast_scale(const float* x, void* y, int64_t n, float scale, void* stream);
/* SiLU on fp32 [n] -> bf16 (ResnetBlock2D time-embedding nonlinearity) */
int fd_silu_f32_to_bf16(const float* x, void* y, int64_t n, void* stream);
/* sinusoidal timestep embedding (UPSTREAM Timesteps(dim, flip_sin_to_cos=True, shift 0)):
 * t [NB] fp32 -> y [NB, dim] bf16 = [cos | sin] */
int fd_timestep_embedding(const float* t, void* y, int32_t NB, int32_t dim, void* stream);

/* GEGLU backward: given pre-activation acc (interleaved layout as written by fd_gemm with
 * geglu=0) [M, N] bf16 and dout [M, N/2], produce dacc [M, N] bf16 */
int fd_geglu_bwd(const void* acc, const void* dout, void* dacc, int64_t M, int32_t N, void* stream);
/* backward of the tanh-GELU the fd_gemm epilogue applies with act = 1 (UPSTREAM FeedForward "gelu-approximate"):
 * acc = recomputed pre-activation, n elements (multiple of 8), all bf16 */
int fd_gelu_tanh_bwd(const void* acc, const void* dout, void* dacc, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Distillation-step elementwise kernels (fp32 latents NCHW [B, C, H, W], n = C*H*W per sample)
 * reference: src/flash/models/flash/flash_diffusion_model.py
 * ------------------------------------------------------------------------------------------ */
/* :243-257 noising: out = sa[b] * z + sg[b] * noise */
int fd_step_add_noise(const float* z, const float* noise, const float* sa, const float* sg,
                      float* out, int32_t B, int64_t n, void* stream);
/* :316-322 CFG combine + DPM-Solver++(2M) update (oracle/schedulers.py):
 *   eps = w*eps_c + (1-w)*eps_u ; x0 = (x - sigma_t*eps)/alpha_t ;
 *   x_next = c_x*x - c_d0*x0 - c_d1*(x0 - x0_prev)/r ; x0_prev <- x0
 * coef = {w, alpha_t, sigma_t, c_x, c_d0, c_d1_over_r} (host scalars) */
int fd_step_cfg_dpm(const float* eps_c, const float* eps_u, float* x, float* x0_prev,
                    const float* coef6, int64_t n, void* stream);
/* :267-280,:328 student_output = c_skip[b]*x_t + c_out[b]*((x_t - sg[b]*eps)/sa[b]) and its
 * gradient factor d(student_output)/d(eps) = -c_out*sg/sa (returned by the host) */
int fd_step_student_output(const float* x_t, const float* eps, const float* sa, const float* sg,
                           const float* c_skip, const float* c_out, float* out, int32_t B,
                           int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLASHB200_H */
