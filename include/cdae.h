/*
 * cdae.h — C ABI of libcdae.so, the sm_100a kernel library behind causaldiffae_b200.
 *
 * The reference (Akomand/CausalDiffAE) is pure Python/PyTorch: it has no FFI, its arithmetic is
 * whatever ATen/cuDNN/cuBLAS does underneath `improved_diffusion/*.py`.  Each entry point below
 * replaces the ATen call sequence of one reference call site (cited file:line, paths relative to
 * the reference root).  Conventions:
 *   - plain pointers and sizes only; every buffer (incl. workspaces) is owned by the caller;
 *   - asynchronous launch on the given stream; no allocation, no host sync, graph-capturable;
 *   - returns 0 on success, a negative cdae_status otherwise; cdae_last_error() gives the text;
 *   - re-entrant / thread-safe (called from Python main and autograd worker threads);
 *   - sm_100a only: cdae_init() fails on any other device.  There is no CPU path.
 * Activations inside the UNet torso are NHWC bf16 ("pixels x channels"); fp32 accumulation.
 */
#ifndef CDAE_H_
#define CDAE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cdae_stream;   /* cudaStream_t */

enum cdae_status {
  CDAE_OK = 0, CDAE_ERR_ARG = -1, CDAE_ERR_SHAPE = -2, CDAE_ERR_CUDA = -3, CDAE_ERR_ARCH = -4
};

int cdae_version(void);
const char* cdae_last_error(void);
/* checks the current device is compute capability 10.x, resolves cuTensorMapEncodeTiled, sets func attributes */
int cdae_init(void);

/* ------------------------------------------------------------------ fused elementwise (HBM bound) */
/* q_sample: x_t = sqrt_ac[t[b]] * x0 + sqrt_1mac[t[b]] * noise      (gaussian_diffusion.py:201-222)
 * tables are device fp32 arrays of length T (fp32(round(float64)), cf. _extract_into_tensor :938-951). */
int cdae_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac,
                  const float* sqrt_1mac, float* x_t, int64_t B, int64_t per_sample, cdae_stream s);
/* per-sample MSE of (target - pred) and its gradient scale            (gaussian_diffusion.py:847, train_util.py:266)
 * mse[b] = mean_chw (target-pred)^2 ; if dpred != NULL: dpred = 2*(pred-target)*gscale[b]*gmul/per_sample
 * (gscale = the schedule sampler's weights, gmul = 1/B: the gradient of mean_b(mse[b] w[b])) */
int cdae_mse_loss(const float* pred, const float* target, float* mse, const float* gscale, float gmul, float* dpred,
                  int64_t B, int64_t per_sample, cdae_stream s);
/* DDIM x_{t-1} update with optional classifier-free guidance combine   (gaussian_diffusion.py:277-285,320-341,506-558)
 * coef_table: device fp32 [T'][8] rows {sqrt_recip_ac, sqrt_recipm1_ac, sqrt(ac_prev), sqrt(1-ac_prev-sigma^2),
 * sigma*[t!=0], clip_denoised, predict_xstart, 0}; the row is chosen on the device by t_idx[b*t_idx_stride]
 * (stride 0: one index for the whole batch, so a captured CUDA graph can be replayed for every step). */
int cdae_ddim_step(const float* x, const float* eps_c, const float* eps_u, float w, int use_w,
                   const float* coef_table, const int32_t* t_idx, int t_idx_stride, const float* noise,
                   float* x_prev, float* pred_xstart, int64_t B, int64_t per_sample, cdae_stream s);
/* fused AdamW + EMA + grad-norm over the flat parameter arena, one pass
 * (train_util.py:276-303 optimize_fp16 / optimize_normal, nn.py:503-513 update_ema, torch.optim.AdamW defaults)
 * hyper: device floats {lr, beta1, beta2, eps, weight_decay, ema_rate, grad_scale}; step: device counter of the steps
 * taken so far (bias corrections of step+1 are evaluated in double on the device; the counter is incremented after the
 * update); g: the gradient arena, fp32 or (g_is_bf16) its bf16 copy as all-reduced; guard: optional device scalar - when
 * it is not finite the launch changes nothing and the counter stays (the reference's "found NaN: skip the step",
 * train_util.py:277-280); gsq_out += sum((g*grad_scale)^2); lognorm (optional, 2 floats) += {sqrt(gsq_out), 1}: the running
 * mean of the gradient norm the logger reports (train_util.py:299-303) without a host read per step. */
int cdae_adam_ema(float* p, const void* g, int g_is_bf16, float* m, float* v, float* ema, const float* hyper,
                  int64_t* step, const float* guard, float* gsq_out, float* lognorm, int64_t n, cdae_stream s);
/* out += sum(g^2) over a flat fp32 (or bf16) buffer (the guard above; train_util.py:277 isfinite check, :299-303 grad norm) */
int cdae_sumsq(const void* g, int g_is_bf16, float* out, int64_t n, cdae_stream s);
/* fp32 -> bf16 (RNE) copy of a flat buffer: the gradient arena as it crosses NVLink (train_util.py:107-126 DDP buckets) */
int cdae_cast_bf16(const float* src, void* dst_bf16, int64_t n, cdae_stream s);
/* extra EMA rates (train_util.py:296-297): ema = rate*ema + (1-rate)*p */
int cdae_ema_update(float* ema, const float* p, float rate, int64_t n, cdae_stream s);
int cdae_zero(void* p, int64_t bytes, cdae_stream s);

/* ------------------------------------------------------------------ layout / packing */
/* NCHW fp32 image -> NHWC bf16 padded to Cpad channels (zeros)          (feeds unet.py:392 stem conv) */
int cdae_nchw_to_nhwc_pad(const float* x, void* out_bf16, int N, int C, int H, int W, int Cpad, cdae_stream s);
/* NHWC bf16 [N,H,W,ld] first C channels -> NCHW fp32 (gradient of the above / debugging) */
int cdae_nhwc_to_nchw(const void* x_bf16, float* out, int N, int C, int H, int W, int ld, cdae_stream s);
/* table-driven weight pack: fp32 master (OHWI physical) -> bf16 forward [Cout_pad][taps][Cin_pad] and
 * bf16 transposed [Cin_pad][taps][Cout_pad] copies used by the implicit-GEMM kernels. One launch for all layers. */
typedef struct {
  int64_t src_off;      /* element offset into fp32 arena */
  int64_t dst_fwd_off;  /* element offset into bf16 arena, -1: skip */
  int64_t dst_tr_off;   /* element offset into bf16 arena, -1: skip */
  int32_t cout, cin, taps, cout_pad, cin_pad;
  int32_t fwd_ld, tr_ld;  /* destination row pitch in elements (0: dense = taps*cin_pad / taps*cout_pad) */
  int32_t _pad;            /* first global tile index of this entry: prefix sum of taps*ceil(cout_pad/32)*ceil(cin_pad/32) */
} cdae_pack_entry;
int cdae_pack_weights(const float* arena, void* bf16_arena, const cdae_pack_entry* entries_dev, int n_entries,
                      int64_t total_tiles, cdae_stream s);
/* nearest x2 upsample NHWC bf16 (unet.py:69-79 F.interpolate) and its adjoint (2x2 sum pool) */
int cdae_upsample2x(const void* x, void* out, int N, int H, int W, int C, cdae_stream s);
int cdae_sumpool2x(const void* dy, void* dx, int N, int H, int W, int C, int accumulate, cdae_stream s);
/* zero-insertion (adjoint gather of a stride-2 conv): out[n,2h,2w,:] = x[n,h,w,:], zeros elsewhere */
int cdae_zero_insert2x(const void* x, void* out, int N, int H, int W, int C, cdae_stream s);
/* column sums of `groups` consecutive [rows, C] bf16 matrices (pitch ld) accumulated into fp32 out[g * out_ld + c]
 * (bias gradients; per-image sums = the gradient of the additive timestep conditioning) */
int cdae_colsum(const void* x_bf16, float* out, int64_t rows, int C, int ld, int groups, int out_ld, cdae_stream s);
/* inverted dropout in place on a bf16 tensor (unet.py:157 nn.Dropout between SiLU and the ResBlock's second conv): element i
 * is kept (and scaled by 1/(1-p)) iff the counter-based generator says so for (seed, base + layer_offset, i); p is read from
 * device memory (0 = identity: eval mode) and the SAME launch on the gradient is the backward.  state: device uint64
 * {seed, base}; n % 8 == 0. */
int cdae_dropout(void* x_bf16, int64_t n, const void* state, int64_t layer_offset, const float* p, cdae_stream s);

/* ------------------------------------------------------------------ GroupNorm32 (+FiLM) (+SiLU)   nn.py:430-437, unet.py:185-198 */
/* x = concat(x0[C0], x1[C1]) NHWC bf16 per sample (x1 may be NULL). y = act(GN(x)*gamma+beta)*(1+scale)+shift ...
 * precisely: u = (xhat*gamma+beta)*(1+scale[b,c]) + shift[b,c]; y = silu ? u*sigmoid(u) : u.  stats -> mean/rstd [B,32].
 * film: fp32 [B, film_ld] with scale at col film_off + c and shift at film_off + C + c (NULL: no FiLM). */
int cdae_gn_fwd(const void* x0, int C0, const void* x1, int C1, int B, int HW,
                const float* gamma, const float* beta, const float* film, int film_ld, int film_off,
                int silu, void* y, float* mean, float* rstd, cdae_stream s);
/* same forward when the producing convolutions already accumulated the per-(image, channel) sums (cdae_igemm_desc.stats):
 * stats0 fp32 [B][C0][2], stats1 fp32 [B][C1][2] = {sum, sum of squares} over the HW pixels.  One streaming pass
 * (2 B read + 2 B written per element), no reduction; mean/rstd [B,32] are still written for the backward.  ab (optional):
 * fp32 [B][C][2] receives the per-(image, channel) constants {a, b} with u (u/2 under SiLU) = a x + b, which the fused
 * backward statistics of the data-gradient convolution read (cdae_igemm_desc.gnb_ab). */
int cdae_gn_apply_fwd(const void* x0, int C0, const float* stats0, const void* x1, int C1, const float* stats1,
                      int B, int HW, const float* gamma, const float* beta, const float* film, int film_ld, int film_off,
                      int silu, void* y, float* mean, float* rstd, float* ab, cdae_stream s);
/* backward: dx split into dx0/dx1; accumulate_dx bit0/bit1: add to the existing contents of dx0/dx1;
 * dadd: optional bf16 [B,HW,C] tensor added to dx (identity-skip gradient of a ResBlock, unet.py:198);
 * dgamma/dbeta (+=, fp32), dfilm (+= into [B, film_ld] at the same columns) */
int cdae_gn_bwd(const void* dy, const void* x0, int C0, const void* x1, int C1, int B, int HW,
                const float* gamma, const float* beta, const float* film, int film_ld, int film_off, int silu,
                const float* mean, const float* rstd, const void* dadd, void* dx0, void* dx1, int accumulate_dx,
                float* dgamma, float* dbeta, float* dfilm, cdae_stream s);

/* the backward as ONE streaming pass, for a GroupNorm whose output gradient was produced by a data-gradient convolution with
 * cdae_igemm_desc.gnb_* set: du = dy * silu'(u) (bf16, same layout as dy) and ws fp32 [B][C][2] = {sum du, sum du*x} per
 * (sample, channel) come from that launch's epilogue; this pass reads du and x once and writes dx = K1 du - K2' - x K3'
 * (+ dadd) (+ old dx) plus dgamma / dbeta / dfilm.  6 B per element, no reduction, no clusters. */
int cdae_gn_bwd_apply(const void* du, const void* x0, int C0, const void* x1, int C1, int B, int HW,
                      const float* gamma, const float* beta, const float* film, int film_ld, int film_off,
                      const float* mean, const float* rstd, const float* ws, const void* dadd, void* dx0, void* dx1,
                      int accumulate_dx, float* dgamma, float* dbeta, float* dfilm, cdae_stream s);

/* ------------------------------------------------------------------ implicit-GEMM convolution on tcgen05/TMEM/TMA
 * Replaces aten::convolution (cuDNN) at unet.py:143-171 (ResBlock convs + 1x1 skip), :69-79, :97-105 (up/down),
 * :216-218 (attention qkv/proj Conv1d), :392, :498 and every nn.Linear that is GEMM shaped.
 *   out[pixel, co] = sum_seg sum_c src[seg.src][pixel shifted by (dh,dw), seg.c0 + c] * wgt[co, seg.wk + c]
 * Sources are NHWC bf16 with identical (N,H,W); a K "segment" is one filter tap of one source (64-channel chunks).
 * The same kernel does forward, data-gradient (transposed weights, negated taps) and plain GEMMs (1 tap).      */
#define CDAE_MAX_SEG 48
typedef struct { int32_t src, dh, dw, c0, nchunk, wk; } cdae_seg;
typedef struct {
  const void* src[4]; int32_t src_c[4];    /* source tensors and their channel counts (row pitch) */
  int32_t nsrc, N, H, W;                   /* source spatial dims */
  int32_t in_stride;                       /* 1, or 2: output pixel p reads input pixel 2p+tap (TMA element stride) */
  int32_t nseg; cdae_seg seg[CDAE_MAX_SEG];
  const void* wgt; int32_t wrows, wk;      /* bf16 [wrows][wk] */
  void* out; int32_t out_mode;             /* 0: NHWC bf16, 1: NCHW fp32, 2: fp32 row-major [pixel][ldo] (GEMM-shaped fp32 results) */
  int32_t OH, OW, ldo, cout;               /* full output dims, row pitch (NHWC), number of real output channels */
  int32_t sps, ooh, oow;                   /* output pixel = (tile_y*sps + ooh, tile_x*sps + oow) */
  const float* bias;                       /* fp32 [cout] or NULL */
  const float* bias2;                      /* second fp32 [cout] bias (fused 1x1 skip conv) or NULL */
  const void* resid; int32_t ldr;          /* bf16 tensor added in the epilogue (same pixel indexing as out) or NULL */
  int32_t bn;                              /* N tile: 16, 32, 64, 128 or 256 (0: auto) */
  float* stats;                            /* optional fp32 [N][cout][2] (16 B aligned), ACCUMULATED (+=): per-(image, channel)
                                              sum and sum of squares of the bf16 output as stored - the GroupNorm statistics of
                                              the consumer (cdae_gn_apply_fwd), produced by the conv epilogue so that the norm
                                              becomes one streaming pass.  Needs out_mode 0, cout % 64 == 0, OH*OW >= 32. */
  /* GroupNorm-backward fusion of a DATA-GRADIENT launch (backward twin of `stats`; nn.py:435-437 + unet.py:185-198 backward):
   * the output is the gradient w.r.t. the OUTPUT of y = [SiLU](FiLM(GroupNorm32(x))), x = concat(gnb_x0 [gnb_c0 channels, pitch
   * gnb_ld0], gnb_x1 [cout - gnb_c0, pitch gnb_ld1]) on the same pixel grid.  The epilogue fetches the x slab (residual TMA
   * path), stores du = dy * silu'(u) instead of dy (u/2 = a x + b, {a, b} = gnb_ab[image][channel][2] written by
   * cdae_gn_apply_fwd) and ACCUMULATES gnb_ws[image][channel][2] += {sum du, sum du*x}: cdae_gn_bwd_apply then is one
   * streaming pass.  Needs out_mode 0, no bias / residual / stats, cout and gnb_c0 % 64 == 0, OH*OW >= 32. */
  float* gnb_ws; const float* gnb_ab; const void* gnb_x0; const void* gnb_x1;
  int32_t gnb_c0, gnb_ld0, gnb_ld1, gnb_silu;
  /* per-(image, channel) additive term of the epilogue: out += bias_img[n * bias_img_ld + co] - the ResBlock's additive
   * timestep conditioning `h + emb_out[..., None, None]` (unet.py:196, use_scale_shift_norm=False).  NHWC output only. */
  const float* bias_img; int32_t bias_img_ld;
  /* GroupNorm(+FiLM)+SiLU applied to a SOURCE while it is loaded (inference: nn.py:430-437 + unet.py:185-198 folded into the
   * operand path of the conv that consumes the activation - no activated tensor is materialised).  gn_ab: fp32
   * [N][gn_c][2] = the per-(image, channel) constants {a, b} of cdae_gn_apply_fwd (u/2 = a x + b; y = u/2 + u/2 tanh(u/2)).
   * gn_off[i] >= 0: source i is GroupNorm input, its channel c uses table column gn_off[i] + c; -1: source i is taken as
   * it is (the 1x1-skip sources).  Zero padding stays zero (the reference pads the ACTIVATED tensor).  3x3 stride-1 layers
   * with cout % 128 == 0 on images that tile into 8 x 32 boxes (the transposed halo kernel); bit-identical to running
   * cdae_gn_apply_fwd first. */
  const float* gn_ab; int32_t gn_c; int32_t gn_off[4];
  /* up2x != 0: the sources are [N, H, W, C] at HALF the output resolution and the conv runs over their nearest-neighbour x2
   * upsampling (Upsample: F.interpolate(scale_factor=2, mode="nearest") + conv3x3, unet.py:69-79) - expanded from a
   * low-resolution halo tile in shared memory, the 4x tensor is never written.  Output [N, 2H, 2W, cout]; 3x3 stride-1
   * layers with cout % 128 == 0 whose OUTPUT tiles into 8 x 32 boxes (the transposed halo kernel). */
  int32_t up2x;
} cdae_igemm_desc;
int cdae_igemm(const cdae_igemm_desc* d, cdae_stream s);

/* weight gradient:  dW[co, tap, ci] += sum_pixels dy[pixel, co] * src[pixel*in_stride + tap, c0 + ci]
 * (aten::convolution_backward wgrad). dy NHWC bf16 [N,OH,OW,ldy]; fp32 accumulation with red.add into dw. */
typedef struct {
  const void* dy; int32_t ldy, cout;       /* cout real rows (<= padded rows present in dy pitch) */
  const void* src; int32_t src_c;          /* source tensor, row pitch */
  int32_t c0, cin;                         /* channel window of the source that maps to dW columns [ci_off, ci_off+cin) */
  int32_t N, H, W, OH, OW, in_stride;
  int32_t ksize;                           /* 1 or 3 */
  float* dw; int32_t dw_ld;                /* fp32 dW[co][tap][dw_ld] ; columns start at ci_off */
  int32_t ci_off, cin_real;                /* only ci < cin_real written */
  int32_t splits;                          /* split-K factor over pixel tiles (0: auto) */
  float* dbias;                            /* fp32 [cout] += column sums of dy (the bias gradient, fused as an N=8 MMA against ones), or NULL */
} cdae_wgrad_desc;
int cdae_wgrad(const cdae_wgrad_desc* d, cdae_stream s);

/* ------------------------------------------------------------------ fused QKV attention            unet.py:239-253
 * qkv: [B, T, 3*C] bf16, head h at channels [h*3*ch, (h+1)*3*ch) = [q_h | k_h | v_h]; out [B, T, C] bf16 (head h at
 * h*ch); softmax in fp32, scale ch^-1/4 on q and k; lse [B, heads, T] fp32 saved for the backward (may be NULL).
 * backward: dqkv in the same interleaved layout (recomputes the probabilities from lse); dsum [B, heads, T] fp32 is
 * caller-owned scratch (rowsum(dout*out), written by the dQ pass and read by the dK/dV pass). ch: multiple of 16, <= 128. */
int cdae_attn_fwd(const void* qkv, void* out, float* lse, int B, int T, int heads, int ch, cdae_stream s);
int cdae_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* dsum, void* dqkv,
                  int B, int T, int heads, int ch, cdae_stream s);

/* ------------------------------------------------------------------ causal DAG mask layer   nn.py:225-240,290-312; unet.py:571-583
 * z_pre[b,i,:] = sum_j A[j,i] u[b,j,:];  z_post[b,i,:] = W2_i leaky_relu(W1_i z_pre[b,i,:] + b1_i) + b2_i + u[b,i,:]
 * u, z_post, dzpost, du: fp32 [B, n, d]; A: fp32 [n, n] row-major; params / grads: DEVICE arrays of 4n pointers
 * {W1_i [D,d], b1_i [D], W2_i [d,D], b2_i [d]} (the reference keeps one MLP per causal variable).  n <= 8.
 * backward: recomputes the hidden layer; parameter gradients are ACCUMULATED (+=) into grads; dzp_ws is a caller-owned
 * fp32 [B, n, d] workspace that must be zero on entry and is left zero on exit; du_add: optional fp32 [B, n, d] added to du
 * (the direct KL gradient w.r.t. mu); d in {64, 128, 256}, D % 32 == 0. */
int cdae_dag_fwd(const float* u, const float* A, const void* const* params, float* zpost, int B, int n, int d, int D,
                 cdae_stream s);
int cdae_dag_bwd(const float* u, const float* A, const void* const* params, const float* dzpost, void* const* grads,
                 float* dzp_ws, float* du, const float* du_add, int B, int n, int d, int D, cdae_stream s);

/* ------------------------------------------------------------------ the [B, 512]-sized representation path in fp32
 * (timestep embedding trunk, FiLM projections, GaussianConvEncoder, reparameterisation / mask / KL).  Replaces, per
 * training step, ~250 ATen / cuBLAS-SIMT / cuDNN launches of nn.py:15-110,440-467,551-569, unet.py:545-616,
 * gaussian_diffusion.py:718-766 and their autograd.  fp32 CUDA-core math (<= 1e-4 vs the oracle): the path is 0.03 % of the
 * step's FLOPs, what it costs is launches.
 *
 * cdae_sgemm: C[M,N] (=, +=, scatter+=) A[M,K] * B[K,N] with pluggable operand views:
 *   a_mode 0 dense A[m*a_sm + k*a_sk] | 1 dense with SiLU applied on load | 2 im2col view of a 3x3 stride-2 pad-1 conv
 *   (m = (b, oh, ow), k = (ci, kh, kw) - the OIHW weight order), source strides g_s*, optional per-input-channel {a, b} table:
 *   the producer's BatchNorm + LeakyReLU(0.01) applied on load (zero padding AFTER the activation, like the reference) |
 *   3 the same view transposed (m = conv k, k = conv pixel: weight gradients);
 *   b_mode 0 dense B[k*b_sk + n*b_sn] | 1 with SiLU on load;
 *   c_mode 0 store C[m*c_sm + n*c_sn] | 1 atomic += | 2 col2im scatter += into the g_s* tensor (data gradient; zero it first);
 *   bias[n] added once; act_out 1 = softplus(v) + 1e-8 (nn.py:108); colstats[n][2] (double) += column sum / sum of squares
 *   of C as stored (the BatchNorm batch statistics of the conv output).  splits: K split factor over grid.z (0 = auto; > 1
 *   needs c_mode 1 or 2). */
typedef struct {
  const float* A; int64_t a_sm, a_sk; int32_t a_mode;
  const float* B; int64_t b_sk, b_sn; int32_t b_mode;
  float* C; int64_t c_sm, c_sn; int32_t c_mode;
  const float* bias; int32_t act_out;
  double* colstats;
  int32_t M, N, K, splits;
  int64_t g_sb, g_sc, g_sh, g_sw; int32_t g_cin, g_h, g_w, g_oh, g_ow; const float* g_ab;
} cdae_sgemm_desc;
int cdae_sgemm(const cdae_sgemm_desc* d, cdae_stream s);
/* timestep_embedding (nn.py:551-569): out[b] = [cos(t_b f) | sin(t_b f)] (+ a zero column when dim is odd); freqs: device
 * fp32 [dim/2] computed on the host exactly like the reference; t int64 or (t_is_float) fp32; map (optional): respaced ->
 * original step table applied first, scale (0 = off): rescale_timesteps factor 1000/T (respace.py:119-124); t_stride 0 = one
 * timestep for the whole batch (a device scalar: the sampling loop's counter) */
int cdae_timestep_embedding(const void* t, int t_is_float, int t_stride, const int64_t* map, float scale, const float* freqs,
                            float* out, int B, int dim, cdae_stream s);
/* device-side step counters of a graph-replayed sampling loop (gaussian_diffusion.py:632-680 `for i in indices`): both
 * (either may be NULL) += delta */
int cdae_step_tick(int64_t* step64, int32_t* step32, int delta, cdae_stream s);
/* counter-based normal / Bernoulli draws (Philox4x32-10 + Box-Muller), replayable inside a CUDA graph: state = device
 * uint64 {seed, offset}; the launch advances the offset.  Replaces th.randn_like / th.bernoulli at gaussian_diffusion.py:790,
 * nn.py:464, unet.py:601 in throughput runs (parity runs inject the reference's CPU-generator draws instead). */
int cdae_randn(float* out, int64_t n, void* state, int bernoulli, float keep_prob, cdae_stream s);
/* y (bf16) = silu ? SiLU(x) : x  - the bf16 A operand of the tensor-core FiLM projection (unet.py:148-154 emb_layers) */
int cdae_silu_cast(const float* x, void* y_bf16, int64_t n, int silu, cdae_stream s);
/* g *= silu'(x)  (backward of the SiLU-on-load operand views) ; g *= softplus'(pre) given var = softplus(pre) + 1e-8 */
int cdae_silu_bwd(float* g, const float* x, int64_t n, cdae_stream s);
int cdae_softplus_bwd(float* g, const float* var, int64_t n, cdae_stream s);
/* x[b,:] += table[idx[b],:] (label_emb, unet.py:549-551); backward: dtable[idx[b],:] += x[b,:] */
int cdae_embed_rows(float* x, const float* table, const int64_t* idx, int B, int D, int backward, float* dtable, cdae_stream s);
/* nn.BatchNorm2d bookkeeping of one encoder layer (nn.py:52-61; eps 1e-5, momentum 0.1): train = batch statistics from the
 * double column sums (count = B*OH*OW) and running-buffer update, else running statistics -> ab[c] = {a, b} with
 * bn(x) = a x + b, mean_rstd[c] = {mean, rstd} */
int cdae_bn_finalize(const double* stats, double count, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, int64_t* num_batches_tracked, int train, float* ab, float* mean_rstd, int C,
                     cdae_stream s);
/* last encoder layer: BatchNorm + LeakyReLU + th.flatten (nn.py:104): NHWC raw [B,P,C] -> [B, C*P]; backward = the permutation */
int cdae_enc_head(const float* in, const float* ab, float* out, int B, int P, int C, int backward, cdae_stream s);
/* BatchNorm(batch stats) + LeakyReLU backward over NHWC [M, C]: dact (gradient w.r.t. the activation) is overwritten by the
 * gradient w.r.t. the raw conv output; dgamma / dbeta += ; sums: double [C][2] scratch */
int cdae_bn_lrelu_bwd(float* dact, const float* raw, const float* ab, const float* mean_rstd, const float* gamma,
                      double* sums, float* dgamma, float* dbeta, int64_t M, int C, cdae_stream s);
/* reparameterize + classifier-free keep mask + closed-form KL of representation_loss (nn.py:440-467, unet.py:590-613,
 * gaussian_diffusion.py:718-766), one launch:  z = (zp + sqrt(var_scale*var) xi) keep, zp_out = zp keep,
 * kld[b] = KL(N(mu,var) || N(0,1)) + [causal] 0.5 sum_i |zp_out_i - c_i|^2.  keep / c / zp_out / kld may be NULL. */
int cdae_latent_fwd(const float* mu, const float* var, const float* zp, const float* xi, const float* keep, const float* c,
                    float* z, float* zp_out, float* kld, int B, int D, int n, int causal, float var_scale, cdae_stream s);
/* its backward: dz [B,D] and dkld [B] (either may be NULL) plus optional external gradients -> dzp, dmu (direct part), dvar */
int cdae_latent_bwd(const float* mu, const float* var, const float* zp, const float* xi, const float* keep, const float* c,
                    const float* dz, const float* dkld, const float* dzp_ext, const float* dmu_ext, const float* dvar_ext,
                    float* dzp, float* dmu, float* dvar, int B, int D, int n, int causal, float var_scale, cdae_stream s);
/* loss assembly of a training step (gaussian_diffusion.py:849-855, train_util.py:262-266): loss[b] = mse[b] + kl_weight *
 * kld_rep (kld_rep = kld[b], or sum(kld keep)/sum(keep) when keep != NULL), total = mean(loss w); emits the gradient scales
 * gscale[b] = w[b]/B (for cdae_mse_loss) and dkld[b] (for cdae_latent_bwd), and (optional) the logger's running sums
 * logsums[20] += {loss w, mse w, kld w, count, 3x4 per-timestep-quartile sums, 4 quartile counts} (train_util.py:401-407) */
int cdae_step_loss(const float* mse, const float* kld, const float* keep, const float* w, const float* kl_weight,
                   const int64_t* t, int num_timesteps, int B, float* loss, float* gscale, float* dkld, float* total,
                   float* logsums, cdae_stream s);

/* ------------------------------------------------------------------ HBM-resident dataset batch assembly
 * (the step before the hot path: image_datasets.py:141-183,241-296,344-392,411-483 = PIL decode + ToTensor + DataLoader
 * collate per item).  images_u8: uint8 [n][H][W][C] resident in HBM (16 B aligned, image stride H*W*C a multiple of 4),
 * labels: fp32 [n][L] (L may be 0), idx: int64 [B] row numbers.  out[b,c,h,w] = images[idx[b],h,w,c] / 255 (fp32 NCHW,
 * IEEE division = torchvision ToTensor, bit-exact; mode 1: u / 127.5 - 1 = ImageDataset :171), out_labels[b,:] =
 * labels[idx[b],:].  C in 1..4, H*W % 4 == 0. */
int cdae_gather_images(const void* images_u8, const float* labels, const int64_t* idx, float* out, float* out_labels,
                       int B, int H, int W, int C, int L, int mode, cdae_stream s);

#ifdef __cplusplus
}
#endif
#endif
