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
 * mse[b] = mean_chw (target-pred)^2 ; if dpred != NULL: dpred = 2*(pred-target)*gscale[b]/per_sample */
int cdae_mse_loss(const float* pred, const float* target, float* mse, const float* gscale, float* dpred,
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
 * train_util.py:277-280); gsq_out += sum((g*grad_scale)^2). */
int cdae_adam_ema(float* p, const void* g, int g_is_bf16, float* m, float* v, float* ema, const float* hyper,
                  int64_t* step, const float* guard, float* gsq_out, int64_t n, cdae_stream s);
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
/* column sums of a [rows, C] bf16 matrix accumulated into fp32 (bias gradients) */
int cdae_colsum(const void* x_bf16, float* out, int64_t rows, int C, int ld, cdae_stream s);

/* ------------------------------------------------------------------ GroupNorm32 (+FiLM) (+SiLU)   nn.py:430-437, unet.py:185-198 */
/* x = concat(x0[C0], x1[C1]) NHWC bf16 per sample (x1 may be NULL). y = act(GN(x)*gamma+beta)*(1+scale)+shift ...
 * precisely: u = (xhat*gamma+beta)*(1+scale[b,c]) + shift[b,c]; y = silu ? u*sigmoid(u) : u.  stats -> mean/rstd [B,32].
 * film: fp32 [B, film_ld] with scale at col film_off + c and shift at film_off + C + c (NULL: no FiLM). */
int cdae_gn_fwd(const void* x0, int C0, const void* x1, int C1, int B, int HW,
                const float* gamma, const float* beta, const float* film, int film_ld, int film_off,
                int silu, void* y, float* mean, float* rstd, cdae_stream s);
/* same forward when the producing convolutions already accumulated the per-(image, channel) sums (cdae_igemm_desc.stats):
 * stats0 fp32 [B][C0][2], stats1 fp32 [B][C1][2] = {sum, sum of squares} over the HW pixels.  One streaming pass
 * (2 B read + 2 B written per element), no reduction; mean/rstd [B,32] are still written for the backward. */
int cdae_gn_apply_fwd(const void* x0, int C0, const float* stats0, const void* x1, int C1, const float* stats1,
                      int B, int HW, const float* gamma, const float* beta, const float* film, int film_ld, int film_off,
                      int silu, void* y, float* mean, float* rstd, cdae_stream s);
/* backward: dx split into dx0/dx1; accumulate_dx bit0/bit1: add to the existing contents of dx0/dx1;
 * dadd: optional bf16 [B,HW,C] tensor added to dx (identity-skip gradient of a ResBlock, unet.py:198);
 * dgamma/dbeta (+=, fp32), dfilm (+= into [B, film_ld] at the same columns) */
int cdae_gn_bwd(const void* dy, const void* x0, int C0, const void* x1, int C1, int B, int HW,
                const float* gamma, const float* beta, const float* film, int film_ld, int film_off, int silu,
                const float* mean, const float* rstd, const void* dadd, void* dx0, void* dx1, int accumulate_dx,
                float* dgamma, float* dbeta, float* dfilm, cdae_stream s);

/* the same backward as two streaming passes (reduce: P_c = sum du, Qx_c = sum du*x per (sample, channel) into ws; apply:
 * dx, walking the tensor in reverse so that the second read of dy / x hits the L2).  ws: caller-owned fp32 [B][2][C],
 * ZERO on entry (it is accumulated into with red.add and left holding the sums). */
int cdae_gn_bwd_stream(const void* dy, const void* x0, int C0, const void* x1, int C1, int B, int HW,
                       const float* gamma, const float* beta, const float* film, int film_ld, int film_off, int silu,
                       const float* mean, const float* rstd, const void* dadd, void* dx0, void* dx1, int accumulate_dx,
                       float* dgamma, float* dbeta, float* dfilm, float* ws, cdae_stream s);

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
  void* out; int32_t out_mode;             /* 0: NHWC bf16, 1: NCHW fp32 */
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
 * fp32 [B, n, d] workspace that must be zero on entry and is left zero on exit; d in {64, 128, 256}, D % 32 == 0. */
int cdae_dag_fwd(const float* u, const float* A, const void* const* params, float* zpost, int B, int n, int d, int D,
                 cdae_stream s);
int cdae_dag_bwd(const float* u, const float* A, const void* const* params, const float* dzpost, void* const* grads,
                 float* dzp_ws, float* du, int B, int n, int d, int D, cdae_stream s);

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
