/* coma_b200 -- C ABI of the B200-native CoMA-UNet hot path (libcoma_b200.so).
 *
 * The reference (mborhi/CoMA-UNet) is pure Python over PyTorch/cuDNN: it has NO native
 * boundary of its own.  This header is the boundary SURVEY.md section 8(b) defines for the
 * path; each entry point names the reference code it replaces (file:line into the reference).
 * The Python host (coma_unet_b200/*.py) binds it with ctypes and keeps the reference's
 * module / criterion / dataset API above it.
 *
 * Conventions
 *  - all tensors are raw DEVICE pointers, activations in NDHWC ("channels last 3d") order:
 *    element (b, d, h, w, c) of a buffer with channel stride cs and channel offset co lives at
 *    ((((b*D + d)*H + h)*W + w) * cs + co + c).  cs/co let a kernel read or write one half of a
 *    concat buffer in place (replaces torch.cat, attn_unet_data_parallel.py:229).
 *  - dtype is the storage type of activations and packed weights (COMA_F32 or COMA_BF16);
 *    accumulation, statistics, scale/shift vectors, losses and weight gradients are fp32.
 *  - every call is asynchronous on `stream`, allocates nothing, keeps no global mutable state
 *    besides a mutex-protected cache of TMA descriptors, and returns 0 on success
 *    (COMA_ERR_* otherwise; coma_last_error() gives the message for the calling thread).
 *  - nothing here falls back to the CPU: without a CUDA device every compute call fails.
 */
#ifndef COMA_B200_H_
#define COMA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* coma_stream_t;

enum { COMA_OK = 0, COMA_ERR_INVALID = 1, COMA_ERR_CUDA = 2, COMA_ERR_UNSUPPORTED = 3, COMA_ERR_WORKSPACE = 4 };
enum { COMA_F32 = 0, COMA_BF16 = 1,
       /* conv fprop / dgrad only: x and w bf16, y stored in fp32 (the accumulator is not rounded).  The split-precision path of the
        * Python host (ops.conv_raw with fp32 tensors, `fp32_tensor_cores=True`): x = x_hi + x_lo and w = w_hi + w_lo in bf16,
        * conv(x, w) ~= conv([x_hi | x_lo | x_hi], [w_hi | w_hi | w_lo]) -- fp32-level accuracy (2^-17 per operand) on tcgen05 */
       COMA_BF16_F32OUT = 2 };
enum { COMA_ACT_NONE = 0, COMA_ACT_RELU = 1, COMA_ACT_LEAKY = 2 /* PReLU(1 param) and LeakyReLU */, COMA_ACT_SIGMOID = 3,
       COMA_ACT_LEAKY_RELU = 4 /* ReLU(PReLU(u)): final_pred_head + final_act, attn_unet_data_parallel.py:654-656 */ };
enum { COMA_NORM_NONE = 0, COMA_NORM_INSTANCE = 1, COMA_NORM_BATCH = 2, COMA_NORM_GIVEN = 3 /* eval BN: running stats */ };
enum { COMA_IMPL_AUTO = 0, COMA_IMPL_SIMT = 1, COMA_IMPL_TCGEN05 = 2 };

int coma_version(void);
const char* coma_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Convolutions.  Replace torch.nn.Conv3d / ConvTranspose3d (cuDNN) inside MONAI `Convolution`
 * (attn_unet_data_parallel.py:20,126,285-306,442,495-497,546-558 and MONAI attentionunet's
 * ConvBlock / UpConv / AttentionLayer.merge).
 *
 * Packed weight layout for every conv kernel: w[tap][Cout][Cin] (tap = (kd*k + kh)*k + kw),
 * same dtype as the activations.  For the transposed conv the tap indexes the ConvTranspose3d
 * kernel position.  `w_bstride` != 0 selects per-sample weights (covariate-routed expert
 * mixture, attn_unet_data_parallel.py:296-306).
 *
 * Fused epilogue (all optional): v = conv + bias;  stats += (v, v*v) per (sample, channel)
 * [feeds the following Instance/BatchNorm];  y = act(scale[b,c]*v + shift[b,c]).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;        /* input activations  */
  const void* w;        /* packed weights     */
  const float* bias;    /* [Cout] fp32 or NULL */
  void* y;              /* output activations */
  const float* scale;   /* [B, Cout] fp32 or NULL */
  const float* shift;   /* [B, Cout] fp32 or NULL */
  const float* slope;   /* device pointer to the negative slope (COMA_ACT_LEAKY) or NULL */
  float* stats;         /* [B, chunks, Cout, 2] fp32 partial sums or NULL; chunks = coma_conv3d_stat_chunks() */
  int32_t B, Di, Hi, Wi, Do, Ho, Wo; /* input and output spatial extents */
  int32_t Cin, Cout;
  int32_t x_cs, x_co;   /* channel stride / offset of the x buffer */
  int32_t y_cs, y_co;   /* channel stride / offset of the y buffer */
  int32_t y_cn;         /* channels actually stored (<= Cout; the rest is MMA padding) */
  int32_t ksize, stride, pad;
  int32_t transposed;   /* 0: Conv3d, 1: ConvTranspose3d (output_padding = stride-1) */
  int64_t w_bstride;    /* elements between per-sample weight sets, 0 = shared */
  int32_t bias_bstride; /* elements between per-sample biases, 0 = shared */
  int32_t act;
  int32_t dtype;
  int32_t impl;         /* COMA_IMPL_* */
  /* Optional input prologue (all NULL / 0 = off): the convolution reads x' = act_in(in_scale[b,ci] * x + in_shift[b,ci]) for
   * every in-bounds input voxel (the zero padding stays zero).  A consumer conv thereby absorbs the Instance/BatchNorm +
   * FiLM + activation that follows its producer (MONAI ADN, attn_unet_data_parallel.py:285-286) and the normalised tensor
   * is never written.  in_act is COMA_ACT_NONE / RELU / LEAKY (slope read from in_slope). */
  const float* in_scale; /* [B, Cin] fp32 or NULL */
  const float* in_shift; /* [B, Cin] fp32 or NULL */
  const float* in_slope; /* device pointer to the negative slope or NULL */
  int32_t in_act;
  int32_t reserved0;
} coma_conv_args;

int coma_conv3d_stat_chunks(const coma_conv_args* a);
/* 1 if the tcgen05/TMA implicit-GEMM path can take this problem (bf16, channel multiples of 16, ...) */
int coma_conv3d_tcgen05_supported(const coma_conv_args* a);
/* the COMA_IMPL_* this problem actually runs on (what COMA_IMPL_AUTO resolves to: few-channel pointwise problems stream on the
 * CUDA-core / mma.sync kernels even where the tcgen05 tile kernel could take them) */
int coma_conv3d_impl(const coma_conv_args* a);
/* 1 if a fused kernel (not the generic CUDA-core gather) applies the input prologue of this problem */
int coma_conv3d_prologue_supported(const coma_conv_args* a);
int coma_conv3d_fprop(const coma_conv_args* a, coma_stream_t stream);
/* ConvTranspose3d forward (UpBlock.up, attn_unet_data_parallel.py:120-131); same struct, transposed=1 */
int coma_convT3d_fprop(const coma_conv_args* a, coma_stream_t stream);
/* Data gradients (autograd's cuDNN dgrad, attn_unet_data_parallel.py:884).  The host passes the
 * adjoint problem: dgrad of a stride-1 conv is a conv with flipped/transposed packed weights, dgrad of a
 * stride-2 conv is the transposed conv and vice versa; these entry points check that pairing. */
int coma_conv3d_dgrad(const coma_conv_args* adjoint, coma_stream_t stream);
int coma_convT3d_dgrad(const coma_conv_args* adjoint, coma_stream_t stream);

/* Weight gradient in conv geometry (i = o*stride + k - pad):
 *   dw[tap][Cg][Cx] (fp32, ACCUMULATED into) += sum_o g[o][cg] * x[i][cx]
 * Conv3d:            g = dy (output grid), x = layer input  -> dw is the packed [tap][Cout][Cin].
 * ConvTranspose3d:   g = layer input (small grid), x = dy (large grid) -> dw is [tap][Cin_t][Cout_t]. */
typedef struct {
  const void* g; const void* x; float* dw;
  int32_t B, Dg, Hg, Wg, Dx, Hx, Wx;
  int32_t Cg, Cx;
  int32_t g_cs, g_co, x_cs, x_co;
  int32_t ksize, stride, pad;
  int32_t dtype;
  int32_t impl;
  /* Optional (NULL / 0 = off): caller-owned, 16-byte aligned scratch of at least coma_conv3d_wgrad_workspace_size() bytes.  With it the tcgen05
   * kernels write one partial [27][Cg][Cx] block per CTA and a second kernel sums the blocks in a fixed order and STORES dw
   * (dw need not be zeroed): bit-identical results from run to run.  Without it the CTAs add into a zeroed dw with fp32 atomics
   * (order-dependent low bits). */
  void* workspace; int64_t workspace_bytes;
} coma_wgrad_args;
/* 1 if a tcgen05 weight-gradient kernel takes this problem: bf16, k3, and either stride 1 (channels multiples of 16, W % 32 == 0
 * or W == 16) or stride 2 (channels multiples of 32, coarse W % 32 == 0 && H % 4 == 0 or W % 16 == 0 && H % 8 == 0) */
int coma_conv3d_wgrad_tcgen05_supported(const coma_wgrad_args* a);
/* bytes of scratch the deterministic path of coma_conv3d_wgrad / coma_convT3d_wgrad wants for this problem (0: no such path) */
int64_t coma_conv3d_wgrad_workspace_size(const coma_wgrad_args* a);
int coma_conv3d_wgrad(const coma_wgrad_args* a, coma_stream_t stream);
int coma_convT3d_wgrad(const coma_wgrad_args* a, coma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Normalisation + covariate modulation (FiLM) + activation.  Replaces BatchNorm3d / InstanceNorm3d
 * + PReLU/ReLU/LeakyReLU kernels of MONAI `ADN` and the missing CondConv modulation
 * (attn_unet_data_parallel.py:126,285-286; oracle/cond_conv.py):
 *     xhat = (x - mean) * rstd          mean/rstd per (b,c) [INSTANCE], per c [BATCH/GIVEN]
 *     y    = act(g[b,c] * xhat + h[b,c])
 * ------------------------------------------------------------------------------------------- */
/* per-(sample, chunk, channel) partial sums (sum x, sum x^2) of an NDHWC buffer */
int coma_norm_stats_chunks(int64_t V);
int coma_norm_stats(const void* x, int32_t B, int64_t V, int32_t C, int32_t cs, int32_t co, int32_t dtype,
                    float* partial /* [B, chunks, C, 2] */, coma_stream_t stream);

typedef struct {
  const float* partial; int32_t chunks;  /* [B, chunks, C, 2] (unused for GIVEN / NONE) */
  int32_t B, C; int64_t V;               /* V = voxels per sample */
  int32_t mode;                          /* COMA_NORM_* */
  const float* given_mean; const float* given_var;  /* [C], mode GIVEN */
  float eps;
  const float* g; const float* h;        /* [B, C] or NULL (g=1, h=0) */
  float* A; float* S;                    /* out [B, C]: y = act(A*x + S) */
  float* mean; float* rstd;              /* out [B, C] (saved for backward) */
  float* running_mean; float* running_var;  /* [C] or NULL; BATCH only */
  float momentum; int32_t n_updates;     /* n_updates=2 applies the reference's duplicated forward
                                            (attn_unet_data_parallel.py:664,666) in closed form */
} coma_norm_finalize_args;
int coma_norm_stats_finalize(const coma_norm_finalize_args* a, coma_stream_t stream);

typedef struct {
  const void* x; void* y;
  const float* A; const float* S;   /* [B, C] */
  const float* slope;               /* device scalar or NULL */
  int32_t B, C; int64_t V;
  int32_t x_cs, x_co, y_cs, y_co;
  int32_t act, dtype;
  const void* r; int32_t r_cs;      /* optional residual added before the activation: y = act(A*x + S + r) */
} coma_affine_act_args;
int coma_norm_film_act_fwd(const coma_affine_act_args* a, coma_stream_t stream);

typedef struct {
  const void* x; const void* dy; void* dx;
  const float* A; const float* S; const float* mean; const float* rstd; const float* g; /* [B, C]; g may be NULL */
  const float* slope;
  int32_t B, C; int64_t V;
  int32_t x_cs, x_co, dy_cs, dy_co, dx_cs, dx_co;
  int32_t act, mode, dtype;
  float* partial;    /* workspace [B, chunks, C, 3] */
  float* dg; float* dh;   /* out [B, C] */
  float* dslope;          /* out [1] (accumulated into) or NULL */
  float* coef;            /* workspace [B, C, 3] */
  const void* r; int32_t r_cs;   /* the forward residual (or NULL) */
  void* dr; int32_t dr_cs;       /* out: gradient of the residual (or NULL) */
} coma_affine_act_bwd_args;
int coma_norm_film_act_bwd(const coma_affine_act_bwd_args* a, coma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Additive attention gate.  Replaces ObservableAttentionBlock.forward
 * (attn_unet_data_parallel.py:139-150): out = x * sigmoid(psi(relu(W_g g + W_x x))).
 * coma_gate_fwd is the single fused kernel for folded (eval-mode) BatchNorm; in training the gate is
 * composed of coma_conv3d_* (k=1), coma_norm_*, and the two elementwise kernels below, with
 * coma_gate_stats giving the batch statistics of the three BatchNorms.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* g; const void* x; void* out; void* psi_out /* optional [B,V] attention coefficients */;
  const float* wg; const float* wx; /* [F, C] fp32, BatchNorm folded in */
  const float* bsum;                /* [F]   folded bias of W_g + W_x */
  const float* wpsi;                /* [F]   folded psi weights */
  float bpsi;                       /* folded psi bias (host value when bpsi_ptr NULL) */
  const float* bpsi_ptr;            /* device scalar, preferred (no host sync) */
  int32_t B, C, F; int64_t V;
  int32_t g_cs, g_co, x_cs, x_co, out_cs, out_co;
  int32_t dtype;
} coma_gate_args;
int coma_gate_fwd(const coma_gate_args* a, coma_stream_t stream);

/* out[b,v,c] = x[b,v,c] * p[b,v]   and its backward (dx = dout*p, dp = sum_c dout*x) */
typedef struct {
  const void* x; const void* p; void* out;
  const void* dout; void* dx; void* dp;
  int32_t B, C; int64_t V;
  int32_t x_cs, x_co, out_cs, out_co;
  int32_t dtype;
} coma_bcast_mul_args;
int coma_gate_apply_fwd(const coma_bcast_mul_args* a, coma_stream_t stream);
int coma_gate_bwd(const coma_bcast_mul_args* a, coma_stream_t stream);
/* statistics of a k=1 conv output without materialising it are not needed by the composed path;
 * coma_gate_stats is coma_norm_stats under the name SURVEY.md 8(b) lists. */
int coma_gate_stats(const void* x, int32_t B, int64_t V, int32_t C, int32_t cs, int32_t co, int32_t dtype,
                    float* partial, coma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ROI painting + prompt selection.  Replaces the 72*B masked index_put_ + .item() loop of
 * forward_modulator_with_uq (attn_unet_data_parallel.py:632-649): writes the NDHWC buffer
 * [prompt(pos|neg by covariate[b,0]==1), saliency, suvr, 0...] that feeds deep_modulator_3c.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* roi;      /* [B, V] fp32 label volume (FreeSurfer ids) */
  const float* mri;      /* [B, V] fp32 model input x (voxels with x < 1e-4 are zeroed) */
  const float* lut;      /* [B, n_roi, 2] fp32 (loc, std) */
  const int32_t* roi_ids;/* [n_roi] */
  const float* is_pos;   /* [B] fp32, 1.0 selects the positive prompt */
  const float* pos_prompt; const float* neg_prompt; /* [V] fp32 */
  void* out;             /* [B, V, out_cs] */
  int32_t B, n_roi, out_cs; int64_t V;
  int32_t dtype;
} coma_roi_paint_args;
int coma_roi_paint(const coma_roi_paint_args* a, coma_stream_t stream);

/* dst[b,v,0] = a[b,v] + (a_add ? a_add[v] : 0); dst[b,v,1] = b ? b[b,v] : 0; dst[b,v,2..] = 0   (replaces the
 * `general_prompt + ...` add and the two torch.cat calls at attn_unet_data_parallel.py:651,654) */
typedef struct {
  const void* a; const float* a_add; const void* b; void* dst;
  int32_t B, dst_cs; int64_t V; int32_t dtype;
} coma_pack2_args;
int coma_pack2_fwd(const coma_pack2_args* a, coma_stream_t stream);
/* backward: da[b,v] = ddst[b,v,0]; db[b,v] = ddst[b,v,1]; d_a_add[v] = sum_b ddst[b,v,0] (fp32, overwritten) */
typedef struct {
  const void* ddst; void* da; void* db; float* d_a_add;
  int32_t B, dst_cs; int64_t V; int32_t dtype;
} coma_unpack2_args;
int coma_pack2_bwd(const coma_unpack2_args* a, coma_stream_t stream);
/* gradient of the painted buffer's prompt channel: dpos[v] = sum_b is_pos[b]*dbuf[b,v,0], dneg likewise */
int coma_roi_paint_bwd(const void* dbuf, const float* is_pos, float* dpos, float* dneg, int32_t B, int64_t V,
                       int32_t cs, int32_t dtype, coma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * RoiMSE (criterions.py:181-211, voxel_wise=False): loss[b] = mean_v(mask_b) * mean_v((pred-gt)^2),
 * mask[v] = w_i where roi[v] == id_i else 0.  One fused reduction; backward writes d pred.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* pred; const float* gt; const float* roi;
  const int32_t* roi_ids; const float* roi_w; int32_t n_roi;
  int32_t B; int64_t V; int32_t dtype;   /* dtype of pred / dpred */
  float* partial;        /* workspace [B, chunks, 2] */
  float* loss;           /* [B] */
  float* sums;           /* [B, 2]: (sum sq err, sum mask), saved for backward */
  const float* dloss;    /* [B] upstream gradient (backward) */
  void* dpred;           /* [B, V] (backward) */
} coma_roi_mse_args;
int coma_roi_mse_chunks(int64_t V);
int coma_roi_mse_fwd(const coma_roi_mse_args* a, coma_stream_t stream);
int coma_roi_mse_bwd(const coma_roi_mse_args* a, coma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Fused evaluation metrics: the step right after inference.  Replaces the batch metric lines
 * attn_unet_data_parallel.py:1214-1231 and calc_roi_metrics :1361-1397 (36 ROIs x ~10 masked full-volume kernels) by ONE
 * pass over pred / tau / roi.  Per sample and per slot (n_roi ROI slots in roi_ids order + one all-voxel slot, index n_roi)
 * eight fp64 sums are ACCUMULATED into `out` (the caller zeroes it), with d = pred - tau:
 *   [0] voxels  [1] sum |d|  [2] sum d^2  [3] sum tau  [4] sum tau^2
 *   [5] sum |d / tau| over the non-NaN ratios (an infinite ratio stays in the sum, like torch.nansum)  [6] NaN ratios (0/0)
 *   [7] sum 100 |d / tau| over |tau| > 1e-8
 * roi holds FreeSurfer label values as floats; labels must be integers in [0, 4096).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* pred;      /* [B, V] fp32 */
  const float* tau;       /* [B, V] fp32 */
  const float* roi;       /* [B, V] fp32 label values */
  const int32_t* roi_ids; /* [n_roi] device */
  int32_t n_roi, B;
  int64_t V;
  double* out;            /* [B, n_roi + 1, 8] fp64, accumulated into */
} coma_eval_metrics_args;
int coma_eval_metrics(const coma_eval_metrics_args* a, coma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * GPU-side input preparation: the step right before the hot path.  Replaces, for one sample, the
 * SimpleITK nearest-neighbour resample to 2 mm (VolumeDataset.py:236-259), torch.nan_to_num (:226), the centred zero padding
 * of data_util.pad_volume (data_util.py:814-828) and `mri[roi == 0] = 0` (VolumeDataset_ADNI_A4_combined.py:68) by ONE
 * kernel over the output voxels.  Arrays are [z, y, x] (the layout of sitk.GetArrayFromImage); per axis a:
 *   resampled size R_a = round(in_size_a * in_spacing_a / out_spacing_a)   (computed by the caller, numpy round-half-even)
 *   output voxel o (after padding) -> resampled index r = o - pad_before_a; outside [0, R_a) -> 0
 *   source index n = floor(r * out_spacing_a / in_spacing_a + 0.5) (ITK's round-half-up nearest index; same origin, same direction,
 *   identity transform); the voxel is inside iff -0.5 <= r * ratio < in_size_a - 0.5, else it takes default_value
 * mri / tau / roi may each be NULL (skipped); mri is masked by the resampled roi when both are given.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const float* mri; const float* tau; const float* roi;   /* [in_size[0], in_size[1], in_size[2]] fp32 */
  float* mri_out; float* tau_out; float* roi_out;           /* [out_size[0], out_size[1], out_size[2]] fp32 */
  int32_t in_size[3], res_size[3], out_size[3], pad_before[3];
  double ratio[3];                                          /* out_spacing / in_spacing per axis */
  float default_value;                                      /* the reference passes volume.GetPixelIDValue() (8 for float32 images) */
} coma_prepare_args;
int coma_prepare_volumes(const coma_prepare_args* a, coma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Small host -> device upload that does not touch a copy engine: `host_values` (n fp32 numbers in ordinary host memory: the
 * per-batch covariates and the B x 36 x 2 ROI table the reference reads from its JSON lookups, attn_unet_data_parallel.py:
 * 708-710,809-810) travel in the kernel parameter space, 960 values per launch.  A cudaMemcpyAsync of these ~2.5 KB queues on the
 * H2D copy engine behind the 134 MB volume upload of the NEXT batch, which stalled the step that needs them by the whole upload
 * (measured: +2 ms per 12 ms inference step).  The host values are consumed before the call returns.
 * ------------------------------------------------------------------------------------------- */
int coma_upload_small(float* dst, const float* host_values, int64_t n, coma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Weight layout transforms between the reference's parameter layout and the tap-major layout every conv entry point above
 * takes.  `param` is the framework's tensor [A][B][T] fp32 with the T = k^3 taps fastest: nn.Conv3d.weight [Cout][Cin][k][k][k]
 * (monai Convolution, attn_unet_data_parallel.py:166-205) or nn.ConvTranspose3d.weight [Cin][Cout][k][k][k] (:207-222).
 * `packed` is [T][R_pad][C_pad] in `dtype`, (r, c) = (a, b) or, with `swap`, (b, a); rows / columns beyond A / B are zero.
 *   forward operand     Conv3d: swap 0            ConvTranspose3d: swap 1           (rows = Cout, cols = Cin)
 *   data-gradient one   Conv3d stride 1: swap 1, flip 1   stride 2: swap 1   ConvTranspose3d: swap 0   (rows = Cin, cols = Cout)
 * `unpack` = 1 goes the other way for a weight gradient: packed fp32 [T][R_pad][C_pad] (coma_conv3d_wgrad's dw) -> param.
 * One launch each; replaces the permute / pad / cast / flip chain of framework kernels (~350 launches per training step).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  void* param;        /* [A][B][T] fp32: read when packing, written when unpacking */
  void* packed;       /* [T][R_pad][C_pad] */
  int A, B, T;        /* T <= 27 */
  int R_pad, C_pad;
  int swap, flip;     /* flip: tap t <-> T - 1 - t (the adjoint of a stride-1 convolution) */
  int dtype;          /* of `packed`: COMA_BF16 or COMA_F32 */
  int unpack;
} coma_weight_layout_args;
int coma_weight_layout(const coma_weight_layout_args* a, coma_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * All FiLM MLPs of a model in one launch, forward and backward.  Every conditioned layer of the reference owns
 * film = Linear(n_cov, 64) -> ReLU -> Linear(64, 2 C) on the per-sample covariates (CondConv call sites, attn_unet_data_parallel.py:
 * 126,285-306,318-325,354-367; specification in DESIGN.md section 4); under autograd these are ~12 framework launches per layer
 * and step (two GEMMs, ReLU, bias sums, chunk / cat) on [B x 6] .. [B x 512] matrices.  One block per layer here.
 *   forward : hid_l = relu(cov[:, :n_l] W1_l^T + b1_l)          [B][64]   (kept for backward)
 *             out_l = hid_l W2_l^T + b2_l, stored as [2][B][C_l]: dgamma block, then beta block
 *   backward: dW1, db1, dW2, db2 of every layer from d_dgamma_l / d_beta_l ([B][C_l] each, NULL = zero); no covariate gradient.
 * Parameter layouts are nn.Linear's: W1 [64][n_l], b1 [64], W2 [2 C_l][64], b2 [2 C_l], all fp32.
 * ------------------------------------------------------------------------------------------- */
#define COMA_FILM_MAX_LAYERS 32
#define COMA_FILM_HIDDEN 64
typedef struct {
  int32_t n_layers, B, cov_stride;         /* cov: [B][cov_stride] fp32, layer l reads its first n_cov[l] columns */
  const float* cov;
  float* hid;                              /* [n_layers][B][64] */
  int32_t n_cov[COMA_FILM_MAX_LAYERS], C[COMA_FILM_MAX_LAYERS];
  const float* W1[COMA_FILM_MAX_LAYERS]; const float* b1[COMA_FILM_MAX_LAYERS];
  const float* W2[COMA_FILM_MAX_LAYERS]; const float* b2[COMA_FILM_MAX_LAYERS];
  float* out[COMA_FILM_MAX_LAYERS];        /* forward: [2][B][C_l] */
  const float* d_dgamma[COMA_FILM_MAX_LAYERS]; const float* d_beta[COMA_FILM_MAX_LAYERS];     /* backward inputs */
  float* dW1[COMA_FILM_MAX_LAYERS]; float* db1[COMA_FILM_MAX_LAYERS];                         /* backward outputs (stored) */
  float* dW2[COMA_FILM_MAX_LAYERS]; float* db2[COMA_FILM_MAX_LAYERS];
} coma_film_args;
int coma_film_mlp_fwd(const coma_film_args* a, coma_stream_t stream);
int coma_film_mlp_bwd(const coma_film_args* a, coma_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* COMA_B200_H_ */
