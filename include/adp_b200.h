/*
 * adp_b200.h -- C ABI of libadp_b200.so: the B200 (sm_100a) implementation of the
 * audio-depth hot path  waveform -> STFT log-magnitude feature -> UNetBaseline ->
 * masked depth loss (forward and backward) -> clip + AdamW.
 *
 * The reference (Kang-ChangWoo/audio-depth-estimation) has no FFI layer: its
 * boundary is the Python API of dataloader/, models/unetbaseline_model.py and
 * utils_loss.py, all of which dispatch into torch/torchaudio/torchvision ops.
 * Each entry point below names the reference call site (file:line, relative to
 * the reference root) whose library op it replaces.
 *
 * Conventions
 *  - plain C symbols, raw DEVICE pointers and sizes, caller-owned memory, no
 *    allocation inside; workspaces are sized with the *_workspace_bytes calls;
 *  - every call enqueues on `stream` (a cudaStream_t passed as void*) and
 *    returns without synchronising;
 *  - return value 0 = ok, negative = error; adp_last_error() gives the text
 *    (thread-local);
 *  - activations are NHWC; Conv2d weights are [Cout][kh][kw][Cin] and
 *    ConvTranspose2d weights [Cin][kh][kw][Cout] -- i.e. the reference's
 *    [Cout,Cin,4,4] / [Cin,Cout,4,4] tensors held in torch.channels_last memory
 *    format, so state_dict shapes are unchanged;
 *  - re-entrant per (device, stream).
 */
#ifndef ADP_B200_H
#define ADP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADP_OK 0
#define ADP_ERR_ARG (-1)
#define ADP_ERR_CUDA (-2)
#define ADP_ERR_UNSUPPORTED (-3)

#define ADP_F32 0
#define ADP_BF16 1

#define ADP_MAX_LEVELS 10

const char* adp_last_error(void);
int adp_version(void);
/* 1 when the running device is sm_100 (B200); the tcgen05 path refuses others. */
int adp_device_is_sm100(void);

/* ------------------------------------------------------------------ feature
 * _get_spectrogram: dataloader/BatvisionV2_Dataset.py:177-185,
 * dataloader/BatvisionV1_Dataset.py:86-95  (T.Spectrogram(n_fft, win_length,
 * power=1, hop_length) -> torch.stft + abs).
 * wave: [rows, L] fp32 with row pitch `wave_pitch` elements (so the V2 cut of
 * BatvisionV2_Dataset.py:102-104 is just L < pitch); spec: [rows, n_fft/2+1, T],
 * T = 1 + L/hop. */
int adp_stft_mag(const float* wave, int rows, int L, int wave_pitch,
                 int n_fft, int win_length, int hop, float* spec, void* stream);

/* Whole audio branch of __getitem__ on a batch:
 * BatvisionV2_Dataset.py:96-135 (log_minmax = 1: log(x+1e-8), per-channel
 * min-max) / BatvisionV1_Dataset.py:70-78 (log_minmax = 0), then
 * utils_dataset.py:18-20 Resize((S,S)) (antialiased bilinear).
 * wave [B*C rows, L] -> out [B*C, S, S] fp32.  workspace: adp_feature_workspace_bytes. */
size_t adp_feature_workspace_bytes(int rows, int L, int n_fft, int hop);
int adp_feature_forward(const float* wave, int rows, int L, int wave_pitch,
                        int n_fft, int win_length, int hop, int log_minmax,
                        int out_size, float* out, void* workspace, size_t workspace_bytes,
                        void* stream);
/* The mel branch of the V2 transform (audio_format 'mel_spectrogram', the default of
 * conf/dataset/batvisionv2.yaml:8): _get_melspectrogram, BatvisionV2_Dataset.py:187-197 ==
 * T.MelSpectrogram(sample_rate, n_fft, win_length, hop = win_length/2, power=1, f_min, f_max,
 * n_mels, norm=None, mel_scale='htk'): |STFT| then the [n_fft/2+1 -> n_mels] triangular
 * filterbank.  mel: [rows, n_mels, T].  workspace: adp_feature_workspace_bytes (or the mel one). */
int adp_mel_spectrogram(const float* wave, int rows, int L, int wave_pitch, int n_fft, int win_length,
                        int hop, int n_mels, float sample_rate, float f_min, float f_max, float* mel,
                        void* workspace, size_t workspace_bytes, void* stream);
/* ... followed by log(x+1e-8), per-channel min-max and Resize((S,S)) (:117-135):
 * wave [B*C rows, L] -> out [B*C, S, S]. */
size_t adp_feature_mel_workspace_bytes(int rows, int L, int n_fft, int hop, int n_mels);
int adp_feature_forward_mel(const float* wave, int rows, int L, int wave_pitch, int n_fft,
                            int win_length, int hop, int n_mels, float sample_rate, float f_min,
                            float f_max, int log_minmax, int out_size, float* out, void* workspace,
                            size_t workspace_bytes, void* stream);
/* utils_dataset.py:18-20 alone: [rows, H, W] -> [rows, S, S]. */
int adp_resize_aa(const float* in, int rows, int H, int W, int out_size, float* out, void* stream);

/* Ground-truth depth preparation of __getitem__ (BatvisionV2_Dataset.py:68-78; V1 :47-65 with
 * nan_to_num = 1 and norm_div = max_depth when depth_norm): raw [rows,H,W] millimetres
 * (raw_dtype 0 = fp32, 1 = uint16) -> out [rows,S,S] fp32 metres, clipped to [0, max_depth]
 * (max_depth <= 0: no upper clip), cv2.INTER_NEAREST sampling, bit-exact with the numpy code. */
int adp_depth_prepare(const void* raw, int raw_dtype, int rows, int H, int W, int out_size,
                      float max_depth, int nan_to_num, float norm_div, float* out, void* stream);

/* --------------------------------------------------------------------- loss
 * train.py:646-669 + utils_loss.py:29-49.  Three phases so that data-parallel
 * ranks can all-reduce the four sufficient statistics in between
 * (sums = {N_valid, sum|p-g|, sum d, sum d^2}, doubles, ACCUMULATED into).
 * use_mask = 1: only pixels with gt != 0 count (train.py:646); 0: every element counts
 * (SIlogLoss / L1Loss called on already-masked vectors, utils_loss.py:29). */
int adp_depth_loss_sums(const float* pred, const float* gt, int64_t n, float scale, float eps,
                        int use_mask, double* sums, void* stream);
/* loss_out[3] = {loss, l1, silog};  loss = l1_w*l1 + silog_w*silog */
int adp_depth_loss_value(const double* sums, float l1_w, float silog_w, float lam,
                         float* loss_out, void* stream);
/* dpred = grad_scale * d loss / d pred  (grad_scale: device scalar or NULL = 1) */
int adp_depth_loss_backward(const float* pred, const float* gt, int64_t n, float scale, float eps,
                            int use_mask, const double* sums, float l1_w, float silog_w, float lam,
                            const float* grad_scale, float* dpred, void* stream);

/* ------------------------------------------------------------------ metrics
 * Validation / test metrics of a batch without leaving the device:
 * utils_criterion.py:6-90 (compute_errors) applied per sample, after the preparation of
 * train.py:807-825 / test.py:251-270 when prepare = 1 (pred, gt scaled by `scale` = max_depth
 * if depth_norm else 1; pred clipped to [clip_lo, clip_hi]; gt >= 0).  prepare = 0, scale = 1
 * is compute_errors itself.  pred, gt: [batch, n_per_sample] fp32.
 * metrics: [batch, 7] doubles = (abs_rel, rmse, a1, a2, a3, log_10, mae). */
size_t adp_depth_metrics_workspace_bytes(int batch);
int adp_depth_metrics(const float* pred, const float* gt, int batch, int64_t n_per_sample, float scale,
                      int prepare, float clip_lo, float clip_hi, double* metrics, void* workspace,
                      size_t workspace_bytes, void* stream);

/* --------------------------------------------------------------------- convs
 * nn.Conv2d(k4,s2,p1,bias=False)  models/unetbaseline_model.py:187 and
 * nn.ConvTranspose2d(k4,s2,p1)    models/unetbaseline_model.py:196,209,218,
 * forward (cudnnConvolutionForward), dgrad (BackwardData) and wgrad (BackwardFilter).
 * dtype = ADP_F32 (fp32 storage, SIMT fp32 kernels) or ADP_BF16 (bf16 storage; tcgen05/TMEM
 * implicit GEMM with fp32 accumulation when `w_op` is given and the shape qualifies,
 * otherwise the SIMT kernel on bf16 storage).  All activations NHWC.
 *   w     : fp32 master weight in the layout stated at the top of this file;
 *   w_op  : optional bf16 tensor-core operand made by adp_weight_operand (NULL = none);
 *   x1/c1 : second half of a channel concat (decoder skip), NULL/0 if none;
 *   dw    : fp32, same layout as w, ACCUMULATED into (zero it first). */

/* w [R][16][C] fp32 -> bf16 [R][16][C]: the tensor-core operand of a weight is its master layout cast to
 * bf16 for all four uses (conv fprop / convT dgrad read it K-major, conv dgrad / convT fprop N-major). */
int adp_weight_operand(const float* w, int R, int C, void* out, void* stream);

int adp_conv2d_k4s2_fprop(int dtype, const void* x, const float* w, const void* w_op, void* y,
                          int B, int Hin, int Win, int Cin, int Cout, void* stream);
int adp_conv2d_k4s2_dgrad(int dtype, const void* dy, const float* w, const void* w_op, void* dx,
                          int B, int Hin, int Win, int Cin, int Cout, void* stream);
int adp_conv2d_k4s2_wgrad(int dtype, const void* x, const void* dy, float* dw,
                          int B, int Hin, int Win, int Cin, int Cout, void* stream);
/* Hin/Win: spatial size of the transposed conv's INPUT (output is 2Hin x 2Win). */
int adp_convT2d_k4s2_fprop(int dtype, const void* x0, int c0, const void* x1, int c1,
                           const float* w, const void* w_op, void* y, int B, int Hin, int Win,
                           int Cout, void* stream);
int adp_convT2d_k4s2_dgrad(int dtype, const void* dy, const float* w, const void* w_op,
                           void* dx0, int c0, void* dx1, int c1, int B, int Hin, int Win, int Cout,
                           void* stream);
int adp_convT2d_k4s2_wgrad(int dtype, const void* x0, int c0, const void* x1, int c1,
                           const void* dy, float* dw, int B, int Hin, int Win, int Cout, void* stream);
/* The two thin layers of the U-Net on tensor cores, the im2col tile built in shared memory (bf16 storage, Cin = 2 /
 * Cout = 1, power-of-two grids at least 16 wide): the outermost nn.Conv2d(2 -> 64, k4 s2 p1) with both activations of its
 * output (models/unetbaseline_model.py:187, :195, :215 -- `a` feeds the next conv, `r` the skip), and the backward of the
 * outermost nn.ConvTranspose2d(128 -> 1, k4 s2 p1) (:196).
 *   x      fp32 NCHW [B,2,H,W] (kept to ~16 mantissa bits: hi/lo bf16 split);   w1 fp32 [64][16][2]
 *   a, r   bf16 NHWC [B,H/2,W/2,64]: a = lrelu(conv, slope0), r = lrelu(conv, slope1)
 *   du     fp32 [B,1,2Hi,2Wi] (rounded to bf16);   wT fp32 [128][16];   g0 | g1, x0 | x1: bf16 [B,Hi,Wi,64] halves
 *   w_scratch: 16 KB of device memory for the padded bf16 weight operand;  dw: ACCUMULATED into (zero it first). */
int adp_first_conv_k4s2_fprop(const float* x, const float* w1, void* w_scratch, void* a, float slope0, void* r,
                              float slope1, int B, int H, int W, void* stream);
int adp_first_conv_k4s2_wgrad(const float* x, const void* g_e, float* dw, int B, int H, int W, void* stream);
/* the same with the activation backward of that level folded in: dL/de = gA * (r > 0 ? 1 : slope) + (r > 0 ? gB : 0)
 * (gA: gradient of the LeakyReLU(slope) branch, gB: of the ReLU skip branch, r = ReLU(e); all bf16 [B,H/2,W/2,64]) is formed
 * in shared memory and never written (models/unetbaseline_model.py:187-192, :231-235). */
int adp_first_conv_k4s2_wgrad_act(const float* x, const void* gA, const void* gB, const void* r, float slope, float* dw,
                                  int B, int H, int W, void* stream);
int adp_last_convT_k4s2_dgrad(const float* du, const float* wT, void* w_scratch, void* g0, void* g1, int B, int Hi,
                              int Wi, void* stream);
/* x1_scale / x1_shift [64] (both or neither): x1 holds t, the tensor the BatchNorm + ReLU in front of the second input
 * half was applied to, and q = ReLU(t * scale + shift) is formed in shared memory (:218-223) -- the forward pass below then
 * never materialises q. */
int adp_last_convT_k4s2_wgrad(const void* x0, const void* x1, const float* x1_scale, const float* x1_shift,
                              const float* du, float* dw, int B, int Hi, int Wi, void* stream);
/* forward of the outermost ConvTranspose2d(128 -> 1) + bias + ReLU | Sigmoid (:196-206), input width 128: one tensor-core
 * product per input row, the 2 x 2 taps of every output pixel combined from the last three rows kept in shared memory (no
 * col2im pass).  y fp32 [B,1,2Hi,256]; w_scratch: 4 KB. */
int adp_last_convT_k4s2_fprop(const void* x0, const void* x1, const float* x1_scale, const float* x1_shift,
                              const float* wT, void* w_scratch, const float* bias, int final_sigmoid, float* y,
                              int B, int Hi, int Wi, void* stream);
/* 1 = use tcgen05 kernels for bf16 tensors where supported (default), 0 = SIMT only.
 * Returns the previous setting.  (Also: environment ADP_TC=0.) */
int adp_set_tensor_core(int on);

/* --------------------------------------------------------------------- U-Net
 * UnetGenerator / UnetSkipConnectionBlock, models/unetbaseline_model.py:123-235. */
typedef struct adp_unet_desc {
  int batch;        /* B */
  int in_ch;        /* input_nc (2) */
  int out_ch;       /* output_nc (1) */
  int ngf;          /* 64 */
  int num_downs;    /* 8 = unet_256, 7 = unet_128 */
  int size;         /* input H = W, a multiple of 2^num_downs */
  int dtype;        /* ADP_F32 | ADP_BF16 : storage + conv arithmetic of the hidden layers */
  int final_sigmoid;/* cfg.dataset.depth_norm: Sigmoid head (1) or ReLU head (0), :201-206 */
  int training;     /* BatchNorm uses batch statistics and updates running stats */
  float bn_eps;     /* 1e-5 */
  float bn_momentum;/* 0.1 */
  int reuse_weight_cache; /* 1: the bf16 weight operands in the workspace are still valid (inference) */
  int inference_only;     /* 1 (with training = 0): no backward pass will follow -- eval-mode BatchNorm scale/shift and the
                           * activations are applied in the convolution epilogues (test.py:231-241) and the workspace
                           * does not keep the raw convolution outputs the backward pass would need */
} adp_unet_desc;

typedef struct adp_unet_level {      /* level 0 = outermost block */
  float* conv_w;       /* [Cout][4][4][Cin] */
  float* convT_w;      /* [CinT][4][4][CoutT] */
  float* convT_bias;   /* [out_ch] (outermost only) or NULL */
  float* bn_down_w; float* bn_down_b; float* bn_down_rm; float* bn_down_rv;  /* NULL where the block has no down-norm */
  float* bn_up_w;   float* bn_up_b;   float* bn_up_rm;   float* bn_up_rv;
  /* Optional bf16 copies of conv_w / convT_w in the same element order, kept up to date by the caller
   * (adp_clip_adamw_step writes them through adp_tensor_ref.p_bf16).  NULL: the forward casts the fp32
   * weights itself.  Ignored for level 0 and in fp32 mode. */
  const void* conv_w_bf16; const void* convT_w_bf16;
} adp_unet_level;

size_t adp_unet_workspace_bytes(const adp_unet_desc* d);
/* x [B,in_ch,S,S] fp32 NCHW -> y [B,out_ch,S,S] fp32.  `params` has num_downs
 * entries.  The workspace keeps what backward needs. */
int adp_unet_forward(const adp_unet_desc* d, const float* x, const adp_unet_level* params,
                     void* workspace, size_t workspace_bytes, float* y, void* stream);
/* dy [B,out_ch,S,S] fp32 -> gradients of every parameter (same layouts, fp32,
 * OVERWRITTEN).  `grads` mirrors `params` (running-stat slots ignored). */
int adp_unet_backward(const adp_unet_desc* d, const float* x, const float* y, const float* dy,
                      const adp_unet_level* params, const adp_unet_level* grads,
                      void* workspace, size_t workspace_bytes, void* stream);
/* The same backward in 2*num_downs stages so that a data-parallel caller can start the
 * all-reduce of finished gradients while later stages run.  Stage s < num_downs: decoder
 * level s (its ConvTranspose2d weight/bias and the up-norm of level s+1 become final);
 * stage num_downs + k: encoder level num_downs-1-k (its Conv2d weight and down-norm).
 * Runs stages [stage_begin, stage_end); stages must be run in order. */
int adp_unet_backward_stages(const adp_unet_desc* d, const float* x, const float* y, const float* dy,
                             const adp_unet_level* params, const adp_unet_level* grads,
                             void* workspace, size_t workspace_bytes, int stage_begin, int stage_end,
                             void* stream);

/* ------------------------------------------------- binaural attention network
 * Building blocks of BASELINE config 4 (models/binaural_attention_model.py), bf16 NHWC
 * activations, tcgen05 kernels shared with the U-Net convolutions.
 * nn.Conv2d(C0+C1, Cout, 3, padding=1, bias=False) (:29,:32) over the channel concat (x0|x1)
 * (:75 torch.cat([x2, x1])): w_bf16 = the weight in channels_last memory [Cout][3][3][C0+C1]
 * cast to bf16 (adp_cast_bf16); the data gradient reads the SAME tensor.  Needs power-of-two
 * H, W and channel counts that are multiples of 64.  scratch: optional fp32 [B*H*W][N] used to
 * split the reduction on small grids (NULL: never split).  wgrad OVERWRITES dw (fp32, weight layout). */
int adp_cast_bf16(const float* src, void* dst, int64_t n, void* stream);
int adp_conv2d_k3s1_fprop(const void* x0, int C0, const void* x1, int C1, const void* w_bf16, void* y,
                          int B, int H, int W, int Cout, void* scratch, size_t scratch_bytes, void* stream);
int adp_conv2d_k3s1_dgrad(const void* dy, int Cout, const void* w_bf16, void* dx0, int C0, void* dx1, int C1,
                          int B, int H, int W, void* scratch, size_t scratch_bytes, void* stream);
int adp_conv2d_k3s1_wgrad(const void* dy, int Cout, const void* x0, int C0, const void* x1, int C1, float* dw,
                          int B, int H, int W, void* stream);
/* Row GEMM C[m][n] = sum_k (A0|A1)[m][k] * (b_kn ? B[k][n] : B[n][k]), bf16 operands, fp32
 * accumulate; output bf16 split over two tensors (N0 | N1) or one fp32 tensor.  The 1x1
 * convolutions (:97-103, :241, :263) and the attention products (:120-131) of config 4. */
int adp_gemm_rows_bf16(const void* a0, int K0, const void* a1, int K1, const void* b, int b_kn,
                       void* c_bf16_0, int N0, void* c_bf16_1, int N1, float* c_f32, int64_t M, void* stream);

/* First convolution of an encoder (:166 DoubleConv(1, C)): one fp32 input plane per sample
 * (x + b * x_batch_stride; the left / right channel of the [B,2,H,W] network input),
 * w fp32 [Cout][3][3]; y bf16 NHWC.  wgrad OVERWRITES dw. */
int adp_conv2d_k3s1_c1_fprop(const float* x, int64_t x_batch_stride, const float* w, void* y,
                             int B, int H, int W, int Cout, void* stream);
int adp_conv2d_k3s1_c1_wgrad(const void* dy, const float* x, int64_t x_batch_stride, float* dw,
                             int B, int H, int W, int Cout, void* stream);
/* nn.BatchNorm2d + ReLU (slope 0) over bf16 rows [rows = B*H*W][C] (:30-34, :242-243).
 * conv_bias (may be NULL): bias of the preceding convolution, NOT added to x -- it cancels in the
 * normalised output and is folded into the running mean.  saved: float [4C] kept for the backward
 * pass; sums_ws: double [2C] scratch.  backward: dy is the gradient of the activation output;
 * dgamma / dbeta are overwritten. */
int adp_bn_act_forward(const void* x, int64_t rows, int C, const float* gamma, const float* beta,
                       const float* conv_bias, float* running_mean, float* running_var, int training,
                       float eps, float momentum, float slope, void* y, float* saved, double* sums_ws,
                       void* stream);
int adp_bn_act_backward(const void* x, int64_t rows, int C, const float* saved, const void* dy, float slope,
                        int training, void* dx, float* dgamma, float* dbeta, double* sums_ws, void* stream);
/* nn.MaxPool2d(2) (:48) on bf16 NHWC, x [B,2Ho,2Wo,C] -> y [B,Ho,Wo,C]; backward recomputes the
 * arg-max (first maximum in scan order, as ATen) from the saved input. */
int adp_maxpool2_forward(const void* x, void* y, int B, int Ho, int Wo, int C, void* stream);
int adp_maxpool2_backward(const void* x, const void* dy, void* dx, int B, int Ho, int Wo, int C, void* stream);
/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) (:62): x [B,H,W,C] -> [B,2H,2W,C];
 * backward is the exact adjoint in gather form (no atomics). */
int adp_upsample2x_forward(const void* x, void* y, int B, int H, int W, int C, void* stream);
int adp_upsample2x_backward(const void* dy, void* dx, int B, int H, int W, int C, void* stream);
/* F.interpolate(mode='bilinear', align_corners=False) of fp32 planes [planes,Hi,Wi] -> [planes,Ho,Wo]
 * (the output resize of :322-328); backward = exact adjoint (dx is zeroed, then accumulated with atomics). */
int adp_bilinear_resize_forward(const float* x, float* y, int64_t planes, int Hi, int Wi, int Ho, int Wo, void* stream);
int adp_bilinear_resize_backward(const float* dy, float* dx, int64_t planes, int Hi, int Wi, int Ho, int Wo, void* stream);
/* nn.ConvTranspose2d(k2, s2) (:65-66, bilinear=False) is one row GEMM with 4N columns (a, b, n) per input pixel plus this
 * rearrangement: inverse = 0: src bf16 [B,H,W,2,2,N] (+ bias[n], may be NULL) -> dst [B,2H,2W,N]; inverse = 1: the reverse copy. */
int adp_pixel_shuffle2(const void* src, const float* bias, void* dst, int B, int H, int W, int N, int inverse, void* stream);
/* bf16 row helpers ([rows][C]).  adp_rows_op: 0: x += bias[c] (a == y, g = bias);
 * 1: y = a + g[0]*b (:134 residual with the learnable gamma); 2: y = a + b; 3: y = g[0]*a.
 * adp_rows_reduce: 0: out[c] = sum_r a[r][c] (bias gradients; sums_ws double [2C]);
 * 1: out[r] = sum_c a[r][c]*b[r][c]; 2: out[0] = sum a*b. */
int adp_rows_op(int op, const void* a, const void* b, const float* g, void* y, int64_t rows, int C, void* stream);
int adp_rows_reduce(int op, const void* a, const void* b, int64_t rows, int C, float* out, double* sums_ws, void* stream);
/* Attention softmax (:119-123) on materialised fp32 scores S [R][N] (row = query):
 * rows: P = softmax(scale*S) in bf16, m = row max of scale*S, l = row sum of exp(scale*S - m);
 * apply: P = exp(scale*S - m)/l with the statistics indexed by row (by_col 0) or column (1: S is
 * the transposed score matrix); backward: dS = scale * P * (dP - delta[query]). */
int adp_softmax_rows(const float* S, int64_t R, int N, float scale, void* P, float* m, float* l, void* stream);
int adp_softmax_apply(const float* S, int64_t R, int N, float scale, const float* m, const float* l, int by_col,
                      void* P, void* stream);
int adp_softmax_backward(const void* P, const float* dP, int64_t R, int N, float scale, const float* delta,
                         int by_col, void* dS, void* stream);
/* The same softmax folded into the epilogue of the score GEMM D[m][n] = sum_k a[m][k]*b[n][k], so the fp32 score matrix
 * is never written: mode 1 running row max of scale*D into stat_m (order-preserving int images, see
 * adp_softmax_stats_init), mode 2 row sums of exp(scale*D - max) into stat_l, mode 3 out = exp(scale*D - m[i]) / l[i]
 * (bf16), mode 4 out = scale * pmat[m][n] * (D - delta[i]) (softmax backward with D = dP).  i = m, or n when by_col
 * (D is the transposed score matrix K Q^T; the statistics stay per query). */
int adp_softmax_stats_init(int* stat_m, float* stat_l, int64_t n, void* stream);
int adp_gemm_rows_softmax(const void* a, int K, const void* b, int N, int64_t M, int mode, int by_col, float scale,
                          int* stat_m, float* stat_l, const float* delta, const void* pmat, void* out_bf16,
                          void* stream);
/* dw[m][n] += sum_r a[r][m]*b[r][n] (fp32 [M][ldd], caller zeroes): 1x1-convolution weight gradients. */
int adp_gemm_tn_bf16(const void* a, int M, const void* b, int N, float* dw, int ldd, int64_t rows, void* stream);
/* Output head (:262-265, :318-332): y = clamp(max_depth * sigmoid(x . w + bias), 0, max_depth), fp32 [rows];
 * backward overwrites dx (bf16), dw [C], db [1]. */
int adp_depth_head_forward(const void* x, const float* w, const float* bias, float max_depth, int64_t rows, int C,
                           float* y, void* stream);
int adp_depth_head_backward(const void* x, const float* w, const float* bias, float max_depth, const float* dy,
                            int64_t rows, int C, void* dx, float* dw, float* db, void* stream);

/* ----------------------------------------------------------------- optimiser
 * clip_grad_norm_(max_norm) + AdamW.step  train.py:471-476, :689-691.
 * Multi-tensor: n tensors described by device-visible pointer tables. */
typedef struct adp_tensor_ref {
  float* p; float* g; float* m; float* v; int64_t n;
  void* p_bf16;   /* optional: bf16 mirror of p (n elements), rewritten by the AdamW step; NULL = none */
} adp_tensor_ref;
/* sumsq[0] += sum g^2 over all tensors (double, device) */
int adp_grad_sumsq(const adp_tensor_ref* refs_host, int n_tensors, double* sumsq, void* stream);
/* p,m,v updated in place; clip coefficient min(1, max_norm/(sqrt(sumsq)+1e-6)) taken
 * from the device scalar; norm_out (device float, may be NULL) receives the total norm. */
int adp_clip_adamw_step(const adp_tensor_ref* refs_host, int n_tensors, const double* sumsq,
                        float max_norm, float lr, float beta1, float beta2, float eps,
                        float weight_decay, int step, float* norm_out, void* stream);

/* ------------------------------------------------------------------ measurement
 * Kernel launches issued by this library since load (bench.py's gpu_launches). */
long long adp_launch_count(void);
/* ... of which tcgen05 tensor-core kernels. */
long long adp_tc_launch_count(void);
/* CUDA-event timing of the convolution kernel families on their launching stream.
 * adp_profile_enable(1) clears and starts, (0) stops; adp_profile_read (after the stream is
 * synchronised) fills 5-entry arrays {gather conv, parity convT, wgrad, thin first/last layers,
 * elementwise}: total ms, total algorithmic FLOP, timed calls. */
int adp_profile_enable(int on);
int adp_profile_read(double* ms, double* work, long long* calls);
/* n = 11 entries: {gather conv, parity convT, wgrad} of the LARGE layers, thin first/last layers (FLOP), BatchNorm +
 * activation passes (work = algorithmic BYTES), {gather, parity, wgrad} of the deep small-M levels (E5-E8, D8-D6), then
 * the feature stage, the loss and clip + AdamW (work = algorithmic bytes).  n = 5 is adp_profile_read. */
int adp_profile_read_n(int n, double* ms, double* work, long long* calls);

/* Run-time form of the ADP_TC_* tuning variables: "tc_halo" (parity kernels load the tile's input window once per
 * channel chunk), "tc_max_bn", "side_stream" (U-Net weight gradients run on a library-owned side stream
 * next to the data-gradient chain), "tc_stats", "tc_alt", "tc_skip_pad_taps", "center", "thin_fused", "d1_fused",
 * "defer_finish" (bit 0 forward, bit 1 backward: split-K sums of the small levels are finished by the single-launch
 * BatchNorm kernel that consumes them instead of a finishing launch).  Returns the previous value, -1 for an unknown
 * name. */
int adp_set_option(const char* name, int value);
/* Hardware probe, not on any product path (tools/probe_umma_offset.py): out[m][n] = sum_k X[m+shift][k]*W[n][k] with the
 * UMMA A descriptor started `shift` rows into a SWIZZLE_128B tile; reports whether shifted windows into one shared-memory
 * tile are usable as MMA operands (base_offset = descriptor bits 49-51, sbo_bytes = stride between 8-row groups). */
int adp_selftest_umma_offset(const void* x, const void* w, int shift, int base_offset, int sbo_bytes, float* out,
                             void* stream);

/* adp_clip_adamw_step with the step counter on the device (int, incremented by the call; scratch = 2 floats):
 * capturable in a CUDA graph. */
int adp_clip_adamw_step_graph(const adp_tensor_ref* refs_host, int n_tensors, const double* sumsq,
                              float max_norm, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int* step_dev, float* scratch, float* norm_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADP_B200_H */
