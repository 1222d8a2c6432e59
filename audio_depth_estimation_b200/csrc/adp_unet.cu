// U-Net orchestration: the whole UnetGenerator forward and backward as one stream of kernel
// launches (no host synchronisation, CUDA-graph capturable), plus the per-layer C ABI.
//
// Dataflow of models/unetbaseline_model.py:123-235, level l = 0 outermost:
//   e[0] = Conv(x);  a[l] = LeakyReLU_0.2(bn(e[l]));  r[l] = ReLU(bn(e[l]))   (the in-place
//   LeakyReLU followed by the parent's in-place ReLU makes the skip ReLU(e), App. D-1)
//   e[l] = Conv(a[l-1]) (+BatchNorm for 0 < l < D-1)
//   t[l-1] = ConvT(r[l] | q[l])  (l = D-1: r only);  q[l-1] = ReLU(BatchNorm(t[l-1]))
//   y = act(ConvT(r[0] | q[0]) + bias)
// torch.cat never materialises: the transposed convolutions read the two halves as two tensors.
#include <string.h>
#include <stdlib.h>
#include "adp_common.cuh"

namespace {

using namespace adp;

struct LevelPlan {
  int cin, cout, hin, hout;      // encoder conv of this level
  int t_c1, t_cout;              // decoder convT: inputs (r: cout | q: t_c1), output channels t_cout
  bool bn_down, bn_up;
  size_t e, a, r, t, q;          // activations (level-l shaped: [B,hout,hout,cout])
  size_t g_a, g_r, g_q, g_e, g_t;
  size_t bn_down_f, bn_up_f;     // float[4*C]: scale, shift, mean, invstd
  size_t sums_down, sums_up;     // double[2C] forward statistics
  size_t bsums_down, bsums_up;   // double[2C] backward reductions
  size_t wb_conv, wb_convT;      // bf16 casts of the two weights in their master layouts
};

struct Plan {
  int D, B, esz;
  LevelPlan lv[ADP_MAX_LEVELS];
  size_t du;                     // float [B,1,S,S]
  size_t p_last, w16_last;       // tensor-core head: P fp32 [B,H/2,W/2,16], bf16 weights [16][Ct]
  size_t xp0, dp0, w1pad, wLpad, dthin;  // tensor-core thin layers: patch rows [pix][64] bf16, padded weights, D fp32 [128][128]
  bool thin_tc;
  // first-level centring (see use_center): per-input-channel sums of x, the constants m[cout0] / T[cout1]
  size_t xsum, center_m, center_T;
  size_t tc_scratch, tc_scratch_bytes;  // fp32 split-K partial sums of the tensor-core convolutions
  size_t tc_scratch_clear;              // leading bytes a split launch can touch
  size_t sums_begin, sums_end;   // forward BN sums region (zeroed every forward)
  size_t bsums_begin, bsums_end; // backward sums region
  size_t total;
};

int make_plan(const adp_unet_desc* d, Plan* p) {
  ADP_CHECK_ARG(d, "unet: null descriptor");
  ADP_CHECK_ARG(d->num_downs >= 2 && d->num_downs <= ADP_MAX_LEVELS, "unet: num_downs %d unsupported", d->num_downs);
  ADP_CHECK_ARG(d->batch > 0 && d->ngf > 0 && d->ngf % 4 == 0, "unet: bad batch/ngf");
  ADP_CHECK_ARG(d->out_ch == 1, "unet: output_nc must be 1 (define_G call sites: train.py:381, test.py:120)");
  ADP_CHECK_ARG(d->in_ch >= 1 && d->in_ch <= 16, "unet: input_nc %d unsupported", d->in_ch);
  ADP_CHECK_ARG(d->size > 0 && d->size % (1 << d->num_downs) == 0, "unet: size %d not divisible by 2^num_downs", d->size);
  ADP_CHECK_ARG(d->dtype == ADP_F32 || d->dtype == ADP_BF16, "unet: bad dtype");
  const int D = d->num_downs;
  p->D = D; p->B = d->batch; p->esz = d->dtype == ADP_F32 ? 4 : 2;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += adp_align_up(bytes, 256); return o; };
  int ch[ADP_MAX_LEVELS];
  for (int l = 0; l < D; ++l) ch[l] = d->ngf * (l < 3 ? (1 << l) : 8);
  for (int l = 0; l < D; ++l) {
    LevelPlan& L = p->lv[l];
    L.cin = l == 0 ? d->in_ch : ch[l - 1];
    L.cout = ch[l];
    L.hin = d->size >> l;
    L.hout = L.hin / 2;
    L.t_c1 = l == D - 1 ? 0 : ch[l];
    L.t_cout = l == 0 ? d->out_ch : ch[l - 1];
    L.bn_down = l > 0 && l < D - 1;
    L.bn_up = l > 0;
  }
  p->sums_begin = off;
  p->xsum = take(sizeof(double) * 16);
  for (int l = 0; l < D; ++l) {
    LevelPlan& L = p->lv[l];
    L.sums_down = take(sizeof(double) * 2 * L.cout);
    L.sums_up = take(sizeof(double) * 2 * (L.t_cout > 4 ? L.t_cout : 4));
  }
  p->sums_end = off;
  p->bsums_begin = off;
  for (int l = 0; l < D; ++l) {
    LevelPlan& L = p->lv[l];
    L.bsums_down = take(sizeof(double) * 2 * L.cout);
    L.bsums_up = take(sizeof(double) * 2 * (L.t_cout > 4 ? L.t_cout : 4));
  }
  p->bsums_end = off;
  for (int l = 0; l < D; ++l) {
    LevelPlan& L = p->lv[l];
    const size_t act = (size_t)d->batch * L.hout * L.hout * L.cout * p->esz;
    L.bn_down_f = take(sizeof(float) * 4 * L.cout);
    L.bn_up_f = take(sizeof(float) * 4 * (L.t_cout > 4 ? L.t_cout : 4));
    // (level 0: room for the explicit one-pixel border of the centred activation)
    L.e = take(act); L.a = take(l == 0 ? (size_t)d->batch * (L.hout + 2) * (L.hout + 2) * L.cout * p->esz : act); L.r = take(act);
    L.g_a = take(act); L.g_r = take(act); L.g_e = take(act);
    if (l < D - 1) { L.t = take(act); L.q = take(act); L.g_q = take(act); L.g_t = take(act); }
    else { L.t = L.q = L.g_q = L.g_t = 0; }
    const size_t wc = (size_t)L.cout * 16 * L.cin * 2, wt = (size_t)(L.cout + L.t_c1) * 16 * L.t_cout * 2;
    L.wb_conv = take(wc);
    L.wb_convT = take(wt);
  }
  p->du = take((size_t)d->batch * d->size * d->size * sizeof(float));
  // a layer only splits K when it has fewer tiles than SMs, i.e. fewer than ~148*128*128 outputs
  p->tc_scratch_bytes = d->dtype == ADP_BF16 ? (size_t)16 << 20 : 0;
  p->tc_scratch = take(p->tc_scratch_bytes);
  // the largest fp32 output a split launch can leave there: what the per-step memset has to cover
  p->tc_scratch_clear = 0;
  for (int l = 1; l < D && p->tc_scratch_bytes; ++l) {
    const LevelPlan& L = p->lv[l];
    const size_t px = (size_t)d->batch * L.hout * L.hout;
    const size_t cand[4] = {px * L.cout * 4,                      // encoder conv output / its input gradient one level down
                            4 * px * L.t_cout * 4,                // decoder convT output
                            px * (L.cout + L.t_c1) * 4,           // decoder convT input gradient
                            4 * px * L.cin * 4};                  // encoder conv input gradient
    for (size_t c : cand)
      if (c <= p->tc_scratch_bytes && c > p->tc_scratch_clear) p->tc_scratch_clear = c;
  }
  p->p_last = take(d->dtype == ADP_BF16 ? (size_t)d->batch * p->lv[0].hout * p->lv[0].hout * 16 * sizeof(float) : 0);
  p->w16_last = take((size_t)16 * (p->lv[0].cout + p->lv[0].t_c1) * 2);
  {
    const LevelPlan& L0 = p->lv[0];
    p->thin_tc = d->dtype == ADP_BF16 && L0.cout == 64 && 16 * L0.cin <= 64 && L0.t_c1 == 64 && d->out_ch == 1 &&
                 ((size_t)d->batch * L0.hout * L0.hout) % 2 == 0;
    const size_t rows = (size_t)d->batch * L0.hout * L0.hout;
    p->xp0 = take(p->thin_tc ? rows * 64 * 2 : 0);
    p->dp0 = take(p->thin_tc ? rows * 64 * 2 : 0);
    p->w1pad = take(64 * 64 * 2);
    p->wLpad = take(128 * 64 * 2);
    p->dthin = take(128 * 128 * sizeof(float));
    p->center_m = take(sizeof(float) * L0.cout);
    p->center_T = take(sizeof(float) * (D > 1 ? p->lv[1].cout : 4));
  }
  p->total = off;
  return ADP_OK;
}

inline char* at(void* ws, size_t off) { return reinterpret_cast<char*>(ws) + off; }

struct BnBuf { float *scale, *shift, *mean, *invstd; };
inline BnBuf bnbuf(void* ws, size_t off, int C) {
  float* f = reinterpret_cast<float*>(at(ws, off));
  return BnBuf{f, f + C, f + 2 * C, f + 3 * C};
}

bool use_tc(int dtype) { return dtype == ADP_BF16 && tc_enabled(); }

// the small-M, weight-streaming levels (at B = 64: E5-E8 / D8-D6, conv outputs of 8 x 8 pixels and below) are reported
// separately by the per-family timers
inline bool deep_level(int conv_out_side) { return conv_out_side <= 8; }

int g_thin_fused = -1;    // "thin_fused" / ADP_THIN_FUSED=0: patch matrices in HBM + pointwise GEMMs (the route for Cin != 2)
bool thin_fused_enabled() {
  if (g_thin_fused < 0) g_thin_fused = getenv("ADP_THIN_FUSED") ? atoi(getenv("ADP_THIN_FUSED")) : 1;
  return g_thin_fused != 0;
}

int g_d1_fused = -1;      // "d1_fused" / ADP_D1_FUSED=0: D1 forward as pointwise GEMM + col2im over a stored q[0]
bool d1_fused_enabled() {
  if (g_d1_fused < 0) g_d1_fused = getenv("ADP_D1_FUSED") ? atoi(getenv("ADP_D1_FUSED")) : 1;
  return g_d1_fused != 0;
}

int g_center = -1;    // "center" / ADP_CENTER=0 switches the first-level centring off
bool center_enabled() {
  if (g_center < 0) g_center = getenv("ADP_CENTER") ? atoi(getenv("ADP_CENTER")) : 1;
  return g_center != 0;
}

// First-level centring (bf16 tensor-core path).  With min-max-normalised log-spectrogram inputs (mean ~0.8, std ~0.05) the
// first activation a[0] = LeakyReLU(Conv(x)) is a large per-channel constant plus a small signal, and the first
// BatchNorm (level 1) divides by the standard deviation of the signal alone: bf16 storage of a[0] and of the raw
// Conv(a[0]) then costs ~2.4e-2 of the depth map (emulated on the CPU oracle and measured on the GPU), above the
// 2e-2 the path promises.  The network function is kept EXACTLY, only the storage changes:
//   * a[0] is stored as a[0] - m[n] (m = bf16(LeakyReLU(W1 . mean(x))), any constant is exact) inside a tensor with an
//     explicit one-pixel border holding -m[n], so the level-1 convolution sees zeros where the reference pads;
//   * its accumulators are then e[1] - T[n'], T = sum_{taps,c} W2[n'][tap][c] m[c]: a per-channel constant the level-1
//     BatchNorm removes (train: the batch mean shifts by -T; running_mean and the eval-mode shift add T back);
//   * the level-1 weight gradient sums over the border too (-m * dL/de[1]), which adds m[c] * sum_pixels dL/de[1][n']:
//     zero with batch statistics (BatchNorm's backward output sums to zero per channel), scale * sum gz in eval mode.
bool use_center(const adp_unet_desc* d, const Plan& p, bool tc) {
  if (!tc || !p.thin_tc || p.D < 3 || !center_enabled()) return false;
  const LevelPlan& L0 = p.lv[0];
  const LevelPlan& L1 = p.lv[1];
  return L1.bn_down && L0.cout == 64 && tc_supported_pointwise16(p.B, L0.hout, L0.hout, L0.cout, L0.t_c1) &&
         tc_supported_gather(p.B, L1.hin, L1.hin, L1.cin, L1.cout, 0) &&
         tc_supported_wgrad(p.B, L1.hout, L1.hout, L1.cout, 0, L1.cin) && d->out_ch == 1;
}

// D1 on the band kernel (adp_thin_tc.cu: thin_tc_last_fwd): its second input half is read as t[0] with the up-norm + ReLU
// applied in shared memory, so q[0] is never written by a forward pass that will be back-propagated; the weight gradient
// of D1 re-forms it the same way.  Forward and backward take the same decision from the descriptor alone.
bool use_d1_fused(const Plan& p, bool tc) {
  if (!tc || !p.thin_tc || p.D < 2 || !thin_fused_enabled() || !d1_fused_enabled()) return false;
  const LevelPlan& L0 = p.lv[0];
  return L0.cout == 64 && L0.t_c1 == 64 && tc_supported_pointwise16(p.B, L0.hout, L0.hout, L0.cout, L0.t_c1) &&
         thin_tc_last_fwd_supported(p.B, L0.hout, L0.hout);
}

// ---- family dispatch: tensor cores when the operands are bf16 and the shape is supported
// (ex: tensor-core extras -- fused BatchNorm statistics, padded input; ignored by the SIMT kernels, whose callers check
// *ex->stats_done and never request padding)
int conv_gather(int dtype, const void* x, const float* w, const void* wb, void* y0, int N0, void* y1, int N1,
                int B, int Hi, int Wi, int C, cudaStream_t s, const ConvExtras* ex = nullptr, bool deep = false) {
  ProfScope prof(deep ? PROF_GATHER_DEEP : PROF_GATHER, s, 2.0 * B * (Hi / 2) * (Wi / 2) * (double)(N0 + N1) * 16.0 * C);
  if (use_tc(dtype) && wb && tc_supported_gather(B, Hi, Wi, C, N0, N1))
    return tc_gather_conv(x, wb, y0, N0, y1, N1, B, Hi, Wi, C, s, ex);
  ADP_CHECK_ARG(!ex || !ex->pad_in, "conv_gather: padded input needs the tensor-core path");
  return simt_gather_conv(dtype, x, w, y0, N0, y1, N1, B, Hi, Wi, C, s);
}
int conv_parity(int dtype, const void* x0, int C0, const void* x1, int C1, const float* w, const void* wb, void* y,
                int B, int Hi, int Wi, int N, cudaStream_t s, const ConvExtras* ex = nullptr, bool deep = false) {
  ProfScope prof(deep ? PROF_PARITY_DEEP : PROF_PARITY, s, 2.0 * B * Hi * Wi * 4.0 * (double)N * 4.0 * (C0 + C1));
  if (use_tc(dtype) && wb && tc_supported_parity(B, Hi, Wi, C0, C1, N))
    return tc_parity_convT(x0, C0, x1, C1, wb, y, B, Hi, Wi, N, s, ex);
  return simt_parity_convT(dtype, x0, C0, x1, C1, w, y, B, Hi, Wi, N, s);
}
int conv_wgrad(int dtype, const void* s0, int M0, const void* s1, int M1, const void* g, int N, float* dw,
               int B, int Hs, int Ws, cudaStream_t s, int g_pad = 0, bool deep = false) {
  ProfScope prof(deep ? PROF_WGRAD_DEEP : PROF_WGRAD, s, 2.0 * B * Hs * Ws * 16.0 * (double)N * (M0 + M1));
  if (use_tc(dtype) && tc_supported_wgrad(B, Hs, Ws, M0, M1, N))
    return tc_wgrad(s0, M0, s1, M1, g, N, dw, B, Hs, Ws, s, g_pad);
  ADP_CHECK_ARG(!g_pad, "conv_wgrad: padded operand needs the tensor-core path");
  return simt_wgrad(dtype, s0, M0, s1, M1, g, N, dw, B, Hs, Ws, s);
}

// "defer_finish" / ADP_DEFER_FINISH (bit 0: forward, bit 1: backward; 0: every split-K launch finishes its own sums --
// the A/B partner of the hand-over to the single-launch BatchNorm kernels)
int g_defer_finish = -1;
int defer_finish_mask() {
  if (g_defer_finish < 0) g_defer_finish = getenv("ADP_DEFER_FINISH") ? (atoi(getenv("ADP_DEFER_FINISH")) & 3) : 3;
  return g_defer_finish;
}

int check_params(const adp_unet_desc* d, const Plan& p, const adp_unet_level* params) {
  ADP_CHECK_ARG(params, "unet: null params");
  for (int l = 0; l < p.D; ++l) {
    const LevelPlan& L = p.lv[l];
    ADP_CHECK_ARG(params[l].conv_w && params[l].convT_w, "unet: level %d missing conv weights", l);
    if (L.bn_down)
      ADP_CHECK_ARG(params[l].bn_down_w && params[l].bn_down_b && params[l].bn_down_rm && params[l].bn_down_rv,
                    "unet: level %d missing down-norm tensors", l);
    if (L.bn_up)
      ADP_CHECK_ARG(params[l].bn_up_w && params[l].bn_up_b && params[l].bn_up_rm && params[l].bn_up_rv,
                    "unet: level %d missing up-norm tensors", l);
  }
  (void)d;
  return ADP_OK;
}

}  // namespace

extern "C" size_t adp_unet_workspace_bytes(const adp_unet_desc* d) {
  Plan p;
  if (make_plan(d, &p) != ADP_OK) return 0;
  return p.total;
}

extern "C" int adp_unet_forward(const adp_unet_desc* d, const float* x, const adp_unet_level* params,
                                void* ws, size_t ws_bytes, float* y, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  Plan p;
  ADP_TRY(make_plan(d, &p));
  ADP_TRY(check_params(d, p, params));
  ADP_CHECK_ARG(x && y && ws && ws_bytes >= p.total, "unet_forward: null pointer or workspace too small (%zu < %zu)",
                ws_bytes, p.total);
  const int D = p.D, B = p.B, dt = d->dtype;
  auto w16 = [&](const void* mirror, size_t off) -> void* { return mirror ? const_cast<void*>(mirror) : (void*)at(ws, off); };
  const bool tc = use_tc(dt);
  // split-K partial sums: cleared once per step here, every split launch (forward and backward) re-zeroes what it used --
  // one memset node instead of one per split layer and pass
  if (tc && p.tc_scratch_clear) ADP_CUDA(cudaMemsetAsync(at(ws, p.tc_scratch), 0, p.tc_scratch_clear, s));
  tc_set_scratch(p.tc_scratch_bytes ? at(ws, p.tc_scratch) : nullptr, p.tc_scratch_bytes, tc && p.tc_scratch_bytes);

  const bool tc_head = tc && d->out_ch == 1 &&
                       tc_supported_pointwise16(B, p.lv[0].hout, p.lv[0].hout, p.lv[0].cout, p.lv[0].t_c1);
  const bool thin_tc = tc && tc_head && p.thin_tc;
  if (tc && !d->reuse_weight_cache) {
    if (thin_tc) {
      ADP_TRY(thin_pad_rows(params[0].conv_w, at(ws, p.w1pad), 64, 16 * p.lv[0].cin, p.lv[0].cin <= 2, s));
      ADP_TRY(thin_pad_rows(params[0].convT_w, at(ws, p.wLpad), 128, 16, 0, s));
    }
    if (tc_head)  // [Ct][16][1] -> [1][16][Ct]
      ADP_TRY(cast_transpose_taps(params[0].convT_w, at(ws, p.w16_last), p.lv[0].cout + p.lv[0].t_c1, 1, s));
    for (int l = 0; l < D; ++l) {
      const LevelPlan& L = p.lv[l];
      if (l > 0) {   // (skipped where the caller maintains a bf16 mirror, e.g. written by the fused AdamW step)
        if (!params[l].conv_w_bf16)
          ADP_TRY(cast_f32_to_bf16(params[l].conv_w, at(ws, L.wb_conv), (long long)L.cout * 16 * L.cin, s));
        if (!params[l].convT_w_bf16)
          ADP_TRY(cast_f32_to_bf16(params[l].convT_w, at(ws, L.wb_convT), (long long)(L.cout + L.t_c1) * 16 * L.t_cout, s));
      }
    }
  }
  ADP_CUDA(cudaMemsetAsync(at(ws, p.sums_begin), 0, p.sums_end - p.sums_begin, s));
  const bool center = thin_tc && use_center(d, p, tc);
  const bool fold = tc && !d->training && d->inference_only;      // BatchNorm + activation in the conv epilogues
  float* cen_m = reinterpret_cast<float*>(at(ws, p.center_m));
  float* cen_T = reinterpret_cast<float*>(at(ws, p.center_T));

  // ---- encoder
  {
    const LevelPlan& L = p.lv[0];
    ProfScope prof(PROF_THIN, s, 2.0 * B * L.hout * L.hout * 16.0 * L.cin * L.cout);
    if (thin_tc) {  // im2col rows (bf16, padded to 64) + pointwise tensor-core GEMM with both activations in the epilogue
      ConvExtras ex;
      memset(&ex, 0, sizeof(ex));
      if (center) {
        double* xsum = reinterpret_cast<double*>(at(ws, p.xsum));
        ADP_TRY(center_input_sums(x, B, L.cin, (long long)L.hin * L.hin, xsum, s));
        ADP_TRY(center_tables(xsum, 1.0 / ((double)B * L.hin * L.hin), params[0].conv_w, L.cin,
                              w16(params[1].conv_w_bf16, p.lv[1].wb_conv), p.lv[1].cout, cen_m, cen_T, at(ws, L.a), B, L.hout,
                              L.hout, s));
        ex.center = cen_m;
        ex.pad_out = 1;
      }
      if (thin_fused_enabled() && L.cin == 2 && thin_tc_supported(B, 2, L.hin, L.hin)) {
        // (patch tile built in shared memory from the fp32 planes: no im2col matrix in HBM)
        ADP_TRY(thin_tc_first_conv(x, at(ws, p.w1pad), at(ws, L.a), 0.2f, at(ws, L.r), 0.f, ex.center, ex.pad_out, B, L.hin,
                                   L.hin, s));
      } else {
        ADP_TRY(thin_patch_rows(x, at(ws, p.xp0), B, L.cin, L.hin, L.hin, L.cin <= 2, s));
        ADP_TRY(tc_pointwise(at(ws, p.xp0), 64, nullptr, 0, at(ws, p.w1pad), at(ws, L.a), 64, at(ws, L.r), 0, 1, 0.2f, 0.f, B,
                             L.hout, L.hout, s, &ex));
      }
    } else {
      ADP_TRY(first_conv_fprop(dt, x, params[0].conv_w, 0.2f, at(ws, L.a), 0.f, at(ws, L.r), B, L.hin, L.hin, L.cin,
                               L.cout, s));
    }
  }
  for (int l = 1; l < D; ++l) {
    const LevelPlan& L = p.lv[l];
    const long long rows = (long long)B * L.hout * L.hout;
    double* sums = reinterpret_cast<double*>(at(ws, L.sums_down));
    int fused = 0;
    ConvExtras ex;
    memset(&ex, 0, sizeof(ex));
    if (L.bn_down && d->training) { ex.stats = sums; ex.stats_done = &fused; }
    // a split-K launch of a small level hands its fp32 sums to the single-launch BatchNorm below (no finishing launch)
    float* part = nullptr;
    if (d->training && (defer_finish_mask() & 1) && (L.bn_down ? bn_small_ok(dt, rows, L.cout, 1) : L.cout % 8 == 0))
      ex.deferred = &part;
    ex.pad_in = (l == 1 && center) ? 1 : 0;
    int folded = 0;
    if (fold && L.bn_down) {
      // inference: scale / shift from the running statistics (cached with the weight operands; level 1 depends on the
      // batch through the centring offset), then BatchNorm + LeakyReLU / ReLU in the convolution epilogue
      BnBuf bn = bnbuf(ws, L.bn_down_f, L.cout);
      if (!d->reuse_weight_cache || (l == 1 && center)) {
        const BnFin fin{sums, 0.0, 1.f, params[l].bn_down_w, params[l].bn_down_b, params[l].bn_down_rm, params[l].bn_down_rv,
                        0, d->bn_eps, d->bn_momentum, bn.scale, bn.shift, bn.mean, bn.invstd, (l == 1 && center) ? cen_T : nullptr};
        ADP_TRY(bn_finalize(fin, L.cout, s));
      }
      ex.bn_scale = bn.scale; ex.bn_shift = bn.shift; ex.slope0 = 0.2f; ex.slope1 = 0.f;
      ex.y_act0 = at(ws, L.a); ex.y_act1 = at(ws, L.r); ex.fold_done = &folded;
    }
    ADP_TRY(conv_gather(dt, at(ws, p.lv[l - 1].a), params[l].conv_w, tc ? w16(params[l].conv_w_bf16, L.wb_conv) : nullptr,
                        at(ws, L.e), L.cout, nullptr, 0, B, L.hin, L.hin, L.cin, s, &ex, deep_level(L.hout)));
    if (folded) continue;
    // (work of the BatchNorm / activation passes = ALGORITHMIC bytes: every input once, every output once)
    ProfScope eprof(PROF_ELEM, s, (double)rows * L.cout * p.esz * (L.bn_down ? 3.0 : 2.0));
    if (L.bn_down) {
      BnBuf bn = bnbuf(ws, L.bn_down_f, L.cout);
      const BnFin fin{sums, 1.0 / (double)rows, rows > 1 ? (float)((double)rows / (double)(rows - 1)) : 1.f, params[l].bn_down_w, params[l].bn_down_b, params[l].bn_down_rm, params[l].bn_down_rv,
                      d->training, d->bn_eps, d->bn_momentum, bn.scale, bn.shift, bn.mean, bn.invstd,
                      (l == 1 && center) ? cen_T : nullptr};
      if (d->training && !fused && bn_small_ok(dt, rows, L.cout, 1)) {
        ADP_TRY(bn_small_fwd(dt, at(ws, L.e), rows, L.cout, fin, 0.2f, at(ws, L.a), 0.f, at(ws, L.r), s, part));
      } else {
        ADP_CHECK_ARG(!part, "unet_forward: un-finished partial sums without a consumer (level %d)", l);
        if (d->training && !fused) ADP_TRY(bn_stats(dt, at(ws, L.e), rows, L.cout, sums, s));
        ADP_TRY(bn_affine_act(dt, at(ws, L.e), rows, L.cout, fin, 0.2f, at(ws, L.a), 0.f, at(ws, L.r), s));
      }
    } else if (part) {   // innermost level: finish the split-K sums and apply the ReLU in one launch
      ADP_TRY(finish_act(dt, part, rows * L.cout, nullptr, 0.f, at(ws, L.e), at(ws, L.r), s));
    } else {
      ADP_TRY(affine_act(dt, at(ws, L.e), rows, L.cout, nullptr, nullptr, 0.f, at(ws, L.r), 0.f, nullptr, s));
    }
  }
  // ---- decoder
  const bool d1_fused = tc_head && use_d1_fused(p, tc);
  bool q0_folded = false;
  for (int l = D - 1; l >= 1; --l) {
    const LevelPlan& L = p.lv[l];
    const LevelPlan& O = p.lv[l - 1];  // output lives at level l-1's resolution
    const long long rows = (long long)B * O.hout * O.hout;
    double* sums = reinterpret_cast<double*>(at(ws, L.sums_up));
    int fused = 0;
    ConvExtras ex;
    memset(&ex, 0, sizeof(ex));
    if (d->training) { ex.stats = sums; ex.stats_done = &fused; }
    float* part = nullptr;
    if (d->training && (defer_finish_mask() & 1) && !(l == 1 && d1_fused) && bn_small_ok(dt, rows, L.t_cout, 1)) ex.deferred = &part;
    int folded = 0;
    if (fold) {
      BnBuf bnf = bnbuf(ws, L.bn_up_f, L.t_cout);
      if (!d->reuse_weight_cache) {
        const BnFin fin{sums, 0.0, 1.f, params[l].bn_up_w, params[l].bn_up_b, params[l].bn_up_rm, params[l].bn_up_rv,
                        0, d->bn_eps, d->bn_momentum, bnf.scale, bnf.shift, bnf.mean, bnf.invstd, nullptr};
        ADP_TRY(bn_finalize(fin, L.t_cout, s));
      }
      ex.bn_scale = bnf.scale; ex.bn_shift = bnf.shift; ex.slope0 = 0.f; ex.slope1 = 0.f;
      ex.y_act0 = at(ws, O.q); ex.y_act1 = nullptr; ex.fold_done = &folded;
    }
    ADP_TRY(conv_parity(dt, at(ws, L.r), L.cout, L.t_c1 ? at(ws, L.q) : nullptr, L.t_c1, params[l].convT_w,
                        tc ? w16(params[l].convT_w_bf16, L.wb_convT) : nullptr, at(ws, O.t), B, L.hout, L.hout, L.t_cout, s, &ex,
                        deep_level(O.hout)));
    if (l == 1) q0_folded = folded != 0;
    if (folded) continue;
    if (l == 1 && d1_fused) {      // only the coefficients: q[0] = ReLU(t[0] * scale + shift) is formed inside D1's kernels
      BnBuf bn = bnbuf(ws, L.bn_up_f, L.t_cout);
      const BnFin fin{sums, 1.0 / (double)rows, rows > 1 ? (float)((double)rows / (double)(rows - 1)) : 1.f, params[l].bn_up_w, params[l].bn_up_b, params[l].bn_up_rm, params[l].bn_up_rv,
                      d->training, d->bn_eps, d->bn_momentum, bn.scale, bn.shift, bn.mean, bn.invstd, nullptr};
      ProfScope eprof(PROF_ELEM, s, d->training && !fused ? (double)rows * L.t_cout * p.esz : 0.0);
      if (d->training && !fused) ADP_TRY(bn_stats(dt, at(ws, O.t), rows, L.t_cout, sums, s));
      ADP_TRY(bn_finalize(fin, L.t_cout, s));
      continue;
    }
    ProfScope eprof(PROF_ELEM, s, (double)rows * L.t_cout * p.esz * 2.0);
    BnBuf bn = bnbuf(ws, L.bn_up_f, L.t_cout);
    const BnFin fin{sums, 1.0 / (double)rows, rows > 1 ? (float)((double)rows / (double)(rows - 1)) : 1.f, params[l].bn_up_w, params[l].bn_up_b, params[l].bn_up_rm, params[l].bn_up_rv,
                    d->training, d->bn_eps, d->bn_momentum, bn.scale, bn.shift, bn.mean, bn.invstd, nullptr};
    if (d->training && !fused && bn_small_ok(dt, rows, L.t_cout, 1)) {
      ADP_TRY(bn_small_fwd(dt, at(ws, O.t), rows, L.t_cout, fin, 0.f, at(ws, O.q), 0.f, nullptr, s, part));
    } else {
      ADP_CHECK_ARG(!part, "unet_forward: un-finished partial sums without a consumer (up level %d)", l);
      if (d->training && !fused) ADP_TRY(bn_stats(dt, at(ws, O.t), rows, L.t_cout, sums, s));
      ADP_TRY(bn_affine_act(dt, at(ws, O.t), rows, L.t_cout, fin, 0.f, at(ws, O.q), 0.f, nullptr, s));
    }
  }
  {
    const LevelPlan& L = p.lv[0];
    ProfScope prof(PROF_THIN, s, 2.0 * B * L.hout * L.hout * 16.0 * (L.cout + L.t_c1));
    if (d1_fused) {
      BnBuf bn = bnbuf(ws, p.lv[1].bn_up_f, 64);
      ADP_TRY(thin_tc_last_fwd(at(ws, L.r), q0_folded ? at(ws, L.q) : at(ws, L.t), q0_folded ? nullptr : bn.scale, bn.shift,
                               at(ws, p.w16_last), params[0].convT_bias, d->final_sigmoid, y, B, L.hout, L.hout, s));
    } else if (tc_head) {
      float* P = reinterpret_cast<float*>(at(ws, p.p_last));
      ADP_TRY(tc_pointwise16(at(ws, L.r), L.cout, at(ws, L.q), L.t_c1, at(ws, p.w16_last), P, B, L.hout, L.hout, s));
      ADP_TRY(last_convT_col2im(P, params[0].convT_bias, d->final_sigmoid, y, B, L.hout, L.hout, s));
    } else {
      ADP_TRY(last_convT_fprop(dt, at(ws, L.r), L.cout, at(ws, L.q), L.t_c1, params[0].convT_w, params[0].convT_bias,
                               d->final_sigmoid, y, B, L.hout, L.hout, s));
    }
  }
  return ADP_OK;
}

// Weight gradients run on a library-owned side stream, concurrently with the data-gradient chain of the same stage
// group: dw(l) only feeds the optimiser, while dx(l) is on the critical path to level l-1, and the deep levels'
// kernels use a fraction of the SMs each.  Forked with an event after the upstream gradient exists, joined before the
// call returns (so callers -- the all-reduce hook, the optimiser, a CUDA-graph capture -- see one stream).
namespace {
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork[32];
  cudaEvent_t join = nullptr;
};
int g_side_stream = -1;    // ADP_SIDE_STREAM=0 keeps everything on the caller's stream
int side_stream_for_device(SideStream** out) {
  static SideStream tab[64];
  *out = nullptr;
  if (g_side_stream < 0) g_side_stream = getenv("ADP_SIDE_STREAM") ? atoi(getenv("ADP_SIDE_STREAM")) : 1;
  if (!g_side_stream) return ADP_OK;
  int dev = 0;
  ADP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return ADP_OK;
  SideStream& t = tab[dev];
  if (!t.stream) {     // (first backward of a process is an eager step, never inside a capture)
    ADP_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
    for (auto& e : t.fork) ADP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ADP_CUDA(cudaEventCreateWithFlags(&t.join, cudaEventDisableTiming));
  }
  *out = &t;
  return ADP_OK;
}
}  // namespace

namespace adp {
int unet_set_option(const char* name, int value) {
  if (!strcmp(name, "center")) {
    const int prev = center_enabled() ? 1 : 0;
    g_center = value ? 1 : 0;
    return prev;
  }
  if (!strcmp(name, "d1_fused")) {
    const int prev = d1_fused_enabled() ? 1 : 0;
    g_d1_fused = value ? 1 : 0;
    return prev;
  }
  if (!strcmp(name, "defer_finish")) {
    const int prev = defer_finish_mask();
    g_defer_finish = value & 3;
    return prev;
  }
  if (!strcmp(name, "thin_fused")) {
    const int prev = thin_fused_enabled() ? 1 : 0;
    g_thin_fused = value ? 1 : 0;
    return prev;
  }
  if (strcmp(name, "side_stream")) return -1;
  if (g_side_stream < 0) g_side_stream = getenv("ADP_SIDE_STREAM") ? atoi(getenv("ADP_SIDE_STREAM")) : 1;
  const int prev = g_side_stream;
  g_side_stream = value ? 1 : 0;
  return prev;
}
}  // namespace adp

extern "C" int adp_unet_backward_stages(const adp_unet_desc* d, const float* x, const float* y, const float* dy,
                                        const adp_unet_level* params, const adp_unet_level* grads, void* ws,
                                        size_t ws_bytes, int stage_begin, int stage_end, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  Plan p;
  ADP_TRY(make_plan(d, &p));
  ADP_TRY(check_params(d, p, params));
  ADP_CHECK_ARG(x && y && dy && grads && ws && ws_bytes >= p.total, "unet_backward: null pointer or workspace too small");
  const int D = p.D, B = p.B, dt = d->dtype;
  auto w16 = [&](const void* mirror, size_t off) -> void* { return mirror ? const_cast<void*>(mirror) : (void*)at(ws, off); };
  ADP_CHECK_ARG(stage_begin >= 0 && stage_end <= 2 * D && stage_begin <= stage_end, "unet_backward: bad stage range");
  const bool tc = use_tc(dt);
  const bool thin_tc_bwd = tc && p.thin_tc && tc_supported_pointwise16(B, p.lv[0].hout, p.lv[0].hout, p.lv[0].cout, p.lv[0].t_c1);
  const int bn_mode = d->training ? 2 : 1;
  // (split-K partial sums: cleared by the forward pass that filled this workspace, left cleared by every split launch)
  tc_set_scratch(p.tc_scratch_bytes ? at(ws, p.tc_scratch) : nullptr, p.tc_scratch_bytes, tc && p.tc_scratch_bytes);
  const bool center = thin_tc_bwd && use_center(d, p, tc);     // (same decision as the forward pass that filled the workspace)

  // BatchNorm + ReLU backward of q[l] (up-norm of level l+1): g_q[l] -> g_t[l]
  // (part: the data-gradient convolution left its split-K sums [rows][cout + t_c1] un-finished -- g_q[l] is taken from
  // them, the skip half is finished into g_r[l] on the way)
  auto up_norm_bwd = [&](int l, float* part) -> int {
    const LevelPlan& L = p.lv[l];
    const LevelPlan& U = p.lv[l + 1];
    const long long rows = (long long)B * L.hout * L.hout;
    const int C = U.t_cout;
    BnBuf bn = bnbuf(ws, U.bn_up_f, C);
    double* bs = reinterpret_cast<double*>(at(ws, U.bsums_up));
    ProfScope eprof(PROF_ELEM, s, (double)rows * C * p.esz * 3.0);      // x, g -> dx
    if (bn_small_ok(dt, rows, C, 3)) {
      const BnSmallPartial pp{part, L.cout + L.t_c1, L.cout, at(ws, L.g_r), L.cout};
      return bn_small_bwd(dt, at(ws, L.t), rows, C, bn.scale, bn.shift, bn.mean, bn.invstd, at(ws, L.g_q), 0.f, nullptr, 0.f,
                          bn_mode, at(ws, L.g_t), grads[l + 1].bn_up_w, grads[l + 1].bn_up_b, bs, s, part ? &pp : nullptr);
    }
    ADP_CHECK_ARG(!part, "unet_backward: un-finished partial sums without a consumer (up level %d)", l);
    ADP_TRY(act_bn_bwd_reduce(dt, at(ws, L.t), rows, C, bn.scale, bn.shift, bn.mean, bn.invstd, at(ws, L.g_q), 0.f,
                              nullptr, 0.f, bs, s));
    ADP_TRY(act_bn_bwd_apply(dt, at(ws, L.t), rows, C, bn.scale, bn.shift, bn.mean, bn.invstd, at(ws, L.g_q), 0.f,
                             nullptr, 0.f, bs, bn_mode, at(ws, L.g_t), grads[l + 1].bn_up_w, grads[l + 1].bn_up_b, s));
    return ADP_OK;
  };

  SideStream* side = nullptr;
  if (tc) ADP_TRY(side_stream_for_device(&side));
  bool forked = false;
  // join on every exit path (also the error returns): everything the side stream did is ordered before whatever the
  // caller enqueues next on s, and a stream capture never ends with an un-joined branch
  struct Join {
    SideStream*& side; bool& forked; cudaStream_t s;
    ~Join() {
      if (forked && side && cudaEventRecord(side->join, side->stream) == cudaSuccess) (void)cudaStreamWaitEvent(s, side->join, 0);
    }
  } join_guard{side, forked, s};
  // returns the stream the weight gradient of stage `st` should use (forks the side stream behind everything queued on s)
  auto wgrad_stream = [&](int st) -> cudaStream_t {
    if (!side) return s;
    if (cudaEventRecord(side->fork[st & 31], s) != cudaSuccess || cudaStreamWaitEvent(side->stream, side->fork[st & 31], 0) != cudaSuccess) {
      (void)cudaGetLastError();
      return s;
    }
    forked = true;
    return side->stream;
  };

  float* part_a = nullptr;      // un-finished split-K sums of g_a[l - 1], handed from encoder stage l to stage l - 1
  float* part_r = nullptr;      // un-finished split-K sums of g_r[D - 1], handed from the innermost decoder stage to the encoder's
  for (int st = stage_begin; st < stage_end; ++st) {
    if (st == 0) {
      const LevelPlan& L = p.lv[0];
      const int Ct = L.cout + L.t_c1;
      ADP_CUDA(cudaMemsetAsync(at(ws, p.bsums_begin), 0, p.bsums_end - p.bsums_begin, s));
      float* du = reinterpret_cast<float*>(at(ws, p.du));
      ADP_CHECK_ARG(grads[0].convT_bias, "unet_backward: level 0 convT bias gradient missing");
      ADP_CUDA(cudaMemsetAsync(grads[0].convT_bias, 0, sizeof(float), s));
      ADP_TRY(head_bwd(y, dy, (long long)B * d->size * d->size, d->final_sigmoid, du, grads[0].convT_bias, s));
      if (thin_tc_bwd) {
        ProfScope prof(PROF_THIN, s, 4.0 * B * L.hout * L.hout * 16.0 * Ct);
        if (thin_fused_enabled() && L.cout == 64 && L.t_c1 == 64 && thin_tc_supported(B, 1, d->size, d->size)) {
          ADP_CUDA(cudaMemsetAsync(grads[0].convT_w, 0, sizeof(float) * 16 * Ct, s));
          if (use_d1_fused(p, tc)) {     // q[0] was never written: re-formed from t[0] on load
            BnBuf bn0 = bnbuf(ws, p.lv[1].bn_up_f, 64);
            ADP_TRY(thin_tc_last_wgrad(at(ws, L.r), at(ws, L.t), du, grads[0].convT_w, B, L.hout, L.hout, s, bn0.scale, bn0.shift));
          } else {
            ADP_TRY(thin_tc_last_wgrad(at(ws, L.r), at(ws, L.q), du, grads[0].convT_w, B, L.hout, L.hout, s));
          }
          ADP_TRY(thin_tc_last_dgrad(du, at(ws, p.wLpad), at(ws, L.g_r), at(ws, L.g_q), B, L.hout, L.hout, s));
        } else {
          float* Dt = reinterpret_cast<float*>(at(ws, p.dthin));
          ADP_TRY(thin_patch_rows(du, at(ws, p.dp0), B, 1, d->size, d->size, 0, s));
          ADP_TRY(tc_gemm_tn(at(ws, L.r), 64, 0, at(ws, L.q), 64, 0, at(ws, p.dp0), 64, 64, (long long)B * L.hout * L.hout, Dt, s));
          ADP_TRY(thin_fold_wgrad(Dt, grads[0].convT_w, 0, 16, s));
          ADP_TRY(tc_pointwise(at(ws, p.dp0), 64, nullptr, 0, at(ws, p.wLpad), at(ws, L.g_r), 64, at(ws, L.g_q), 64, 0, 0.f,
                               0.f, B, L.hout, L.hout, s));
        }
      } else {
        ADP_CUDA(cudaMemsetAsync(grads[0].convT_w, 0, sizeof(float) * 16 * Ct, s));
        ADP_TRY(last_convT_wgrad(dt, at(ws, L.r), L.cout, at(ws, L.q), L.t_c1, du, grads[0].convT_w, B, L.hout, L.hout, s));
        ADP_TRY(last_convT_dgrad(dt, du, params[0].convT_w, at(ws, L.g_r), L.cout, at(ws, L.g_q), L.t_c1, B, L.hout,
                                 L.hout, s));
      }
      ADP_TRY(up_norm_bwd(0, nullptr));
    } else if (st < D) {
      const int l = st;
      const LevelPlan& L = p.lv[l];
      const LevelPlan& O = p.lv[l - 1];
      const int Ct = L.cout + L.t_c1;
      {
        cudaStream_t sw = wgrad_stream(st);
        ADP_CUDA(cudaMemsetAsync(grads[l].convT_w, 0, sizeof(float) * 16 * (size_t)Ct * L.t_cout, sw));
        ADP_TRY(conv_wgrad(dt, at(ws, L.r), L.cout, L.t_c1 ? at(ws, L.q) : nullptr, L.t_c1, at(ws, O.g_t), L.t_cout,
                           grads[l].convT_w, B, L.hout, L.hout, sw, 0, deep_level(O.hout)));
      }
      float* part = nullptr;
      ConvExtras gex;
      memset(&gex, 0, sizeof(gex));
      if (l < D - 1 && (defer_finish_mask() & 2) && L.t_c1 == p.lv[l + 1].t_cout &&
          bn_small_ok(dt, (long long)B * L.hout * L.hout, L.t_c1, 3))
        gex.deferred = &part;
      // innermost level: g_r[D-1] has one reader, the activation backward that opens the next stage
      if (l == D - 1 && (defer_finish_mask() & 2) && st + 1 < stage_end && L.t_c1 == 0 && L.cout % 8 == 0) gex.deferred = &part_r;
      ADP_TRY(conv_gather(dt, at(ws, O.g_t), params[l].convT_w, tc ? w16(params[l].convT_w_bf16, L.wb_convT) : nullptr, at(ws, L.g_r),
                          L.cout, L.t_c1 ? at(ws, L.g_q) : nullptr, L.t_c1, B, O.hout, O.hout, L.t_cout, s, &gex,
                          deep_level(O.hout)));
      if (l < D - 1) ADP_TRY(up_norm_bwd(l, part));
    } else {
      const int l = 2 * D - 1 - st;
      const LevelPlan& L = p.lv[l];
      const long long rows = (long long)B * L.hout * L.hout;
      // level 0 on the shared-memory patch route: dL/de[0] = gA * lrelu' + gB * relu' is formed inside the weight-gradient
      // kernel (its only consumer) and never written
      const bool fuse_act0 = l == 0 && thin_tc_bwd && thin_fused_enabled() && L.cin == 2 && L.cout == 64 &&
                             thin_tc_supported(B, 2, L.hin, L.hin);
      if (!fuse_act0) {
      ProfScope eprof(PROF_ELEM, s, (double)rows * L.cout * p.esz * (l == D - 1 ? 3.0 : 4.0));   // x, gA[, gB] -> dx
      if (l == D - 1 && part_r) {
        ADP_TRY(finish_act(dt, part_r, rows * L.cout, at(ws, L.e), 0.f, nullptr, at(ws, L.g_e), s));
        part_r = nullptr;
      } else if (l == D - 1) {
        ADP_TRY(act_bn_bwd_apply(dt, at(ws, L.e), rows, L.cout, nullptr, nullptr, nullptr, nullptr, at(ws, L.g_r), 0.f,
                                 nullptr, 0.f, nullptr, 0, at(ws, L.g_e), nullptr, nullptr, s));
      } else if (L.bn_down) {
        BnBuf bn = bnbuf(ws, L.bn_down_f, L.cout);
        double* bs = reinterpret_cast<double*>(at(ws, L.bsums_down));
        if (bn_small_ok(dt, rows, L.cout, 3)) {
          const BnSmallPartial pp{part_a, L.cout, 0, nullptr, 0};
          ADP_TRY(bn_small_bwd(dt, at(ws, L.e), rows, L.cout, bn.scale, bn.shift, bn.mean, bn.invstd, at(ws, L.g_a), 0.2f,
                               at(ws, L.g_r), 0.f, bn_mode, at(ws, L.g_e), grads[l].bn_down_w, grads[l].bn_down_b, bs, s,
                               part_a ? &pp : nullptr));
          part_a = nullptr;
        } else {
          ADP_TRY(act_bn_bwd_reduce(dt, at(ws, L.e), rows, L.cout, bn.scale, bn.shift, bn.mean, bn.invstd,
                                    at(ws, L.g_a), 0.2f, at(ws, L.g_r), 0.f, bs, s));
          ADP_TRY(act_bn_bwd_apply(dt, at(ws, L.e), rows, L.cout, bn.scale, bn.shift, bn.mean, bn.invstd, at(ws, L.g_a),
                                   0.2f, at(ws, L.g_r), 0.f, bs, bn_mode, at(ws, L.g_e), grads[l].bn_down_w,
                                   grads[l].bn_down_b, s));
        }
      } else {  // level 0: no norm; e > 0 <=> r = ReLU(e) > 0 (a[0] may be stored centred, r[0] never is)
        ADP_TRY(act_bn_bwd_apply(dt, at(ws, L.r), rows, L.cout, nullptr, nullptr, nullptr, nullptr, at(ws, L.g_a), 0.2f,
                                 at(ws, L.g_r), 0.f, nullptr, 0, at(ws, L.g_e), nullptr, nullptr, s));
      }
      }
      cudaStream_t sw = l > 0 ? wgrad_stream(st) : s;
      ADP_CUDA(cudaMemsetAsync(grads[l].conv_w, 0, sizeof(float) * 16 * (size_t)L.cout * L.cin, sw));
      if (l == 0 && thin_tc_bwd) {
        ProfScope prof(PROF_THIN, s, 2.0 * B * L.hout * L.hout * 16.0 * L.cin * L.cout);
        if (fuse_act0) {       // (dw zeroed above; e > 0 <=> r = ReLU(e) > 0: a[0] may be stored centred, r[0] never is)
          ADP_TRY(thin_tc_first_wgrad_act(x, at(ws, L.g_a), at(ws, L.g_r), at(ws, L.r), 0.2f, grads[0].conv_w, B, L.hin, L.hin, s));
        } else if (thin_fused_enabled() && L.cin == 2 && thin_tc_supported(B, 2, L.hin, L.hin)) {
          ADP_TRY(thin_tc_first_wgrad(x, at(ws, L.g_e), grads[0].conv_w, B, L.hin, L.hin, s));
        } else {
          float* Dt = reinterpret_cast<float*>(at(ws, p.dthin));
          // pixel pairs folded into 128 "channels": D[(h,n)][(h',t)], the two diagonal blocks are the gradient
          ADP_TRY(tc_gemm_tn(at(ws, L.g_e), 128, 0, at(ws, L.g_e), 128, 64, at(ws, p.xp0), 128, 128,
                             (long long)B * L.hout * L.hout / 2, Dt, s));
          ADP_TRY(thin_fold_wgrad(Dt, grads[0].conv_w, L.cin <= 2 ? 2 : 1, 16 * L.cin, s));
        }
      } else if (l == 0) {
        ADP_TRY(first_conv_wgrad(dt, x, at(ws, L.g_e), grads[0].conv_w, B, L.hin, L.hin, L.cin, L.cout, s));
      } else {
        const LevelPlan& I = p.lv[l - 1];
        const bool cen = l == 1 && center;
        ADP_TRY(conv_wgrad(dt, at(ws, L.g_e), L.cout, nullptr, 0, at(ws, I.a), L.cin, grads[l].conv_w, B, L.hout,
                           L.hout, sw, cen ? 1 : 0, deep_level(L.hout)));
        if (cen && !d->training) {   // eval-mode BatchNorm: sum_pixels dL/de = scale * sum gz is not zero
          BnBuf bn = bnbuf(ws, L.bn_down_f, L.cout);
          ADP_TRY(center_wgrad_fix(grads[l].conv_w, reinterpret_cast<const float*>(at(ws, p.center_m)), bn.scale,
                                   reinterpret_cast<const double*>(at(ws, L.bsums_down)), L.cout, L.cin, sw));
        }
        // the consumer of g_a[l - 1] is the next stage's single-launch BatchNorm backward, when it runs in this call
        ConvExtras pex;
        memset(&pex, 0, sizeof(pex));
        ADP_CHECK_ARG(!part_a, "unet_backward: un-finished partial sums without a consumer (level %d)", l);
        if (st + 1 < stage_end && l - 1 >= 1 && I.bn_down && (defer_finish_mask() & 2) &&
            bn_small_ok(dt, (long long)B * I.hout * I.hout, I.cout, 3))
          pex.deferred = &part_a;
        ADP_TRY(conv_parity(dt, at(ws, L.g_e), L.cout, nullptr, 0, params[l].conv_w, tc ? w16(params[l].conv_w_bf16, L.wb_conv) : nullptr,
                            at(ws, I.g_a), B, L.hout, L.hout, L.cin, s, &pex, deep_level(L.hout)));
      }
    }
  }
  return ADP_OK;      // (join_guard joins the side stream)
}

extern "C" int adp_unet_backward(const adp_unet_desc* d, const float* x, const float* y, const float* dy,
                                 const adp_unet_level* params, const adp_unet_level* grads, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  if (!d) { adp_set_error("unet_backward: null descriptor"); return ADP_ERR_ARG; }
  return adp_unet_backward_stages(d, x, y, dy, params, grads, workspace, workspace_bytes, 0, 2 * d->num_downs, stream);
}

// ------------------------------------------------------------------ per-layer C ABI
// (no workspace argument: the tensor-core kernels run without K-splitting here)
extern "C" int adp_weight_operand(const float* w, int R, int C, void* out, void* stream) {
  ADP_CHECK_ARG(w && out && R > 0 && C > 0, "weight_operand: bad arguments");
  ADP_CHECK_ARG(((long long)R * 16 * C) % 4 == 0, "weight_operand: size must be a multiple of 4");
  return cast_f32_to_bf16(w, out, (long long)R * 16 * C, (cudaStream_t)stream);
}

extern "C" int adp_conv2d_k4s2_fprop(int dtype, const void* x, const float* w, const void* w_op, void* y, int B,
                                     int Hin, int Win, int Cin, int Cout, void* stream) {
  ADP_CHECK_ARG(x && w && y, "conv2d_fprop: null pointer");
  tc_set_scratch(nullptr, 0);
  return conv_gather(dtype, x, w, w_op, y, Cout, nullptr, 0, B, Hin, Win, Cin, (cudaStream_t)stream);
}
extern "C" int adp_conv2d_k4s2_dgrad(int dtype, const void* dy, const float* w, const void* w_op, void* dx, int B,
                                     int Hin, int Win, int Cin, int Cout, void* stream) {
  ADP_CHECK_ARG(dy && w && dx, "conv2d_dgrad: null pointer");
  tc_set_scratch(nullptr, 0);
  return conv_parity(dtype, dy, Cout, nullptr, 0, w, w_op, dx, B, Hin / 2, Win / 2, Cin, (cudaStream_t)stream);
}
extern "C" int adp_conv2d_k4s2_wgrad(int dtype, const void* x, const void* dy, float* dw, int B, int Hin, int Win,
                                     int Cin, int Cout, void* stream) {
  ADP_CHECK_ARG(x && dy && dw, "conv2d_wgrad: null pointer");
  return conv_wgrad(dtype, dy, Cout, nullptr, 0, x, Cin, dw, B, Hin / 2, Win / 2, (cudaStream_t)stream);
}
extern "C" int adp_convT2d_k4s2_fprop(int dtype, const void* x0, int c0, const void* x1, int c1, const float* w,
                                      const void* w_op, void* y, int B, int Hin, int Win, int Cout, void* stream) {
  ADP_CHECK_ARG(x0 && w && y && (c1 == 0 || x1), "convT2d_fprop: null pointer");
  tc_set_scratch(nullptr, 0);
  return conv_parity(dtype, x0, c0, x1, c1, w, w_op, y, B, Hin, Win, Cout, (cudaStream_t)stream);
}
extern "C" int adp_convT2d_k4s2_dgrad(int dtype, const void* dy, const float* w, const void* w_op, void* dx0, int c0,
                                      void* dx1, int c1, int B, int Hin, int Win, int Cout, void* stream) {
  ADP_CHECK_ARG(dy && w && dx0 && (c1 == 0 || dx1), "convT2d_dgrad: null pointer");
  tc_set_scratch(nullptr, 0);
  return conv_gather(dtype, dy, w, w_op, dx0, c0, dx1, c1, B, 2 * Hin, 2 * Win, Cout, (cudaStream_t)stream);
}
extern "C" int adp_convT2d_k4s2_wgrad(int dtype, const void* x0, int c0, const void* x1, int c1, const void* dy,
                                      float* dw, int B, int Hin, int Win, int Cout, void* stream) {
  ADP_CHECK_ARG(x0 && dy && dw && (c1 == 0 || x1), "convT2d_wgrad: null pointer");
  return conv_wgrad(dtype, x0, c0, x1, c1, dy, Cout, dw, B, Hin, Win, (cudaStream_t)stream);
}

// ---- thin layers (adp_thin_tc.cu)
extern "C" int adp_first_conv_k4s2_fprop(const float* x, const float* w1, void* w_scratch, void* a, float slope0, void* r,
                                         float slope1, int B, int H, int W, void* stream) {
  ADP_CHECK_ARG(x && w1 && w_scratch && a && r, "first_conv_fprop: null pointer");
  ADP_CHECK_ARG(thin_tc_supported(B, 2, H, W), "first_conv_fprop: unsupported shape %dx%dx%d (power-of-two grid >= 16 wide on sm_100)", B, H, W);
  cudaStream_t s = (cudaStream_t)stream;
  ADP_TRY(thin_pad_rows(w1, w_scratch, 64, 32, 1, s));
  return thin_tc_first_conv(x, w_scratch, a, slope0, r, slope1, nullptr, 0, B, H, W, s);
}
extern "C" int adp_first_conv_k4s2_wgrad(const float* x, const void* g_e, float* dw, int B, int H, int W, void* stream) {
  ADP_CHECK_ARG(x && g_e && dw, "first_conv_wgrad: null pointer");
  ADP_CHECK_ARG(thin_tc_supported(B, 2, H, W), "first_conv_wgrad: unsupported shape %dx%dx%d", B, H, W);
  return thin_tc_first_wgrad(x, g_e, dw, B, H, W, (cudaStream_t)stream);
}
extern "C" int adp_first_conv_k4s2_wgrad_act(const float* x, const void* gA, const void* gB, const void* r, float slope, float* dw,
                                             int B, int H, int W, void* stream) {
  ADP_CHECK_ARG(x && gA && gB && r && dw, "first_conv_wgrad_act: null pointer");
  ADP_CHECK_ARG(thin_tc_supported(B, 2, H, W), "first_conv_wgrad_act: unsupported shape %dx%dx%d", B, H, W);
  return thin_tc_first_wgrad_act(x, gA, gB, r, slope, dw, B, H, W, (cudaStream_t)stream);
}
extern "C" int adp_last_convT_k4s2_dgrad(const float* du, const float* wT, void* w_scratch, void* g0, void* g1, int B, int Hi,
                                         int Wi, void* stream) {
  ADP_CHECK_ARG(du && wT && w_scratch && g0 && g1, "last_convT_dgrad: null pointer");
  ADP_CHECK_ARG(thin_tc_supported(B, 1, 2 * Hi, 2 * Wi), "last_convT_dgrad: unsupported shape %dx%dx%d", B, Hi, Wi);
  cudaStream_t s = (cudaStream_t)stream;
  ADP_TRY(thin_pad_rows(wT, w_scratch, 128, 16, 0, s));
  return thin_tc_last_dgrad(du, w_scratch, g0, g1, B, Hi, Wi, s);
}
extern "C" int adp_last_convT_k4s2_wgrad(const void* x0, const void* x1, const float* x1_scale, const float* x1_shift,
                                         const float* du, float* dw, int B, int Hi, int Wi, void* stream) {
  ADP_CHECK_ARG(x0 && x1 && du && dw && ((x1_scale != nullptr) == (x1_shift != nullptr)), "last_convT_wgrad: bad pointers");
  ADP_CHECK_ARG(thin_tc_supported(B, 1, 2 * Hi, 2 * Wi), "last_convT_wgrad: unsupported shape %dx%dx%d", B, Hi, Wi);
  return thin_tc_last_wgrad(x0, x1, du, dw, B, Hi, Wi, (cudaStream_t)stream, x1_scale, x1_shift);
}
extern "C" int adp_last_convT_k4s2_fprop(const void* x0, const void* x1, const float* x1_scale, const float* x1_shift,
                                         const float* wT, void* w_scratch, const float* bias, int final_sigmoid, float* y,
                                         int B, int Hi, int Wi, void* stream) {
  ADP_CHECK_ARG(x0 && x1 && wT && w_scratch && y && ((x1_scale != nullptr) == (x1_shift != nullptr)), "last_convT_fprop: bad pointers");
  ADP_CHECK_ARG(thin_tc_last_fwd_supported(B, Hi, Wi), "last_convT_fprop: unsupported shape %dx%dx%d (input width must be 128)", B, Hi, Wi);
  cudaStream_t s = (cudaStream_t)stream;
  ADP_TRY(cast_transpose_taps(wT, w_scratch, 128, 1, s));         // [128][16] -> bf16 [16][128]
  return thin_tc_last_fwd(x0, x1, x1_scale, x1_shift, w_scratch, bias, final_sigmoid, y, B, Hi, Wi, s);
}
