// Waveform -> |STFT| -> (log, per-channel min-max) -> antialiased bilinear resize.
//
// Replaces the library ops reached by the reference at
//   dataloader/BatvisionV2_Dataset.py:96-135, :177-185   (T.Spectrogram, log, min-max, Resize)
//   dataloader/BatvisionV1_Dataset.py:70-78,  :86-95     (T.Spectrogram, Resize)
//   dataloader/utils_dataset.py:18-20                    (transforms.Resize((S,S)))
//
// T.Spectrogram(n_fft, win_length, hop, power=1) == torch.stft(center=True, reflect pad,
// periodic Hann of win_length zero-padded (centred) to n_fft, onesided) + abs.  Because the
// window has only win_length non-zero taps, frame t touches x[t*hop + n - win/2], n < win, and
//   S[k,t] = | sum_n hann[n] * xr[t*hop + n - win/2] * exp(-2*pi*i*k*n/n_fft) |
// (the centring offset is a unit phase factor).  This file evaluates that pruned DFT directly:
// one thread owns one bin k and FRAMES_PER_THREAD frames, the windowed twiddle w[n]*W^(kn) is
// read once per tap and reused over the frames, samples are staged in shared memory and read
// as broadcast float4.
#include "adp_common.cuh"

namespace {

constexpr int STFT_THREADS = 256;
constexpr int FRAMES_PER_THREAD = 16;
constexpr int FRAME_GROUPS = 4;
constexpr int FRAMES_PER_BLOCK = FRAMES_PER_THREAD * FRAME_GROUPS;  // 64

// grid: (ceil(T / 64), rows).  smem: twiddle[n_fft] float2 | window[win] | samples[(64-1)*hop + win]
__global__ void __launch_bounds__(STFT_THREADS)
stft_mag_kernel(const float* __restrict__ wave, int L, int pitch, int n_fft, int win, int hop, int T,
                float* __restrict__ spec, int log_mode, int* __restrict__ minmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* tw = reinterpret_cast<float2*>(smem_raw);
  float* wnd = reinterpret_cast<float*>(tw + n_fft);
  float* xs = wnd + win;  // 16-byte aligned: n_fft*8 + win*4 with win % 4 == 0

  const int row = blockIdx.y;
  const int t0 = blockIdx.x * FRAMES_PER_BLOCK;
  const int nsamp = (FRAMES_PER_BLOCK - 1) * hop + win;
  const int F = n_fft / 2 + 1;
  const float* x = wave + (size_t)row * pitch;

  for (int i = threadIdx.x; i < n_fft; i += STFT_THREADS) {
    float s, c;
    sincospif(2.0f * (float)i / (float)n_fft, &s, &c);
    tw[i] = make_float2(c, -s);
  }
  for (int i = threadIdx.x; i < win; i += STFT_THREADS)
    wnd[i] = 0.5f - 0.5f * cospif(2.0f * (float)i / (float)win);
  for (int i = threadIdx.x; i < nsamp; i += STFT_THREADS) {
    int j = t0 * hop + i - win / 2;
    if (j < 0) j = -j;
    if (j >= L) j = 2 * (L - 1) - j;
    j = min(max(j, 0), L - 1);
    xs[i] = x[j];
  }
  __syncthreads();

  float vmin = INFINITY, vmax = -INFINITY;
  const int items = F * FRAME_GROUPS;
  for (int item = threadIdx.x; item < items; item += STFT_THREADS) {
    const int g = item / F;  // frame group; consecutive lanes -> consecutive k
    const int k = item - g * F;
    const int tb = g * FRAMES_PER_THREAD;
    if (t0 + tb >= T) continue;
    float re[FRAMES_PER_THREAD], im[FRAMES_PER_THREAD];
#pragma unroll
    for (int f = 0; f < FRAMES_PER_THREAD; ++f) re[f] = im[f] = 0.f;
    const float* xb = xs + tb * hop;
    int idx = 0;  // (k * n) mod n_fft
    for (int n = 0; n < win; n += 4) {
      float2 w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float2 t = tw[idx];
        float wn = wnd[n + q];
        w[q] = make_float2(t.x * wn, t.y * wn);
        idx += k;
        if (idx >= n_fft) idx -= n_fft;
      }
#pragma unroll
      for (int f = 0; f < FRAMES_PER_THREAD; ++f) {
        float4 v = *reinterpret_cast<const float4*>(xb + f * hop + n);
        re[f] = fmaf(v.x, w[0].x, re[f]);
        im[f] = fmaf(v.x, w[0].y, im[f]);
        re[f] = fmaf(v.y, w[1].x, re[f]);
        im[f] = fmaf(v.y, w[1].y, im[f]);
        re[f] = fmaf(v.z, w[2].x, re[f]);
        im[f] = fmaf(v.z, w[2].y, im[f]);
        re[f] = fmaf(v.w, w[3].x, re[f]);
        im[f] = fmaf(v.w, w[3].y, im[f]);
      }
    }
    float* out = spec + ((size_t)row * F + k) * T + t0 + tb;
#pragma unroll
    for (int f = 0; f < FRAMES_PER_THREAD; ++f) {
      if (t0 + tb + f < T) {
        float m = sqrtf(re[f] * re[f] + im[f] * im[f]);
        if (log_mode) {
          m = logf(m + 1e-8f);
          vmin = fminf(vmin, m);
          vmax = fmaxf(vmax, m);
        }
        out[f] = m;
      }
    }
  }
  if (log_mode) {
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0 && vmin <= vmax) {
      atomicMin(&minmax[2 * row], float_to_ordered(vmin));
      atomicMax(&minmax[2 * row + 1], float_to_ordered(vmax));
    }
  }
}

__global__ void init_minmax_kernel(int* mm, int rows) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) {
    mm[2 * i] = float_to_ordered(INFINITY);
    mm[2 * i + 1] = float_to_ordered(-INFINITY);
  }
}

// torchaudio.functional.melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate, norm=None, mel_scale="htk")
// followed by MelScale's matmul: mel[m,t] = sum_f spec[f,t] * fb[f,m]   (BatvisionV2_Dataset.py:187-197).
// grid (ceil(T/64), rows); smem: fb[F][n_mels] | f_pts[n_mels+2] | band lo/hi [2*n_mels].  The triangular bank is
// rebuilt per block (a few thousand flops) so that the call needs no persistent state.
constexpr int MEL_THREADS = 256;
constexpr int MEL_FRAMES = 64;

__global__ void __launch_bounds__(MEL_THREADS)
mel_contract_kernel(const float* __restrict__ spec, int F, int T, int n_mels, float nyquist, float f_min, float f_max,
                    float* __restrict__ mel, int log_mode, int* __restrict__ minmax) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* fb = reinterpret_cast<float*>(smem_raw);
  float* fpts = fb + (size_t)F * n_mels;
  int* band = reinterpret_cast<int*>(fpts + n_mels + 2);
  const int row = blockIdx.y, t0 = blockIdx.x * MEL_FRAMES;

  for (int i = threadIdx.x; i < n_mels + 2; i += MEL_THREADS) {
    const double m_lo = 2595.0 * log10(1.0 + (double)f_min / 700.0), m_hi = 2595.0 * log10(1.0 + (double)f_max / 700.0);
    const float m = (float)(m_lo + (m_hi - m_lo) * (double)i / (double)(n_mels + 1));
    fpts[i] = (float)(700.0 * (pow(10.0, (double)m / 2595.0) - 1.0));
  }
  __syncthreads();
  const float fstep = nyquist / (float)(F - 1);
  for (int i = threadIdx.x; i < F * n_mels; i += MEL_THREADS) {
    const int f = i / n_mels, m = i - f * n_mels;
    const float freq = f < F / 2 ? fstep * (float)f : nyquist - fstep * (float)(F - 1 - f);    // torch.linspace's rule
    const float down = (freq - fpts[m]) / (fpts[m + 1] - fpts[m]);
    const float up = (fpts[m + 2] - freq) / (fpts[m + 2] - fpts[m + 1]);
    fb[i] = fmaxf(0.f, fminf(down, up));
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n_mels; m += MEL_THREADS) {
    int lo = F, hi = -1;
    for (int f = 0; f < F; ++f)
      if (fb[f * n_mels + m] != 0.f) { lo = min(lo, f); hi = f; }
    band[2 * m] = lo;
    band[2 * m + 1] = hi;
  }
  __syncthreads();

  const int t = t0 + (threadIdx.x & (MEL_FRAMES - 1));
  float vmin = INFINITY, vmax = -INFINITY;
  if (t < T) {
    const float* src = spec + (size_t)row * F * T + t;
    for (int m = threadIdx.x / MEL_FRAMES; m < n_mels; m += MEL_THREADS / MEL_FRAMES) {
      float acc = 0.f;
      for (int f = band[2 * m]; f <= band[2 * m + 1]; ++f) acc = fmaf(src[(size_t)f * T], fb[f * n_mels + m], acc);
      if (log_mode) {
        acc = logf(acc + 1e-8f);
        vmin = fminf(vmin, acc);
        vmax = fmaxf(vmax, acc);
      }
      mel[((size_t)row * n_mels + m) * T + t] = acc;
    }
  }
  if (log_mode) {
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0 && vmin <= vmax) {
      atomicMin(&minmax[2 * row], float_to_ordered(vmin));
      atomicMax(&minmax[2 * row + 1], float_to_ordered(vmax));
    }
  }
}

// aten::_upsample_bilinear2d_aa index/weight rule for one axis (align_corners = False)
struct AxisTaps {
  int lo, size;
  float center, invscale, total;
};
__device__ __forceinline__ float aa_w(const AxisTaps& a, int j) {
  float t = fabsf(((float)(j + a.lo) - a.center + 0.5f) * a.invscale);
  return t < 1.f ? 1.f - t : 0.f;
}
__device__ __forceinline__ AxisTaps aa_axis(int i, int n_in, int n_out) {
  AxisTaps a;
  float scale = (float)n_in / (float)n_out;
  float support = scale >= 1.f ? scale : 1.f;
  a.invscale = scale >= 1.f ? 1.f / scale : 1.f;
  a.center = scale * ((float)i + 0.5f);
  a.lo = max((int)(a.center - support + 0.5f), 0);
  a.size = min((int)(a.center + support + 0.5f), n_in) - a.lo;
  float tot = 0.f;
  for (int j = 0; j < a.size; ++j) tot += aa_w(a, j);
  a.total = tot;
  return a;
}

// in [rows,H,W] -> out [rows,S,S]; optional per-row (x - min)/(max - min).  The taps of an output pixel depend on
// (p, o) only, so a thread computes them once (registers) and applies them to RESIZE_PLANES consecutive planes.
constexpr int RESIZE_PLANES = 8;
constexpr int RESIZE_HT = 6, RESIZE_VT = 4;   // register-resident taps; larger footprints take the generic loop

__global__ void __launch_bounds__(256)
resize_aa_kernel(const float* __restrict__ in, int rows, int H, int W, int S, float* __restrict__ out,
                 const int* __restrict__ minmax) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int p = blockIdx.y;
  if (o >= S) return;
  const AxisTaps ah = aa_axis(p, H, S);
  const AxisTaps aw = aa_axis(o, W, S);
  const float ih = ah.total != 0.f ? 1.f / ah.total : 1.f, iw = aw.total != 0.f ? 1.f / aw.total : 1.f;
  const bool fast = aw.size <= RESIZE_HT && ah.size <= RESIZE_VT;
  float wh[RESIZE_VT], ww[RESIZE_HT];
  int cw[RESIZE_HT];
#pragma unroll
  for (int j = 0; j < RESIZE_VT; ++j) wh[j] = j < ah.size ? aa_w(ah, j) * ih : 0.f;
#pragma unroll
  for (int j = 0; j < RESIZE_HT; ++j) {
    ww[j] = j < aw.size ? aa_w(aw, j) * iw : 0.f;
    cw[j] = min(aw.lo + j, W - 1);          // clamped: taps beyond the footprint carry weight 0
  }
  const int row0 = blockIdx.z * RESIZE_PLANES;
  if (fast) {
    // all planes of this thread at once: RESIZE_PLANES x (<= 4 x 6) independent loads in flight before the first use
    float acc[RESIZE_PLANES];
#pragma unroll
    for (int k = 0; k < RESIZE_PLANES; ++k) {
      acc[k] = 0.f;
      const int row = min(row0 + k, rows - 1);                     // (clamped: the surplus planes of the last block are not stored)
      const float* src = in + (size_t)row * H * W;
#pragma unroll
      for (int jh = 0; jh < RESIZE_VT; ++jh) {
        if (jh < ah.size) {
          const float* line = src + (size_t)(ah.lo + jh) * W;
          float hacc = 0.f;
#pragma unroll
          for (int jw = 0; jw < RESIZE_HT; ++jw)
            if (jw < aw.size) hacc = fmaf(__ldg(line + cw[jw]), ww[jw], hacc);
          acc[k] = fmaf(hacc, wh[jh], acc[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RESIZE_PLANES; ++k) {
      const int row = row0 + k;
      if (row < rows) {
        float v = acc[k];
        if (minmax) {
          // the weights sum to one, so the per-channel min-max commutes with the resize
          const float mn = ordered_to_float(minmax[2 * row]), mx = ordered_to_float(minmax[2 * row + 1]);
          v = mx > mn ? (v - mn) / (mx - mn) : 0.f;
        }
        out[((size_t)row * S + p) * S + o] = v;
      }
    }
    return;
  }
  const int row_end = min(rows, row0 + RESIZE_PLANES);
  for (int row = row0; row < row_end; ++row) {
    const float* src = in + (size_t)row * H * W;
    float acc = 0.f;
    for (int jh = 0; jh < ah.size; ++jh) {
      const float* line = src + (size_t)(ah.lo + jh) * W + aw.lo;
      float hacc = 0.f;
      for (int jw = 0; jw < aw.size; ++jw) hacc = fmaf(line[jw], aa_w(aw, jw) * iw, hacc);
      acc = fmaf(hacc, aa_w(ah, jh) * ih, acc);
    }
    if (minmax) {
      const float mn = ordered_to_float(minmax[2 * row]), mx = ordered_to_float(minmax[2 * row + 1]);
      acc = mx > mn ? (acc - mn) / (mx - mn) : 0.f;
    }
    out[((size_t)row * S + p) * S + o] = acc;
  }
}

int check_stft_args(int rows, int L, int pitch, int n_fft, int win, int hop) {
  ADP_CHECK_ARG(rows > 0 && L > 0 && pitch >= L, "stft: bad rows/L/pitch (%d,%d,%d)", rows, L, pitch);
  ADP_CHECK_ARG(n_fft >= 8 && n_fft <= 4096 && n_fft % 2 == 0, "stft: n_fft %d unsupported", n_fft);
  ADP_CHECK_ARG(win > 0 && win <= n_fft && win % 4 == 0, "stft: win_length %d must be a multiple of 4 and <= n_fft", win);
  ADP_CHECK_ARG(hop > 0 && hop % 4 == 0, "stft: hop %d must be a positive multiple of 4", hop);
  ADP_CHECK_ARG(n_fft / 2 < L, "stft: reflect padding needs n_fft/2 < L");
  return ADP_OK;
}

size_t mel_smem(int F, int n_mels) { return ((size_t)F * n_mels + n_mels + 2) * 4 + (size_t)n_mels * 8; }

size_t stft_smem(int n_fft, int win, int hop) {
  return (size_t)n_fft * 8 + (size_t)win * 4 + (size_t)((FRAMES_PER_BLOCK - 1) * hop + win) * 4;
}

int launch_stft(const float* wave, int rows, int L, int pitch, int n_fft, int win, int hop, float* spec,
                int log_mode, int* minmax, cudaStream_t s) {
  const int T = 1 + L / hop;
  size_t smem = stft_smem(n_fft, win, hop);
  ADP_CHECK_ARG(smem <= 200 * 1024, "stft: shared memory %zu too large", smem);
  if (smem > 48 * 1024)
    ADP_CUDA(cudaFuncSetAttribute(stft_mag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(adp_cdiv(T, FRAMES_PER_BLOCK), rows);
  stft_mag_kernel<<<grid, STFT_THREADS, smem, s>>>(wave, L, pitch, n_fft, win, hop, T, spec, log_mode, minmax);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

}  // namespace

extern "C" int adp_stft_mag(const float* wave, int rows, int L, int wave_pitch, int n_fft, int win_length,
                            int hop, float* spec, void* stream) {
  ADP_TRY(check_stft_args(rows, L, wave_pitch, n_fft, win_length, hop));
  ADP_CHECK_ARG(wave && spec, "stft: null pointer");
  return launch_stft(wave, rows, L, wave_pitch, n_fft, win_length, hop, spec, 0, nullptr, (cudaStream_t)stream);
}

extern "C" size_t adp_feature_workspace_bytes(int rows, int L, int n_fft, int hop) {
  if (rows <= 0 || L <= 0 || n_fft <= 0 || hop <= 0) return 0;
  size_t T = 1 + (size_t)L / hop, F = (size_t)n_fft / 2 + 1;
  return adp_align_up((size_t)rows * F * T * 4, 1024) + adp_align_up((size_t)rows * 8, 1024) +
         adp::tc_stft_workspace_bytes(rows, L, n_fft, hop);
}

extern "C" int adp_resize_aa(const float* in, int rows, int H, int W, int out_size, float* out, void* stream) {
  ADP_CHECK_ARG(in && out && rows > 0 && H > 0 && W > 0 && out_size > 0, "resize: bad arguments");
  ADP_CHECK_ARG(rows <= 65535 && out_size <= 65535, "resize: rows/out_size too large for one launch");
  dim3 grid(adp_cdiv(out_size, 256), out_size, adp_cdiv(rows, RESIZE_PLANES));
  resize_aa_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, rows, H, W, out_size, out, nullptr);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

namespace {

struct FeatureWs {
  float* spec;
  int* minmax;
  void* tc_ws;
  float* mel;
};

FeatureWs carve_feature_ws(void* workspace, int rows, int L, int n_fft, int hop) {
  const size_t T = 1 + (size_t)L / hop, F = (size_t)n_fft / 2 + 1;
  char* wsb = reinterpret_cast<char*>(workspace);
  FeatureWs w;
  w.spec = reinterpret_cast<float*>(wsb);
  w.minmax = reinterpret_cast<int*>(wsb + adp_align_up((size_t)rows * F * T * 4, 1024));
  w.tc_ws = wsb + adp_align_up((size_t)rows * F * T * 4, 1024) + adp_align_up((size_t)rows * 8, 1024);
  w.mel = reinterpret_cast<float*>(wsb + adp_feature_workspace_bytes(rows, L, n_fft, hop));
  return w;
}

int run_stft(const float* wave, int rows, int L, int pitch, int n_fft, int win, int hop, const FeatureWs& w, int log_mode,
             cudaStream_t s) {
  if (adp::tc_enabled() && adp::tc_supported_stft(rows, L, n_fft, win, hop))
    return adp::tc_stft_mag(wave, rows, L, pitch, n_fft, hop, w.spec, log_mode, w.minmax, w.tc_ws, s);
  return launch_stft(wave, rows, L, pitch, n_fft, win, hop, w.spec, log_mode, w.minmax, s);
}

int check_mel_args(int n_fft, int n_mels, float sample_rate, float f_min, float f_max) {
  ADP_CHECK_ARG(n_mels >= 1 && n_mels <= 512, "mel: n_mels %d unsupported", n_mels);
  ADP_CHECK_ARG(sample_rate > 0.f && f_min >= 0.f && f_max > f_min, "mel: bad sample_rate/f_min/f_max");
  ADP_CHECK_ARG(mel_smem(n_fft / 2 + 1, n_mels) <= 200 * 1024, "mel: filterbank %d x %d does not fit shared memory",
                n_fft / 2 + 1, n_mels);
  return ADP_OK;
}

int launch_mel(const float* spec, int rows, int F, int T, int n_mels, float sample_rate, float f_min, float f_max, float* mel,
               int log_mode, int* minmax, cudaStream_t s) {
  const size_t smem = mel_smem(F, n_mels);
  if (smem > 48 * 1024)
    ADP_CUDA(cudaFuncSetAttribute(mel_contract_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const float nyquist = (float)((long long)sample_rate / 2);      // MelScale uses sample_rate // 2
  dim3 grid(adp_cdiv(T, MEL_FRAMES), rows);
  mel_contract_kernel<<<grid, MEL_THREADS, smem, s>>>(spec, F, T, n_mels, nyquist, f_min, f_max, mel, log_mode, minmax);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

}  // namespace

extern "C" int adp_feature_forward(const float* wave, int rows, int L, int wave_pitch, int n_fft, int win_length,
                                   int hop, int log_minmax, int out_size, float* out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ADP_TRY(check_stft_args(rows, L, wave_pitch, n_fft, win_length, hop));
  ADP_CHECK_ARG(wave && out && workspace, "feature: null pointer");
  ADP_CHECK_ARG(out_size > 0 && out_size <= 65535 && rows <= 65535, "feature: bad out_size/rows");
  adp::ProfScope prof(adp::PROF_FEATURE, s, (double)rows * ((double)L + (double)out_size * out_size) * 4.0);   // waveform in, feature out
  ADP_CHECK_ARG(workspace_bytes >= adp_feature_workspace_bytes(rows, L, n_fft, hop),
                "feature: workspace too small (%zu)", workspace_bytes);
  const int T = 1 + L / hop, F = n_fft / 2 + 1;
  const FeatureWs w = carve_feature_ws(workspace, rows, L, n_fft, hop);
  if (log_minmax) {
    init_minmax_kernel<<<adp_cdiv(rows, 256), 256, 0, s>>>(w.minmax, rows);
    ADP_LAUNCH_CHECK();
  }
  ADP_TRY(run_stft(wave, rows, L, wave_pitch, n_fft, win_length, hop, w, log_minmax ? 1 : 0, s));
  dim3 grid(adp_cdiv(out_size, 256), out_size, adp_cdiv(rows, RESIZE_PLANES));
  resize_aa_kernel<<<grid, 256, 0, s>>>(w.spec, rows, F, T, out_size, out, log_minmax ? w.minmax : nullptr);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

extern "C" size_t adp_feature_mel_workspace_bytes(int rows, int L, int n_fft, int hop, int n_mels) {
  if (rows <= 0 || L <= 0 || n_fft <= 0 || hop <= 0 || n_mels <= 0) return 0;
  return adp_feature_workspace_bytes(rows, L, n_fft, hop) + adp_align_up((size_t)rows * n_mels * (1 + (size_t)L / hop) * 4, 1024);
}

extern "C" int adp_mel_spectrogram(const float* wave, int rows, int L, int wave_pitch, int n_fft, int win_length, int hop,
                                   int n_mels, float sample_rate, float f_min, float f_max, float* mel, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ADP_TRY(check_stft_args(rows, L, wave_pitch, n_fft, win_length, hop));
  ADP_TRY(check_mel_args(n_fft, n_mels, sample_rate, f_min, f_max));
  ADP_CHECK_ARG(wave && mel && workspace, "mel: null pointer");
  ADP_CHECK_ARG(rows <= 65535, "mel: too many rows");
  ADP_CHECK_ARG(workspace_bytes >= adp_feature_workspace_bytes(rows, L, n_fft, hop), "mel: workspace too small (%zu)",
                workspace_bytes);
  const FeatureWs w = carve_feature_ws(workspace, rows, L, n_fft, hop);
  ADP_TRY(run_stft(wave, rows, L, wave_pitch, n_fft, win_length, hop, w, 0, s));
  return launch_mel(w.spec, rows, n_fft / 2 + 1, 1 + L / hop, n_mels, sample_rate, f_min, f_max, mel, 0, nullptr, s);
}

extern "C" int adp_feature_forward_mel(const float* wave, int rows, int L, int wave_pitch, int n_fft, int win_length,
                                       int hop, int n_mels, float sample_rate, float f_min, float f_max, int log_minmax,
                                       int out_size, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  ADP_TRY(check_stft_args(rows, L, wave_pitch, n_fft, win_length, hop));
  ADP_TRY(check_mel_args(n_fft, n_mels, sample_rate, f_min, f_max));
  ADP_CHECK_ARG(wave && out && workspace, "feature_mel: null pointer");
  ADP_CHECK_ARG(out_size > 0 && out_size <= 65535 && rows <= 65535, "feature_mel: bad out_size/rows");
  ADP_CHECK_ARG(workspace_bytes >= adp_feature_mel_workspace_bytes(rows, L, n_fft, hop, n_mels),
                "feature_mel: workspace too small (%zu)", workspace_bytes);
  const int T = 1 + L / hop, F = n_fft / 2 + 1;
  const FeatureWs w = carve_feature_ws(workspace, rows, L, n_fft, hop);
  if (log_minmax) {
    init_minmax_kernel<<<adp_cdiv(rows, 256), 256, 0, s>>>(w.minmax, rows);
    ADP_LAUNCH_CHECK();
  }
  ADP_TRY(run_stft(wave, rows, L, wave_pitch, n_fft, win_length, hop, w, 0, s));   // log/min-max follow the mel contraction
  ADP_TRY(launch_mel(w.spec, rows, F, T, n_mels, sample_rate, f_min, f_max, w.mel, log_minmax ? 1 : 0, w.minmax, s));
  dim3 grid(adp_cdiv(out_size, 256), out_size, adp_cdiv(rows, RESIZE_PLANES));
  resize_aa_kernel<<<grid, 256, 0, s>>>(w.mel, rows, n_mels, T, out_size, out, log_minmax ? w.minmax : nullptr);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Ground-truth depth preparation (BatvisionV2_Dataset.py:68-78, BatvisionV1_Dataset.py:47-65): mm -> m, clip to
// [0, max_depth], cv2.INTER_NEAREST resize to S x S, optional / max_depth.  Every step but the resize is
// element-wise, so the kernel gathers first; fp32 IEEE division keeps it bit-exact with the numpy code.
namespace {

template <class T>
__global__ void __launch_bounds__(256)
depth_prepare_kernel(const T* __restrict__ raw, int H, int W, int S, float max_depth, int nan_to_num, float norm_div,
                     float* __restrict__ out) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y, row = blockIdx.z;
  if (o >= S) return;
  // cv2 resizeNN: sx = min(floor(dx * (1 / ((double)S / W))), W - 1)
  const double ify = 1.0 / ((double)S / (double)H), ifx = 1.0 / ((double)S / (double)W);
  const int sy = min((int)floor((double)p * ify), H - 1), sx = min((int)floor((double)o * ifx), W - 1);
  float d = (float)raw[((size_t)row * H + sy) * W + sx];
  if (nan_to_num) {                                   // np.nan_to_num: nan -> 0, +-inf -> +-float max (V1 :49-52)
    if (d != d) d = 0.f;
    else if (isinf(d)) d = d > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  }
  d = d / 1000.0f;
  if (max_depth > 0.f && d > max_depth) d = max_depth;
  if (d < 0.f) d = 0.f;
  if (norm_div > 0.f) d = d / norm_div;
  out[((size_t)row * S + p) * S + o] = d;
}

}  // namespace

extern "C" int adp_depth_prepare(const void* raw, int raw_dtype, int rows, int H, int W, int out_size, float max_depth,
                                 int nan_to_num, float norm_div, float* out, void* stream) {
  ADP_CHECK_ARG(raw && out && rows > 0 && rows <= 65535 && H > 0 && W > 0 && out_size > 0 && out_size <= 65535,
                "depth_prepare: bad arguments");
  ADP_CHECK_ARG(raw_dtype == 0 || raw_dtype == 1, "depth_prepare: raw_dtype %d (0 = fp32, 1 = uint16)", raw_dtype);
  dim3 grid(adp_cdiv(out_size, 256), out_size, rows);
  cudaStream_t s = (cudaStream_t)stream;
  if (raw_dtype == 0)
    depth_prepare_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(raw), H, W, out_size, max_depth,
                                                     nan_to_num, norm_div, out);
  else
    depth_prepare_kernel<unsigned short><<<grid, 256, 0, s>>>(reinterpret_cast<const unsigned short*>(raw), H, W, out_size,
                                                              max_depth, 0, norm_div, out);
  ADP_LAUNCH_CHECK();
  return ADP_OK;
}
