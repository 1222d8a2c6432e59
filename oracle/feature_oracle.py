"""numpy restatement of the reference's waveform -> [2,S,S] feature transform.

Test infrastructure only (see oracle/__init__.py).  Parity: pinned against the
reference's torchaudio / torchvision call chain by tests/test_oracle_golden.py.

The arithmetic the reference reaches lives in third-party libraries that are not
under /root/reference and that the reference does not pin (no requirements
file): torchaudio.transforms.Spectrogram -> torch.stft (torchaudio 2.11.0 /
torch 2.11.0 in this image) and torchvision.transforms.Resize ->
aten::_upsample_bilinear2d_aa (torchvision 0.26.0).  Their published algorithms
are restated here; parity is anchored on the reference's own call sites:

* dataloader/BatvisionV2_Dataset.py:96-135, :177-185  (cut, STFT params, log,
  per-channel min-max, Resize)
* dataloader/BatvisionV2_Dataset.py:111-114, :187-197 (mel branch: T.MelSpectrogram -> MelScale ->
  torchaudio.functional.melscale_fbanks, HTK scale, norm=None)
* dataloader/BatvisionV1_Dataset.py:70-78, :86-95     (STFT, Resize; no log)
* dataloader/utils_dataset.py:10-28                   (Resize((S,S)))
"""
import numpy as np


def stft_params(max_depth):
    """(n_fft, win_length, hop_length) -- BatvisionV2_Dataset.py:96-108."""
    if max_depth:
        return 512, 64, 64 // 4
    return 400, 200, 100


def cut_length(max_depth, sr=44100):
    """Samples kept by the V2 cut -- BatvisionV2_Dataset.py:102-104."""
    return int((2 * max_depth / 340) * sr)


def hann_periodic(win_length):
    n = np.arange(win_length, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)


def stft_mag(wave, n_fft=400, win_length=400, hop_length=100):
    """|STFT| as T.Spectrogram(n_fft, win_length, power=1.0, hop_length) computes it.

    Follows _get_spectrogram (BatvisionV2_Dataset.py:177-185,
    BatvisionV1_Dataset.py:86-95): periodic Hann window of win_length zero-padded
    (centred) to n_fft, center=True with reflect padding n_fft//2, onesided,
    not normalised, magnitude (power=1).  wave: [..., L] -> [..., n_fft//2+1, 1+L//hop].
    Evaluated in float64 and rounded to float32.
    """
    wave = np.asarray(wave)
    lead = wave.shape[:-1]
    x = wave.reshape(-1, wave.shape[-1]).astype(np.float64)
    L = x.shape[-1]
    pad = n_fft // 2
    xp = np.pad(x, ((0, 0), (pad, pad)), mode="reflect")
    T = 1 + L // hop_length
    left = (n_fft - win_length) // 2
    win = np.zeros(n_fft, dtype=np.float64)
    win[left:left + win_length] = hann_periodic(win_length)
    idx = np.arange(T)[:, None] * hop_length + np.arange(n_fft)[None, :]
    frames = xp[:, idx] * win                      # [C, T, n_fft]
    # only the window support contributes; restrict the DFT to it
    n = np.arange(left, left + win_length)
    k = np.arange(n_fft // 2 + 1)
    ang = -2.0 * np.pi * np.outer(n, k) / n_fft    # [win, F]
    fr = frames[:, :, left:left + win_length]
    re = fr @ np.cos(ang)
    im = fr @ np.sin(ang)
    mag = np.sqrt(re * re + im * im)               # [C, T, F]
    out = np.transpose(mag, (0, 2, 1)).astype(np.float32)
    return out.reshape(*lead, n_fft // 2 + 1, T)


def log_minmax(spec):
    """log(spec+1e-8) then per-channel min-max -- BatvisionV2_Dataset.py:122-132.

    spec: [C, F, T] float32.  fp32 arithmetic like the reference.
    """
    y = np.log(spec.astype(np.float32) + np.float32(1e-8)).astype(np.float32)
    out = np.empty_like(y)
    for c in range(y.shape[0]):
        lo = y[c].min()
        hi = y[c].max()
        if hi > lo:
            out[c] = (y[c] - lo) / (hi - lo)
        else:
            out[c] = 0.0
    return out


def aa_resize_weights(n_in, n_out):
    """Index/weight table of aten::_upsample_bilinear2d_aa for one axis
    (align_corners=False), the op transforms.Resize reaches for tensors
    (utils_dataset.py:18-20).  Returns (lo[n_out] int32, size[n_out] int32,
    w[n_out, K] float32 zero-padded)."""
    scale = n_in / n_out
    support = scale if scale >= 1.0 else 1.0
    inv = 1.0 / scale if scale >= 1.0 else 1.0
    K = int(np.ceil(support)) * 2 + 1
    lo = np.zeros(n_out, dtype=np.int32)
    size = np.zeros(n_out, dtype=np.int32)
    w = np.zeros((n_out, K), dtype=np.float32)
    for i in range(n_out):
        c = scale * (i + 0.5)
        xmin = max(int(c - support + 0.5), 0)
        xsize = min(int(c + support + 0.5), n_in) - xmin
        ws = np.zeros(K, dtype=np.float64)
        for j in range(xsize):
            t = abs((j + xmin - c + 0.5) * inv)
            ws[j] = 1.0 - t if t < 1.0 else 0.0
        tot = ws.sum()
        if tot != 0.0:
            ws /= tot
        lo[i] = xmin
        size[i] = xsize
        w[i] = ws.astype(np.float32)
    return lo, size, w


def aa_resize_matrix(n_in, n_out):
    lo, size, w = aa_resize_weights(n_in, n_out)
    m = np.zeros((n_out, n_in), dtype=np.float32)
    for i in range(n_out):
        m[i, lo[i]:lo[i] + size[i]] = w[i, :size[i]]
    return m


def resize(spec, out_size):
    """transforms.Resize((S,S)) on a float tensor [C,H,W] (antialiased bilinear).
    ATen runs the horizontal pass first, then the vertical one."""
    C, H, W = spec.shape
    mw = aa_resize_matrix(W, out_size)
    mh = aa_resize_matrix(H, out_size)
    tmp = np.einsum("chw,ow->cho", spec.astype(np.float32), mw).astype(np.float32)
    return np.einsum("cho,ph->cpo", tmp, mh).astype(np.float32)


def feature_v2(wave, max_depth=30.0, images_size=256, sr=44100, do_resize=True):
    """BatvisionV2Dataset.__getitem__ audio branch, 'spectrogram' format
    (BatvisionV2_Dataset.py:92-135).  wave [2, L_raw] -> [2, S, S]."""
    n_fft, win, hop = stft_params(max_depth)
    if max_depth:
        wave = wave[:, :cut_length(max_depth, sr)]
    spec = stft_mag(wave, n_fft, win, hop)
    spec = log_minmax(spec)
    return resize(spec, images_size) if do_resize else spec


def feature_v1(wave, images_size=256, do_resize=True):
    """BatvisionV1Dataset.__getitem__ audio branch (BatvisionV1_Dataset.py:68-78):
    Spectrogram(512, 64, hop 16), no log, no min-max, Resize."""
    spec = stft_mag(wave, 512, 64, 16)
    return resize(spec, images_size) if do_resize else spec


def mel_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate):
    """torchaudio.functional.melscale_fbanks(..., norm=None, mel_scale='htk') in float32, the bank
    T.MelSpectrogram builds for _get_melspectrogram (BatvisionV2_Dataset.py:187-197).  [n_freqs, n_mels]."""
    f32 = np.float32
    all_freqs = np.linspace(0, sample_rate // 2, n_freqs).astype(f32)
    m_min = 2595.0 * np.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * np.log10(1.0 + f_max / 700.0)
    m_pts = np.linspace(m_min, m_max, n_mels + 2).astype(f32)
    f_pts = (f32(700.0) * (np.power(f32(10.0), m_pts / f32(2595.0)) - f32(1.0))).astype(f32)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return np.maximum(f32(0.0), np.minimum(down, up)).astype(f32)


def mel_spectrogram(wave, n_fft=400, win_length=400, f_min=20.0, f_max=20000.0, n_mels=32, sample_rate=44100,
                    hop_length=None):
    """T.MelSpectrogram(sample_rate, n_fft, win_length, power=1.0, f_min, f_max, n_mels): hop defaults to
    win_length // 2; mel = (spec^T @ fb)^T.  wave [..., L] -> [..., n_mels, 1 + L // hop]."""
    hop = hop_length if hop_length else win_length // 2
    spec = stft_mag(wave, n_fft, win_length, hop).astype(np.float64)
    fb = mel_fbanks(n_fft // 2 + 1, f_min, f_max, n_mels, sample_rate).astype(np.float64)
    return np.einsum("...ft,fm->...mt", spec, fb).astype(np.float32)


def feature_v2_mel(wave, max_depth=30.0, images_size=256, sr=44100, do_resize=True):
    """BatvisionV2Dataset.__getitem__ audio branch, 'mel_spectrogram' format (BatvisionV2_Dataset.py:92-135)."""
    n_fft, win, _ = stft_params(max_depth)
    if max_depth:
        wave = wave[:, :cut_length(max_depth, sr)]
    spec = log_minmax(mel_spectrogram(wave, n_fft, win))
    return resize(spec, images_size) if do_resize else spec
