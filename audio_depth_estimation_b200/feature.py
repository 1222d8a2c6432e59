"""Batched GPU feature transform: binaural waveform -> [B,2,S,S] network input.

Host-side mirror of the audio branch of the reference datasets, running on the collated
batch through libadp_b200 (adp_stft_mag / adp_mel_spectrogram / adp_feature_forward[_mel] / adp_resize_aa):
  dataloader/BatvisionV2_Dataset.py:92-137, :177-197   (cut, STFT params, [mel bank,] log, min-max, Resize)
  dataloader/BatvisionV1_Dataset.py:68-81,  :86-95     (STFT, Resize; no log / min-max)
  dataloader/utils_dataset.py:10-28                    (Resize((S,S)), antialiased bilinear)
"""
import torch

from . import _lib


def stft_params(max_depth):
    """(n_fft, win_length, hop_length) chosen at BatvisionV2_Dataset.py:96-108."""
    if max_depth:
        return 512, 64, 64 // 4
    return 400, 200, 100


def cut_length(max_depth, sr=44100):
    """Number of samples kept by the V2 cut (BatvisionV2_Dataset.py:102-104)."""
    return int((2 * max_depth / 340) * sr)


def _rows(wave):
    _lib.require_cuda(wave, "waveform", torch.float32)
    if wave.dim() < 1 or wave.stride(-1) != 1:
        wave = wave.contiguous()
    lead = wave.shape[:-1]
    flat = wave.reshape(-1, wave.shape[-1])
    if flat.stride(-1) != 1 or (flat.shape[0] > 1 and flat.stride(0) < flat.shape[1]):
        flat = flat.contiguous()
    return flat, lead


def spectrogram(waveform, n_fft=400, power=1.0, win_length=400, hop_length=100, length=None):
    """|STFT| with the semantics of T.Spectrogram(n_fft, win_length, power=1.0, hop_length)
    (the reference's _get_spectrogram).  waveform [..., L] CUDA fp32 -> [..., n_fft//2+1, 1+L//hop].
    `length` < L evaluates the transform on the first `length` samples without a copy."""
    if power != 1.0:
        raise NotImplementedError("only power=1.0 (magnitude) is on the hot path (BatvisionV2_Dataset.py:117)")
    flat, lead = _rows(waveform)
    rows, full = flat.shape
    L = full if length is None else min(int(length), full)
    pitch = flat.stride(0) if rows > 1 else full
    T = 1 + L // hop_length
    spec = torch.empty((rows, n_fft // 2 + 1, T), device=flat.device, dtype=torch.float32)
    lib = _lib.load()
    with _lib.on_device(flat):
        _lib.check(lib.adp_stft_mag(flat.data_ptr(), rows, L, pitch, n_fft, win_length, hop_length,
                                    spec.data_ptr(), _lib.stream_ptr()))
    return spec.reshape(*lead, n_fft // 2 + 1, T)


MEL_DEFAULTS = dict(sample_rate=44100, f_min=20.0, f_max=20000.0, n_mels=32)      # BatvisionV2_Dataset.py:187-196


def melspectrogram(waveform, n_fft=400, power=1.0, win_length=400, f_min=20.0, f_max=20000.0, n_mels=32,
                   sample_rate=44100, hop_length=None, length=None):
    """T.MelSpectrogram(sample_rate, n_fft, win_length, power=1.0, f_min, f_max, n_mels) (the reference's
    _get_melspectrogram; hop defaults to win_length // 2, HTK mel scale, no filter normalisation).
    waveform [..., L] CUDA fp32 -> [..., n_mels, 1 + L // hop]."""
    if power != 1.0:
        raise NotImplementedError("only power=1.0 (magnitude) is on the hot path (BatvisionV2_Dataset.py:114)")
    hop = int(hop_length) if hop_length else win_length // 2
    flat, lead = _rows(waveform)
    rows, full = flat.shape
    L = full if length is None else min(int(length), full)
    pitch = flat.stride(0) if rows > 1 else full
    T = 1 + L // hop
    lib = _lib.load()
    ws = torch.empty(lib.adp_feature_workspace_bytes(rows, L, n_fft, hop), device=flat.device, dtype=torch.uint8)
    mel = torch.empty((rows, n_mels, T), device=flat.device, dtype=torch.float32)
    with _lib.on_device(flat):
        _lib.check(lib.adp_mel_spectrogram(flat.data_ptr(), rows, L, pitch, n_fft, win_length, hop, n_mels,
                                           float(sample_rate), float(f_min), float(f_max), mel.data_ptr(),
                                           ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    return mel.reshape(*lead, n_mels, T)


def resize(spec, size):
    """transforms.Resize((size,size)) on [..., H, W] CUDA fp32 (utils_dataset.py:18-20)."""
    _lib.require_cuda(spec, "spectrogram", torch.float32)
    spec = spec.contiguous()
    H, W = spec.shape[-2:]
    lead = spec.shape[:-2]
    rows = int(spec.numel() // (H * W))
    out = torch.empty((rows, size, size), device=spec.device, dtype=torch.float32)
    with _lib.on_device(spec):
        _lib.check(_lib.load().adp_resize_aa(spec.data_ptr(), rows, H, W, size, out.data_ptr(), _lib.stream_ptr()))
    return out.reshape(*lead, size, size)


class SpectrogramTransform:
    """waveform [B,C,L_raw] (CUDA fp32) -> feature [B,C,S,S] in one fused call.

    log_minmax=True reproduces BatVision V2 ('spectrogram' audio_format): cut to the echo
    window, STFT(512,64,16), log(x+1e-8), per-channel min-max, Resize.  log_minmax=False is
    BatVision V1: STFT, Resize.  mel=dict(sample_rate, f_min, f_max, n_mels) inserts the mel
    filterbank after the magnitude ('mel_spectrogram' audio_format; hop is then win // 2).
    """

    def __init__(self, images_size=256, max_depth=30.0, log_minmax=True, cut=True, sample_rate=44100,
                 stft=None, mel=None):
        self.size = int(images_size)
        self.log_minmax = bool(log_minmax)
        self.n_fft, self.win, self.hop = stft if stft is not None else stft_params(max_depth)
        self.mel = dict(MEL_DEFAULTS, **mel) if mel is not None else None
        if self.mel is not None and stft is None:
            self.hop = self.win // 2               # T.MelSpectrogram default; the reference passes no hop (:114)
        self.cut = cut_length(max_depth, sample_rate) if (cut and max_depth) else None
        self._ws = None

    @classmethod
    def for_cfg(cls, cfg):
        name = str(getattr(cfg.dataset, "name", "batvisionv2")).lower()
        fmt = str(getattr(cfg.dataset, "audio_format", "spectrogram"))
        if "v1" in name:
            return cls(cfg.dataset.images_size, cfg.dataset.max_depth, log_minmax=False, cut=False,
                       stft=(512, 64, 16))
        return cls(cfg.dataset.images_size, cfg.dataset.max_depth, log_minmax=True, cut=True,
                   mel={} if "mel" in fmt else None)

    def __call__(self, waveform):
        flat, lead = _rows(waveform)
        rows, full = flat.shape
        L = full if self.cut is None else min(self.cut, full)
        pitch = flat.stride(0) if rows > 1 else full
        lib = _lib.load()
        if self.mel is not None:
            need = lib.adp_feature_mel_workspace_bytes(rows, L, self.n_fft, self.hop, self.mel["n_mels"])
        else:
            need = lib.adp_feature_workspace_bytes(rows, L, self.n_fft, self.hop)
        if self._ws is None or self._ws.numel() < need or self._ws.device != flat.device:
            self._ws = torch.empty(need, device=flat.device, dtype=torch.uint8)
        out = torch.empty((rows, self.size, self.size), device=flat.device, dtype=torch.float32)
        with _lib.on_device(flat):
            if self.mel is not None:
                m = self.mel
                _lib.check(lib.adp_feature_forward_mel(flat.data_ptr(), rows, L, pitch, self.n_fft, self.win, self.hop,
                                                       int(m["n_mels"]), float(m["sample_rate"]), float(m["f_min"]),
                                                       float(m["f_max"]), 1 if self.log_minmax else 0, self.size,
                                                       out.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                                       _lib.stream_ptr()))
            else:
                _lib.check(lib.adp_feature_forward(flat.data_ptr(), rows, L, pitch, self.n_fft, self.win, self.hop,
                                                   1 if self.log_minmax else 0, self.size, out.data_ptr(),
                                                   self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr()))
        return out.reshape(*lead, self.size, self.size)


class DepthTransform:
    """Raw depth maps [B,H,W] in millimetres (CUDA fp32, or uint16 viewed as int16/uint16) -> gt_depth [B,1,S,S]:
    the depth half of __getitem__ (BatvisionV2_Dataset.py:68-78; BatvisionV1_Dataset.py:47-65) for a whole batch,
    bit-exact with the reference's numpy + cv2.INTER_NEAREST code."""

    def __init__(self, images_size=256, max_depth=30.0, depth_norm=False, nan_to_num=False):
        self.size = int(images_size)
        self.max_depth = float(max_depth) if max_depth else 0.0
        self.norm_div = self.max_depth if depth_norm else 0.0
        self.nan_to_num = bool(nan_to_num)

    @classmethod
    def for_cfg(cls, cfg):
        v1 = "v1" in str(getattr(cfg.dataset, "name", "batvisionv2")).lower()
        # V2 never normalises the depth in the dataset (its __getitem__ has no depth_norm branch)
        return cls(cfg.dataset.images_size, cfg.dataset.max_depth, depth_norm=v1 and bool(cfg.dataset.depth_norm),
                   nan_to_num=v1)

    def __call__(self, raw):
        if not raw.is_cuda:
            raise RuntimeError("DepthTransform needs a CUDA tensor (no CPU fallback)")
        if raw.dtype == torch.float32:
            code = 0
        elif raw.dtype in (torch.uint16, torch.int16):
            code = 1
        else:
            raise TypeError("raw depth must be float32 or uint16, got %s" % raw.dtype)
        raw = raw.contiguous()
        H, W = raw.shape[-2:]
        rows = raw.numel() // (H * W)
        out = torch.empty((rows, 1, self.size, self.size), device=raw.device, dtype=torch.float32)
        with _lib.on_device(raw):
            _lib.check(_lib.load().adp_depth_prepare(raw.data_ptr(), code, rows, H, W, self.size, self.max_depth,
                                                     1 if self.nan_to_num else 0, self.norm_div, out.data_ptr(),
                                                     _lib.stream_ptr()))
        return out
