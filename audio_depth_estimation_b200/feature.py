"""Batched GPU feature transform: binaural waveform -> [B,2,S,S] network input.

Host-side mirror of the audio branch of the reference datasets, running on the collated
batch through libadp_b200 (adp_stft_mag / adp_feature_forward / adp_resize_aa):
  dataloader/BatvisionV2_Dataset.py:92-137, :177-185   (cut, STFT params, log, min-max, Resize)
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
    _lib.check(lib.adp_stft_mag(flat.data_ptr(), rows, L, pitch, n_fft, win_length, hop_length,
                                spec.data_ptr(), _lib.stream_ptr()))
    return spec.reshape(*lead, n_fft // 2 + 1, T)


def resize(spec, size):
    """transforms.Resize((size,size)) on [..., H, W] CUDA fp32 (utils_dataset.py:18-20)."""
    _lib.require_cuda(spec, "spectrogram", torch.float32)
    spec = spec.contiguous()
    H, W = spec.shape[-2:]
    lead = spec.shape[:-2]
    rows = int(spec.numel() // (H * W))
    out = torch.empty((rows, size, size), device=spec.device, dtype=torch.float32)
    _lib.check(_lib.load().adp_resize_aa(spec.data_ptr(), rows, H, W, size, out.data_ptr(), _lib.stream_ptr()))
    return out.reshape(*lead, size, size)


class SpectrogramTransform:
    """waveform [B,C,L_raw] (CUDA fp32) -> feature [B,C,S,S] in one fused call.

    log_minmax=True reproduces BatVision V2 ('spectrogram' audio_format): cut to the echo
    window, STFT(512,64,16), log(x+1e-8), per-channel min-max, Resize.  log_minmax=False is
    BatVision V1: STFT, Resize.
    """

    def __init__(self, images_size=256, max_depth=30.0, log_minmax=True, cut=True, sample_rate=44100,
                 stft=None):
        self.size = int(images_size)
        self.log_minmax = bool(log_minmax)
        self.n_fft, self.win, self.hop = stft if stft is not None else stft_params(max_depth)
        self.cut = cut_length(max_depth, sample_rate) if (cut and max_depth) else None
        self._ws = None

    @classmethod
    def for_cfg(cls, cfg):
        name = str(getattr(cfg.dataset, "name", "batvisionv2")).lower()
        if "v1" in name:
            return cls(cfg.dataset.images_size, cfg.dataset.max_depth, log_minmax=False, cut=False,
                       stft=(512, 64, 16))
        return cls(cfg.dataset.images_size, cfg.dataset.max_depth, log_minmax=True, cut=True)

    def __call__(self, waveform):
        flat, lead = _rows(waveform)
        rows, full = flat.shape
        L = full if self.cut is None else min(self.cut, full)
        pitch = flat.stride(0) if rows > 1 else full
        lib = _lib.load()
        need = lib.adp_feature_workspace_bytes(rows, L, self.n_fft, self.hop)
        if self._ws is None or self._ws.numel() < need or self._ws.device != flat.device:
            self._ws = torch.empty(need, device=flat.device, dtype=torch.uint8)
        out = torch.empty((rows, self.size, self.size), device=flat.device, dtype=torch.float32)
        _lib.check(lib.adp_feature_forward(flat.data_ptr(), rows, L, pitch, self.n_fft, self.win, self.hop,
                                           1 if self.log_minmax else 0, self.size, out.data_ptr(),
                                           self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr()))
        return out.reshape(*lead, self.size, self.size)
