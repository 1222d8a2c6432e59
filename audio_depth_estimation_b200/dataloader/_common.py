"""Host-side pieces shared by the two BatVision dataset mirrors (file I/O stays on the host)."""
import os
import wave as _wave

import numpy as np
import torch


def nearest_resize(depth, size):
    """cv2.resize(..., interpolation=cv2.INTER_NEAREST) index rule: src = min(floor(dst*src/dst_n), src_n-1)."""
    h, w = depth.shape
    iy = np.minimum(np.floor(np.arange(size) * (h / size)).astype(np.int64), h - 1)
    ix = np.minimum(np.floor(np.arange(size) * (w / size)).astype(np.int64), w - 1)
    return depth[iy][:, ix]


def in_worker_process():
    """True inside a torch DataLoader worker (a forked process must not touch CUDA)."""
    from torch.utils.data import get_worker_info
    return get_worker_info() is not None


def load_audio(path):
    """Returns (waveform [C,L] float32 tensor, sample_rate).  .npy and 8/16/32-bit PCM .wav are read directly, anything
    else goes through torchaudio.load as in the reference (BatvisionV2_Dataset.py:92-95)."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        arr = np.load(path).astype(np.float32)
        if arr.ndim == 1:
            arr = arr[None]
        return torch.from_numpy(arr), 44100
    if ext == ".wav":
        try:
            with _wave.open(path, "rb") as f:
                sr, ch, width, n = f.getframerate(), f.getnchannels(), f.getsampwidth(), f.getnframes()
                raw = f.readframes(n)
        except _wave.Error:                       # float / extensible wav files: what the reference uses throughout
            import torchaudio
            return torchaudio.load(path)
        if width == 3:
            import torchaudio
            return torchaudio.load(path)
        if width == 2:
            data = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
        elif width == 4:
            data = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
        elif width == 1:
            data = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        else:
            raise ValueError("unsupported wav sample width %d in %s" % (width, path))
        return torch.from_numpy(data.reshape(-1, ch).T.copy()), sr
    import torchaudio
    wav, sr = torchaudio.load(path)
    return wav, sr
