"""Seeded synthetic BatVision-shaped data.

Pure numpy so the same seed gives the same arrays on every machine; used by the
golden-vector generator, the parity tests and bench.py.
"""
import numpy as np

V2_LEN = 7782      # int(2*30/340*44100), BatvisionV2_Dataset.py:102-104
V1_LEN = 3200      # BatVision-V1 recordings (72.5 ms @ 44.1 kHz)


def waveform(batch, length=V2_LEN, seed=1234, echo=False):
    """U(-1,1) float32 [B,2,L]; echo=True gives a chirp + decaying-noise variant
    with a wide log dynamic range."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1.0, 1.0, size=(batch, 2, length)).astype(np.float32)
    if echo:
        t = np.arange(length, dtype=np.float64) / 44100.0
        chirp = np.sin(2 * np.pi * (2000.0 + 4.0e6 * t) * t) * (t < 3e-3)
        env = np.exp(-t / 0.02)
        x = (0.7 * chirp[None, None, :] + 0.3 * x * env[None, None, :]).astype(np.float32)
    return x


def gt_depth(batch, size=256, max_depth=30.0, seed=4321, invalid_frac=0.15, normalised=False):
    """U(0,max_depth) float32 [B,1,S,S] with exact zeros where U(0,1) < invalid_frac."""
    rng = np.random.default_rng(seed)
    d = rng.uniform(0.0, max_depth, size=(batch, 1, size, size)).astype(np.float32)
    inv = rng.uniform(0.0, 1.0, size=d.shape) < invalid_frac
    d[inv] = 0.0
    if normalised:
        d = (d / np.float32(max_depth)).astype(np.float32)
    return d


def feature_like(batch, size=256, seed=99):
    """U(0,1) float32 [B,2,S,S] stand-in for the feature tensor."""
    rng = np.random.default_rng(seed)
    return rng.uniform(0.0, 1.0, size=(batch, 2, size, size)).astype(np.float32)
