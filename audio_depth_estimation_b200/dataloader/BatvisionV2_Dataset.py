"""Mirror of dataloader/BatvisionV2_Dataset.py (reference :12-197) for the audio hot path.

Same constructor, `instances` DataFrame, `__getitem__ -> (input[2,S,S], gt_depth[1,S,S])` and
`_get_spectrogram` signature.  The spectrogram / log / min-max / Resize chain (:96-135) runs on
the GPU through libadp_b200; depth files and audio are read on the host exactly as the reference
does (:65-78, :142-175).  `audio_format='waveform'` returns the cut waveform so that the batched
transform (feature.SpectrogramTransform.for_cfg(cfg)) can run once per collated batch.
"""
import os

import numpy as np
import pandas as pd
import torch
from torch.utils.data import Dataset

from .. import feature
from ._common import load_audio, nearest_resize
from .utils_dataset import get_transform  # noqa: F401  (re-exported like the reference module)


class BatvisionV2Dataset(Dataset):
    def __init__(self, cfg, annotation_file, location_blacklist=None, use_image=False):
        self.cfg = cfg
        self.root_dir = cfg.dataset.dataset_dir
        self.audio_format = cfg.dataset.audio_format
        self.use_image = use_image
        self.device = torch.device("cuda") if torch.cuda.is_available() else None
        # (os.listdir order, as the reference :22-26 -- the row order of .instances follows it)
        locations = [d for d in os.listdir(self.root_dir)
                     if os.path.isdir(os.path.join(self.root_dir, d))
                     and not d.startswith((".", "__")) and not d.endswith("_unzipped")]
        if location_blacklist:
            locations = [d for d in locations if d not in location_blacklist]
        frames = []
        for loc in locations:
            csv = os.path.join(self.root_dir, loc, annotation_file)
            if os.path.exists(csv):
                frames.append(pd.read_csv(csv))
            else:
                print("Warning: %s not found, skipping location %s" % (csv, loc))
        if not frames:
            raise ValueError("No valid locations found with %s in %s" % (annotation_file, self.root_dir))
        self.instances = pd.concat(frames)

    def __len__(self):
        return len(self.instances)

    def _depth(self, inst):
        d = np.load(os.path.join(self.root_dir, inst["depth path"], inst["depth file name"])).astype(np.float32)
        d = d / 1000.0
        if self.cfg.dataset.max_depth:
            d[d > self.cfg.dataset.max_depth] = self.cfg.dataset.max_depth
        d[d < 0] = 0
        d = nearest_resize(d, self.cfg.dataset.images_size)
        return torch.from_numpy(np.ascontiguousarray(d)).unsqueeze(0)

    def __getitem__(self, idx):
        inst = self.instances.iloc[idx]
        gt_depth = self._depth(inst)
        if self.use_image:      # camera branch (reference :87-90): the RGB frame instead of the echo, host-side only
            return self._load_image(os.path.join(self.root_dir, inst["camera path"], inst["camera file name"])), gt_depth
        waveform, sr = load_audio(os.path.join(self.root_dir, inst["audio path"], inst["audio file name"]))
        n_fft, win_length, hop_length = 400, 200, 100
        if self.cfg.dataset.max_depth:
            waveform = waveform[:, :feature.cut_length(self.cfg.dataset.max_depth, sr)]
            n_fft, win_length, hop_length = feature.stft_params(self.cfg.dataset.max_depth)
        if "spectrogram" in self.audio_format:
            if "resize" not in str(self.cfg.dataset.preprocess):
                raise NotImplementedError("the fused feature kernel always resizes (cfg.dataset.preprocess='resize')")
            # STFT + log + per-channel min-max + Resize (reference :117-135) as one fused library call
            # ('mel_spectrogram': the mel bank sits between magnitude and log, hop = win_length // 2, :111-114)
            mel = "mel" in self.audio_format
            key = (n_fft, win_length, hop_length, mel)
            if getattr(self, "_fused_key", None) != key:          # one transform (and workspace) for the whole dataset
                self._fused = feature.SpectrogramTransform(self.cfg.dataset.images_size, self.cfg.dataset.max_depth,
                                                           log_minmax=True, cut=False, mel={} if mel else None,
                                                           stft=(n_fft, win_length, win_length // 2 if mel else hop_length))
                self._fused_key = key
            return self._fused(self._to_device(waveform)), gt_depth
        if "waveform" in self.audio_format:
            return waveform, gt_depth
        raise ValueError("unknown audio_format %r" % (self.audio_format,))

    def _to_device(self, waveform):
        from ._common import in_worker_process
        if in_worker_process():
            raise RuntimeError("audio_format=%r computes the feature on the GPU inside __getitem__, which cannot run in a "
                               "DataLoader worker process (num_workers > 0): use audio_format='waveform' and apply "
                               "feature.SpectrogramTransform.for_cfg(cfg) to the collated batch (TrainStep does), or "
                               "num_workers=0" % (self.audio_format,))
        if self.device is None:
            raise RuntimeError("BatvisionV2Dataset needs a CUDA device for the spectrogram transform "
                               "(no CPU fallback); use audio_format='waveform' in CPU worker processes")
        return waveform.to(self.device, dtype=torch.float32)

    def _load_image(self, image_path):
        """Reference :199-210: cv2.imread -> RGB -> cv2.resize (bilinear) to images_size -> float32 / 255 -> [3, S, S]."""
        import cv2
        image = cv2.imread(image_path)
        if image is None:
            raise RuntimeError("Could not load image file %s" % image_path)
        image = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
        image = cv2.resize(image, (self.cfg.dataset.images_size, self.cfg.dataset.images_size))
        return torch.from_numpy(image.astype(np.float32) / 255.0).permute(2, 0, 1)

    def _get_spectrogram(self, waveform, n_fft=400, power=1.0, win_length=400, hop_length=100):
        return feature.spectrogram(waveform, n_fft=n_fft, power=power, win_length=win_length, hop_length=hop_length)

    def _get_melspectrogram(self, waveform, n_fft=400, power=1.0, win_length=400, f_min=20.0, f_max=20000.0):
        return feature.melspectrogram(waveform, n_fft=n_fft, power=power, win_length=win_length, f_min=f_min, f_max=f_max,
                                      n_mels=32, sample_rate=44100)
