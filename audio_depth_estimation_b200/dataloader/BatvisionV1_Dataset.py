"""Mirror of dataloader/BatvisionV1_Dataset.py (reference :13-95) for the audio hot path.

Single annotation csv (:22), depth clean-up / clip / NEAREST resize / optional max_depth
normalisation (:47-66), left+right .npy waveforms stacked [2,L] (:70-72), Spectrogram(512, 64,
hop 16) with no log and no min-max (:75-78), Resize.  The spectrogram runs on the GPU.
"""
import os

import numpy as np
import pandas as pd
import torch
from torch.utils.data import Dataset

from .. import feature
from ._common import nearest_resize
from .utils_dataset import get_transform


class BatvisionV1Dataset(Dataset):
    def __init__(self, cfg, annotation_file, location_blacklist=None):
        self.cfg = cfg
        self.root_dir = cfg.dataset.dataset_dir
        self.audio_format = cfg.dataset.audio_format
        self.device = torch.device("cuda") if torch.cuda.is_available() else None
        self.instances = pd.read_csv(os.path.join(self.root_dir, annotation_file))
        if location_blacklist:      # reference :25-31 filters on the left-ear audio path
            for location in location_blacklist:
                self.instances = self.instances[~self.instances["audio path left"].str.contains(location)]

    def __len__(self):
        return len(self.instances)

    def __getitem__(self, idx):
        inst = self.instances.iloc[idx]
        d = np.load(os.path.join(self.root_dir, inst["depth path"])).astype(np.float32)
        d = np.nan_to_num(d)          # reference :49-52: nan -> 0, +-inf -> +-float max (then clipped below)
        d[d == -np.inf] = 0
        d[d == np.inf] = 0
        d = d / 1000
        if self.cfg.dataset.max_depth:
            d[d > self.cfg.dataset.max_depth] = self.cfg.dataset.max_depth
        d[d < 0] = 0
        d = nearest_resize(d, self.cfg.dataset.images_size)
        if self.cfg.dataset.depth_norm and self.cfg.dataset.max_depth:
            d = d / np.float32(self.cfg.dataset.max_depth)
        gt_depth = torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32)).unsqueeze(0)

        left = np.load(os.path.join(self.root_dir, inst["audio path left"])).astype(np.float32)
        right = np.load(os.path.join(self.root_dir, inst["audio path right"])).astype(np.float32)
        waveform = torch.from_numpy(np.stack((left, right)))
        if "spectrogram" in self.audio_format:
            if self.device is None:
                raise RuntimeError("BatvisionV1Dataset needs a CUDA device for the spectrogram transform "
                                   "(no CPU fallback); use audio_format='waveform' in CPU worker processes")
            spec = self._get_spectrogram(waveform.to(self.device), n_fft=512, power=1.0, win_length=64,
                                         hop_length=64 // 4)
            if "resize" in str(self.cfg.dataset.preprocess):
                spec = get_transform(self.cfg, convert=False)(spec)
            return spec, gt_depth
        return waveform, gt_depth

    def _get_spectrogram(self, waveform, n_fft=400, power=1.0, win_length=400, hop_length=100):
        return feature.spectrogram(waveform, n_fft=n_fft, power=power, win_length=win_length, hop_length=hop_length)
