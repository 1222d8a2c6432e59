"""Mirror of dataloader/utils_dataset.py:10-49: get_transform(cfg, convert, depth_norm), ToTensor and MinMaxNorm.

The returned callable applies Resize((S,S)) with torchvision's tensor semantics (antialiased
bilinear) through adp_resize_aa; it expects CUDA fp32 tensors [..., H, W].  depth_norm=True appends the
reference's MinMaxNorm(0, max_depth) (:23-27, :31-49; used by the sibling sparse-depth dataset, not by the V1 / V2 audio
path, whose depth normalisation lives in __getitem__).
"""
import torch

from .. import feature


class MinMaxNorm(torch.nn.Module):
    """(x - min) / (max - min); `min` / `max` are floats, or 2-tuples applied per channel of a [2, ...] tensor
    (reference :31-49, same assertion and the same channel rule)."""

    def __init__(self, min, max):
        super().__init__()
        assert isinstance(min, (float, tuple)) and isinstance(max, (float, tuple))
        self.min = torch.tensor(min)
        self.max = torch.tensor(max)

    def forward(self, tensor):
        lo, hi = self.min.to(tensor.device), self.max.to(tensor.device)
        if tensor.shape[0] == 2:
            return torch.stack([(tensor[c] - lo[c]) / (hi[c] - lo[c]) for c in range(2)], dim=0)
        return (tensor - lo) / (hi - lo)


class ToTensor:
    """torchvision.transforms.ToTensor for what the reference feeds it (:14-16): a PIL image or an HWC / HW numpy array
    becomes a CHW float32 tensor, uint8 input scaled by 1/255; other dtypes keep their values.  The tensor is moved to the
    current CUDA device when there is one, because the Resize that follows runs there."""

    def __call__(self, pic):
        import numpy as np
        arr = np.asarray(pic)
        if arr.ndim == 2:
            arr = arr[:, :, None]
        if arr.ndim != 3:
            raise ValueError("ToTensor expects an image with 2 or 3 dimensions, got shape %s" % (arr.shape,))
        t = torch.from_numpy(np.ascontiguousarray(arr.transpose(2, 0, 1)))
        t = t.to(torch.float32).div(255) if arr.dtype == np.uint8 else t
        return t.cuda() if torch.cuda.is_available() else t


class Resize:
    def __init__(self, size):
        self.size = size if isinstance(size, int) else size[0]

    def __call__(self, x):
        return feature.resize(x, self.size)


class Compose:
    def __init__(self, transforms):
        self.transforms = list(transforms)

    def __call__(self, x):
        for t in self.transforms:
            x = t(x)
        return x


def get_transform(cfg, convert=False, depth_norm=False):
    steps = []
    if convert:
        steps.append(ToTensor())
    if "resize" in str(cfg.dataset.preprocess):
        steps.append(Resize((cfg.dataset.images_size, cfg.dataset.images_size)))
    if depth_norm:
        steps.append(MinMaxNorm(min=0.0, max=cfg.dataset.max_depth))
    return Compose(steps)
