"""Mirror of dataloader/utils_dataset.py:10-28: get_transform(cfg, convert, depth_norm).

The returned callable applies Resize((S,S)) with torchvision's tensor semantics (antialiased
bilinear) through adp_resize_aa; it expects CUDA fp32 tensors [..., H, W].
"""
from .. import feature


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
        raise NotImplementedError("convert=True (ToTensor on PIL images) is the image branch, outside the audio hot path")
    if "resize" in str(cfg.dataset.preprocess):
        steps.append(Resize((cfg.dataset.images_size, cfg.dataset.images_size)))
    return Compose(steps)
