"""Checkpoint files in the reference's format (train.py:1003-1017 writes, train.py:600-606 and test.py:150-205 read):

    ./checkpoints/<experiment_name>/checkpoint_<epoch>.pth = {'epoch', 'state_dict', 'optimizer'}

`state_dict` keys carry the `module.` prefix when the reference trained under nn.DataParallel; UnetGenerator's
load_state_dict accepts both, and `save_checkpoint(..., data_parallel_keys=True)` writes the prefixed form so that the
reference's own test.py (which wraps the model in DataParallel when gpu_ids is non-empty) can read the file.  Unlike
the reference, `load_checkpoint` can also restore the optimizer moments (the reference saves but never reloads them).
"""
import os

import torch


def checkpoint_path(experiment_name, epoch, root="./checkpoints"):
    return os.path.join(root, experiment_name, "checkpoint_%s.pth" % (epoch,))


def save_checkpoint(path, epoch, model, optimizer=None, data_parallel_keys=False):
    sd = {k: v.detach().cpu().contiguous() for k, v in model.state_dict().items()}
    if data_parallel_keys:
        sd = {"module." + k: v for k, v in sd.items()}
    state = {"epoch": int(epoch), "state_dict": sd}
    if optimizer is not None:
        osd = optimizer.state_dict()
        for ent in osd.get("state", {}).values():
            for k in ("exp_avg", "exp_avg_sq"):
                ent[k] = ent[k].cpu()
        state["optimizer"] = osd
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(state, path)
    return path


def load_checkpoint(path, model, optimizer=None, map_location="cpu"):
    """Returns the epoch to resume FROM (saved epoch + 1, train.py:606)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(ckpt["state_dict"])
    if optimizer is not None and ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(ckpt["optimizer"])
    return int(ckpt["epoch"]) + 1
