"""Mirror of config_loader.py (reference :43-97): load_config(dataset_name, mode, experiment_name,
model_name) -> SimpleNamespace(dataset, mode, model) read from conf/{dataset,mode,model}/*.yaml with
the reference's key names.  Uses PyYAML when present, otherwise a small
`key: value` reader that understands the same scalar forms the reference's fallback does."""
import os
from types import SimpleNamespace

CONF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "conf")


def _scalar(text):
    text = text.split("#", 1)[0].strip()
    if text == "" or text.lower() in ("null", "none", "~"):
        return None
    if text.lower() in ("true", "false"):
        return text.lower() == "true"
    if len(text) >= 2 and text[0] == text[-1] and text[0] in "\"'":
        return text[1:-1]
    for cast in (int, float):
        try:
            return cast(text)
        except ValueError:
            pass
    return text


def _read_flat_yaml(path):
    out = {}
    with open(path, "r") as f:
        for raw in f:
            line = raw.strip()
            if not line or line.startswith("#") or ":" not in line:
                continue
            key, value = line.split(":", 1)
            out[key.strip()] = _scalar(value)
    return out


def _read(path):
    if not os.path.exists(path):
        raise FileNotFoundError("config file not found: %s" % path)
    try:
        import yaml
    except ImportError:
        return _read_flat_yaml(path)
    with open(path, "r") as f:
        return yaml.safe_load(f) or {}


def load_config(dataset_name="batvisionv2", mode="train", experiment_name="default", model_name="unet_baseline",
                conf_dir=None):
    conf_dir = conf_dir or CONF_DIR
    groups = {}
    for group, name in (("dataset", dataset_name), ("mode", mode), ("model", model_name)):
        path = os.path.join(conf_dir, group, name + ".yaml")
        if group == "model" and not os.path.exists(path):       # reference :73-76: unknown model -> unet_baseline.yaml
            path = os.path.join(conf_dir, "model", "unet_baseline.yaml")
        groups[group] = SimpleNamespace(**_read(path))
    groups["mode"].mode = mode                                  # reference :93
    groups["mode"].experiment_name = experiment_name
    return SimpleNamespace(**groups)
