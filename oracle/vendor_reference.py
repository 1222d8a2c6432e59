"""Copy the reference's own modules for the hot path into baseline/_ref/ (git-ignored, shipped to the GPU box by
gpurun) so that `bench.py --impl reference` and the `cpu_baseline` leg can time the UNMODIFIED reference code
 instead of the oracle port.  Test / measurement infrastructure only: nothing under
audio_depth_estimation_b200/ imports it, and no reference source enters the git history.

    python oracle/vendor_reference.py            (run by __graft_entry__.build() when /root/reference exists)
"""
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(REPO, "baseline", "_ref")
FILES = ["config_loader.py", "utils_loss.py", "utils_criterion.py", "models/unetbaseline_model.py",
         "models/binaural_attention_model.py", "dataloader/BatvisionV1_Dataset.py", "dataloader/BatvisionV2_Dataset.py",
         "dataloader/utils_dataset.py"]


def vendor(verbose=True):
    if not os.path.isdir(SRC):
        return False
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
    for pkg in ("models", "dataloader"):
        init = os.path.join(SRC, pkg, "__init__.py")
        dst = os.path.join(DST, pkg, "__init__.py")
        if os.path.exists(init):
            shutil.copyfile(init, dst)
        elif not os.path.exists(dst):
            open(dst, "w").close()
    if os.path.isdir(os.path.join(SRC, "conf")):
        shutil.copytree(os.path.join(SRC, "conf"), os.path.join(DST, "conf"), dirs_exist_ok=True)
    if verbose:
        print("reference modules vendored into", DST)
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
