"""Recipe for oracle/_ref/: the UNMODIFIED reference implementation of the path (pure Python over torch), staged from
/root/reference so that it travels to the GPU box with the snapshot (oracle/_ref/ is git-ignored: no reference source
enters the history).  TEST / BASELINE INFRASTRUCTURE ONLY: `bench.py --impl reference` imports it to time the
reference's own CPU path (`cpu_baseline.kind = "reference"`); nothing under neural_marionette_b200/ may.

    python oracle/make_ref.py          # BUILD container only (needs /root/reference); __graft_entry__.build() calls it
"""
from __future__ import annotations

import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
# the files the path imports: model/, modules/, the four utils modules, the shipped hyper-parameters
FILES = ["model/neural_marionette.py", "model/kypt_detector.py", "model/hsvrnn_bvh.py", "modules/vox_modules.py",
         "utils/dataset_utils.py", "utils/kypt_detector_utils.py", "utils/geo_utils.py", "utils/dyna_utils.py",
         "pretrained/aist/opt.pickle"]


def stage() -> bool:
    if not os.path.isdir(REF):
        return os.path.isdir(DST)
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    for pkg in ("model", "modules", "utils"):
        src_init = os.path.join(REF, pkg, "__init__.py")
        dst_init = os.path.join(DST, pkg, "__init__.py")
        if os.path.exists(src_init):
            shutil.copyfile(src_init, dst_init)
    return True


if __name__ == "__main__":
    print("staged" if stage() else "no /root/reference here", DST)
