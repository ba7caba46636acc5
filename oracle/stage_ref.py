"""Stage the reference's own hot-path modules for the GPU box (build container only).

`bench.py --impl reference` should time the REFERENCE's PyTorch CPU path, not our port of it.  The reference is
pure Python and /root/reference does not exist on the GPU box, so `__graft_entry__.build()` calls `stage()`
here: it copies the eleven reference files the path imports (found by importing the reference through
oracle/ref_loader.py and listing `sys.modules`) plus its vocab files into `baseline/_ref/MolNexTR/`.  That
directory is git-ignored -- reference sources never enter this repository's history -- but it is not
gpurun-ignored, so it travels with the snapshot like the built .so.  The third-party classes the reference
imports (OpenNMT-py 2.2.0, timm 0.4.12) are not installed anywhere and stay restated in oracle/ref_shims.
"""
from __future__ import annotations

import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/MolNexTR"
DST = os.path.join(ROOT, "baseline", "_ref", "MolNexTR")
FILES = ["components.py", "tokenization.py", "utils.py", "decoding/__init__.py", "decoding/beam_search.py",
         "decoding/decode_strategy.py", "decoding/greedy_search.py", "models/__init__.py", "models/decoder.py",
         "models/embedding.py", "models/transformers.py", "vocab/vocab_chars.json", "vocab/vocab_uspto.json"]


def stage() -> bool:
    """Copy the files if the reference is present; returns True when baseline/_ref is usable afterwards."""
    if os.path.isdir(SRC):
        for rel in FILES:
            src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
            if not os.path.exists(src):
                continue
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst) or os.path.getsize(src) != os.path.getsize(dst):
                shutil.copyfile(src, dst)
    return os.path.exists(os.path.join(DST, "components.py"))


if __name__ == "__main__":
    print("staged" if stage() else "reference not available", DST)
