"""Recipe: stage the UNMODIFIED reference modules of the hot path under the git-ignored ``oracle/_ref/``.

    python -m oracle.build_ref          # run in the build container (needs /root/reference)

The reference is pure Python, so "building" it is a byte-for-byte copy of the files the path imports
(``modules/rrt.py`` -> ``rmsa.py``, ``emb_position.py``, ``datten.py``, ``nystrom_attention.py``) plus the MIL
hosts the drop-in GPU test runs (``attmil.py``, ``mean_max.py``, ``dsmil.py``).  Nothing is written anywhere
else: ``oracle/_ref/`` is listed in ``.gitignore`` (the sources never enter this repository's history) but not
in ``.gpurunignore``, so the staged tree travels to the GPU box with the snapshot, exactly like the built
``librrt_b200.so``.  There ``bench.py --impl reference`` times the real ``modules.rrt.RRTEncoder`` (CPU, all host
threads; plus eager fp32 and fp16-autocast on the GPU as informational extras) and ``tests/test_gpu_dropin.py``
runs the reference's own hosts with this repository's encoder inside.

TEST / BASELINE INFRASTRUCTURE ONLY -- the product (``rrt_mil_b200/``) never imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("RRT_REFERENCE_SOURCE", "/root/reference")
FILES = ["modules/__init__.py", "modules/rrt.py", "modules/rmsa.py", "modules/emb_position.py",
         "modules/datten.py", "modules/nystrom_attention.py", "modules/attmil.py", "modules/mean_max.py",
         "modules/dsmil.py"]


def staged() -> bool:
    return os.path.isfile(os.path.join(DEST, "modules", "rrt.py"))


def build(verbose: bool = False) -> bool:
    """Copy the files; returns False (and leaves any earlier staging alone) when the reference tree is absent."""
    if not os.path.isfile(os.path.join(SOURCE, "modules", "rrt.py")):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    commit = None
    sub = os.path.join(SOURCE, ".SUBMODULES.json")
    if os.path.isfile(sub):
        try:
            commit = json.load(open(sub)).get("commit")
        except Exception:
            commit = None
    json.dump({"source": SOURCE, "commit": commit, "sha256": manifest},
              open(os.path.join(DEST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"staged {len(FILES)} reference files under {DEST}")
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    if not ok:
        print(f"no reference tree under {SOURCE}; nothing staged", file=sys.stderr)
        sys.exit(0 if staged() else 1)
