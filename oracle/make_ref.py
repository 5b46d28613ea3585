"""Recipe for oracle/_ref: a verbatim, git-ignored copy of the reference files the hot path lives in.

Test / measurement infrastructure.  The reference is pure Python, so "building" it means making the few modules of
the path importable where bench.py runs: the GPU box has no /root/reference, but oracle/_ref travels there with the
snapshot (it is listed in .gitignore, so no reference source ever enters the history, and NOT in .gpurunignore).

    python oracle/make_ref.py          # also run by __graft_entry__.build() when /root/reference is present

Copies, unmodified:  src/model/{squeezedet,modules}.py, src/engine/detector.py, src/utils/{boxes,image,misc}.py
Used only by bench.py --impl reference / its cpu_baseline leg (kind "reference") and never by the product package."""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SQD_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref")
FILES = ["src/model/squeezedet.py", "src/model/modules.py", "src/engine/detector.py", "src/utils/boxes.py",
         "src/utils/image.py", "src/utils/misc.py"]


def make(verbose=True):
    if not os.path.isdir(os.path.join(REF, "src")):
        if verbose:
            print(f"{REF} not present: oracle/_ref left as it is")
        return None
    manifest = {}
    for rel in FILES:
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": "hazenai/SqueezeDet-PyTorch @ " + REF, "sha256": manifest}, f, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference files copied")
    return OUT


def load():
    """Import the reference's modules from oracle/_ref (None when the copy is absent).  Returns a namespace with
    `model` (model.squeezedet), `detector` (engine.detector), `boxes` (utils.boxes)."""
    src = os.path.join(OUT, "src")
    if not os.path.isfile(os.path.join(src, "model", "squeezedet.py")):
        return None
    import importlib
    import types
    if src not in sys.path:
        sys.path.insert(0, src)
    for name in ("model", "engine", "utils"):       # the reference has no __init__.py: namespace packages, ours first
        mod = sys.modules.get(name)
        if mod is not None and not any(str(p).startswith(src) for p in getattr(mod, "__path__", [])):
            del sys.modules[name]
    ns = types.SimpleNamespace()
    ns.model = importlib.import_module("model.squeezedet")
    ns.detector = importlib.import_module("engine.detector")
    ns.boxes = importlib.import_module("utils.boxes")
    return ns


if __name__ == "__main__":
    make()
