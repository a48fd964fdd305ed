#!/usr/bin/env python
"""Recipe for `oracle/_ref/`: the reference's own model sources, made importable as the CPU arm of the benchmark and as a second
checker -- TEST / MEASUREMENT INFRASTRUCTURE, never imported by the product package.

The reference is plain Python with no build step: "building" it means placing `src/model.py` and `src/optimized_model.py`
(the two files of the hot path) where the GPU box can import them.  The copies live ONLY under
`oracle/_ref/` (git-ignored, so they never enter this repository's history; not gpurun-ignored, so they travel to the GPU box
with the snapshot).  Run here, in the build container, where `/root/reference` exists:

    python oracle/build_ref.py          # also called by __graft_entry__.build()

On the GPU box `/root/reference` does not exist: `bench.py --impl reference` and the smoke test use `oracle/_ref/` when it is
present (`kind: "reference"`) and the oracle port `oracle/torch_unet.py` otherwise (`kind: "port"`).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DG_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["src/model.py", "src/optimized_model.py"]


def build():
    if not os.path.isdir(REF):
        return None
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(DST, os.path.basename(f)))
    with open(os.path.join(DST, "README"), "w") as fh:
        fh.write("Unmodified copies of the reference's src/model.py and src/optimized_model.py, placed here by oracle/build_ref.py.\n"
                 "Git-ignored on purpose: measurement infrastructure, not part of this repository.\n")
    return DST


def load(name="model"):
    """Import oracle/_ref/<name>.py by path (as scripts/export_to_onnx.py:26-37 loads model classes); None if absent."""
    path = os.path.join(DST, name + ".py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("dg_reference_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    out = build()
    print("oracle/_ref:", out if out else f"skipped ({REF} not present)")
    sys.exit(0)
