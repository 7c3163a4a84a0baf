"""Recipe that installs the UNMODIFIED reference model as ``oracle/_ref`` (test/bench infrastructure).

    python oracle/build_ref.py            # in the build container, where /root/reference exists

The reference hot path is one pure-Python file, ``models/FastEGNN.py``; its only third-party import is
``torch_geometric.nn.global_mean_pool`` (PyG is not in the image), which ``oracle/ref_loader.py`` stands in
for with a scatter-mean stub.  This script copies that one file — byte for byte, sha256
recorded — from where it lies under ``/root/reference`` into ``oracle/_ref/models/`` so that the CPU arm of
``bench.py`` (``cpu_baseline.kind == "reference"``, ``--impl reference``) and the tests can time/evaluate the
reference itself instead of the restatement in ``fastegnn_oracle.py``.

``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored, so the installed copy travels to the GPU box like the built ``.so``.  On the GPU box
``/root/reference`` does not exist: there this script is a no-op and the prebuilt copy (if any) is used.

Only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU arm may import anything under ``oracle/``.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ["models/FastEGNN.py"]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose: bool = False) -> bool:
    """Install (or refresh) oracle/_ref from /root/reference.  Returns True if oracle/_ref is usable."""
    manifest_path = os.path.join(DST, "MANIFEST.json")
    if not os.path.isdir(REF_ROOT):
        return os.path.exists(manifest_path)
    os.makedirs(os.path.join(DST, "models"), exist_ok=True)
    manifest = {"source": REF_ROOT, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(DST, rel)
        if not os.path.exists(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest["files"][rel] = _sha(dst)
    open(os.path.join(DST, "models", "__init__.py"), "a").close()
    with open(manifest_path, "w") as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print(json.dumps(manifest, indent=1))
    return True


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref", "ready" if ok else "NOT available (no /root/reference and no prebuilt copy)")
