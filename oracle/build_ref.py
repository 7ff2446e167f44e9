"""Recipe that stages the UNMODIFIED reference implementation of the hot path under oracle/_ref/ so that the
reference itself — not a restatement — can be timed as the CPU baseline on the GPU box (`bench.py --impl
reference`, cpu_baseline.kind == "reference") and used to pin oracle/step.py (tests/test_ref_step.py).

The reference is pure Python (no native code: nothing to compile); "building" it = copying the nine hot-path
module files byte for byte from /root/reference into oracle/_ref/ (git-ignored, so they never enter this
repository's history; NOT gpurun-ignored, so they travel to the GPU box like the built .so).  Runs in the build
container only (`__graft_entry__.build()` calls it when /root/reference is present).  Test / measurement
infrastructure, never the product."""
from __future__ import annotations

import hashlib
import os
import shutil
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("GRAPHECHO_REFERENCE", "/root/reference"))
OUT = HERE / "_ref"

# the files the hot path lives in
FILES = ("models/fpnseg.py", "models/graph_matching.py", "models/affinity_layer.py", "models/transformer.py",
         "models/TGCN.py", "models/vig.py", "models/gradient_reversal.py", "utils/sinkhorn_distance.py",
         "utils/losses.py")


def available() -> bool:
    return all((OUT / f).exists() for f in FILES)


def build(verbose: bool = False) -> bool:
    """Copy the reference files (if the reference tree is mounted).  Returns True when oracle/_ref is complete."""
    if not REF.exists():
        return available()
    manifest = []
    for f in FILES:
        src, dst = REF / f, OUT / f
        dst.parent.mkdir(parents=True, exist_ok=True)
        if not dst.exists() or dst.read_bytes() != src.read_bytes():
            shutil.copyfile(src, dst)
        manifest.append(f"{hashlib.sha256(dst.read_bytes()).hexdigest()}  {f}")
    (OUT / "MANIFEST.sha256").write_text("\n".join(manifest) + "\n")
    if verbose:
        print(f"staged {len(FILES)} reference files under {OUT}")
    return available()


if __name__ == "__main__":
    print("oracle/_ref complete:", build(verbose=True))
