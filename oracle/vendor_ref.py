"""Recipe that makes the reference's OWN hot-path modules available to the GPU box.

    python -m oracle.vendor_ref          (also run by __graft_entry__.build())

The reference (szbonaldo/FedMLP) is pure Python; there is nothing to compile.  The GPU box has no
/root/reference, so — exactly like a compiled oracle/_ref/*.so would — the UNMODIFIED files are copied
from where they lie under /root/reference into oracle/_ref/ (git-ignored, not gpurun-ignored: they travel
with the snapshot but never enter the repository history).  `bench.py --impl reference` and
tests/test_oracle_vs_reference.py import them through oracle/ref_loader.py; nothing under fedmlp_b200/
ever does.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

SRC = Path("/root/reference")
DST = Path(__file__).resolve().parent / "_ref"
# utils/local_training.py imports evaluations / feature_visual / FedNoRo / utils; evaluations imports
# multilabel_metrixs.  Nothing else of the reference is on the path.
FILES = ["utils/FedAvg.py", "utils/FedNoRo.py", "utils/utils.py", "utils/local_training.py", "utils/evaluations.py",
         "utils/feature_visual.py", "utils/multilabel_metrixs.py"]


def vendor(verbose: bool = False) -> bool:
    """Copy the files if the reference is mounted; returns True when oracle/_ref is complete."""
    if (SRC / "utils" / "FedAvg.py").is_file():
        manifest = {}
        for rel in FILES:
            dst = DST / rel
            dst.parent.mkdir(parents=True, exist_ok=True)
            shutil.copyfile(SRC / rel, dst)
            manifest[rel] = hashlib.sha256(dst.read_bytes()).hexdigest()
        (DST / "MANIFEST.json").write_text(json.dumps({"source": str(SRC), "sha256": manifest}, indent=1))
        if verbose:
            print(f"[oracle] vendored {len(FILES)} unmodified reference files into {DST}", file=sys.stderr)
    return all((DST / rel).is_file() for rel in FILES)


if __name__ == "__main__":
    print("oracle/_ref complete:", vendor(verbose=True))
