"""Build recipe for libfedmlp_b200.so (nvcc, sm_100a only, in-tree).

    python -m fedmlp_b200._build [--force]

The shared library is written to fedmlp_b200/lib/libfedmlp_b200.so so that it travels with the
source tree (it is git-ignored, not gpurun-ignored).  nvcc cross-compiles without a GPU.
No --use_fast_math: the kernels reproduce the reference's fp32 arithmetic (IEEE divide, expf).
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
INCLUDE = ROOT / "include"
LIB_DIR = PKG / "lib"
LIB_PATH = LIB_DIR / "libfedmlp_b200.so"
OBJ_DIR = PKG / "build"

SOURCES = ["cabi.cu", "adam.cu", "fedavg.cu", "fedavg_allreduce.cu", "fedavg_allreduce_q.cu", "proto.cu", "tag_sim.cu", "pool_tag.cu", "tag_select.cu", "loss.cu", "eval.cu"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
# keep the statically linked CUDA runtime private to this library (torch ships its own libcudart)
LINK_FLAGS = ["-Xlinker", "--exclude-libs=ALL", "-Xlinker", "-Bsymbolic"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _deps(src: Path) -> list[Path]:
    return [src, CSRC / "common.cuh", INCLUDE / "fedmlp_b200.h", Path(__file__)]


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def _compile(nvcc: str, src: Path, obj: Path) -> None:
    cmd = [nvcc, *ARCH_FLAGS, *NVCC_FLAGS, f"-I{INCLUDE}", f"-I{CSRC}", "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link libfedmlp_b200.so.  Returns its path."""
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    LIB_DIR.mkdir(exist_ok=True)
    jobs = []
    for name in SOURCES:
        src = CSRC / name
        obj = OBJ_DIR / (src.stem + ".o")
        if force or _stale(obj, _deps(src)):
            jobs.append((src, obj))
    if jobs:
        if verbose:
            print(f"[fedmlp_b200] nvcc: {', '.join(s.name for s, _ in jobs)}", file=sys.stderr)
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            for f in [ex.submit(_compile, nvcc, s, o) for s, o in jobs]:
                f.result()
    objs = [OBJ_DIR / (Path(n).stem + ".o") for n in SOURCES]
    if force or jobs or _stale(LIB_PATH, objs):
        tmp = LIB_PATH.with_suffix(".so.tmp")
        cmd = [nvcc, *ARCH_FLAGS, "-shared", *LINK_FLAGS, "-o", str(tmp), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB_PATH)
        if verbose:
            print(f"[fedmlp_b200] linked {LIB_PATH}", file=sys.stderr)
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p)
