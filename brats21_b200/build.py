"""In-tree build of libb21.so (the C-ABI kernel library) and the C oracle helpers.

nvcc cross-compiles for sm_100a without a GPU; the resulting .so is git-ignored but travels with gpurun snapshots.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libb21.so"
STAMP = PKG_DIR / ".libb21.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-Xptxas", "-v",
]


# B21_NVCC_EXTRA: extra nvcc flags (e.g. -DB21_WAIT_HINT_NS=0 for an A/B build); B21_BUILD_OUT: library path of that build
import shlex  # noqa: E402

NVCC_FLAGS += shlex.split(os.environ.get("B21_NVCC_EXTRA", ""))
if os.environ.get("B21_BUILD_OUT"):
    LIB_PATH = Path(os.environ["B21_BUILD_OUT"]).resolve()
    STAMP = LIB_PATH.with_suffix(".stamp")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) +
                    [PKG_DIR.parent / "include" / "b21.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into one shared library. Rebuilds only when sources changed."""
    digest = _digest()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB_PATH
    objs = []
    build_dir = PKG_DIR / ("build" if not os.environ.get("B21_BUILD_OUT") else "build_" + LIB_PATH.stem)
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in _sources():
        obj = build_dir / (src.stem + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        log.append(f"==== {src.name}\n{out}")
        if pr.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src.name}")
        objs.append(str(obj))
    (build_dir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(cmd, check=True)
    STAMP.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    p = build_lib(force="--force" in sys.argv, verbose=True)
    print("built", p)
