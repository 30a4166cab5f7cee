"""TEST / BASELINE INFRASTRUCTURE ONLY — recipe for oracle/_ref.

The reference is pure Python (no native code to compile): the "build" of oracle/_ref packs its .py files, read where
they lie under /root/reference (read-only, build container), into ONE binary artefact oracle/_ref/reference_py.tar.gz —
git-ignored, so no reference source ever enters the repository, but not gpurun-ignored, so the UNMODIFIED reference
travels to the GPU box where /root/reference does not exist.  oracle/ref_loader.py unpacks it into a scratch directory
at import time.  There it serves (with the MONAI shim) as
  * the CPU arm of bench.py (`--impl reference`, cpu_baseline.kind = "reference"),
  * the Engine that tests/test_gpu_reference_engine.py drives with brats21_b200 modules patched in.
Called by __graft_entry__.build() when /root/reference is present;   python -m oracle.make_ref   does the same."""
from __future__ import annotations

import io
import os
import tarfile

SRC = "/root/reference"
DST_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
ARCHIVE = os.path.join(DST_DIR, "reference_py.tar.gz")
PACKAGES = ("networks", "utils", "tta", "learning", "src")


def make() -> str | None:
    """Returns the archive path (None when neither the reference tree nor a previously built archive exists)."""
    if not os.path.isdir(os.path.join(SRC, "networks")):
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    os.makedirs(DST_DIR, exist_ok=True)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:gz") as tar:
        for pkg in PACKAGES:
            for root, dirs, files in os.walk(os.path.join(SRC, pkg)):
                dirs.sort()
                for f in sorted(files):
                    if f.endswith(".py"):
                        full = os.path.join(root, f)
                        info = tar.gettarinfo(full, arcname=os.path.relpath(full, SRC))
                        info.mtime, info.uid, info.gid, info.uname, info.gname = 0, 0, 0, "", ""
                        with open(full, "rb") as fh:
                            tar.addfile(info, fh)
    data = buf.getvalue()
    if not os.path.exists(ARCHIVE) or open(ARCHIVE, "rb").read() != data:
        with open(ARCHIVE, "wb") as fh:
            fh.write(data)
    return ARCHIVE


if __name__ == "__main__":
    print(make())
