"""TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference hot-path modules — from /root/reference (read-only, build container) or from the
verbatim copy oracle/make_ref.py places under oracle/_ref (git-ignored; travels to the GPU box with gpurun snapshots) —
with the MONAI shim and py3.12 compatibility aliases.  Used by tests/golden/make_golden*.py, by the tests that
cross-check the oracle port / drive the reference Engine, and by bench.py's `--impl reference` / cpu_baseline legs.
"""
from __future__ import annotations

import collections
import collections.abc
import contextlib
import io
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ARCHIVE = os.path.join(_HERE, "_ref", "reference_py.tar.gz")


def _reference_root() -> str:
    """/root/reference exists only in the build container; elsewhere (the GPU box) the archive oracle/make_ref.py
    built there (oracle/_ref/reference_py.tar.gz: git-ignored, NOT gpurun-ignored) is unpacked into a scratch dir."""
    env = os.environ.get("B21_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/networks"):
        return "/root/reference"
    if os.path.exists(_ARCHIVE):
        import hashlib
        import tarfile
        import tempfile
        tag = hashlib.sha256(open(_ARCHIVE, "rb").read()).hexdigest()[:16]
        dst = os.path.join(tempfile.gettempdir(), f"b21_reference_{tag}")
        if not os.path.isdir(os.path.join(dst, "networks")):
            tmp = dst + f".{os.getpid()}"
            with tarfile.open(_ARCHIVE) as tar:
                tar.extractall(tmp, filter="data")
            try:
                os.rename(tmp, dst)
            except OSError:  # another process won the race
                import shutil
                shutil.rmtree(tmp, ignore_errors=True)
        return dst
    return "/root/reference"


REFERENCE_ROOT = _reference_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "networks"))


def load():
    """Returns a namespace with the reference modules: equiunet2020, equiunet2021, inferers, tta, optimizer."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    from . import monai_shim
    monai_shim.install()
    # utils/misc.py:6 does `from collections import Sequence` (removed in py3.10)
    for name in ("Sequence", "Iterable", "Mapping"):
        if not hasattr(collections, name):
            setattr(collections, name, getattr(collections.abc, name))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import types
    ns = types.SimpleNamespace()
    with contextlib.redirect_stdout(io.StringIO()):
        import networks.equiunet2020 as e20
        import networks.equiunet2021 as e21
        import networks.factory as factory
        import utils.inferers as inferers
        import utils.misc as misc
        import tta as tta
        import learning.optimizer as optimizer
    ns.equiunet2020, ns.equiunet2021, ns.factory = e20, e21, factory
    ns.inferers, ns.misc, ns.tta, ns.optimizer = inferers, misc, tta, optimizer
    return ns


@contextlib.contextmanager
def quiet():
    """The reference constructors print; keep test output clean."""
    import warnings
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


# ------------------------------------------------------------------------------------------ whole Engine
class _EnumMeta(type):
    """Class attributes in CAPITALS resolve to their lower-case name (LossReduction.MEAN -> "mean"): enough for the
    default-argument expressions evaluated while the reference modules are imported."""

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return name.lower()


def _placeholder(name):
    return _EnumMeta(name, (), {"__init__": lambda self, *a, **k: None})


def _install_import_stubs():
    """IMPORT-ONLY stand-ins for third-party packages absent from this image (SimpleITK, skimage, nibabel, openpyxl,
    oyaml, tensorboard, ranger21 and the parts of monai outside oracle/monai_shim.py): learning/engine.py and
    src/definer.py import them at module scope, the functions the tests call never touch them."""
    import importlib.abc
    import importlib.machinery
    import types

    class Stub(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            v = _placeholder(name)
            setattr(self, name, v)
            return v

    # consulted LAST on sys.meta_path, i.e. only for modules the image really lacks (sklearn and pandas exist in the
    # build container but not on every GPU box)
    tops = {"SimpleITK", "skimage", "nibabel", "openpyxl", "oyaml", "tensorboard", "tensorboardX", "ranger21", "monai",
            "matplotlib", "seaborn", "apex", "medpy", "sklearn", "pandas", "scipy", "yaml", "tqdm"}

    class Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, name, path, target=None):
            if name.split(".")[0] in tops and name not in sys.modules:
                return importlib.machinery.ModuleSpec(name, self, is_package=True)
            return None

        def create_module(self, spec):
            m = Stub(spec.name)
            m.__path__ = []
            return m

        def exec_module(self, module):
            pass

    if any(type(f).__name__ == "Finder" and getattr(f, "_b21", False) for f in sys.meta_path):
        return
    for k, m in list(sys.modules.items()):
        if k.split(".")[0] == "monai":
            if not hasattr(m, "__path__"):
                m.__path__ = []

            def ga(name, m=m):
                if name.startswith("__") or name.islower():
                    raise AttributeError(name)
                v = _placeholder(name)
                setattr(m, name, v)
                return v
            m.__getattr__ = ga
    f = Finder()
    f._b21 = True
    sys.meta_path.append(f)


def load_engine():
    """The UNMODIFIED ``learning.engine`` (class Engine) and ``src.definer`` modules of the reference."""
    ns = load()
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # utils/transforms.py:503 uses the alias removed in numpy 1.24
    _install_import_stubs()
    with contextlib.redirect_stdout(io.StringIO()):
        import learning.engine as engine
        import src.definer as definer
    ns.engine, ns.definer = engine, definer
    return ns
