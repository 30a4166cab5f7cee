"""TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference hot-path modules from /root/reference (read-only, present only in the build
container — never on the GPU box) with the MONAI shim and py3.12 compatibility aliases.  Used solely by
tests/golden/make_golden.py and by CPU tests that cross-check the oracle port when the reference is present.
"""
from __future__ import annotations

import collections
import collections.abc
import contextlib
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("B21_REFERENCE_ROOT", "/root/reference")


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
