"""CPU: the C-ABI library builds/loads, exports every symbol include/b21.h declares, and the host-side logic
(window grid, TTA variants, state_dict contract, factories, error behaviour) matches the reference's."""
import argparse
import ctypes
import os
import re

import pytest
import torch

from brats21_b200 import _lib, definer, inferers, networks, tta
from oracle import inference as oinf
from oracle import nets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    from brats21_b200.build import build_lib
    path = build_lib()
    lib = ctypes.CDLL(str(path))
    header = open(os.path.join(ROOT, "include", "b21.h")).read()
    declared = set(re.findall(r"\b(b21_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in b21.h but not exported"
    assert declared == set(_lib.exported_symbols())
    lib.b21_last_error.restype = ctypes.c_char_p
    assert lib.b21_version() >= 100
    # argument validation happens before any CUDA call: bad taps -> B21_ERR_BAD_ARG with a message
    lib.b21_conv3d_fwd.argtypes = _lib._SIGNATURES["b21_conv3d_fwd"]
    rc = lib.b21_conv3d_fwd(1, 8, 1, None, 1, 8, None, 1, 8, 8, 8, 8, 8, 5, 1, None)
    assert rc == -1 and b"taps" in lib.b21_last_error()
    assert lib.b21_conv_cout_padded(48) == 48 and lib.b21_conv_cout_padded(24) == 32
    assert lib.b21_conv_cout_padded(384) == 384 and lib.b21_conv_cout_padded(3) == 16


def test_window_grid_matches_oracle():
    for img, roi, ov in (((240, 240, 155), (128,) * 3, 0.25), ((240, 240, 160), (128,) * 3, 0.25),
                         ((40, 36, 29), (16,) * 3, 0.25), ((40, 36, 29), (48, 32, 16), 0.5), ((16, 16, 16), (16,) * 3, 0.0)):
        image = tuple(max(i, r) for i, r in zip(img, roi))
        assert inferers.window_origins(image, roi, ov) == oinf.window_grid(image, roi, ov)
    assert len(inferers.window_origins((240, 240, 155), (128,) * 3, 0.25)) == 18
    for n in (16, 32, 128):
        a = inferers.importance_profiles((n,) * 3, "gaussian", 0.125, "cpu")[0]
        assert torch.allclose(a, oinf.gaussian_profile(n).float(), atol=1e-7)
    with pytest.raises(ValueError):
        inferers._get_scan_interval((8, 8), (4, 4, 4), 3, 0.25)


def test_tta_variants_are_signed_permutations():
    comp = tta.get_tta_transforms()
    assert len(comp) == 16
    x = torch.arange(2 * 3 * 4 * 5, dtype=torch.float32).reshape(1, 2, 3, 4, 5)
    ovars = oinf.reference_tta()
    seen = set()
    for tr, ov in zip(comp, ovars):
        a = tr.augment_image(x)
        assert torch.equal(a, ov.augment_image(x))
        assert torch.equal(tr.deaugment_mask(a), x)
        perm, flip = tr.variant
        seen.add((perm, flip))
        # (perm, flip) reproduces the augmentation: A[a] = V[s], s_j = flip_j ? dim_j-1-a[perm_j] : a[perm_j]
        dims = x.shape[2:]
        idx = torch.meshgrid(*[torch.arange(s) for s in a.shape[2:]], indexing="ij")
        src = [(dims[j] - 1 - idx[perm[j]]) if flip[j] else idx[perm[j]] for j in range(3)]
        assert torch.equal(a, x[:, :, src[0], src[1], src[2]])
    assert len(seen) == 16
    flips = [tr.variant for tr in tta.get_flip8_transforms()]
    assert len(set(flips)) == 8 and all(p == (0, 1, 2) for p, _ in flips)


def test_state_dict_contract_and_factories():
    args = argparse.Namespace(model="equiunet", width=16, norm="group", act="relu", dropout=0, num_classes=3)
    v1 = definer.get_model(args)
    assert isinstance(v1, networks.EquiUnet)
    assert [k for k, _ in nets.v1_param_shapes(16)] == list(v1.state_dict().keys())
    for k, shape in nets.v1_param_shapes(16):
        assert tuple(v1.state_dict()[k].shape) == shape
    args.model = "equiunet_assp_evo"
    with pytest.warns(UserWarning):
        v2 = definer.get_network(args)
    assert [k for k, _ in nets.v2_param_shapes(16)] == list(v2.state_dict().keys())
    for k, shape in nets.v2_param_shapes(16):
        assert tuple(v2.state_dict()[k].shape) == shape
    assert sum(p.numel() for p in definer.get_model(argparse.Namespace(
        model="equiunet", width=48, norm="group", act="relu", dropout=0, num_classes=3)).parameters()) == 23154735
    args.model = "vnet"
    with pytest.raises(NameError):
        definer.get_model(args)
    # no CPU fallback: a CPU tensor is refused loudly
    with pytest.raises(RuntimeError):
        v1(torch.zeros(1, 4, 8, 8, 8))
    with pytest.raises(RuntimeError):
        inferers.sliding_window_inference(torch.zeros(1, 4, 8, 8, 8), 8, 1, v1)


def test_header_arity_matches_ctypes_signatures():
    """Every prototype of include/b21.h has as many parameters as its ctypes binding (brats21_b200/_lib.py)."""
    header = open(os.path.join(ROOT, "include", "b21.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"\b(?:int|long long|const char\*)\s+(b21_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", header)
    assert len(protos) >= 50
    seen = set()
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        seen.add(name)
        if name in ("b21_last_error", "b21_version"):
            assert n == 0
            continue
        assert name in _lib._SIGNATURES, name
        assert n == len(_lib._SIGNATURES[name]), f"{name}: header has {n} parameters, ctypes binding {len(_lib._SIGNATURES[name])}"
    assert seen == set(_lib.exported_symbols())


def test_pre_post_entry_points_validate_arguments_before_any_cuda_call():
    lib = _lib.load()
    err = lambda: lib.b21_last_error().decode()  # noqa: E731
    assert lib.b21_keep_components(None, None, 8, 8, 8, 10, None) == -1 and "null" in err()
    assert lib.b21_keep_components(1, 1, 0, 8, 8, 10, None) == -1 and "dims" in err()
    assert lib.b21_replace_rare_labels(1, 1, 8, 8, 8, 20, 5, None) == -1 and "axis" in err()
    assert lib.b21_replace_rare_labels(1, 1, 8, 8, 8, -1, 2, None) == -1 and "thresh" in err()
    assert lib.b21_labels_to_channels(None, None, 8, None) == -1
    assert lib.b21_foreground_bbox(None, 4, 8, 8, 8, None, None) == -1 and "null" in err()
    assert lib.b21_nonzero_stats(1, 0, 8, 8, 8, 1, 1, None) == -1 and "dims" in err()
    assert lib.b21_normalize_crop_pad(1, 1, 17, 8, 8, 8, 1, 1, 8, 8, 8, 0, 0, 0, 0.0, None) == -1
    assert lib.b21_grad_centralize(None, 0, None) == -1
    assert lib.b21_keep_components_workspace_bytes(1000) == 8016
    assert lib.b21_replace_rare_workspace_bytes(20) == 4 * (258 + 2 * 256 * 20)
    # split-input conv: the second tensor is mandatory and the channel split must be chunk-aligned
    assert lib.b21_conv3d_march_fwd_fold2(1, 24, 24, None, 24, 1, 0, None, None, 1, 48, None, None, 1, 1, 8, 16, 16, 48, 48,
                                          None) == -1 and "second input" in err()
