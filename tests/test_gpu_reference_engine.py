"""INTEGRATION.md §1, executed: the UNMODIFIED reference ``learning.engine.Engine`` (``train``, ``evaluate``,
``_compute_output``, ``_compute_loss``, ``_apply_tta``) and the reference's own ``src.definer.get_tta_transforms`` /
``tta`` classes, with the brats21_b200 network, criterion, optimizer and ``sliding_window_inference`` patched in
exactly where a maintainer would put them — nothing else moves.

The reference tree is loaded from /root/reference (build container) or from oracle/_ref/reference_py.tar.gz (the GPU
box; built by oracle/make_ref.py during __graft_entry__.build()).  Third-party imports absent from the image are
import-only stubs (oracle/ref_loader.py); MONAI pieces the engine really calls come from oracle/monai_shim.py."""
import argparse
import contextlib
import io
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
WIDTH, SHAPE, ROI = 16, (40, 32, 48), (32, 32, 32)


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree / oracle/_ref archive not present (python -m oracle.make_ref in the build container)")
    ns = ref_loader.load_engine()
    from brats21_b200 import inferers
    ns.engine.sliding_window_inference = inferers.sliding_window_inference  # the one-line engine patch
    return ns


def _net(seed=93, train=False):
    from brats21_b200 import networks
    from oracle import synth
    params = {k: v.to(DEV) for k, v in synth.make_params(2, WIDTH, seed).items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = networks.EquiUnetASSPEvo(4, 3, [WIDTH * 2 ** i for i in range(4)], norm_layer="group", act="relu",
                                       deep_supervision=True).to(DEV)
    net.load_state_dict(params)
    return (net.train() if train else net.eval()), params


def _args(**kw):
    base = dict(no_amp=False, sliding_window_inference=True, sliding_window_size=ROI, criterion="dice",
                gradient_accumulation_iter=None, gradient_clipping=False, adaptive_gradient_clipping=False,
                log_train_metrics=False, log_val_metrics=False, log_train_interval=1, log_val_interval=1,
                no_tensorboard=True, swa_start=None, save_path=None)
    base.update(kw)
    return argparse.Namespace(**base)


class _Writer:
    def add_scalar(self, *a, **k):
        pass


def _engine(ref, model, criterion=None):
    eng = object.__new__(ref.engine.Engine)  # the constructor only moves modules to CUDA and builds metric callables
    eng.model, eng.criterion, eng.swa_model = model, criterion, None
    eng.key_metric = eng.additional_metrics = None
    eng.summary_writer, eng.labels = _Writer(), None
    eng.train_step = eng.val_step = eng.test_step = 0
    return eng


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def test_reference_compute_output_and_apply_tta_run_on_b200_modules(ref):
    from oracle import inference as oinf
    from oracle import nets, synth
    net, params = _net()
    img = synth.volume(seed=11, shape=SHAPE).to(DEV)
    eng = _engine(ref, net)
    fwd = lambda z: nets.equiunet_v2_forward(params, z.contiguous().to(DEV))[0]  # noqa: E731
    with torch.no_grad():
        # Engine._compute_output: autocast + sliding_window_inference(img, roi, sw_batch_size=1, predictor=model, device=cpu)
        out = ref.engine.Engine._compute_output(_args(), net, img)
        assert out.device.type == "cpu" and out.shape == (1, 3) + SHAPE
        want = oinf.sliding_window_inference(img, ROI, 1, fwd, 0.25, "constant").cpu()
        assert _rel(out, want) <= 2.5e-2
        # full-volume branch (the reference default): model(img) -> (out, [deeps])
        full = ref.engine.Engine._compute_output(_args(sliding_window_inference=False), net, img)
        assert isinstance(full, tuple) and len(full[1]) == 2
        assert _rel(full[0], fwd(img)) <= 2.5e-2
        # Engine._apply_tta with the reference's OWN 16-variant Compose (tta/base.py, tta/transforms.py unmodified)
        comp = ref.definer.get_tta_transforms()
        assert len(comp) == 16
        outs = eng._apply_tta(_args(), net, img, comp)
        assert len(outs) == 16 and all(o.device.type == "cpu" and o.shape == (1, 3) + SHAPE for o in outs)
        wants = oinf.apply_tta(lambda z: oinf.sliding_window_inference(z.contiguous(), ROI, 1, fwd, 0.25, "constant"),
                               img, oinf.reference_tta())
        for o, w in zip(outs, wants):
            assert _rel(o, w.cpu()) <= 2.5e-2


def test_reference_evaluate_loop_matches_fused_predict_volume(ref):
    """Engine.evaluate(use_tta=True) over a one-case loader with two models: shape_to_divisible -> eval_mode ->
    _apply_tta per model -> sigmoid -> mean -> post_trans -> remove_background_voxels, all reference code; the mean
    probability it hands to post_trans must agree with engine.predict_volume (the fused path bench.py times)."""
    from brats21_b200 import engine as b21engine
    from brats21_b200 import tta as b21tta
    from oracle import synth
    net_a, _ = _net(93)
    net_b, _ = _net(7)
    img = synth.volume(seed=12, shape=(40, 30, 45)).to(DEV)  # not divisible by 8: exercises shape_to_divisible
    seen = {}

    def post_trans(p):
        seen["prob"] = p.clone()
        return (p >= 0.5).float()

    eng = _engine(ref, [net_a, net_b])
    loader = [{"img": img.cpu(), "patient_id": ["case0"]}]
    with contextlib.redirect_stdout(io.StringIO()):
        losses, *_ = eng.evaluate(loader, 0, _args(), use_tta=True, post_trans=post_trans, activation=torch.sigmoid)
    assert tuple(seen["prob"].shape) == (1, 3, 40, 32, 48)
    from oracle import inference as oinf
    vol, _, _ = oinf.shape_to_divisible(img.cpu(), 8)
    _, _, prob = b21engine.predict_volume([net_a, net_b], vol.to(DEV), b21tta.get_tta_transforms(), True, ROI, 1, 0.25,
                                          "constant", return_prob=True)
    assert (prob.cpu()[None] - seen["prob"].cpu()).abs().max().item() <= 0.02


def test_reference_train_loop_steps_b200_modules(ref):
    """Engine.train (zero_grad -> _compute_output -> _compute_loss over all heads -> GradScaler.scale(loss).backward()
    -> scaler.step(optimizer) -> scaler.update()) for three batches with the brats21_b200 network, DiceLoss and fused
    Ranger2020, against engine.train_step on an identical second copy."""
    from brats21_b200 import engine as b21engine
    from brats21_b200.losses import DiceLoss
    from brats21_b200.optimizer import Ranger2020
    from oracle import synth
    from oracle import train as otrain
    from oracle import nets
    tgt = synth.target(shape=ROI)
    batches = [{"img": synth.volume(seed=30 + i, shape=ROI), "seg": tgt} for i in range(3)]

    def fresh():
        net, params = _net(train=True)
        opt = Ranger2020([p for n, p in net.named_parameters() if not n.endswith(".v")], lr=1e-3, weight_decay=1e-5,
                         use_gc=False)
        return net, opt, params

    net_r, opt_r, params = fresh()
    eng = _engine(ref, net_r, DiceLoss())
    scaler = torch.cuda.amp.GradScaler()
    with contextlib.redirect_stdout(io.StringIO()):
        losses, *_ = eng.train(batches, opt_r, scaler, 0, _args(sliding_window_inference=False))
    assert eng.train_step == 3 and losses.count == 3
    net_o, opt_o, _ = fresh()
    mine = [b21engine.train_step(None, net_o, DiceLoss(), opt_o, b["img"].to(DEV), b["seg"].to(DEV)).item()
            for b in batches]
    assert max(abs(a - b) for a, b in zip(losses.all_val, mine)) <= 2e-3
    for (n, p), q in zip(net_r.named_parameters(), net_o.parameters()):
        # same kernels; only the loss scale (a power of two) and the order of atomics differ
        assert _rel(p.detach(), q.detach()) <= 2e-3, n
    # and the first loss is the oracle's deep-supervision Dice on the initial weights
    with torch.no_grad():
        out, deeps = nets.equiunet_v2_forward(params, batches[0]["img"].to(DEV))
        want = otrain.deep_supervision_loss([out] + list(deeps), tgt.to(DEV)).item()
    assert abs(losses.all_val[0] - want) <= 5e-3
