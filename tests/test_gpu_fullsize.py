"""Parity AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs 2-4, north_star): width 48, 4 x 128^3 window
batches through the CUDA-graph inference path, the whole 4x240x240x155 sliding-window (+ 8-flip TTA) pipeline, and
one 128^3 training step — against the fp32 oracle run on the same GPU (TF32 off).

Stated tolerances (bf16 storage, fp32 accumulation):
  * logits: relative L2 <= 2.5e-2 and max-abs <= 4e-2 * max|logit| (same as tests/test_gpu_networks.py);
  * blended pipeline output: the logit tolerance is 2e-2 * max|blended reference logit| (tighter than the per-window
    4e-2: blending averages overlapping windows); a sigmoid has slope <= 1/4, so the mean probability map must agree
    within a quarter of that (V2: ~0.01, V1 with its kaiming-init logits of std 4: ~0.1), and the hard labels must be
    bit-exact wherever the oracle's probability is further than that from the threshold;
  * per-region (TC, WT, ET) Dice agreement of the label maps >= 0.999, or — for random-init networks whose logits
    crowd the threshold — a Dice deficit no larger than 1.5x the deficit torch's own autocast(bf16) run of the
    reference code shows against its fp32 run on the same input (both numbers are recorded);
  * training: loss within 5e-3; every parameter gradient within 5 % (V2) / 12 % (V1) relative L2 of the bf16-storage
    emulation of the oracle (median < 1 %; V1's ReLU masks and max-pool winners flip on bf16 ties, which the first
    layer's gradient integrates over 2 M voxels), and within 1.6x torch-autocast's own drift (+6 %) of the plain fp32
    oracle (measured: V2 2.8 % vs autocast 3.5 %, V1 14.6 % vs 15.3 %, worst tensor).
Every measured figure is written to gpurun_out/parity_full_size.json (committed copy: profiles/r02_parity_full_size.json,
quoted by bench.py's `parity` block)."""
import json
import os
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out", "parity_full_size.json")
REL_L2_TOL, MAX_ABS_TOL, PIPE_LOGIT_TOL = 2.5e-2, 4e-2, 2e-2
ROI = (128, 128, 128)


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.cuda.empty_cache()


def _record(key, value):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    data = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            data = json.load(f)
    data[key] = value
    with open(OUT, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


def _build(ver, seed, train=False):
    from brats21_b200 import networks
    from oracle import synth
    params = {k: v.to(DEV) for k, v in synth.make_params(ver, 48, seed).items()}
    cls = networks.EquiUnet if ver == 1 else networks.EquiUnetASSPEvo
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = cls(4, 3, [48 * 2 ** i for i in range(4)], norm_layer="group", act="relu", deep_supervision=True).to(DEV)
    net.load_state_dict(params, strict=True)
    return (net.train() if train else net.eval()), params


def _fwd(ver):
    from oracle import nets
    return nets.equiunet_v1_forward if ver == 1 else nets.equiunet_v2_forward


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _dice(a, b):
    a, b = a.bool(), b.bool()
    den = a.sum().item() + b.sum().item()
    return 2.0 * (a & b).sum().item() / den if den else 1.0


@pytest.mark.parametrize("ver,seed,batch", [(2, 93, 9), (1, 123, 4)])
def test_w48_window_batch_graph_forward_matches_oracle(ver, seed, batch):
    """The exact code path bench.py times per window batch: pack -> forward_infer (CUDA graph; V2: folded EvoNorm on
    levels 1-3, split level-1 concat, 128-plane marches) on a batch of 128^3 windows (9 for v2_tta8, 4 for v1_sw), vs
    the fp32 oracle."""
    from brats21_b200 import ops
    from oracle import synth
    net, params = _build(ver, seed)
    x = torch.cat([synth.volume(seed=s, shape=ROI) for s in range(batch)]).to(DEV)
    assert ops.use_graphs and ops.use_fold and ops.fold_level3 and ops.split_concat
    with torch.no_grad():
        x8 = net.pack_input(x)
        first = net.forward_infer(x8).clone()   # warm-up + capture + first replay
        out = net.forward_infer(x8).clone()     # pure replay
        assert len(net._graphs) == 1
        ref = torch.cat([_fwd(ver)(params, x[i:i + 1], deep_supervision=False) for i in range(batch)])
        with torch.autocast("cuda", dtype=torch.bfloat16):
            auto = torch.cat([_fwd(ver)(params, x[i:i + 1], deep_supervision=False).float() for i in range(batch)])
    rel, mx = _rel(out, ref), ((out - ref).abs().max() / ref.abs().max()).item()
    rel_a, mx_a = _rel(auto, ref), ((auto - ref).abs().max() / ref.abs().max()).item()
    margin = ref.abs() > MAX_ABS_TOL * ref.abs().max()
    flips = (((out >= 0) != (ref >= 0)) & margin).sum().item()
    dice = [_dice(out[:, c] >= 0, ref[:, c] >= 0) for c in range(3)]
    dice_a = [_dice(auto[:, c] >= 0, ref[:, c] >= 0) for c in range(3)]
    _record(f"forward_v{ver}_w48_{batch}x128", dict(
        rel_l2=rel, max_abs_over_max=mx, replay_vs_first_rel_l2=_rel(out, first), sign_flips_outside_margin=flips,
        dice_logit_sign=dice, torch_autocast_bf16=dict(rel_l2=rel_a, max_abs_over_max=mx_a, dice_logit_sign=dice_a),
        tol=dict(rel_l2=REL_L2_TOL, max_abs_over_max=MAX_ABS_TOL)))
    assert rel <= REL_L2_TOL and mx <= MAX_ABS_TOL, (rel, mx)
    assert _rel(out, first) <= 1.5e-2  # replays differ only by the order of the statistics atomics
    assert flips == 0


@pytest.mark.parametrize("name,ver,seed,tta_name,mode,sw_batch", [("v2_tta8", 2, 93, "flip8", "gaussian", 9),
                                                                  ("v1_sw", 1, 123, None, "constant", 4)])
def test_full_volume_pipeline_matches_oracle(name, ver, seed, tta_name, mode, sw_batch):
    """BASELINE configs 3 and 2 end to end: one synthetic 4x240x240x155 volume -> shape_to_divisible ->
    (8 flips x) 18 windows of 128^3 in bench.py's batches (9 / 4) -> blend -> sigmoid -> mean -> threshold -> background
    removal -> label map, exactly bench.py's step, against the oracle pipeline in fp32 on the GPU."""
    from brats21_b200 import engine, tta
    from oracle import inference as oinf
    from oracle import synth
    net, params = _build(ver, seed)
    vol0 = synth.volume(seed=1000, shape=(240, 240, 155))
    vol, pb, pa = oinf.shape_to_divisible(vol0, 8)
    vol = vol.to(DEV)
    assert tuple(vol.shape[2:]) == (240, 240, 160)
    comp = tta.get_flip8_transforms() if tta_name else None
    ovar = oinf.flip8_tta() if tta_name else [oinf.Variant("id", lambda x: x, lambda x: x)]
    onehot, label, prob = engine.predict_volume([net], vol, comp, True, ROI, sw_batch, 0.25, mode, return_prob=True)
    fwd = lambda z: _fwd(ver)(params, z.contiguous(), deep_supervision=False)  # noqa: E731

    def oracle_pipeline():
        with torch.no_grad():
            outs = oinf.apply_tta(lambda z: oinf.sliding_window_inference(z.contiguous(), ROI, 1, fwd, 0.25, mode), vol,
                                  ovar)
            p, hard = oinf.ensemble_mean_threshold(outs)
            lmax = max(o.abs().max().item() for o in outs)
        return p, oinf.remove_background_voxels(vol, hard), lmax

    prob_ref, hard_ref, logit_max = oracle_pipeline()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        prob_auto, hard_auto, _ = oracle_pipeline()
    PROB_TOL = 0.25 * PIPE_LOGIT_TOL * logit_max
    lab_ref = oinf.brats_label_map(hard_ref)
    agree_auto = (oinf.brats_label_map(hard_auto) == lab_ref).float().mean().item()
    dprob = (prob[None] - prob_ref).abs().max().item()
    dprob_auto = (prob_auto - prob_ref).abs().max().item()
    margin = (prob_ref - 0.5).abs() > PROB_TOL
    mism = ((onehot.float() != hard_ref) & margin).sum().item()
    total_mism = (onehot.float() != hard_ref).sum().item()
    dice = [_dice(onehot[0, c], hard_ref[0, c]) for c in range(3)]
    dice_auto = [_dice(hard_auto[0, c], hard_ref[0, c]) for c in range(3)]
    agree = (label == lab_ref).float().mean().item()
    _record(f"pipeline_{name}_240x240x155", dict(
        windows=18 * len(ovar), sw_batch_size=sw_batch, prob_max_abs=dprob, label_bits_differing_outside_margin=mism,
        label_bits_differing_total=total_mism, voxels=int(prob_ref[0, 0].numel()), label_map_agreement=agree,
        dice_tc_wt_et=dice, region_voxels=[int(hard_ref[0, c].sum().item()) for c in range(3)],
        reference_logit_max_abs=logit_max,
        torch_autocast_bf16=dict(prob_max_abs=dprob_auto, dice_tc_wt_et=dice_auto, label_map_agreement=agree_auto,
                                 label_bits_differing_total=int((hard_auto != hard_ref).sum().item())),
        tol=dict(logit_max_abs_over_max=PIPE_LOGIT_TOL, prob_max_abs=PROB_TOL, dice=0.999,
                 deficit_vs_autocast=1.5)))
    assert dprob <= PROB_TOL, (dprob, PROB_TOL)
    assert mism == 0, mism
    assert agree >= 0.999 or (1.0 - agree) <= 1.5 * (1.0 - agree_auto) + 1e-4, (agree, agree_auto)
    for c in range(3):
        if hard_ref[0, c].sum().item() >= 5000:
            assert dice[c] >= 0.999 or (1.0 - dice[c]) <= 1.5 * (1.0 - dice_auto[c]) + 1e-4, (c, dice, dice_auto)
    assert (label[0, 0][(vol[0] == 0).all(0)] == 0).all()
    # crop back to the original 240x240x155 grid exactly as bench.py's e2e step does
    lab155 = oinf.shape_to_original(label, pb, pa)
    assert tuple(lab155.shape[2:]) == (240, 240, 155)


def _reference_step(ver, params, x, tgt, jaccard=False):
    from oracle import train as otrain
    ps = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith("running_var"))
          for k, v in params.items()}
    out, deeps = _fwd(ver)(ps, x)
    loss = otrain.deep_supervision_loss([out] + list(deeps), tgt, jaccard)
    loss.backward()
    grads = {k: v.grad for k, v in ps.items() if v.grad is not None}
    return loss.detach(), grads, out.detach()


@pytest.mark.parametrize("ver,seed", [(2, 93), (1, 123)])
def test_w48_128cube_training_step_matches_oracle(ver, seed):
    """BASELINE config 4's step at its real size: forward, Dice over all heads, backward — loss and every parameter
    gradient (conv_wgrad_march 48x48 @128^3, conv_wgrad at levels 3-4, norm/SE/pool/upsample adjoints)."""
    from test_gpu_train import _bf16_oracle
    from brats21_b200 import engine
    from brats21_b200.losses import DiceLoss
    from oracle import synth
    net, params = _build(ver, seed, train=True)
    x = synth.volume(seed=2000, shape=ROI).to(DEV)
    tgt = synth.target(shape=ROI).to(DEV)
    net.zero_grad()
    outputs = net(x)
    _, loss = engine.compute_loss(None, DiceLoss(), outputs, tgt)
    loss.backward()
    torch.cuda.synchronize()
    got = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
    out = outputs[0].detach().clone()
    del outputs
    ref_loss, ref_grads, ref_out = _reference_step(ver, params, x, tgt)
    with _bf16_oracle():
        emu_loss, emu_grads, _ = _reference_step(ver, params, x, tgt)
    ps2 = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith("running_var"))
           for k, v in params.items()}
    from oracle import train as otrain
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o2, d2 = _fwd(ver)(ps2, x)
    otrain.deep_supervision_loss([o2.float()] + [t.float() for t in d2], tgt, False).backward()
    auto_grads = {k: v.grad for k, v in ps2.items() if v.grad is not None}
    assert set(got) == set(ref_grads)
    e_fp32 = {k: _rel(got[k], g) for k, g in ref_grads.items()}
    e_emu = {k: _rel(got[k], g) for k, g in emu_grads.items()}
    e_auto = {k: _rel(auto_grads[k], g) for k, g in ref_grads.items()}
    srt = lambda d: sorted(d.items(), key=lambda kv: -kv[1])  # noqa: E731
    med = lambda d: sorted(d.values())[len(d) // 2]  # noqa: E731
    _record(f"train_step_v{ver}_w48_128", dict(
        loss=loss.item(), loss_fp32_oracle=ref_loss.item(), loss_bf16_emulating_oracle=emu_loss.item(),
        logits_rel_l2=_rel(out, ref_out), tensors=len(got),
        grad_rel_l2_vs_fp32=dict(max=srt(e_fp32)[0][1], median=med(e_fp32), worst=srt(e_fp32)[:3]),
        grad_rel_l2_vs_bf16_emulation=dict(max=srt(e_emu)[0][1], median=med(e_emu), worst=srt(e_emu)[:3]),
        torch_autocast_grad_rel_l2_vs_fp32=dict(max=srt(e_auto)[0][1], median=med(e_auto), worst=srt(e_auto)[:3])))
    assert _rel(out, ref_out) <= 3e-2
    assert abs(loss.item() - ref_loss.item()) <= 5e-3
    bad = {k: v for k, v in e_emu.items() if v > (0.05 if ver == 2 else 0.12)}
    assert not bad, f"gradient mismatch vs the bf16-emulating oracle: {srt(bad)[:8]}"
    assert med(e_emu) <= 0.01, med(e_emu)
    bad = {k: (v, e_auto[k]) for k, v in e_fp32.items() if v > 1.6 * e_auto[k] + 0.06}
    assert not bad, f"gradient error vs fp32 beyond torch-autocast's own drift: {sorted(bad.items(), key=lambda kv: -kv[1][0])[:8]}"


@pytest.mark.parametrize("cin,cout,size", [(48, 48, 128), (96, 96, 64), (8, 48, 128)])
def test_wgrad_march_at_benchmark_shapes(cin, cout, size):
    """conv_wgrad_march on the level-1 / level-2 shapes of the width-48 training step vs fp32 autograd of F.conv3d
    on the same bf16-rounded operands (fp32 accumulation over 2.1 M voxels: 1e-3 relative)."""
    import torch.nn.functional as F
    from brats21_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(cin + cout)
    x = torch.randn((1, size, size, size, cin), device=DEV, generator=g).to(torch.bfloat16)
    dz = torch.randn((1, size, size, size, cout), device=DEV, generator=g).to(torch.bfloat16)
    cin_true = 4 if cin == 8 else cin
    if cin == 8:
        x[..., 4:] = 0
    dw = torch.zeros((cout, cin_true, 3, 3, 3), device=DEV)
    assert ops.use_wgrad_march
    ops.conv3d_wgrad(x, dz, dw)
    wt = torch.zeros((cout, cin_true, 3, 3, 3), device=DEV, requires_grad=True)
    xn = x.float().permute(0, 4, 1, 2, 3)[:, :cin_true].contiguous()
    y = F.conv3d(xn, wt, None, padding=1)
    y.backward(dz.float().permute(0, 4, 1, 2, 3).contiguous())
    rel = _rel(dw, wt.grad)
    _record(f"wgrad_march_{cin}x{cout}_{size}", dict(rel_l2=rel))
    assert rel <= 1e-3, rel
