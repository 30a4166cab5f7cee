"""GPU bring-up check of the full networks and the inference wrappers against the oracle (torch fp32 on the GPU,
TF32 off).  python tools/gpu_net_check.py [--big]"""
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import engine, inferers, networks, tta  # noqa: E402
from oracle import inference as oinf  # noqa: E402
from oracle import nets, synth  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda"


def stats(name, got, ref):
    got, ref = got.float(), ref.float()
    diff = (got - ref)
    rel = (diff.norm() / ref.norm().clamp_min(1e-12)).item()
    print(f"  {name}: max_abs {diff.abs().max().item():.4e} rel_l2 {rel:.4e} ref_absmax {ref.abs().max().item():.3f}",
          flush=True)
    return rel


def net_case(ver, width, shape, n=1, seed=123):
    print(f"== V{ver} width {width} shape {shape} n {n}")
    params = {k: v.to(DEV) for k, v in synth.make_params(ver, width, seed).items()}
    feats = [width * 2 ** i for i in range(4)]
    cls = networks.EquiUnet if ver == 1 else networks.EquiUnetASSPEvo
    net = cls(4, 3, feats, norm_layer="group", act="relu", deep_supervision=True).to(DEV).eval()
    net.load_state_dict(params)
    x = torch.cat([synth.volume(seed=s, shape=shape) for s in range(n)]).to(DEV)
    fwd = nets.equiunet_v1_forward if ver == 1 else nets.equiunet_v2_forward
    with torch.no_grad():
        ref_out, ref_deeps = fwd(params, x)
        out, deeps = net(x)
    torch.cuda.synchronize()
    r = stats("out", out, ref_out)
    for i, (a, b) in enumerate(zip(deeps, ref_deeps)):
        stats(f"deep{i}", a, b)
    flips = ((out >= 0) != (ref_out >= 0)).float().mean().item()
    print(f"  sign flips {flips:.4%}")
    return net, params, r


def main():
    ok = True
    try:
        net_case(1, 16, (32, 32, 32))
        net_case(2, 16, (32, 32, 32), seed=93)
        net_case(1, 16, (16, 24, 40), n=2)
        net_case(2, 16, (16, 24, 40), n=2, seed=93)
        if "--big" in sys.argv:
            net_case(1, 48, (64, 64, 64))
            net_case(2, 48, (64, 64, 64), seed=93)
    except Exception:  # noqa: BLE001
        traceback.print_exc()
        ok = False

    # ---- sliding window + TTA + labels on a small volume
    try:
        width, roi = 16, (32, 32, 32)
        net, params, _ = net_case(2, width, (32, 32, 32), seed=93)
        vol = synth.volume(seed=1, shape=(48, 40, 56)).to(DEV)
        fwd = lambda z: nets.equiunet_v2_forward(params, z)  # noqa: E731
        for mode in ("constant", "gaussian"):
            with torch.no_grad():
                ref = oinf.sliding_window_inference(vol.cpu(), roi, 2, lambda z: fwd(z.to(DEV))[0].cpu(), 0.25, mode)
                got = inferers.sliding_window_inference(vol, roi, 2, net, overlap=0.25, mode=mode)
            stats(f"sliding_window[{mode}]", got.cpu(), ref)
        for name, comp, ovar in (("tta16", tta.get_tta_transforms(), oinf.reference_tta()),
                                 ("flip8", tta.get_flip8_transforms(), oinf.flip8_tta())):
            with torch.no_grad():
                outs = oinf.apply_tta(lambda z: oinf.sliding_window_inference(
                    z.contiguous().cpu(), roi, 2, lambda q: fwd(q.to(DEV))[0].cpu(), 0.25, "gaussian"), vol.cpu(), ovar)
                prob_ref, hard_ref = oinf.ensemble_mean_threshold(outs)
                hard_ref = oinf.remove_background_voxels(vol.cpu(), hard_ref)
                lab_ref = oinf.brats_label_map(hard_ref)
                onehot, label, prob = engine.predict_volume([net], vol, comp, True, roi, 2, 0.25, "gaussian",
                                                            return_prob=True)
            stats(f"{name} mean prob", prob.cpu()[None], prob_ref)
            margin = (prob_ref - 0.5).abs() > 0.02
            mism = ((onehot.cpu().float() != hard_ref) & margin).sum().item()
            print(f"  {name}: onehot mismatches outside margin: {mism}; label agreement "
                  f"{(label.cpu() == lab_ref).float().mean().item():.5f}")
            ok = ok and mism == 0
    except Exception:  # noqa: BLE001
        traceback.print_exc()
        ok = False
    print("NET_CHECK_OK" if ok else "NET_CHECK_FAILED")


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("elapsed", time.time() - t0)
