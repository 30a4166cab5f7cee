"""Diagnostic: per-parameter relative gradient error of one training step (V1 / V2, width 16, 32^3) vs torch fp32
autograd on the oracle networks.  Run under gpurun:  python tools/gpu_train_check.py [1|2] [size]"""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from brats21_b200 import engine, networks  # noqa: E402
from brats21_b200.losses import DiceLoss  # noqa: E402
from oracle import nets, synth  # noqa: E402
from oracle import train as otrain  # noqa: E402

version = int(sys.argv[1]) if len(sys.argv) > 1 else 1
size = int(sys.argv[2]) if len(sys.argv) > 2 else 32
DEV = "cuda"
width = 16
params = {k: v.to(DEV) for k, v in synth.make_params(version, width, 123).items()}
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    cls = networks.EquiUnet if version == 1 else networks.EquiUnetASSPEvo
    net = cls(4, 3, [width * 2 ** i for i in range(4)], norm_layer="group", deep_supervision=True).to(DEV)
net.load_state_dict(params)
net.train()
x = synth.volume(seed=5, shape=(size,) * 3).to(DEV)
tgt = synth.target(shape=(size,) * 3).to(DEV)
net.zero_grad()
outputs = net(x)
_, loss = engine.compute_loss(None, DiceLoss(), outputs, tgt)
loss.backward()
ps = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith("running_var")) for k, v in params.items()}
fwd = nets.equiunet_v1_forward if version == 1 else nets.equiunet_v2_forward
out, deeps = fwd(ps, x)
ref_loss = otrain.deep_supervision_loss([out] + list(deeps), tgt, False)
ref_loss.backward()
# the same reference with bf16-rounded activations would be the fair comparison; print fp32 reference errors
rel = lambda a, b: ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()  # noqa: E731
print("loss", loss.item(), ref_loss.item(), "out rel", rel(outputs[0].detach(), out.detach()))
got = dict(net.named_parameters())
for name, p in ps.items():
    if p.grad is None:
        continue
    print(f"{name:45s} {rel(got[name].grad, p.grad):.4f}  |ref| {p.grad.norm().item():.3e}")

# how far does torch's own bf16 autocast drift from fp32 on the same step?  (context for the tolerances)
ps2 = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith("running_var")) for k, v in params.items()}
with torch.autocast("cuda", dtype=torch.bfloat16):
    out2, deeps2 = fwd(ps2, x)
loss2 = otrain.deep_supervision_loss([out2.float()] + [t.float() for t in deeps2], tgt, False)
loss2.backward()
print("\ntorch bf16-autocast vs fp32 (same oracle code):")
names = [n for n, p in ps.items() if p.grad is not None]
for name in names[:12] + names[-6:]:
    print(f"{name:45s} autocast {rel(ps2[name].grad, ps[name].grad):.4f}   ours {rel(got[name].grad, ps[name].grad):.4f}"
          f"   ours-vs-autocast {rel(got[name].grad, ps2[name].grad):.4f}")
