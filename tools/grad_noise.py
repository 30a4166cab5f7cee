"""Run-to-run noise of the flat gradient of one w16 / 32^3 training step (fp32 atomics order + bf16 roundings):
the tolerance of tests/test_gpu_train.py::test_data_parallel_gradients_are_rank_sums is set against this."""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import engine, networks, ops  # noqa: E402
from brats21_b200.losses import DiceLoss  # noqa: E402
from oracle import synth  # noqa: E402

DEV = torch.device("cuda:0")


def grads(seed):
    width = 16
    params = {k: v.to(DEV) for k, v in synth.make_params(2, width, 93).items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = networks.EquiUnetASSPEvo(4, 3, [width * 2 ** i for i in range(4)], deep_supervision=True).to(DEV)
    net.load_state_dict(params)
    net.train()
    tgt = synth.target(shape=(32, 32, 32)).to(DEV)
    net.zero_grad()
    x = synth.volume(seed=seed, shape=(32, 32, 32)).to(DEV)
    _, loss = engine.compute_loss(None, DiceLoss(), net(x), tgt)
    loss.backward()
    torch.cuda.synchronize()
    return net.grad_store().flat.clone(), float(loss)


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


for use in (True, False):
    ops.use_input = use
    runs = [grads(10) for _ in range(6)]
    ref = runs[0][0]
    print(f"use_input={use}: loss {[round(r[1], 6) for r in runs]}")
    print("   rel vs run 0:", [f"{rel(r[0], ref):.2e}" for r in runs[1:]], flush=True)
ops.use_input = True
a = grads(10)[0]
ops.use_input = False
b = grads(10)[0]
print(f"input kernel vs march kernel: rel {rel(a, b):.2e}")
