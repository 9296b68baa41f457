"""Rounding-noise budget of the 16-bit path, on the CPU (no GPU needed): tests/noise_model.py restates the network with a
rounding hook at every point where the CUDA path stores an activation / gradient in 16 bits, and this script switches
those sites on one at a time and in the combinations that were considered for round 2.

    python tools/noise_budget.py [H=64] [seed=5]  > profiles/r02_noise_budget.txt

Metric: max|got - ref| / max|ref| per tensor against the unrounded fp32 run (= the oracle); gradient columns are over
the 225 parameter tensors of config/baseline.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _parity import BASELINE_KW, make_inputs, rel_err  # noqa: E402
import noise_model as NM  # noqa: E402
from oracle import cesm_oracle as O  # noqa: E402


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    from cesm_emulator_b200.model import UNet
    torch.manual_seed(0)
    sd = {k: v.detach().float() for k, v in UNet(**BASELINE_KW).state_dict().items()}
    cfg, buf = O.OracleConfig.from_unet_kwargs(**BASELINE_KW), O.diffusion_buffers(1000)
    x0, cond, t, noise = make_inputs(2, 3, H, H, seed=seed)

    def run(rnd, S=1.0):
        leaves = {}
        for k, v in sd.items():
            v = v.detach().clone()
            if v.is_floating_point() and not k.endswith("rotary_emb.freqs"):
                v.requires_grad_(True)
            leaves[k] = v
        x_t = O.q_sample(buf, x0, t, noise).unsqueeze(2).expand(-1, -1, cond.shape[2], -1, -1)
        out = NM.unet3d_forward(leaves, cfg, x_t, t, cond, rnd)
        eps = out[:, :, out.shape[2] // 2]
        loss = F.mse_loss(eps, noise)
        names = [k for k, v in leaves.items() if v.requires_grad]
        grads = torch.autograd.grad(loss * S, [leaves[k] for k in names])
        return eps.detach(), {k: g / S for k, g in zip(names, grads)}

    e0, g0 = run(NM.Rounder(()))

    def row(name, rnd, S=1.0):
        e, g = run(rnd, S)
        errs = np.array(sorted(rel_err(g[k], g0[k]) for k in g))
        print(f"{name:58s} fwd {rel_err(e, e0):.2e} | grad median {np.median(errs):.2e} p90 {errs[int(.9 * len(errs))]:.2e} "
              f"max {errs[-1]:.2e}  <=1e-2: {(errs <= 1e-2).mean():.2f}", flush=True)

    acts = [c for c in NM.ALL if c != "w"]
    bf = {c: "bf16" for c in acts}
    h16 = {c: "fp16" for c in NM.ALL}
    print(f"# config/baseline, B=2, K=3, {H}x{H}, input seed {seed}, t={t.tolist()}; categories: {NM.__doc__.split('Categories')[1].strip()}")
    print("# --- bf16 everywhere (round 1's path), one rounding site at a time, then all of them ---")
    for c in NM.ALL:
        row(f"bf16, only '{c}' rounded", NM.Rounder((c,)))
    row("bf16, all sites (= round 1: measured 0.9-1.05e-2 on B200)", NM.Rounder(NM.ALL))
    print("# --- alternatives considered ---")
    row("fp16 weights, bf16 activations and gradients", NM.Rounder({**bf, "w": "fp16"}, bf))
    row("  + residual stream kept as (hi, lo) bf16 pairs", NM.Rounder({**bf, "w": "fp16"}, bf, ("res",)))
    row("  + conv outputs too", NM.Rounder({**bf, "w": "fp16"}, bf, ("res", "conv")))
    row("fp16 forward (weights + activations), bf16 gradients", NM.Rounder(h16, bf))
    row("fp16 forward, UNSCALED fp16 gradients (underflow)", NM.Rounder(h16, {c: "fp16" for c in acts}))
    print("# --- chosen: fp16 everywhere + loss scaling (the reference's autocast + GradScaler recipe, train.py:853-867) ---")
    for k in (16, 20, 24):
        row(f"fp16 forward, fp16 gradients, loss scale 2^{k}", NM.Rounder(h16, {c: "fp16" for c in acts}), 2.0 ** k)


if __name__ == "__main__":
    main()
