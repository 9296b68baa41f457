"""Control experiment for the loss-curve parity record: how far do TWO fp32 runs of the reference algorithm (the CPU
oracle) drift apart over N steps when the initial weights differ by a relative 1e-5 (far below any 16-bit rounding)?
Same data, (t, noise), clip and AdamW as tools/loss_curve_parity.py.  If fp32-vs-fp32 already decorrelates, a
per-step relative loss difference says nothing about precision; window means do.

    python tools/loss_curve_control.py [steps=200] [H=128] [perturbation=1e-5] [out.json]      (CPU only, ~45 min)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _parity import BASELINE_KW  # noqa: E402
from oracle import cesm_oracle as O  # noqa: E402
from cesm_emulator_b200.synthetic import SyntheticEnsemble  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    H = W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    pert = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-5
    B = 2
    ds = SyntheticEnsemble(members=4, times=16, lat=H, lon=W, seed=7, K=3)
    g = torch.Generator().manual_seed(11)
    batches = []
    for _ in range(steps):
        cond, x0 = ds.batch(torch.randint(0, len(ds), (B,), generator=g).tolist(), augment=False)
        batches.append((cond, x0, torch.randint(0, 1000, (B,), generator=g), torch.randn(B, 1, H, W, generator=g)))
    cfg, buf = O.OracleConfig.from_unet_kwargs(**BASELINE_KW), O.diffusion_buffers(1000)
    from cesm_emulator_b200.model import UNet   # initial weights only (identical to the reference's under the same seed)
    torch.manual_seed(0)
    init = {k: v.detach().float().clone() for k, v in UNet(**BASELINE_KW).state_dict().items()}
    hp = dict(lr=2e-4, betas=(0.9, 0.999), weight_decay=1e-4, eps=1e-8)

    def run(p):
        sd = {k: v.clone() for k, v in init.items()}
        gg = torch.Generator().manual_seed(99)
        if p:
            for k, v in sd.items():
                if v.is_floating_point() and not k.endswith("freqs"):
                    v.mul_(1 + p * torch.randn(v.shape, generator=gg))
        names = [k for k, v in sd.items() if v.is_floating_point() and not k.endswith("rotary_emb.freqs")]
        leaves = [sd[k].requires_grad_(True) for k in names]
        opt = torch.optim.AdamW(leaves, **hp)
        out = []
        for cond, x0, t, noise in batches:
            loss, grads = O.loss_and_grads(sd, cfg, buf, x0, cond, t, noise)
            for k, q in zip(names, leaves):
                q.grad = grads[k]
            torch.nn.utils.clip_grad_norm_(leaves, 1.0)
            opt.step()
            out.append(loss.item())
        return np.array(out)

    a, b = run(0.0), run(pert)
    rel = np.abs(a - b) / np.abs(a)
    print(f"# fp32 oracle vs fp32 oracle with initial weights perturbed by {pert:g} (relative), {steps} steps, {H}x{W}, B={B}")
    for i in range(0, steps, 20):
        print(f"step {i:4d}  run A {a[i]:.6f}  run B {b[i]:.6f}  rel {rel[i]:.2e}")
    print(f"per-step |rel diff|: mean {rel.mean():.3e} median {np.median(rel):.3e} p90 {np.percentile(rel, 90):.3e} max {rel.max():.3e}; "
          f"mean loss A {a.mean():.6f} B {b.mean():.6f} (rel {abs(a.mean() - b.mean()) / a.mean():.2e}); "
          f"last-20 mean A {a[-20:].mean():.6f} B {b[-20:].mean():.6f} (rel {abs(a[-20:].mean() - b[-20:].mean()) / a[-20:].mean():.2e})")
    if len(sys.argv) > 4:
        json.dump({"a": a.tolist(), "b": b.tolist()}, open(sys.argv[4], "w"))


if __name__ == "__main__":
    main()
