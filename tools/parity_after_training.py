"""Parity of the B200 path at TRAINED weights (not just at initialisation): train N steps with the reference's AMP
loop on the GPU, then compare one step (forward, loss, every gradient) with the fp32 oracle on the same weights
for several (t, noise) draws, with the loss scale the GradScaler holds at that point.

    python tools/parity_after_training.py [steps=80] [H=64] [W=64]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _parity import BASELINE_KW, module_loss_and_grads, oracle_loss_and_grads, rel_err  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 80
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    B = 2
    from cesm_emulator_b200.model import Diffusion, UNet
    from cesm_emulator_b200.synthetic import SyntheticEnsemble
    torch.set_num_threads(os.cpu_count() or 1)
    ds = SyntheticEnsemble(members=4, times=16, lat=H, lon=W, seed=7, K=3)
    g = torch.Generator().manual_seed(11)
    torch.manual_seed(0)
    diff = Diffusion(UNet(**BASELINE_KW)).cuda()
    diff.train()
    params = [p for p in diff.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=2e-4, weight_decay=1e-4)
    scaler = torch.amp.GradScaler("cuda")

    def report(tag):
        for tt in ([5, 40], [300, 700], [950, 999]):
            cond, x0 = ds.batch(torch.randint(0, len(ds), (B,), generator=g).tolist(), augment=False)
            t = torch.tensor(tt)
            noise = torch.randn(B, 1, H, W, generator=g)
            args = [v.cuda() for v in (x0, cond, t, noise)]
            eps, loss, grads = module_loss_and_grads(diff, *args, loss_scale=scaler.get_scale() if scaler._scale is not None else 65536.0)
            ref_eps, ref_loss, ref_grads = oracle_loss_and_grads(diff.model, BASELINE_KW, *args)
            errs = {k: rel_err(grads[k], ref_grads[k]) for k in ref_grads}
            v = np.array(list(errs.values()))
            w = max(errs, key=errs.get)
            print(f"{tag} t={tt}: loss {ref_loss.item():.5f} (rel {abs(loss.item()-ref_loss.item())/ref_loss.item():.1e}) "
                  f"fwd {rel_err(eps, ref_eps):.2e} grad med {np.median(v):.2e} p90 {np.percentile(v, 90):.2e} "
                  f"max {v.max():.2e} ({w}) |eps| {ref_eps.abs().max():.2f}", flush=True)

    report("init")
    for s in range(steps):
        cond, x0 = ds.batch(torch.randint(0, len(ds), (B,), generator=g).tolist(), augment=False)
        t = torch.randint(0, 1000, (B,), generator=g)
        noise = torch.randn(B, 1, H, W, generator=g)
        opt.zero_grad(set_to_none=True)
        loss = diff.loss(x0.cuda(), cond.cuda(), t=t.cuda(), noise=noise.cuda())
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        scaler.step(opt)
        scaler.update()
        if (s + 1) % 40 == 0:
            report(f"after {s + 1} steps")


if __name__ == "__main__":
    main()
