"""Print per-tensor parity of the B200 path against the fp32 CPU oracle (run on the GPU box).

    python tools/parity_report.py [B K H W]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _parity import BASELINE_KW, make_inputs, module_loss_and_grads, oracle_loss_and_grads, rel_err  # noqa: E402


def main():
    B, K, H, W = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (2, 3, 32, 32)
    from cesm_emulator_b200.model import Diffusion, UNet
    torch.manual_seed(0)
    unet = UNet(**BASELINE_KW)
    diff = Diffusion(unet).cuda()
    x0, cond, t, noise = make_inputs(B, K, H, W, seed=int(os.environ.get("SEED", "5")), device="cuda")
    eps, loss, grads = module_loss_and_grads(diff, x0, cond, t, noise)
    ref_eps, ref_loss, ref_grads = oracle_loss_and_grads(unet, BASELINE_KW, x0, cond, t, noise)
    print(f"shape B={B} K={K} {H}x{W}")
    print(f"eps   rel err {rel_err(eps, ref_eps):.3e}")
    print(f"loss  got {loss.item():.6f} ref {ref_loss.item():.6f} rel {abs(loss.item()-ref_loss.item())/abs(ref_loss.item()):.3e}")
    errs = {k: rel_err(grads[k], ref_grads[k]) for k in ref_grads}
    vals = np.array(list(errs.values()))
    print(f"grads n={len(vals)} median {np.median(vals):.3e} p90 {np.percentile(vals, 90):.3e} max {vals.max():.3e} "
          f"frac<=1e-2 {(vals <= 1e-2).mean():.3f}")
    for k in sorted(errs, key=errs.get, reverse=True)[:25]:
        print(f"  {errs[k]:.3e}  {k}  |ref|max {ref_grads[k].abs().max().item():.3e}")


if __name__ == "__main__":
    main()
