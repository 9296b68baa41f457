"""Robustness sweep on the GPU box: unusual batch / window / grid shapes of the whole path against the fp32 oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _parity import BASELINE_KW, make_inputs, module_loss_and_grads, oracle_loss_and_grads, rel_err  # noqa: E402


def main():
    from cesm_emulator_b200.model import Diffusion, UNet
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    unet = UNet(**BASELINE_KW)
    diff = Diffusion(unet).cuda()
    shapes = [(1, 2, 24, 40), (3, 3, 40, 56), (1, 4, 16, 24), (1, 5, 16, 16), (2, 16, 16, 24), (1, 1, 32, 32), (5, 3, 16, 16),
              (1, 3, 20, 28), (1, 33, 8, 12), (2, 7, 12, 20)]
    print(f"{'shape (B,K,H,W)':18s} {'fwd':>9s} {'loss':>9s} {'grad med':>9s} {'grad max':>9s}  status")
    bad = 0
    for shape in shapes:
        try:
            x0, cond, t, noise = make_inputs(*shape, seed=11, device="cuda")
            eps, loss, grads = module_loss_and_grads(diff, x0, cond, t, noise)
            ref_eps, ref_loss, ref_grads = oracle_loss_and_grads(unet, BASELINE_KW, x0, cond, t, noise)
            v = np.array([rel_err(grads[k], ref_grads[k]) for k in ref_grads])
            e_f, e_l = rel_err(eps, ref_eps), abs(loss.item() - ref_loss.item()) / abs(ref_loss.item())
            ok = e_f < 1e-2 and e_l < 1e-2 and v.max() < 1e-2
            bad += not ok
            print(f"{str(shape):18s} {e_f:9.2e} {e_l:9.1e} {np.median(v):9.2e} {v.max():9.2e}  {'ok' if ok else 'FAIL'}", flush=True)
        except Exception as exc:  # noqa: BLE001
            bad += 1
            print(f"{str(shape):18s} ERROR {type(exc).__name__}: {str(exc)[:160]}", flush=True)
    print("failures:", bad)


if __name__ == "__main__":
    main()
