"""Parity matrix on the GPU box: shapes x input seeds x reruns of the B200 path against the fp32 CPU oracle.

    python tools/parity_matrix.py [--quick] > profiles/r02_parity_matrix.txt

Per case: forward eps error, loss error, and over the 225 parameter-gradient tensors the median / p90 / max of
max|got-ref|/max|ref| and the fraction within 1e-2 (north_star's per-tensor bar).  The oracle runs once per
(shape, seed); the CUDA path is re-run `reruns` times on the same inputs (fp32 atomics reorder sums, so reruns are
not bitwise equal -- the spread is what is reported).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _parity import BASELINE_KW, make_inputs, module_loss_and_grads, oracle_loss_and_grads, rel_err  # noqa: E402


def main():
    quick = "--quick" in sys.argv
    from cesm_emulator_b200.model import Diffusion, UNet
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    unet = UNet(**BASELINE_KW)
    diff = Diffusion(unet).cuda()
    shapes = [(2, 3, 128, 128), (1, 3, 48, 72), (1, 3, 192, 288)]
    seeds = (5,) if quick else (5, 6, 7)
    reruns = 1 if quick else 3
    print(f"# config/baseline architecture, weights torch.manual_seed(0); metric max|got-ref|/max|ref| per tensor; "
          f"host cores {os.cpu_count()}")
    print(f"{'shape (B,K,H,W)':18s} {'seed':>4s} {'run':>3s} {'fwd eps':>9s} {'loss':>9s} {'grad med':>9s} {'grad p90':>9s} "
          f"{'grad max':>9s} {'<=1e-2':>7s}  worst tensor")
    worst_fwd = worst_grad = 0.0
    for shape in shapes:
        for seed in seeds:
            x0, cond, t, noise = make_inputs(*shape, seed=seed, device="cuda")
            t0 = time.time()
            ref_eps, ref_loss, ref_grads = oracle_loss_and_grads(unet, BASELINE_KW, x0, cond, t, noise)
            dt = time.time() - t0
            for run in range(reruns):
                eps, loss, grads = module_loss_and_grads(diff, x0, cond, t, noise)
                e_f = rel_err(eps, ref_eps)
                e_l = abs(loss.item() - ref_loss.item()) / abs(ref_loss.item())
                errs = {k: rel_err(grads[k], ref_grads[k]) for k in ref_grads}
                v = np.array(list(errs.values()))
                w = max(errs, key=errs.get)
                worst_fwd, worst_grad = max(worst_fwd, e_f), max(worst_grad, v.max())
                print(f"{str(shape):18s} {seed:4d} {run:3d} {e_f:9.2e} {e_l:9.1e} {np.median(v):9.2e} {np.percentile(v, 90):9.2e} "
                      f"{v.max():9.2e} {(v <= 1e-2).mean():7.3f}  {w}", flush=True)
            print(f"#   oracle fwd+bwd on the CPU took {dt:.1f} s", flush=True)
    print(f"# worst forward error {worst_fwd:.2e}, worst gradient tensor {worst_grad:.2e} (bars: 1e-2 / 1e-2)")


if __name__ == "__main__":
    main()
