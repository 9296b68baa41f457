"""Small-shape pass over the tcgen05 / TMA / mbarrier kernels for compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py

Covers igemm2 (halo 3x3 + fused GroupNorm statistics, streamed and resident weights, strided / phase
variants), wgrad (3x3 and 1x1, split-K red.add), the fused q/k/v backward, the attention and normalisation
kernels and one optimizer step: every mbarrier protocol, TMEM alloc/dealloc and red.add path of the library.
Records: profiles/r02_compute_sanitizer_*.txt.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from _parity import BASELINE_KW, LOSS_SCALE
    from cesm_emulator_b200 import kernels as K, ops
    from cesm_emulator_b200.engine import TrainEngine
    from cesm_emulator_b200.model import Diffusion, UNet
    dev = "cuda"
    torch.manual_seed(0)
    h16 = torch.float16
    # igemm2: halo conv with GN sums; 1x1 projection (resident weights); wide cout (streamed weights)
    x = torch.randn(2, 16, 16, 64, device=dev).to(h16)
    w9 = (torch.randn(64, 9 * 64, device=dev) * 0.05).to(h16)
    sums = torch.zeros(1, 8, 2, device=dev)
    K.igemm(x, w9, taps=K.TAPS_3x3, bias=torch.zeros(64, device=dev), gn_sums=sums, gn_frames=2)
    w1 = (torch.randn(768, 64, device=dev) * 0.05).to(h16)
    qkv = K.igemm(x, w1)
    x5 = torch.randn(2, 8, 8, 512, device=dev).to(h16)
    K.igemm(x5, (torch.randn(512, 9 * 512, device=dev) * 0.02).to(h16), taps=K.TAPS_3x3)
    # weight gradients
    dy = torch.randn(2, 16, 16, 64, device=dev).to(h16)
    K.wgrad(x, dy, taps=K.TAPS_3x3)
    K.wgrad(x, qkv)
    # fused q/k/v backward
    K.qkv_bwd(qkv, x, w1.t().contiguous())
    torch.cuda.synchronize()
    # the whole step at a tiny grid (every kernel of the library, engine mode, eager)
    d = Diffusion(UNet(**BASELINE_KW)).to(dev)
    d.train()
    eng = TrainEngine(d, (1, 1, 16, 16), (1, 1, 3, 16, 16), use_graph=False)
    for _ in range(2):
        loss = eng.step(torch.randn(1, 1, 16, 16), torch.randn(1, 1, 3, 16, 16))
    torch.cuda.synchronize()
    print("sanitize_small: loss", float(loss), "loss scale", float(eng.opt.loss_scale), "(expected", LOSS_SCALE, ")")
    ops.set_grad_sink(None)


if __name__ == "__main__":
    main()
