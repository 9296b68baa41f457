"""Fused q/k/v-projection backward (cesm_qkv_bwd) against the two separate kernels (run on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cesm_emulator_b200 import kernels as K
from tools.bench_igemm import timeit

for name, n, h, w in (("L0", 6, 192, 288), ("L1", 6, 96, 144), ("L2", 6, 48, 72)):
    if name != "L0":
        break  # C = 64 only at L0 (the fused kernel is specialised for 64 input channels)
    rows, cout = n * h * w, 768
    dy = torch.randn(n, h, w, cout, device="cuda").half()
    x = torch.randn(n, h, w, 64, device="cuda").half()
    wt = (torch.randn(64, cout, device="cuda") * 0.1).half()
    into = torch.zeros(cout, 1, 64, device="cuda")
    acc = torch.zeros(cout, 64, device="cuda")
    t_d = timeit(lambda: K.igemm(dy, wt))
    t_w = timeit(lambda: K.wgrad(x, dy, into=into, layout=(64, 1, [0])))
    t_f = timeit(lambda: K.qkv_bwd(dy, x, wt, dw_into=acc))
    by = 2.0 * rows * (cout + 128)
    print(f"{name} dgrad {t_d*1e3:.1f} us + wgrad {t_w*1e3:.1f} us = {(t_d+t_w)*1e3:.1f} us | fused {t_f*1e3:.1f} us "
          f"({by/t_f/1e6:.0f} GB/s, {4.0*rows*cout*64/t_f/1e9:.0f} TFLOP/s)")
