"""CUDA-event timing of the attention cores at the bench shapes (run on the GPU box).
Inputs at L0 are 510 MB (> 126 MB L2), so consecutive reps do not hit in L2."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cesm_emulator_b200 import kernels as K


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    H, D = 8, 32
    B, F = 2, 3
    for name, n in (("L0", 192 * 288), ("L1", 96 * 144), ("L2", 48 * 72)):
        NI = B * F
        torch.manual_seed(0)
        qkv = torch.randn(NI * n, 3 * H * D, device="cuda").half()
        dout = torch.randn(NI * n, H * D, device="cuda").half()
        bias = torch.randn(H, F, F, device="cuda")
        freqs = (1.0 / (10000 ** (torch.arange(0, D, 2).float() / D))).cuda()
        ang = torch.arange(F, device="cuda", dtype=torch.float32)[:, None] * freqs[None]
        cs, sn = ang.cos().contiguous(), ang.sin().contiguous()
        rows = NI * n
        out, ws = K.linattn_fwd(qkv, NI, n, H, D, D ** -0.5)
        t = timeit(lambda: K.linattn_fwd(qkv, NI, n, H, D, D ** -0.5))
        print(f"{name} linattn_fwd  {t:8.1f} us  {rows * 2048 / t / 1e3:7.0f} GB/s (alg 2 KB/row: q,k,v once + out)")
        t = timeit(lambda: K.linattn_bwd(qkv, ws, dout, NI, n, H, D, D ** -0.5))
        print(f"{name} linattn_bwd  {t:8.1f} us  {rows * 3584 / t / 1e3:7.0f} GB/s (alg 3.5 KB/row: qkv + dout in, dqkv out)")
        t = timeit(lambda: K.tattn_fwd(qkv, bias, cs, sn, B, F, n, H, D, D ** -0.5))
        print(f"{name} tattn_fwd    {t:8.1f} us  {rows * 2048 / t / 1e3:7.0f} GB/s")
        t = timeit(lambda: K.tattn_bwd(qkv, bias, cs, sn, None, None, dout, B, F, n, H, D, D ** -0.5))
        print(f"{name} tattn_bwd    {t:8.1f} us  {rows * 3584 / t / 1e3:7.0f} GB/s")


if __name__ == "__main__":
    main()
