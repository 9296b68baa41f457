"""CUDA-event timing of the GroupNorm / LayerNorm kernels at the bench shapes (run on the GPU box).
PLAIN=1: three plain calls of each (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cesm_emulator_b200 import kernels as K
from tools.time_attn import timeit


def main():
    B, F, G = 2, 3, 8
    plain = os.environ.get("PLAIN")
    for name, hw, C in (("L0 C64", 192 * 288, 64), ("L0 C128", 192 * 288, 128), ("L1 C128", 96 * 144, 128), ("L2 C256", 48 * 72, 256)):
        P = F * hw
        torch.manual_seed(0)
        x = torch.randn(B * P, C, device="cuda").half()
        dout = torch.randn(B * P, C, device="cuda").half()
        res = torch.randn(B * P, C, device="cuda").half()
        gamma = torch.randn(C, device="cuda"); beta = torch.randn(C, device="cuda")
        film = torch.randn(B, 2 * C, device="cuda") * 0.1
        lg = torch.randn(C, device="cuda")
        sums = K.gn_stats(x, B, G)
        fns = {
            "gn_stats": (lambda: K.gn_stats(x, B, G), 2),
            "gn_apply_fwd": (lambda: K.gn_apply_fwd(x, sums, gamma, beta, film, None, B, G, 1e-5), 4),
            "gn_apply_fwd+res": (lambda: K.gn_apply_fwd(x, sums, gamma, beta, film, res, B, G, 1e-5), 6),
            "gn_bwd": (lambda: K.gn_bwd(x, dout, sums, gamma, beta, film, B, G, 1e-5, conv_bias_grad=True), 10),
            "ln_fwd": (lambda: K.ln_fwd(x, lg, 1e-5), 4),
            "ln_bwd": (lambda: K.ln_bwd(x, lg, dout, res, 1e-5), 8),
        }
        for k, (fn, bpe) in fns.items():
            if plain:
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                continue
            t = timeit(fn, reps=20)
            print(f"{name:8s} {k:18s} {t:8.1f} us  {x.numel() * bpe / t / 1e3:7.0f} GB/s ({bpe} B/elt)")
        if plain:
            break


if __name__ == "__main__":
    main()
