"""Time / profile the attention cores at the L0 shape of the full-grid bench (run on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cesm_emulator_b200 import kernels as K

def main():
    NI, n, H, D = 6, 192 * 288, 8, 32
    B, F = 2, 3
    torch.manual_seed(0)
    qkv = torch.randn(NI * n, 3 * H * D, device="cuda").half()
    dout = torch.randn(NI * n, H * D, device="cuda").half()
    bias = torch.randn(H, F, F, device="cuda")
    freqs = (1.0 / (10000 ** (torch.arange(0, D, 2).float() / D))).cuda()
    ang = torch.arange(F, device="cuda", dtype=torch.float32)[:, None] * freqs[None]
    cs, sn = ang.cos().contiguous(), ang.sin().contiguous()
    reps = int(os.environ.get("REPS", "5"))
    for _ in range(reps):
        out, ws = K.linattn_fwd(qkv, NI, n, H, D, D ** -0.5)
        K.linattn_bwd(qkv, ws, dout, NI, n, H, D, D ** -0.5)
        o2, _ = K.tattn_fwd(qkv, bias, cs, sn, B, F, n, H, D, D ** -0.5)
        K.tattn_bwd(qkv, bias, cs, sn, None, None, dout, B, F, n, H, D, D ** -0.5)
    torch.cuda.synchronize()

if __name__ == "__main__":
    main()
