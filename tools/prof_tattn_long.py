"""One forward + backward of the long-window temporal attention at F frames on a small pixel count (for ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cesm_emulator_b200 import kernels as K  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
HW = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 2 * 8
H, D = 8, 32
torch.manual_seed(0)
qkv = (torch.randn(F * HW, 3 * H * D, device="cuda") * 0.5).half()
dout = (torch.randn(F * HW, H * D, device="cuda") * 0.1).half()
diag = torch.randn(H, 2 * F - 1, device="cuda") * 0.1
i = torch.arange(F, device="cuda")
bias = diag[:, (i[None, :] - i[:, None]) + F - 1].contiguous()
freqs = (1.0 / (10000 ** (torch.arange(0, D, 2).float() / D))).cuda()
ang = torch.arange(F, device="cuda", dtype=torch.float32)[:, None] * freqs[None]
cs, sn = ang.cos().contiguous(), ang.sin().contiguous()
for _ in range(2):
    out, lse = K.tattn_fwd(qkv, bias, cs, sn, 1, F, HW, H, D, D ** -0.5)
    K.tattn_bwd(qkv, bias, cs, sn, out, lse, dout, 1, F, HW, H, D, D ** -0.5)
torch.cuda.synchronize()
print("ok")
