"""Compare gradients: engine (direct delivery) vs plain autograd vs fp32 CPU oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from _parity import BASELINE_KW, rel_err, oracle_loss_and_grads
from cesm_emulator_b200 import ops
from cesm_emulator_b200.engine import TrainEngine
from cesm_emulator_b200.model import Diffusion, UNet

B, K, H, W = 2, 3, 32, 48
torch.manual_seed(0)
d = Diffusion(UNet(**BASELINE_KW)).cuda(); d.train()
g = torch.Generator().manual_seed(3)
x0, cond = torch.randn(B, 1, H, W, generator=g).cuda(), torch.randn(B, 1, K, H, W, generator=g).cuda()
t = torch.tensor([100, 700], device="cuda"); noise = torch.randn(B, 1, H, W, generator=g).cuda()

def run_autograd():
    ops.set_grad_sink(None)
    d.zero_grad(set_to_none=True)
    loss = d.loss(x0, cond, t=t, noise=noise); loss.backward()
    return loss.item(), {k: p.grad.clone() for k, p in d.named_parameters() if p.grad is not None}

l1, a1 = run_autograd()
l2, a2 = run_autograd()
eng = TrainEngine(d, (B, 1, H, W), (B, 1, K, H, W), lr=0.0, weight_decay=0.0, max_grad_norm=None, use_graph=False)
orig_loss = d.loss
d.loss = lambda x, c: orig_loss(x, c, t=t, noise=noise)
le = eng.step(x0, cond).item()
e1 = {k: p.grad.clone() for k, p in d.named_parameters() if p.requires_grad}
_, lo, og = oracle_loss_and_grads(d.model, BASELINE_KW, x0, cond, t, noise)
og = {"model." + k: v for k, v in og.items()}
print("loss autograd", l1, l2, "engine", le, "oracle", lo.item())
def stats(name, A, Bd):
    errs = {k: rel_err(A[k], Bd[k]) for k in Bd if k in A}
    v = np.array(list(errs.values())); w = max(errs, key=errs.get)
    print(f"{name:28s} median {np.median(v):.2e} p90 {np.percentile(v,90):.2e} max {v.max():.2e} ({w})")
stats("autograd run1 vs run2", a1, a2)
stats("engine vs autograd", e1, a1)
stats("autograd vs oracle", a1, og)
stats("engine vs oracle", e1, og)
