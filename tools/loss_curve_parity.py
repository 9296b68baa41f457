"""200-step loss-curve parity (north_star: loss within 1 % over 200 steps), run on the GPU box.

Trains the same initial weights on the same synthetic batches with the same (t, noise) draws twice:
  * fp32 CPU oracle (oracle/cesm_oracle.py) + torch AdamW + global-norm clip  -- the reference's step
  * the B200 path (fp16 kernels through the C ABI) driven by the reference's own AMP loop (train.py:853-867:
    stock torch.amp.GradScaler + torch AdamW + clip_grad_norm_) with the same optimizer settings
and prints the per-step relative loss difference (mean / last-20 / max).

    python tools/loss_curve_parity.py [steps=200] [H=128] [W=128] [B=2]      (config/baseline's crop and batch)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _parity import BASELINE_KW  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    from cesm_emulator_b200 import ops
    from cesm_emulator_b200.model import Diffusion, UNet
    from cesm_emulator_b200.synthetic import SyntheticEnsemble
    from oracle import cesm_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    ds = SyntheticEnsemble(members=4, times=16, lat=H, lon=W, seed=7, K=3)
    g = torch.Generator().manual_seed(11)
    batches = []
    for s in range(steps):
        idx = torch.randint(0, len(ds), (B,), generator=g).tolist()
        cond, x0 = ds.batch(idx, augment=False)
        t = torch.randint(0, 1000, (B,), generator=g)
        noise = torch.randn(B, 1, H, W, generator=g)
        batches.append((cond, x0, t, noise))

    hp = dict(lr=2e-4, betas=(0.9, 0.999), weight_decay=1e-4, eps=1e-8)
    # ---- B200 path ----
    torch.manual_seed(0)
    diff = Diffusion(UNet(**BASELINE_KW)).cuda()
    diff.train()
    init = {k: v.detach().float().cpu().clone() for k, v in diff.model.state_dict().items()}
    params = [p for p in diff.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, **hp)
    scaler = torch.amp.GradScaler("cuda", init_scale=float(os.environ.get("INIT_SCALE", "65536")))
    gpu_losses = []
    for cond, x0, t, noise in batches:
        opt.zero_grad(set_to_none=True)
        loss = diff.loss(x0.cuda(), cond.cuda(), t=t.cuda(), noise=noise.cuda())
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        scaler.step(opt)
        scaler.update()
        gpu_losses.append(loss.item())
    print(f"b200 path done: final loss scale {scaler.get_scale():.0f} (65536 = no overflow / skipped step)", flush=True)

    # ---- fp32 CPU oracle ----
    cfg = O.OracleConfig.from_unet_kwargs(**BASELINE_KW)
    buf = O.diffusion_buffers(1000)
    sd = {k: v.clone() for k, v in init.items()}
    names = [k for k, v in sd.items() if v.is_floating_point() and not k.endswith("rotary_emb.freqs")]
    leaves = [sd[k].requires_grad_(True) for k in names]
    opt_c = torch.optim.AdamW(leaves, **hp)
    cpu_losses = []
    for i, (cond, x0, t, noise) in enumerate(batches):
        loss, grads = O.loss_and_grads(sd, cfg, buf, x0, cond, t, noise)
        for k, p in zip(names, leaves):
            p.grad = grads[k]
        torch.nn.utils.clip_grad_norm_(leaves, 1.0)
        opt_c.step()
        cpu_losses.append(loss.item())
        if i % 20 == 0:
            print(f"step {i:4d}  oracle {cpu_losses[-1]:.6f}  b200 {gpu_losses[i]:.6f}  "
                  f"rel {abs(gpu_losses[i] - cpu_losses[-1]) / abs(cpu_losses[-1]):.2e}", flush=True)
    rel = [abs(a - b) / abs(b) for a, b in zip(gpu_losses, cpu_losses)]
    if os.environ.get("DUMP"):
        import json
        with open(os.environ["DUMP"], "w") as f:
            json.dump({"oracle": cpu_losses, "b200": gpu_losses}, f)
    srel = sorted(rel)
    print(f"per-step |rel diff|: median {srel[len(srel) // 2]:.3e}  p90 {srel[int(0.9 * len(srel))]:.3e}; "
          f"per-step |abs diff|: mean {sum(abs(a - b) for a, b in zip(gpu_losses, cpu_losses)) / steps:.3e}; "
          f"mean loss over all steps oracle {sum(cpu_losses) / steps:.6f} b200 {sum(gpu_losses) / steps:.6f}")
    mean_c, mean_g = sum(cpu_losses[-20:]) / 20, sum(gpu_losses[-20:]) / 20
    print(f"steps {steps} grid {H}x{W} B={B}: max rel diff {max(rel):.3e}, mean rel diff {sum(rel)/len(rel):.3e}, "
          f"last-20 mean oracle {mean_c:.6f} b200 {mean_g:.6f} (rel {abs(mean_g-mean_c)/mean_c:.2e}); "
          f"loss {cpu_losses[0]:.4f} -> {cpu_losses[-1]:.4f}")
    _ = ops


if __name__ == "__main__":
    main()
