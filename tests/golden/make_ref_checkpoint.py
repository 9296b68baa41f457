"""Write a checkpoint exactly the way the reference's train.py does (train.py:1154-1165, DDP branch) from the
UNMODIFIED reference classes, plus the reference's outputs on fixed inputs:

    python tests/golden/make_ref_checkpoint.py     ->  tests/golden/ref_ckpt_tiny.pt, ref_ckpt_tiny_expect.npz

The model is a small instance of the reference architecture (base_ch 8, two levels) so that the file stays small;
every module type of config/baseline is present.  The optimizer went through two real AdamW steps of the
reference's loop body so that its state dict carries real moments.  Run in the build container only
(/root/reference is not on the GPU box).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

KW = dict(in_channels=2, out_channels=1, base_ch=8, ch_mults=(1, 2), num_res_blocks=2, time_dim=32, groups=8,
          dropout=0.0, use_checkpoint=False)


def main():
    ref_model, _ = import_reference()
    torch.manual_seed(3)
    unet = ref_model.UNet(**KW)
    diffusion = ref_model.Diffusion(unet, timesteps=1000, beta_schedule="linear")
    diffusion.train()
    optimizer = torch.optim.AdamW(diffusion.parameters(), lr=2e-4, betas=(0.9, 0.999), weight_decay=1e-4)
    g = torch.Generator().manual_seed(4)
    B, K, H, W = 2, 3, 16, 16
    for _ in range(2):  # train.py:858-867 without AMP (CPU)
        x0, cond = torch.randn(B, 1, H, W, generator=g), torch.randn(B, 1, K, H, W, generator=g)
        optimizer.zero_grad(set_to_none=True)
        loss = diffusion.loss(x0, cond)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(diffusion.parameters(), 1.0)
        optimizer.step()
    cfg = {"unet": {**KW, "ch_mults": list(KW["ch_mults"])}, "diffusion": {"timesteps": 1000, "beta_schedule": "linear"},
           "dataset": {"K": K}, "train": {"batch_size": B}}
    # ---- train.py:1154-1165 verbatim semantics ----
    full_sd = diffusion.state_dict()
    unet_sd = {k[len("model."):]: v for k, v in full_sd.items() if k.startswith("model.")}
    buffers = {k: v for k, v in full_sd.items() if not k.startswith("model.")}
    ckpt = {"epoch": 7, "model": unet_sd, "diffusion_buffers": buffers, "optimizer": optimizer.state_dict(), "config": cfg}
    torch.save(ckpt, os.path.join(HERE, "ref_ckpt_tiny.pt"))
    # ---- what the reference computes with these weights ----
    diffusion.eval()
    x_t, cond = torch.randn(B, 1, H, W, generator=g), torch.randn(B, 1, K, H, W, generator=g)
    t = torch.tensor([17, 803])
    with torch.no_grad():
        eps = unet(x_t, cond, t)
        eps_f1 = unet(x_t, cond[:, :, 1], t)
    np.savez(os.path.join(HERE, "ref_ckpt_tiny_expect.npz"), x_t=x_t.numpy(), cond=cond.numpy(), t=t.numpy(),
             eps=eps.numpy(), eps_f1=eps_f1.numpy(),
             exp_avg_0=optimizer.state_dict()["state"][0]["exp_avg"].numpy(),
             n_params=np.array(len(optimizer.state_dict()["state"])))
    print("wrote", os.path.getsize(os.path.join(HERE, "ref_ckpt_tiny.pt")), "bytes")


if __name__ == "__main__":
    main()
