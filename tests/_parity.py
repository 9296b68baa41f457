"""Shared helpers for the model-level parity tests and the parity report tool.

Error metric (the one north_star's tolerance is stated in): per tensor,
    max|got - ref| / max|ref|.
"""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")

BASELINE_KW = dict(in_channels=2, out_channels=1, base_ch=64, ch_mults=(1, 2, 4), num_res_blocks=2, time_dim=124,
                   groups=8, dropout=0.0, use_checkpoint=False)


def load_golden(name):
    z = np.load(os.path.join(GOLD, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def rel_err(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return ((got - ref).abs().max() / (ref.abs().max() + 1e-30)).item()


def make_inputs(B, K, H, W, seed, T=1000, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    x0 = torch.randn(B, 1, H, W, generator=g)
    cond = torch.randn(B, 1, K, H, W, generator=g)
    t = torch.randint(0, T, (B,), generator=g)
    noise = torch.randn(B, 1, H, W, generator=g)
    return tuple(v.to(device) for v in (x0, cond, t, noise))


def oracle_loss_and_grads(unet, unet_kwargs, x0, cond, t, noise, T=1000):
    """fp32 CPU oracle on the module's own weights -> (eps, loss, {param name: grad})."""
    from oracle import cesm_oracle as O
    cfg = O.OracleConfig.from_unet_kwargs(**unet_kwargs)
    sd = {k: v.detach().float().cpu() for k, v in unet.state_dict().items()}
    buf = O.diffusion_buffers(T)
    x0, cond, t, noise = (v.cpu() for v in (x0, cond, t, noise))
    with torch.no_grad():
        eps = O.unet_forward(sd, cfg, O.q_sample(buf, x0, t, noise), cond, t)
    loss, grads = O.loss_and_grads(sd, cfg, buf, x0, cond, t, noise)
    return eps, loss, grads


# The activations and the gradient stream are fp16 (the reference's autocast dtype, train.py:853), so -- exactly as
# in the reference's step (train.py:862-864: scaler.scale(loss).backward(); scaler.unscale_(opt)) -- the loss is
# scaled before backward and the fp32 parameter gradients are unscaled afterwards.  65536 is GradScaler's
# initial scale.
LOSS_SCALE = 65536.0


def module_loss_and_grads(diffusion, x0, cond, t, noise, loss_scale=LOSS_SCALE):
    diffusion.zero_grad(set_to_none=True)
    x_t, _ = diffusion.q_sample(x0, t, noise)
    eps = diffusion.model(x_t, cond, t)
    from cesm_emulator_b200 import ops
    loss = ops.MseLossFn.apply(eps, noise)
    (loss * loss_scale).backward()
    grads = {k: p.grad / loss_scale for k, p in diffusion.model.named_parameters() if p.grad is not None}
    return eps.detach(), loss.detach(), grads
