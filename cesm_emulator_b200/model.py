"""2-D facing wrapper and DDPM maths of the CESM emulator (mirror of the reference's model.py).

`UNet` and `Diffusion` keep the reference's constructor signatures, attribute names
(`UNet.net`, `Diffusion.model`, `.T`, the eight schedule buffers) and call surface
(`forward(x_t, cond, t)`, `loss`, `q_sample`, `p_sample`, `sample`), so train.py / inference.py
style drivers and reference checkpoints work unchanged.  All tensor work runs in the sm_100a
kernels of libcesm_b200.so; there is no CPU path.
"""
from __future__ import annotations

import torch
from torch import nn

from . import kernels as K
from . import ops
from .video_net import UNetModel3D


class UNet(nn.Module):
    """model.py:37-134.  `in_channels`, `num_res_blocks`, `time_dim` and `dropout` are accepted
    and ignored exactly as in the reference (model.py:46-52)."""

    def __init__(self, in_channels: int = 2, out_channels: int = 1, base_ch: int = 64, ch_mults=(1, 2, 4),
                 num_res_blocks: int = 2, time_dim: int = 256, groups: int = 8, dropout: float = 0.0,
                 attn_heads: int = 8, attn_dim_head: int = 32, use_sparse_linear_attn: bool = True,
                 use_mid_attn: bool = False, init_kernel_size: int = 7, use_checkpoint: bool = False,
                 use_temp_attn: bool = True, day_cond: bool = False, year_cond: bool = False,
                 cond_map: bool = True):
        super().__init__()
        self.net = UNetModel3D(
            n_vars=out_channels, model_dim=base_ch, dim_mults=tuple(ch_mults), attn_heads=attn_heads,
            attn_dim_head=attn_dim_head, use_sparse_linear_attn=use_sparse_linear_attn, use_mid_attn=use_mid_attn,
            init_kernel_size=init_kernel_size, resnet_groups=groups, use_checkpoint=use_checkpoint,
            use_temp_attn=use_temp_attn, day_cond=day_cond, year_cond=year_cond, cond_map=cond_map)

    def forward(self, x_t, cond, t):
        """x_t: [B,1,H,W] or [B,1,F,H,W]; cond: [B,1,H,W] or [B,1,F,H,W]; t: [B] -> [B,1,H,W].

        The reference expands the single-frame side to F frames, runs the 3-D net on all frames
        and keeps frame F//2 (model.py:110-130).  Here the broadcast happens inside the input-conv
        kernel and the 1x1x1 output conv is evaluated for the centre frame only."""
        if x_t.ndim == 4:
            x_t = x_t.unsqueeze(2)
        elif x_t.ndim != 5:
            raise ValueError(f"x_t must be 4D or 5D, got {x_t.ndim}D")
        if cond is None:
            raise ValueError("cond must be provided")
        if cond.ndim == 4:
            cond = cond.unsqueeze(2)
        elif cond.ndim != 5:
            raise ValueError(f"cond must be 4D or 5D, got {cond.ndim}D")
        Fx, Fc = x_t.shape[2], cond.shape[2]
        if Fx != Fc and not (Fx == 1 or Fc == 1):
            raise ValueError(f"Frame mismatch: x_t F={Fx}, cond F={Fc}")
        F = max(Fx, Fc)
        out = self.net.forward_frames(x_t, t, cond, frames=[F // 2])
        return out.squeeze(2)


class Diffusion(nn.Module):
    """model.py:141-208."""

    def __init__(self, model, img_channels=1, timesteps=1000, beta_schedule="linear"):
        super().__init__()
        self.model = model
        self.img_channels = img_channels
        self.T = timesteps
        if beta_schedule == "linear":
            beta_start, beta_end = 1e-4, 2e-2
            betas = torch.linspace(beta_start, beta_end, timesteps)
        else:
            raise ValueError("Only 'linear' beta_schedule implemented")
        alphas = 1.0 - betas
        alphas_cumprod = torch.cumprod(alphas, dim=0)
        alphas_cumprod_prev = torch.cat([torch.tensor([1.0]), alphas_cumprod[:-1]], dim=0)
        self.register_buffer("betas", betas)
        self.register_buffer("alphas", alphas)
        self.register_buffer("alphas_cumprod", alphas_cumprod)
        self.register_buffer("alphas_cumprod_prev", alphas_cumprod_prev)
        self.register_buffer("sqrt_alphas_cumprod", torch.sqrt(alphas_cumprod))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", torch.sqrt(1.0 - alphas_cumprod))
        self.register_buffer("sqrt_recip_alphas", torch.sqrt(1.0 / alphas))
        self.register_buffer("posterior_variance", betas * (1.0 - alphas_cumprod_prev) / (1.0 - alphas_cumprod))

    # -- sampling ------------------------------------------------------------------------------
    def p_sample(self, x_t, cond, t, noise=None):
        """model.py:168-183.  One reverse step.  posterior_variance[0] == 0, so the reference's
        `(t == 0).all()` branch (a host sync per step) is the same arithmetic as always adding
        sqrt(var_t) * z; the fused kernel does that without the sync.  `noise` may be supplied
        for deterministic parity tests."""
        with torch.no_grad():
            eps_theta = self.model(x_t, cond, t)
            z = torch.randn_like(x_t) if noise is None else noise
            return K.p_sample(x_t.float().contiguous(), eps_theta, z.float().contiguous(), t.contiguous(),
                              self.betas, self.sqrt_one_minus_alphas_cumprod, self.sqrt_recip_alphas,
                              self.posterior_variance)

    def sample(self, cond, shape, device):
        """model.py:186-194."""
        with torch.no_grad():
            B = shape[0]
            x = torch.randn(shape, device=device)
            for tt in reversed(range(self.T)):
                t_tensor = torch.full((B,), tt, device=device, dtype=torch.long)
                x = self.p_sample(x, cond, t_tensor)
            return x

    # -- training ------------------------------------------------------------------------------
    def q_sample(self, x0, t, noise=None):
        """model.py:196-201."""
        if noise is None:
            noise = torch.randn_like(x0)
        x_t = K.q_sample(x0.float().contiguous(), noise.float().contiguous(), t.contiguous(),
                         self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod)
        return x_t, noise

    def loss(self, x0, cond, t=None, noise=None):
        """model.py:203-208.  `t` / `noise` default to the reference's draws (randint, then
        randn_like); passing them makes parity tests deterministic."""
        B = x0.size(0)
        if t is None:
            t = torch.randint(0, self.T, (B,), device=x0.device).long()
        x_t, noise = self.q_sample(x0, t, noise)
        eps_pred = self.model(x_t, cond, t)
        return ops.MseLossFn.apply(eps_pred, noise.float())
