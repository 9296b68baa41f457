"""CPU model of WHERE the B200 path rounds to bf16 (test/dev infrastructure, never imported by the product).

The oracle's network (oracle/cesm_oracle.py) restated with a rounding hook `rnd(category, tensor)` at
every point where the CUDA path stores an activation (and, in backward, its gradient) as bf16.  With all
categories off it is the fp32 oracle; switching categories on one at a time gives each rounding site's
contribution to the forward and gradient error, which is how the precision plan in DESIGN.md was chosen
without spending GPU time.

Categories
  w        GEMM weights rounded to bf16 (unavoidable for a bf16 tensor-core operand)
  res      the residual stream: outputs of ResnetBlock / attention blocks / down / up / input conv
  conv     pre-GroupNorm conv outputs y1, y2
  gn       GroupNorm+FiLM+SiLU outputs (the next conv's operand)
  ln       LayerNorm outputs (the projection's operand)
  qkv      q/k/v projections
  o        attention-core outputs (the out-projection's operand)
  rc       res_conv (1x1x1) outputs before the add
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from oracle import cesm_oracle as O

ALL = ("w", "res", "conv", "gn", "ln", "qkv", "o", "rc")


_DT = {"bf16": torch.bfloat16, "fp16": torch.float16}


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x.to(_DT[fwd]).float() if fwd else x

    @staticmethod
    def backward(ctx, g):
        return (g.to(_DT[ctx.bwd]).float() if ctx.bwd else g), None, None


class Rounder:
    """on: categories rounded in forward; on_bwd: categories whose gradient is rounded (default: same)."""

    def __init__(self, on=(), on_bwd=None, split=()):
        """on / on_bwd: iterable of categories (bf16) or dict category -> "bf16" | "fp16"."""
        self.on = dict(on) if isinstance(on, dict) else {c: "bf16" for c in on}
        on_bwd = self.on if on_bwd is None else on_bwd
        self.on_bwd = dict(on_bwd) if isinstance(on_bwd, dict) else {c: "bf16" for c in on_bwd}
        self.split = set(split)   # categories kept as hi+lo bf16 pairs (error 2^-17): treated as exact

    def __call__(self, cat, x):
        if cat in self.split:
            return x
        f, b = self.on.get(cat), self.on_bwd.get(cat)
        if not (f or b):
            return x
        return _Round.apply(x, f, b)

    def w(self, t):
        return _Round.apply(t, self.on["w"], None) if "w" in self.on else t


def block(sd, pre, x, groups, rnd, scale_shift=None):
    x = F.conv3d(x, rnd.w(sd[pre + "proj.weight"]), sd[pre + "proj.bias"], padding=(0, 1, 1))
    x = rnd("conv", x)
    x = F.group_norm(x, groups, sd[pre + "norm.weight"], sd[pre + "norm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def resnet_block(sd, pre, x, temb, groups, rnd, x_op=None):
    """x: residual-stream value (may be more exact than the operand); x_op: what the convs read."""
    x_op = x if x_op is None else x_op
    scale_shift = None
    if (pre + "mlp.1.weight") in sd:
        e = F.linear(F.silu(temb), sd[pre + "mlp.1.weight"], sd[pre + "mlp.1.bias"])
        scale_shift = e[:, :, None, None, None].chunk(2, dim=1)
    h = rnd("gn", block(sd, pre + "block1.", x_op, groups, rnd, scale_shift))
    h = block(sd, pre + "block2.", h, groups, rnd)
    if (pre + "res_conv.weight") in sd:
        res = rnd("rc", F.conv3d(x_op, rnd.w(sd[pre + "res_conv.weight"]), sd[pre + "res_conv.bias"]))
    else:
        res = x
    return rnd("res", h + res)


def spatial_block(sd, pre, x, heads, rnd):
    y = rnd("ln", O.channel_layer_norm(x, sd[pre + "fn.norm.gamma"]))
    p = pre + "fn.fn."
    b, c, f, h, w = y.shape
    xf = y.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    qkv = rnd("qkv", F.conv2d(xf, rnd.w(sd[p + "to_qkv.weight"])))
    hidden = qkv.shape[1] // 3
    d = hidden // heads
    q, k, v = (t.reshape(b * f, heads, d, h * w) for t in qkv.chunk(3, dim=1))
    q = q.softmax(dim=-2) * d ** -0.5
    k = k.softmax(dim=-1)
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", context, q)
    out = rnd("o", out.reshape(b * f, hidden, h, w))
    out = F.conv2d(out, rnd.w(sd[p + "to_out.weight"]), sd[p + "to_out.bias"])
    return rnd("res", out.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4) + x)


def temporal_block(sd, pre, x, heads, pos_bias, rnd):
    b, c, f, h, w = x.shape
    y = rnd("ln", O.channel_layer_norm(x, sd[pre + "fn.norm.gamma"]))
    y = y.permute(0, 3, 4, 2, 1).reshape(b, h * w, f, c)
    p = pre + "fn.fn.fn."
    qkv = rnd("qkv", F.linear(y, rnd.w(sd[p + "to_qkv.weight"])))
    hidden = qkv.shape[-1] // 3
    d = hidden // heads

    def split(t):
        return t.reshape(*t.shape[:-1], heads, d).transpose(-2, -3)

    q, k, v = (split(t) for t in qkv.chunk(3, dim=-1))
    q = q * d ** -0.5
    ang = O.rotary_angles(sd[p + "rotary_emb.freqs"], f)
    q, k = O.apply_rotary(q, ang), O.apply_rotary(k, ang)
    sim = torch.einsum("...hid,...hjd->...hij", q, k) + pos_bias
    attn = sim.softmax(dim=-1)
    out = torch.einsum("...hij,...hjd->...hid", attn, v)
    out = rnd("o", out.transpose(-2, -3).reshape(b, h * w, f, hidden))
    y = F.linear(out, rnd.w(sd[p + "to_out.weight"]))
    y = y.reshape(b, h, w, f, c).permute(0, 4, 3, 1, 2)
    return rnd("res", y + x)


def unet3d_forward(sd, cfg, x, timesteps, cond_map, rnd, pre="net."):
    heads, groups = cfg.attn_heads, cfg.resnet_groups
    pos_bias = O.rel_pos_bias(sd[pre + "time_rel_pos_bias.relative_attention_bias.weight"], x.shape[2], cfg)
    x = torch.cat([x, cond_map], dim=1)
    pad = cfg.init_kernel_size // 2
    x = rnd("res", F.conv3d(x, sd[pre + "input_conv.weight"], sd[pre + "input_conv.bias"], padding=(0, pad, pad)))
    x = temporal_block(sd, pre + "input_temp_op.", x, heads, pos_bias, rnd)
    r = x
    t = O.sinusoidal_pos_emb(timesteps, cfg.model_dim)
    t = F.linear(t, sd[pre + "time_mlp.1.weight"], sd[pre + "time_mlp.1.bias"])
    t = F.linear(F.silu(t), sd[pre + "time_mlp.3.weight"], sd[pre + "time_mlp.3.bias"])
    n_levels = len(cfg.dim_mults)
    skips = []
    for lvl in range(n_levels):
        p = f"{pre}downs.{lvl}."
        x = resnet_block(sd, p + "0.", x, t, groups, rnd)
        x = resnet_block(sd, p + "1.", x, t, groups, rnd)
        if (p + "2.fn.norm.gamma") in sd:
            x = spatial_block(sd, p + "2.", x, heads, rnd)
        x = temporal_block(sd, p + "3.", x, heads, pos_bias, rnd)
        skips.append(x)
        if (p + "4.weight") in sd:
            x = rnd("res", F.conv3d(x, rnd.w(sd[p + "4.weight"]), sd[p + "4.bias"], stride=(1, 2, 2), padding=(0, 1, 1)))
    x = resnet_block(sd, pre + "mid_block1.", x, t, groups, rnd)
    x = temporal_block(sd, pre + "mid_temporal_attn.", x, heads, pos_bias, rnd)
    x = resnet_block(sd, pre + "mid_block2.", x, t, groups, rnd)
    for lvl in range(n_levels):
        p = f"{pre}ups.{lvl}."
        x = torch.cat((x, skips.pop()), dim=1)
        x = resnet_block(sd, p + "0.", x, t, groups, rnd)
        x = resnet_block(sd, p + "1.", x, t, groups, rnd)
        if (p + "2.fn.norm.gamma") in sd:
            x = spatial_block(sd, p + "2.", x, heads, rnd)
        x = temporal_block(sd, p + "3.", x, heads, pos_bias, rnd)
        if (p + "4.weight") in sd:
            x = rnd("res", F.conv_transpose3d(x, rnd.w(sd[p + "4.weight"]), sd[p + "4.bias"], stride=(1, 2, 2),
                                              padding=(0, 1, 1)))
    x = torch.cat((x, r), dim=1)
    x = resnet_block(sd, pre + "out_conv.0.", x, None, groups, rnd)
    return F.conv3d(x, sd[pre + "out_conv.1.weight"], sd[pre + "out_conv.1.bias"])


def loss_and_grads(sd, cfg, buf, x0, cond, t, noise, rnd):
    leaves = {}
    for k, v in sd.items():
        v = v.detach().clone()
        if v.is_floating_point() and not k.endswith("rotary_emb.freqs"):
            v.requires_grad_(True)
        leaves[k] = v
    x_t = O.q_sample(buf, x0, t, noise).unsqueeze(2).expand(-1, -1, cond.shape[2], -1, -1)
    out = unet3d_forward(leaves, cfg, x_t, t, cond, rnd)
    eps = out[:, :, out.shape[2] // 2]
    loss = F.mse_loss(eps, noise)
    names = [k for k, v in leaves.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return eps.detach(), loss.detach(), {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, grads)}
