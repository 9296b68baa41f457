"""CPU oracle for the CESM-emulator hot path (TEST INFRASTRUCTURE -- never imported by the product).

A functional fp32 restatement of the reference's space-time U-Net + DDPM wrapper
(/root/reference/video_net.py, model.py, rotary_embedding.py).  Every function takes the
reference's own `state_dict` (same key names and tensor layouts) plus plain tensors and cites
the reference lines it follows.  It is written with torch fp32 CPU ops because the path is
floating point; nothing here touches CUDA kernels of this repo.

Parity status: PINNED.  tests/golden/*.npz hold outputs of the unmodified reference modules
(imported from /root/reference with an `einops_exts` shim) for fixed seeds; the generating
script is tests/golden/make_golden.py and tests/test_oracle_golden.py checks this file against
them (forward, loss, every parameter gradient, the sampling chain), plus the known-answer
vectors of SURVEY.md section 4.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module, and only as the checker or the timed CPU baseline.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


@dataclass
class OracleConfig:
    """Constructor arguments that shape the network (model.py:44-64, video_net.py:562-578)."""
    n_vars: int = 1
    model_dim: int = 64
    dim_mults: Sequence[int] = (1, 2, 4)
    attn_heads: int = 8
    attn_dim_head: int = 32
    use_sparse_linear_attn: bool = True
    use_mid_attn: bool = False
    init_kernel_size: int = 7
    resnet_groups: int = 8
    cond_map: bool = True
    rel_pos_num_buckets: int = 32     # RelativePositionBias default (video_net.py:269)
    rel_pos_max_distance: int = 32    # UNetModel3D passes max_distance=32 (video_net.py:630-632)
    spatial_dim_head: int = 32        # SpatialLinearAttention default dim_head (video_net.py:314)

    @property
    def dims(self) -> List[int]:
        return [self.model_dim, *[int(self.model_dim * m) for m in self.dim_mults]]  # video_net.py:646

    @classmethod
    def from_unet_kwargs(cls, **kw) -> "OracleConfig":
        """Map model.UNet kwargs (model.py:44-83) to the 3-D network's arguments."""
        return cls(
            n_vars=kw.get("out_channels", 1), model_dim=kw.get("base_ch", 64),
            dim_mults=tuple(kw.get("ch_mults", (1, 2, 4))), attn_heads=kw.get("attn_heads", 8),
            attn_dim_head=kw.get("attn_dim_head", 32),
            use_sparse_linear_attn=kw.get("use_sparse_linear_attn", True),
            use_mid_attn=kw.get("use_mid_attn", False), init_kernel_size=kw.get("init_kernel_size", 7),
            resnet_groups=kw.get("groups", 8), cond_map=kw.get("cond_map", True))


# --------------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------------
def sinusoidal_pos_emb(t: Tensor, dim: int) -> Tensor:
    """video_net.py:106-113: cat(sin, cos) of t * exp(-i*ln(1e4)/(dim/2-1))."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, device=t.device) * -(math.log(10000) / (half - 1)))
    arg = t[:, None] * freq[None, :]
    return torch.cat((arg.sin(), arg.cos()), dim=-1)


def relative_position_bucket(rel: Tensor, num_buckets: int = 32, max_distance: int = 128) -> Tensor:
    """video_net.py:276-300 (T5 bidirectional buckets)."""
    n = -rel
    half = num_buckets // 2
    ret = (n < 0).long() * half
    n = n.abs()
    max_exact = half // 2
    large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact)
                         * (half - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, half - 1))
    return ret + torch.where(n < max_exact, n, large)


def rel_pos_bucket_table(n: int, num_buckets: int = 32, max_distance: int = 32) -> Tensor:
    """[n, n] bucket index for rel = k_pos - q_pos (video_net.py:302-308)."""
    pos = torch.arange(n, dtype=torch.long)
    rel = pos[None, :] - pos[:, None]
    return relative_position_bucket(rel, num_buckets, max_distance)


def rel_pos_bias(emb_weight: Tensor, n: int, cfg: OracleConfig) -> Tensor:
    """video_net.py:302-310: Embedding gather -> [heads, n, n]."""
    idx = rel_pos_bucket_table(n, cfg.rel_pos_num_buckets, cfg.rel_pos_max_distance).to(emb_weight.device)
    return emb_weight[idx].permute(2, 0, 1)


def rotary_angles(freqs: Tensor, seq_len: int) -> Tensor:
    """rotary_embedding.py:143-144,275-278: angle[p, 2i] = angle[p, 2i+1] = p * freqs[i]."""
    pos = torch.arange(seq_len, device=freqs.device, dtype=freqs.dtype)
    return (pos[:, None] * freqs[None, :]).repeat_interleave(2, dim=-1)


def apply_rotary(t: Tensor, angles: Tensor) -> Tensor:
    """rotary_embedding.py:29-48: t*cos + rotate_half(t)*sin on the first rot_dim features;
    rotate_half maps interleaved pairs (x0, x1) -> (-x1, x0).  t: [..., seq, d]."""
    rot = angles.shape[-1]
    head, tail = t[..., :rot], t[..., rot:]
    pairs = head.reshape(*head.shape[:-1], rot // 2, 2)
    rotated = torch.stack((-pairs[..., 1], pairs[..., 0]), dim=-1).reshape(head.shape)
    out = head * angles.cos() + rotated * angles.sin()
    return torch.cat((out, tail), dim=-1)


def channel_layer_norm(x: Tensor, gamma: Tensor, eps: float = 1e-5) -> Tensor:
    """video_net.py:84-87: normalise over dim 1, biased variance, eps inside the sqrt, gain only."""
    var = x.var(dim=1, unbiased=False, keepdim=True)
    mean = x.mean(dim=1, keepdim=True)
    return (x - mean) / (var + eps).sqrt() * gamma


# --------------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------------
def block(sd: StateDict, pre: str, x: Tensor, groups: int, scale_shift=None) -> Tensor:
    """video_net.py:219-227: Conv3d(1,3,3) -> GroupNorm -> optional FiLM -> SiLU."""
    x = F.conv3d(x, sd[pre + "proj.weight"], sd[pre + "proj.bias"], padding=(0, 1, 1))
    x = F.group_norm(x, groups, sd[pre + "norm.weight"], sd[pre + "norm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def resnet_block(sd: StateDict, pre: str, x: Tensor, temb: Optional[Tensor], groups: int) -> Tensor:
    """video_net.py:254-265."""
    scale_shift = None
    if (pre + "mlp.1.weight") in sd:
        assert temb is not None, "time emb must be passed in"
        e = F.linear(F.silu(temb), sd[pre + "mlp.1.weight"], sd[pre + "mlp.1.bias"])
        e = e[:, :, None, None, None]
        scale_shift = e.chunk(2, dim=1)
    h = block(sd, pre + "block1.", x, groups, scale_shift)
    h = block(sd, pre + "block2.", h, groups)
    if (pre + "res_conv.weight") in sd:
        res = F.conv3d(x, sd[pre + "res_conv.weight"], sd[pre + "res_conv.bias"])
    else:
        res = x
    return h + res


def spatial_linear_attention(sd: StateDict, pre: str, x: Tensor, heads: int) -> Tensor:
    """video_net.py:331-347 (per-frame linear attention, softmax(q) over d, softmax(k) over n)."""
    b, c, f, h, w = x.shape
    xf = x.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
    qkv = F.conv2d(xf, sd[pre + "to_qkv.weight"])
    hidden = qkv.shape[1] // 3
    d = hidden // heads
    q, k, v = (t.reshape(b * f, heads, d, h * w) for t in qkv.chunk(3, dim=1))
    q = q.softmax(dim=-2) * d ** -0.5
    k = k.softmax(dim=-1)
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", context, q)
    out = out.reshape(b * f, hidden, h, w)
    out = F.conv2d(out, sd[pre + "to_out.weight"], sd[pre + "to_out.bias"])
    return out.reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)


def attention(sd: StateDict, pre: str, x: Tensor, heads: int, pos_bias: Optional[Tensor],
              rotary: bool) -> Tensor:
    """video_net.py:395-454 with an all-False focus_present_mask.  x: [..., n, c]."""
    qkv = F.linear(x, sd[pre + "to_qkv.weight"])
    hidden = qkv.shape[-1] // 3
    d = hidden // heads
    n = x.shape[-2]

    def split(t):  # "... n (h d) -> ... h n d"
        return t.reshape(*t.shape[:-1], heads, d).transpose(-2, -3)

    q, k, v = (split(t) for t in qkv.chunk(3, dim=-1))
    q = q * d ** -0.5
    if rotary:
        ang = rotary_angles(sd[pre + "rotary_emb.freqs"], n)
        q, k = apply_rotary(q, ang), apply_rotary(k, ang)
    sim = torch.einsum("...hid,...hjd->...hij", q, k)
    if pos_bias is not None:
        sim = sim + pos_bias
    sim = sim - sim.amax(dim=-1, keepdim=True).detach()
    attn = sim.softmax(dim=-1)
    out = torch.einsum("...hij,...hjd->...hid", attn, v)
    out = out.transpose(-2, -3).reshape(*x.shape[:-1], hidden)
    return F.linear(out, sd[pre + "to_out.weight"])


def temporal_attention_block(sd: StateDict, pre: str, x: Tensor, heads: int, pos_bias: Tensor) -> Tensor:
    """Residual(PreNorm(EinopsToAndFrom('b c f h w' -> 'b (h w) f c', Attention)))
    (video_net.py:69-75, 90-98, 357-365, 611-643).  `pre` is the Residual module's prefix."""
    b, c, f, h, w = x.shape
    y = channel_layer_norm(x, sd[pre + "fn.norm.gamma"])
    y = y.permute(0, 3, 4, 2, 1).reshape(b, h * w, f, c)
    y = attention(sd, pre + "fn.fn.fn.", y, heads, pos_bias, rotary=True)
    y = y.reshape(b, h, w, f, c).permute(0, 4, 3, 1, 2)
    return y + x


def spatial_attention_block(sd: StateDict, pre: str, x: Tensor, heads: int) -> Tensor:
    """Residual(PreNorm(SpatialLinearAttention)) (video_net.py:688-697)."""
    y = channel_layer_norm(x, sd[pre + "fn.norm.gamma"])
    return spatial_linear_attention(sd, pre + "fn.fn.", y, heads) + x


def mid_spatial_attention_block(sd: StateDict, pre: str, x: Tensor, heads: int) -> Tensor:
    """use_mid_attn=True branch (video_net.py:713-719): full softmax attention over h*w per frame,
    no rotary, no positional bias."""
    b, c, f, h, w = x.shape
    y = channel_layer_norm(x, sd[pre + "fn.norm.gamma"])
    y = y.permute(0, 2, 3, 4, 1).reshape(b, f, h * w, c)
    y = attention(sd, pre + "fn.fn.fn.", y, heads, None, rotary=False)
    y = y.reshape(b, f, h, w, c).permute(0, 4, 1, 2, 3)
    return y + x


# --------------------------------------------------------------------------------------------
# the network
# --------------------------------------------------------------------------------------------
def unet3d_forward(sd: StateDict, cfg: OracleConfig, x: Tensor, timesteps: Tensor, cond_map: Optional[Tensor],
                   pre: str = "net.") -> Tensor:
    """video_net.py:766-871 with days/years/lowres_cond None and prob_focus_present 0."""
    heads, groups = cfg.attn_heads, cfg.resnet_groups
    pos_bias = rel_pos_bias(sd[pre + "time_rel_pos_bias.relative_attention_bias.weight"], x.shape[2], cfg)
    if cond_map is not None:
        x = torch.cat([x, cond_map], dim=1)
    pad = cfg.init_kernel_size // 2
    x = F.conv3d(x, sd[pre + "input_conv.weight"], sd[pre + "input_conv.bias"], padding=(0, pad, pad))
    x = temporal_attention_block(sd, pre + "input_temp_op.", x, heads, pos_bias)
    r = x
    t = sinusoidal_pos_emb(timesteps, cfg.model_dim)
    t = F.linear(t, sd[pre + "time_mlp.1.weight"], sd[pre + "time_mlp.1.bias"])
    t = F.linear(F.silu(t), sd[pre + "time_mlp.3.weight"], sd[pre + "time_mlp.3.bias"])

    n_levels = len(cfg.dim_mults)
    skips = []
    for lvl in range(n_levels):
        p = f"{pre}downs.{lvl}."
        x = resnet_block(sd, p + "0.", x, t, groups)
        x = resnet_block(sd, p + "1.", x, t, groups)
        if (p + "2.fn.norm.gamma") in sd:
            x = spatial_attention_block(sd, p + "2.", x, heads)
        x = temporal_attention_block(sd, p + "3.", x, heads, pos_bias)
        skips.append(x)
        if (p + "4.weight") in sd:  # Downsample, video_net.py:61-62
            x = F.conv3d(x, sd[p + "4.weight"], sd[p + "4.bias"], stride=(1, 2, 2), padding=(0, 1, 1))

    x = resnet_block(sd, pre + "mid_block1.", x, t, groups)
    if (pre + "mid_spatial_attn.fn.norm.gamma") in sd:
        x = mid_spatial_attention_block(sd, pre + "mid_spatial_attn.", x, heads)
    x = temporal_attention_block(sd, pre + "mid_temporal_attn.", x, heads, pos_bias)
    x = resnet_block(sd, pre + "mid_block2.", x, t, groups)

    for lvl in range(n_levels):
        p = f"{pre}ups.{lvl}."
        x = torch.cat((x, skips.pop()), dim=1)
        x = resnet_block(sd, p + "0.", x, t, groups)
        x = resnet_block(sd, p + "1.", x, t, groups)
        if (p + "2.fn.norm.gamma") in sd:
            x = spatial_attention_block(sd, p + "2.", x, heads)
        x = temporal_attention_block(sd, p + "3.", x, heads, pos_bias)
        if (p + "4.weight") in sd:  # Upsample, video_net.py:65-66
            x = F.conv_transpose3d(x, sd[p + "4.weight"], sd[p + "4.bias"], stride=(1, 2, 2), padding=(0, 1, 1))

    x = torch.cat((x, r), dim=1)
    x = resnet_block(sd, pre + "out_conv.0.", x, None, groups)
    return F.conv3d(x, sd[pre + "out_conv.1.weight"], sd[pre + "out_conv.1.bias"])


def unet_forward(sd: StateDict, cfg: OracleConfig, x_t: Tensor, cond: Tensor, t: Tensor, pre: str = "net.") -> Tensor:
    """model.py:85-134: frame alignment, 3-D net, centre frame."""
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
    fx, fc = x_t.shape[2], cond.shape[2]
    if fx != fc:
        if fx == 1 and fc > 1:
            x_t = x_t.expand(-1, -1, fc, -1, -1)
        elif fc == 1 and fx > 1:
            cond = cond.expand(-1, -1, fx, -1, -1)
        else:
            raise ValueError(f"Frame mismatch: x_t F={fx}, cond F={fc}")
    out = unet3d_forward(sd, cfg, x_t, t, cond, pre)
    fo = out.shape[2]
    return out.squeeze(2) if fo == 1 else out[:, :, fo // 2]


# --------------------------------------------------------------------------------------------
# DDPM wrapper
# --------------------------------------------------------------------------------------------
def diffusion_buffers(timesteps: int = 1000, beta_schedule: str = "linear") -> Dict[str, Tensor]:
    """model.py:148-165."""
    if beta_schedule != "linear":
        raise ValueError("Only 'linear' beta_schedule implemented")
    betas = torch.linspace(1e-4, 2e-2, timesteps)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = torch.cat([torch.tensor([1.0]), ac[:-1]], dim=0)
    return {
        "betas": betas, "alphas": alphas, "alphas_cumprod": ac, "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac), "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "sqrt_recip_alphas": torch.sqrt(1.0 / alphas),
        "posterior_variance": betas * (1.0 - ac_prev) / (1.0 - ac),
    }


def q_sample(buf: Dict[str, Tensor], x0: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """model.py:196-201."""
    a = buf["sqrt_alphas_cumprod"][t].view(-1, 1, 1, 1)
    s = buf["sqrt_one_minus_alphas_cumprod"][t].view(-1, 1, 1, 1)
    return a * x0 + s * noise


def diffusion_loss(sd: StateDict, cfg: OracleConfig, buf: Dict[str, Tensor], x0: Tensor, cond: Tensor, t: Tensor,
                   noise: Tensor, pre: str = "net.") -> Tensor:
    """model.py:203-208 with t and noise supplied by the caller (the reference draws them with
    torch.randint / torch.randn_like, in that order)."""
    x_t = q_sample(buf, x0, t, noise)
    return F.mse_loss(unet_forward(sd, cfg, x_t, cond, t, pre), noise)


def p_sample(sd: StateDict, cfg: OracleConfig, buf: Dict[str, Tensor], x_t: Tensor, cond: Tensor, t: Tensor,
             noise: Optional[Tensor], pre: str = "net.") -> Tensor:
    """model.py:168-183; `noise` is ignored (may be None) when every t == 0."""
    beta = buf["betas"][t].view(-1, 1, 1, 1)
    s1m = buf["sqrt_one_minus_alphas_cumprod"][t].view(-1, 1, 1, 1)
    sra = buf["sqrt_recip_alphas"][t].view(-1, 1, 1, 1)
    eps = unet_forward(sd, cfg, x_t, cond, t, pre)
    mean = sra * (x_t - beta / s1m * eps)
    if bool((t == 0).all()):
        return mean
    return mean + torch.sqrt(buf["posterior_variance"][t].view(-1, 1, 1, 1)) * noise


@torch.no_grad()
def sample(sd: StateDict, cfg: OracleConfig, buf: Dict[str, Tensor], cond: Tensor, shape, pre: str = "net.",
           generator: Optional[torch.Generator] = None) -> Tensor:
    """model.py:186-194: x ~ N(0,1), then T reverse steps drawing randn_like(x) per step with t > 0."""
    T = buf["betas"].shape[0]
    x = torch.randn(shape, generator=generator)
    for tt in reversed(range(T)):
        t = torch.full((shape[0],), tt, dtype=torch.long)
        z = torch.randn(shape, generator=generator) if tt > 0 else None
        x = p_sample(sd, cfg, buf, x, cond, t, z, pre)
    return x


def loss_and_grads(sd: StateDict, cfg: OracleConfig, buf: Dict[str, Tensor], x0: Tensor, cond: Tensor, t: Tensor,
                   noise: Tensor, pre: str = "net.") -> Tuple[Tensor, Dict[str, Tensor]]:
    """Loss and d(loss)/d(param) for every floating-point entry of `sd` that the reference trains
    (everything except the rotary `freqs`, rotary_embedding.py:107)."""
    leaves = {}
    for k, v in sd.items():
        v = v.detach().clone()
        if v.is_floating_point() and not k.endswith("rotary_emb.freqs"):
            v.requires_grad_(True)
        leaves[k] = v
    loss = diffusion_loss(leaves, cfg, buf, x0, cond, t, noise, pre)
    names = [k for k, v in leaves.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return loss.detach(), {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names, grads)}


# --------------------------------------------------------------------------------------------
# a state dict of the reference's layout without the reference (timing legs of bench.py)
# --------------------------------------------------------------------------------------------
def random_state_dict(cfg: OracleConfig, seed: int = 0, time_dim: Optional[int] = None, pre: str = "net.") -> StateDict:
    """Random weights in the key / shape layout of model.UNet(...).state_dict() (video_net.py:562-764, SURVEY.md
    section 8(b)); scales follow PyTorch's default fan-in initialisation so that activations stay O(1).  Used where
    only the arithmetic is timed (bench.py's CPU / stock-PyTorch legs): the product package is not imported there."""
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    dim, heads, dh = cfg.model_dim, cfg.attn_heads, cfg.attn_dim_head
    hidden, sp_hidden = heads * dh, heads * cfg.spatial_dim_head
    tdim = dim * 4 if time_dim is None else time_dim  # video_net.py:650
    ks = cfg.init_kernel_size

    def w(key, *shape, fan_in=None):
        fan = fan_in if fan_in is not None else max(1, int(torch.tensor(shape[1:]).prod()))
        bound = 1.0 / math.sqrt(fan)
        sd[pre + key] = (torch.rand(*shape, generator=g) * 2 - 1) * bound

    def conv(key, cout, cin, kh, kw, bias=True):
        w(key + ".weight", cout, cin, 1, kh, kw)
        if bias:
            w(key + ".bias", cout, fan_in=cin * kh * kw)

    def temporal(key, c):
        sd[pre + key + ".fn.norm.gamma"] = torch.ones(1, c, 1, 1, 1)
        w(key + ".fn.fn.fn.to_qkv.weight", 3 * hidden, c)
        w(key + ".fn.fn.fn.to_out.weight", c, hidden)
        sd[pre + key + ".fn.fn.fn.rotary_emb.freqs"] = 1.0 / (10000 ** (torch.arange(0, min(32, dh), 2).float() / min(32, dh)))

    def spatial(key, c):
        sd[pre + key + ".fn.norm.gamma"] = torch.ones(1, c, 1, 1, 1)
        w(key + ".fn.fn.to_qkv.weight", 3 * sp_hidden, c, 1, 1)
        w(key + ".fn.fn.to_out.weight", c, sp_hidden, 1, 1)
        w(key + ".fn.fn.to_out.bias", c, fan_in=sp_hidden)

    def resnet(key, cin, cout, temb=True):
        if temb:
            w(key + ".mlp.1.weight", 2 * cout, tdim)
            w(key + ".mlp.1.bias", 2 * cout, fan_in=tdim)
        conv(key + ".block1.proj", cout, cin, 3, 3)
        sd[pre + key + ".block1.norm.weight"], sd[pre + key + ".block1.norm.bias"] = torch.ones(cout), torch.zeros(cout)
        conv(key + ".block2.proj", cout, cout, 3, 3)
        sd[pre + key + ".block2.norm.weight"], sd[pre + key + ".block2.norm.bias"] = torch.ones(cout), torch.zeros(cout)
        if cin != cout:
            conv(key + ".res_conv", cout, cin, 1, 1)

    cin0 = cfg.n_vars + (1 if cfg.cond_map else 0)
    conv("input_conv", dim, cin0, ks, ks)
    sd[pre + "time_rel_pos_bias.relative_attention_bias.weight"] = torch.randn(cfg.rel_pos_num_buckets, heads, generator=g)
    temporal("input_temp_op", dim)
    w("time_mlp.1.weight", tdim, dim)
    w("time_mlp.1.bias", tdim, fan_in=dim)
    w("time_mlp.3.weight", tdim, tdim)
    w("time_mlp.3.bias", tdim, fan_in=tdim)
    dims = cfg.dims
    in_out = list(zip(dims[:-1], dims[1:]))
    n = len(in_out)
    for i, (ci, co) in enumerate(in_out):
        last = i >= n - 1
        resnet(f"downs.{i}.0", ci, co)
        resnet(f"downs.{i}.1", co, co)
        if cfg.use_sparse_linear_attn:
            spatial(f"downs.{i}.2", co)
        temporal(f"downs.{i}.3", co)
        if not last:
            conv(f"downs.{i}.4", co, co, 4, 4)
    mid = dims[-1]
    resnet("mid_block1", mid, mid)
    temporal("mid_temporal_attn", mid)
    resnet("mid_block2", mid, mid)
    for i, (ci, co) in enumerate(reversed(in_out)):
        last = i >= n - 1
        resnet(f"ups.{i}.0", co * 2, ci)
        resnet(f"ups.{i}.1", ci, ci)
        if cfg.use_sparse_linear_attn:
            spatial(f"ups.{i}.2", ci)
        temporal(f"ups.{i}.3", ci)
        if not last:
            # ConvTranspose3d weight is [cin, cout, 1, 4, 4] (video_net.py:65-66)
            w(f"ups.{i}.4.weight", ci, ci, 1, 4, 4)
            w(f"ups.{i}.4.bias", ci, fan_in=ci * 16)
    resnet("out_conv.0", dim * 2, dim, temb=False)
    conv("out_conv.1", cfg.n_vars, dim, 1, 1)
    return sd
