"""torch.autograd.Function wrappers: one per fused CUDA op, forward and backward both in the C ABI.

Internal activation layout is channels-last bf16 `[N, H, W, C]` with N = batch*frames (the
reference's NCDHW tensors, permuted once at the network boundary).  Parameters stay fp32 in the
reference's layouts; bf16 GEMM operands are re-packed from them on the fly.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import kernels as K

BF16 = torch.bfloat16

# sub-pixel decomposition of a k4/s2/p1 transposed conv: output phase -> [(input offset, kernel index)]
_PHASE_TAPS = {0: [(0, 1), (-1, 3)], 1: [(1, 0), (0, 2)]}


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.contiguous()


def _phase(ph: int, pw: int):
    taps, koff = [], []
    for dh, kh in _PHASE_TAPS[ph]:
        for dw, kw in _PHASE_TAPS[pw]:
            taps.append((dh, dw))
            koff.append(kh * 4 + kw)
    return taps, koff


# Fused / capturable optimizers update parameters without bumping `Tensor._version`, so the
# packed-operand caches are also keyed on a global epoch that every torch optimizer step advances.
_WEIGHT_EPOCH = [0]


def invalidate_weight_cache(*_args, **_kwargs) -> None:
    """Call after modifying parameters by any means torch cannot see (raw pointer writes)."""
    _WEIGHT_EPOCH[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_hook
    _reg_hook(invalidate_weight_cache)
except Exception:  # pragma: no cover - very old torch
    pass


class PackCache:
    """bf16 GEMM-operand copies of one layer's fp32 parameter, keyed by layout tag and re-made
    when the parameter changes (optimizer step bumps `_version`).  While a CUDA graph of a TRAINING
    step is being captured a trainable parameter is always re-packed, so that the pack kernel is
    part of the graph and replays see the current weights; a no-grad (sampling) capture uses the
    cached copy, which stays valid as long as the weights are not modified."""

    def __init__(self):
        self._d = {}

    @staticmethod
    def must_repack(weight: torch.Tensor) -> bool:
        """True in the forward of a training step that is being graph-captured."""
        return weight.requires_grad and torch.is_grad_enabled() and torch.cuda.is_current_stream_capturing()

    def get(self, weight: torch.Tensor, tag, make, force: bool = False):
        key = (weight.data_ptr(), weight._version, _WEIGHT_EPOCH[0])
        if not force:
            hit = self._d.get(tag)
            if hit is not None and hit[0] == key:
                return hit[1]
        val = make()
        self._d[tag] = (key, val)
        return val

    def __deepcopy__(self, memo):
        return PackCache()


# ------------------------------------------------------------------------------------------------
# convolutions / linears on the tcgen05 implicit GEMM
# ------------------------------------------------------------------------------------------------
class ConvFn(torch.autograd.Function):
    """Stride-1 conv with a square k x k kernel (k in {1, 3}) or a linear layer, over one or two
    concatenated channels-last sources, with fused bias and residual.

    weight: fp32 [cout, cin, (1,) k, k] (Conv3d / Conv2d) or [cout, cin] (Linear).
    Reference: video_net.py:215 (Block.proj), :246 (res_conv), :322-323, :380-381.
    """

    @staticmethod
    def forward(ctx, x0, x1, weight, bias, residual, ksize: int, cache: PackCache):
        x0, x1, residual = _c(x0), _c(x1), _c(residual)
        cout = weight.shape[0]
        c0 = x0.shape[-1]
        c1 = 0 if x1 is None else x1.shape[-1]
        kk = ksize * ksize
        assert weight.numel() == cout * (c0 + c1) * kk, (weight.shape, c0, c1, ksize)
        r = ksize // 2
        taps = [(kh - r, kw - r) for kh in range(ksize) for kw in range(ksize)]
        ctx.force = PackCache.must_repack(weight)
        wt = cache.get(weight, "fwd",
                       lambda: K.pack_weight(weight, cout, kk, c0 + c1, (c0 + c1) * kk, kk, list(range(kk))), ctx.force)
        y = K.igemm(x0, wt, a1=x1, taps=taps, bias=bias, residual=residual)
        ctx.save_for_backward(x0, x1, weight)
        ctx.ksize, ctx.has_bias, ctx.has_res = ksize, bias is not None, residual is not None
        ctx.taps, ctx.cache = taps, cache
        return y

    @staticmethod
    def backward(ctx, dy):
        x0, x1, weight = ctx.saved_tensors
        dy = dy.contiguous()
        cout = weight.shape[0]
        c0 = x0.shape[-1]
        c1 = 0 if x1 is None else x1.shape[-1]
        ctot, kk = c0 + c1, ctx.ksize * ctx.ksize
        ntaps = [(-dh, -dw) for dh, dw in ctx.taps]
        dx0 = dx1 = dwt = db = None
        wflat = weight.reshape(-1)
        if ctx.needs_input_grad[0]:
            wd = ctx.cache.get(weight, "dgrad0",
                               lambda: K.pack_weight(wflat, c0, kk, cout, kk, ctot * kk, list(range(kk))), ctx.force)
            dx0 = K.igemm(dy, wd, taps=ntaps)
        if x1 is not None and ctx.needs_input_grad[1]:
            wd = ctx.cache.get(weight, "dgrad1",
                               lambda: K.pack_weight(wflat[c0 * kk:], c1, kk, cout, kk, ctot * kk, list(range(kk))),
                               ctx.force)
            dx1 = K.igemm(dy, wd, taps=ntaps)
        if ctx.needs_input_grad[2]:
            g = K.wgrad(x0, dy, x1=x1, taps=ctx.taps)  # [cout, kk, ctot]
            dwt = torch.empty_like(weight)
            K.unpack_wgrad(g, dwt, cout, kk, ctot, ctot * kk, kk, list(range(kk)))
        if ctx.has_bias and ctx.needs_input_grad[3]:
            db = K.colsum(dy)
        dres = dy if (ctx.has_res and ctx.needs_input_grad[4]) else None
        return dx0, dx1, dwt, db, dres, None, None


class DownsampleFn(torch.autograd.Function):
    """Conv3d(dim, dim, (1,4,4), (1,2,2), (0,1,1)) -- video_net.py:61-62."""

    TAPS = [(kh - 1, kw - 1) for kh in range(4) for kw in range(4)]

    @staticmethod
    def forward(ctx, x, weight, bias, cache: PackCache):
        x = x.contiguous()
        c = x.shape[-1]
        ctx.force = PackCache.must_repack(weight)
        wt = cache.get(weight, "fwd", lambda: K.pack_weight(weight, c, 16, c, c * 16, 16, list(range(16))), ctx.force)
        y = K.igemm(x, wt, taps=DownsampleFn.TAPS, stride=2, bias=bias)
        ctx.save_for_backward(x, weight)
        ctx.cache = cache
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        n, h, w, c = x.shape
        dx = dwt = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            for ph in (0, 1):
                for pw in (0, 1):
                    taps, koff = _phase(ph, pw)
                    # [ci][t][co] = W[co, ci, kh_t, kw_t]
                    wd = ctx.cache.get(weight, ("dgrad", ph, pw),
                                       lambda: K.pack_weight(weight, c, 4, c, 16, c * 16, koff), ctx.force)
                    K.igemm(dy, wd, taps=taps, out=dx, out_hw=(h // 2, w // 2), out_place=(2, 2, ph, pw))
        if ctx.needs_input_grad[1]:
            g = K.wgrad(x, dy, taps=DownsampleFn.TAPS, stride=2)
            dwt = torch.empty_like(weight)
            K.unpack_wgrad(g, dwt, c, 16, c, c * 16, 16, list(range(16)))
        if ctx.needs_input_grad[2]:
            db = K.colsum(dy)
        return dx, dwt, db, None


class UpsampleFn(torch.autograd.Function):
    """ConvTranspose3d(dim, dim, (1,4,4), (1,2,2), (0,1,1)) as four sub-pixel 2x2 convs -- video_net.py:65-66."""

    @staticmethod
    def forward(ctx, x, weight, bias, cache: PackCache):
        x = x.contiguous()
        n, h, w, c = x.shape
        out = torch.empty((n, 2 * h, 2 * w, c), dtype=BF16, device=x.device)
        ctx.force = PackCache.must_repack(weight)
        for ph in (0, 1):
            for pw in (0, 1):
                taps, koff = _phase(ph, pw)
                # [co][t][ci] = W[ci, co, kh_t, kw_t]
                wt = cache.get(weight, ("fwd", ph, pw), lambda: K.pack_weight(weight, c, 4, c, 16, c * 16, koff),
                               ctx.force)
                K.igemm(x, wt, taps=taps, out=out, out_hw=(h, w), out_place=(2, 2, ph, pw), bias=bias)
        ctx.save_for_backward(x, weight)
        ctx.cache = cache
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        n, h, w, c = x.shape
        dx = dwt = db = None
        if ctx.needs_input_grad[0]:
            # [ci][t][co] = W[ci, co, kh, kw]
            wd = ctx.cache.get(weight, "dgrad", lambda: K.pack_weight(weight, c, 16, c, c * 16, 16, list(range(16))),
                               ctx.force)
            dx = K.igemm(dy, wd, taps=DownsampleFn.TAPS, stride=2)
        if ctx.needs_input_grad[1]:
            dwt = torch.empty_like(weight)
            for ph in (0, 1):
                for pw in (0, 1):
                    taps, koff = _phase(ph, pw)
                    g = K.wgrad(x, dy, taps=taps, grid_hw=(h, w), dy_place=(2, 2, ph, pw))  # [co, 4, ci]
                    K.unpack_wgrad(g, dwt, c, 4, c, 16, c * 16, koff)
        if ctx.needs_input_grad[2]:
            db = K.colsum(dy)
        return dx, dwt, db, None


# ------------------------------------------------------------------------------------------------
# normalisation
# ------------------------------------------------------------------------------------------------
class GroupNormSiLUFn(torch.autograd.Function):
    """GroupNorm -> optional FiLM (x*(scale+1)+shift) -> SiLU -> optional residual add
    (video_net.py:221-227, :265).  x: [B*F, H, W, C]; film: fp32 [B, 2C]."""

    @staticmethod
    def forward(ctx, x, gamma, beta, film, residual, B: int, G: int, eps: float):
        x, film, residual = x.contiguous(), _c(film), _c(residual)
        sums = K.gn_stats(x, B, G)
        out = K.gn_apply_fwd(x, sums, gamma, beta, film, residual, B, G, eps)
        ctx.save_for_backward(x, sums, gamma, beta, film)
        ctx.cfg = (B, G, eps, residual is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, sums, gamma, beta, film = ctx.saved_tensors
        B, G, eps, has_res = ctx.cfg
        dout = dout.contiguous()
        dx, dgamma, dbeta, dfilm = K.gn_bwd(x, dout, sums, gamma, beta, film, B, G, eps)
        return dx, dgamma, dbeta, dfilm, (dout if has_res else None), None, None, None


class LayerNormFn(torch.autograd.Function):
    """Channel LayerNorm with gain only (video_net.py:84-87); gamma is the [1,C,1,1,1] parameter."""

    @staticmethod
    def forward(ctx, x, gamma, eps: float):
        x = x.contiguous()
        g = gamma.reshape(-1)
        ctx.save_for_backward(x, g)
        ctx.eps, ctx.gshape = eps, gamma.shape
        return K.ln_fwd(x, g, eps)

    @staticmethod
    def backward(ctx, dy):
        x, g = ctx.saved_tensors
        dx, dg = K.ln_bwd(x, g, dy.contiguous(), None, ctx.eps)
        return dx, dg.view(ctx.gshape), None


# ------------------------------------------------------------------------------------------------
# attention cores
# ------------------------------------------------------------------------------------------------
class TemporalAttnCoreFn(torch.autograd.Function):
    """q*scale, RoPE, q.k + bias, softmax over frames, .v (video_net.py:413-453).  qkv: [B*F*HW, 3*H*D]."""

    @staticmethod
    def forward(ctx, qkv, pos_bias, cs, sn, B: int, F: int, HW: int, H: int, D: int):
        qkv, pos_bias = qkv.contiguous(), pos_bias.contiguous().float()
        scale = D ** -0.5
        out, lse = K.tattn_fwd(qkv, pos_bias, cs, sn, B, F, HW, H, D, scale)
        if F <= 4:  # the small-window kernel recomputes the softmax in the backward
            ctx.save_for_backward(qkv, pos_bias, cs, sn)
        else:
            ctx.save_for_backward(qkv, pos_bias, cs, sn, out, lse)
        ctx.dims = (B, F, HW, H, D, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, F, HW, H, D, scale = ctx.dims
        if F <= 4:
            (qkv, pos_bias, cs, sn), out, lse = ctx.saved_tensors, None, None
        else:
            qkv, pos_bias, cs, sn, out, lse = ctx.saved_tensors
        dqkv, dbias = K.tattn_bwd(qkv, pos_bias, cs, sn, out, lse, dout.contiguous(), B, F, HW, H, D, scale)
        return dqkv, dbias, None, None, None, None, None, None, None


class RelPosBiasFn(torch.autograd.Function):
    """Embedding gather of the T5 bucket table -> [heads, n, n] (video_net.py:302-310).  The
    table has 32 x heads entries; gather and scatter-add are index ops on a few hundred floats."""

    @staticmethod
    def forward(ctx, weight, idx):
        ctx.save_for_backward(idx)
        ctx.shape = weight.shape
        return weight.detach()[idx].permute(2, 0, 1).float().contiguous()

    @staticmethod
    def backward(ctx, dbias):
        (idx,) = ctx.saved_tensors
        dw = torch.zeros(ctx.shape, dtype=dbias.dtype, device=dbias.device)
        dw.index_add_(0, idx.reshape(-1), dbias.permute(1, 2, 0).reshape(-1, ctx.shape[1]))
        return dw, None


class LinearAttnCoreFn(torch.autograd.Function):
    """softmax(q) over d, softmax(k) over pixels, ctx = k^T v, out = ctx^T q (video_net.py:338-344)."""

    @staticmethod
    def forward(ctx, qkv, NI: int, n: int, H: int, D: int):
        qkv = qkv.contiguous()
        scale = D ** -0.5
        out, ws = K.linattn_fwd(qkv, NI, n, H, D, scale)
        ctx.save_for_backward(qkv, ws)
        ctx.dims = (NI, n, H, D, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, ws = ctx.saved_tensors
        NI, n, H, D, scale = ctx.dims
        return K.linattn_bwd(qkv, ws, dout.contiguous(), NI, n, H, D, scale), None, None, None, None


# ------------------------------------------------------------------------------------------------
# network boundary
# ------------------------------------------------------------------------------------------------
class InputConvFn(torch.autograd.Function):
    """cat([x, cond_map], dim=1) -> Conv3d(2, C, (1,k,k)) with frame broadcast folded in
    (video_net.py:808-815, model.py:110-121).  x/cond: fp32 [B,1,Fx,H,W]; out: bf16 [B*F,H,W,C]."""

    @staticmethod
    def forward(ctx, x, cond, weight, bias, F: int):
        x, cond = x.contiguous().float(), cond.contiguous().float()
        B, H, W = x.shape[0], x.shape[-2], x.shape[-1]
        ks = weight.shape[-1]
        out = K.input_conv_fwd(x, cond, weight, bias, B, F, H, W, ks)
        ctx.save_for_backward(x, cond)
        ctx.dims = (B, F, H, W, ks)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, cond = ctx.saved_tensors
        B, F, H, W, ks = ctx.dims
        dw, db = K.input_conv_wgrad(x, cond, dy.contiguous(), B, F, H, W, ks)
        return None, None, dw, db, None


class OutConvFn(torch.autograd.Function):
    """Conv3d(C, 1, 1) evaluated on the centre frame (video_net.py:763 + model.py:129-130).
    a: bf16 [B*F, H, W, 64] -> fp32 [B, 1, H, W]."""

    @staticmethod
    def forward(ctx, a, weight, bias, B: int, F: int, mid: int):
        a = a.contiguous()
        H, W = a.shape[1], a.shape[2]
        ctx.save_for_backward(a, weight)
        ctx.dims = (B, F, H, W, mid)
        return K.out_conv_fwd(a, weight, bias, B, F, H, W, mid)

    @staticmethod
    def backward(ctx, deps):
        a, weight = ctx.saved_tensors
        B, F, H, W, mid = ctx.dims
        da, dw, db = K.out_conv_bwd(a, weight, deps.contiguous().float(), B, F, H, W, mid)
        return da, dw, db, None, None, None


class SmallLinearFn(torch.autograd.Function):
    """fp32 y = act(x) W^T + b for the time-embedding MLP and FiLM projections
    (video_net.py:651-656, 238-241); act = SiLU on the input when `act_in`."""

    @staticmethod
    def forward(ctx, x, weight, bias, act_in: bool):
        x = x.contiguous()
        ctx.save_for_backward(x, weight)
        ctx.act_in = act_in
        return K.small_linear_fwd(x, weight, bias, act_in)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dx, dW, db = K.small_linear_bwd(x, weight, dy.contiguous(), ctx.act_in, ctx.needs_input_grad[0])
        return dx, dW, db, None


class MseLossFn(torch.autograd.Function):
    """F.mse_loss(eps, noise) (model.py:208) with the gradient 2*(eps-noise)/N produced on device."""

    @staticmethod
    def forward(ctx, eps, noise):
        loss, diff = K.mse_fwd(eps.contiguous(), noise.contiguous())
        ctx.save_for_backward(diff)
        return loss

    @staticmethod
    def backward(ctx, g):
        (diff,) = ctx.saved_tensors
        g = g.reshape(1).float().contiguous()
        return K.scale_by_scalar(diff, g, 2.0 / diff.numel()), None


# ------------------------------------------------------------------------------------------------
# layout helpers (network boundary and module-level drop-in calls only)
# ------------------------------------------------------------------------------------------------
def to_cl(x: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
    """[B, C, F, H, W] (any float dtype) -> bf16 [B*F, H, W, C] contiguous, plus (B, F)."""
    B, C, F, H, W = x.shape
    return x.permute(0, 2, 3, 4, 1).reshape(B * F, H, W, C).to(BF16).contiguous(), B, F


def from_cl(y: torch.Tensor, B: int, F: int) -> torch.Tensor:
    """bf16 [B*F, H, W, C] -> [B, C, F, H, W] view."""
    n, H, W, C = y.shape
    return y.view(B, F, H, W, C).permute(0, 4, 1, 2, 3)
