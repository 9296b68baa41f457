"""Thin tensor-level wrappers over the C ABI: validate, take data_ptr()/current stream, call.

Nothing here computes with PyTorch; torch only owns the memory and the stream.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import IgemmArgs

BF16 = torch.bfloat16


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.CesmError("cesm_emulator_b200 kernels need CUDA tensors (no CPU fallback)")
        if t is not None and not t.is_contiguous():
            raise _lib.CesmError("cesm_emulator_b200 kernels need contiguous tensors")


TAPS_3x3 = [(dh, dw) for dh in (-1, 0, 1) for dw in (-1, 0, 1)]
TAPS_1x1 = [(0, 0)]


def igemm(a0: torch.Tensor, wt: torch.Tensor, *, taps: Sequence[Tuple[int, int]] = TAPS_1x1,
          a1: Optional[torch.Tensor] = None, stride: int = 1,
          out: Optional[torch.Tensor] = None, out_hw: Optional[Tuple[int, int]] = None,
          out_place: Tuple[int, int, int, int] = (1, 1, 0, 0),
          bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
          out_dtype: torch.dtype = BF16) -> torch.Tensor:
    """out[n,oh,ow,:] = sum_t A[n, oh*stride+dh_t, ow*stride+dw_t, :] @ wt[:, t, :]^T (+bias)(+residual).

    a0/a1: bf16 [N,H,W,C]; wt: bf16 [cout, len(taps)*(C0+C1)].  `out_hw` is the iterated output
    grid (defaults to H//stride, W//stride); `out_place` = (o_sh, o_sw, o_h0, o_w0) scatters the
    grid into a larger `out` tensor (transposed-conv phases).
    """
    _req_cuda(a0, a1, wt, out, bias, residual)
    assert a0.dtype == BF16 and wt.dtype == BF16 and a0.dim() == 4
    n, h, w, c0 = a0.shape
    c1 = 0 if a1 is None else a1.shape[-1]
    cout = wt.shape[0]
    assert wt.shape[1] == len(taps) * (c0 + c1), (wt.shape, len(taps), c0, c1)
    oh, ow = out_hw if out_hw is not None else (h // stride, w // stride)
    o_sh, o_sw, o_h0, o_w0 = out_place
    if out is None:
        assert out_place == (1, 1, 0, 0)
        out = torch.empty((n, oh, ow, cout), dtype=out_dtype, device=a0.device)
    args = IgemmArgs()
    args.a0, args.a1, args.c0, args.c1 = _ptr(a0), _ptr(a1), c0, c1
    args.n, args.h, args.w, args.stride = n, h, w, stride
    args.num_taps = len(taps)
    for i, (dh, dw) in enumerate(taps):
        args.tap_dh[i], args.tap_dw[i] = dh, dw
    args.wt, args.cout = _ptr(wt), cout
    args.oh, args.ow = oh, ow
    args.out, args.out_fp32, args.ldo = _ptr(out), int(out.dtype == torch.float32), out.shape[-1]
    args.out_h, args.out_w = out.shape[1], out.shape[2]
    args.o_sh, args.o_sw, args.o_h0, args.o_w0 = o_sh, o_sw, o_h0, o_w0
    args.bias = _ptr(bias)
    args.residual = _ptr(residual)
    args.ldr = 0 if residual is None else residual.shape[-1]
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == cout
    if residual is not None:
        assert residual.dtype == BF16 and residual.shape[:3] == out.shape[:3]
    _lib.call("cesm_igemm", ctypes.byref(args), _stream())
    return out
