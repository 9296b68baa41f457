"""Rotary position embedding for the temporal attention (reference: rotary_embedding.py:62-163).

Only the mode the emulator uses exists here: `freqs_for='lang'`, fixed (non-learned) frequencies
`theta^(-2i/dim)`, positions `arange(seq_len)`, interleaved pairing (rotary_embedding.py:29-48).
The module owns the same `freqs` Parameter (requires_grad=False) the reference registers, so
state dicts interchange; the rotation itself runs inside the fused attention kernel
(csrc/attn.cu), which consumes the per-frame cos/sin tables built by `tables()`.
xpos / learned / pixel / axial modes are never reached from video_net.py and are rejected.
"""
from __future__ import annotations

import torch
from torch import nn


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, custom_freqs=None, freqs_for="lang", theta=10000, max_freq=10, num_freqs=1,
                 learned_freq=False, use_xpos=False, xpos_scale_base=512, interpolate_factor=1.0,
                 theta_rescale_factor=1.0, seq_before_head_dim=False, cache_if_possible=True):
        super().__init__()
        if freqs_for != "lang" or learned_freq or use_xpos or custom_freqs is not None or interpolate_factor != 1.0:
            raise NotImplementedError(
                "cesm_emulator_b200.RotaryEmbedding implements the fixed 'lang' frequencies only "
                "(the one mode video_net.py:601 constructs)")
        theta = theta * theta_rescale_factor ** (dim / (dim - 2))
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: (dim // 2)].float() / dim))  # rotary_embedding.py:96
        self.dim = dim
        self.freqs = nn.Parameter(freqs, requires_grad=False)
        self.default_seq_dim = -2
        self._tables = {}

    def tables(self, seq_len: int):
        """fp32 cos/sin of `p * freqs[i]`, each [seq_len, dim/2], on the device of `freqs`."""
        key = (seq_len, self.freqs.device, self.freqs._version, self.freqs.data_ptr())
        hit = self._tables.get(key)
        if hit is None:
            with torch.no_grad():
                pos = torch.arange(seq_len, device=self.freqs.device, dtype=torch.float32)
                ang = pos[:, None] * self.freqs.detach().float()[None, :]
                hit = (ang.cos().contiguous(), ang.sin().contiguous())
            self._tables = {key: hit}
        return hit

    def rotate_queries_or_keys(self, t: torch.Tensor, seq_dim=None, offset=0) -> torch.Tensor:
        """Stand-alone rotation of [..., seq, dim] (rotary_embedding.py:146-163); the network
        never calls this -- its rotation is fused into the attention kernel -- but the reference
        exposes it, so it is kept (plain tensor math, fp32, any device) for API parity."""
        if offset != 0 or (seq_dim is not None and seq_dim != -2):
            raise NotImplementedError("only seq_dim=-2, offset=0 (the call made by video_net.py:418-419)")
        n = t.shape[-2]
        pos = torch.arange(n, device=t.device, dtype=torch.float32)
        ang = (pos[:, None] * self.freqs.float().to(t.device)[None, :]).repeat_interleave(2, dim=-1)
        rot = ang.shape[-1]
        head, tail = t[..., :rot].float(), t[..., rot:]
        pairs = head.reshape(*head.shape[:-1], rot // 2, 2)
        rotated = torch.stack((-pairs[..., 1], pairs[..., 0]), dim=-1).reshape(head.shape)
        out = head * ang.cos() + rotated * ang.sin()
        return torch.cat((out.to(t.dtype), tail), dim=-1)
