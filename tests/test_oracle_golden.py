"""Pin the CPU oracle (oracle/cesm_oracle.py) to outputs of the unmodified reference.

Fixtures: tests/golden/*.npz, written by tests/golden/make_golden.py from /root/reference.
Tolerance: both sides are fp32 on CPU and differ only in op order -> 2e-5 relative to max|ref|.
"""
import os

import numpy as np
import pytest
import torch

from oracle import cesm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 2e-5


def _load(name):
    z = np.load(os.path.join(GOLD, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _close(a, b, tol=TOL):
    a, b = _t(np.asarray(a)).double(), _t(np.asarray(b)).double()
    denom = b.abs().max().item() + 1e-30
    return (a - b).abs().max().item() / denom <= tol


@pytest.fixture(scope="module")
def tiny():
    z = _load("tiny_unet_seed0.npz")
    sd = {k[3:]: _t(v) for k, v in z.items() if k.startswith("sd/")}
    cfg = O.OracleConfig.from_unet_kwargs(base_ch=8, ch_mults=(1, 2), groups=4, attn_heads=2, attn_dim_head=8)
    return z, sd, cfg


def test_known_answer_vectors():
    # SURVEY.md section 4 item 6, derived from video_net.py:276-300 and rotary_embedding.py:96
    b12 = O.rel_pos_bucket_table(12, 32, 32)
    assert b12[0].tolist() == [0, 17, 18, 19, 20, 21, 22, 23, 24, 24, 25, 25]
    assert b12[:, 0].tolist() == [0, 1, 2, 3, 4, 5, 6, 7, 8, 8, 9, 9]
    b64 = O.rel_pos_bucket_table(64, 32, 32)
    assert b64[0, 26].item() < 31 and b64[0, 27:].eq(31).all()
    assert sorted(set(O.rel_pos_bucket_table(3, 32, 32).flatten().tolist())) == [0, 1, 2, 17, 18]
    freqs = 1.0 / (10000 ** (torch.arange(0, 32, 2).float() / 32))
    assert torch.allclose(freqs[:4], torch.tensor([1.0, 0.56234, 0.31623, 0.17783]), atol=1e-5)
    x = torch.randn(2, 4, 32)
    ang = O.rotary_angles(freqs, 4)
    y = O.apply_rotary(x, ang)
    assert torch.equal(y[:, 0], x[:, 0])  # position 0 is the identity
    c, s = math_cos_sin(freqs[0].item())
    assert torch.allclose(y[:, 1, 0], x[:, 1, 0] * c - x[:, 1, 1] * s, atol=1e-6)
    assert torch.allclose(y[:, 1, 1], x[:, 1, 1] * c + x[:, 1, 0] * s, atol=1e-6)


def math_cos_sin(a):
    import math
    return math.cos(a), math.sin(a)


def test_pieces_match_reference():
    z = _load("pieces.npz")
    cfg = O.OracleConfig()
    for n in (1, 3, 12, 64):
        assert np.array_equal(O.rel_pos_bucket_table(n, 32, 32).numpy(), z[f"rpb_bucket_n{n}"])
        assert _close(O.rel_pos_bias(_t(z["rpb_weight"]), n, cfg), z[f"rpb_bias_n{n}"])
    assert _close(O.apply_rotary(_t(z["rot_in"]), O.rotary_angles(_t(z["rot_freqs"]), 7)), z["rot_out"])
    sd = {"a.to_qkv.weight": _t(z["attn_qkv_w"]), "a.to_out.weight": _t(z["attn_out_w"]),
          "a.rotary_emb.freqs": _t(z["rot_freqs"])}
    y = O.attention(sd, "a.", _t(z["attn_x"]), 8, O.rel_pos_bias(_t(z["rpb_weight"]), 7, cfg), rotary=True)
    assert _close(y, z["attn_y"])
    sd = {"s.to_qkv.weight": _t(z["sla_qkv_w"]), "s.to_out.weight": _t(z["sla_out_w"]), "s.to_out.bias": _t(z["sla_out_b"])}
    assert _close(O.spatial_linear_attention(sd, "s.", _t(z["sla_x"]), 8), z["sla_y"])


def test_f1_attention_is_value_projection():
    # with one frame, temporal attention == to_out(v)
    torch.manual_seed(0)
    sd = {"a.to_qkv.weight": torch.randn(96, 16), "a.to_out.weight": torch.randn(16, 32),
          "a.rotary_emb.freqs": 1.0 / (10000 ** (torch.arange(0, 8, 2).float() / 8))}
    x = torch.randn(3, 5, 1, 16)
    y = O.attention(sd, "a.", x, 4, torch.randn(4, 1, 1), rotary=True)
    v = (x @ sd["a.to_qkv.weight"].t())[..., 64:]
    assert torch.allclose(y, v @ sd["a.to_out.weight"].t(), atol=1e-5)


def test_schedule_buffers(tiny):
    z, _, _ = tiny
    buf = O.diffusion_buffers(6)
    for k, v in buf.items():
        assert _close(v, z["buf/" + k], 1e-6), k
    b = O.diffusion_buffers(1000)
    assert abs(b["betas"][0].item() - 1e-4) < 1e-9 and abs(b["betas"][-1].item() - 2e-2) < 1e-8
    assert b["alphas_cumprod_prev"][0].item() == 1.0 and b["posterior_variance"][0].item() == 0.0
    with pytest.raises(ValueError):
        O.diffusion_buffers(10, "cosine")


def test_tiny_forward_loss_grads(tiny):
    z, sd, cfg = tiny
    buf = O.diffusion_buffers(6)
    x0, cond, t, noise = _t(z["x0"]), _t(z["cond"]), _t(z["t"]), _t(z["noise"])
    assert _close(O.q_sample(buf, x0, t, noise), z["x_t"])
    eps = O.unet_forward(sd, cfg, _t(z["x_t"]), cond, t)
    assert _close(eps, z["eps"])
    loss, grads = O.loss_and_grads(sd, cfg, buf, x0, cond, t, noise)
    assert abs(loss.item() - float(z["loss"])) <= 1e-5 * abs(float(z["loss"]))
    ref_grads = {k[5:]: v for k, v in z.items() if k.startswith("grad/")}
    assert set(ref_grads) == set(grads)
    bad = [k for k in grads if not _close(grads[k], ref_grads[k], 2e-4)]
    assert not bad, bad
    # inference-shaped call (4-D cond, one frame)
    assert _close(O.unet_forward(sd, cfg, _t(z["x_t"]), cond[:, :, 1], t), z["eps_f1"])


def test_tiny_sampling_chain(tiny):
    z, sd, cfg = tiny
    buf = O.diffusion_buffers(6)
    torch.manual_seed(2)
    x = O.sample(sd, cfg, buf, _t(z["cond"])[:, :, 1], (2, 1, 16, 16))
    assert _close(x, z["sample"], 1e-4)


def test_input_validation(tiny):
    z, sd, cfg = tiny
    x, c, t = _t(z["x_t"]), _t(z["cond"]), _t(z["t"])
    with pytest.raises(ValueError):
        O.unet_forward(sd, cfg, x[0, 0], c, t)
    with pytest.raises(ValueError):
        O.unet_forward(sd, cfg, x, None, t)
    with pytest.raises(ValueError):
        O.unet_forward(sd, cfg, x.unsqueeze(2).expand(-1, -1, 2, -1, -1), c, t)
