"""CPU-side checks of the drop-in boundary: module tree, state-dict layout, initial weights,
argument validation and the C-ABI symbol table.  No compute kernels run here.

Golden data: tests/golden/*.npz (from the unmodified reference, tests/golden/make_golden.py).
"""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from _parity import BASELINE_KW, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_baseline_state_dict_and_init_match_reference():
    from cesm_emulator_b200.model import UNet
    g = load_golden("baseline_arch_seed0.npz")
    torch.manual_seed(0)
    unet = UNet(**BASELINE_KW)
    sd = unet.state_dict()
    assert list(sd.keys()) == list(g["state_dict_keys"])
    assert [",".join(map(str, v.shape)) for v in sd.values()] == list(g["state_dict_shapes"])
    assert len(sd) == 233
    # torch.manual_seed(0) + construction reproduces the reference's initial weights exactly
    for k, v in sd.items():
        ref = float(g["wsum/" + k])
        assert abs(float(v.double().sum()) - ref) <= 1e-5 * max(1.0, abs(ref)), k
    assert sum(p.numel() for p in unet.parameters()) == 10_327_889
    assert sum(p.numel() for p in unet.parameters() if p.requires_grad) == 10_327_873


def test_more_blocks_state_dict_matches_reference():
    from cesm_emulator_b200.model import UNet
    g = load_golden("more_blocks_layout.npz")
    unet = UNet(base_ch=64, ch_mults=(1, 2, 4, 8), groups=8)
    sd = unet.state_dict()
    assert list(sd.keys()) == list(g["state_dict_keys"])
    assert [",".join(map(str, v.shape)) for v in sd.values()] == list(g["state_dict_shapes"])
    assert sum(p.numel() for p in unet.parameters()) == int(g["n_params"]) == 35_186_257


def test_diffusion_buffers_and_api():
    from cesm_emulator_b200.model import Diffusion, UNet
    from oracle import cesm_oracle as O
    d = Diffusion(UNet(**BASELINE_KW), timesteps=1000)
    buf = O.diffusion_buffers(1000)
    for k, v in buf.items():
        assert torch.allclose(getattr(d, k), v, rtol=0, atol=0), k
    assert d.T == 1000 and d.img_channels == 1
    keys = [k for k in d.state_dict() if not k.startswith("model.")]
    assert keys == list(buf.keys())  # checkpoint "diffusion_buffers" layout, train.py:1157-1158
    with pytest.raises(ValueError):
        Diffusion(UNet(**BASELINE_KW), beta_schedule="cosine")
    for name in ("loss", "q_sample", "p_sample", "sample"):
        assert callable(getattr(d, name))


def test_unet_argument_validation():
    from cesm_emulator_b200.model import UNet
    u = UNet(**BASELINE_KW)
    x, c, t = torch.zeros(2, 1, 8, 8), torch.zeros(2, 1, 3, 8, 8), torch.zeros(2, dtype=torch.long)
    with pytest.raises(ValueError):
        u(x[0, 0], c, t)
    with pytest.raises(ValueError):
        u(x, None, t)
    with pytest.raises(ValueError):
        u(x, c[0, 0, 0], t)
    with pytest.raises(ValueError):
        u(x.unsqueeze(2).expand(-1, -1, 2, -1, -1), c, t)


def test_no_cpu_fallback():
    """The product path must fail loudly without CUDA instead of computing on the CPU."""
    from cesm_emulator_b200 import _lib
    from cesm_emulator_b200.model import UNet
    u = UNet(**BASELINE_KW)
    x, c, t = torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 3, 8, 8), torch.zeros(1, dtype=torch.long)
    with pytest.raises(_lib.CesmError):
        u(x, c, t)


def test_rel_pos_bucket_table_known_answers():
    from cesm_emulator_b200.video_net import RelativePositionBias
    rpb = RelativePositionBias(heads=8, max_distance=32)
    b12 = rpb.bucket_table(12, "cpu")
    assert b12[0].tolist() == [0, 17, 18, 19, 20, 21, 22, 23, 24, 24, 25, 25]
    assert b12[:, 0].tolist() == [0, 1, 2, 3, 4, 5, 6, 7, 8, 8, 9, 9]
    g = load_golden("pieces.npz")
    for n in (1, 3, 12, 64):
        assert np.array_equal(rpb.bucket_table(n, "cpu").numpy(), g[f"rpb_bucket_n{n}"])


def test_rotary_module_matches_reference():
    from cesm_emulator_b200.rotary_embedding import RotaryEmbedding
    g = load_golden("pieces.npz")
    rot = RotaryEmbedding(32)
    assert np.allclose(rot.freqs.detach().numpy(), g["rot_freqs"], rtol=0, atol=0)
    y = rot.rotate_queries_or_keys(torch.from_numpy(g["rot_in"]))
    assert np.abs(y.numpy() - g["rot_out"]).max() <= 2e-5 * np.abs(g["rot_out"]).max()
    cs, sn = rot.tables(7)
    assert cs.shape == (7, 16) and torch.all(cs[0] == 1) and torch.all(sn[0] == 0)


def test_c_abi_exports_every_declared_symbol():
    from cesm_emulator_b200 import _lib
    header = open(os.path.join(ROOT, "include", "cesm_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cesm_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed from include/cesm_b200.h"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.exported_symbols()) == declared
    lib.cesm_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.cesm_version()


def test_loads_a_checkpoint_written_by_the_reference():
    """tests/golden/ref_ckpt_tiny.pt was saved by the UNMODIFIED reference classes with train.py:1154-1165's
    layout (tests/golden/make_ref_checkpoint.py).  This repo's resume path (train.load_checkpoint, the
    train.py:915-946 semantics) must take it as is: every key lands (strict), the buffers match, the epoch
    advances, the optimizer state indexes the same parameters, and the loaded weights compute what the
    reference computed (checked through the fp32 oracle; the CUDA path needs channel counts that are
    multiples of 64, so the GPU side of checkpoint interchange is tests/test_engine_gpu.py's job)."""
    import os
    import numpy as np
    import torch
    import train as train_entry
    from cesm_emulator_b200.model import Diffusion
    from oracle import cesm_oracle as O
    gold = os.path.join(os.path.dirname(__file__), "golden")
    ckpt = torch.load(os.path.join(gold, "ref_ckpt_tiny.pt"), map_location="cpu", weights_only=False)
    assert set(ckpt) == {"epoch", "model", "diffusion_buffers", "optimizer", "config"}
    diffusion = Diffusion(train_entry.build_model_from_config(ckpt["config"]["unet"]), timesteps=1000)
    start = train_entry.load_checkpoint(os.path.join(gold, "ref_ckpt_tiny.pt"), diffusion, engine=None, device="cpu")
    assert start == ckpt["epoch"] + 1
    ours = diffusion.model.state_dict()
    assert set(ours) == set(ckpt["model"])
    for k, v in ckpt["model"].items():
        assert torch.equal(ours[k], v), k
    for k, v in ckpt["diffusion_buffers"].items():
        assert torch.equal(getattr(diffusion, k), v), k
    # optimizer: torch AdamW over this repo's module.parameters() accepts the reference's optimizer state
    params = list(diffusion.parameters())   # train.py:1078: AdamW(diffusion.parameters()), frozen rotary freqs included
    opt = torch.optim.AdamW(params, lr=2e-4)
    opt.load_state_dict(ckpt["optimizer"])
    z = np.load(os.path.join(gold, "ref_ckpt_tiny_expect.npz"))
    assert len(opt.state_dict()["state"]) == int(z["n_params"]) == len([p for p in params if p.requires_grad])
    assert 3 not in opt.state_dict()["state"] and not params[3].requires_grad  # rotary_emb.freqs holds index 3, no state
    assert torch.equal(opt.state[params[0]]["exp_avg"], torch.from_numpy(z["exp_avg_0"]))
    assert opt.state[params[0]]["exp_avg"].shape == params[0].shape
    # the loaded weights reproduce the reference's outputs
    cfg = O.OracleConfig.from_unet_kwargs(**{**ckpt["config"]["unet"], "ch_mults": tuple(ckpt["config"]["unet"]["ch_mults"])})
    sd = {k: v.float() for k, v in ours.items()}
    with torch.no_grad():
        eps = O.unet_forward(sd, cfg, torch.from_numpy(z["x_t"]), torch.from_numpy(z["cond"]), torch.from_numpy(z["t"]))
        eps1 = O.unet_forward(sd, cfg, torch.from_numpy(z["x_t"]), torch.from_numpy(z["cond"][:, :, 1]), torch.from_numpy(z["t"]))
    assert (eps - torch.from_numpy(z["eps"])).abs().max() < 2e-5 * np.abs(z["eps"]).max()
    assert (eps1 - torch.from_numpy(z["eps_f1"])).abs().max() < 2e-5 * np.abs(z["eps_f1"]).max()
