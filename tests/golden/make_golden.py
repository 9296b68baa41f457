"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (the reference lives at /root/reference, read-only; it is not on the
GPU box, which is why the outputs are committed):

    python tests/golden/make_golden.py

The reference imports `einops_exts` (absent here); its only use is `rearrange_many`, a pure
re-layout helper, so an exact shim is installed in sys.modules before importing.
"""
import os
import sys
import types

import numpy as np
import torch
from einops import rearrange

REF = os.environ.get("CESM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    shim = types.ModuleType("einops_exts")
    shim.rearrange_many = lambda ts, pat, **kw: tuple(rearrange(t, pat, **kw) for t in ts)
    sys.modules.setdefault("einops_exts", shim)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import model as ref_model  # noqa
    import video_net as ref_video_net  # noqa
    return ref_model, ref_video_net


def to_np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def run_case(ref_model, unet_kwargs, B, K, H, W, seed, T=1000, with_sample=False):
    torch.manual_seed(seed)
    unet = ref_model.UNet(**unet_kwargs)
    diff = ref_model.Diffusion(unet, timesteps=T)
    diff.train()
    g = torch.Generator().manual_seed(seed + 1)
    x0 = torch.randn(B, 1, H, W, generator=g)
    cond = torch.randn(B, 1, K, H, W, generator=g)
    t = torch.randint(0, T, (B,), generator=g)
    noise = torch.randn(B, 1, H, W, generator=g)
    x_t, _ = diff.q_sample(x0, t, noise)
    eps = unet(x_t, cond, t)
    loss = torch.nn.functional.mse_loss(eps, noise)
    loss.backward()
    out = {"x0": x0, "cond": cond, "t": t, "noise": noise, "x_t": x_t, "eps": eps, "loss": loss}
    grads = {"grad/" + k: p.grad for k, p in unet.named_parameters() if p.grad is not None}
    # inference-shaped call: 4-D cond, F = 1 (inference.py:221-229)
    unet.eval()
    with torch.no_grad():
        cond1 = cond[:, :, K // 2]
        eps_f1 = unet(x_t, cond1, t)
    out["eps_f1"] = eps_f1
    if with_sample:
        torch.manual_seed(seed + 2)
        with torch.no_grad():
            out["sample"] = diff.sample(cond1, (B, 1, H, W), torch.device("cpu"))
    return unet, diff, out, grads


def main():
    ref_model, ref_vn = import_reference()
    torch.set_num_threads(8)

    # ---- case 1: tiny network, everything stored (weights, inputs, outputs, all gradients) ----
    kw = dict(base_ch=8, ch_mults=(1, 2), groups=4, attn_heads=2, attn_dim_head=8)
    unet, diff, out, grads = run_case(ref_model, kw, B=2, K=3, H=16, W=16, seed=0, T=6, with_sample=True)
    blob = {"sd/" + k: v for k, v in unet.state_dict().items()}
    blob.update(out)
    blob.update(grads)
    blob.update({"buf/" + k: v for k, v in diff.state_dict().items() if not k.startswith("model.")})
    np.savez_compressed(os.path.join(HERE, "tiny_unet_seed0.npz"), **to_np(blob))
    print("tiny: loss", float(out["loss"]), "params", sum(p.numel() for p in unet.parameters()))

    # ---- case 2: config/baseline architecture at 32x32; weights are NOT stored: they are the
    # torch default init under torch.manual_seed(0), which the package's modules must reproduce ----
    kw = dict(in_channels=2, out_channels=1, base_ch=64, ch_mults=(1, 2, 4), num_res_blocks=2, time_dim=124,
              groups=8, dropout=0.0, use_checkpoint=False)
    unet, diff, out, grads = run_case(ref_model, kw, B=1, K=3, H=32, W=32, seed=0)
    blob = dict(out)
    blob.update({"gradnorm/" + k[5:]: g.norm() for k, g in grads.items()})
    blob.update({"gradhead/" + k[5:]: g.flatten()[:8] for k, g in grads.items()})
    sd = unet.state_dict()
    blob.update({"wsum/" + k: v.double().sum().float() for k, v in sd.items()})
    blob["state_dict_keys"] = np.array(list(sd.keys()))
    blob["state_dict_shapes"] = np.array([",".join(map(str, v.shape)) for v in sd.values()])
    np.savez_compressed(os.path.join(HERE, "baseline_arch_seed0.npz"),
                        **{k: (v if isinstance(v, np.ndarray) else v.detach().cpu().numpy()) for k, v in blob.items()})
    print("baseline-arch: loss", float(out["loss"]), "tensors", len(sd))

    # ---- case 3: config/more_blocks architecture: state-dict layout only ----
    kw = dict(base_ch=64, ch_mults=(1, 2, 4, 8), groups=8)
    torch.manual_seed(0)
    unet = ref_model.UNet(**kw)
    sd = unet.state_dict()
    np.savez_compressed(os.path.join(HERE, "more_blocks_layout.npz"),
                        state_dict_keys=np.array(list(sd.keys())),
                        state_dict_shapes=np.array([",".join(map(str, v.shape)) for v in sd.values()]),
                        n_params=np.array(sum(p.numel() for p in unet.parameters())))
    print("more_blocks: tensors", len(sd), "params", sum(p.numel() for p in unet.parameters()))

    # ---- case 4: isolated pieces with non-trivial frame counts ----
    torch.manual_seed(3)
    rpb = ref_vn.RelativePositionBias(heads=8, max_distance=32)
    pieces = {"rpb_weight": rpb.relative_attention_bias.weight}
    for n in (1, 3, 12, 64):
        pieces[f"rpb_bias_n{n}"] = rpb(n, device="cpu")
        q = torch.arange(n)
        pieces[f"rpb_bucket_n{n}"] = ref_vn.RelativePositionBias._relative_position_bucket(
            q[None, :] - q[:, None], num_buckets=32, max_distance=32)
    from rotary_embedding import RotaryEmbedding
    rot = RotaryEmbedding(32)
    tq = torch.randn(2, 5, 8, 7, 32)
    pieces["rot_in"] = tq
    pieces["rot_out"] = rot.rotate_queries_or_keys(tq)
    pieces["rot_freqs"] = rot.freqs
    attn = ref_vn.Attention(64, heads=8, dim_head=32, rotary_emb=rot)
    xa = torch.randn(2, 9, 7, 64)
    pieces["attn_x"] = xa
    pieces["attn_qkv_w"] = attn.to_qkv.weight
    pieces["attn_out_w"] = attn.to_out.weight
    pieces["attn_y"] = attn(xa, pos_bias=rpb(7, device="cpu"))
    sla = ref_vn.SpatialLinearAttention(64, heads=8)
    xs = torch.randn(1, 64, 2, 6, 5)
    pieces["sla_x"] = xs
    pieces["sla_qkv_w"], pieces["sla_out_w"], pieces["sla_out_b"] = sla.to_qkv.weight, sla.to_out.weight, sla.to_out.bias
    pieces["sla_y"] = sla(xs)
    np.savez_compressed(os.path.join(HERE, "pieces.npz"), **to_np(pieces))
    print("pieces done")


if __name__ == "__main__":
    main()
