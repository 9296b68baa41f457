"""Worker of tests/test_multigpu_gpu.py: run under `python -m torch.distributed.run --nproc-per-node N` on N GPUs.

(1) N ranks train 5 steps (2 eager + capture + 2 replays) on DIFFERENT data with the graph-captured bucketed NCCL
    all-reduce; afterwards every rank must hold bit-identical parameters (what DDP guarantees, train.py:1076).
(2) With lr = 0 the all-reduced gradient of the N per-rank batches must equal the gradient of ONE process stepping
    on the concatenated global batch, per tensor.
Rank 0 prints one JSON line.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _parity import BASELINE_KW, rel_err  # noqa: E402


def signature(flat):
    bits = flat.view(torch.int32).to(torch.int64)
    return torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=flat.device) % 8191 + 1)).sum()])


def main():
    from cesm_emulator_b200 import ops
    from cesm_emulator_b200.engine import TrainEngine
    from cesm_emulator_b200.model import Diffusion, UNet
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    solo = [dist.new_group([r]) for r in range(world)][rank]   # a 1-rank group: the single-process reference
    B, K, H, W = 1, 3, 32, 48
    kw = dict(BASELINE_KW)
    g = torch.Generator().manual_seed(100)
    x0_all = torch.randn(world * B, 1, H, W, generator=g)
    cond_all = torch.randn(world * B, 1, K, H, W, generator=g)
    t_all = torch.randint(0, 1000, (world * B,), generator=g)
    noise_all = torch.randn(world * B, 1, H, W, generator=g)
    mine = slice(rank * B, (rank + 1) * B)
    out = {"world": world}

    # ---- (1) replicas stay bit-identical through graph-replayed steps ----
    torch.manual_seed(rank)  # different initial weights per rank: the engine must broadcast rank 0's
    d = Diffusion(UNet(**kw)).to(dev)
    d.train()
    eng = TrainEngine(d, (B, 1, H, W), (B, 1, K, H, W), lr=2e-4)
    torch.manual_seed(1234 + rank)  # different (t, noise) draws per rank
    losses = [eng.step(x0_all[mine] + 0.1 * s, cond_all[mine]).item() for s in range(5)]
    assert eng.graph is not None
    sig = signature(eng.opt.p)
    sigs = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    out["ranks_in_sync"] = all(torch.equal(sigs[0], s) for s in sigs)
    out["losses_rank0"] = losses
    out["skipped_steps"] = int(eng.opt.state[5])
    for h in eng.buckets._hooks:
        h.remove()
    ops.set_grad_sink(None)
    del eng, d

    # ---- (2) all-reduced gradient == global-batch gradient ----
    def grads_of(pg, sl, batch):
        torch.manual_seed(0)
        m = Diffusion(UNet(**kw)).to(dev)
        m.train()
        e = TrainEngine(m, (batch, 1, H, W), (batch, 1, K, H, W), lr=0.0, weight_decay=0.0, max_grad_norm=None,
                        use_graph=False, process_group=pg)
        plain = m.loss
        t, nz = t_all[sl].to(dev), noise_all[sl].to(dev)
        m.loss = lambda x, c: plain(x, c, t=t, noise=nz)
        e.step(x0_all[sl], cond_all[sl])
        S = float(e.opt.loss_scale)
        gr = {k: (p.grad / S).clone() for k, p in m.named_parameters() if p.requires_grad}
        for h in e.buckets._hooks:
            h.remove()
        ops.set_grad_sink(None)
        return gr

    g_ddp = grads_of(None, mine, B)                       # default group: N ranks, mean of per-rank gradients
    g_one = grads_of(solo, slice(0, world * B), world * B)  # this rank alone on the concatenated batch
    errs = {k: rel_err(g_ddp[k], g_one[k]) for k in g_one}
    worst = max(errs, key=errs.get)
    out["grad_vs_global_batch"] = {"max": errs[worst], "worst": worst,
                                   "median": sorted(errs.values())[len(errs) // 2], "tensors": len(errs)}
    bad = sorted((k for k in errs if errs[k] > 1e-2), key=errs.get, reverse=True)
    out["bad_tensors"] = {k: [errs[k], float(g_ddp[k].norm()), float(g_one[k].norm())] for k in bad[:40]}
    out["n_bad"] = len(bad)
    # the per-rank gradient alone must NOT match (the data differ): the check above is not vacuous
    g_local = grads_of(solo, mine, B)
    out["bad_vs_half_local"] = {k: rel_err(g_ddp[k], 0.5 * g_local[k]) for k in bad[:40]}
    out["local_only_vs_global_batch_median"] = sorted(rel_err(g_local[k], g_one[k]) for k in g_one)[len(g_one) // 2]
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out), flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    os._exit(0)


if __name__ == "__main__":
    main()
