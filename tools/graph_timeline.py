"""Kernel timeline of ONE CUDA-graph replay of the training step (CUPTI through torch.profiler; run on
the GPU box): per-kernel busy time, and the idle gaps between consecutive kernels grouped by the kernel
that FOLLOWS the gap -- what a launch list under ncu (serialised, every kernel alone) cannot show."""
import collections, os, re, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cesm_emulator_b200.engine import TrainEngine
from cesm_emulator_b200.model import Diffusion, UNet
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _parity import BASELINE_KW


def main():
    B, K, H, W = 2, 3, 192, 288
    torch.manual_seed(0)
    diff = Diffusion(UNet(**BASELINE_KW), timesteps=1000).to("cuda")
    diff.train()
    if len(sys.argv) > 1 and sys.argv[1] == "sample":  # one reverse-diffusion step of a 16-field batch
        from cesm_emulator_b200.engine import SampleEngine
        diff.eval()
        for p_ in diff.parameters():
            p_.requires_grad_(False)
        eng = SampleEngine(diff, (16, 1, H, W))
        eng.cond.normal_(); eng.x.normal_(); eng.t.fill_(999)
        eng.step_resident = eng.step
    else:
        eng = TrainEngine(diff, (B, 1, H, W), (B, 1, K, H, W))
        eng.x0.normal_(); eng.cond.normal_()
    for _ in range(6):
        eng.step_resident()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        eng.step_resident()
        eng.step_resident()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    # second replay only: split at the largest gap
    gaps = [(ks[i + 1][0] - ks[i][1], i) for i in range(len(ks) - 1)]
    cut = max(gaps)[1] + 1
    ks = ks[cut:]
    short = lambda n: re.sub(r"\(.*", "", n).replace("void ", "").replace("cesm::", "")[:48]
    busy = collections.defaultdict(lambda: [0, 0.0])
    gap_by = collections.defaultdict(lambda: [0, 0.0])
    for i, (s, e, n) in enumerate(ks):
        b = busy[short(n)]; b[0] += 1; b[1] += e - s
        if i:
            g = max(0.0, s - ks[i - 1][1])
            gb = gap_by[short(n)]; gb[0] += 1; gb[1] += g
    span = ks[-1][1] - ks[0][0]
    tb, tg = sum(v[1] for v in busy.values()), sum(v[1] for v in gap_by.values())
    print(f"# one graph replay: {len(ks)} kernels, span {span:.1f} us, busy {tb:.1f} us, idle gaps {tg:.1f} us")
    print(f"{'kernel':50s} {'n':>4s} {'busy us':>9s} {'gap-before us':>14s} {'gap/launch':>10s}")
    for k, (n, t) in sorted(busy.items(), key=lambda kv: -kv[1][1]):
        g = gap_by.get(k, [0, 0.0])
        print(f"{k:50s} {n:4d} {t:9.1f} {g[1]:14.1f} {g[1] / max(1, g[0]):10.2f}")


if __name__ == "__main__":
    main()
