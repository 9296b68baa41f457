"""Timing probe (developer tool): is there concurrency to gain by running the two samples of the config/baseline
batch as two independent chains on two streams?  Compares one B=2 graph-replayed training step with two B=1 steps
(separate models) replayed concurrently on two streams.  Numbers only; nothing here is a product path."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import BASELINE_KW  # noqa: E402
from cesm_emulator_b200.engine import TrainEngine  # noqa: E402
from cesm_emulator_b200.model import Diffusion, UNet  # noqa: E402


def make(B, H, W, seed):
    torch.manual_seed(seed)
    d = Diffusion(UNet(**BASELINE_KW), timesteps=1000).cuda()
    d.train()
    e = TrainEngine(d, (B, 1, H, W), (B, 1, 3, H, W))
    e.x0.normal_()
    e.cond.normal_()
    for _ in range(4):
        e.step_resident()
    torch.cuda.synchronize()
    return e


def main():
    H, W = 192, 288
    n = 20
    e2 = make(2, H, W, 0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        e2.graph.replay()
    ev[1].record()
    torch.cuda.synchronize()
    t2 = ev[0].elapsed_time(ev[1]) / n
    print(f"one B=2 step              {t2:7.3f} ms")
    del e2
    ea, eb = make(1, H, W, 1), make(1, H, W, 2)
    ev[0].record()
    for _ in range(n):
        ea.graph.replay()
    ev[1].record()
    torch.cuda.synchronize()
    t1 = ev[0].elapsed_time(ev[1]) / n
    print(f"one B=1 step              {t1:7.3f} ms")
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(n):
        sa.wait_stream(torch.cuda.current_stream())
        sb.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(sa):
            ea.graph.replay()
        with torch.cuda.stream(sb):
            eb.graph.replay()
        torch.cuda.current_stream().wait_stream(sa)
        torch.cuda.current_stream().wait_stream(sb)
    ev[1].record()
    torch.cuda.synchronize()
    tc = ev[0].elapsed_time(ev[1]) / n
    print(f"two B=1 steps, 2 streams  {tc:7.3f} ms   (vs B=2: {100 * (tc / t2 - 1):+.1f} %)")


if __name__ == "__main__":
    main()
