"""Timing probe (developer tool): ensemble generation as ONE chain of 16 fields vs TWO chains of 8 fields replayed
concurrently on two streams (memory-bound kernels of one chain could overlap tensor-bound kernels of the other)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import BASELINE_KW  # noqa: E402
from cesm_emulator_b200.engine import SampleEngine  # noqa: E402
from cesm_emulator_b200.model import Diffusion, UNet  # noqa: E402


def make(d, B, H, W):
    e = SampleEngine(d, (B, 1, H, W))
    e.refresh_operands()
    e.cond.normal_()
    e.x.normal_()
    e.t.fill_(d.T - 1)
    for _ in range(3):
        e.step()
    torch.cuda.synchronize()
    return e


def timed(fn, n=10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n


def main():
    H, W = 192, 288
    torch.manual_seed(0)
    d = Diffusion(UNet(**BASELINE_KW), timesteps=1000).cuda().eval()
    for nf in (16, 32):
        e = make(d, nf, H, W)
        t1 = timed(e.graph.replay)
        del e
        ea, eb = make(d, nf // 2, H, W), make(d, nf // 2, H, W)
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

        def both():
            cur = torch.cuda.current_stream()
            sa.wait_stream(cur)
            sb.wait_stream(cur)
            with torch.cuda.stream(sa):
                ea.graph.replay()
            with torch.cuda.stream(sb):
                eb.graph.replay()
            cur.wait_stream(sa)
            cur.wait_stream(sb)

        t2 = timed(both)
        print(f"{nf} fields: one chain {t1:7.3f} ms/step, two chains of {nf // 2} on two streams {t2:7.3f} ms/step "
              f"({100 * (t2 / t1 - 1):+.1f} %)")
        del ea, eb


if __name__ == "__main__":
    main()
