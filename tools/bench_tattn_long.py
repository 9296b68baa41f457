"""Time the long-window temporal attention kernels in isolation: python tools/bench_tattn_long.py [F ...] (B=1, 192x288)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

frames = tuple(int(a) for a in sys.argv[1:]) or (32, 64, 128)
res = bench.measure_long_window_attention(torch.device("cuda"), 192, 288, frames=frames)
for k, v in res["by_frames"].items():
    print(k, f"fwd {v['fwd_ms']:.3f} ms {v['roofline_fwd']['achieved']:.0f} GB/s ({v['roofline_fwd']['frac']:.2f}) "
             f"{v['roofline_fwd']['tensor_tflops']:.0f} TF | bwd {v['bwd_ms']:.3f} ms {v['roofline_bwd']['achieved']:.0f} GB/s "
             f"({v['roofline_bwd']['frac']:.2f}) {v['roofline_bwd']['tensor_tflops']:.0f} TF")
