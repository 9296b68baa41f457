"""Pure-write / pure-read / copy bandwidth of this GPU (torch kernels; run on the GPU box)."""
import torch
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
N = 1 << 30
a = torch.empty(N, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
af = a.view(torch.float32); bf = b.view(torch.float32)
print(f"memset (pure write)  {N / t(lambda: a.zero_()) / 1e12:.2f} TB/s")
print(f"fill fp32 (write)    {N / t(lambda: af.fill_(1.0)) / 1e12:.2f} TB/s")
print(f"sum fp32 (pure read) {N / t(lambda: af.sum()) / 1e12:.2f} TB/s")
print(f"copy (read+write)    {2 * N / t(lambda: b.copy_(a)) / 1e12:.2f} TB/s")
