"""Summarise ncu outputs into small text files for profiles/ (run here; needs only the ncu CLI).

    python tools/summarize_ncu.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python tools/summarize_ncu.py report   gpurun_out/prof.ncu-rep  > profiles/rNN_kernel_metrics.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__m_xbar2l1tex_read_bytes.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:70]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.1f} us total (ncu-serialised, cold-cache: compare shares)")
    print(f"{'kernel':72s} {'n':>5s} {'us':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} {n:5d} {t:10.1f} {t / tot:7.1%}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    for r in rows[2:]:
        print(f"kernel: {r[idx['Kernel Name']]}  grid {r[idx.get('Grid Size', 0)]}")
        for k in KEYS:
            if k in idx:
                print(f"  {k:85s} {r[idx[k]]:>14s} {units[idx[k]]}")
        print()


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
