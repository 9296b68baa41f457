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
        if row.get("Metric Name", "gpu__time_duration.sum") != "gpu__time_duration.sum":
            continue  # logs that also carry dram__bytes rows (see `traffic`)
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


def traffic(path):
    """CSV from `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`: the LAST
    training step in the log (from its weight re-pack to its AdamW kernel), per kernel: launches, time,
    DRAM bytes moved and the resulting GB/s.  Also writes <path>.json with per-launch averages."""
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        d = per.setdefault(row["ID"], {"name": re.sub(r"\(.*", "", row["Kernel Name"])[:70]})
        d[row["Metric Name"]] = v
    ids = list(per)
    starts = [i for i, k in enumerate(ids) if "relayout_tiled_kernel<0>" in per[k]["name"]]
    ends = [i for i, k in enumerate(ids) if "adamw_clip_kernel" in per[k]["name"]]
    lo, hi = 0, len(ids)
    if starts and ends:   # the last COMPLETE step (a capture cut short ends inside a step)
        before = [i for i in starts if i < ends[-1]]
        if before:
            lo, hi = before[-1], ends[-1] + 1
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for k in ids[lo:hi]:
        d = per[k]
        a = agg[d["name"]]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot_t, tot_b = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values())
    print(f"# {path}: one training step = launches {lo}..{hi - 1}: {hi - lo} launches, {tot_t:.1f} us, "
          f"{tot_b / 1e9:.2f} GB of DRAM traffic (ncu-serialised, cold-cache: compare shares)")
    print(f"{'kernel':72s} {'n':>5s} {'us':>10s} {'share':>7s} {'DRAM MB':>10s} {'GB/s':>8s}")
    for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} {n:5d} {t:10.1f} {t / tot_t:7.1%} {b / 1e6:10.1f} {b / t / 1e3 if t else 0:8.0f}")
    json.dump({k: {"launches": n, "us_per_launch": t / n, "dram_bytes_per_launch": b / n} for k, (n, t, b) in agg.items()},
              open(path + ".json", "w"), indent=1)


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
    {"launches": launches, "report": report, "traffic": traffic}[sys.argv[1]](sys.argv[2])
