"""Print per-launch durations from an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv, sys

def main(path, filt=""):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i
            break
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    for r in rows[start + 2:]:
        if filt in r[ki]:
            print(f"{float(r[vi].replace(',', '')) / 1e3:9.1f} us  {r[gi]:>16}  {r[ki][:80]}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
