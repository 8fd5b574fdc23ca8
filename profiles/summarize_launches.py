"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel table: ms/step, launches/step, share.
    python profiles/summarize_launches.py gpurun_out/rNN_launches.csv [steps_captured]"""
import csv
import re
import sys
from collections import defaultdict


def main(path, steps=3):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for name, ns in rows:
        name = re.sub(r"<.*", "", name)
        name = re.sub(r"\(.*", "", name)
        name = name.replace("at::", "").strip()
        tot[name] += ns
        cnt[name] += 1
    total = sum(tot.values())
    print(f"{len(rows)} launches, {total / 1e6:.1f} ms of kernel time over {steps} captured steps = {total / 1e6 / steps:.2f} ms/step\n")
    print("| ms/step | launches/step | share | kernel |\n|---:|---:|---:|---|")
    for name in sorted(tot, key=lambda n: -tot[n])[:45]:
        print(f"| {tot[name] / 1e6 / steps:.3f} | {cnt[name] / steps:.0f} | {100 * tot[name] / total:.1f}% | `{name[:90]}` |")
    own = sum(v for k, v in tot.items() if "nsb::" in k)
    print(f"\nOwn kernels (nsb::*): {own / 1e6 / steps:.2f} ms/step = {100 * own / total:.1f}% of kernel time")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
