"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel / grid / block.

usage: python profiles/summarise_launches.py launches.csv "command that was profiled" > summary.txt
"""

from __future__ import annotations

import csv
import re
import sys
from collections import defaultdict


def main() -> None:
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "?")
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^(void )?otf::", "", r["Kernel Name"])
        name = re.sub(r"\(.*$", "", name)
        rows.append((name, r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3))
    total = sum(t for *_, t in rows)
    ours = sum(t for n, *_, t in rows if not n.startswith(("at::", "void at::", "void ")))
    agg: dict[tuple, list[float]] = defaultdict(list)
    for n, g, b, t in rows:
        agg[(n, g, b)].append(t)
    print(f"# ncu launch list of `{cmd}`")
    print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 ; per-launch times are cold-cache and")
    print("# serialised: compare SHARES with bench.py's stage_ms, not absolutes.")
    print(f"# {len(rows)} launches captured, {total:.1f} us total, {100 * ours / total:.1f}% of it in libotf_b200 kernels")
    print(f"{'kernel':<62} {'grid':>16} {'block':>12} {'n':>4} {'mean_us':>9} {'share':>7}")
    for (n, g, b), ts in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{n[:62]:<62} {g:>16} {b:>12} {len(ts):>4} {sum(ts) / len(ts):>9.1f} {100 * sum(ts) / total:>6.1f}%")


if __name__ == "__main__":
    main()
