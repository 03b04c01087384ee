"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X ...`).

    python tests/summarize_launches.py launches.csv [summary.csv] [launch_list.csv]
"""
import collections
import csv
import re
import sys


def main():
    rows = []
    with open(sys.argv[1], newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        unit = r.get("Metric Unit", "ns")
        val = float(r["Metric Value"].replace(",", ""))
        ns = val * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        name = re.sub(r"\(.*$", "", r["Kernel Name"]).strip()
        rows.append((int(r["ID"]), name, r.get("Grid Size", ""), r.get("Block Size", ""), ns))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for _, name, _, _, ns in rows:
        short = re.sub(r"<.*$", "", name)
        tot[short][0] += 1
        tot[short][1] += ns
    total = sum(v[1] for v in tot.values())
    out = ["share_pct,launches,avg_us,total_us,kernel"]
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{100 * ns / total:.1f},{n},{ns / n / 1e3:.1f},{ns / 1e3:.1f},{k[:90].replace(',', ';')}")
    text = f"# {len(rows)} launches, {total / 1e6:.2f} ms of serialised kernel time\n" + "\n".join(out)
    print(text)
    if len(sys.argv) > 2:
        with open(sys.argv[2], "w") as f:
            f.write(text + "\n")
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as f:
            f.write("id,kernel,grid,block,duration_ns\n")
            for i, name, grid, block, ns in rows:
                f.write(f'{i},{name[:120].replace(",", ";")},"{grid}","{block}",{ns:.0f}\n')


if __name__ == "__main__":
    main()
