"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown)."""
import csv, re, sys
from collections import defaultdict

def main(path, title, command, anchor=None):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
    note = f"{len(rows)} consecutive kernel launches"
    if anchor:   # restrict to exactly one training step: from one launch of the anchor kernel to the next
        at = [i for i, (n, _) in enumerate(rows) if anchor in n]
        if len(at) >= 2:
            note = f"launches {at[0]}..{at[1] - 1} of {len(rows)} captured = exactly one training step ({at[1] - at[0]} launches, `{anchor}` to `{anchor}`)"
            rows = rows[at[0]:at[1]]
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        short = re.sub(r"^(void )?(<unnamed>|\(anonymous namespace\))::", "", name)
        short = re.sub(r"\(.*$", "", short)
        agg[short][0] += 1
        agg[short][1] += ns / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"# {title}\n\nCommand: `{command}`\n({note}; cold-cache and serialised under ncu: compare SHARES). "
          f"Total {tot / 1e3:.3f} ms.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us:.1f} | {100 * us / tot:.1f} % |")

if __name__ == "__main__":
    main(*sys.argv[1:5])
