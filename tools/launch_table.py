"""ncu launch list (gpu__time_duration.sum CSV) -> markdown: per-kernel totals and the launch sequence of the LAST forward.
    python tools/launch_table.py gpurun_out/launches.csv [launches_per_forward]"""
import csv, re, sys
from collections import OrderedDict


def load(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        name = re.sub(r"\b(float|__half)\b", lambda m: {"float": "f32", "__half": "f16"}[m.group(1)], name)
        rows.append((name, r["Grid Size"], us))
    return rows


def main():
    rows = load(sys.argv[1])
    aero = [r for r in rows if r[0].startswith("aero::")]
    # one forward = from one stft kernel to the next
    starts = [i for i, r in enumerate(aero) if "stft512_kernel" in r[0] and "istft" not in r[0] or r[0].startswith("aero::stft_kernel")]
    segs = [aero[a:b] for a, b in zip(starts, starts[1:] + [len(aero)])]
    full = max(len(sg) for sg in segs)
    last = [sg for sg in segs if len(sg) == full][-1]           # the last COMPLETE forward
    tot = sum(r[2] for r in last)
    print(f"last forward: {len(last)} launches, {tot/1e3:.3f} ms\n")
    agg = OrderedDict()
    for name, grid, us in last:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {n} | {us/1e3:.3f} | {100*us/tot:.1f}% |")
    print("\n| # | kernel | grid | us |\n|---|---|---|---|")
    for i, (name, grid, us) in enumerate(last):
        print(f"| {i} | `{name}` | {grid} | {us:.1f} |")


if __name__ == "__main__":
    main()
