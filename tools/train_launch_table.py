"""ncu launch list of one adversarial training step (gpu__time_duration.sum CSV) -> markdown: per-kernel totals of ONE step (the launches
between two discriminator `adam_kernel` launches), split into the generator half and the discriminator half.
    python tools/train_launch_table.py gpurun_out/launches_train_gan_p1.csv "TF32 mode" """
import sys
from collections import OrderedDict

from launch_table import load


def table(seg, title, out):
    tot = sum(r[2] for r in seg)
    aero = sum(r[2] for r in seg if r[0].startswith("aero::"))
    out.append(f"## {title}: {len(seg)} launches, {tot/1e3:.1f} ms ({aero/1e3:.1f} ms in aero:: kernels, {(tot-aero)/1e3:.1f} ms in torch glue kernels)\n")
    agg = OrderedDict()
    for n, g, us in seg:
        x = agg.setdefault(n, [0, 0.0])
        x[0] += 1
        x[1] += us
    out.append("| kernel | launches | ms | share |\n|---|---|---|---|")
    other = [0, 0.0]
    for n, (k, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if us / tot < 0.004:
            other[0] += k
            other[1] += us
            continue
        nm = n if len(n) < 110 else n[:107] + "..."
        out.append(f"| `{nm}` | {k} | {us/1e3:.2f} | {100*us/tot:.1f}% |")
    out.append(f"| (kernels under 0.4 % each) | {other[0]} | {other[1]/1e3:.2f} | {100*other[1]/tot:.1f}% |\n")


def main():
    rows = load(sys.argv[1])
    mode = sys.argv[2] if len(sys.argv) > 2 else ""
    idx = [i for i, r in enumerate(rows) if "adam_kernel" in r[0]]
    # adam launches alternate generator / discriminator; a step = (D adam, G adam] + (G adam, D adam]; find a D adam followed by two more
    gaps = [b - a for a, b in zip(idx, idx[1:])]
    k = 0 if gaps[0] > gaps[1] else 1                     # the longer gap is the generator half
    a, b, c = idx[k] + 1, idx[k + 1] + 1, idx[k + 2] + 1
    out = [f"### {mode}\n"] if mode else []
    table(rows[a:c], "whole step", out)
    table(rows[a:b], "generator half (G forward, MR-STFT, D forward on the estimate, backward through D and G, Adam)", out)
    table(rows[b:c], "discriminator half (D forward on real + detached estimate, D backward, Adam)", out)
    print("\n".join(out))


if __name__ == "__main__":
    sys.path.insert(0, __import__("os").path.dirname(__file__))
    main()
