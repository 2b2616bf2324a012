#!/usr/bin/env python
"""Probe: one forward of B clips vs two concurrent forwards of B/2 clips on two streams (each its own engine / CUDA graph).
The LSTM recurrence is latency-bound (200 dependent steps on a fraction of the SMs); if the other half-batch's GEMMs fill the idle SMs the
pair finishes sooner than the single batch.  python tools/split_probe.py [--batch 32]"""
import argparse
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    cfg = B.CONFIGS["4-16"]
    dev = torch.device("cuda")
    m1 = B.build_model(cfg).to(dev).eval()
    m2 = copy.deepcopy(m1)
    x = torch.randn(args.batch, 1, cfg["length"], device=dev)
    xa, xb = x[: args.batch // 2].contiguous(), x[args.batch // 2:].contiguous()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def single():
        return m1(x)

    def pair():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            ya = m1(xa)
        with torch.cuda.stream(s2):
            yb = m2(xb)
        cur.wait_stream(s1); cur.wait_stream(s2)
        return ya, yb

    for name, fn in (("single", single), ("pair", pair), ("single", single), ("pair", pair)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / args.iters:.3f} ms per {args.batch} clips")
    y = single()
    ya, yb = pair()
    torch.cuda.synchronize()
    print("max abs diff pair vs single:", float((torch.cat([ya, yb]) - y).abs().max()))


if __name__ == "__main__":
    main()
