#!/usr/bin/env python
"""Run the other BASELINE.json configurations once at their named sizes (sanity + timing): configs[2] 12->48 kHz B=16 x 2 s,
configs[4] 11.025->44.1 kHz stereo 10 s clips (B=2 per GPU), plus a large-batch 4->16 run."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from util import SEED, trained_like_, white_noise  # noqa: E402
from aero_b200 import Aero, aero_kwargs  # noqa: E402

for exp, B, C, L, secs in (("aero_12-48_512_128", 16, 1, 24000, 2.0), ("aero_11-44_512_64", 2, 2, 110250, 10.0),
                           ("aero_4-16_512_64", 96, 1, 8000, 2.0)):
    torch.manual_seed(SEED)
    m = Aero(**aero_kwargs(exp)).eval()
    m.load_state_dict(trained_like_(m.state_dict()))
    m = m.cuda()
    x = white_noise((B, C, L)).cuda()
    for _ in range(4):                 # warm-up: workspaces, weight packing and the CUDA-graph capture of this shape
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        y = m(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{exp}: in {tuple(x.shape)} -> out {tuple(y.shape)} finite={bool(torch.isfinite(y).all())} {ms:.2f} ms/forward "
          f"-> {B * secs / ms * 1e3:.0f} audio-s/s, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    del m, x, y
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
