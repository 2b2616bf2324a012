#!/usr/bin/env python
"""Single-clip latency of Aero.forward (BASELINE.json configs[0] shape: B=1, 2 s) -- eager launches vs CUDA-graph replay."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from util import SEED, rel_l2, trained_like_, white_noise  # noqa: E402
from aero_b200 import Aero, aero_kwargs  # noqa: E402

torch.manual_seed(SEED)
m = Aero(**aero_kwargs("aero_4-16_512_64")).eval()
m.load_state_dict(trained_like_(m.state_dict()))
m = m.cuda()
for B in (1, 4):
    x = white_noise((B, 1, 8000)).cuda()
    res = {}
    for mode in ("eager", "graph"):
        m.use_cuda_graph(mode == "graph")
        for _ in range(5):
            y = m(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 30
        for _ in range(n):
            y = m(x)
        torch.cuda.synchronize()
        res[mode] = ((time.perf_counter() - t0) / n * 1e3, y.clone())
    print(f"B={B}: eager {res['eager'][0]:.2f} ms/forward ({B*2/res['eager'][0]*1e3:.0f} audio-s/s), graph {res['graph'][0]:.2f} ms "
          f"({B*2/res['graph'][0]*1e3:.0f} audio-s/s); outputs rel_l2 {rel_l2(res['graph'][1].cpu(), res['eager'][1].cpu()):.1e}")
