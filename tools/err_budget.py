#!/usr/bin/env python
"""Where does the TF32 error come from?  End-to-end rel-L2 vs the reference golden vector with groups of
tap-GEMMs forced back to exact fp32 (engine.fp32_tags)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from util import SEED, rel_l2, trained_like_, white_noise  # noqa: E402
from aero_b200 import Aero, aero_kwargs  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "c1_4-16_hop64_b2.npz"))
torch.manual_seed(SEED)
m = Aero(**aero_kwargs(str(g["exp"]))).eval()
m.load_state_dict(trained_like_(m.state_dict()))
m = m.cuda()
mix = white_noise((int(g["B"]), 1, int(g["L"]))).cuda()
eng = m._engine()
m(mix)
names = sorted(set(eng._wname.values()))
groups = {
    "none (all tf32)": (),
    "all encoders fp32": ("encoder.",),
    "all decoders fp32": ("decoder.",),
    "enc0 fp32": ("encoder.0",), "enc1 fp32": ("encoder.1",), "enc2 fp32": ("encoder.2",), "enc3 fp32": ("encoder.3",),
    "dec0 fp32": ("decoder.0",), "dec1 fp32": ("decoder.1",), "dec2 fp32": ("decoder.2",), "dec3 fp32": ("decoder.3",),
    "dec rewrites fp32": tuple(f"decoder.{j}.rw" for j in range(4)),
    "dec conv_tr fp32": tuple(f"decoder.{j}.ct" for j in range(4)),
    "ftb fp32": tuple(n for n in names if ".ftb" in n),
    "dconv fp32": tuple(n for n in names if ".dc" in n),
    "enc conv+rw fp32": tuple(n for n in names if n.endswith((".conv", ".rw")) and n.startswith("encoder")),
}
for label, tags in groups.items():
    eng.fp32_tags = tuple(tags)
    out = m(mix)
    print(f"{label:24s} rel_l2 {rel_l2(out.cpu(), g['out']):.3e}")
eng.fp32_tags = ()
eng.precision = 0
print(f"{'precision 0':24s} rel_l2 {rel_l2(m(mix).cpu(), g['out']):.3e}")
