#!/usr/bin/env python
"""Algorithmic HBM bytes and FLOPs of every launch of one forward (dry run of the host sequence on CPU, B scaled),
joined with an ncu launch list: shows which launches sit far from their roofline."""
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from aero_b200 import Aero, aero_kwargs, cabi  # noqa: E402
from aero_b200.engine import AeroEngine  # noqa: E402

B_REAL = 32
HBM, TF32 = 6483.3e9, 715.75e12


class DryEngine(AeroEngine):
    def __init__(self, model):
        self.model, self.geom, self.lib = model, model.geom, None
        self._packed = self._packed_key = None
        self._bufs, self._windows, self._stats = {}, {}, None
        self.precision, self.fp32_tags = 1, ()
        self._prof, self._prof_tags = None, set()
        self._wk, self._wname = {}, {}
        self.use_graph, self._graphs = False, {}
        self.log = []

    def _require(self, x): pass
    def _stream(self): return None

    def _gemm(self, out, w, *, B, F_out, T, N, C1, a1=None, a2=None, C2=0, F_in=None, T_in=None, mode=0, kf=1, kt=1, stride_f=1,
              glu=0, residual=None, tag=None, **kw):
        tag = tag or self._wname.get(w.data_ptr(), "?")
        F_in = F_out if F_in is None else F_in
        T_in = T if T_in is None else T_in
        ntaps = kf // stride_f if mode == cabi.TAPS_CONVT else kf * kt
        K = C1 + C2
        if mode == cabi.TAPS_MIX:
            flops = 2.0 * B * T * N * C1
            byt = 4.0 * (B * C1 * T + B * N * T)
        else:
            flops = 2.0 * B * F_out * T * N * K * ntaps
            n_out = N // 2 if glu else N
            byt = 4.0 * (B * F_in * T_in * K + B * F_out * T * n_out * (2 if residual is not None else 1) + K * N * ntaps)
        self.log.append(("gemm:" + tag, flops, byt))
        return out

    def _norm_act(self, x, stats, gamma, beta, y, *, B, F_in, T, C_, groups, scope, op, F_out=None, f_off=0, residual=None, **kw):
        F_out = F_in if F_out is None else F_out
        co = C_ // 2 if op in (cabi.NA_GLU, cabi.NA_GLU_SCALE_RES) else C_
        self.log.append(("norm_act", 0.0, 4.0 * B * T * (F_out * C_ + F_out * co * (2 if residual is not None else 1))))
        return y

    def _lstm_rec(self, gin, bias_pad, whh, hout, *, rows, T, H, n_win, steps, **kw):
        self.log.append(("lstm", 2.0 * 2 * rows * n_win * steps * 4 * H * H, 4.0 * (gin.numel() + hout.numel())))

    def _attn(self, qkvd, out, *, rows, T, H, heads, **kw):
        self.log.append(("attn", 4.0 * rows * T * T * H, 4.0 * (qkvd.numel() + out.numel())))

    def _sample_norm(self, x, stats, y, affine, B, per_sample):
        self.log.append(("sample_norm", 0.0, 8.0 * B * per_sample))

    def stft_into(self, x, z, stats, **kw):
        self.log.append(("stft", 0.0, 4.0 * (x.numel() + z.numel())))

    def istft_into(self, z, y, **kw):
        self.log.append(("istft", 0.0, 4.0 * (z.numel() + y.numel())))


def main():
    m = Aero(**aero_kwargs("aero_4-16_512_64")).eval()
    eng = DryEngine(m)
    object.__setattr__(m, "_engine_obj", eng)
    m(torch.zeros(1, 1, 8000))
    log = eng.log
    times = None
    if len(sys.argv) > 1:
        lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
        rows = list(csv.DictReader(lines))[-len(log):]
        times = [float(r["Metric Value"].replace(",", "")) / 1e3 for r in rows]
        names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("void aero::", "")[:28] for r in rows]
    print(f"{'#':>3} {'op':34s} {'GFLOP':>8s} {'MB':>8s} {'t_hbm us':>9s} {'t_tc us':>8s} {'meas us':>8s} {'x roof':>6s}")
    tot_roof = tot_meas = 0.0
    for i, (tag, fl, by) in enumerate(log):
        fl, by = fl * B_REAL, by * B_REAL if not tag.startswith("gemm") else by * B_REAL
        t_h, t_c = by / HBM * 1e6, fl / TF32 * 1e6
        roof = max(t_h, t_c)
        meas = times[i] if times else float("nan")
        tot_roof += roof
        tot_meas += meas if times else 0
        print(f"{i:3d} {tag[:34]:34s} {fl/1e9:8.1f} {by/1e6:8.1f} {t_h:9.1f} {t_c:8.1f} {meas:8.1f} {meas/roof if times else 0:6.1f}"
              + (f"  {names[i]}" if times else ""))
    print(f"sum of per-launch roofline times {tot_roof/1e3:.2f} ms; measured {tot_meas/1e3:.2f} ms")


if __name__ == "__main__":
    main()
