"""Host-side logic (weight packing, launch sequence, strides, LSTM windowing bookkeeping) checked on
CPU: the product's AeroEngine drives tests/cpu_emu.EmuEngine's statement of each kernel contract,
and the result must match the oracle / the reference's golden vectors."""
import os

import numpy as np
import pytest
import torch

from cpu_emu import emulated
from util import SEED, rel_l2, trained_like_, white_noise

from aero_b200 import Aero, aero_kwargs
from oracle import aero_oracle as O


def make(exp):
    torch.manual_seed(SEED)
    m = Aero(**aero_kwargs(exp)).eval()
    m.load_state_dict(trained_like_(m.state_dict()))
    return emulated(m)


@pytest.mark.parametrize("case", ["c2_4-16_hop256_ragged", "c6_4-16_hop64_short", "c5_8-24_nonpow2"])
def test_engine_sequence_matches_reference_golden(golden_dir, case):
    g = np.load(os.path.join(golden_dir, case + ".npz"))
    m = make(str(g["exp"]))
    mix = white_noise((int(g["B"]), m.in_channels, int(g["L"])))
    out, zc, zl = m(mix, return_spec=True, return_lr_spec=True)
    assert out.shape == g["out"].shape
    assert rel_l2(out, g["out"]) < 2e-5
    zc_r = torch.view_as_real(zc.contiguous()).reshape(-1)[torch.from_numpy(g["spec_idx"].astype(np.int64))]
    assert rel_l2(zc_r, g["spec_val"]) < 2e-5
    zl_r = torch.view_as_real(zl.contiguous()).reshape(-1)[torch.from_numpy(g["lrspec_idx"].astype(np.int64))]
    assert rel_l2(zl_r, g["lrspec_val"]) < 2e-5


def test_engine_windowed_lstm_and_stereo_match_oracle():
    # T > 200 exercises the overlapping-window LSTM bookkeeping; stereo exercises channel strides
    m = make("aero_11-44_512_64")
    mix = white_noise((1, 2, 3400))        # T = 1 + 3408/16 = 214 frames -> 3 windows
    with torch.no_grad():
        ref = O.aero_forward(m.state_dict(), m.geom, mix)
    out = m(mix)
    assert rel_l2(out, ref) < 2e-5
    kinds = {c[0] for c in m._engine_obj.calls}
    assert {"stft", "istft", "tapgemm", "norm_act", "lstm", "attn", "sample_norm"} <= kinds


@pytest.mark.parametrize("precision", [2, 1])
def test_storage_type_plumbing_of_the_tensor_core_engines(precision):
    """engine.precision 2 / 1 on the CPU emulation: the real host logic picks FP16 / fp32 buffers, K-major weight twins, the
    MN-major frequency mix and the fused layer-0 path; the emulation stores through those dtypes, so the FP16 activation
    rounding is real (the arithmetic is exact): the error budget of the default engine, measured without a GPU."""
    m = make("aero_4-16_512_256")
    eng = m._engine_obj
    eng.precision = precision
    mix = white_noise((2, 1, 4100))
    with torch.no_grad():
        ref = O.aero_forward(m.state_dict(), m.geom, mix)
    out = m(mix)
    err = rel_l2(out, ref)
    dts = {t.dtype for t in eng._bufs.values()}
    print(f"precision {precision}: rel_l2 {err:.2e}, buffer dtypes {sorted(str(d) for d in dts)}")
    assert (torch.float16 in dts) == (precision == 2)
    assert err < (1e-3 if precision == 2 else 2e-5)


@pytest.mark.parametrize("exp,C,L", [("aero_12-48_512_128", 1, 3000), ("aero_8-24_512_64", 1, 2100), ("aero_11-44_512_64", 2, 2300),
                                     ("aero_4-16_512_128", 1, 2050)])
def test_default_engine_plumbing_on_every_shipped_geometry(exp, C, L):
    """precision 2 on the CPU emulation for the other experiment files (odd hops, stereo, ragged lengths): buffer dtypes, row
    padding, frequency-mix variants and the fused layer 0 must line up for every geometry the YAMLs describe."""
    m = make(exp)
    m._engine_obj.precision = 2
    mix = white_noise((1, C, L))
    with torch.no_grad():
        ref = O.aero_forward(m.state_dict(), m.geom, mix)
    out = m(mix)
    assert out.shape == ref.shape
    assert rel_l2(out, ref) < 1e-3


def test_spec_roundtrip_api_shapes():
    m = make("aero_4-16_512_64")
    x = white_noise((2, 1, 4000))
    z = m._spec(x)
    assert z.shape == (2, 1, 256, 251) and z.is_complex()
    with torch.no_grad():
        assert rel_l2(torch.view_as_real(z), torch.view_as_real(O.spec(x, m.geom))) < 1e-5
    zs = m._spec(white_noise((2, 1, 16000)), scale=True)
    assert zs.shape == (2, 1, 256, 251)
    y = m._ispec(zs)
    assert y.shape == (2, 1, 64 * 250)


def test_train_mode_and_cpu_fail_loudly():
    torch.manual_seed(0)
    m = Aero(**aero_kwargs("aero_4-16_512_256"))
    with pytest.raises(RuntimeError, match="CUDA only"):
        m.eval()(torch.zeros(1, 1, 4000))
    # training mode has its own CUDA path (aero_b200/train_engine.py): a CPU tensor fails just as loudly there
    m.train()
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.zeros(1, 1, 4000))
    # ... and the INFERENCE engine never runs a model that is in training mode (it would silently use eval BatchNorm)
    with pytest.raises(NotImplementedError, match="eval"):
        emulated(m)._engine().forward(torch.zeros(1, 1, 4000))


def test_enhance_long_equals_serial_chunks():
    """Batched long-file inference == the reference's chunk-by-chunk loop (predict.py:61-80)."""
    from aero_b200.enhance import enhance_long, get_estimate
    m = make("aero_4-16_512_256")
    sig = white_noise((1, 1, 9000))[0]                     # [C=1, L]; chunks of 0.5 s = 2000 samples -> 4 full + 1000 tail
    got = enhance_long(m, sig, sr=4000, segment_sec=0.5, max_batch=3)
    serial = []
    with torch.no_grad():
        for i in range(0, 9000, 2000):
            serial.append(O.aero_forward(m.state_dict(), m.geom, sig[None, :, i:i + 2000])[0])
    ref = torch.cat(serial, -1)
    assert got.shape == ref.shape and rel_l2(got, ref) < 2e-5
    assert torch.equal(get_estimate(m, sig[None, :, :2000]), m(sig[None, :, :2000]))
