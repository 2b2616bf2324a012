"""torch.library registration of the path (aero_b200/ops.py): ops exist, shape inference (fake tensors) agrees with the
reference's shapes, and there is no CPU kernel behind them."""
import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from aero_b200 import Aero, aero_kwargs
from aero_b200 import ops


def test_ops_are_registered_with_fake_shapes():
    assert hasattr(torch.ops.aero_b200, "stft") and hasattr(torch.ops.aero_b200, "istft")
    assert hasattr(torch.ops.aero_b200, "generator_forward")
    m = Aero(**aero_kwargs("aero_4-16_512_64")).eval()
    h = ops.register_model(m)
    with FakeTensorMode():
        x = torch.empty(3, 2, 8000, device="cuda")
        z = torch.ops.aero_b200.stft(x, 512, 16, 128)
        assert z.shape == (3, 2, 257, 501, 2) and z.dtype == torch.float32          # reference spec.py:9-22
        y = torch.ops.aero_b200.istft(z, 64, 512, 32000)
        assert y.shape == (3, 2, 32000)
        out = torch.ops.aero_b200.generator_forward(torch.empty(5, 1, 7777, device="cuda"), h)
        assert out.shape == (5, 1, 31108)                                            # int(L * scale), aero.py:513
        out = torch.ops.aero_b200.generator_forward(torch.empty(2, 1, 8000, device="cuda"), h)
        assert out.shape == (2, 1, 32000)


def test_no_cpu_kernel_behind_the_ops():
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.aero_b200.stft(torch.zeros(1, 4000), 512, 16, 128)
