"""``torch.library`` custom ops over the C ABI ("torch.library.custom_op + ctypes shim ...
register_fake for shape inference").

Importing this module registers, for the CUDA dispatch key only (there is no CPU kernel: a CPU tensor fails in the
dispatcher), the ops

  * ``aero_b200::stft(x, n_fft, hop, win) -> float32 [..., n_fft/2+1, 1+L//hop, 2]``   (``aero_stft_fwd``; reference spec.py:9-22)
  * ``aero_b200::istft(z, hop, win, length) -> float32 [..., length]``                  (``aero_istft_fwd``; spec.py:25-38)
  * ``aero_b200::generator_forward(mix, handle) -> float32 [B, C_out, L*scale]``        (the whole launch sequence of
    ``Aero.forward`` for the model registered under ``handle``; aero.py:446-523)

each with a fake (meta) implementation, so the path can sit inside ``torch.compile`` / ``torch.export`` graphs as opaque
nodes with known output shapes.  ``Aero.forward`` itself keeps calling the engine directly; these ops are the registration
a framework integrator asks for.  Inference only: no autograd formulas are registered (training kernels are SURVEY.md 8f rank 1).
"""
from __future__ import annotations

import weakref

import torch

from . import spec

_models = weakref.WeakValueDictionary()


def register_model(model):
    """Handle under which ``generator_forward`` finds ``model`` (an ``aero_b200.Aero``)."""
    h = id(model)
    _models[h] = model
    return h


@torch.library.custom_op("aero_b200::stft", mutates_args=(), device_types="cuda")
def stft(x: torch.Tensor, n_fft: int, hop: int, win: int) -> torch.Tensor:
    return torch.view_as_real(spec.spectro(x, n_fft, hop, win_length=win)).contiguous()


@stft.register_fake
def _(x, n_fft, hop, win):
    return x.new_empty((*x.shape[:-1], n_fft // 2 + 1, 1 + x.shape[-1] // hop, 2), dtype=torch.float32)


@torch.library.custom_op("aero_b200::istft", mutates_args=(), device_types="cuda")
def istft(z: torch.Tensor, hop: int, win: int, length: int) -> torch.Tensor:
    return spec.ispectro(torch.view_as_complex(z.contiguous()), hop, length=length, win_length=win)


@istft.register_fake
def _(z, hop, win, length):
    return z.new_empty((*z.shape[:-3], length), dtype=torch.float32)


@torch.library.custom_op("aero_b200::generator_forward", mutates_args=(), device_types="cuda")
def generator_forward(mix: torch.Tensor, handle: int) -> torch.Tensor:
    model = _models.get(handle)
    if model is None:
        raise RuntimeError("aero_b200::generator_forward: unknown model handle (aero_b200.ops.register_model)")
    return model(mix)


@generator_forward.register_fake
def _(mix, handle):
    model = _models.get(handle)
    if model is None:
        raise RuntimeError("aero_b200::generator_forward: unknown model handle (aero_b200.ops.register_model)")
    g = model.geom
    length = mix.shape[-1]
    padded = length + (-length) % g.hop_in
    frames = 1 + padded // g.hop_in
    out_len = min(int(length * g.scale), g.hop_out * (frames - 1))
    return mix.new_empty((mix.shape[0], g.kw["out_channels"], out_len), dtype=torch.float32)
