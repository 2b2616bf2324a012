"""CUDA drop-ins for the reference's ``src/models/spec.py``: ``spectro`` (9-22) and ``ispectro`` (25-38).

Same signatures and shapes: ``spectro(x[..., L]) -> complex [..., n_fft/2+1, 1+L//hop]`` (normalized, centred
reflect, periodic Hann of ``win_length`` zero-padded to ``n_fft``), ``ispectro`` its inverse.  Inputs must be
CUDA fp32 / complex64 tensors; the work is done by ``aero_stft_fwd`` / ``aero_istft_fwd`` (include/aero_b200.h)."""
from __future__ import annotations

import ctypes as C

import torch

from . import cabi

_windows = {}


def _window(win, device):
    w = _windows.get((win, device))
    if w is None:
        w = torch.hann_window(win).to(device)      # host fp32 evaluation, like reference spec.py:15
        _windows[(win, device)] = w
    return w


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"aero_b200.{what}: CUDA tensors only (no CPU fallback); got {t.device}")


@torch.no_grad()
def spectro(x, n_fft=512, hop_length=None, pad=0, win_length=None):
    _need_cuda(x, "spectro")
    lib = cabi.load()
    *other, length = x.shape
    n = n_fft * (1 + pad)
    hop = hop_length or n_fft // 4
    win = win_length or n_fft
    x2 = x.reshape(-1, length).to(torch.float32).contiguous()
    bins, frames = n // 2 + 1, 1 + length // hop
    z = torch.empty(x2.shape[0], bins, frames, 2, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        p = cabi.StftParams(n, hop, win, x2.shape[0], 1, length, frames, bins, bins * frames * 2, 0, frames * 2, 2)
        cabi.check(lib.aero_stft_fwd(C.c_void_p(x2.data_ptr()), C.c_void_p(_window(win, x.device).data_ptr()),
                                     C.c_void_p(z.data_ptr()), None, C.byref(p), _stream()), lib)
    return torch.view_as_complex(z).view(*other, bins, frames)


@torch.no_grad()
def ispectro(z, hop_length=None, length=None, pad=0, win_length=None):
    _need_cuda(z, "ispectro")
    lib = cabi.load()
    *other, bins, frames = z.shape
    n_fft = 2 * bins - 2
    hop = hop_length or n_fft // 2
    win = win_length or n_fft // (1 + pad)
    zr = torch.view_as_real(z.reshape(-1, bins, frames).to(torch.complex64).contiguous())
    full = hop * (frames - 1)
    out_len = full if length is None else min(length, full)
    y = torch.empty(zr.shape[0], out_len, dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        p = cabi.IstftParams(n_fft, hop, win, zr.shape[0], 1, frames, bins, out_len, bins * frames * 2, 0, frames * 2, 2)
        cabi.check(lib.aero_istft_fwd(C.c_void_p(zr.data_ptr()), C.c_void_p(_window(win, z.device).data_ptr()),
                                      C.c_void_p(y.data_ptr()), C.byref(p), _stream()), lib)
    if length is not None and length > full:
        y = torch.nn.functional.pad(y, (0, length - full))
    return y.view(*other, y.shape[-1])
