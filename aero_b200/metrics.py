"""GPU metrics next to the hot path.

``get_lsd(ref_sig, out_sig)`` has the call shape of reference ``src/metrics.py:59-70`` (log-spectral distance with
``STFTMag(2048, 512)``: centred reflect STFT, periodic Hann 2048, magnitude) but runs the two STFTs with
``aero_stft_fwd`` and the distance with ``aero_lsd_fwd`` on the device: no D2H copy of the waveforms and no CPU STFT
per file as in reference ``src/evaluate.py:54-97``.  (The reference's own ``STFTMag`` calls ``torch.stft`` without
``return_complex`` and raises on torch >= 2; the semantics implemented here are the intended ones.)
"""
from __future__ import annotations

import ctypes as C

import torch

from . import cabi
from .spec import spectro

LSD_NFFT, LSD_HOP = 2048, 512


@torch.no_grad()
def get_lsd(ref_sig, out_sig, n_fft=LSD_NFFT, hop=LSD_HOP):
    """ref_sig, out_sig: CUDA tensors [B, T] or [T].  Returns a 0-dim CUDA tensor (fp32)."""
    if not (ref_sig.is_cuda and out_sig.is_cuda):
        raise RuntimeError("aero_b200.metrics.get_lsd: CUDA tensors only (no CPU fallback)")
    if ref_sig.shape != out_sig.shape:
        raise ValueError(f"shape mismatch {tuple(ref_sig.shape)} vs {tuple(out_sig.shape)}")
    lib = cabi.load()
    r2 = ref_sig.reshape(-1, ref_sig.shape[-1]).float()
    o2 = out_sig.reshape(-1, out_sig.shape[-1]).float()
    zr = torch.view_as_real(spectro(r2, n_fft, hop, win_length=n_fft)).contiguous()      # [B, bins, frames, 2], x n_fft^-1/2
    ze = torch.view_as_real(spectro(o2, n_fft, hop, win_length=n_fft)).contiguous()
    B, bins, frames = zr.shape[:3]
    acc = torch.zeros(1, dtype=torch.float64, device=zr.device)
    with torch.cuda.device(zr.device):
        cabi.check(lib.aero_lsd_fwd(C.c_void_p(zr.data_ptr()), C.c_void_p(ze.data_ptr()), C.c_void_p(acc.data_ptr()),
                                    B, bins, frames, n_fft, C.c_void_p(torch.cuda.current_stream().cuda_stream)), lib)
    return (acc[0] / (B * frames)).float()
