"""Host side of the tcgen05 DFT-stage prototype (kapre_b200/csrc/tc_dft.cuh) -- measured experiment only.

Stage 1 of the 32 x 32 Cooley-Tukey factorisation of the n_fft = 1024 real FFT that
``tf.signal.stft`` performs for kapre.STFT (kapre/time_frequency.py:174-182), computed as a
tensor-core GEMM straight from the hop-overlapped sample buffer.
"""
from __future__ import annotations

import ctypes

import torch

from .. import _native

N_FFT, HOP = 1024, 256


def num_frames(length: int) -> int:
    return 1 + (length - N_FFT) // HOP


def pack_matrix(F):
    """(32 n1, 32 cols) float64/32 matrix -> the (2, 8, 32, 4) hi / lo K-major operand the kernel stages."""
    import numpy as np
    F = np.asarray(F, dtype=np.float32)
    hi = (F.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
    lo = F - hi
    out = np.empty((2, 8, 32, 4), dtype=np.float32)
    for h, M in enumerate((hi, lo)):
        out[h] = M.reshape(8, 4, 32).transpose(0, 2, 1)       # [kc][col][e] = M[4 kc + e][col]
    return out


def dft_stage1(x: torch.Tensor, store: bool = True, fmat=None):
    """x: (items, length) float32 CUDA tensor.  store=True -> (items, T, 32, 32) stage-1 output
    (column 2*k1 / 2*k1+1 = Re / Im for k1 = 0..15, column 1 = the real k1 = 16 sum); store=False
    (timing) -> the per-thread checksums."""
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    items, length = x.shape
    T = num_frames(length)
    if store:
        out = torch.empty((items, T, 32, 32), dtype=torch.float32, device=x.device)
    else:
        out = torch.empty((2 * 256 * torch.cuda.get_device_properties(x.device).multi_processor_count,),
                          dtype=torch.float32, device=x.device)
    grid = ctypes.c_int(0)
    fptr = None
    if fmat is not None:
        fmat = torch.as_tensor(fmat, dtype=torch.float32).contiguous().to(x.device)
        assert fmat.numel() == 2048
        fptr = fmat.data_ptr()
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _native.check(_native.lib().kapre_tc_dft_stage1(x.data_ptr(), items, x.stride(0), length, out.data_ptr(),
                                                       1 if store else 0, ctypes.byref(grid), fptr, st))
    return out if store else out[:grid.value * 256]
