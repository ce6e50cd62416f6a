"""Multi-threaded CPU port of the reference op chain (torch CPU ops, fp32) -- the CPU baseline.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the "kapre-CPU restatement"
of BASELINE.md section 2: TensorFlow and librosa are not installed anywhere in this sandbox, so
the reference itself cannot be timed; this port executes the same graph op by op the way
kapre-on-TF does -- materialised frames (tf.signal.frame), window multiply, rFFT with right
zero-pad, abs, dense matmul with the mel matrix, log10 chain with the per-item maximum -- on all
host cores.  It is validated against oracle.reference in tests/test_oracle.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import reference as R


class MelSpectrogramCPU:
    """kapre.composed.get_melspectrogram_layer (composed.py:138-261) restated with torch CPU ops."""

    def __init__(self, n_fft=2048, win_length=None, hop_length=None, window_name=None, pad_begin=False,
                 pad_end=False, sample_rate=22050, n_mels=128, mel_f_min=0.0, mel_f_max=None, mel_htk=False,
                 mel_norm='slaney', return_decibel=False, db_amin=1e-5, db_ref_value=1.0, db_dynamic_range=80.0,
                 input_data_format='default', output_data_format='default'):
        self.n_fft = n_fft
        self.win = win_length or n_fft
        self.hop = hop_length or self.win // 4
        self.pad_begin, self.pad_end = pad_begin, pad_end
        self.idf = R.resolve_data_format(input_data_format)
        self.odf = R.resolve_data_format(output_data_format)
        self.window = torch.from_numpy(R.get_window(window_name, self.win, dtype=np.float32))
        self.fb = torch.from_numpy(R.filterbank_mel(sample_rate, n_fft // 2 + 1, n_mels, mel_f_min, mel_f_max,
                                                    mel_htk, mel_norm, dtype=np.float32))
        self.return_decibel = return_decibel
        self.amin, self.ref, self.dr = db_amin, db_ref_value, db_dynamic_range

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if self.idf == R.CH_LAST:
            x = x.transpose(1, 2)                                   # time_frequency.py:164-167
        if self.pad_begin:
            x = torch.nn.functional.pad(x, (self.n_fft - self.hop, 0))   # :169-172
        L = x.shape[-1]
        T = R.num_frames(L, self.win, self.hop, self.pad_end)
        need = (T - 1) * self.hop + self.win
        if need > L:
            x = torch.nn.functional.pad(x, (0, need - L))
        frames = x.unfold(-1, self.win, self.hop)[..., :T, :].contiguous()   # tf.signal.frame materialises
        frames = frames * self.window                                        # window Mul
        spec = torch.fft.rfft(frames, n=self.n_fft, dim=-1)                  # RFFT (right zero-pad / crop)
        mag = spec.abs()                                                     # Magnitude, :351-359
        out = torch.matmul(mag, self.fb)                                     # (b, ch, t, mel) tensordot, :544
        if self.odf == R.CH_LAST:
            out = out.permute(0, 2, 3, 1).contiguous()
        if self.return_decibel:                                              # backend.py:186-192
            out = 10.0 * torch.log10(torch.clamp(out, min=self.amin)) - 10.0 * math.log10(max(self.amin, self.ref))
            mx = out.reshape(out.shape[0], -1).max(dim=1).values.reshape(-1, *([1] * (out.dim() - 1)))
            out = torch.maximum(out, mx - self.dr)
        return out
