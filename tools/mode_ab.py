"""Kernel time of the single-channel register kernel per output mode (cfg2-sized batch, n_fft 1024 / hop 256, mono,
channels_last); run under KAPRE_B200_PAIRED=0 / 1 for the pair-step A/B."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kapre_b200 as K
from kapre_b200 import _native


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    _native.profile_read()
    _native.profile_enable(True)
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    _native.profile_enable(False)
    ms, n = _native.profile_read()
    return ms / max(n, 1)


torch.cuda.set_device(0)
x = torch.rand((256, 110250, 1), device='cuda') * 2 - 1
rows = []
for n_fft, hop in ((1024, 256), (512, 128), (256, 64)):
    layers = {
        'complex': K.STFT(n_fft=n_fft, hop_length=hop),
        'magnitude': K.get_stft_magnitude_layer(n_fft=n_fft, hop_length=hop),
        'magnitude dB': K.get_stft_magnitude_layer(n_fft=n_fft, hop_length=hop, return_decibel=True),
        'mag + phase': K.get_stft_mag_phase(input_shape=(110250, 1), n_fft=n_fft, hop_length=hop),
    }
    for name, l in layers.items():
        ms = timeit(lambda: l(x))
        rows.append(dict(n_fft=n_fft, hop=hop, mode=name, kernel_ms=round(ms, 4), launch=_native.last_launch_info()))
print(json.dumps(dict(paired=os.environ.get('KAPRE_B200_PAIRED', '1'), rows=rows), indent=1))
