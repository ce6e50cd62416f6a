"""Fused log-mel kernel on a few common front-end configurations (mono, device-resident)."""
import json
import sys

import torch

sys.path.insert(0, '.')
import kapre_b200 as K
from kapre_b200 import _native

PEAK = 6569.6
cfgs = [  # name, B, L, sr, n_fft, win, hop, n_mels
    ('speech 16k win400 hop160 80mel', 1024, 160000, 16000, 512, 400, 160, 80),
    ('speech 16k n512 hop128 64mel', 1024, 160000, 16000, 512, 512, 128, 64),
    ('music 44.1k n2048 hop512 128mel', 256, 441000, 44100, 2048, 2048, 512, 128),
    ('music 22k n2048 hop512 96mel', 256, 220500, 22050, 2048, 2048, 512, 96),
    ('n256 hop64 40mel', 2048, 64000, 8000, 256, 256, 64, 40),
    ('cfg2 n1024 hop256 128mel', 256, 110250, 22050, 1024, 1024, 256, 128),
]
res = []
for name, B, L, sr, n_fft, win, hop, n_mels in cfgs:
    x = torch.empty((B, L, 1), device='cuda').uniform_(-1, 1)
    layer = K.get_melspectrogram_layer(n_fft=n_fft, win_length=win, hop_length=hop, sample_rate=sr, n_mels=n_mels,
                                       return_decibel=True, input_data_format='channels_last',
                                       output_data_format='channels_last')
    for _ in range(3):
        y = layer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = layer(x)
    e1.record()
    torch.cuda.synchronize()
    k_ms = e0.elapsed_time(e1) / 10          # per call: fused kernel + clamp launch
    frames = B * y.shape[1]
    nbytes = x.numel() * 4 + y.numel() * 4
    res.append(dict(cfg=name, frames=frames, ms_per_call=round(k_ms, 4), frames_per_s=round(frames / k_ms * 1e3),
                    GBs=round(nbytes / k_ms / 1e6, 1), frac=round(nbytes / k_ms / 1e6 / PEAK, 4),
                    ns_per_frame=round(k_ms * 1e6 / frames, 3), launch=_native.last_launch_info()))
    del x, y
print(json.dumps(res, indent=1))
