"""Stereo / 4-channel log-mel in both data formats: how much do the interleaved tensors cost the
single-channel-tile filterbank kernel?"""
import json
import sys

import torch

sys.path.insert(0, '.')
import kapre_b200 as K
from kapre_b200 import _native


def timed(layer, x, n=30):
    for _ in range(3):
        y = layer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        y = layer(x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = []
for B, L, C in ((256, 110250, 2), (128, 110250, 4), (512, 110250, 1)):
    row = {'B': B, 'L': L, 'C': C}
    for fmt in ('channels_last', 'channels_first'):
        shape = (B, L, C) if fmt == 'channels_last' else (B, C, L)
        x = torch.empty(shape, device='cuda').uniform_(-1, 1)
        layer = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128,
                                           return_decibel=True, input_data_format=fmt, output_data_format=fmt)
        row[fmt + '_ms'] = round(timed(layer, x), 4)
        row[fmt + '_launch'] = _native.last_launch_info()
        del x
    res.append(row)
print(json.dumps(res, indent=1))
