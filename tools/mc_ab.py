"""A/B of the multi-channel tile kernel (stft_mc_core.cuh) against the single-channel kernel with
strided access (KAPRE_B200_NOMC=1) on channels_last tensors."""
import json
import os
import sys

import torch

sys.path.insert(0, '.')
import kapre_b200 as K
from kapre_b200 import _native


def timed(layer, x, n=20):
    for _ in range(3):
        y = layer(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        y = layer(x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, y


def main():
    shapes = [  # B, L, C, n_fft, hop, decibel
        (1024, 44100, 6, 2048, 1024, True),
        (512, 110250, 2, 1024, 256, True),
        (256, 160000, 2, 512, 128, False),
        (512, 48000, 4, 2048, 512, True),
        (2048, 16000, 3, 256, 64, False),
        (64, 441000, 2, 2048, 512, True),
    ]
    res = []
    for B, L, C, n_fft, hop, db in shapes:
        x = torch.empty((B, L, C), device='cuda').uniform_(-1, 1)
        layer = K.get_stft_magnitude_layer(n_fft=n_fft, hop_length=hop, return_decibel=db,
                                           input_data_format='channels_last', output_data_format='channels_last')
        row = dict(shape=[B, L, C], n_fft=n_fft, hop=hop, db=db)
        outs = {}
        for nomc in ('1', '0'):
            os.environ['KAPRE_B200_NOMC'] = nomc
            ms, y = timed(layer, x)
            key = 'single_channel_tiles' if nomc == '1' else 'multi_channel_tiles'
            nbytes = x.numel() * 4 + y.numel() * 4
            row[key] = dict(ms=round(ms, 4), GBs=round(nbytes / ms / 1e6, 1), launch=_native.last_launch_info())
            outs[nomc] = y
        row['max_abs_diff'] = float((outs['0'] - outs['1']).abs().max())
        row['speedup'] = round(row['single_channel_tiles']['ms'] / row['multi_channel_tiles']['ms'], 3)
        res.append(row)
        del x, outs, y
        torch.cuda.empty_cache()
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
