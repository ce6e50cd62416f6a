"""STFT-magnitude throughput against n_fft (register kernel for 64*{4,8,16,32}, mixed-radix Stockham otherwise)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kapre_b200 as K
from kapre_b200 import _native

torch.cuda.set_device(0)
B, L = 64, 160000
xs = [torch.rand((B, 1, L), device='cuda') * 2 - 1 for _ in range(4)]     # 4 x 41 MB
rows = []
for n_fft in (256, 400, 512, 1000, 1024, 2000, 2048, 4096, 8192, 16384, 1022):
    hop = n_fft // 4
    layer = K.get_stft_magnitude_layer(n_fft=n_fft, hop_length=hop, input_data_format='channels_first',
                                       output_data_format='channels_first')
    for _ in range(2):
        y = layer(xs[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3 if n_fft == 1022 else 10
    e0.record()
    for i in range(n):
        y = layer(xs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    frames = B * y.shape[2]
    rows.append({'n_fft': n_fft, 'hop': hop, 'ms': ms, 'frames_per_s': frames / (ms * 1e-3),
                 'samples_per_s': B * L / (ms * 1e-3), 'launch': _native.last_launch_info()})
    print(rows[-1], flush=True)
json.dump(rows, open('gpurun_out/r2_nfft_table.json', 'w'), indent=1)
