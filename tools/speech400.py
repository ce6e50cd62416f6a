"""Log-mel front ends whose n_fft is not 64 * 2^k (speech: n_fft 400 / hop 160 / 80 mel at 16 kHz; n_fft 1000 as in the
reference's tests): frames/s per call of get_melspectrogram_layer(return_decibel=True) and kernel launches per call."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kapre_b200 as K
from kapre_b200 import _native

torch.cuda.set_device(0)
rows = []
for (sr, n_fft, hop, n_mels, B, secs) in ((16000, 400, 160, 80, 256, 30), (16000, 512, 160, 80, 256, 30),
                                          (22050, 1000, 250, 128, 256, 5), (22050, 1024, 256, 128, 256, 5),
                                          (44100, 4096, 1024, 128, 64, 10)):
    L = sr * secs
    xs = [torch.rand((B, L, 1), device='cuda') * 2 - 1 for _ in range(3)]
    layer = K.get_melspectrogram_layer(input_shape=(L, 1), n_fft=n_fft, hop_length=hop, sample_rate=sr, n_mels=n_mels,
                                       return_decibel=True, input_data_format='channels_last',
                                       output_data_format='channels_last')
    for _ in range(2):
        y = layer(xs[0])
    torch.cuda.synchronize()
    l0 = _native.launch_count() if hasattr(_native, 'launch_count') else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for i in range(n):
        y = layer(xs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    l1 = _native.launch_count() if hasattr(_native, 'launch_count') else 0
    ms = e0.elapsed_time(e1) / n
    frames = B * y.shape[1]
    nbytes = B * L * 4 + y.numel() * 4
    rows.append({'sr': sr, 'n_fft': n_fft, 'hop': hop, 'n_mels': n_mels, 'batch': B, 'seconds': secs, 'ms': ms,
                 'frames_per_s': frames / (ms * 1e-3), 'algorithmic_GBs': nbytes / ms / 1e6,
                 'frac_of_hbm': nbytes / ms / 1e6 / 6569.6, 'launches_per_call': (l1 - l0) / n,
                 'launch': _native.last_launch_info()})
    print(rows[-1], flush=True)
    del xs, y
os.makedirs('gpurun_out', exist_ok=True)
json.dump(rows, open('gpurun_out/r2_speech400.json', 'w'), indent=1)
