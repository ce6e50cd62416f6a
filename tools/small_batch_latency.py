"""cfg1 (batch 4, 16 kHz x 1 s, n_fft 512 hop 256, 64 mel dB): eager vs CUDA-graph replay, per-call time."""
import json
import sys
import time

import torch

sys.path.insert(0, '.')
import kapre_b200 as K

layer = K.get_melspectrogram_layer(n_fft=512, hop_length=256, sample_rate=16000, n_mels=64, return_decibel=True,
                                   input_data_format='channels_last', output_data_format='channels_last')
x = torch.empty((4, 16000, 1), device='cuda').uniform_(-1, 1)
cap = layer.capture(x)


def per_call(fn, n=2000):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


res = {'eager_us': per_call(lambda: layer(x)), 'graph_replay_us': per_call(lambda: cap.graph.replay()),
       'graph_call_with_input_copy_us': per_call(lambda: cap(x))}
res['frames_per_call'] = 4 * 61
print(json.dumps(res))
