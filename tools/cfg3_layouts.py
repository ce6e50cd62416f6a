"""cfg3 (batch 1024, 6 ch, 44.1 kHz 1 s, n_fft 2048 hop 1024, magnitude dB) in both data formats."""
import json
import sys

import torch

sys.path.insert(0, '.')
import kapre_b200 as K
from kapre_b200 import _native


def timed(layer, x, n=20):
    for _ in range(3):
        y = layer(x)
    torch.cuda.synchronize()
    _native.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        y = layer(x)
    e1.record()
    torch.cuda.synchronize()
    _native.profile_enable(False)
    kms, kn = _native.profile_read()
    return e0.elapsed_time(e1) / n, kms / max(kn, 1), kn / n, tuple(y.shape)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    res = {}
    for fmt in ('channels_last', 'channels_first'):
        shape = (B, 44100, 6) if fmt == 'channels_last' else (B, 6, 44100)
        x = torch.empty(shape, device='cuda').uniform_(-1, 1)
        layer = K.get_stft_magnitude_layer(n_fft=2048, hop_length=1024, return_decibel=True,
                                           input_data_format=fmt, output_data_format=fmt)
        step_ms, k_ms, k_per_step, oshape = timed(layer, x)
        nbytes = x.numel() * 4 + 4 * int(torch.tensor(oshape).prod())
        res[fmt] = dict(step_ms=step_ms, kernel_ms_avg=k_ms, kernels_per_step=k_per_step, out=oshape,
                        algorithmic_GBs=nbytes / step_ms / 1e6, launch=_native.last_launch_info())
        del x
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
