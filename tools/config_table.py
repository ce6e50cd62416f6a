"""Times every BASELINE.json configuration on one GPU (device-resident, CUDA events) and prints a
JSON table with frames/s and the fraction of the measured HBM roofline (dev tool for DESIGN.md)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kapre_b200 as K
from kapre_b200 import _native

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')) else 6650.0


def timeit(fn, iters):
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


def main():
    torch.cuda.set_device(0)
    rows = []

    def add(name, frames, bytes_alg, ms):
        rows.append(dict(config=name, frames=frames, kernel_ms=ms, frames_per_s=frames / ms * 1e3,
                         algorithmic_MB=bytes_alg / 1e6, achieved_GBs=bytes_alg / ms / 1e6,
                         frac_of_measured_hbm=bytes_alg / ms / 1e6 / PEAK, launch=_native.last_launch_info()))

    # cfg1
    x = torch.rand((4, 16000, 1), device='cuda') * 2 - 1
    l = K.get_melspectrogram_layer(n_fft=512, hop_length=256, sample_rate=16000, n_mels=64, return_decibel=True)
    T = 61
    add('cfg1 B4 16k x1s N512 H256 M64 dB', 4 * T, 4 * (4 * ((T - 1) * 256 + 512) + 4 * T * 64), timeit(lambda: l(x), 50))
    # cfg2
    x = torch.rand((256, 110250, 1), device='cuda') * 2 - 1
    l = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128, return_decibel=True)
    T = 427
    add('cfg2 B256 22.05k x5s N1024 H256 M128 dB', 256 * T, 256 * (4 * ((T - 1) * 256 + 1024) + 4 * T * 128), timeit(lambda: l(x), 30))
    # cfg3
    x = torch.rand((1024, 44100, 6), device='cuda') * 2 - 1
    l = K.get_stft_magnitude_layer(n_fft=2048, hop_length=1024, return_decibel=True, input_data_format='channels_last',
                                   output_data_format='channels_last')
    T = 42
    add('cfg3 B1024 C6 44.1k x1s ch_last N2048 H1024 mag dB', 1024 * 6 * T,
        1024 * 6 * (4 * ((T - 1) * 1024 + 2048) + 4 * T * 1025), timeit(lambda: l(x), 10))
    del x
    # cfg4 forward + inverse
    x = torch.rand((128, 16000, 1), device='cuda') * 2 - 1
    stft, istft = K.get_perfectly_reconstructing_stft_istft(1024, 256, 'channels_last', 'channels_last')
    S = stft(x)
    T = S.shape[1]
    add('cfg4 STFT B128 16k x1s N1024 H256 pad (complex out)', 128 * T, 128 * (4 * 16000 + 8 * T * 513), timeit(lambda: stft(x), 50))
    add('cfg4 ISTFT', 128 * T, 128 * (8 * T * 513 + 4 * ((T - 1) * 256 + 1024)), timeit(lambda: istft(S), 50))
    y = istft(S)
    d = y[:, 768:768 + 16000, :] - x
    rows.append(dict(config='cfg4 round trip', mse=float((d * d).mean()), max_abs=float(d.abs().max())))
    # cfg5 shard
    x = torch.rand((1024, 160000, 1), device='cuda') * 2 - 1
    l = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=16000, n_mels=128, return_decibel=True)
    T = 622
    add('cfg5 shard B1024 16k x10s N1024 H256 M128 dB', 1024 * T, 1024 * (4 * ((T - 1) * 256 + 1024) + 4 * T * 128), timeit(lambda: l(x), 10))
    print(json.dumps(rows, indent=1))


if __name__ == '__main__':
    main()
