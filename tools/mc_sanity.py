"""Small multi-channel (channels_last) runs for compute-sanitizer: memcheck / racecheck targets."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import kapre_b200 as K
import oracle as O
from kapre_b200 import _native

rng = np.random.default_rng(0)
for n_fft, hop, C, L in ((2048, 1024, 6, 9000), (1024, 256, 2, 6000), (512, 128, 3, 4000), (256, 64, 5, 3000)):
    x = rng.uniform(-1, 1, size=(3, L, C)).astype(np.float32)
    for fmt_out in ('channels_last', 'channels_first'):
        layer = K.get_stft_magnitude_layer(n_fft=n_fft, hop_length=hop, return_decibel=True, pad_end=True,
                                           input_data_format='channels_last', output_data_format=fmt_out)
        y = layer(torch.from_numpy(x).cuda()).cpu().numpy()
        ref = O.stft_magnitude_layer(x, n_fft=n_fft, hop_length=hop, return_decibel=True, pad_end=True,
                                     input_data_format='channels_last', output_data_format=fmt_out)
        print(n_fft, hop, C, fmt_out, _native.last_launch_info(), 'max err dB %.2e' % np.abs(y - ref).max())
for n_fft, hop, C, L, n_mels in ((1024, 256, 2, 9000, 128), (2048, 512, 4, 12000, 96), (512, 128, 3, 5000, 40)):
    x = rng.uniform(-1, 1, size=(3, L, C)).astype(np.float32)
    for fmt_out in ('channels_last', 'channels_first'):
        kw = dict(n_fft=n_fft, hop_length=hop, sample_rate=22050, n_mels=n_mels, return_decibel=True, pad_end=True,
                  input_data_format='channels_last', output_data_format=fmt_out)
        y = K.get_melspectrogram_layer(**kw)(torch.from_numpy(x).cuda()).cpu().numpy()
        ref = O.melspectrogram_layer(x, **kw)
        print('mel', n_fft, hop, C, fmt_out, _native.last_launch_info(), 'max err dB %.2e' % np.abs(y - ref).max())
