"""Runs a few launches of one cfg2 configuration (for ncu).  MODE env: meldb | mel | mag | stft | istft | istft_big | music."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kapre_b200 as K

mode = os.environ.get('MODE', 'meldb')
n = int(os.environ.get('N', '3'))
torch.cuda.set_device(0)
B, L = 256, 110250
x = torch.rand((B, L, 1), device='cuda') * 2 - 1
if mode == 'meldb':
    layer = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128, return_decibel=True)
elif mode == 'mel':
    layer = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128)
elif mode == 'mag':
    layer = K.get_stft_magnitude_layer(n_fft=1024, hop_length=256)
elif mode == 'stft':
    layer = K.STFT(n_fft=1024, hop_length=256)
elif mode == 'music':      # n_fft 2048 log-mel: the 16-warp CTA variant
    B, L = 64, 441000
    x = torch.rand((B, L, 1), device='cuda') * 2 - 1
    layer = K.get_melspectrogram_layer(n_fft=2048, hop_length=512, sample_rate=44100, n_mels=128, return_decibel=True)
elif mode == 'cfg3':      # BASELINE configs[2]: 6-channel channels_last magnitude dB, n_fft 2048 / hop 1024
    x = torch.rand((1024, 44100, 6), device='cuda') * 2 - 1
    layer = K.get_stft_magnitude_layer(n_fft=2048, hop_length=1024, return_decibel=True, input_data_format='channels_last',
                                       output_data_format='channels_last')
elif mode == 'istft':
    stft, layer = K.get_perfectly_reconstructing_stft_istft(1024, 256, 'channels_last', 'channels_last')
    x = stft(x[:128, :16000])
elif mode == 'istft_big':
    stft, layer = K.get_perfectly_reconstructing_stft_istft(1024, 256, 'channels_last', 'channels_last')
    x = stft(x)
for _ in range(n):
    y = layer(x)
torch.cuda.synchronize()
print(mode, tuple(y.shape), K._native.last_launch_info())
