"""A few launches of the mixed-radix STFT-magnitude kernel for ncu.  NFFT env (default 400)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kapre_b200 as K

n_fft = int(os.environ.get('NFFT', '400'))
torch.cuda.set_device(0)
x = torch.rand((64, 1, 160000), device='cuda') * 2 - 1
layer = K.get_stft_magnitude_layer(n_fft=n_fft, hop_length=n_fft // 4, input_data_format='channels_first',
                                   output_data_format='channels_first')
for _ in range(4):
    y = layer(x)
torch.cuda.synchronize()
print(n_fft, tuple(y.shape), K._native.last_launch_info())
