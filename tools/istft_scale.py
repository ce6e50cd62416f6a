"""InverseSTFT kernel time at cfg4 size and at a cfg2-sized batch."""
import json
import sys

import torch

sys.path.insert(0, '.')
import kapre_b200 as K
from kapre_b200 import _native

res = []
for B, L in ((128, 16000), (256, 110250), (1024, 160000)):
    stft, istft = K.get_perfectly_reconstructing_stft_istft(1024, 256, 'channels_last', 'channels_last')
    x = torch.empty((B, L, 1), device='cuda').uniform_(-1, 1)
    X = stft(x)
    for _ in range(3):
        y = istft(X)
    torch.cuda.synchronize()
    _native.profile_read()
    _native.profile_enable(True)
    for _ in range(10):
        y = istft(X)
    torch.cuda.synchronize()
    _native.profile_enable(False)
    ms, n = _native.profile_read()
    nbytes = X.numel() * 8 + y.numel() * 4
    res.append(dict(B=B, L=L, kernel_ms=ms / n, GBs=nbytes / (ms / n) / 1e6, frac=nbytes / (ms / n) / 1e6 / 6569.6,
                    launch=_native.last_launch_info()))
    del x, X, y
print(json.dumps(res, indent=1))
