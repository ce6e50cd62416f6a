"""Parity of the tcgen05 fused log-mel kernel (KAPRE_B200_TC=1): complex spectrum dump + final output vs the oracle."""
import ctypes
import os
import sys

import numpy as np
import torch

os.environ['KAPRE_B200_TC'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import kapre_b200 as K
from kapre_b200 import _native

torch.cuda.set_device(0)
np.set_printoptions(linewidth=200, precision=5, suppress=True)
rng = np.random.default_rng(0)
B, L = 3, 1024 + 256 * 37 + 11
x = rng.uniform(-1, 1, size=(B, 1, L)).astype(np.float32)
x[2] *= 1e-3
kw = dict(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128, input_data_format='channels_first',
          output_data_format='channels_first')
T = 1 + (L - 1024) // 256
dbg = torch.zeros((B, T, 513), dtype=torch.complex64, device='cuda')
_native.lib().kapre_tc_set_debug(ctypes.c_void_p(dbg.data_ptr()))
mel = K.get_melspectrogram_layer(**kw)(torch.from_numpy(x).cuda())
torch.cuda.synchronize()
print('launch:', _native.last_launch_info())
spec = dbg.cpu().numpy()
ref = O.stft_layer(x, 1024, None, 256, None, False, False, 'channels_first', 'channels_first')[:, 0]   # (B, T, 513)
for b in range(B):
    sc = np.abs(ref[b]).max()
    err = np.abs(spec[b] - ref[b]) / sc
    k = np.arange(513)
    cls = {'k%32==0': k % 32 == 0, 'k%32==16': k % 32 == 16, 'k%32 in 1..15': (k % 32 >= 1) & (k % 32 <= 15),
           'k%32 in 17..31': k % 32 >= 17}
    print('item', b, 'spectrum max rel err', err.max(), {n: float(err[:, m].max()) for n, m in cls.items()},
          'worst frame', int(err.max(axis=1).argmax()), 'per-frame max', np.round(err.max(axis=1)[:20], 7))
refm = O.melspectrogram_layer(x, **kw)
got = mel.cpu().numpy()
for b in range(B):
    print('item', b, 'mel max rel err', np.abs(got[b] - refm[b]).max() / np.abs(refm[b]).max())
db = K.get_melspectrogram_layer(return_decibel=True, **kw)(torch.from_numpy(x).cuda()).cpu().numpy()
refdb = O.melspectrogram_layer(x, return_decibel=True, **kw)
print('dB max abs err', np.abs(db - refdb).max())
# pad_begin / pad_end and hop 128
kw2 = dict(kw, hop_length=128, pad_begin=True, pad_end=True)
got2 = K.get_melspectrogram_layer(**kw2)(torch.from_numpy(x).cuda()).cpu().numpy()
print('launch:', _native.last_launch_info())
ref2 = O.melspectrogram_layer(x, **kw2)
print('hop128 pad: shape', got2.shape, ref2.shape, 'max rel err', [float(np.abs(got2[b] - ref2[b]).max() / np.abs(ref2[b]).max()) for b in range(B)])
# timing
xb = [torch.rand((256, 1, 110250), device='cuda') * 2 - 1 for _ in range(3)]
layer = K.get_melspectrogram_layer(return_decibel=True, **kw)
for env in ('1', '0'):
    os.environ['KAPRE_B200_TC'] = env
    for _ in range(3):
        layer(xb[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        layer(xb[i % 3])
    e1.record()
    torch.cuda.synchronize()
    print('TC=%s: %.4f ms/step  %s' % (env, e0.elapsed_time(e1) / 20, _native.last_launch_info()))
