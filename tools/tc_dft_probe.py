"""Parity + timing of the tcgen05 DFT stage-1 prototype on a cfg2-sized batch (dev tool; MODE=prof: few launches for ncu)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
from kapre_b200.experimental import tc_dft

torch.cuda.set_device(0)
res = {}
# parity on a small batch
rng = np.random.default_rng(1)
x = rng.uniform(-1, 1, size=(3, 1024 + 256 * 37 + 11)).astype(np.float32)
got = tc_dft.dft_stage1(torch.from_numpy(x).cuda()).cpu().numpy()
ref = O.dft_stage1_32x32(x)
exp = np.empty(got.shape)
exp[..., 0::2] = ref[..., :16].real
exp[..., 1::2] = ref[..., :16].imag
exp[..., 1] = ref[..., 16].real
res['parity_max_rel'] = float(np.abs(got - exp).max() / np.abs(exp).max())
print('parity', res['parity_max_rel'], flush=True)
if os.environ.get('MODE') == 'prof':
    xb = torch.rand((256, 110250), device='cuda') * 2 - 1
    for _ in range(3):
        tc_dft.dft_stage1(xb, store=False)
    torch.cuda.synchronize()
    sys.exit(0)

B, L = 256, 110250
xs = [torch.rand((B, L), device='cuda') * 2 - 1 for _ in range(3)]
frames = B * tc_dft.num_frames(L)
for store in (False, True):
    for _ in range(3):
        y = tc_dft.dft_stage1(xs[0], store=store)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for i in range(n):
        y = tc_dft.dft_stage1(xs[i % 3], store=store)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    # 3 split products x 2 * 128 * 32 * 32 flop per 4 frames
    flops = frames / 4 * 3 * 2 * 128 * 32 * 32
    res['store' if store else 'nostore'] = {'ms': ms, 'frames_per_s': frames / (ms * 1e-3), 'tf32_tflops': flops / (ms * 1e-3) / 1e12}
print(json.dumps(res))
