"""Layout debugging of the tcgen05 prototype: identity / one-hot operands reveal the operand indexing."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kapre_b200.experimental import tc_dft

torch.cuda.set_device(0)
np.set_printoptions(linewidth=200, precision=4, suppress=True)
L = 1024 + 256 * 7
x = np.arange(L, dtype=np.float32)[None, :] % 4096          # x[s] = s: exactly representable, hi part only
I = np.eye(32, dtype=np.float32)
got = tc_dft.dft_stage1(torch.from_numpy(x).cuda(), fmat=tc_dft.pack_matrix(I)).cpu().numpy()   # (1, 8, 32, 32)
exp = np.empty_like(got)
for f in range(8):
    for n2 in range(32):
        for c in range(32):
            exp[0, f, n2, c] = x[0, 256 * f + 32 * c + n2]
print('identity: max abs diff', np.abs(got - exp).max(), 'nonzero frac', float((got != 0).mean()))
print('got[0,0,:4,:8]\n', got[0, 0, :4, :8])
print('exp[0,0,:4,:8]\n', exp[0, 0, :4, :8])
print('got[0,1,:4,:8]\n', got[0, 1, :4, :8])
print('got[0,5,30:,24:]\n', got[0, 5, 30:, 24:])
print('exp[0,5,30:,24:]\n', exp[0, 5, 30:, 24:])
# where does each output come from?  values are sample indices (for x[s] = s)
bad = np.argwhere(got != exp)
print('mismatches', len(bad), 'first', bad[:10].tolist())
for (i, f, n2, c) in bad[:10]:
    print((f, n2, c), 'got sample', got[i, f, n2, c], 'expected', exp[i, f, n2, c])
# B side: ones input, F[n1, c] = n1 * 32 + c  -> D[., c] = sum_n1 F[n1, c] = 32 * 496 + 32 c ... with x = one-hot on n1
x1 = np.zeros((1, L), dtype=np.float32)
x1[0, 32 * 3:32 * 4] = 1.0                                   # frame 0: n1 = 3, all n2
Fm = (np.arange(32)[:, None] * 32 + np.arange(32)[None, :]).astype(np.float32)
g2 = tc_dft.dft_stage1(torch.from_numpy(x1).cuda(), fmat=tc_dft.pack_matrix(Fm)).cpu().numpy()
print('one-hot n1=3: got[0,0,0,:8]', g2[0, 0, 0, :8], 'expected', Fm[3, :8])
print('one-hot n1=3: got[0,0,5,24:]', g2[0, 0, 5, 24:], 'expected', Fm[3, 24:])
