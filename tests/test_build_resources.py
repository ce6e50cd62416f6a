"""Static resource checks on the built library (cuobjdump, no GPU needed).  Register pressure is what sets
the occupancy of the FFT kernels -- an instantiation that silently grew past 128 registers once halved the
resident warps and cost 2x -- so the limits the launch configurations assume are asserted here, together
with the SASS evidence that the sm_100a paths are in use (TMA bulk copy, mbarrier, packed fp32x2, cp.async)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'kapre_b200', '_lib', 'libkapre_b200.so')

pytestmark = pytest.mark.skipif(shutil.which('cuobjdump') is None or not os.path.exists(LIB),
                                reason='needs cuobjdump and the built library')


def _resources():
    out = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True, check=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r'Function (\S+):', line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r'REG:(\d+) STACK:(\d+)', line)
        if m and cur:
            res[cur] = (int(m.group(1)), int(m.group(2)))
            cur = None
    return res


def test_register_and_spill_budgets():
    res = _resources()
    fused = {k: v for k, v in res.items() if re.search(r'kb_stft_(kernel|mc_kernel|mcfb_kernel|kernel_w16)I', k)}
    assert len(fused) >= 4 * 6 + 4 * 4 + 4 * 2 + 2          # every (Q, mode) instantiation is in the binary
    for name, (regs, stack) in fused.items():
        assert regs <= 128, (name, regs)        # 2 CTAs x 8 warps (or 1 x 16) per SM need <= 128 registers per thread
        assert stack <= 16, (name, stack)       # no meaningful spilling
    inverse = {k: v for k, v in res.items() if 'kb_istft_kernel' in k}
    assert len(inverse) == 4
    for name, (regs, stack) in inverse.items():
        assert regs <= 168 and stack <= 64, (name, regs, stack)     # __launch_bounds__(128, 3)


def test_sm100a_instructions_present():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    assert 'sm_100a' in sass or 'SM100' in sass.upper() or 'EF_CUDA_SM100' in sass
    for mnemonic in ('UBLKCP', 'SYNCS', 'FADD2', 'FFMA2', 'FMUL2', 'LDGSTS', 'CREDUX'):
        assert mnemonic in sass, mnemonic
