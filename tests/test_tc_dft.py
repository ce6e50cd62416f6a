"""tcgen05 DFT-stage prototype (kapre_b200/csrc/tc_dft.cuh): parity of the tensor-core stage-1 outputs with
the float64 partial DFT of the oracle.  Tolerance: 2e-6 of the largest output (3xTF32 split: fp32-grade)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

import oracle as O  # noqa: E402


def test_oracle_stage1_completes_to_the_rfft():
    """Checker of the checker: finishing the 32 x 32 split in float64 gives numpy's rfft of the frames."""
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, size=(2, 1024 + 3 * 256))
    S = O.dft_stage1_32x32(x)                                   # (i, f, n2, k1), k1 = 0..16
    n2 = np.arange(32)[:, None]
    k1 = np.arange(17)[None, :]
    Tw = S * np.exp(-2j * np.pi * n2 * k1 / 1024.0)
    k2 = np.arange(32)
    G = np.exp(-2j * np.pi * np.arange(32)[:, None] * k2[None, :] / 32.0)   # (n2, k2)
    X = np.einsum('ifbk,bc->ifkc', Tw, G)                       # X[k1 + 32 k2]
    idx = 256 * np.arange(4)[:, None] + np.arange(1024)[None, :]
    ref = np.fft.fft(x[:, idx], axis=-1)
    for k1v in range(17):
        for k2v in range(32):
            assert np.abs(X[:, :, k1v, k2v] - ref[:, :, k1v + 32 * k2v]).max() < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize('items,length', [(1, 1024), (3, 1024 + 256 * 37 + 11), (2, 110250)])
def test_tc_dft_stage1_vs_oracle(items, length):
    from kapre_b200.experimental import tc_dft
    rng = np.random.default_rng(length)
    x = rng.uniform(-1, 1, size=(items, length)).astype(np.float32)
    x[-1] *= 1e-3
    got = tc_dft.dft_stage1(torch.from_numpy(x).cuda()).cpu().numpy()      # (i, f, n2, 32)
    ref = O.dft_stage1_32x32(x)                                              # (i, f, n2, 17) complex
    exp = np.empty(got.shape, dtype=np.float64)
    exp[..., 0::2] = ref[..., :16].real
    exp[..., 1::2] = ref[..., :16].imag
    exp[..., 1] = ref[..., 16].real
    for i in range(items):
        scale = np.abs(exp[i]).max()
        assert np.abs(got[i] - exp[i]).max() < 2e-6 * scale, (i, np.abs(got[i] - exp[i]).max() / scale)
