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


@pytest.mark.gpu
@pytest.mark.parametrize('hop,pads', [(256, (False, False)), (128, (True, True)), (256, (True, False))])
def test_tc_fused_logmel_kernel_vs_oracle(monkeypatch, hop, pads):
    """The complete tensor-core pipeline (kapre_b200/csrc/tc_mel.cuh, opt-in with KAPRE_B200_TC=1): both FFT stages as
    tcgen05 GEMMs, the Hann window as a 3-tap filter between them.  Its complex spectrum (debug dump), mel and dB outputs
    against the float64 oracle: 2e-6 of the item's largest value (3xTF32 split with round-to-nearest hi parts)."""
    import ctypes
    import kapre_b200 as K
    from kapre_b200 import _native
    monkeypatch.setenv('KAPRE_B200_TC', '1')
    rng = np.random.default_rng(hop)
    B, L = 3, 1024 + 256 * 41 + 13
    x = rng.uniform(-1, 1, size=(B, 1, L)).astype(np.float32)
    x[2] *= 1e-3
    kw = dict(n_fft=1024, hop_length=hop, sample_rate=22050, n_mels=128, pad_begin=pads[0], pad_end=pads[1],
              input_data_format='channels_first', output_data_format='channels_first')
    ref_spec = O.stft_layer(x, 1024, None, hop, None, pads[0], pads[1], 'channels_first', 'channels_first')[:, 0]
    T = ref_spec.shape[1]
    dbg = torch.zeros((B, T, 513), dtype=torch.complex64, device='cuda')
    _native.lib().kapre_tc_set_debug(ctypes.c_void_p(dbg.data_ptr()))
    mel = K.get_melspectrogram_layer(**kw)(torch.from_numpy(x).cuda()).cpu().numpy()
    assert _native.last_launch_info().startswith('TC tcgen05')
    spec = dbg.cpu().numpy()
    refm = O.melspectrogram_layer(x, **kw)
    for b in range(B):
        assert np.abs(spec[b] - ref_spec[b]).max() < 2e-6 * np.abs(ref_spec[b]).max()
        assert np.abs(mel[b] - refm[b]).max() < 2e-6 * np.abs(refm[b]).max()
    db = K.get_melspectrogram_layer(return_decibel=True, **kw)(torch.from_numpy(x).cuda()).cpu().numpy()
    refdb = O.melspectrogram_layer(x, return_decibel=True, **kw)
    assert np.abs(db - refdb).max() < 2e-4
    # a hamming window takes the same path (a - b cos with a = 0.54, b = 0.46)
    kwh = dict(kw, window_name='hamming_window')
    melh = K.get_melspectrogram_layer(**kwh)(torch.from_numpy(x).cuda()).cpu().numpy()
    assert _native.last_launch_info().startswith('TC tcgen05')
    refh = O.melspectrogram_layer(x, **kwh)
    for b in range(B):
        assert np.abs(melh[b] - refh[b]).max() < 2e-6 * np.abs(refh[b]).max()
