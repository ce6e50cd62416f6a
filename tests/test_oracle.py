"""CPU tier: pins the oracle (oracle/reference.py) against the committed golden fixtures,
independent implementations available in the container (torch.stft, torchaudio, scipy) and
analytic known-answer tests.  See oracle/__init__.py for what "pinned" can mean here."""
import json
import os

import numpy as np
import pytest

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def test_golden_checks_file_is_tight():
    checks = json.load(open(os.path.join(HERE, 'golden', 'kapre_cases_checks.json')))
    assert len(checks) >= 10
    for k, v in checks.items():
        assert v < 1e-5, (k, v)


def test_oracle_reproduces_golden(golden):
    src = golden['audio'].astype(np.float64)
    for key in golden.files:
        if key.startswith('stft_') and 'win512' not in key:
            _, n_fft, hop, wname = key.split('_', 3)
            wname = None if wname == 'default' else wname
            s = O.stft_frames(src, int(n_fft), int(n_fft), int(hop), O.get_window(wname, int(n_fft)), False)
            assert np.abs(s - golden[key]).max() < 1e-6 * max(1.0, np.abs(s).max())
    for hop in (128, 256):
        mel = O.melspectrogram_layer(src[None, :, None], n_fft=512, hop_length=hop, sample_rate=22050, n_mels=40,
                                     mel_f_max=8000)[0, :, :, 0]
        np.testing.assert_allclose(mel, golden['mel_512_%d' % hop], rtol=1e-12, atol=1e-15)


def test_stft_vs_torch_and_dft_matrix():
    torch = pytest.importorskip('torch')
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, 5000)
    for n_fft, hop in ((1000, 250), (512, 128), (1024, 256)):
        w = O.get_window(None, n_fft)
        s = O.stft_frames(x, n_fft, n_fft, hop, w, False)
        t = torch.stft(torch.from_numpy(x), n_fft, hop, n_fft, window=torch.from_numpy(w), center=False,
                       return_complex=True).numpy().T
        assert np.abs(s - t).max() < 1e-12
        # the reference's own matmul restatement (kapre/tflite_compatible_stft.py)
        assert np.abs(s - O.stft_by_dft_matrix(x, n_fft, n_fft, hop, w, False)).max() < 1e-11
    # right zero-pad (win < n_fft) and pad_end
    w = O.get_window(None, 400)
    for pad_end in (False, True):
        s = O.stft_frames(x, 1000, 400, 200, w, pad_end)
        assert s.shape[0] == O.num_frames(5000, 400, 200, pad_end)
        assert np.abs(s - O.stft_by_dft_matrix(x, 1000, 400, 200, w, pad_end)).max() < 1e-11


def test_windows_vs_scipy():
    sig = pytest.importorskip('scipy.signal')
    for W in (8, 512, 1000, 1024):
        assert np.abs(O.get_window(None, W) - sig.get_window('hann', W, fftbins=True)).max() < 1e-15
        assert np.abs(O.get_window('hamming_window', W) - sig.get_window('hamming', W, fftbins=True)).max() < 1e-15
    for W in (7, 1001):  # TF's "periodic" window of odd length is symmetric
        assert np.abs(O.get_window(None, W) - sig.get_window('hann', W, fftbins=False)).max() < 1e-15
    assert O.get_window(None, 1).tolist() == [1.0]
    with pytest.raises(NotImplementedError):
        O.get_window('wrong_window_name', 16)


def test_analytic_known_answers():
    n_fft, hop = 512, 128
    w = O.get_window(None, n_fft)
    # unit impulse at frame position n0: |X[k]| == w[n0] for every k
    x = np.zeros(2048)
    x[700] = 1.0
    s = O.stft_frames(x, n_fft, n_fft, hop, w, False)
    for t in range(s.shape[0]):
        n0 = 700 - t * hop
        expect = w[n0] if 0 <= n0 < n_fft else 0.0
        assert np.abs(np.abs(s[t]) - expect).max() < 1e-12
    # DC: X[0] = sum(w), bin-centred cosine: |X[k0]| = sum(w)/2
    s = O.stft_frames(np.ones(2048), n_fft, n_fft, hop, w, False)
    assert np.abs(s[:, 0] - w.sum()).max() < 1e-9
    k0 = 37
    s = O.stft_frames(np.cos(2 * np.pi * k0 * np.arange(2048) / n_fft), n_fft, n_fft, hop, w, False)
    assert np.abs(np.abs(s[:, k0]) - w.sum() / 2).max() < 1e-9
    # Parseval per frame
    rng = np.random.default_rng(1)
    x = rng.normal(size=4096)
    fr = O.frame_signal(x, n_fft, hop, False) * w
    s = O.stft_frames(x, n_fft, n_fft, hop, w, False)
    full = np.concatenate([s, np.conj(s[:, -2:0:-1])], axis=1)
    assert np.abs((np.abs(full) ** 2).sum(1) / n_fft - (fr ** 2).sum(1)).max() < 1e-9


def test_frame_counts():
    # tests/test_time_frequency.py:32-39
    assert O.num_frames(8000, 1000, 250, False) == (8000 - (1000 - 250)) // 250
    assert O.num_frames(8000, 1000, 250, True) == int(np.ceil(8000 / 250))
    assert O.num_frames(100, 512, 128, False) == 0


def test_mel_filterbank_vs_torchaudio():
    ta = pytest.importorskip('torchaudio')
    for sr, nf, nm, fmin, fmax, htk in ((22050, 513, 128, 0.0, None, False), (44100, 1025, 32, 200.0, 11025.0, True),
                                        (22050, 257, 40, 0.0, 8000.0, False)):
        fb = O.filterbank_mel(sr, nf, nm, fmin, fmax, htk, 'slaney')
        ref = ta.functional.melscale_fbanks(nf, fmin, fmax if fmax else sr / 2, nm, sr, norm='slaney',
                                            mel_scale='htk' if htk else 'slaney').numpy()
        assert fb.shape == (nf, nm) and fb.dtype == np.float32
        assert np.abs(fb - ref).max() < 1e-6 * np.abs(ref).max() * 10
        assert ((fb != 0).sum(axis=1) <= 2).all()  # <= 2 non-zeros per frequency row
    # norm=None keeps unit peaks, numeric norm L1-normalises each band
    fb = O.filterbank_mel(22050, 257, 20, norm=None)
    assert fb.max() <= 1.0 + 1e-6
    fb1 = O.filterbank_mel(22050, 257, 20, norm=1.0)
    np.testing.assert_allclose(np.abs(fb1).sum(0), 1.0, rtol=1e-5)


def test_decibel_literal_matrix():
    """tests/test_backend.py:15-40: per-row maximum, closed form 10*log10."""
    x = np.array([[1e-20, 1e-5, 1e-3, 5e-2], [0.3, 1.0, 20.5, 9999]])
    for dr in (80.0, 120.0):
        ref = 10 * np.log10(np.maximum(x, 1e-5))
        ref = np.maximum(ref, ref.max(axis=1, keepdims=True) - dr)
        np.testing.assert_allclose(O.magnitude_to_decibel(x, 1.0, 1e-5, dr), ref, atol=1e-12)
    # per-item maximum is over ALL non-batch axes (channels share it)
    y = O.magnitude_to_decibel(np.array([[[1.0, 1e-9]], [[1e-3, 1e-9]]]), 1.0, 1e-5, 20.0)
    np.testing.assert_allclose(y[:, 0, :], [[0.0, -20.0], [-30.0, -50.0]])
    for bad in (dict(ref_value=0.0), dict(amin=-1.0), dict(dynamic_range=0.0)):
        with pytest.raises(ValueError):
            O.magnitude_to_decibel(x, **bad)


@pytest.mark.parametrize('n_fft,hop', [(2048, 1024), (2048, 512), (2048, 256), (1024, 256)])
def test_istft_roundtrip_identity(golden, n_fft, hop):
    """tests/test_time_frequency.py:472-486 in float64: exact after trimming n_fft - hop."""
    src = golden['audio'].astype(np.float64)
    S = O.stft_layer(src[None, :, None], n_fft, n_fft, hop, 'hann_window', True, True, 'channels_last', 'channels_last')
    y = O.istft_layer(S, n_fft, n_fft, hop, 'hann_window', 'channels_last', 'channels_last')
    assert y.shape[1] == (S.shape[1] - 1) * hop + n_fft
    assert np.abs(y[0, n_fft - hop:n_fft - hop + 8000, 0] - src).max() < 1e-15


def test_cpu_port_matches_oracle():
    torch = pytest.importorskip('torch')
    from oracle.fast_cpu import MelSpectrogramCPU
    rng = np.random.default_rng(3)
    for fmt in ('channels_first', 'channels_last'):
        x = rng.uniform(-1, 1, (2, 2, 6000) if fmt == 'channels_first' else (2, 6000, 2)).astype(np.float32)
        kw = dict(n_fft=512, hop_length=128, sample_rate=16000, n_mels=40, return_decibel=True, pad_begin=True,
                  pad_end=True, input_data_format=fmt, output_data_format=fmt)
        y = MelSpectrogramCPU(**kw)(torch.from_numpy(x)).numpy()
        ref = O.melspectrogram_layer(x, **kw)
        assert y.shape == ref.shape
        assert np.abs(y - ref).max() < 5e-4


def test_concat_frequency_map_oracle():
    """kapre/time_frequency.py:707-733: the extra channel is linspace(0, 1, n_freq) on the frequency axis."""
    x = np.zeros((2, 5, 9, 3), dtype=np.float32)
    y = O.concat_frequency_map(x, 'channels_last')
    assert y.shape == (2, 5, 9, 4)
    np.testing.assert_allclose(y[1, 2, :, 3], np.linspace(0, 1, 9), atol=1e-7)
    assert y[..., 3].min() == 0.0 and y[..., 3].max() == 1.0
    y = O.concat_frequency_map(np.ones((1, 2, 4, 6), np.float32), 'channels_first')
    assert y.shape == (1, 3, 4, 6) and (y[0, :2] == 1).all()
    np.testing.assert_allclose(y[0, 2, 3], np.linspace(0, 1, 6), atol=1e-7)


def test_spec_augment_oracle_and_mask_sampler():
    """kapre/augmentation.py:205-208: width in [0, param), start in [0, limit - width): the sampler's ranges, and the oracle's
    inclusive [start, start + width] masking."""
    import kapre_b200.augmentation as A
    rng = np.random.default_rng(0)
    m = A.draw_masks(rng, 2000, 3, 7, 20)
    assert m.shape == (2000, 3, 2) and m.dtype == np.int32
    assert set(np.unique(m[..., 1])) == set(range(7))
    assert (m[..., 0] >= 0).all() and (m[..., 0] + m[..., 1] <= 19).all()
    assert m[..., 0][m[..., 1] == 0].max() == 19          # a width-0 mask may start at the last index
    x = np.ones((1, 6, 5, 1), np.float32)
    y = O.spec_augment(x, [[[1, 1]]], [[[4, 0]]], 9.0, 'channels_last')
    assert (y[0, 1:3] == 9).all() and (y[0, :, 4] == 9).all() and y[0, 0, 0, 0] == 1 and y[0, 3, 3, 0] == 1
    with pytest.raises(ValueError):
        A.draw_masks(rng, 1, 1, 30, 20)
