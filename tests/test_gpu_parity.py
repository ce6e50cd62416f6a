"""GPU parity tests: the CUDA path (through the C ABI) against the float64 oracle and the
committed golden fixtures.  Tolerances are written next to each assertion; the reference's own
tolerances (tests/test_time_frequency.py:65-69,120,256,265-267,486; tests/test_backend.py:12,40)
are asserted as well where a reference test is mirrored."""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')

import oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def K():
    assert torch.cuda.is_available(), 'GPU tier needs a CUDA device'
    import kapre_b200
    from kapre_b200 import _native
    _native.lib()  # the CUDA library must be the thing that runs: fail loudly if it is missing
    return kapre_b200


def nerr(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(float(np.abs(b).max()), 1e-30))


def wave(rng, B, C, L, fmt):
    shape = (B, C, L) if fmt == 'channels_first' else (B, L, C)
    return rng.uniform(-1, 1, size=shape).astype(np.float32)


# ------------------------------------------------------------------------------- STFT
@pytest.mark.parametrize('n_fft,win,hop', [(256, 256, 64), (512, 512, 256), (1024, 1024, 256), (2048, 2048, 1024),
                                           (2048, 2018, 1024), (1024, 400, 160), (512, 512, 125), (1000, 1000, 250),
                                           (1000, 512, 256), (96, 96, 24)])
@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
@pytest.mark.parametrize('pad', [(False, False), (True, True)])
def test_stft_complex_vs_oracle(K, n_fft, win, hop, fmt, pad):
    rng = np.random.default_rng(n_fft + hop)
    x = wave(rng, 3, 2, 6000, fmt)
    layer = K.STFT(n_fft=n_fft, win_length=win, hop_length=hop, pad_begin=pad[0], pad_end=pad[1],
                   input_data_format=fmt, output_data_format=fmt)
    got = layer(x)  # numpy in -> numpy out
    ref = O.stft_layer(x, n_fft, win, hop, None, pad[0], pad[1], fmt, fmt)
    assert got.shape == ref.shape and got.dtype == np.complex64
    assert nerr(got, ref) < 2e-6  # fp32 FFT vs float64 truth, normalised max-abs


@pytest.mark.parametrize('mode', ['complex', 'magnitude', 'mel_db'])
def test_pair_step_forms_agree(K, monkeypatch, mode):
    """n_fft 1024 carries both forms of the real-FFT pair step (paired columns: the default; natural order:
    KAPRE_B200_PAIRED=0, the A/B alternative).  Both against the oracle and against each other."""
    from kapre_b200 import _native
    rng = np.random.default_rng(7)
    x = wave(rng, 3, 1, 22050, 'channels_last')
    xt = torch.from_numpy(x).cuda()
    if mode == 'complex':
        layer = K.STFT(n_fft=1024, hop_length=256, pad_begin=True)
        ref = O.stft_layer(x, 1024, 1024, 256, None, True, False, 'channels_last', 'channels_last')
    elif mode == 'magnitude':
        layer = K.get_stft_magnitude_layer(n_fft=1024, hop_length=256)
        ref = O.stft_magnitude_layer(x, 1024, None, 256, input_data_format='channels_last', output_data_format='channels_last')
    else:
        layer = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128, return_decibel=True,
                                           db_dynamic_range=1e9)
        ref = O.melspectrogram_layer(x, n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128, return_decibel=True,
                                     db_dynamic_range=1e9)
    outs = {}
    for paired in ('1', '0'):
        monkeypatch.setenv('KAPRE_B200_PAIRED', paired)
        y = layer(xt).cpu().numpy()
        info = _native.last_launch_info()
        assert info.endswith('v0' if paired == '1' else 'v2'), info
        outs[paired] = y
    if mode == 'mel_db':
        assert np.abs(outs['1'] - ref).max() < 2e-3 and np.abs(outs['0'] - ref).max() < 2e-3      # dB
        assert np.abs(outs['1'] - outs['0']).max() < 2e-3
    else:
        assert nerr(outs['1'], ref) < 2e-6 and nerr(outs['0'], ref) < 2e-6
        assert nerr(outs['1'], outs['0']) < 5e-7


@pytest.mark.parametrize('ifmt', ['channels_first', 'channels_last'])
@pytest.mark.parametrize('ofmt', ['channels_first', 'channels_last'])
def test_stft_mixed_formats(K, ifmt, ofmt):
    rng = np.random.default_rng(5)
    x = wave(rng, 2, 3, 5000, ifmt)
    got = K.STFT(n_fft=512, hop_length=128, input_data_format=ifmt, output_data_format=ofmt)(x)
    ref = O.stft_layer(x, 512, None, 128, None, False, False, ifmt, ofmt)
    assert got.shape == ref.shape
    assert nerr(got, ref) < 2e-6


@pytest.mark.parametrize('wname', [None, 'hann_window', 'hamming_window'])
def test_stft_windows_golden(K, golden, wname):
    """Mirror of tests/test_time_frequency.py:128-185 on the reference's speech fixture."""
    src = golden['audio']
    x = np.tile(src[None, :, None], (1, 1, 2))
    got = K.STFT(n_fft=512, win_length=512, hop_length=256, window_name=wname, input_data_format='channels_last',
                 output_data_format='channels_last')(x)[0]
    ref = golden['stft_512_256_%s' % (wname or 'hann_window')]
    for ch in range(2):
        g = got[:, :, ch]
        np.testing.assert_allclose(np.abs(g), np.abs(ref), rtol=1e-5, atol=1e-3)   # reference tolerance
        np.testing.assert_allclose(g.real, ref.real, rtol=1e-5, atol=1e-3)
        np.testing.assert_allclose(g.imag, ref.imag, rtol=1e-5, atol=1e-3)
        assert nerr(g, ref) < 2e-6                                                # ours


@pytest.mark.parametrize('hop', [250, 256])
@pytest.mark.parametrize('n_ch', [1, 2, 6])
@pytest.mark.parametrize('fmt', ['default', 'channels_first', 'channels_last'])
@pytest.mark.parametrize('batch', [1, 10])
def test_spectrogram_correctness_n1000(K, golden, hop, n_ch, fmt, batch):
    """Mirror of tests/test_time_frequency.py:72-125 (n_fft=1000: the direct-DFT kernel)."""
    src = golden['audio']
    x = np.tile(src[None, :, None], (batch, 1, n_ch))
    if fmt == 'channels_first':
        x = np.ascontiguousarray(np.transpose(x, (0, 2, 1)))
    got = K.STFT(n_fft=1000, win_length=1000, hop_length=None if hop == 250 else hop, input_data_format=fmt,
                 output_data_format=fmt)(x)[0]
    ref = golden['stft_1000_%d_default' % hop]
    ref = np.tile(ref[:, :, None], (1, 1, n_ch))
    if fmt == 'channels_first':
        ref = np.transpose(ref, (2, 0, 1))
    assert got.shape == ref.shape
    np.testing.assert_allclose(np.abs(got), np.abs(ref), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(got.real, ref.real, rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(got.imag, ref.imag, rtol=1e-5, atol=1e-3)
    assert nerr(got, ref) < 5e-6
    mag = K.Sequential([K.STFT(n_fft=1000, hop_length=hop, input_data_format=fmt, output_data_format=fmt),
                        K.Magnitude()])(x)[0]
    np.testing.assert_allclose(np.abs(ref), mag, atol=2e-4)
    ph = K.Sequential([K.STFT(n_fft=1000, hop_length=hop, input_data_format=fmt, output_data_format=fmt),
                       K.Phase()])(x)[0]
    np.testing.assert_allclose(np.sin(np.angle(got)), np.sin(ph), atol=1e-3)
    np.testing.assert_allclose(np.cos(np.angle(got)), np.cos(ph), atol=1e-3)


@pytest.mark.parametrize('pad_end', [False, True])
def test_stft_short_window_golden(K, golden, pad_end):
    """win_length < n_fft is RIGHT zero-padded (tests/test_time_frequency.py:270-357)."""
    x = golden['audio'][None, :, None]
    got = K.STFT(n_fft=1000, win_length=512, hop_length=250, pad_end=pad_end, input_data_format='channels_last',
                 output_data_format='channels_last')(x)[0, :, :, 0]
    ref = golden['stft_1000_250_win512_padend%d' % pad_end]
    assert got.shape == ref.shape
    assert nerr(got, ref) < 5e-6


@pytest.mark.parametrize('n_fft,win,hop,path', [(400, 400, 160, 'MR'), (1000, 1000, 250, 'MR'), (4096, 4096, 1024, 'MR'),
                                                (8192, 8192, 2048, 'MR'), (16384, 16384, 4096, 'MR'), (75, 75, 25, 'MR'),
                                                (480, 300, 120, 'MR'), (98, 98, 49, 'DFT'), (1022, 1022, 511, 'DFT')])
@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
def test_stft_any_n_fft_mixed_radix(K, n_fft, win, hop, path, fmt):
    """Every n_fft outside 64 * {4, 8, 16, 32}: the mixed-radix Stockham kernel when n_fft / 2 (even) or n_fft (odd)
    is 5-smooth -- incl. 4096 / 8192 / 16384, which the direct-DFT kernel could not stage -- and the
    direct DFT otherwise (a prime factor above 5)."""
    rng = np.random.default_rng(n_fft)
    x = wave(rng, 2, 2, max(6000, 2 * n_fft + 1000), fmt)
    layer = K.STFT(n_fft=n_fft, win_length=win, hop_length=hop, pad_begin=True, pad_end=True,
                   input_data_format=fmt, output_data_format=fmt)
    got = layer(x)
    info = K._native.last_launch_info()
    assert info.startswith('MR ') == (path == 'MR'), info
    ref = O.stft_layer(x, n_fft, win, hop, None, True, True, fmt, fmt)
    assert got.shape == ref.shape
    assert nerr(got, ref) < 3e-6
    mag = K.get_stft_magnitude_layer(n_fft=n_fft, win_length=win, hop_length=hop, pad_begin=True, pad_end=True,
                                     input_data_format=fmt, output_data_format=fmt)(x)
    assert nerr(mag, np.abs(ref)) < 3e-6


@pytest.mark.parametrize('n_fft,hop,n_mels,sr', [(400, 160, 80, 16000), (1000, 250, 128, 22050), (4096, 1024, 128, 44100),
                                                  (480, 120, 33, 16000), (8192, 2048, 64, 44100)])
@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
def test_melspectrogram_fused_any_n_fft(K, n_fft, hop, n_mels, sr, fmt):
    """Log-mel / dB front ends at n_fft outside 64 * 2^k (speech 400 / 160 / 80; the reference's own n_fft = 1000,
    tests/test_time_frequency.py:72): ONE mixed-radix launch with the filterbank + decibel tail (+ the clamp pass),
    equal to the layer-by-layer chain and to the oracle."""
    rng = np.random.default_rng(n_fft)
    x = wave(rng, 3, 2, max(9000, 3 * n_fft), fmt)
    x[1] *= 1e-3
    kw = dict(n_fft=n_fft, hop_length=hop, sample_rate=sr, n_mels=n_mels, input_data_format=fmt, output_data_format=fmt)
    xt = torch.from_numpy(x).cuda()
    n0 = K._native.launch_count()
    mel_layer = K.get_melspectrogram_layer(**kw)
    mel = mel_layer(xt).cpu().numpy()
    # n_fft 8192: two 4096-point buffers + a 4097-bin magnitude tile exceed shared memory -> the layer chain runs instead
    fused = bool(mel_layer.layers[0].plan.supports_mode(K._native.OUT_FB))
    assert fused == (n_fft != 8192)
    if fused:
        assert K._native.launch_count() - n0 == 1
        info = K._native.last_launch_info()
        assert info.startswith('MR ') and 'frt' in info and not info.endswith('frt0'), info
    ref = O.melspectrogram_layer(x, **kw)
    assert mel.shape == ref.shape
    assert nerr(mel, ref) < 3e-6
    seq = K.get_melspectrogram_layer(**kw)
    y = xt
    for layer in seq.layers:
        y = layer(y)
    assert nerr(y.cpu().numpy(), ref) < 3e-6
    n0 = K._native.launch_count()
    db = K.get_melspectrogram_layer(return_decibel=True, **kw)(xt).cpu().numpy()
    assert K._native.launch_count() - n0 == (2 if fused else 5)     # fused kernel + clamp pass | STFT, filterbank, dB chain
    refdb = O.melspectrogram_layer(x, return_decibel=True, **kw)
    assert np.abs(db - refdb).max() < 2e-4
    np.testing.assert_allclose(refdb, db, rtol=3e-3, atol=1e-5)
    # an active clamp (40 dB) and the magnitude-dB tail
    db40 = K.get_melspectrogram_layer(return_decibel=True, db_dynamic_range=40.0, **kw)(xt).cpu().numpy()
    assert np.abs(db40 - O.melspectrogram_layer(x, return_decibel=True, db_dynamic_range=40.0, **kw)).max() < 2e-4
    mdb = K.get_stft_magnitude_layer(n_fft=n_fft, hop_length=hop, return_decibel=True, db_dynamic_range=60.0,
                                     input_data_format=fmt, output_data_format=fmt)(xt).cpu().numpy()
    assert K._native.last_launch_info().endswith('frt0')           # magnitude + dB tail: fused at every size
    refm = O.stft_magnitude_layer(x, n_fft, None, hop, return_decibel=True, db_dynamic_range=60.0, input_data_format=fmt,
                                  output_data_format=fmt)
    # bins 80 dB under the item's peak carry the fp32 FFT's leakage, not signal: compare where there is signal
    lin = O.stft_magnitude_layer(x, n_fft, None, hop, input_data_format=fmt, output_data_format=fmt)
    sig = lin > 1e-4 * lin.reshape(3, -1).max(1).reshape(3, 1, 1, 1)
    assert np.abs(mdb - refm)[sig].max() < 2e-3
    assert np.abs(mdb - refm).max() < 0.05


# ------------------------------------------------------------------------------- fused magnitude / mel / dB
@pytest.mark.parametrize('n_fft,hop,n_mels,sr', [(512, 256, 64, 16000), (1024, 256, 128, 22050), (2048, 512, 128, 44100),
                                                  (256, 128, 20, 8000), (1024, 160, 80, 16000)])
@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
@pytest.mark.parametrize('fbmma', ['0', '1'])   # filterbank phase on the CUDA cores / as the mma.sync 3xTF32 GEMM
def test_melspectrogram_fused_vs_oracle(K, monkeypatch, n_fft, hop, n_mels, sr, fmt, fbmma):
    monkeypatch.setenv('KAPRE_B200_FBMMA', fbmma)
    rng = np.random.default_rng(n_fft)
    x = wave(rng, 3, 2, 9000, fmt)
    x[1] *= 1e-3  # a quiet item: the per-item maximum must not leak between items
    kw = dict(n_fft=n_fft, hop_length=hop, sample_rate=sr, n_mels=n_mels, input_data_format=fmt, output_data_format=fmt)
    xt = torch.from_numpy(x).cuda()
    mel = K.get_melspectrogram_layer(**kw)(xt).cpu().numpy()
    if fmt == 'channels_first':   # planar tensors run the single-channel tiles, which honour the switch
        assert ('fbmma' + fbmma) in K._native.last_launch_info()
    ref = O.melspectrogram_layer(x, **kw)
    assert mel.shape == ref.shape
    assert nerr(mel, ref) < 2e-6
    # fused == layer-by-layer (stand-alone kernels) to rounding
    seq = K.get_melspectrogram_layer(**kw)
    y = xt
    for layer in seq.layers:
        y = layer(y)
    assert nerr(y.cpu().numpy(), ref) < 2e-6
    # decibel, default dynamic range (clamp inactive for the loud items, active nowhere)
    db = K.get_melspectrogram_layer(return_decibel=True, **kw)(xt).cpu().numpy()
    refdb = O.melspectrogram_layer(x, return_decibel=True, **kw)
    assert np.abs(db - refdb).max() < 2e-4  # dB, absolute (fp32 mel of ~1e-7 relative -> ~1e-5 dB; log2 approx)
    # reference tolerance, test_time_frequency.py:265-267 (atol: a dB value within 1e-5 of zero has no relative scale)
    np.testing.assert_allclose(refdb, db, rtol=3e-3, atol=1e-5)


@pytest.mark.parametrize('amin', [1e-5, 1e-3])
@pytest.mark.parametrize('dynamic_range', [120.0, 80.0])
@pytest.mark.parametrize('fmt', ['default', 'channels_first', 'channels_last'])
@pytest.mark.parametrize('hop', [128, 256])
def test_melspectrogram_correctness_golden(K, golden, amin, dynamic_range, fmt, hop):
    """Mirror of tests/test_time_frequency.py:188-267 (the reference's primary parity test)."""
    src = golden['audio']
    x = np.tile(src[None, :, None], (1, 1, 2))
    if fmt == 'channels_first':
        x = np.ascontiguousarray(np.transpose(x, (0, 2, 1)))

    def shape_ref(r):
        r = np.tile(r[:, :, None], (1, 1, 2))
        return np.transpose(r, (2, 0, 1)) if fmt == 'channels_first' else r

    kw = dict(n_fft=512, sample_rate=22050, n_mels=40, mel_f_min=0.0, mel_f_max=8000, win_length=512,
              hop_length=None if hop == 128 else hop, input_data_format=fmt, output_data_format=fmt)
    S = K.get_melspectrogram_layer(return_decibel=False, **kw).predict(x)[0]
    S_ref = shape_ref(golden['mel_512_%d' % hop])
    np.testing.assert_allclose(S_ref, S, atol=1e-4)          # reference tolerance (:256)
    assert nerr(S, S_ref) < 2e-6
    Sdb = K.get_melspectrogram_layer(return_decibel=True, db_amin=amin, db_dynamic_range=dynamic_range, **kw).predict(x)[0]
    Sdb_ref = shape_ref(golden['meldb_512_%d_amin%g_dr%g' % (hop, amin, dynamic_range)])
    np.testing.assert_allclose(Sdb_ref, Sdb, rtol=3e-3)      # reference tolerance (:265-267)
    assert np.abs(Sdb - Sdb_ref).max() < 2e-4


@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
def test_stft_magnitude_db_and_clamp(K, fmt):
    """get_stft_magnitude_layer(return_decibel=True) incl. a case where the dynamic-range clamp
    binds (full-scale sine, n_fft=2048: peak ~ +30 dB, floor -50 dB, range 60 dB)."""
    rng = np.random.default_rng(3)
    B, C, L = 3, 2, 12000
    t = np.arange(L) / 16000.0
    x = np.zeros((B, C, L), np.float32)
    x[0] = np.sin(2 * np.pi * 1000.0 * t)[None, :]            # loud tone -> clamp binds
    x[1] = 1e-4 * rng.uniform(-1, 1, size=(C, L))             # quiet -> amin floor
    x[2, 0] = rng.uniform(-1, 1, size=L)                      # one loud channel sets the item max ...
    x[2, 1] = 1e-6 * rng.uniform(-1, 1, size=L)               # ... the other one gets clamped by it
    if fmt == 'channels_last':
        x = np.ascontiguousarray(np.transpose(x, (0, 2, 1)))
    kw = dict(n_fft=2048, hop_length=512, return_decibel=True, db_dynamic_range=60.0, input_data_format=fmt,
              output_data_format=fmt)
    got = K.get_stft_magnitude_layer(**kw)(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = O.stft_magnitude_layer(x, **kw)
    assert got.shape == ref.shape
    # the clamp must actually have been exercised
    unclamped = O.stft_magnitude_layer(x, **dict(kw, db_dynamic_range=1e9))
    assert (np.abs(ref - unclamped) > 1.0).any()
    # fp32 rounding noise sits far below the clamp threshold (peak - 60 dB), so after the clamp
    # the whole tensor is comparable: 5e-3 dB absolute
    assert np.abs(got - ref).max() < 5e-3


def test_decibel_clamp_long_items_nan_and_workspace_reuse(K):
    """The clamp pass runs several CTAs per long item (here 3 and 2 chunks), keeps the workspace self-cleaning across
    calls of different batch sizes, and an item that contains a NaN comes out all-NaN while its neighbours are untouched
    (tf.maximum / tf.reduce_max propagate NaN, kapre/backend.py:186-192)."""
    rng = np.random.default_rng(5)
    L = 16000 * 4
    t = np.arange(L) / 16000.0
    kw = dict(n_fft=2048, hop_length=512, return_decibel=True, db_dynamic_range=40.0, input_data_format='channels_first',
              output_data_format='channels_first')
    layer = K.get_stft_magnitude_layer(**kw)
    for B in (3, 1, 4):                          # item = 122 frames x 1025 bins = 125 050 values -> 4 chunks
        # noise 40 dB under the tone: well above the fp32 leakage of the tone (1e-7 of its peak), partly under the clamp
        x = (1e-2 * rng.uniform(-1, 1, size=(B, 1, L))).astype(np.float32)
        x[0, 0] += np.sin(2 * np.pi * 440.0 * t)                  # loud tone: the clamp binds in item 0
        got = layer(torch.from_numpy(x).cuda()).cpu().numpy()
        ref = O.stft_magnitude_layer(x, **kw)
        assert (np.abs(ref - O.stft_magnitude_layer(x, **dict(kw, db_dynamic_range=1e9))) > 1.0).any()
        assert np.abs(got - ref).max() < 5e-3
    # fused log-mel and the stand-alone layer: NaN in one item
    x = rng.uniform(-1, 1, size=(3, 1, 9000)).astype(np.float32)
    x[1, 0, 4321] = np.nan
    mel = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, n_mels=64, return_decibel=True,
                                     input_data_format='channels_first', output_data_format='channels_first')
    y = mel(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.isnan(y[1]).all() and np.isfinite(y[0]).all() and np.isfinite(y[2]).all()
    y2 = mel(torch.from_numpy(np.nan_to_num(x)).cuda()).cpu().numpy()         # the workspace was left clean
    assert np.isfinite(y2).all() and np.array_equal(y2[0], y[0])
    m = np.abs(rng.normal(size=(3, 40, 33))).astype(np.float32)
    m[2, 7, 5] = np.nan
    d = K.backend.magnitude_to_decibel(m)
    assert np.isnan(d[2]).all() and np.isfinite(d[:2]).all()


@pytest.mark.parametrize('dynamic_range', [80.0, 120.0])
def test_magnitude_to_decibel_literal(K, golden, dynamic_range):
    """Mirror of tests/test_backend.py:15-40 (per-row maximum)."""
    x = golden['db_literal_in'].astype(np.float32)
    got = K.backend.magnitude_to_decibel(x, ref_value=1.0, amin=1e-5, dynamic_range=dynamic_range)
    np.testing.assert_allclose(got, golden['db_literal_dr%g' % dynamic_range], atol=1e-5)  # reference TOL
    got1d = K.backend.magnitude_to_decibel(x[1], ref_value=1.0, amin=1e-5, dynamic_range=dynamic_range)
    np.testing.assert_allclose(got1d, golden['db_literal_dr%g' % dynamic_range][1], atol=1e-5)
    with pytest.raises(ValueError):
        K.backend.magnitude_to_decibel(x, amin=0.0)
    with pytest.raises(ValueError):
        K.MagnitudeToDecibel(dynamic_range=-1.0)(torch.from_numpy(x).cuda())


@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
@pytest.mark.parametrize('kind', ['mel', 'log'])
def test_apply_filterbank_standalone(K, fmt, kind):
    rng = np.random.default_rng(11)
    B, C, T, F = 2, 3, 45, 513
    shape = (B, C, T, F) if fmt == 'channels_first' else (B, T, F, C)
    x = rng.uniform(0, 10, size=shape).astype(np.float32)
    if kind == 'mel':
        kwargs = {'sample_rate': 22050, 'n_freq': F, 'n_mels': 96, 'f_min': 30.0, 'f_max': 9000.0, 'htk': True, 'norm': None}
        fb = O.filterbank_mel(22050, F, 96, 30.0, 9000.0, True, None)
    else:
        kwargs = {'sample_rate': 22050, 'n_freq': F, 'n_bins': 84, 'bins_per_octave': 12, 'f_min': None, 'spread': 0.125}
        fb = O.filterbank_log(22050, F, 84, 12)
    layer = K.ApplyFilterbank(type=kind, filterbank_kwargs=kwargs, data_format=fmt)
    np.testing.assert_allclose(layer.filterbank, fb, rtol=1e-6, atol=1e-9)
    got = layer(x)
    ref = O.apply_filterbank(x.astype(np.float64), fb.astype(np.float64), fmt)
    assert got.shape == ref.shape
    assert nerr(got, ref) < 2e-6


@pytest.mark.parametrize('fbmma', ['0', '1'])
def test_log_frequency_layer_fused(K, monkeypatch, fbmma):
    monkeypatch.setenv('KAPRE_B200_FBMMA', fbmma)
    rng = np.random.default_rng(12)
    x = wave(rng, 2, 1, 8000, 'channels_last')
    seq = K.get_log_frequency_spectrogram_layer(n_fft=1024, hop_length=256, return_decibel=True)
    got = seq(torch.from_numpy(x).cuda()).cpu().numpy()
    s = np.abs(O.stft_layer(x, 1024, None, 256))
    ref = O.magnitude_to_decibel(O.apply_filterbank(s, O.filterbank_log(22050, 513).astype(np.float64), 'default'))
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 5e-4


# ------------------------------------------------------------------------------- inverse STFT
@pytest.mark.parametrize('n_fft,win,hop', [(1024, 1024, 256), (2048, 2048, 1024), (2048, 2048, 256), (512, 400, 100),
                                           (256, 256, 64), (1000, 1000, 250), (1000, 600, 200),
                                           (2048, 2048, 64), (512, 512, 31), (1024, 1024, 128), (256, 200, 7)])
@pytest.mark.parametrize('ifmt', ['channels_first', 'channels_last'])
@pytest.mark.parametrize('ofmt', ['channels_first', 'channels_last'])
def test_istft_vs_oracle(K, n_fft, win, hop, ifmt, ofmt):
    rng = np.random.default_rng(n_fft + hop)
    B, C, T, F = 2, 2, 41, n_fft // 2 + 1
    shape = (B, C, T, F) if ifmt == 'channels_first' else (B, T, F, C)
    X = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64)
    got = K.InverseSTFT(n_fft=n_fft, win_length=win, hop_length=hop, input_data_format=ifmt, output_data_format=ofmt)(X)
    ref = O.istft_layer(X, n_fft, win, hop, None, ifmt, ofmt)
    assert got.shape == ref.shape and got.dtype == np.float32
    assert nerr(got, ref) < 2e-6


@pytest.mark.parametrize('wfmt', ['default', 'channels_first', 'channels_last'])
@pytest.mark.parametrize('sfmt', ['default', 'channels_first', 'channels_last'])
@pytest.mark.parametrize('hop_ratio', [0.5, 0.25, 0.125])
def test_perfectly_reconstructing_stft_istft(K, golden, wfmt, sfmt, hop_ratio):
    """Mirror of tests/test_time_frequency.py:447-486 (STFT -> ISTFT, atol 1e-5)."""
    src = golden['audio']
    x = src[None, :, None] if wfmt != 'channels_first' else src[None, None, :]
    n_fft, hop = 2048, int(2048 * hop_ratio)
    stft, istft = K.get_perfectly_reconstructing_stft_istft(n_fft, hop, wfmt, sfmt)
    rec = K.Sequential([stft, istft])(x)
    lead = n_fft - hop
    rec = rec[:, :, lead:lead + 8000] if wfmt == 'channels_first' else rec[:, lead:lead + 8000, :]
    np.testing.assert_allclose(x, rec, atol=1e-5)


def test_cfg4_roundtrip_mse(K):
    """BASELINE cfg4: batch 128 mono 16 kHz, n_fft 1024 hop 256, reconstruction error."""
    g = torch.Generator(device='cuda').manual_seed(1234)
    x = torch.rand((128, 16000, 1), generator=g, device='cuda') * 2 - 1
    stft, istft = K.get_perfectly_reconstructing_stft_istft(1024, 256, 'channels_last', 'channels_last')
    y = istft(stft(x))
    assert y.shape == (128, 17664, 1)
    d = y[:, 768:768 + 16000, :] - x
    assert float(d.abs().max()) < 5e-6 and float((d * d).mean()) < 1e-12


# ------------------------------------------------------------------------------- full-size properties (cfg2)
def test_cfg2_full_size_properties(K):
    """BASELINE cfg2 (batch 256, 22.05 kHz x 5 s, n_fft 1024, hop 256, 128 mel): spot-check items
    against the oracle, plus size-independent properties on the whole batch."""
    g = torch.Generator(device='cuda').manual_seed(1234)
    x = torch.rand((256, 110250, 1), generator=g, device='cuda') * 2 - 1
    mel = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128,
                                     input_data_format='channels_last', output_data_format='channels_last')
    y = mel(x)
    assert y.shape == (256, 427, 128, 1)
    assert bool(torch.isfinite(y).all())
    for i in (0, 101, 255):
        ref = O.melspectrogram_layer(x[i:i + 1].cpu().numpy(), n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128)
        assert nerr(y[i:i + 1].cpu().numpy(), ref) < 2e-6
    # homogeneity: mel(2x) == 2 mel(x) exactly (power-of-two scaling commutes with every fp32 op on the path)
    assert torch.equal(mel(2.0 * x), 2.0 * y)
    # batch sharding: any split of the batch gives bit-identical items (the multi-GPU contract)
    parts = torch.cat([mel(x[:100]), mel(x[100:228]), mel(x[228:])], dim=0)
    assert torch.equal(parts, y)
    # time shift by one hop moves the frames by one
    ys = mel(x[:8, 256:, :])
    assert torch.equal(ys[:, :425], y[:8, 1:426])
    # decibel path: monotone map of the linear path, clamp inactive (max-80 < floor) -> identical to formula
    ydb = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128, return_decibel=True,
                                     input_data_format='channels_last', output_data_format='channels_last')(x)
    expect = 10.0 * torch.log10(torch.clamp(y.double(), min=1e-5))
    assert float((ydb.double() - expect).abs().max()) < 1e-4


def test_full_tensor_blocks_of_16_items_cfg2_and_cfg3(K):
    """Full-tensor (every element, not spot checks) oracle comparison of 16-item blocks at the BASELINE sizes:
    cfg2 log-mel (16 x 5 s, n_fft 1024, hop 256, 128 mel, dB) and cfg3 (16 x 6 channels x 1 s channels_last, n_fft 2048,
    hop 1024, magnitude + dB).  Linear outputs at 2e-6 of the block maximum; dB values wherever the linear value is above
    1e-6 of the item peak (below that the fp32 FFT noise floor under a logarithm is not a parity statement) at 1e-3 dB."""
    g = torch.Generator(device='cuda').manual_seed(77)
    x2 = (torch.rand((16, 110250, 1), generator=g, device='cuda') * 2 - 1)
    x2[3] *= 1e-3
    kw2 = dict(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128)
    lin = K.get_melspectrogram_layer(**kw2)(x2).cpu().numpy()
    db = K.get_melspectrogram_layer(return_decibel=True, **kw2)(x2).cpu().numpy()
    xn = x2.cpu().numpy()
    ref_lin = O.melspectrogram_layer(xn, **kw2)
    ref_db = O.melspectrogram_layer(xn, return_decibel=True, **kw2)
    assert lin.shape == ref_lin.shape == (16, 427, 128, 1)
    for i in range(16):
        assert np.abs(lin[i] - ref_lin[i]).max() < 2e-6 * np.abs(ref_lin[i]).max(), i
        sig = ref_lin[i] > 1e-6 * ref_lin[i].max()
        assert sig.mean() > 0.99
        assert np.abs(db[i] - ref_db[i])[sig].max() < 1e-3, i
    x3 = (torch.rand((16, 44100, 6), generator=g, device='cuda') * 2 - 1)
    x3[:, :, 4] *= 1e-2
    kw3 = dict(n_fft=2048, hop_length=1024, input_data_format='channels_last', output_data_format='channels_last')
    mag = K.get_stft_magnitude_layer(**kw3)(x3).cpu().numpy()
    mdb = K.get_stft_magnitude_layer(return_decibel=True, **kw3)(x3).cpu().numpy()
    xn3 = x3.cpu().numpy()
    ref_mag = O.stft_magnitude_layer(xn3, **kw3)
    ref_mdb = O.stft_magnitude_layer(xn3, return_decibel=True, **kw3)
    assert mag.shape == ref_mag.shape == (16, 42, 1025, 6)
    for i in range(16):
        assert np.abs(mag[i] - ref_mag[i]).max() < 2e-6 * np.abs(ref_mag[i]).max(), i
        sig = ref_mag[i] > 1e-5 * ref_mag[i].max()
        assert np.abs(mdb[i] - ref_mdb[i])[sig].max() < 2e-3, i


# ------------------------------------------------------------------------------- C ABI behaviour
def test_abi_errors_and_config(K):
    from kapre_b200 import _native, ops
    with pytest.raises(_native.KapreNativeError):
        ops.stft_forward(torch.zeros(1, 100, 1), K.STFT(n_fft=512).plan, 'channels_last', 'channels_last', False, False)
    with pytest.raises(ValueError):
        K.InverseSTFT(n_fft=512)(torch.zeros((1, 3, 100, 1), dtype=torch.complex64).cuda())  # wrong bin count
    # too-short input -> zero frames, like tf.signal.frame
    out = K.STFT(n_fft=512, hop_length=128)(torch.zeros(2, 100, 1).cuda())
    assert out.shape == (2, 0, 257, 1)
    # config round trip
    m = K.get_melspectrogram_layer(n_fft=512, hop_length=128, n_mels=32, return_decibel=True)
    m2 = K.Sequential.from_config(m.get_config())
    x = torch.rand(2, 4000, 1).cuda()
    assert torch.equal(m(x), m2(x))
    n0 = _native.launch_count()
    m(x)
    assert _native.launch_count() - n0 == 2  # fused kernel + clamp kernel


# ------------------------------------------------------------------------------- other BASELINE configs at full size
def test_cfg1_reference_cpu_case(K):
    """BASELINE cfg1: batch 4, mono 16 kHz x 1 s, n_fft 512 hop 256 64 mel -- whole tensor vs oracle."""
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, size=(4, 16000, 1)).astype(np.float32)
    kw = dict(n_fft=512, hop_length=256, sample_rate=16000, n_mels=64, return_decibel=True)
    got = K.get_melspectrogram_layer(**kw).predict(x)
    ref = O.melspectrogram_layer(x, **kw)
    assert got.shape == ref.shape == (4, 61, 64, 1)
    assert np.abs(got - ref).max() < 2e-4


def test_cfg3_full_size_properties(K):
    """BASELINE cfg3: batch 1024, 6 channels, 44.1 kHz x 1 s, channels_last, n_fft 2048 hop 1024,
    magnitude + decibel (get_stft_magnitude_layer).  Spot-check items against the oracle and check
    size-independent properties on the whole 1 GB tensor."""
    g = torch.Generator(device='cuda').manual_seed(4321)
    x = torch.rand((1024, 44100, 6), generator=g, device='cuda') * 2 - 1
    x[7] *= 1e-3
    kw = dict(n_fft=2048, hop_length=1024, return_decibel=True, input_data_format='channels_last',
              output_data_format='channels_last')
    layer = K.get_stft_magnitude_layer(**kw)
    y = layer(x)
    assert y.shape == (1024, 42, 1025, 6)
    assert bool(torch.isfinite(y).all())
    lin = K.get_stft_magnitude_layer(**dict(kw, return_decibel=False))(x[:8])
    for i in (0, 7):
        ref = O.stft_magnitude_layer(x[i:i + 1].cpu().numpy(), **kw)
        sig = lin[i:i + 1].cpu().numpy() > 1e-4 * float(lin[i].max())
        assert np.abs(y[i:i + 1].cpu().numpy() - ref)[sig].max() < 2e-3
    # sharding property: any split of the batch reproduces the items bit for bit
    assert torch.equal(layer(x[512:640]), y[512:640])
    # channels are independent up to the shared per-item clamp (inactive here): permuting channels permutes outputs
    perm = [3, 0, 5, 1, 4, 2]
    assert torch.equal(layer(x[:16][:, :, perm]), y[:16][:, :, :, perm])
    # channels_first output of the same data is the transpose
    yf = K.get_stft_magnitude_layer(**dict(kw, output_data_format='channels_first'))(x[:16])
    assert torch.equal(yf.permute(0, 2, 3, 1), y[:16])


def test_cfg5_shard_full_size_properties(K):
    """BASELINE cfg5, one GPU's shard: batch 1024, mono 16 kHz x 10 s, log-mel (n_fft 1024, hop 256, 128 mel)."""
    g = torch.Generator(device='cuda').manual_seed(99)
    x = torch.rand((1024, 160000, 1), generator=g, device='cuda') * 2 - 1
    kw = dict(n_fft=1024, hop_length=256, sample_rate=16000, n_mels=128, return_decibel=True)
    layer = K.get_melspectrogram_layer(**kw)
    y = layer(x)
    assert y.shape == (1024, 622, 128, 1)
    assert bool(torch.isfinite(y).all())
    for i in (0, 1023):
        ref = O.melspectrogram_layer(x[i:i + 1].cpu().numpy(), **kw)
        assert np.abs(y[i:i + 1].cpu().numpy() - ref).max() < 2e-4
    # 8-way shard boundaries (the 8-GPU job) are bit-identical to the unsharded run
    for r in (0, 3, 7):
        assert torch.equal(layer(x[r * 128:(r + 1) * 128]), y[r * 128:(r + 1) * 128])
    # scaling the waveform by 2 adds 10*log10(2) dB everywhere above the amin floor
    y2 = layer(2.0 * x[:32])
    above = y[:32] > -49.0
    assert float(((y2 - y[:32]) - 10.0 * np.log10(2.0)).abs()[above].max()) < 1e-4


def test_predict_pipelined_host_path(K):
    """predict() with host buffers (chunked H2D / kernels / D2H on three streams) == device call."""
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, size=(37, 9000, 2)).astype(np.float32)
    layer = K.get_melspectrogram_layer(n_fft=512, hop_length=128, n_mels=40, return_decibel=True, db_dynamic_range=30.0)
    full = layer(torch.from_numpy(x).cuda()).cpu().numpy()
    for bs in (None, 5, 64):
        np.testing.assert_array_equal(layer.predict(x, batch_size=bs), full)
    pinned = torch.from_numpy(x).pin_memory()
    np.testing.assert_array_equal(layer.predict(pinned), full)


# ------------------------------------------------------------------------------- "next" row: get_stft_mag_phase
@pytest.mark.parametrize('fmt', ['default', 'channels_first', 'channels_last'])
@pytest.mark.parametrize('n_fft,hop', [(512, 256), (2048, 512), (1000, 250)])
@pytest.mark.parametrize('db', [False, True])
def test_stft_mag_phase(K, golden, fmt, n_fft, hop, db):
    """Mirror of tests/test_time_frequency.py:390-444 (magnitude atol 2e-4) plus the phase half,
    which the reference leaves untested, against the oracle (kapre/composed.py:420-511)."""
    rng = np.random.default_rng(n_fft)
    C = 2
    x = rng.uniform(-1, 1, size=(3, C, 7000) if fmt == 'channels_first' else (3, 7000, C)).astype(np.float32)
    x[1] *= 1e-2
    kw = dict(n_fft=n_fft, hop_length=hop, return_decibel=db, db_dynamic_range=40.0, input_data_format=fmt,
              output_data_format=fmt)
    layer = K.get_stft_mag_phase(input_shape=x.shape[1:], **kw)
    got = layer(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = O.stft_mag_phase_layer(x, **kw)
    assert got.shape == ref.shape
    ax = 1 if fmt == 'channels_first' else 3
    gm, gp = np.split(got, 2, axis=ax)
    rm, rp = np.split(ref, 2, axis=ax)
    lin = np.abs(O.stft_layer(x, n_fft, None, hop, input_data_format=fmt, output_data_format=fmt))
    sig = lin > 1e-3 * lin.reshape(3, -1).max(1).reshape(3, 1, 1, 1)
    if db:
        assert np.abs(gm - rm)[sig].max() < 2e-3          # dB
        assert np.abs(gm - rm).max() < 5e-2               # incl. clamp-floor region (fp32 noise bins)
    else:
        np.testing.assert_allclose(gm, rm, atol=2e-4)     # reference tolerance
    d = np.angle(np.exp(1j * (gp - rp)))
    assert np.abs(d)[sig].max() < 2e-4                    # radians, where the bin is not rounding noise
    # speech fixture, magnitude half only (what the reference test checks)
    xs = golden['audio'][None, :, None] if fmt != 'channels_first' else golden['audio'][None, None, :]
    g2 = K.get_stft_mag_phase(input_shape=xs.shape[1:], n_fft=512, win_length=512, hop_length=256,
                              input_data_format=fmt, output_data_format=fmt)(xs)
    mag = np.take(g2[0], [0], axis=0 if fmt == 'channels_first' else 2)
    refm = np.abs(golden['stft_512_256_hann_window'])
    refm = refm[None] if fmt == 'channels_first' else refm[:, :, None]
    np.testing.assert_allclose(mag, refm, atol=2e-4)


# ------------------------------------------------------------------------------- "next" rows: Delta, Frame, Energy, MFCC
def test_delta_literal_and_modes(K):
    """tests/test_time_frequency.py:375-387 (literal KAT) + all modes / formats vs the oracle."""
    x = np.array([1.0, 2.0, 3.0, 4.0], dtype=np.float32).reshape(1, -1, 1, 1)
    got = K.Delta(win_length=3, data_format='channels_last')(x)
    np.testing.assert_allclose(got, np.array([0.5, 1.0, 1.0, 0.5], np.float32).reshape(1, -1, 1, 1))
    rng = np.random.default_rng(0)
    for fmt in ('channels_first', 'channels_last'):
        shape = (2, 3, 17, 40) if fmt == 'channels_first' else (2, 17, 40, 3)
        x = rng.normal(size=shape).astype(np.float32)
        for mode in ('symmetric', 'reflect', 'constant'):
            for win in (3, 5, 9):
                got = K.Delta(win_length=win, mode=mode, data_format=fmt)(x)
                ref = O.delta(x, win, mode, fmt)
                np.testing.assert_allclose(got, ref, atol=2e-6)
    with pytest.raises(ValueError):
        K.Delta(win_length=4)
    with pytest.raises(ValueError):
        K.Delta(mode='wrap')


@pytest.mark.parametrize('fmt', ['default', 'channels_first', 'channels_last'])
def test_concatenate_frequency_map(K, fmt):
    """Mirror of the reference's test (tests/test_time_frequency.py: ConcatenateFrequencyMap): the added channel is
    linspace(0, 1, n_freq) on the frequency axis, the other channels are untouched; float equality."""
    rng = np.random.default_rng(4)
    shape = (2, 3, 7, 33) if fmt == 'channels_first' else (2, 7, 33, 3)
    x = rng.normal(size=shape).astype(np.float32)
    layer = K.ConcatenateFrequencyMap(data_format=fmt)
    got = layer(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = O.concat_frequency_map(x, fmt)
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got, ref)
    assert layer.get_config()['data_format'] == fmt
    # after a spectrogram layer, as in the reference's docstring example
    seq = K.Sequential([K.STFT(n_fft=256, hop_length=128), K.Magnitude(), K.ConcatenateFrequencyMap()])
    y = seq(torch.from_numpy(rng.normal(size=(2, 2048, 1)).astype(np.float32)).cuda())
    assert tuple(y.shape) == (2, 15, 129, 2)
    np.testing.assert_allclose(y[0, 3, :, 1].cpu().numpy(), np.linspace(0.0, 1.0, 129), atol=1e-7)


@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
def test_spec_augment(K, fmt):
    """Mirror of the reference's SpecAugment tests (tests/test_augmentation.py): identity unless training=True, depth must be 1,
    masked elements equal mask_value and every other element is untouched; the applied masks (drawn on the host) reproduce
    the output through the oracle exactly; widths / starts stay inside the ranges tf.random.uniform would give."""
    rng = np.random.default_rng(8)
    shape = (5, 1, 40, 30) if fmt == 'channels_first' else (5, 40, 30, 1)
    x = rng.normal(size=shape).astype(np.float32) + 10.0
    layer = K.SpecAugment(freq_mask_param=5, time_mask_param=10, n_freq_masks=3, n_time_masks=2, mask_value=-1.0,
                          data_format=fmt, seed=3)
    xt = torch.from_numpy(x).cuda()
    assert layer(xt) is xt and layer(xt, training=False) is xt
    y = layer(xt, training=True).cpu().numpy()
    tm, fm = layer.last_masks
    assert tm.shape == (5, 2, 2) and fm.shape == (5, 3, 2)
    assert (tm[..., 1] >= 0).all() and (tm[..., 1] < 10).all() and (tm[..., 0] >= 0).all() and (tm[..., 0] + tm[..., 1] < 40).all()
    assert (fm[..., 1] >= 0).all() and (fm[..., 1] < 5).all() and (fm[..., 0] + fm[..., 1] < 30).all()
    ref = O.spec_augment(x, tm, fm, -1.0, fmt)
    np.testing.assert_array_equal(y, ref)
    assert (y == -1.0).any() and (y[y != -1.0] > 0).all()
    cfg = layer.get_config()
    assert cfg['freq_mask_param'] == 5 and cfg['n_time_masks'] == 2 and cfg['data_format'] == fmt
    with pytest.raises(RuntimeError):
        K.SpecAugment(freq_mask_param=0, time_mask_param=3)
    with pytest.raises(RuntimeError):
        bad = torch.zeros((2, 2, 40, 30) if fmt == 'channels_first' else (2, 40, 30, 2), device='cuda')
        layer(bad, training=True)
    with pytest.raises(ValueError):
        K.SpecAugment(freq_mask_param=50, time_mask_param=3, data_format=fmt)(xt, training=True)


@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
@pytest.mark.parametrize('pad_end', [False, True])
def test_frame_and_energy(K, fmt, pad_end):
    """Mirrors of tests/test_signal.py (Frame / Energy) against the oracle."""
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, size=(2, 3, 5000) if fmt == 'channels_first' else (2, 5000, 3)).astype(np.float32)
    got = K.Frame(frame_length=1024, hop_length=300, pad_end=pad_end, pad_value=0.25, data_format=fmt)(x)
    ref = O.frame_layer(x, 1024, 300, pad_end, 0.25, fmt)
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got, ref.astype(np.float32))
    e = K.Energy(sample_rate=16000, ref_duration=0.05, frame_length=800, hop_length=400, pad_end=pad_end,
                 data_format=fmt)(x)
    eref = O.energy_layer(x, 16000, 0.05, 800, 400, pad_end, 0, fmt)
    assert e.shape == eref.shape
    np.testing.assert_allclose(e, eref, rtol=2e-6)
    with pytest.raises(ValueError):
        K.Frame(frame_length=100, hop_length=200)


@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
def test_logmel_to_mfcc(K, fmt):
    """Mirror of tests/test_signal.py:79-106: HTK-scaled DCT-II of the log-mel spectrogram."""
    rng = np.random.default_rng(2)
    x = rng.uniform(-1, 1, size=(3, 9000, 2)).astype(np.float32)
    kw = dict(n_fft=512, hop_length=128, sample_rate=16000, n_mels=40, return_decibel=True)
    mel = K.get_melspectrogram_layer(output_data_format=fmt, **kw)
    seq = K.Sequential(list(mel.layers) + [K.LogmelToMFCC(n_mfccs=13, data_format=fmt)])
    got = seq(torch.from_numpy(x).cuda()).cpu().numpy()
    ref = O.logmel_to_mfcc(O.melspectrogram_layer(x, output_data_format=fmt, **kw), 13, fmt)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 5e-4


def test_cuda_graph_capture_matches_eager(K):
    """Sequential.capture: the same kernels recorded into a CUDA graph; replays with new inputs must equal
    the eager path bit for bit (cfg1 shape: log-mel with the fused dB clamp)."""
    rng = np.random.default_rng(11)
    layer = K.get_melspectrogram_layer(n_fft=512, hop_length=256, sample_rate=16000, n_mels=64, return_decibel=True,
                                       input_data_format='channels_last', output_data_format='channels_last')
    x0 = torch.from_numpy(rng.uniform(-1, 1, size=(4, 16000, 1)).astype(np.float32)).cuda()
    cap = layer.capture(x0)
    for scale in (1.0, 1e-3, 0.3):
        x = torch.from_numpy((scale * rng.uniform(-1, 1, size=(4, 16000, 1))).astype(np.float32)).cuda()
        want = layer(x)
        got = cap(x).clone()
        assert torch.equal(got, want)
    got_host = cap.predict(x.cpu().numpy())
    assert np.array_equal(got_host, want.cpu().numpy())
    with pytest.raises(ValueError):
        cap(torch.zeros((2, 16000, 1), device='cuda'))


def test_more_than_2_31_elements(K):
    """Maximum sizes: a waveform tensor with > 2^31 elements (8.9 GB) -- every global offset must be
    64-bit.  Items at both ends of the batch are checked against a small run of the same items."""
    B, L = 8192, 270000
    assert B * L > 2 ** 31
    free, _ = torch.cuda.mem_get_info()
    if free < 64 * 2 ** 30:
        pytest.skip('needs 64 GB of free device memory (8.9 GB input, 35 GB complex output, copies)')
    layer = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=16000, n_mels=128, return_decibel=True,
                                       input_data_format='channels_last', output_data_format='channels_last')
    x = torch.empty((B, L, 1), device='cuda')
    g = torch.Generator(device='cuda').manual_seed(3)
    for i in range(0, B, 1024):          # fill in slabs (keeps the RNG call below 2^31 elements)
        x[i:i + 1024].uniform_(-1, 1, generator=g)
    x[B - 1] *= 1e-2
    y = layer(x)
    assert y.shape == (B, 1051, 128, 1)
    for sl in (slice(0, 2), slice(B - 2, B), slice(B // 2, B // 2 + 1)):
        small = layer(x[sl].clone())
        assert torch.equal(y[sl], small)
    assert torch.isfinite(y[::257]).all()
    spec = K.STFT(n_fft=1024, hop_length=256)(x)           # complex output: 8192 * 1051 * 513 * 8 B = 35 GB
    small = K.STFT(n_fft=1024, hop_length=256)(x[B - 1:].clone())
    assert torch.equal(spec[B - 1:], small)


def test_c_client_matches_oracle(tmp_path):
    """The C ABI driven from plain C (tests/abi_c/abi_example.c: cudaMalloc / cudaMemcpy + libkapre_b200.so, no
    Python or torch in that process) reproduces the oracle's log-mel spectrogram."""
    import shutil
    import subprocess
    import sys
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_host_api import _build_c_client
    try:
        exe = _build_c_client(str(tmp_path / 'abi_example'))
    except (subprocess.CalledProcessError, OSError) as e:      # toolchain / headers missing on this box
        pytest.skip('cannot build the C client here: %s' % (e,))
    rng = np.random.default_rng(21)
    B, L, n_fft, hop, n_mels, sr = 3, 20000, 1024, 256, 128, 22050
    x = rng.uniform(-1, 1, size=(B, L)).astype(np.float32)
    x[1] *= 1e-2
    win = O.get_window(None, n_fft).astype(np.float32)
    fb = O.filterbank_mel(sr, n_fft // 2 + 1, n_mels, 0.0, None, False, 'slaney').astype(np.float32)
    x.tofile(tmp_path / 'wave.f32')
    win.tofile(tmp_path / 'window.f32')
    np.ascontiguousarray(fb).tofile(tmp_path / 'fb.f32')
    out = subprocess.run([exe, str(tmp_path / 'wave.f32'), str(B), str(L), str(tmp_path / 'window.f32'), str(n_fft),
                          str(hop), str(tmp_path / 'fb.f32'), str(n_mels), str(tmp_path / 'out.f32')],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    T = 1 + (L - n_fft) // hop
    got = np.fromfile(tmp_path / 'out.f32', dtype=np.float32).reshape(B, T, n_mels)
    ref = O.melspectrogram_layer(x[:, :, None], n_fft=n_fft, hop_length=hop, sample_rate=sr, n_mels=n_mels,
                                 return_decibel=True, input_data_format='channels_last',
                                 output_data_format='channels_last')[..., 0]
    assert np.abs(got - ref).max() < 1e-3
    assert 'launches=4' in out.stdout          # 2 calls x (fused kernel + clamp)


def test_pad_begin_with_hop_larger_than_n_fft_is_rejected(K):
    """kapre pads by n_fft - hop_length (time_frequency.py:169-172); tf.pad raises on a negative amount."""
    x = np.zeros((1, 4000, 1), dtype=np.float32)
    with pytest.raises(ValueError):
        K.STFT(n_fft=256, hop_length=300, pad_begin=True)(x)
    y = K.STFT(n_fft=256, hop_length=300, pad_begin=False)(x)      # hop > n_fft itself is fine
    assert y.shape == (1, 1 + (4000 - 256) // 300, 129, 1)
