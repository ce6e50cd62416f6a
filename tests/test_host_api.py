"""CPU tier: host-side logic of the kapre-compatible API (no GPU compute) and the C ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def K():
    import kapre_b200
    return kapre_b200


def test_import_surface(K):
    # hot-path subset of kapre/__init__.py:8-36
    for name in ('STFT', 'InverseSTFT', 'Magnitude', 'MagnitudeToDecibel', 'ApplyFilterbank', 'Phase',
                 'get_melspectrogram_layer', 'get_stft_magnitude_layer', 'get_perfectly_reconstructing_stft_istft',
                 'get_log_frequency_spectrogram_layer', 'backend', 'composed'):
        assert hasattr(K, name)
    import kapre
    assert kapre.STFT is K.STFT and kapre.composed.get_melspectrogram_layer is K.get_melspectrogram_layer


def test_stft_defaults_and_config(K):
    l = K.STFT()
    assert (l.n_fft, l.win_length, l.hop_length) == (2048, 2048, 512)  # hop = win // 4, time_frequency.py:128-129
    assert l.input_data_format == 'channels_last' and l.output_data_format == 'channels_last'
    l = K.STFT(n_fft=1000, win_length=400, name='foo', input_data_format='channels_first')
    assert l.hop_length == 100
    cfg = l.get_config()
    for key in ('n_fft', 'win_length', 'hop_length', 'window_name', 'pad_begin', 'pad_end', 'input_data_format',
                'output_data_format', 'name', 'trainable', 'dtype'):
        assert key in cfg
    assert cfg['output_data_format'] == 'default' and cfg['input_data_format'] == 'channels_first'
    l2 = K.STFT.from_config(cfg)
    assert l2.get_config() == cfg
    i = K.InverseSTFT(n_fft=512, hop_length=128, forward_window_name='hamming_window')
    assert set(i.get_config()) >= {'n_fft', 'win_length', 'hop_length', 'forward_window_name', 'input_data_format',
                                   'output_data_format'}
    d = K.MagnitudeToDecibel(ref_value=2.0, amin=1e-3, dynamic_range=50.0).get_config()
    assert (d['ref_value'], d['amin'], d['dynamic_range']) == (2.0, 1e-3, 50.0)


def test_validation_errors(K):
    for cls in (K.STFT, K.InverseSTFT):
        with pytest.raises(ValueError):
            cls(input_data_format='weird_string')        # tests/test_time_frequency.py:645-654
        with pytest.raises(ValueError):
            cls(output_data_format='weird_string')
        with pytest.raises(TypeError):
            cls(input_data_format={'config': 'channels_last'})  # validated before the dict unwrap
    with pytest.raises(ValueError):
        K.ApplyFilterbank(type='mel', filterbank_kwargs={'sample_rate': 22050, 'n_freq': 257}, data_format='weird')
    with pytest.raises(NotImplementedError):
        K.backend.get_window_fn('wrong_window_name')     # tests/test_backend.py:127-129
    with pytest.raises(ValueError):
        K.backend.validate_data_format_str('weird_string')
    with pytest.raises(RuntimeError):
        K.backend.filterbank_log(sample_rate=22050, n_freq=513, n_bins=300, bins_per_octave=12)
    with pytest.raises(ValueError):
        K.get_melspectrogram_layer(input_data_format='nope')
    # parameters of MagnitudeToDecibel are only checked when it is called (backend.py:168-173)
    K.MagnitudeToDecibel(amin=-1.0)
    with pytest.raises(ValueError):
        K.backend.magnitude_to_decibel(np.ones(4, np.float32), amin=-1.0)


@pytest.mark.parametrize('sample_rate', [44100, 22050])
@pytest.mark.parametrize('n_freq', [1025, 257])
@pytest.mark.parametrize('n_mels', [32, 128])
@pytest.mark.parametrize('f_min', [0.0, 200])
@pytest.mark.parametrize('f_max_ratio', [1.0, 0.5])
@pytest.mark.parametrize('htk', [True, False])
@pytest.mark.parametrize('norm', [None, 'slaney', 1.0])
def test_mel(K, sample_rate, n_freq, n_mels, f_min, f_max_ratio, htk, norm):
    """Mirror of tests/test_backend.py:43-75 with the oracle's librosa restatement as reference."""
    f_max = int(f_max_ratio * (sample_rate // 2))
    fb = K.backend.filterbank_mel(sample_rate=sample_rate, n_freq=n_freq, n_mels=n_mels, f_min=f_min, f_max=f_max,
                                  htk=htk, norm=norm)
    ref = O.filterbank_mel(sample_rate, n_freq, n_mels, f_min, f_max, htk, norm)
    assert fb.dtype == np.float32 and fb.shape == (n_freq, n_mels)
    np.testing.assert_allclose(ref, fb, rtol=1e-6, atol=1e-10)


def test_log_filterbank_and_windows(K):
    fb = K.backend.filterbank_log(22050, 513, 84, 12)
    assert fb.shape == (513, 84) and fb.dtype == np.float32   # tests/test_backend.py:78-96
    np.testing.assert_allclose(fb, O.filterbank_log(22050, 513, 84, 12), rtol=1e-6, atol=1e-12)
    for name in (None, 'hann_window', 'hamming_window'):
        for W in (1, 7, 512, 1000):
            np.testing.assert_allclose(K.backend.get_window_fn(name)(W), O.get_window(name, W), atol=1e-7)
    for name in ('kaiser_window', 'kaiser_bessel_derived_window', 'vorbis_window'):
        w = K.backend.get_window_fn(name)(512)
        assert w.shape == (512,) and np.isfinite(w).all() and w.max() <= 1.0 + 1e-6
    dual = K.backend.inverse_stft_window_fn(256, K.backend.get_window_fn(None))(1024)
    np.testing.assert_allclose(dual, O.inverse_stft_window(1024, 256, O.get_window(None, 1024)), atol=1e-6)


def test_composed_structure_and_fusion_plan(K):
    m = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, return_decibel=True, name='mel')
    assert [type(l).__name__ for l in m.layers] == ['STFT', 'Magnitude', 'ApplyFilterbank', 'MagnitudeToDecibel']
    assert m.name == 'mel' and m.layers[2].filterbank.shape == (513, 128)
    m2 = K.Sequential.from_config(m.get_config())
    assert m2.get_config() == m.get_config()
    s = K.get_stft_magnitude_layer(n_fft=512)
    assert [type(l).__name__ for l in s.layers] == ['STFT', 'Magnitude']
    stft, istft = K.get_perfectly_reconstructing_stft_istft(1024, 256, 'channels_last', 'channels_first')
    assert stft.pad_begin and stft.pad_end and stft.window_name == 'hann_window'
    assert istft.input_data_format == 'channels_first' and istft.output_data_format == 'channels_last'


def test_no_cpu_fallback(K):
    """The product path must fail loudly without a GPU instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from kapre_b200 import _native
    with pytest.raises(_native.KapreNativeError):
        K.get_melspectrogram_layer(n_fft=512)(np.zeros((1, 4000, 1), np.float32))
    with pytest.raises(_native.KapreNativeError):
        K.backend.magnitude_to_decibel(np.ones((2, 3), np.float32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'kapre_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(import|from)\s+oracle', text, re.M), f


def test_abi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports everything include/kapre_b200.h declares."""
    from kapre_b200 import _native
    lib = _native.lib()
    header = open(os.path.join(ROOT, 'include', 'kapre_b200.h')).read()
    declared = set(re.findall(r'\b(kapre_[a-z_0-9]+)\s*\(', header))
    declared -= {'kapre_wave_desc', 'kapre_spec_desc', 'kapre_db_cfg'}
    bound = {name for name, _, _ in _native.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.kapre_version() == 200   # round 2: 8-byte-per-item decibel workspace, sampled profiling, tcgen05 probe
    assert lib.kapre_launch_count() == 0 or lib.kapre_launch_count() > 0
    # argument validation happens before any CUDA call
    out = ctypes.c_void_p()
    w = np.ones(8, np.float32)
    rc = lib.kapre_stft_plan_create(0, 8, 2, w.ctypes.data_as(ctypes.c_void_p), ctypes.byref(out))
    assert rc == -1 and b'out of range' in lib.kapre_last_error()


def test_adjacent_layers_config_and_validation():
    """Delta / Frame / Energy / LogmelToMFCC mirror the reference's arguments and errors
    (kapre/time_frequency.py:563-644, kapre/signal.py:22-233, :365-447)."""
    import kapre_b200 as K
    d = K.Delta(win_length=7, mode='reflect', data_format='channels_first')
    cfg = d.get_config()
    assert (cfg['win_length'], cfg['mode'], cfg['data_format']) == (7, 'reflect', 'channels_first')
    assert d.n == 3 and d.denom == 2 * (1 + 4 + 9)
    for bad in (dict(win_length=2), dict(win_length=4), dict(mode='wrap'), dict(data_format='weird')):
        with pytest.raises(ValueError):
            K.Delta(**bad)
    f = K.Frame(frame_length=512, hop_length=128, pad_end=True, pad_value=1, data_format='channels_last')
    assert K.Frame.from_config(f.get_config()).get_config() == f.get_config()
    e = K.Energy(sample_rate=8000, frame_length=400, hop_length=200)
    assert e.get_config()['ref_duration'] == 0.1
    m = K.LogmelToMFCC(n_mfccs=13)
    assert m.get_config()['n_mfccs'] == 13
    import kapre
    assert kapre.signal.Frame is K.Frame and kapre.Delta is K.Delta


def test_dct_matrix_matches_oracle_and_scipy():
    import scipy.fft
    from kapre_b200.signal import dct2_htk_matrix
    from oracle import reference as O
    x = np.random.default_rng(0).normal(size=(2, 5, 40, 1))
    ref = O.logmel_to_mfcc(x, 13, 'channels_last')
    mat = dct2_htk_matrix(40, 13).astype(np.float64)
    got = np.einsum('btmc,mk->btkc', x, mat)
    np.testing.assert_allclose(got, ref, atol=1e-5)
    sp = scipy.fft.dct(x, type=2, axis=2)[:, :, :13] / np.sqrt(2 * 40)
    np.testing.assert_allclose(ref, sp, atol=1e-10)


def test_oracle_delta_frame_energy_known_answers():
    from oracle import reference as O
    x = np.array([1.0, 2.0, 3.0, 4.0]).reshape(1, -1, 1, 1)
    np.testing.assert_allclose(O.delta(x, 3, 'symmetric', 'channels_last').ravel(), [0.5, 1.0, 1.0, 0.5])
    w = np.arange(10, dtype=np.float64).reshape(1, 10, 1)
    fr = O.frame_layer(w, 4, 3, False, 0, 'channels_last')
    assert fr.shape == (1, 3, 4, 1)
    np.testing.assert_array_equal(fr[0, :, :, 0], [[0, 1, 2, 3], [3, 4, 5, 6], [6, 7, 8, 9]])
    fr = O.frame_layer(w, 4, 3, True, -1, 'channels_last')
    assert fr.shape == (1, 4, 4, 1)
    np.testing.assert_array_equal(fr[0, 3, :, 0], [9, -1, -1, -1])
    en = O.energy_layer(np.ones((1, 8, 1)), sample_rate=4, ref_duration=1.0, frame_length=4, hop_length=4,
                        data_format='channels_last')
    np.testing.assert_allclose(en.ravel(), [4.0, 4.0])


def test_predict_result_buffers_are_recycled_only_when_released(monkeypatch):
    """Sequential.predict hands out page-locked result buffers from a pool; a buffer may be reused only
    after the array returned for it (and every view of it) is gone."""
    import torch
    import kapre_b200.composed as C
    real_empty = torch.empty

    def cpu_empty(*a, **k):          # no CUDA in this tier: drop the pinning request
        k.pop('pin_memory', None)
        return real_empty(*a, **k)
    monkeypatch.setattr(torch, 'empty', cpu_empty)
    seq = C.Sequential([])

    def get():
        return seq._pinned_result((4, 5), torch.float32)[1]
    n1, n2 = get(), get()
    assert n1.ctypes.data != n2.ctypes.data
    p1 = n1.ctypes.data
    view = n1[1:3]
    del n1
    n3 = get()
    assert n3.ctypes.data not in (p1, n2.ctypes.data)     # the view still pins buffer 1
    del view
    n4 = get()
    assert n4.ctypes.data == p1                           # released -> recycled, no new allocation
    assert len(seq._result_pool) == 3


def _build_c_client(out_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, 'kapre_b200', '_lib')
    cuda = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    cmd = ['gcc', '-std=c11', '-O1', '-Wall', '-Werror', '-I' + os.path.join(root, 'include'),
           '-I' + os.path.join(cuda, 'include'), os.path.join(root, 'tests', 'abi_c', 'abi_example.c'), '-o', out_path,
           '-L' + lib_dir, '-lkapre_b200', '-L' + os.path.join(cuda, 'lib64'), '-lcudart',
           '-Wl,-rpath,' + lib_dir, '-Wl,-rpath,' + os.path.join(cuda, 'lib64')]
    subprocess.check_call(cmd)
    return out_path


def test_header_is_plain_c_and_c_client_links(tmp_path):
    """include/kapre_b200.h must be usable from C (no C++ / torch / CUDA types): the plain-C client of
    tests/abi_c compiles with gcc -std=c11 -Wall -Werror and links against libkapre_b200.so."""
    import shutil
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    exe = _build_c_client(str(tmp_path / 'abi_example'))
    assert os.path.exists(exe)


def test_host_copy_threads_and_views():
    """composed._host_copy (the pageable -> page-locked staging copy of Sequential.predict): single memcpy below the
    threshold, slices over the copy threads above it, odd sizes, and a destination that is a view of a larger buffer."""
    import numpy as np
    import torch
    from kapre_b200.composed import _host_copy
    rng = np.random.default_rng(0)
    for shape, min_bytes in (((3, 1001, 1), 1 << 20), ((5, 40001, 2), 1 << 12), ((1, 7), 4), ((4, 250000), 1 << 18)):
        x = torch.from_numpy(rng.normal(size=shape).astype(np.float32))
        raw = torch.full((x.numel() * 4 + 64,), 0x7f, dtype=torch.uint8)
        dst = raw[: x.numel() * 4].view(torch.float32).view(shape)
        _host_copy(dst, x, min_bytes)
        assert torch.equal(dst, x)
        assert bool((raw[x.numel() * 4:] == 0x7f).all())      # nothing written past the view
