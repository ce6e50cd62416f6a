"""Fixtures produced by RUNNING THE REFERENCE's own layer code (tests/golden/make_golden_ref.py: the
unmodified /root/reference/kapre sources over a NumPy stand-in for TensorFlow / librosa, float64).

CPU tier: the oracle must reproduce every fixture (pins the oracle's reading of kapre's semantics).
GPU tier: the CUDA path, driven through the same layer names and keyword arguments, must reproduce
them within the fp32 tolerances below.  Nothing here reads /root/reference at run time.
"""
import json
import os

import numpy as np
import pytest

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
_npz = np.load(os.path.join(HERE, 'golden', 'kapre_ref_cases.npz'))
with open(os.path.join(HERE, 'golden', 'kapre_ref_cases.json')) as _f:
    MANIFEST = json.load(_f)['cases']
CASES = {c['key']: c for c in MANIFEST}


def _input(case, dtype):
    k = case['input']
    if k is None:
        return None
    if k.startswith('stft:'):
        _, name, n_fft, hop = k.split(':')
        x = _npz['in_' + name].astype(np.float64)
        return O.stft_layer(x, n_fft=int(n_fft), hop_length=int(hop), input_data_format='channels_last',
                            output_data_format='channels_last')
    if k == 'magnitude_0[0,:,5,0]':
        return _npz['magnitude_0'][0, :, 5, 0].astype(dtype)
    if ('in_' + k) in _npz:
        x = _npz['in_' + k]
    else:
        x = _npz[k]
    if case.get('batch'):
        x = x[:case['batch']]
    return x.astype(np.complex128 if np.iscomplexobj(x) else dtype)


# ------------------------------------------------------------------------------- oracle side
def _oracle(case):
    kind, kw = case['kind'], dict(case['kwargs'])
    x = _input(case, np.float64)
    if kind in ('STFT', 'STFTTflite'):
        return O.stft_layer(x, **kw)
    if kind == 'Magnitude':
        return O.magnitude(x)
    if kind == 'Phase':
        return O.phase(x)
    if kind in ('MagnitudeToDecibel', 'backend.magnitude_to_decibel'):
        return O.magnitude_to_decibel(x, **kw)
    if kind == 'backend.filterbank_mel':
        return O.filterbank_mel(kw['sample_rate'], kw['n_freq'], kw['n_mels'], kw['f_min'], kw['f_max'], kw['htk'],
                                kw['norm'])
    if kind == 'backend.filterbank_log':
        return O.filterbank_log(kw['sample_rate'], kw['n_freq'], kw['n_bins'], kw['bins_per_octave'], kw['f_min'],
                                kw['spread'])
    if kind == 'get_stft_magnitude_layer':
        return O.stft_magnitude_layer(x, **kw)
    if kind == 'get_melspectrogram_layer':
        return O.melspectrogram_layer(x, **kw)
    if kind == 'get_log_frequency_spectrogram_layer':
        n_fft, sr = kw['n_fft'], kw['sample_rate']
        s = O.stft_magnitude_layer(x, n_fft=n_fft, hop_length=kw['hop_length'], return_decibel=False,
                                   input_data_format=kw['input_data_format'],
                                   output_data_format=kw['output_data_format'])
        fb = O.filterbank_log(sr, n_fft // 2 + 1, kw['log_n_bins'], 12, None, 0.125)
        y = O.apply_filterbank(s, fb, kw['output_data_format'])
        return O.magnitude_to_decibel(y) if kw['return_decibel'] else y
    if kind == 'get_stft_mag_phase':
        kw.pop('input_shape')
        return O.stft_mag_phase_layer(x, **kw)
    if kind == 'InverseSTFT':
        return O.istft_layer(x, **kw)
    if kind == 'get_perfectly_reconstructing_stft_istft':
        s = O.stft_layer(x, n_fft=kw['n_fft'], win_length=kw['n_fft'], hop_length=kw['hop_length'],
                         window_name='hann_window', pad_begin=True, pad_end=True,
                         input_data_format=kw['waveform_data_format'], output_data_format=kw['stft_data_format'])
        return O.istft_layer(s, n_fft=kw['n_fft'], win_length=kw['n_fft'], hop_length=kw['hop_length'],
                             forward_window_name='hann_window', input_data_format=kw['stft_data_format'],
                             output_data_format=kw['waveform_data_format'])
    if kind == 'Delta':
        return O.delta(x, kw['win_length'], kw['mode'], kw['data_format'])
    if kind == 'LogmelToMFCC':
        return O.logmel_to_mfcc(x, kw['n_mfccs'], kw['data_format'])
    if kind == 'Frame':
        return O.frame_layer(x, kw['frame_length'], kw['hop_length'], kw.get('pad_end', False), kw.get('pad_value', 0),
                             kw['data_format'])
    if kind == 'Energy':
        return O.energy_layer(x, kw['sample_rate'], kw['ref_duration'], kw['frame_length'], kw['hop_length'],
                              kw.get('pad_end', False), kw.get('pad_value', 0), kw['data_format'])
    raise KeyError(kind)


def _tolerance(case):
    """(absolute tolerance as a fraction of max|fixture|, absolute floor).  The fixtures are stored as
    float32, so 1e-6 of the maximum is their own resolution; the tflite cases were computed by the
    reference in float32 with a complex64 DFT matrix (tflite_compatible_stft.py:14-35)."""
    if case['kind'] == 'STFTTflite':
        return 3e-6, 0.0
    if case['kind'] in ('Delta', 'LogmelToMFCC'):
        return 0.0, 2e-5     # their input is the float32-rounded log-mel fixture (values up to ~80 dB)
    return 2e-7, 1e-7


@pytest.mark.parametrize('key', [c['key'] for c in MANIFEST])
def test_oracle_reproduces_reference_run(key):
    case = CASES[key]
    want = _npz[key]
    got = np.asarray(_oracle(case))
    assert list(got.shape) == case['shape'] == list(want.shape)
    rel, floor = _tolerance(case)
    tol = rel * float(np.abs(want).max()) + floor
    if case['kind'] == 'Phase':
        # angles of bins whose magnitude is at the float32 noise floor are not determined
        mag = np.abs(_input(case, np.float64))
        ok = mag > 1e-6 * mag.max()
        d = np.abs(np.angle(np.exp(1j * (got - want))))
        assert d[ok].max() < 1e-5
        return
    assert float(np.abs(got - want).max()) <= tol, (key, float(np.abs(got - want).max()), tol)


# ------------------------------------------------------------------------------- CUDA side
def _cuda(case, K):
    kind, kw = case['kind'], dict(case['kwargs'])
    x = _input(case, np.float32)
    if x is not None and np.iscomplexobj(x):
        x = x.astype(np.complex64)
    if kind in ('STFT', 'STFTTflite'):
        return K.STFT(**kw)(x)
    if kind in ('Magnitude', 'Phase', 'MagnitudeToDecibel', 'InverseSTFT', 'Delta', 'LogmelToMFCC', 'Frame',
                'Energy'):
        return getattr(K, kind)(**kw)(x)
    if kind == 'backend.magnitude_to_decibel':
        return K.backend.magnitude_to_decibel(x, **kw)
    if kind == 'backend.filterbank_mel':
        return np.asarray(K.backend.filterbank_mel(**kw))
    if kind == 'backend.filterbank_log':
        return np.asarray(K.backend.filterbank_log(**kw))
    if kind in ('get_stft_magnitude_layer', 'get_melspectrogram_layer', 'get_log_frequency_spectrogram_layer'):
        return getattr(K, kind)(**kw)(x)
    if kind == 'get_stft_mag_phase':
        kw['input_shape'] = tuple(kw['input_shape'])
        return K.get_stft_mag_phase(**kw)(x)
    if kind == 'get_perfectly_reconstructing_stft_istft':
        stft, istft = K.get_perfectly_reconstructing_stft_istft(**kw)
        return istft(stft(x))
    raise KeyError(kind)


def _gpu_tolerance(case, want):
    """fp32 CUDA path against the float64 reference run.  Linear outputs: 2e-6 of the largest value
    (FFT round-off is relative to the frame norm).  Decibel outputs: 2e-3 dB (north_star: 1e-4 relative
    on the linear scale = 4.3e-4 dB; the clamp floor sits 80 dB below the maximum)."""
    kind, kw = case['kind'], case['kwargs']
    if kind in ('MagnitudeToDecibel', 'backend.magnitude_to_decibel') or kw.get('return_decibel'):
        return 2e-3
    if kind in ('Delta', 'LogmelToMFCC'):
        return 1e-4      # consumes dB-scaled input stored as float32
    return 3e-6 * float(np.abs(want).max()) + 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize('key', [c['key'] for c in MANIFEST])
def test_cuda_reproduces_reference_run(key):
    import kapre_b200 as K
    case = CASES[key]
    want = _npz[key]
    got = _cuda(case, K)
    got = got.cpu().numpy() if hasattr(got, 'cpu') else np.asarray(got)
    assert list(got.shape) == list(want.shape)
    if case['kind'] == 'Phase':
        mag = np.abs(_input(case, np.float64))
        ok = mag > 1e-4 * mag.max()
        d = np.abs(np.angle(np.exp(1j * (got.astype(np.float64) - want))))
        assert d[ok].max() < 2e-3
        return
    kind, kw = case['kind'], case['kwargs']
    if kind == 'get_stft_mag_phase':
        # magnitude half: linear / decibel tolerance; phase half: angular distance where the bin is not at the
        # fp32 round-off floor (the angle of a numerically-zero bin is not determined)
        ax = 1 if kw['output_data_format'] == 'channels_first' else 3
        gm, gp = np.split(got.astype(np.float64), 2, axis=ax)
        wm, wp = np.split(want.astype(np.float64), 2, axis=ax)
        lin_w = 10.0 ** (wm / 10.0) if kw['return_decibel'] else wm
        lin_g = 10.0 ** (gm / 10.0) if kw['return_decibel'] else gm
        assert np.abs(lin_g - lin_w).max() <= 3e-6 * lin_w.max() + 1e-7
        if kw['return_decibel']:
            strong = lin_w > 1e-3 * lin_w.max()
            assert np.abs(gm - wm)[strong].max() < 2e-3
        strong = lin_w > 1e-3 * lin_w.max()
        assert np.abs(np.angle(np.exp(1j * (gp - wp))))[strong].max() < 2e-3
        return
    tol = _gpu_tolerance(case, want)
    diff = np.abs(got.astype(want.dtype) - want)
    if kind in ('MagnitudeToDecibel', 'backend.magnitude_to_decibel') or kw.get('return_decibel'):
        # decibel outputs: values whose LINEAR magnitude sits at the fp32 round-off floor of the FFT
        # (5e-7 of the item's peak, e.g. the far skirts of a pure tone around amin) are compared on
        # the linear scale; everything else must agree to `tol` dB
        lin_g, lin_w = 10.0 ** (got.astype(np.float64) / 10.0), 10.0 ** (want.astype(np.float64) / 10.0)
        floor = 5e-7 * lin_w.reshape(lin_w.shape[0], -1).max(axis=1).reshape((-1,) + (1,) * (lin_w.ndim - 1)) \
            if lin_w.ndim > 1 else 5e-7 * lin_w.max()
        bad = (diff > tol) & (np.abs(lin_g - lin_w) > floor)
        assert not bad.any(), (key, float(diff[bad].max()), tol)
        return
    err = float(diff.max())
    assert err <= tol, (key, err, tol)
