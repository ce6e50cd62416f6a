"""NumPy restatement of the kapre hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to /root/reference).
The arithmetic that lives inside TensorFlow / librosa (not vendored in the reference) is
restated from their published algorithms (tensorflow>=2.16,<2.21 ``tf.signal``; librosa>=0.11
``filters.mel``), see SURVEY.md Appendix A.

All functions take ``dtype`` (default float64): float64 is the "truth" the GPU path is
compared with; float32 reproduces the reference's working precision for the CPU baseline.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    'CH_FIRST', 'CH_LAST', 'CH_DEFAULT', 'resolve_data_format', 'get_window', 'num_frames',
    'frame_signal', 'stft_frames', 'stft_by_dft_matrix', 'stft_layer', 'magnitude',
    'hz_to_mel', 'mel_to_hz', 'filterbank_mel', 'filterbank_log', 'apply_filterbank',
    'magnitude_to_decibel', 'inverse_stft_window', 'inverse_stft_frames', 'istft_layer',
    'melspectrogram_layer', 'stft_magnitude_layer', 'phase', 'stft_mag_phase_layer',
    'delta', 'frame_layer', 'energy_layer', 'logmel_to_mfcc', 'dft_stage1_32x32', 'concat_frequency_map', 'spec_augment',
]

CH_FIRST = 'channels_first'
CH_LAST = 'channels_last'
CH_DEFAULT = 'default'


def resolve_data_format(fmt: str) -> str:
    """kapre/time_frequency.py:142-144 -- 'default' -> K.image_data_format() == 'channels_last'
    (the Keras default; kapre/backend.py:21-37 falls back to 'channels_last' too)."""
    if not isinstance(fmt, str):
        raise TypeError('data_format must be a string')  # kapre/backend.py:114-117
    if fmt not in (CH_DEFAULT, CH_FIRST, CH_LAST):
        raise ValueError('bad data_format %r' % (fmt,))  # kapre/backend.py:119-123
    return CH_LAST if fmt == CH_DEFAULT else fmt


# --------------------------------------------------------------------------- windows
def get_window(window_name, length: int, dtype=np.float64) -> np.ndarray:
    """kapre/backend.py:58-100 -> tf.signal.{hann,hamming}_window(length, periodic=True).

    TF's raised-cosine window: w[n] = a - b*cos(2*pi*n / D), D = W + periodic*even - 1 with
    even = 1 - W%2, i.e. periodic for even W but SYMMETRIC for odd W.
    W == 1 -> ones.
    """
    if window_name is None:
        window_name = 'hann_window'
    W = int(length)
    if window_name in ('kaiser_window', 'kaiser_bessel_derived_window', 'vorbis_window'):
        return _other_window(window_name, W).astype(dtype)     # backend.py:82-87
    coeffs = {'hann_window': (0.5, 0.5), 'hamming_window': (0.54, 0.46)}
    if window_name not in coeffs:
        raise NotImplementedError('Window name %s is not supported' % window_name)  # backend.py:89-98
    a, b = coeffs[window_name]
    if W == 1:
        return np.ones(1, dtype=dtype)
    even = 1 - W % 2
    D = W + even - 1
    n = np.arange(W, dtype=np.float64)
    return (a - b * np.cos(2.0 * np.pi * n / D)).astype(dtype)


def _bessel_i0(x):
    """Modified Bessel function I0 by its power series sum_k ((x/2)^k / k!)^2 (converges fast for x <= ~30)."""
    x = np.asarray(x, dtype=np.float64)
    term = np.ones_like(x)
    total = np.ones_like(x)
    for k in range(1, 200):
        term = term * (x / (2.0 * k)) ** 2
        total = total + term
    return total


def _other_window(name, W):
    """tf.signal.kaiser_window / kaiser_bessel_derived_window / vorbis_window with TF's default beta=12
    (kapre passes only the length, time_frequency.py:178).  Kaiser: I0(beta*sqrt(1-r^2))/I0(beta),
    r = 2n/(W-1)-1; KBD: sqrt of the normalised running sum of a (W/2+1)-point Kaiser, mirrored;
    Vorbis: sin(pi/2 * sin^2(pi*(n+1/2)/W))."""
    beta = 12.0

    def kaiser(m):
        if m == 1:
            return np.ones(1)
        r = 2.0 * np.arange(m, dtype=np.float64) / (m - 1) - 1.0
        return _bessel_i0(beta * np.sqrt(np.maximum(0.0, 1.0 - r * r))) / _bessel_i0(beta)

    if name == 'kaiser_window':
        return kaiser(W)
    if name == 'kaiser_bessel_derived_window':
        cs = np.cumsum(kaiser(W // 2 + 1))
        half = np.sqrt(cs[:-1] / cs[-1])
        return np.concatenate([half, half[::-1]])
    n = np.arange(W, dtype=np.float64) + 0.5
    return np.sin(np.pi / 2.0 * np.sin(np.pi * n / W) ** 2)


# --------------------------------------------------------------------------- framing / STFT
def num_frames(length: int, win_length: int, hop: int, pad_end: bool) -> int:
    """tf.signal.frame frame count; restated in-repo at kapre/tflite_compatible_stft.py:101
    (pad_end=False) and :179-182 (pad_end=True); tests/test_time_frequency.py:32-39."""
    if pad_end:
        return -(-length // hop)
    return max(0, 1 + (length - win_length) // hop)


def frame_signal(x: np.ndarray, win_length: int, hop: int, pad_end: bool) -> np.ndarray:
    """(..., L) -> (..., T, W).  kapre/tflite_compatible_stft.py:95-150 / :176-182."""
    L = x.shape[-1]
    T = num_frames(L, win_length, hop, pad_end)
    need = (T - 1) * hop + win_length if T > 0 else 0
    if need > L:
        pad = [(0, 0)] * (x.ndim - 1) + [(0, need - L)]
        x = np.pad(x, pad)
    idx = np.arange(T)[:, None] * hop + np.arange(win_length)[None, :]
    return x[..., idx]


def stft_frames(x, n_fft, win_length, hop, window, pad_end, dtype=np.float64):
    """tf.signal.stft on the last axis: frame, window, rfft with RIGHT zero-pad (or crop) to
    n_fft.  kapre/time_frequency.py:174-182; in-repo restatement
    kapre/tflite_compatible_stft.py:61-68 (pad) and :185-190."""
    x = np.asarray(x, dtype=dtype)
    frames = frame_signal(x, win_length, hop, pad_end) * np.asarray(window, dtype=dtype)
    spec = np.fft.rfft(frames.astype(np.float64), n=n_fft, axis=-1)  # n<W crops, n>W right-pads
    cdt = np.complex128 if np.dtype(dtype) == np.float64 else np.complex64
    return spec.astype(cdt)


def stft_by_dft_matrix(x, n_fft, win_length, hop, window, pad_end):
    """The reference's own matmul-DFT restatement (kapre/tflite_compatible_stft.py:14-35,
    61-75, 175-192): one-sided DFT matrix exp(-2*pi*i*k*n/N), right zero-pad, matmul.
    float64.  Used only to cross-check ``stft_frames``."""
    x = np.asarray(x, dtype=np.float64)
    frames = frame_signal(x, win_length, hop, pad_end) * np.asarray(window, dtype=np.float64)
    if win_length < n_fft:
        pad = [(0, 0)] * (frames.ndim - 1) + [(0, n_fft - win_length)]
        frames = np.pad(frames, pad)
    else:
        frames = frames[..., :n_fft]
    k = np.arange(n_fft // 2 + 1)
    n = np.arange(n_fft)
    mat = np.exp(np.outer(-2j * np.pi / n_fft * k, n)).T  # (n, k)
    return frames @ mat


def stft_layer(x, n_fft=2048, win_length=None, hop_length=None, window_name=None,
               pad_begin=False, pad_end=False, input_data_format='default',
               output_data_format='default', dtype=np.float64):
    """kapre.STFT.call, kapre/time_frequency.py:146-187 (+ __init__ defaults :126-129)."""
    idf = resolve_data_format(input_data_format)
    odf = resolve_data_format(output_data_format)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    x = np.asarray(x, dtype=dtype)
    if idf == CH_LAST:
        x = np.transpose(x, (0, 2, 1))  # :164-167
    if pad_begin:
        if hop_length > n_fft:   # tf.pad rejects negative paddings (InvalidArgumentError in the reference)
            raise ValueError('pad_begin needs hop_length <= n_fft: the padding is n_fft - hop_length = %d' % (n_fft - hop_length))
        x = np.pad(x, [(0, 0), (0, 0), (int(n_fft - hop_length), 0)])  # :169-172 (n_fft, not win)
    window = get_window(window_name, win_length, dtype=dtype)
    s = stft_frames(x, n_fft, win_length, hop_length, window, pad_end, dtype=dtype)  # (b, c, t, f)
    if odf == CH_LAST:
        s = np.transpose(s, (0, 2, 3, 1))  # :184-185
    return s


def magnitude(x):
    """kapre.Magnitude.call, kapre/time_frequency.py:351-359 (tf.abs)."""
    return np.abs(x)


# --------------------------------------------------------------------------- filterbanks
def hz_to_mel(f, htk=False):
    """librosa.hz_to_mel (Slaney / HTK)."""
    f = np.asanyarray(f, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide='ignore', invalid='ignore'):
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def mel_to_hz(m, htk=False):
    """librosa.mel_to_hz (Slaney / HTK)."""
    m = np.asanyarray(m, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def _normalize(S, norm, axis):
    """librosa.util.normalize(S, norm=p, axis) for finite p>0, fill=None."""
    mag = np.abs(S).astype(np.float64 if S.dtype == np.float64 else S.dtype)
    if norm == np.inf:
        length = np.max(mag, axis=axis, keepdims=True)
    else:
        length = np.sum(mag ** norm, axis=axis, keepdims=True) ** (1.0 / norm)
    small = length < np.finfo(S.dtype).tiny
    length = np.where(small, 1.0, length)
    return S / length


def filterbank_mel(sample_rate, n_freq, n_mels=128, f_min=0.0, f_max=None, htk=False,
                   norm='slaney', dtype=np.float32):
    """kapre/backend.py:197-231 -> librosa.filters.mel(sr, n_fft=(n_freq-1)*2, ...).astype(floatx).T

    librosa >= 0.11 algorithm: triangular ramps on the FFT-bin frequencies, computed in
    float64 and stored into a float32 array; 'slaney' area normalisation 2/(f[i+2]-f[i]);
    numeric ``norm`` -> librosa.util.normalize along the frequency axis.  Returns (n_freq, n_mels).
    """
    n_fft = (n_freq - 1) * 2
    if f_max is None:
        f_max = float(sample_rate) / 2
    weights = np.zeros((int(n_mels), n_freq), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sample_rate)
    mels = np.linspace(hz_to_mel(f_min, htk), hz_to_mel(f_max, htk), int(n_mels) + 2)
    mel_f = mel_to_hz(mels, htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(int(n_mels)):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if isinstance(norm, str):
        if norm != 'slaney':
            raise ValueError('Unsupported norm=%r' % (norm,))
        enorm = 2.0 / (mel_f[2:int(n_mels) + 2] - mel_f[:int(n_mels)])
        weights *= enorm[:, np.newaxis]
    elif norm is not None:
        weights = _normalize(weights, norm, axis=-1)
    return np.ascontiguousarray(weights.astype(dtype).T)


def filterbank_log(sample_rate, n_freq, n_bins=84, bins_per_octave=12, f_min=None,
                   spread=0.125, dtype=np.float32):
    """kapre/backend.py:234-299 (log-normal constant-Q approximation).  Returns (n_freq, n_bins)."""
    if f_min is None:
        f_min = 32.70319566
    f_max = f_min * 2 ** (n_bins / bins_per_octave)
    if f_max > sample_rate // 2:
        raise RuntimeError('Maximum frequency of log filterbank should be lower or equal to the '
                           'maximum frequency of the input')  # backend.py:267-275
    sigma = float(spread) / bins_per_octave
    basis = np.zeros((n_bins, n_freq))
    fftf = np.fft.rfftfreq(n=(n_freq - 1) * 2, d=1.0 / sample_rate)
    log_freqs = np.log2(fftf[1:])
    for i in range(n_bins):
        c_freq = f_min * (2.0 ** (float(i) / bins_per_octave))
        basis[i, 1:] = np.exp(-0.5 * ((log_freqs - np.log2(c_freq)) / sigma) ** 2
                              - np.log2(sigma) - log_freqs)
    basis = _normalize(basis, 1, axis=1)
    return np.ascontiguousarray(basis.astype(dtype).T)


def apply_filterbank(x, filterbank, data_format='default'):
    """kapre.ApplyFilterbank.call, kapre/time_frequency.py:535-548:
    tensordot over the frequency axis (3 for channels_first, 2 for channels_last) and, for
    channels_last, the (0,1,3,2) transpose back."""
    df = resolve_data_format(data_format)
    fb = np.asarray(filterbank, dtype=x.dtype)
    if df == CH_FIRST:
        return np.tensordot(x, fb, axes=(3, 0))  # (b, ch, t, new_f)
    out = np.tensordot(x, fb, axes=(2, 0))  # (b, t, ch, new_f)
    return np.transpose(out, (0, 1, 3, 2))


def magnitude_to_decibel(x, ref_value=1.0, amin=1e-5, dynamic_range=80.0):
    """kapre/backend.py:126-194.  10*log10(max(x, amin)) - 10*log10(max(amin, ref_value)),
    then clamp to (per-batch-item max over ALL non-batch axes) - dynamic_range.  It is
    10*log10 of a *magnitude* on purpose -- do not 'fix'."""
    if ref_value <= 0:
        raise ValueError('ref_value must be positive, got: %s' % ref_value)
    if amin <= 0:
        raise ValueError('amin must be positive, got: %s' % amin)
    if dynamic_range <= 0:
        raise ValueError('dynamic_range must be positive, got: %s' % dynamic_range)
    x = np.asarray(x)
    dt = x.dtype
    ln10 = np.log(np.asarray(10, dtype=dt))
    amin_t = np.asarray(amin, dtype=dt)
    log_spec = (10.0 * (np.log(np.maximum(x, amin_t)) / ln10)).astype(dt)
    log_spec = log_spec - (10.0 * (np.log(np.maximum(amin_t, np.asarray(ref_value, dtype=dt))) / ln10)).astype(dt)
    axis = tuple(range(x.ndim))[1:] if x.ndim > 1 else None
    mx = np.max(log_spec, axis=axis, keepdims=True)
    return np.maximum(log_spec, mx - np.asarray(dynamic_range, dtype=dt)).astype(dt)


# --------------------------------------------------------------------------- inverse STFT
def inverse_stft_window(win_length, hop, forward_window):
    """tf.signal.inverse_stft_window_fn(frame_step, forward_window_fn) -- the dual window
    w / sum_k w^2[n + k*hop] used at kapre/time_frequency.py:278-280."""
    w = np.asarray(forward_window, dtype=np.float64)
    overlaps = -(-win_length // hop)
    den = np.zeros(overlaps * hop)
    den[:win_length] = w ** 2
    den = den.reshape(overlaps, hop).sum(0, keepdims=True)
    den = np.tile(den, (overlaps, 1)).reshape(overlaps * hop)
    with np.errstate(divide='ignore', invalid='ignore'):     # zero window samples give inf / nan, as in TF
        return w / den[:win_length]


def inverse_stft_frames(stfts, n_fft, win_length, hop, dual_window, dtype=np.float64):
    """tf.signal.inverse_stft on (..., T, F): irfft(n_fft) -> keep first win_length samples
    (zero-pad if n_fft < win_length) -> * dual window -> overlap-add.
    kapre/time_frequency.py:307-314.  Output length (T-1)*hop + win_length."""
    s = np.asarray(stfts)
    T = s.shape[-2]
    fr = np.fft.irfft(s.astype(np.complex128), n=n_fft, axis=-1)
    if n_fft >= win_length:
        fr = fr[..., :win_length]
    else:
        fr = np.pad(fr, [(0, 0)] * (fr.ndim - 1) + [(0, win_length - n_fft)])
    fr = fr * np.asarray(dual_window, dtype=np.float64)
    out_len = (T - 1) * hop + win_length if T > 0 else 0
    out = np.zeros(s.shape[:-2] + (out_len,), dtype=np.float64)
    for t in range(T):
        out[..., t * hop:t * hop + win_length] += fr[..., t, :]
    return out.astype(dtype)


def istft_layer(x, n_fft=2048, win_length=None, hop_length=None, forward_window_name=None,
                input_data_format='default', output_data_format='default', dtype=np.float64):
    """kapre.InverseSTFT.call, kapre/time_frequency.py:289-319 (+ __init__ :269-280)."""
    idf = resolve_data_format(input_data_format)
    odf = resolve_data_format(output_data_format)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    x = np.asarray(x)
    if idf == CH_LAST:
        x = np.transpose(x, (0, 3, 1, 2))  # :304-305 -> (b, ch, t, f)
    fwd = get_window(forward_window_name, win_length, dtype=np.float64)
    dual = inverse_stft_window(win_length, hop_length, fwd)
    y = inverse_stft_frames(x, n_fft, win_length, hop_length, dual, dtype=dtype)  # (b, ch, time)
    if odf == CH_LAST:
        y = np.transpose(y, (0, 2, 1))  # :316-317
    return y


# --------------------------------------------------------------------------- composed
def stft_magnitude_layer(x, n_fft=2048, win_length=None, hop_length=None, window_name=None,
                         pad_begin=False, pad_end=False, return_decibel=False, db_amin=1e-5,
                         db_ref_value=1.0, db_dynamic_range=80.0, input_data_format='default',
                         output_data_format='default', dtype=np.float64):
    """kapre.composed.get_stft_magnitude_layer, kapre/composed.py:32-135."""
    s = magnitude(stft_layer(x, n_fft, win_length, hop_length, window_name, pad_begin, pad_end,
                             input_data_format, output_data_format, dtype=dtype))
    if return_decibel:
        s = magnitude_to_decibel(s, db_ref_value, db_amin, db_dynamic_range)
    return s


def melspectrogram_layer(x, n_fft=2048, win_length=None, hop_length=None, window_name=None,
                         pad_begin=False, pad_end=False, sample_rate=22050, n_mels=128,
                         mel_f_min=0.0, mel_f_max=None, mel_htk=False, mel_norm='slaney',
                         return_decibel=False, db_amin=1e-5, db_ref_value=1.0,
                         db_dynamic_range=80.0, input_data_format='default',
                         output_data_format='default', dtype=np.float64):
    """kapre.composed.get_melspectrogram_layer, kapre/composed.py:138-261."""
    s = magnitude(stft_layer(x, n_fft, win_length, hop_length, window_name, pad_begin, pad_end,
                             input_data_format, output_data_format, dtype=dtype))
    fb = filterbank_mel(sample_rate, n_fft // 2 + 1, n_mels, mel_f_min, mel_f_max, mel_htk,
                        mel_norm, dtype=np.float32)  # floatx constant, :241-252
    s = apply_filterbank(s, fb.astype(s.dtype), output_data_format)
    if return_decibel:
        s = magnitude_to_decibel(s, db_ref_value, db_amin, db_dynamic_range)
    return s


def phase(x):
    """kapre.Phase.call with approx_atan_accuracy=None, kapre/time_frequency.py:391-402 (tf.math.angle)."""
    return np.angle(x)


def stft_mag_phase_layer(x, n_fft=2048, win_length=None, hop_length=None, window_name=None, pad_begin=False,
                         pad_end=False, return_decibel=False, db_amin=1e-5, db_ref_value=1.0,
                         db_dynamic_range=80.0, input_data_format='default', output_data_format='default',
                         dtype=np.float64):
    """kapre.composed.get_stft_mag_phase, kapre/composed.py:420-511: magnitude (optionally dB) and
    phase concatenated on the channel axis (1 for 'channels_first', else 3 -- note the reference
    compares the *unresolved* string at :504, so 'default' concatenates on axis 3 = channels_last)."""
    s = stft_layer(x, n_fft, win_length, hop_length, window_name, pad_begin, pad_end, input_data_format,
                   output_data_format, dtype=dtype)
    mag = magnitude(s)
    ph = phase(s)
    if return_decibel:
        mag = magnitude_to_decibel(mag, db_ref_value, db_amin, db_dynamic_range)
    ch_axis = 1 if output_data_format == CH_FIRST else 3
    return np.concatenate([mag, ph.astype(mag.dtype)], axis=ch_axis)


def delta(x, win_length=5, mode='symmetric', data_format='default'):
    """kapre.Delta.call, kapre/time_frequency.py:613-636: tf.pad(mode) over time, correlate with
    arange(-n, n+1), divide by 2*sum(m^2)."""
    df = resolve_data_format(data_format)
    x = np.asarray(x, dtype=np.float64)
    if df == CH_FIRST:
        x = np.transpose(x, (0, 2, 3, 1))
    n = (win_length - 1) // 2
    denom = 2 * sum(m * m for m in range(1, n + 1))
    xp = np.pad(x, [(0, 0), (n, n), (0, 0), (0, 0)], mode=mode.lower())
    T = x.shape[1]
    out = np.zeros_like(x)
    for j, m in enumerate(range(-n, n + 1)):
        out += m * xp[:, j:j + T]
    out /= denom
    if df == CH_FIRST:
        out = np.transpose(out, (0, 3, 1, 2))
    return out


def frame_layer(x, frame_length, hop_length, pad_end=False, pad_value=0, data_format='default'):
    """kapre.Frame.call, kapre/signal.py:88-104 (tf.signal.frame along the time axis)."""
    df = resolve_data_format(data_format)
    x = np.asarray(x, dtype=np.float64)
    if df == CH_LAST:
        x = np.transpose(x, (0, 2, 1))
    L = x.shape[-1]
    T = num_frames(L, frame_length, hop_length, pad_end)
    need = (T - 1) * hop_length + frame_length if T > 0 else 0
    if need > L:
        x = np.pad(x, [(0, 0), (0, 0), (0, need - L)], constant_values=pad_value)
    idx = np.arange(T)[:, None] * hop_length + np.arange(frame_length)[None, :]
    fr = x[..., idx]                       # (b, ch, T, frame_length)
    if df == CH_LAST:
        fr = np.transpose(fr, (0, 2, 3, 1))  # (b, T, frame_length, ch)
    return fr


def energy_layer(x, sample_rate=22050, ref_duration=0.1, frame_length=2205, hop_length=1102, pad_end=False,
                 pad_value=0, data_format='default'):
    """kapre.Energy.call, kapre/signal.py:181-212."""
    df = resolve_data_format(data_format)
    fr = frame_layer(x, frame_length, hop_length, pad_end, pad_value, data_format)
    e = np.sum(fr ** 2, axis=2 if df == CH_LAST else 3)
    return ref_duration / (frame_length / sample_rate) * e


def logmel_to_mfcc(x, n_mfccs=20, data_format='default'):
    """kapre.LogmelToMFCC.call, kapre/signal.py:418-437 -> tf.signal.mfccs_from_log_mel_spectrograms:
    dct(type=2)(x) * rsqrt(2 * n_mels), i.e. sqrt(2/N) * sum_n x[n] cos(pi k (2n+1) / (2N)); first n_mfccs."""
    df = resolve_data_format(data_format)
    x = np.asarray(x, dtype=np.float64)
    if df == CH_LAST:
        x = np.transpose(x, (0, 1, 3, 2))
    N = x.shape[-1]
    n = np.arange(N)[:, None]
    k = np.arange(N)[None, :]
    mat = 2.0 * np.cos(np.pi * k * (2 * n + 1) / (2.0 * N)) / np.sqrt(2.0 * N)
    y = (x @ mat)[..., :n_mfccs]
    if df == CH_LAST:
        y = np.transpose(y, (0, 1, 3, 2))
    return y


def dft_stage1_32x32(x, n_fft=1024, hop=256):
    """First stage of the 32 x 32 Cooley-Tukey split of the real FFT inside ``tf.signal.stft``
    (kapre/time_frequency.py:174-182), float64: S[i, f, n2, k1] = sum_n1 x[i, hop f + 32 n1 + n2] exp(-2 pi i n1 k1 / 32),
    k1 = 0..16.  Checker for the tensor-core prototype (kapre_b200/csrc/tc_dft.cuh); window-free by design."""
    x = np.asarray(x, dtype=np.float64)
    T = 1 + (x.shape[1] - n_fft) // hop
    idx = hop * np.arange(T)[:, None] + np.arange(n_fft)[None, :]
    fr = x[:, idx].reshape(x.shape[0], T, 32, 32)            # (i, f, n1, n2)
    n1 = np.arange(32)[:, None]
    k1 = np.arange(17)[None, :]
    F = np.exp(-2j * np.pi * n1 * k1 / 32.0)                  # (n1, k1)
    return np.einsum('ifab,ak->ifbk', fr, F)                  # (i, f, n2, k1)


def concat_frequency_map(x, data_format=CH_DEFAULT):
    """kapre/time_frequency.py:707-733: one more channel holding linspace(0, 1, n_freq) (float32, as the reference casts it)
    along the frequency axis of a (b, t, f, ch) / (b, ch, t, f) batch."""
    x = np.asarray(x)
    fmt = resolve_data_format(data_format)
    F = x.shape[2] if fmt == CH_LAST else x.shape[3]
    # tf.linspace on Python floats is a float32 op (LinSpace kernel): start + step * i in float32, last element = stop
    step = np.float32(1.0) / np.float32(F - 1) if F > 1 else np.float32(0.0)
    m = (np.arange(F, dtype=np.float32) * step).astype(np.float32)
    if F > 1:
        m[-1] = np.float32(1.0)
    if fmt == CH_LAST:
        B, T, _, C = x.shape
        return np.concatenate([x, np.broadcast_to(m.reshape(1, 1, F, 1), (B, T, F, 1)).astype(x.dtype)], axis=3)
    B, C, T, _ = x.shape
    return np.concatenate([x, np.broadcast_to(m.reshape(1, 1, 1, F), (B, 1, T, F)).astype(x.dtype)], axis=1)


def spec_augment(x, time_masks, freq_masks, mask_value=0.0, data_format=CH_DEFAULT):
    """kapre/augmentation.py:205-260 with the random draws given: per item, elements whose time index lies in
    [start, start + width] of any of its time masks, or whose frequency index lies in such a range of any of its frequency
    masks, become mask_value.  Masks: (batch, n, 2) integer (start, width)."""
    x = np.array(x, copy=True)
    fmt = resolve_data_format(data_format)
    t_ax, f_ax = (1, 2) if fmt == CH_LAST else (2, 3)
    T, F = x.shape[t_ax], x.shape[f_ax]
    for b in range(x.shape[0]):
        mt = np.zeros(T, dtype=bool)
        for s0, w in np.asarray(time_masks)[b].reshape(-1, 2):
            mt |= (np.arange(T) >= s0) & (np.arange(T) <= s0 + w)
        mf = np.zeros(F, dtype=bool)
        for s0, w in np.asarray(freq_masks)[b].reshape(-1, 2):
            mf |= (np.arange(F) >= s0) & (np.arange(F) <= s0 + w)
        m = mt[:, None] | mf[None, :]
        if fmt == CH_LAST:
            x[b][m] = mask_value
        else:
            x[b][:, m] = mask_value
    return x

