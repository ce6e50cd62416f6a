"""A NumPy stand-in for the slice of TensorFlow / Keras / librosa that kapre's hot path calls.

TEST INFRASTRUCTURE ONLY, used by make_golden_ref.py in the build container: TensorFlow and librosa
are not installed (no network), so the reference package cannot be imported as is.  With these
modules registered in ``sys.modules`` the UNMODIFIED reference sources under /root/reference/kapre
import and run, so that every line of kapre's own logic (the transposes, ``pad_begin``, the call
arguments of ``tf.signal.stft``, the decibel formula and its per-item clamp, the filterbank
contraction axis, the dual window wiring, Delta's padding + correlation, ...) is executed by the
reference itself.  What is NOT the reference here is the inside of ``tf.signal.*`` / ``librosa.*``:
those are re-stated below from the published algorithms (tensorflow 2.16-2.20 ``tf.signal``, librosa
0.11 ``filters.mel``) and are cross-checked against torch.stft / torchaudio / scipy in
make_golden.py.  Arithmetic follows the input dtype, so float64 inputs give float64 "truth".
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np


def _mod(name):
    m = types.ModuleType(name)
    m.__path__ = []   # let "import a.b" work
    sys.modules[name] = m
    return m


# ------------------------------------------------------------------------------------ tf core
def _asarray(x, dtype=None):
    return np.asarray(x, dtype=dtype)


def _np_dtype(d):
    return np.dtype(d) if d is not None else None


def _pad(x, paddings, mode='CONSTANT', constant_values=0):
    mode = mode.lower()
    paddings = [tuple(int(v) for v in p) for p in np.asarray(paddings)]
    if mode == 'constant':
        return np.pad(x, paddings, mode='constant', constant_values=constant_values)
    return np.pad(x, paddings, mode=mode)   # 'reflect' / 'symmetric' mean the same in tf.pad and np.pad


def _tensordot(a, b, axes):
    return np.tensordot(a, b, axes=axes)


def _angle(x):
    return np.angle(x)


# ------------------------------------------------------------------------------------ tf.signal
def _periodic_raised_cosine(n, a, b, periodic, dtype):
    # tf.signal.window_ops._raised_cosine_window: even = 1 - n % 2; N = n + periodic * even - 1
    if n == 1:
        return np.ones([1], dtype=dtype)
    even = 1 - n % 2
    denom = n + int(periodic) * even - 1
    k = np.arange(n, dtype=np.float64)
    return (a - b * np.cos(2.0 * np.pi * k / denom)).astype(dtype)


def hann_window(window_length, periodic=True, dtype=np.float32, name=None):
    return _periodic_raised_cosine(int(window_length), 0.5, 0.5, periodic, _np_dtype(dtype))


def hamming_window(window_length, periodic=True, dtype=np.float32, name=None):
    return _periodic_raised_cosine(int(window_length), 0.54, 0.46, periodic, _np_dtype(dtype))


def kaiser_window(window_length, beta=12.0, dtype=np.float32, name=None):
    n = int(window_length)
    if n == 1:
        return np.ones([1], dtype=dtype)
    k = np.arange(n, dtype=np.float64)
    r = 2.0 * k / (n - 1) - 1.0
    return (np.i0(beta * np.sqrt(np.maximum(0.0, 1.0 - r * r))) / np.i0(beta)).astype(dtype)


def kaiser_bessel_derived_window(window_length, beta=12.0, dtype=np.float32, name=None):
    n = int(window_length)
    half = n // 2
    kw = kaiser_window(half + 1, beta, np.float64)
    cs = np.cumsum(kw)
    h = np.sqrt(cs[:-1] / cs[-1])
    return np.concatenate([h, h[::-1]]).astype(dtype)


def vorbis_window(window_length, dtype=np.float32, name=None):
    n = int(window_length)
    k = np.arange(n, dtype=np.float64)
    return np.sin(np.pi / 2.0 * np.sin(np.pi / n * (k + 0.5)) ** 2).astype(dtype)


def frame(signal, frame_length, frame_step, pad_end=False, pad_value=0, axis=-1, name=None):
    """tf.signal.frame: N = 1 + (L - frame_length) // step, or ceil(L / step) with pad_end."""
    x = np.asarray(signal)
    axis = axis % x.ndim
    x = np.moveaxis(x, axis, -1)
    length = x.shape[-1]
    frame_length, frame_step = int(frame_length), int(frame_step)
    if pad_end:
        n = -(-length // frame_step)
        need = (n - 1) * frame_step + frame_length if n > 0 else 0
        if need > length:
            x = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(0, need - length)], constant_values=pad_value)
    else:
        n = max(0, 1 + (length - frame_length) // frame_step) if length >= frame_length else 0
    idx = np.arange(n)[:, None] * frame_step + np.arange(frame_length)[None, :]
    out = x[..., idx]                              # (..., n, frame_length)
    # put (n, frame_length) where `axis` was
    src = list(range(out.ndim))
    lead = src[:axis]
    rest = src[axis:-2]
    order = lead + [out.ndim - 2, out.ndim - 1] + rest
    return np.transpose(out, order)


def _complex_of(dtype):
    return np.complex128 if np.dtype(dtype) == np.float64 else np.complex64


def stft(signals, frame_length, frame_step, fft_length=None, window_fn=hann_window, pad_end=False, name=None):
    """tf.signal.stft: frame -> * window_fn(frame_length) -> rfft(fft_length) (zero-pad or crop on the right)."""
    x = np.asarray(signals)
    if fft_length is None:
        fft_length = 1 << int(math.ceil(math.log2(frame_length)))
    fr = frame(x, frame_length, frame_step, pad_end=pad_end)
    if window_fn is not None:
        fr = fr * np.asarray(window_fn(frame_length, dtype=x.dtype))
    if x.dtype == np.float64:
        return np.fft.rfft(fr, n=int(fft_length), axis=-1)
    # float32 working precision like TF's cuFFT / Eigen rfft (only used by the CPU-baseline comparisons)
    import scipy.fft
    return scipy.fft.rfft(fr.astype(np.float32), n=int(fft_length), axis=-1)


def overlap_and_add(signal, frame_step):
    x = np.asarray(signal)
    n, fl = x.shape[-2], x.shape[-1]
    out = np.zeros(x.shape[:-2] + ((n - 1) * frame_step + fl,), dtype=x.dtype)
    for i in range(n):
        out[..., i * frame_step:i * frame_step + fl] += x[..., i, :]
    return out


def inverse_stft(stfts, frame_length, frame_step, fft_length=None, window_fn=hann_window, name=None):
    """tf.signal.inverse_stft: irfft(fft_length) -> first frame_length samples (or zero-pad) -> * window ->
    overlap_and_add(frame_step)."""
    s = np.asarray(stfts)
    if fft_length is None:
        fft_length = 1 << int(math.ceil(math.log2(frame_length)))
    real = np.fft.irfft(s, n=int(fft_length), axis=-1)
    if s.dtype == np.complex64:
        real = real.astype(np.float32)
    if fft_length >= frame_length:
        real = real[..., :frame_length]
    else:
        real = np.pad(real, [(0, 0)] * (real.ndim - 1) + [(0, frame_length - fft_length)])
    if window_fn is not None:
        real = real * np.asarray(window_fn(frame_length, dtype=real.dtype))
    return overlap_and_add(real, int(frame_step))


def inverse_stft_window_fn(frame_step, forward_window_fn=hann_window, name=None):
    """tf.signal.inverse_stft_window_fn: forward window / (sum over the overlapping hops of its square)."""

    def inverse_window(frame_length, dtype=np.float32):
        fw = np.asarray(forward_window_fn(frame_length, dtype=dtype))
        den = fw * fw
        overlaps = -(-frame_length // frame_step)
        den = np.pad(den, (0, overlaps * frame_step - frame_length))
        den = den.reshape(overlaps, frame_step).sum(0, keepdims=True)
        den = np.tile(den, (overlaps, 1)).reshape(overlaps * frame_step)
        return fw / den[:frame_length]

    return inverse_window


def mfccs_from_log_mel_spectrograms(log_mel_spectrograms, name=None):
    """tf.signal.mfccs_from_log_mel_spectrograms: dct(type=2) * rsqrt(2 * num_mel_bins)."""
    import scipy.fft
    x = np.asarray(log_mel_spectrograms)
    n = x.shape[-1]
    return scipy.fft.dct(x, type=2, axis=-1) / np.sqrt(2.0 * n)


# ------------------------------------------------------------------------------------ keras
class _Sym:
    """Node of a functional-API graph (keras.Input / layer(sym)): evaluated lazily by Model.__call__."""

    def __init__(self, layer=None, parents=()):
        self.layer, self.parents = layer, tuple(parents)

    def evaluate(self, feed):
        if self in feed:
            return feed[self]
        args = [p.evaluate(feed) for p in self.parents]
        val = self.layer.call(args if getattr(self.layer, '_takes_list', False) else args[0])
        feed[self] = val
        return val


def Input(shape=None, **kwargs):
    return _Sym()


class Layer:
    _counters = {}

    def __init__(self, name=None, trainable=True, dtype=None, input_shape=None, **kwargs):
        cls = type(self).__name__.lower()
        if name is None:
            i = Layer._counters.get(cls, 0)
            Layer._counters[cls] = i + 1
            name = cls if i == 0 else '%s_%d' % (cls, i)
        self.name = name
        self.trainable = trainable

    def __call__(self, x, *a, **k):
        if isinstance(x, _Sym):
            return _Sym(self, [x])
        if isinstance(x, (list, tuple)) and x and all(isinstance(v, _Sym) for v in x):
            return _Sym(self, list(x))
        return self.call(x, *a, **k)

    def get_config(self):
        return {'name': self.name, 'trainable': self.trainable}


class Concatenate(Layer):
    _takes_list = True

    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, xs):
        return np.concatenate(list(xs), axis=self.axis)


class Model(Layer):
    """keras.Model(inputs=sym, outputs=sym): runs the recorded layer graph eagerly."""

    def __init__(self, inputs=None, outputs=None, name=None):
        super().__init__(name=name)
        self.inputs, self.outputs = inputs, outputs

    def call(self, x):
        return self.outputs.evaluate({self.inputs: x})


class Sequential(Layer):
    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        self.layers = list(layers or [])

    def add(self, layer):
        self.layers.append(layer)

    def call(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


def register_keras_serializable(package='Custom', name=None):
    def deco(cls):
        return cls
    return deco


def _conv2d(x, kernel, strides=(1, 1), padding='valid', data_format=None, dilation_rate=(1, 1)):
    """K.conv2d for the one case kapre uses (Delta): a (k, 1, 1, 1) kernel, 'valid', channels_last.
    Keras convolution is cross-correlation."""
    x = np.asarray(x)
    kernel = np.asarray(kernel)
    assert kernel.shape[1:] == (1, 1, 1) and padding == 'valid' and data_format == 'channels_last'
    assert x.shape[-1] == 1 or True
    k = kernel.shape[0]
    t_out = x.shape[1] - k + 1
    out = np.zeros((x.shape[0], t_out) + x.shape[2:], dtype=np.result_type(x, kernel))
    for j in range(k):
        out += kernel[j, 0, 0, 0] * x[:, j:j + t_out]
    return out


# ------------------------------------------------------------------------------------ librosa
def _hz_to_mel(f, htk=False):
    f = np.asanyarray(f, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m, htk=False):
    m = np.asanyarray(m, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def _fft_frequencies(sr=22050, n_fft=2048):
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def _librosa_mel(*, sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False, norm='slaney', dtype=np.float32):
    """librosa.filters.mel (0.10 / 0.11): triangles on the mel scale, 'slaney' area normalisation."""
    if fmax is None:
        fmax = float(sr) / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=dtype)
    fftfreqs = _fft_frequencies(sr=sr, n_fft=n_fft)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin, htk), _hz_to_mel(fmax, htk), n_mels + 2), htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    if norm == 'slaney':
        enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
        weights *= enorm[:, np.newaxis]
    return weights


def _librosa_normalize(S, norm=np.inf, axis=0):
    S = np.asarray(S)
    mag = np.abs(S).astype(np.float64)
    if norm == 1:
        length = mag.sum(axis=axis, keepdims=True)
    elif norm == np.inf:
        length = mag.max(axis=axis, keepdims=True)
    else:
        length = (mag ** norm).sum(axis=axis, keepdims=True) ** (1.0 / norm)
    tiny = np.finfo(S.dtype if np.issubdtype(S.dtype, np.floating) else np.float32).tiny
    length = np.where(length < tiny, 1.0, length)
    return S / length


# ------------------------------------------------------------------------------------ install
def install():
    """Register the stand-in modules.  Call BEFORE importing the reference package."""
    if 'tensorflow' in sys.modules and not getattr(sys.modules['tensorflow'], '_kapre_standin', False):
        raise RuntimeError('a real tensorflow is already imported')
    tf = _mod('tensorflow')
    tf._kapre_standin = True
    tf.Tensor = np.ndarray
    tf.float32, tf.float64, tf.int32, tf.complex64 = np.float32, np.float64, np.int32, np.complex64
    tf.constant = lambda v, dtype=None, **k: _asarray(v, _np_dtype(dtype))
    tf.convert_to_tensor = lambda v, dtype=None, **k: _asarray(v, _np_dtype(dtype))
    tf.cast = lambda x, dtype: _asarray(x).astype(_np_dtype(dtype))
    tf.transpose = lambda x, perm=None, **k: np.transpose(x, perm)
    tf.pad = lambda x, paddings, mode='CONSTANT', constant_values=0, **k: _pad(x, paddings, mode, constant_values)
    tf.shape = lambda x: np.asarray(np.shape(x))
    tf.reshape = lambda x, s: np.reshape(x, tuple(int(v) for v in s))
    tf.tile = lambda x, m: np.tile(x, tuple(int(v) for v in m))
    tf.concat = lambda xs, axis: np.concatenate(xs, axis=axis)
    tf.linspace = lambda start, stop, num: np.linspace(start, stop, int(num))
    tf.tensordot = _tensordot
    tf.abs = np.abs
    tf.print = lambda *a, **k: None

    def _function(fn=None, **k):      # tf.function, with or without arguments: run eagerly
        return fn if fn is not None else (lambda f: f)
    tf.function = _function
    tf.bool, tf.float16, tf.int64 = np.bool_, np.float16, np.int64
    tf.where = lambda c, a, b: np.where(c, a, b)
    tf.logical_and, tf.equal = np.logical_and, np.equal
    tf.zeros = lambda shape, dtype=np.float32: np.zeros(tuple(int(v) for v in np.atleast_1d(shape)), _np_dtype(dtype))
    tf.repeat = lambda x, repeats, axis=None: np.repeat(x, repeats, axis=axis)
    tf.gather = lambda x, idx, axis=0, **k: np.take(x, np.asarray(idx), axis=axis)
    tf.range = lambda start, limit=None, delta=1, dtype=None: np.arange(start, limit, delta, dtype=_np_dtype(dtype))
    tf.matmul = np.matmul
    tf.expand_dims = lambda x, axis: np.expand_dims(x, axis)
    tf.stack = lambda xs, axis=0: np.stack(xs, axis=axis)
    tf.slice = lambda x, begin, size: np.asarray(x)[tuple(slice(int(b), int(b) + int(n)) for b, n in zip(begin, size))]
    tf.rank = lambda x: np.ndim(x)
    tf.norm = lambda x, **k: np.linalg.norm(x)
    m = _mod('tensorflow.math')
    tf.math = m
    m.log, m.log1p, m.exp, m.abs, m.sign, m.square = np.log, np.log1p, np.exp, np.abs, np.sign, np.square
    m.maximum = np.maximum
    m.real, m.imag, m.angle = np.real, np.imag, _angle
    m.reduce_max = lambda x, axis=None, keepdims=False: np.max(x, axis=axis, keepdims=keepdims)
    m.reduce_any = lambda x, axis=None: np.any(x, axis=axis)
    m.reduce_sum = lambda x, axis=None, keepdims=False: np.sum(x, axis=axis, keepdims=keepdims)
    s = _mod('tensorflow.signal')
    tf.signal = s
    for fn in (hann_window, hamming_window, kaiser_window, kaiser_bessel_derived_window, vorbis_window, frame,
               stft, inverse_stft, inverse_stft_window_fn, overlap_and_add, mfccs_from_log_mel_spectrograms):
        setattr(s, fn.__name__, fn)

    keras = _mod('tensorflow.keras')
    tf.keras = keras
    K = _mod('tensorflow.keras.backend')
    keras.backend = K
    K.image_data_format = lambda: 'channels_last'
    K.floatx = lambda: 'float32'
    K.ndim = lambda x: np.ndim(x)
    K.permute_dimensions = lambda x, pattern: np.transpose(x, pattern)
    K.arange = lambda start, stop=None, step=1, dtype='int32': np.arange(start, stop, step, dtype=dtype)
    K.reshape = lambda x, shape: np.reshape(x, shape)
    K.conv2d = _conv2d
    K.cast_to_floatx = lambda x: np.asarray(x, dtype=np.float32)
    layers = _mod('tensorflow.keras.layers')
    keras.layers = layers
    layers.Layer = Layer
    utils = _mod('tensorflow.keras.utils')
    keras.utils = utils
    utils.register_keras_serializable = register_keras_serializable
    keras.Sequential = Sequential
    keras.Model = Model
    keras.Input = Input
    layers.Concatenate = Concatenate
    sys.modules['keras'] = keras   # "from tensorflow import keras" / "import keras.config" fallbacks

    librosa = _mod('librosa')
    filters = _mod('librosa.filters')
    librosa.filters = filters
    filters.mel = _librosa_mel
    util = _mod('librosa.util')
    librosa.util = util
    util.normalize = _librosa_normalize
    librosa.fft_frequencies = lambda sr=22050, n_fft=2048: _fft_frequencies(sr=sr, n_fft=n_fft)
    return tf
