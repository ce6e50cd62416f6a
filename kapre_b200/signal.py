"""Mirrors of the waveform-side layers of ``kapre/signal.py`` that sit next to the hot path
(SURVEY section 8f): ``Frame`` (reference :22-119), ``Energy`` (:123-233) and ``LogmelToMFCC``
(:365-447).  Same constructor arguments, validation and ``get_config`` keys; ``call`` runs CUDA
kernels through the C ABI."""
from __future__ import annotations

import numpy as np

from . import backend, ops
from .backend import _CH_DEFAULT_STR, _CH_FIRST_STR, _CH_LAST_STR
from .time_frequency import Layer, _unwrap_format, register_keras_serializable

__all__ = ['Frame', 'Energy', 'LogmelToMFCC']


def _resolve(fmt):
    return backend.image_data_format() if fmt == _CH_DEFAULT_STR else fmt


@register_keras_serializable(package='Kapre')
class Frame(Layer):
    """``tf.signal.frame`` along the time axis (reference kapre/signal.py:22-119):
    ``(batch, time, ch)`` -> ``(batch, frames, frame_length, ch)`` or
    ``(batch, ch, time)`` -> ``(batch, ch, frames, frame_length)``."""

    def __init__(self, frame_length, hop_length, pad_end=False, pad_value=0, data_format='default', **kwargs):
        super().__init__(**kwargs)
        backend.validate_data_format_str(data_format)
        if frame_length <= 0:
            raise ValueError('frame_length must be positive, got: %s' % (frame_length,))
        if hop_length <= 0:
            raise ValueError('hop_length must be positive, got: %s' % (hop_length,))
        if frame_length < hop_length:
            raise ValueError('frame_length (%s) must be >= hop_length (%s)' % (frame_length, hop_length))
        self.frame_length, self.hop_length = frame_length, hop_length
        self.pad_end, self.pad_value = pad_end, pad_value
        self.data_format_str = data_format
        self.data_format = _resolve(data_format)
        self.time_axis = 2 if self.data_format == _CH_FIRST_STR else 1

    def call(self, x):
        return ops.frame(x, self.frame_length, self.hop_length, self.pad_end, self.pad_value, self.data_format)

    _config_fields = (
        ('frame_length', 'frame_length'),
        ('hop_length', 'hop_length'),
        ('pad_end', 'pad_end'),
        ('pad_value', 'pad_value'),
        ('data_format', 'data_format_str'),
    )


@register_keras_serializable(package='Kapre')
class Energy(Layer):
    """Energy per frame, normalised to ``ref_duration`` (reference kapre/signal.py:123-233):
    ``ref_duration / (frame_length / sample_rate) * sum(frame ** 2)``.  Output ``(batch, frames, ch)`` or
    ``(batch, ch, frames)``.  (The reference's stray ``tf.print`` debugging, :191-207, is not reproduced.)"""

    def __init__(self, sample_rate=22050, ref_duration=0.1, frame_length=2205, hop_length=1102, pad_end=False,
                 pad_value=0, data_format='default', **kwargs):
        super().__init__(**kwargs)
        backend.validate_data_format_str(data_format)
        self.sample_rate, self.ref_duration = sample_rate, ref_duration
        self.frame_length, self.hop_length = frame_length, hop_length
        self.pad_end, self.pad_value = pad_end, pad_value
        self.data_format_str = data_format
        self.data_format = _resolve(data_format)
        self.time_axis = 2 if self.data_format == _CH_FIRST_STR else 1

    def call(self, x):
        nor_coeff = self.ref_duration / (self.frame_length / self.sample_rate)
        return ops.energy(x, self.frame_length, self.hop_length, self.pad_end, self.pad_value, nor_coeff,
                          self.data_format)

    _config_fields = (
        ('sample_rate', 'sample_rate'),
        ('ref_duration', 'ref_duration'),
        ('frame_length', 'frame_length'),
        ('hop_length', 'hop_length'),
        ('pad_end', 'pad_end'),
        ('pad_value', 'pad_value'),
        ('data_format', 'data_format_str'),
    )


def dct2_htk_matrix(n_mels: int, n_mfccs: int) -> np.ndarray:
    """``tf.signal.mfccs_from_log_mel_spectrograms`` as a matrix: unnormalised DCT-II
    ``2 sum_n x[n] cos(pi k (2n+1) / (2N))`` scaled by ``1/sqrt(2N)`` (HTK convention), first
    ``n_mfccs`` coefficients.  Shape ``(n_mels, n_mfccs)`` so that it plugs into the filterbank kernel."""
    n = np.arange(n_mels, dtype=np.float64)[:, None]
    k = np.arange(n_mfccs, dtype=np.float64)[None, :]
    return (np.sqrt(2.0 / n_mels) * np.cos(np.pi * k * (2.0 * n + 1.0) / (2.0 * n_mels))).astype(np.float32)


@register_keras_serializable(package='Kapre')
class LogmelToMFCC(Layer):
    """MFCC from a log-mel spectrogram (reference kapre/signal.py:365-447): HTK-scaled DCT-II along the
    mel axis, first ``n_mfccs`` coefficients.  ``(b, time, mel, ch)`` -> ``(b, time, n_mfccs, ch)`` or
    ``(b, ch, time, mel)`` -> ``(b, ch, time, n_mfccs)``.  The DCT is a dense matrix on the mel axis, so it
    runs on the same kernel as ``ApplyFilterbank``."""

    def __init__(self, n_mfccs=20, data_format='default', **kwargs):
        super().__init__(**kwargs)
        backend.validate_data_format_str(data_format)
        self.n_mfccs = n_mfccs
        self.data_format_str = data_format
        self.data_format = _resolve(data_format)
        self.permutation = (0, 1, 3, 2) if self.data_format == _CH_LAST_STR else None
        self._fb = {}

    def call(self, log_melgrams):
        n_mels = log_melgrams.shape[2] if self.data_format == _CH_LAST_STR else log_melgrams.shape[3]
        fb = self._fb.get(n_mels)
        if fb is None:
            fb = self._fb[n_mels] = ops.Filterbank(dct2_htk_matrix(n_mels, min(self.n_mfccs, n_mels)))
        return ops.apply_filterbank(log_melgrams, fb, self.data_format)

    _config_fields = (('n_mfccs', 'n_mfccs'), ('data_format', 'data_format_str'))
