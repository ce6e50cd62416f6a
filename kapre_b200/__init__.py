"""kapre_b200 -- B200-native (sm_100a CUDA) implementation of kapre's STFT -> magnitude ->
mel filterbank -> decibel hot path and the matching inverse STFT, behind kapre's layer API.

Import surface mirrors the hot-path part of ``kapre/__init__.py`` (reference :8-36).
"""
__version__ = '0.1.0'
VERSION = __version__

from . import backend  # noqa: E402
from . import composed  # noqa: E402
from .time_frequency import (  # noqa: E402
    STFT,
    InverseSTFT,
    Magnitude,
    Phase,
    MagnitudeToDecibel,
    ApplyFilterbank,
    Delta,
    ConcatenateFrequencyMap,
    Layer,
)
from .signal import Frame, Energy, LogmelToMFCC  # noqa: E402
from .augmentation import SpecAugment  # noqa: E402
from . import augmentation  # noqa: E402
from .composed import (  # noqa: E402
    Sequential,
    get_stft_magnitude_layer,
    get_melspectrogram_layer,
    get_log_frequency_spectrogram_layer,
    get_perfectly_reconstructing_stft_istft,
    get_stft_mag_phase,
)

__all__ = [
    '__version__', 'VERSION', 'backend', 'composed',
    'STFT', 'InverseSTFT', 'Magnitude', 'Phase', 'MagnitudeToDecibel', 'ApplyFilterbank',
    'Delta', 'ConcatenateFrequencyMap', 'Frame', 'Energy', 'LogmelToMFCC', 'SpecAugment', 'augmentation', 'Layer', 'Sequential',
    'get_stft_magnitude_layer', 'get_melspectrogram_layer', 'get_log_frequency_spectrogram_layer',
    'get_perfectly_reconstructing_stft_istft', 'get_stft_mag_phase',
]
