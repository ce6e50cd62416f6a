"""Import alias so that unmodified call sites (``import kapre``; ``from kapre import STFT``;
``from kapre.composed import get_melspectrogram_layer``) resolve to the B200-native package."""
import sys

import kapre_b200
from kapre_b200 import *  # noqa: F401,F403
from kapre_b200 import augmentation, backend, composed, signal, time_frequency  # noqa: F401

__version__ = kapre_b200.__version__
sys.modules[__name__ + '.backend'] = backend
sys.modules[__name__ + '.composed'] = composed
sys.modules[__name__ + '.time_frequency'] = time_frequency
sys.modules[__name__ + '.signal'] = signal
sys.modules[__name__ + '.augmentation'] = augmentation
