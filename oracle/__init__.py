"""CPU oracle for the kapre STFT -> |.| -> filterbank -> dB hot path (and InverseSTFT).

TEST INFRASTRUCTURE ONLY.  Nothing under ``kapre_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs use it, and only as the checker / the baseline, never as the product.

PARITY PINNING STATUS: *unpinned by the reference's own outputs*.  The reference
(kapre 0.4.0) is pure Python over TensorFlow + librosa; neither is installed in the build
container nor on the GPU box and there is no network, and the reference repository stores
no golden output vectors (its tests compare against librosa evaluated at test time).  The
restatement is therefore pinned against what IS available:

* the reference's literal dB test matrix (``tests/test_backend.py:20-22``) evaluated with
  the closed-form ``10*log10`` formula,
* independent implementations present in the container: ``torch.stft(center=False)``,
  ``torchaudio.functional.melscale_fbanks``, ``scipy.signal.get_window``, ``numpy.fft``,
* analytic known-answer tests (impulse, DC, bin-centred cosine, Parseval, STFT->ISTFT identity),
* the in-repo matmul restatement of ``tf.signal.stft`` (``kapre/tflite_compatible_stft.py``),
  whose arithmetic ``oracle.reference.stft_by_dft_matrix`` follows line by line.

See ``tests/golden/make_golden.py`` for the script that produced the committed fixtures.
"""
from .reference import *  # noqa: F401,F403
