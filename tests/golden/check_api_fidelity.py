"""Constructor / get_config fidelity of the layer mirrors against the reference's own classes.

Runs only where /root/reference exists (the build container): the unmodified reference package is imported
over tf_standin.py, and for a few hundred random keyword sets -- valid and invalid -- every mirrored layer is
constructed in both packages; the exception TYPE (if any) and the ``get_config()`` dictionaries must agree.
Invoked by tests/test_reference_fidelity.py in a subprocess (the reference package and the alias package
are both called ``kapre``).  Exit code 0 = no disagreement."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_standin  # noqa: E402

tf_standin.install()
sys.path.insert(0, '/root/reference')
import kapre as R  # noqa: E402  (the reference)

sys.path.append(os.path.join(HERE, '..', '..'))
import kapre_b200 as M  # noqa: E402

assert R.__file__.startswith('/root/reference/'), R.__file__


def compare(name, kw):
    ca = cb = ea = eb = None
    try:
        ca = getattr(R, name)(**kw).get_config()
    except Exception as e:  # noqa: BLE001
        ea = type(e).__name__
    try:
        cb = getattr(M, name)(**kw).get_config()
    except Exception as e:  # noqa: BLE001
        eb = type(e).__name__
    if ea or eb:
        if ea != eb:
            print('exception differs:', name, kw, ea, eb)
            return 1
        return 0
    for k in ('name', 'trainable', 'dtype'):
        ca.pop(k, None)
        cb.pop(k, None)
    if ca != cb:
        print('get_config differs:', name, kw, ca, cb)
        return 1
    return 0


def main(n):
    rng = np.random.default_rng(1)

    def pick(options):
        v = options[int(rng.integers(len(options)))]
        return v

    bad = 0
    for _ in range(n):
        fmt_i = pick(['channels_first', 'channels_last', 'default', 'weird', 3])
        fmt_o = pick(['channels_first', 'channels_last', 'default', 'weird', 3])
        n_fft = pick([256, 512, 1000])
        win, hop = pick([None, 200, 256]), pick([None, 64, 100])
        wn = pick([None, 'hann_window', 'hamming_window', 'nope'])
        flip = lambda: bool(rng.random() < 0.5)  # noqa: E731
        bad += compare('STFT', dict(n_fft=n_fft, win_length=win, hop_length=hop, window_name=wn, pad_begin=flip(),
                                    pad_end=flip(), input_data_format=fmt_i, output_data_format=fmt_o))
        bad += compare('InverseSTFT', dict(n_fft=n_fft, win_length=win, hop_length=hop, forward_window_name=wn,
                                           input_data_format=fmt_i, output_data_format=fmt_o))
        bad += compare('MagnitudeToDecibel', dict(ref_value=pick([1.0, 0.1]), amin=pick([1e-5, 1e-10]),
                                                  dynamic_range=pick([80.0, 40.0])))
        bad += compare('Delta', dict(win_length=pick([2, 3, 4, 5, 9]),
                                     mode=pick(['symmetric', 'reflect', 'constant', 'wrap', 'SYMMETRIC']),
                                     data_format=fmt_i))
        bad += compare('Frame', dict(frame_length=pick([100, 512]), hop_length=pick([50, 200]), pad_end=flip(),
                                     pad_value=pick([0, 1.5]), data_format=fmt_i))
        bad += compare('Energy', dict(sample_rate=pick([8000, 22050]), ref_duration=pick([0.1, 0.05]),
                                      frame_length=pick([100, 2205]), hop_length=pick([50, 1102]), pad_end=flip(),
                                      data_format=fmt_i))
        bad += compare('LogmelToMFCC', dict(n_mfccs=pick([13, 20]), data_format=fmt_i))
        bad += compare('Magnitude', {})
        bad += compare('Phase', {})
        bad += compare('ApplyFilterbank', dict(type=pick(['mel', 'log', 'tri']),
                                               filterbank_kwargs=dict(sample_rate=16000, n_freq=n_fft // 2 + 1, n_mels=8),
                                               data_format=fmt_i))
    # composed factories: same layer classes in the same order, each with the same config
    from kapre import composed as RC

    def layer_list(model):
        out = []
        for lay in model.layers:
            cfg = lay.get_config()
            for k in ('name', 'trainable', 'dtype'):
                cfg.pop(k, None)
            cfg = dict((k, (np.asarray(v).shape if hasattr(v, 'shape') else v)) for k, v in cfg.items())
            out.append((type(lay).__name__, cfg))
        return out

    for _ in range(max(10, n // 5)):
        kw = dict(n_fft=pick([256, 512, 1024]), win_length=pick([None, 200]), hop_length=pick([None, 64]),
                  window_name=pick([None, 'hamming_window']), pad_begin=flip(), pad_end=flip(),
                  return_decibel=flip(), db_amin=pick([1e-5, 1e-7]), db_ref_value=pick([1.0, 0.5]),
                  db_dynamic_range=pick([80.0, 50.0]), input_data_format=pick(['channels_first', 'channels_last', 'default']),
                  output_data_format=pick(['channels_first', 'channels_last', 'default']))
        for fn, extra in (('get_stft_magnitude_layer', {}),
                          ('get_melspectrogram_layer', dict(sample_rate=pick([16000, 22050]), n_mels=pick([40, 128]),
                                                            mel_f_min=pick([0.0, 30.0]), mel_f_max=pick([None, 7000.0]),
                                                            mel_htk=flip(), mel_norm=pick(['slaney', None]))),
                          ('get_log_frequency_spectrogram_layer', dict(sample_rate=22050, log_n_bins=pick([48, 84]),
                                                                       log_f_min=pick([None, 55.0]),
                                                                       log_bins_per_octave=pick([12, 24]),
                                                                       log_spread=pick([0.125, 0.25])))):
            a = layer_list(getattr(RC, fn)(**dict(kw, **extra)))
            b = layer_list(getattr(M, fn)(**dict(kw, **extra)))
            if a != b:
                print('factory differs:', fn, dict(kw, **extra), a, b)
                bad += 1
    # filterbank matrices: the reference's backend functions (librosa restated in the stand-in) vs this repository's
    from kapre import backend as RB
    for _ in range(20):
        sr, n_freq = pick([8000, 16000, 44100]), pick([129, 257, 513])
        kw = dict(sample_rate=sr, n_freq=n_freq, n_mels=pick([10, 40, 96]), f_min=pick([0.0, 50.0]),
                  f_max=pick([None, sr / 2.5]), htk=flip(), norm=pick(['slaney', None]))
        a, b = np.asarray(RB.filterbank_mel(**kw)), np.asarray(M.backend.filterbank_mel(**kw))
        if a.shape != b.shape or np.abs(a - b).max() > 2e-6 * max(1.0, np.abs(a).max()):
            print('filterbank_mel differs:', kw)
            bad += 1
        kw = dict(sample_rate=sr, n_freq=n_freq, n_bins=pick([24, 60]), bins_per_octave=pick([12, 24]),
                  f_min=pick([None, 40.0]), spread=pick([0.125, 0.3]))
        a, b = np.asarray(RB.filterbank_log(**kw)), np.asarray(M.backend.filterbank_log(**kw))
        if a.shape != b.shape or np.abs(a - b).max() > 2e-6 * max(1.0, np.abs(a).max()):
            print('filterbank_log differs:', kw)
            bad += 1
    print('api fidelity: %d keyword sets per layer, %d disagreements' % (n, bad))
    return bad


if __name__ == '__main__':
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 200) else 0)
