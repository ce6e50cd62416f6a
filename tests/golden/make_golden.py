"""Generates tests/golden/kapre_cases.npz (run in the build container, where /root/reference exists).

The reference (kapre on TensorFlow + librosa) cannot be imported here (neither dependency is
installed, no network), and its repository holds no golden outputs -- its tests compare with
librosa at test time.  So the fixtures are produced by the float64 oracle on the reference's own
test input (first 8000 samples of tests/speech_test_file.npz, tests/utils.py:10-15) for the
configurations of tests/test_time_frequency.py:72-267 and tests/test_backend.py:15-40, and each
one is cross-checked at generation time against an independent implementation available in the
container (torch.stft, torchaudio mel filterbank, closed-form dB).  The measured agreement is
stored alongside the vectors.
"""
import json
import os
import sys

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
import oracle as O  # noqa: E402

REF_FIXTURE = '/root/reference/tests/speech_test_file.npz'


def main():
    src = np.load(REF_FIXTURE)['audio_data'].astype(np.float32)[:8000]
    out = {'audio': src}
    checks = {}
    x64 = src.astype(np.float64)

    def torch_stft(n_fft, hop, win):
        w = torch.from_numpy(np.asarray(win, dtype=np.float64))
        return torch.stft(torch.from_numpy(x64), n_fft, hop, len(win), window=w, center=False,
                          return_complex=True).numpy().T

    # tests/test_time_frequency.py:72-125 (n_fft=1000, hop None->250 / 256) and :128-185 (512/256, windows)
    for n_fft, hop, wname in ((1000, 250, None), (1000, 256, None), (512, 256, 'hann_window'),
                              (512, 256, 'hamming_window'), (1024, 256, None), (2048, 512, None)):
        win = O.get_window(wname, n_fft)
        s = O.stft_frames(x64, n_fft, n_fft, hop, win, False)
        key = 'stft_%d_%d_%s' % (n_fft, hop, wname or 'default')
        out[key] = s.astype(np.complex64)  # fixture size; 6e-8 relative rounding
        checks[key + '_vs_torch_stft'] = float(np.abs(s - torch_stft(n_fft, hop, win)).max())
    # win_length < n_fft and pad_end (tests/test_time_frequency.py:270-357 pins this against the
    # in-repo matmul DFT, kapre/tflite_compatible_stft.py)
    for pad_end in (False, True):
        win = O.get_window(None, 512)
        s = O.stft_frames(x64, 1000, 512, 250, win, pad_end)
        d = O.stft_by_dft_matrix(x64, 1000, 512, 250, win, pad_end)
        key = 'stft_1000_250_win512_padend%d' % pad_end
        out[key] = s.astype(np.complex64)  # fixture size; 6e-8 relative rounding
        checks[key + '_vs_dft_matrix'] = float(np.abs(s - d).max())
    # tests/test_time_frequency.py:188-267: mel (n_fft 512, sr 22050, 40 mels, fmax 8000), hop None->128 / 256
    fb = O.filterbank_mel(22050, 257, 40, 0.0, 8000, False, 'slaney')
    ta = torchaudio.functional.melscale_fbanks(257, 0.0, 8000.0, 40, 22050, norm='slaney', mel_scale='slaney').numpy()
    out['melfb_22050_257_40_8000'] = fb
    checks['melfb_vs_torchaudio'] = float(np.abs(fb - ta).max())
    for hop in (128, 256):
        mel = O.melspectrogram_layer(x64[None, :, None], n_fft=512, hop_length=hop, sample_rate=22050, n_mels=40,
                                     mel_f_min=0.0, mel_f_max=8000, input_data_format='channels_last',
                                     output_data_format='channels_last')[0, :, :, 0]
        out['mel_512_%d' % hop] = mel
        t = np.abs(torch_stft(512, hop, O.get_window(None, 512))) @ ta.astype(np.float64)
        checks['mel_512_%d_vs_torch' % hop] = float(np.abs(mel - t).max())
        for amin in (1e-5, 1e-3):
            for dr in (120.0, 80.0):
                db = O.magnitude_to_decibel(mel[None], 1.0, amin, dr)[0]
                out['meldb_512_%d_amin%g_dr%g' % (hop, amin, dr)] = db
    # tests/test_backend.py:20-22 literal matrix; closed form 10*log10 with a per-row maximum
    lit = np.array([[1e-20, 1e-5, 1e-3, 5e-2], [0.3, 1.0, 20.5, 9999]], dtype=np.float64)
    out['db_literal_in'] = lit
    for dr in (80.0, 120.0):
        got = O.magnitude_to_decibel(lit, 1.0, 1e-5, dr)
        closed = 10.0 * np.log10(np.maximum(lit, 1e-5))
        closed = np.maximum(closed, closed.max(axis=1, keepdims=True) - dr)
        out['db_literal_dr%g' % dr] = got
        checks['db_literal_dr%g_vs_closed_form' % dr] = float(np.abs(got - closed).max())
    # filterbank grid of tests/test_backend.py:43-75 vs torchaudio (slaney-normalised subset)
    worst = 0.0
    for sr in (44100, 22050):
        for nf in (1025, 257):
            for nm in (32, 128):
                for fmin in (0.0, 200.0):
                    for ratio in (1.0, 0.5):
                        for htk in (True, False):
                            fmax = int(ratio * (sr // 2))
                            a = O.filterbank_mel(sr, nf, nm, fmin, fmax, htk, 'slaney')
                            b = torchaudio.functional.melscale_fbanks(nf, fmin, float(fmax), nm, sr, norm='slaney',
                                                                       mel_scale='htk' if htk else 'slaney').numpy()
                            worst = max(worst, float(np.abs(a - b).max() / np.abs(b).max()))
    checks['melfb_grid_vs_torchaudio_rel'] = worst
    np.savez_compressed(os.path.join(HERE, 'kapre_cases.npz'), **out)
    with open(os.path.join(HERE, 'kapre_cases_checks.json'), 'w') as f:
        json.dump(checks, f, indent=1, sort_keys=True)
    for k, v in sorted(checks.items()):
        print('%-45s %.3g' % (k, v))
        assert v < 1e-5, (k, v)


if __name__ == '__main__':
    main()
