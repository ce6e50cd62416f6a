"""Randomised differential tests of the CUDA kernel bodies (through the CPU phase emulator) against the
float64 oracle: random hop / length / padding / tile shapes / data formats, so that tile boundaries, the
16-byte-rounded staging, partial rounds and odd strides are hit in combinations the hand-written cases miss.
CPU tier; the example budget keeps it to a few seconds."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import emu_harness as E
import oracle as O

FMT = st.sampled_from(['channels_first', 'channels_last'])
# KAPRE_FUZZ_EXAMPLES=N switches to N fresh random examples per test (bug hunting); the default is a small
# deterministic set so that the CPU tier stays fast and reproducible
_N = int(os.environ.get('KAPRE_FUZZ_EXAMPLES', '0'))
COMMON = dict(deadline=None, max_examples=_N or 60, suppress_health_check=[HealthCheck.too_slow, HealthCheck.filter_too_much],
              derandomize=not _N)


def _wave(seed, B, C, L, fmt):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, size=(B, C, L) if fmt == 'channels_first' else (B, L, C)).astype(np.float32)


def _nerr(a, b):
    return float(np.abs(a - b).max() / max(float(np.abs(b).max()), 1e-30))


@settings(**COMMON)
@given(n_fft=st.sampled_from([256, 512]), hop=st.integers(17, 300), L=st.integers(600, 2600), C=st.integers(1, 3),
       win_frac=st.sampled_from([1.0, 0.8, 0.37]), pad_begin=st.booleans(), pad_end=st.booleans(), fmt=FMT,
       tf_log=st.integers(0, 3), nw=st.sampled_from([1, 2, 4]), bulk=st.booleans(), seed=st.integers(0, 10 ** 6))
def test_fuzz_single_channel_tiles(n_fft, hop, L, C, win_frac, pad_begin, pad_end, fmt, tf_log, nw, bulk, seed):
    if pad_begin and hop > n_fft:
        with pytest.raises(ValueError):          # the reference's tf.pad rejects the negative padding
            O.stft_layer(np.zeros((1, 1, L), np.float32), n_fft, n_fft, hop, None, True, pad_end, fmt, fmt)
        return
    win = max(2, int(n_fft * win_frac))
    fpw = 32 // (n_fft // 64)
    TF = nw * fpw * (1 << tf_log)
    if TF > 64:
        TF = nw * fpw
    x = _wave(seed, 2, C, L, fmt)
    w = O.get_window('hamming_window' if seed % 2 else None, win).astype(np.float32)
    ref = O.stft_layer(x, n_fft, win, hop, 'hamming_window' if seed % 2 else None, pad_begin, pad_end, fmt, fmt)
    if ref.size == 0:
        return
    out, _ = E.emu_stft(x, n_fft, win, hop, w, pad_begin, pad_end, E.MODE_COMPLEX, fmt, fmt, TF=TF, n_warps=nw,
                        n_cta=1 + seed % 3, bulk=int(bulk))
    assert out.shape == ref.shape
    assert not np.isnan(out.view(np.float32)).any()
    assert _nerr(out, ref) < 2e-6


@settings(**COMMON)
@given(n_fft=st.sampled_from([256, 512]), hop=st.integers(16, 260), L=st.integers(700, 2400), C=st.integers(2, 5),
       pad_begin=st.booleans(), pad_end=st.booleans(), ifmt=FMT, ofmt=FMT, TF=st.integers(1, 6),
       nw=st.sampled_from([1, 2, 3, 4]), mode=st.sampled_from([E.MODE_COMPLEX, E.MODE_MAG]), seed=st.integers(0, 10 ** 6))
def test_fuzz_all_channel_tiles(n_fft, hop, L, C, pad_begin, pad_end, ifmt, ofmt, TF, nw, mode, seed):
    if nw * 32 < C or (pad_begin and hop > n_fft):
        return
    x = _wave(seed, 2, C, L, ifmt)
    w = O.get_window(None, n_fft).astype(np.float32)
    ref = O.stft_layer(x, n_fft, n_fft, hop, None, pad_begin, pad_end, ifmt, ofmt)
    if ref.size == 0:
        return
    out, _ = E.emu_stft_mc(x, n_fft, n_fft, hop, w, pad_begin, pad_end, mode, ifmt, ofmt, TF=TF, n_warps=nw,
                           n_cta=1 + seed % 4)
    want = ref if mode == E.MODE_COMPLEX else np.abs(ref)
    assert out.shape == want.shape
    assert not np.isnan(out.view(np.float32)).any()
    assert _nerr(out, want) < 2e-6


@settings(**COMMON)
@given(n_fft=st.sampled_from([256, 512, 1024]), hop_div=st.sampled_from([2, 4, 8]), frames=st.integers(1, 40), C=st.integers(1, 3),
       ifmt=FMT, ofmt=FMT, nw=st.sampled_from([1, 2, 4]), seed=st.integers(0, 10 ** 6))
def test_fuzz_inverse(n_fft, hop_div, frames, C, ifmt, ofmt, nw, seed):
    hop = n_fft // hop_div
    rng = np.random.default_rng(seed)
    F = n_fft // 2 + 1
    shape = (2, frames, F, C) if ifmt == 'channels_last' else (2, C, frames, F)
    X = (rng.normal(size=shape) + 1j * rng.normal(size=shape)).astype(np.complex64)
    dual = O.inverse_stft_window(n_fft, hop, O.get_window(None, n_fft)).astype(np.float32)
    ref = O.istft_layer(X, n_fft, n_fft, hop, None, ifmt, ofmt)
    fpw = 32 // (n_fft // 64)
    R = -(-n_fft // hop)
    y = E.emu_istft(X, n_fft, n_fft, hop, dual, ifmt, ofmt, TFc=R * nw * fpw, n_warps=nw, n_cta=1 + seed % 3)
    assert y.shape == ref.shape
    assert _nerr(y, ref) < 3e-6
    y2 = E.emu_istft(X, n_fft, n_fft, hop, dual, ifmt, ofmt, seg=1 + seed % 50, n_warps=nw, n_cta=1 + seed % 3)   # streaming body
    assert not np.isnan(y2).any()
    assert _nerr(y2, ref) < 3e-6


@settings(**COMMON)
@given(n_fft=st.sampled_from([256, 512]), hop=st.integers(32, 260), L=st.integers(800, 2600), C=st.integers(1, 4),
       n_mels=st.integers(3, 48), htk=st.booleans(), pad_end=st.booleans(), ifmt=FMT, ofmt=FMT,
       nw=st.sampled_from([1, 2, 4, 8]), all_channels=st.booleans(), seed=st.integers(0, 10 ** 6))
def test_fuzz_filterbank_modes(n_fft, hop, L, C, n_mels, htk, pad_end, ifmt, ofmt, nw, all_channels, seed):
    """Mel + dB epilogue on both tile kinds: chunk lists for arbitrary band counts, partial rounds, copy-out."""
    fpw = 32 // (n_fft // 64)
    if nw * fpw > 32:
        nw = 32 // fpw
    x = _wave(seed, 2, C, L, ifmt)
    x[1] *= 1e-2
    w = O.get_window(None, n_fft).astype(np.float32)
    fb = O.filterbank_mel(16000, n_fft // 2 + 1, n_mels, 0.0, None, htk, 'slaney')
    ref = O.melspectrogram_layer(x, n_fft=n_fft, hop_length=hop, sample_rate=16000, n_mels=n_mels, mel_htk=htk,
                                 pad_end=pad_end, return_decibel=False, input_data_format=ifmt, output_data_format=ofmt)
    if ref.size == 0:
        return
    if all_channels and C > 1 and nw * fpw >= C:
        out, item_max = E.emu_stft_mc(x, n_fft, n_fft, hop, w, False, pad_end, E.MODE_FB_DB, ifmt, ofmt, fb=fb,
                                      TF=(nw * fpw) // C, n_warps=nw, n_cta=1 + seed % 3)
    else:
        out, item_max = E.emu_stft(x, n_fft, n_fft, hop, w, False, pad_end, E.MODE_FB_DB, ifmt, ofmt, fb=fb,
                                   TF=nw * fpw, n_warps=nw, n_cta=1 + seed % 3, bulk=seed % 2, fb_mma=(0, 1, 2, 4)[(seed // 2) % 4])   # flat chunk lists, tensor-core GEMM, band descriptors, paired-column pair step
    assert out.shape == ref.shape
    ref_db = 10.0 * np.log10(np.maximum(ref, 1e-5))
    assert np.abs(out - ref_db).max() < 2e-3
    assert np.allclose(item_max, np.maximum(ref, 1e-5).reshape(2, -1).max(axis=1), rtol=5e-6)


@settings(**COMMON)
@given(n_fft=st.integers(2, 130), win_frac=st.sampled_from([1.0, 0.6]), hop=st.integers(1, 90), L=st.integers(1, 700),
       C=st.integers(1, 3), pad_begin=st.booleans(), pad_end=st.booleans(), fmt=FMT, seed=st.integers(0, 10 ** 6))
def test_fuzz_generic_n_fft(n_fft, win_frac, hop, L, C, pad_begin, pad_end, fmt, seed):
    """Direct-DFT bodies (any n_fft, even or odd): forward against the oracle, then the inverse of that spectrum."""
    win = max(1, int(n_fft * win_frac))
    if pad_begin and hop > n_fft:
        return
    x = _wave(seed, 2, C, L, fmt)
    w = O.get_window(None, win).astype(np.float32)
    ref = O.stft_layer(x, n_fft, win, hop, None, pad_begin, pad_end, fmt, fmt)
    if ref.size == 0:
        return
    out = E.emu_dft(x, n_fft, win, hop, w, pad_begin, pad_end, E.MODE_COMPLEX, fmt, fmt, n_cta=1 + seed % 3)
    assert out.shape == ref.shape
    assert _nerr(out, ref) < 5e-6
    grp = (1, 2, 4)[seed % 3]
    mr = E.emu_mr(x, n_fft, win, hop, w, pad_begin, pad_end, E.MODE_COMPLEX, fmt, fmt, n_warps=grp * (1 + seed % 4),
                  fpw=1 + seed % 3, n_cta=1 + seed % 3, group=grp)   # the Stockham kernel, when n_fft is 5-smooth
    if mr is not None:
        assert mr.shape == ref.shape
        assert _nerr(mr, ref) < 5e-6
    if hop <= win:                  # the dual window needs overlapping (or abutting) frames
        dual = O.inverse_stft_window(win, hop, O.get_window(None, win))
        if np.isfinite(dual).all() and np.abs(dual).max() < 20:      # barely overlapping frames: 1/w blows up round-off
            spec = ref.astype(np.complex64)
            want = O.istft_layer(spec, n_fft, win, hop, None, fmt, fmt)
            y = E.emu_idft(spec, n_fft, win, hop, dual, fmt, fmt, n_cta=1 + seed % 2)
            assert y.shape == want.shape
            # fp32 sums of n_fft terms, times a dual window of up to 20: 2e-5 of the output's maximum, plus the rounding of the
            # sums themselves (~eps sqrt(n_fft) |X|max / n_fft per term, times the dual window, x16) -- which is all that is
            # left of the bound for degenerate inputs (two-sample signals under a window's zero end: outputs of 1e-7 and less)
            wmax, smax = float(np.abs(want).max()), float(np.abs(spec).max())
            tol = 2e-5 * wmax + 1e-6 * smax * max(1.0, float(np.abs(dual).max())) / np.sqrt(n_fft)
            assert float(np.abs(y - want).max()) <= tol
