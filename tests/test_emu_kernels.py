"""CPU tier: the CUDA kernel bodies (the exact source nvcc compiles, run through the phase
emulation of tests/emu) against the float64 oracle.  No GPU needed."""
import numpy as np
import pytest

import emu_harness as E
import oracle as O


def nerr(a, b):
    return float(np.abs(a - b).max() / max(float(np.abs(b).max()), 1e-30))


def wave(rng, B, C, L, fmt):
    return rng.uniform(-1, 1, size=(B, C, L) if fmt == 'channels_first' else (B, L, C)).astype(np.float32)


@pytest.mark.parametrize('n_fft,hop,win', [(1024, 256, 1024), (512, 256, 512), (2048, 1024, 2018), (256, 64, 200),
                                           (1024, 250, 1024), (512, 125, 400)])
@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
@pytest.mark.parametrize('pads', [(False, False), (True, True)])
def test_emu_stft_complex_and_mag(n_fft, hop, win, fmt, pads):
    rng = np.random.default_rng(n_fft + hop)
    x = wave(rng, 2, 3, 5003, fmt)  # odd length: item bases hit every 4 B alignment class
    w = O.get_window(None, win).astype(np.float32)
    ref = O.stft_layer(x, n_fft, win, hop, None, pads[0], pads[1], fmt, fmt)
    out, _ = E.emu_stft(x, n_fft, win, hop, w, pads[0], pads[1], E.MODE_COMPLEX, fmt, fmt, TF=16, n_warps=4)
    assert out.shape == ref.shape
    assert nerr(out, ref) < 1e-6
    mag, _ = E.emu_stft(x, n_fft, win, hop, w, pads[0], pads[1], E.MODE_MAG, fmt, fmt, TF=8, n_warps=2, dbuf=0)
    assert nerr(mag, np.abs(ref)) < 1e-6
    # the natural-order pair step (A/B alternative of the paired-column form that n_fft <= 1024 runs by default)
    out_n, _ = E.emu_stft(x, n_fft, win, hop, w, pads[0], pads[1], E.MODE_COMPLEX, fmt, fmt, TF=16, n_warps=4, fb_mma=4)
    assert nerr(out_n, ref) < 1e-6
    assert nerr(out_n, out) < 3e-7


@pytest.mark.parametrize('dbuf,bulk', [(0, 1), (0, 0)])
@pytest.mark.parametrize('L', [4000, 4001, 4002, 4003])
def test_emu_staging_paths(dbuf, bulk, L):
    """TMA-style 16 B-rounded staging (emulated as memcpy) for every alignment of the item base,
    against the cooperative loader; pad regions next to the valid range must stay zero."""
    rng = np.random.default_rng(L)
    x = wave(rng, 3, 2, L, 'channels_first')
    w = O.get_window('hamming_window', 512).astype(np.float32)
    ref = O.stft_layer(x, 512, 512, 128, 'hamming_window', True, True, 'channels_first', 'channels_first')
    out, _ = E.emu_stft(x, 512, 512, 128, w, True, True, E.MODE_COMPLEX, 'channels_first', 'channels_first',
                        TF=8, n_warps=2, n_cta=2, dbuf=dbuf, bulk=bulk)
    assert nerr(out, ref) < 1e-6


@pytest.mark.parametrize('n_fft,hop,n_mels,sr,TF,nw', [(1024, 256, 128, 22050, 16, 8), (1024, 256, 128, 22050, 8, 4),
                                                        (512, 256, 40, 22050, 16, 4), (2048, 512, 64, 44100, 8, 8),
                                                        (2048, 512, 128, 44100, 4, 4), (256, 64, 20, 16000, 32, 4),
                                                        (512, 128, 64, 16000, 8, 2), (1024, 256, 80, 16000, 4, 2),
                                                        (512, 256, 64, 16000, 4, 1), (1024, 256, 128, 22050, 2, 1)])   # 1-warp CTAs: small batches
@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
def test_emu_mel_and_db(n_fft, hop, n_mels, sr, TF, nw, fmt):
    rng = np.random.default_rng(n_mels)
    B = 2
    x = wave(rng, B, 2, 6001, fmt)
    x[1] *= 1e-3
    w = O.get_window(None, n_fft).astype(np.float32)
    fb = O.filterbank_mel(sr, n_fft // 2 + 1, n_mels, 0.0, None, False, 'slaney')
    kw = dict(n_fft=n_fft, hop_length=hop, sample_rate=sr, n_mels=n_mels, input_data_format=fmt, output_data_format=fmt)
    ref = O.melspectrogram_layer(x, **kw)
    out, _ = E.emu_stft(x, n_fft, n_fft, hop, w, False, False, E.MODE_FB, fmt, fmt, fb=fb, TF=TF, n_warps=nw)
    assert nerr(out, ref) < 1e-6
    # two-level walk over band descriptors: the same sums in the same order -> bit-identical
    out2, _ = E.emu_stft(x, n_fft, n_fft, hop, w, False, False, E.MODE_FB, fmt, fmt, fb=fb, TF=TF, n_warps=nw, fb_mma=2)
    assert np.array_equal(out, out2)
    # natural-order pair step (the A/B alternative of the default paired-column form): same magnitudes up to the rounding
    # of the twiddles of bins >= P/2
    out3, _ = E.emu_stft(x, n_fft, n_fft, hop, w, False, False, E.MODE_FB, fmt, fmt, fb=fb, TF=TF, n_warps=nw, fb_mma=4)
    assert nerr(out3, ref) < 1e-6
    assert nerr(out3, out) < 3e-7
    refdb = O.melspectrogram_layer(x, return_decibel=True, db_dynamic_range=1e9, **kw)
    outdb, imax = E.emu_stft(x, n_fft, n_fft, hop, w, False, False, E.MODE_FB_DB, fmt, fmt, fb=fb, TF=TF, n_warps=nw)
    assert np.abs(outdb - refdb).max() < 5e-5
    np.testing.assert_allclose(imax, np.maximum(ref.reshape(B, -1).max(1), 1e-5), rtol=1e-6)
    sm = O.stft_magnitude_layer(x, n_fft, None, hop, return_decibel=True, db_dynamic_range=1e9,
                                input_data_format=fmt, output_data_format=fmt)
    smo, imax2 = E.emu_stft(x, n_fft, n_fft, hop, w, False, False, E.MODE_MAG_DB, fmt, fmt, TF=2 * TF, n_warps=nw)
    lin = O.stft_magnitude_layer(x, n_fft, None, hop, input_data_format=fmt, output_data_format=fmt)
    sig = lin > 1e-4 * lin.reshape(B, -1).max(1).reshape(B, 1, 1, 1)
    assert np.abs(smo - sm)[sig].max() < 1e-3
    np.testing.assert_allclose(imax2, lin.reshape(B, -1).max(1), rtol=1e-6)


@pytest.mark.parametrize('n_fft,hop,n_mels,sr,TF,nw', [(1024, 256, 128, 22050, 16, 8), (1024, 256, 128, 22050, 8, 4),
                                                        (512, 256, 40, 22050, 16, 4), (512, 128, 64, 16000, 32, 8),
                                                        (2048, 512, 128, 44100, 8, 8), (2048, 512, 64, 44100, 16, 16),
                                                        (256, 64, 20, 16000, 32, 4), (1024, 256, 77, 16000, 4, 2)])
@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
def test_emu_mel_tensor_core_filterbank(n_fft, hop, n_mels, sr, TF, nw, fmt):
    """Filterbank phase as the block-banded mma.sync 3xTF32 GEMM (kb_fb_mma_phase): same tolerances as the
    CUDA-core phase, incl. a band count that is no multiple of 8 and a half-filled 16-row tile (n_fft 2048, 8 warps)."""
    rng = np.random.default_rng(n_mels + 1)
    B = 2
    x = wave(rng, B, 2, 6001, fmt)
    x[1] *= 1e-3
    w = O.get_window(None, n_fft).astype(np.float32)
    fb = O.filterbank_mel(sr, n_fft // 2 + 1, n_mels, 0.0, None, False, 'slaney')
    kw = dict(n_fft=n_fft, hop_length=hop, sample_rate=sr, n_mels=n_mels, input_data_format=fmt, output_data_format=fmt)
    ref = O.melspectrogram_layer(x, **kw)
    out, _ = E.emu_stft(x, n_fft, n_fft, hop, w, False, False, E.MODE_FB, fmt, fmt, fb=fb, TF=TF, n_warps=nw, fb_mma=1)
    assert nerr(out, ref) < 1e-6
    refdb = O.melspectrogram_layer(x, return_decibel=True, db_dynamic_range=1e9, **kw)
    outdb, imax = E.emu_stft(x, n_fft, n_fft, hop, w, True, True, E.MODE_FB_DB, fmt, fmt, fb=fb, TF=TF, n_warps=nw,
                             fb_mma=1)
    refdb2 = O.melspectrogram_layer(x, return_decibel=True, db_dynamic_range=1e9, pad_begin=True, pad_end=True, **kw)
    assert np.abs(outdb - refdb2).max() < 5e-5


def test_emu_log_filterbank_tensor_core():
    rng = np.random.default_rng(10)
    x = wave(rng, 1, 1, 4000, 'channels_first')
    fb = O.filterbank_log(22050, 513, 84, 12)
    ref = O.apply_filterbank(np.abs(O.stft_layer(x, 1024, None, 256, input_data_format='channels_first',
                                                 output_data_format='channels_first')), fb.astype(np.float64), 'channels_first')
    out, _ = E.emu_stft(x, 1024, 1024, 256, O.get_window(None, 1024), False, False, E.MODE_FB, 'channels_first',
                        'channels_first', fb=fb, TF=16, n_warps=8, fb_mma=1)
    assert nerr(out, ref) < 1e-6


def test_emu_log_filterbank_dense_bands():
    rng = np.random.default_rng(9)
    x = wave(rng, 1, 1, 4000, 'channels_first')
    fb = O.filterbank_log(22050, 513, 84, 12)
    ref = O.apply_filterbank(np.abs(O.stft_layer(x, 1024, None, 256, input_data_format='channels_first',
                                                 output_data_format='channels_first')), fb.astype(np.float64), 'channels_first')
    out, _ = E.emu_stft(x, 1024, 1024, 256, O.get_window(None, 1024), False, False, E.MODE_FB, 'channels_first',
                        'channels_first', fb=fb, TF=8, n_warps=4)
    assert nerr(out, ref) < 1e-6


@pytest.mark.parametrize('n_fft,hop,win,TFc,nw', [(1024, 256, 1024, 16, 4), (2048, 1024, 2048, 8, 4), (2048, 256, 2048, 24, 4),
                                                   (512, 128, 400, 12, 2), (256, 100, 256, 9, 2)])
@pytest.mark.parametrize('ifmt', ['channels_first', 'channels_last'])
@pytest.mark.parametrize('ofmt', ['channels_first', 'channels_last'])
def test_emu_istft(n_fft, hop, win, TFc, nw, ifmt, ofmt):
    rng = np.random.default_rng(n_fft + hop)
    B, C, T, F = 2, 2, 37, n_fft // 2 + 1
    shp = (B, C, T, F) if ifmt == 'channels_first' else (B, T, F, C)
    X = (rng.normal(size=shp) + 1j * rng.normal(size=shp)).astype(np.complex64)
    dual = O.inverse_stft_window(win, hop, O.get_window(None, win))
    ref = O.istft_layer(X, n_fft, win, hop, None, ifmt, ofmt)
    y = E.emu_istft(X, n_fft, win, hop, dual, ifmt, ofmt, TFc=TFc, n_warps=nw, n_cta=2)
    assert y.shape == ref.shape
    assert nerr(y, ref) < 1e-6


@pytest.mark.parametrize('n_fft,hop,win,seg,nw', [(1024, 256, 1024, 5, 4), (1024, 256, 1024, 13, 2), (2048, 1024, 2048, 3, 4),
                                                   (2048, 256, 2048, 9, 4), (512, 128, 400, 29, 2), (256, 100, 256, 7, 2),
                                                   (256, 300, 256, 4, 1), (512, 37, 512, 100, 1), (1024, 256, 1024, 1, 4),
                                                   (512, 128, 512, 41, 4), (256, 64, 250, 6, 3)])
@pytest.mark.parametrize('ifmt', ['channels_first', 'channels_last'])
@pytest.mark.parametrize('ofmt', ['channels_first', 'channels_last'])
def test_emu_istft_streaming(n_fft, hop, win, seg, nw, ifmt, ofmt):
    """istft_core.cuh kb_istft2_cta: frames left in the exchange regions, gather-sum with a carry between rounds; hop > win
    (gaps), odd hops (scalar gather), a window shorter than n_fft, tiles of one hop and of the whole signal."""
    rng = np.random.default_rng(n_fft + hop + seg)
    B, C, T, F = 2, 2, 37, n_fft // 2 + 1
    shp = (B, C, T, F) if ifmt == 'channels_first' else (B, T, F, C)
    X = (rng.normal(size=shp) + 1j * rng.normal(size=shp)).astype(np.complex64)
    if hop <= win:
        dual = O.inverse_stft_window(win, hop, O.get_window(None, win))
        ref = O.istft_layer(X, n_fft, win, hop, None, ifmt, ofmt)
    else:                      # no overlap: any window will do, compare with the class-ordered body
        dual = O.get_window(None, win)
        ref = E.emu_istft(X, n_fft, win, hop, dual, ifmt, ofmt, TFc=4, n_warps=1, n_cta=1)
    y = E.emu_istft(X, n_fft, win, hop, dual, ifmt, ofmt, seg=seg, n_warps=nw, n_cta=2)
    assert y.shape == ref.shape
    assert not np.isnan(y).any()
    assert nerr(y, ref) < 1e-6


# ------------------------------------------------------------------------------- multi-channel tiles
@pytest.mark.parametrize('n_fft,hop,win,C,TF,nw', [(2048, 1024, 2048, 6, 1, 6), (2048, 1024, 2048, 6, 2, 4),
                                                  (1024, 256, 1024, 2, 8, 8), (512, 128, 400, 3, 3, 2),
                                                  (256, 64, 256, 5, 4, 3), (1024, 255, 1024, 2, 2, 1),
                                                  (512, 256, 512, 4, 5, 4), (256, 64, 256, 12, 2, 4),
                                                  (512, 128, 512, 31, 1, 2)])
@pytest.mark.parametrize('ifmt', ['channels_last', 'channels_first'])
@pytest.mark.parametrize('ofmt', ['channels_last', 'channels_first'])
def test_emu_mc_complex_and_mag(n_fft, hop, win, C, TF, nw, ifmt, ofmt):
    """stft_mc_core.cuh: all-channel tiles, de-interleaving loader, cooperative (interleaved output)
    and per-warp (planar output) pair step; odd hop exercises the scalar window path."""
    rng = np.random.default_rng(n_fft + hop + C)
    x = wave(rng, 2, C, 7003, ifmt)
    w = O.get_window(None, win).astype(np.float32)
    for pads in ((False, False), (True, True)):
        ref = O.stft_layer(x, n_fft, win, hop, None, pads[0], pads[1], ifmt, ofmt)
        out, _ = E.emu_stft_mc(x, n_fft, win, hop, w, pads[0], pads[1], E.MODE_COMPLEX, ifmt, ofmt, TF=TF, n_warps=nw)
        assert out.shape == ref.shape
        assert not np.isnan(out).any()
        assert nerr(out, ref) < 1e-6
    mag, _ = E.emu_stft_mc(x, n_fft, win, hop, w, True, False, E.MODE_MAG, ifmt, ofmt, TF=TF, n_warps=nw, n_cta=5)
    ref = O.stft_layer(x, n_fft, win, hop, None, True, False, ifmt, ofmt)
    assert nerr(mag, np.abs(ref)) < 1e-6


@pytest.mark.parametrize('fmt', ['channels_last', 'channels_first'])
def test_emu_mc_db_and_phase(fmt):
    rng = np.random.default_rng(5)
    B, C = 3, 6
    x = wave(rng, B, C, 9000, fmt)
    x[1] *= 1e-3
    w = O.get_window('hamming_window', 2048).astype(np.float32)
    spec = O.stft_layer(x, 2048, 2048, 1024, 'hamming_window', False, True, fmt, fmt)
    mag = np.abs(spec)
    out, item_max = E.emu_stft_mc(x, 2048, 2048, 1024, w, False, True, E.MODE_MAG_DB, fmt, fmt, TF=1, n_warps=6)
    ref_db = 10.0 * np.log10(np.maximum(mag, 1e-5))       # un-clamped: the clamp is a second kernel
    assert np.abs(out - ref_db).max() < 1e-3     # fp32 FFT round-off on the weakest bins of the 1e-3-scaled item
    want_max = np.maximum(mag, 1e-5).reshape(B, -1).max(axis=1)
    assert np.allclose(item_max, want_max, rtol=2e-6)
    both, item_max = E.emu_stft_mc(x, 2048, 2048, 1024, w, False, True, E.MODE_MAG_PHASE, fmt, fmt, TF=2, n_warps=4,
                                   db_on=1)
    ch_axis = 3 if fmt == 'channels_last' else 1
    m, ph = np.split(both, 2, axis=ch_axis)
    assert np.abs(m - ref_db).max() < 1e-3
    ok = mag > 1e-3 * mag.max()
    d = np.abs(np.angle(np.exp(1j * (ph - np.angle(spec)))))
    assert d[ok].max() < 1e-4
    assert np.allclose(item_max, want_max, rtol=2e-6)


@pytest.mark.parametrize('n_fft,hop,n_mels,sr,C,TF,nw', [(1024, 256, 128, 22050, 2, 8, 8), (1024, 256, 80, 16000, 3, 5, 8),
                                                          (512, 128, 40, 16000, 4, 4, 4), (2048, 512, 96, 44100, 2, 2, 4),
                                                          (2048, 1024, 64, 44100, 6, 1, 8), (256, 64, 20, 8000, 5, 6, 4)])
@pytest.mark.parametrize('ifmt', ['channels_last', 'channels_first'])
@pytest.mark.parametrize('ofmt', ['channels_last', 'channels_first'])
def test_emu_mc_mel_and_db(n_fft, hop, n_mels, sr, C, TF, nw, ifmt, ofmt):
    """Filterbank modes on all-channel tiles (kb_stft_mcfb_cta): partial rounds (TF*C < columns), interleaved
    and planar copy-out, per-item maximum."""
    rng = np.random.default_rng(n_mels + C)
    B = 2
    x = wave(rng, B, C, 6001, ifmt)
    x[1] *= 1e-3
    w = O.get_window(None, n_fft).astype(np.float32)
    fb = O.filterbank_mel(sr, n_fft // 2 + 1, n_mels, 0.0, None, False, 'slaney')
    kw = dict(n_fft=n_fft, hop_length=hop, sample_rate=sr, n_mels=n_mels, pad_end=True, input_data_format=ifmt,
              output_data_format=ofmt)
    ref = O.melspectrogram_layer(x, return_decibel=False, **kw)
    out, _ = E.emu_stft_mc(x, n_fft, n_fft, hop, w, False, True, E.MODE_FB, ifmt, ofmt, fb=fb, TF=TF, n_warps=nw)
    assert out.shape == ref.shape
    assert not np.isnan(out).any()
    assert nerr(out, ref) < 2e-6
    out, item_max = E.emu_stft_mc(x, n_fft, n_fft, hop, w, False, True, E.MODE_FB_DB, ifmt, ofmt, fb=fb, TF=TF,
                                  n_warps=nw, n_cta=2)
    ref_db = 10.0 * np.log10(np.maximum(ref, 1e-5))
    assert np.abs(out - ref_db).max() < 1e-3
    assert np.allclose(item_max, np.maximum(ref, 1e-5).reshape(B, -1).max(axis=1), rtol=3e-6)


def test_emu_sixteen_warp_cta_n2048_mel():
    """n_fft = 2048 filterbank modes run 16-warp CTAs (kb_stft_kernel_w16): TF = 16 frames per tile."""
    rng = np.random.default_rng(16)
    x = wave(rng, 2, 1, 30000, 'channels_last')
    x[1] *= 1e-2
    w = O.get_window(None, 2048).astype(np.float32)
    fb = O.filterbank_mel(44100, 1025, 128, 0.0, None, False, 'slaney')
    kw = dict(n_fft=2048, hop_length=512, sample_rate=44100, n_mels=128, input_data_format='channels_last',
              output_data_format='channels_last')
    ref = O.melspectrogram_layer(x, return_decibel=False, **kw)
    out, _ = E.emu_stft(x, 2048, 2048, 512, w, False, False, E.MODE_FB, 'channels_last', 'channels_last', fb=fb,
                        TF=16, n_warps=16, n_cta=2)
    assert nerr(out, ref) < 2e-6
    out, item_max = E.emu_stft(x, 2048, 2048, 512, w, False, False, E.MODE_FB_DB, 'channels_last', 'channels_last',
                               fb=fb, TF=16, n_warps=16, n_cta=3)
    assert np.abs(out - 10.0 * np.log10(np.maximum(ref, 1e-5))).max() < 1e-3
    assert np.allclose(item_max, np.maximum(ref, 1e-5).reshape(2, -1).max(axis=1), rtol=3e-6)


# ------------------------------------------------------------------------------- mixed-radix Stockham body
@pytest.mark.parametrize('n_fft,win,hop', [(1000, 1000, 250), (1000, 512, 250), (400, 400, 160), (4096, 4096, 1024),
                                           (100, 64, 33), (6, 6, 2), (2, 2, 1), (75, 75, 25), (135, 100, 40), (480, 480, 120),
                                           (8192, 8192, 4096), (128, 128, 64), (3, 3, 1), (1, 1, 1)])
@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
def test_emu_mixed_radix_forward(n_fft, win, hop, fmt):
    """mr_core.cuh: Stockham FFT with radices 2/3/4/5 for n_fft outside {256..2048}; packed real FFT for even
    n_fft, complex FFT for odd (the reference's tests use n_fft = 1000, tests/test_time_frequency.py:72-125)."""
    rng = np.random.default_rng(n_fft + hop)
    x = wave(rng, 2, 2, max(3000, 2 * n_fft + 900) if n_fft > 50 else 200, fmt)
    w = O.get_window(None, win).astype(np.float32)
    for pads in ((False, False), (True, True)):
        ref = O.stft_layer(x, n_fft, win, hop, None, pads[0], pads[1], fmt, fmt)
        nw = 2 if n_fft >= 4096 else 4
        out = E.emu_mr(x, n_fft, win, hop, w, pads[0], pads[1], E.MODE_COMPLEX, fmt, fmt, n_warps=nw, fpw=2)
        assert out is not None and out.shape == ref.shape
        assert nerr(out, ref) < 3e-6
        mag = E.emu_mr(x, n_fft, win, hop, w, pads[0], pads[1], E.MODE_MAG, fmt, fmt, n_warps=nw, fpw=1, n_cta=2)
        assert nerr(mag, np.abs(ref)) < 3e-6
        # several warps per frame (the launch shape of the large sizes): bit-identical to one warp per frame, the
        # butterflies are the same, only their assignment to threads differs
        for nw_g, g in ((4, 2), (8, 8), (16, 16), (8, 4)):
            grp = E.emu_mr(x, n_fft, win, hop, w, pads[0], pads[1], E.MODE_COMPLEX, fmt, fmt, n_warps=nw_g, fpw=1, n_cta=2,
                           group=g)
            assert np.array_equal(grp.view(np.float32), out.view(np.float32))


@pytest.mark.parametrize('n_fft,hop,n_mels,sr,nw,group,frt', [(400, 160, 80, 16000, 8, 1, 32), (400, 160, 80, 16000, 4, 2, 8),
                                                                (1000, 250, 128, 22050, 8, 4, 16), (4096, 1024, 128, 44100, 8, 8, 8),
                                                                (480, 120, 33, 16000, 2, 1, 16), (75, 25, 10, 8000, 4, 1, 32),
                                                                (1000, 250, 40, 22050, 16, 2, 32)])
@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
def test_emu_mixed_radix_mel_and_db(n_fft, hop, n_mels, sr, nw, group, frt, fmt):
    """mr_core.cuh fused tail: magnitudes into a (bin x frame) tile, banded filterbank, decibel + per-item maximum --
    log-mel front ends whose n_fft is not 64 * 2^k (speech: 400 / 160 / 80) in one launch, like the register kernel."""
    rng = np.random.default_rng(n_mels + n_fft)
    B = 2
    x = wave(rng, B, 2, max(3 * n_fft + 700, 2500), fmt)
    x[1] *= 1e-3
    w = O.get_window(None, n_fft).astype(np.float32)
    fb = O.filterbank_mel(sr, n_fft // 2 + 1, n_mels, 0.0, None, False, 'slaney')
    kw = dict(n_fft=n_fft, hop_length=hop, sample_rate=sr, n_mels=n_mels, input_data_format=fmt, output_data_format=fmt)
    for pads in ((False, False), (True, True)):
        ref = O.melspectrogram_layer(x, pad_begin=pads[0], pad_end=pads[1], **kw)
        out = E.emu_mr(x, n_fft, n_fft, hop, w, pads[0], pads[1], E.MODE_FB, fmt, fmt, n_warps=nw, group=group, fb=fb, frt=frt)
        assert out.shape == ref.shape and not np.isnan(out).any()
        assert nerr(out, ref) < 2e-6
        refdb = O.melspectrogram_layer(x, return_decibel=True, db_dynamic_range=1e9, pad_begin=pads[0], pad_end=pads[1], **kw)
        outdb, imax = E.emu_mr(x, n_fft, n_fft, hop, w, pads[0], pads[1], E.MODE_FB_DB, fmt, fmt, n_warps=nw, group=group,
                               fb=fb, frt=frt, with_item_max=True)
        assert np.abs(outdb - refdb).max() < 5e-5
        np.testing.assert_allclose(imax, np.maximum(ref.reshape(B, -1).max(1), 1e-5), rtol=2e-6)
    lin = O.stft_magnitude_layer(x, n_fft, None, hop, input_data_format=fmt, output_data_format=fmt)
    sm = O.stft_magnitude_layer(x, n_fft, None, hop, return_decibel=True, db_dynamic_range=1e9, input_data_format=fmt,
                                output_data_format=fmt)
    smo, imax2 = E.emu_mr(x, n_fft, n_fft, hop, w, False, False, E.MODE_MAG_DB, fmt, fmt, n_warps=nw, group=group, fpw=2,
                          with_item_max=True)
    sig = lin > 1e-4 * lin.reshape(B, -1).max(1).reshape(B, 1, 1, 1)
    assert np.abs(smo - sm)[sig].max() < 1e-3
    np.testing.assert_allclose(imax2, lin.reshape(B, -1).max(1), rtol=2e-6)


def test_emu_mixed_radix_rejects_large_prime_factors():
    x = np.zeros((1, 1, 100), dtype=np.float32)
    for n_fft in (14, 22, 98, 7):       # 7 | P or 11 | P
        assert E.emu_mr(x, n_fft, n_fft, 1, np.ones(n_fft, np.float32), False, False, E.MODE_MAG, 'channels_first',
                        'channels_first') is None


# ------------------------------------------------------------------------------- stand-alone / generic-n_fft bodies
@pytest.mark.parametrize('n_fft,win,hop', [(1000, 1000, 250), (1000, 512, 250), (400, 400, 160), (100, 64, 33), (6, 6, 2)])
@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
def test_emu_generic_n_fft_forward_and_inverse(n_fft, win, hop, fmt):
    """aux_core.cuh: direct DFT / inverse DFT for n_fft outside {256..2048} (the reference's tests use 1000)."""
    rng = np.random.default_rng(n_fft + hop)
    x = wave(rng, 2, 2, 3000 if n_fft > 50 else 200, fmt)
    w = O.get_window(None, win).astype(np.float32)
    for pads in ((False, False), (True, True)):
        ref = O.stft_layer(x, n_fft, win, hop, None, pads[0], pads[1], fmt, fmt)
        out = E.emu_dft(x, n_fft, win, hop, w, pads[0], pads[1], E.MODE_COMPLEX, fmt, fmt)
        assert out.shape == ref.shape and nerr(out, ref) < 3e-6
        mag = E.emu_dft(x, n_fft, win, hop, w, pads[0], pads[1], E.MODE_MAG, fmt, fmt, n_cta=2)
        assert nerr(mag, np.abs(ref)) < 3e-6
    spec = O.stft_layer(x, n_fft, win, hop, None, True, True, fmt, fmt).astype(np.complex64)
    dual = O.inverse_stft_window(win, hop, O.get_window(None, win))
    ref = O.istft_layer(spec, n_fft, win, hop, None, fmt, fmt)
    y = E.emu_idft(spec, n_fft, win, hop, dual, fmt, fmt)
    assert y.shape == ref.shape and nerr(y, ref) < 3e-6


@pytest.mark.parametrize('fmt', ['channels_first', 'channels_last'])
@pytest.mark.parametrize('F,M,kind', [(257, 40, 'mel'), (513, 128, 'mel'), (501, 24, 'mel'), (257, 84, 'log'), (129, 7, 'dense')])
def test_emu_standalone_filterbank(fmt, F, M, kind):
    """kb_fb_cta (ApplyFilterbank on its own): banded mel, overlapping log-frequency bands, a dense matrix."""
    rng = np.random.default_rng(F + M)
    if kind == 'mel':
        fb = O.filterbank_mel(16000, F, M, 0.0, None, False, 'slaney')
    elif kind == 'log':
        fb = O.filterbank_log(22050, F, M, 12, None, 0.125)
    else:
        fb = rng.normal(size=(F, M))
    shape = (2, 45, F, 3) if fmt == 'channels_last' else (2, 3, 45, F)
    x = np.abs(rng.normal(size=shape)).astype(np.float32)
    ref = O.apply_filterbank(x, fb, fmt)
    out = E.emu_fb(x, fb, fmt)
    assert out.shape == ref.shape
    assert nerr(out, ref) < 2e-6
    for R in (16, 4, 1):          # shorter tiles: what long spectra (n_fft >= 4096) fall back to
        assert nerr(E.emu_fb(x, fb, fmt, n_cta=2, R=R), ref) < 2e-6


def test_emu_atan2_accuracy_and_special_values():
    """kb_atan2 (octant reduction + degree-7 minimax polynomial) against float64 arctan2: as accurate as a float32 atan2
    can be near pi (1 ulp = 2.4e-7), quadrants and signed zeros as atan2f (tf.math.angle, kapre/time_frequency.py:402)."""
    rng = np.random.default_rng(3)
    n = 1_000_000
    x = (rng.normal(size=n) * 10.0 ** rng.integers(-6, 6, n)).astype(np.float32)
    y = (rng.normal(size=n) * 10.0 ** rng.integers(-6, 6, n)).astype(np.float32)
    y[: n // 4] = x[: n // 4] * (1 + 0.3 * rng.normal(size=n // 4)).astype(np.float32)      # around the octant boundaries
    got = E.emu_atan2(y, x).astype(np.float64)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    assert np.abs(got - ref).max() < 4e-7
    ys = np.array([0.0, 0.0, -0.0, 1.0, -1.0, 0.0, -0.0, 0.0, 3.0, -3.0], np.float32)
    xs = np.array([1.0, -1.0, -1.0, 0.0, 0.0, 0.0, 0.0, -0.0, -0.0, 0.0], np.float32)
    np.testing.assert_allclose(E.emu_atan2(ys, xs), np.arctan2(ys, xs), atol=3e-7)
    assert np.array_equal(np.signbit(E.emu_atan2(ys, xs)), np.signbit(np.arctan2(ys, xs)))
