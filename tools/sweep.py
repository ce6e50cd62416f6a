"""Quick device-time sweep of the fused kernels over launch configurations (dev tool)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kapre_b200 as K
from kapre_b200 import _native


try:
    import pynvml
    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(0)

    def sm_clock():
        return pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
except Exception:  # pragma: no cover
    def sm_clock():
        return -1


def time_it(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    clk = sm_clock()
    torch.cuda.synchronize()
    time_it.clk = clk
    return e0.elapsed_time(e1) / iters


def main():
    torch.cuda.set_device(0)
    print(torch.cuda.get_device_name(0), flush=True)
    B, L = 256, 110250
    xs = [torch.rand((B, L, 1), device='cuda') * 2 - 1 for _ in range(3)]  # 3 x 113 MB > L2
    frames = B * 427
    cnt = [0]
    for name, layer in (
            ('meldb', K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128, return_decibel=True)),
            ('mel', K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128)),
            ('mag', K.get_stft_magnitude_layer(n_fft=1024, hop_length=256)),
            ('stft', K.STFT(n_fft=1024, hop_length=256))):
        fb = name.startswith('mel')
        for nw, tf in [(nw, tf) for nw in (2, 4, 8) for tf in ((0,) if fb else (8, 16, 32))]:
            os.environ['KAPRE_B200_TF'] = str(tf)
            os.environ['KAPRE_B200_NW'] = str(nw)

            def fn():
                cnt[0] += 1
                return layer(xs[cnt[0] % 3])
            try:
                ms = time_it(fn)
                print('%-6s TF=%2d NW=%d  %.3f ms  %.3e frames/s  clk=%d [%s]' % (name, tf, nw, ms, frames / ms * 1e3, time_it.clk,
                                                                          _native.last_launch_info()), flush=True)
            except Exception as e:  # config does not fit
                print('%-6s TF=%2d NW=%d n/a (%s)' % (name, tf, nw, str(e)[:60]), flush=True)
    os.environ.pop('KAPRE_B200_TF'); os.environ.pop('KAPRE_B200_NW')
    # inverse
    stft, istft = K.get_perfectly_reconstructing_stft_istft(1024, 256, 'channels_last', 'channels_last')
    x = torch.rand((128, 16000, 1), device='cuda') * 2 - 1
    S = stft(x)
    for inw in (2, 4, 8):
        os.environ['KAPRE_B200_INW'] = str(inw)
        ms = time_it(lambda: istft(S))
        print('istft INW=%d %.3f ms %.3e frames/s' % (inw, ms, 128 * S.shape[1] / ms * 1e3), flush=True)
    # cfg3-like: 6ch channels_last, n_fft 2048 hop 1024, mag + dB
    x3 = torch.rand((256, 44100, 6), device='cuda') * 2 - 1
    l3 = K.get_stft_magnitude_layer(n_fft=2048, hop_length=1024, return_decibel=True, input_data_format='channels_last',
                                    output_data_format='channels_last')
    ms = time_it(lambda: l3(x3), iters=5)
    print('cfg3/4 (B=256) %.3f ms %.3e frames/s [%s]' % (ms, 256 * 6 * 42 / ms * 1e3, _native.last_launch_info()), flush=True)


if __name__ == '__main__':
    main()
