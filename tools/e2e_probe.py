"""Which part of the bench's end-to-end pattern costs time?  Variants of Sequential.predict usage."""
import json
import sys
import time

import torch

sys.path.insert(0, '.')
import kapre_b200 as K


def main():
    B, L = 256, 110250
    layer = K.get_melspectrogram_layer(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128,
                                       return_decibel=True, input_data_format='channels_last',
                                       output_data_format='channels_last')
    a = torch.empty((B, L, 1), dtype=torch.float32, pin_memory=True).uniform_(-1, 1)
    b = torch.empty((B, L, 1), dtype=torch.float32, pin_memory=True).uniform_(-1, 1)
    res = {'pinned': [a.is_pinned(), b.is_pinned()]}

    def run(name, fn, n=20):
        for i in range(4):
            fn(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            fn(i)
        torch.cuda.synchronize()
        res.setdefault(name, []).append(round((time.perf_counter() - t0) / n * 1e3, 3))

    state = {}

    def v_same_discard(i):
        layer.predict(a)

    def v_same_keep(i):
        state['y'] = layer.predict(a)

    def v_alt_discard(i):
        layer.predict(a if i % 2 == 0 else b)

    def v_alt_keep(i):
        state['y'] = layer.predict(a if i % 2 == 0 else b)

    def v_explicit_bs(i):
        layer.predict(a, batch_size=32)

    def v_same_keep2(i):
        state['y%d' % (i % 2)] = layer.predict(a)

    for rep in range(2):
        run('same_discard', v_same_discard)
        run('same_keep', v_same_keep)
        run('alt_discard', v_alt_discard)
        run('alt_keep', v_alt_keep)
        run('explicit_bs32', v_explicit_bs)
        run('same_keep_two_alive', v_same_keep2)
        state.clear()
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
