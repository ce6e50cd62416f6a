import os, sys
import torch
os.environ['KAPRE_B200_TC'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kapre_b200 as K
torch.cuda.set_device(0)
kw = dict(n_fft=1024, hop_length=256, sample_rate=22050, n_mels=128, input_data_format='channels_first', output_data_format='channels_first')
xb = [torch.rand((256, 1, 110250), device='cuda') * 2 - 1 for _ in range(3)]
layer = K.get_melspectrogram_layer(return_decibel=True, **kw)
for ab in [int(a) for a in os.environ.get('ABL', '0,1,2,3,4,5,6,7,8').split(',')]:
    os.environ['KAPRE_B200_TC_ABLATE'] = str(ab)
    for _ in range(2):
        layer(xb[0])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10):
        layer(xb[i % 3])
    e1.record()
    torch.cuda.synchronize()
    print('ablate %d: %.4f ms/step' % (ab, e0.elapsed_time(e1) / 10), flush=True)
