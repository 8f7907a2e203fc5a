"""Device time of the two CUDA graphs of one bench step (trunk graph, window graph), V videos."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import configs, synth  # noqa: E402
from hvrnet_b200.runtime import GraphRunner  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 7
dev = torch.device('cuda:0')
m, sd, w = configs.build_workload('hrnmp', dev)
T = w['t_dim']
m.enable_cuda_graphs(True)
meta = synth.make_img_meta()
img = synth.make_frames(V, seed=0).to(dev)
c4 = m(img=img, img_meta=[meta] * V, backbone_feat=True)[0]
wins = [[t for _ in range(T)] for t in GraphRunner.per_frame(c4)]
m.forward_feat_batch(wins, [meta] * T, rescale=True)
torch.cuda.synchronize()


def timeit(fn, n=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


t_trunk = timeit(lambda: m(img=img, img_meta=[meta] * V, backbone_feat=True))
t_win = timeit(lambda: m.forward_feat_batch(wins, [meta] * T, rescale=True))
print('V=%d  trunk graph %.3f ms (%.3f / key frame)   window graph + D2H %.3f ms (%.3f / key frame)'
      % (V, t_trunk, t_trunk / V, t_win, t_win / V))
