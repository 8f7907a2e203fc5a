"""Device time of the three graphs of the inter-video step (A: C5 / RPN / proposals / RoIAlign / stages 1-3 + send buffer,
B: branch post-processing + k_4 projection, C: support rows + stage 4 + post-processing) against the single graph of the
same batch without the inter-video stage, on one GPU (world size 1, no collective).   python scripts/inter_graph_times.py [V]"""
import os
import sys
from collections import deque

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import configs, synth  # noqa: E402
from hvrnet_b200.runtime import GraphRunner  # noqa: E402

dev = torch.device('cuda:0')
V = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model, sd, w = configs.build_workload('hrnmp_inter', dev)
T = w['t_dim']
metas = [synth.make_img_meta() for _ in range(T)]
frames = synth.make_frames(T + 1, seed=100)
model.enable_cuda_graphs(True)
dqs = [deque(maxlen=T) for _ in range(V)]
for i in range(T):
    c4 = model(img=torch.cat([frames[(i + v) % (T + 1)][None] for v in range(V)]).to(dev), img_meta=[metas[0]] * V,
               backbone_feat=True)[0]
    for v, t in enumerate(GraphRunner.per_frame(c4)):
        dqs[v].append(t)
wins = [list(d) for d in dqs]
for _ in range(2):
    model.forward_feat_intervideo(wins, metas, n_support=4, rescale=True)
    model.forward_feat_batch(wins, metas, rescale=True)
torch.cuda.synchronize()
r = model._runner
ci = r.last_inter
cb = [c for k, c in r._window.items() if k[0] == 'intra'][0]


def t_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ga, gb, gc = ci.graphs
ta, tb, tc = t_ms(lambda: r._replay(ga)), t_ms(lambda: r._replay(gb)), t_ms(lambda: r._replay(gc))
ti = t_ms(lambda: r._replay(cb))
print('V = %d key frames: graph A %.2f ms (%d launches), B %.2f ms (%d), C %.2f ms (%d): %.2f ms;  intra graph %.2f ms (%d launches);  '
      'difference %.2f ms' % (V, ta, ga.launches, tb, gb.launches, tc, gc.launches, ta + tb + tc, ti, cb.launches, ta + tb + tc - ti))
