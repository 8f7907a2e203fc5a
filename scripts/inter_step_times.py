"""Per-step device times of the inter-video loop of bench.py (world size 1) with the next step's trunk prefetched on the side
stream, to see whether slow runs are a few outlier steps or uniformly slow.   python scripts/inter_step_times.py [V] [steps]"""
import os
import sys
from collections import deque

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import configs, synth  # noqa: E402
from hvrnet_b200.runtime import GraphRunner  # noqa: E402

dev = torch.device('cuda:0')
V = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
prefetch = os.environ.get('HVR_NO_PREFETCH') != '1'
model, sd, w = configs.build_workload('hrnmp_inter', dev)
T = w['t_dim']
pool = 3
metas = [synth.make_img_meta() for _ in range(T)]
frames = synth.make_frames(T + pool, seed=100)
devV = [torch.cat([frames[(i + v) % (T + pool)][None] for v in range(V)]).to(dev) for i in range(T + pool)]
model.enable_cuda_graphs(True)
dqs = [deque(maxlen=T) for _ in range(V)]
for i in range(T):
    c4 = model(img=devV[i], img_meta=[metas[0]] * V, backbone_feat=True)[0]
    for v, t in enumerate(GraphRunner.per_frame(c4)):
        dqs[v].append(t)


phases = []


def step(i, inter):
    t0 = time.perf_counter()
    c4 = model(img=devV[T + i % pool], img_meta=[metas[0]] * V, backbone_feat=True)[0]
    t1 = time.perf_counter()
    if prefetch:
        model._runner.prefetch(devV[T + (i + 1) % pool])
    t2 = time.perf_counter()
    for v, t in enumerate(GraphRunner.per_frame(c4)):
        dqs[v].append(t)
    wins = [list(d) for d in dqs]
    t3 = time.perf_counter()
    if inter:
        r = model.forward_feat_intervideo(wins, metas, n_support=4, rescale=True)
    else:
        r = model.forward_feat_batch(wins, metas, rescale=True)
    t4 = time.perf_counter()
    phases.append(tuple(round((b - a) * 1e3, 1) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4))))
    return r


import gc
import time
if os.environ.get('HVR_NO_GC') == '1':
    gc.disable()
trace = []
_orig_replay = model._runner._replay


def _traced_replay(c):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    t0 = time.perf_counter()
    _orig_replay(c)
    hms = (time.perf_counter() - t0) * 1e3
    b.record()
    trace.append((a, b, c.launches, hms))


model._runner._replay = _traced_replay
ext = []
_r = model._runner


_orig_extract = _r.extract
_orig_empty = torch.empty
_orig_copy = torch.Tensor.copy_
fine = {}


def _empty(*a, **k):                                          # host time and cudaMalloc count of every torch.empty inside extract
    n0 = torch.cuda.memory_stats()['segment.all.allocated']
    t0 = time.perf_counter()
    r = _orig_empty(*a, **k)
    fine['empty_ms'] = fine.get('empty_ms', 0.0) + (time.perf_counter() - t0) * 1e3
    fine['cudaMalloc'] = fine.get('cudaMalloc', 0) + torch.cuda.memory_stats()['segment.all.allocated'] - n0
    return r


def _copy(self, *a, **k):
    t0 = time.perf_counter()
    r = _orig_copy(self, *a, **k)
    fine.setdefault('copy_ms', []).append(round((time.perf_counter() - t0) * 1e3, 1))
    return r


def _timed_extract(img):                                      # GraphRunner.extract under host timers
    fine.clear()
    torch.empty, torch.Tensor.copy_ = _empty, _copy
    t0 = time.perf_counter()
    try:
        out = _orig_extract(img)
    finally:
        torch.empty, torch.Tensor.copy_ = _orig_empty, _orig_copy
    ext.append((round((time.perf_counter() - t0) * 1e3, 1), dict(fine)))
    return out


_r.extract = _timed_extract
for inter in (False, True, False, True):
    for i in range(2):
        step(i, inter)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    del trace[:]
    del phases[:]
    del ext[:]
    host = []
    evs[0].record()
    for i in range(steps):
        t0 = time.perf_counter()
        step(2 + i, inter)
        host.append((time.perf_counter() - t0) * 1e3)
        evs[i + 1].record()
    torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    st = torch.cuda.memory_stats()
    print('%s prefetch=%s: mean %.1f ms; steps: %s | reserved %.1f GB, alloc retries %d, cudaMalloc calls %d' % (
        'inter' if inter else 'intra', prefetch, sum(ts) / len(ts), ' '.join('%.0f' % t for t in ts),
        torch.cuda.memory_reserved() / 2**30, st['num_alloc_retries'], st['segment.all.allocated']))
    per = len(trace) // steps
    slow = [i for i, t in enumerate(ts) if t > 1.15 * sorted(ts)[len(ts) // 2]]
    for i in slow[:4] + [steps - 1]:
        print('   step %d host phases (trunk pick-up, prefetch, lists, window call) ms: %s; GraphRunner.extract: %s ms' % (i, phases[i], ext[i]))
        print('   step %d: %.0f ms device, %.0f ms host call; graph replays (launches: ms): %s' % (
            i, ts[i], host[i], ', '.join('%d: %.1f (host %.1f)' % (n, a.elapsed_time(b), h) for a, b, n, h in trace[i * per:(i + 1) * per])))

