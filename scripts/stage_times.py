"""Stage-by-stage device times of one bench step (eager, CUDA events), V videos x T frames."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import configs, engine, ops, synth  # noqa: E402

V = int(sys.argv[1]) if len(sys.argv) > 1 else 7
dev = torch.device('cuda:0')
m, sd, w = configs.build_workload('hrnmp', dev)
T, P = w['t_dim'], 300
meta = synth.make_img_meta()
img = synth.make_frames(V, seed=0).to(dev)
c4_one = m.backbone.forward_split(img)                       # [V,...]
win = ops.Split(c4_one.hi.repeat(T, 1, 1, 1).contiguous(), c4_one.lo.repeat(T, 1, 1, 1).contiguous())
packed = m.bbox_head.packed(dev)
N, Npad, s = T * P, ops.round_up(T * P, 64), m.key_dim * P


def stages():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    ev[0].record()
    m.backbone.forward_split(img)
    ev[1].record()
    maps = m.rpn_head.forward_maps(win)
    ev[2].record()
    props, counts = m.rpn_head.proposals_from_maps(maps, meta['img_shape'], m.test_cfg.rpn)
    ev[3].record()
    c5 = m.shared_head.forward_nhwc(win)
    ev[4].record()
    fidx = torch.arange(V * T, device=dev, dtype=torch.float32).view(V * T, 1, 1).expand(V * T, P, 1)
    rois = torch.cat([fidx, props[..., :4]], -1).view(-1, 5).contiguous()
    rois_p = torch.zeros((V, Npad, 5), device=dev)
    rois_p[:, :N] = rois.view(V, N, 5)
    rows = m.bbox_roi_extractor.roi_layers[0].forward_nhwc_split(c5, rois_p.view(-1, 5))
    ev[5].record()
    o1, o2 = engine.hrnmp_forward_batched(packed, rows, V, N, Npad, s, P)
    ev[6].record()
    for v in range(V):
        rk = rois[v * N + s:v * N + s + P].clone()
        rk[:, 0] = 0
        for o in (o1, o2):
            c, r = m.bbox_head._split_out(o[v * P:(v + 1) * P])
            m.bbox_head.get_det_bboxes(rk, c, r, meta['img_shape'], 1.0, rescale=True, cfg=m.test_cfg.rcnn)
    ev[7].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(7)]


for _ in range(2):
    stages()
ts = [stages() for _ in range(3)]
names = ['trunk (V new frames)', 'rpn convs (V*T frames)', 'proposals', 'c5 convs (V*T frames)', 'roi_align',
         'relation head (batched)', 'decode + multiclass nms']
best = [min(t[i] for t in ts) for i in range(7)]
print('V=%d  T=%d  (ms per step / per key frame)' % (V, T))
for n, t in zip(names, best):
    print('%-28s %8.3f %8.3f' % (n, t, t / V))
print('%-28s %8.3f %8.3f' % ('sum', sum(best), sum(best) / V))
