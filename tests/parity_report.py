"""Prints the measured relative errors (max |a-b| / max |b|) of the CUDA path against the CPU
oracle at the BASELINE.json frame size - the numbers quoted in DESIGN.md."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import configs, engine, ops, synth  # noqa: E402
from oracle import cref, ref_torch as R  # noqa: E402


def rel(a, b):
    return float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max())


dev = torch.device('cuda:0')
m, sd, w = configs.build_workload('hrnmp', dev)
m.key_dim = 1
frames = synth.make_frames(3, seed=0)
metas = [synth.make_img_meta() for _ in range(3)]
with torch.no_grad():
    c4_ref = R.trunk_forward(sd, frames)
    c5_ref = R.c5_forward(sd, c4_ref)
    cls_ref, reg_ref = R.rpn_forward(sd, c4_ref)
c4 = m.backbone.forward_split(frames.to(dev))
print('C4 (trunk, 91 convs)        %.2e' % rel(ops.nhwc_split_to_nchw(c4), c4_ref))
c4o = ops.nchw_to_nhwc_split(c4_ref.to(dev))
print('C5 (from oracle C4)         %.2e' % rel(m.shared_head.forward_nhwc(c4o).permute(0, 3, 1, 2), c5_ref))
print('C5 (end to end)             %.2e' % rel(m.shared_head.forward_nhwc(c4).permute(0, 3, 1, 2), c5_ref))
o = engine.rpn_forward(m.rpn_head.packed(dev), c4)
print('RPN logits (end to end)     %.2e' % rel(o[..., :12].permute(0, 3, 1, 2), cls_ref))
print('RPN deltas (end to end)     %.2e' % rel(o[..., 12:60].permute(0, 3, 1, 2), reg_ref))
with torch.no_grad():
    ref, raux = R.hnmb_forward_feat(sd, list(c4_ref.split(1)), metas, 1, roi_align_fn=cref.roi_align, return_aux=True)
c4s = [m(img=frames[i:i + 1].to(dev), img_meta=[metas[i]], backbone_feat=True)[0] for i in range(3)]
res, aux = m(x=c4s, img=None, img_meta=metas, forward_feat=True, return_loss=False, rescale=True,
             proposals=[p.to(dev) for p in raux['proposals']], return_aux=True)
for n, a, b in zip(['cls branch', 'cls final', 'reg branch', 'reg final'], aux['cls'] + aux['reg'], raux['cls'] + raux['reg']):
    print('head %-22s %.2e' % (n + ' (oracle proposals)', rel(a, b)))
