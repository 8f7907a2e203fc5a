"""Generates tests/golden/ref_golden.pt from the REFERENCE's own code, run in the build
container (where /root/reference exists).  The reference package itself cannot be imported
(no mmcv, broken in-tree imports - SURVEY.md 8c), so the pure-Python files on the hot path
are loaded standalone with empty stand-ins for the modules they import but do not need:

  mmdet/core/anchor/anchor_generator.py   AnchorGenerator.grid_anchors        (R4)
  mmdet/core/bbox/transforms.py           delta2bbox, bbox2roi, bbox2result   (R6, R13)
  mmdet/core/post_processing/bbox_nms.py  multiclass_nms                      (R12)
  mmdet/ops/nms/src/nms_cpu.cpp           nms (compiled unmodified, oracle/_ref) (R7)
  mmdet/core/evaluation/mean_ap.py        eval_map as tools/vid_eval.py calls it (next row N3)

multiclass_nms runs on the reference's CPU NMS, i.e. the `>=` threshold semantic
(nms_cpu.cpp:55); the oracle reproduces it with strict_gt=False.  Inputs are seeded; the
fixture stores inputs and outputs so the tests need neither /root/reference nor this script.

    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference/mmdet'


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    from oracle import build
    nms_cpu = build.load_ref() or (build.build_ref() and build.load_ref())
    assert nms_cpu is not None, 'oracle/_ref not built'
    sys.modules.setdefault('mmcv', types.ModuleType('mmcv'))          # imported, unused on these paths
    anchor = load('core/anchor/anchor_generator.py', 'ref_anchor_generator')
    transforms = load('core/bbox/transforms.py', 'ref_transforms')

    # stand-in for mmdet.ops.nms.nms_wrapper: the CPU branch of nms_wrapper.py:50-61
    wrapper = types.ModuleType('mmdet.ops.nms.nms_wrapper')

    def nms(dets, iou_thr, device_id=None):
        inds = dets.new_zeros(0, dtype=torch.long) if dets.shape[0] == 0 else nms_cpu.nms(dets, iou_thr)
        return dets[inds, :], inds
    wrapper.nms = nms
    pkg = types.ModuleType('mmdet.ops.nms')
    pkg.nms_wrapper = wrapper
    for n, m in (('mmdet', types.ModuleType('mmdet')), ('mmdet.ops', types.ModuleType('mmdet.ops')),
                 ('mmdet.ops.nms', pkg), ('mmdet.ops.nms.nms_wrapper', wrapper)):
        sys.modules.setdefault(n, m)
    bbox_nms = load('core/post_processing/bbox_nms.py', 'ref_bbox_nms')

    g = torch.Generator().manual_seed(1234)
    out = {}
    # ---- anchors (the config's generator and the doctest's)
    ag = anchor.AnchorGenerator(16, [4, 8, 16, 32], [0.5, 1.0, 2.0])
    out['base_anchors'] = ag.base_anchors.clone()
    out['grid_anchors_38x63'] = ag.grid_anchors((38, 63), 16, device='cpu').clone()
    out['grid_anchors_doctest'] = anchor.AnchorGenerator(9, [1.], [1.]).grid_anchors((2, 2), device='cpu').clone()
    # ---- delta2bbox: RPN style (stds 1) and RCNN style (stds .1 .1 .2 .2), clamp to 600x1000
    n = 500
    xy = torch.rand(n, 2, generator=g) * torch.tensor([900., 500.])
    wh = torch.rand(n, 2, generator=g) * 300 + 2
    rois = torch.cat([xy, xy + wh], 1)
    deltas = torch.randn(n, 4, generator=g) * torch.tensor([0.5, 0.5, 2.5, 2.5])      # some hit the wh clip
    out['d2b_rois'], out['d2b_deltas'] = rois, deltas
    out['d2b_rpn'] = transforms.delta2bbox(rois, deltas, [0., 0., 0., 0.], [1., 1., 1., 1.], (600, 1000))
    out['d2b_rcnn'] = transforms.delta2bbox(rois, deltas * 4, [0., 0., 0., 0.], [0.1, 0.1, 0.2, 0.2], (600, 1000))
    out['d2b_doctest'] = transforms.delta2bbox(
        torch.Tensor([[0., 0., 1., 1.], [0., 0., 1., 1.], [0., 0., 1., 1.], [5., 5., 5., 5.]]),
        torch.Tensor([[0., 0., 0., 0.], [1., 1., 1., 1.], [0., 0., 2., -1.], [0.7, -1.9, -0.5, 0.3]]),
        max_shape=(32, 32))
    # ---- bbox2roi / bbox2result
    bl = [torch.rand(5, 5, generator=g), torch.zeros(0, 5), torch.rand(3, 5, generator=g)]
    out['b2roi_in'] = bl
    out['b2roi'] = transforms.bbox2roi(bl)
    # ---- NMS: the reference's own nms_cpu.cpp (>=), ascending kept indices
    c = torch.rand(800, 2, generator=g) * 300
    s = torch.rand(800, 2, generator=g) * 100 + 4
    dets = torch.cat([c - s / 2, c + s / 2, torch.rand(800, 1, generator=g)], 1)
    out['nms_dets'] = dets
    for thr in (0.3, 0.5, 0.7):
        out['nms_keep_%g' % thr] = nms_cpu.nms(dets, thr)
    out['nms_doctest_dets'] = torch.tensor(
        [[49.1, 32.4, 51.0, 35.9, 0.9], [49.3, 32.9, 51.0, 35.3, 0.9], [49.2, 31.8, 51.0, 35.4, 0.5],
         [35.1, 11.5, 39.1, 15.7, 0.5], [35.6, 11.8, 39.3, 14.2, 0.5], [35.3, 11.5, 39.9, 14.5, 0.4],
         [35.2, 11.7, 39.7, 15.7, 0.3]])
    out['nms_doctest_keep'] = nms_cpu.nms(out['nms_doctest_dets'], 0.7)
    # ---- multiclass_nms (class-agnostic boxes, 31 classes) in both branches of bbox_nms.py:57-61
    boxes = dets[:300, :4].contiguous()
    for tag, shift in (('few', -4.0), ('many', 1.0)):
        scores = torch.softmax(torch.randn(300, 31, generator=g) * 2 + torch.cat(
            [torch.zeros(1), torch.full((30,), shift)]), 1)
        d, l = bbox_nms.multiclass_nms(boxes, scores, 0.001, dict(type='nms', iou_thr=0.3), 300)
        out['mc_%s_scores' % tag], out['mc_%s_dets' % tag], out['mc_%s_labels' % tag] = scores, d, l
    out['mc_boxes'] = boxes
    # ---- eval_map (next row N3): the reference's mean_ap.py as a package with stubbed printing deps
    sys.modules.setdefault('terminaltables', types.ModuleType('terminaltables'))
    sys.modules['terminaltables'].AsciiTable = object
    ev_pkg = types.ModuleType('ref_eval')
    ev_pkg.__path__ = [os.path.join(REF, 'core/evaluation')]
    sys.modules['ref_eval'] = ev_pkg
    for name in ('bbox_overlaps', 'class_names', 'mean_ap'):
        spec = importlib.util.spec_from_file_location('ref_eval.' + name, os.path.join(REF, 'core/evaluation', name + '.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules['ref_eval.' + name] = mod
        try:
            spec.loader.exec_module(mod)
        except Exception:
            if name != 'class_names':
                raise
    mean_ap = sys.modules['ref_eval.mean_ap']
    rs = np.random.RandomState(99)
    n_img, n_cls = 40, 6
    gts, gls, igs, dets = [], [], [], []
    for _ in range(n_img):
        k = rs.randint(0, 4)
        xy = rs.rand(k, 2) * 300
        wh = rs.rand(k, 2) * 150 + 10
        g = np.hstack([xy, xy + wh]).astype(np.float32)
        gts.append(g)
        gls.append(rs.randint(1, n_cls + 1, size=k))
        igs.append(rs.rand(k) < 0.15)
        per = []
        for c in range(n_cls):
            own = g[gls[-1] == c + 1]
            jit = own + rs.randn(*own.shape).astype(np.float32) * 12 if own.shape[0] else np.zeros((0, 4), np.float32)
            dup = jit[:1] + 3 if jit.shape[0] else np.zeros((0, 4), np.float32)
            m = rs.randint(0, 3)
            xy2 = rs.rand(m, 2) * 300
            rnd = np.hstack([xy2, xy2 + rs.rand(m, 2) * 150 + 10]).astype(np.float32)
            b = np.vstack([jit, dup, rnd]).astype(np.float32)
            per.append(np.hstack([b, rs.rand(b.shape[0], 1).astype(np.float32)]))
        dets.append(per)
    m_ap, res = mean_ap.eval_map(dets, gts, gls, gt_ignore=igs, scale_ranges=None, iou_thr=0.5,
                                 dataset=('a', 'b', 'c', 'd', 'e', 'f'), print_summary=False)
    m_ap2, res2 = mean_ap.eval_map(dets, gts, gls, gt_ignore=None, scale_ranges=None, iou_thr=0.5,
                                   dataset=('a', 'b', 'c', 'd', 'e', 'f'), print_summary=False)
    out['ev_dets'], out['ev_gts'], out['ev_labels'], out['ev_ignore'] = dets, gts, gls, igs
    out['ev_map'], out['ev_map_noignore'] = m_ap, m_ap2
    out['ev_cls'] = [dict(num_gts=int(r['num_gts']), num_dets=int(r['num_dets']), ap=float(r['ap']),
                          recall=np.asarray(r['recall']), precision=np.asarray(r['precision'])) for r in res]
    torch.save(out, os.path.join(HERE, 'ref_golden.pt'))
    print('wrote', os.path.join(HERE, 'ref_golden.pt'), sorted(out.keys()))


if __name__ == '__main__':
    main()
