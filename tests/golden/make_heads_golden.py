"""Generates tests/golden/ref_heads_golden.pt from the REFERENCE's own head and proposal code, run in the
build container where /root/reference exists.  The reference package cannot be imported (no mmcv, private
pytorch_metric_learning fork, HRNMPBBoxHead.__init__ unpacks 4 modules into 6 names - SURVEY.md 8c), so the
methods on the inference path are cut out of the files' syntax trees and mounted, unmodified, on bare
nn.Module harnesses whose __init__ sets the attributes the reference's __init__ sets:

  mmdet/models/bbox_heads/hrnmp_bbox_head.py   _add_selsa_with_fc (:134-189), forward_single_selsa (:216-355),
                                               forward_test (:800-909)                                (R9, R10)
  mmdet/models/bbox_heads/selsa_bbox_head.py   _add_selsa_with_fc (:58-88), forward_single_selsa (:108-200),
                                               forward (:203-261)                                     (R10s)
  mmdet/models/anchor_heads/rpn_head.py        get_bboxes_single (:55-104) on the reference's own delta2bbox
                                               (core/bbox/transforms.py) and nms_cpu.cpp (oracle/_ref) (R5)
  mmdet/models/bbox_heads/hrnmp_bbox_head.py   get_det_bboxes (:1009-1052) on the reference's own delta2bbox and
                                               multiclass_nms (core/post_processing/bbox_nms.py)      (R11)

Sizes are reduced (fixture size); weights and inputs are seeded; the q/k projections are scaled so that the
attention is far from uniform.  The fixture stores state dicts, inputs and outputs, so the tests need neither
/root/reference nor this script.

    python tests/golden/make_heads_golden.py
"""
import ast
import importlib.util
import math
import os
import sys
import types
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference/mmdet'


def cut_methods(path, cls_name, names, extra_globals=None):
    """The named methods of class `cls_name` in `path`, compiled from the file's own syntax tree."""
    src = os.path.join(REF, path)
    tree = ast.parse(open(src).read())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == cls_name)
    ns = dict(torch=torch, nn=nn, F=F, math=math, OrderedDict=OrderedDict)
    ns.update(extra_globals or {})
    out = {}
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), src, 'exec'), ns)
            out[node.name] = ns[node.name]
    assert set(out) == set(names), (set(names) - set(out))
    return out


def make_head(path, cls_name, fwd, n_stage, in_channels, roi_feat_size, num_classes, d, t_dim, sampler_num):
    methods = cut_methods(path, cls_name, ['_add_selsa_with_fc', 'forward_single_selsa', fwd])

    def init(self):
        # what BBoxHead.__init__ (bbox_head.py:18-61) and the head's __init__ set for this configuration
        nn.Module.__init__(self)
        self.in_channels, self.roi_feat_area, self.num_classes = in_channels, roi_feat_size * roi_feat_size, num_classes
        self.with_cls = self.with_reg = True
        self.reg_class_agnostic = True
        self.feat_dim = in_channels * self.roi_feat_area
        self.sampler_num, self.t_dim, self.nongt_dim = sampler_num, t_dim, sampler_num * t_dim
        self.fc_feat_dim, self.dim = d, (d, d, d)
        self.non_cur_space = self.output_cur_only = False
        self.conv_z, self.conv_g = [True] * 8, [False] * 8
        mods = self._add_selsa_with_fc(self.feat_dim, d, self.dim, self.conv_z, self.conv_g)
        assert len(mods) == n_stage
        for i, m in enumerate(mods):
            setattr(self, 'selsa_%d' % (i + 1), m)
        self.relu = nn.ReLU(inplace=True)
        self.fc_cls, self.fc_reg = nn.Linear(d, num_classes), nn.Linear(d, 4)
        if n_stage == 4:
            self.fc_cls_2, self.fc_reg_2 = nn.Linear(d, num_classes), nn.Linear(d, 4)
    cls = type('Ref' + cls_name, (nn.Module,), dict(methods, __init__=init))
    return cls()


def seed_weights(head, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in head.named_parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.5 if ('q_data' in name or 'k_data' in name) else 0.08))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    out = {'hrnmp': [], 'selsa': [], 'rpn': [], 'det': []}
    # ---- relation heads
    for seed, (t_dim, P, key, C, d) in enumerate([(3, 6, 1, 4, 32), (5, 4, 0, 2, 16), (4, 5, 3, 4, 32), (1, 7, 0, 4, 16)]):
        for kind in ('hrnmp', 'selsa'):
            if kind == 'hrnmp':
                head = make_head('models/bbox_heads/hrnmp_bbox_head.py', 'HRNMPBBoxHead', 'forward_test', 4,
                                 C, 7, 31, d, t_dim, P)
            else:
                head = make_head('models/bbox_heads/selsa_bbox_head.py', 'SelsaBBoxHead', 'forward', 2,
                                 C, 7, 31, d, t_dim, P)
            seed_weights(head, 100 + seed)
            g = torch.Generator().manual_seed(200 + seed)
            n = t_dim * P - (seed % 2)                       # one frame one proposal short on odd seeds
            feats = torch.rand(n, C, 7, 7, generator=g)
            rng = dict(start=key * P, length=P if key * P + P <= n else n - key * P)
            with torch.no_grad():
                if kind == 'hrnmp':
                    cls, reg = head.forward_test(feats.clone(), cur_range_s=[rng], key_dim=key)
                    res = dict(cls=[c.clone() for c in cls], reg=[r.clone() for r in reg])
                else:
                    cls, reg, _ = head.forward(feats.clone(), cur_range=rng, key_dim=key)
                    res = dict(cls=cls.clone(), reg=reg.clone())
            out[kind].append(dict(sd={k: v.clone() for k, v in head.state_dict().items()}, feats=feats,
                                  start=rng['start'], length=rng['length'], **res))
    # ---- RPN proposal generation on the reference's own delta2bbox and CPU NMS
    from oracle import build
    nms_cpu = build.load_ref() or (build.build_ref() and build.load_ref())
    assert nms_cpu is not None, 'oracle/_ref not built'
    sys.modules.setdefault('mmcv', types.ModuleType('mmcv'))
    transforms = load('core/bbox/transforms.py', 'ref_transforms')
    anchor = load('core/anchor/anchor_generator.py', 'ref_anchor_generator')

    def nms(dets, iou_thr, device_id=None):                   # the CPU branch of nms_wrapper.py:50-61
        inds = dets.new_zeros(0, dtype=torch.long) if dets.shape[0] == 0 else nms_cpu.nms(dets, iou_thr)
        return dets[inds, :], inds
    fn = cut_methods('models/anchor_heads/rpn_head.py', 'RPNHead', ['get_bboxes_single'],
                     dict(delta2bbox=transforms.delta2bbox, nms=nms))['get_bboxes_single']
    me = types.SimpleNamespace(use_sigmoid_cls=True, target_means=[.0, .0, .0, .0], target_stds=[1.0, 1.0, 1.0, 1.0])
    ag = anchor.AnchorGenerator(16, [4, 8, 16, 32], [0.5, 1.0, 2.0])
    for seed, (fh, fw, nms_pre, max_num, min_size) in enumerate([(12, 20, 600, 300, 0), (38, 63, 6000, 300, 0),
                                                                 (9, 14, 100, 40, 16), (6, 6, 6000, 300, 0)]):
        g = torch.Generator().manual_seed(300 + seed)
        cls = torch.randn(12, fh, fw, generator=g) * 2
        reg = torch.randn(48, fh, fw, generator=g) * 0.3
        anchors = ag.grid_anchors((fh, fw), 16, device='cpu')
        cfg = types.SimpleNamespace(nms_across_levels=False, nms_pre=nms_pre, nms_post=max_num, max_num=max_num,
                                    nms_thr=0.7, min_bbox_size=min_size)
        img_shape = (fh * 16 - 8, fw * 16 - 8, 3)
        props = fn(me, [cls], [reg], [anchors], img_shape, 1.0, cfg, False)
        out['rpn'].append(dict(cls=cls, reg=reg, img_shape=img_shape, nms_pre=nms_pre, max_num=max_num,
                               min_bbox_size=min_size, proposals=props.clone()))
    # ---- decode + multiclass NMS of the two head outputs
    wrapper = types.ModuleType('mmdet.ops.nms.nms_wrapper')
    wrapper.nms = nms
    for nme, m in (('mmdet', types.ModuleType('mmdet')), ('mmdet.ops', types.ModuleType('mmdet.ops')),
                   ('mmdet.ops.nms', types.ModuleType('mmdet.ops.nms')), ('mmdet.ops.nms.nms_wrapper', wrapper)):
        sys.modules.setdefault(nme, m)
    sys.modules['mmdet.ops.nms'].nms_wrapper = wrapper
    bbox_nms = load('core/post_processing/bbox_nms.py', 'ref_bbox_nms')
    det = cut_methods('models/bbox_heads/hrnmp_bbox_head.py', 'HRNMPBBoxHead', ['get_det_bboxes'],
                      dict(delta2bbox=transforms.delta2bbox, multiclass_nms=bbox_nms.multiclass_nms,
                           force_fp32=lambda *a, **k: (lambda f: f)))['get_det_bboxes']
    me = types.SimpleNamespace(target_means=[0., 0., 0., 0.], target_stds=[0.1, 0.1, 0.2, 0.2])
    for seed, (n, sf, rescale, shift) in enumerate([(300, 1.0, False, -3.0), (300, 1.6, True, 0.5), (17, 0.625, True, -7.0)]):
        g = torch.Generator().manual_seed(400 + seed)
        xy = torch.rand(n, 2, generator=g) * torch.tensor([800., 450.])
        wh = torch.rand(n, 2, generator=g) * 250 + 8
        rois = torch.cat([torch.zeros(n, 1), xy, xy + wh], 1)
        cls = [torch.randn(n, 31, generator=g) * 2 + torch.cat([torch.zeros(1), torch.full((30,), shift)]) for _ in range(2)]
        reg = [torch.randn(n, 4, generator=g) * torch.tensor([1., 1., 2., 2.]) for _ in range(2)]
        cfg = types.SimpleNamespace(score_thr=0.001, nms=dict(type='nms', iou_thr=0.3), max_per_img=300)
        dets, labels = det(me, rois, cls, reg, (600, 1000, 3), sf, rescale, cfg)
        out['det'].append(dict(rois=rois, cls=cls, reg=reg, scale_factor=sf, rescale=rescale,
                               dets=[d.clone() for d in dets], labels=[l.clone() for l in labels]))
    torch.save(out, os.path.join(HERE, 'ref_heads_golden.pt'))
    print('wrote', {k: len(v) for k, v in out.items()}, [p['proposals'].shape[0] for p in out['rpn']],
          [[d.shape[0] for d in c['dets']] for c in out['det']])


if __name__ == '__main__':
    main()
