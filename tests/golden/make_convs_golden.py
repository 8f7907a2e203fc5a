"""Generates tests/golden/ref_convs_golden.pt from the REFERENCE's own convolution stacks, run in the build
container where /root/reference exists:

  mmdet/models/backbones/resnet.py        ResNet(depth=101, num_stages=3, strides=(1,2,2), out_indices=(2,),
                                          style='caffe', norm_eval=True)                   (R1, hrnmp cfg:40-50)
  mmdet/models/shared_heads/res_layer.py  ResLayer(depth=101, stage=3, stride=1, dilation=2, style='caffe',
                                          external_conv=True)                              (R2, hrnmp cfg:51-60)
  mmdet/models/anchor_heads/rpn_head.py   RPNHead._init_layers / forward_single (:18-35)   (R3)

`import mmdet` fails here (no mmcv; models/__init__ pulls in pycocotools and the broken head modules), so the
three files are imported as themselves inside a skeleton package: the real mmdet/utils, models/registry.py
and models/utils/* are loaded from the reference tree, while mmcv (init helpers, load_checkpoint), mmdet.ops
and mmdet.models.plugins (DCN / attention plugins, unused by this configuration) and mmdet.core.auto_fp16
are empty stand-ins.  The modules are built with the config's kwargs, loaded with hvrnet_b200.synth's
seeded state dict (ResNet-101 weights are far too large for a fixture; the tests regenerate the identical
state dict from the same seed) and run in eval mode on small seeded inputs.  The fixture stores inputs and
outputs only.

    python tests/golden/make_convs_golden.py
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference/mmdet'


def skeleton():
    def pkg(name, path=None, **attrs):
        m = types.ModuleType(name)
        if path is not None:
            m.__path__ = [path]
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    noop = lambda *a, **k: None
    pkg('mmcv', None, is_str=lambda s: isinstance(s, str))
    pkg('mmcv.cnn', None, constant_init=noop, kaiming_init=noop, normal_init=noop, xavier_init=noop, uniform_init=noop)
    pkg('mmcv.runner', None, load_checkpoint=noop)
    sys.modules['mmcv'].cnn, sys.modules['mmcv'].runner = sys.modules['mmcv.cnn'], sys.modules['mmcv.runner']
    pkg('mmdet', REF)
    dummy = type('Unused', (nn.Module,), {})
    pkg('mmdet.ops', None, ContextBlock=dummy, DeformConv=dummy, ModulatedDeformConv=dummy, nms=noop)
    pkg('mmdet.core', None, auto_fp16=lambda *a, **k: (lambda f: f), delta2bbox=noop)
    pkg('mmdet.models', os.path.join(REF, 'models'))
    pkg('mmdet.models.plugins', None, GeneralizedAttention=dummy)
    importlib.import_module('mmdet.utils')                 # real: Registry, build_from_cfg
    importlib.import_module('mmdet.models.registry')       # real
    importlib.import_module('mmdet.models.utils')          # real: ConvModule, build_conv_layer, build_norm_layer
    bb = pkg('mmdet.models.backbones', os.path.join(REF, 'models', 'backbones'))
    resnet = importlib.import_module('mmdet.models.backbones.resnet')      # the real file
    bb.ResNet, bb.make_res_layer = resnet.ResNet, resnet.make_res_layer
    pkg('mmdet.models.shared_heads', os.path.join(REF, 'models', 'shared_heads'))
    res_layer = importlib.import_module('mmdet.models.shared_heads.res_layer')
    ah = pkg('mmdet.models.anchor_heads', os.path.join(REF, 'models', 'anchor_heads'))
    pkg('mmdet.models.anchor_heads.anchor_head', None, AnchorHead=type('AnchorHead', (nn.Module,), {}))
    rpn = importlib.import_module('mmdet.models.anchor_heads.rpn_head')
    return resnet.ResNet, res_layer.ResLayer, rpn.RPNHead


def sub_state(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def main():
    from hvrnet_b200 import synth
    ResNet, ResLayer, RPNHead = skeleton()
    norm_cfg = dict(type='BN', requires_grad=False)
    sd = synth.make_state_dict('hrnmp', seed=0)
    out = {'seed': 0, 'head': 'hrnmp'}
    g = torch.Generator().manual_seed(77)

    trunk = ResNet(depth=101, num_stages=3, strides=(1, 2, 2), dilations=(1, 1, 1), out_indices=(2,),
                   frozen_stages=1, style='caffe', norm_eval=True, norm_cfg=norm_cfg)
    r = trunk.load_state_dict(sub_state(sd, 'backbone.'), strict=False)
    assert not r.unexpected_keys and all('num_batches_tracked' in k for k in r.missing_keys), r
    trunk.eval()
    img = torch.rand(1, 3, 96, 144, generator=g) * 255 - 110
    with torch.no_grad():
        c4 = trunk(img)
    assert isinstance(c4, tuple) and len(c4) == 1
    out['img'], out['c4'] = img, c4[0].clone()

    c5m = ResLayer(depth=101, stage=3, stride=1, dilation=2, style='caffe', norm_eval=True, norm_cfg=norm_cfg,
                   external_conv=True)
    r = c5m.load_state_dict(sub_state(sd, 'shared_head.'), strict=False)
    assert not r.unexpected_keys and all('num_batches_tracked' in k for k in r.missing_keys), r
    c5m.eval()
    x = out['c4']
    with torch.no_grad():
        out['c5'] = c5m(x).clone()

    rpn = RPNHead.__new__(RPNHead)                      # AnchorHead.__init__ needs the anchor / loss machinery;
    nn.Module.__init__(rpn)                             # the three layers are what forward_single uses
    rpn.in_channels, rpn.feat_channels, rpn.num_anchors, rpn.cls_out_channels = 1024, 512, 12, 1
    rpn._init_layers()
    r = rpn.load_state_dict(sub_state(sd, 'rpn_head.'), strict=False)
    assert not r.unexpected_keys and not r.missing_keys, r
    rpn.eval()
    with torch.no_grad():
        cls, reg = rpn.forward_single(x)
    out['rpn_cls'], out['rpn_reg'] = cls.clone(), reg.clone()
    torch.save(out, os.path.join(HERE, 'ref_convs_golden.pt'))
    print('wrote', {k: (tuple(v.shape) if hasattr(v, 'shape') else v) for k, v in out.items()})


if __name__ == '__main__':
    main()
