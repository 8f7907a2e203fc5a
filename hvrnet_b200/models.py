"""Host-side mirror of the reference's plugin modules for the per-key-frame inference path.

Every class registers under the reference's type name and accepts the reference's
constructor kwargs (configs/faster_rcnn_r101_hrnmp_c5.py:37-95, faster_rcnn_r101_selsa_c5.py:
16-72), keeps the reference's parameter names (so an mmdet-format state_dict loads), and
exposes the reference's call surface:

    model(img=..., img_meta=..., backbone_feat=True)                    -> (C4,)
    model(x=[C4...], img=None, img_meta=[...], forward_feat=True,
          return_loss=False, rescale=True)                              -> [bbox_result, ...]
    model(img=[tensor], img_meta=[[meta]], return_loss=False)           -> simple_test

(mmdet/models/detectors/base.py:106-132, hnmb_rcnn.py:195-222, selsa_rcnn.py:56-83,
two_stage.py:280-299).  The torch ``nn`` sub-modules below are parameter containers only:
their ``forward`` is never called; all arithmetic runs in libhvr_b200.so through
``engine`` / ``ops``.  Inference only: ``forward_train`` raises.
"""
from collections import abc

import numpy as np
import torch
from torch import nn

from . import engine, ops
from ._lib import HvrError
from .builder import build_backbone, build_head, build_loss, build_roi_extractor, build_shared_head
from .registry import BACKBONES, DETECTORS, HEADS, LOSSES, ROI_EXTRACTORS, SHARED_HEADS

# ----------------------------------------------------------------------------------------
# losses: constructed by the heads (bbox_head.py:48-49, anchor_head.py:71-72), never called
# at inference.  Accepted and ignored.
# ----------------------------------------------------------------------------------------


class _InferenceOnlyLoss(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.cfg = kwargs

    def forward(self, *a, **k):
        raise HvrError('hvrnet_b200 covers the inference path only (losses are training-time)')


@LOSSES.register_module
class CrossEntropyLoss(_InferenceOnlyLoss):
    pass


@LOSSES.register_module
class SmoothL1Loss(_InferenceOnlyLoss):
    pass


# ----------------------------------------------------------------------------------------
# parameter containers with the reference's names
# ----------------------------------------------------------------------------------------
class Bottleneck(nn.Module):
    """resnet.py:86-266 (parameters only; caffe/pytorch style only moves the stride)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=False, style='caffe'):
        super().__init__()
        assert style == 'caffe', 'the path covers the caffe-style bottleneck of both reference configs'
        self.conv1 = nn.Conv2d(inplanes, planes, 1, stride=stride, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        if downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                                            nn.BatchNorm2d(planes * 4))


def make_res_layer(inplanes, planes, blocks, stride=1, dilation=1, style='caffe'):
    """resnet.py:269-329."""
    layers = [Bottleneck(inplanes, planes, stride, dilation, stride != 1 or inplanes != planes * 4, style)]
    for _ in range(1, blocks):
        layers.append(Bottleneck(planes * 4, planes, 1, dilation, False, style))
    return nn.Sequential(*layers)


class _Packed(nn.Module):
    """Mixin: lazily packs the state_dict for the CUDA kernels; repacks after a load."""

    def __init__(self):
        super().__init__()
        self._hvr_packed = None
        self._hvr_version = 0      # bumped whenever the parameters may have changed (CUDA graphs hold raw pointers
                                   # into the packed copies: runtime.GraphRunner re-captures on a version change)

    def _load_from_state_dict(self, *args, **kwargs):
        self._hvr_packed = None
        self._hvr_version += 1
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *a, **k):
        self._hvr_packed = None
        self._hvr_version += 1
        return super()._apply(fn, *a, **k)

    def packed(self, device):
        if self._hvr_packed is None or self._hvr_packed[0] != device:
            sd = {k: v.detach().float().cpu() for k, v in self.state_dict().items()}
            self._hvr_packed = (device, self._pack(sd, device))
        return self._hvr_packed[1]

    def init_weights(self, pretrained=None):
        pass


@BACKBONES.register_module
class ResNet(_Packed):
    """mmdet/models/backbones/resnet.py:332-543 (depth 101 C4 trunk of both configs)."""
    arch_settings = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}

    def __init__(self, depth, in_channels=3, num_stages=4, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1),
                 out_indices=(0, 1, 2, 3), style='pytorch', frozen_stages=-1, conv_cfg=None,
                 norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, dcn=None, stage_with_dcn=None,
                 gcb=None, stage_with_gcb=None, gen_attention=None, stage_with_gen_attention=None,
                 with_cp=False, zero_init_residual=True):
        super().__init__()
        if depth not in self.arch_settings:
            raise KeyError('invalid depth {} for resnet'.format(depth))
        assert dcn is None and gcb is None and gen_attention is None, 'plugins are outside the hot path'
        assert in_channels == 3 and len(strides) == len(dilations) == num_stages
        assert tuple(out_indices) == (num_stages - 1,), 'single-level (C4) output only'
        assert norm_cfg.get('type', 'BN') == 'BN' and norm_eval, 'frozen eval-mode BN only'
        self.depth, self.num_stages, self.strides, self.dilations = depth, num_stages, tuple(strides), tuple(dilations)
        self.out_indices = tuple(out_indices)
        blocks = self.arch_settings[depth][:num_stages]
        self.stage_blocks = blocks
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        self.res_layers = []
        for i, nb in enumerate(blocks):
            planes = 64 * 2 ** i
            name = 'layer{}'.format(i + 1)
            self.add_module(name, make_res_layer(inplanes, planes, nb, strides[i], dilations[i], style))
            inplanes = planes * 4
            self.res_layers.append(name)
        self.feat_dim = inplanes
        self.eval()

    def _pack(self, sd, device):
        P = engine.pack_trunk(sd, device, prefix='', strides=self.strides, dilations=self.dilations)
        assert tuple(len(l) for l in P['layers']) == tuple(self.stage_blocks)
        return P

    def forward_split(self, img):
        return engine.trunk_forward(self.packed(img.device), img)

    def forward(self, x):
        """[B,3,H,W] fp32 CUDA -> (C4 [B,1024,H/16,W/16] fp32,), resnet.py:522-533.  The NHWC
        split copy rides along as ``._hvr_split`` so later stages skip the re-conversion."""
        s = self.forward_split(x)
        out = ops.nhwc_split_to_nchw(s)
        out._hvr_split = s
        return (out,)


@SHARED_HEADS.register_module
class ResLayer(_Packed):
    """mmdet/models/shared_heads/res_layer.py:16-74 (layer4 on the whole map + new_layer_1)."""

    def __init__(self, depth, stage=3, stride=2, dilation=1, style='pytorch', norm_cfg=dict(type='BN'),
                 norm_eval=True, with_cp=False, dcn=None, external_conv=False):
        super().__init__()
        assert dcn is None
        self.stage, self.stride, self.dilation, self.external_conv = stage, stride, dilation, external_conv
        blocks = ResNet.arch_settings[depth][stage]
        planes = 64 * 2 ** stage
        inplanes = 64 * 2 ** (stage - 1) * 4
        self.add_module('layer{}'.format(stage + 1), make_res_layer(inplanes, planes, blocks, stride, dilation, style))
        if external_conv:
            self.new_layer_1 = nn.Module()
            self.new_layer_1.conv = nn.Conv2d(2048, 256, 1)      # ConvModule(2048,256,1): bias, no norm, ReLU
        self.eval()

    def _pack(self, sd, device):
        return engine.pack_c5(sd, device, prefix='', stride=self.stride, dilation=self.dilation)

    def forward_nhwc(self, c4_split):
        return engine.c5_forward(self.packed(c4_split.hi.device), c4_split)

    def forward(self, x):
        s = getattr(x, '_hvr_split', None)
        if s is None:
            s = ops.nchw_to_nhwc_split(x)
        f = self.forward_nhwc(s)
        out = ops.nhwc_to_nchw(f)
        out._hvr_nhwc = f
        return out


def gen_base_anchors(base_size, scales, ratios):
    """anchor_generator.py:29-56 (ratio-major, scale-minor, rounded)."""
    scales = torch.tensor(scales, dtype=torch.float32)
    ratios = torch.tensor(ratios, dtype=torch.float32)
    w = h = float(base_size)
    xc, yc = 0.5 * (w - 1), 0.5 * (h - 1)
    hr = torch.sqrt(ratios)
    wr = 1 / hr
    ws = (w * wr[:, None] * scales[None, :]).reshape(-1)
    hs = (h * hr[:, None] * scales[None, :]).reshape(-1)
    return torch.stack([xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1), xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)],
                       dim=-1).round()


@HEADS.register_module
class RPNHead(_Packed):
    """mmdet/models/anchor_heads/rpn_head.py:14-104 + anchor_head.py:24-98,209-278 (inference)."""

    def __init__(self, in_channels, feat_channels=256, anchor_scales=[8, 16, 32], anchor_ratios=[0.5, 1.0, 2.0],
                 anchor_strides=[4, 8, 16, 32, 64], anchor_base_sizes=None, target_means=(.0, .0, .0, .0),
                 target_stds=(1.0, 1.0, 1.0, 1.0), loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=True),
                 loss_bbox=dict(type='SmoothL1Loss', beta=1.0 / 9.0), **kwargs):
        super().__init__()
        assert len(anchor_strides) == 1, 'single-level (C4) RPN only'
        assert loss_cls.get('use_sigmoid', False), 'sigmoid objectness (cls_out_channels = 1)'
        assert tuple(target_means) == (0., 0., 0., 0.) and tuple(target_stds) == (1., 1., 1., 1.)
        self.in_channels, self.feat_channels = in_channels, feat_channels
        self.anchor_scales, self.anchor_ratios, self.anchor_strides = anchor_scales, anchor_ratios, anchor_strides
        self.anchor_base_sizes = list(anchor_strides) if anchor_base_sizes is None else anchor_base_sizes
        self.num_anchors = len(anchor_ratios) * len(anchor_scales)
        self.loss_cls, self.loss_bbox = build_loss(loss_cls), build_loss(loss_bbox)
        self.rpn_conv = nn.Conv2d(in_channels, feat_channels, 3, padding=1)
        self.rpn_cls = nn.Conv2d(feat_channels, self.num_anchors, 1)
        self.rpn_reg = nn.Conv2d(feat_channels, self.num_anchors * 4, 1)
        self.base_anchors = gen_base_anchors(self.anchor_base_sizes[0], anchor_scales, anchor_ratios)
        self.eval()

    def _pack(self, sd, device):
        P = engine.pack_rpn(sd, device, prefix='')
        P['base'] = self.base_anchors.to(device)
        return P

    def forward_maps(self, c4_split):
        """rpn_head.py:30-35 for all frames -> fp32 [T,h,w,64]: columns [0,A) logits, [A,5A) deltas."""
        return engine.rpn_forward(self.packed(c4_split.hi.device), c4_split)

    def proposals_from_maps(self, o, img_shape, cfg, want_idx=False):
        """anchor_head.py:209-278 + rpn_head.py:55-104, all frames in one pass ->
        proposals [T,max_num,5], counts [T] (device)."""
        P = self.packed(o.device)
        T, h, w, ld = o.shape
        A = self.num_anchors
        assert not cfg.get('nms_across_levels', False) and cfg.get('min_bbox_size', 0) == 0
        return ops.rpn_proposals(o, o.view(-1)[A:], ld, ld, T, h, w, A, P['base'], self.anchor_strides[0],
                                 img_shape[:2], cfg['nms_pre'], cfg['nms_post'], cfg['max_num'], cfg['nms_thr'],
                                 want_idx=want_idx)

    def get_proposals(self, c4_split, img_shape, cfg, want_idx=False):
        return self.proposals_from_maps(self.forward_maps(c4_split), img_shape, cfg, want_idx)


class RoIAlign(nn.Module):
    """mmdet/ops/roi_align/roi_align.py:59-80.  features NCHW fp32 CUDA, rois [n,5] ->
    [n,C,out,out] (reference layout).  No CPU path (as in the reference, :24-28)."""

    def __init__(self, out_size, spatial_scale, sample_num=0, use_torchvision=False):
        super().__init__()
        assert not use_torchvision
        self.out_size = (out_size, out_size) if isinstance(out_size, int) else tuple(out_size)
        assert self.out_size[0] == self.out_size[1]
        self.spatial_scale = float(spatial_scale)
        self.sample_num = int(sample_num)

    def forward(self, features, rois):
        if not features.is_cuda:
            raise NotImplementedError   # roi_align.py:27-28
        return ops.roi_align(features, rois, self.out_size[0], self.spatial_scale, self.sample_num)

    # arithmetic of the pipeline variant below: 'fast' (separable FMA evaluation, 1e-5 relative to the strict
    # one) or 'strict' (bit-exact with the reference kernel).  forward() - the reference-facing op - is always strict.
    pipeline_arithmetic = 'fast'

    def forward_nhwc_split(self, feat_nhwc, rois):
        """Pipeline variant: NHWC fp32 map -> Split [n, out*out*C] rows for fc_new_1."""
        return ops.roi_align(feat_nhwc, rois, self.out_size[0], self.spatial_scale, self.sample_num, feat_nhwc=True,
                             out_nhwc=True, want_split=True, want_f32=False, arithmetic=self.pipeline_arithmetic)[1]

    def __repr__(self):
        return '{}(out_size={}, spatial_scale={}, sample_num={})'.format(self.__class__.__name__, self.out_size,
                                                                         self.spatial_scale, self.sample_num)


ROI_LAYERS = {'RoIAlign': RoIAlign}     # stands in for getattr(mmdet.ops, layer_type), single_level.py:45-52


@ROI_EXTRACTORS.register_module
class SingleRoIExtractor(nn.Module):
    """mmdet/models/roi_extractors/single_level.py:11-107 (single-level fast path :90-92)."""

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56):
        super().__init__()
        cfg = dict(roi_layer)
        layer_type = cfg.pop('type')
        assert layer_type in ROI_LAYERS, layer_type
        assert len(featmap_strides) == 1, 'single-level (C5) extraction only'
        self.roi_layers = nn.ModuleList([ROI_LAYERS[layer_type](spatial_scale=1 / s, **cfg) for s in featmap_strides])
        self.out_channels, self.featmap_strides, self.finest_scale = out_channels, featmap_strides, finest_scale

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def init_weights(self):
        pass

    def forward(self, feats, rois, roi_scale_factor=None):
        assert len(feats) == 1
        return self.roi_layers[0](feats[0], rois)


class BBoxHead(_Packed):
    """mmdet/models/bbox_heads/bbox_head.py:14-169 (inference parts)."""
    kind = None

    def __init__(self, with_avg_pool=False, with_cls=True, with_reg=True, roi_feat_size=7, in_channels=256,
                 num_classes=81, target_means=[0., 0., 0., 0.], target_stds=[0.1, 0.1, 0.2, 0.2],
                 reg_class_agnostic=False, loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0),
                 loss_bbox=dict(type='SmoothL1Loss', beta=1.0, loss_weight=1.0)):
        super().__init__()
        assert with_cls and with_reg and not with_avg_pool
        assert reg_class_agnostic, 'class-agnostic regression (both reference configs)'
        assert tuple(target_means) == (0., 0., 0., 0.)
        self.roi_feat_size = (roi_feat_size, roi_feat_size) if isinstance(roi_feat_size, int) else roi_feat_size
        self.roi_feat_area = self.roi_feat_size[0] * self.roi_feat_size[1]
        self.in_channels, self.num_classes = in_channels, num_classes
        self.target_means, self.target_stds = target_means, target_stds
        self.reg_class_agnostic = reg_class_agnostic
        self.loss_cls, self.loss_bbox = build_loss(loss_cls), build_loss(loss_bbox)

    def _pack(self, sd, device):
        return engine.pack_head(sd, device, self.kind, prefix='', roi_channels=self.in_channels,
                                roi_size=self.roi_feat_size[0])

    def _as_rows(self, roi_feats):
        """Accepts the pipeline's Split [N, h*w*C] rows or a reference-layout fp32 [N,C,h,w]."""
        if isinstance(roi_feats, ops.Split):
            return roi_feats
        x = roi_feats.permute(0, 2, 3, 1).reshape(roi_feats.shape[0], -1)
        return ops.split(x)

    def _split_out(self, o):
        nc = self.num_classes
        return o[:, :nc], o[:, nc:nc + 4]

    def get_det_bboxes(self, rois, cls_score, bbox_pred, img_shape, scale_factor, rescale=False, cfg=None):
        """bbox_head.py:132-169 / hrnmp_bbox_head.py:1009-1052 -> device tensors
        (dets [max,5], labels [max], n [1]) per head output."""
        single = not isinstance(cls_score, (list, tuple))
        cs, bp = ([cls_score], [bbox_pred]) if single else (cls_score, bbox_pred)
        if isinstance(scale_factor, (np.ndarray, list, tuple)):
            scale_factor = float(np.asarray(scale_factor).reshape(-1)[0])
        outs = []
        for c, r in zip(cs, bp):
            nms_cfg = dict(cfg['nms'])
            assert nms_cfg.pop('type', 'nms') == 'nms'
            outs.append(ops.det_postprocess(rois, c, r, img_shape[:2], scale_factor, rescale, self.target_stds,
                                            cfg['score_thr'], nms_cfg['iou_thr'], cfg['max_per_img'],
                                            n_cls=self.num_classes))
        return outs[0] if single else outs

    def get_det_bboxes_batched(self, rois, cls_score, bbox_pred, G, img_shape, scale_factor, rescale=False, cfg=None):
        """get_det_bboxes for the key frames of G videos at once (rows of video g = [g*n, (g+1)*n)):
        one launch per post-processing stage; per video the same bits as get_det_bboxes.
        Returns (dets [G,max,5], labels [G,max], n [G])."""
        if isinstance(scale_factor, (np.ndarray, list, tuple)):
            scale_factor = float(np.asarray(scale_factor).reshape(-1)[0])
        nms_cfg = dict(cfg['nms'])
        assert nms_cfg.pop('type', 'nms') == 'nms'
        return ops.det_postprocess_batched(rois, cls_score, bbox_pred, G, img_shape[:2], scale_factor, rescale,
                                           self.target_stds, cfg['score_thr'], nms_cfg['iou_thr'],
                                           cfg['max_per_img'], n_cls=self.num_classes)


@HEADS.register_module
class SharedFCBBoxHead(BBoxHead):
    """mmdet/models/bbox_heads/convfc_bbox_head.py:170-185 (2 shared fcs) - config 1."""
    kind = 'shared_fc'

    def __init__(self, num_fcs=2, fc_out_channels=1024, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert num_fcs == 2
        feat = self.in_channels * self.roi_feat_area
        self.shared_fcs = nn.ModuleList([nn.Linear(feat, fc_out_channels), nn.Linear(fc_out_channels, fc_out_channels)])
        self.fc_cls = nn.Linear(fc_out_channels, self.num_classes)
        self.fc_reg = nn.Linear(fc_out_channels, 4)

    def forward(self, roi_feats):
        x = self._as_rows(roi_feats)
        return self._split_out(engine.shared_fc_forward(self.packed(x.hi.device), x))


def _selsa_block(k, fc_feat_dim, dim):
    return nn.ModuleDict({'q_data_fc_%d' % k: nn.Linear(fc_feat_dim, dim[0]),
                          'k_data_fc_%d' % k: nn.Linear(fc_feat_dim, dim[1]),
                          'linear_out_%d' % k: nn.Conv2d(dim[2], dim[2], 1)})


@HEADS.register_module
class SelsaBBoxHead(BBoxHead):
    """mmdet/models/bbox_heads/selsa_bbox_head.py:16-261 (2 relation stages)."""
    kind = 'selsa'
    stages = 2

    def __init__(self, sampler_num, t_dim, imgs_per_video=None, fc_feat_dim=1024, non_cur_space=False,
                 dim=(1024, 1024, 1024), output_cur_only=False, conv_z=None, conv_g=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert tuple(dim) == (fc_feat_dim,) * 3 and not non_cur_space
        # what _add_selsa_with_fc builds with the configs' defaults (hrnmp_bbox_head.py:146-149,339-347): no value
        # projection v_data_fc_k (conv_g False), output projection linear_out_k present (conv_z True).  Other
        # settings would need layers this path does not evaluate: refuse them instead of computing something else.
        assert conv_z is None or all(bool(z) for z in conv_z), 'conv_z=False (no linear_out_k) is not on the hot path'
        assert conv_g is None or not any(bool(g) for g in conv_g), 'conv_g=True (v_data_fc_k) is not on the hot path'
        assert not output_cur_only, 'output_cur_only is not on the hot path'
        self.feat_dim = self.in_channels * self.roi_feat_area
        self.sampler_num, self.t_dim, self.imgs_per_video = sampler_num, t_dim, imgs_per_video
        self.nongt_dim = sampler_num * t_dim
        self.fc_feat_dim, self.dim = fc_feat_dim, dim
        for k in range(1, self.stages + 1):
            setattr(self, 'fc_new_%d' % k, nn.Linear(self.feat_dim if k == 1 else dim[2], fc_feat_dim))
            setattr(self, 'selsa_%d' % k, _selsa_block(k, fc_feat_dim, dim))
        self.fc_cls = nn.Linear(dim[2], self.num_classes)
        self.fc_reg = nn.Linear(dim[2], 4)

    def forward(self, roi_feats, cur_range, key_dim=None):
        """selsa_bbox_head.py:203-261 -> (cls [len,n_cls], reg [len,4], None)."""
        x = self._as_rows(roi_feats)
        assert self.nongt_dim >= x.shape[0]                            # :126
        r = cur_range[0] if isinstance(cur_range, (list, tuple)) else cur_range
        o = engine.selsa_forward(self.packed(x.hi.device), x, int(r['start']), int(r['length']))
        cls, reg = self._split_out(o)
        return cls, reg, None


@HEADS.register_module
class HRNMPBBoxHead(SelsaBBoxHead):
    """mmdet/models/bbox_heads/hrnmp_bbox_head.py:56-1052 (4 stages; what _add_selsa_with_fc
    builds, :134-189 - the 6-way unpack at :100-103 is a reference bug)."""
    kind = 'hrnmp'
    stages = 4

    def __init__(self, sampler_num, t_dim, imgs_per_video, *args, **kwargs):
        super().__init__(sampler_num, t_dim, imgs_per_video, *args, **kwargs)
        self.fc_cls_2 = nn.Linear(self.dim[2], self.num_classes)
        self.fc_reg_2 = nn.Linear(self.dim[2], 4)

    def forward(self, *a, **k):
        raise HvrError('HRNMPBBoxHead.forward is the training path (hrnmp_bbox_head.py:609-795); use forward_test')

    def forward_test(self, roi_feats, cur_range, key_dim=None, all_res=False, support=None, return_support=False):
        """hrnmp_bbox_head.py:800-909 -> ([cls_branch, cls], [reg_branch, reg]).
        ``support``: Split [M, 1024] post-fc_new_4 key rows of other videos appended to stage 4's
        key/value set (inter-video definition)."""
        x = self._as_rows(roi_feats)
        assert self.nongt_dim >= x.shape[0]                            # :249
        r = cur_range[0] if isinstance(cur_range, (list, tuple)) else cur_range
        o1, o2, f4k = engine.hrnmp_forward_test(self.packed(x.hi.device), x, int(r['start']), int(r['length']),
                                                support=support)
        c1, r1 = self._split_out(o1)
        c2, r2 = self._split_out(o2)
        if return_support:
            return [c1, c2], [r1, r2], f4k
        return [c1, c2], [r1, r2]


# ----------------------------------------------------------------------------------------
# detectors
# ----------------------------------------------------------------------------------------
def bbox2roi(bbox_list):
    """transforms.py:149-168."""
    out = []
    for i, b in enumerate(bbox_list):
        if b.shape[0] > 0:
            out.append(torch.cat([b.new_full((b.shape[0], 1), i), b[:, :4]], dim=-1))
        else:
            out.append(b.new_zeros((0, 5)))
    return torch.cat(out, 0)


def bbox2result(bboxes, labels, num_classes):
    """transforms.py:181-199 -> list of (num_classes-1) float32 [k,5] arrays."""
    if bboxes.shape[0] == 0:
        return [np.zeros((0, 5), dtype=np.float32) for _ in range(num_classes - 1)]
    b = bboxes.cpu().numpy()
    lab = labels.cpu().numpy()
    return [b[lab == i, :] for i in range(num_classes - 1)]


def _result_from_device(dets, labels, n, num_classes):
    """One D2H per output (the reference's bbox2result D2H, hnmb_rcnn.py:214-218)."""
    k = int(n.item())
    return bbox2result(dets[:k], labels[:k], num_classes)


class TwoStageDetector(nn.Module):
    """two_stage.py:20-97 + base.py:106-132 (inference)."""

    def __init__(self, backbone, neck=None, shared_head=None, rpn_head=None, bbox_roi_extractor=None, bbox_head=None,
                 mask_roi_extractor=None, mask_head=None, train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__()
        assert neck is None and mask_head is None and mask_roi_extractor is None
        self.backbone = build_backbone(backbone)
        if shared_head is not None:
            self.shared_head = build_shared_head(shared_head)
        if rpn_head is not None:
            self.rpn_head = build_head(rpn_head)
        if bbox_head is not None:
            bbox_roi_extractor = dict(bbox_roi_extractor)
            self.feat_from_shared_head = bbox_roi_extractor.pop('feat_from_shared_head', False)   # two_stage.py:46
            self.bbox_roi_extractor = build_roi_extractor(bbox_roi_extractor)
            self.bbox_head = build_head(bbox_head)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.init_weights(pretrained=pretrained)
        self.eval()

    with_neck = False
    with_mask = False

    @property
    def with_shared_head(self):
        return hasattr(self, 'shared_head')

    @property
    def with_rpn(self):
        return hasattr(self, 'rpn_head')

    @property
    def with_bbox(self):
        return hasattr(self, 'bbox_head')

    def init_weights(self, pretrained=None):
        pass

    _runner = None

    def weights_version(self):
        """Changes whenever a load_state_dict / .to() / _apply touched any packed module."""
        packed = self.__dict__.get('_hvr_packed_modules')
        if packed is None:
            packed = [m for m in self.modules() if isinstance(m, _Packed)]
            self.__dict__['_hvr_packed_modules'] = packed       # the module tree of a built detector is fixed
        return tuple(m._hvr_version for m in packed)

    def enable_cuda_graphs(self, flag=True, capture=True):
        """Run the trunk and the window stage through captured CUDA graphs (runtime.GraphRunner):
        same kernels and results, two graph launches and one device->host read per key frame.
        capture=False keeps the runner's batched launch sequence but issues it eagerly."""
        from .runtime import GraphRunner
        self._runner = GraphRunner(self, capture=capture) if flag else None
        return self

    _streamer = None

    def enable_streaming(self, n_videos=1, window=None, flag=True, capture=True):
        """Streaming scheduler behind the detector's own call surface: afterwards
        ``model(img=frames [V,3,H,W], img_meta=[meta], stream=True, rescale=...)`` advances V video streams by one
        frame each and returns None while the windows fill, then the V forward_feat results of the windows' key
        frames - the same detections, bit for bit, as backbone_feat=True + forward_feat=True on the same windows,
        with the per-frame stages (trunk, C5, RPN, proposals, RoIAlign, fc_new_1) computed once per frame instead of
        once per window (runtime.StreamGraphRunner: two CUDA graphs and one device->host read per step)."""
        from .runtime import StreamGraphRunner
        self._streamer = StreamGraphRunner(self, n_videos, window=window, capture=capture) if flag else None
        return self

    def extract_feat(self, img):
        """two_stage.py:91-97 -> (C4,).  `img` may live in pinned host memory when CUDA graphs
        are enabled (it is copied straight into the trunk graph's input buffer)."""
        if self._runner is not None:
            return self._runner.extract(img)
        if not img.is_cuda:
            raise HvrError('hvrnet_b200 has no CPU path: move the model inputs to a CUDA device')
        return self.backbone(img)

    def forward_train(self, *a, **k):
        raise HvrError('hvrnet_b200 covers the per-key-frame inference path only')

    def forward_test(self, imgs, img_metas, **kwargs):
        """base.py:77-104."""
        for var, name in [(imgs, 'imgs'), (img_metas, 'img_metas')]:
            if not isinstance(var, list):
                raise TypeError('{} must be a list, but got {}'.format(name, type(var)))
        if len(imgs) != len(img_metas):
            raise ValueError('num of augmentations ({}) != num of image meta ({})'.format(len(imgs), len(img_metas)))
        assert imgs[0].size(0) == 1
        assert len(imgs) == 1, 'test-time augmentation is outside the hot path'
        return self.simple_test(imgs[0], img_metas[0], **kwargs)

    @torch.no_grad()
    def forward(self, img=None, img_meta=None, return_loss=True, backbone_feat=False, forward_feat=False, stream=False,
                **kwargs):
        """base.py:106-132 (+ stream=True: enable_streaming's per-frame entry)."""
        if stream:
            if self._streamer is None:
                raise HvrError('call enable_streaming(n_videos, window) first')
            meta = img_meta[0] if isinstance(img_meta, (list, tuple)) else img_meta
            return self._streamer.push(img, meta, rescale=kwargs.get('rescale', True))
        if backbone_feat:
            if isinstance(img, list):
                assert len(img) == len(img_meta), 'img and img_meta should have same number!'
                return [self.extract_feat(im_) for im_ in img]
            return self.extract_feat(img)
        if forward_feat:
            return self.forward_feat(img_meta=img_meta, **kwargs)
        if return_loss:
            return self.forward_train(img, img_meta, **kwargs)
        return self.forward_test(img, img_meta, **kwargs)

    # ---- shared by the window detectors -------------------------------------------------
    def _window_split(self, x):
        """list of per-frame C4 tensors (or one stacked tensor) -> Split NHWC [T,h,w,C]."""
        if isinstance(x, torch.Tensor):
            x = [x]
        parts = []
        for t in x:
            s = getattr(t, '_hvr_split', None)
            if s is None:
                s = ops.nchw_to_nhwc_split(t)
            parts.append(s)
        if len(parts) == 1:
            return parts[0]
        return ops.Split(torch.cat([p.hi for p in parts], 0), torch.cat([p.lo for p in parts], 0))

    def _detect(self, c4, img_meta, V, T, key_dim, rescale, proposals=None, return_aux=False):
        """window.detect_windows + the one device->host read of the step -> per video the list (one entry per head
        output) of bbox2result lists; optionally the intermediate tensors (tests / parity reports)."""
        from . import window
        result, fs, outs = window.detect_windows(self, c4, img_meta, V, T, key_dim, rescale, proposals=proposals)
        counts, per_video = result.parse(result.buf.cpu())
        nc = self.bbox_head.num_classes
        res = [[bbox2result(d, l, nc) for d, l in pv] for pv in per_video]
        if not return_aux:
            return res
        assert V == 1, 'return_aux is a single-window facility'
        P = fs.P
        n = counts[key_dim]
        rois = torch.cat([fs.rois[t * P:t * P + counts[t]] for t in range(T)], 0)     # the reference's compact roi list
        aux = dict(c5=fs.c5, proposals=fs.props, counts=counts, rois=rois, start=int(sum(counts[:key_dim])), length=n,
                   cls=[self.bbox_head._split_out(o)[0][:n] for o in outs],
                   reg=[self.bbox_head._split_out(o)[1][:n] for o in outs],
                   dets=[(d[0], l[0], k[0:1]) for d, l, k in result.outs], frame_stages=fs)
        return res, aux


@DETECTORS.register_module
class FasterRCNN(TwoStageDetector):
    """detectors/faster_rcnn.py + two_stage.py:280-299 + test_mixins.py:9-13,40-69 (config 1)."""

    key_dim = 0          # a "window" of one frame whose key frame is itself (runtime.GraphRunner.detect)

    def simple_test(self, img, img_meta, proposals=None, rescale=False):
        c4 = self.extract_feat(img)[0]._hvr_split
        if self._runner is not None and proposals is None:
            d, l = self._runner.detect([[c4]], img_meta, rescale)[0][0]
            return bbox2result(d, l, self.bbox_head.num_classes)
        return self._detect(c4, img_meta, 1, 1, 0, rescale, proposals=proposals)[0][0]

    def simple_test_batch(self, img, img_meta, rescale=False):
        """Throughput extension (the reference's test loop feeds one image per call, two_stage.py:280-299): V images
        [V,3,H,W] of the same shape in one call -> list of V simple_test results (per image the same arithmetic);
        with CUDA graphs enabled: one trunk graph + one graph for everything after it."""
        c4 = self.extract_feat(img)[0]
        V = c4.shape[0]
        if self._runner is not None:
            from .runtime import GraphRunner
            outs = self._runner.detect([[t._hvr_split] for t in GraphRunner.per_frame(c4)], img_meta, rescale)
            return [bbox2result(out[0][0], out[0][1], self.bbox_head.num_classes) for out in outs]
        return [r[0] for r in self._detect(c4._hvr_split, img_meta, V, 1, 0, rescale)]


class _WindowRCNN(TwoStageDetector):

    def _key_range(self, cnt):
        start = int(sum(cnt[:self.key_dim]))                             # hnmb_rcnn.py:586 (int repair)
        return [dict(start=start, length=int(cnt[self.key_dim]))]

    def forward_feat(self, x=None, img_meta=None, proposals=None, rescale=False, support=None, return_aux=False):
        """hnmb_rcnn.py:195-222 / selsa_rcnn.py:56-83: one key frame from a window of C4 maps."""
        assert x is not None and img_meta is not None
        if isinstance(x, abc.Sequence):
            assert len(x) == len(img_meta)
            assert isinstance(x[0], torch.Tensor)
        if (self._runner is not None and proposals is None and support is None and not return_aux
                and not isinstance(x, torch.Tensor) and all(hasattr(t, '_hvr_split') for t in x)):
            out = self._runner.detect([[t._hvr_split for t in x]], img_meta, rescale)[0]
            return [bbox2result(d, l, self.bbox_head.num_classes) for d, l in out]
        if support is not None:
            raise HvrError('forward_feat(support=...) was replaced by forward_feat_intervideo (fixed-size exchange)')
        c4 = self._window_split(x)
        T = c4.shape[0]
        out = self._detect(c4, img_meta, 1, T, self.key_dim, rescale, proposals=proposals, return_aux=return_aux)
        if return_aux:
            return out[0][0], out[1]
        return out[0]

    def forward_feat_batch(self, xs, img_meta, rescale=False):
        """Throughput extension (not in the reference, whose driver handles one video per
        process): V windows of V different videos in one call -> list of V forward_feat results.
        Each video's arithmetic is exactly forward_feat's; with CUDA graphs enabled the V*T frames
        share the C5 / RPN / RoIAlign launches."""
        V, T = len(xs), len(xs[0])
        assert all(len(x) == T for x in xs)
        if self._runner is not None and all(hasattr(t, '_hvr_split') for x in xs for t in x):
            outs = self._runner.detect([[t._hvr_split for t in x] for x in xs], img_meta, rescale)
            return [[bbox2result(d, l, self.bbox_head.num_classes) for d, l in out] for out in outs]
        c4 = self._window_split([t for x in xs for t in x])
        return self._detect(c4, img_meta, V, T, self.key_dim, rescale)

    def simple_test(self, img, img_meta, proposals=None, rescale=False):
        pass                                                             # hnmb_rcnn.py:615-616


@DETECTORS.register_module
class HNMBRCNN(_WindowRCNN):
    """mmdet/models/detectors/hnmb_rcnn.py:16-48,195-222,571-613."""

    def __init__(self, backbone, rpn_head, bbox_roi_extractor, bbox_head, train_cfg, test_cfg, neck=None,
                 shared_head=None, pretrained=None, loss_frames=1):
        super().__init__(backbone=backbone, neck=neck, shared_head=shared_head, rpn_head=rpn_head,
                         bbox_roi_extractor=bbox_roi_extractor, bbox_head=bbox_head, train_cfg=train_cfg,
                         test_cfg=test_cfg, pretrained=pretrained)
        if self.train_cfg is not None:
            self.key_dim = int(self.train_cfg.rcnn.key_dim)
        else:
            self.key_dim = int(self.test_cfg.bbox_head.key_dim)
            self.bbox_head.t_dim = int(test_cfg.bbox_head.t_dim)
            self.bbox_head.sampler_num = int(test_cfg.bbox_head.sampler_num)
            self.bbox_head.nongt_dim = self.bbox_head.t_dim * self.bbox_head.sampler_num

    def _head(self, rows, cur_range, support):
        return self.bbox_head.forward_test(rows, cur_range, key_dim=self.key_dim, all_res=False, support=support)


    def forward_feat_intervideo(self, xs, img_meta, n_support=4, rescale=False, group=None, return_aux=False,
                                proposals=None, support_select='ring'):
        """BASELINE.json configs 4-5 (oracle-defined, parity unpinned by the
        reference): V local key frames, one window each; stage 4 of every key frame also attends
        to the post-fc_new_4 key rows of `n_support` other key frames (ring order over all ranks,
        intervideo.support_indices) gathered with ONE all-gather.  Frames may yield fewer than max_num
        proposals: every key frame travels as a fixed block of rows plus its count, and the receivers mask
        the support keys with the counts (window.inter_stage_a/b/c).
        support_select='similarity' (next row N4): the supports of a key frame are the n_support videos
        whose descriptor - max over the window's frames of the spatially averaged C5 map, as
        get_triplet_patches builds it (hnmb_rcnn.py:76-101) - is most similar to its own; the descriptors
        travel inside the same all-gather."""
        import torch.distributed as dist
        from . import intervideo, window
        if support_select not in ('ring', 'similarity'):
            raise ValueError('support_select must be "ring" or "similarity", got %r' % (support_select,))
        V, T = len(xs), len(xs[0])
        assert all(len(x) == T for x in xs)
        if (self._runner is not None and proposals is None and not return_aux and support_select == 'ring'
                and all(hasattr(t, '_hvr_split') for x in xs for t in x)):
            outs = self._runner.detect_inter([[t._hvr_split for t in x] for x in xs], img_meta, rescale, n_support, group)
            return [[bbox2result(d, l, self.bbox_head.num_classes) for d, l in out] for out in outs]
        on = dist.is_available() and dist.is_initialized()
        world, rank = (dist.get_world_size(group), dist.get_rank(group)) if on else (1, 0)
        c4 = self._window_split([t for x in xs for t in x])
        flat = None if proposals is None else [p for pv in proposals for p in pv]
        st = window.inter_stage_a(self, c4, img_meta, V, T, self.key_dim, n_support, support_select, proposals=flat)
        recv, _ = intervideo.exchange(st.send, group)
        window.inter_stage_b(self, st, img_meta, rescale)
        window.inter_stage_c(self, st, recv, img_meta, rescale, world, rank)
        counts, per_video = st.result.parse(st.result.buf.cpu())
        nc = self.bbox_head.num_classes
        results = [[bbox2result(d, l, nc) for d, l in pv] for pv in per_video]
        if not return_aux:
            return results
        P, S = st.fs.P, n_support
        kc = [counts[v * T + self.key_dim] for v in range(V)]
        aux = []
        for v in range(V):
            n = kc[v]
            c1, r1 = self.bbox_head._split_out(st.out1[v * P:v * P + n])
            c2, r2 = self.bbox_head._split_out(st.out2[v * P:v * P + n])
            sup = ops.Split(st.sup_rows.hi[v * S * P:(v + 1) * S * P], st.sup_rows.lo[v * S * P:(v + 1) * S * P])
            aux.append(dict(cls=[c1, c2], reg=[r1, r2], support=sup, selected=st.sel[v].tolist(),
                            support_counts=st.fs.seg[v, T:].tolist()))
        return results, aux


@DETECTORS.register_module
class SelsaRCNN(_WindowRCNN):
    """mmdet/models/detectors/selsa_rcnn.py:18-83,281-317."""

    def __init__(self, backbone, rpn_head, bbox_roi_extractor, bbox_head, train_cfg, test_cfg, neck=None,
                 shared_head=None, pretrained=None):
        super().__init__(backbone=backbone, neck=neck, shared_head=shared_head, rpn_head=rpn_head,
                         bbox_roi_extractor=bbox_roi_extractor, bbox_head=bbox_head, train_cfg=train_cfg,
                         test_cfg=test_cfg, pretrained=pretrained)
        if self.train_cfg is not None:
            self.key_dim = int(self.train_cfg.rcnn.key_dim)
        else:
            self.key_dim = int(self.test_cfg.relation_setup.frame_interval)   # selsa_rcnn.py:40
            if 'bbox_head' in self.test_cfg:
                self.bbox_head.t_dim = int(test_cfg.bbox_head.t_dim)
                self.bbox_head.sampler_num = int(test_cfg.bbox_head.sampler_num)
                self.bbox_head.nongt_dim = self.bbox_head.t_dim * self.bbox_head.sampler_num

    def _head(self, rows, cur_range, support):
        cls, reg, _ = self.bbox_head(rows, cur_range, key_dim=self.key_dim)   # 3-tuple (repair 2)
        return [cls], [reg]
