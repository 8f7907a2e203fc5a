"""Config dicts of the BASELINE.json workloads, in the reference's config schema
(configs/faster_rcnn_r101_hrnmp_c5.py:37-160, faster_rcnn_r101_selsa_c5.py:16-140), built
programmatically so the GPU box (which has no /root/reference) can construct the detectors.
The reference's own config files load through ``hvrnet_b200.config.Config.fromfile`` too
(tests/test_host.py does that when /root/reference is present)."""
from .config import Config


def model_cfg(net_type='HNMBRCNN', t_dim=15, key_dim=7, nms_pos=300, sampler_num=128, num_classes=31):
    norm_cfg = dict(type='BN', requires_grad=False)
    bbox_type = {'HNMBRCNN': 'HRNMPBBoxHead', 'SelsaRCNN': 'SelsaBBoxHead', 'FasterRCNN': 'SharedFCBBoxHead'}[net_type]
    head = dict(type=bbox_type, with_avg_pool=False, in_channels=256, roi_feat_size=7, num_classes=num_classes,
                target_means=[0., 0., 0., 0.], target_stds=[0.1, 0.1, 0.2, 0.2], reg_class_agnostic=True,
                loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=False, loss_weight=1.0),
                loss_bbox=dict(type='SmoothL1Loss', beta=1.0, loss_weight=1.0))
    if net_type == 'HNMBRCNN':
        head.update(sampler_num=sampler_num, imgs_per_video=3, t_dim=9, fc_feat_dim=1024)
    elif net_type == 'SelsaRCNN':
        head.update(sampler_num=sampler_num, t_dim=3, fc_feat_dim=1024)
    else:
        head.update(num_fcs=2, fc_out_channels=1024)
    model = dict(
        type=net_type,
        backbone=dict(type='ResNet', depth=101, num_stages=3, strides=(1, 2, 2), dilations=(1, 1, 1),
                      out_indices=(2,), frozen_stages=1, style='caffe', norm_eval=True, norm_cfg=norm_cfg),
        shared_head=dict(type='ResLayer', depth=101, stage=3, stride=1, dilation=2, style='caffe', norm_eval=True,
                         norm_cfg=norm_cfg, external_conv=True),
        rpn_head=dict(type='RPNHead', in_channels=1024, feat_channels=512, anchor_scales=[4, 8, 16, 32],
                      anchor_ratios=[0.5, 1.0, 2.0], anchor_strides=[16], target_means=[.0, .0, .0, .0],
                      target_stds=[1.0, 1.0, 1.0, 1.0],
                      loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0),
                      loss_bbox=dict(type='SmoothL1Loss', beta=1.0 / 9.0, loss_weight=1.0)),
        bbox_roi_extractor=dict(type='SingleRoIExtractor', roi_layer=dict(type='RoIAlign', out_size=7, sample_num=2),
                                out_channels=1024, featmap_strides=[16], feat_from_shared_head=True),
        bbox_head=head)
    test_cfg = dict(
        rpn=dict(nms_across_levels=False, nms_pre=6000, nms_post=nms_pos, max_num=nms_pos, nms_thr=0.7,
                 min_bbox_size=0),
        rcnn=dict(score_thr=0.001, nms=dict(type='nms', iou_thr=0.3), max_per_img=300, key_dim=key_dim),
        bbox_head=dict(sampler_num=nms_pos, t_dim=t_dim, key_dim=key_dim),
        relation_setup=dict(shuffle=False, video_shuffle=True, has_rpn=True, frame_interval=key_dim, frame_stride=1))
    return Config(dict(model=model, test_cfg=test_cfg, train_cfg=None))


WORKLOADS = {
    # BASELINE.json configs[0..3]
    'faster_rcnn': dict(net_type='FasterRCNN', t_dim=1, key_dim=0, head='shared_fc'),
    'selsa': dict(net_type='SelsaRCNN', t_dim=3, key_dim=1, head='selsa'),
    'hrnmp': dict(net_type='HNMBRCNN', t_dim=15, key_dim=7, head='hrnmp'),
    'hrnmp_inter': dict(net_type='HNMBRCNN', t_dim=15, key_dim=7, head='hrnmp', support_videos=4),
}


def build_workload(name, device, seed=0):
    """Detector of a BASELINE.json workload with the seeded synthetic weights, on `device`."""
    from . import models, synth  # noqa: F401  (registers the modules)
    from .builder import build_detector
    w = WORKLOADS[name]
    cfg = model_cfg(w['net_type'], w['t_dim'], w['key_dim'])
    model = build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
    sd = synth.make_state_dict(w['head'], seed=seed)
    model.load_state_dict(sd, strict=False)
    return model.to(device), sd, w
