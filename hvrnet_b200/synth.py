"""Seeded synthetic weights and frames for the hot path (no checkpoint or data ships
with the reference; configs/faster_rcnn_r101_hrnmp_c5.py:356-359 points at
author-local files).  Parameter names are the reference's, so a real mmdet-format
checkpoint could be loaded instead (checkpoint row).

Recipe: convs Kaiming-normal (fan_out), frozen BN with
non-trivial affine and running statistics, RPN / fc layers N(0, 0.01) as the
reference initialises them (rpn_head.py:25-28, hrnmp_bbox_head.py:192-214), except
that the relation q/k projections are scaled up so the attention logits have a
standard deviation around 2 - otherwise softmax is uniform and a test could not see
an attention bug.
"""
import math

import torch

R101_BLOCKS = (3, 4, 23, 3)
IMG_MEAN = (103.06, 115.90, 123.15)     # hrnmp cfg:166-167 (BGR, std 1, no RGB swap)


def _conv(g, cout, cin, k, gain=1.0):
    std = gain * math.sqrt(2.0 / (cout * k * k))
    return torch.randn(cout, cin, k, k, generator=g) * std


def _bn(g, sd, name, c, wlo=0.5, whi=1.5):
    sd[name + '.weight'] = torch.rand(c, generator=g) * (whi - wlo) + wlo
    sd[name + '.bias'] = torch.randn(c, generator=g) * 0.1
    sd[name + '.running_mean'] = torch.randn(c, generator=g) * 0.1
    sd[name + '.running_var'] = torch.rand(c, generator=g) + 0.5


def _res_layer(g, sd, p, inplanes, planes, blocks):
    for i in range(blocks):
        q = '%s%d.' % (p, i)
        cin = inplanes if i == 0 else planes * 4
        sd[q + 'conv1.weight'] = _conv(g, planes, cin, 1)
        _bn(g, sd, q + 'bn1', planes)
        sd[q + 'conv2.weight'] = _conv(g, planes, planes, 3)
        _bn(g, sd, q + 'bn2', planes)
        sd[q + 'conv3.weight'] = _conv(g, planes * 4, planes, 1)
        # a damped last BN keeps 33 stacked residual blocks in a sane numeric range
        _bn(g, sd, q + 'bn3', planes * 4, 0.15, 0.35)
        if i == 0:
            sd[q + 'downsample.0.weight'] = _conv(g, planes * 4, cin, 1)
            _bn(g, sd, q + 'downsample.1', planes * 4, 0.5, 1.0)


def _linear(g, sd, name, cout, cin, std=0.01, bias_std=0.0):
    sd[name + '.weight'] = torch.randn(cout, cin, generator=g) * std
    sd[name + '.bias'] = torch.randn(cout, generator=g) * bias_std


def make_state_dict(head='hrnmp', seed=0, num_classes=31, fc_dim=1024, roi_channels=256,
                    roi_size=7, qk_std=None, trunk=True):
    """head in {'hrnmp', 'selsa', 'shared_fc'}.  Returns a flat fp32 CPU state_dict."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    if trunk:
        sd['backbone.conv1.weight'] = _conv(g, 64, 3, 7, gain=0.04)   # inputs are +-128
        _bn(g, sd, 'backbone.bn1', 64)
        inpl = 64
        for s in range(3):
            _res_layer(g, sd, 'backbone.layer%d.' % (s + 1), inpl, 64 * 2 ** s, R101_BLOCKS[s])
            inpl = 64 * 2 ** s * 4
        _res_layer(g, sd, 'shared_head.layer4.', 1024, 512, R101_BLOCKS[3])
        sd['shared_head.new_layer_1.conv.weight'] = _conv(g, roi_channels, 2048, 1, gain=0.25)
        sd['shared_head.new_layer_1.conv.bias'] = torch.randn(roi_channels, generator=g) * 0.1
        sd['rpn_head.rpn_conv.weight'] = torch.randn(512, 1024, 3, 3, generator=g) * 0.004
        sd['rpn_head.rpn_conv.bias'] = torch.randn(512, generator=g) * 0.01
        sd['rpn_head.rpn_cls.weight'] = torch.randn(12, 512, 1, 1, generator=g) * 0.12
        sd['rpn_head.rpn_cls.bias'] = torch.randn(12, generator=g) * 0.01
        sd['rpn_head.rpn_reg.weight'] = torch.randn(48, 512, 1, 1, generator=g) * 0.012
        sd['rpn_head.rpn_reg.bias'] = torch.randn(48, generator=g) * 0.01
    feat_dim = roi_channels * roi_size * roi_size
    p = 'bbox_head.'
    if head == 'shared_fc':
        _linear(g, sd, p + 'shared_fcs.0', fc_dim, feat_dim, 0.01, 0.01)
        _linear(g, sd, p + 'shared_fcs.1', fc_dim, fc_dim, 0.03, 0.01)
        _linear(g, sd, p + 'fc_cls', num_classes, fc_dim, 0.05, 0.01)
        _linear(g, sd, p + 'fc_reg', 4, fc_dim, 0.01, 0.01)
        return sd
    stages = 4 if head == 'hrnmp' else 2
    qk = (2.6 / math.sqrt(fc_dim)) if qk_std is None else qk_std
    for k in range(1, stages + 1):
        _linear(g, sd, p + 'fc_new_%d' % k, fc_dim, feat_dim if k == 1 else fc_dim,
                0.01 if k == 1 else 0.03, 0.01)
        s = p + 'selsa_%d.' % k
        _linear(g, sd, s + 'q_data_fc_%d' % k, fc_dim, fc_dim, qk, 0.01)
        _linear(g, sd, s + 'k_data_fc_%d' % k, fc_dim, fc_dim, qk, 0.01)
        sd[s + 'linear_out_%d.weight' % k] = torch.randn(fc_dim, fc_dim, 1, 1, generator=g) * 0.03
        sd[s + 'linear_out_%d.bias' % k] = torch.randn(fc_dim, generator=g) * 0.01
    _linear(g, sd, p + 'fc_cls', num_classes, fc_dim, 0.05, 0.01)
    _linear(g, sd, p + 'fc_reg', 4, fc_dim, 0.01, 0.01)
    if head == 'hrnmp':
        _linear(g, sd, p + 'fc_cls_2', num_classes, fc_dim, 0.05, 0.01)
        _linear(g, sd, p + 'fc_reg_2', 4, fc_dim, 0.01, 0.01)
    return sd


def make_frames(n, seed=0, h=600, w=1000, pad_to=16, noise=4.0, device='cpu'):
    """n frames of one synthetic video: frame 0 is U[0,255) BGR, later frames are
    frame 0 + N(0, noise) (keeps proposals overlapping across the window); mean
    subtracted (std 1), then zero-padded bottom/right to a multiple of 16
    (pipelines/transforms.py:240-322 order: Normalize before Pad)."""
    g = torch.Generator().manual_seed(1000 + seed)
    base = torch.rand(3, h, w, generator=g) * 255.0
    frames = [base]
    for _ in range(1, n):
        frames.append(base + torch.randn(3, h, w, generator=g) * noise)
    x = torch.stack(frames) - torch.tensor(IMG_MEAN).view(1, 3, 1, 1)
    hp = (h + pad_to - 1) // pad_to * pad_to
    wp = (w + pad_to - 1) // pad_to * pad_to
    out = torch.zeros(n, 3, hp, wp)
    out[:, :, :h, :w] = x
    return out.to(device)


def make_img_meta(h=600, w=1000, pad_to=16, scale_factor=1.0):
    hp = (h + pad_to - 1) // pad_to * pad_to
    wp = (w + pad_to - 1) // pad_to * pad_to
    return dict(img_shape=(h, w, 3), pad_shape=(hp, wp, 3), scale_factor=scale_factor, flip=False)
