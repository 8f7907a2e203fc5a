"""CPU checks of tests/parity_tools.py (the index-parity instrumentation used by the GPU tests and the parity
report): its traces are the oracle's own results, identical inputs give no divergence, and a perturbation of
the size of the CUDA path's arithmetic error is classified as a near-tie."""
import torch

from oracle import ref_torch as R
from tests import parity_tools as PT


def _maps(seed, h=20, w=24, A=12):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(A, h, w, generator=g) * 2, torch.randn(4 * A, h, w, generator=g) * 0.3


def test_proposal_trace_is_the_oracle():
    cls, reg = _maps(0)
    anchors = PT.anchors_for(20, 24)
    cfg = dict(nms_pre=1500, nms_post=300, max_num=300, nms_thr=0.7)
    t = PT.proposal_trace(cls, reg, anchors, (300, 380), cfg)
    props, idx = R.rpn_proposals_single(cls, reg, anchors, (300, 380), return_aux=True, **cfg)
    assert torch.equal(t['props'], props) and torch.equal(t['anchor'], idx)
    assert PT.explain_proposal_divergence(t, t) is None


def test_proposal_near_tie_classification():
    cls, reg = _maps(1)
    anchors = PT.anchors_for(20, 24)
    cfg = dict(nms_pre=1500, nms_post=300, max_num=300, nms_thr=0.7)
    g = torch.Generator().manual_seed(7)
    found = 0
    for _ in range(20):
        cls2 = cls + torch.randn(cls.shape, generator=g) * 2e-4          # the size of the measured logit error
        reg2 = reg + torch.randn(reg.shape, generator=g) * 2e-5
        to, td = PT.proposal_trace(cls, reg, anchors, (300, 380), cfg), PT.proposal_trace(cls2, reg2, anchors, (300, 380), cfg)
        e = PT.explain_proposal_divergence(to, td)
        if e is not None:
            found += 1
            assert e['near_tie'], e
    # a gross perturbation is NOT explained as a near-tie
    td = PT.proposal_trace(cls.flip(1), reg, anchors, (300, 380), cfg)
    e = PT.explain_proposal_divergence(PT.proposal_trace(cls, reg, anchors, (300, 380), cfg), td)
    assert e is not None and not e['near_tie']


def test_det_trace_is_the_oracle():
    g = torch.Generator().manual_seed(3)
    n = 200
    x1, y1 = torch.rand(n, generator=g) * 700, torch.rand(n, generator=g) * 400
    rois = torch.stack([torch.zeros(n), x1, y1, x1 + 30 + torch.rand(n, generator=g) * 200,
                        y1 + 30 + torch.rand(n, generator=g) * 150], 1)
    cls = torch.randn(n, 31, generator=g) * 2
    reg = torch.randn(n, 4, generator=g) * 0.5
    t = PT.det_trace(rois, cls, reg, (600, 1000), 1.0)
    dets, labels = R.get_det_bboxes(rois, [cls], [reg], (600, 1000), 1.0, True)
    assert torch.equal(t['dets'], dets[0]) and torch.equal(t['labels'], labels[0])
    assert torch.equal(t['boxes'][t['rows']], t['dets'][:, :4])
    assert PT.explain_det_divergence(t, t) is None
    found = 0
    for i in range(20):
        t2 = PT.det_trace(rois, cls + torch.randn(cls.shape, generator=g) * 3e-4, reg, (600, 1000), 1.0)
        e = PT.explain_det_divergence(t, t2)
        if e is not None:
            found += 1
            assert e['near_tie'], e
    t3 = PT.det_trace(rois, cls.flip(0), reg, (600, 1000), 1.0)
    e = PT.explain_det_divergence(t, t3)
    assert e is not None and not e['near_tie']
