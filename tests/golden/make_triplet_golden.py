"""Generates tests/golden/ref_triplet_golden.pt from the REFERENCE's own video-similarity code
(next row N4 of SURVEY.md 8f: similarity-based support-video selection), run in the build
container where /root/reference exists.

mmdet/models/detectors/hnmb_rcnn.py cannot be imported (mmcv and in-tree breakages),
so the one method needed - HNMBRCNN.get_triplet_patches (:76-101), which touches no other member -
is cut out of the file's syntax tree and compiled on its own with the names it uses
(torch, math, adaptive_avg_pool2d, softmax) bound to the real torch functions.  Inputs are seeded;
the fixture stores inputs and outputs, so the tests need neither /root/reference nor this script.

    python tests/golden/make_triplet_golden.py
"""
import ast
import math
import os

import torch
from torch.nn.functional import adaptive_avg_pool2d, softmax

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/mmdet/models/detectors/hnmb_rcnn.py'


def load_method(name):
    tree = ast.parse(open(SRC).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = dict(torch=torch, math=math, adaptive_avg_pool2d=adaptive_avg_pool2d, softmax=softmax)
            exec(compile(mod, SRC, 'exec'), ns)
            return ns[name]
    raise KeyError(name)


def main():
    fn = load_method('get_triplet_patches')
    cases = []
    for seed in range(18):
        g = torch.Generator().manual_seed(1000 + seed)
        video_per_cls = 2 + seed % 3                 # 2..4 videos of the key class
        extra = 1 + seed % 5                         # 1..5 videos of other classes
        imgs = 2 + seed % 3 if seed % 4 else 3       # frames per video (3 in the hrnmp config; 1 breaks the reference's squeeze())
        C, h, w = (256, 2, 3) if seed % 6 == 0 else (32, 3, 4)
        # post-ReLU-like maps with a per-video channel signature so that similarities differ
        feats = []
        for _ in range(video_per_cls + extra):
            sig = torch.rand(1, C, 1, 1, generator=g) * 6
            feats.append([(torch.randn(imgs, C, h, w, generator=g) + sig).clamp(min=0)])
        ids = fn(None, feats, key_video=0, imgs_per_video=imgs, extra_cls=extra, video_per_cls=video_per_cls)
        cases.append(dict(c5=[f[0] for f in feats], video_per_cls=video_per_cls, ids=[int(i) for i in ids]))
    torch.save(cases, os.path.join(HERE, 'ref_triplet_golden.pt'))
    print('wrote', len(cases), 'cases;', [c['ids'] for c in cases])


if __name__ == '__main__':
    main()
