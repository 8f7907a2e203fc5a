"""Generates tests/golden/ref_shard_golden.json: how the REFERENCE's own VIDSeqDataset.get_indices
(mmdet/datasets/imagenet_vid_sequence.py:117-158, cut out of the file's syntax tree and run on a bare object)
distributes whole videos over the ranks of its distributed test.  Seeded video lengths; the fixture stores,
per case, the lengths, the world size and the video indices each rank received.

    python tests/golden/make_shard_golden.py
"""
import ast
import json
import os
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/mmdet/datasets/imagenet_vid_sequence.py'


def main():
    tree = ast.parse(open(SRC).read())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == 'VIDSeqDataset')
    node = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == 'get_indices')
    ns = dict(np=np, print=lambda *a, **k: None)
    exec(compile(ast.Module(body=[node], type_ignores=[]), SRC, 'exec'), ns)
    rs = np.random.RandomState(7)
    cases = []
    for world in (1, 2, 3, 4, 8):
        for n_videos in (1, 5, 8, 40):
            seg = [int(x) for x in rs.randint(1, 300, size=n_videos)]
            ds = types.SimpleNamespace(test_mode=True, size=int(sum(seg)),
                                       img_infos=[dict(frame_seg_len=L, frame_id=1 + int(sum(seg[:i]))) for i, L in enumerate(seg)])
            ns['get_indices'](ds, world)
            # per rank: the global video index of every frame it received, de-duplicated in order
            per_rank = []
            for idx in ds.indices_list:
                vids = []
                for i in idx:
                    v = ds.global_video_list[int(i)]
                    if not vids or vids[-1] != v:
                        vids.append(int(v))
                per_rank.append(vids)
            cases.append(dict(seg_lens=seg, world=world, videos=per_rank,
                              frames=[int(x) for x in ds.local_frame_size_list]))
    json.dump(cases, open(os.path.join(HERE, 'ref_shard_golden.json'), 'w'))
    print(len(cases), 'cases; e.g.', cases[9]['world'], cases[9]['frames'])


if __name__ == '__main__':
    main()
