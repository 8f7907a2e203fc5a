"""Generates tests/golden/ref_loop_golden.pt: a trace of the REFERENCE's own sliding-window test loop (R15), run in
the build container where /root/reference exists.

Executed unmodified, cut out of the files' syntax trees:
  tools/hnl_test.py                         multi_hnl_gpu_test (:309-475), pre_padding_imgs (:293-307)
  mmdet/datasets/imagenet_vid_sequence.py   VIDSeqDataset.prepare_test_img (:192-243), __getitem__ (:280-293;
                                            test mode) - the code that sets key_frame_flag / frame_offset and,
                                            with video_shuffle=True (hrnmp cfg:156), visits the frames of a video
                                            in a np.random.shuffle order

around stand-ins that only RECORD: the model returns, for backbone_feat=True, a [n,1] tensor holding the frame
offsets it was given and, for forward_feat=True, the list of offsets in the window; the dataset's image
pipeline returns a one-element tensor holding the frame offset; mmcv.ProgressBar / collate / get_dist_info
and collect_selsa_results_cpu are trivial.  `np.int` (removed from numpy) is bound to int.  The trace is, per
global frame id, the window (frame offsets, padding frames included) whose detection the loop filed there.

    python tests/golden/make_loop_golden.py
"""
import ast
import os
import sys
import time
import types
from collections import deque

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'


def cut(path, names, ns, cls_name=None):
    src = os.path.join(REF, path)
    tree = ast.parse(open(src).read())
    body = tree.body if cls_name is None else next(
        n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == cls_name).body
    out = {}
    for node in body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), src, 'exec'), ns)
            out[node.name] = ns[node.name]
    assert set(out) == set(names), set(names) - set(out)
    return out


class Box:                                  # stands in for mmcv's DataContainer: just a .data attribute
    def __init__(self, data):
        self.data = data


def collate(items, samples_per_gpu=1):
    return dict(img=[torch.cat([it['img'] for it in items], 0)],
                img_meta=Box([[it['img_meta'].data for it in items]]))


def run(seg_lens, window, video_shuffle, seed):
    np_shim = types.SimpleNamespace(**{k: getattr(np, k) for k in ('zeros', 'max', 'arange', 'random', 'sum')}, int=int)
    ns = dict(np=np_shim, deque=deque, time=time, torch=torch, collate=collate,
              get_dist_info=lambda: (0, 1),
              mmcv=types.SimpleNamespace(ProgressBar=lambda n: types.SimpleNamespace(update=lambda: None)),
              collect_selsa_results_cpu=lambda part, size, tmpdir=None: part)
    fns = cut('tools/hnl_test.py', ['multi_hnl_gpu_test', 'pre_padding_imgs'], ns)
    ds_fns = cut('mmdet/datasets/imagenet_vid_sequence.py', ['prepare_test_img', '__getitem__'], dict(np=np),
                 cls_name='VIDSeqDataset')
    n_frames = int(sum(seg_lens))
    starts = [1 + int(sum(seg_lens[:v])) for v in range(len(seg_lens))]
    ds = types.SimpleNamespace(
        test_mode=True, video_shuffle=video_shuffle, proposals=None, cur_tid=0, cur_seg_len=0, key_frame_flag=-1,
        img_infos=[dict(frame_seg_len=int(L), frame_id=starts[v]) for v, L in enumerate(seg_lens)],
        global_video_list=[v for v, L in enumerate(seg_lens) for _ in range(L)],
        local_frame_size_list=[n_frames], local_video_list=[[v for v, L in enumerate(seg_lens) for _ in range(L)]],
        global_video_size_list=[len(seg_lens)],
        make_img_info_anno_info=lambda video, offsets: ([dict(offset=int(o)) for o in offsets], None, None),
        pre_pipeline=lambda results: None,
        pipeline=lambda results: dict(img=torch.tensor([[float(results['img_info']['offset'])]]), img_meta=Box({})))
    for k, f in ds_fns.items():
        setattr(ds, k, types.MethodType(f, ds))

    class Loader:                           # DataLoader(batch_size=1, num_workers=0, shuffle=False) + collate
        dataset = ds

        def __len__(self):
            return n_frames

        def __iter__(self):
            for i in range(n_frames):
                yield collate([ds.__getitem__(i)], 1)

    calls = []

    def model(**kw):
        if kw.get('backbone_feat'):
            return (kw['img'][0].clone(),)
        assert kw.get('forward_feat') and kw['img'] is None and kw['return_loss'] is False
        rec = dict(window=[int(t.item()) for t in kw['x']], n_meta=len(kw['img_meta']))
        calls.append(rec)
        return rec
    model.eval = lambda: None
    np.random.seed(seed)
    # len(dataset) is used for the progress bar and the final slice
    results = fns['multi_hnl_gpu_test'](model, _LenProxy(Loader(), n_frames), window, tmpdir='unused')
    return dict(seg_lens=list(map(int, seg_lens)), window=window, video_shuffle=video_shuffle, seed=seed,
                filed=[None if r is None else r['window'] for r in results], n_calls=len(calls),
                calls=[c['window'] for c in calls])


class _LenProxy:
    """data_loader whose .dataset supports len(): the loop calls len(dataset) and len(data_loader)."""

    def __init__(self, loader, n):
        self._loader, self._n = loader, n
        ds = loader.dataset

        class DS:
            def __len__(s):
                return n

            def __getattr__(s, k):
                return getattr(ds, k)

            def __setattr__(s, k, v):
                setattr(ds, k, v)
        self.dataset = DS()

    def __len__(self):
        return self._n

    def __iter__(self):
        return iter(self._loader)


def main():
    cases = []
    for seed, (seg_lens, window, shuffle) in enumerate([
            ([20], 15, False), ([20], 15, True), ([9, 31, 16], 15, True), ([40, 8], 21, True), ([6, 12], 5, True),
            ([17, 3, 25], 3, True), ([30], 15, True), ([16, 16], 15, False)]):
        cases.append(run(seg_lens, window, shuffle, 500 + seed))
        c = cases[-1]
        print(c['seg_lens'], c['window'], c['video_shuffle'], 'calls', c['n_calls'], 'filed', sum(f is not None for f in c['filed']))
    torch.save(cases, os.path.join(HERE, 'ref_loop_golden.pt'))


if __name__ == '__main__':
    main()
