"""Sliding-window video loop (R15): the control flow of ``multi_hnl_gpu_test``
(tools/hnl_test.py:359-463, pre_padding_imgs :293-307) for one video, restated around the
registered detectors.

  first frame   window = [(W-1)/2 randomly chosen frames of the same video] + [frame 0]
                (np.random.shuffle + np.random.choice, hnl_test.py:294-296 - same RNG calls)
  middle frames append; once the window holds W maps every new frame yields the detection of
                the window's CENTRE frame (offset list index (W-1)/2)
  last frame    appended (W+1)/2 times (clipped to the video length), one detection per repeat;
                short videos are topped up with more random frames first

With ``video_shuffle=True`` (the hrnmp config's relation_setup, cfg:156; the only mode the reference's
loop files results for) the frames of a video ARRIVE in a np.random.shuffle order
(imagenet_vid_sequence.py:203-210), drawn before the loop's own padding draws, so a window holds
frames from all over the video - SELSA's global aggregation.  The default here is temporal order.

``window_schedule`` is the pure-Python schedule (frame indices only) - the CPU tests compare it
with the oracle's restatement and with a trace of the reference's own loop
(tests/golden/ref_loop_golden.pt); ``detect_video`` drives a detector with it.
"""
import numpy as np


def pre_padding_indices(seg_len, num, rng=np.random):
    """hnl_test.py:293-296: which frames of the video pad the window."""
    video_index = np.arange(seg_len).tolist()
    rng.shuffle(video_index)
    return rng.choice(video_index, num, replace=num > seg_len).tolist()


def window_schedule(seg_len, window, rng=np.random, video_shuffle=False):
    """Yields (window_frame_indices, window_offsets, key_offset) for every detection of a video
    of `seg_len` frames, in the order the reference produces them.  Offsets are -1 for padding
    frames (their detections are never emitted)."""
    half = int((window - 1) / 2)
    frames, offs = [], []
    order = np.arange(seg_len).tolist()                      # arrival order of the frame offsets
    if video_shuffle:
        rng.shuffle(order)                                   # imagenet_vid_sequence.py:203-205

    def push(f, o):
        frames.append(f)
        offs.append(o)
        if len(frames) > window:
            frames.pop(0)
            offs.pop(0)

    for t in range(seg_len):
        last = t == seg_len - 1
        if t == 0:                                           # key_frame_flag == 0
            frames, offs = [], []
            pad = pre_padding_indices(seg_len, half, rng)
            for f in pad:
                push(f, -1)
            push(order[0], order[0])
            if not last:
                continue
        if not last:                                         # key_frame_flag == 2
            full = len(frames) >= window - 1
            push(order[t], order[t])
            if full:
                yield list(frames), list(offs), offs[half]
            continue
        # key_frame_flag == 1 (also the only frame of a 1-frame video, which the reference flags 0 then ends)
        if t == 0:
            frames.pop()
            offs.pop()
        end_counter = 0
        while end_counter < min(seg_len, int((window + 1) / 2)):
            push(order[t], order[t])
            end_counter += 1
            if len(frames) < window - 1:
                for f in pre_padding_indices(seg_len, window - len(frames), rng):
                    push(f, -1)
            yield list(frames), list(offs), offs[half]


def detect_video(model, frames, img_meta, window=None, rng=np.random, rescale=True, video_shuffle=False):
    """frames: sequence of preprocessed [1,3,H,W] CUDA tensors of ONE video.  Returns
    {frame_offset: result} with one entry per emitted key frame (forward_feat results).

    C4 maps are kept only while a scheduled window can still use them: the schedule is drawn first (it is pure
    index arithmetic on the RNG), the last use of every frame index is looked up, and a map is dropped right
    after it - at most `window` + the current pad frames are alive (the reference's deque holds `window` maps,
    tools/hnl_test.py:359-463), instead of one ~20 MB map per frame of a 2000-frame video.
    Deviation from the reference's loop (documented next to the golden-trace test of window_schedule): windows
    whose centre is a padding frame (offset -1), or that are shorter than `window`, are skipped here; the
    reference still runs forward_feat on them and files the result under offset -1, where the later frames
    overwrite it and the evaluation never reads it (tools/hnl_test.py:421-433)."""
    window = int(window or model.bbox_head.t_dim)
    sched = [(idxs, key) for idxs, offs, key in window_schedule(len(frames), window, rng, video_shuffle)
             if key >= 0 and len(idxs) == window]
    last_use = {}
    for n, (idxs, _) in enumerate(sched):
        for i in idxs:
            last_use[i] = n
    cache = {}

    def feat(i):
        if i not in cache:
            cache[i] = model(img=frames[i], img_meta=[img_meta], backbone_feat=True)[0]
        return cache[i]

    out = {}
    metas = [img_meta] * window
    for n, (idxs, key) in enumerate(sched):
        out[key] = model(x=[feat(i) for i in idxs], img=None, img_meta=metas, forward_feat=True, return_loss=False,
                         rescale=rescale)
        for i in set(idxs):
            if last_use[i] == n:
                del cache[i]
    return out
