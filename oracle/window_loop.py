"""Oracle restatement (TEST INFRASTRUCTURE ONLY) of the reference's sliding-window test loop,
tools/hnl_test.py:359-463, as a trace: the detector calls are replaced by recording which
frames were in the deque and which frame offset the result was filed under.  The data loader
is replaced by its observable behaviour: frames arrive in order with key_frame_flag
0 (first), 2 (middle), 1 (last) (mmdet/datasets/imagenet_vid_sequence.py:192-243)."""
from collections import deque

import numpy as np


def trace(seg_len, all_frame_interval, rng=np.random, video_shuffle=False):
    """video_shuffle (hrnmp cfg:156): the data set hands the frames of a video out in a shuffled order drawn
    when its first frame is fetched (imagenet_vid_sequence.py:203-210), i.e. before the loop's padding draws.
    Pinned against a trace of the reference's own loop, tests/golden/ref_loop_golden.pt."""
    events = []
    feat_list = frame_offset_list = None
    video_index = np.arange(seg_len).tolist()
    if video_shuffle:
        rng.shuffle(video_index)

    def pre_padding_imgs(num):                                             # :293-296
        video_index = np.arange(seg_len).tolist()
        rng.shuffle(video_index)
        return rng.choice(video_index, num, replace=num > seg_len).tolist()

    for tid in range(seg_len):
        frame_offset = video_index[tid]
        if seg_len == 1:
            flags = [0, 1]                    # a single frame is both the first and the last of its video
        else:
            flags = [0] if tid == 0 else ([1] if tid == seg_len - 1 else [2])
        for key_frame_flag in flags:
            if key_frame_flag == 0:                                            # :359-383
                feat_list = deque(maxlen=all_frame_interval)
                frame_offset_list = deque(maxlen=all_frame_interval)
                pre = pre_padding_imgs(int((all_frame_interval - 1) / 2))
                feat_list.extend(pre)
                feat_list.append(frame_offset)
                frame_offset_list.extend([-1] * len(pre))
                frame_offset_list.append(frame_offset)
                if seg_len == 1:                 # the lone frame is re-appended by the flag-1 branch below
                    feat_list.pop()
                    frame_offset_list.pop()
            elif key_frame_flag == 2:                                          # :384-421
                if len(feat_list) < all_frame_interval - 1:
                    feat_list.append(frame_offset)
                    frame_offset_list.append(frame_offset)
                else:
                    feat_list.append(frame_offset)
                    frame_offset_list.append(frame_offset)
                    events.append((list(feat_list), list(frame_offset_list),
                                   frame_offset_list[int((all_frame_interval - 1) / 2)]))
            else:                                                              # :422-463
                end_counter = 0
                while end_counter < min(seg_len, int((all_frame_interval + 1) / 2)):
                    feat_list.append(frame_offset)
                    frame_offset_list.append(frame_offset)
                    end_counter += 1
                    if len(feat_list) < all_frame_interval - 1:
                        pre = pre_padding_imgs(all_frame_interval - len(feat_list))
                        feat_list.extend(pre)
                        frame_offset_list.extend([-1] * len(pre))
                    events.append((list(feat_list), list(frame_offset_list),
                                   frame_offset_list[int((all_frame_interval - 1) / 2)]))
    return events
