"""CUDA-graph executor of the per-key-frame path.

The eager modules (models.py) issue ~200 kernel launches per key frame from Python; on a
B200 the host cannot enqueue them as fast as the GPU retires them.  ``GraphRunner`` captures
the same calls - unchanged kernels, unchanged order - into two CUDA graphs per input shape:

  trunk graph    image [1,3,H,W] (static buffer) -> C4 split NHWC + the NCHW fp32 copy
  window graph   window of T C4 maps (static buffer) -> C5, RPN, proposals, RoIAlign,
                 relation head, decode + multiclass NMS -> one packed result buffer

so a key frame costs two graph launches and ONE device->host read.  The window graph is
captured for the common case "every frame yields max_num proposals" (row offsets are then
static); the per-frame counts travel back with the result and, if any frame produced fewer
proposals, the frame is recomputed on the eager path with the actual counts - results never
depend on the speculation.
"""
import torch

from . import _lib, ops


class _Captured:
    __slots__ = ('graph', 'fn', 'inputs', 'outputs', 'launches', 'perm', 'perm_host', 'perm_pin', 'ring')


class WindowRing:
    """Host bookkeeping of the window graph's static input buffer [V*T frames].  The reference's driver
    passes, for every key frame, the whole window as a list of per-frame C4 maps of which all but one were
    in the previous window (tools/hnl_test.py:359-463); concatenating them again costs 1 GB of copies per step
    at V = 7, T = 15.  The buffer is therefore a ring per video: a frame that is already held (the same
    tensor object, not written in place since it was copied) stays in its slot and only new frames are
    copied.  C5 / RPN / proposal generation are per-frame and run in slot order; the returned permutation
    (window position -> slot) puts the proposals back into window order and is the frame index RoIAlign
    reads, so everything downstream sees the window exactly as the caller ordered it."""

    def __init__(self, n_videos, n_frames):
        self.V, self.T = n_videos, n_frames
        self.slots = [[None] * n_frames for _ in range(n_videos)]   # slot -> (hi tensor, its version, lo's version)

    def place(self, windows):
        """windows: V lists of T per-frame Splits.  Returns (copies, perm): copies = [(slot index into the
        buffer, Split to copy there)], perm[v*T + t] = slot index holding window position t of video v."""
        T = self.T
        copies, perm = [], []
        for v, w in enumerate(windows):
            sl = self.slots[v]
            held = {id(q[0]): j for j, q in enumerate(sl) if q is not None}   # ids are unique while the refs are held
            where = []
            for p in w:
                j = held.get(id(p.hi))
                if j is not None and (sl[j][1] != p.hi._version or sl[j][2] != p.lo._version):
                    sl[j] = None                                # written in place since it was copied
                    j = None
                where.append(j)
            used = {j for j in where if j is not None}
            for t, p in enumerate(w):
                if where[t] is None:
                    j = next((where[u] for u in range(t) if w[u].hi is p.hi), None)   # frame repeated in the window
                    if j is None:
                        j = next(j for j in range(T) if j not in used)
                        copies.append((v * T + j, p))
                        sl[j] = (p.hi, p.hi._version, p.lo._version)
                        used.add(j)
                    where[t] = j
            perm += [v * T + j for j in where]
        return copies, perm


class GraphRunner:

    def __init__(self, model, capture=True):
        """capture=False: the same two closures (static buffers, forked branches, batched head) are
        re-issued eagerly every step instead of being replayed as graphs - identical launches, so
        bench.py's roofline leg can bracket each one with CUDA events."""
        self.m = model
        self.capture = capture
        self._trunk = {}
        self._window = {}
        self.replayed_launches = 0      # kernels launched through graph replays (bench.py gpu_launches)
        self._copy_stream = None
        self._staged = None             # (the prefetched tensor, trunk already run, ready event)
        self._version = model.weights_version()

    def _check_weights(self):
        """The captured graphs read the packed weight tensors through raw pointers; a load_state_dict / .to() /
        load_checkpoint after the capture frees them (models._Packed).  Drop every capture then: the next call
        re-packs and re-captures."""
        v = self.m.weights_version()
        if v != self._version:
            torch.cuda.synchronize()
            self._trunk.clear()
            self._window.clear()
            self._staged = None
            self._version = v

    # ------------------------------------------------------------------ input prefetch
    def prefetch(self, img, trunk=True):
        """Software pipelining of the NEXT step: on a side stream, copy its frames host->device and
        (trunk=True) replay the trunk graph on them, so both overlap the current step's window
        graph; ``extract`` picks the result up when it is handed the same tensor.  All of the work
        still happens once per step, inside whatever region the caller times."""
        self._check_weights()
        key = tuple(img.shape)
        c = self._trunk.get(key)
        if c is None:
            return                                   # trunk graph not captured yet: first call goes through extract()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        side = self._copy_stream
        side.wait_stream(torch.cuda.current_stream())   # previous consumers of the static buffers are enqueued
        with torch.cuda.stream(side):
            c.inputs.copy_(img, non_blocking=True)
            if trunk:
                self._replay(c)
            ev = torch.cuda.Event()
            ev.record()
        self._staged = (img, trunk, ev)             # holds the tensor: identity cannot be recycled while staged

    # ------------------------------------------------------------------ capture helper
    def _capture(self, fn):
        """Warm up eagerly (lazy weight packing, func attributes, tensor-map cache), then capture."""
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        c = _Captured()
        c.fn = fn
        if not getattr(self, 'capture', True):
            c.graph, c.outputs, c.launches = None, None, 0
            return c
        c.graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        with torch.cuda.graph(c.graph):
            c.outputs = fn()
        c.launches = _lib.launch_count() - l0
        return c

    @staticmethod
    def _replay(c):
        if c.graph is None:
            c.outputs = c.fn()          # eager re-issue (capture=False): counted by _lib.launch_count()
        else:
            c.graph.replay()

    # ------------------------------------------------------------------ trunk
    def extract(self, img):
        self._check_weights()
        key = tuple(img.shape)
        c = self._trunk.get(key)
        if c is None:
            buf = torch.zeros(img.shape, dtype=torch.float32, device='cuda:%d' % torch.cuda.current_device())

            def fn():
                s = self.m.backbone.forward_split(buf)
                return s, ops.nhwc_split_to_nchw(s)
            c = self._capture(fn)
            c.inputs = buf
            self._trunk[key] = c
        staged, self._staged = self._staged, None
        if staged is not None:
            torch.cuda.current_stream().wait_event(staged[2])   # the side stream is done with the static buffers
        if staged is not None and staged[0] is img and tuple(staged[0].shape) == key:
            if not staged[1]:                                   # frames staged by prefetch(), trunk still to run
                self._replay(c)
        else:
            c.inputs.copy_(img, non_blocking=True)      # H2D (pinned host) or D2D into the static buffer
            self._replay(c)
        self.replayed_launches += c.launches
        s, nchw = c.outputs
        out = nchw.clone()                              # the caller keeps C4 maps in its window deque
        out._hvr_split = ops.Split(s.hi.clone(), s.lo.clone())
        return (out,)

    @staticmethod
    def per_frame(c4):
        """Split a batched C4 (B frames) into B single-frame tensors that keep their split view."""
        outs = []
        for i in range(c4.shape[0]):
            t = c4[i:i + 1]
            t._hvr_split = c4._hvr_split[i:i + 1]
            outs.append(t)
        return outs

    # ------------------------------------------------------------------ window(s)
    def detect(self, windows, img_meta, rescale):
        """windows: list of V windows, each a list of T per-frame C4 Splits [1,h,w,C] (V key
        frames of V different videos are batched through one graph: C5 / RPN / proposals /
        RoIAlign run over all V*T frames at once, the relation head once per video).
        Returns, per video, the list of per-output (dets, labels) host tensors, or None for
        the videos whose speculation (every frame yields max_num proposals) failed."""
        self._check_weights()
        m = self.m
        V, T = len(windows), len(windows[0])
        meta = img_meta[0]
        sf = meta['scale_factor']
        sf = float(sf if not hasattr(sf, '__len__') else sf[0])
        key = (V, T, tuple(windows[0][0].shape), tuple(meta['img_shape'][:2]), sf, bool(rescale), m.key_dim)
        c = self._window.get(key)
        P = m.test_cfg.rpn['max_num']
        M = m.test_cfg.rcnn['max_per_img']

        def fill(c):
            """Bring the static window buffer up to date: only the frames WindowRing has not seen in the
            previous calls are copied; `perm` (window position -> ring slot) is a graph input."""
            copies, perm = c.ring.place(windows)
            for slot, p in copies:
                c.inputs.hi[slot].copy_(p.hi[0])
                c.inputs.lo[slot].copy_(p.lo[0])
            if perm != c.perm_host:
                # pinned staging buffer: the previous step's copy has completed (every step ends with a
                # device->host read), so it can be rewritten here
                c.perm_pin.copy_(torch.tensor(perm, dtype=torch.int64))
                c.perm.copy_(c.perm_pin, non_blocking=True)
                c.perm_host = perm

        if c is None:
            dev = windows[0][0].hi.device
            _, h, w, C = windows[0][0].shape
            win = ops.Split.zeros((V * T, h, w, C), dev)
            perm = torch.arange(V * T, device=dev)          # window position -> ring slot (static graph input)

            side = [torch.cuda.Stream(), torch.cuda.Stream()]

            def fn():
                main = torch.cuda.current_stream()
                # RPN maps first; proposal generation (sort + decode + greedy NMS: 15 busy CTAs, latency
                # bound) then runs on a forked branch UNDER the C5 convolutions instead of after them
                maps = m.rpn_head.forward_maps(win)
                side[0].wait_stream(main)
                with torch.cuda.stream(side[0]):
                    props, counts = m.rpn_head.proposals_from_maps(maps, meta['img_shape'], m.test_cfg.rpn)
                c5 = m.shared_head.forward_nhwc(win) if m.feat_from_shared_head else ops.merge(win)
                main.wait_stream(side[0])
                props, counts = props.index_select(0, perm), counts.index_select(0, perm)   # slot -> window order
                fidx = perm.to(torch.float32).view(V * T, 1, 1).expand(V * T, P, 1)
                rois = torch.cat([fidx, props[..., :4]], -1).view(-1, 5).contiguous()
                flat = [counts.float()]
                forked, per_video = set(), []
                s = m.key_dim * P
                N = T * P
                head = m.bbox_head
                if V > 1:
                    # batched head: rows of video v live at [v*Npad, v*Npad + N) (Npad = N rounded up to 64);
                    # the pad rows pool a dummy 1-pixel RoI and never reach a result
                    from . import engine
                    Npad = ops.round_up(N, 64)
                    rois_p = torch.zeros((V, Npad, 5), device=dev)
                    rois_p[:, :N] = rois.view(V, N, 5)
                    rows = m.bbox_roi_extractor.roi_layers[0].forward_nhwc_split(c5, rois_p.view(-1, 5))
                    packed = head.packed(dev)
                    # key-frame rois of every video, batch index 0 as bbox2roi([props_key]) gives them
                    rois_key = rois.view(V, N, 5)[:, s:s + P].reshape(V * P, 5).clone()
                    rois_key[:, 0] = 0

                    def post(j, h):
                        """decode + multiclass NMS of head output j for all V key frames: one launch per
                        stage (hvr_det_postprocess_batched), on a forked branch"""
                        st = side[j % 2]
                        st.wait_stream(main)
                        forked.add(st)
                        cls_, reg_ = head._split_out(h)
                        with torch.cuda.stream(st):
                            return head.get_det_bboxes_batched(rois_key, cls_, reg_, V, meta['img_shape'], sf,
                                                               rescale=rescale, cfg=m.test_cfg.rcnn)
                    if head.kind == 'hrnmp':
                        # the branch output is post-processed under stages 3-4
                        out1, f4, f4T = engine.hrnmp_stage123_batched(packed, rows, V, N, Npad, s, P)
                        b1 = post(0, out1)
                        a4 = engine.relation_batched(packed, 4, f4, f4T, V, N, Npad, q_range=(s, P),
                                                     res=engine._key_rows(f4, V, Npad, s, P))
                        _, out2, _ = engine.lin(a4, packed['out2'], want_split=False, want_f32=True)
                        batched = [b1, post(1, out2)]
                        keep = [maps, props, counts, c5, rois, rois_p, rows, out1, f4, f4T, a4, out2, rois_key, batched]
                    else:
                        out1 = engine.selsa_forward_batched(packed, rows, V, N, Npad, s, P)
                        batched = [post(0, out1)]
                        keep = [maps, props, counts, c5, rois, rois_p, rows, out1, rois_key, batched]
                    for v in range(V):
                        per_video.append([(d[v], l[v], k[v:v + 1]) for d, l, k in batched])
                else:
                    rows = m.bbox_roi_extractor.roi_layers[0].forward_nhwc_split(c5, rois)
                    keep = [maps, props, counts, c5, rois, rows]
                    cls, reg = m._head(rows[0:T * P], [dict(start=s, length=P)], None)
                    rois_key = rois[s:s + P].clone()
                    rois_key[:, 0] = 0
                    # the head outputs are post-processed on parallel branches (tiny, latency-bound kernels)
                    outs = []
                    for j, (c_, r_) in enumerate(zip(cls, reg)):
                        st = side[j % 2]
                        st.wait_stream(main)
                        forked.add(st)
                        with torch.cuda.stream(st):
                            outs.append(m.bbox_head.get_det_bboxes(rois_key, c_, r_, meta['img_shape'], sf,
                                                                   rescale=rescale, cfg=m.test_cfg.rcnn))
                    keep += [cls, reg, rois_key, outs]
                    per_video.append(outs)
                for st in forked:                       # join only the branches that were forked
                    main.wait_stream(st)
                for outs in per_video:
                    for d, l, k in outs:
                        flat += [k.float(), d.reshape(-1), l.float()]
                return torch.cat(flat), keep
            c = _Captured()
            c.inputs, c.perm, c.perm_host = win, perm, list(range(V * T))
            c.perm_pin = torch.empty(V * T, dtype=torch.int64).pin_memory()
            c.ring = WindowRing(V, T)
            fill(c)                                     # real data for the warm-up pass
            cap = self._capture(fn)
            c.graph, c.fn, c.outputs, c.launches = cap.graph, cap.fn, cap.outputs, cap.launches
            self._window[key] = c
        fill(c)
        self._replay(c)
        self.replayed_launches += c.launches
        host = c.outputs[0].cpu()                       # the one device->host read of the step
        n_out = (host.numel() - V * T) // (V * (1 + 6 * M))
        res, o = [], V * T
        for v in range(V):
            ok = bool((host[v * T:(v + 1) * T] == P).all())
            outs = []
            for _ in range(n_out):
                k = int(host[o])
                d = host[o + 1:o + 1 + 5 * M].view(M, 5)[:k]
                l = host[o + 1 + 5 * M:o + 1 + 6 * M][:k].long()
                outs.append((d, l))
                o += 1 + 6 * M
            res.append(outs if ok else None)
        return res


class StreamGraphRunner:
    """CUDA-graph executor of the streaming scheduler (streaming.py) for V
    video streams advanced in lock-step.  Per step: ONE frame graph (trunk, C5, RPN, proposals,
    RoIAlign, fc_new_1 for the V new frames) and ONE head graph (relation stages, decode, NMS for
    the V key frames on the cached rows), one device->host read.  Captured for the common case
    of max_num proposals per frame; a short frame raises (use streaming.StreamingDetector then)."""

    def __init__(self, model, n_videos, window=None):
        from collections import deque
        self.m = model
        self.V = n_videos
        self.T = int(window or model.bbox_head.t_dim)
        self.P = model.test_cfg.rpn['max_num']
        self.cache = [deque(maxlen=self.T) for _ in range(n_videos)]      # per video: (props, f1, f1T)
        self._frame = None
        self._head = None
        self.replayed_launches = 0
        self._cap = GraphRunner._capture

    def _frame_graph(self, img, meta):
        m, V, P = self.m, self.V, self.P
        buf = torch.zeros(img.shape, dtype=torch.float32, device='cuda:%d' % torch.cuda.current_device())
        dev = buf.device

        def fn():
            c4 = m.backbone.forward_split(buf)
            maps = m.rpn_head.forward_maps(c4)
            props, counts = m.rpn_head.proposals_from_maps(maps, meta['img_shape'], m.test_cfg.rpn)
            c5 = m.shared_head.forward_nhwc(c4)
            fidx = torch.arange(V, device=dev, dtype=torch.float32).view(V, 1, 1).expand(V, P, 1)
            rois = torch.cat([fidx, props[..., :4]], -1).view(-1, 5).contiguous()
            rows = m.bbox_roi_extractor.roi_layers[0].forward_nhwc_split(c5, rois)
            f1, f1T = engine_head_fc1(m, rows)
            return props, counts, f1, f1T, (c4, maps, c5, rois, rows)
        c = self._cap(self, fn)
        c.inputs = buf
        return c

    def _head_graph(self, meta, rescale):
        m, V, T, P = self.m, self.V, self.T, self.P
        dev = self._frame.inputs.device
        D = self._frame.outputs[2].shape[1]
        N = T * P
        f1w = [ops.Split.zeros((N, D), dev) for _ in range(V)]
        f1Tw = [ops.Split.zeros((D, ops.round_up(N, 64)), dev) for _ in range(V)]
        rk = [torch.zeros((P, 5), device=dev) for _ in range(V)]
        counts = self._frame.outputs[1]
        sf = meta['scale_factor']
        sf = float(sf if not hasattr(sf, '__len__') else sf[0])
        head = m.bbox_head
        side = [torch.cuda.Stream(), torch.cuda.Stream()]

        def fn():
            main = torch.cuda.current_stream()
            packed = head.packed(dev)
            flat, keep = [counts.float()], []
            forked, per_video = set(), []
            s = m.key_dim * P
            from . import engine
            for v in range(V):
                if head.kind == 'hrnmp':
                    o1, o2, _ = engine.hrnmp_forward_test(packed, None, s, P, f1=f1w[v], f1T=f1Tw[v])
                    outs = [o1, o2]
                else:
                    outs = [engine.selsa_forward(packed, None, s, P, f1=f1w[v], f1T=f1Tw[v])]
                dets, used = [], []
                for j, o in enumerate(outs):
                    cls, reg = head._split_out(o)
                    st = side[j % 2]
                    st.wait_stream(main)
                    used.append(st)
                    with torch.cuda.stream(st):
                        dets.append(head.get_det_bboxes(rk[v], cls, reg, meta['img_shape'], sf, rescale=rescale,
                                                        cfg=m.test_cfg.rcnn))
                forked.update(used)
                keep += [outs, dets]
                per_video.append(dets)
            for st in forked:
                main.wait_stream(st)
            for dets in per_video:
                for d, l, k in dets:
                    flat += [k.float(), d.reshape(-1), l.float()]
            return torch.cat(flat), keep
        c = self._cap(self, fn)
        c.inputs = (f1w, f1Tw, rk)
        return c

    def push(self, img, meta, rescale=True):
        """img [V,3,H,W] (device or pinned host): the next frame of each of the V streams.
        Returns None while the windows fill, then a list of V forward_feat-style results."""
        V, T, P = self.V, self.T, self.P
        if self._frame is None:
            self._frame = self._frame_graph(img, meta)
        fr = self._frame
        fr.inputs.copy_(img, non_blocking=True)
        fr.graph.replay()
        self.replayed_launches += fr.launches
        props, counts, f1, f1T = fr.outputs[:4]
        for v in range(V):
            self.cache[v].append((props[v].clone(), f1.hi[v * P:(v + 1) * P].clone(), f1.lo[v * P:(v + 1) * P].clone(),
                                  f1T.hi[:, v * P:(v + 1) * P].clone(), f1T.lo[:, v * P:(v + 1) * P].clone()))
        if len(self.cache[0]) < T:
            return None
        if self._head is None:
            self._head = self._head_graph(meta, rescale)
        hd = self._head
        f1w, f1Tw, rk = hd.inputs
        N = T * P
        for v in range(V):
            fl = list(self.cache[v])
            torch.cat([f[1] for f in fl], 0, out=f1w[v].hi)
            torch.cat([f[2] for f in fl], 0, out=f1w[v].lo)
            torch.cat([f[3] for f in fl], 1, out=f1Tw[v].hi[:, :N])
            torch.cat([f[4] for f in fl], 1, out=f1Tw[v].lo[:, :N])
            rk[v][:, 1:] = fl[self.m.key_dim][0][:, :4]
        hd.graph.replay()
        self.replayed_launches += hd.launches
        host = hd.outputs[0].cpu()
        if not bool((host[:V] == P).all()):
            raise RuntimeError('a frame produced fewer than %d proposals: use streaming.StreamingDetector' % P)
        M = self.m.test_cfg.rcnn['max_per_img']
        n_out = (host.numel() - V) // (V * (1 + 6 * M))
        from .models import bbox2result
        res, o = [], V
        for v in range(V):
            outs = []
            for _ in range(n_out):
                k = int(host[o])
                d = host[o + 1:o + 1 + 5 * M].view(M, 5)[:k]
                l = host[o + 1 + 5 * M:o + 1 + 6 * M][:k].long()
                outs.append(bbox2result(d, l, self.m.bbox_head.num_classes))
                o += 1 + 6 * M
            res.append(outs)
        return res


def engine_head_fc1(model, rows):
    from . import engine
    return engine.head_fc1(model.bbox_head.packed(rows.hi.device), rows)
