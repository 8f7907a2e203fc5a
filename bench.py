#!/usr/bin/env python
"""Benchmark of the HVRNet per-key-frame inference hot path (BASELINE.json metric:
VID key frames / sec at 1000x600, 300 proposals per frame).

    python bench.py --gpus N --steps K --warmup W [--workload hrnmp] [--impl reference]

One "step" = one key frame exactly as the reference executes it (tools/hnl_test.py:384-421):
the new frame goes through the R101 trunk (HNMBRCNN.forward(backbone_feat=True)), its C4 map
enters the window deque, and forward_feat runs C5 + RPN + proposals + RoIAlign for all T
frames of the window, the relation head and box decode + multiclass NMS for the key frame
(2113 GFLOP at T=15).  No per-frame caching across windows.

value  = key frames / s with the frame already resident in HBM (CUDA events, max over ranks)
e2e    = the same through the reference call surface with the frame in pinned HOST memory:
         H2D of the frame and the D2H of the detections are inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONV_GFLOP_PER_NEW_FRAME = 166.99          # trunk, BASELINE.md section 3
CONV_GFLOP_PER_WINDOW_FRAME = 74.05 + 22.74  # C5 + RPN




_REAL_STDOUT = None


def _stdout_to_stderr():
    """Point file descriptor 1 at stderr for the rest of the process and keep the real stdout aside: native
    libraries (NCCL's version banner) write to fd 1 directly, and the driver reads ONE JSON line from stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def _emit(text):
    out = _REAL_STDOUT or sys.stdout
    out.write(text + '\n')
    out.flush()


def _measured_peak(peaks, must, prefer, lo, hi):
    """Pick a figure from the driver-written MEASURED_PEAKS.json (schema not under our control): the first
    numeric entry whose (nested) key contains every word of `must` and whose value lies in [lo, hi] - GB/s
    or TFLOP/s - preferring keys that contain a word of `prefer`.  (None, None) if there is none."""
    found = []

    def walk(d, prefix):
        if isinstance(d, dict):
            for k, v in d.items():
                walk(v, prefix + '.' + str(k) if prefix else str(k))
        elif isinstance(d, (int, float)) and not isinstance(d, bool):
            found.append((prefix, float(d)))
    walk(peaks or {}, '')
    ok = [(k, v) for k, v in found if all(m in k.lower() for m in must) and lo <= v <= hi]
    if not ok:
        return None, None
    ok.sort(key=lambda kv: 0 if any(p_ in kv[0].lower() for p_ in prefer) else 1)
    return ok[0]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference', 'torch_cuda'],
                    help="reference: the oracle port on the host cores; torch_cuda: the same port's dense stages on cuDNN / cuBLAS "
                         "(SURVEY.md 8d: the existing-GPU-implementation bar; not run by the driver)")
    ap.add_argument('--workload', default='hrnmp', choices=['hrnmp', 'selsa', 'faster_rcnn', 'hrnmp_inter'])
    ap.add_argument('--support-select', default='ring', choices=['ring', 'similarity'],
                    help='hrnmp_inter: inter-video supports by ring order (BASELINE.json config 5) or by video-descriptor similarity (SURVEY.md 8f N4)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--videos-per-gpu', type=int, default=0, help='key frames (of different videos) batched per step (default 7: 133 of 148 SMs busy in the trunk)')
    ap.add_argument('--no-streaming', action='store_true', help='skip the extra (labelled) streaming-scheduler figure')
    ap.add_argument('--no-other-workloads', action='store_true', help='skip the short side measurements of the other BASELINE.json configs')
    ap.add_argument('--eager', action='store_true', help='disable CUDA graphs (per-kernel Python launches)')
    ap.add_argument('--no-inter-video', action='store_true', help='N > 1: skip the config-5 inter-video block')
    ap.add_argument('--inter-keys-per-gpu', type=int, default=32, help='key frames per rank of the config-5 block')
    ap.add_argument('--gemm-report', default=None, help='write a per-shape table of the igemm launches (csv)')
    return ap.parse_args()


def workload_name(w, T):
    return {'hrnmp': 'HVRNet intra-video (faster_rcnn_r101_hrnmp_c5), 1 key + %d ref frames, 300 proposals' % (T - 1),
            'selsa': 'SELSA R101 (faster_rcnn_r101_selsa_c5), 1 key + %d ref frames, 300 proposals' % (T - 1),
            'hrnmp_inter': 'HVRNet intra+inter-video, 1 key + %d ref frames + 4 inter-video support videos, '
                           '300 proposals each' % (T - 1),
            'faster_rcnn': 'Faster-RCNN R101-C5, single 600x1000 frame'}[w]


def config_dict(workload, T, V, world, launch):
    return {'workload': workload_name(workload, T), 'frames_per_window': T, 'proposals_per_frame': 300,
            'videos_per_gpu': V, 'key_frames_per_step': V, 'input': '%dx3x608x1008 fp32 per step' % V,
            'l2': 'working set (305 MB split weights + >1 GB activations per step) exceeds the 126 MB L2; '
                  'no explicit flush',
            'parallelism': 'replicas over videos, dp%d' % world, 'launch': launch}


# ----------------------------------------------------------------------------------------
# CPU reference leg (oracle): rank 0 only
# ----------------------------------------------------------------------------------------
def cpu_key_frame_seconds(workload, max_seconds=200.0, steps=1, warm=True):
    """Times the oracle (oracle/ref_torch.py, the CPU restatement of the reference's PyTorch
    path + oracle/c RoIAlign) on the same workload: one step = trunk on the new frame +
    forward_feat over the T-frame window.  Returns (seconds per key frame, steps executed)."""
    from hvrnet_b200 import configs, synth
    from oracle import cref, ref_torch as R
    torch.set_num_threads(os.cpu_count())
    w = configs.WORKLOADS[workload]
    T, key = w['t_dim'], w['key_dim']
    # hrnmp_inter: the CPU arm times the intra-video key frame; the 1200 extra stage-4 key rows add
    # 4 GFLOP to ~2100, i.e. < 0.2 %
    sd = synth.make_state_dict(w['head'])
    frames = synth.make_frames(2, seed=0)
    metas = [synth.make_img_meta() for _ in range(T)]
    with torch.no_grad():
        if warm:   # page in MKL/oneDNN and build the C oracle: one trunk pass on a quarter-size frame
            R.trunk_forward(sd, frames[:1, :, :304, :504])
        c4 = R.trunk_forward(sd, frames[:1])
        window = [c4 + 0.01 * i for i in range(T)]
        done, t0 = 0, time.perf_counter()
        for _ in range(steps):
            c4n = R.trunk_forward(sd, frames[1:2])
            window = window[1:] + [c4n]
            if workload == 'faster_rcnn':
                R.faster_rcnn_simple_test(sd, frames[1:2], metas[0], roi_align_fn=cref.roi_align)
            else:
                R.hnmb_forward_feat(sd, window, metas, key, head=w['head'], roi_align_fn=cref.roi_align)
            done += 1
            if time.perf_counter() - t0 > max_seconds:
                break
        dt = time.perf_counter() - t0
    return dt / done, done


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from hvrnet_b200 import configs
    T = configs.WORKLOADS[args.workload]['t_dim']
    sec, done = cpu_key_frame_seconds(args.workload, max_seconds=150.0, steps=max(1, args.steps), warm=True)
    fps = 1.0 / sec
    line = {
        'impl': 'reference', 'metric': 'VID key frames/sec (1000x600, 300 proposals)', 'value': fps,
        'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps, 'steps_executed': done, 'warmup': args.warmup,
        'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': dict(config_dict(args.workload, T, args.videos_per_gpu or (5 if args.workload == 'hrnmp_inter' else 7),
                                   args.gpus, 'torch CPU, %d threads' % os.cpu_count()),
                       note='reference package is not importable (SURVEY.md 8c): oracle port of its PyTorch-CPU '
                            'path, all host threads; each step = ONE key frame (a bounded sample of the '
                            'key_frames_per_step batch), value = key frames / s'),
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': os.cpu_count(), 'kind': 'port',
                         'sample': '%d key frame(s), trunk on 1 new frame + forward_feat over %d frames' % (done, T)},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def run_torch_cuda(args):
    """SURVEY.md 8d, last row: the reference-style torch-CUDA path as the `existing GPU implementation` bar.
    The reference package cannot be imported, so this runs the oracle port's dense stages - the calls the
    reference makes into cuDNN / cuBLAS (F.conv2d, F.linear, torch.mm, softmax) - on cuda:0 in the same
    as-executed schedule as our arm: trunk on the V new frames, C5 + RPN convolutions on every frame of each
    window, the relation head on T*300 pooled rows per video.  Proposal generation, RoIAlign and NMS are NOT
    included (the oracle has them on the CPU only; < 10 % of our own step), so the figure is an upper bound
    for the library path.  Measured twice: strict fp32 (the precision class of our split-bf16 kernels) and
    with TF32 allowed (about 1e-3 relative, cuDNN / cuBLAS default for convolutions)."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    import torch
    from hvrnet_b200 import configs, synth
    from oracle import ref_torch as R
    dev = torch.device('cuda:0')
    w = configs.WORKLOADS['hrnmp' if args.workload == 'hrnmp_inter' else args.workload]
    T, V = w['t_dim'], args.videos_per_gpu or 7
    K, W = max(1, min(args.steps, 5)), max(1, min(args.warmup, 3))
    sd = {k: v.to(dev) for k, v in synth.make_state_dict(w['head']).items()}
    frames = synth.make_frames(V, seed=0).to(dev)
    g = torch.Generator().manual_seed(0)
    feats = torch.rand(T * 300, 256, 7, 7, generator=g).to(dev)
    key = w['key_dim'] * 300
    torch.backends.cudnn.benchmark = True

    def step():
        with torch.no_grad():
            c4 = R.trunk_forward(sd, frames)
            for v in range(V):
                win = c4[v:v + 1].expand(T, -1, -1, -1).contiguous()      # values do not matter for the timing
                R.c5_forward(sd, win)
                R.rpn_forward(sd, win)
                if w['head'] == 'hrnmp':
                    R.hrnmp_forward_test(sd, feats, key, 300)
                elif w['head'] == 'selsa':
                    R.selsa_forward(sd, feats, key, 300)
                else:
                    R.shared_fc_forward(sd, feats)
    out = {}
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for _ in range(W):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(K):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / K
        out['tf32' if tf32 else 'fp32'] = {'value': V / (ms / 1e3), 'unit': 'frames/s', 'ms_per_step': ms}
    print(json.dumps({
        'impl': 'torch_cuda', 'metric': 'VID key frames/sec (1000x600, 300 proposals)', 'value': out['fp32']['value'],
        'unit': 'frames/s', 'n_gpus': 1, 'steps': K, 'warmup': W, 'ms_per_step': out['fp32']['ms_per_step'],
        'higher_is_better': True, 'dtype': 'f32 (cuDNN / cuBLAS, TF32 off)', 'tf32_allowed': out['tf32'],
        'data': 'synthetic', 'config': config_dict(args.workload, T, V, 1, 'eager torch ops'),
        'note': 'oracle port of the reference path on cuDNN / cuBLAS, dense stages only (no proposal generation, '
                'RoIAlign, NMS): an upper bound for the existing library implementation'}))


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.f.read().splitlines():
            c = [x.strip() for x in ln.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def inter_video_block(args, dev, world, rank, dist, steps=8, warm=None):
    """BASELINE.json configs[4]: `--inter-keys-per-gpu` key frames per rank, every key frame's stage 4 also attends to
    the key rows of 4 other key frames of the whole job (ring order).  Returns the block added to the JSON line."""
    from collections import deque
    from hvrnet_b200 import configs, synth
    from hvrnet_b200.runtime import GraphRunner
    torch.cuda.empty_cache()
    model, sd, w = configs.build_workload('hrnmp_inter', dev)
    T = w['t_dim']
    V = args.inter_keys_per_gpu
    metas = [synth.make_img_meta() for _ in range(T)]
    pool = 3
    frames = synth.make_frames(T + pool, seed=100 + rank)
    devV = [torch.cat([frames[(i + v) % (T + pool)][None] for v in range(V)]).to(dev) for i in range(T + pool)]
    model.enable_cuda_graphs(True)
    # Warm-up = one full turn of the window deques after each switch between the two captured paths (allocator steady
    # state; DESIGN.md section 6 has the trace of what a shorter warm-up used to measure).
    warm = T + 3 if warm is None else warm
    prefetch = os.environ.get('HVR_NO_PREFETCH') != '1'      # (experiments: the next step's trunk on the side stream on / off)
    dqs = [deque(maxlen=T) for _ in range(V)]
    for i in range(T):
        c4 = model(img=devV[i], img_meta=[metas[0]] * V, backbone_feat=True)[0]
        for v, t in enumerate(GraphRunner.per_frame(c4)):
            dqs[v].append(t)

    def step(i):
        c4 = model(img=devV[T + i % pool], img_meta=[metas[0]] * V, backbone_feat=True)[0]
        # the next step's trunk runs on the side stream under this step's window graphs (as in the headline loop): it also
        # keeps the GPU busy while graph C waits for the slowest rank to reach the all-gather
        if prefetch:
            model._runner.prefetch(devV[T + (i + 1) % pool])
        for v, t in enumerate(GraphRunner.per_frame(c4)):
            dqs[v].append(t)
        return model.forward_feat_intervideo([list(d) for d in dqs], metas, n_support=4, rescale=True)
    def timed(fn):
        for i in range(warm):
            fn(i)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        extra = []
        e0.record()
        for i in range(steps):
            fn(warm + i)
            extra.append(model._runner.last_all_gather_ms() or 0.0)   # events on the communication stream; the step has ended
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        return e0.elapsed_time(e1), extra
    # the same batch WITHOUT the inter-video stage (plain forward_feat_batch on the same windows): what the exchange
    # and the longer stage-4 key set cost at this batch size.  Measured before and after the inter-video pass (order
    # effects between two large captures on one device showed up as +-10 %); the better of the two is reported.
    def step_intra(i):
        c4 = model(img=devV[T + i % pool], img_meta=[metas[0]] * V, backbone_feat=True)[0]
        if prefetch:
            model._runner.prefetch(devV[T + (i + 1) % pool])
        for v, t in enumerate(GraphRunner.per_frame(c4)):
            dqs[v].append(t)
        return model.forward_feat_batch([list(d) for d in dqs], metas, rescale=True)
    ms_intra_a, _ = timed(step_intra)
    ms, ag = timed(step)
    ms_intra_b, _ = timed(step_intra)
    ms_intra = min(ms_intra_a, ms_intra_b)
    # the collective alone: back-to-back replays of the same all_gather_into_tensor after a barrier (no rank skew)
    c = model._runner.last_inter
    reps = 10
    torch.cuda.synchronize()
    dist.barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ag_iso = 0.0
    if world > 1:                                             # (world 1: scripts/inter_loss_1gpu.py, no collective)
        dist.all_gather_into_tensor(c.recv_flat, c.state['st'].send)
        a0.record()
        for _ in range(reps):
            dist.all_gather_into_tensor(c.recv_flat, c.state['st'].send)
        a1.record()
        torch.cuda.synchronize()
        ag_iso = a0.elapsed_time(a1) / reps
    t = torch.tensor([ms, max(ag), sum(ag) / len(ag), ms_intra, ag_iso], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ag_max, ag_mean, ms_intra, ag_iso = [float(x) for x in t.tolist()]
    st = c.state['st']
    send_bytes = st.send.numel() * 2
    recv_bytes = send_bytes * (world - 1)
    model.enable_cuda_graphs(False)
    del model
    torch.cuda.empty_cache()
    fps, fps_intra = world * V * steps / (ms / 1e3), world * V * steps / (ms_intra / 1e3)
    return {'workload': 'HVRNet hrnmp batched inference, %d key-frames/GPU, NCCL all-gather of inter-video proposal features' % V,
            'value': fps, 'unit': 'frames/s', 'ms_per_step': ms / steps, 'steps': steps,
            'key_frames_per_gpu': V, 'n_support': 4, 'launch': 'cuda graphs: trunk + three window graphs around the all-gather',
            'intra_same_batch': {'value': fps_intra, 'unit': 'frames/s', 'ms_per_step': ms_intra / steps,
                                 'ms_per_step_before_and_after': [ms_intra_a / steps, ms_intra_b / steps],
                                 'note': 'forward_feat_batch on the same %d windows per rank (no inter-video stage)' % V},
            'loss_vs_intra_same_batch': 1.0 - fps / fps_intra,
            'all_gather': {'ms_isolated': ag_iso, 'achieved_GBs_per_rank': recv_bytes / (ag_iso / 1e3) / 1e9 if ag_iso > 0 else 0.0,
                           'nvlink_reference_GBs': 770.0, 'bytes_sent_per_rank': send_bytes,
                           'bytes_received_per_rank': recv_bytes, 'share_of_step_isolated': ag_iso / (ms / steps),
                           'ms_in_step_mean': ag_mean, 'ms_in_step_max': ag_max,
                           'note': 'ms_isolated: the same all_gather_into_tensor replayed back to back after a barrier. '
                                   'ms_in_step: device time between events on the communication stream around the ONE collective '
                                   'of a step - it includes waiting for the slowest rank to reach it (rank skew) and runs under '
                                   'graph B (branch post-processing, k_4 projection)'},
            'input': 'device-resident frames (%dx3x608x1008 fp32 per step per rank)' % V}


# ----------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == 'torch_cuda':
        return run_torch_cuda(args)
    if args.impl == 'reference':
        return run_reference(args)
    import torch.distributed as dist
    from hvrnet_b200 import _lib, configs, ops, synth
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback; use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        _stdout_to_stderr()      # NCCL prints its version banner on fd 1; stdout must carry the one JSON line only
        dist.init_process_group('nccl', device_id=dev)

    model, sd, w = configs.build_workload(args.workload, dev)
    T = w['t_dim']
    metas = [synth.make_img_meta() for _ in range(T)]
    V = args.videos_per_gpu or (5 if args.workload == 'hrnmp_inter' else 7)
    if args.workload == 'faster_rcnn':
        V = 1
    inter = args.workload == 'hrnmp_inter'
    pool = 4                                             # distinct "new" frames cycled through
    frames = synth.make_frames(T + pool, seed=rank)      # every rank streams its own synthetic video(s)
    host = [frames[i:i + 1].contiguous().pin_memory() for i in range(T + pool)]
    devf = [h.to(dev) for h in host]
    # V videos per GPU: video v is the same synthetic clip shifted by v frames
    hostV = [torch.cat([frames[(i + v) % (T + pool)][None] for v in range(V)]).contiguous().pin_memory()
             for i in range(T + pool)] if (V > 1 or inter) else None
    devV = [h.to(dev) for h in hostV] if hostV is not None else None
    frame_bytes = host[0].numel() * 4 * V
    from hvrnet_b200.runtime import GraphRunner

    def prefill():
        from collections import deque
        dqs = [deque(maxlen=T) for _ in range(V)]
        for i in range(T):
            if V == 1 and not inter:
                dqs[0].append(model(img=devf[i], img_meta=[metas[0]], backbone_feat=True)[0])
            else:
                c4 = model(img=devV[i], img_meta=[metas[0]] * V, backbone_feat=True)[0]
                for v, t in enumerate(GraphRunner.per_frame(c4)):
                    dqs[v].append(t)
        return dqs

    def step(dqs, i, from_host):
        j = T + i % pool
        if V > 1 or inter:
            img = hostV[j] if from_host else devV[j]
            if from_host and model._runner is None:
                img = img.to(dev, non_blocking=True)
            c4 = model(img=img, img_meta=[metas[0]] * V, backbone_feat=True)[0]
            # next step's H2D + trunk overlap this step's window graph (side stream); not in the roofline leg,
            # whose per-launch event timings must not see a second stream's kernels
            if model._runner is not None and model._runner.capture and os.environ.get('HVR_NO_PREFETCH') != '1':
                model._runner.prefetch((hostV if from_host else devV)[T + (i + 1) % pool])
            for v, t in enumerate(GraphRunner.per_frame(c4)):
                dqs[v].append(t)
            if inter:   # configs 4-5: one all-gather of the post-fc_new_4 key rows, ring-order supports
                return model.forward_feat_intervideo([list(d) for d in dqs], metas, n_support=4, rescale=True,
                                                     support_select=args.support_select)
            return model.forward_feat_batch([list(d) for d in dqs], metas, rescale=True)
        dq = dqs[0]
        if from_host:
            img = host[j] if model._runner is not None else host[j].to(dev, non_blocking=True)
        else:
            img = devf[j]
        if args.workload == 'faster_rcnn':
            return model(img=[img], img_meta=[[metas[0]]], return_loss=False, rescale=True)
        dq.append(model(img=img, img_meta=[metas[0]], backbone_feat=True)[0])
        return model(x=list(dq), img=None, img_meta=metas, forward_feat=True, return_loss=False, rescale=True)

    def timed(from_host, K, W, profile=False):
        # the roofline leg brackets individual launches with events, which needs the eager path
        # (the runner's launch sequence - batched head, forked branches - re-issued eagerly: capture=False)
        graphs_ok = not args.eager and args.workload != 'faster_rcnn'
        model.enable_cuda_graphs(graphs_ok, capture=not profile)
        dq = prefill()
        for i in range(W):
            res = step(dq, i, from_host)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if profile:
            ops.PROFILE = []
        l0 = _lib.launch_count() + (model._runner.replayed_launches if model._runner is not None else 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ncu_range = os.environ.get('HVR_NCU_RANGE') == '1'     # `ncu --profile-from-start off`: exactly the K timed steps
        if ncu_range:
            torch.cuda.profiler.start()
        e0.record()
        for i in range(K):
            res = step(dq, W + i, from_host)
        e1.record()
        torch.cuda.synchronize()
        if ncu_range:
            torch.cuda.profiler.stop()
            clocks.stop()
            print(json.dumps({'ncu_range': 'timed region of %d steps profiled; not a bench value' % K}))
            sys.exit(0)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() + (model._runner.replayed_launches if model._runner is not None else 0) - l0
        prof, ops.PROFILE = ops.PROFILE, None
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        def nbytes(r):
            return int(r.nbytes) if hasattr(r, 'nbytes') else sum(nbytes(x) for x in r)
        d2h = nbytes(res)
        return ms, launches, prof, d2h

    K, W = args.steps, max(args.warmup, 3)
    clocks = ClockSampler(local)
    ms, launches, _, d2h = timed(False, K, W)
    clk = clocks.stop()
    ms_e2e, _, _, d2h = timed(True, K, W)
    # roofline leg: the same K steps again with every hvr_igemm launch bracketed by CUDA events on
    # the launching stream (kept out of the headline passes: 2 x 133 event records per step are
    # host work)
    ms_prof, _, prof, _ = timed(False, K, W, profile=True)

    fps = world * V * K / (ms / 1e3)
    fps_e2e = world * V * K / (ms_e2e / 1e3)
    # roofline of the dominant kernel (igemm_tc_kernel): algorithmic FLOPs / summed launch durations
    gemm_ms = sum(p[0].elapsed_time(p[1]) for p in prof)
    gemm_flops = sum(p[2] for p in prof)
    if args.gemm_report and rank == 0:
        import collections
        agg = collections.OrderedDict()
        for e0_, e1_, f_, shp in prof:
            a_ = agg.setdefault(shp, [0, 0.0, 0.0])
            a_[0] += 1; a_[1] += e0_.elapsed_time(e1_); a_[2] += f_
        with open(args.gemm_report, 'w') as fh:
            fh.write('M,N,K,launches_per_step,ms_per_step,algorithmic_TFLOPs\n')
            for shp, (n_, t_, f_) in agg.items():
                fh.write('%d,%d,%d,%.1f,%.4f,%.1f\n' % (shp[0], shp[1], shp[2], n_ / K, t_ / K, f_ / t_ / 1e9))
        # the launches of the LAST profiled step in issue order: start offset from the step's first igemm launch, duration
        per_step = len(prof) // K
        last = prof[-per_step:]
        with open(args.gemm_report.replace('.csv', '') + '_timeline.csv', 'w') as fh:
            fh.write('launch,M,N,K,start_us,us,algorithmic_TFLOPs\n')
            for i_, (e0_, e1_, f_, shp) in enumerate(last):
                t_ = e0_.elapsed_time(e1_)
                fh.write('%d,%d,%d,%d,%.1f,%.1f,%.1f\n' % (i_, shp[0], shp[1], shp[2], last[0][0].elapsed_time(e0_) * 1e3,
                                                             t_ * 1e3, f_ / max(t_, 1e-6) / 1e9))
    peaks, peak_src = None, 'fallback (B200_PROFILING.md: 1590 TFLOP/s burst)'
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    peak = 1590.0
    # GEMMs are timed inside a long step -> the sustained figure; RoIAlign below is timed alone -> burst
    k_, v_ = _measured_peak(peaks, ('bf16',), ('sustain',), 200.0, 5000.0)
    if k_:
        peak, peak_src = v_, 'MEASURED_PEAKS.json ' + k_
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(ROOT, 'profiles', 'igemm_ncu_step.json')))
    except (OSError, ValueError):
        pass

    # second named figure of the metric: RoIAlign achieved HBM GB/s on the step's batched launch
    # (T frames x 300 proposals, 256-channel 38x63 maps, NHWC in -> split rows out)
    def roi_align_roofline():
        g = torch.Generator().manual_seed(5)
        Tn = T * V
        feat = torch.randn(Tn, 38, 63, 256, generator=g).to(dev)
        x1 = torch.rand(Tn * 300, generator=g) * 800
        y1 = torch.rand(Tn * 300, generator=g) * 450
        wh = torch.rand(Tn * 300, 2, generator=g) * 380 + 16
        rois = torch.stack([(torch.arange(Tn * 300) // 300).float(), x1, y1, (x1 + wh[:, 0]).clamp(max=999),
                            (y1 + wh[:, 1]).clamp(max=599)], 1).to(dev)
        def time_us(fn, reps=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps * 1e3
        # the pipeline's variant (separable FMA evaluation, 1e-5 relative to the strict one) and its bit-exact twin
        us = time_us(lambda: ops.roi_align(feat, rois, feat_nhwc=True, out_nhwc=True, want_split=True, want_f32=False,
                                           arithmetic='fast'))
        us_strict = time_us(lambda: ops.roi_align(feat, rois, feat_nhwc=True, out_nhwc=True, want_split=True,
                                                  want_f32=False, arithmetic='strict'))
        nbytes = 17510256.0 * Tn           # write 15 052 800 + map 2 451 456 + rois 6 000 per frame
        hbm_peak, src = 6650.0, 'fallback (B200_PROFILING.md: 6.65 TB/s)'
        k_, v_ = _measured_peak(peaks, ('hbm',), ('burst', 'copy'), 1000.0, 10000.0)
        if k_:
            hbm_peak, src = v_, 'MEASURED_PEAKS.json ' + k_
        gbs = nbytes / us / 1e3
        out = {'bound': 'hbm', 'kernel': 'roi_align_sepp_kernel (one CTA per RoI, thread = output column x 8 lane-interleaved channels, '
                                         'per-RoI row program in shared memory, rows interpolated once along x with merged column taps, '
                                         'coalesced LDG.128, FMA)', 'frames': Tn,
               'rois': Tn * 300, 'us_per_launch': us, 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s',
               'frac': gbs / hbm_peak, 'peak_source': src, 'algorithmic_bytes': nbytes,
               'strict_twin': {'kernel': 'roi_align_sn2_kernel (bit-exact with the reference built -fmad=false)',
                               'us_per_launch': us_strict, 'achieved': nbytes / us_strict / 1e3,
                               'frac': nbytes / us_strict / 1e3 / hbm_peak}}
        # SURVEY.md 8d last row, "existing GPU kernel" bar: the REFERENCE's own RoIAlign / NMS CUDA ops, compiled
        # unmodified for sm_100a into oracle/_ref (checker code: timed here beside ours, after every timed region of
        # the product; never on the product path), on the same launch.  NCHW fp32 in / out as the reference lays it out.
        try:
            from oracle import build as obuild
            ref_roi, ref_nms = obuild.load_ref_roi_align(True), obuild.load_ref_nms_cuda()
        except Exception:                                     # noqa: BLE001
            ref_roi = ref_nms = None
        if ref_roi is not None:
            fc = feat.permute(0, 3, 1, 2).contiguous()
            o_ref = fc.new_zeros(Tn * 300, 256, 7, 7)
            us_ref = time_us(lambda: ref_roi.forward(fc, rois, 7, 7, 1 / 16., 2, o_ref), reps=5)
            del o_ref, fc
            out['vs_reference_kernel'] = {'reference_us_per_launch': us_ref, 'speedup': us_ref / us,
                                          'speedup_strict_twin': us_ref / us_strict,
                                          'reference': 'mmdet/ops/roi_align/src/roi_align_kernel.cu built unmodified for sm_100a '
                                                       '(oracle/_ref), NCHW fp32 output'}
        if ref_nms is not None:
            n = 6000
            c = torch.rand(n, 2, generator=g) * torch.tensor([900., 500.])
            wh_ = torch.rand(n, 2, generator=g) * 200 + 8
            d = torch.cat([c, c + wh_, torch.rand(n, 1, generator=g)], 1).to(dev)
            us_nms_ref = time_us(lambda: ref_nms.nms(d, 0.7), reps=5)
            us_nms = time_us(lambda: ops.nms(d, 0.7), reps=5)
            out['nms_6000'] = {'us_per_call': us_nms, 'reference_us_per_call': us_nms_ref, 'speedup': us_nms_ref / us_nms,
                               'note': 'hvr_nms vs mmdet/ops/nms/src/nms_kernel.cu (oracle/_ref), n = 6000, IoU 0.7, both '
                                       'include their count read-back; identical kept indices (tests)'}
        return out
    roi_rf = roi_align_roofline()

    # extra, clearly separate figure: the streaming scheduler - per-frame caches of
    # proposals and fc_new_1 rows, bit-identical detections, ~570 instead of 2020 GFLOP per key frame.
    # NOT the headline: `value` / `e2e` above time the path as the reference executes it.
    streaming = None
    if args.workload in ('hrnmp', 'selsa') and not args.no_streaming and not args.eager:
        model.enable_cuda_graphs(False)
        src_h, src_d = (hostV, devV) if V > 1 else (host, devf)
        n_src = T + pool

        def stream_pass(src, capture=True, profile=False):
            """K timed steps of the streaming scheduler through the detector's call surface (stream=True)."""
            model.enable_streaming(V, window=T, capture=capture)
            for i in range(T + W):                                 # fill the windows + warm-up (captures both graphs)
                model(img=src[i % n_src], img_meta=[metas[0]], stream=True, rescale=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            if profile:
                ops.PROFILE = []
            l0 = _lib.launch_count() + model._streamer.replayed_launches
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for i in range(K):
                res = model(img=src[(T + W + i) % n_src], img_meta=[metas[0]], stream=True, rescale=True)
            s1.record()
            torch.cuda.synchronize()
            sms = s0.elapsed_time(s1)
            n_launch = _lib.launch_count() + model._streamer.replayed_launches - l0
            prof, ops.PROFILE = ops.PROFILE, None
            if world > 1:
                t = torch.tensor([sms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sms = float(t.item())
            model.enable_streaming(flag=False)
            return sms, prof, n_launch, res
        sms_dev, _, s_launch, s_res = stream_pass(src_d)
        sms_e2e, _, _, _ = stream_pass(src_h)
        sms_prof, sprof, _, _ = stream_pass(src_d, capture=False, profile=True)
        s_gemm_ms = sum(p_[0].elapsed_time(p_[1]) for p_ in sprof)
        s_gemm_flops = sum(p_[2] for p_ in sprof)
        s_ach = s_gemm_flops / (s_gemm_ms / 1e3) / 1e12 if s_gemm_ms > 0 else 0.0

        def nbytes_(r):
            return int(r.nbytes) if hasattr(r, 'nbytes') else sum(nbytes_(x) for x in r)
        streaming = {'value': world * V * K / (sms_dev / 1e3), 'unit': 'frames/s', 'ms_per_step': sms_dev / K,
                     'e2e': {'value': world * V * K / (sms_e2e / 1e3), 'unit': 'frames/s', 'ms_per_step': sms_e2e / K,
                             'h2d_bytes_per_step': frame_bytes, 'd2h_bytes_per_step': nbytes_(s_res)},
                     'gpu_launches': int(s_launch),
                     'executed_gflop_per_key_frame': s_gemm_flops / K / V / 1e9,
                     'roofline': {'bound': 'tensor', 'kernel': 'igemm_tc_kernel (tcgen05 split-bf16 implicit GEMM)',
                                  'achieved': s_ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': s_ach / peak,
                                  'tensor_work_frac': 3.0 * s_ach / peak, 'peak_source': peak_src,
                                  'kernel_ms_per_step': s_gemm_ms / K, 'share_of_step': s_gemm_ms / sms_prof,
                                  'launches_per_step': len(sprof) / K},
                     'call': 'model.enable_streaming(V, window=T); model(img=frames, img_meta=[meta], stream=True)',
                     'note': 'streaming scheduler with per-frame caches (SURVEY 8f N1): same detections bit for '
                             'bit (ragged frames included), less work per key frame than the reference executes; '
                             'reported for information, not the headline'}
    line = {
        'metric': 'VID key frames/sec (1000x600, 300 proposals)', 'value': fps, 'unit': 'frames/s', 'n_gpus': world,
        'steps': K, 'warmup': W, 'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'bf16x3 (split-bf16 operands, 3 tcgen05 products, fp32 accumulate)',
        'data': 'synthetic',
        'config': config_dict(args.workload, T, V, world, 'eager' if (args.eager or args.workload == 'faster_rcnn')
                              else 'cuda graphs (trunk + window)'),
        'e2e': {'value': fps_e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': frame_bytes, 'd2h_bytes_per_step': d2h,
                'ms_per_step': ms_e2e / K},
        'gpu_launches': int(launches),
        'clocks': clk,
        'roofline': {'bound': 'tensor', 'kernel': 'igemm_tc_kernel (tcgen05 split-bf16 implicit GEMM)',
                     'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
                     'tensor_work_tflops': 3.0 * achieved, 'tensor_work_frac': 3.0 * achieved / peak,
                     'peak_source': peak_src,
                     # ncu dram__bytes_read.sum + dram__bytes_write.sum, average per igemm launch of one step, and the
                     # time-weighted tensor-pipe activity: read from the summary file scripts/ncu_metrics_summary.py writes
                     # from an ncu range capture of exactly this command's timed region (never constants in this file)
                     'traffic': ncu.get('traffic_bytes_per_launch') if (args.workload == 'hrnmp' and V == 7) else None,
                     'traffic_source': ncu.get('source'),
                     'ncu_tensor_pipe_active_pct': ncu.get('tensor_pipe_active_pct') if (args.workload == 'hrnmp' and V == 7) else None,
                     'algorithmic_gflop_per_step': gemm_flops / K / 1e9, 'launches_per_step': len(prof) / K,
                     'kernel_ms_per_step': gemm_ms / K, 'share_of_step': gemm_ms / ms_prof,
                     'profiled_ms_per_step': ms_prof / K,
                     'note': 'achieved counts algorithmic fp32-equivalent FLOPs; the kernel issues 3 bf16 MMAs per '
                             'product (tensor-pipe work = 3x), so frac <= 1/3 by construction'},
    }
    if args.workload == 'hrnmp_inter':
        line['config']['support_select'] = args.support_select
    line['roi_align'] = roi_rf
    # the other single-GPU configurations of BASELINE.json, measured briefly through the same public
    # call surface (one video per GPU, device-resident frames) so every config has a number on record
    if world == 1 and args.workload == 'hrnmp' and not args.no_other_workloads and not args.eager:
        others = {}
        for wl in ('selsa', 'faster_rcnn'):
            model = None
            torch.cuda.empty_cache()
            model, _, w2 = configs.build_workload(wl, dev)
            T2 = w2['t_dim']
            metas2 = [synth.make_img_meta() for _ in range(T2)]
            model.enable_cuda_graphs(True)
            from collections import deque

            def run(Vb):
                """Vb videos per step through the public call surface, device-resident frames; -> key frames / s"""
                src = devV if Vb > 1 else devf
                dqs2 = [deque(maxlen=T2) for _ in range(Vb)]

                def one(i):
                    img = src[i % len(src)]
                    if wl == 'faster_rcnn':
                        if Vb == 1:
                            return model(img=[img], img_meta=[[metas2[0]]], return_loss=False, rescale=True)
                        return model.simple_test_batch(img, [metas2[0]], rescale=True)
                    c4 = model(img=img, img_meta=[metas2[0]] * Vb, backbone_feat=True)[0]
                    for v_, t_ in enumerate(GraphRunner.per_frame(c4)):
                        dqs2[v_].append(t_)
                    if len(dqs2[0]) < T2:
                        return None
                    if Vb == 1:
                        return model(x=list(dqs2[0]), img=None, img_meta=metas2, forward_feat=True, return_loss=False,
                                     rescale=True)
                    return model.forward_feat_batch([list(d) for d in dqs2], metas2, rescale=True)
                for i in range(T2 + 3):
                    one(i)
                torch.cuda.synchronize()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                for i in range(10):
                    one(T2 + 3 + i)
                b_.record()
                torch.cuda.synchronize()
                return Vb * 10 / (a_.elapsed_time(b_) / 1e3)
            one_v, batched = run(1), run(V)
            others[wl] = {'value': one_v, 'unit': 'frames/s', 'videos_per_gpu': 1, 'workload': workload_name(wl, T2),
                          'batched': {'value': batched, 'unit': 'frames/s', 'videos_per_gpu': V,
                                      'call': 'simple_test_batch' if wl == 'faster_rcnn' else 'forward_feat_batch'}}
        line['other_workloads'] = others
        model = None
    if streaming is not None:
        line['streaming'] = streaming
    # BASELINE.json configs[4] at N > 1: 32 key frames per rank, inter-video stage with ONE NCCL all-gather
    # of the post-fc_new_4 key rows per step, three CUDA graphs around it (runtime.GraphRunner.detect_inter).  Device-
    # resident frames; timed like the headline (events, barrier, max over ranks).
    if world > 1 and args.workload == 'hrnmp' and not args.eager and not args.no_inter_video:
        line['inter_video'] = inter_video_block(args, dev, world, rank, dist)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            sec, done = cpu_key_frame_seconds(args.workload, max_seconds=60.0, steps=1)
            line['cpu_baseline'] = {'value': 1.0 / sec, 'unit': 'frames/s', 'cores': os.cpu_count(), 'kind': 'port',
                                    'sample': '%d key frame, trunk on 1 new frame + forward_feat over %d frames '
                                              '(oracle port of the reference PyTorch-CPU path)' % (done, T)}
        _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
