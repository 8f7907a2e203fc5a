"""Per-launch device times (CUDA events, best of 5) and algorithmic DRAM bytes of ONE relation stage at the bench's
shapes (V = 7 videos, N = 4500 keys per video, D = 1024), for the two kinds of stage:

  all-row stage (1, 3):  queries = all V*Npad rows           key-only stage (2, 4):  queries = the 300 key rows per video

and the measurements behind DESIGN.md's answer to the north_star's "one fused linear -> QK^T -> softmax -> PV kernel":
  (a) q_data_fc | k_data_fc as ONE N = 2048 GEMM (shipped, engine.FUSE_QK) against the two-GEMM evaluation;
  (b) what a fused kernel could save: the logit round trip (S written fp32 by the QK^T epilogue, read by the softmax,
      P written split, read by P.V) = the softmax launch + the S / P bytes;
  (c) what a flash-style fused kernel would pay: its QK^T phase has to run on single-CTA 128 x 128 tiles with the query
      tile re-streamed for every key tile (a 128 x 1024 split query tile is 512 KB: it cannot stay in shared memory, and
      the 128 x 1024 fp32 output accumulator is 1024 TMEM columns, twice what a CTA has).  That operand-ingest pattern is
      exactly what hvr_igemm's single-CTA BN = 128 kernel does, so it is MEASURED here (hvr_debug_force_bn(128)) against
      the CTA-pair 256 x 256 kernel the un-fused path uses - an upper bound for the fused kernel's S phase.

    python scripts/relation_stage_bench.py > profiles/r02_relation_stage_bench.txt
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hvrnet_b200 import _lib, configs, engine, ops  # noqa: E402

dev = torch.device('cuda:0')
V, T, P, D = 7, 15, 300, 1024
N, Npad = T * P, ops.round_up(T * P, 64)
m, sd, w = configs.build_workload('hrnmp', dev)
Pk = m.bbox_head.packed(dev)
g = torch.Generator().manual_seed(0)
X = ops.split((torch.randn(V * Npad, D, generator=g) * 0.5).to(dev))
XT = ops.transpose_split(X, D)
seg = torch.full((V, T), P, dtype=torch.int32, device=dev)


def best_us(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return min(ts)


def row(name, us, flops=0.0, nbytes=0.0):
    print('  %-46s %9.1f us %8.1f TFLOP/s %8.2f GB  %7.2f TB/s' % (name, us, flops / us / 1e6 if flops else 0.0, nbytes / 1e9,
                                                                   nbytes / us / 1e6 if nbytes else 0.0))
    return us


def stage(kind):
    nq = Npad if kind == 'all' else P
    Mq = V * nq
    print('\n== %s stage: %d query rows (%d per video) x %d keys per video, D = %d' % (
        'all-row' if kind == 'all' else 'key-only', Mq, nq, N, D))
    Xq = X if kind == 'all' else engine.key_rows(X, V, Npad, 7 * P, P)
    sp = 4.0                                                           # bytes per split element (hi + lo)
    tot = {}
    if kind == 'all':
        fl = 2.0 * V * Npad * D * D
        tot['qk2'] = row('q_data_fc + k_data_fc, two GEMMs (N = 1024 each)',
                         best_us(lambda: (engine.lin(X, Pk['q1']), engine.lin(X, Pk['k1']))), 2 * fl,
                         2 * V * Npad * D * sp + 2 * V * Npad * D * sp)
        tot['qk1'] = row('q | k as ONE GEMM (N = 2048)  [shipped]', best_us(lambda: engine._qk(Pk, 1, X)), 2 * fl,
                         V * Npad * D * sp + 2 * V * Npad * D * sp)
        Q, K = engine._qk(Pk, 1, X)
    else:
        tot['qk2'] = row('q_data_fc (key rows) + k_data_fc (all rows)',
                         best_us(lambda: (engine.lin(Xq, Pk['q2']), engine.lin(X, Pk['k2']))),
                         2.0 * (Mq + V * Npad) * D * D, (Mq + V * Npad) * D * sp * 2)
        Q, _, _ = engine.lin(Xq, Pk['q2'])
        K, _, _ = engine.lin(X, Pk['k2'])
    fl_s = 2.0 * Mq * N * D
    bmm_s = lambda: ops.bmm(Q, K, V, N, Npad * K.hi.stride(0), alpha=1.0 / math.sqrt(D), want_split=False, want_f32=True)
    tot['s'] = row('S = Q K^T / sqrt(D) -> fp32 logits (CTA pairs)', best_us(bmm_s), fl_s,
                   (Mq + V * Npad) * D * sp + Mq * N * 4.0)
    _, S = bmm_s()
    sm = lambda: ops.softmax_rows_split(S, N, ld_p=Npad, seg_counts=seg, slot=P, rows_per_problem=nq)
    tot['sm'] = row('row softmax -> split probabilities', best_us(sm), 0.0, Mq * N * 4.0 + Mq * Npad * sp)
    row('  (the same without the key mask)', best_us(lambda: ops.softmax_rows_split(S, N, ld_p=Npad)), 0.0,
        Mq * N * 4.0 + Mq * Npad * sp)
    Pm = sm()
    tot['pv'] = row('O = P X (values = un-projected rows)', best_us(lambda: ops.bmm(Pm, XT, V, D, Npad)), 2.0 * Mq * Npad * D,
                    Mq * Npad * sp + V * Npad * D * sp + Mq * D * sp)
    O, _ = ops.bmm(Pm, XT, V, D, Npad)
    tot['o'] = row('linear_out + residual + ReLU', best_us(lambda: engine.lin(O, Pk['o1'], relu=True, res=Xq)),
                   2.0 * Mq * D * D, 3 * Mq * D * sp)
    tot['xt'] = row('X^T (split transpose, once per fc_new_k output)', best_us(lambda: ops.transpose_split(X, D)), 0.0,
                    2 * V * Npad * D * sp)
    base = tot['qk2'] + tot['s'] + tot['sm'] + tot['pv'] + tot['o']
    now = tot.get('qk1', tot['qk2']) + tot['s'] + tot['sm'] + tot['pv'] + tot['o']
    print('  stage total: %.1f us with separate q / k projections, %.1f us as shipped' % (base, now))
    rt_bytes = Mq * N * 4.0 * 2 + Mq * Npad * sp * 2
    print('  (b) logit round trip a fused kernel would save: softmax launch %.1f us; S write + read and P write + read = %.2f GB'
          ' (%.1f us at 6.5 TB/s, already overlapped with the GEMMs\' own traffic)' % (tot['sm'], rt_bytes / 1e9,
                                                                                     rt_bytes / 6.5e12 * 1e6))
    # (c) the S phase of a flash-style kernel: single-CTA 128 x 128 tiles, query tile re-streamed per key tile
    L = _lib.lib()
    L.hvr_debug_force_bn(128)
    try:
        us128 = best_us(bmm_s)
    finally:
        L.hvr_debug_force_bn(0)
    print('  (c) the same QK^T on single-CTA 128 x 128 tiles (the fused kernel\'s operand-ingest pattern): %.1f us = %.1f TFLOP/s,'
          ' %.2fx the CTA-pair kernel' % (us128, fl_s / us128 / 1e6, us128 / tot['s']))
    print('      -> fusing would trade a %.1f us softmax launch for +%.1f us in the QK^T phase alone, before the P.V phase pays the'
          ' same tile shape and the 2-fold QK^T recomputation the 1024-column accumulator forces (512 TMEM columns per CTA)'
          % (tot['sm'], us128 - tot['s']))


print('relation stage at the bench shapes on one B200 (CUDA events, best of 5); bytes are algorithmic (each operand once)')
stage('all')
stage('key')
