"""bench.py's inter_video block at world size 1 (no peer, the all-gather degenerates): what the three-graph split and the longer
stage-4 key set cost by themselves, without rank skew.   python scripts/inter_loss_1gpu.py [keys_per_gpu] [steps]"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
os.environ.setdefault('MASTER_PORT', '29577')
dist.init_process_group('nccl', rank=0, world_size=1)
dev = torch.device('cuda:0')
torch.cuda.set_device(dev)


class A:
    inter_keys_per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 32


steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
out = bench.inter_video_block(A, dev, 1, 0, dist, steps=steps)
print(json.dumps({k: out[k] for k in ('value', 'ms_per_step', 'intra_same_batch', 'loss_vs_intra_same_batch')}))
dist.destroy_process_group()
