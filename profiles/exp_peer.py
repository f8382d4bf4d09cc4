#!/usr/bin/env python
"""Tuning experiment of the peer-memory push kernels (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/exp_peer.py
Times cledb_gather_maps_dev (k_push_records with a bounded number of blocks) on a random 1024^2-pixel
partition and cledb_allgather_dev into the window (k_push_block) on large blocks, CUDA events, max over ranks."""
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200'), os.path.join(REPO, 'tests')]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from cledb_b200 import native  # noqa: E402
from helpers import device_view  # noqa: E402

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ctx = native.Context(local)
ids = [native.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.comm_init(rank, world, ids[0])
print('rank', rank, 'transport', ctx.comm_transport(), flush=True)
lib = native.load_library()
npix, k = 1024 * 1024, 8
rng = np.random.default_rng(7)
order = rng.permutation(npix).astype(np.int32)
bounds = np.linspace(0, npix, world + 1).astype(np.int64)
n = int(bounds[rank + 1] - bounds[rank])
maxcount = int(np.diff(bounds).max())
lay, mlay = np.zeros(5, np.int64), np.zeros(5, np.int64)
lib.cledb_shard_layout(maxcount, k, lay.ctypes.data_as(ctypes.c_void_p))
lib.cledb_map_layout(npix, k, mlay.ctypes.data_as(ctypes.c_void_p))
send = torch.randint(0, 256, (int(lay[4]),), dtype=torch.uint8, device=dev)
mine = torch.from_numpy(order[bounds[rank]:bounds[rank + 1]].copy()).to(dev)
win = ctx.comm_window(int(mlay[4]))
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=6):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = torch.tensor([float(np.mean(ts)), float(np.min(ts))], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


for blocks in (4, 8, 16, 24, 32, 64, 148, 0):
    for root in (-1, 0):
        os.environ['CLEDB_B200_PUSH_BLOCKS'] = str(blocks)
        mean, best = timed(lambda: ctx.gather_maps_dev(npix, n, mine.data_ptr(), maxcount, k, send.data_ptr(), win, root, stream.cuda_stream))
        in_mb = (npix - n) * (k * 76 + 5) / 1e6 if (root < 0 or root == rank) else 0.0
        if rank == 0:
            print('push with %3d blocks (0 = 8 per SM) root %2d: %.3f ms mean  %.3f ms min   %.0f MB into rank 0 -> %.0f GB/s' %
                  (blocks, root, mean, best, in_mb, in_mb / best if best else 0), flush=True)
os.environ.pop('CLEDB_B200_PUSH_BLOCKS', None)
# contiguous blocks
for mb in (4, 32, 160):
    nb = mb << 20
    w = ctx.comm_window(max(int(mlay[4]), world * nb))
    src = torch.randint(0, 256, (nb,), dtype=torch.uint8, device=dev)
    mean, best = timed(lambda: ctx.allgather_dev(src.data_ptr(), w, nb, stream.cuda_stream))
    want = torch.empty(world * nb, dtype=torch.uint8, device=dev)
    mean2, best2 = timed(lambda: dist.all_gather_into_tensor(want, src))
    torch.cuda.synchronize()
    same = bool(torch.equal(device_view(w, world * nb, dev), want))
    if rank == 0:
        print('all-gather of %3d MB blocks: window push %.3f ms (%.0f GB/s out), torch/NCCL %.3f ms; equal %s' %
              (mb, best, nb * (world - 1) / best / 1e6, best2, same), flush=True)
dist.barrier()
ctx.comm_destroy()
dist.destroy_process_group()
