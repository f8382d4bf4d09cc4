#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on the GPU box, one rank per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/dist_check.py
Every rank inverts its pixel shard with CLEDB_PROC.cledb_invproc on its own GPU (cledb_b200.dist.invproc_sharded); the
gathered maps must equal the single-GPU result of rank 0 bit for bit, and a seeded subset must match the CPU oracle."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from cledb_b200 import synth, dist as cdist, codec, native  # noqa: E402
import CLEDB_PROC.CLEDB_PROC as procinv  # noqa: E402
import oracle.cledb_oracle as orc  # noqa: E402

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
procinv.OPTIONS.update(device=local, dbencoding=native.ENC_I16, precision=0)

nx = 96
dbs = [synth.make_database(5 + 3 * h, 30 + 2 * d) for h in range(2) for d in range(3)]
enc = (np.arange(nx)[:, None] * 2 // nx * 3 + np.random.default_rng(5).integers(0, 3, (nx, nx))).astype(np.int32)
obs = synth.make_observation(nx, nx, dbs, db_enc=enc, seed=21)
hdr = synth.make_dbhdr()
args = (obs['sobs_totrot'], obs['sobs_dopp'], dbs, obs['db_enc'], obs['yobs'], obs['aobs'], obs['dobs'], obs['snr'])
tail = (hdr, obs['keyvals'], 8, 1000, 0, False, 1, False, 0)
iss = obs['issuemask'].copy()
inv, sf = cdist.invproc_sharded(procinv.cledb_invproc, *args, iss, *tail, device=torch.device('cuda', local))
ok = True
if rank == 0:
    iss1 = obs['issuemask'].copy()
    inv1, sf1 = procinv.cledb_invproc(*args, iss1, *tail)
    same = np.array_equal(inv, inv1, equal_nan=True) and np.array_equal(sf, sf1, equal_nan=True) and np.array_equal(iss, iss1)
    print('world %d: gathered maps == single-GPU maps: %s' % (world, same))
    dec = [codec.decoded_database(d, native.ENC_I16).reshape(d.shape) for d in dbs]
    rng = np.random.default_rng(2)
    pix = [(int(p // nx), int(p % nx)) for p in rng.choice(nx * nx, 48, replace=False)]
    iss2 = obs['issuemask'].copy()
    ref, _ = orc.invert_map(obs['sobs_totrot'], obs['sobs_dopp'], dec, obs['db_enc'], obs['yobs'], obs['aobs'], obs['dobs'], obs['snr'],
                            iss2, hdr, obs['keyvals'], 8, 1000, 0, False, pixels=pix)
    nbad = 0
    for (x, y) in pix:
        a, b = inv[x, y], ref[x, y]
        if not np.array_equal(a[:, 0], b[:, 0]):
            # allowed only inside chi^2 near-ties (1e-5 relative)
            chi = b[:, 1]
            if not np.allclose(np.sort(a[:, 1]), np.sort(chi), rtol=2e-5, atol=1e-9):
                nbad += 1
        elif not np.allclose(a, b, rtol=2e-5, atol=1e-9, equal_nan=True):
            nbad += 1
    print('oracle subset: %d pixels, %d mismatching' % (len(pix), nbad))
    ok = same and nbad == 0
flag = torch.tensor([1 if ok else 0], device='cuda')
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
