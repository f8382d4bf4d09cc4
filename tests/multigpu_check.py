#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs, one rank per GPU; tests/test_multigpu.py launches it):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
cledb_invproc as a collective (cledb_b200.dist.invproc_sharded, device='cuda'): every rank searches and builds its pixel
shard and the library gathers the finished records into the full maps of every rank (cledb_invert_sharded) - once with its
own peer-memory transport (k_push_records: each record stored straight at its pixel in every rank's window over NVLink,
free / done flags) where CUDA IPC works on the box, once with ncclAllGather + k_unshard.  With both transports the gathered
maps and issue masks must equal the single-GPU result bit for bit on EVERY rank, in both modes, and a seeded subset must
match the CPU oracle (near-tie proof member by member).  The window all-gather (cledb_allgather_dev into the window) is
checked against torch's all_gather."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200'), os.path.join(REPO, 'tests')]
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

from oracle.compare import check_invout_against_reference  # noqa: E402

nx = 96
dbs = [synth.make_database(5 + 3 * h, 30 + 2 * d) for h in range(2) for d in range(3)]
enc = (np.arange(nx)[:, None] * 2 // nx * 3 + np.random.default_rng(5).integers(0, 3, (nx, nx))).astype(np.int32)
obs = synth.make_observation(nx, nx, dbs, db_enc=enc, seed=21)
hdr = synth.make_dbhdr()
args = (obs['sobs_totrot'], obs['sobs_dopp'], dbs, obs['db_enc'], obs['yobs'], obs['aobs'], obs['dobs'], obs['snr'])
ok = True
info = None
ctx = cdist.init_comm()
default_transport = ctx.comm_transport()
print('rank %d/%d: default transport %s' % (rank, world, default_transport), flush=True)
transports = ['peer', 'nccl'] if default_transport == 'peer' else ['nccl']
if os.environ.get('CLEDB_B200_REQUIRE_PEER') == '1':
    ok = ok and default_transport == 'peer'
# 'peer/3': the pipelined gather with the shard in three pieces (a map this small goes through in one piece by default)
variants = (['peer', 'peer/3'] if default_transport == 'peer' else []) + ['nccl']
for transport, iqud, bcalc in [(t, q, b) for t in variants for q, b in ((False, 0), (True, 3))]:
    ctx.comm_transport(transport.split('/')[0])
    os.environ.pop('CLEDB_B200_PIECES', None)
    if '/' in transport:
        os.environ['CLEDB_B200_PIECES'] = transport.split('/')[1]
    tail = (hdr, obs['keyvals'], 8, 1000, bcalc, iqud, 1, False, 0)
    iss = obs['issuemask'].copy()
    iss[::3, :, 1] = 1024                                    # a code that is already set
    iss_in = iss.copy()
    inv, sf = cdist.invproc_sharded(procinv.cledb_invproc, *args, iss, *tail, device=torch.device('cuda', local))
    info = procinv._context().comm_info()
    iss1 = iss_in.copy()
    inv1, sf1 = procinv.cledb_invproc(*args, iss1, *tail)   # the same map on this rank's GPU alone
    same = np.array_equal(inv, inv1, equal_nan=True) and np.array_equal(sf, sf1, equal_nan=True) and np.array_equal(iss, iss1)
    print('rank %d/%d iqud=%s transport=%s: gathered maps == single-GPU maps: %s (NCCL %s)' % (rank, world, iqud, transport, same, info), flush=True)
    ok = ok and same and info['nranks'] == world and (world == 1 or info['nccl'])
    if rank == 0 and transport == variants[0]:
        dec = [codec.decoded_database(d, native.ENC_I16).reshape(-1, 8) for d in dbs]
        pick = np.random.default_rng(2).choice(nx * nx, 48, replace=False)
        pix = [(int(p // nx), int(p % nx)) for p in pick]
        iss2 = iss_in.copy()
        ref, ref_sf = orc.invert_map(obs['sobs_totrot'], obs['sobs_dopp'], [d.reshape(dbs[0].shape) for d in dec], obs['db_enc'], obs['yobs'],
                                     obs['aobs'], obs['dobs'], obs['snr'], iss2, hdr, obs['keyvals'], 8, 1000, bcalc, iqud, pixels=pix)
        n = nx * nx
        slots, subs = check_invout_against_reference(inv.reshape(n, 8, 11)[pick], ref.reshape(n, 8, 11)[pick], int(iqud),
                                                     obs['sobs_totrot'].reshape(n, 8)[pick], obs['snr'].reshape(n, 8)[pick], dec,
                                                     enc.reshape(-1)[pick], 1000, sf.reshape(n, 8, 8)[pick], ref_sf.reshape(n, 8, 8)[pick],
                                                     iss.reshape(n, 2)[pick], iss2.reshape(n, 2)[pick])
        print('oracle subset: %d slots, %d near-tie substitutions' % (slots, subs), flush=True)
        ok = ok and subs <= 8
os.environ.pop('CLEDB_B200_PIECES', None)
# root-only gather: the other ranks get (None, None); with every transport the root's maps are the single-GPU maps
tail = (hdr, obs['keyvals'], 8, 1000, 0, False, 1, False, 0)
inv1, sf1 = procinv.cledb_invproc(*args, obs['issuemask'].copy(), *tail)
for transport in transports:
    ctx.comm_transport(transport)
    inv_r, sf_r = cdist.invproc_sharded(procinv.cledb_invproc, *args, obs['issuemask'].copy(), *tail, device='cuda', root=world - 1)
    ok = ok and ((inv_r is not None) == (rank == world - 1))
    if inv_r is not None:
        ok = ok and np.array_equal(inv_r, inv1, equal_nan=True) and np.array_equal(sf_r, sf1, equal_nan=True)
# all-gather of contiguous blocks into the window against torch's own all_gather, several rounds back to back
from helpers import device_view  # noqa: E402
for transport in transports:
    ctx.comm_transport(transport)
    nb = 3 << 20
    win = ctx.comm_window(world * nb + 4096)
    for rep in range(4):
        mine = torch.randint(0, 256, (nb,), dtype=torch.uint8, device='cuda')
        want = torch.empty(world * nb, dtype=torch.uint8, device='cuda')
        dist.all_gather_into_tensor(want, mine)
        torch.cuda.synchronize()
        ctx.allgather_dev(mine.data_ptr(), win + 4096, nb, None)
        ctx.synchronize()
        good = bool(torch.equal(device_view(win + 4096, world * nb, torch.device('cuda', local)), want))
        ok = ok and good
    print('rank %d/%d transport=%s: window all-gather == torch all_gather: %s' % (rank, world, transport, good), flush=True)
flag = torch.tensor([1 if ok else 0], device='cuda')
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
