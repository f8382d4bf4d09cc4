"""World-size-2 test (gloo, CPU) of the multi-GPU host logic: pixel partition balanced by candidate count and the
gather of result maps.  The single-device compute function is the CPU oracle here; on the GPU box the same driver
wraps CLEDB_PROC.cledb_invproc (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cledb_b200.synth as synth
from cledb_b200 import dist as cdist
import oracle.cledb_oracle as orc


def make_case():
    dbs = [synth.make_database(8, 39, ngx=3, nbphi=12, nbtheta=8), synth.make_database(20, 70, ngx=4, nbphi=12, nbtheta=8)]
    nx, ny = 6, 5
    enc = (np.arange(nx * ny).reshape(nx, ny) % 3 == 0).astype(np.int32)
    obs = synth.make_observation(nx, ny, dbs, db_enc=enc, seed=11, frac_nan=0.1, frac_zero=0.05)
    hdr = synth.make_dbhdr(3, 12, 8)
    return dbs, obs, hdr


def oracle_invproc(sobs, dopp, database, db_enc, yobs, aobs, dobs, snr, issuemask, dbhdr, keyvals, nsearch, maxchisq, bcalc,
                   iqud, ncpu, reduced, verbose):
    return orc.invert_map(sobs, dopp, database, db_enc, yobs, aobs, dobs, snr, issuemask, dbhdr, keyvals, nsearch, maxchisq,
                          bcalc, iqud)


def worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dbs, obs, hdr = make_case()
    iss = obs['issuemask'].copy()
    inv, sf = cdist.invproc_sharded(oracle_invproc, obs['sobs_totrot'], obs['sobs_dopp'], dbs, obs['db_enc'], obs['yobs'],
                                    obs['aobs'], obs['dobs'], obs['snr'], iss, hdr, obs['keyvals'], 4, 1000, 0, False, 1, False, 0)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), inv=inv, sf=sf, iss=iss)
    dist.destroy_process_group()


def test_partition_covers_and_balances():
    rng = np.random.default_rng(0)
    enc = rng.integers(0, 5, (40, 30))
    w = [1000, 10, 500, 2000, 1]
    for world in (1, 2, 3, 8):
        parts = cdist.partition(enc, w, world)
        allp = np.concatenate(parts)
        assert sorted(allp.tolist()) == list(range(enc.size))
        loads = [np.asarray(w)[enc.reshape(-1)[p]].sum() for p in parts]
        assert max(loads) - min(loads) <= max(w) * 2
        for p in parts:                                     # contiguous in database order: few databases per rank
            e = enc.reshape(-1)[p]
            assert (np.diff(e) >= 0).all()


def test_sharded_inversion_world2(tmp_path):
    world, port = 2, 29531 + os.getpid() % 200
    mp.spawn(worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    dbs, obs, hdr = make_case()
    iss = obs['issuemask'].copy()
    ref_inv, ref_sf = oracle_invproc(obs['sobs_totrot'], obs['sobs_dopp'], dbs, obs['db_enc'], obs['yobs'], obs['aobs'], obs['dobs'],
                                     obs['snr'], iss, hdr, obs['keyvals'], 4, 1000, 0, False, 1, False, 0)
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r))
        assert np.array_equal(g['inv'], ref_inv, equal_nan=True)
        assert np.array_equal(g['sf'], ref_sf, equal_nan=True)
        assert np.array_equal(g['iss'], iss)


def test_c_abi_partition_equals_the_numpy_statement():
    """cledb_partition (pure host code of the library, no GPU) cuts the pixels exactly as cledb_b200.dist.partition does."""
    from cledb_b200 import native
    rng = np.random.default_rng(1)
    for ndb, shape in ((5, (40, 30)), (247, (64, 64)), (1, (9, 7))):
        enc = rng.integers(0, ndb, shape).astype(np.int32)
        w = rng.integers(1, 3000, ndb).astype(np.float64)
        for world in (1, 2, 3, 8):
            order, bounds = native.partition(enc.reshape(-1), w, world)
            parts = cdist.partition(enc, w, world)
            assert bounds[0] == 0 and bounds[-1] == enc.size and len(bounds) == world + 1
            for r in range(world):
                assert np.array_equal(order[bounds[r]:bounds[r + 1]], parts[r])


def test_shard_and_map_layouts():
    """cledb_shard_layout / cledb_map_layout (pure host code): the four blocks of the packed records and of the full maps follow
    one another, 256-byte aligned, each large enough for its rows - what cledb_gather_maps_dev and the bench rely on."""
    import ctypes
    from cledb_b200 import native
    lib = native.load_library()
    for fn in (lib.cledb_shard_layout, lib.cledb_map_layout):
        for n, k in ((0, 8), (1, 1), (3000, 3), (131071, 8), (1 << 20, 32)):
            lay = np.zeros(5, np.int64)
            assert fn(n, k, lay.ctypes.data_as(ctypes.c_void_p)) == 0
            o_inv, o_sf, o_st, o_iss, total = (int(v) for v in lay)
            assert o_inv == 0 and all(v % 256 == 0 for v in lay)
            assert o_sf - o_inv >= n * k * 44 and o_st - o_sf >= n * k * 32 and o_iss - o_st >= n and total - o_iss >= n * 4
            assert total < n * (k * 76 + 5) + 4 * 256 + 1
        assert fn(-1, 8, lay.ctypes.data_as(ctypes.c_void_p)) < 0 and fn(10, 0, lay.ctypes.data_as(ctypes.c_void_p)) < 0
