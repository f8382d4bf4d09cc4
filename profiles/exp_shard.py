#!/usr/bin/env python
"""Where the time of a shard search goes (one GPU): rank 0's shard of the config-5 map as cledb_partition cuts it for 8 ranks
(131 Ki pixels, 31 databases), searched + built whole, in halves and in quarters.  python profiles/exp_shard.py [nranks]"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from cledb_b200 import native, synth  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
nx, ndb, k = 1024, 247, 8
npix = nx * nx
nh = max(1, int(round(ndb ** 0.5)))
nd = (ndb + nh - 1) // nh
grid = [(h, d) for h in range(nh) for d in range(nd)][:ndb]
rng = np.random.default_rng(4321)
band = (np.arange(nx) * nh // nx)[:, None]
enc = np.minimum(band * nd + rng.integers(0, nd, (nx, nx)), ndb - 1).astype(np.int32).reshape(-1)
order, bounds = native.partition(enc, np.full(ndb, 340200.0), world)
mine = order[bounds[0]:bounds[1]]
n = int(mine.size)
emine = enc[mine]
used = np.unique(emine)
dev = torch.device('cuda', 0)
ctx = native.Context(0)
ctx.set_pruning(True)
sobs, snr = np.empty((n, 8), np.float32), np.empty((n, 8), np.float32)
lut = np.full(ndb, -1, np.int32)
for j, e in enumerate(used):
    h, d = grid[int(e)]
    db = synth.make_database((5 + 3 * h) % 99, (30 + 2 * d) % 121)
    sel = np.flatnonzero(emine == e)
    o = synth.make_observation(sel.size, 1, [db], seed=1000 + int(e))
    sobs[sel], snr[sel] = o['sobs_totrot'].reshape(-1, 8), o['snr'].reshape(-1, 8)
    ctx.db_put(j, db, native.ENC_I16)
    lut[e] = j
ids = lut[emine]
dt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_sobs, d_snr, d_ids = dt(sobs), dt(snr), dt(ids)
zeros = torch.zeros(n, dtype=torch.float32, device=dev)
ones = torch.ones(n, dtype=torch.float32, device=dev)
hdr = synth.make_dbhdr()
g = native.DbGrid.from_dbhdr(hdr)
tabs = [dt(t) for t in g._tables]
g.sin_bphi, g.cos_bphi, g.sin_btheta, g.cos_btheta = (t.data_ptr() for t in tabs)
inv = torch.empty((n, k, 11), dtype=torch.float32, device=dev)
sf = torch.empty((n, k, 8), dtype=torch.float32, device=dev)
st = torch.empty(n, dtype=torch.uint8, device=dev)
iss = torch.empty(n, dtype=torch.int32, device=dev)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(j0, m):
    ctx.invert_dev(native.MODE_IQUV, m, d_sobs.data_ptr() + j0 * 32, d_snr.data_ptr() + j0 * 32, d_ids.data_ptr() + j0 * 4, k,
                   ones.data_ptr(), ones.data_ptr(), ones.data_ptr(), zeros.data_ptr(), None, 0, g, 1000, 0,
                   inv.data_ptr() + j0 * k * 44, sf.data_ptr() + j0 * k * 32, st.data_ptr() + j0, iss.data_ptr() + j0 * 4, stream.cuda_stream)


print('shard of rank 0 of %d: %d pixels, %d databases' % (world, n, used.size))
for pieces in (1, 2, 4, 8):
    per = (n + pieces - 1) // pieces
    for _ in range(2):
        for i in range(pieces):
            run(i * per, min(per, n - i * per))
    torch.cuda.synchronize()
    ts, ks = [], []
    for _ in range(reps):
        flush.zero_()
        ctx.profile_begin()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for i in range(pieces):
            run(i * per, min(per, n - i * per))
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
        ks.append(float(np.sum(ctx.profile_collect())))
    print('%d piece(s) of %6d pixels: %.3f ms per shard (search kernels %.3f ms, everything else %.3f ms)' %
          (pieces, per, np.mean(ts), np.mean(ks), np.mean(ts) - np.mean(ks)), flush=True)
