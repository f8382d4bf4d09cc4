#!/usr/bin/env python
"""Small driver for ncu: BASELINE config 2 (256x256 IQUV map, one database, int16) through cledb_invert_dev (search, k_finalize,
k_build) followed by cledb_gather_maps_dev on a one-rank communicator (k_push_records storing every record at its pixel in
the window).  Usage (GPU box):  ncu --set full --clock-control none --import-source on -k regex:k_build -s 2 -c 1 -o out python profiles/prof_build.py [nx]"""
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from cledb_b200 import native, synth  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
db = synth.make_database(9, 40)
obs = synth.make_observation(nx, nx, [db], seed=7)
ctx = native.Context(0)
ctx.db_put(0, db, native.ENC_I16)
ctx.comm_init(0, 1)
dev = torch.device('cuda', 0)
npix, k = nx * nx, 8
dt = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
sobs, snr = dt(obs['sobs_totrot'].reshape(-1, 8)), dt(obs['snr'].reshape(-1, 8))
ids = torch.zeros(npix, dtype=torch.int32, device=dev)
ang = obs['aobs'].reshape(-1)
y, d, rc, rs = dt(obs['yobs'].reshape(-1)), dt(obs['dobs'].reshape(-1)), dt(np.cos(-2. * ang)), dt(np.sin(-2. * ang))
g = native.DbGrid.from_dbhdr(synth.make_dbhdr())
tabs = [dt(t) for t in g._tables]
g.sin_bphi, g.cos_bphi, g.sin_btheta, g.cos_btheta = (t.data_ptr() for t in tabs)
lib = native.load_library()
lay, mlay = np.zeros(5, np.int64), np.zeros(5, np.int64)
lib.cledb_shard_layout(npix, k, lay.ctypes.data_as(ctypes.c_void_p))
lib.cledb_map_layout(npix, k, mlay.ctypes.data_as(ctypes.c_void_p))
send = torch.empty(int(lay[4]), dtype=torch.uint8, device=dev)
sp = send.data_ptr()
win = ctx.comm_window(int(mlay[4]))
order = torch.from_numpy(np.random.default_rng(1).permutation(npix).astype(np.int32)).to(dev)
s = torch.cuda.Stream(device=dev)
torch.cuda.synchronize()
for i in range(4):
    ctx.invert_dev(native.MODE_IQUV, npix, sobs.data_ptr(), snr.data_ptr(), ids.data_ptr(), k, y.data_ptr(), d.data_ptr(), rc.data_ptr(),
                   rs.data_ptr(), None, 0, g, 1000, 0, sp + int(lay[0]), sp + int(lay[1]), sp + int(lay[2]), sp + int(lay[3]), s.cuda_stream)
    ctx.gather_maps_dev(npix, npix, order.data_ptr(), npix, k, sp, win, -1, s.cuda_stream)
torch.cuda.synchronize()
print('done: %d pixels, %d launches' % (npix, ctx.launch_count()))
