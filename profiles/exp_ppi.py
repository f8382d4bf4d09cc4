#!/usr/bin/env python
"""Pixels per work item of the pruned search: kernel time for maps of 32 Ki .. 1 Mi pixels (one database) at P = 4, 8, 16, 32
and the library's own choice (0).  Usage: python profiles/exp_ppi.py"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np, torch
from cledb_b200 import native, synth
dev = torch.device('cuda', 0)
ctx = native.Context(0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
db = synth.make_database(9, 40)
ctx.db_put(0, db, native.ENC_I16)
obs = synth.make_observation(1024, 1024, [db], seed=7)
S, N = obs['sobs_totrot'].reshape(-1, 8), obs['snr'].reshape(-1, 8)
k = 8
for npix in (32768, 65536, 98304, 131072, 262144):
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sobs, snr, ids = d(S[:npix]), d(N[:npix]), d(np.zeros(npix, np.int32))
    idx = torch.empty((npix, k), dtype=torch.int32, device=dev); chi = torch.empty((npix, k), dtype=torch.float32, device=dev)
    st = torch.empty(npix, dtype=torch.uint8, device=dev)
    for mode in (0, 1):
        row = []
        for ppi in (0, 1, 2, 3, 4, 6, 8):
            ctx.set_tuning(ppi, 0)
            f = lambda: ctx.match_dev(mode, npix, sobs.data_ptr(), snr.data_ptr(), ids.data_ptr(), k, idx.data_ptr(), chi.data_ptr(), None, None, None, st.data_ptr(), None, stream.cuda_stream)
            f(); f(); torch.cuda.synchronize()
            ctx.profile_begin()
            for _ in range(5): f()
            row.append(float(np.mean(ctx.profile_collect())))
        print('%8d px %s  kernel ms at P = auto/1/2/3/4/6/8: %s' % (npix, 'IQUD' if mode else 'IQUV', '  '.join('%.3f' % v for v in row)), flush=True)
ctx.close()
