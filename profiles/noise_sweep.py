#!/usr/bin/env python
"""How the exact pruning degrades when observations fit no database entry well: the 256x256 map of the bench with every
Stokes component multiplied by (1 + s*N(0,1)); kernel time and evaluated share, pruned vs the two brute-force variants
(all three return the same bits: asserted).
Usage (GPU box): python profiles/noise_sweep.py"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from cledb_b200 import native, synth  # noqa: E402

nx = 256
db = synth.make_database(9, 40)
obs = synth.make_observation(nx, nx, [db], seed=7)
ctx = native.Context(0)
dev = torch.device('cuda', 0)
npix, k = nx * nx, 8
rng = np.random.default_rng(3)
base = obs['sobs_totrot'].reshape(-1, 8)
snr = torch.from_numpy(obs['snr'].reshape(-1, 8)).to(dev)
ids = torch.zeros(npix, dtype=torch.int32, device=dev)
idx = torch.empty((npix, k), dtype=torch.int32, device=dev)
chi = torch.empty((npix, k), dtype=torch.float32, device=dev)
st = torch.empty(npix, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream(device=dev)
for mode, name in ((native.MODE_IQUV, 'IQUV'), (native.MODE_IQUD, 'IQUD')):
    for scale in (0.0, 0.02, 0.1, 0.3, 1.0):
        pert = (base * (1.0 + scale * rng.standard_normal(base.shape))).astype(np.float32)
        sobs = torch.from_numpy(pert).to(dev)
        row = []
        keep = None
        for prune in (1, 2, 0):
            ctx.set_pruning(prune)
            ctx.db_put(0, db, native.ENC_I16)
            torch.cuda.synchronize()
            ctx.profile_begin()
            for i in range(4):
                ctx.match_dev(mode, npix, sobs.data_ptr(), snr.data_ptr(), ids.data_ptr(), k, idx.data_ptr(), chi.data_ptr(),
                              status_ptr=st.data_ptr(), stream=s.cuda_stream)
            torch.cuda.synchronize()
            ms = float(np.mean(ctx.profile_collect()[1:]))
            stats = ctx.pruning_stats() if prune == 1 else None
            cur = (idx.cpu().numpy().copy(), chi.cpu().numpy().copy())
            if keep is not None:
                assert np.array_equal(keep[0], cur[0]) and np.array_equal(keep[1], cur[1], equal_nan=True)
            keep = cur
            row.append((ms, stats))
            ctx.db_drop(0)
        ev = row[0][1]
        print('%s perturbation %.2f: pruned %.3f ms (%.1f %% of the (pixel, tile) pairs evaluated), brute force seeded %.3f ms, brute force plain %.3f ms, '
              'median chi2 of the best match %.3g' % (name, scale, row[0][0], 100.0 * ev['evaluated'] / max(1, ev['evaluated'] + ev['skipped']),
                                                      row[1][0], row[2][0], float(np.nanmedian(keep[1][:, 0]))))
