#!/usr/bin/env python
"""End-to-end time of cledb_invert (pinned host buffers) on the 256x256 and 512x512 maps as a function of the pipeline chunk
($CLEDB_B200_CHUNK pixels per chunk; search of chunk c+1 under the device-to-host copy of chunk c)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np, torch
from cledb_b200 import native, synth
ctx = native.Context(0)
db = synth.make_database(9, 40)
ctx.db_put(0, db, native.ENC_I16)
hdr = synth.make_dbhdr()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
for nx in (256, 512):
    obs = synth.make_observation(nx, nx, [db], seed=7)
    npix, k = nx * nx, 8
    out = ctx.result_buffers(npix, k)
    s_h, n_h, ids = pin(obs['sobs_totrot'].reshape(-1, 8)), pin(obs['snr'].reshape(-1, 8)), pin(np.zeros(npix, np.int32))
    y, d, a = pin(obs['yobs'].reshape(-1)), pin(obs['dobs'].reshape(-1)), obs['aobs'].reshape(-1)
    ref = None
    for chunk in (0, npix // 2, (npix + 2) // 3, npix // 4, npix // 8):
        if chunk:
            os.environ['CLEDB_B200_CHUNK'] = str(chunk)
        else:
            os.environ.pop('CLEDB_B200_CHUNK', None)
        for mode, bcalc in ((0, 0), (1, 3)):
            dopp = pin(obs['sobs_dopp'].reshape(npix, -1)[:, 0]) if mode else None
            f = lambda: ctx.invert(mode, s_h, n_h, ids, k, y, d, a, dopp, hdr, 1000, bcalc, out=out)
            for _ in range(3): f()
            t0 = time.perf_counter()
            for _ in range(20): f()
            ms = (time.perf_counter() - t0) / 20 * 1e3
            chk = int(out['invout'].view(np.int32).astype(np.int64).sum())
            print('%4d^2 %s chunk %7s: %.3f ms  %.1f Mpx/s  checksum %x' % (nx, 'IQUD' if mode else 'IQUV', chunk or 'default', ms, npix / ms / 1e3, chk & 0xffffffffff), flush=True)
ctx.close()
