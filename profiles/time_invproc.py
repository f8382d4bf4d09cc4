#!/usr/bin/env python
"""Wall-clock of the reference-shaped Python call CLEDB_PROC.cledb_invproc on the GPU box (256x256 map by default),
split into the C-ABI search and the host-side solution building."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np
from cledb_b200 import synth, native, solution
import CLEDB_PROC.CLEDB_PROC as procinv

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
procinv.OPTIONS.update(dbencoding=native.ENC_I16)
db = synth.make_database(9, 40)
obs = synth.make_observation(nx, nx, [db], seed=7)
hdr = synth.make_dbhdr()
for mode, iqud, bcalc in (('IQUV', False, 0), ('IQUD', True, 3)):
    for rep in range(3):
        iss = obs['issuemask'].copy()
        t0 = time.perf_counter()
        inv, sf = procinv.cledb_invproc(obs['sobs_totrot'], obs['sobs_dopp'], [db], obs['db_enc'], obs['yobs'], obs['aobs'], obs['dobs'],
                                        obs['snr'], iss, hdr, obs['keyvals'], 8, 1000, bcalc, iqud, 1, False, 0)
        t1 = time.perf_counter()
    ctx = native.get_context(0)
    sobs = obs['sobs_totrot'].reshape(-1, 8); snr = obs['snr'].reshape(-1, 8)
    t2 = time.perf_counter()
    raw = ctx.match(1 if iqud else 0, sobs, snr, np.zeros(nx * nx, np.int32), 8)
    t3 = time.perf_counter()
    print('%s %dx%d: cledb_invproc %.1f ms (%.3e px/s); of which C-ABI search %.1f ms' % (mode, nx, nx, (t1 - t0) * 1e3, nx * nx / (t1 - t0), (t3 - t2) * 1e3))
import cProfile, pstats
iss = obs['issuemask'].copy()
pr = cProfile.Profile(); pr.enable()
procinv.cledb_invproc(obs['sobs_totrot'], obs['sobs_dopp'], [db], obs['db_enc'], obs['yobs'], obs['aobs'], obs['dobs'],
                      obs['snr'], iss, hdr, obs['keyvals'], 8, 1000, 0, False, 1, False, 0)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
