#!/usr/bin/env python
"""Where the wall-clock of the reference-shaped Python call CLEDB_PROC.cledb_invproc goes on the GPU box (256x256 map):
per-call times of 30 calls, then the stages of the binding timed one by one on the same inputs."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, 'solar-coronal-inversion_b200')]
import numpy as np
if len(sys.argv) > 2 and sys.argv[2] == 'torch':
    import torch
    torch.zeros(1, device='cuda')
from cledb_b200 import synth, native, solution
import CLEDB_PROC.CLEDB_PROC as procinv

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
procinv.OPTIONS.update(dbencoding=native.ENC_I16)
db = synth.make_database(9, 40)
obs = synth.make_observation(nx, nx, [db], seed=7)
hdr = synth.make_dbhdr()
npix = nx * nx
for mode, iqud, bcalc in (('IQUV', False, 0), ('IQUD', True, 3)):
    ts = []
    iss = obs['issuemask'].copy()
    for rep in range(30):
        t0 = time.perf_counter()
        inv, sf = procinv.cledb_invproc(obs['sobs_totrot'], obs['sobs_dopp'], [db], obs['db_enc'], obs['yobs'], obs['aobs'], obs['dobs'],
                                        obs['snr'], iss, hdr, obs['keyvals'], 8, 1000, bcalc, iqud, 1, False, 0)
        ts.append((time.perf_counter() - t0) * 1e3)
    print('%s %dx%d cledb_invproc ms per call: first %.2f, then median %.3f min %.3f max %.3f' % (mode, nx, nx, ts[0], np.median(ts[2:]), min(ts[2:]), max(ts[2:])))
    print('   ', ' '.join('%.2f' % t for t in ts))
ctx = native.get_context(0)
sobs, snr = obs['sobs_totrot'].reshape(-1, 8), obs['snr'].reshape(-1, 8)
enc = obs['db_enc'].reshape(-1)


def T(name, fn, n=20):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
    print('  %-46s %.3f ms' % (name, (time.perf_counter() - t0) / n * 1e3))
    return r

print('stages of one IQUV call:')
T('_used_databases (bincount)', lambda: procinv._used_databases(enc, 1))
T('_upload (fingerprint of the resident database)', lambda: procinv._upload([db], [0]))
ang = obs['aobs'].reshape(-1)
T('cos/sin(-2 aobs) float32', lambda: (np.cos(-2. * ang), np.sin(-2. * ang)))
T('DbGrid.from_dbhdr', lambda: native.DbGrid.from_dbhdr(hdr))
out = ctx.result_buffers(npix, 8)
T('result_buffers from the pinned pool (4 arrays)', lambda: ctx.result_buffers(npix, 8))
T('ctx.invert, outputs given (pinned)', lambda: ctx.invert(0, sobs, snr, enc, 8, obs['yobs'], obs['dobs'], obs['aobs'], None, hdr, 1000, 0, out=out))
pag = dict(invout=np.empty((npix, 8, 11), np.float32), sfound=np.empty((npix, 8, 8), np.float32), status=np.empty(npix, np.uint8), issue=np.empty((npix, 2), np.uint8))
T('ctx.invert, outputs given (pageable)', lambda: ctx.invert(0, sobs, snr, enc, 8, obs['yobs'], obs['dobs'], obs['aobs'], None, hdr, 1000, 0, out=pag))
T('ctx.invert, fresh pinned outputs', lambda: ctx.invert(0, sobs, snr, enc, 8, obs['yobs'], obs['dobs'], obs['aobs'], None, hdr, 1000, 0))
codes = T('issue_codes', lambda: native.issue_codes(out['issue'], out['invout']))
iss = obs['issuemask'].copy()
T('apply_issue_codes', lambda: solution.apply_issue_codes(iss, codes.reshape(nx, nx)))
T('update_issuemask (NumPy statement, for comparison)', lambda: solution.update_issuemask(iss, out['invout'].reshape(nx, nx, 8, 11), obs['snr'], nx, nx), n=3)
