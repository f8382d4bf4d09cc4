#!/usr/bin/env python
"""Golden vectors for CLE-type databases (dbtype 0) from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

    python oracle/run_reference_cle.py            # (re)writes tests/golden/cle_type.npz

A CLE-type database carries the electron density as its leading axis - ``[ned, ngx, nbphi, nbtheta, 8]`` - and the
solution's density comes from the Baumbach profile scaled by ``xed[i]`` (``cledb_elecdens``, CLEDB_PROC.py:1688-1697,
``cledb_physcle`` :1744-1747) instead of from the observation.  CLE databases are deprecated upstream, but the branch is
on the path, so it is pinned here like the PyCELP one: the reference's own ``cledb_invproc``
runs on a small seeded map in IQUV (bcalc 0, 2) and IQUD (bcalc 3) mode.
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'oracle'))
import run_reference as rr  # noqa: E402  (sets up sys.path: stubs, /root/reference, product for the input generator)

import numpy as np  # noqa: E402


def main():
    import cledb_b200.synth as synth
    rr.reference_first()
    scratch = rr.scratch_layout()
    os.chdir(scratch)
    import CLEDB_PROC.CLEDB_PROC as procinv
    assert procinv.__file__.startswith(rr.REF), procinv.__file__

    ned, ngx, nbphi, nbtheta = 3, 4, 12, 8
    # one PyCELP-shaped block per density index, stacked along the leading axis
    db = np.stack([synth.make_database(7 + 2 * i, 35 + 9 * i, ngx=ngx, nbphi=nbphi, nbtheta=nbtheta) for i in range(ned)])
    assert db.shape == (ned, ngx, nbphi, nbtheta, 8)
    hdr = synth.make_dbhdr(ngx, nbphi, nbtheta, dbtype=0)
    g = hdr[0].copy()
    g[1] = ned
    xed = np.array([1.0, 2.5, 6.0], dtype=np.float32)                # density scalings of the three blocks
    hdr = (g, np.int32(ned)) + tuple(hdr[2:5]) + (xed,) + tuple(hdr[6:])
    nx, ny = 6, 5
    flat_db = db.reshape(ned * ngx, nbphi, nbtheta, 8)               # the observation generator wants [.., nbphi, nbtheta, 8]
    obs = synth.make_observation(nx, ny, [flat_db], seed=77, frac_nan=0.07, frac_zero=0.04)
    kv = obs['keyvals']
    out = dict(db=db, hdr_g=g, xed=xed, sobs=obs['sobs_totrot'], sobs_dopp=obs['sobs_dopp'], snr=obs['snr'], yobs=obs['yobs'],
               aobs=obs['aobs'], dobs=obs['dobs'], cdelt=np.float64(kv[11]))
    for tag, iqud, bcalc, k in (('iquv_b0', False, 0, 6), ('iquv_b2', False, 2, 3), ('iqud', True, 3, 6)):
        iss = np.zeros((nx, ny, 2), np.int32)
        inv, sf = procinv.cledb_invproc(obs['sobs_totrot'], obs['sobs_dopp'], [db], obs['db_enc'], obs['yobs'], obs['aobs'],
                                        obs['dobs'], obs['snr'], iss, hdr, kv, k, 1000, bcalc, iqud, 2, False, 0)
        out['invout_' + tag], out['sfound_' + tag], out['issuemask_' + tag] = inv, sf, iss
        out['cfg_' + tag] = np.array([int(iqud), bcalc, k, 1000])
        print(tag, 'solutions:', int((inv[..., 0] >= 0).sum()), 'density column range', float(inv[..., 2].min()), float(inv[..., 2].max()),
              'density index of the winners:', sorted(set((inv[..., 0][inv[..., 0] >= 0] // (ngx * nbphi * nbtheta)).astype(int).tolist())))
    np.savez_compressed(os.path.join(rr.GOLDEN, 'cle_type.npz'), **out)


if __name__ == '__main__':
    main()
