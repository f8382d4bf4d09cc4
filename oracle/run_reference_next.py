#!/usr/bin/env python
"""Golden vectors for the rows SURVEY.md section 8(f) marks "next", from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

    python oracle/run_reference_next.py        # (re)writes tests/golden/dbselect.npz and tests/golden/reduced_maps.npz

Runs only in the build container (needs /root/reference).  Same recipe as oracle/run_reference.py: astropy stub, scratch
working directory that mirrors the relative paths the reference opens.

dbselect.npz
  * the reference's ``sdb_findelongationanddens`` on a file table shaped like the full 99 x 121 (height, density) grid in the
    order ``sorted(glob(...))`` gives it (``d1000`` before ``d600``), for 4 000 seeded (height, density) pairs including exact
    grid points, mid-points and out-of-range values;
  * the reference's ``sdb_preprocess`` (verbose = 4, its unit-test switch) over a 9 x 8 map on a scratch directory of
    4 heights x 5 densities of small synthetic PyCELP-type files: db_enc, db_enc_flnm, db_uniq, file names and a SHA-256 of
    every normalised database it returns.

integrate.npz / density.npz
  * ``obs_integrate`` on a 5 x 6 map of noisy, shifted copies of the reference's own test spectra
    (tests/sample_test_stokesspectra.npz; pixel [0,0] is the unit-test pixel itself), and ``obs_dens`` on a 7 x 8 map with a
    small synthetic CHIANTI-shaped look-up table.

reduced_maps.npz
  * ``cledb_invproc(..., reduced=True)`` of the reference (IQUV and IQUD, nsearch 8 / 4 / 3) on an 8 x 9 map over two small
    databases, with its ``np.argsort`` order of an all-equal array on this machine (``tie_order``: the phi = 0 plane of the
    presort is one big tie and NumPy's default sort is not stable).
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, 'oracle'))
import run_reference as rr  # noqa: E402  (sets sys.path for the reference + stub, provides the scratch layout)
import numpy as np  # noqa: E402

MINI_HEIGHTS = (106, 109, 112, 120)
MINI_DENS = (600, 795, 895, 1000, 1200)
MINI_SHAPE = dict(ngx=3, nbphi=12, nbtheta=8)


def mini_files(root, synth):
    """4 x 5 small synthetic per-line files DB_h###_d###.npy under root/fe-xiii_107{4,9}/ (un-normalised, V per Angstrom)."""
    rng = np.random.default_rng(4242)
    for h in MINI_HEIGHTS:
        for d in MINI_DENS:
            l1, l2 = synth.make_raw_line_pair(h - 101, (d - 600) // 5, rng, **MINI_SHAPE)
            name = 'DB_h%03d_d%d.npy' % (h, d)
            np.save(os.path.join(root, 'fe-xiii_1074', name), l1)
            np.save(os.path.join(root, 'fe-xiii_1079', name), l2)


def mini_map():
    rng = np.random.default_rng(77)
    nx, ny = 9, 8
    yobs = rng.uniform(1.03, 1.23, (nx, ny)).astype(np.float32)
    dobs = rng.uniform(5.5, 12.5, (nx, ny)).astype(np.float32)
    yobs[0, 0], dobs[0, 0] = np.float32(1.09), np.float32(7.95)        # exact grid point
    yobs[0, 1], dobs[0, 1] = np.float32(1.075), np.float32(8.45)       # mid-points (ties in float32 or not: as NumPy decides)
    yobs[0, 2], dobs[0, 2] = np.float32(1.2), np.float32(12.0)         # the last density of a height is never selected
    yobs[0, 3], dobs[0, 3] = np.float32(0.9), np.float32(3.0)          # below every grid value
    yobs[0, 4], dobs[0, 4] = np.float32(3.0), np.float32(20.0)         # above every grid value
    dobs[0, 5] = np.nan                                                # NaN density: argmin returns the first NaN
    return nx, ny, yobs, dobs


def main():
    import cledb_b200.synth as synth
    rr.reference_first()
    scratch = rr.scratch_layout()
    os.chdir(scratch)
    import CLEDB_PREPINV.CLEDB_PREPINV as prepinv
    import ctrlparams
    assert prepinv.__file__.startswith(rr.REF), prepinv.__file__
    out = {}

    # ---- selection on a full-grid-shaped table
    names = sorted('DB_h%03d_d%d.npy' % (101 + h, 600 + 5 * d) for h in range(99) for d in range(121))
    at = names[0].find('DB_h')
    table = np.empty((len(names), 3), dtype=np.float32)
    for i, n in enumerate(names):
        table[i] = (i, np.float32(n[at + 4:at + 7]), np.float32(n[at + 9:-4]))
    rng = np.random.default_rng(31)
    ny = 4000
    ys = rng.uniform(0.95, 2.05, ny).astype(np.float32)
    ds = rng.uniform(5.5, 12.5, ny).astype(np.float32)
    ys[:200] = (np.float32(1.01) + np.float32(0.01) * rng.integers(0, 99, 200)).astype(np.float32)     # on the grid
    ds[100:300] = (np.float32(6.0) + np.float32(0.05) * rng.integers(0, 121, 200)).astype(np.float32)
    ys[300:400] = (np.float32(1.015) + np.float32(0.01) * rng.integers(0, 98, 100)).astype(np.float32)  # between two heights
    ds[400:500] = (np.float32(6.025) + np.float32(0.05) * rng.integers(0, 120, 100)).astype(np.float32)
    sel = np.array([prepinv.sdb_findelongationanddens(0, 0, ys[i], ds[i], table)[2] for i in range(ny)], dtype=np.int32)
    out.update(grid_names=np.array(names), grid_table=table, grid_y=ys, grid_d=ds, grid_sel=sel)
    print('grid selection: %d pairs, %d distinct files' % (ny, np.unique(sel).size))

    # ---- sdb_preprocess on the mini grid
    for line in ('fe-xiii_1074', 'fe-xiii_1079'):
        for f in os.listdir(os.path.join('tests', 'dbfiles', line)):
            if f.endswith('.npy'):
                os.remove(os.path.join('tests', 'dbfiles', line, f))
    mini_files(os.path.join(scratch, 'tests', 'dbfiles'), synth)
    nx, nyy, yobs, dobs = mini_map()
    params = ctrlparams.ctrlparams()
    params.verbose = 4
    params.ncpu = 4
    keyvals = synth.make_keyvals(nx, nyy)
    wlarr = np.array([[1074.2571372 + 0.00538654 * i for i in range(4)], [1079.4203968 + 0.00541243 * i for i in range(4)]])
    db_enc, database, dbhdr, dbnames, db_enc_flnm, db_uniq = prepinv.sdb_preprocess(yobs, dobs, keyvals, wlarr, params)
    out.update(mini_yobs=yobs, mini_dobs=dobs, mini_db_enc=db_enc, mini_db_enc_flnm=db_enc_flnm, mini_db_uniq=db_uniq,
               mini_names=np.array([os.path.basename(n) for n in dbnames[0]]), mini_wlarr=wlarr, mini_hdr_g=dbhdr[0],
               mini_sha=np.array([rr.sha(np.asarray(d)) for d in database]),
               mini_shape=np.array(np.asarray(database[0]).shape))
    print('sdb_preprocess: %d distinct databases of shape %s, db_enc[0] = %s' % (len(database), np.asarray(database[0]).shape, db_enc[0]))
    np.savez_compressed(os.path.join(rr.GOLDEN, 'dbselect.npz'), **out)

    # ---- reduced-database mode (reduced=True) of cledb_invproc, small databases
    import CLEDB_PROC.CLEDB_PROC as procinv
    assert procinv.__file__.startswith(rr.REF), procinv.__file__
    small = synth.make_database(8, 39, ngx=5, nbphi=36, nbtheta=18)
    small2 = synth.make_database(20, 70, ngx=5, nbphi=36, nbtheta=18)
    hdr = prepinv.sdb_parseheader('./tests/dbfiles/fe-xiii_1074/db.hdr')
    gg = hdr[0].copy()
    gg[2], gg[3], gg[4] = 5, 36, 18
    hdr = (gg, hdr[1], np.int32(5), np.int32(36), np.int32(18)) + tuple(hdr[5:])
    dbs = [small, small2]
    nx, ny = 8, 9
    enc = (np.arange(nx * ny).reshape(nx, ny) % 2).astype(np.int32)
    obs = synth.make_observation(nx, ny, dbs, db_enc=enc, seed=9, frac_nan=0.04, frac_zero=0.03)
    rng = np.random.default_rng(5)
    obs['sobs_dopp'][:, :, 1] = rng.uniform(-1.5, 1.5, (nx, ny)).astype(np.float32)        # wave angle, used only when reduced
    s = obs['sobs_totrot']
    s[1, 1, 2] = 0.0                                   # U = 0: Phi_B = 0 or pi/2
    s[1, 2, 1] = 0.0                                   # Q = 0
    obs['sobs_dopp'][2, 2, 1] = 0.0                    # wave angle 0: tan = 0, every theta ties at tan(theta)*sin(phi) ~ 0 rows
    red = {}
    for tag, iqud, bcalc, k in (('iquv_k8', False, 0, 8), ('iquv_k4', False, 2, 4), ('iqud_k8', True, 3, 8), ('iqud_k3', True, 3, 3)):
        ss = s.copy()
        if iqud:
            ss[np.isnan(ss).any(axis=2) & ~np.isnan(ss).all(axis=2)] = np.nan
        iss = obs['issuemask'].copy()
        inv, sf = procinv.cledb_invproc(ss, obs['sobs_dopp'], dbs, enc, obs['yobs'], obs['aobs'], obs['dobs'], obs['snr'], iss, hdr,
                                        obs['keyvals'], k, 1000, bcalc, iqud, 6, True, 0)
        red['invout_' + tag], red['sfound_' + tag], red['issuemask_' + tag], red['sobs_' + tag] = inv, sf, iss, ss
        red['cfg_' + tag] = np.array([int(iqud), bcalc, k, 1000], dtype=np.float64)
        print('reduced', tag, 'searched', int((inv[..., 0, 0] >= 0).sum()), 'idx[0,0]', inv[0, 0, :, 0])
    # one pixel's subset, for debugging the device-side construction
    dsel, ored = procinv.cledb_getsubsetiquv(s[0, 0], hdr, dbs[enc[0, 0]], 8)
    np.savez_compressed(os.path.join(rr.GOLDEN, 'reduced_maps.npz'), db0=small, db1=small2, db_enc=enc, hdr_g=hdr[0],
                        sobs_dopp=obs['sobs_dopp'], snr=obs['snr'], yobs=obs['yobs'], aobs=obs['aobs'], dobs=obs['dobs'],
                        cdelt=np.float64(obs['keyvals'][11]), tie_order=np.argsort(np.zeros(18)).astype(np.int32),
                        outred_00=ored, datasel_00_sha=rr.sha(dsel), **red)
    # ---- spectral integration (obs_integrate) on a small map of noisy spectra built from the reference's own test profiles
    data = np.load('./tests/sample_test_stokesspectra.npz')
    rng = np.random.default_rng(17)
    nx, ny, nw = 5, 6, 134
    cubes = []
    for key, wkey in (('aas', 'awl'), ('bbs', 'bwl')):
        base = data[key]                                                   # [134,4] float64, the profile tests/test_functions.py uses
        amp = rng.uniform(0.3, 3.0, (nx, ny, 1, 1))
        shift = rng.integers(-12, 13, (nx, ny))
        cube = np.empty((nx, ny, nw, 4))
        for xx in range(nx):
            for yy in range(ny):
                cube[xx, yy] = np.roll(base, shift[xx, yy], axis=0) * amp[xx, yy, 0]
        noise = rng.normal(0.0, 1.0, cube.shape) * np.abs(base[:, 0]).max() * np.array([2e-3, 2e-3, 2e-3, 1e-4])
        cube = cube + noise * rng.uniform(0.0, 3.0, (nx, ny, 1, 1))
        cube[0, 0] = base                                                  # the reference's own unit-test pixel, untouched
        cube[0, 1] = 0.0                                                   # no counts
        cube[0, 2, 5, 1] = np.nan                                          # a NaN sample
        cube[0, 3] = base * 1e-3 + noise[0, 3] * 30                        # noise dominated: SNR < 1 issue codes
        cubes.append(cube)
    wl = np.zeros((nw, 2), dtype=np.float32)                               # as obs_headstructproc builds it (CLEDB_PREPINV.py:720, 852)
    for i, (crval3, cdelt3) in enumerate(((1074.2571372, 0.00538654), (1079.4203968, 0.00541243))):
        for j in range(nw):
            wl[j, i] = crval3 + (j * cdelt3)
    kv = synth.make_keyvals(nx, ny)
    kv[2] = nw
    tot, snr, bkg, iss = prepinv.obs_integrate(cubes, wl, kv, 0, 4, False)
    print('obs_integrate: tot[0,0]', tot[0, 0], 'issues', np.unique(iss))
    np.savez_compressed(os.path.join(rr.GOLDEN, 'integrate.npz'), cube0=cubes[0], cube1=cubes[1], wl=wl, sobs_tot=tot, snr=snr,
                        background=bkg, issuemask=iss, aat=data['aat'], bbt=data['bbt'])
    # ---- density from the line ratio (obs_dens) with a small synthetic CHIANTI-shaped look-up table
    nh, nd = 12, 40
    tab_h = 1.02 + 0.08 * np.arange(nh)
    tab_den = 10 ** np.linspace(6.0, 12.0, nd)
    xx_ = np.linspace(0.0, 1.0, nd)
    tab_rat = np.stack([0.9 + (14.0 + 0.5 * i) * (1 - xx_) ** (1.5 + 0.05 * i) for i in range(nh)])       # decreasing with density
    np.savez('lookup_small.npz', h=tab_h, den=tab_den, rat=tab_rat)
    rng = np.random.default_rng(23)
    nx, ny = 7, 8
    st = np.zeros((nx, ny, 8), dtype=np.float32)
    st[:, :, 4] = rng.uniform(0.5e-9, 1.5e-9, (nx, ny))
    st[:, :, 0] = st[:, :, 4] * rng.uniform(0.5, 18.0, (nx, ny))                                         # ratios inside and outside the table
    yo = rng.uniform(1.0, 2.1, (nx, ny)).astype(np.float32)                                              # some above the last table height
    st[0, 0, 4] = 0.0
    st[0, 1, 0] = np.nan
    st[0, 2, 0] = -st[0, 2, 0]
    params = ctrlparams.ctrlparams()
    params.verbose = 0
    params.ncpu = 4
    params.lookuptb = './lookup_small.npz'
    dobs_ref, issd = prepinv.obs_dens(st, yo, synth.make_keyvals(nx, ny), params)
    print('obs_dens: dobs[0]', dobs_ref[0], 'issues', np.unique(issd))
    np.savez_compressed(os.path.join(rr.GOLDEN, 'density.npz'), h=tab_h, den=tab_den, rat=tab_rat, sobs=st, yobs=yo, dobs=dobs_ref, iss=issd)
    import shutil
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == '__main__':
    main()
