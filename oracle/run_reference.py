#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Runs only in the build container, where the reference checkout exists at /root/reference (read-only):

    python oracle/run_reference.py            # (re)writes tests/golden/*.npz

What it does
  * puts a 3-file ``astropy`` stub (oracle/ref_stubs) and /root/reference on sys.path and imports the
    reference's CLEDB_PREPINV / CLEDB_PROC / constants / ctrlparams as they are;
  * chdir()s into a scratch directory that mirrors the relative paths the reference opens
    (./tests/dbfiles/, ./tests/sample_test_stokesspectra.npz, ./CLEDB_BUILD/*.npz);
  * runs the reference's own unit-test chain (tests/test_functions.py) and a set of seeded synthetic maps
    through ``cledb_invproc`` (IQUV and IQUD), ``blos_proc`` and ``obs_qurotate``/``obs_calcheight``;
  * stores inputs + reference outputs as small compressed .npz fixtures.

Nothing here is imported by the product, the GPU tests, smoke() or bench.py.
"""
import hashlib
import os
import shutil
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
GOLDEN = os.path.join(REPO, 'tests', 'golden')

sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REPO, 'oracle', 'ref_stubs'))
sys.path.insert(0, os.path.join(REPO, 'solar-coronal-inversion_b200'))

import numpy as np  # noqa: E402


def scratch_layout():
    """Scratch cwd with the files the reference resolves relative to the working directory."""
    d = tempfile.mkdtemp(prefix='cledb_ref_')
    os.makedirs(os.path.join(d, 'tests', 'dbfiles', 'fe-xiii_1074'))
    os.makedirs(os.path.join(d, 'tests', 'dbfiles', 'fe-xiii_1079'))
    os.makedirs(os.path.join(d, 'CLEDB_BUILD'))
    os.symlink(os.path.join(REF, 'tests', 'sample_test_stokesspectra.npz'),
               os.path.join(d, 'tests', 'sample_test_stokesspectra.npz'))
    for f in os.listdir(os.path.join(REF, 'CLEDB_BUILD')):
        if f.endswith('.npz'):
            os.symlink(os.path.join(REF, 'CLEDB_BUILD', f), os.path.join(d, 'CLEDB_BUILD', f))
    for line in ('fe-xiii_1074', 'fe-xiii_1079'):
        shutil.copy(os.path.join(REF, 'tests', 'dbfiles', line, 'db.hdr'),
                    os.path.join(d, 'tests', 'dbfiles', line, 'db.hdr'))
    return d


def raw_line_files(d, synth):
    """The six DB_h*_d*.npy blobs are missing from the checkout (.MISSING_LARGE_BLOBS): write synthetic
    un-normalised per-line IQUV files of the same names/shape so the reference's sdb_preprocess can run."""
    rng = np.random.default_rng(99)
    for hi, name in ((5, 'DB_h106_d895.npy'), (8, 'DB_h109_d795.npy'), (8, 'DB_h109_d895.npy')):
        di = 59 if 'd895' in name else 39
        l1, l2 = synth.make_raw_line_pair(hi, di, rng)   # files hold V per Angstrom; x0.1 happens on load
        np.save(os.path.join(d, 'tests', 'dbfiles', 'fe-xiii_1074', name), l1)
        np.save(os.path.join(d, 'tests', 'dbfiles', 'fe-xiii_1079', name), l2)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def reference_first():
    """The product mirrors the reference's module names (CLEDB_PROC, CLEDB_PREPINV, constants, ctrlparams): once the
    input generator is imported, take the product directory off sys.path so that those names resolve to /root/reference."""
    pkg = os.path.join(REPO, 'solar-coronal-inversion_b200')
    sys.path[:] = [p for p in sys.path if os.path.abspath(p) != pkg]
    for name in list(sys.modules):
        if name.split('.')[0] in ('CLEDB_PROC', 'CLEDB_PREPINV', 'constants', 'ctrlparams'):
            del sys.modules[name]


def main():
    import cledb_b200.synth as synth
    reference_first()
    scratch = scratch_layout()
    os.chdir(scratch)
    raw_line_files(scratch, synth)

    import CLEDB_PROC.CLEDB_PROC as procinv
    import CLEDB_PREPINV.CLEDB_PREPINV as prepinv
    import constants as consts
    import ctrlparams
    assert procinv.__file__.startswith(REF), procinv.__file__
    params = ctrlparams.ctrlparams()
    params.verbose = 0
    os.makedirs(GOLDEN, exist_ok=True)

    # ------------------------------------------------------------------ case 1: the reference's unit-test pixel
    data = np.load('./tests/sample_test_stokesspectra.npz')
    sobs_in = [data['aas'].reshape(1, 1, 134, 4), data['bbs'].reshape(1, 1, 134, 4)]
    head_in = [{'CRPIX1': 0, 'CRPIX2': 0, 'CRPIX3': 0, 'CRVAL1': -0.30, 'CRVAL2': 1.05, 'CRVAL3': 1074.2571372,
                'CDELT1': 0.0001, 'CDELT2': 0.0001, 'CDELT3': 0.00538654, 'LINEWAV': 1074.6153, 'INSTRUME': 'PyCELP'},
               {'CRPIX1': 0, 'CRPIX2': 0, 'CRPIX3': 0, 'CRVAL1': -0.30, 'CRVAL2': 1.05, 'CRVAL3': 1079.4203968,
                'CDELT1': 0.0001, 'CDELT2': 0.0001, 'CDELT3': 0.00541243, 'LINEWAV': 1079.7803, 'INSTRUME': 'PyCELP'}]
    sobs_tot, yobs, snr, background, issuemask, wlarr, keyvals, sobs_totrot, aobs, dobs = \
        prepinv.sobs_preprocess(sobs_in, head_in, params)
    params.verbose = 4
    db_enc, database, dbhdr, dbnames, db_enc_fln, db_u = prepinv.sdb_preprocess(yobs, dobs, keyvals, wlarr, params)
    params.verbose = 0
    iss_blos = issuemask.copy()
    blosout = procinv.blos_proc(sobs_tot, snr, iss_blos, keyvals, consts, params)
    iss_v = issuemask.copy()
    inv_v, sf_v = procinv.cledb_invproc(sobs_totrot, 0, database, db_enc, yobs, aobs, dobs, snr, iss_v, dbhdr,
                                        keyvals, params.nsearch, params.maxchisq, 0, False, params.ncpu, False, 0)
    iss_d = issuemask.copy()
    dopp = np.array((4.60)).reshape(1, 1, 1)
    inv_d, sf_d = procinv.cledb_invproc(sobs_totrot, dopp, database, db_enc, yobs, aobs, dobs, snr, iss_d, dbhdr,
                                        keyvals, params.nsearch, params.maxchisq, 3, True, params.ncpu, False, 0)
    raw1 = np.load(dbnames[0][db_u[0]])
    raw2 = np.load(dbnames[1][db_u[0]])
    np.savez_compressed(
        os.path.join(GOLDEN, 'unit_pixel.npz'),
        sobs_tot=sobs_tot, sobs_totrot=sobs_totrot, snr=snr, yobs=yobs, aobs=aobs, dobs=dobs,
        issuemask_in=issuemask, keyvals_num=np.array([keyvals[0], keyvals[1], keyvals[2], keyvals[3]]),
        crval=np.array([keyvals[8], keyvals[9]], dtype=np.float64), cdelt=np.array([keyvals[11], keyvals[12]], dtype=np.float64),
        hdr_g=dbhdr[0], db_u=db_u, db_file=os.path.basename(dbnames[0][db_u[0]]),
        db_sha=sha(np.asarray(database[0])), raw1_sha=sha(raw1), raw2_sha=sha(raw2), wl0=np.float64(wlarr[0, 0]),
        db_hd=np.array([8, 39]),
        blosout=blosout, issuemask_blos=iss_blos,
        invout_iquv=inv_v, sfound_iquv=sf_v, issuemask_iquv=iss_v,
        dopp=dopp, invout_iqud=inv_d, sfound_iqud=sf_d, issuemask_iqud=iss_d)
    print('unit_pixel: idx', inv_v[0, 0, :, 0], 'B', inv_v[0, 0, 0, 5], 'blos', blosout[0, 0])

    # ------------------------------------------------------------------ case 2: small-DB maps with edge cases
    small = synth.make_database(8, 39, ngx=5, nbphi=36, nbtheta=18)
    hdr_small = prepinv.sdb_parseheader('./tests/dbfiles/fe-xiii_1074/db.hdr')
    g = hdr_small[0].copy()
    g[2], g[3], g[4] = 5, 36, 18
    hdr_small = (g, hdr_small[1], np.int32(5), np.int32(36), np.int32(18)) + tuple(hdr_small[5:])
    small2 = synth.make_database(20, 70, ngx=5, nbphi=36, nbtheta=18)
    # quirk / tie database: first rows are near-copies of far rows so that one of the first nsearch kept
    # rows ranks inside the top-nsearch out of turn; plus exact duplicate rows (ties)
    quirk = small.copy().reshape(-1, 8)
    src = [1500, 1501, 2100, 2101, 1502, 2102]
    for dst, s in enumerate(src):
        quirk[dst] = quirk[s] * np.float32(1.0 + 1e-4 * (dst + 1))
        quirk[dst, 0] = 1.0
    quirk[2000:2010] = quirk[1990:2000]          # exact duplicates -> exact chi^2 ties
    quirk = quirk.reshape(small.shape)
    dbs = [small, small2, quirk]
    nx, ny = 10, 12
    enc = (np.arange(nx * ny).reshape(nx, ny) % 3).astype(np.int32)
    obs = synth.make_observation(nx, ny, dbs, db_enc=enc, seed=7, frac_nan=0.03, frac_zero=0.03)
    s = obs['sobs_totrot']
    flatq = quirk.reshape(-1, 8)
    # engineered pixels (all on the quirk database, enc == 2)
    qpix = [(xx, yy) for xx in range(nx) for yy in range(ny) if enc[xx, yy] == 2]
    for n_, (xx, yy) in zip(src + [1995, 2005, 1995], qpix):
        row = flatq[n_].astype(np.float64)
        row[3] *= 7.0
        row[7] *= 7.0
        s[xx, yy] = (row * 2e-9).astype(np.float32)
    s[0, 1, 3] = np.nan                           # partly-NaN pixel (IQUV: null; IQUD: reference raises -> excluded there)
    s[0, 4, 5] = 0.0                              # one zero component: sigma = 0 -> 0/0 and x/0 paths
    s[1, 1, 3] = 0.0                              # V == 0: sign 0 keeps no rows (IQUV: reference raises)
    obs['snr'][2, 2, 1] = 0.0                     # zero SNR on one component -> sigma = inf -> residual 0
    obs['snr'][3, 3, :] = 0.0                     # integrated=True style all-zero SNR: all chi^2 equal
    obs['snr'][4, 4, 0] = 0.5                     # snr < 1 -> issue 2048
    cases = {}
    ok_iquv = np.ones((nx, ny), bool)
    ok_iqud = np.ones((nx, ny), bool)
    ok_iquv[1, 1] = False                         # reference raises ValueError (argmin of empty)
    ok_iqud[0, 1] = False                         # reference raises IndexError (max is NaN)

    def run_ref(mode_iqud, bcalc, nsearch, maxchisq, okmask, tag):
        # the reference cannot survive pixels that make it raise; feed those as all-NaN (null) pixels
        ss = s.copy()
        ss[~okmask] = np.nan
        iss = obs['issuemask'].copy()
        inv, sf = procinv.cledb_invproc(ss, obs['sobs_dopp'], dbs, enc, obs['yobs'], obs['aobs'], obs['dobs'],
                                        obs['snr'], iss, hdr_small, obs['keyvals'], nsearch, maxchisq, bcalc,
                                        mode_iqud, 6, False, 0)
        cases['sobs_' + tag] = ss
        cases['invout_' + tag] = inv
        cases['sfound_' + tag] = sf
        cases['issuemask_' + tag] = iss
        cases['cfg_' + tag] = np.array([int(mode_iqud), bcalc, nsearch, maxchisq], dtype=np.float64)
        return inv

    inv = run_ref(False, 0, 8, 1000, ok_iquv, 'iquv_b0')
    run_ref(False, 1, 8, 1000, ok_iquv, 'iquv_b1')
    run_ref(False, 2, 4, 1000, ok_iquv, 'iquv_b2_k4')
    run_ref(False, 0, 8, 3.0, ok_iquv, 'iquv_b0_chi3')
    run_ref(False, 0, 16, 1000, ok_iquv, 'iquv_b0_k16')
    invd = run_ref(True, 3, 8, 1000, ok_iqud, 'iqud')
    run_ref(True, 3, 8, 3.0, ok_iqud, 'iqud_chi3')
    run_ref(True, 0, 8, 1000, ok_iqud, 'iqud_badbcalc')     # guard CLEDB_PROC.py:174
    ndup = sum(len(set(inv[xx, yy, :, 0][inv[xx, yy, :, 0] >= 0])) < (inv[xx, yy, :, 0] >= 0).sum()
               for xx in range(nx) for yy in range(ny))
    ndupd = sum(len(set(invd[xx, yy, :, 0][invd[xx, yy, :, 0] >= 0])) < (invd[xx, yy, :, 0] >= 0).sum()
                for xx in range(nx) for yy in range(ny))
    print('small maps: pixels with duplicated indices (partsort quirk): iquv', ndup, 'iqud', ndupd)
    assert ndup > 0 and ndupd > 0, 'engineered quirk pixels did not trigger'
    np.savez_compressed(os.path.join(GOLDEN, 'small_maps.npz'), db0=small, db1=small2, db2=quirk, db_enc=enc,
                        hdr_g=hdr_small[0], sobs_dopp=obs['sobs_dopp'], snr=obs['snr'], yobs=obs['yobs'],
                        aobs=obs['aobs'], dobs=obs['dobs'], cdelt=np.float64(obs['keyvals'][11]), **cases)

    # ------------------------------------------------------------------ case 3: full-size database, a few pixels
    big = synth.make_database(9, 40)
    hdr_big = prepinv.sdb_parseheader('./tests/dbfiles/fe-xiii_1074/db.hdr')
    nx, ny = 6, 6
    obsb = synth.make_observation(nx, ny, [big], seed=5, frac_nan=0.03, frac_zero=0.03)
    outb = {}
    for tag, iqud, bcalc in (('iquv', False, 0), ('iqud', True, 3)):
        iss = obsb['issuemask'].copy()
        inv, sf = procinv.cledb_invproc(obsb['sobs_totrot'], obsb['sobs_dopp'], [big], obsb['db_enc'], obsb['yobs'],
                                        obsb['aobs'], obsb['dobs'], obsb['snr'], iss, hdr_big, obsb['keyvals'],
                                        8, 1000, bcalc, iqud, 6, False, 0)
        outb['invout_' + tag], outb['sfound_' + tag], outb['issuemask_' + tag] = inv, sf, iss
    np.savez_compressed(os.path.join(GOLDEN, 'full_db_pixels.npz'), db_hd=np.array([9, 40]), db_sha=sha(big),
                        hdr_g=hdr_big[0], sobs=obsb['sobs_totrot'], sobs_dopp=obsb['sobs_dopp'], snr=obsb['snr'],
                        yobs=obsb['yobs'], aobs=obsb['aobs'], dobs=obsb['dobs'], cdelt=np.float64(obsb['keyvals'][11]),
                        **outb)
    print('full_db_pixels: idx[0,0]', outb['invout_iquv'][0, 0, :, 0])

    # ------------------------------------------------------------------ case 4: elementwise neighbours
    nx, ny = 24, 20
    one = synth.make_oneline_map(nx, ny, seed=11, frac_bad=0.05)
    one[3, 3, 1] = 0.0           # Q == 0 -> flag 32 (and 64 when U/I < 1e-6)
    one[3, 4, 1] = 0.0
    one[3, 4, 2] = -abs(one[3, 4, 2])
    one[3, 5, 2] = one[3, 5, 0] * np.float32(5e-7)
    two = np.concatenate([one, synth.make_oneline_map(nx, ny, seed=12, frac_bad=0.05)], axis=-1)
    el = {}
    for nline, arr, lines in ((1, one, ['si-x_1430', 0]), (2, two, ['fe-xiii_1074', 'fe-xiii_1079'])):
        kv = synth.make_keyvals(nx, ny, nline=nline)
        kv[4] = lines
        iss = np.zeros((nx, ny, nline), dtype=np.int32)
        params.iqud = False
        params.ncpu = 6
        el['blos%d' % nline] = procinv.blos_proc(arr, np.ones_like(arr), iss, kv, consts, params)
        el['blos%d_iss' % nline] = iss
    el['one'], el['two'] = one, two
    raw = synth.make_twoline_raw(nx, ny, seed=21)
    variants = {'std': (synth.make_keyvals(nx, ny), None),
                'crpix_linpol': (synth.make_keyvals(nx, ny), None),
                'hpxy': (synth.make_keyvals(nx, ny), 'grid')}
    variants['crpix_linpol'][0][5], variants['crpix_linpol'][0][6] = 3, 5
    variants['crpix_linpol'][0][14] = 0.7
    variants['crpix_linpol'][0][8] = 0.05        # crosses x = 0 so that a < linpolref happens
    for tag, (kv, grid) in variants.items():
        if grid is None:
            hp = np.zeros((2, nx, ny), dtype=np.float32)
        else:
            rg = np.random.default_rng(3)
            hp = rg.uniform(-1.5, 1.5, (2, nx, ny)).astype(np.float32)
        rot, ao = prepinv.obs_qurotate(raw, kv, hp)
        yo = prepinv.obs_calcheight(kv, hp)
        el['rot_' + tag], el['aobs_' + tag], el['yobs_' + tag], el['hp_' + tag] = rot, ao, yo, hp
        el['kv_' + tag] = np.array([kv[5], kv[6], kv[8], kv[9], kv[11], kv[12], kv[14]], dtype=np.float64)
    el['raw'] = raw
    np.savez_compressed(os.path.join(GOLDEN, 'elementwise.npz'), **el)
    print('elementwise: blos1[0,0]', el['blos1'][0, 0, 0], 'flags set', int((el['blos1_iss'] > 0).sum()))
    # known answer of tests/debug_unittests.ipynb cells 9-10
    ka = procinv.blos_proc_1pix(0, 0, 0, np.array([2.7957714e-09, 3.3037444e-11, -5.7222532e-11, -1.2028063e-13],
                                                   dtype=np.float32), consts.Constants('fe-xiii_1074'), None)
    print('known answer blos:', ka[3])
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == '__main__':
    main()
