"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/*.npz, made by
oracle/run_reference.py) and against the known answers of the reference's own tests.  CPU only."""
import hashlib
import os

import numpy as np
import pytest

import oracle.cledb_oracle as orc
import cledb_b200.synth as synth
import constants as consts


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hdr_from_g(g):
    """dbhdr 15-tuple from the stored 16 header numbers (layout of CLEDB_PREPINV.py:482-508)."""
    h = synth.make_dbhdr(int(g[2]), int(g[3]), int(g[4]), dbtype=int(g[-1]))
    assert np.array_equal(h[0], g)
    return h


def same(a, b):
    """bit-for-bit equality that treats NaN == NaN"""
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def test_unit_pixel_chain(golden_dir):
    """The reference's own test pixel (tests/test_functions.py) through IQUV and IQUD search + BLOS."""
    g = np.load(os.path.join(golden_dir, 'unit_pixel.npz'))
    rng = np.random.default_rng(99)
    pairs = {}
    for hi, name in ((5, 'DB_h106_d895.npy'), (8, 'DB_h109_d795.npy'), (8, 'DB_h109_d895.npy')):
        pairs[name] = synth.make_raw_line_pair(hi, 59 if 'd895' in name else 39, rng)
    l1, l2 = pairs[str(g['db_file'])]
    assert sha(l1) == str(g['raw1_sha']) and sha(l2) == str(g['raw2_sha'])
    db = orc.normalise_database(l1, l2, np.float32(g['wl0']))
    assert sha(db) == str(g['db_sha']), 'normalise_database differs from the reference sdb_preprocess'
    assert np.all(db[..., 0] == 1.0)
    hdr = hdr_from_g(g['hdr_g'])
    kv = synth.make_keyvals(1, 1, cdelt=float(g['cdelt'][0]))
    for tag, iqud, bcalc, dopp in (('iquv', False, 0, np.zeros((1, 1, 1), np.float32)), ('iqud', True, 3, g['dopp'])):
        iss = g['issuemask_in'].copy()
        inv, sf = orc.invert_map(g['sobs_totrot'], dopp, [db], np.zeros((1, 1), np.int32), g['yobs'], g['aobs'],
                                 g['dobs'], g['snr'], iss, hdr, kv, 8, 1000, bcalc, iqud)
        assert same(inv, g['invout_' + tag]), tag
        assert same(sf, g['sfound_' + tag]), tag
        assert same(iss, g['issuemask_' + tag]), tag
    # relations the reference's own test asserts (tests/test_functions.py:130-133)
    inv = g['invout_iquv']
    assert inv[0, 0, 0, 2] == g['dobs'][0, 0] and inv[0, 0, 0, 3] == g['yobs'][0, 0]
    # BLOS
    kv2 = synth.make_keyvals(1, 1)
    iss = g['issuemask_in'].copy()
    blos = orc.blos_map(g['sobs_tot'], iss, kv2, consts)
    assert same(blos, g['blosout']) and same(iss, g['issuemask_blos'])
    np.testing.assert_allclose(blos[:, :, :, 3] * 180 / np.pi, -30, atol=1e-2)      # tests/test_functions.py:109-110


def test_blos_known_answer():
    """tests/debug_unittests.ipynb cells 9-10 (author's run, real data)."""
    out, flags = orc.blos_pixel(np.array([2.7957714e-09, 3.3037444e-11, -5.7222532e-11, -1.2028063e-13], np.float32),
                                consts.Constants('fe-xiii_1074'))
    assert np.array_equal(out, np.array([5.4486594, 5.197059, 5.319886, -0.5235988], np.float32))
    assert flags == [0, 0]


def test_qurotate_known_answer(golden_dir):
    """sobs_totrot / yobs / aobs pinned by tests/test_functions.py:61-66."""
    g = np.load(os.path.join(golden_dir, 'unit_pixel.npz'))
    kv = synth.make_keyvals(1, 1, cdelt=1e-4)
    rot, aobs, yobs = orc.qurotate_map(g['sobs_tot'], kv, np.zeros((2, 1, 1), np.float32))
    assert same(rot, g['sobs_totrot']) and same(aobs, g['aobs']) and same(yobs, g['yobs'])
    np.testing.assert_allclose(rot, np.array([2.7957712e-09, 5.8281428e-11, -3.1131424e-11, -1.2028062e-13,
                                              9.8659880e-10, 1.8144947e-12, -9.6922470e-13, -4.1924003e-14]).reshape(1, 1, 8),
                               atol=1e-14)
    np.testing.assert_allclose(yobs, 1.09, atol=1e-2)
    np.testing.assert_allclose(aobs, 0.27, atol=1e-2)


SMALL_TAGS = ['iquv_b0', 'iquv_b1', 'iquv_b2_k4', 'iquv_b0_chi3', 'iquv_b0_k16', 'iqud', 'iqud_chi3', 'iqud_badbcalc']


@pytest.mark.parametrize('tag', SMALL_TAGS)
def test_small_maps(golden_dir, tag):
    """10x12 maps over three small databases with engineered edge pixels (null, zero sigma, zero SNR,
    partsort-quirk duplicates, exact ties, chi^2 cut-offs)."""
    g = np.load(os.path.join(golden_dir, 'small_maps.npz'))
    dbs = [g['db0'], g['db1'], g['db2']]
    hdr = hdr_from_g(g['hdr_g'])
    iqud, bcalc, nsearch, maxchisq = g['cfg_' + tag]
    kv = synth.make_keyvals(10, 12, cdelt=float(g['cdelt']))
    iss = np.zeros((10, 12, 2), np.int32)
    inv, sf = orc.invert_map(g['sobs_' + tag], g['sobs_dopp'], dbs, g['db_enc'], g['yobs'], g['aobs'], g['dobs'],
                             g['snr'], iss, hdr, kv, int(nsearch), maxchisq, int(bcalc), bool(iqud))
    assert same(inv, g['invout_' + tag])
    assert same(sf, g['sfound_' + tag])
    assert same(iss, g['issuemask_' + tag])


def test_full_db_pixels(golden_dir):
    """6x6 pixels against the full-size [21,180,90,8] database (340 200 entries)."""
    g = np.load(os.path.join(golden_dir, 'full_db_pixels.npz'))
    db = synth.make_database(*[int(v) for v in g['db_hd']])
    if sha(db) != str(g['db_sha']):
        pytest.skip('synthetic database generator is not bit-reproducible on this CPU')
    hdr = hdr_from_g(g['hdr_g'])
    kv = synth.make_keyvals(6, 6, cdelt=float(g['cdelt']))
    for tag, iqud, bcalc in (('iquv', False, 0), ('iqud', True, 3)):
        iss = np.zeros((6, 6, 2), np.int32)
        inv, sf = orc.invert_map(g['sobs'], g['sobs_dopp'], [db], np.zeros((6, 6), np.int32), g['yobs'], g['aobs'],
                                 g['dobs'], g['snr'], iss, hdr, kv, 8, 1000, bcalc, iqud, workers=4)
        assert same(inv, g['invout_' + tag]) and same(sf, g['sfound_' + tag]) and same(iss, g['issuemask_' + tag])


def test_elementwise(golden_dir):
    g = np.load(os.path.join(golden_dir, 'elementwise.npz'))
    nx, ny = g['one'].shape[:2]
    for nline, key, lines in ((1, 'one', ['si-x_1430', 0]), (2, 'two', ['fe-xiii_1074', 'fe-xiii_1079'])):
        kv = synth.make_keyvals(nx, ny, nline=nline)
        kv[4] = lines
        iss = np.zeros((nx, ny, nline), np.int32)
        blos = orc.blos_map(g[key], iss, kv, consts)
        assert same(blos, g['blos%d' % nline]) and same(iss, g['blos%d_iss' % nline])
    for tag in ('std', 'crpix_linpol', 'hpxy'):
        v = g['kv_' + tag]
        kv = synth.make_keyvals(nx, ny)
        kv[5], kv[6], kv[8], kv[9], kv[11], kv[12], kv[14] = (int(v[0]), int(v[1]), float(v[2]), float(v[3]),
                                                              float(v[4]), float(v[5]), float(v[6]))
        rot, aobs, yobs = orc.qurotate_map(g['raw'], kv, g['hp_' + tag])
        assert same(rot, g['rot_' + tag]) and same(aobs, g['aobs_' + tag]) and same(yobs, g['yobs_' + tag]), tag


def test_selection_quirk_examples():
    """SURVEY 8a row a4 probes of cledb_partsort."""
    assert orc.selection_positions(np.array([1, 5, 6, 0, 7.]), 2).tolist() == [3, 3]
    assert orc.selection_positions(np.array([2, 1, 6, 0, 7, 0.5]), 4).tolist() == [3, 5, 5, 3]
    assert orc.stable_topk(np.array([2, 1, 6, 0, 7, 0.5]), 4).tolist() == [3, 5, 1, 0]


def test_database_selection(golden_dir):
    """sdb_findelongationanddens / sdb_preprocess of the unmodified reference (oracle/run_reference_next.py): file table
    in sorted-glob order, nearest (height, density) file per pixel including the never-selected last density, the list of
    distinct files, the per-pixel position in it and the normalised databases."""
    g = np.load(os.path.join(golden_dir, 'dbselect.npz'))
    names = [str(n) for n in g['grid_names']]
    table = orc.file_table(names)
    assert same(table, g['grid_table'])
    sel = np.array([orc.nearest_database(g['grid_y'][i], g['grid_d'][i], table) for i in range(0, g['grid_y'].size, 5)], np.int32)
    assert np.array_equal(sel, g['grid_sel'][::5])
    mini = [str(n) for n in g['mini_names']]
    assert mini == sorted(mini) and mini.index('DB_h106_d1000.npy') < mini.index('DB_h106_d600.npy')
    flnm = orc.select_databases(g['mini_yobs'], g['mini_dobs'], orc.file_table(mini))
    assert np.array_equal(flnm, g['mini_db_enc_flnm'])
    uniq = np.unique(flnm)
    assert np.array_equal(uniq, g['mini_db_uniq'])
    assert np.array_equal(np.searchsorted(uniq, flnm), g['mini_db_enc'])
    last_of_height = [i for i, n in enumerate(mini) if i + 1 == len(mini) or mini[i + 1][:8] != n[:8]]
    assert not np.isin(flnm, last_of_height).any(), 'the reference never selects the last file of a height'
    rng = np.random.default_rng(4242)                        # same stream as oracle/run_reference_next.py:mini_files
    shas = {}
    for h in (106, 109, 112, 120):
        for d in (600, 795, 895, 1000, 1200):
            l1, l2 = synth.make_raw_line_pair(h - 101, (d - 600) // 5, rng, ngx=3, nbphi=12, nbtheta=8)
            shas['DB_h%03d_d%d.npy' % (h, d)] = sha(orc.normalise_database(l1, l2, g['mini_wlarr'][0, 0]))
    assert [shas[mini[u]] for u in uniq] == [str(x) for x in g['mini_sha']]


@pytest.mark.parametrize('tag', ['iquv_k8', 'iquv_k4', 'iqud_k8', 'iqud_k3'])
def test_reduced_mode(golden_dir, tag):
    """reduced=True (cledb_getsubsetiquv / cledb_getsubsetiqud presort + search over the per-pixel subset, with its duplicate
    rows, all-zero phi = 0 plane, NaN chi^2 and index mapping back to the full database) against the unmodified reference."""
    g = np.load(os.path.join(golden_dir, 'reduced_maps.npz'))
    hdr = hdr_from_g(g['hdr_g'])
    kv = synth.make_keyvals(8, 9, cdelt=float(g['cdelt']))
    iqud, bcalc, k, mx = g['cfg_' + tag]
    iss = np.zeros((8, 9, 2), np.int32)
    inv, sf = orc.invert_map(g['sobs_' + tag], g['sobs_dopp'], [g['db0'], g['db1']], g['db_enc'], g['yobs'], g['aobs'], g['dobs'],
                             g['snr'], iss, hdr, kv, int(k), mx, int(bcalc), bool(iqud), reduced=True)
    assert same(inv, g['invout_' + tag]) and same(sf, g['sfound_' + tag]) and same(iss, g['issuemask_' + tag])
    if tag == 'iquv_k8':
        dsel, ored = orc.reduced_subset(0, g['sobs_' + tag][0, 0], None, hdr, g['db%d' % g['db_enc'][0, 0]], 8)
        assert same(ored, g['outred_00']) and sha(dsel) == str(g['datasel_00_sha'])


def test_spectral_integration(golden_dir):
    """obs_integrate of the unmodified reference on a 5 x 6 map built from the reference's own test spectra
    (tests/sample_test_stokesspectra.npz): bit-exact, and pixel [0,0] reproduces the PyCELP emission coefficients the
    reference's own unit test pins (tests/test_functions.py:60, atol 1e-12)."""
    g = np.load(os.path.join(golden_dir, 'integrate.npz'))
    kv = synth.make_keyvals(5, 6)
    kv[2] = 134
    tot, snr, bkg, iss = orc.integrate_map([g['cube0'], g['cube1']], g['wl'], kv)
    assert same(tot, g['sobs_tot']) and same(snr, g['snr']) and same(bkg, g['background']) and same(iss, g['issuemask'])
    np.testing.assert_allclose(tot[0, 0, :4], g['aat'], atol=1e-12, rtol=0)
    np.testing.assert_allclose(tot[0, 0, 4:], g['bbt'], atol=1e-12, rtol=0)
    assert not tot[0, 1].any() and not tot[0, 2].any() and set(np.unique(iss)) == {0, 2, 3}


def test_density_lookup(golden_dir):
    """obs_dens of the unmodified reference (ratio -> density through a quadratic spline of the look-up table row above the
    pixel, issue 4 beyond the table, the 0 / NaN ratio conventions)."""
    pytest.importorskip('scipy')
    g = np.load(os.path.join(golden_dir, 'density.npz'))
    dobs, iss = orc.density_map(g['sobs'], g['yobs'], synth.make_keyvals(7, 8), dict(h=g['h'], den=g['den'], rat=g['rat']))
    assert same(dobs, g['dobs']) and same(iss, g['iss'])
    assert dobs[0, 0] == 0 and dobs[0, 1] == 0 and 4 in iss


def test_unit_test_database_choice():
    """tests/test_functions.py:91-93 of the reference: its test pixel (height 1.092, log density 7.936) picks
    DB_h109_d795.npy out of the three files of tests/dbfiles."""
    names = ['DB_h106_d895.npy', 'DB_h109_d795.npy', 'DB_h109_d895.npy']
    table = orc.file_table(names)
    assert names[int(orc.nearest_database(np.float32(1.0920165), np.float32(7.936), table))] == 'DB_h109_d795.npy'


@pytest.mark.parametrize('tag', ['iquv_b0', 'iquv_b2', 'iqud'])
def test_cle_type_database(golden_dir, tag):
    """CLE-type header (dbtype 0): the density index is the leading database axis and the solution's density comes from
    cledb_elecdens (CLEDB_PROC.py:1688-1697, 1744-1747).  Golden = the unmodified reference (oracle/run_reference_cle.py)."""
    g = np.load(os.path.join(golden_dir, 'cle_type.npz'))
    hdr = cle_header(g)
    iqud, bcalc, k, mx = (int(v) for v in g['cfg_' + tag])
    nx, ny = g['sobs'].shape[:2]
    kv = synth.make_keyvals(nx, ny, cdelt=float(g['cdelt']))
    iss = np.zeros((nx, ny, 2), np.int32)
    inv, sf = orc.invert_map(g['sobs'], g['sobs_dopp'], [g['db']], np.zeros((nx, ny), np.int32), g['yobs'], g['aobs'], g['dobs'],
                             g['snr'], iss, hdr, kv, k, mx, bcalc, bool(iqud))
    assert same(inv, g['invout_' + tag]) and same(sf, g['sfound_' + tag]) and same(iss, g['issuemask_' + tag])
    assert (inv[..., 0] >= 0).sum() > 20
    won = inv[..., 0][inv[..., 0] >= 0] // (4 * 12 * 8)
    assert len(set(won.astype(int).tolist())) >= 2, 'winners in more than one density block'


def cle_header(g):
    hdr = synth.make_dbhdr(int(g['hdr_g'][2]), int(g['hdr_g'][3]), int(g['hdr_g'][4]), dbtype=0)
    return (g['hdr_g'], np.int32(g['hdr_g'][1])) + tuple(hdr[2:5]) + (g['xed'],) + tuple(hdr[6:])
