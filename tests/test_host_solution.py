"""CPU tests of the host-side solution building (cledb_b200.solution): fed with the oracle's raw search result
it must reproduce the reference's invout / sfound / issuemask of the golden fixtures bit-for-bit."""
import os

import numpy as np
import pytest

from cledb_b200 import native, solution
import cledb_b200.synth as synth
from helpers import hdr_from_g, same, oracle_raw

TAGS = ['iquv_b0', 'iquv_b1', 'iquv_b2_k4', 'iquv_b0_chi3', 'iquv_b0_k16', 'iqud', 'iqud_chi3']


@pytest.mark.parametrize('tag', TAGS)
def test_build_solutions_matches_reference(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, 'small_maps.npz'))
    dbs = [g['db0'], g['db1'], g['db2']]
    hdr = hdr_from_g(g['hdr_g'])
    iqud, bcalc, nsearch, maxchisq = g['cfg_' + tag]
    mode, k = int(iqud), int(nsearch)
    nx, ny = 10, 12
    sobs = g['sobs_' + tag].reshape(-1, 8)
    snr = g['snr'].reshape(-1, 8)
    raw = oracle_raw(mode, sobs, snr, dbs, g['db_enc'].reshape(-1), k)
    inv, sf = solution.build_solutions(raw, mode, sobs, g['sobs_dopp'].reshape(nx * ny, -1), g['yobs'].reshape(-1),
                                       g['aobs'].reshape(-1), g['dobs'].reshape(-1), hdr, k, maxchisq, int(bcalc))
    inv, sf = inv.reshape(nx, ny, k, 11), sf.reshape(nx, ny, k, 8)
    ref_inv, ref_sf = g['invout_' + tag], g['sfound_' + tag]
    assert np.array_equal(inv[..., 0], ref_inv[..., 0]), 'indices (incl. partsort-quirk duplicates)'
    assert np.array_equal(inv[..., 1], ref_inv[..., 1]), 'chi^2'
    assert np.array_equal(inv[..., [2, 3, 4, 6, 7]], ref_inv[..., [2, 3, 4, 6, 7]]), 'ne, y, gx, phi, theta'
    np.testing.assert_allclose(inv[..., [5, 8, 9, 10]], ref_inv[..., [5, 8, 9, 10]], rtol=1e-6, atol=0)
    np.testing.assert_allclose(sf, ref_sf, rtol=1e-6, atol=0)
    iss = np.zeros((nx, ny, 2), np.int32)
    solution.update_issuemask(iss, inv, g['snr'], nx, ny)
    assert same(iss, g['issuemask_' + tag])


def test_replay_partsort_examples():
    """[1,5,6,0,7],k=2 -> [3,3]; [2,1,6,0,7,0.5],k=4 -> [3,5,5,3]."""
    # true top-2 of [1,5,6,0,7]: positions [3,0]
    sel = solution.replay_partsort(np.array([[3, 0]]), np.array([[0., 1.]]), np.ones((1, 2), bool))
    assert np.array([[3, 0]])[0, sel[0]].tolist() == [3, 3]
    # true top-4 of [2,1,6,0,7,0.5]: positions [3,5,1,0]
    sel = solution.replay_partsort(np.array([[3, 5, 1, 0]]), np.array([[0., 0.5, 1., 2.]]), np.ones((1, 4), bool))
    assert np.array([3, 5, 1, 0])[sel[0]].tolist() == [3, 5, 5, 3]


def test_replay_partsort_random():
    """Replay on (positions, values) of the true top-k equals the literal partial selection sort."""
    import oracle.cledb_oracle as orc
    rng = np.random.default_rng(0)
    for trial in range(300):
        n, k = int(rng.integers(12, 40)), int(rng.integers(1, 9))
        chi = rng.random(n)
        lit = orc.selection_positions(chi, k)
        top = orc.stable_topk(chi, k)
        sel = solution.replay_partsort(top[None, :].astype(np.int32), chi[top][None, :], np.ones((1, k), bool))
        assert top[sel[0]].tolist() == lit.tolist()


@pytest.mark.parametrize('tag', ['iquv_b0', 'iquv_b2', 'iqud'])
def test_build_solutions_cle_type(golden_dir, tag):
    """CLE-type databases (dbtype 0) are built by the host path: density index from the flat row, density from
    cledb_elecdens (CLEDB_PROC.py:1688-1697, 1744-1747).  Fed with the oracle's raw search result it must give the
    reference's own output."""
    from test_oracle_golden import cle_header
    g = np.load(os.path.join(golden_dir, 'cle_type.npz'))
    hdr = cle_header(g)
    iqud, bcalc, k, mx = (int(v) for v in g['cfg_' + tag])
    nx, ny = g['sobs'].shape[:2]
    sobs, snr = g['sobs'].reshape(-1, 8), g['snr'].reshape(-1, 8)
    raw = oracle_raw(iqud, sobs, snr, [g['db']], np.zeros(nx * ny, np.int32), k)
    inv, sf = solution.build_solutions(raw, iqud, sobs, g['sobs_dopp'].reshape(nx * ny, -1), g['yobs'].reshape(-1),
                                       g['aobs'].reshape(-1), g['dobs'].reshape(-1), hdr, k, mx, bcalc)
    inv, sf = inv.reshape(nx, ny, k, 11), sf.reshape(nx, ny, k, 8)
    ref_inv, ref_sf = g['invout_' + tag], g['sfound_' + tag]
    assert np.array_equal(inv[..., [0, 1, 3, 4, 6, 7]], ref_inv[..., [0, 1, 3, 4, 6, 7]])
    np.testing.assert_allclose(inv[..., 2], ref_inv[..., 2], rtol=1e-6, atol=0)          # log10 of the Baumbach density
    np.testing.assert_allclose(inv[..., [5, 8, 9, 10]], ref_inv[..., [5, 8, 9, 10]], rtol=1e-6, atol=0)
    np.testing.assert_allclose(sf, ref_sf, rtol=1e-6, atol=0)
    iss = np.zeros((nx, ny, 2), np.int32)
    solution.update_issuemask(iss, inv, g['snr'], nx, ny)
    assert same(iss, g['issuemask_' + tag])


def test_issue_codes_from_device_flags():
    """native.issue_codes + solution.apply_issue_codes == solution.update_issuemask: the flags k_build writes (restated here
    in float64, with the 1e-3 deg ambiguity band handed back to NumPy) give the reference's issue mask, and a code that is
    already set is not added twice."""
    rng = np.random.default_rng(3)
    npix, k = 4000, 3
    inv = np.zeros((npix, k, 11), np.float32)
    inv[:, :, 3] = rng.uniform(1.0, 1.5, (npix, 1))
    inv[:, :, 4] = rng.uniform(-0.5, 0.5, (npix, k))
    inv[:, :, 6] = rng.uniform(0, 6.2, (npix, k))
    inv[:, :, 7] = rng.uniform(0, 3.1, (npix, k))
    inv[:, 0, 1] = rng.uniform(0, 20, npix)
    inv[::7] = 0                                                 # rejected pixels: zeros
    snr = rng.uniform(0.5, 3, (npix, 1, 8)).astype(np.float32)
    o = inv[:, -1, :].astype(np.float64)
    alpha = -np.arctan2(o[:, 4], o[:, 3])
    xb, yb, zb = np.sin(o[:, 7]) * np.cos(o[:, 6]), np.sin(o[:, 7]) * np.sin(o[:, 6]), np.cos(o[:, 7])
    cb = 6.123233995736766e-17
    yyb, yzb = cb * yb - zb, yb + cb * zb
    xxb, zzb = np.cos(alpha) * xb + np.sin(alpha) * yzb, -np.sin(alpha) * xb + np.cos(alpha) * yzb
    deg = np.arctan2(np.sqrt(xxb ** 2 + yyb ** 2), zzb) * 180.0 / np.pi
    flags = np.zeros(npix, np.int32)
    flags |= (((deg >= 53) & (deg <= 56)).astype(np.int32) << 9) | ((np.abs(deg - 53) < 1e-3) | (np.abs(deg - 56) < 1e-3)).astype(np.int32)
    flags |= ((inv[:, 0, 1] > 10).astype(np.int32) << 10) | ((snr[:, 0, 0] < 1).astype(np.int32) << 11)
    flags[:5] |= 1                                               # force the NumPy re-evaluation on a few pixels
    codes = native.issue_codes(flags.copy(), inv)
    assert not (codes & 1).any()
    a = rng.choice([0, 32, 512, 1024 + 4, 2048 + 512], (npix, 1, 2)).astype(np.int32)
    b = a.copy()
    solution.apply_issue_codes(a, codes.reshape(npix, 1))
    solution.update_issuemask(b, inv.reshape(npix, 1, k, 11), snr, npix, 1)
    assert np.array_equal(a, b)
    assert (codes & 512).any() and (codes & 1024).any() and (codes & 2048).any()


def test_apply_issue_codes_on_every_mask_type():
    """The vectorised paths of apply_issue_codes (integer masks: one OR per line; float masks - the reference's own issuemask is
    float64, CLEDB_PREPINV.py:159 - one integer view and one addition per line) equal the reference's arithmetic (add a code
    only where its bit is not set, CLEDB_PROC.py:347-355) on masks that already hold arbitrary codes; a third line is untouched."""
    rng = np.random.default_rng(0)
    for dt in (np.float64, np.float32, np.int32, np.int64, np.int16):
        nx, ny = 37, 256
        codes = rng.choice([0, 512, 1024, 1536, 2048, 2560, 3072, 3584], (nx, ny)).astype(np.int32)
        a = rng.choice([0, 1, 32, 64, 512, 1024 + 4, 2048 + 512, 3584, 4096 + 1024], (nx, ny, 3)).astype(dt)
        b = a.copy()
        solution.apply_issue_codes(a, codes)
        for line in range(2):
            mm = b[:, :, line]
            for code in (512, 1024, 2048):
                add = ((codes & code) != 0) & ((mm.astype(np.int64) & code) == 0)
                mm[add] += code
        assert np.array_equal(a, b), dt


def test_issue_codes_finds_every_marked_pixel():
    """native.issue_codes looks for the marked pixels (bit 0) row by row of a 2-D view when the map allows it: the same pixels
    as a flat search, whatever the map size."""
    rng = np.random.default_rng(1)
    for n in (65536, 4096, 5000, 256 * 300, 7):
        codes = rng.choice([0, 512, 1024, 2048], n).astype(np.int32)
        idx = rng.choice(n, min(5, n), replace=False)
        codes[idx] |= 1
        inv = np.zeros((n, 2, 11), np.float32)
        inv[:, :, 3] = 1
        inv[:, :, 7] = rng.uniform(0, 3.1, (n, 1))
        inv[:, :, 6] = rng.uniform(0, 6, (n, 1))
        want = codes.copy()
        amb = np.flatnonzero(want & 1)
        want &= ~np.int32(1)
        vv = solution.van_vleck_last(inv[amb])
        want[amb] = (want[amb] & ~np.int32(512)) | (vv.astype(np.int32) << 9)
        assert np.array_equal(native.issue_codes(codes.copy(), inv), want), n
