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
