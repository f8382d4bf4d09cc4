"""Pixel-by-pixel comparison of a raw GPU search result with the CPU oracle.  TEST INFRASTRUCTURE ONLY
(used by tests/ and by __graft_entry__.smoke())."""
import numpy as np

from . import cledb_oracle as orc

CHI_RTOL, CHI_ATOL = 2e-5, 1e-9
PIX_SEARCHED, PIX_NULL, PIX_ALLNAN, PIX_REFRAISES, PIX_ALLINF = 0, 1, 2, 3, 4


def check_raw_against_oracle(raw, mode, sobs, snr, dbs_decoded, enc, k, exact=False):
    """Compare cledb_match output with the oracle pixel by pixel.  Returns (#pixels searched, #pixels whose index
    list differs from the oracle's by a near-tie reordering)."""
    npix = sobs.shape[0]
    searched = reordered = 0
    for p in range(npix):
        db = dbs_decoded[int(enc[p])]
        ref = orc.raw_topk(mode, sobs[p], snr[p], db, k, return_all=True)
        if ref is None:
            assert raw['status'][p] == PIX_NULL, p
            assert (raw['idx'][p] == -1).all()
            continue
        ridx, rchi, rpos, ncand, allchi = ref
        if ncand == 0:
            assert raw['status'][p] == PIX_REFRAISES, p
            continue
        if np.isnan(allchi).any():
            assert raw['status'][p] == PIX_ALLNAN, (p, raw['status'][p])
            continue
        if np.isinf(rchi).all():
            assert raw['status'][p] == PIX_ALLINF, (p, raw['status'][p])
            continue
        assert raw['status'][p] == PIX_SEARCHED, (p, raw['status'][p])
        assert raw['ncand'][p] == ncand
        searched += 1
        n = min(k, ncand)
        gidx, gchi = raw['idx'][p], raw['chi'][p]
        assert (gidx[n:] == -1).all()
        if exact:
            assert np.array_equal(gidx[:n], ridx[:n]), (p, gidx, ridx)
            assert np.array_equal(raw['chi64'][p][:n], rchi[:n]), (p, raw['chi64'][p], rchi)
            assert np.array_equal(raw['kpos'][p][:n], rpos[:n])
        else:
            np.testing.assert_allclose(gchi[:n], rchi[:n], rtol=CHI_RTOL, atol=CHI_ATOL, err_msg='pixel %d' % p)
            if not np.array_equal(gidx[:n], ridx[:n]):
                reordered += 1
                # every index the GPU reports must be a genuine near-tie member: its ORACLE chi^2 matches the slot
                keep = np.flatnonzero(np.sign(db[:, 3]) == np.sign(sobs[p, 3])) if mode == 0 else np.flatnonzero(db[:, 3] == db[:, 3])
                lookup = dict(zip(keep.tolist(), allchi.tolist()))
                for r in range(n):
                    assert gidx[r] in lookup, (p, r, gidx[r])
                    np.testing.assert_allclose(lookup[int(gidx[r])], rchi[r], rtol=CHI_RTOL, atol=CHI_ATOL)
                assert len(set(gidx[:n].tolist())) == n
        # decoded rows returned by the kernel are the database rows of the reported indices
        assert np.array_equal(raw['rows'][p][:n], db[gidx[:n]], equal_nan=True)
    return searched, reordered
