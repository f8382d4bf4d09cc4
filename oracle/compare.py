"""Pixel-by-pixel comparison of a raw GPU search result with the CPU oracle.  TEST INFRASTRUCTURE ONLY
(used by tests/ and by __graft_entry__.smoke())."""
import numpy as np

from . import cledb_oracle as orc

# north_star / SURVEY 8(d): chi^2 of the fp32 kernel within 1e-5 relative of the reference's float64 value; an index may
# differ from the reference's only where the reference's own chi^2 of the two candidates agree within that tolerance
CHI_RTOL, CHI_ATOL = 1e-5, 1e-30
_W = np.array((1, 1, 1, 0, 0.01, 1, 1, 0), dtype=np.float64)           # wtg_fact, CLEDB_PROC.py:981


def chi_tolerance(chi_ref, snr_p):
    """Allowed |chi2_fp32 - chi2_reference| for one pixel: 1e-5 * chi2 (north_star) plus the float32 resolution floor.

    The fp32 kernel forms a residual as fma(d, a_c, nb_c) with the per-pixel constants a_c = w_c/sigma_c and
    nb_c = -w_c*s~_c/sigma_c = -w_c*snr_c rounded to float32; the reference subtracts first, (D - s~)/sigma, which is exact
    where D and s~ nearly cancel.  One float32 rounding (2^-24 relative) of nb_c and of the product moves r_c by at most
    eta_c = 2^-23*|w_c*snr_c|, and chi2 = sum r_c^2 / 2 by at most sqrt(2*chi2)*||eta|| + ||eta||^2/2 (first and second order:
    the second is all that is left where the reference's chi2 is exactly 0).  The terms only matter for matches far better
    than the noise (chi2 << 1e-3, e.g. an observation that IS a database row): at chi2 = 1 they are 1e-6 relative for the
    S/N of the benchmark maps."""
    eta = 2.0 ** -23 * float(np.sqrt(np.sum((_W * np.asarray(snr_p, dtype=np.float64)) ** 2)))
    c = np.abs(np.asarray(chi_ref, dtype=np.float64))
    with np.errstate(invalid='ignore'):
        return CHI_RTOL * c + np.sqrt(2.0 * c) * eta + 0.5 * eta * eta + CHI_ATOL


def assert_chi_close(chi, chi_ref, snr_p, msg=''):
    chi, chi_ref = np.asarray(chi, dtype=np.float64), np.asarray(chi_ref, dtype=np.float64)
    tol = chi_tolerance(chi_ref, snr_p)
    both_inf = np.isinf(chi) & np.isinf(chi_ref) & (np.sign(chi) == np.sign(chi_ref))
    with np.errstate(invalid='ignore'):
        bad = ~((np.abs(chi - chi_ref) <= tol) | both_inf | (np.isnan(chi) & np.isnan(chi_ref)))
    assert not bad.any(), '%s chi^2 outside the fp32 tolerance: got %s, reference %s, allowed %s' % (msg, chi[bad], chi_ref[bad], tol[bad])


def assert_chi_close_map(chi, chi_ref, snr):
    """The same for [npix, k] arrays with snr[npix, 8]."""
    chi, chi_ref = np.asarray(chi).reshape(snr.reshape(-1, 8).shape[0], -1), np.asarray(chi_ref).reshape(snr.reshape(-1, 8).shape[0], -1)
    for p in range(chi.shape[0]):
        assert_chi_close(chi[p], chi_ref[p], snr.reshape(-1, 8)[p], 'pixel %d:' % p)
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
            assert_chi_close(gchi[:n], rchi[:n], snr[p], 'pixel %d:' % p)
            if not np.array_equal(gidx[:n], ridx[:n]):
                reordered += 1
                # every index the GPU reports must be a genuine near-tie member: its ORACLE chi^2 matches the slot
                keep = np.flatnonzero(np.sign(db[:, 3]) == np.sign(sobs[p, 3])) if mode == 0 else np.flatnonzero(db[:, 3] == db[:, 3])
                lookup = dict(zip(keep.tolist(), allchi.tolist()))
                for r in range(n):
                    assert gidx[r] in lookup, (p, r, gidx[r])
                    assert_chi_close(lookup[int(gidx[r])], rchi[r], snr[p], 'pixel %d slot %d: index %d is not a near-tie member:' % (p, r, gidx[r]))
                assert len(set(gidx[:n].tolist())) == n
        # decoded rows returned by the kernel are the database rows of the reported indices
        assert np.array_equal(raw['rows'][p][:n], db[gidx[:n]], equal_nan=True)
    return searched, reordered


def check_invout_against_reference(inv, ref_inv, mode, sobs, snr, dbs_decoded, enc, maxchisq, sf=None, ref_sf=None, iss=None,
                                   ref_iss=None):
    """cledb_invproc-level comparison of an fp32-kernel result with the reference's (golden or oracle) output.

    ``inv`` / ``ref_inv``: [npix,k,11].  Equal indices: chi^2 within CHI_RTOL, ne / y / gx / phi / theta bit-equal, B and
    its projections and ``sf`` within 1e-6.  A differing index is accepted ONLY with the member-by-member near-tie proof:
    the GPU's entry is a candidate of the pixel and its ORACLE chi^2 agrees with the reference's chi^2 of that slot within
    CHI_RTOL (so the two candidates tie at the parity tolerance); a slot filled on one side and empty on the other is
    accepted only when the chi^2 in question sits within CHI_RTOL of a ``maxchisq`` cut-off.  Issue masks may differ only in
    pixels with such a slot.  Returns (slots compared, slots with a near-tie substitution)."""
    npix, k = inv.shape[:2]
    nsub = 0
    limit = maxchisq * 10 if mode == 0 else maxchisq
    for p in range(npix):
        a, b = inv[p], ref_inv[p]
        same_idx = a[:, 0] == b[:, 0]
        tainted = False
        if not same_idx.all():
            db = dbs_decoded[int(enc[p])].reshape(-1, 8)
            ref = orc.raw_topk(mode, sobs[p], snr[p], db, k, return_all=True)
            assert ref is not None, p
            ridx, rchi, rpos, ncand, allchi = ref
            keep = np.flatnonzero(np.sign(db[:, 3]) == np.sign(sobs[p, 3])) if mode == 0 else np.flatnonzero(db[:, 3] == db[:, 3])
            lookup = dict(zip(keep.tolist(), allchi.tolist()))
            for r in np.flatnonzero(~same_idx):
                nsub += 1
                tainted = True
                if a[r, 0] >= 0 and b[r, 0] >= 0:
                    assert int(a[r, 0]) in lookup, (p, r, a[r, 0])
                    assert_chi_close(lookup[int(a[r, 0])], b[r, 1], snr[p], 'pixel %d slot %d: index %d is not a near-tie of %d:' % (p, r, a[r, 0], b[r, 0]))
                else:                                   # one side rejected the slot: only at a cut-off
                    c = a[r, 1] if a[r, 0] >= 0 else b[r, 1]
                    c0 = max(a[0, 1], b[0, 1])
                    near = lambda v, t: abs(v - t) <= chi_tolerance(t, snr[p])
                    assert near(c, maxchisq) or near(c0, limit), (p, r, a[r], b[r])
        m = same_idx & (a[:, 0] >= 0)
        assert_chi_close(a[same_idx, 1], b[same_idx, 1], snr[p], 'pixel %d:' % p)
        assert np.array_equal(a[m][:, [2, 3, 4, 6, 7]], b[m][:, [2, 3, 4, 6, 7]], equal_nan=True), (p, a[m][:, [2, 3, 4, 6, 7]], b[m][:, [2, 3, 4, 6, 7]])
        np.testing.assert_allclose(a[m][:, [5, 8, 9, 10]], b[m][:, [5, 8, 9, 10]], rtol=1e-6, atol=0, err_msg='B of pixel %d' % p)
        if sf is not None:
            np.testing.assert_allclose(sf[p][m], ref_sf[p][m], rtol=1e-6, atol=0, err_msg='sfound of pixel %d' % p)
        if iss is not None and not tainted:
            assert np.array_equal(iss[p], ref_iss[p]), (p, iss[p], ref_iss[p])
    return npix * k, nsub
