"""Host-side solution building: everything ``cledb_matchiquv`` / ``cledb_matchiqud`` / ``cledb_invproc`` do
AFTER the chi^2 search, vectorised over ``[npix, nsearch]`` with NumPy.

The GPU returns, per pixel, the stable top-``nsearch`` (original row index, chi^2, position in the candidate
list, decoded row).  From that this module reproduces, in the reference's own dtypes and operation order:

* the index quirk of ``cledb_partsort`` (CLEDB_PROC.py:1620-1625) - optional, ``strict_topk`` switches it off;
* the ``maxchisq`` cut-offs (CLEDB_PROC.py:1027-1032, 1044 / 1202-1207, 1219);
* the field strength for ``bcalc`` 0/1/2/3 (CLEDB_PROC.py:1049-1059, 1225, 1712-1713);
* ``cledb_physcle`` / ``cledb_phys`` / ``cledb_params`` (CLEDB_PROC.py:1634-1659, 1704-1749);
* ``cledb_quderotate`` (CLEDB_PROC.py:1575-1588);
* the Van-Vleck test and issue codes 512 / 1024 / 2048 of ``cledb_invproc`` (CLEDB_PROC.py:323-355).

It is O(npix * nsearch) work; the O(npix * N_db) part ran on the GPU.
"""
import numpy as np

from . import native

F32 = np.float32


def replay_partsort(kpos, chi, valid):
    """Which winner the reference reports in each output slot.

    ``kpos[npix,k]``: positions of the true top-k inside the pixel's candidate list (ascending chi^2, ties by
    position).  The reference's partial selection sort returns ``argmin(tmp[i:]) + i`` of an array it is
    swapping in place (CLEDB_PROC.py:1620-1625); when one of the first ``k`` candidates is itself a winner but
    not at its turn, the value it was swapped to is found again at the *old* position of an earlier winner, so
    that earlier winner is reported twice.  Only the winners' positions and the first ``k`` positions are ever
    touched, so the sort can be replayed exactly on that sparse set.  Returns ``sel[npix,k]`` (index of the
    winner shown in slot i; identity when the quirk does not trigger).
    """
    npix, k = kpos.shape
    sel = np.broadcast_to(np.arange(k, dtype=np.int64), (npix, k)).copy()
    touched = np.flatnonzero(((kpos >= 0) & (kpos < k) & valid).any(axis=1))
    for p in touched:
        if not valid[p].all():
            continue        # fewer than k candidates: the reference raises; keep the plain order
        q = kpos[p].astype(np.int64)
        positions = np.union1d(np.arange(k), q)            # sorted; the first k entries are 0..k-1
        work = np.full(positions.size, np.inf)
        slot_of = {int(pos): m for m, pos in enumerate(q)}
        for m, pos in enumerate(q):
            work[np.searchsorted(positions, pos)] = chi[p, m]
        for i in range(k):
            a = int(np.argmin(work[i:])) + i
            sel[p, i] = slot_of.get(int(positions[a]), i)
            work[i], work[a] = work[a], work[i]
    return sel


def unravel_index(index, dbhdr):
    """flat index -> (i, j, k, l), CLEDB_PROC.py:1634-1659 (int64 arithmetic)."""
    dbcgrid = dbhdr[0]
    dbtype = int(dbhdr[14])
    ngx, nbphi, nbtheta = int(dbcgrid[2]), int(dbcgrid[3]), int(dbcgrid[4])
    n5, n4, n3 = ngx * nbphi * nbtheta, nbphi * nbtheta, nbtheta
    index = index.astype(np.int64)
    i = index // n5 if dbtype == 0 else np.zeros_like(index)
    j = (index - i * n5) // n4
    k = (index - i * n5 - j * n4) // n3
    l = index - i * n5 - j * n4 - k * n3
    return i, j, k, l


def entry_physics(index, yobs, dobs, dbhdr, bcalc, b):
    """Columns 2..10 of ``invout`` for arrays of winners; CLEDB_PROC.py:1704-1749.

    ``index`` int64 [M]; ``yobs``/``dobs`` [M]; ``b`` [M] in the dtype the reference would carry (float32, or the
    dtype of ``sobs_dopp`` for ``bcalc == 3``).  The grid arithmetic is float64 because the reference multiplies a
    NumPy int64 index with float32 header values; the result is rounded to float32 exactly as there.
    """
    dbcgrid, ned, ngx, nbphi, nbtheta, xed, gxmin, gxmax, bphimin, bphimax, bthetamin, bthetamax, nline, wavel, dbtype = dbhdr
    i, j, k, l = unravel_index(index, dbhdr)
    with np.errstate(all='ignore'):
        gx = np.float64(gxmin) + (j.astype(np.float64) * np.float64(F32(gxmax) - F32(gxmin))) / np.float64(int(ngx) - 1)
        bphi = np.float64(bphimin) + (k.astype(np.float64) * np.float64(F32(bphimax) - F32(bphimin))) / np.float64(int(nbphi) - 1)
        btheta = np.float64(bthetamin) + (l.astype(np.float64) * np.float64(F32(bthetamax) - F32(bthetamin))) / np.float64(int(nbtheta) - 1)
        if int(dbtype) == 0:
            # CLE databases: density from the Baumbach profile, CLEDB_PROC.py:1689-1697, 1745
            r = np.sqrt(yobs.astype(F32) * yobs.astype(F32) + gx * gx)
            dens = (3.e8 * np.exp(-(r - 1.) / 7.18401074e-02) + 1.e8 * (0.036 / r ** 1.5 + 1.55 / r ** 6.)).astype(F32)
            ed = np.log10(np.asarray(xed, dtype=F32)[i] * dens)
        else:
            ed = dobs
        ed = np.asarray(ed).astype(F32)
        gx, bphi, btheta = gx.astype(F32), bphi.astype(F32), btheta.astype(F32)
        if bcalc == 3:
            b = b / np.sqrt(np.sin(btheta) ** 2 * np.sin(bphi) ** 2 + np.cos(btheta) ** 2)
        bx = np.abs(b) * np.sin(btheta) * np.cos(bphi)
        by = np.abs(b) * np.sin(btheta) * np.sin(bphi)
        bz = np.abs(b) * np.cos(btheta)
    out = np.empty((index.shape[0], 9), dtype=F32)
    for c, col in enumerate((ed, yobs, gx, b, bphi, btheta, bx, by, bz)):
        out[:, c] = col
    return out


def derotate_rows(rows, ang, nf):
    """``cledb_quderotate`` (CLEDB_PROC.py:1575-1588) for [M,8] rows, [M] angles and [M] normalisation factors."""
    with np.errstate(all='ignore'):
        nf = nf.astype(F32)[:, None]
        out = rows * nf
        c = np.cos(-2. * ang.astype(F32))
        s = np.sin(-2. * ang.astype(F32))
        n1 = nf[:, 0]
        for q in (1, 5):
            out[:, q] = rows[:, q] * n1 * c - rows[:, q + 1] * n1 * s
            out[:, q + 1] = rows[:, q] * n1 * s + rows[:, q + 1] * n1 * c
    return out


def van_vleck_last(invout_flat):
    """Van-Vleck proximity flag per pixel, computed from the LAST solution only, exactly as the reference's loop
    leaves it (``flg = 0`` sits inside the loop over solutions, CLEDB_PROC.py:323-343).  ``invout_flat[npix,k,11]``."""
    o = invout_flat[:, -1, :]
    with np.errstate(all='ignore'):
        alpha = -np.arctan2(o[:, 4], o[:, 3])
        beta = np.pi / 2
        xb = np.sin(o[:, 7]) * np.cos(o[:, 6])
        yb = np.sin(o[:, 7]) * np.sin(o[:, 6])
        zb = np.cos(o[:, 7])
        yyb = np.cos(beta) * yb - np.sin(beta) * zb
        yzb = np.sin(beta) * yb + np.cos(beta) * zb
        xxb = np.cos(alpha) * xb + np.sin(alpha) * yzb
        zzb = -np.sin(alpha) * xb + np.cos(alpha) * yzb
        vartheta = np.arctan2(np.sqrt(xxb ** 2 + yyb ** 2), zzb)
        deg = vartheta * 180 / np.pi
        return (53 <= deg) & (deg <= 56)


def build_solutions(raw, mode, sobs, sobs_dopp, yobs, aobs, dobs, dbhdr_of_pixel, nsearch, maxchisq, bcalc,
                    strict_topk=False):
    """From the raw GPU search result to ``(invout[npix,k,11], sfound[npix,k,8])`` float32.

    ``raw``: dict from ``native.Context.match`` (needs rows and kpos).  ``sobs[npix,8]``; ``sobs_dopp[npix,>=1]`` or
    None; ``yobs, aobs, dobs [npix]``; ``dbhdr_of_pixel``: a single dbhdr tuple (all databases of one call share
    the header, CLEDB_PREPINV.py:298-301).
    """
    k = int(nsearch)
    npix = sobs.shape[0]
    idx, chi, kpos, rows, status = raw['idx'], raw['chi'], raw['kpos'], raw['rows'], raw['status']
    chi_cmp = raw['chi64'] if raw.get('chi64') is not None else chi.astype(np.float64)
    valid = idx >= 0
    searched = status == native.PIX_SEARCHED
    sel = np.broadcast_to(np.arange(k), (npix, k)) if strict_topk else replay_partsort(kpos, chi_cmp, valid)
    ar = np.arange(npix)[:, None]
    shift = raw.get('nan_slots')
    if shift is not None and (shift > 0).any():
        # reduced mode: candidates with a NaN chi^2 come first in the reference's selection sort and leave their output slots
        # empty (they fail both chi^2 tests); the finite winners follow
        sel = np.array(sel)
        sel[shift > 0] = np.arange(k)
        src = np.arange(k)[None, :] - shift[:, None]
        sel = np.where(src >= 0, np.take_along_axis(sel, np.maximum(src, 0), axis=1), 0)
        nan_lead = src < 0
    else:
        nan_lead = np.zeros((npix, k), bool)
    idx_s, chi_s, rows_s, valid_s = idx[ar, sel], chi_cmp[ar, sel], rows[ar, sel], valid[ar, sel] & ~nan_lead

    invout = np.zeros((npix, k, 11), dtype=F32)
    invout[:, :, 0] = -1
    sfound = np.zeros((npix, k, 8), dtype=F32)
    with np.errstate(all='ignore'):
        limit = maxchisq * 10 if mode == native.MODE_IQUV else maxchisq
        best = chi_s[:, 0]
        rejected = searched & valid_s[:, 0] & (best > limit)             # CLEDB_PROC.py:1027 / 1202 (a leading NaN is not > limit)
        invout[rejected, :, 1] = best[rejected, None].astype(F32)
        allinf = status == native.PIX_ALLINF                             # every chi^2 is +inf in the reference
        invout[allinf, :, 1] = np.inf
        keep = searched[:, None] & ~rejected[:, None] & valid_s & (chi_s <= maxchisq)   # :1044 / :1219
        pp, ss = np.nonzero(keep)
        if pp.size:
            r = rows_s[pp, ss]
            s = sobs[pp]
            if mode == native.MODE_IQUV:
                nf = s[:, 0]
                def ratio(c):
                    return np.where(np.abs(r[:, c]) > F32(1e-7), s[:, c] / nf / r[:, c], F32(0)).astype(F32)
                if bcalc == 0:
                    b = ratio(3)
                elif bcalc == 1:
                    b = ratio(7)
                else:
                    b = (F32(0.5) * (ratio(7) + ratio(3))).astype(F32)
            else:
                nf = np.max(s, axis=1)
                b = np.asarray(sobs_dopp)[pp, 0]
                if b.dtype != np.float64:
                    b = b.astype(F32)
            phys = entry_physics(idx_s[pp, ss], yobs[pp].astype(F32), dobs[pp], dbhdr_of_pixel, bcalc, b)
            invout[pp, ss, 0] = idx_s[pp, ss]
            invout[pp, ss, 1] = chi_s[pp, ss]
            invout[pp, ss, 2:] = phys
            sfound[pp, ss, :] = derotate_rows(r, aobs[pp], nf)
    return invout, sfound


def update_issuemask(issuemask, invout, snr, nx, ny):
    """Issue codes 512 / 1024 / 2048 for both lines; CLEDB_PROC.py:347-355.  ``invout[nx,ny,k,11]``; in place."""
    k = invout.shape[2]
    flat = invout.reshape(nx * ny, k, 11)
    vv = van_vleck_last(flat).reshape(nx, ny)
    with np.errstate(all='ignore'):
        poor = (flat[:, 0, 1] > 10).reshape(nx, ny)
        noisy = (snr[:, :, 0] < 1)
    for line in range(2):
        m = issuemask[:, :, line]
        for cond, code in ((vv, 512), (poor, 1024), (noisy, 2048)):
            bits = m.astype(np.int64)
            add = cond & ((bits & code) == 0)
            m[add] += code
    return issuemask


def apply_issue_codes(issuemask, codes):
    """Add the per-pixel issue codes (a sum of 512 / 1024 / 2048, ``native.issue_codes``) to both lines of ``issuemask``,
    each code only where it is not yet set (CLEDB_PROC.py:347-355).  Adding a power of two whose bit is clear is an OR."""
    if isinstance(issuemask, np.ndarray) and np.issubdtype(issuemask.dtype, np.integer):
        c = codes if codes.dtype == issuemask.dtype else codes.astype(issuemask.dtype)
        for line in range(2):                              # one strided pass per line: 2.5x faster than a broadcast over [:, :, 0:2]
            issuemask[:, :, line] |= c
        return issuemask
    if isinstance(issuemask, np.ndarray) and np.issubdtype(issuemask.dtype, np.floating):
        # the reference's own mask is float64 (np.zeros without a dtype, CLEDB_PREPINV.py:159) holding integers: one integer view
        # of a line, the bits that are not yet set, one addition
        c = codes.astype(np.int64) & 3584
        for line in range(2):
            mm = issuemask[:, :, line]
            mm += c & ~mm.astype(np.int64)
        return issuemask
    for line in range(2):                                  # exotic mask types: the reference's own arithmetic
        mm = issuemask[:, :, line]
        for code in (512, 1024, 2048):
            add = ((codes & code) != 0) & ((mm.astype(np.int64) & code) == 0)
            mm[add] += code
    return issuemask
