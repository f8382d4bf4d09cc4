"""CPU oracle for the CLEDB 2-line database-search hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of what the reference (arparaschiv/solar-coronal-inversion, CLEDB 0.7.0-beta)
computes on the path ``cledb_invproc -> cledb_matchiquv / cledb_matchiqud`` and of its two elementwise
neighbours ``blos_proc_1pix`` and ``obs_qurotate``.  It is the checker the CUDA path is compared against.
Nothing in the product (``solar-coronal-inversion_b200/``) imports it; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

Parity status: PINNED.  ``oracle/run_reference.py`` imports the unmodified reference from /root/reference
(in the build container, where it exists), runs it on seeded synthetic inputs and on the reference's own
unit-test pixel, and stores its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` asserts that
this restatement reproduces those outputs bit-for-bit (indices, chi^2, all 11 physics columns, sfound,
issuemask) and reproduces the known answers of tests/test_functions.py / tests/debug_unittests.ipynb.

Every function cites the reference lines it follows (paths relative to the reference root).
"""
import numpy as np

# chi^2 weights, CLEDB_PROC.py:981 / 1156 (float64 because of the 0.01 literal)
WEIGHTS = np.array((1, 1, 1, 0, 0.01, 1, 1, 0), dtype=np.float64)


# --------------------------------------------------------------------------------------------------
# index helpers and physics of one database entry
# --------------------------------------------------------------------------------------------------
def db_unravel(index, dbcgrid, dbtype):
    """flat index -> (i, j, k, l) of a database, CLEDB_PROC.py:1634-1659."""
    ngx, nbphi, nbtheta = np.int32(dbcgrid[2]), np.int32(dbcgrid[3]), np.int32(dbcgrid[4])
    per_j = nbphi * nbtheta
    per_i = ngx * per_j
    i = index // per_i if dbtype == 0 else 0
    rem = index - i * per_i
    j = rem // per_j
    rem2 = index - i * per_i - j * per_j
    k = rem2 // nbtheta
    l = index - i * per_i - j * per_j - k * nbtheta
    return i, j, k, l


def baumbach_density(r):
    """CLEDB_PROC.py:1689-1697 (only reached for CLE-type databases)."""
    return np.float32(3.e8 * np.exp(-(r - 1.) / 7.18401074e-02) + 1.e8 * (0.036 / r ** 1.5 + 1.55 / r ** 6.))


def entry_physics(index, yobs, dobs, dbhdr, bcalc, b):
    """[ne, y, gx, B, phi, theta, Bx, By, Bz] of database entry ``index``; CLEDB_PROC.py:1704-1749.

    ``index`` must be a NumPy int64 scalar as in the reference (CLEDB_PROC.py:1082), which makes the
    grid arithmetic float64 before the final float32 cast.
    """
    dbcgrid, ned, ngx, nbphi, nbtheta, xed, gxmin, gxmax, bphimin, bphimax, \
        bthetamin, bthetamax, nline, wavel, dbtype = dbhdr
    i, j, k, l = db_unravel(index, dbcgrid, dbtype)
    gx = gxmin + j * (gxmax - gxmin) / (ngx - 1)
    bphi = bphimin + k * (bphimax - bphimin) / (nbphi - 1)
    btheta = bthetamin + l * (bthetamax - bthetamin) / (nbtheta - 1)
    if dbtype == 0:
        ed = np.log10(xed[i] * baumbach_density(np.sqrt(yobs * yobs + gx * gx)))
    else:
        ed = dobs
    phs = np.array((ed, gx, bphi, btheta), dtype=np.float32)
    if bcalc == 3:     # plane-of-sky field from the Doppler/wave observation, CLEDB_PROC.py:1712-1713
        b = b / np.sqrt(np.sin(phs[3]) ** 2 * np.sin(phs[2]) ** 2 + np.cos(phs[3]) ** 2)
    bx = np.abs(b) * np.sin(phs[3]) * np.cos(phs[2])
    by = np.abs(b) * np.sin(phs[3]) * np.sin(phs[2])
    bz = np.abs(b) * np.cos(phs[3])
    return np.array((phs[0], yobs, phs[1], b, phs[2], phs[3], bx, by, bz), dtype=np.float32)


def derotate_entry(row, ang, nf):
    """Matched database row scaled back by ``nf`` and Q,U rotated by -2*ang; CLEDB_PROC.py:1575-1588."""
    out = np.copy(row) * nf
    c, s = np.cos(-2. * ang), np.sin(-2. * ang)
    out[1] = row[1] * nf * c - row[2] * nf * s
    out[2] = row[1] * nf * s + row[2] * nf * c
    out[5] = row[5] * nf * c - row[6] * nf * s
    out[6] = row[5] * nf * s + row[6] * nf * c
    return out


# --------------------------------------------------------------------------------------------------
# chi^2 and selection
# --------------------------------------------------------------------------------------------------
def chi2_of_rows(rows, sobs, snr, nf):
    """Reduced chi^2 of every row of ``rows`` [M,8] against one observation; CLEDB_PROC.py:989-1005.

    float32 residual ratio, float64 weighting/square/sum (NumPy pairwise order over 8 terms), /2, 15 decimals.
    """
    with np.errstate(all='ignore'):
        target = sobs / nf
        sigma = target / snr
        resid = (rows[:, :8] - target) / sigma            # float32
        wres = resid * WEIGHTS                            # float64
        wres *= wres
        return np.around(np.sum(wres, axis=1) / 2, 15)


def selection_positions(chi, nsearch):
    """Positions returned by the reference's partial selection sort, including its index quirk.

    CLEDB_PROC.py:1595-1627: position ``argmin(tmp[i:]) + i`` of a copy that is being swapped in place, so a
    value displaced by an earlier swap is reported at its *new* position.
    For nsearch > 100 the reference switches to a full argsort (CLEDB_PROC.py:1010-1013).
    """
    if nsearch > 100:
        return np.argsort(chi)[0:nsearch]
    work = np.array(chi, copy=True)
    found = np.zeros(nsearch, dtype=np.int32)
    for i in range(nsearch):
        p = int(np.argmin(work[i:])) + i
        found[i] = p
        work[i], work[p] = work[p], work[i]
    return found


def stable_topk(chi, nsearch):
    """The mathematically intended answer: the nsearch smallest, ties to the lowest position."""
    return np.argsort(chi, kind='stable')[:nsearch]


def _null_result(nsearch, chi=None):
    out = np.zeros((nsearch, 11), dtype=np.float32)
    out[:, 0] = -1
    if chi is not None:
        out[:, 1] = chi
    return out, np.zeros((nsearch, 8), dtype=np.float32)


def reduced_subset(mode, sobs, sobsd, dbhdr, database, nsearch):
    """The per-pixel reduced database of ``reduced=True``; CLEDB_PROC.py:1401-1478 (IQUV) / 1487-1568 (IQUD).

    For every phi index the ``nsearch`` theta entries whose ``tan(theta) sin(phi)`` is closest to the observed
    ``tan(Phi_B)`` and the ``nsearch`` closest to the degenerate branch ``tan(pi - Phi_B)`` are kept (for all gx, and all
    densities of a CLE-type database).  Returns ``(datasel[M,8] float32, outredindex[nbphi, 2 nsearch] int32)``; rows of
    the phi = 0 plane stay zero when both branches start at theta index 0 (CLEDB_PROC.py:1446, 1466).  Ties are
    ordered by NumPy's default ``argsort`` exactly as in the reference (not stable: machine dependent).
    """
    dbcgrid, ned, ngx, nbphi, nbtheta, xed, gxmin, gxmax, bphimin, bphimax, bthetamin, bthetamax, nline, wavel, dbtype = dbhdr
    kk = np.arange(0, nbphi)
    bphir = bphimin + kk * (bphimax - bphimin) / (nbphi)
    ll = np.arange(0, nbtheta)
    bthetar = bthetamin + ll * (bthetamax - bthetamin) / (nbtheta)
    tt = np.tan(bthetar)
    with np.errstate(all='ignore'):
        phib = 0.5 * np.arctan2(sobs[2], sobs[1]) if mode == 0 else sobsd[1]      # :1428 / :1516
        tmain = np.tan(phib)
        tdeg = np.tan(np.pi - phib)
    db = np.asarray(database)
    lead = db.shape[:-3]                                  # (ngx,) for PyCELP, (ned, ngx) for CLE
    datasel = np.zeros(lead + (nbphi, 2 * nsearch, 8), dtype=np.float32)
    outredindex = np.zeros((nbphi, 2 * nsearch), dtype=np.int32)
    with np.errstate(all='ignore'):
        for ir in range(bphir.shape[0]):
            ttp = tt * np.sin(bphir[ir])
            srta = np.argsort(np.abs(tmain - ttp))[0:nsearch]
            srtb = np.argsort(np.abs(tdeg - ttp))[0:nsearch]
            if ir + srta[0] > 0 or ir + srtb[0] > 0:
                datasel[..., ir, :nsearch, :] = db[..., ir, srta, :]
                datasel[..., ir, nsearch:, :] = db[..., ir, srtb, :]
                outredindex[ir, :nsearch] = srta
                outredindex[ir, nsearch:] = srtb
    return np.reshape(datasel, (-1, 8)), outredindex


def reduced_original_index(pos, outredindex, dbhdr):
    """Position in the reduced database -> row of the full database; CLEDB_PROC.py:1070-1071 (cledb_params on the grid with
    ``nbtheta`` replaced by ``2 nsearch``, then cledb_invparams with the stored theta index)."""
    dbcgrid, dbtype = dbhdr[0], dbhdr[14]
    red = np.array((dbcgrid[0], dbcgrid[1], dbcgrid[2], dbcgrid[3], outredindex.shape[1], 0, 0, 0, 0, 0, 0, 0, 0, 0), dtype=np.float32)
    i, j, k, l = db_unravel(pos, red, dbtype)
    return np.int32(i * dbcgrid[2] * dbcgrid[3] * dbcgrid[4] + j * dbcgrid[3] * dbcgrid[4] + k * dbcgrid[4] + outredindex[k, l])


def match_pixel(mode, sobs, sobsd, yobs, aobs, dobs, database, dbhdr, snr, nsearch, maxchisq, bcalc,
                strict_topk=False, return_raw=False, reduced=False):
    """One pixel of the 2-line search.  mode 0 = IQUV (CLEDB_PROC.py:921-1087), 1 = IQUD (1094-1252).

    Full-database mode (``reduced=False``).  Returns ``(out[nsearch,11], smatch[nsearch,8])`` float32; with
    ``return_raw`` also the dict {keep, chi, positions} used by the GPU parity tests.
    """
    sobs = np.asarray(sobs)
    if mode == 0:
        empty = np.isnan(sobs).any() or np.count_nonzero(sobs) == 0          # :944
    else:
        empty = np.isnan(sobs).all() or np.count_nonzero(sobs) == 0          # :1118
    if empty:
        res = _null_result(nsearch)
        return res + (None,) if return_raw else res

    outredindex = None
    if reduced:
        flat, outredindex = reduced_subset(mode, sobs, sobsd, dbhdr, database, nsearch)   # :954 / :1128
    else:
        flat = np.reshape(database, (-1, 8))                                  # :969-972 / 1141-1144
    with np.errstate(all='ignore'):
        if mode == 0:
            keep = np.flatnonzero(np.sign(flat[:, 3]) == np.sign(sobs[3]))    # :974
            nf = sobs[0]                                                      # :980
        else:
            keep = np.flatnonzero(flat[:, 3] == flat[:, 3])                   # :1148 (drops NaN rows only)
            nf = sobs[np.argwhere(sobs == np.max(sobs))[0, 0]]                # :1155
    rows = flat[keep]
    chi = chi2_of_rows(rows, sobs, snr, nf)
    pos = stable_topk(chi, nsearch) if strict_topk else selection_positions(chi, nsearch)
    raw = dict(keep=keep, chi=chi, positions=pos, nf=nf)

    best = chi[pos[0]]
    limit = maxchisq * 10 if mode == 0 else maxchisq                          # :1027 vs :1202
    if best > limit:
        res = _null_result(nsearch, best)
        return res + (raw,) if return_raw else res

    out = np.zeros((nsearch, 11), dtype=np.float32)
    out[:, 0] = -1
    smatch = np.zeros((nsearch, 8), dtype=np.float32)
    with np.errstate(all='ignore'):
        for si in range(nsearch):
            p = pos[si]
            c = chi[p]
            if not (c <= maxchisq):                                           # :1044 / 1219
                continue
            row = rows[p]
            if mode == 1:
                b = sobsd[0]                                                  # :1225
            else:
                def ratio(comp):                                              # :1051-1059
                    return (np.abs(row[comp]) > 1e-7 and sobs[comp] / nf / row[comp]) or 0
                if bcalc == 0:
                    b = ratio(3)
                elif bcalc == 1:
                    b = ratio(7)
                else:
                    b = 0.5 * (ratio(7) + ratio(3))
            smatch[si, :] = derotate_entry(row, aobs, nf)                     # :1064
            ix = reduced_original_index(keep[p], outredindex, dbhdr) if reduced else keep[p]   # :1070-1071
            out[si, 0] = ix                                                   # :1080
            out[si, 1] = c
            out[si, 2:] = entry_physics(ix, yobs, dobs, dbhdr, bcalc, b)      # :1082
    return (out, smatch, raw) if return_raw else (out, smatch)


# --------------------------------------------------------------------------------------------------
# map driver: cledb_invproc
# --------------------------------------------------------------------------------------------------
def bits_of(mask):
    """Set of power-of-two issue codes present in one mask value; CLEDB_PREPINV.py:1563-1576."""
    x = int(mask)
    return {1 << b for b in range(x.bit_length()) if (x >> b) & 1} if x > 0 else set()


def van_vleck_flag(out):
    """Van-Vleck proximity of the LAST solution only (``flg`` is reset inside the loop); CLEDB_PROC.py:323-343."""
    flg = 0
    for j in range(out.shape[0]):
        flg = 0
        alpha = -np.arctan2(out[j, 4], out[j, 3])
        beta = np.pi / 2
        xb = np.sin(out[j, 7]) * np.cos(out[j, 6])
        yb = np.sin(out[j, 7]) * np.sin(out[j, 6])
        zb = np.cos(out[j, 7])
        yyb = np.cos(beta) * yb - np.sin(beta) * zb
        yzb = np.sin(beta) * yb + np.cos(beta) * zb
        xxb = np.cos(alpha) * xb + np.sin(alpha) * yzb
        zzb = -np.sin(alpha) * xb + np.cos(alpha) * yzb
        vartheta = np.arctan2(np.sqrt(xxb ** 2 + yyb ** 2), zzb)
        if 53 <= vartheta * 180 / np.pi <= 56:
            flg = 1
    return flg


def _pixel_job(args):
    (mode, xx, yy, sobs, sobsd, yobs, aobs, dobs, dbk, dbhdr, snr, nsearch, maxchisq, bcalc, strict, reduced) = args
    out, smatch = match_pixel(mode, sobs, sobsd, yobs, aobs, dobs, _DBS[dbk], dbhdr, snr, nsearch, maxchisq,
                              bcalc, strict_topk=strict, reduced=reduced)
    return xx, yy, out, smatch


_DBS = None


def invert_map(sobs_totrot, sobs_dopp, database, db_enc, yobs, aobs, dobs, snr, issuemask, dbhdr, keyvals,
               nsearch, maxchisq, bcalc, iqud, workers=1, strict_topk=False, pixels=None, reduced=False):
    """``cledb_invproc`` (CLEDB_PROC.py:93-363), full-database mode.

    ``workers`` > 1 uses a fork pool whose children inherit the databases (the "IPC-free" CPU baseline of
    BASELINE.md section 3); the reference itself pickles the database into every task (CLEDB_PROC.py:306).
    ``pixels``: optional list of (xx,yy) to restrict the work to (others stay zero) - used for sub-sampled
    parity checks on big maps.  Mutates ``issuemask`` like the reference.
    """
    global _DBS
    nx, ny = keyvals[0:2]
    invout = np.zeros((nx, ny, nsearch, 11), dtype=np.float32)
    sfound = np.zeros((nx, ny, nsearch, 8), dtype=np.float32)
    if np.shape(sobs_totrot)[-1] != 8 or database[0].shape[-1] != 8:          # :166-171
        invout[:, :, :, 0] = -1
        return invout, sfound
    if (iqud == True and bcalc != 3) or (iqud != True and bcalc == 3):        # :174-178
        invout[:, :, :, 0] = -1
        return invout, sfound
    mode = 1 if iqud == True else 0
    if pixels is None:
        pixels = [(xx, yy) for xx in range(nx) for yy in range(ny)]
    jobs = [(mode, xx, yy, sobs_totrot[xx, yy, :], sobs_dopp[xx, yy, :] if (mode == 1 or (reduced and np.ndim(sobs_dopp) >= 3)) else None,
             yobs[xx, yy], aobs[xx, yy], dobs[xx, yy], int(db_enc[xx, yy]), dbhdr, snr[xx, yy, :],
             nsearch, maxchisq, bcalc, strict_topk, reduced) for (xx, yy) in pixels]
    _DBS = database
    if workers > 1:
        import multiprocessing
        with multiprocessing.get_context('fork').Pool(workers) as pool:
            results = pool.map(_pixel_job, jobs, chunksize=max(1, len(jobs) // (workers * 8)))
    else:
        results = [_pixel_job(j) for j in jobs]
    with np.errstate(all='ignore'):
        for xx, yy, out, smatch in results:
            invout[xx, yy] = out
            sfound[xx, yy] = smatch
            flg = van_vleck_flag(out)
            for line in range(2):                                             # :347-355
                if flg and 512 not in bits_of(issuemask[xx, yy, line]):
                    issuemask[xx, yy, line] += 512
                if out[0, 1] > 10 and 1024 not in bits_of(issuemask[xx, yy, line]):
                    issuemask[xx, yy, line] += 1024
                if snr[xx, yy, 0] < 1 and 2048 not in bits_of(issuemask[xx, yy, line]):
                    issuemask[xx, yy, line] += 2048
    return invout, sfound


# --------------------------------------------------------------------------------------------------
# raw search (what the CUDA kernel returns before solution building) for big parity cases
# --------------------------------------------------------------------------------------------------
def raw_topk(mode, sobs, snr, database, nsearch, return_all=False):
    """(orig_index[k], chi[k] float64, kept_position[k], ncand) of the stable top-k of one pixel, or None for
    a null pixel.  Same arithmetic as ``match_pixel``; this is the quantity ``cledb_match`` (C-ABI) returns."""
    sobs = np.asarray(sobs)
    if (np.isnan(sobs).any() if mode == 0 else np.isnan(sobs).all()) or np.count_nonzero(sobs) == 0:
        return None
    flat = np.reshape(database, (-1, 8))
    with np.errstate(all='ignore'):
        if mode == 0:
            keep = np.flatnonzero(np.sign(flat[:, 3]) == np.sign(sobs[3]))
            nf = sobs[0]
        else:
            keep = np.flatnonzero(flat[:, 3] == flat[:, 3])
            nf = sobs[np.argwhere(sobs == np.max(sobs))[0, 0]]
    chi = chi2_of_rows(flat[keep], sobs, snr, nf)
    pos = stable_topk(chi, nsearch)
    if return_all:
        return keep[pos], chi[pos], pos, keep.size, chi
    return keep[pos], chi[pos], pos, keep.size


# --------------------------------------------------------------------------------------------------
# database side: what sdb_preprocess hands to the search
# --------------------------------------------------------------------------------------------------
def normalise_database(line1, line2, wl0, dbtype=1):
    """Concatenate the two per-line IQUV arrays and divide by I1; PyCELP Stokes V additionally scaled by the
    nm/Angstrom factor.  CLEDB_PREPINV.py:306-317 (component 0 is divided last and becomes exactly 1)."""
    db = np.append(line1, line2, axis=-1)
    with np.errstate(all='ignore'):
        scale = 10 ** np.rint(np.log10(wl0 / 10000))
        for comp in range(7, -1, -1):
            if dbtype == 1 and comp in (3, 7):
                db[..., comp] = scale * db[..., comp] / db[..., 0]
            else:
                db[..., comp] = db[..., comp] / db[..., 0]
    return db


def nearest_database(y, d, dbynumbers):
    """File index of the database closest in (height, density); CLEDB_PREPINV.py:458-468, including its
    slice that never selects the last density of a height."""
    dist = np.abs(dbynumbers[:, 1] / 100. - y)
    rows = np.argwhere(dist == np.min(dist))
    first, last = rows[0][0], rows[-1][0]
    dd = np.argmin(np.abs(dbynumbers[first:last, 2] / 100 - d))
    return np.int32(dbynumbers[first + dd, 0])


def file_table(names):
    """(index, 100*height, 100*log10 density) of a sorted list of database file names; CLEDB_PREPINV.py:402-412."""
    at = str.find(names[0], 'DB_h')
    table = np.empty((len(names), 3), dtype=np.float32)
    for i in range(len(names)):
        table[i, 0] = i
        table[i, 1] = np.float32(names[i][at + 4:at + 7])
        table[i, 2] = np.float32(names[i][at + 9:-4])
    return table


def select_databases(yobs, dobs, dbynumbers):
    """``db_enc_flnm`` of sdb_preprocess (CLEDB_PREPINV.py:262-283): ``nearest_database`` for every pixel of a map."""
    yobs, dobs = np.asarray(yobs), np.asarray(dobs)
    out = np.zeros(yobs.shape, dtype=np.int32)
    for pos in np.ndindex(*yobs.shape):
        out[pos] = nearest_database(yobs[pos], dobs[pos], dbynumbers)
    return out


# --------------------------------------------------------------------------------------------------
# elementwise neighbours
# --------------------------------------------------------------------------------------------------
def blos_pixel(iquv, const):
    """Analytic 1-line B_LOS: 3 magnetograph variants + azimuth and the two issue flags; CLEDB_PROC.py:874-914."""
    out = np.zeros(4, dtype=np.float32)
    flags = [0, 0]
    with np.errstate(all='ignore'):
        if (iquv[0] != 0) and (np.isnan(iquv[:]).any() == False):
            lpol = np.sqrt(iquv[1] ** 2 + iquv[2] ** 2)
            kk = 1.e9 * (const.planckconst * const.l_speed / const.bohrmagneton / (const.line_ref ** 2))
            out[0] = kk * (-iquv[3] / (const.g_eff * (iquv[0] - lpol) - 0.66 * const.F_factor * (lpol / 1.)))
            out[1] = kk * (-iquv[3] / (const.g_eff * (iquv[0] + lpol) + 0.66 * const.F_factor * (lpol / 1.)))
            out[2] = kk * (-iquv[3] / (const.g_eff * iquv[0]))
            out[3] = 0.5 * np.arctan2(iquv[2], iquv[1])
            qi, ui = iquv[1] / iquv[0], iquv[2] / iquv[0]
            if (np.abs(qi) < 1e-6) or (np.abs(ui) < 1e-6):
                flags[0] = 1
                if qi < 1e-6 and ui < 1e-6:
                    flags[1] = 1
    return out, flags


def blos_map(sobs_tot, issuemask, keyvals, consts_module):
    """``blos_proc`` (CLEDB_PROC.py:388-486) for IQUV data: blosout[nx,ny,nline,4]; issue codes 32 / 64."""
    nx, ny = keyvals[0:2]
    nline, tline = keyvals[3], keyvals[4]
    blosout = np.zeros((nx, ny, nline, 4), dtype=np.float32)
    for zz in range(nline):
        const = consts_module.Constants(tline[zz])
        for xx in range(nx):
            for yy in range(ny):
                vals, flags = blos_pixel(sobs_tot[xx, yy, 4 * zz:4 * (zz + 1)], const)
                blosout[xx, yy, zz, :] = vals
                if flags[0] != 0 and 32 not in bits_of(issuemask[xx, yy, zz]):
                    issuemask[xx, yy, zz] += 32
                if flags[1] != 0 and 64 not in bits_of(issuemask[xx, yy, zz]):
                    issuemask[xx, yy, zz] += 64
    return blosout


def qurotate_map(sobs_tot, keyvals, hpxygrid):
    """``obs_qurotate`` + ``obs_calcheight`` (CLEDB_PREPINV.py:1277-1350, 897-926): (sobs_totrot, aobs, yobs)."""
    nx, ny = keyvals[0:2]
    crpix1, crpix2 = keyvals[5:7]
    crval1, crval2 = keyvals[8:10]
    cdelt1, cdelt2 = keyvals[11:13]
    linpolref = keyvals[14]
    if crpix1 != 0:
        crval1 = crval1 - (crpix1 * cdelt1)
    if crpix2 != 0:
        crval2 = crval2 - (crpix2 * cdelt2)
    aobs = np.empty((nx, ny), dtype=np.float32)
    yobs = np.empty((nx, ny), dtype=np.float32)
    rot = np.copy(sobs_tot)
    grid = np.sum(hpxygrid) != 0.
    for xx in range(nx):
        for yy in range(ny):
            if grid:
                px, py = hpxygrid[0, xx, yy], hpxygrid[1, xx, yy]
            else:
                px, py = crval1 + xx * cdelt1, crval2 + yy * cdelt2
            yobs[xx, yy] = np.sqrt(px ** 2 + py ** 2)
            aobs[xx, yy] = -np.arctan2(px, py)
            if aobs[xx, yy] < 0:
                aobs[xx, yy] = 2 * np.pi + aobs[xx, yy]
            if aobs[xx, yy] < linpolref:
                aobs[xx, yy] = 2 * np.pi - aobs[xx, yy]
            else:
                aobs[xx, yy] = aobs[xx, yy] + linpolref
            c, s = np.cos(2 * aobs[xx, yy]), np.sin(2 * aobs[xx, yy])
            for q in (1, 5):
                rot[xx, yy, q] = sobs_tot[xx, yy, q] * c - sobs_tot[xx, yy, q + 1] * s
                rot[xx, yy, q + 1] = sobs_tot[xx, yy, q] * s + sobs_tot[xx, yy, q + 1] * c
    return rot, aobs, yobs


# --------------------------------------------------------------------------------------------------
# spectral integration: obs_integrate / obs_integrate_work_1pix / obs_cdf
# --------------------------------------------------------------------------------------------------
def spectrum_cdf(spectr):
    """Normalised running sum of one Stokes I spectrum, float32; CLEDB_PREPINV.py:1356-1364 (every prefix is its own
    ``np.sum``, i.e. NumPy's pairwise summation of that prefix, rounded to float32)."""
    cdf = np.zeros((spectr.shape[0]), dtype=np.float32)
    for i in range(0, spectr.shape[0]):
        cdf[i] = np.sum(spectr[0:i + 1])
    return cdf[:] / cdf[-1]


def integrate_pixel(nw, wlarr, wcont, dwv, spec):
    """One pixel of one line: (sobs_tot[4] float64, snr[4], background[4], issue code); CLEDB_PREPINV.py:1164-1270.

    ``spec[nw,4]``: Stokes IQUV spectra.  Background = mean of three continuum samples at ``wcont``; I, Q, U are summed
    times the dispersion ``dwv``; V is the median over the inner half of the window of V / (d(I/Itot)/dlambda + 1e-36);
    SNR = rms of the middle third of the signal window (0.3 % .. 99.7 % of the Stokes I running sum) over the rms outside it.
    """
    zeros = (np.zeros((4), dtype=np.float32), np.zeros((4), dtype=np.float32), np.zeros((4), dtype=np.float32), 0)
    if not ((np.count_nonzero(spec[:, 0])) and (np.isfinite(spec[:, :]).all())):
        return zeros
    with np.errstate(all='ignore'):
        cdf = spectrum_cdf(spec[:, 0])
        lr0 = np.argwhere((np.abs(cdf) > 0.003) & (np.abs(cdf) < 0.997))
        background = np.mean(spec[wcont:wcont + 3:, :], axis=0)
        tmp2 = spec[:, :] - background
        tmp2[tmp2[:, 0] < 0, 0] = 0
        tot = np.zeros((4), dtype=np.float64)
        tot[0] = np.sum(tmp2[:, 0] * dwv)
        slope = np.gradient(tmp2[:, 0] / tot[0], wlarr)
        lo, hi = np.int32(0.25 * nw), np.int32(0.75 * nw)
        vscl = tmp2[lo:hi, 3] / (slope[lo:hi] + 1e-36)
        tot[3] = np.median(vscl[np.isfinite(vscl)])
        tot[1] = np.sum(tmp2[:, 1] * dwv)
        tot[2] = np.sum(tmp2[:, 2] * dwv)
        rms_noise = np.sqrt(np.mean(np.concatenate((spec[0:lr0[0, 0] + 1, :], spec[lr0[-1, 0]:, :])) ** 2))
        a = np.round(lr0[0, 0] + lr0.shape[0] / 3).astype(np.int16)
        b = np.round(lr0[0, 0] + lr0.shape[0] * 2 / 3).astype(np.int16)
        rms_signal = np.sqrt(np.mean(spec[a:b, :] ** 2, axis=0))
        snr = rms_signal / rms_noise
    iss = 0
    for kk in range(3):
        if snr[kk] < 1:
            iss = 1
    if snr[3] < 1:
        iss += 2
    return tot, snr, background, iss


def continuum_samples(wlarr, nline):
    """Per line: index of the clean continuum window; CLEDB_PREPINV.py:968-976.  None if a line is not Fe XIII 1074/1079."""
    out = []
    for zz in range(nline):
        if 1074.65 > wlarr[:, zz].min() and 1074.65 < wlarr[:, zz].max():
            out.append(np.argmin(np.abs(wlarr[:, zz] - 1074.)))
        elif 1079.8 > wlarr[:, zz].min() and 1079.8 < wlarr[:, zz].max():
            out.append(np.argmin(np.abs(wlarr[:, zz] - 1080.2)))
    return out if len(out) else None


def integrate_map(sobs_in, wlarr, keyvals):
    """``obs_integrate`` without the atmospheric reduction branch (CLEDB_PREPINV.py:935-1063): (sobs_tot, snr, background,
    issuemask) for a list of per-line arrays ``[nx,ny,nw,4]``."""
    nx, ny, nw = keyvals[0:3]
    nline = keyvals[3]
    sobs_tot = np.zeros((nx, ny, nline * 4), dtype=np.float32)
    background = np.zeros((nx, ny, nline * 4), dtype=np.float32)
    snr = np.ones((nx, ny, nline * 4), dtype=np.float32)
    iss = np.zeros((nx, ny, nline), dtype=np.int32)
    wcont = continuum_samples(wlarr, nline)
    if wcont is None:
        return -1, -1, -1
    for zz in range(nline):
        dwv = np.mean(np.diff(wlarr[:, zz]))
        for xx in range(nx):
            for yy in range(ny):
                t, s, b, i = integrate_pixel(nw, wlarr[:, zz], wcont[zz], dwv, sobs_in[zz][xx, yy, :, :])
                sobs_tot[xx, yy, 4 * zz:4 * (zz + 1)], snr[xx, yy, 4 * zz:4 * (zz + 1)] = t, s
                background[xx, yy, 4 * zz:4 * (zz + 1)], iss[xx, yy, zz] = b, i
    return sobs_tot, snr, background, iss


def density_map(sobs_totrot, yobs, keyvals, table):
    """``obs_dens`` (CLEDB_PREPINV.py:1371-1479): log10 density from the 1074/1079 Stokes I ratio and the CHIANTI table
    ``{'h','den','rat'}``, one quadratic ``interp1d`` per pixel exactly like the reference.  Needs scipy."""
    import scipy.interpolate
    nx, ny = keyvals[0:2]
    dobs = np.empty((nx, ny), dtype=np.float32)
    iss = np.zeros((nx, ny), dtype=np.int32)
    with np.errstate(all='ignore'):
        for xx in range(nx):
            for yy in range(ny):
                a_obs, b_obs, y_obs = sobs_totrot[xx, yy, 0], sobs_totrot[xx, yy, 4], yobs[xx, yy]
                if b_obs == 0:
                    dobs[xx, yy] = 1
                    continue
                rat_obs = a_obs / b_obs
                if np.isnan(rat_obs) or np.isinf(rat_obs):
                    dobs[xx, yy] = 0
                    continue
                subh = np.argwhere(table['h'] > y_obs)
                if len(subh) == 0:
                    subh = [-1]
                    iss[xx, yy] = 4
                ifunc = scipy.interpolate.interp1d(table['rat'][subh[0], :].flatten(), table['den'], kind="quadratic",
                                                   fill_value="extrapolate")
                dobs[xx, yy] = ifunc(rat_obs)
        dobs[dobs <= 0] = 1
        return np.round(np.log10(dobs), 3), iss
