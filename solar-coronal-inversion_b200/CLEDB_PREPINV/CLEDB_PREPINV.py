"""B200-native drop-in for the hot-path neighbours in the reference's ``CLEDB_PREPINV/CLEDB_PREPINV.py``.

    obs_qurotate(sobs_tot, keyvals, hpxygrid) -> (sobs_totrot, aobs)      CLEDB_PREPINV.py:1277-1350
    obs_calcheight(keyvals, hpxygrid)          -> yobs                    CLEDB_PREPINV.py:897-926
    sdb_preprocess(yobs, dobs, keyvals, wlarr, params) -> (db_enc, database, dbhdr)   CLEDB_PREPINV.py:201-367
    sdb_fileingest / sdb_findelongationanddens / sdb_read / sdb_parseheader / sdb_lcgrid   CLEDB_PREPINV.py:380-569
    obs_integrate(sobs_in, wlarr, keyvals, verbose, ncpu, atmred) -> (sobs_tot, snr, background, issuemask)   CLEDB_PREPINV.py:935-1063
    obs_dens(sobs_totrot, yobs, keyvals, params) -> (dobs, iss_dens)      CLEDB_PREPINV.py:1371-1430
    iss_block(x)                                                          CLEDB_PREPINV.py:1563-1576

The rotation and the height map are one fused elementwise kernel (72 B/pixel) behind the C-ABI.  ``sdb_preprocess``
picks the (height, density) database of every pixel with one kernel over the whole map (``cledb_db_select``) instead of
a pool task per pixel, loads and normalises only the distinct files, and uploads them to HBM once, so that
``cledb_invproc`` finds them resident.  Instrument header ingestion, spectral integration and the density look-up
of ``sobs_preprocess`` are outside the scope of this build.
"""
import glob
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from cledb_b200 import native
from CLEDB_PROC.CLEDB_PROC import OPTIONS, iss_block  # noqa: F401  (same helper lives in both reference modules)


def _pointing(keyvals):
    """Per-row / per-column coordinates evaluated in the scalar types the header carries (Python floats in the
    reference's own tests), so that the float64 arithmetic of CLEDB_PREPINV.py:1288-1301 is reproduced exactly."""
    nx, ny = keyvals[0:2]
    crpix1, crpix2 = keyvals[5:7]
    crval1, crval2 = keyvals[8:10]
    cdelt1, cdelt2 = keyvals[11:13]
    if crpix1 != 0:
        crval1 = crval1 - (crpix1 * cdelt1)
    if crpix2 != 0:
        crval2 = crval2 - (crpix2 * cdelt2)
    xs = np.array([crval1 + xx * cdelt1 for xx in range(nx)], dtype=np.float64)
    ys = np.array([crval2 + yy * cdelt2 for yy in range(ny)], dtype=np.float64)
    return nx, ny, xs, ys


def _rotate(sobs_tot, keyvals, hpxygrid):
    nx, ny, xs, ys = _pointing(keyvals)
    hp = None
    if hpxygrid is not None and np.sum(hpxygrid) != 0.:
        hp = np.ascontiguousarray(hpxygrid, dtype=np.float32)[:2]
    ctx = native.get_context(OPTIONS['device'])
    return ctx.qurotate(np.ascontiguousarray(sobs_tot, dtype=np.float32).reshape(nx, ny, 8), xs, ys, keyvals[14], hp)


def obs_qurotate(sobs_tot, keyvals, hpxygrid):
    """Rotate Q,U of both lines into the database frame; returns ``(sobs_totrot, aobs)``."""
    rot, aobs, _ = _rotate(sobs_tot, keyvals, hpxygrid)
    return rot, aobs


def obs_calcheight(keyvals, hpxygrid):
    """Apparent height of every pixel, ``yobs[nx,ny]`` float32."""
    nx, ny = keyvals[0:2]
    _, _, yobs = _rotate(np.zeros((nx, ny, 8), dtype=np.float32), keyvals, hpxygrid)
    return yobs


def obs_qurotate_height(sobs_tot, keyvals, hpxygrid):
    """Both products of the fused kernel in one pass: ``(sobs_totrot, aobs, yobs)``."""
    return _rotate(sobs_tot, keyvals, hpxygrid)


def sdb_lcgrid(mmin, mmax, n):
    """Logarithmic density grid of a CLE database header; CLEDB_PREPINV.py:560-569."""
    ff = 10. ** ((mmax - mmin) / (n - 1))
    xf = np.empty(n, np.float32)
    xf[0] = 10. ** mmin
    for k in range(1, n):
        xf[k] = xf[k - 1] * ff
    return xf


def sdb_parseheader(dbheaderfile):
    """Parse ``db.hdr`` into the 15-tuple every hot-path function receives as ``dbhdr``."""
    g = np.fromfile(dbheaderfile, dtype=np.float32, sep=' ')
    ned, ngx, nbphi, nbtheta = (np.int32(g[i]) for i in (1, 2, 3, 4))
    nline = np.int32(g[13])
    wavel = np.array([g[14 + k] for k in range(nline)], dtype=np.float32)
    return (g, ned, ngx, nbphi, nbtheta, sdb_lcgrid(np.float32(g[5]), np.float32(g[6]), ned),
            np.float32(g[7]), np.float32(g[8]), np.float32(g[9]), np.float32(g[10]), np.float32(g[11]), np.float32(g[12]),
            nline, wavel, np.int32(g[-1]))


# ---------------------------------------------------------------------------------------------------------
# spectral integration of the observation
# ---------------------------------------------------------------------------------------------------------
def obs_integrate(sobs_in, wlarr, keyvals, verbose, ncpu, atmred):
    """Background subtraction, wavelength integration and signal statistics of every pixel; CLEDB_PREPINV.py:935-1063.

    ``sobs_in``: list of per-line arrays ``[nx,ny,nw,4]``; ``wlarr[nw,nline]``.  Returns ``(sobs_tot[nx,ny,4*nline],
    snr, background)`` float32 and ``issuemask_obsint[nx,ny,nline]`` int32, or ``(-1,-1,-1)`` for an unknown wavelength
    range (:978-980).  One kernel launch per line replaces the pool task per pixel (cledb_obs_integrate); the result is
    bit-identical for float64 spectra.  The telluric/photospheric reduction branch (``atmred=True``) is unfinished in the
    reference and needs downloads; it is passed through to the reference module.
    """
    if atmred == True:
        from cledb_b200.passthrough import reference_attr
        return reference_attr('CLEDB_PREPINV/CLEDB_PREPINV.py', 'obs_integrate')(sobs_in, wlarr, keyvals, verbose, ncpu, atmred)
    nx, ny, nw = keyvals[0:3]
    nline = keyvals[3]
    wlarr = np.asarray(wlarr)
    wcont = []
    for zz in range(nline):                                                              # :968-976
        lo, hi = wlarr[:, zz].min(), wlarr[:, zz].max()
        if 1074.65 > lo and 1074.65 < hi:
            wcont.append(np.argmin(np.abs(wlarr[:, zz] - 1074.)))
        elif 1079.8 > lo and 1079.8 < hi:
            wcont.append(np.argmin(np.abs(wlarr[:, zz] - 1080.2)))
    if len(wcont) == 0:
        if verbose >= 1:
            print("OBS_INTEGRATE: FATAL! Wavelength case not handled for atlas fitting!")
        return -1, -1, -1
    sobs_tot = np.zeros((nx, ny, nline * 4), dtype=np.float32)
    background = np.zeros((nx, ny, nline * 4), dtype=np.float32)
    snr = np.ones((nx, ny, nline * 4), dtype=np.float32)
    issuemask_obsint = np.zeros((nx, ny, nline), dtype=np.int32)
    ctx = native.get_context(OPTIONS['device'])
    if verbose >= 1:
        print("OBS_INTEGRATE: Integrating the Stokes IQUV spectra:")
    for zz in range(nline):
        spec = np.ascontiguousarray(sobs_in[zz], dtype=np.float64).reshape(nx * ny, nw, 4)
        tot, s, b, iss, status = ctx.obs_integrate(spec, wlarr[:, zz], wcont[zz])
        if (status == 1).any():
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")      # lr0[0,0] of an empty window, :1254
        sobs_tot[:, :, 4 * zz:4 * (zz + 1)] = tot.reshape(nx, ny, 4)
        snr[:, :, 4 * zz:4 * (zz + 1)] = s.reshape(nx, ny, 4)
        background[:, :, 4 * zz:4 * (zz + 1)] = b.reshape(nx, ny, 4)
        issuemask_obsint[:, :, zz] = iss.reshape(nx, ny)
    return sobs_tot, snr, background, issuemask_obsint


def obs_dens(sobs_totrot, yobs, keyvals, params):
    """Line-of-sight plasma density from the Fe XIII 1074/1079 ratio and the CHIANTI look-up table; CLEDB_PREPINV.py:1371-1479.

    Returns ``(log10 density rounded to 3 decimals [nx,ny] float32, iss_dens[nx,ny] int32)`` (issue 4: no table height above
    the pixel).  The reference builds one quadratic spline per PIXEL; here the ``nh`` splines of the table are built once with
    the same scipy call and every pixel is evaluated by one kernel (cledb_obs_dens).
    """
    nx, ny = keyvals[0:2]
    dobs = np.empty((nx, ny), dtype=np.float32)
    iss_dens = np.zeros((nx, ny), dtype=np.int32)
    if params.lookuptb[-4:] != ".npz":
        if params.verbose >= 1:
            print("OBS_DENS: FATAL! CHIANTI look-up table not found. Is the path correct?")
        return dobs, iss_dens
    import scipy.interpolate
    table = dict(np.load(params.lookuptb))
    splines = [scipy.interpolate.interp1d(table['rat'][i, :].flatten(), table['den'], kind="quadratic", fill_value="extrapolate")._spline
               for i in range(table['h'].shape[0])]                                       # :1461, once per height
    knots = np.stack([np.asarray(sp.t, dtype=np.float64) for sp in splines])
    coefs = np.stack([np.asarray(sp.c, dtype=np.float64).reshape(-1) for sp in splines])
    if params.verbose >= 1:
        print("OBS_DENS: Calculating observation LOS plasma density:")
    s = np.asarray(sobs_totrot)
    dens, iss = native.get_context(OPTIONS['device']).obs_dens(s[:, :, 0], s[:, :, 4], yobs, table['h'], knots, coefs)
    dobs[...] = dens.reshape(nx, ny)
    iss_dens[...] = iss.reshape(nx, ny)
    dobs[dobs <= 0] = 1
    with np.errstate(all='ignore'):
        return np.round(np.log10(dobs), 3), iss_dens


# ---------------------------------------------------------------------------------------------------------
# database side: which file for which pixel, load the distinct ones, make them resident
# ---------------------------------------------------------------------------------------------------------
_ION_DIRS = ("fe-xiii_1074/", "fe-xiii_1079/", "si-x_1430/", "si-ix_3934/", "mg-viii_3028/")
_INGEST_ERROR = ([None, None], None, 'Ingest Error')


def sdb_fileingest(dbdir, nline, tline, verbose):
    """File names and the (index, 100*height, 100*log density) table of a database directory; CLEDB_PREPINV.py:380-451.

    Returns ``([names_line1, names_line2], dbynumbers[nfiles,3] float32, [subdir1, subdir2])`` or the ingest-error triple
    ``([None, None], None, 'Ingest Error')``.  File order is Python's lexicographic ``sorted`` of the glob, exactly as in
    the reference (so ``d1000`` sorts before ``d600``); heights are characters 4..6 and densities characters 9..-4 after
    ``DB_h``.  Which sub-directories are taken follows the reference's own pairing rule (:402-405).
    """
    present = [os.path.isdir(dbdir + d) for d in _ION_DIRS]
    npresent = int(np.sum(present))
    if nline == 2:
        if npresent >= 2:
            where = np.where(present)[0]
            subdirs = [_ION_DIRS[where[i]] for i in range(npresent) if _ION_DIRS[i][:-1] in (tline[0], tline[1])]
            names = [sorted(glob.glob(dbdir + subdirs[0] + "DB_*")), sorted(glob.glob(dbdir + subdirs[1] + "DB_*"))]
            at = names[0][0].find('DB_h')
            table = np.empty((len(names[0]), 3), dtype=np.float32)
            for i, name in enumerate(names[0]):
                table[i] = (i, np.float32(name[at + 4:at + 7]), np.float32(name[at + 9:-4]))
            return names, table, subdirs
        if verbose >= 2 and verbose < 4:
            print("SDB_FILEINGEST: FATAL! Two line observation provided! Requires two individual ion databases in directory. "
                  "Only one database is computed." if npresent == 1 else
                  "SDB_FILEINGEST: FATAL! No database or incomplete calculations found in directory ")
        return _INGEST_ERROR
    # one-line database ingestion is unfinished upstream (CLEDB_PREPINV.py:434-446 cannot run); report it as an ingest error
    if verbose >= 1 and verbose < 4:
        print("SDB_FILEINGEST: one-line observations are solved analytically (blos_proc); no database is loaded.")
    return _INGEST_ERROR


def _select_files(yobs, dobs, dbynumbers):
    """Nearest database file of every pixel (cledb_db_select); raises what the reference raises for impossible pixels."""
    sel = native.get_context(OPTIONS['device']).db_select(yobs, dobs, dbynumbers)
    if (sel == -1).any():
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")      # NaN height: CLEDB_PREPINV.py:465-466
    if (sel == -2).any():
        raise ValueError("attempt to get argmin of an empty sequence")           # one file at that height: the empty slice of :466
    return sel


def sdb_findelongationanddens(xx, yy, y, d, dbynumbers):
    """Database file index for one pixel; CLEDB_PREPINV.py:458-468.  Returns ``(xx, yy, index)``."""
    return xx, yy, np.int32(_select_files(np.float32(y), np.float32(d), dbynumbers)[0])


def sdb_read(fildb, dbhdr, verbose):
    """One database file as the generator wrote it; CLEDB_PREPINV.py:514-553 (PyCELP ``.npy`` or raw CLE float32)."""
    dbcgrid, ned, ngx, nbphi, nbtheta, xed, gxmin, gxmax, bphimin, bphimax, bthetamin, bthetamax, nline, wavel, dbtype = dbhdr
    if verbose >= 2 and verbose < 4:
        print("INDIVIDUAL DB file location:", fildb)
    if dbtype == 0:
        return np.reshape(np.fromfile(fildb, dtype=np.float32), (ned, ngx, nbphi, nbtheta, nline * 4))
    return np.load(fildb)


def _two_line_database(file1, file2, dbhdr, wl0, verbose):
    """Both lines side by side, every component divided by I1 (component 0 last, so that it ends up exactly 1); PyCELP
    Stokes V additionally scaled by the nm/Angstrom factor; CLEDB_PREPINV.py:306-317."""
    db = np.append(sdb_read(file1, dbhdr, verbose), sdb_read(file2, dbhdr, verbose), axis=-1)
    with np.errstate(all='ignore'):
        if dbhdr[-1] == 1:
            vscale = 10 ** np.rint(np.log10(wl0 / 10000))
            for comp in range(7, -1, -1):
                if comp in (3, 7):
                    db[..., comp] = vscale * db[..., comp] / db[..., 0]
                else:
                    db[..., comp] = db[..., comp] / db[..., 0]
        else:
            for comp in range(7, -1, -1):
                db[..., comp] = db[..., comp] / db[..., 0]
    return db


def sdb_preprocess(yobs, dobs, keyvals, wlarr, params):
    """Find, load and make resident every database the observation needs; CLEDB_PREPINV.py:201-367.

    Returns ``(db_enc[nx,ny] int32, database (list of float32 arrays), dbhdr)``; with ``params.verbose == 4`` (the
    reference's unit-test switch, which also redirects to ``./tests/dbfiles/``) additionally ``dbnames, db_enc_flnm,
    db_uniq``; ``(None, None, None)`` when no database directory can be ingested.
    """
    chatty = params.verbose >= 1 and params.verbose < 4
    if chatty:
        print('------------------------------------\n----SDB_PREPROCESS - READ START-----\n------------------------------------')
    nx, ny, nline, tline = keyvals[0], keyvals[1], keyvals[3], keyvals[4]
    dbdir = params.dbdir if params.verbose != 4 else "./tests/dbfiles/"
    dbnames, dbynumbers, dbsubdirs = sdb_fileingest(dbdir, nline, tline, params.verbose)
    if dbsubdirs == 'Ingest Error':
        if chatty:
            print("SDB_PREPROCESS: FATAL! Could not read database.")
        return None, None, None

    # one kernel over the map instead of one pool task per pixel
    db_enc_flnm = _select_files(np.asarray(yobs, dtype=np.float32).reshape(-1), np.asarray(dobs, dtype=np.float32).reshape(-1),
                                dbynumbers).reshape(nx, ny).astype(np.int32)
    db_uniq = np.unique(db_enc_flnm)
    if chatty:
        print("Available CLEDB databases cover a span of", len(np.unique(dbynumbers[:, 1])), "solar heights between",
              dbynumbers[0, 1] / 100, "-", dbynumbers[-1, 1] / 100, " radius")
        print("Load ", db_uniq.shape[0], " heights x densities  DB datafiles in memory for each of ", nline,
              "line(s).\n------------------------------------")
    dbhdr = sdb_parseheader(dbdir + dbsubdirs[0] + 'db.hdr')
    if nline != 2:
        raise NotImplementedError("cledb_b200: only two-line databases are searched (one-line data is solved by blos_proc)")
    wl0 = wlarr[0, 0]
    with ThreadPoolExecutor(max_workers=max(1, min(8, db_uniq.shape[0]))) as pool:      # file reads overlap
        database = list(pool.map(lambda u: _two_line_database(dbnames[0][u], dbnames[1][u], dbhdr, wl0, params.verbose), db_uniq))
    db_enc = np.searchsorted(db_uniq, db_enc_flnm).astype(np.int32)        # position of a pixel's file in the loaded list

    # make the distinct databases resident now: cledb_invproc recognises the arrays and does not upload them again
    import CLEDB_PROC.CLEDB_PROC as procinv
    procinv.make_resident(database)
    if chatty:
        print('------------------------------------\n--SDB_PREPROCESS - READ FINALIZED---\n------------------------------------')
    if params.verbose == 4:
        return db_enc, database, dbhdr, dbnames, db_enc_flnm, db_uniq
    return db_enc, database, dbhdr


def __getattr__(name):
    """Names outside the accelerated path (e.g. sobs_preprocess) come from the reference's own module, see cledb_b200/passthrough.py."""
    if name.startswith('__'):
        raise AttributeError(name)
    from cledb_b200.passthrough import reference_attr
    return reference_attr('CLEDB_PREPINV/CLEDB_PREPINV.py', name, globals())
