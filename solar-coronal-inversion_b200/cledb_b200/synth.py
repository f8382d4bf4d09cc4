"""Synthetic PyCELP-shaped databases and observation maps.

The 33 GB Fe XIII database of the reference is a download and is not available offline, so every
benchmark / parity case runs on data of the same *shape* generated here with fixed seeds:

* a database is ``[ngx=21, nbphi=180, nbtheta=90, 8]`` float32, already in the form that the
  reference's ``sdb_preprocess`` hands to ``cledb_invproc`` (CLEDB_PREPINV.py:306-317): every
  component divided by I1, the two Stokes V components additionally scaled by 0.1;
* an observation map is built from hidden database rows times a field strength plus noise, with the
  header of the reference's own unit test (tests/test_functions.py:39-41) scaled to the map size.

This module only *creates inputs*.  It is used by bench.py, the tests and the golden-vector script;
it contains no part of the inversion itself.
"""
import math
import numpy as np

NGX, NBPHI, NBTHETA = 21, 180, 90
N_DB = NGX * NBPHI * NBTHETA            # 340 200 entries

# the test header of the reference: tests/dbfiles/fe-xiii_1074/db.hdr
HDR_VALUES = np.array([99, 121, 21, 180, 90, 6.0, 12.0, -0.5, 0.5, 0.0, 6.248279, 0.0, 3.106686,
                       1, 10746.15342348, 1], dtype=np.float32)


def _gauss(rng, shape):
    """Unit-variance noise from four uniforms.  Only + - * are used so that the stream is bit-reproducible
    on any CPU (ziggurat normals call libm, whose last bit depends on the CPU dispatch)."""
    u = rng.random(shape) + rng.random(shape) + rng.random(shape) + rng.random(shape)
    return (u - 2.0) * 1.7320508075688772


def _table(fn, values):
    """1-D table through Python's libm wrapper (same image => same glibc => same bits), not NumPy SIMD."""
    return np.array([fn(float(v)) for v in values], dtype=np.float64)


def make_dbhdr(ngx=NGX, nbphi=NBPHI, nbtheta=NBTHETA, dbtype=1):
    """The 15-tuple that CLEDB_PREPINV.sdb_parseheader returns (CLEDB_PREPINV.py:482-508)."""
    g = HDR_VALUES.copy()
    g[2], g[3], g[4] = ngx, nbphi, nbtheta
    g[-1] = dbtype
    ned = np.int32(g[1])
    emin, emax = np.float32(g[5]), np.float32(g[6])
    ff = 10. ** ((emax - emin) / (ned - 1))
    xed = np.empty(ned, np.float32)
    xed[0] = 10. ** emin
    for k in range(1, ned):
        xed[k] = xed[k - 1] * ff
    nline = np.int32(g[13])
    wavel = np.empty(nline, dtype=np.float32)
    for k in range(nline):
        wavel[k] = g[14 + k]
    return (g, ned, np.int32(g[2]), np.int32(g[3]), np.int32(g[4]), xed,
            np.float32(g[7]), np.float32(g[8]), np.float32(g[9]), np.float32(g[10]),
            np.float32(g[11]), np.float32(g[12]), nline, wavel, np.int32(g[-1]))


def make_database(h_index=9, d_index=40, ngx=NGX, nbphi=NBPHI, nbtheta=NBTHETA):
    """One normalised 2-line database ``[ngx,nbphi,nbtheta,8]`` float32 for grid point (height, density).

    Component layout as CLEDB_PREPINV.py:306-317 leaves it: [1, Q1, U1, 0.1*V1, I2, Q2, U2, 0.1*V2]/I1.
    ``h_index`` in 0..98 (1.01..2.00 R_sun), ``d_index`` in 0..120 (log ne 6.00..12.00).
    """
    rng = np.random.default_rng(20260101 + 1000 * int(h_index) + int(d_index))
    gx = (np.arange(ngx) / max(ngx - 1, 1) - 0.5)[:, None, None]
    phis = np.arange(nbphi) * (6.248279 / max(nbphi - 1, 1))
    ths = np.arange(nbtheta) * (3.106686 / max(nbtheta - 1, 1))
    cos2phi = _table(math.cos, 2 * phis)[None, :, None]
    sin2phi = _table(math.sin, 2 * phis)[None, :, None]
    cosphi = _table(math.cos, phis)[None, :, None]
    sinth = _table(math.sin, ths)[None, None, :]
    costh = _table(math.cos, ths)[None, None, :]
    hh = 1.01 + 0.01 * h_index
    dd = 6.0 + 0.05 * d_index
    p1 = 0.01 + 0.05 * (0.5 + 0.5 * math.cos(0.7 * hh + 0.3 * dd))        # in [0.01, 0.06]
    fgx = 1.0 / (1.0 + 2.0 * gx * gx)
    s2 = sinth * sinth
    db = np.empty((ngx, nbphi, nbtheta, 8), dtype=np.float64)
    db[..., 0] = 1.0
    db[..., 1] = p1 * s2 * cos2phi * fgx
    db[..., 2] = p1 * s2 * sin2phi * fgx
    i2 = 0.2 + 0.4 * (0.5 + 0.5 * math.tanh(0.8 * (dd - 8.5))) * (0.75 + 0.25 * costh * costh) * (1 - 0.3 * gx * gx)
    db[..., 4] = i2 + 0.0 * cosphi
    p2 = 0.1 * p1
    db[..., 5] = p2 * s2 * cos2phi * fgx * i2
    db[..., 6] = p2 * s2 * sin2phi * fgx * i2
    v = 1e-5 * 10 ** (0.5 + 0.5 * math.sin(1.3 * hh + 0.9 * dd))          # in [1e-5, 1e-4]
    ggx = 1.0 / (1.0 + gx * gx)
    db[..., 3] = 0.1 * v * costh * cosphi * ggx
    db[..., 7] = 0.35 * db[..., 3]
    noise = 1.0 + 1e-3 * _gauss(rng, db.shape)
    noise[..., 0] = 1.0
    db *= noise
    return np.ascontiguousarray(db.astype(np.float32))


def make_raw_line_pair(h_index, d_index, rng, **shape):
    """Un-normalised per-line IQUV arrays ``[ngx,nbphi,nbtheta,4]`` as the PyCELP generator stores them on disk
    (CLEDB_BUILD/rundb_1line_with_PyCELP.py:129-169): absolute intensities, Stokes V per Angstrom.  Loading them
    through ``sdb_preprocess`` (CLEDB_PREPINV.py:306-317) gives back ``make_database`` up to float32 rounding."""
    db = make_database(h_index, d_index, **shape).astype(np.float64)
    i1 = rng.uniform(1e-9, 3e-9, db.shape[:3])
    l1 = db[..., 0:4] * i1[..., None]
    l2 = db[..., 4:8] * i1[..., None]
    l1[..., 3] *= 10.0
    l2[..., 3] *= 10.0
    return l1.astype(np.float32), l2.astype(np.float32)


def make_keyvals(nx, ny, cdelt=None, nline=2):
    """Observation keyword list in the slot layout of CLEDB_PREPINV.obs_headstructproc (tests/test_functions.py:68-83)."""
    if cdelt is None:
        cdelt = 1e-4 * (1024.0 / max(nx, ny))
    tline = ['fe-xiii_1074', 'fe-xiii_1079'] if nline == 2 else ['fe-xiii_1074', 0]
    return [nx, ny, 1, nline, tline, 0, 0, [0, 0], -0.30, 1.05, [1074.2571, 1079.4204],
            cdelt, cdelt, [np.float32(0.00538654), np.float32(0.00541243)], 0, 0]


def geometry(nx, ny, keyvals):
    """Reference-exact ``aobs``/``yobs`` of the rectangular-grid branch (CLEDB_PREPINV.py:1300-1311, 917-920).

    float64 per-pixel scalar arithmetic rounded into float32 arrays, exactly as the reference's loops do.
    """
    crval1, crval2 = keyvals[8:10]
    cdelt1, cdelt2 = keyvals[11:13]
    linpolref = keyvals[14]
    xl = [crval1 + xx * cdelt1 for xx in range(nx)]
    yl = [crval2 + yy * cdelt2 for yy in range(ny)]
    xs = np.array(xl, dtype=np.float64)[:, None]
    ys = np.array(yl, dtype=np.float64)[None, :]
    yobs = np.sqrt(xs ** 2 + ys ** 2).astype(np.float32)
    a = np.array([[-math.atan2(x, y) for y in yl] for x in xl], dtype=np.float64).astype(np.float32)
    neg = a < 0
    a[neg] = (2 * np.pi + a[neg]).astype(np.float32)
    low = a < linpolref
    a = np.where(low, (2 * np.pi - a).astype(np.float32), (a + linpolref).astype(np.float32)).astype(np.float32)
    return a, yobs


SNR_TRUE = np.array([300, 5, 5, 0.5, 300, 1, 1, 0.3], dtype=np.float64)


def make_observation(nx, ny, databases, db_enc=None, seed=7, frac_nan=0.01, frac_zero=0.01, dobs_value=8.0):
    """Synthetic 2-line observation map drawn from hidden rows of ``databases`` (list of [..,8] arrays).

    Returns a dict with the per-pixel inputs of ``cledb_invproc`` (CLEDB_PROC.py:93):
    sobs_totrot, sobs_dopp, snr, yobs, aobs, dobs, db_enc, keyvals, issuemask and the hidden truth.
    """
    rng = np.random.default_rng(seed)
    npix = nx * ny
    if db_enc is None:
        db_enc = np.zeros((nx, ny), dtype=np.int32)
    enc = db_enc.reshape(-1)
    sobs = np.empty((npix, 8), dtype=np.float64)
    bpos = np.empty(npix, dtype=np.float64)
    hidden = np.empty(npix, dtype=np.int64)
    bmag = rng.uniform(1.0, 30.0, npix)
    nf = rng.uniform(1e-9, 3e-9, npix)
    hdr = make_dbhdr(*databases[0].shape[:3])
    for k, db in enumerate(databases):
        sel = np.flatnonzero(enc == k)
        if sel.size == 0:
            continue
        flat = db.reshape(-1, 8)
        rows = rng.integers(0, flat.shape[0], sel.size)
        hidden[sel] = rows
        s = flat[rows].astype(np.float64)
        s[:, 3] *= bmag[sel]
        s[:, 7] *= bmag[sel]
        sobs[sel] = s * nf[sel, None]
        # plane-of-sky field of the hidden row for the IQUD (Doppler) mode, CLEDB_PROC.py:1713
        nbth = db.shape[2]
        nbph = db.shape[1]
        ll = rows % nbth
        kk = (rows // nbth) % nbph
        pht = float(hdr[8]) + np.arange(nbph) * (float(hdr[9]) - float(hdr[8])) / max(nbph - 1, 1)
        tht = float(hdr[10]) + np.arange(nbth) * (float(hdr[11]) - float(hdr[10])) / max(nbth - 1, 1)
        sp, st, ct = _table(math.sin, pht)[kk], _table(math.sin, tht)[ll], _table(math.cos, tht)[ll]
        bpos[sel] = bmag[sel] * np.sqrt(st * st * sp * sp + ct * ct)
    sobs = sobs + np.abs(sobs) / SNR_TRUE[None, :] * _gauss(rng, sobs.shape)
    snr = (SNR_TRUE[None, :] * rng.uniform(0.9, 1.1, (npix, 8))).astype(np.float32)
    sobs = sobs.astype(np.float32)
    u = rng.random(npix)
    sobs[u < frac_nan] = np.nan
    sobs[(u >= frac_nan) & (u < frac_nan + frac_zero)] = 0.0
    keyvals = make_keyvals(nx, ny)
    aobs, yobs = geometry(nx, ny, keyvals)
    dopp = np.zeros((npix, 2), dtype=np.float32)
    dopp[:, 0] = bpos
    return dict(sobs_totrot=sobs.reshape(nx, ny, 8), sobs_dopp=dopp.reshape(nx, ny, 2),
                snr=snr.reshape(nx, ny, 8), yobs=yobs, aobs=aobs,
                dobs=np.full((nx, ny), dobs_value, dtype=np.float32), db_enc=db_enc.astype(np.int32),
                keyvals=keyvals, issuemask=np.zeros((nx, ny, 2), dtype=np.int32),
                hidden_row=hidden.reshape(nx, ny), hidden_b=bmag.reshape(nx, ny))


def make_oneline_map(nx, ny, seed=11, frac_bad=0.005):
    """Config-4 input: 1-line IQUV map ``[nx,ny,4]`` for the analytic B_LOS kernel."""
    rng = np.random.default_rng(seed)
    i = rng.uniform(1e-9, 3e-9, (nx, ny))
    s = np.empty((nx, ny, 4), dtype=np.float64)
    s[..., 0] = i
    s[..., 1] = i * 0.02 * _gauss(rng, (nx, ny))
    s[..., 2] = i * 0.02 * _gauss(rng, (nx, ny))
    s[..., 3] = i * 1e-4 * _gauss(rng, (nx, ny))
    s = s.astype(np.float32)
    u = rng.random((nx, ny))
    s[u < frac_bad / 2] = np.nan
    s[(u >= frac_bad / 2) & (u < frac_bad)] = 0.0
    return s


def make_twoline_raw(nx, ny, seed=11):
    """Config-4 input of the QU-rotation kernel: un-rotated 2-line map ``[nx,ny,8]`` float32."""
    a = make_oneline_map(nx, ny, seed=seed, frac_bad=0.0)
    b = make_oneline_map(nx, ny, seed=seed + 1, frac_bad=0.0) * np.float32(0.35)
    return np.ascontiguousarray(np.concatenate([a, b], axis=-1))
