"""B200-native drop-in for the hot-path neighbours in the reference's ``CLEDB_PREPINV/CLEDB_PREPINV.py``.

    obs_qurotate(sobs_tot, keyvals, hpxygrid) -> (sobs_totrot, aobs)      CLEDB_PREPINV.py:1277-1350
    obs_calcheight(keyvals, hpxygrid)          -> yobs                    CLEDB_PREPINV.py:897-926
    sdb_parseheader(dbheaderfile)              -> dbhdr 15-tuple          CLEDB_PREPINV.py:475-508
    iss_block(x)                                                          CLEDB_PREPINV.py:1563-1576

The rotation and the height map are one fused elementwise kernel (72 B/pixel) behind the C-ABI.  Instrument
header ingestion, spectral integration and the density look-up of ``sobs_preprocess`` are outside the scope of
this build.
"""
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
