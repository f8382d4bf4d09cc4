"""B200-native drop-in for the hot path of the reference's ``CLEDB_PROC/CLEDB_PROC.py``.

Same function names, positional arguments, return arity and error conventions as the reference
(arparaschiv/solar-coronal-inversion, CLEDB 0.7.0-beta):

    cledb_invproc(...)      CLEDB_PROC.py:93-363    2-line inversion of a map
    cledb_matchiquv(...)    CLEDB_PROC.py:921-1087  one pixel, full Stokes
    cledb_matchiqud(...)    CLEDB_PROC.py:1094-1252 one pixel, IQU + Doppler
    blos_proc(...)          CLEDB_PROC.py:388-486   1-line analytic B_LOS

The chi^2 search over the database runs in hand-written sm_100a kernels through the C-ABI of
``libcledb_b200.so``; there is no CPU fallback (a missing library or GPU raises).  ``ncpu`` and the Numba flags
are accepted and ignored.  Everything after the search (thresholds, field strength, geometry, issue codes) is
built on the device too (``cledb_invert``: solutions and the tests behind the issue codes); ``cledb_b200.solution`` is the NumPy
statement of the same steps, used for CLE-type headers and as a cross-check.  Result arrays live in pinned host memory
(``cledb_b200.native.PinnedPool``); they are the caller's like any fresh NumPy array.

Module-level options (not part of the reference surface; the defaults reproduce the reference):
    OPTIONS['dbencoding']  0 = database kept as float32 in HBM (lossless, default here), 1 = int16 codes
    OPTIONS['precision']   0 = fp32 kernel, 1 = fp64 validation kernel (reference operation order)
    OPTIONS['strict_topk'] False = reproduce the cledb_partsort index quirk, True = true top-nsearch
    OPTIONS['solution']    'device' = solutions built on the GPU right after the search (cledb_invert; PyCELP-type databases),
                           'host' = NumPy on the raw search result (cledb_b200.solution); both give the same bits
    OPTIONS['pruning']     True / 1 (default) = exact tile pruning in the search (k-d tile layout, bit-identical results, several
                           times faster), False / 0 = brute-force evaluation of every (pixel, entry) pair on the plain layout,
                           2 = brute force on the k-d layout with every pixel started from its seed tile
    OPTIONS['device']      CUDA ordinal (default: $CLEDB_B200_DEVICE, else $LOCAL_RANK, else 0)
    OPTIONS['sharded']     True = cledb_invproc is a collective over the ranks of cledb_b200.dist.init_comm(): pixels sharded,
                           the finished maps gathered over NVLink inside the library; OPTIONS['shard_root'] = rank that receives the maps
                           (-1, default: every rank)
"""
import os

import numpy as np

from cledb_b200 import native, solution

OPTIONS = dict(reduced_tie_order=None,      # theta order of an all-equal presort row; None = this machine's np.argsort
               dbencoding=int(os.environ.get('CLEDB_B200_DBENCODING', '0')),
               precision=int(os.environ.get('CLEDB_B200_PRECISION', '0')),
               strict_topk=False,
               solution=os.environ.get('CLEDB_B200_SOLUTION', 'device'),
               pruning=os.environ.get('CLEDB_B200_PRUNING', '1') != '0',
               device=int(os.environ.get('CLEDB_B200_DEVICE', os.environ.get('LOCAL_RANK', '0'))))


# ---------------------------------------------------------------------------------------------------------
# database residency: upload each distinct database array once and keep it in HBM across calls
# ---------------------------------------------------------------------------------------------------------
from cledb_b200.residency import Resident

OPTIONS.setdefault('verify_resident', os.environ.get('CLEDB_B200_VERIFY_RESIDENT', 'sample'))   # 'sample' | 'full'
OPTIONS.setdefault('hbm_budget_bytes', int(float(os.environ.get('CLEDB_B200_HBM_BUDGET', 120e9))))

_resident = {}


def _context():
    return native.get_context(OPTIONS['device'])


def release_databases():
    """Drop every database this module has made resident (frees HBM)."""
    dev = OPTIONS['device']
    if dev in _resident:
        _resident[dev].clear(_context())


def _upload(database, used):
    ctx = _context()
    res = _resident.get(OPTIONS['device'])
    if res is None:
        res = _resident[OPTIONS['device']] = Resident()
    res.verify, res.budget_bytes = OPTIONS['verify_resident'], OPTIONS['hbm_budget_bytes']
    want = int(OPTIONS['pruning'])          # 0 plain layout / brute force, 1 exact tile pruning, 2 seeded brute force (k-d layout)
    if res.entries and (res.encoding != OPTIONS['dbencoding'] or res.pruning != (want != 0)):
        res.clear(ctx)                      # one encoding and one tile layout at a time live on the device
    ctx.set_pruning(want)
    res.pruning, res.encoding = want != 0, OPTIONS['dbencoding']
    return ctx, res.ensure_all(ctx, {int(k): database[int(k)] for k in used}, OPTIONS['dbencoding'])


def make_resident(database):
    """Upload every array of ``database`` to HBM now (used by ``sdb_preprocess``); a later ``cledb_invproc`` with the same
    arrays finds them resident."""
    _upload(database, range(len(database)))


def _used_databases(enc, ndb):
    """Distinct values of ``db_enc`` (a bincount: cheaper than a sort for the small integer range)."""
    if enc.size == 0:
        return np.zeros(0, np.int64)
    if np.issubdtype(enc.dtype, np.integer) and enc.min() >= 0 and enc.max() < max(ndb, 1) + 4096:
        return np.flatnonzero(np.bincount(enc, minlength=1))
    return np.unique(enc)


def _search_and_build(mode, sobs, sobs_dopp, database, db_enc, yobs, aobs, dobs, snr, dbhdr, nsearch, maxchisq, bcalc,
                      reduced=False):
    """Flattened pixels in, (invout[npix,k,11], sfound[npix,k,8], issue codes [npix] or None) out."""
    sobs = np.ascontiguousarray(sobs, dtype=np.float32).reshape(-1, 8)
    snr = np.ascontiguousarray(snr, dtype=np.float32).reshape(-1, 8)
    enc = np.asarray(db_enc).reshape(-1)
    if OPTIONS.get('sharded') and getattr(_context(), 'nranks', 1) > 1:
        return _search_and_build_sharded(_context(), mode, sobs, sobs_dopp, database, enc, yobs, aobs, dobs, snr, dbhdr, nsearch,
                                         maxchisq, bcalc, reduced)
    if enc.size and (int(enc.max()) >= len(database) or int(enc.min()) < 0):
        raise IndexError('db_enc refers to database %d, the list holds %d' % (int(enc.max()), len(database)))   # as the reference's database[db_enc[xx,yy]]
    used = np.zeros(1, np.int64) if len(database) == 1 else _used_databases(enc, len(database))
    ctx, ids = _upload(database, used)
    if len(ids) == 1 and list(ids.values())[0] == 0 and not enc.any() and enc.dtype == np.int32:
        dbid = np.ascontiguousarray(enc)             # one database, id 0: db_enc itself is the id map
    else:
        lut = np.zeros(int(enc.max()) + 1, dtype=np.int32)
        for k, v in ids.items():
            lut[k] = v
        dbid = lut[enc]
    if reduced:                                      # cledb_getsubsetiquv / cledb_getsubsetiqud + search over the subset
        tie = OPTIONS['reduced_tie_order']
        if OPTIONS['solution'] == 'device' and int(dbhdr[14]) == 1:
            inv, sf, _ = ctx.invert_reduced(mode, sobs, snr, dbid, nsearch, yobs, dobs, aobs, sobs_dopp, dbhdr, maxchisq, bcalc,
                                            strict_topk=OPTIONS['strict_topk'], precision=OPTIONS['precision'], tie_order=tie)
            return inv, sf, native.issue_codes(ctx.last_issue, inv)
        raw = ctx.match_reduced(mode, sobs, snr, dbid, nsearch, sobs_dopp, dbhdr, precision=OPTIONS['precision'], tie_order=tie)
        return solution.build_solutions(raw, mode, sobs, sobs_dopp, np.asarray(yobs).reshape(-1), np.asarray(aobs).reshape(-1),
                                        np.asarray(dobs).reshape(-1), dbhdr, nsearch, maxchisq, bcalc,
                                        strict_topk=OPTIONS['strict_topk']) + (None,)
    if OPTIONS['solution'] == 'device' and int(dbhdr[14]) == 1:
        dopp0 = None if sobs_dopp is None else np.asarray(sobs_dopp).reshape(sobs.shape[0], -1)[:, 0]
        inv, sf, _ = ctx.invert(mode, sobs, snr, dbid, nsearch, yobs, dobs, aobs, dopp0, dbhdr, maxchisq, bcalc,
                                strict_topk=OPTIONS['strict_topk'], precision=OPTIONS['precision'])
        return inv, sf, native.issue_codes(ctx.last_issue, inv)
    raw = ctx.match(mode, sobs, snr, dbid, nsearch, precision=OPTIONS['precision'], want_rows=True,
                    want_chi64=OPTIONS['precision'] == native.PREC_FP64)
    return solution.build_solutions(raw, mode, sobs, sobs_dopp, np.asarray(yobs).reshape(-1), np.asarray(aobs).reshape(-1),
                                    np.asarray(dobs).reshape(-1), dbhdr, nsearch, maxchisq, bcalc,
                                    strict_topk=OPTIONS['strict_topk']) + (None,)


def _search_and_build_sharded(ctx, mode, sobs, sobs_dopp, database, enc, yobs, aobs, dobs, snr, dbhdr, nsearch, maxchisq, bcalc,
                              reduced):
    """The same over every rank of the context's communicator (cledb_b200.dist.init_comm; OPTIONS['sharded']): every rank is
    called with the full map, takes its share of the pixel partition - contiguous ranges of the database-sorted pixel list,
    balanced by database rows, so it uploads only the databases of its range -, and the library's gather over NVLink (its
    peer-memory push kernel, or ncclAllGather + scatter) returns the finished maps (cledb_invert_sharded).  Collective."""
    if reduced or OPTIONS['solution'] != 'device' or int(dbhdr[14]) != 1:
        raise native.CledbError('sharded inversion covers the full search with device-side solutions of PyCELP-type databases')
    npix = sobs.shape[0]
    enc = np.ascontiguousarray(enc, dtype=np.int32)
    weights = np.array([float(np.prod(np.shape(d)[:-1])) for d in database])
    order, bounds = native.partition(enc, weights, ctx.nranks)
    mine = order[bounds[ctx.rank]:bounds[ctx.rank + 1]]
    emine = enc[mine]
    _, ids = _upload(database, _used_databases(emine, len(database)))
    lut = np.zeros(len(database), dtype=np.int32)
    for k, v in ids.items():
        lut[k] = v
    flat = lambda a: np.asarray(a).reshape(-1)[mine]
    dopp0 = None if sobs_dopp is None else np.asarray(sobs_dopp).reshape(npix, -1)[mine, 0]
    inv, sf, _ = ctx.invert_sharded(mode, sobs[mine], snr[mine], lut[emine], nsearch, flat(yobs), flat(dobs), flat(aobs), dopp0, dbhdr,
                                    maxchisq, bcalc, npix, order, bounds, root=int(OPTIONS.get('shard_root', -1)),
                                    strict_topk=OPTIONS['strict_topk'], precision=OPTIONS['precision'])
    if inv is None:                                  # a rank that does not receive the maps (shard_root names another)
        return None, None, None
    return inv, sf, native.issue_codes(ctx.last_issue, inv)


# ---------------------------------------------------------------------------------------------------------
# the reference's call surface
# ---------------------------------------------------------------------------------------------------------
def cledb_invproc(sobs_totrot, sobs_dopp, database, db_enc, yobs, aobs, dobs, snr, issuemask, dbhdr, keyvals,
                  nsearch, maxchisq, bcalc, iqud, ncpu, reduced, verbose):
    """2-line inversion of a preprocessed map; reference CLEDB_PROC.py:93-363.

    Returns ``(invout[nx,ny,nsearch,11], sfound[nx,ny,nsearch,8])`` float32 and updates ``issuemask`` in place
    with codes 512 / 1024 / 2048.  Fatal input problems do not raise: they return ``invout[...,0] == -1``.
    """
    if verbose >= 1:
        print('--------------------------------------\n----CLEDB_INVPROC - INVERSION START---\n--------------------------------------')
    nx, ny = keyvals[0:2]

    def aborted():
        invout = np.zeros((nx, ny, nsearch, 11), dtype=np.float32)
        invout[:, :, :, 0] = -1
        return invout, np.zeros((nx, ny, nsearch, 8), dtype=np.float32)

    if np.shape(sobs_totrot)[-1] != 8 or database[0].shape[-1] != 8:                      # CLEDB_PROC.py:166-171
        if verbose >= 1:
            print('CLEDB_INVPROC: FATAL! Must have a two-line observation and a corresponding two-line database; Aborting!')
            print("Shapes are: ", np.shape(sobs_totrot), " and ", database[0].shape)
        return aborted()
    if (iqud == True and bcalc != 3) or (iqud != True and bcalc == 3):                    # CLEDB_PROC.py:174-178
        if verbose >= 1:
            print("CLEDB_INVPROC: FATAL! Field strength calculation incompatible with requested input data! Check bcalc and iqud keywords; Aborting!")
        return aborted()
    if verbose >= 2:
        print("CLEDB_INVPROC: Using a %s database search of size:" % ("reduced" if reduced == True else "full"),
              dbhdr[1] * dbhdr[2] * dbhdr[3] * dbhdr[4])

    mode = native.MODE_IQUD if iqud == True else native.MODE_IQUV
    dopp = None
    if mode == native.MODE_IQUD:
        dopp = np.asarray(sobs_dopp)
        dopp = dopp.reshape(nx * ny, -1)
    inv, sf, codes = _search_and_build(mode, sobs_totrot, dopp, database, db_enc, yobs, aobs, dobs, snr, dbhdr, nsearch, maxchisq,
                                       bcalc, reduced=reduced == True)
    if inv is None:                                            # sharded call on a rank other than OPTIONS['shard_root']
        return None, None
    invout = np.ascontiguousarray(inv, dtype=np.float32).reshape(nx, ny, nsearch, 11)
    sfound = np.ascontiguousarray(sf, dtype=np.float32).reshape(nx, ny, nsearch, 8)
    if codes is None:                                          # solutions built on the host (CLE-type header, OPTIONS['solution'])
        solution.update_issuemask(issuemask, invout, np.asarray(snr), nx, ny)
    else:                                                      # tests evaluated by k_build on the device (CLEDB_PROC.py:323-355)
        solution.apply_issue_codes(issuemask, codes.reshape(nx, ny))
    if verbose >= 1:
        print('--------------------------------------\n--CLEDB_INVPROC - INVERSION FINALIZED-\n--------------------------------------')
    return invout, sfound


def cledb_matchiquv(xx, yy, sobs_1pix, yobs_1pix, aobs_1pix, dobs_1pix, database_in, dbhdr, snr, nsearch, maxchisq,
                    bcalc, reduced, verbose):
    """One pixel, full Stokes IQUV; reference CLEDB_PROC.py:921-1087.  Returns ``(xx, yy, out, smatch)``."""
    inv, sf, _ = _search_and_build(native.MODE_IQUV, np.asarray(sobs_1pix).reshape(1, 8), None, [database_in], np.zeros(1, np.int32),
                                np.array([yobs_1pix]), np.array([aobs_1pix]), np.array([dobs_1pix]), np.asarray(snr).reshape(1, 8),
                                dbhdr, nsearch, maxchisq, bcalc, reduced=reduced == True)
    return xx, yy, inv[0], sf[0]


def cledb_matchiqud(xx, yy, sobs_1pix, sobsd_1pix, yobs_1pix, aobs_1pix, dobs_1pix, database_in, dbhdr, snr, nsearch,
                    maxchisq, bcalc, reduced, verbose):
    """One pixel, IQU + Doppler; reference CLEDB_PROC.py:1094-1252.  Returns ``(xx, yy, out, smatch)``."""
    inv, sf, _ = _search_and_build(native.MODE_IQUD, np.asarray(sobs_1pix).reshape(1, 8), np.asarray(sobsd_1pix).reshape(1, -1),
                                [database_in], np.zeros(1, np.int32), np.array([yobs_1pix]), np.array([aobs_1pix]),
                                np.array([dobs_1pix]), np.asarray(snr).reshape(1, 8), dbhdr, nsearch, maxchisq, bcalc,
                                reduced=reduced == True)
    return xx, yy, inv[0], sf[0]


def blos_proc(sobs_tot, snr, issuemask, keyvals, consts, params):
    """1-line analytic B_LOS for every line of the observation; reference CLEDB_PROC.py:388-486.

    Returns ``blosout[nx,ny,nline,4]`` float32 and adds issue codes 32 / 64 to ``issuemask``.  ``consts`` is the
    constants *module* (``consts.Constants(line)`` is instantiated per line).  Returns 0 for ``params.iqud``.
    """
    if params.iqud == True:
        print("BLOS_PROC: FATAL! No Stokes V data available to compute analytical line-of-sight projected magnetic fields.")
        return 0
    if params.verbose >= 1:
        print('--------------------------------------\n---BLOS_PROC: B LOS ESTIMATION START--\n--------------------------------------')
    nx, ny = keyvals[0:2]
    nline, tline = keyvals[3], keyvals[4]
    blosout = np.zeros((nx, ny, nline, 4), dtype=np.float32)
    ctx = _context()
    sobs_tot = np.asarray(sobs_tot)
    for zz in range(nline):
        const = consts.Constants(tline[zz])
        kfac = 1.e9 * (const.planckconst * const.l_speed / const.bohrmagneton / (const.line_ref ** 2))
        iquv = np.ascontiguousarray(sobs_tot[:, :, 4 * zz:4 * (zz + 1)], dtype=np.float32).reshape(-1, 4)
        out, flags = ctx.blos(iquv, kfac, const.g_eff, const.F_factor)
        blosout[:, :, zz, :] = out.reshape(nx, ny, 4)
        flags = flags.reshape(nx, ny, 2)
        m = issuemask[:, :, zz]
        for col, code in ((0, 32), (1, 64)):                                               # CLEDB_PROC.py:478-479
            add = (flags[:, :, col] != 0) & ((m.astype(np.int64) & code) == 0)
            m[add] += code
    if params.verbose >= 1:
        print('--------------------------------------\n-BLOS_PROC: BLOS ESTIMATION FINALIZED-\n--------------------------------------')
    return blosout


def iss_block(x):
    """Issue codes set in one issuemask value (list of powers of two, zeros for clear bits); CLEDB_PREPINV.py:1563-1576."""
    x = int(x)
    return [(1 << b) if (x >> b) & 1 else 0 for b in range(x.bit_length())] if x > 0 else []


def __getattr__(name):
    """Names outside the accelerated path (e.g. spectro_proc) come from the reference's own module, see cledb_b200/passthrough.py."""
    if name.startswith('__'):
        raise AttributeError(name)
    from cledb_b200.passthrough import reference_attr
    return reference_attr('CLEDB_PROC/CLEDB_PROC.py', name, globals())
