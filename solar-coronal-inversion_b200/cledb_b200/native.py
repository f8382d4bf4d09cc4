"""ctypes binding of libcledb_b200.so (include/cledb_b200.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device is present, creating a
:class:`Context` raises.  NumPy arrays are passed as C-contiguous host buffers; the ``*_dev`` methods take raw
device addresses (e.g. ``torch.Tensor.data_ptr()``) and a CUDA stream handle.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('CLEDB_B200_LIB') or os.path.join(os.path.dirname(_HERE), 'lib', 'libcledb_b200.so')

MODE_IQUV, MODE_IQUD = 0, 1
ENC_F32, ENC_I16 = 0, 1
PREC_FP32, PREC_FP64 = 0, 1
MAX_NSEARCH = 32
PIX_SEARCHED, PIX_NULL, PIX_ALLNAN, PIX_REFRAISES, PIX_ALLINF = 0, 1, 2, 3, 4
SELECT_NAN, SELECT_EMPTY = -1, -2


def lib_const(name):
    """Value of a ``#define`` of include/cledb_b200.h by name (the tests read the constants through here)."""
    return {'CLEDB_SELECT_NAN': SELECT_NAN, 'CLEDB_SELECT_EMPTY': SELECT_EMPTY}[name]

# every symbol include/cledb_b200.h declares, with (restype, argtypes)
_c = ctypes
_P = ctypes.c_void_p
SIGNATURES = {
    'cledb_version': (_c.c_char_p, []),
    'cledb_create': (_c.c_int, [_c.c_int, _c.POINTER(_P)]),
    'cledb_destroy': (_c.c_int, [_P]),
    'cledb_last_error': (_c.c_char_p, [_P]),
    'cledb_device_props': (_c.c_int, [_P, _P]),
    'cledb_synchronize': (_c.c_int, [_P]),
    'cledb_launch_count': (_c.c_int64, [_P]),
    'cledb_profile_begin': (_c.c_int, [_P]),
    'cledb_profile_collect': (_c.c_int, [_P, _P, _c.c_int]),
    'cledb_measure_fp32_peak': (_c.c_int, [_P, _P]),
    'cledb_db_put': (_c.c_int, [_P, _c.c_int, _P, _c.c_int64, _c.c_int]),
    'cledb_db_drop': (_c.c_int, [_P, _c.c_int]),
    'cledb_db_info': (_c.c_int, [_P, _c.c_int, _P]),
    'cledb_db_decoded': (_c.c_int, [_P, _c.c_int, _P]),
    'cledb_gather_rows': (_c.c_int, [_P, _c.c_int, _P, _c.c_int64, _P]),
    'cledb_db_select': (_c.c_int, [_P, _c.c_int64, _P, _P, _c.c_int, _P, _P]),
    'cledb_db_select_dev': (_c.c_int, [_P, _c.c_int64, _P, _P, _c.c_int, _P, _P, _P]),
    'cledb_obs_integrate': (_c.c_int, [_P, _c.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    'cledb_obs_dens': (_c.c_int, [_P, _c.c_int64, _P, _P, _P, _c.c_int, _P, _c.c_int, _P, _c.c_int, _P, _P, _P]),
    'cledb_match': (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int64, _P, _P, _P, _c.c_int, _P, _P, _P, _P, _P, _P, _P]),
    'cledb_match_dev': (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int64, _P, _P, _P, _c.c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    'cledb_set_tuning': (_c.c_int, [_P, _c.c_int, _c.c_int]),
    'cledb_set_pruning': (_c.c_int, [_P, _c.c_int]),
    'cledb_pruning_stats': (_c.c_int, [_P, _P]),
    'cledb_invert': (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int64, _P, _P, _P, _c.c_int, _P, _P, _P, _P, _P, _c.c_int,
                                _c.c_int64, _P, _c.c_double, _c.c_int, _c.c_int, _P, _P, _P, _P]),
    'cledb_match_reduced': (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int64, _P, _P, _P, _c.c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    'cledb_match_reduced_dev': (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int64, _P, _P, _P, _c.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'cledb_invert_reduced': (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int64, _P, _P, _P, _c.c_int, _P, _P, _P, _P, _P, _P, _c.c_int,
                                        _c.c_int64, _P, _c.c_double, _c.c_int, _c.c_int, _P, _P, _P, _P]),
    'cledb_invert_dev': (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int64, _P, _P, _P, _c.c_int, _P, _P, _P, _P, _P, _c.c_int,
                                    _c.c_int64, _P, _c.c_double, _c.c_int, _c.c_int, _P, _P, _P, _P, _P]),
    'cledb_invert_sharded': (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int64, _P, _P, _P, _c.c_int, _P, _P, _P, _P, _P, _c.c_int,
                                        _c.c_int64, _P, _c.c_double, _c.c_int, _c.c_int, _c.c_int64, _P, _P, _c.c_int, _P, _P, _P, _P]),
    'cledb_comm_unique_id': (_c.c_int, [_P]),
    'cledb_comm_init': (_c.c_int, [_P, _c.c_int, _c.c_int, _P]),
    'cledb_comm_destroy': (_c.c_int, [_P]),
    'cledb_comm_info': (_c.c_int, [_P, _P]),
    'cledb_comm_transport': (_c.c_int, [_P, _c.c_int, _P]),
    'cledb_comm_push_blocks': (_c.c_int, [_P, _c.c_int]),
    'cledb_comm_window': (_c.c_int, [_P, _c.c_int64, _c.POINTER(_P)]),
    'cledb_allgather_dev': (_c.c_int, [_P, _P, _P, _c.c_int64, _P]),
    'cledb_map_layout': (_c.c_int, [_c.c_int64, _c.c_int, _P]),
    'cledb_gather_maps_dev': (_c.c_int, [_P, _c.c_int64, _c.c_int64, _P, _c.c_int64, _c.c_int, _P, _P, _c.c_int, _P]),
    'cledb_invert_gather_dev': (_c.c_int, [_P, _c.c_int, _c.c_int, _c.c_int64, _P, _P, _P, _c.c_int, _P, _P, _P, _P, _P, _c.c_int, _P,
                                           _c.c_double, _c.c_int, _c.c_int, _c.c_int64, _P, _P, _c.c_int, _c.c_int, _P]),
    'cledb_partition': (_c.c_int, [_c.c_int64, _P, _c.c_int, _P, _c.c_int, _P, _P]),
    'cledb_shard_layout': (_c.c_int, [_c.c_int64, _c.c_int, _P]),
    'cledb_unshard_dev': (_c.c_int, [_P, _c.c_int, _P, _P, _c.c_int64, _c.c_int64, _c.c_int, _P, _P, _P, _P, _P, _P]),
    'cledb_host_alloc': (_c.c_int, [_c.c_int64, _c.POINTER(_P)]),
    'cledb_host_free': (_c.c_int, [_P]),
    'cledb_blos': (_c.c_int, [_P, _c.c_int64, _P, _c.c_double, _c.c_double, _c.c_double, _P, _P]),
    'cledb_blos_dev': (_c.c_int, [_P, _c.c_int64, _P, _c.c_double, _c.c_double, _c.c_double, _P, _P, _P]),
    'cledb_qurotate': (_c.c_int, [_P, _c.c_int, _c.c_int, _P, _P, _c.c_double, _P, _P, _P, _P, _P]),
    'cledb_qurotate_dev': (_c.c_int, [_P, _c.c_int, _c.c_int, _P, _P, _c.c_double, _P, _P, _P, _P, _P, _P]),
    'cledb_codec_shift': (_c.c_int, [_P, _c.c_int64, _c.c_int64]),
    'cledb_codec_encode': (None, [_P, _c.c_int64, _c.c_int, _P]),
    'cledb_codec_decode': (None, [_P, _c.c_int64, _c.c_int, _P]),
}

_lib = None


class CledbError(RuntimeError):
    pass


class DbGrid(ctypes.Structure):
    """``cledb_dbgrid`` of the header: the part of ``dbhdr`` (CLEDB_PREPINV.py:507-508) that solution building needs,
    plus sin/cos tables of the float32 grid angles evaluated here with NumPy (the reference's own bits)."""
    _fields_ = [('dbtype', _c.c_int32), ('ngx', _c.c_int32), ('nbphi', _c.c_int32), ('nbtheta', _c.c_int32),
                ('gxmin', _c.c_float), ('gxmax', _c.c_float), ('bphimin', _c.c_float), ('bphimax', _c.c_float),
                ('bthetamin', _c.c_float), ('bthetamax', _c.c_float),
                ('sin_bphi', _P), ('cos_bphi', _P), ('sin_btheta', _P), ('cos_btheta', _P)]

    @classmethod
    def from_dbhdr(cls, dbhdr):
        dbcgrid, ned, ngx, nbphi, nbtheta, xed, gxmin, gxmax, bphimin, bphimax, bthetamin, bthetamax, nline, wavel, dbtype = dbhdr
        f32 = np.float32

        def angles(lo, hi, n):      # cledb_physcle, CLEDB_PROC.py:1737-1739: float64 arithmetic on an int64 index, then float32
            k = np.arange(int(n), dtype=np.int64).astype(np.float64)
            with np.errstate(all='ignore'):
                return (np.float64(lo) + (k * np.float64(f32(hi) - f32(lo))) / np.float64(int(n) - 1)).astype(f32)

        g = cls()
        g.dbtype, g.ngx, g.nbphi, g.nbtheta = int(dbtype), int(ngx), int(nbphi), int(nbtheta)
        g.gxmin, g.gxmax, g.bphimin, g.bphimax = float(f32(gxmin)), float(f32(gxmax)), float(f32(bphimin)), float(f32(bphimax))
        g.bthetamin, g.bthetamax = float(f32(bthetamin)), float(f32(bthetamax))
        phi, th = angles(bphimin, bphimax, nbphi), angles(bthetamin, bthetamax, nbtheta)
        g._tables = [np.ascontiguousarray(t, dtype=f32) for t in (np.sin(phi), np.cos(phi), np.sin(th), np.cos(th))]
        g.sin_bphi, g.cos_bphi, g.sin_btheta, g.cos_btheta = (t.ctypes.data for t in g._tables)
        return g


def load_library(path=None):
    """Load the shared library (once) and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise CledbError('%s is missing: build it with `python __graft_entry__.py` (or make -C '
                         'solar-coronal-inversion_b200/csrc); there is no CPU fallback' % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _as(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# ---- int16 codec (host only) ---------------------------------------------------------------------------------
def codec_shift(x):
    x = _as(x, np.float32).reshape(-1)
    return int(load_library().cledb_codec_shift(_ptr(x), x.size, 1))


def codec_encode(x, shift):
    x = _as(x, np.float32)
    out = np.empty(x.shape, dtype=np.int16)
    load_library().cledb_codec_encode(_ptr(x), x.size, int(shift), _ptr(out))
    return out


def codec_decode(codes, shift):
    codes = _as(codes, np.int16)
    out = np.empty(codes.shape, dtype=np.float32)
    load_library().cledb_codec_decode(_ptr(codes), codes.size, int(shift), _ptr(out))
    return out


# ---- pinned host memory ---------------------------------------------------------------------------------------
class _PinnedBlock:
    """Owner of one page-locked allocation; NumPy arrays made from it keep it alive through their ``base`` chain."""

    def __init__(self, ptr, nbytes):
        self.__array_interface__ = dict(data=(ptr, False), shape=(nbytes,), typestr='|u1', version=3)


class PinnedPool:
    """Result arrays in page-locked host memory (cledb_host_alloc): device-to-host copies into them run at full PCIe speed.

    ``empty(shape, dtype)`` returns a NumPy array like ``np.empty``.  The caller owns it like any other fresh array; when
    the array and every view of it have been garbage collected the block goes back to the pool (up to ``keep_bytes`` are
    kept for reuse, the rest is freed), so a loop that drops its results reuses the same pinned blocks and a caller that
    keeps them gets new ones - results of an earlier call are never overwritten."""

    def __init__(self, keep_bytes=4 << 30):
        self.free = {}                 # size class -> [ptr]
        self.kept = 0
        self.keep_bytes = keep_bytes
        self.allocs = 0                # statistics (tests)

    @staticmethod
    def _size_class(nbytes):
        c = 1 << 16
        while c < nbytes:
            c <<= 1
        return c if c - nbytes <= c // 4 else (nbytes + (1 << 20) - 1) >> 20 << 20

    def _release(self, cls, ptr):
        if self.kept + cls <= self.keep_bytes:
            self.free.setdefault(cls, []).append(ptr)
            self.kept += cls
        else:
            try:
                load_library().cledb_host_free(ctypes.c_void_p(ptr))
            except Exception:
                pass

    def empty(self, shape, dtype):
        import weakref
        dtype = np.dtype(dtype)
        nbytes = max(1, int(np.prod(shape, dtype=np.int64)) * dtype.itemsize)
        cls = self._size_class(nbytes)
        lst = self.free.get(cls)
        if lst:
            ptr = lst.pop()
            self.kept -= cls
        else:
            h = ctypes.c_void_p()
            if load_library().cledb_host_alloc(cls, ctypes.byref(h)) != 0 or not h.value:
                return np.empty(shape, dtype)          # no pinned memory left: an ordinary array (slower copies, same result)
            ptr = h.value
            self.allocs += 1
        block = _PinnedBlock(ptr, cls)
        weakref.finalize(block, self._release, cls, ptr)
        return np.asarray(block)[:nbytes].view(dtype).reshape(shape)

    def trim(self):
        for cls, lst in self.free.items():
            for ptr in lst:
                load_library().cledb_host_free(ctypes.c_void_p(ptr))
        self.free.clear()
        self.kept = 0


pinned = PinnedPool()


def issue_codes(codes, invout=None):
    """Issue codes 512 / 1024 / 2048 of cledb_invproc (CLEDB_PROC.py:347-355) per pixel from the device's ``issue_out[npix]``
    (see cledb_invert in the header).  Pixels whose Van-Vleck angle lies within 2e-3 deg of a limit (bit 0) are re-evaluated
    with NumPy from ``invout[npix,k,11]`` - the reference's own float32 chain decides those."""
    from . import solution
    # marked pixels are about 1 in 25 000: find the rows of a 2-D view that hold one first (an OR-reduction per row), then
    # look inside those rows only - a fifth of the cost of a flat search on a 256 x 256 map
    n = codes.size
    w = 256 if n % 256 == 0 and n >= 4096 and codes.flags.c_contiguous else 0
    if w:
        c2 = codes.reshape(-1, w)
        rows = np.flatnonzero(np.bitwise_or.reduce(c2, axis=1) & 1)
        if rows.size:
            r, c = np.nonzero(c2[rows] & 1)
            amb = rows[r] * w + c
        else:
            amb = rows
    else:
        amb = np.flatnonzero(codes & 1)
    if amb.size:
        codes = codes & ~np.int32(1)
        if invout is not None:
            vv = solution.van_vleck_last(invout[amb])
            codes[amb] = (codes[amb] & ~np.int32(512)) | (vv.astype(np.int32) << 9)
    return codes


class Reduced(ctypes.Structure):
    """``cledb_reduced`` of the header: what the reduced-database presort (cledb_getsubsetiquv / cledb_getsubsetiqud,
    CLEDB_PROC.py:1401-1568) needs, with every transcendental evaluated here by NumPy in the reference's own dtypes."""
    _fields_ = [('outer', _c.c_int32), ('nbphi', _c.c_int32), ('nbtheta', _c.c_int32),
                ('tan_btheta', _P), ('sin_bphi', _P), ('tie_rank', _P), ('tan_phib', _P)]

    @classmethod
    def build(cls, mode, sobs, sobs_dopp, dbhdr, tie_order=None):
        dbcgrid, ned, ngx, nbphi, nbtheta, xed, gxmin, gxmax, bphimin, bphimax, bthetamin, bthetamax, nline, wavel, dbtype = dbhdr
        with np.errstate(all='ignore'):
            kk = np.arange(0, nbphi)
            bphir = bphimin + kk * (bphimax - bphimin) / (nbphi)                 # :1417-1418
            ll = np.arange(0, nbtheta)
            bthetar = bthetamin + ll * (bthetamax - bthetamin) / (nbtheta)       # :1420-1421
            if mode == MODE_IQUV:
                phib = 0.5 * np.arctan2(sobs[:, 2], sobs[:, 1])                  # :1428 (float32)
            else:
                phib = np.asarray(sobs_dopp)[:, 1]                               # :1516 (wave angle, dtype of the input)
            tans = np.stack([np.tan(phib), np.tan(np.pi - phib)], axis=1).astype(np.float64)
            tt, sp = np.tan(bthetar).astype(np.float64), np.sin(bphir).astype(np.float64)
        order = np.argsort(np.zeros(int(nbtheta))) if tie_order is None else np.asarray(tie_order)
        rank = np.empty(int(nbtheta), dtype=np.int32)
        rank[order] = np.arange(int(nbtheta), dtype=np.int32)
        r = cls()
        r.outer = int(ngx) if int(dbtype) == 1 else int(ned) * int(ngx)
        r.nbphi, r.nbtheta = int(nbphi), int(nbtheta)
        r._keep = [np.ascontiguousarray(tt), np.ascontiguousarray(sp), rank, np.ascontiguousarray(tans)]
        r.tan_btheta, r.sin_bphi, r.tie_rank, r.tan_phib = (a.ctypes.data for a in r._keep)
        return r


class Spectral(ctypes.Structure):
    """``cledb_spectral`` of the header: everything ``obs_integrate_work_1pix`` derives from the wavelength axis of one line,
    evaluated here with NumPy in the dtype the axis has (float32 in the reference, CLEDB_PREPINV.py:720)."""
    _fields_ = [('nw', _c.c_int32), ('wcont', _c.c_int32), ('lo', _c.c_int32), ('hi', _c.c_int32), ('uniform', _c.c_int32),
                ('dwv', _c.c_double), ('h2', _c.c_double), ('dx_first', _c.c_double), ('dx_last', _c.c_double),
                ('ga', _P), ('gb', _P), ('gc', _P)]

    @classmethod
    def build(cls, wl, wcont):
        wl = np.asarray(wl)
        if np.issubdtype(wl.dtype, np.integer):
            wl = wl.astype(np.float64)
        nw = wl.shape[0]
        r = cls()
        r.nw, r.wcont = int(nw), int(wcont)
        r.lo, r.hi = int(np.int32(0.25 * nw)), int(np.int32(0.75 * nw))          # :1210
        with np.errstate(all='ignore'):
            r.dwv = float(np.mean(np.diff(wl)))                                   # :1046
            dx = np.diff(wl)                                                      # numpy.gradient, non-uniform spacing
            r.uniform = 1 if (dx == dx[0]).all() else 0
            r.h2 = float(2. * dx[0])
            r.dx_first, r.dx_last = float(dx[0]), float(dx[-1])
            dx1, dx2 = dx[0:-1], dx[1:]
            a = -(dx2) / (dx1 * (dx1 + dx2))
            b = (dx2 - dx1) / (dx1 * dx2)
            c = dx1 / (dx2 * (dx1 + dx2))
        r._keep = [np.ascontiguousarray(t, dtype=np.float64) for t in (a, b, c)]
        r.ga, r.gb, r.gc = (t.ctypes.data for t in r._keep)
        return r


def partition(key, weight_of_key, nranks):
    """cledb_partition: ``(order[npix] int32, bounds[nranks+1] int64)``; rank r owns the pixels ``order[bounds[r]:bounds[r+1]]``.
    ``key``: per-pixel database index (db_enc), ``weight_of_key``: cost of one pixel of each database (its rows) or None."""
    key = _as(np.asarray(key).reshape(-1), np.int32)
    w = None if weight_of_key is None else _as(weight_of_key, np.float64)
    nkeys = int(w.size) if w is not None else (int(key.max()) + 1 if key.size else 1)
    order = np.empty(key.size, np.int32)
    bounds = np.empty(int(nranks) + 1, np.int64)
    rc = load_library().cledb_partition(key.size, _ptr(key), nkeys, _ptr(w), int(nranks), _ptr(order), _ptr(bounds))
    if rc != 0:
        raise CledbError('cledb_partition: bad argument (a key outside 0..nkeys-1?)')
    return order, bounds


def comm_unique_id():
    """128-byte NCCL id for ``Context.comm_init`` (made on rank 0, shipped to the other ranks by the launcher)."""
    buf = ctypes.create_string_buffer(128)
    rc = load_library().cledb_comm_unique_id(buf)
    if rc != 0:
        raise CledbError('cledb_comm_unique_id failed (%d): NCCL could not be loaded' % rc)
    return buf.raw


class Context:
    """One CUDA device.  Owns the resident databases and the search workspace."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.cledb_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise CledbError(self.lib.cledb_last_error(None).decode())
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, 'h', None):
            self.lib.cledb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise CledbError('cledb error %d: %s' % (rc, self.lib.cledb_last_error(self.h).decode()))
        return rc

    # ---- introspection
    def device_props(self):
        p = np.zeros(8, dtype=np.int64)
        self._check(self.lib.cledb_device_props(self.h, _ptr(p)))
        keys = ('sm_count', 'sm_clock_khz', 'hbm_bytes', 'l2_bytes', 'cc', 'match_ctas_per_sm', 'match_regs', 'match_smem')
        return dict(zip(keys, (int(v) for v in p)))

    def synchronize(self):
        self._check(self.lib.cledb_synchronize(self.h))

    def launch_count(self):
        return int(self.lib.cledb_launch_count(self.h))

    def profile_begin(self):
        self._check(self.lib.cledb_profile_begin(self.h))

    def profile_collect(self, cap=256):
        ms = np.zeros(cap, dtype=np.float32)
        n = self._check(self.lib.cledb_profile_collect(self.h, _ptr(ms), cap))
        return ms[:n].copy()

    def measure_fp32_peak(self):
        v = ctypes.c_double(0.0)
        self._check(self.lib.cledb_measure_fp32_peak(self.h, ctypes.byref(v)))
        return float(v.value)

    def set_tuning(self, pixels_per_item=32, split=0):
        self._check(self.lib.cledb_set_tuning(self.h, int(pixels_per_item), int(split)))

    def set_pruning(self, on):
        """Search variant: 0 / False = plain layout, brute force; 1 / True = exact tile pruning (default); 2 = k-d layout,
        seeded brute force (see the header); results do not change."""
        self._check(self.lib.cledb_set_pruning(self.h, int(on)))

    def pruning_stats(self):
        v = np.zeros(2, dtype=np.int64)
        self._check(self.lib.cledb_pruning_stats(self.h, _ptr(v)))
        return dict(evaluated=int(v[0]), skipped=int(v[1]))

    # ---- databases
    def db_put(self, db_id, db, encoding=ENC_I16):
        flat = _as(db, np.float32).reshape(-1, 8)
        self._check(self.lib.cledb_db_put(self.h, int(db_id), _ptr(flat), flat.shape[0], int(encoding)))

    def db_drop(self, db_id):
        self._check(self.lib.cledb_db_drop(self.h, int(db_id)))

    def db_info(self, db_id):
        v = np.zeros(16, dtype=np.int64)
        self._check(self.lib.cledb_db_info(self.h, int(db_id), _ptr(v)))
        return dict(n=int(v[0]), npos=int(v[1]), nneg=int(v[2]), nzero=int(v[3]), nnan=int(v[4]), encoding=int(v[5]),
                    hbm_bytes=int(v[6]), tile=int(v[7]), shift=[int(s) for s in v[8:16]])

    def db_decoded(self, db_id):
        n = self.db_info(db_id)['n']
        out = np.empty((n, 8), dtype=np.float32)
        self._check(self.lib.cledb_db_decoded(self.h, int(db_id), _ptr(out)))
        return out

    def gather_rows(self, db_id, idx):
        idx = _as(idx, np.int32).reshape(-1)
        out = np.empty((idx.size, 8), dtype=np.float32)
        self._check(self.lib.cledb_gather_rows(self.h, int(db_id), _ptr(idx), idx.size, _ptr(out)))
        return out

    def db_select(self, yobs, dobs, dbynumbers):
        """File index of the nearest (height, density) database for every pixel; sdb_findelongationanddens.
        Negative entries mark the pixels for which the reference raises (see the header)."""
        y = _as(np.asarray(yobs).reshape(-1), np.float32)
        d = _as(np.asarray(dobs).reshape(-1), np.float32)
        t = _as(dbynumbers, np.float32).reshape(-1, 3)
        out = np.empty(y.size, dtype=np.int32)
        self._check(self.lib.cledb_db_select(self.h, y.size, _ptr(y), _ptr(d), t.shape[0], _ptr(t), _ptr(out)))
        return out

    def obs_integrate(self, spec, wl, wcont):
        """obs_integrate_work_1pix for ``spec[npix,nw,4]`` float64 spectra of one line with wavelength axis ``wl[nw]``.
        Returns ``(tot, snr, background)`` float32 [npix,4], ``issue`` int32 [npix], ``status`` uint8 [npix]."""
        spec = _as(spec, np.float64)
        npix, nw = spec.shape[0], spec.shape[1]
        sp = Spectral.build(wl, wcont)
        tot, snr, bkg = (np.empty((npix, 4), np.float32) for _ in range(3))
        iss, st = np.empty(npix, np.int32), np.empty(npix, np.uint8)
        self._check(self.lib.cledb_obs_integrate(self.h, npix, _ptr(spec), ctypes.addressof(sp), _ptr(tot), _ptr(snr), _ptr(bkg),
                                                 _ptr(iss), _ptr(st)))
        return tot, snr, bkg, iss, st

    def obs_dens(self, a_obs, b_obs, yobs, heights, knots, coefs):
        """obs_dens_work_1pix for flat arrays of the two Stokes I and the heights; ``knots[nh,nknots]``/``coefs[nh,ncoef]``:
        the quadratic splines of every table height.  Returns ``(density float32 [npix], issue int32 [npix])``."""
        a = _as(np.asarray(a_obs).reshape(-1), np.float32)
        b = _as(np.asarray(b_obs).reshape(-1), np.float32)
        y = _as(np.asarray(yobs).reshape(-1), np.float32)
        h, t, c = _as(heights, np.float64), _as(knots, np.float64), _as(coefs, np.float64)
        dens, iss = np.empty(a.size, np.float32), np.empty(a.size, np.int32)
        self._check(self.lib.cledb_obs_dens(self.h, a.size, _ptr(a), _ptr(b), _ptr(y), h.size, _ptr(h), t.shape[1], _ptr(t),
                                            c.shape[1], _ptr(c), _ptr(dens), _ptr(iss)))
        return dens, iss

    # ---- search
    def match(self, mode, sobs, snr, db_id, nsearch, precision=PREC_FP32, want_rows=True, want_chi64=False, out=None):
        """Raw top-``nsearch`` of every pixel.  Returns a dict of NumPy arrays (see cledb_match in the header).
        ``out``: optional dict of preallocated (e.g. pinned) output arrays with the same keys, reused across calls."""
        sobs = _as(sobs, np.float32).reshape(-1, 8)
        snr = _as(snr, np.float32).reshape(-1, 8)
        npix = sobs.shape[0]
        db_id = _as(np.broadcast_to(np.asarray(db_id, dtype=np.int32).reshape(-1), (npix,)), np.int32)
        k = int(nsearch)
        if out is None:
            out = dict(idx=np.empty((npix, k), np.int32), chi=np.empty((npix, k), np.float32),
                       kpos=np.empty((npix, k), np.int32), status=np.empty(npix, np.uint8),
                       ncand=np.empty(npix, np.int64),
                       rows=np.empty((npix, k, 8), np.float32) if want_rows else None,
                       chi64=np.empty((npix, k), np.float64) if want_chi64 else None)
        self._check(self.lib.cledb_match(self.h, int(mode), int(precision), npix, _ptr(sobs), _ptr(snr), _ptr(db_id), k,
                                         _ptr(out['idx']), _ptr(out['chi']), _ptr(out['chi64']), _ptr(out['kpos']),
                                         _ptr(out['rows']), _ptr(out['status']), _ptr(out['ncand'])))
        return out

    def _invert_inputs(self, sobs, snr, db_id, yobs, dobs, aobs, dopp):
        sobs = _as(sobs, np.float32).reshape(-1, 8)
        snr = _as(snr, np.float32).reshape(-1, 8)
        npix = sobs.shape[0]
        db_id = np.asarray(db_id)
        if db_id.dtype != np.int32 or db_id.size != npix or not db_id.flags.c_contiguous:
            db_id = _as(np.broadcast_to(np.asarray(db_id, dtype=np.int32).reshape(-1), (npix,)), np.int32)
        yobs = _as(np.asarray(yobs).reshape(-1), np.float32)
        dobs = _as(np.asarray(dobs).reshape(-1), np.float32)
        ang = _as(np.asarray(aobs).reshape(-1), np.float32)
        with np.errstate(all='ignore'):
            twice = -2. * ang                     # cledb_quderotate, CLEDB_PROC.py:1584-1587, in the reference's float32
            rot_cos = np.cos(twice)
            rot_sin = np.sin(twice)
        is64 = 0
        if dopp is not None:
            dopp = np.asarray(dopp).reshape(-1)
            is64 = 1 if dopp.dtype == np.float64 else 0
            dopp = _as(dopp, np.float64 if is64 else np.float32)
        return sobs, snr, npix, db_id, yobs, dobs, rot_cos, rot_sin, dopp, is64

    @staticmethod
    def result_buffers(npix, k):
        """Fresh result arrays in pinned host memory (see PinnedPool)."""
        return dict(invout=pinned.empty((npix, k, 11), np.float32), sfound=pinned.empty((npix, k, 8), np.float32),
                    status=pinned.empty((npix,), np.uint8), issue=pinned.empty((npix,), np.int32))

    def invert(self, mode, sobs, snr, db_id, nsearch, yobs, dobs, aobs, dopp, dbhdr, maxchisq, bcalc, strict_topk=False,
               precision=PREC_FP32, out=None):
        """Search + solution building on the device (cledb_invert).  Returns ``(invout[npix,k,11], sfound[npix,k,8], status)``.
        ``dopp``: per-pixel B_POS (float32 or float64, [npix]) for IQUD, else None.  ``out``: dict of result arrays to fill
        (keys invout, sfound, status and optionally issue); default: fresh arrays in pinned memory.  The issue flags of the
        call are left in ``self.last_issue`` ([npix,2] uint8, see the header)."""
        sobs, snr, npix, db_id, yobs, dobs, rot_cos, rot_sin, dopp, is64 = self._invert_inputs(sobs, snr, db_id, yobs, dobs, aobs, dopp)
        k = int(nsearch)
        grid = DbGrid.from_dbhdr(dbhdr)
        if out is None:
            out = self.result_buffers(npix, k)
        issue = out.get('issue')
        self._check(self.lib.cledb_invert(self.h, int(mode), int(precision), npix, _ptr(sobs), _ptr(snr), _ptr(db_id), k,
                                          _ptr(yobs), _ptr(dobs), _ptr(rot_cos), _ptr(rot_sin), _ptr(dopp), is64, 1,
                                          ctypes.addressof(grid), float(maxchisq), int(bcalc), 1 if strict_topk else 0,
                                          _ptr(out['invout']), _ptr(out['sfound']), _ptr(out['status']), _ptr(issue)))
        self.last_issue = issue
        return out['invout'], out['sfound'], out['status']

    # ---- multi-GPU (one process per GPU)
    def comm_init(self, rank, nranks, unique_id=None):
        """Collective.  ``unique_id``: the 128 bytes rank 0 got from ``comm_unique_id()`` (shipped by the launcher)."""
        buf = None
        if nranks > 1:
            buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._check(self.lib.cledb_comm_init(self.h, int(rank), int(nranks), buf))
        self.rank, self.nranks = int(rank), int(nranks)

    def comm_destroy(self):
        self._check(self.lib.cledb_comm_destroy(self.h))

    def comm_info(self):
        v = np.zeros(4, np.int32)
        self._check(self.lib.cledb_comm_info(self.h, _ptr(v)))
        return dict(rank=int(v[0]), nranks=int(v[1]), nccl_version=int(v[2]), nccl=bool(v[3]))

    def comm_transport(self, want=None):
        """The transport of the gathers: ``'peer'`` (the library's own stores into the peers' windows over NVLink - default where
        CUDA IPC works) or ``'nccl'``.  ``want``: ``None`` queries, else one of the two names; every rank must ask for the same.
        Raises ``CledbError`` with the reason when peer memory is asked for and not available."""
        code = -1 if want is None else {'nccl': 0, 'peer': 1}[want]
        act = ctypes.c_int32(-1)
        self._check(self.lib.cledb_comm_transport(self.h, code, ctypes.byref(act)))
        return 'peer' if act.value == 1 else 'nccl'

    def comm_push_blocks(self, blocks):
        """Bound of the grid of the peer transport's push kernels (0 = none): a gather that runs on a second stream beside a
        search should leave the SMs to the search."""
        self._check(self.lib.cledb_comm_push_blocks(self.h, int(blocks)))

    def comm_window(self, nbytes):
        """Collective: device address of this rank's receive window of at least ``nbytes`` bytes (valid until a later call
        asks for more; ``invert_sharded`` uses the start of the same window)."""
        p = ctypes.c_void_p()
        self._check(self.lib.cledb_comm_window(self.h, int(nbytes), ctypes.byref(p)))
        return p.value

    def allgather_dev(self, send_ptr, recv_ptr, nbytes, stream=None):
        self._check(self.lib.cledb_allgather_dev(self.h, send_ptr, recv_ptr, int(nbytes), stream))

    def gather_maps_dev(self, npix, nlocal, my_order_ptr, maxcount, nsearch, send_ptr, maps_ptr, root=-1, stream=None):
        """cledb_gather_maps_dev (peer transport): all-gather + scatter of the packed records in one kernel."""
        self._check(self.lib.cledb_gather_maps_dev(self.h, int(npix), int(nlocal), my_order_ptr, int(maxcount), int(nsearch), send_ptr,
                                                   maps_ptr, int(root), stream))

    def unshard_dev(self, nranks, bounds_ptr, order_ptr, npix, maxcount, nsearch, recv_ptr, invout_ptr, sfound_ptr, status_ptr,
                    issue_ptr, stream=None):
        self._check(self.lib.cledb_unshard_dev(self.h, int(nranks), bounds_ptr, order_ptr, int(npix), int(maxcount), int(nsearch),
                                               recv_ptr, invout_ptr, sfound_ptr, status_ptr, issue_ptr, stream))

    def invert_gather_dev(self, mode, nlocal, sobs_ptr, snr_ptr, dbid_ptr, nsearch, yobs_ptr, dobs_ptr, rc_ptr, rs_ptr, dopp_ptr, dopp_is_f64,
                          grid, maxchisq, bcalc, npix, my_order_ptr, maps_ptr, root=-1, pieces=0, stream=None, strict_topk=False,
                          precision=PREC_FP32):
        """cledb_invert_gather_dev (peer transport): the shard searched and built in ``pieces`` pieces, the records of a piece
        stored at their pixels in the maps of every rank while the next piece is searched."""
        self._check(self.lib.cledb_invert_gather_dev(self.h, int(mode), int(precision), int(nlocal), sobs_ptr, snr_ptr, dbid_ptr, int(nsearch),
                                                     yobs_ptr, dobs_ptr, rc_ptr, rs_ptr, dopp_ptr, int(dopp_is_f64), ctypes.addressof(grid),
                                                     float(maxchisq), int(bcalc), 1 if strict_topk else 0, int(npix), my_order_ptr,
                                                     maps_ptr, int(root), int(pieces), stream))

    def invert_dev(self, mode, npix, sobs_ptr, snr_ptr, dbid_ptr, nsearch, yobs_ptr, dobs_ptr, rc_ptr, rs_ptr, dopp_ptr, dopp_is_f64,
                   grid, maxchisq, bcalc, invout_ptr, sfound_ptr, status_ptr, issue_ptr, stream=None, strict_topk=False,
                   precision=PREC_FP32):
        """cledb_invert_dev; ``grid``: a ``DbGrid`` whose four table pointers are device addresses."""
        self._check(self.lib.cledb_invert_dev(self.h, int(mode), int(precision), int(npix), sobs_ptr, snr_ptr, dbid_ptr, int(nsearch),
                                              yobs_ptr, dobs_ptr, rc_ptr, rs_ptr, dopp_ptr, int(dopp_is_f64), 1, ctypes.addressof(grid),
                                              float(maxchisq), int(bcalc), 1 if strict_topk else 0, invout_ptr, sfound_ptr, status_ptr,
                                              issue_ptr, stream))

    def invert_sharded(self, mode, sobs, snr, db_id, nsearch, yobs, dobs, aobs, dopp, dbhdr, maxchisq, bcalc, npix, order, bounds,
                       root=-1, strict_topk=False, precision=PREC_FP32, out=None):
        """cledb_invert_sharded: one map of ``npix`` pixels over all ranks of the communicator (collective).  The arrays are
        those of THIS rank's pixels ``order[bounds[rank]:bounds[rank+1]]``.  Returns ``(invout, sfound, status)`` of the FULL
        map on the ranks that receive it (``root < 0``: all), else Nones."""
        sobs, snr, nlocal, db_id, yobs, dobs, rot_cos, rot_sin, dopp, is64 = self._invert_inputs(sobs, snr, db_id, yobs, dobs, aobs, dopp)
        k = int(nsearch)
        grid = DbGrid.from_dbhdr(dbhdr)
        order, bounds = _as(order, np.int32), _as(bounds, np.int64)
        want = root < 0 or root == getattr(self, 'rank', 0)
        if out is None and want:
            out = self.result_buffers(int(npix), k)
        o = out or dict(invout=None, sfound=None, status=None, issue=None)
        self._check(self.lib.cledb_invert_sharded(self.h, int(mode), int(precision), nlocal, _ptr(sobs), _ptr(snr), _ptr(db_id), k,
                                                  _ptr(yobs), _ptr(dobs), _ptr(rot_cos), _ptr(rot_sin), _ptr(dopp), is64, 1,
                                                  ctypes.addressof(grid), float(maxchisq), int(bcalc), 1 if strict_topk else 0,
                                                  int(npix), _ptr(order), _ptr(bounds), int(root),
                                                  _ptr(o['invout']), _ptr(o['sfound']), _ptr(o['status']), _ptr(o.get('issue'))))
        self.last_issue = o.get('issue')
        return o['invout'], o['sfound'], o['status']

    def match_reduced(self, mode, sobs, snr, db_id, nsearch, sobs_dopp, dbhdr, precision=PREC_FP32, tie_order=None):
        """Raw top-``nsearch`` of every pixel over its reduced database (``reduced=True``); dict as ``match`` plus
        ``nan_slots`` (leading output slots the reference leaves empty)."""
        sobs = _as(sobs, np.float32).reshape(-1, 8)
        snr = _as(snr, np.float32).reshape(-1, 8)
        npix = sobs.shape[0]
        db_id = _as(np.broadcast_to(np.asarray(db_id, dtype=np.int32).reshape(-1), (npix,)), np.int32)
        k = int(nsearch)
        red = Reduced.build(mode, sobs, sobs_dopp, dbhdr, tie_order)
        out = dict(idx=np.empty((npix, k), np.int32), chi=np.empty((npix, k), np.float32), kpos=np.empty((npix, k), np.int32),
                   status=np.empty(npix, np.uint8), rows=np.empty((npix, k, 8), np.float32), nan_slots=np.empty(npix, np.int32),
                   chi64=np.empty((npix, k), np.float64) if precision == PREC_FP64 else None, ncand=None)
        self._check(self.lib.cledb_match_reduced(self.h, int(mode), int(precision), npix, _ptr(sobs), _ptr(snr), _ptr(db_id), k,
                                                 ctypes.addressof(red), _ptr(out['idx']), _ptr(out['chi']), _ptr(out['chi64']),
                                                 _ptr(out['kpos']), _ptr(out['rows']), _ptr(out['status']), _ptr(out['nan_slots'])))
        return out

    def match_reduced_dev(self, mode, npix, sobs_ptr, snr_ptr, dbid_ptr, nsearch, red, idx_ptr, chi_ptr, chi64_ptr, kpos_ptr, rows_ptr,
                          status_ptr, nan_ptr, stream=None, precision=PREC_FP32):
        """cledb_match_reduced_dev; ``red``: a ``Reduced`` whose four table pointers are device addresses."""
        self._check(self.lib.cledb_match_reduced_dev(self.h, int(mode), int(precision), int(npix), sobs_ptr, snr_ptr, dbid_ptr, int(nsearch),
                                                     ctypes.addressof(red), idx_ptr, chi_ptr, chi64_ptr, kpos_ptr, rows_ptr, status_ptr,
                                                     nan_ptr, stream))

    def invert_reduced(self, mode, sobs, snr, db_id, nsearch, yobs, dobs, aobs, sobs_dopp, dbhdr, maxchisq, bcalc,
                       strict_topk=False, precision=PREC_FP32, tie_order=None):
        """Reduced-database search + solution building on the device (cledb_invert_reduced)."""
        sobs = _as(sobs, np.float32).reshape(-1, 8)
        snr = _as(snr, np.float32).reshape(-1, 8)
        npix = sobs.shape[0]
        db_id = _as(np.broadcast_to(np.asarray(db_id, dtype=np.int32).reshape(-1), (npix,)), np.int32)
        k = int(nsearch)
        yobs = _as(np.asarray(yobs).reshape(-1), np.float32)
        dobs = _as(np.asarray(dobs).reshape(-1), np.float32)
        ang = _as(np.asarray(aobs).reshape(-1), np.float32)
        with np.errstate(all='ignore'):
            rot_cos, rot_sin = np.cos(-2. * ang), np.sin(-2. * ang)
        dopp0, is64 = None, 0
        if mode == MODE_IQUD:
            dopp0 = np.asarray(sobs_dopp)[:, 0]
            is64 = 1 if dopp0.dtype == np.float64 else 0
            dopp0 = _as(dopp0, np.float64 if is64 else np.float32)
        red = Reduced.build(mode, sobs, sobs_dopp, dbhdr, tie_order)
        grid = DbGrid.from_dbhdr(dbhdr)
        invout, sfound = np.empty((npix, k, 11), np.float32), np.empty((npix, k, 8), np.float32)
        status, issue = np.empty(npix, np.uint8), np.empty(npix, np.int32)
        self._check(self.lib.cledb_invert_reduced(self.h, int(mode), int(precision), npix, _ptr(sobs), _ptr(snr), _ptr(db_id), k,
                                                  ctypes.addressof(red), _ptr(yobs), _ptr(dobs), _ptr(rot_cos), _ptr(rot_sin),
                                                  _ptr(dopp0), is64, 1, ctypes.addressof(grid), float(maxchisq), int(bcalc),
                                                  1 if strict_topk else 0, _ptr(invout), _ptr(sfound), _ptr(status), _ptr(issue)))
        self.last_issue = issue
        return invout, sfound, status

    def match_dev(self, mode, npix, sobs_ptr, snr_ptr, dbid_ptr, nsearch, idx_ptr, chi_ptr, chi64_ptr=None, kpos_ptr=None,
                  rows_ptr=None, status_ptr=None, ncand_ptr=None, stream=None, precision=PREC_FP32):
        self._check(self.lib.cledb_match_dev(self.h, int(mode), int(precision), int(npix), sobs_ptr, snr_ptr, dbid_ptr,
                                             int(nsearch), idx_ptr, chi_ptr, chi64_ptr, kpos_ptr, rows_ptr, status_ptr,
                                             ncand_ptr, stream))

    # ---- elementwise
    def blos(self, iquv, kfac, g_eff, ffac):
        iquv = _as(iquv, np.float32).reshape(-1, 4)
        out = np.empty_like(iquv)
        flags = np.empty((iquv.shape[0], 2), dtype=np.uint8)
        self._check(self.lib.cledb_blos(self.h, iquv.shape[0], _ptr(iquv), float(kfac), float(g_eff), float(ffac),
                                        _ptr(out), _ptr(flags)))
        return out, flags

    def blos_dev(self, npix, in_ptr, kfac, g_eff, ffac, out_ptr, flags_ptr, stream=None):
        self._check(self.lib.cledb_blos_dev(self.h, int(npix), in_ptr, float(kfac), float(g_eff), float(ffac), out_ptr,
                                            flags_ptr, stream))

    def qurotate(self, sobs_tot, xs, ys, linpolref, hpxy=None):
        s = _as(sobs_tot, np.float32)
        nx, ny = s.shape[0], s.shape[1]
        out = np.empty_like(s)
        aobs = np.empty((nx, ny), np.float32)
        yobs = np.empty((nx, ny), np.float32)
        xs = None if xs is None else _as(xs, np.float64)
        ys = None if ys is None else _as(ys, np.float64)
        hp = None if hpxy is None else _as(hpxy, np.float32)
        self._check(self.lib.cledb_qurotate(self.h, nx, ny, _ptr(xs), _ptr(ys), float(linpolref), _ptr(hp), _ptr(s),
                                            _ptr(out), _ptr(aobs), _ptr(yobs)))
        return out, aobs, yobs

    def qurotate_dev(self, nx, ny, xs_ptr, ys_ptr, linpolref, hpxy_ptr, in_ptr, out_ptr, aobs_ptr, yobs_ptr, stream=None):
        self._check(self.lib.cledb_qurotate_dev(self.h, int(nx), int(ny), xs_ptr, ys_ptr, float(linpolref), hpxy_ptr,
                                                in_ptr, out_ptr, aobs_ptr, yobs_ptr, stream))


_contexts = {}


def get_context(device=0):
    """Process-wide context per device (the reference's functions are stateless; the GPU state lives here)."""
    ctx = _contexts.get(device)
    if ctx is None or ctx.h is None:
        ctx = _contexts[device] = Context(device)
    return ctx
