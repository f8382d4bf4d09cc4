"""Shared helpers of the test-suite (CPU and GPU)."""
import numpy as np

import oracle.cledb_oracle as orc
import cledb_b200.synth as synth
from cledb_b200 import native


def hdr_from_g(g):
    h = synth.make_dbhdr(int(g[2]), int(g[3]), int(g[4]), dbtype=int(g[-1]))
    assert np.array_equal(h[0], g)
    return h


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)


def oracle_raw(mode, sobs, snr, dbs, enc, nsearch):
    """The raw search result (what cledb_match returns) computed by the oracle, in the layout of native.Context.match."""
    npix = sobs.shape[0]
    k = nsearch
    out = dict(idx=np.full((npix, k), -1, np.int32), chi=np.zeros((npix, k), np.float32), chi64=np.zeros((npix, k)),
               kpos=np.full((npix, k), -1, np.int32), rows=np.zeros((npix, k, 8), np.float32),
               status=np.zeros(npix, np.uint8), ncand=np.zeros(npix, np.int64))
    for p in range(npix):
        db = np.reshape(dbs[int(enc[p])], (-1, 8))
        r = orc.raw_topk(mode, sobs[p], snr[p], db, k, return_all=True)
        if r is None:
            out['status'][p] = native.PIX_NULL
            continue
        idx, chi, pos, ncand, allchi = r
        out['ncand'][p] = ncand
        if ncand == 0:
            out['status'][p] = native.PIX_REFRAISES
            continue
        if np.isnan(allchi).any():         # argmin lands on a NaN: every slot fails `chi <= maxchisq`
            out['status'][p] = native.PIX_ALLNAN
            continue
        if np.isinf(chi).all():
            out['status'][p] = native.PIX_ALLINF
            continue
        n = idx.size
        out['idx'][p, :n], out['chi'][p, :n], out['chi64'][p, :n], out['kpos'][p, :n] = idx, chi, chi, pos
        out['rows'][p, :n] = db[idx]
    return out


class _DeviceBytes:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = dict(shape=(int(nbytes),), typestr='|u1', data=(int(ptr), False), version=2)


def device_view(ptr, nbytes, device):
    """uint8 torch tensor over ``nbytes`` of device memory at ``ptr`` (memory the library owns, e.g. the receive window)."""
    import torch
    return torch.as_tensor(_DeviceBytes(ptr, nbytes), device=device)
