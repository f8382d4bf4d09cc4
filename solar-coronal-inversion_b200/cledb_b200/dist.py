"""Multi-GPU driver of the 2-line inversion: one process per GPU, pixels sharded, database replicated.

The search has no exchange step - every pixel is independent (the reference dispatches one pool task per pixel,
CLEDB_PROC.py:304-313) - so the only collective is the gather of the finished result maps at the end
(``torch.distributed.all_gather`` over NCCL/NVLink on GPUs, gloo in the CPU tests).

    rank r:  pixels = partition(...)[r]            contiguous ranges of the bucket-sorted pixel list, balanced by
                                                    the number of candidate evaluations, so a rank touches few databases
             invout_r, sfound_r = cledb_invproc(sub-map of its pixels)      (GPU search + host solution building)
             all_gather(invout_r, sfound_r, issuemask_r)  ->  full [nx,ny,...] maps on every rank
"""
import numpy as np


def partition(db_enc, weight_of_db, world):
    """Split the pixels of a map into ``world`` index arrays.

    Pixels are ordered by database (so that a rank needs only the databases of its range resident in HBM) and cut
    into contiguous ranges of equal total weight; ``weight_of_db[k]`` = candidate evaluations one pixel of database
    ``k`` costs (its entry count).  Returns a list of int64 arrays of flat pixel indices; their union is every pixel
    exactly once."""
    enc = np.asarray(db_enc).reshape(-1)
    order = np.argsort(enc, kind='stable')
    w = np.asarray(weight_of_db, dtype=np.float64)[enc[order]]
    csum = np.cumsum(w)
    total = csum[-1] if csum.size else 0.0
    cuts = [int(np.searchsorted(csum, total * r / world, side='left')) for r in range(1, world)]
    bounds = [0] + cuts + [order.size]
    return [order[bounds[r]:bounds[r + 1]].astype(np.int64) for r in range(world)]


def _all_gather_rows(local, counts, dist, device):
    """all_gather of row blocks with different row counts (padded to the largest); returns the list of blocks."""
    import torch
    mx = max(counts)
    pad = np.zeros((mx,) + local.shape[1:], dtype=local.dtype)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(pad).to(device)
    outs = [torch.empty_like(t) for _ in counts]
    dist.all_gather(outs, t)
    return [o.cpu().numpy()[:c] for o, c in zip(outs, counts)]


def invproc_sharded(invproc, sobs_totrot, sobs_dopp, database, db_enc, yobs, aobs, dobs, snr, issuemask, dbhdr, keyvals,
                    nsearch, maxchisq, bcalc, iqud, ncpu, reduced, verbose, rank=None, world=None, device='cpu'):
    """``cledb_invproc`` over every rank of the default process group.  ``invproc`` is the single-device function
    (``CLEDB_PROC.cledb_invproc``; the CPU tests inject the oracle).  Every rank receives the full maps and returns
    the full ``(invout, sfound)``; ``issuemask`` is updated in place on every rank."""
    import torch.distributed as dist
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    nx, ny = keyvals[0:2]
    npix = nx * ny
    weights = [int(np.prod(np.shape(d)[:-1])) for d in database]
    parts = partition(db_enc, weights, world)
    mine = parts[rank]
    n = mine.size
    flat = lambda a, c: np.asarray(a).reshape(npix, c)[mine].reshape(n, 1, c)     # my pixels as an [n,1] map
    sub_kv = list(keyvals)
    sub_kv[0], sub_kv[1] = n, 1
    sub_iss = np.asarray(issuemask).reshape(npix, -1)[mine].reshape(n, 1, -1).copy()
    dopp = np.asarray(sobs_dopp)
    sub_dopp = dopp.reshape(npix, -1)[mine].reshape(n, 1, -1) if dopp.ndim >= 2 and dopp.size >= npix else sobs_dopp
    if n:
        inv, sf = invproc(flat(sobs_totrot, 8), sub_dopp, database, np.asarray(db_enc).reshape(-1)[mine].reshape(n, 1),
                          np.asarray(yobs).reshape(-1)[mine].reshape(n, 1), np.asarray(aobs).reshape(-1)[mine].reshape(n, 1),
                          np.asarray(dobs).reshape(-1)[mine].reshape(n, 1), flat(snr, 8), sub_iss, dbhdr, sub_kv,
                          nsearch, maxchisq, bcalc, iqud, ncpu, reduced, verbose)
    else:
        inv = np.zeros((0, 1, nsearch, 11), np.float32)
        sf = np.zeros((0, 1, nsearch, 8), np.float32)
    counts = [p.size for p in parts]
    inv_blocks = _all_gather_rows(inv.reshape(n, nsearch * 11), counts, dist, device)
    sf_blocks = _all_gather_rows(sf.reshape(n, nsearch * 8), counts, dist, device)
    iss_blocks = _all_gather_rows(sub_iss.reshape(n, -1).astype(np.int64), counts, dist, device)
    invout = np.zeros((npix, nsearch, 11), np.float32)
    sfound = np.zeros((npix, nsearch, 8), np.float32)
    iss_flat = np.asarray(issuemask).reshape(npix, -1)
    for p, a, b, c in zip(parts, inv_blocks, sf_blocks, iss_blocks):
        invout[p] = a.reshape(-1, nsearch, 11)
        sfound[p] = b.reshape(-1, nsearch, 8)
        iss_flat[p] = c.astype(iss_flat.dtype)
    np.asarray(issuemask).reshape(npix, -1)[...] = iss_flat
    return invout.reshape(nx, ny, nsearch, 11), sfound.reshape(nx, ny, nsearch, 8)
