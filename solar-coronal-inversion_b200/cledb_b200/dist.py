"""Multi-GPU driver of the 2-line inversion: one process per GPU, pixels sharded, databases replicated per range.

The search has no exchange step - every pixel is independent (the reference dispatches one pool task per pixel,
CLEDB_PROC.py:304-313) - so the only collective is the gather of the finished result maps at the end.

On GPUs (``device='cuda'``) that gather happens inside the library: ``init_comm()`` creates the communicator on the
process's context (``cledb_comm_init``; the 128-byte NCCL id travels through ``torch.distributed``; the ranks map each other's
receive windows with CUDA IPC), and ``cledb_invproc`` with ``OPTIONS['sharded']`` becomes a collective:

    rank r:  order, bounds = cledb_partition(db_enc, rows per database)   the same on every rank
             pixels r      = order[bounds[r] : bounds[r+1]]               contiguous in database order: few databases per rank
             cledb_invert_sharded: search + solutions of the shard on the device, packed records in the send buffer
               peer-memory transport (default): k_push_records stores every record at its pixel in the maps of every
                 rank over NVLink / NVSwitch (from six ranks on in two pieces, the first crossing NVLink while the second is searched)
               NCCL transport (``ctx.comm_transport('nccl')``): ONE ncclAllGather -> k_unshard scatters the records -> full maps

Nothing goes through NumPy or host memory between the search and the gathered maps.  The CPU path below (``device='cpu'``,
gloo) runs the same partition around an injected single-device function and gathers one packed buffer per rank; the CPU
tests use it with the oracle as that function.
"""
import numpy as np


def partition(db_enc, weight_of_db, world):
    """Split the pixels of a map into ``world`` index arrays (NumPy statement of ``cledb_partition``).

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


def init_comm(ctx=None):
    """Create the library's communicator over the ranks of ``torch.distributed``'s default group (collective).  ``ctx``:
    the ``native.Context`` to attach it to; default: the context ``CLEDB_PROC`` uses on this process."""
    import torch.distributed as dist
    from . import native
    if ctx is None:
        import CLEDB_PROC.CLEDB_PROC as procinv
        ctx = procinv._context()
    if getattr(ctx, 'nranks', 0) == dist.get_world_size():
        return ctx
    rank, world = dist.get_rank(), dist.get_world_size()
    ids = [native.comm_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(rank, world, ids[0])
    return ctx


def invproc_sharded(invproc, sobs_totrot, sobs_dopp, database, db_enc, yobs, aobs, dobs, snr, issuemask, dbhdr, keyvals,
                    nsearch, maxchisq, bcalc, iqud, ncpu, reduced, verbose, rank=None, world=None, device='cpu', root=-1):
    """``cledb_invproc`` over every rank of the default process group.  Every rank is called with the full maps and
    returns the full ``(invout, sfound)`` (``root >= 0``: only that rank, the others get ``(None, None)``); ``issuemask``
    is updated in place on the receiving ranks.

    ``device='cuda'``: ``invproc`` must be ``CLEDB_PROC.cledb_invproc``; the shard search, the all-gather (NCCL) and the
    scatter into the full maps run inside the library (see the module docstring).  ``device='cpu'``: ``invproc`` is any
    single-device function with the reference's signature (the CPU tests inject the oracle); one gloo ``all_gather`` of a
    packed per-rank buffer."""
    if str(device).startswith('cuda'):
        import CLEDB_PROC.CLEDB_PROC as procinv
        init_comm()
        old = procinv.OPTIONS.get('sharded'), procinv.OPTIONS.get('shard_root', -1)
        procinv.OPTIONS.update(sharded=True, shard_root=root)
        try:
            return procinv.cledb_invproc(sobs_totrot, sobs_dopp, database, db_enc, yobs, aobs, dobs, snr, issuemask, dbhdr, keyvals,
                                         nsearch, maxchisq, bcalc, iqud, ncpu, reduced, verbose)
        finally:
            procinv.OPTIONS.update(sharded=old[0], shard_root=old[1])

    import torch
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
    nl = np.asarray(issuemask).reshape(npix, -1).shape[1]
    sub_iss = np.asarray(issuemask).reshape(npix, nl)[mine].reshape(n, 1, nl).copy()
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
    # one packed record per pixel - invout | sfound | issue mask (bit patterns in float32 lanes) - and ONE collective
    counts = [p.size for p in parts]
    width = nsearch * 19 + nl
    pack = np.zeros((max(counts), width), np.float32)
    pack[:n, :nsearch * 11] = inv.reshape(n, nsearch * 11)
    pack[:n, nsearch * 11:nsearch * 19] = sf.reshape(n, nsearch * 8)
    pack[:n, nsearch * 19:] = sub_iss.reshape(n, nl).astype(np.int32).view(np.float32)
    t = torch.from_numpy(pack)
    gathered = torch.empty((world * pack.shape[0], width), dtype=t.dtype)
    dist.all_gather_into_tensor(gathered, t)
    if root >= 0 and root != rank:
        return None, None
    g = gathered.numpy().reshape(world, pack.shape[0], width)
    invout = np.zeros((npix, nsearch, 11), np.float32)
    sfound = np.zeros((npix, nsearch, 8), np.float32)
    iss_flat = np.asarray(issuemask).reshape(npix, nl)
    for r, p in enumerate(parts):
        blk = g[r, :p.size]
        invout[p] = blk[:, :nsearch * 11].reshape(-1, nsearch, 11)
        sfound[p] = blk[:, nsearch * 11:nsearch * 19].reshape(-1, nsearch, 8)
        iss_flat[p] = np.ascontiguousarray(blk[:, nsearch * 19:]).view(np.int32).astype(iss_flat.dtype)
    np.asarray(issuemask).reshape(npix, nl)[...] = iss_flat
    return invout.reshape(nx, ny, nsearch, 11), sfound.reshape(nx, ny, nsearch, 8)
