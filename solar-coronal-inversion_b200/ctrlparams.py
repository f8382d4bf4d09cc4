"""Inversion control parameters with the attribute names and defaults of the reference's
``ctrlparams.ctrlparams`` (ctrlparams.py:6-42), plus the B200-specific switches (all optional).

``ncpu`` and the three ``jit*`` flags are accepted for source compatibility and ignored: the search runs
on the GPU(s), there is no Numba and no process pool.
"""

_REFERENCE_DEFAULTS = dict(
    dbdir='./CLEDB_BUILD/', verbose=2, ncpu=100,
    integrated=False, dblinpolref=0, instwidth=0, atmred=False,
    nsearch=8, maxchisq=1000, gaussfit=2, bcalc=0, reduced=False, iqud=False,
    jitparallel=False, jitcache=False, jitdisable=True)

_B200_DEFAULTS = dict(
    device=0,            # CUDA device ordinal of this process
    dbencoding=1,        # 1 = int16 codes in HBM (scaled binary16 pattern), 0 = lossless float32
    precision=0,         # 0 = fp32 production kernel, 1 = fp64 validation kernel (reference operation order)
    strict_topk=False)   # False = reproduce the reference's cledb_partsort index quirk, True = true top-nsearch


class ctrlparams:
    def __init__(self):
        for k, v in _REFERENCE_DEFAULTS.items():
            setattr(self, k, v)
        self.lookuptb = self.dbdir + 'chianti_v10.1_pycelp_fe13_h99_d120_ratio.npz'
        for k, v in _B200_DEFAULTS.items():
            setattr(self, k, v)
