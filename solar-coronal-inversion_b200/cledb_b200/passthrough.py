"""Pass-through of the reference's functions that are OUTSIDE the accelerated path.

The drop-in modules ``CLEDB_PROC/CLEDB_PROC.py`` and ``CLEDB_PREPINV/CLEDB_PREPINV.py`` shadow the reference's modules of
the same name when this package is ahead of the reference checkout on ``sys.path``.  They define only the hot path
(SURVEY.md section 8); every other name a driver script asks for (``sobs_preprocess``, ``spectro_proc``, ``obs_integrate``
...) is looked up in the reference's own module, loaded by file path from

    $CLEDB_REFERENCE_ROOT            if set, else
    the first later ``sys.path`` entry that holds the same relative file.

Nothing of the hot path ever goes through here: names the drop-in modules define are never looked up in the reference.
"""
import importlib.util
import os
import sys
import types

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_loaded = {}


def _find(relpath):
    root = os.environ.get('CLEDB_REFERENCE_ROOT')
    roots = [root] if root else [p for p in sys.path if p and os.path.abspath(p) != _PKG]
    for r in roots:
        f = os.path.join(r, relpath)
        if os.path.isfile(f) and os.path.abspath(r) != _PKG:
            return f
    return None


def reference_attr(relpath, name, dropin=None):
    """Attribute ``name`` of the reference module at ``relpath`` (e.g. 'CLEDB_PREPINV/CLEDB_PREPINV.py').

    ``dropin``: the globals of the drop-in module.  Every function it defines replaces the function of the same name
    inside the loaded reference module, so that reference code which calls a hot-path function internally
    (``sobs_preprocess`` -> ``obs_qurotate`` / ``obs_calcheight``) runs it on the GPU as well."""
    mod = _loaded.get(relpath)
    if mod is None:
        f = _find(relpath)
        if f is None:
            raise AttributeError(
                "'%s' is outside the B200-accelerated path and the reference checkout was not found: set CLEDB_REFERENCE_ROOT "
                "(or keep the reference on sys.path after this package) to pass it through" % name)
        # registered in sys.modules under an importable name: the reference pickles its worker functions by module name
        # when it dispatches them to its multiprocessing pool
        if 'cledb_reference' not in sys.modules:
            parent = types.ModuleType('cledb_reference')
            parent.__path__ = []
            sys.modules['cledb_reference'] = parent
        full = 'cledb_reference.' + os.path.basename(relpath)[:-3]
        spec = importlib.util.spec_from_file_location(full, f)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        try:
            spec.loader.exec_module(mod)
        except BaseException:
            del sys.modules[full]
            raise
        if dropin is not None:
            for key, val in dropin.items():
                if callable(val) and getattr(val, '__module__', None) == dropin.get('__name__') and hasattr(mod, key) \
                        and not key.startswith('_'):
                    setattr(mod, key, val)
        _loaded[relpath] = mod
    try:
        return getattr(mod, name)
    except AttributeError:
        raise AttributeError("neither the B200 drop-in nor the reference module %s defines '%s'" % (relpath, name)) from None
