mkdir -p gpurun_out/exp
python - <<'PY'
import sys, time
sys.path[:0] = ['.', 'solar-coronal-inversion_b200']
from cledb_b200 import native, synth
db = synth.make_database(9, 40)
ctx = native.Context(0)
for prune in (False, True, False, True):
    ctx.set_pruning(prune)
    t = time.perf_counter(); ctx.db_put(0, db, native.ENC_I16); dt = time.perf_counter() - t
    print('db_put pruning=%s: %.1f ms' % (prune, dt * 1e3)); ctx.db_drop(0)
PY
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-prune-pass > gpurun_out/exp/b.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/exp/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-prune-pass > gpurun_out/exp/ncu.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "chunks or sdb_preprocess or invproc" 2>&1 | tail -3
