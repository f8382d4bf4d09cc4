#!/bin/bash
# quick GPU pass: parity tests + both benches, no ncu.  Usage: gpurun -- 'bash profiles/quick.sh tag'
tag=${1:-q}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest.log
for m in iquv iqud; do
timeout 600 python bench.py --steps 20 --warmup 3 --mode $m --no-cpu-baseline > $out/bench_$m.json 2> $out/bench_$m.err; echo "bench $m exit $?"
python - <<PY
import json
try:
    j=json.loads(open('$out/bench_$m.json').read().strip().splitlines()[-1])
    print('$m value %.4e e2e %.4e frac %.4f kernel_ms %.3f step_ms %.3f'%(j['value'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['kernel_ms'], j['ms_per_step']))
except Exception as e: print('bench parse failed', e); print(open('$out/bench_$m.err').read()[-2000:])
PY
done
