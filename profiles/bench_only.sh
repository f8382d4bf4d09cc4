#!/bin/bash
# both benches, no tests.  Usage: gpurun -- 'bash profiles/bench_only.sh tag'
tag=${1:-b}; out=gpurun_out/$tag; mkdir -p $out
for m in iquv iqud; do
timeout 600 python bench.py --steps 20 --warmup 3 --mode $m --no-cpu-baseline > $out/bench_$m.json 2> $out/bench_$m.err
python - <<PY
import json
try:
    j=json.loads(open('$out/bench_$m.json').read().strip().splitlines()[-1])
    print('$m value %.4e e2e %.4e frac %.4f kernel_ms %.3f step_ms %.3f regs %s'%(j['value'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['kernel_ms'], j['ms_per_step'], j['device']['match_regs']))
except Exception as e: print('bench parse failed', e); print(open('$out/bench_$m.err').read()[-2000:])
PY
done
