python -m pytest tests -m gpu -x -q -k "pruning" 2>&1 | tail -5
mkdir -p gpurun_out/exp
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline "${@:2}" > gpurun_out/exp/$1.json 2>gpurun_out/exp/$1.err
  python -c "
import json
try:
    j=json.loads(open('gpurun_out/exp/$1.json').read().strip().splitlines()[-1])
    print('%-22s value %.4e e2e %.4e kernel_ms %.3f frac %.3f'%('$1', j['value'], j['e2e']['value'], j['roofline']['kernel_ms'], j['roofline']['frac']), j.get('pruning'))
except Exception as e: print('$1 failed', e); print(open('gpurun_out/exp/$1.err').read()[-800:])"
}
for p in 4 8 16 32; do run pruned_p$p --prune --ppi $p; done
run pruned_iqud --prune --mode iqud
run pruned_1024 --prune --nx 1024
run pruned_f32 --prune --encoding 0
