mkdir -p gpurun_out/exp
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline "${@:2}" > gpurun_out/exp/$1.json 2>gpurun_out/exp/$1.err
  python -c "
import json
try:
    j=json.loads(open('gpurun_out/exp/$1.json').read().strip().splitlines()[-1])
    print('%-22s kernel_ms %.3f frac %.3f'%('$1', j['roofline']['kernel_ms'], j['roofline']['frac']))
except Exception as e: print('$1 failed', e); print(open('gpurun_out/exp/$1.err').read()[-800:])"
}
for p in 18 24 30 34 36 38 40; do run iquv_p$p --ppi $p --split 1; done
for p in 18 24 30 34 36 38 40; do run iqud_p$p --mode iqud --ppi $p --split 1; done
run iquv_auto
run iqud_auto --mode iqud
