mkdir -p gpurun_out/exp
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline "${@:2}" > gpurun_out/exp/$1.json 2>gpurun_out/exp/$1.err
  python -c "
import json
try:
    j=json.loads(open('gpurun_out/exp/$1.json').read().strip().splitlines()[-1])
    print('%-22s value %.4e kernel_ms %.3f frac %.3f ctas/sm %s regs %s'%('$1', j['value'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['device']['match_ctas_per_sm'], j['device']['match_regs']))
except Exception as e: print('$1 failed', e); print(open('gpurun_out/exp/$1.err').read()[-800:])"
}
run nx256
run nx1024 --nx 1024
run nx1024iqud --nx 1024 --mode iqud
