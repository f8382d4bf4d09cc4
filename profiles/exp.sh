mkdir -p gpurun_out/exp
run() { # name, args
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline "${@:2}" > gpurun_out/exp/$1.json 2>gpurun_out/exp/$1.err
  python -c "
import json
try:
    j=json.loads(open('gpurun_out/exp/$1.json').read().strip().splitlines()[-1])
    print('%-22s kernel_ms %.3f frac %.3f  ctas/sm %s regs %s'%('$1', j['roofline']['kernel_ms'], j['roofline']['frac'], j['device']['match_ctas_per_sm'], j['device']['match_regs']))
except Exception as e: print('$1 failed', e)"
}
run base
run dbg2 --dbg 2
run iqud --mode iqud
run nx512 --nx 512
