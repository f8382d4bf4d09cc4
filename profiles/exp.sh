python -m pytest tests -m gpu -x -q 2>&1 | tail -3
mkdir -p gpurun_out/exp
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline "${@:2}" > gpurun_out/exp/$1.json 2>gpurun_out/exp/$1.err
  python -c "
import json
try:
    j=json.loads(open('gpurun_out/exp/$1.json').read().strip().splitlines()[-1])
    r=j['roofline']; rp=j.get('roofline_pruned') or {}
    print('%-12s value %.4e e2e %.4e step_ms %.3f | pruned kernel %.3f ms frac %.3f alg %.2f | brute kernel %.3f frac %.3f'%('$1', j['value'], j['e2e']['value'], j['ms_per_step'], rp.get('kernel_ms',0), rp.get('frac',0), rp.get('frac_algorithmic',0), r['kernel_ms'], r['frac']))
except Exception as e: print('$1 failed', e); print(open('gpurun_out/exp/$1.err').read()[-800:])"
}
run p256
run iqud --mode iqud
run p1024 --nx 1024
