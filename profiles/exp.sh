python -m pytest tests -m gpu -x -q -k "pruning" 2>&1 | tail -5
mkdir -p gpurun_out/exp
run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-prune-pass "${@:2}" > gpurun_out/exp/$1.json 2>gpurun_out/exp/$1.err
  python -c "
import json
try:
    j=json.loads(open('gpurun_out/exp/$1.json').read().strip().splitlines()[-1])
    print('%-22s value %.4e e2e %.4e kernel_ms %.3f step_ms %.3f'%('$1', j['value'], j['e2e']['value'], j['roofline']['kernel_ms'], j['ms_per_step']), j.get('pruning'))
except Exception as e: print('$1 failed', e); print(open('gpurun_out/exp/$1.err').read()[-800:])"
}
run p256
run p256_p8 --ppi 8
run p256_p16 --ppi 16
run p512 --nx 512
run p512_p16 --nx 512 --ppi 16
run p1024 --nx 1024
run p1024_p16 --nx 1024 --ppi 16
run iqud --mode iqud
run f32 --encoding 0
