mkdir -p gpurun_out/exp
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline "${@:2}" > gpurun_out/exp/$1.json 2>gpurun_out/exp/$1.err
  python -c "
import json
try:
    j=json.loads(open('gpurun_out/exp/$1.json').read().strip().splitlines()[-1])
    print('%-22s value %.4e e2e %.4e kernel_ms %.3f frac %.3f smem %s'%('$1', j['value'], j['e2e']['value'], j['roofline']['kernel_ms'], j['roofline']['frac'], j['device']['match_smem']))
except Exception as e: print('$1 failed', e); print(open('gpurun_out/exp/$1.err').read()[-800:])"
}
run i16_iquv
run i16_iqud --mode iqud
run f32_iquv --encoding 0
run f32_iqud --encoding 0 --mode iqud
run i16_nx1024 --nx 1024
run f32_nx1024 --nx 1024 --encoding 0
