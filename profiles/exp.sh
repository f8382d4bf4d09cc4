mkdir -p gpurun_out/exp
for d in 0 2; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dbg $d > gpurun_out/exp/dbg$d.json 2>&1
  python -c "
import json,sys
j=json.loads(open('gpurun_out/exp/dbg$d.json').read().strip().splitlines()[-1])
print('dbg $d kernel_ms %.3f frac %.3f'%(j['roofline']['kernel_ms'], j['roofline']['frac']))"
done
