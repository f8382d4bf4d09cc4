timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in iquv iqud; do python bench.py --steps 20 --no-cpu-baseline --no-prune-pass --mode $m | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('$m value %.4e e2e %.4e step %.3f | pruned kernel %.3f'%(j['value'], j['e2e']['value'], j['ms_per_step'], r['kernel_ms']))"; done
