python -m pytest tests -m gpu -x -q -k "pruning" 2>&1 | tail -3
for m in iquv iqud; do python bench.py --steps 20 --no-cpu-baseline --no-prune-pass --mode $m | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('$m value %.4e e2e %.4e step %.3f | pruned kernel %.3f'%(j['value'], j['e2e']['value'], j['ms_per_step'], r['kernel_ms']), j['pruning'])"; done
python bench.py --steps 10 --no-cpu-baseline --no-prune-pass --nx 1024 | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('1024 value %.4e e2e %.4e step %.3f | pruned kernel %.3f'%(j['value'], j['e2e']['value'], j['ms_per_step'], r['kernel_ms']), j['pruning'])"
