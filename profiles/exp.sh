timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python profiles/noise_sweep.py 2>&1 | tail -10
python bench.py --steps 20 --no-cpu-baseline | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']; rp=j['roofline_pruned']
print('value %.4e e2e %.4e step %.3f | pruned kernel %.3f | brute kernel %.3f frac %.3f'%(j['value'], j['e2e']['value'], j['ms_per_step'], rp['kernel_ms'], r['kernel_ms'], r['frac']))"
