#!/usr/bin/env python
"""Short summary of a bench.py JSON line: python profiles/show_bench.py gpurun_out/<tag>/bench.json"""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads([l for l in open(path) if l.startswith('{')][-1])
    print('%s: N=%d value %.1f M px/s  step %.4f ms  e2e %.1f M  kernel %.3f ms  launches %d' % (
        path, d['n_gpus'], d['value'] / 1e6, d['ms_per_step'], d['e2e']['value'] / 1e6, d['roofline']['kernel_ms'], d['gpu_launches']))
    if d.get('gather_overlap'):
        print('   serial step %.4f ms' % d['gather_overlap']['ms_per_step_serial'])
    py = d['e2e'].get('python')
    if py:
        print('   e2e %.3f ms; python median %.3f ms; stream of maps %.3f ms' % (d['e2e']['ms_per_step'], py['ms_per_call_median'],
                                                                              d['e2e'].get('stream_of_maps', {}).get('ms_per_map', float('nan'))))
    for name, r in d['by_config'].items():
        roof = r.get('roofline') or {}
        extra = ''
        if 'step_breakdown_ms' in r:
            extra = ' parts %s' % {k: round(v, 3) for k, v in r['step_breakdown_ms'].items() if k != 'note'}
            extra += ' transport %s' % r.get('transport')
            if r.get('nccl_transport'):
                extra += ' | nccl %.3f ms' % r['nccl_transport']['ms_per_step']
            if r.get('bruteforce'):
                extra += ' | brute %.2f ms frac %.3f' % (r['bruteforce']['ms_per_step'], r['bruteforce']['roofline']['frac'] or 0)
        e2e = r.get('e2e') or {}
        print('   %-36s %10.2f M/s step %8.4f ms  kernel %s frac %s  e2e %s ms%s' % (
            name, r['value'] / 1e6, r['ms_per_step'], ('%.3f' % roof['kernel_ms']) if roof.get('kernel_ms') else '-',
            ('%.3f' % roof['frac']) if roof.get('frac') else '-', ('%.3f' % e2e['ms_per_step']) if e2e.get('ms_per_step') else '-', extra))
