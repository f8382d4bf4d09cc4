#!/usr/bin/env python
"""profiles/kernel_traffic.json from the committed-run ncu reports: dram__bytes_read.sum + dram__bytes_write.sum per launch.
Usage (CPU box, after a profiling pass): python profiles/make_traffic.py name=path.ncu-rep ... > profiles/kernel_traffic.json"""
import csv, io, json, subprocess, sys
out = {'comment': 'dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none (profiles/gpu_profile_r02.sh); '
                  'search kernels: BASELINE config 2/3 (256x256 map, one 340200-entry int16 database, nsearch 8); elementwise: 2048x2048 map'}
for arg in sys.argv[1:]:
    name, path = arg.split('=')
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    def val(metric):
        i = h.index(metric)
        x = float(v[i].replace(',', ''))
        return x * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u[i]]
    out[name] = int(val('dram__bytes_read.sum') + val('dram__bytes_write.sum'))
print(json.dumps(out, indent=1))
