#!/bin/bash
# Round-2 profiling pass (one GPU): launch list of the bench command, full captures of the search kernels and the elementwise
# kernels, noise sweep.  Usage: gpurun -- 'bash profiles/gpu_profile_r02.sh tag'.  Reports land under gpurun_out/<tag>/.
tag=${1:-prof}; out=gpurun_out/$tag; mkdir -p $out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs 2b,4 > $out/bench_plain.json 2> $out/bench_plain.err; echo "bench (no ncu) exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs 2b,4 > $out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
cap() {  # name, kernel regex, skip, script args...
  name=$1; kre=$2; skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o $out/$name "$@" > $out/ncu_$name.log 2>&1; echo "ncu $name exit $?"
}
cap k_match_pruned_iquv  k_match_pruned 2 python profiles/prof_match.py iquv 256 prune
cap k_match_pruned_iqud  k_match_pruned 2 python profiles/prof_match.py iqud 256 prune
cap k_match_plain_iquv   '^k_match$'   2 python profiles/prof_match.py iquv 256 brute
cap k_match_seeded_iquv  '^k_match$'   2 python profiles/prof_match.py iquv 256 seeded
cap k_match_seeded_1024  '^k_match$'   1 python profiles/prof_match.py iquv 1024 seeded
cap k_blos               k_blos         1 python profiles/prof_elementwise.py
cap k_qurotate           k_qurotate     1 python profiles/prof_elementwise.py
python profiles/noise_sweep.py > $out/noise_sweep.txt 2>&1; echo "noise sweep exit $?"
ls -la $out
