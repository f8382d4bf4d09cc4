#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both modes), then the ncu launch list and one full capture of the
# search kernel.  Everything lands under gpurun_out/<tag>/.  Usage: gpurun -- 'bash profiles/gpu_check.sh r01a'
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?" | tee -a $out/pytest.log
tail -3 $out/pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; echo "smoke exit $?" | tee -a $out/smoke.log
timeout 600 python bench.py > $out/bench_iquv.json 2> $out/bench_iquv.err; echo "bench iquv exit $?"
cat $out/bench_iquv.json
timeout 600 python bench.py --steps 20 --warmup 3 --mode iqud --no-cpu-baseline > $out/bench_iqud.json 2> $out/bench_iqud.err; echo "bench iqud exit $?"
cat $out/bench_iqud.json
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
if [ -z "$NO_NCU_FULL" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_match_pruned -s 2 -c 1 -f -o $out/k_match_pruned_iquv \
    python profiles/prof_match.py iquv 256 prune > $out/ncu_full_pruned_iquv.log 2>&1; echo "ncu full pruned iquv exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_match_pruned -s 2 -c 1 -f -o $out/k_match_pruned_iqud \
    python profiles/prof_match.py iqud 256 prune > $out/ncu_full_pruned_iqud.log 2>&1; echo "ncu full pruned iqud exit $?"
fi
if [ -n "$NCU_BRUTE" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_match -s 2 -c 1 -f -o $out/k_match_iquv \
    python profiles/prof_match.py iquv 256 brute > $out/ncu_full_iquv.log 2>&1; echo "ncu full iquv exit $?"
fi
fi
ls -la $out
