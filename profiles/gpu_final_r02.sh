#!/bin/bash
# Final round-2 pass (one GPU): parity tests, smoke, the default bench line, the launch list of the bench command and full captures of
# the kernels this round's last changes touched.  Usage: gpurun -- 'bash profiles/gpu_final_r02.sh tag'.  Reports land under gpurun_out/<tag>/.
tag=${1:-final}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/pytest.log
timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' > $out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $out/smoke.log
timeout 400 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench exit $?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $out/ref.json 2> $out/ref.err; echo "reference arm exit $?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs 2b,4 > $out/bench_plain.json 2> $out/bench_plain.err; echo "bench (short, no ncu) exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs 2b,4 > $out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
cap() {  # name, kernel regex, skip, script args...
  name=$1; kre=$2; skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o $out/$name "$@" > $out/ncu_$name.log 2>&1; echo "ncu $name exit $?"
}
python profiles/prof_build.py 1024 > $out/prof_build_plain.log 2>&1; echo "prof_build (no ncu) exit $?"
cap k_push_records_1024 k_push_records 1 python profiles/prof_build.py 1024
cap k_build_1024        k_build        1 python profiles/prof_build.py 1024
cap k_finalize_1024     k_finalize     1 python profiles/prof_build.py 1024
ls -la $out
