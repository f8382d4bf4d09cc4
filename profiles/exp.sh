mkdir -p gpurun_out/final
python bench.py > gpurun_out/final/bench_iquv.json 2> gpurun_out/final/bench_iquv.err; echo exit $?
python bench.py --mode iqud --steps 20 --no-cpu-baseline > gpurun_out/final/bench_iqud.json 2> gpurun_out/final/bench_iqud.err; echo exit $?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu.log 2>&1; echo exit $?
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
