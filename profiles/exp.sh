for m in iquv iqud; do python bench.py --steps 20 --no-cpu-baseline --mode $m > gpurun_out/exp_$m.json 2> gpurun_out/exp_$m.err; echo exit $?; tail -3 gpurun_out/exp_$m.err; done
python bench.py --steps 10 --no-cpu-baseline --nx 1024 > gpurun_out/exp_1024.json 2> gpurun_out/exp_1024.err; echo exit $?
