timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
mkdir -p gpurun_out/multidb
python bench.py --steps 5 --no-cpu-baseline --ndb 256 --nx 1024 > gpurun_out/multidb/ndb256_1024.json 2> gpurun_out/multidb/ndb256_1024.err; echo exit $?; tail -2 gpurun_out/multidb/ndb256_1024.err
python bench.py --steps 10 --no-cpu-baseline --ndb 16 > gpurun_out/multidb/ndb16_256.json 2> gpurun_out/multidb/ndb16_256.err; echo exit $?
python bench.py --steps 20 --no-cpu-baseline > gpurun_out/multidb/default.json 2> gpurun_out/multidb/default.err; echo exit $?
