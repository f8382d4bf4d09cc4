mkdir -p gpurun_out/r01g
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/dist_check.py > gpurun_out/r01g/dist_check.log 2>&1; echo "dist_check exit $?"; grep -E "world|oracle|Error|error" gpurun_out/r01g/dist_check.log | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01g/bench_n2.json 2> gpurun_out/r01g/bench_n2.err; echo "bench n2 exit $?"
tail -1 gpurun_out/r01g/bench_n2.json | cut -c1-400
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01g/bench_n1.json 2> gpurun_out/r01g/bench_n1.err; tail -1 gpurun_out/r01g/bench_n1.json | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | tail -1 | cut -c1-300
