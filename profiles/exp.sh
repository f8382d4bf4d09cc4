mkdir -p gpurun_out/r01n8
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01n8/bench_n8.json 2> gpurun_out/r01n8/bench_n8.err; echo "bench n8 exit $?"
tail -1 gpurun_out/r01n8/bench_n8.json | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('n8 value %.4e e2e %.4e ms %.3f frac %.3f'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac']))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01n8/bench_n4.json 2> gpurun_out/r01n8/bench_n4.err; echo "bench n4 exit $?"
tail -1 gpurun_out/r01n8/bench_n4.json | python -c "
import sys,json
j=json.loads(sys.stdin.read()); print('n4 value %.4e e2e %.4e ms %.3f frac %.3f'%(j['value'], j['e2e']['value'], j['ms_per_step'], j['roofline']['frac']))"
tail -3 gpurun_out/r01n8/bench_n8.err
