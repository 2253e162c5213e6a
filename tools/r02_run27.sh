python -m pytest tests/test_gpu_comm.py tests/test_gpu_api.py -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 --no-weak > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err
echo "rc $?"; tail -2 gpurun_out/r02_bench_n2b.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2b.json')); print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'kernel', d['roofline']['kernel_ms_per_launch'], 'e2e', d['e2e']['ms_per_step'])"
