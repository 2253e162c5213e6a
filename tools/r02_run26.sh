python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
echo "rc $?"; tail -3 gpurun_out/r02_bench_n8.err; cat gpurun_out/r02_bench_n8.json | cut -c1-900
python -m pytest tests/test_gpu_comm.py -q 2>&1 | tail -3
