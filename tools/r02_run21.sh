python -m pytest tests/test_gpu_comm.py -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
echo "rc $?"; tail -5 gpurun_out/r02_bench_n2.err; cat gpurun_out/r02_bench_n2.json | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r02_bench_n2_ref.json 2>> gpurun_out/r02_bench_n2.err
echo "rc $?"; cat gpurun_out/r02_bench_n2_ref.json | cut -c1-400
