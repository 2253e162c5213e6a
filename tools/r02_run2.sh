set -x
python -m pytest tests/test_gpu_pair.py -x -q 2>&1 | tail -25
python -m pytest tests/test_gpu_parity.py -x -q -k "fast" 2>&1 | tail -8
for band in 0 4; do for blk in 64 4000; do
  H9_BENCH_BAND=$band H9_BLOCK=$blk python bench.py --grid band8 --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band $band block $blk', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3), 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'], 3))" >> gpurun_out/r02_pair.txt
done; done
cat gpurun_out/r02_pair.txt; tail -3 gpurun_out/r02_b.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; tail -3 gpurun_out/r02_bench1.err; cat gpurun_out/r02_bench1.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench1_ref.json 2>> gpurun_out/r02_bench1.err; cat gpurun_out/r02_bench1_ref.json
