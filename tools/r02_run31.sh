python -m pytest tests/test_gpu_pair.py tests/test_gpu_parity.py -q -k "fast or pair" 2>&1 | tail -3
rm -f gpurun_out/r02_v13.txt
for band in 0 7 2; do for blk in 4000 64; do
  H9_BENCH_BAND=$band H9_BLOCK=$blk python bench.py --grid band8 --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band $band of 8 block $blk', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3))" >> gpurun_out/r02_v13.txt
done; done
cat gpurun_out/r02_v13.txt
