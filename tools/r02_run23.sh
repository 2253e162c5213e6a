H9_BENCH_BAND=7 ncu --set full --clock-control none --import-source on -k regex:days_kernel --launch-skip 8 -c 1 -f -o gpurun_out/r02_ncu_v9_band7_pair python bench.py --grid band8 --days 365 --steps 1 --warmup 8 --no-cpu --no-e2e > gpurun_out/r02_ncu_v9_band7_pair.log 2>&1
tail -1 gpurun_out/r02_ncu_v9_band7_pair.log
for band in 0 1 2 3; do for blk in 64 1128; do
  H9_BENCH_NBANDS=4 H9_BENCH_BAND=$band H9_BLOCK=$blk python bench.py --grid band8 --steps 4 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band $band of 4 block $blk', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3))" >> gpurun_out/r02_v10.txt
done; done
cat gpurun_out/r02_v10.txt
