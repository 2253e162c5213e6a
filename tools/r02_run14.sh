python -m pytest tests/test_gpu_parity.py tests/test_gpu_pair.py tests/test_ref_golden.py -q 2>&1 | tail -5
rm -f gpurun_out/r02_v7.txt
for band in 0 2 4 6; do for blk in 64 4000; do
  H9_BENCH_BAND=$band H9_BLOCK=$blk python bench.py --grid band8 --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band $band of 8 block $blk', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3), 'shallow', d['config']['share_cells_water_table_in_soil_column_at_end'])" >> gpurun_out/r02_v7.txt
done; done
for band in 0 1 2 3; do for blk in 64 4000; do
  H9_BENCH_NBANDS=4 H9_BENCH_BAND=$band H9_BLOCK=$blk python bench.py --grid band8 --steps 4 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band $band of 4 block $blk', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3), 'shallow', d['config']['share_cells_water_table_in_soil_column_at_end'])" >> gpurun_out/r02_v7.txt
done; done
for band in 0 1; do for blk in 64 1064 4000; do
  H9_BENCH_NBANDS=2 H9_BENCH_BAND=$band H9_BLOCK=$blk H9_PAIR_MAX_CELLS=0 python bench.py --grid band8 --steps 4 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band $band of 2 block $blk', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3), 'shallow', d['config']['share_cells_water_table_in_soil_column_at_end'])" >> gpurun_out/r02_v7.txt
done; done
for blk in 1064 64; do
H9_BLOCK=$blk python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('0.5deg block $blk', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3))" >> gpurun_out/r02_v7.txt
done
cat gpurun_out/r02_v7.txt
