python -m pytest tests/test_gpu_pair.py tests/test_gpu_parity.py -q -k "pair or two_lanes" 2>&1 | tail -4
rm -f gpurun_out/r02_v12.txt
for band in 0 2 4 7; do for blk in 4000; do
  H9_BENCH_BAND=$band H9_BLOCK=$blk python bench.py --grid band8 --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band $band of 8 block $blk', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3), 'shallow', d['config']['share_cells_water_table_in_soil_column_at_end'])" >> gpurun_out/r02_v12.txt
done; done
cat gpurun_out/r02_v12.txt
