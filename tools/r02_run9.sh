python -m pytest tests/test_gpu_parity.py tests/test_gpu_pair.py tests/test_ref_golden.py -x -q 2>&1 | tail -4
rm -f gpurun_out/r02_v4.txt
for band in 4 0 2 6; do for blk in 64; do
  H9_BENCH_BAND=$band H9_BLOCK=$blk python bench.py --grid band8 --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band $band block $blk', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3), 'shallow', d['config']['share_cells_water_table_in_soil_column_at_end'])" >> gpurun_out/r02_v4.txt
done; done
for blk in 1064; do
H9_BLOCK=$blk python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('0.5deg block $blk', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3), 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'], 3))" >> gpurun_out/r02_v4.txt
done
cat gpurun_out/r02_v4.txt
python tools/cycle_budget.py --band 4 --cell 0 --days 60 --coarse 2>&1 | grep -v Warning | grep -v stddev | grep -v corr > gpurun_out/r02_cycle_budget_band4_v4.txt
cat gpurun_out/r02_cycle_budget_band4_v4.txt
