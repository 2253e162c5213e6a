python -m pytest tests -m gpu -q 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
rm -f gpurun_out/r02_v9.txt
for band in 0 1 2 3 4 5 6 7; do
  H9_BENCH_BAND=$band python bench.py --grid band8 --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band $band of 8 auto', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3), 'shallow', d['config']['share_cells_water_table_in_soil_column_at_end'])" >> gpurun_out/r02_v9.txt
done
H9_BENCH_BAND=0 H9_BLOCK=64 python bench.py --grid band8 --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/r02_b.err | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band 0 of 8 block 64', d['config']['kernel_variant'], 'ms', round(d['ms_per_step'], 3))" >> gpurun_out/r02_v9.txt
cat gpurun_out/r02_v9.txt
