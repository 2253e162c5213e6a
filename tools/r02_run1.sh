set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
./tools/ubench/lat > gpurun_out/r02_ubench_lat.txt 2>&1
cat gpurun_out/r02_ubench_lat.txt
for cfg in "64 32" "128 32" "64 16" "128 16" "128 8" "128 12" "64 24"; do
  set -- $cfg
  H9_BLOCK=$1 H9_LANES=$2 python bench.py --grid band8 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band8 block $1 lanes $2', 'ms', round(d['ms_per_step'], 3), 'kernel_ms', round(d['roofline']['kernel_ms_per_launch'], 3))" >> gpurun_out/r02_lanes.txt
done
H9_BENCH_BAND=4 H9_BLOCK=128 H9_LANES=16 python bench.py --grid band8 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band4 block 128 lanes 16', 'ms', round(d['ms_per_step'], 3))" >> gpurun_out/r02_lanes.txt
H9_BENCH_BAND=4 python bench.py --grid band8 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band4 default', 'ms', round(d['ms_per_step'], 3))" >> gpurun_out/r02_lanes.txt
cat gpurun_out/r02_lanes.txt
python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench0.json 2> gpurun_out/r02_bench0.err; cat gpurun_out/r02_bench0.json
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
