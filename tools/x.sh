out=gpurun_out/pass3.txt; : > $out
run() { H9_BENCH_NBANDS=$1 H9_BENCH_BAND=$2 python bench.py --grid band8 --block $3 --steps 4 --warmup 3 --no-cpu --no-e2e --no-weak 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('band $2 of $1 block $3', d['config'].get('kernel_variant'), 'ms', round(d['ms_per_step'], 3))" >> $out; }
run 8 4 0; run 8 0 0; run 8 7 0; run 8 2 0
python -m pytest tests/test_gpu_pair.py tests/test_gpu_fullsize.py -q -x -k "not multi_decade" 2>&1 | tail -3 >> $out
cat $out
